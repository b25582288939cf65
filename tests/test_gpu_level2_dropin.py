"""INTEGRATION.md Level 2 as a test: the unmodified reference optimizer
(baseline/_ref, the pip-installed copy that travels with the repo) driving
skypilot_b200's catalog function table on the GPU must give the records of
the unmodified reference on its own pandas catalog (tests/golden/<catalog>.json)
-- whole suites: plans, ordered candidate tables, objectives, error texts."""
import json
import os
import subprocess
import sys

import pytest

from tests import scenario_runner as runner
from tests import scenarios

pytestmark = pytest.mark.gpu

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('catalog', ['multi6k', 'three4k', 'gpuclouds'])
def test_reference_optimizer_on_our_catalog_table(tmp_path, catalog):
    ref = os.path.join(_REPO, 'baseline', '_ref', 'sky')
    if not os.path.isdir(ref):
        pytest.skip('baseline/_ref (the pip-installed reference) is absent')
    payload = runner.load_golden(catalog)
    golden = {r['name']: r for r in payload['records']}
    suite = {s['name']: s for s in scenarios.SUITES[catalog]()}
    if catalog != 'gpuclouds':
        # every catalog call is a device round trip here (tens of thousands
        # for a whole suite: it runs green, 100 + 95 records, in four
        # minutes): the test keeps every third scenario and the chains
        names = list(suite)
        keep = set(names[::3]) | {n for n in names if 'chain' in n or
                                  'diamond' in n or 'cfg' in n}
        suite = {n: suite[n] for n in names if n in keep}
    out = tmp_path / 'level2.json'
    proc = subprocess.run(
        [sys.executable, os.path.join(_REPO, 'tests', 'level2_child.py'),
         catalog, str(out)] + (list(suite) if catalog != 'gpuclouds' else []),
        cwd=_REPO, capture_output=True, text=True, timeout=1500, check=False)
    assert proc.returncode == 0, proc.stdout[-1500:] + proc.stderr[-3000:]
    got = json.loads(out.read_text())
    assert got['catalog_calls'] > 20  # the table really was ours
    assert len(got['records']) == len(suite)
    failures = []
    for rec in got['records']:
        sc = suite[rec['name']]
        unordered = any(
            t.get('resources_kind') == 'set' for t in sc['tasks'])
        want = golden[rec['name']]
        diffs = runner.compare(want, rec, unordered_candidates=unordered)
        if diffs:
            failures.append((rec['name'], diffs[:2],
                             rec.get('error', {}).get('traceback', '')[-600:]))
    assert not failures, failures
