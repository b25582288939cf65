"""Regenerate tests/golden/<catalog>.json from the UNMODIFIED reference.

Build container only (needs /root/reference). One subprocess per catalog
(the reference binds its catalog directory at import time):

    python oracle/ref_harness/gen_golden.py [catalog ...]

The fixtures are committed; the GPU box never runs this script.
"""
import json
import os
import subprocess
import sys
import tempfile

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(os.path.dirname(_HERE))
sys.path.insert(0, _REPO)

from tests import scenarios  # noqa: E402  pylint: disable=wrong-import-position


def main():
    """`gen_golden.py [catalog ...]` regenerates the optimizer fixtures,
    `gen_golden.py --listings [catalog ...]` the accelerator listings
    (tests/golden/accel_<catalog>.json), `--job-groups` the JobGroup plans
    (tests/golden/jobgroup_<catalog>.json), `--calls` the catalog function
    calls (tests/golden/calls_<catalog>.json)."""
    argv = sys.argv[1:]
    listings = '--listings' in argv
    groups = '--job-groups' in argv
    calls = '--calls' in argv
    argv = [a for a in argv
            if a not in ('--listings', '--job-groups', '--calls')]
    suites = (scenarios.LISTING_SUITES if listings else
              scenarios.JOB_GROUP_SUITES if groups else
              scenarios.CALL_SUITES if calls else
              dict(scenarios.ALL_SUITES, **scenarios.EXTRA_GOLDEN_SUITES))
    prefix = ('accel_' if listings else 'jobgroup_' if groups else
              'calls_' if calls else '')
    wanted = argv or [k for k in suites
                      if k not in scenarios.EXTRA_GOLDEN_SUITES or not argv]
    out_dir = os.path.join(_REPO, 'tests', 'golden')
    os.makedirs(out_dir, exist_ok=True)
    for name in wanted:
        spec = dict(scenarios.CATALOGS[name])
        suite = suites[name]()
        with tempfile.NamedTemporaryFile('w', suffix='.json',
                                         delete=False) as f:
            json.dump(suite, f)
            sc_path = f.name
        out_path = os.path.join(out_dir, f'{prefix}{name}.json')
        tmp_out = out_path + '.tmp'
        cmd = [
            sys.executable,
            os.path.join(_HERE, 'run_reference.py'), '--catalog',
            json.dumps(spec), '--scenarios', sc_path, '--out', tmp_out
        ]
        print('[gen_golden]', name, len(suite), 'scenarios', flush=True)
        proc = subprocess.run(cmd, cwd='/tmp', stdout=subprocess.DEVNULL,
                              check=False)
        os.unlink(sc_path)
        if proc.returncode != 0:
            raise SystemExit(f'{name}: reference harness failed')
        with open(tmp_out, encoding='utf-8') as f:
            records = json.load(f)
        os.unlink(tmp_out)
        # compact candidates: [cloud, instance_type, region, zone, value]
        for rec in records:
            if 'candidates' in rec:
                rec['candidates'] = [[[
                    c['cloud'], c['instance_type'], c['region'], c['zone'],
                    c['value']
                ] for c in cands] for cands in rec['candidates']]
        payload = {
            'catalog': spec,
            'generated_by': 'oracle/ref_harness/gen_golden.py',
            'reference': 'skypilot-org/skypilot @ 7808630 (unmodified)',
            'records': records
        }
        with open(out_path, 'w', encoding='utf-8') as f:
            json.dump(payload, f, separators=(',', ':'), sort_keys=True)
        n_err = sum('error' in r for r in records)
        print(f'[gen_golden] wrote {out_path}: {len(records)} records, '
              f'{n_err} errors', flush=True)


if __name__ == '__main__':
    main()
