"""The catalog function table (skypilot_b200.catalog, the drop-in for
sky/catalog/__init__.py) against answers of the unmodified reference
(tests/golden/calls_three4k.json, oracle/ref_harness/gen_golden.py --calls):
GCP host attachability, accelerator counts, image tags, scalar look-ups --
results, exception classes and exception texts."""
import math

import pytest

import skypilot_b200 as sky
from tests import scenario_runner as runner
from tests import scenarios

pytestmark = pytest.mark.gpu

_CASES = scenarios.catalog_call_cases()


def _plain(x):
    if isinstance(x, (list, tuple)):
        return [_plain(v) for v in x]
    if isinstance(x, dict):
        return {str(k): _plain(v) for k, v in x.items()}
    if x is None or isinstance(x, (bool, str)):
        return x
    if isinstance(x, (int, float)) or hasattr(x, 'dtype'):
        v = float(x)
        if math.isnan(v):
            return None
        return int(v) if isinstance(x, int) else v
    if hasattr(x, 'name'):
        return x.name
    return str(x)


def _same(a, b) -> bool:
    if isinstance(a, list) and isinstance(b, list):
        return len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    if isinstance(a, dict) and isinstance(b, dict):
        # (the fixture file is written with sorted keys: no order to compare)
        return (sorted(a.keys()) == sorted(b.keys()) and
                all(_same(a[k], b[k]) for k in a))
    if isinstance(a, bool) or isinstance(b, bool) or a is None or b is None \
            or isinstance(a, str) or isinstance(b, str):
        return a == b and type(a) is type(b)  # pylint: disable=unidiomatic-typecheck
    return math.isclose(float(a), float(b), rel_tol=1e-6, abs_tol=1e-12)


@pytest.fixture(scope='module')
def golden():
    payload = runner.load_golden('calls_three4k')
    assert payload['catalog'] == scenarios.CATALOGS['three4k']
    runner.activate_catalog(payload['catalog'])
    return {r['name']: r for r in payload['records']}


@pytest.mark.parametrize('case', _CASES, ids=lambda c: c['name'])
def test_catalog_call_matches_reference(golden, case):
    want = golden[case['name']]
    fn = getattr(sky.catalog, case['fn'])
    try:
        got = {'result': _plain(fn(*case['args'], **case['kwargs']))}
    except Exception as e:  # pylint: disable=broad-except
        got = {'error': {'type': type(e).__name__, 'message': str(e)}}
    if 'error' in want:
        assert 'error' in got, (want['error'], got)
        assert got['error']['type'] == want['error']['type'], got
        if want['error']['type'] != 'AssertionError':
            assert got['error']['message'] == want['error']['message']
        return
    assert 'error' not in got, got
    assert _same(want['result'], got['result']), (want['result'],
                                                  got['result'])
