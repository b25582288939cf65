"""Property test on CPU: for seeded random requests (other seeds than the
reference-pinned `fuzz6k` / `fuzzmany` fixtures) the constraint vectors the
product states pick the instance type the oracle's `feasible()` picks, cloud
by cloud. The oracle is pinned against the unmodified reference by
tests/test_oracle.py; see tests/test_host_statement.py for the method."""
import pytest

from oracle import optimizer_oracle as oo
from tests import scenario_runner as runner
from tests import scenarios
from tests import test_host_statement as ths

_MANY_REGIONS = {
    'aws': ['us-east-1', 'ap-south-1'], 'gcp': ['asia-east1'],
    'lambda': ['us-east-1', 'me-west-1'], 'runpod': ['US', 'NL'],
    'cudo': ['no-luster-1', 'us-newyork-1'], 'ibm': ['us-south', 'eu-de'],
    'oci': ['us-ashburn-1', 'ap-tokyo-1'], 'nebius': ['eu-north1'],
    'scp': ['KR-EAST-3', 'KOREA-EAST-1-SCP-B001'], 'verda': ['FIN-01'],
}


@pytest.mark.parametrize('catalog,seed', [('fuzz6k', 101), ('fuzz6k', 102),
                                          ('fuzzmany', 201),
                                          ('fuzzmany', 202)])
def test_random_requests_are_stated_like_the_oracle(catalog, seed):
    spec = scenarios.CATALOGS[catalog]
    runner.activate_catalog(spec)
    cat = oo.catalog_for(spec)
    regions = _MANY_REGIONS if catalog == 'fuzzmany' else None
    failures = []
    for sc in scenarios.fuzz_scenarios(seed=seed, n=60, regions=regions,
                                       prefix=f's{seed}_'):
        try:
            ths.test_stated_queries_pick_the_oracles_instance(catalog, sc)
        except ValueError as e:
            # a TPU request with a memory *ratio* makes the reference itself
            # raise ValueError (float('8x'), sky/clouds/gcp.py:796-806); the
            # oracle and the product inherit that
            assert 'could not convert string to float' in str(e), sc
            assert any('tpu' in str(r.get('accelerators', ''))
                       for t in sc['tasks'] for r in t['resources']), sc
        except BaseException as e:  # pylint: disable=broad-except
            if isinstance(e, (KeyboardInterrupt, SystemExit)):
                raise
            failures.append((sc['name'], str(e)[:200]))
    del cat
    assert not failures, failures
