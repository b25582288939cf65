"""CPU tests of the oracle: it must reproduce (1) every record the UNMODIFIED
reference produced for tests/scenarios.py (tests/golden/*.json) and (2) the
inline known-answer vectors of the reference's own unit tests."""
import os

import pytest

from oracle import catalog_oracle as co
from oracle import optimizer_oracle as oo
from tests import reference_vectors as rv
from tests import scenario_runner as runner
from tests import scenarios

# the 50k-row suites take ~15 s each on CPU; keep one, sample the other
_FULL = ['multi6k', 'three4k', 'aws50k', 'gpuclouds', 'ibm5k', 'hyperprime',
         'latecl', 'fuzz6k', 'oci5k', 'nebvast', 'scp4k', 'fuzzmany', 'vsphere3k', 'seeweb3k', 'shade3k', 'fuzzdag']
_SAMPLED = {'multi50k': 4}


def _cases():
    out = []
    for catalog in scenarios.ALL_SUITES:
        if not os.path.exists(
                os.path.join(runner.GOLDEN_DIR, f'{catalog}.json')):
            continue
        suite = scenarios.ALL_SUITES[catalog]()
        if catalog in _SAMPLED:
            suite = suite[::_SAMPLED[catalog]]
        elif catalog not in _FULL:
            continue
        for sc in suite:
            out.append(pytest.param(catalog, sc, id=f'{catalog}:{sc["name"]}'))
    return out


_golden = {}


def _records(catalog):
    if catalog not in _golden:
        payload = runner.load_golden(catalog)
        _golden[catalog] = (payload['catalog'],
                            {r['name']: r for r in payload['records']})
    return _golden[catalog]


@pytest.mark.parametrize('catalog,scenario', _cases())
def test_oracle_matches_reference_fixture(catalog, scenario):
    spec, records = _records(catalog)
    got = oo.run_scenario(spec, scenario)
    unordered = any(
        t.get('resources_kind') == 'set' for t in scenario['tasks'])
    # the port restates the algorithm, not the error text: the class of the
    # error is compared, the message only for the product (GPU tests)
    diffs = runner.compare(records[scenario['name']], got, unordered,
                           check_message=False)
    assert not diffs, '\n'.join(diffs)


@pytest.mark.parametrize('cpus,memory,region,zone,expected', rv.AZ_CASES)
def test_cpus_mem_with_az(cpus, memory, region, zone, expected):
    assert co.instance_type_for_cpus_mem(rv.az_frame(), cpus, memory, region,
                                         zone) == expected


@pytest.mark.parametrize('cpus,memory,region,expected', rv.NO_AZ_CASES)
def test_cpus_mem_no_az(cpus, memory, region, expected):
    assert co.instance_type_for_cpus_mem(rv.no_az_frame(), cpus, memory,
                                         region) == expected


def test_hourly_cost_is_python_float():
    df = rv.price_frame()
    for spot, want in ((False, 1.5), (True, 0.5)):
        cost = co.hourly_cost(df, 'test-instance', spot, None, None)
        assert type(cost) is float and cost == want  # pylint: disable=unidiomatic-typecheck


@pytest.mark.parametrize('local_disk,expected', rv.LOCAL_DISK_CASES)
def test_local_disk_selection(local_disk, expected):
    df = co.filter_with_local_disk(rv.local_disk_frame(), local_disk)
    assert co.instance_type_for_cpus_mem(df, '1+', None) == expected


@pytest.mark.parametrize('local_disk,expected', rv.LOCAL_DISK_SETS)
def test_local_disk_sets(local_disk, expected):
    df = co.filter_with_local_disk(rv.local_disk_frame(), local_disk)
    assert sorted(df['InstanceType'].tolist()) == sorted(expected)


def test_egress_tariffs_known_values():
    # hand-evaluated from the piecewise definitions (aws.py:667-688,
    # gcp.py:395-404, azure.py:142-165)
    assert oo.egress_tariff('aws', 0.5) == 0.0
    assert oo.egress_tariff('aws', 500) == pytest.approx(499 * 0.09)
    assert oo.egress_tariff('aws', 20 * 1024) == pytest.approx(
        (20480 - 10240) * 0.085 + (10240 - 1) * 0.09)
    assert oo.egress_tariff('aws', 200 * 1024) == pytest.approx(0.05 * 204800)
    assert oo.egress_tariff('gcp', 100) == pytest.approx(12.0)
    assert oo.egress_tariff('gcp', 5000) == pytest.approx(550.0)
    assert oo.egress_tariff('gcp', 20000) == pytest.approx(1600.0)
    assert oo.egress_tariff('azure', 500) == pytest.approx(499 * 0.0875)
    assert oo.egress_tariff('lambda', 500) == 0.0
