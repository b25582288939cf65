"""GPU tests: `skypilot_b200.catalog.list_accelerators` (device reduction
`skyopt_list_offerings` + GCP host-VM scan) against the reference fixtures
and, where the reference errored under pandas 3.0, against the oracle."""
import pytest

import skypilot_b200 as sky
from oracle import listing_oracle as lo
from tests import scenario_runner as runner
from tests import scenarios
from tests.test_oracle_listing import cases, load, normalise

pytestmark = pytest.mark.gpu

_active = {}


def activate(catalog):
    if _active.get('name') != catalog:
        spec, _, _ = load(catalog)
        runner.activate_catalog(dict(spec))
        _active['name'] = catalog


@pytest.mark.parametrize('catalog,case', cases())
def test_listing_matches_reference(catalog, case):
    _, frames, records = load(catalog)
    activate(catalog)
    rec = records[case['name']]
    got = sky.catalog.list_accelerators(**case['kwargs'])
    as_lists = {k: [list(i) for i in v] for k, v in got.items()}
    if 'error' in rec:
        want = lo.list_accelerators(frames, **case['kwargs'])
        assert list(got.keys()) == list(want.keys())
        assert normalise(as_lists) == normalise(want)
        return
    assert list(got.keys()) == rec['order']
    assert normalise(as_lists) == rec['listing']


def test_accelerator_counts():
    activate('multi6k')
    _, frames, _ = load('multi6k')
    clouds = ['aws', 'gcp', 'azure', 'lambda']
    got = sky.catalog.list_accelerator_counts(clouds=clouds)
    want = {}
    for name, infos in lo.list_accelerators(frames, clouds=clouds,
                                            require_price=False).items():
        want[name] = sorted({i[3] for i in infos})
    assert got == want
