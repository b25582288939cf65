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


def _random_listing_kwargs(rng, clouds):
    names = [None, None, 'A100', 'a100', 'V100|T4', '^H', 'tpu', 'tpu-v[34]',
             'L4$', 'A10', 'K80', '100', 'Trainium', '.*']
    regions = [None, None, None, 'us', 'us-east', 'europe|eu-', 'west', 'EAST',
               'asia', 'central1$']
    kw = {}
    name = names[int(rng.integers(len(names)))]
    if name is not None:
        kw['name_filter'] = name
    region = regions[int(rng.integers(len(regions)))]
    if region is not None:
        kw['region_filter'] = region
    if rng.uniform() < 0.35:
        kw['quantity_filter'] = int(rng.choice([1, 2, 4, 8, 16]))
    kw['gpus_only'] = bool(rng.uniform() < 0.8)
    kw['case_sensitive'] = bool(rng.uniform() < 0.6)
    kw['all_regions'] = bool(rng.uniform() < 0.4)
    kw['require_price'] = bool(rng.uniform() < 0.7)
    n = int(rng.integers(1, len(clouds) + 1))
    kw['clouds'] = [clouds[i] for i in sorted(
        rng.choice(len(clouds), size=n, replace=False))]
    return kw


@pytest.mark.parametrize('seed', range(40))
def test_random_listings_match_the_oracle(seed):
    import numpy as np
    _, frames, _ = load('multi6k')
    activate('multi6k')
    rng = np.random.default_rng(1000 + seed)
    kw = _random_listing_kwargs(rng, ['aws', 'gcp', 'azure', 'lambda'])
    want = lo.list_accelerators(frames, **kw)
    got = sky.catalog.list_accelerators(**kw)
    assert list(got.keys()) == list(want.keys()), kw
    as_lists = {k: [list(i) for i in v] for k, v in got.items()}
    assert normalise(as_lists) == normalise(want), kw
