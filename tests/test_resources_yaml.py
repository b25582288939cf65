"""Resources.from_yaml_config / _parse_accelerators_from_str (host logic; the
behaviour of sky/resources.py:2209-2411). CPU only."""
import pandas as pd
import pytest

import skypilot_b200 as sky
from skypilot_b200 import synth
from skypilot_b200.catalog.store import CatalogStore

R = sky.Resources


@pytest.fixture(scope='module', autouse=True)
def store():
    st = CatalogStore.from_frames(synth.make_catalogs(seed=7, n_rows=2000,
                                                      clouds=['aws', 'gcp']))
    st.set_accelerator_metadata(synth.accelerator_metadata())
    sky.catalog.set_store(st)
    yield st


def _accs(resources):
    return sorted(list(r.accelerators.items())[0] for r in resources)


def test_plain_name_is_one_user_specified_alternative():
    assert R._parse_accelerators_from_str('A100:8') == [('A100:8', True)]  # pylint: disable=protected-access


@pytest.mark.parametrize('spec,want', [
    ('80GB', [('A100-80GB', 1), ('H100', 1)]),
    ('80GB+', [('A100-80GB', 1), ('B200', 1), ('H100', 1), ('H200', 1),
               ('MI300X', 1)]),
    ('16GB:4', [('P100', 4), ('T4', 4), ('V100', 4)]),
    ('nvidia:24GB', [('A10', 1), ('A10G', 1), ('L4', 1), ('RTX6000', 1)]),
    ('AMD:100GB+:8', [('MI300X', 8)]),
])
def test_memory_specs_expand_through_the_metadata_table(spec, want):
    got = R.from_yaml_config({'accelerators': spec})
    assert isinstance(got, set)
    assert _accs(got) == want
    assert all(r.no_missing_accel_warnings for r in got)


def test_list_of_accelerators_keeps_its_order_and_set_does_not():
    got = R.from_yaml_config({'accelerators': ['H100:8', 'A100:8'],
                              'use_spot': True})
    assert isinstance(got, list)
    assert [list(r.accelerators.items())[0] for r in got] == [('H100', 8),
                                                              ('A100', 8)]
    assert all(r.use_spot for r in got)
    got = R.from_yaml_config({'accelerators': {'V100': 1, 'T4': 2}})
    assert isinstance(got, set) and _accs(got) == [('T4', 2), ('V100', 1)]


def test_any_of_and_ordered():
    got = R.from_yaml_config({'cpus': 8, 'any_of': [{'cloud': 'aws'},
                                                    {'cloud': 'gcp',
                                                     'memory': '32+'}]})
    assert isinstance(got, set) and len(got) == 2
    assert sorted(str(r.cloud) for r in got) == ['AWS', 'GCP']
    assert all(r.cpus == '8' for r in got)
    got = R.from_yaml_config({'ordered': [{'accelerators': 'A100:8'},
                                          {'accelerators': 'V100:8',
                                           'use_spot': True}]})
    assert isinstance(got, list)
    assert [r.use_spot for r in got] == [False, True]
    # several accelerators inside any_of multiply out
    got = R.from_yaml_config({'any_of': [{'accelerators': {'T4': 1, 'L4': 1}},
                                         {'cloud': 'aws', 'cpus': '4+'}]})
    assert len(got) == 3


def test_errors():
    with pytest.raises(ValueError, match='both "any_of" and "ordered"'):
        R.from_yaml_config({'any_of': [{}], 'ordered': [{}]})
    with pytest.raises(ValueError, match='multiple "accelerators" with '
                       '"ordered"'):
        R.from_yaml_config({'accelerators': ['A100', 'V100'],
                            'ordered': [{'cloud': 'aws'}]})
    with pytest.raises(ValueError, match='preferred order'):
        R.from_yaml_config({'accelerators': ['A100', 'V100'],
                            'any_of': [{'cloud': 'aws'}]})
    with pytest.raises(ValueError, match='both gpus and accelerators'):
        R.from_yaml_config({'gpus': 'T4', 'accelerators': 'V100'})
    with pytest.raises(AssertionError, match='Invalid resource args'):
        R.from_yaml_config({'no_such_field': 1})


def test_none_and_alias():
    got = R.from_yaml_config(None)
    assert len(got) == 1 and list(got)[0].is_empty()
    got = R.from_yaml_config({'gpus': 'L4:2'})
    assert _accs(got) == [('L4', 2)]
