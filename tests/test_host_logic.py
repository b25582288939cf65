"""CPU tests of the host side: catalog ingest, constraint-vector
construction, Resources / Dag semantics, egress tariffs and the C ABI
(symbols and struct layouts only -- no kernel runs without a GPU)."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

import skypilot_b200 as sky
from skypilot_b200 import _native
from skypilot_b200 import engine
from skypilot_b200 import synth
from skypilot_b200.catalog import rules
from skypilot_b200.catalog.store import CatalogStore
from skypilot_b200.utils import registry
from tests import reference_vectors as rv
from tests import scenario_runner as runner
from tests import scenarios

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def store():
    return runner.activate_catalog(scenarios.CATALOGS['multi6k'])


# ---- C ABI -----------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    """Every function include/skyopt.h declares is exported by libskyopt.so."""
    with open(os.path.join(_REPO, 'include', 'skyopt.h'),
              encoding='utf-8') as f:
        header = f.read()
    declared = set(re.findall(r'\b(skyopt_[a-z_]+)\s*\(', header))
    assert declared == set(_native.EXPORTS)
    lib = _native.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.skyopt_abi_version() == _native.ABI_VERSION


def test_struct_layouts_match_header():
    """numpy dtypes mirror the C structs (sizes / key offsets)."""
    assert _native.QUERY_DTYPE.itemsize == 88
    assert _native.QUERY_DTYPE.fields['cpus'][1] == 56
    assert _native.QUERY_DTYPE.fields['max_price'][1] == 80
    assert _native.SLOT_DTYPE.itemsize == 80
    assert _native.SLOT_DTYPE.fields["region_set"][1] == 48
    assert _native.SLOT_DTYPE.fields["hours"][1] == 56
    assert _native.CANDIDATE_DTYPE.fields['hourly'][1] == 16
    assert _native.DAG_RESULT_DTYPE.fields['objective'][1] == 8
    assert ctypes.sizeof(_native.Stats) == 56
    assert _native.Stats.scan_form.offset == 48


def test_no_device_fails_loudly():
    """Without a GPU the product path raises; it never falls back to CPU."""
    if _native.device_count() > 0:
        pytest.skip('a GPU is present')
    frames = synth.make_catalogs(5, 600, clouds=['aws'])
    st = CatalogStore.from_frames(frames)
    with pytest.raises(_native.SkyoptError):
        st.handle(0)


# ---- ingest ----------------------------------------------------------------
def test_ingest_layout(store):
    c = store.columns
    offs = c['cloud_row_offsets']
    assert offs[0] == 0 and offs[-1] == store.n_rows
    assert all(o % 8 == 0 for o in offs)
    assert store.n_real_rows == sum(t.n_rows for t in store.clouds)
    for name in ('price', 'spot_price', 'vcpus', 'mem'):
        assert c[name].dtype == np.float64 and len(c[name]) == store.n_rows
    for name in ('acc_key', 'region_id', 'zone_id', 'flags'):
        assert c[name].dtype == np.uint16 and len(c[name]) == store.n_rows
    # padding rows are invalid
    for t, b, e in zip(store.clouds, offs[:-1], offs[1:]):
        assert np.all(c['flags'][b:b + t.n_rows] & _native.F_VALID)
        assert not np.any(c['flags'][b + t.n_rows:e] & _native.F_VALID)


def test_ingest_columns_round_trip(store):
    """Decoding the SoA columns gives back the CSV frame of every cloud."""
    c = store.columns
    for t in store.clouds:
        b, e = t.row_begin, t.row_end
        df = t.frame
        assert np.array_equal(
            np.nan_to_num(c['price'][b:e], nan=-1.0),
            np.nan_to_num(df['Price'].to_numpy(dtype=float), nan=-1.0))
        regions = [t.region_names[i] for i in c['region_id'][b:e]]
        assert regions == list(df['Region'])
        inst = [
            None if i < 0 else store.inst_names[i] for i in c['inst_id'][b:e]
        ]
        want = [None if x != x or x is None else x
                for x in df['InstanceType'].astype(object)]
        assert inst == want
        if t.has_zone_column:
            zones = [
                None if z == _native.NONE16 else t.zone_names[z]
                for z in c['zone_id'][b:e]
            ]
            wantz = [None if x != x or x is None else x
                     for x in df['AvailabilityZone'].astype(object)]
            assert zones == wantz


def test_region_ids_are_ranks_in_string_order(store):
    for t in store.clouds:
        assert t.region_names == sorted(t.region_names)
        assert t.zone_names == sorted(t.zone_names)


def test_csr_groups_rows_by_instance_type(store):
    c = store.columns
    off, rows, iid = c['inst_row_offsets'], c['inst_rows'], c['inst_id']
    for inst in range(0, len(store.inst_names), 37):
        seg = rows[off[inst]:off[inst + 1]]
        assert len(seg) > 0
        assert np.all(iid[seg] == inst)
        assert np.all(np.diff(seg) > 0)
    assert off[-1] == int(np.sum(iid >= 0))


def test_flag_rules(store):
    aws = store.cloud('aws')
    c = store.columns
    for name, want in (('m6i.2xlarge', True), ('c7i.large', True),
                       ('m5.large', False), ('p3.2xlarge', False)):
        row = aws.row_begin + int(aws.inst_first_row[aws.inst_index[name] -
                                                     aws.inst_begin])
        assert bool(c['flags'][row] & _native.F_DEFAULT_FAMILY) is want
    assert rules.azure_instance_family('Standard_D8s_v5') == 'Ds_v5'
    assert rules.azure_instance_family('Standard_NC4as_T4_v3') == 'NCas_T4_v3'
    assert rules.azure_instance_family('Standard_E4-2ds_v4') == 'E_ds_v4'
    assert rules.azure_is_s_series('Standard_D8s_v5')
    assert not rules.azure_is_s_series('Standard_D8_v5')
    gcp = store.cloud('gcp')
    row = gcp.row_begin + int(
        gcp.inst_first_row[gcp.inst_index['a2-highgpu-8g'] - gcp.inst_begin])
    assert (c['flags'][row] >> 8) == rules.GCP_GROUP_IDS[('A100', 8)]


def test_zone_map_summarises_every_chunk(store):
    from skypilot_b200.catalog import store as store_lib
    c = store.columns
    zm = c['zone_map']
    zr = _native.ZONE_ROWS
    assert len(zm) == store.n_rows // zr
    lib = _native.load()
    for v in (0.0, 0.0001, 0.526, 3.06, 98.32, 1e6, -1.5):
        assert int(store_lib.price_keys(np.array([v]))[0]) == (
            lib.skyopt_price_key(v))
    keys_sorted = store_lib.price_keys(np.array([-2.0, -1.0, 0.0, 0.5, 7.0]))
    assert list(keys_sorted) == sorted(keys_sorted)
    none = np.uint64(0xFFFFFFFFFFFFFFFF)
    for z in range(0, len(zm), 7):
        rows = slice(z * zr, (z + 1) * zr)
        fl = c['flags'][rows]
        valid = (fl & _native.F_VALID) != 0
        assert zm['flags_or'][z] == int(
            np.bitwise_or.reduce(np.where(valid, fl & 0xFF, 0)))
        sig = 0
        for k, ok in zip(c['acc_key'][rows], valid):
            if ok and k != _native.NONE16:
                sig |= 1 << (int(k) % 64)
        assert (int(zm['sig_hi'][z]) << 32 | int(zm['sig_lo'][z])) == sig
        for col, name in enumerate(('price', 'spot_price')):
            p = c[name][rows]
            ok = valid & ~np.isnan(p)
            want = (store_lib.price_keys(p[ok]).min() if ok.any() else none)
            assert zm['min_key'][z, col] == want


# ---- constraint vectors ------------------------------------------------------
def test_cpus_memory_parsing():
    assert engine.parse_cpus('8+') == (_native.OP_GE, 8.0)
    assert engine.parse_cpus('4') == (_native.OP_EQ, 4.0)
    assert engine.parse_cpus(None) == (_native.OP_NONE, 0.0)
    assert engine.parse_memory('16+') == (_native.OP_GE, 16.0)
    assert engine.parse_memory('4x') == (_native.OP_RATIO, 4.0)
    assert engine.parse_memory('32') == (_native.OP_EQ, 32.0)
    with pytest.raises(ValueError):
        engine.parse_cpus('many')


def test_accelerator_sets_follow_pandas_semantics(store):
    exact, fuzzy, strict = engine.accelerator_sets(store, 'a100', 8)
    names = lambda words: sorted(  # noqa: E731
        store.acc_keys[k] for k in range(len(store.acc_keys))
        if (words[k >> 5] >> (k & 31)) & 1)
    assert names(exact) == [('A100', 8.0)]
    assert names(strict) == [('A100', 8.0)]
    got = names(fuzzy)
    assert ('A100-80GB', 8.0) in got and ('A100', 16.0) in got
    assert all('a100' in n.lower() and c >= 8 for n, c in got)
    exact, _, strict = engine.accelerator_sets(store, 'A10', 0.17)
    assert names(exact) == [('A10', 0.167)]  # |count - c| <= 0.01
    assert names(strict) == []               # count == c


def test_problem_for_a_chain(store):
    del store
    sc = next(s for s in scenarios.basic_scenarios()
              if s['name'] == 'chain2_egress_big')
    dag, tasks = runner.build_dag(sc)
    import networkx as nx
    from skypilot_b200.optimizer import Optimizer
    graph = dag.get_graph()
    topo = list(nx.topological_sort(graph))
    problem = Optimizer._state_problem(graph, topo, True, [], True)  # pylint: disable=protected-access
    p = problem.builder.pack()
    # T4 on 4 clouds (GCP needs a gate + a host query) and cpus=8+ on 4
    assert (p.n_tasks, p.n_slots, p.n_queries, p.n_dags) == (2, 8, 9, 1)
    assert list(p.tasks['n_parents'][:2]) == [0, 1]
    aws_i = problem.builder.store.cloud_index['aws']
    assert p.tariffs[aws_i] == pytest.approx(
        sky.AWS().get_egress_cost(500))
    q_cpu = p.queries[p.slots['query'][4]]
    assert q_cpu['cpus_op'] == _native.OP_GE and q_cpu['cpus'] == 8.0
    assert q_cpu['mem_op'] == _native.OP_RATIO and q_cpu['mem'] == 4.0
    assert q_cpu['flags_require'] & _native.F_DEFAULT_FAMILY
    assert math.isinf(q_cpu['max_price'])
    del tasks


# ---- Resources / Dag -----------------------------------------------------------
def _res(kwargs):
    kwargs = dict(kwargs)
    kwargs['cloud'] = registry.CLOUD_REGISTRY.from_str(kwargs['cloud'])
    return sky.Resources(**kwargs)


@pytest.mark.parametrize('cand,blocked,expected', rv.BLOCKED_CASES)
def test_should_be_blocked_by(cand, blocked, expected):
    assert _res(cand).should_be_blocked_by(_res(blocked)) is expected


@pytest.mark.parametrize('edges,n,expected', rv.CHAIN_CASES)
def test_dag_is_chain(edges, n, expected):
    with sky.Dag() as dag:
        tasks = [sky.Task(f't{i}') for i in range(n)]
        for u, v in edges:
            tasks[u] >> tasks[v]  # pylint: disable=pointless-statement
    assert dag.is_chain() is expected


def test_resources_parsing_and_copy():
    r = sky.Resources(accelerators='V100:4', cpus='8+', memory=32,
                      use_spot=True)
    assert r.accelerators == {'V100': 4}
    assert r.cpus == '8+' and r.memory == '32'
    c = r.copy(cloud=sky.AWS(), instance_type='p3.8xlarge', region='us-east-1')
    assert c.is_launchable() and c.use_spot and c.region == 'us-east-1'
    assert not r.is_launchable()
    assert sky.Resources(accelerators='A10:0.5').accelerators == {'A10': 0.5}
    assert sky.Resources(infra='aws/us-east-1/us-east-1a').zone == 'us-east-1a'
    with pytest.raises(ValueError):
        sky.Resources(cpus='-1')
    with pytest.raises(ValueError):
        sky.Resources(accelerators='V100:x')


def test_validate_canonicalises_and_infers(store):
    del store
    r = sky.Resources(accelerators='a100-80gb:8')
    r.validate()
    assert r.accelerators == {'A100-80GB': 8}
    r = sky.Resources(instance_type='p3.8xlarge')
    r.validate()
    assert isinstance(r.cloud, sky.AWS)
    r = sky.Resources(cloud=sky.AWS(), zone='us-east-1a')
    r.validate()
    assert r.region == 'us-east-1'
    with pytest.raises(ValueError):
        sky.Resources(cloud=sky.AWS(), region='mars-1').validate()


def test_egress_tariffs_match_the_oracle():
    from oracle import optimizer_oracle as oo
    for cloud, name in ((sky.AWS(), 'aws'), (sky.GCP(), 'gcp'),
                        (sky.Azure(), 'azure'), (sky.Lambda(), 'lambda')):
        for g in (0.5, 1, 80, 1024, 5000, 10240, 20000, 51200, 60000, 153600,
                  200000):
            assert cloud.get_egress_cost(g) == oo.egress_tariff(name, g)


def test_dummy_nodes_round_trip():
    from skypilot_b200.optimizer import Optimizer
    with sky.Dag() as dag:
        a, b = sky.Task('a'), sky.Task('b')
        a >> b  # pylint: disable=pointless-statement
    Optimizer._add_dummy_source_sink_nodes(dag)  # pylint: disable=protected-access
    assert len(dag.tasks) == 4 and dag.is_chain()
    Optimizer._remove_dummy_source_sink_nodes(dag)  # pylint: disable=protected-access
    assert dag.tasks == [a, b]


# ---------------------------------------------------------------------------
# columnar catalog cache (CatalogStore.save / load / from_directory)
def _write_catalog_dir(tmp_path, frames):
    for cloud, df in frames.items():
        (tmp_path / cloud).mkdir()
        df.to_csv(tmp_path / cloud / 'vms.csv', index=False)


def _assert_same_store(a, b):
    assert set(a.columns) == set(b.columns)
    for k, v in a.columns.items():
        w = b.columns[k]
        if v is None:
            assert w is None
            continue
        assert v.dtype == w.dtype and v.shape == w.shape, k
        assert v.tobytes() == w.tobytes(), k  # NaN payloads included
    assert a.acc_keys == b.acc_keys and a.inst_names == b.inst_names
    assert a.n_rows == b.n_rows and a.max_group_rows == b.max_group_rows
    for ta, tb in zip(a.clouds, b.clouds):
        assert ta.name == tb.name
        assert ta.region_names == tb.region_names
        assert ta.zone_names == tb.zone_names
        assert ta.zone_region == tb.zone_region
        assert ta.inst_index == tb.inst_index
        assert ta.gpu_info_unique == tb.gpu_info_unique
        assert ta.gpu_info_any_nan == tb.gpu_info_any_nan


def test_catalog_cache_round_trip(tmp_path):
    from skypilot_b200 import synth
    from skypilot_b200.catalog.store import CatalogStore
    frames = synth.make_catalogs(seed=5, n_rows=3000,
                                 clouds=['aws', 'gcp', 'azure', 'lambda'])
    _write_catalog_dir(tmp_path, frames)
    parsed = CatalogStore.from_directory(str(tmp_path), use_cache=False)
    first = CatalogStore.from_directory(str(tmp_path))    # writes the cache
    cached = CatalogStore.from_directory(str(tmp_path))   # reads it
    caches = os.listdir(tmp_path / '.skyopt_cache')
    assert len(caches) == 1
    _assert_same_store(parsed, first)
    _assert_same_store(parsed, cached)
    # metadata look-ups work on the per-type frame of a cached store
    row_a = parsed.instance_row('aws', 'p3.2xlarge')
    row_b = cached.instance_row('aws', 'p3.2xlarge')
    for col in ('InstanceType', 'AcceleratorName', 'AcceleratorCount', 'vCPUs',
                'MemoryGiB', 'Arch'):
        assert row_a[col] == row_b[col]


def test_catalog_cache_is_invalidated_by_an_edited_csv(tmp_path):
    from skypilot_b200 import synth
    from skypilot_b200.catalog.store import CatalogStore
    frames = synth.make_catalogs(seed=6, n_rows=2000, clouds=['aws'])
    _write_catalog_dir(tmp_path, frames)
    CatalogStore.from_directory(str(tmp_path))
    df = frames['aws'].copy()
    df.loc[0, 'Price'] = 123.456
    df.to_csv(tmp_path / 'aws' / 'vms.csv', index=False)
    st = os.stat(tmp_path / 'aws' / 'vms.csv')
    os.utime(tmp_path / 'aws' / 'vms.csv', ns=(st.st_atime_ns,
                                                st.st_mtime_ns + 10**9))
    again = CatalogStore.from_directory(str(tmp_path))
    assert again.columns['price'][0] == 123.456
    assert len(os.listdir(tmp_path / '.skyopt_cache')) == 2


# ---------------------------------------------------------------------------
# host-side statement memos (plan templates, per-task replay) never change
# what is sent to the device
def _packed_bytes(dag):
    import networkx as nx
    from skypilot_b200 import optimizer as opt_lib
    graph = dag.get_graph()
    topo = list(nx.topological_sort(graph))
    problem = opt_lib.Optimizer._state_problem(graph, topo, True, [],  # pylint: disable=protected-access
                                               dag.is_chain())
    p = problem.builder.pack()
    return (p.queries[:p.n_queries].tobytes(), p.slots[:p.n_slots].tobytes(),
            p.tasks[:p.n_tasks].tobytes(), p.n_slots,
            [i.cloud.canonical_name() for i in problem.slot_info])


def _chain(specs, num_nodes=1):
    import skypilot_b200 as sky
    with sky.Dag() as dag:
        prev = None
        for i, spec in enumerate(specs):
            t = sky.Task(f't{i}', num_nodes=num_nodes).set_resources(
                sky.Resources(**spec))
            t.set_outputs('s3://bucket/x', estimated_size_gigabytes=10 * (i + 1))
            if prev is not None:
                prev >> t  # pylint: disable=pointless-statement
            prev = t
    return dag


def test_statement_memos_are_transparent(store):
    import skypilot_b200 as sky
    specs = [dict(accelerators='V100'), dict(cpus='8+'),
             dict(accelerators='T4', use_spot=True), dict(memory='32+')]
    dag = _chain(specs)
    first = _packed_bytes(dag)
    assert _packed_bytes(dag) == first            # replayed from the memos
    assert _packed_bytes(_chain(specs)) == first  # fresh objects, same request
    # a changed request is stated afresh
    dag.tasks[1].set_resources(sky.Resources(cpus='16+', memory='64+'))
    specs2 = list(specs)
    specs2[1] = dict(cpus='16+', memory='64+')
    assert _packed_bytes(dag) == _packed_bytes(_chain(specs2))
    # so is a changed node count
    for t in dag.tasks:
        t.num_nodes = 3
    assert _packed_bytes(dag) == _packed_bytes(_chain(specs2, num_nodes=3))
    # and a changed set of enabled clouds
    try:
        sky.check.set_enabled_clouds(['aws', 'gcp'])
        limited = _packed_bytes(dag)
        assert limited == _packed_bytes(_chain(specs2, num_nodes=3))
        assert set(limited[4]) <= {'aws', 'gcp'}
    finally:
        sky.check.set_enabled_clouds(sky.check.ALL_CATALOG_CLOUDS)
    full = _packed_bytes(dag)
    assert full == _packed_bytes(_chain(specs2, num_nodes=3))
    # dropping the memos changes nothing either
    sky.catalog.clear_request_level_cache()
    assert _packed_bytes(dag) == full


def test_preferred_regions_follow_the_clouds_rule():
    """`us_first` on the device reads the per-region flag computed at ingest:
    'us-*' names by default (aws_catalog.py:327-336), names containing 'SCP'
    on SCP (scp_catalog.py:118-126)."""
    store = runner.activate_catalog(scenarios.CATALOGS['scp4k'])
    cols = store.columns
    off = cols['cloud_region_offsets']
    for i, table in enumerate(store.clouds):
        flags = [int(f) for f in cols['region_is_us'][off[i]:off[i + 1]]]
        if table.name == 'scp':
            want = [int('SCP' in n) for n in table.region_names]
        else:
            want = [int(n.startswith('us-')) for n in table.region_names]
        assert flags == want, table.name


def test_free_clouds_are_stated_with_zero_hours():
    """vSphere instances cost 0.0 per hour whatever the catalog says
    (vsphere.py:128-135): their slots carry `hours = 0`, also when the task
    is replayed from the statement memo; the other clouds' slots do not."""
    import networkx as nx
    import numpy as np
    from skypilot_b200 import optimizer as opt_lib
    runner.activate_catalog(scenarios.CATALOGS['vsphere3k'])
    dag = _chain([dict(accelerators='V100'), dict(cpus='8+')])
    for t in dag.tasks:
        t.set_time_estimator(lambda r: 7200)
    graph = dag.get_graph()
    topo = list(nx.topological_sort(graph))
    packs = []
    for _ in range(2):  # stated afresh, then replayed
        problem = opt_lib.Optimizer._state_problem(  # pylint: disable=protected-access
            graph, topo, True, [], True)
        p = problem.builder.pack()
        slots = p.slots[:p.n_slots]
        names = [i.cloud.canonical_name() for i in problem.slot_info]
        assert 'vsphere' in names and 'aws' in names
        for s, name in zip(slots, names):
            assert s['hours'] == (0.0 if name == 'vsphere' else 2.0), name
            assert s['time_value'] == 7200.0
        packs.append(slots.tobytes())
    assert packs[0] == packs[1]
    assert np.isclose(
        sky.clouds.Vsphere().instance_type_to_hourly_cost('cpu_8', False), 0.0)


def test_no_cloud_is_enabled_until_someone_says_so(store):
    """check.py: nothing configured -> NoCloudAccessError (never 'every cloud
    of the catalog'); an explicit list, a provider callback, or the explicit
    catalog-wide opt-in enable clouds."""
    from skypilot_b200 import check
    sky.catalog.set_store(store)
    try:
        check.set_enabled_clouds(None)
        check.set_enabled_clouds_provider(None)
        assert check.get_cached_enabled_clouds_or_refresh() == []
        with pytest.raises(sky.exceptions.NoCloudAccessError):
            check.get_cached_enabled_clouds_or_refresh(
                raise_if_no_cloud_access=True)
        check.set_enabled_clouds_provider(lambda **kw: ['AWS', 'gcp'])
        assert [str(c) for c in check.get_cached_enabled_clouds_or_refresh()
               ] == ['AWS', 'GCP']
        check.set_enabled_clouds(['azure'])  # an explicit list wins
        assert [str(c) for c in check.get_cached_enabled_clouds_or_refresh()
               ] == ['Azure']
        check.set_enabled_clouds(check.ALL_CATALOG_CLOUDS)
        assert len(check.get_cached_enabled_clouds_or_refresh()) == len(
            store.clouds)
    finally:
        check.set_enabled_clouds_provider(None)
        check.set_enabled_clouds(check.ALL_CATALOG_CLOUDS)


def test_image_tags(store):
    import pandas as pd
    sky.catalog.set_store(store)
    store.set_images('aws', pd.DataFrame({
        'Tag': ['skypilot:gpu-ubuntu-2004', 'skypilot:gpu-ubuntu-2004', 'x'],
        'Region': ['us-east-1', 'us-west-2', 'us-east-1'],
        'ImageId': ['ami-1', 'ami-2', None]
    }))
    cat = sky.catalog
    assert cat.get_image_id_from_tag('skypilot:gpu-ubuntu-2004', 'US-EAST-1',
                                     clouds='aws') == 'ami-1'
    assert cat.get_image_id_from_tag('x', 'us-east-1', clouds='aws') is None
    assert not cat.is_image_tag_valid('x', 'us-east-1', clouds='aws')
    assert cat.is_image_tag_valid('skypilot:gpu-ubuntu-2004', None,
                                  clouds='aws')
    with pytest.raises(AssertionError, match='Multiple images'):
        cat.get_image_id_from_tag('skypilot:gpu-ubuntu-2004', None,
                                  clouds='aws')
    assert cat.get_image_id_from_tag('nope', None, clouds='gcp') is None


def test_timeline_hook(tmp_path, monkeypatch):
    from skypilot_b200.utils import timeline
    path = tmp_path / 'trace.json'
    monkeypatch.setenv('SKYPILOT_TIMELINE_FILE_PATH', str(path))

    @timeline.event
    def work(x):
        return x + 1

    with timeline.Event('outer', 'msg'):
        assert work(1) == 2
    timeline.save()
    import json
    events = json.loads(path.read_text())['traceEvents']
    names = [(e['name'], e['ph']) for e in events[-4:]]
    assert names == [('outer', 'B'), (f'{__name__}.test_timeline_hook.<locals>.work', 'B'),
                     (f'{__name__}.test_timeline_hook.<locals>.work', 'E'), ('outer', 'E')]
    assert sky.Optimizer.optimize.__wrapped__ is not None


# ---- region allow-lists (sky/resources.py:1210-1246) ------------------------
def test_ssh_proxy_config_restricts_regions(tmp_path):
    from skypilot_b200 import skypilot_config
    gen = skypilot_config.generation()
    try:
        skypilot_config.set_config(
            {'aws': {'ssh_proxy_command': {'us-east-1': 'ssh a',
                                           'eu-west-1': 'ssh b'}},
             'gcp': {'ssh_proxy_command': 'ssh -W %h:%p jump'}})
        assert skypilot_config.generation() > gen
        assert skypilot_config.allowed_regions_by_ssh_proxy('aws') == {
            'us-east-1', 'eu-west-1'}
        # one command for every region restricts nothing; unset neither
        assert skypilot_config.allowed_regions_by_ssh_proxy('gcp') is None
        assert skypilot_config.allowed_regions_by_ssh_proxy('azure') is None
        r = sky.Resources(cloud=sky.clouds.AWS(),
                          image_id={'us-east-1': 'ami-1', 'us-west-2': 'ami-2'})
        assert r.allowed_region_names() == {'us-east-1'}
        assert sky.Resources(cloud=sky.clouds.GCP()).allowed_region_names() \
            is None
        # a config file is read like the reference's ~/.sky/config.yaml
        path = tmp_path / 'config.yaml'
        path.write_text('aws:\n  ssh_proxy_command:\n    us-west-2: ssh c\n')
        skypilot_config.load(str(path))
        assert skypilot_config.allowed_regions_by_ssh_proxy('aws') == {
            'us-west-2'}
    finally:
        skypilot_config.set_config(None)
    assert skypilot_config.allowed_regions_by_ssh_proxy('aws') is None


def test_config_change_invalidates_pinned_plans(store):
    """Plans pinned on a Resources object are not replayed after the
    SkyPilot config changed (the region allow-list is part of the plan)."""
    from skypilot_b200 import engine, skypilot_config
    sky.catalog.set_store(store)
    r = sky.Resources(cloud=sky.clouds.AWS(), cpus='4+')
    aws = sky.clouds.AWS()
    try:
        plan_a, _ = aws.plan_cached(engine.ProblemBuilder(store), r)
        rec_a = r.__dict__['_plan_templates'][1][(sky.clouds.AWS, False)][1]
        skypilot_config.set_config(
            {'aws': {'ssh_proxy_command': {'us-east-1': 'ssh a'}}})
        plan_b, _ = aws.plan_cached(engine.ProblemBuilder(store), r)
        rec_b = r.__dict__['_plan_templates'][1][(sky.clouds.AWS, False)][1]
        assert plan_a is not plan_b
        assert rec_a.slot_recs != rec_b.slot_recs  # region_set differs
    finally:
        skypilot_config.set_config(None)
