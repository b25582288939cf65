"""GPU tests of the catalog function table (the drop-in boundary at the
catalog layer): the reference's own inline known-answer vectors and the
per-cloud functions, checked against the pandas oracle."""
import numpy as np
import pytest

import skypilot_b200 as sky
from oracle import catalog_oracle as co
from skypilot_b200.catalog import common
from tests import reference_vectors as rv
from tests import scenario_runner as runner
from tests import scenarios

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('cpus,memory,region,zone,expected', rv.AZ_CASES)
def test_cpus_mem_with_az(cpus, memory, region, zone, expected):
    df = rv.az_frame()
    assert common.get_instance_type_for_cpus_mem_impl(
        df, cpus=cpus, memory_gb_or_ratio=memory, region=region,
        zone=zone) == expected


@pytest.mark.parametrize('cpus,memory,region,expected', rv.NO_AZ_CASES)
def test_cpus_mem_no_az(cpus, memory, region, expected):
    df = rv.no_az_frame()
    assert common.get_instance_type_for_cpus_mem_impl(
        df, cpus=cpus, memory_gb_or_ratio=memory, region=region) == expected


def test_hourly_cost_is_python_float():
    df = rv.price_frame()
    for spot, want in ((False, 1.5), (True, 0.5)):
        cost = common.get_hourly_cost_impl(df, 'test-instance', use_spot=spot,
                                           region=None, zone=None)
        assert type(cost) is float and cost == want  # pylint: disable=unidiomatic-typecheck


@pytest.mark.parametrize('local_disk,expected', rv.LOCAL_DISK_CASES)
def test_local_disk_selection(local_disk, expected):
    view = common.filter_with_local_disk(rv.local_disk_frame(), local_disk)
    assert common.get_instance_type_for_cpus_mem_impl(
        view, cpus='1+', memory_gb_or_ratio=None, region=None) == expected


@pytest.fixture(scope='module')
def frames():
    store = runner.activate_catalog(scenarios.CATALOGS['multi6k'])
    return {t.name: t.frame for t in store.clouds}


ACC_CASES = [
    ('aws', 'V100', 1, {}), ('aws', 'T4', 1, {'cpus': '8+'}),
    ('aws', 'A100', 8, {'use_spot': True}), ('aws', 'A100', 3, {}),
    ('aws', 'T4', 1, {'region': 'us-west-2'}),
    ('aws', 'V100', 4, {'max_hourly_cost': 1.0}),
    ('azure', 'A10', 0.5, {}), ('azure', 'V100', 2, {'memory': '200+'}),
    ('lambda', 'A100', 1, {}), ('lambda', 'H100', 3, {}),
    ('aws', 'NoSuch', 1, {}), ('azure', 'K80', 1, {'use_spot': True}),
]


@pytest.mark.parametrize('cloud,acc,count,kw', ACC_CASES)
def test_instance_type_for_accelerator_lists(frames, cloud, acc, count, kw):
    """(sorted instance types, fuzzy candidates) == the pandas oracle's."""
    want = co.instance_type_for_accelerator(frames[cloud], acc, count, **kw)
    got = sky.catalog.get_instance_type_for_accelerator(acc, count,
                                                        clouds=cloud, **kw)
    assert got[0] == want[0]
    assert got[1] == want[1]


@pytest.mark.parametrize('mode', ['queue', 'queue32'])
def test_accelerator_lists_through_the_queue_scan(frames, mode):
    """The sorted list / fuzzy tables come out of the scan kernel; the queue
    form must fill them like the one-tile-per-block form."""
    store = sky.catalog.get_store()
    store.set_scan_mode(mode)
    try:
        for cloud, acc, count, kw in ACC_CASES:
            want = co.instance_type_for_accelerator(frames[cloud], acc, count,
                                                    **kw)
            got = sky.catalog.get_instance_type_for_accelerator(
                acc, count, clouds=cloud, **kw)
            assert (got[0], got[1]) == (want[0], want[1]), (cloud, acc, kw)
    finally:
        store.set_scan_mode('auto')


@pytest.mark.parametrize('cloud,instance_type,spot', [
    ('aws', 'p3.2xlarge', False), ('aws', 'g4dn.xlarge', True),
    ('gcp', 'n2-standard-8', False), ('gcp', 'n1-highmem-8', True),
    ('azure', 'Standard_NC6s_v3', True), ('lambda', 'gpu_1x_a100', False),
])
def test_region_zones_and_hourly_cost(frames, cloud, instance_type, spot):
    df = frames[cloud]
    want = co.region_zones(df[df['InstanceType'] == instance_type], spot)
    if cloud in ('aws', 'lambda'):
        want = co.us_first(want)
    got = sky.catalog.get_region_zones_for_instance_type(instance_type, spot,
                                                         clouds=cloud)
    assert [(r.name, None if r.zones is None else [z.name for z in r.zones])
            for r in got] == want
    for region, zones in want[:3]:
        assert sky.catalog.get_hourly_cost(
            instance_type, spot, region, None, clouds=cloud) == (
                co.hourly_cost(df, instance_type, spot, region, None))
        if zones:
            assert sky.catalog.get_hourly_cost(
                instance_type, spot, region, zones[0], clouds=cloud) == (
                    co.hourly_cost(df, instance_type, spot, region, zones[0]))
    assert sky.catalog.get_hourly_cost(
        instance_type, spot, None, None,
        clouds=cloud) == co.hourly_cost(df, instance_type, spot, None, None)


@pytest.mark.parametrize('acc,count,spot', [('T4', 1, False), ('V100', 2, True),
                                            ('A100', 8, False),
                                            ('tpu-v3-8', 1, True)])
def test_gcp_accelerator_functions(frames, acc, count, spot):
    df = frames['gcp']
    want = co.region_zones(co.gcp_accelerator_rows(df, acc, count, None), spot)
    got = sky.catalog.get_region_zones_for_accelerators(acc, count, spot,
                                                        clouds='gcp')
    assert [(r.name, [z.name for z in r.zones]) for r in got] == want
    for region, zones in want[:2]:
        w = co.gcp_accelerator_hourly_cost(df, acc, count, spot, region,
                                           zones[0])
        g = sky.catalog.get_accelerator_hourly_cost(acc, count, spot, region,
                                                    zones[0], clouds='gcp')
        assert g == pytest.approx(float(w), rel=1e-12)
    inst, fuzzy = sky.catalog.get_instance_type_for_accelerator(
        acc, count, clouds='gcp') if not acc.startswith('tpu') else ([None],
                                                                     [])
    if not acc.startswith('tpu'):
        w_inst, w_fuzzy = co.gcp_instance_type_for_accelerator(
            df, acc, count, None, None, False, None, None, None)
        assert inst == w_inst and fuzzy == w_fuzzy


def test_feasible_resources_object_path(frames):
    """Cloud.get_feasible_launchable_resources / _fill_in_launchable_resources
    (the non-fused API) agree with the fused optimizer's candidate tables."""
    del frames
    from skypilot_b200 import optimizer as opt_lib
    task = sky.Task('t')
    task.set_resources(sky.Resources(accelerators='T4'))
    launchable, per_cloud, fuzzy, hints = (
        opt_lib._fill_in_launchable_resources(task, None, quiet=True))  # pylint: disable=protected-access
    objs = list(launchable.values())[0]
    with sky.Dag() as dag:
        dag.add(task)
    opt_lib.Optimizer._add_dummy_source_sink_nodes(dag)  # pylint: disable=protected-access
    try:
        import networkx as nx
        topo = list(nx.topological_sort(dag.get_graph()))
        cost_map, _ = opt_lib.Optimizer._estimate_nodes_cost_or_time(topo)  # pylint: disable=protected-access
    finally:
        opt_lib.Optimizer._remove_dummy_source_sink_nodes(dag)  # pylint: disable=protected-access
    fused = list(cost_map[task].keys())
    key = lambda r: (str(r.cloud), r.instance_type, r.region, r.zone)  # noqa: E731
    assert [key(r) for r in objs] == [key(r) for r in fused]
    for r, v in cost_map[task].items():
        assert r.get_cost(3600) == pytest.approx(v, rel=1e-12)
    assert not fuzzy and not hints
    assert all(len(v) >= 1 for v in per_cloud.values())
    bad = sky.Task('bad')
    bad.set_resources(sky.Resources(accelerators='A100:3'))
    _, _, fuzzy, _ = opt_lib._fill_in_launchable_resources(bad, None,  # pylint: disable=protected-access
                                                            quiet=True)
    assert 'A100:8' in fuzzy or 'A100:4' in fuzzy


def test_batch_of_independent_dags_matches_single_calls():
    """cfg5 shape: many single-task DAGs in ONE device call == one by one."""
    import networkx as nx
    from skypilot_b200 import engine
    from skypilot_b200 import optimizer as opt_lib
    store = runner.activate_catalog(scenarios.CATALOGS['multi6k'])
    rng = np.random.default_rng(4)
    accs = [None, 'V100', 'T4', 'A100:8', 'L4', 'H100:8', 'A10G', 'K80']
    specs = []
    for _ in range(64):
        spec = {}
        acc = accs[int(rng.integers(len(accs)))]
        if acc:
            spec['accelerators'] = acc
        cpus = [None, '2+', '8+', '32+'][int(rng.integers(4))]
        if cpus:
            spec['cpus'] = cpus
        if rng.uniform() < 0.3:
            spec['use_spot'] = True
        specs.append(spec)
    b = engine.ProblemBuilder(store)
    problems, singles = [], []
    for spec in specs:
        with sky.Dag() as dag:
            t = sky.Task('t')
            t.set_resources(sky.Resources(**spec))
        graph = dag.get_graph()
        problems.append(
            opt_lib.Optimizer._state_problem(graph, [t], True, [], True,  # pylint: disable=protected-access
                                             builder=b))
        try:
            sky.optimize(dag, quiet=True)
            r = t.best_resources
            singles.append((str(r.cloud), r.instance_type, r.region, r.zone))
        except sky.exceptions.ResourcesUnavailableError:
            singles.append(None)
    sol = engine.solve(b)
    assert len(sol.dag) == len(specs)
    for i, prob in enumerate(problems):
        if sol.dag[i]['status'] != 0:
            assert singles[i] is None
            continue
        r = prob.launchable(sol.chosen[i])
        assert (str(r.cloud), r.instance_type, r.region,
                r.zone) == singles[i]
    del nx


def _random_single_task_dags(n, seed):
    rng = np.random.default_rng(seed)
    accs = [None, 'V100', 'T4', 'A100:8', 'L4', 'H100:8', 'A10G', 'K80',
            'A100:3', 'T4:4']
    dags, tasks = [], []
    for _ in range(n):
        spec = {}
        acc = accs[int(rng.integers(len(accs)))]
        if acc:
            spec['accelerators'] = acc
        cpus = [None, '2+', '8+', '32+'][int(rng.integers(4))]
        if cpus:
            spec['cpus'] = cpus
        mem = [None, '16+', '4x'][int(rng.integers(3))]
        if mem and not acc:
            spec['memory'] = mem
        if rng.uniform() < 0.3:
            spec['use_spot'] = True
        with sky.Dag() as dag:
            t = sky.Task('t')
            t.set_resources(sky.Resources(**spec))
        dags.append(dag)
        tasks.append(t)
    return dags, tasks


def test_optimize_batch_matches_one_by_one():
    """BASELINE config 5 shape: a batch of independent single-task DAGs,
    sharded over every visible GPU, equals optimizing them one at a time."""
    from skypilot_b200 import _native
    runner.activate_catalog(scenarios.CATALOGS['multi6k'])
    dags, tasks = _random_single_task_dags(200, seed=4)
    want = []
    for dag, t in zip(dags, tasks):
        try:
            sky.optimize(dag, quiet=True)
            r = t.best_resources
            want.append((str(r.cloud), r.instance_type, r.region, r.zone))
        except sky.exceptions.ResourcesUnavailableError as e:
            want.append(('unavailable', str(e)))
        t.best_resources = None
    devices = list(range(_native.device_count()))
    out = sky.optimize_batch(dags, devices=devices, return_exceptions=True)
    got = []
    for o, t in zip(out, tasks):
        if isinstance(o, Exception):
            # worded from one scan for the whole batch: same text as the
            # one-DAG path (hints, fuzzy candidates)
            got.append(('unavailable', str(o)))
        else:
            r = t.best_resources
            got.append((str(r.cloud), r.instance_type, r.region, r.zone))
    assert got == want
    assert any(w[0] == 'unavailable' for w in want)
    assert any(w[0] != 'unavailable' for w in want)
    with pytest.raises(sky.exceptions.ResourcesUnavailableError):
        sky.optimize_batch(dags, devices=devices)


def test_optimize_on_a_cache_loaded_catalog(tmp_path):
    """A store read back from the columnar cache plans like the parsed one."""
    from skypilot_b200 import synth
    from skypilot_b200.catalog.store import CatalogStore
    spec = dict(scenarios.CATALOGS['multi6k'])
    spec.pop('enabled', None)
    frames = synth.make_catalogs(**spec)
    for cloud, df in frames.items():
        (tmp_path / cloud).mkdir()
        df.to_csv(tmp_path / cloud / 'vms.csv', index=False)
    names = list(frames.keys())
    plans = []
    for _ in range(2):  # first: parse + write the cache, second: cache hit
        store = CatalogStore.from_directory(str(tmp_path), clouds=names)
        sky.catalog.set_store(store)
        sky.check.set_enabled_clouds(sky.check.ALL_CATALOG_CLOUDS)
        dag, tasks = runner.build_dag(scenarios.basic_scenarios()[0])
        sky.Optimizer.optimize(dag, quiet=True)
        plans.append([runner.res_record(t.best_resources) for t in tasks])
        listing = sky.catalog.list_accelerators(clouds=names,
                                                name_filter='V100')
        plans.append({k: [tuple(i) for i in v] for k, v in listing.items()})
    assert plans[0] == plans[2]
    assert repr(plans[1]) == repr(plans[3])


def _failover_blocked(sky_mod, plan_record):
    """The wildcard a failed launch of `plan_record` adds: zone, else region
    (sky/backends/cloud_vm_ray_backend.py:332-339)."""
    from skypilot_b200.utils import registry
    kw = dict(cloud=registry.CLOUD_REGISTRY.from_str(plan_record['cloud']),
              instance_type=plan_record['instance_type'],
              region=plan_record['region'])
    if plan_record['zone'] is not None:
        kw['zone'] = plan_record['zone']
    r = sky_mod.Resources(**kw)
    r._use_spot_specified = False  # pylint: disable=protected-access
    return r


@pytest.mark.parametrize('name', ['acc_V100', 'acc_A100x8', 'cpu_default', 'cfg2_chain8',
                                  'chain8_spot', 'cfg3_diamond'])
def test_session_follows_full_reoptimisation(name):
    """OptimizerSession (candidate sets resident, re-mask + re-solve) gives
    the plans of a full Optimizer.optimize for a growing blocked list."""
    runner.activate_catalog(scenarios.CATALOGS['multi6k'])
    by_name = {s['name']: s for s in scenarios.basic_scenarios()}
    if name not in by_name:
        pytest.skip(f'no scenario {name}')
    sc = by_name[name]
    dag_a, tasks_a = runner.build_dag(sc)
    dag_b, tasks_b = runner.build_dag(sc)
    blocked = []
    with sky.Optimizer.session(dag_a) as session:
        for step in range(6):
            err_a = err_b = None
            try:
                session.optimize(blocked)
            except sky.exceptions.ResourcesUnavailableError as e:
                err_a = str(e)
            try:
                sky.Optimizer.optimize(dag_b, blocked_resources=blocked,
                                       quiet=True)
            except sky.exceptions.ResourcesUnavailableError as e:
                err_b = str(e)
            assert (err_a is None) == (err_b is None), (step, err_a, err_b)
            if err_a is not None:
                assert err_a == err_b
                break
            plan_a = [runner.res_record(t.best_resources) for t in tasks_a]
            plan_b = [runner.res_record(t.best_resources) for t in tasks_b]
            assert plan_a == plan_b, step
            # the launch of the first task's placement "fails"
            blocked.append(_failover_blocked(sky, plan_a[0]))


def test_concurrent_optimize_calls_share_one_catalog():
    """The C ABI is reentrant: calls from several host threads (the
    reference fans out over a ThreadPool, sky/optimizer.py:1712-1715) borrow
    private streams / workspaces from the same catalog handle."""
    import threading
    runner.activate_catalog(scenarios.CATALOGS['multi6k'])
    cases = [s for s in scenarios.basic_scenarios()
             if s['name'] in ('acc_V100', 'cpu_default', 'cfg2_chain8',
                              'chain8_spot', 'acc_T4', 'mem_4x')]
    want = []
    for sc in cases:
        dag, tasks = runner.build_dag(sc)
        sky.Optimizer.optimize(dag, quiet=True)
        want.append([runner.res_record(t.best_resources) for t in tasks])
    errors = []

    def work(k):
        try:
            for it in range(15):
                i = (k + it) % len(cases)
                dag, tasks = runner.build_dag(cases[i])
                sky.Optimizer.optimize(dag, quiet=True)
                got = [runner.res_record(t.best_resources) for t in tasks]
                assert got == want[i], (k, it)
        except BaseException as e:  # pylint: disable=broad-except
            errors.append(e)

    threads = [threading.Thread(target=work, args=(k,)) for k in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[0]


@pytest.mark.parametrize('seed', range(5))
def test_session_random_blocked_lists(seed):
    """Random wildcards of varying specificity, derived from the current
    plan, accumulate; the session and a full re-optimisation stay in step."""
    from skypilot_b200.utils import registry
    rng = np.random.default_rng(50 + seed)
    runner.activate_catalog(scenarios.CATALOGS['multi6k'])
    by_name = {s['name']: s for s in scenarios.basic_scenarios()}
    sc = by_name[['cfg2_chain8', 'chain3_mixed', 'chain8_spot',
                  'chain2_egress_big', 'cfg2_chain8'][seed]]
    dag_a, tasks_a = runner.build_dag(sc)
    dag_b, tasks_b = runner.build_dag(sc)
    blocked = []
    with sky.Optimizer.session(dag_a) as session:
        for step in range(8):
            err_a = err_b = None
            try:
                session.optimize(blocked)
            except sky.exceptions.ResourcesUnavailableError as e:
                err_a = str(e)
            try:
                sky.Optimizer.optimize(dag_b, blocked_resources=blocked,
                                       quiet=True)
            except sky.exceptions.ResourcesUnavailableError as e:
                err_b = str(e)
            assert err_a == err_b, step
            if err_a is not None:
                break
            plan_a = [runner.res_record(t.best_resources) for t in tasks_a]
            plan_b = [runner.res_record(t.best_resources) for t in tasks_b]
            assert plan_a == plan_b, step
            rec = plan_a[int(rng.integers(len(plan_a)))]
            kind = int(rng.integers(4))
            kw = dict(cloud=registry.CLOUD_REGISTRY.from_str(rec['cloud']))
            if kind >= 1:
                kw['region'] = rec['region']
            if kind >= 2:
                kw['instance_type'] = rec['instance_type']
            if kind >= 3 and rec['zone'] is not None:
                kw['zone'] = rec['zone']
            r = sky.Resources(**kw)
            r._use_spot_specified = False  # pylint: disable=protected-access
            blocked.append(r)
