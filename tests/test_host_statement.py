"""CPU check of the HOST side of the product: the constraint vectors that
`Cloud.plan_feasible` states for the device.

The device kernels are not run here (the product path has no CPU fallback).
Instead a small numpy model of what the scan kernel computes for ONE query
(filter + first cheapest row, skypilot_b200/csrc/skyopt_kernels.cuh
`score_rows`) turns every stated (task, request, cloud) slot into the
instance type the device will pick, and that is compared with the oracle's
`feasible()` -- which tests/test_oracle.py pins against the unmodified
reference. A divergence means a cloud's rules (defaults, which arguments its
look-ups receive, feature gates) are stated wrongly.
"""
import networkx as nx
import numpy as np
import pytest

from oracle import optimizer_oracle as oo
from skypilot_b200 import _native
from skypilot_b200 import optimizer as opt_lib
from tests import scenario_runner as runner
from tests import scenarios

_SUITES = ['multi6k', 'three4k', 'aws50k', 'multi50k', 'gpuclouds', 'ibm5k',
           'hyperprime', 'latecl', 'fuzz6k', 'oci5k', 'nebvast', 'scp4k', 'fuzzmany', 'vsphere3k', 'seeweb3k', 'shade3k', 'fuzzdag']


def _stage1(cols, sets, q, lo, hi):
    """Rows of [lo, hi) that pass the integer part of query `q`."""
    fl = cols['flags'][lo:hi].astype(np.uint32)
    req = (int(q['flags_require']) | _native.F_VALID) & 0xFF
    ok = (fl & req) == req
    if q['group'] != 0:
        ok &= (fl >> 8) == (int(q['group']) & 0xFF)
    if q['region_id'] >= 0:
        ok &= cols['region_id'][lo:hi] == q['region_id']
    if q['zone_id'] >= 0:
        ok &= cols['zone_id'][lo:hi] == q['zone_id']
    if q['qflags'] & _native.Q_ACC:
        key = cols['acc_key'][lo:hi].astype(np.int64)
        words = sets[int(q['acc_set'])]
        has = key != _native.NONE16
        k = np.where(has, key, 0)
        ok &= has & (((words[k >> 5] >> (k & 31)) & 1) == 1)
    return ok, fl


def _best_row(cols, sets, q, offsets):
    lo, hi = int(offsets[q['cloud']]), int(offsets[q['cloud'] + 1])
    ok, fl = _stage1(cols, sets, q, lo, hi)
    any1 = bool(ok.any())
    f2 = int(q['flags_require2'])
    ok &= (fl & f2) == f2
    vc, mm = cols['vcpus'][lo:hi], cols['mem'][lo:hi]
    if q['cpus_op'] == _native.OP_EQ:
        ok &= vc == q['cpus']
    elif q['cpus_op'] == _native.OP_GE:
        ok &= vc >= q['cpus']
    if q['mem_op'] == _native.OP_EQ:
        ok &= mm == q['mem']
    elif q['mem_op'] == _native.OP_GE:
        ok &= mm >= q['mem']
    elif q['mem_op'] == _native.OP_RATIO:
        ok &= mm >= vc * q['mem']
    if q['disk_op'] != 0:
        total = cols['disk_total'][lo:hi]
        if q['disk_op'] == _native.DISK_GE:
            ok &= total >= q['disk_size']
        else:
            ok &= np.abs(total - q['disk_size']) < 1.0
    price = cols['spot_price' if q['price_col'] else 'price'][lo:hi]
    with np.errstate(invalid='ignore'):
        ok &= price <= q['max_price']  # NaN: False
    if not ok.any():
        return -1, any1
    masked = np.where(ok, price, np.inf)
    return lo + int(np.argmin(masked)), any1  # first minimum: lowest row


def _stated_instances(store, scenario):
    """{(task index, request index, cloud): instance type or None}."""
    dag, tasks = runner.build_dag(scenario)
    O = opt_lib.Optimizer
    O._add_dummy_source_sink_nodes(dag)  # pylint: disable=protected-access
    graph = dag.get_graph()
    topo = [t for t in nx.topological_sort(graph)
            if not opt_lib._is_dummy(t)]  # pylint: disable=protected-access
    problem = O._state_problem(  # pylint: disable=protected-access
        graph, topo, True, [], dag.is_chain())
    packed = problem.builder.pack()
    queries = packed.queries[:packed.n_queries]
    slots = packed.slots[:packed.n_slots]
    sets = np.asarray(packed.acc_sets,
                      dtype=np.uint32).reshape(-1, _native.ACC_SET_WORDS)
    cols = store.columns
    offsets = cols['cloud_row_offsets']
    out = {}
    for s, info in zip(slots, problem.slot_info):
        ti = tasks.index(info.task)
        ri = list(info.task.resources).index(info.resources)
        gate_ok = True
        if s['gate_query'] >= 0:
            _, gate_ok = _best_row(cols, sets, queries[int(s['gate_query'])],
                                   offsets)
        if not gate_ok:
            inst = -1
        elif s['query'] >= 0:
            row, _ = _best_row(cols, sets, queries[int(s['query'])], offsets)
            inst = int(cols['inst_id'][row]) if row >= 0 else -1
        else:
            inst = int(s['inst_id'])  # explicit instance type, -2 = TPU-VM
        name = ('TPU-VM' if inst == -2 else
                store.inst_names[inst] if inst >= 0 else None)
        out[(ti, ri, str(info.cloud).lower())] = name
    return out


def _cases():
    out = []
    for catalog in _SUITES:
        for sc in scenarios.ALL_SUITES[catalog]():
            if any(t.get('resources_kind', 'single') != 'single'
                   for t in sc['tasks']):
                continue
            out.append(pytest.param(catalog, sc, id=f'{catalog}:{sc["name"]}'))
    return out


@pytest.mark.parametrize('catalog,scenario', _cases())
def test_stated_queries_pick_the_oracles_instance(catalog, scenario):
    spec = scenarios.CATALOGS[catalog]
    store = runner.activate_catalog(spec)
    cat = oo.catalog_for(spec)
    try:
        got = _stated_instances(store, scenario)
    except Exception as e:  # pylint: disable=broad-except
        # requests the product rejects while stating them must be ones the
        # oracle finds infeasible everywhere
        got = None
        error = e
    for ti, tspec in enumerate(scenario['tasks']):
        req = oo.normalize_request(cat, tspec['resources'][0])
        clouds = [req['cloud']] if req['cloud'] is not None else cat.enabled
        for cloud in clouds:
            if cloud not in cat.enabled:
                continue
            launchables, _ = oo.feasible(cat, cloud, dict(req, cloud=cloud),
                                         tspec.get('num_nodes', 1))
            want = launchables[0]['instance_type'] if launchables else None
            if got is None:
                assert want is None, (cloud, want, repr(error))
                continue
            have = got.get((ti, 0, cloud))
            if have != want and want is not None and have is not None:
                pytest.fail(f'task {ti} on {cloud}: stated queries pick '
                            f'{have}, the oracle {want}')
            if want is None:
                # a slot may exist and expand to nothing on the device (no
                # region left): only a definite pick must agree
                continue
            if have is None and not oo.regions_with_offering(
                    cat, launchables[0]):
                # stated without a slot because nothing can be offered (IBM
                # spot: ibm.py:88-92) -- no candidates either way
                continue
            assert have == want, (ti, cloud, have, want)


# ---------------------------------------------------------------------------
# The fast statement path (Cloud.plan_fast) writes the bytes of the generic
# one (_feature_hint + plan_feasible).
def _both_statements(store, cloud, res, num_nodes):
    from skypilot_b200 import engine
    from skypilot_b200.clouds import cloud as cloud_lib
    fast = cloud.plan_fast(store, res, num_nodes)
    plan = cloud_lib.SlotPlan()
    recorder = engine.ProblemBuilder(store)
    plan.hint = cloud._feature_hint(res, num_nodes)  # pylint: disable=protected-access
    if plan.hint is None:
        plan = cloud.plan_feasible(recorder, res)
    return fast, (plan, recorder)


@pytest.mark.parametrize('suite', _SUITES + ['altres', 'regfilter'])
def test_fast_statement_is_byte_identical(suite):
    import skypilot_b200 as sky
    from skypilot_b200 import workloads
    all_suites = dict(scenarios.ALL_SUITES, **scenarios.EXTRA_GOLDEN_SUITES)
    store = runner.activate_catalog(scenarios.CATALOGS[suite])
    enabled = sky.check.get_cached_enabled_clouds_or_refresh()
    n_fast = n_all = 0
    for sc in all_suites[suite]():
        if sc.get('config') is not None:
            continue  # a SkyPilot config always takes the generic path
        try:
            _, tasks = workloads.build_dag(sc)
        except ValueError:
            continue  # an invalid request (e.g. a bad image tag)
        for task in tasks:
            for res in task.resources:
                clouds_list = [res.cloud] if res.cloud is not None else enabled
                for cloud in clouds_list:
                    if not store.has_cloud(cloud.canonical_name()):
                        continue
                    fast, (plan, rec) = _both_statements(
                        store, cloud, res, task.num_nodes)
                    n_all += 1
                    if fast is None:
                        continue
                    n_fast += 1
                    fplan, frec = fast
                    where = (sc['name'], repr(res), repr(cloud))
                    assert fplan.hint is None and plan.hint is None, where
                    for attr in ('slot', 'list_query', 'fuzzy_query',
                                 'gate_query', 'explicit_instance',
                                 'no_fuzzy_when_matched'):
                        assert getattr(fplan, attr) == getattr(plan, attr), (
                            where, attr)
                    assert frec.query_recs == rec.query_recs, where
                    assert frec.slot_recs == rec.slot_recs, where
                    assert (fplan.make is None) == (plan.make is None), where
                    if plan.make is not None and plan.list_query is not None:
                        # same launchable for some instance type of the cloud
                        name = store.cloud(
                            cloud.canonical_name()).inst_names[0]
                        a, b = fplan.make(name, res), plan.make(name, res)
                        assert repr(a) == repr(b) and (
                            a is None or
                            (a.cpus, a.memory, a.accelerators, a.use_spot,
                             a.region, a.zone) ==
                            (b.cpus, b.memory, b.accelerators, b.use_spot,
                             b.region, b.zone)), where
    assert n_all > 0
    if suite in ('multi6k', 'multi50k', 'fuzz6k', 'fuzzmany'):
        assert n_fast > n_all // 2, (n_fast, n_all)  # the path is really taken
