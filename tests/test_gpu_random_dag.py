"""Property test in the spirit of the reference's
tests/test_optimizer_random_dag.py (:20-183): random general DAGs, the
optimizer's objective must equal the exhaustive optimum computed with the
independent host functions Optimizer._compute_total_cost / _time."""
import itertools

import networkx as nx
import numpy as np
import pytest

import skypilot_b200 as sky
from skypilot_b200 import optimizer as opt_lib
from tests import scenario_runner as runner
from tests import scenarios

pytestmark = pytest.mark.gpu

_POOL = [
    {'accelerators': 'V100'}, {'accelerators': 'T4'}, {'accelerators': 'L4'},
    {'accelerators': 'A100:8'}, {'accelerators': 'H100:8'}, {'cpus': '8+'},
    {'cpus': '32+', 'memory': '128+'}, {'accelerators': 'A10G'},
    {'accelerators': 'T4:4', 'use_spot': True}, {'cpus': '4+', 'use_spot': True},
    {'accelerators': 'K80'}, {'memory': '64+'},
]


def _random_dag(seed: int, n_tasks: int):
    rng = np.random.default_rng(seed)
    with sky.Dag() as dag:
        tasks = []
        for i in range(n_tasks):
            t = sky.Task(f't{i}', num_nodes=int(rng.integers(1, 3)))
            k = 2 if rng.uniform() < 0.3 else 1
            specs = [dict(_POOL[int(j)])
                     for j in rng.choice(len(_POOL), size=k, replace=False)]
            res = [sky.Resources(**s) for s in specs]
            t.set_resources(set(res) if k > 1 else res[0])
            if rng.uniform() < 0.8:
                t.set_outputs('x', float(rng.choice([0.5, 20, 300, 2500])))
            if rng.uniform() < 0.5:
                secs = int(rng.integers(600, 9000))
                t.set_time_estimator(lambda r, s=secs: s)
            tasks.append(t)
        for j in range(1, n_tasks):
            for i in rng.choice(j, size=int(rng.integers(1, min(j, 2) + 1)),
                                replace=False):
                dag.add_edge(tasks[int(i)], tasks[j])
    return dag, tasks


def _exhaustive(graph, topo, cost_map, minimize_cost):
    """Optimum over one representative per (task, cloud): within a cloud the
    cheapest candidate dominates because egress depends on clouds only."""
    options = []
    for t in topo:
        best = {}
        for r, v in cost_map[t].items():
            c = str(r.cloud)
            if c not in best or v < cost_map[t][best[c]]:
                best[c] = r
        options.append(list(best.values()))
    fn = (opt_lib.Optimizer._compute_total_cost
          if minimize_cost else opt_lib.Optimizer._compute_total_time)
    best_val = np.inf
    for combo in itertools.product(*options):
        plan = dict(zip(topo, combo))
        best_val = min(best_val, fn(graph, topo, plan))
    return best_val


@pytest.mark.parametrize('minimize_cost', [True, False])
@pytest.mark.parametrize('seed', range(6))
def test_random_dag_objective_is_optimal(seed, minimize_cost):
    runner.activate_catalog(scenarios.CATALOGS['three4k'])
    dag, tasks = _random_dag(seed, n_tasks=5)
    O = opt_lib.Optimizer
    O._add_dummy_source_sink_nodes(dag)
    try:
        graph = dag.get_graph()
        topo = list(nx.topological_sort(graph))
        plan = O._optimize_dag(dag, minimize_cost, quiet=True)
        fn = O._compute_total_cost if minimize_cost else O._compute_total_time
        objective = fn(graph, topo, plan)
        cost_map, _ = O._estimate_nodes_cost_or_time(topo, minimize_cost)
        want = _exhaustive(graph, topo, cost_map, minimize_cost)
        assert objective == pytest.approx(want, rel=1e-9, abs=1e-9)
        # the stand-alone operators on the same cost map agree
        if dag.is_chain():
            _, obj2 = O._optimize_by_dp(topo, cost_map, minimize_cost)
        else:
            _, obj2 = O._optimize_by_ilp(graph, topo, cost_map, minimize_cost)
        assert obj2 == pytest.approx(want, rel=1e-9, abs=1e-9)
    finally:
        O._remove_dummy_source_sink_nodes(dag)
    del tasks


@pytest.mark.parametrize('name', ['cfg2_chain8', 'chain2_egress_big',
                                  'chain2_time', 'cfg3_diamond',
                                  'diamond_time'])
def test_standalone_dp_and_dag_search_match_the_fused_path(name):
    """_optimize_by_dp / _optimize_by_ilp on the candidate tables reproduce
    the plan of the fused device call."""
    runner.activate_catalog(scenarios.CATALOGS['multi6k'])
    sc = next(s for s in scenarios.basic_scenarios() if s['name'] == name)
    minimize_cost = sc.get('minimize', 'cost') == 'cost'
    dag, tasks = runner.build_dag(sc)
    O = opt_lib.Optimizer
    O._add_dummy_source_sink_nodes(dag)
    try:
        graph = dag.get_graph()
        topo = list(nx.topological_sort(graph))
        fused = O._optimize_dag(dag, minimize_cost, quiet=True)
        cost_map, _ = O._estimate_nodes_cost_or_time(topo, minimize_cost)
        if dag.is_chain():
            plan, _ = O._optimize_by_dp(topo, cost_map, minimize_cost)
        else:
            plan, _ = O._optimize_by_ilp(graph, topo, cost_map, minimize_cost)
        key = lambda r: (str(r.cloud), r.instance_type, r.region, r.zone)  # noqa: E731
        assert [key(plan[t]) for t in tasks] == [key(fused[t]) for t in tasks]
    finally:
        O._remove_dummy_source_sink_nodes(dag)


@pytest.mark.parametrize('name', ['diamonds19', 'stages26', 'sparse18',
                                  'join22'])
def test_large_general_dag_matches_reference(name):
    """General DAGs of 18-26 tasks: beyond the device enumeration, placed by
    the exact elimination of dag_solver.py. The candidate tables and the
    optimum are the reference's (tests/golden/bigdag.json: its tables, the
    frontier DP of oracle/dag_oracle.py over them)."""
    payload = runner.load_golden('bigdag')
    assert payload['catalog'] == scenarios.CATALOGS['bigdag']
    golden = next(r for r in payload['records'] if r['name'] == name)
    sc = next(s for s in scenarios.big_dag_scenarios() if s['name'] == name)
    runner.activate_catalog(payload['catalog'])
    got = runner.run_scenario(sc)
    assert 'error' not in got, got
    # identical ordered candidate tables; the optimum within 1e-6 relative
    # (plans among equal-cost optima are not compared)
    plan_free = dict(golden)
    plan_free['plan'] = got['plan']
    diffs = runner.compare(plan_free, got)
    assert not diffs, diffs
    assert runner.close(got['objective'], golden['objective'])
