"""Size-independent properties at the benchmark's full size (the 32-task
chain on the 1 M-row catalog): checks that need no fixture and would catch a
wrong plan whatever the reference says.

  * the device's objective equals the cost recomputed on the host from the
    plan it returned (a checksum of the whole DP);
  * every task's candidate table is its slots' launchables in request order,
    values consistent with hourly price x runtime x nodes;
  * blocking the chosen launchable of a task never lowers the objective, and
    the new plan avoids it;
  * the same request again gives the same plan (idempotence), through the
    session (candidates kept on the device) as well.
"""
import math

import networkx as nx
import pytest

import skypilot_b200 as sky
from skypilot_b200 import optimizer as opt_lib
from skypilot_b200 import workloads
from tests import scenario_runner as runner
from tests import scenarios

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def chain():
    runner.activate_catalog(scenarios.CATALOGS['cfg4_1m'])
    dag, tasks = workloads.build_dag(workloads.chain_scenario(32))
    sky.optimize(dag, quiet=True)
    return dag, tasks


def _objective(dag, tasks):
    O = opt_lib.Optimizer
    O._add_dummy_source_sink_nodes(dag)  # pylint: disable=protected-access
    try:
        graph = dag.get_graph()
        topo = list(nx.topological_sort(graph))
        plan = {t: (t.best_resources if not opt_lib._is_dummy(t)  # pylint: disable=protected-access
                    else list(t.resources)[0]) for t in topo}
        return float(O._compute_total_cost(graph, topo, plan))  # pylint: disable=protected-access
    finally:
        O._remove_dummy_source_sink_nodes(dag)  # pylint: disable=protected-access


def test_objective_is_the_cost_of_the_plan(chain):
    dag, tasks = chain
    O = opt_lib.Optimizer
    O._add_dummy_source_sink_nodes(dag)  # pylint: disable=protected-access
    try:
        graph = dag.get_graph()
        topo = [t for t in nx.topological_sort(graph)
                if not opt_lib._is_dummy(t)]  # pylint: disable=protected-access
        problem = O._state_problem(graph, topo, True, [], True)  # pylint: disable=protected-access
        sol = O._solve(problem, want_tables=True)  # pylint: disable=protected-access
    finally:
        O._remove_dummy_source_sink_nodes(dag)  # pylint: disable=protected-access
    assert sol.dag[0]['status'] == 0
    device = float(sol.dag[0]['objective'])
    host = _objective(dag, tasks)
    assert math.isclose(device, host, rel_tol=1e-9), (device, host)
    # the chosen record of every task is a row of its table, and the table is
    # the concatenation of its slots in request order
    for i, task in enumerate(topo):
        table = sol.task_table(i)
        assert len(table) == int(sol.task_n[i]) > 0
        chosen = sol.chosen[i]
        row = table[int(sol.chosen_index[i])]
        for field in ('slot', 'inst_id', 'region_id', 'zone_id', 'hourly',
                      'value'):
            assert row[field] == chosen[field], (i, field)
        slots = table['slot']
        assert (slots[1:] >= slots[:-1]).all(), i
        runtime = 3600.0 if task.time_estimator_func is None else float(
            task.estimate_runtime(task.best_resources))
        want = table['hourly'] * (runtime / 3600) * max(task.num_nodes, 0)
        assert all(math.isclose(a, b, rel_tol=1e-12, abs_tol=1e-15)
                   for a, b in zip(table['value'], want)), i


def test_blocking_the_choice_never_lowers_the_objective(chain):
    dag, tasks = chain
    base = _objective(dag, tasks)
    first_plan = [workloads.res_record(t.best_resources) for t in tasks]
    for k in (0, 7, 18, 31):
        blocked = [tasks[k].best_resources.copy()]
        dag2, tasks2 = workloads.build_dag(workloads.chain_scenario(32))
        sky.optimize(dag2, blocked_resources=blocked, quiet=True)
        again = _objective(dag2, tasks2)
        assert again >= base * (1 - 1e-12), (k, again, base)
        got = workloads.res_record(tasks2[k].best_resources)
        assert got != first_plan[k], k
        assert not tasks2[k].best_resources.should_be_blocked_by(blocked[0])


def test_same_request_same_plan(chain):
    dag, tasks = chain
    first = [workloads.res_record(t.best_resources) for t in tasks]
    base = _objective(dag, tasks)
    dag2, tasks2 = workloads.build_dag(workloads.chain_scenario(32))
    sky.optimize(dag2, quiet=True)
    assert [workloads.res_record(t.best_resources) for t in tasks2] == first
    with sky.Optimizer.session(dag2) as session:
        session.optimize([])
        assert [workloads.res_record(t.best_resources)
                for t in tasks2] == first
        blocked = [tasks2[3].best_resources.copy()]
        session.optimize(blocked)
        assert _objective(dag2, tasks2) >= base * (1 - 1e-12)
        # ... and equals a fresh optimisation under the same blocked list
        dag3, tasks3 = workloads.build_dag(workloads.chain_scenario(32))
        sky.optimize(dag3, blocked_resources=blocked, quiet=True)
        assert [workloads.res_record(t.best_resources) for t in tasks2] == [
            workloads.res_record(t.best_resources) for t in tasks3]
