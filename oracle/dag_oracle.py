"""Exact optimum of the reference's general-DAG objective for DAGs too large
to enumerate -- test infrastructure (only tests/ and oracle/ref_harness import
this; the product's solver is skypilot_b200/dag_solver.py, written
independently as a bucket elimination).

The reference minimises, over one launchable candidate per task
(sky/optimizer.py:605-620, COST):

    sum_t cost[t][x_t]  +  sum_(u -> v) egress(u, x_u, v, x_v)

with PuLP + CBC. CBC is not installed here, and
tests/test_optimizer_random_dag.py:110-148 of the reference checks the ILP
against brute force, which stops at a dozen tasks. This restatement is a
*frontier* dynamic program: tasks are visited in a topological order; the
state is the choice made for every visited task that still has an unvisited
child (the frontier). Exact; exponential only in the frontier width.
"""
from typing import Callable, Dict, List, Sequence, Tuple


def frontier_dp(order: Sequence, children: Dict, parents: Dict,
                options: Dict, node_cost: Callable, edge_cost: Callable
                ) -> Tuple[float, Dict]:
    """order: tasks in topological order; children / parents: adjacency;
    options[t]: candidate choices of t; node_cost(t, x); edge_cost(u, xu, v,
    xv). Returns (optimum, {task: choice})."""
    pos = {t: i for i, t in enumerate(order)}
    last_child = {t: max([pos[c] for c in children[t]] or [-1]) for t in order}
    # state: tuple of (task, choice) for frontier tasks -> (value, back-pointer)
    states: Dict[Tuple, Tuple[float, Tuple]] = {(): (0.0, None)}
    history: List[Dict] = []
    for i, t in enumerate(order):
        new_states: Dict[Tuple, Tuple[float, Tuple]] = {}
        for key, (val, _) in states.items():
            chosen = dict(key)
            for x in options[t]:
                v = val + node_cost(t, x)
                for p in parents[t]:
                    v += edge_cost(p, chosen[p], t, x)
                nxt = dict(chosen)
                nxt[t] = x
                # drop tasks whose children are all visited
                front = tuple(sorted(
                    ((u, xu) for u, xu in nxt.items() if last_child[u] > i),
                    key=lambda kv: pos[kv[0]]))
                old = new_states.get(front)
                if old is None or v < old[0]:
                    new_states[front] = (v, (key, t, x))
        history.append(new_states)
        states = new_states
    best_key = min(states, key=lambda k: states[k][0])
    best = states[best_key][0]
    plan: Dict = {}
    key = best_key
    for i in range(len(order) - 1, -1, -1):
        _, back = history[i][key]
        key, t, x = back
        plan[t] = x
    return best, plan
