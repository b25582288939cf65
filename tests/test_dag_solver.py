"""skypilot_b200/dag_solver.py (exact elimination for general DAGs beyond the
device enumeration) against brute force and against the independent frontier
DP of oracle/dag_oracle.py. CPU only."""
import itertools

import numpy as np
import pytest

from oracle import dag_oracle
from skypilot_b200 import dag_solver
from skypilot_b200 import exceptions


def _random_problem(rng, n, c, p_edge=0.4, p_inf=0.3):
    unary = [rng.uniform(1, 10, c) for _ in range(n)]
    for u in unary:
        if rng.uniform() < p_inf:
            u[int(rng.integers(c))] = np.inf
    edges, mats = [], {}
    children = {t: [] for t in range(n)}
    parents = {t: [] for t in range(n)}
    for v in range(1, n):
        for u in range(max(0, v - 4), v):
            if rng.uniform() < p_edge:
                tar = rng.uniform(0, 3, c)
                m = np.repeat(tar[:, None], c, axis=1)
                np.fill_diagonal(m, 0.0)
                edges.append((u, v, m))
                mats[(u, v)] = m
                children[u].append(v)
                parents[v].append(u)
    return unary, edges, mats, children, parents


@pytest.mark.parametrize('seed', range(40))
def test_elimination_equals_brute_force(seed):
    rng = np.random.default_rng(seed)
    n, c = int(rng.integers(2, 9)), int(rng.integers(2, 5))
    unary, edges, _, _, _ = _random_problem(rng, n, c)
    _, assignment = dag_solver.min_sum(unary, edges)
    got = dag_solver.evaluate_cost(unary, edges, assignment)
    best = min(dag_solver.evaluate_cost(unary, edges, a)
               for a in itertools.product(range(c), repeat=n))
    assert got == pytest.approx(best, rel=1e-12) or (np.isinf(best) and
                                                     np.isinf(got))


@pytest.mark.parametrize('seed', range(10))
def test_elimination_equals_frontier_dp_on_large_dags(seed):
    """24-40 tasks, 3-6 clouds: far beyond enumeration; two independent exact
    methods must agree."""
    rng = np.random.default_rng(100 + seed)
    n, c = int(rng.integers(24, 41)), int(rng.integers(3, 7))
    unary, edges, mats, children, parents = _random_problem(
        rng, n, c, p_edge=0.3, p_inf=0.1)
    _, assignment = dag_solver.min_sum(unary, edges)
    got = dag_solver.evaluate_cost(unary, edges, assignment)
    want, _ = dag_oracle.frontier_dp(
        list(range(n)), children, parents,
        {t: list(range(c)) for t in range(n)},
        lambda t, x: unary[t][x], lambda u, xu, v, xv: mats[(u, v)][xu, xv])
    assert got == pytest.approx(want, rel=1e-12)


def test_solve_cost_dag_picks_first_cheapest_candidate_per_cloud():
    # two tasks, two clouds; task 0 has two equal-priced candidates in cloud 0
    values = [[1.0, 1.0, 5.0], [2.0, 0.5]]
    clouds = [[0, 0, 1], [0, 1]]
    parents = [[], [0]]
    edge = [[], [[10.0, 10.0]]]
    chosen, objective = dag_solver.solve_cost_dag(values, clouds, parents,
                                                  edge, [None, None], 2)
    assert chosen == [0, 0] and objective == pytest.approx(3.0)


def test_dense_dag_raises_the_limit_error_not_unavailable():
    n, c = 12, 8
    unary = [np.ones(c) for _ in range(n)]
    m = np.ones((c, c))
    edges = [(u, v, m) for v in range(n) for u in range(v)]  # a clique
    with pytest.raises(exceptions.OptimizerLimitError):
        dag_solver.min_sum(unary, edges, max_table=1 << 16)
    assert not issubclass(exceptions.OptimizerLimitError,
                          exceptions.ResourcesUnavailableError)
