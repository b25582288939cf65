"""Exact placement of general (non-chain) DAGs beyond the device enumeration.

The reference solves general DAGs with an ILP through PuLP + CBC
(sky/optimizer.py:490-637; no size limit). On the device `solve_kernel`
enumerates cloud assignments, which is exact but exponential in the number of
tasks. This module removes that wall for the COST objective.

Within one cloud the cheapest candidate of a task dominates (egress depends
on the two clouds only, sky/optimizer.py:75-104), so the ILP's objective
(sky/optimizer.py:605-620) is a min-sum problem over one variable per task
(its cloud) with unary terms (the task's cheapest value in that cloud, plus
the egress of a source task's inputs) and pairwise terms per DAG edge (the
parent's tariff when the clouds differ). That is solved exactly by variable
(bucket) elimination: eliminate one task at a time, replacing every term that
mentions it by their minimum over its cloud; cost O(T * C^(w+1)) for clouds C
and induced width w of the elimination order -- workflow DAGs (chains of
diamonds, fan-out / fan-in stages) have w of 2 or 3. A greedy min-fill order
is used; a term of more than `max_table` entries raises OptimizerLimitError.

The TIME objective (critical path; min-max, not a sum) does not decompose this
way: beyond the device enumeration it raises OptimizerLimitError.
"""
from typing import Dict, List, Sequence, Tuple

import numpy as np

from skypilot_b200 import exceptions

DEFAULT_MAX_TABLE = 1 << 24


def _order(n: int, neighbours: List[set]) -> List[int]:
    """Greedy min-fill elimination order (ties: smaller degree, lower id)."""
    adj = [set(s) for s in neighbours]
    left = set(range(n))
    order = []
    while left:
        best, best_key = None, None
        for v in left:
            nb = adj[v] & left
            fill = 0
            nbl = list(nb)
            for i, a in enumerate(nbl):
                for b in nbl[i + 1:]:
                    if b not in adj[a]:
                        fill += 1
            key = (fill, len(nb), v)
            if best_key is None or key < best_key:
                best, best_key = v, key
        nb = adj[best] & left
        for a in nb:
            adj[a] |= nb - {a}
        left.discard(best)
        order.append(best)
    return order


def min_sum(unary: Sequence[np.ndarray],
            pairwise: Sequence[Tuple[int, int, np.ndarray]],
            max_table: int = DEFAULT_MAX_TABLE) -> Tuple[float, List[int]]:
    """argmin_x  sum_t unary[t][x_t] + sum_(u,v,M) M[x_u, x_v].

    unary[t]: C floats (inf = the task has no candidate in that cloud);
    pairwise: (u, v, CxC matrix indexed [x_u, x_v]). Returns (objective,
    assignment). Ties resolve to the lowest cloud index at every elimination.
    """
    n = len(unary)
    if n == 0:
        return 0.0, []
    c = len(unary[0])
    neighbours: List[set] = [set() for _ in range(n)]
    for u, v, _ in pairwise:
        if u != v:
            neighbours[u].add(v)
            neighbours[v].add(u)
    # factors: (scope tuple in ascending task order, array with one axis per
    # scope variable)
    factors: List[Tuple[Tuple[int, ...], np.ndarray]] = []
    for t, vec in enumerate(unary):
        factors.append(((t,), np.asarray(vec, dtype=np.float64)))
    for u, v, mat in pairwise:
        mat = np.asarray(mat, dtype=np.float64)
        if u == v:
            continue
        if u < v:
            factors.append(((u, v), mat))
        else:
            factors.append(((v, u), mat.T))
    order = _order(n, neighbours)
    decisions: List[Tuple[int, Tuple[int, ...], np.ndarray]] = []
    constant = 0.0
    for x in order:
        mine = [f for f in factors if x in f[0]]
        factors = [f for f in factors if x not in f[0]]
        scope = sorted(set(v for f in mine for v in f[0]) - {x})
        if c**(len(scope) + 1) > max_table:
            raise exceptions.OptimizerLimitError(
                f'general-DAG placement: eliminating a task couples '
                f'{len(scope)} others over {c} clouds '
                f'({c ** (len(scope) + 1)} table entries > {max_table}); the '
                'DAG is too densely connected for the exact solver.')
        axes = scope + [x]
        pos = {v: i for i, v in enumerate(axes)}
        total = np.zeros((c,) * len(axes), dtype=np.float64)
        for fscope, arr in mine:
            shape = [1] * len(axes)
            src_axes = [pos[v] for v in fscope]
            # bring the factor's axes into the joint order
            perm = np.argsort(src_axes)
            arr_t = np.transpose(arr, perm)
            for v in fscope:
                shape[pos[v]] = c
            total = total + arr_t.reshape(shape)
        arg = np.argmin(total, axis=-1)
        new = np.min(total, axis=-1)
        decisions.append((x, tuple(scope), arg))
        if scope:
            factors.append((tuple(scope), new))
        else:
            constant += float(new)
    assignment = [-1] * n
    for x, scope, arg in reversed(decisions):
        idx = tuple(assignment[v] for v in scope)
        assignment[x] = int(arg[idx]) if scope else int(arg)
    return constant, assignment


def evaluate_cost(unary, pairwise, assignment) -> float:
    """The objective of `assignment`, summed in the order the reference's
    plan cost is (nodes in order, then edges)."""
    total = 0.0
    for t, vec in enumerate(unary):
        total += float(vec[assignment[t]])
    for u, v, mat in pairwise:
        total += float(mat[assignment[u], assignment[v]])
    return total


def solve_cost_dag(values: Sequence[Sequence[float]],
                   clouds: Sequence[Sequence[int]],
                   parents: Sequence[Sequence[int]],
                   edge_tariffs: Sequence[Sequence[Sequence[float]]],
                   src_tariffs: Sequence,
                   n_clouds: int,
                   max_table: int = DEFAULT_MAX_TABLE
                   ) -> Tuple[List[int], float]:
    """COST placement of one DAG from its candidate tables.

    values[t][k], clouds[t][k]: value and cloud index of candidate k of task t
    (the reference's dictionary order); parents[t]: task indices;
    edge_tariffs[t][j][g]: egress of the edge parents[t][j] -> t when the
    parent runs in cloud g and the child elsewhere; src_tariffs[t]: None or the
    per-cloud egress of a source task's inputs. Returns (chosen candidate index
    per task, objective)."""
    inf = float('inf')
    unary, first = [], []
    for t, (vals, cls) in enumerate(zip(values, clouds)):
        best = np.full(n_clouds, inf)
        idx = np.full(n_clouds, -1, dtype=np.int64)
        for k, (v, g) in enumerate(zip(vals, cls)):
            if v < best[g]:  # strict: the first cheapest candidate of a cloud
                best[g] = v
                idx[g] = k
        if src_tariffs[t] is not None:
            best = best + np.asarray(src_tariffs[t], dtype=np.float64)
        unary.append(best)
        first.append(idx)
    pairwise = []
    for t, ps in enumerate(parents):
        for j, p in enumerate(ps):
            tar = np.asarray(edge_tariffs[t][j], dtype=np.float64)
            mat = np.repeat(tar[:, None], n_clouds, axis=1)
            np.fill_diagonal(mat, 0.0)
            pairwise.append((p, t, mat))
    _, assignment = min_sum(unary, pairwise, max_table)
    objective = evaluate_cost(unary, pairwise, assignment)
    if not np.isfinite(objective):
        raise exceptions.ResourcesUnavailableError(
            'No launchable resource found for a task of the DAG.')
    chosen = [int(first[t][assignment[t]]) for t in range(len(values))]
    return chosen, objective
