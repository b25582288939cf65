"""Run scenario specs through the UNMODIFIED reference optimizer.

Test infrastructure only (build container only; see bootstrap.py). One process
per catalog: the reference reads `~/.sky/catalogs` lazily at module import, so
a process is bound to the scratch $HOME it was started with.

A *scenario* is a JSON-able dict (schema documented in tests/scenarios.py):
tasks with `resources` kwargs, chain / DAG edges, optional blocked resources,
inputs/outputs sizes and a declarative time estimator. For each scenario the
harness records

  * the plan: per task (cloud, instance_type, region, zone, accelerators,
    use_spot) of `task.best_resources` after `Optimizer.optimize`;
  * the ordered candidate table the optimizer built
    (`_estimate_nodes_cost_or_time`, reference sky/optimizer.py:239-426):
    per task the list of (cloud, instance_type, region, zone, cost_or_time);
  * the objective, and total cost / time of the plan
    (`_compute_total_cost/_time`, sky/optimizer.py:640-698);
  * for non-chain DAGs (the reference needs PuLP/CBC, absent here): the exact
    optimum by exhaustive search over the reference's own candidate table and
    `_egress_cost_or_time`, as tests/test_optimizer_random_dag.py:110-148 does;
  * the error class/message when the reference raises.

usage: run_reference.py --catalog '{"seed": 1, "n_rows": 50000, ...}'
                        --scenarios in.json --out out.json
"""
import argparse
import itertools
import json
import math
import os
import sys
import tempfile
import time

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(os.path.dirname(_HERE))
sys.path.insert(0, _REPO)
sys.path.insert(0, _HERE)

import bootstrap  # noqa: E402  pylint: disable=wrong-import-position


def _make_time_estimator(spec):
    """Declarative estimator -> callable(resources) (sky/task.py:1361-1379)."""
    by_acc = spec.get('by_acc', {})
    default = spec.get('default', 3600)
    by_cloud = spec.get('by_cloud', {})

    def estimate(resources):
        seconds = default
        accs = resources.accelerators
        if accs:
            name = list(accs.keys())[0]
            seconds = by_acc.get(name, seconds)
        if resources.cloud is not None:
            seconds = by_cloud.get(str(resources.cloud).lower(), seconds)
        return seconds

    return estimate


def _resources_kwargs(sky, spec):
    kwargs = dict(spec)
    cloud = kwargs.pop('cloud', None)
    if cloud is not None:
        from sky.utils import registry
        kwargs['cloud'] = registry.CLOUD_REGISTRY.from_str(cloud)
    del sky
    return kwargs


def _task_extras(task, tspec):
    if 'outputs_gb' in tspec:
        task.set_outputs(tspec.get('outputs', 'CLOUD://out'),
                         estimated_size_gigabytes=tspec['outputs_gb'])
    if 'inputs' in tspec:
        task.set_inputs(tspec['inputs'][0],
                        estimated_size_gigabytes=tspec['inputs'][1])
    if 'time_est' in tspec:
        task.set_time_estimator(_make_time_estimator(tspec['time_est']))


def build_dag(sky, scenario):
    tasks = []
    with sky.Dag() as dag:
        for i, tspec in enumerate(scenario['tasks']):
            task = sky.Task(name=tspec.get('name', f't{i}'),
                            num_nodes=tspec.get('num_nodes', 1))
            if 'resources_yaml' in tspec:
                # request alternatives the way a task YAML states them
                # (sky/resources.py:2264-2411)
                import copy as _copy
                task.set_resources(sky.Resources.from_yaml_config(
                    _copy.deepcopy(tspec['resources_yaml'])))
                tasks.append(task)
                _task_extras(task, tspec)
                continue
            res = [
                sky.Resources(**_resources_kwargs(sky, r))
                for r in tspec['resources']
            ]
            kind = tspec.get('resources_kind', 'single')
            if kind == 'single':
                assert len(res) == 1
                task.set_resources(res[0])
            elif kind == 'list':
                task.set_resources(res)
            else:
                # NOTE: a Python set of identity-hashed Resources iterates in
                # address order (reference sky/task.py:1292-1312); the harness
                # records the realised order next to the candidates.
                task.set_resources(set(res))
            _task_extras(task, tspec)
            tasks.append(task)
        for u, v in scenario.get('edges', []):
            dag.add_edge(tasks[u], tasks[v])
    return dag, tasks


def _res_record(r):
    accs = r.accelerators
    return {
        'cloud': None if r.cloud is None else str(r.cloud).lower(),
        'instance_type': r.instance_type,
        'region': r.region,
        'zone': r.zone,
        'accelerators': None if accs is None else
                        {k: float(v) for k, v in accs.items()},
        'use_spot': bool(r.use_spot),
    }


def _blocked(sky, scenario):
    out = []
    for spec in scenario.get('blocked', []):
        kwargs = _resources_kwargs(sky, spec)
        spot = kwargs.pop('use_spot', None)
        r = sky.Resources(**kwargs)
        if spot is not None:
            r = r.copy(use_spot=spot)
        else:
            # Resources() turns use_spot=None into False; blocked wildcards
            # are built by failover handlers with an explicit None
            # (reference sky/backends/cloud_vm_ray_backend.py:332-339).
            r._use_spot_specified = False  # pylint: disable=protected-access
        out.append(r)
    return out or None


def _finite(x):
    x = float(x)
    return x if math.isfinite(x) else repr(x)


def _plain(x):
    """JSON-able scalar: NaN / None -> None, numpy scalars -> Python."""
    if x is None:
        return None
    if isinstance(x, str):
        return x
    x = float(x)
    return None if math.isnan(x) else x


def run_listing(sky, scenario):
    """`sky.catalog.list_accelerators(**kwargs)` (sky/catalog/__init__.py:
    56-85) -> {name: [InstanceTypeInfo fields ...]} in the reference's order."""
    from sky import catalog
    bootstrap.clear_request_cache()
    result = catalog.list_accelerators(**scenario.get('kwargs', {}))
    out = {}
    for name, infos in result.items():
        out[name] = [[
            i.cloud, _plain(i.instance_type), i.accelerator_name,
            _plain(i.accelerator_count), _plain(i.cpu_count),
            _plain(i.device_memory), _plain(i.memory), _plain(i.price),
            _plain(i.spot_price), i.region
        ] for i in infos]
    return {'name': scenario['name'], 'listing': out,
            'order': list(result.keys())}


def _jsonable(x):
    if isinstance(x, (list, tuple)):
        return [_jsonable(v) for v in x]
    if isinstance(x, dict):
        return {str(k): _jsonable(v) for k, v in x.items()}
    if x is None or isinstance(x, (bool, str)):
        return x
    if isinstance(x, (int, float)) or hasattr(x, 'dtype'):
        v = float(x)
        if math.isnan(v):
            return None
        return int(v) if isinstance(x, int) else v
    if hasattr(x, 'name'):  # Region / Zone
        return x.name
    return str(x)


def run_call(sky, scenario):
    """One call of the catalog function table (sky/catalog/__init__.py)."""
    from sky import catalog
    bootstrap.clear_request_cache()
    fn = getattr(catalog, scenario['fn'])
    try:
        result = fn(*scenario.get('args', []), **scenario.get('kwargs', {}))
    except Exception as e:  # pylint: disable=broad-except
        return {'name': scenario['name'],
                'error': {'type': type(e).__name__, 'message': str(e)}}
    return {'name': scenario['name'], 'result': _jsonable(result)}


def run_job_group(sky, scenario):
    """`Optimizer.optimize_job_group` (sky/optimizer.py:1039-1200) on a
    parallel-execution DAG. Besides the plan, records the common infras the
    reference found: with several of them and more than one job its choice is
    the first element of a Python set (see skypilot_b200/optimizer.py)."""
    from sky import dag as dag_lib
    from sky import exceptions
    from sky import optimizer as opt_lib
    from sky.utils import common as sky_common
    bootstrap.clear_request_cache()
    minimize_cost = scenario.get('minimize', 'cost') == 'cost'
    target = (sky_common.OptimizeTarget.COST
              if minimize_cost else sky_common.OptimizeTarget.TIME)
    dag, tasks = build_dag(sky, scenario)
    dag.name = scenario['name']
    dag.set_execution(dag_lib.DagExecution.PARALLEL)
    record = {'name': scenario['name']}
    # the common infras, computed the reference's way
    try:
        launchables = {}
        for t in tasks:
            per_req, _, _, _ = opt_lib._fill_in_launchable_resources(
                t, blocked_resources=None, quiet=True)
            by_cloud = {}
            for lst in per_req.values():
                for res in lst:
                    by_cloud.setdefault(res.cloud, []).append(res)
            launchables[t] = by_cloud
        if all(launchables[t] for t in tasks):
            common = opt_lib.Optimizer._find_common_infras(launchables)
            record['common_infras'] = sorted(
                [str(c).lower(), r] for c, r in common)
    except exceptions.ResourcesUnavailableError:
        pass
    bootstrap.clear_request_cache()
    try:
        opt_lib.Optimizer.optimize_job_group(dag, target, None, quiet=True)
    except exceptions.ResourcesUnavailableError as e:
        record['error'] = {'type': 'ResourcesUnavailableError',
                           'message': str(e)}
        return record
    record['plan'] = [_res_record(t.best_resources) for t in tasks]
    record['overrides'] = [[
        [None if r.cloud is None else str(r.cloud).lower(), r.region]
        for r in list(t.resources)
    ] for t in tasks]
    return record


def run_scenario(sky, scenario):
    if scenario.get('config') is not None and not scenario.get('_in_config'):
        # a SkyPilot config for this scenario only (e.g. a per-region
        # ssh_proxy_command: sky/resources.py:1210-1246)
        from sky import skypilot_config
        from sky.utils import config_utils
        inner = dict(scenario, _in_config=True)
        with skypilot_config.replace_skypilot_config(
                config_utils.Config.from_dict(scenario['config'])):
            return run_scenario(sky, inner)
    if scenario.get('kind') == 'list_accelerators':
        return run_listing(sky, scenario)
    if scenario.get('kind') == 'job_group':
        return run_job_group(sky, scenario)
    if scenario.get('kind') == 'catalog_call':
        return run_call(sky, scenario)
    from sky import exceptions
    from sky import optimizer as opt_lib
    from sky.utils import common as sky_common
    import networkx as nx
    Optimizer = opt_lib.Optimizer
    bootstrap.clear_request_cache()
    minimize_cost = scenario.get('minimize', 'cost') == 'cost'
    target = (sky_common.OptimizeTarget.COST
              if minimize_cost else sky_common.OptimizeTarget.TIME)
    dag, tasks = build_dag(sky, scenario)
    blocked = _blocked(sky, scenario)
    record = {'name': scenario['name']}
    is_chain = dag.is_chain()
    record['is_chain'] = bool(is_chain)

    has_list = any(
        t.get('resources_kind') == 'list' or isinstance(tk.resources, list)
        for t, tk in zip(scenario['tasks'], tasks))
    if has_list:
        # Ordered resources are resolved by _optimize_dag's pre-pass
        # (sky/optimizer.py:1403-1448); only the end-to-end plan is recorded.
        assert is_chain
        try:
            Optimizer.optimize(dag, minimize=target,
                               blocked_resources=blocked, quiet=True)
        except exceptions.ResourcesUnavailableError as e:
            record['error'] = {
                'type': 'ResourcesUnavailableError',
                'message': str(e)
            }
            return record
        record['plan'] = [_res_record(t.best_resources) for t in tasks]
        record['ordered'] = True
        return record

    # Candidate table + (for general DAGs) exhaustive optimum, on a dag with
    # the dummy source/sink attached exactly as Optimizer.optimize does.
    Optimizer._add_dummy_source_sink_nodes(dag)
    try:
        graph = dag.get_graph()
        topo = list(nx.topological_sort(graph))
        try:
            cost_map, _ = Optimizer._estimate_nodes_cost_or_time(
                topo, minimize_cost, blocked, quiet=True)
        except exceptions.ResourcesUnavailableError as e:
            record['error'] = {
                'type': 'ResourcesUnavailableError',
                'message': str(e)
            }
            return record
        cands = []
        for t in tasks:
            cands.append([
                dict(_res_record(r), value=_finite(v))
                for r, v in cost_map[t].items()
            ])
        record['candidates'] = cands
        if is_chain:
            plan, objective = Optimizer._optimize_by_dp(topo, cost_map,
                                                        minimize_cost)
        else:
            # Exhaustive search (PuLP/CBC is not installed here). Egress
            # depends on the two clouds only (sky/optimizer.py:75-104), so
            # inside one cloud the first cheapest candidate of a task
            # dominates; the product runs over those representatives unless
            # SKYOPT_FULL_EXHAUSTIVE=1 asks for every combination.
            full = os.environ.get('SKYOPT_FULL_EXHAUSTIVE') == '1'
            names = []
            for n in topo:
                keys = list(cost_map[n].keys())
                if not full:
                    best_of = {}
                    for r in keys:
                        c = str(r.cloud)
                        if c not in best_of or (cost_map[n][r] <
                                                cost_map[n][best_of[c]]):
                            best_of[c] = r
                    keys = [r for r in keys if best_of[str(r.cloud)] is r]
                names.append(keys)
            record['exhaustive'] = 'full' if full else 'per-cloud'
            n_combos = 1
            for keys in names:
                n_combos *= max(len(keys), 1)
            best, best_plan = None, None
            if n_combos > 3_000_000 and minimize_cost:
                # too many assignments to enumerate: the exact frontier DP of
                # oracle/dag_oracle.py over the same tables and egress terms
                from oracle import dag_oracle
                record['exhaustive'] = 'frontier-dp'
                children = {n: list(graph.successors(n)) for n in topo}
                parents_of = {n: list(graph.predecessors(n)) for n in topo}
                best, best_plan = dag_oracle.frontier_dp(
                    topo, children, parents_of, dict(zip(topo, names)),
                    lambda n, r: cost_map[n][r],
                    lambda u, ru, v, rv: Optimizer._egress_cost_or_time(
                        True, u, ru, v, rv))
                names = None
            for combo in (itertools.product(*names) if names is not None else []):
                plan_try = dict(zip(topo, combo))
                if minimize_cost:
                    total = 0.0
                    for n in topo:
                        total += cost_map[n][plan_try[n]]
                    for u, v in graph.edges():
                        total += Optimizer._egress_cost_or_time(
                            True, u, plan_try[u], v, plan_try[v])
                else:
                    finish = {}
                    for n in topo:
                        start = 0
                        for p in graph.predecessors(n):
                            start = max(
                                start, finish[p] +
                                Optimizer._egress_cost_or_time(
                                    False, p, plan_try[p], n, plan_try[n]))
                        finish[n] = cost_map[n][plan_try[n]] + start
                    total = finish[topo[-1]]
                if best is None or total < best:
                    best, best_plan = total, plan_try
            plan, objective = best_plan, best
            for n, r in plan.items():
                n.best_resources = r
        record['objective'] = _finite(objective)
        record['plan'] = [_res_record(plan[t]) for t in tasks]
        record['plan_index'] = [
            list(cost_map[t].keys()).index(plan[t]) for t in tasks
        ]
        record['total_cost'] = _finite(
            Optimizer._compute_total_cost(graph, topo, plan))
        record['total_time'] = _finite(
            Optimizer._compute_total_time(graph, topo, plan))
    finally:
        Optimizer._remove_dummy_source_sink_nodes(dag)

    # End-to-end through the public entry point (chain DAGs only).
    if is_chain:
        for t in tasks:
            t.best_resources = None
        bootstrap.clear_request_cache()
        Optimizer.optimize(dag, minimize=target, blocked_resources=blocked,
                           quiet=True)
        e2e = [_res_record(t.best_resources) for t in tasks]
        assert e2e == record['plan'], (e2e, record['plan'])
    return record


def time_scenario(sky, scenario, warmup, iters):
    """p50/p90 of `Optimizer.optimize(dag, quiet=True)`, cache cleared."""
    from sky import optimizer as opt_lib
    dag, tasks = build_dag(sky, scenario)
    samples = []
    for i in range(warmup + iters):
        for t in tasks:
            t.best_resources = None
        bootstrap.clear_request_cache()
        t0 = time.perf_counter()
        opt_lib.Optimizer.optimize(dag, quiet=True)
        dt = time.perf_counter() - t0
        if i >= warmup:
            samples.append(dt)
    samples.sort()
    return {
        'p50_ms': 1e3 * samples[len(samples) // 2],
        'p90_ms': 1e3 * samples[int(len(samples) * 0.9)],
        'iters': iters
    }


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--catalog', required=True)
    parser.add_argument('--scenarios', required=True)
    parser.add_argument('--out', required=True)
    parser.add_argument('--time', type=int, default=0,
                        help='also time each chain scenario this many times')
    args = parser.parse_args()
    from skypilot_b200 import synth
    cat = json.loads(args.catalog)
    enabled = cat.pop('enabled', None) or cat.get('clouds') or [
        'aws', 'gcp', 'azure', 'lambda'
    ]
    catalogs = synth.make_catalogs(**cat)
    with open(args.scenarios, encoding='utf-8') as f:
        scenarios = json.load(f)
    out = []
    with tempfile.TemporaryDirectory(prefix='skyref_home_') as home:
        bootstrap.write_catalogs(home, catalogs)
        sky = bootstrap.import_reference(home, enabled)
        for sc in scenarios:
            try:
                rec = run_scenario(sky, sc)
            except Exception as e:  # pylint: disable=broad-except
                rec = {
                    'name': sc['name'],
                    'error': {
                        'type': type(e).__name__,
                        'message': str(e)
                    }
                }
            if args.time and rec.get('is_chain') and 'error' not in rec:
                rec['timing'] = time_scenario(sky, sc, 2, args.time)
            out.append(rec)
            print(f'[ref] {sc["name"]}: '
                  f'{"ERROR " + rec["error"]["type"] if "error" in rec else rec.get("objective")}',
                  file=sys.stderr)
    with open(args.out, 'w', encoding='utf-8') as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == '__main__':
    main()
