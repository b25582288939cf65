"""Synthetic workloads of BASELINE.json `configs` / SURVEY.md section 8d as
declarative, JSON-able *scenarios*, plus the few helpers that turn a scenario
into `sky.Dag` objects of this package.

The same dicts drive every implementation that must agree: the unmodified
reference (oracle/ref_harness, build container -> tests/golden/*.json and
`baseline/_ref` on the GPU box), the portable pandas oracle (oracle/) and the
CUDA path. bench.py, __graft_entry__.smoke() and tests/ all read the
workloads from here, so the thing that is timed is the thing that is pinned.

Scenario schema: see tests/scenarios.py (name, minimize, tasks[resources,
resources_kind, num_nodes, outputs_gb, inputs, time_est], edges, blocked).
Catalog specs are arguments of skypilot_b200.synth.make_catalogs.
"""
from typing import Any, Dict, List, Optional

CLOUDS4 = ['aws', 'gcp', 'azure', 'lambda']

# name -> synth.make_catalogs(**spec)
CATALOGS: Dict[str, Dict[str, Any]] = {
    # cfg2 / cfg3 / cfg5: multi-cloud, ~50k rows (seed 1)
    'cfg2': {'seed': 1, 'n_rows': 50000, 'clouds': list(CLOUDS4)},
    # cfg4: synthetic 1M-row catalog (seed 3)
    'cfg4': {'seed': 3, 'n_rows': 1000000, 'clouds': list(CLOUDS4)},
}

# The cfg2 constraint set.
CFG2_TASKS = [
    {'accelerators': 'V100', 'outputs_gb': 10},
    {'accelerators': 'T4', 'outputs_gb': 10},
    {'accelerators': 'A100:8', 'outputs_gb': 10},
    {'accelerators': 'H100:8', 'outputs_gb': 10},
    {'accelerators': 'L4', 'outputs_gb': 10},
    {'cpus': '8+', 'outputs_gb': 10},
    {'cpus': '32+', 'memory': '128+', 'outputs_gb': 10},
    {'memory': '4x', 'outputs_gb': 10},
]

_TASK_KEYS = ('num_nodes', 'outputs_gb', 'inputs', 'time_est')


def single(name: str, **res) -> Dict[str, Any]:
    """One task, one requested Resources."""
    extra = {}
    for key in _TASK_KEYS:
        if key in res:
            extra[key] = res.pop(key)
    task = {'resources': [res]}
    task.update(extra)
    return {'name': name, 'tasks': [task]}


def chain(name: str, specs, **kw) -> Dict[str, Any]:
    """A chain DAG t0 -> t1 -> ... of the given task specs."""
    tasks = []
    for i, spec in enumerate(specs):
        spec = dict(spec)
        task = {'name': f't{i}'}
        for key in _TASK_KEYS + ('resources_kind',):
            if key in spec:
                task[key] = spec.pop(key)
        task['resources'] = spec.pop('resources', None) or [spec]
        tasks.append(task)
    sc = {'name': name, 'tasks': tasks,
          'edges': [[i, i + 1] for i in range(len(tasks) - 1)]}
    sc.update(kw)
    return sc


def chain_scenario(n_tasks: int) -> Dict[str, Any]:
    """cfg2 (n=8) / cfg4 (n=32): the cfg2 constraint set, cycled with varying
    cpus / memory thresholds and alternating spot requests."""
    specs = []
    for i in range(n_tasks):
        spec = dict(CFG2_TASKS[i % len(CFG2_TASKS)])
        round_ = i // len(CFG2_TASKS)
        if round_ and 'cpus' in spec:
            spec['cpus'] = f'{int(spec["cpus"].rstrip("+")) * (round_ + 1)}+'
        if round_ and 'memory' in spec and spec['memory'].endswith('+'):
            spec['memory'] = f'{int(spec["memory"][:-1]) * (round_ + 1)}+'
        if round_ and 'accelerators' in spec and round_ % 2 == 1:
            spec['use_spot'] = True
        specs.append(spec)
    return chain(f'chain{n_tasks}', specs)


# cfg5 (BASELINE.json configs[4]): independent
# single-task DAGs with constraint vectors drawn from
#   acc in 12 names + None, count in {1,2,4,8}, cpus in {None,2+,8+,32+},
#   memory in {None,16+,4x}, use_spot in {F,T}, region in {None, 10 regions}.
CFG5_ACCS = ['V100', 'T4', 'A100', 'A100-80GB', 'H100', 'L4', 'A10G', 'K80',
             'A10', 'P100', 'H200', 'L40S']
CFG5_REGIONS = [('aws', 'us-east-1'), ('aws', 'us-west-2'),
                ('aws', 'eu-west-1'), ('aws', 'ap-northeast-1'),
                ('gcp', 'us-central1'), ('gcp', 'europe-west4'),
                ('gcp', 'asia-east1'), ('azure', 'eastus'),
                ('azure', 'westeurope'), ('lambda', 'us-east-1')]


def cfg5_scenarios(n: int = 10000, seed: int = 4) -> List[Dict[str, Any]]:
    import numpy as np  # pylint: disable=import-outside-toplevel
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        spec: Dict[str, Any] = {}
        a = int(rng.integers(len(CFG5_ACCS) + 1))
        count = [1, 2, 4, 8][int(rng.integers(4))]
        cpus = [None, '2+', '8+', '32+'][int(rng.integers(4))]
        mem = [None, '16+', '4x'][int(rng.integers(3))]
        spot = bool(rng.integers(2))
        region = int(rng.integers(2 * len(CFG5_REGIONS)))
        if a < len(CFG5_ACCS):
            spec['accelerators'] = f'{CFG5_ACCS[a]}:{count}'
        if cpus:
            spec['cpus'] = cpus
        if mem:
            spec['memory'] = mem
        if spot:
            spec['use_spot'] = True
        if region < len(CFG5_REGIONS):
            spec['cloud'], spec['region'] = CFG5_REGIONS[region]
        out.append(single(f'd{i}', **spec))
    return out


# --------------------------------------------------------------------------
# scenario -> Dag (this package's types)


def make_time_estimator(spec):
    """Declarative estimator -> callable(resources) (sky/task.py:1361-1379)."""
    by_acc = spec.get('by_acc', {})
    default = spec.get('default', 3600)
    by_cloud = spec.get('by_cloud', {})

    def estimate(resources):
        seconds = default
        accs = resources.accelerators
        if accs:
            seconds = by_acc.get(list(accs.keys())[0], seconds)
        if resources.cloud is not None:
            seconds = by_cloud.get(str(resources.cloud).lower(), seconds)
        return seconds

    return estimate


def make_resources(spec):
    import skypilot_b200 as sky  # pylint: disable=import-outside-toplevel
    from skypilot_b200.utils import registry  # pylint: disable=import-outside-toplevel
    kwargs = dict(spec)
    cloud = kwargs.pop('cloud', None)
    if cloud is not None:
        kwargs['cloud'] = registry.CLOUD_REGISTRY.from_str(cloud)
    return sky.Resources(**kwargs)


def _task_extras(task, tspec) -> None:
    if 'outputs_gb' in tspec:
        task.set_outputs('CLOUD://out', tspec['outputs_gb'])
    if 'inputs' in tspec:
        task.set_inputs(tspec['inputs'][0], tspec['inputs'][1])
    if 'time_est' in tspec:
        task.set_time_estimator(make_time_estimator(tspec['time_est']))


def build_dag(scenario):
    """-> (Dag, [Task ...]) of skypilot_b200 types."""
    import skypilot_b200 as sky  # pylint: disable=import-outside-toplevel
    tasks = []
    with sky.Dag() as dag:
        for i, tspec in enumerate(scenario['tasks']):
            task = sky.Task(name=tspec.get('name', f't{i}'),
                            num_nodes=tspec.get('num_nodes', 1))
            if 'resources_yaml' in tspec:
                import copy  # pylint: disable=import-outside-toplevel
                task.set_resources(sky.Resources.from_yaml_config(
                    copy.deepcopy(tspec['resources_yaml'])))
                _task_extras(task, tspec)
                tasks.append(task)
                continue
            res = [make_resources(r) for r in tspec['resources']]
            kind = tspec.get('resources_kind', 'single')
            if kind == 'single':
                task.set_resources(res[0])
            elif kind == 'list':
                task.set_resources(res)
            else:
                task.set_resources(set(res))
            _task_extras(task, tspec)
            tasks.append(task)
        for u, v in scenario.get('edges', []):
            dag.add_edge(tasks[u], tasks[v])
    return dag, tasks


def blocked_list(scenario) -> Optional[List[Any]]:
    out = [make_resources(spec) for spec in scenario.get('blocked', [])]
    return out or None


def res_record(r) -> Dict[str, Any]:
    """The fields of a chosen Resources the parity checks compare."""
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


# name -> catalog spec and chain length of the two chain workloads (tools/)
WORKLOADS = {
    'cfg2': {'catalog': CATALOGS['cfg2'], 'tasks': 8},
    'cfg4': {'catalog': CATALOGS['cfg4'], 'tasks': 32},
}
