"""bench.py -- candidates scored / second of the placement-optimizer hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun ... bench.py --gpus N ...          (one rank per GPU, N > 1)

One *step* = one pass of the hot path over one DAG: every task's constraint
vector is scored against every catalog row (filter + argmin), the winners are
expanded to region/zone candidates and costed, and the chain DP picks the
plan. Workloads (BASELINE.json `configs`, SURVEY.md section 8d;
skypilot_b200/workloads.py):

  cfg4   32-task chain DAG x synthetic 1M-row catalog -- the configuration the
         throughput / HBM-roofline target is quoted on (configs[3]). The
         headline of every line: each GPU optimises its own replica of the
         DAG against its own replica of the catalog (weak scaling, no
         collective -- the path only shards over independent DAGs).
  cfg5   10 000 seeded single-task DAGs on the cfg2 catalog sharded over the
         GPUs (configs[4]): DAG i -> rank i mod N, one device problem per
         rank, plans gathered on rank 0 and compared with the reference's
         records. Reported in every line under "scaling_cfg5" (strong scaling:
         the total work is fixed as N grows).
  cfg2   8-task chain x ~50k-row multi-cloud catalog, the optimize() latency
         configuration (configs[1]) -- N = 1 only, under "latency_cfg2".
  stress the cfg4 catalog with every row repeated 8 times (9.3M rows: larger
         than the 126 MB L2) scanned WITHOUT pruning -- N = 1 only, under
         "hbm_stress": what the scan kernel sustains when it must stream.

Keys: `value` = candidates/s with the problem resident in HBM (CUDA events
around the step, L2 flushed by a 192 MB write before every step); `e2e` = the
same metric through the public API, `Optimizer.optimize(dag)`, with the
host-side request memos dropped before every call (a fresh request), host
buffers in and out; `roofline` = the scan kernel (scan2_kernel launched on its
own, CUDA events) against MEASURED_PEAKS.json; `cpu_baseline` / `--impl
reference` = the UNMODIFIED reference (baseline/_ref, else /root/reference)
through oracle/ref_harness on this box's host cores, on a bounded sample.
"""
import argparse
import contextlib
import hashlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

_REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _REPO)

METRIC = 'candidate (task,instance) placements scored/sec'


# --------------------------------------------------------------------------
def measured_peaks():
    path = os.path.join(_REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path, encoding='utf-8') as f:
            peaks = json.load(f)
        return float(peaks['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def committed_traffic(key: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of one scan launch from
    the committed ncu capture (profiles/round2_traffic.json), or None."""
    try:
        with open(os.path.join(_REPO, 'profiles', 'round2_traffic.json'),
                  encoding='utf-8') as f:
            return json.load(f)[key]['traffic']
    except (OSError, KeyError, ValueError, TypeError):
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the GPU is being timed. One
    low-rate thread (it forks a process per sample: it must not sit in the
    latency-timed loops of eight ranks)."""

    def __init__(self, index: int, period: float = 0.1):
        self.index = index
        self.period = period
        self.samples = []
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, daemon=True)

    def sample(self):
        query = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
                 'clocks_event_reasons.hw_thermal_slowdown,'
                 'clocks_event_reasons.sw_thermal_slowdown,'
                 'clocks_event_reasons.sw_power_cap')
        try:
            out = subprocess.run([
                'nvidia-smi', f'--query-gpu={query}',
                '--format=csv,noheader,nounits', '-i',
                str(self.index)
            ], capture_output=True, text=True, timeout=5,
                                 check=False).stdout.strip()
            if out:
                self.samples.append([v.strip() for v in out.split(',')])
        except Exception:  # pylint: disable=broad-except
            pass

    def _run(self):
        while not self._stop.is_set():
            self.sample()
            self._stop.wait(self.period)

    def __enter__(self):
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                 'sw_power_cap']
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx.append(float(s[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, s[2:]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {
            'sm_mhz': statistics.median(sm) if sm else None,
            'sm_max_mhz': max(mx) if mx else None,
            'reasons': sorted(reasons),
            'samples': len(sm),
        }


# --------------------------------------------------------------------------
# The CPU arm: the unmodified reference through oracle/ref_harness.


def reference_root():
    for path in (os.path.join(_REPO, 'baseline', '_ref'), '/root/reference'):
        if os.path.isdir(os.path.join(path, 'sky')):
            return path
    return None


@contextlib.contextmanager
def stdout_to_stderr():
    """The reference writes progress payloads to stdout; this program's
    stdout carries exactly one JSON line."""
    sys.stdout.flush()
    saved = os.dup(1)
    try:
        os.dup2(2, 1)
        yield
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


class ReferenceArm:
    """`sky.Optimizer.optimize` of the reference on one catalog spec (one per
    process: the reference binds its catalog directory at import)."""

    def __init__(self, spec):
        from oracle.ref_harness import bootstrap  # pylint: disable=import-outside-toplevel
        from oracle.ref_harness import run_reference  # pylint: disable=import-outside-toplevel
        from skypilot_b200 import synth  # pylint: disable=import-outside-toplevel
        root = reference_root()
        if root is None:
            raise RuntimeError('no reference tree (baseline/_ref, '
                               '/root/reference)')
        bootstrap.REFERENCE_ROOT = root
        self.root = root
        self._bootstrap = bootstrap
        self._run = run_reference
        spec = dict(spec)
        enabled = spec.pop('enabled', None) or spec.get('clouds')
        self._home = tempfile.TemporaryDirectory(prefix='skyref_home_')
        frames = synth.make_catalogs(**spec)
        self.n_rows = synth.total_rows(frames)
        bootstrap.write_catalogs(self._home.name, frames)
        with stdout_to_stderr():
            self.sky = bootstrap.import_reference(self._home.name, enabled)
        self.phase_s = {}
        self._wrap_phases()

    def _wrap_phases(self):
        """Per-phase split of the reference's optimize:
        seconds inside `_fill_in_launchable_resources`, `Resources.get_cost`
        and `_optimize_by_dp`, accumulated by thin wrappers (a perf_counter
        pair per call; the functions are millisecond-scale)."""
        from sky import optimizer as opt_lib  # pylint: disable=import-outside-toplevel
        from sky import resources as res_lib  # pylint: disable=import-outside-toplevel
        acc = self.phase_s

        def timed(fn, label):
            def inner(*a, **k):
                t = time.perf_counter()
                try:
                    return fn(*a, **k)
                finally:
                    acc[label] = acc.get(label, 0.0) + time.perf_counter() - t
            return inner

        opt_lib._fill_in_launchable_resources = timed(  # pylint: disable=protected-access
            opt_lib._fill_in_launchable_resources,  # pylint: disable=protected-access
            'fill_in_launchable_resources')
        res_lib.Resources.get_cost = timed(res_lib.Resources.get_cost,
                                           'get_cost')
        opt_lib.Optimizer._optimize_by_dp = staticmethod(timed(  # pylint: disable=protected-access
            opt_lib.Optimizer._optimize_by_dp, 'optimize_by_dp'))  # pylint: disable=protected-access

    def optimize_seconds(self, scenario):
        """One cold `Optimizer.optimize(dag, quiet=True)` (request cache
        cleared, as tests/conftest.py:51-52 of the reference does)."""
        from sky import optimizer as opt_lib  # pylint: disable=import-outside-toplevel
        dag, tasks = self._run.build_dag(self.sky, scenario)
        self._bootstrap.clear_request_cache()
        with stdout_to_stderr():
            t0 = time.perf_counter()
            opt_lib.Optimizer.optimize(dag, quiet=True)
            dt = time.perf_counter() - t0
        return dt, [self._run._res_record(t.best_resources) for t in tasks]  # pylint: disable=protected-access


def port_seconds(spec, scenario):
    """Fallback when no reference tree travelled: the pandas port."""
    from oracle import optimizer_oracle as oo  # pylint: disable=import-outside-toplevel
    oo.catalog_for(spec)
    t0 = time.perf_counter()
    rec = oo.run_scenario(spec, scenario)
    return time.perf_counter() - t0, rec.get('plan')


def sub_chain(scenario, n_tasks):
    sample = dict(scenario)
    sample['tasks'] = scenario['tasks'][:n_tasks]
    sample['edges'] = [[i, i + 1] for i in range(n_tasks - 1)]
    return sample


def run_reference_arm(args, spec, scenario, workload_name, budget_s):
    """Times the reference on a bounded sample: W warm-up and K timed steps,
    each the first `m` tasks of the workload's chain, `m` sized so that the
    whole run fits `budget_s`."""
    n_tasks = len(scenario['tasks'])
    kind = 'reference'
    try:
        arm = ReferenceArm(spec)
        run = arm.optimize_seconds
        n_rows = arm.n_rows
    except Exception as e:  # pylint: disable=broad-except
        kind = f'port (reference unavailable: {type(e).__name__}: {e})'[:200]
        from skypilot_b200 import synth  # pylint: disable=import-outside-toplevel
        n_rows = synth.total_rows(synth.make_catalogs(**spec))
        run = lambda sc: port_seconds(spec, sc)  # noqa: E731
    # one task (its time is not reported) sizes the sample
    t1, _ = run(sub_chain(scenario, 1))
    steps_total = max(1, args.steps + args.warmup)
    m = int(max(1, min(n_tasks, (budget_s / steps_total) // max(t1, 1e-3))))
    sample = sub_chain(scenario, m)
    for _ in range(args.warmup):
        run(sample)
    times, plan = [], None
    phases = getattr(arm, 'phase_s', None) if kind == 'reference' else None
    if phases is not None:
        phases.clear()
    for _ in range(args.steps):
        dt, plan = run(sample)
        times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = n_rows * m / (ms / 1e3)
    return {
        'value': value, 'ms_per_step': ms,
        'p50_ms': 1e3 * statistics.median(times),
        'cpu_baseline': {
            'value': value, 'unit': 'candidates/s', 'cores': os.cpu_count(),
            'kind': kind,
            'threads': ('the reference\'s own: single-threaded pandas plus a '
                        'ThreadPool of max(4, cores-1) across clouds '
                        '(sky/optimizer.py:1712-1715)'),
            'sample': (f'{m} of {n_tasks} tasks of the {workload_name} chain '
                       f'per step, {args.steps} steps, full catalog '
                       f'({n_rows} rows), request cache cleared every call'),
            'ms_per_step': ms,
            'phases_ms_per_step': ({k: 1e3 * v / len(times)
                                    for k, v in phases.items()}
                                   if phases else None),
        },
        'plan': plan, 'sample_tasks': m,
    }


# --------------------------------------------------------------------------
def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--gpus', type=int, default=1)
    parser.add_argument('--steps', type=int, default=20)
    parser.add_argument('--warmup', type=int, default=5)
    parser.add_argument('--workload', default='cfg4', choices=['cfg2', 'cfg4'])
    parser.add_argument('--impl', default='ours',
                        choices=['ours', 'reference'])
    parser.add_argument('--cpu-baseline-budget', type=float, default=25.0,
                        help='seconds of reference work in the cpu_baseline leg')
    parser.add_argument('--reference-budget', type=float, default=200.0,
                        help='seconds the --impl reference run may take')
    parser.add_argument('--no-extras', action='store_true',
                        help='skip cfg2 latency, the HBM stress row, cfg5 and '
                        'the CPU baseline (tuning runs)')
    parser.add_argument('--cfg5-dags', type=int, default=10000)
    parser.add_argument('--scan-mode', default='auto',
                        help='kernel selection for the headline loop (tests / '
                        'tuning; see CatalogStore.set_scan_mode)')
    args = parser.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))

    from skypilot_b200 import workloads  # pylint: disable=import-outside-toplevel
    spec = workloads.CATALOGS[args.workload]
    n_tasks = 32 if args.workload == 'cfg4' else 8
    scenario = workloads.chain_scenario(n_tasks)
    desc = (f'{args.workload}: {n_tasks}-task chain DAG, synthetic '
            f'{"1M" if args.workload == "cfg4" else "~50k"}-row catalog')

    if args.impl == 'reference':
        if rank != 0:
            return
        ref = run_reference_arm(args, spec, scenario, args.workload,
                                args.reference_budget)
        line = {
            'metric': METRIC, 'value': ref['value'], 'unit': 'candidates/s',
            'impl': 'reference', 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ref['ms_per_step'],
            'p50_ms': ref['p50_ms'], 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic',
            'config': {'workload': desc, 'catalog': spec, 'tasks': n_tasks},
            'cpu_baseline': ref['cpu_baseline'],
            'e2e': {'value': ref['value'], 'unit': 'candidates/s',
                    'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0,
        }
        print(json.dumps(line), flush=True)
        return

    dist = None
    if world > 1:
        import torch  # pylint: disable=import-outside-toplevel
        import torch.distributed as dist  # pylint: disable=import-outside-toplevel
        torch.cuda.set_device(local_rank)
        dist.init_process_group('nccl', device_id=torch.device(
            'cuda', local_rank))
        # keep each rank's host threads on its own share of the cores
        try:
            cores = sorted(os.sched_getaffinity(0))
            share = max(1, len(cores) // world)
            os.sched_setaffinity(
                0, cores[local_rank * share:(local_rank + 1) * share] or cores)
        except (AttributeError, OSError):
            pass

    import numpy as np  # pylint: disable=import-outside-toplevel
    import networkx as nx  # pylint: disable=import-outside-toplevel
    import skypilot_b200 as sky  # pylint: disable=import-outside-toplevel
    from skypilot_b200 import engine, sharding, synth  # pylint: disable=import-outside-toplevel
    from skypilot_b200 import optimizer as opt_lib  # pylint: disable=import-outside-toplevel
    from skypilot_b200.catalog.store import CatalogStore  # pylint: disable=import-outside-toplevel
    Optimizer = opt_lib.Optimizer

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        return sharding.max_over_ranks(x, dist, device='cuda')

    def load_store(catalog_spec):
        """Rank 0 ingests the synthetic catalog and leaves it in the columnar
        cache; the other ranks map the cache (CatalogStore.save / load)."""
        key = hashlib.sha256(json.dumps(
            catalog_spec, sort_keys=True).encode()).hexdigest()[:16]
        cache = os.path.join(tempfile.gettempdir(), f'skyopt_bench_{key}')
        store = None
        if rank == 0:
            frames = synth.make_catalogs(**catalog_spec)
            store = CatalogStore.from_frames(frames)
            if world > 1:
                store.save(cache)
        barrier()
        if store is None:
            store = CatalogStore.load(cache)
        sky.catalog.set_store(store, local_rank)
        sky.check.set_enabled_clouds(sky.check.ALL_CATALOG_CLOUDS)
        store.handle(local_rank)
        return store

    def drop_memos(tasks):
        """A fresh request: nothing stated before is replayed (SURVEY.md
        section 8d item 2, "request cache cleared each call")."""
        sky.catalog.clear_request_level_cache()
        for t in tasks:
            t.__dict__.pop('_stated', None)
            for r in t.resources:
                for k in ('_request_key', '_validated_store',
                          '_plan_templates'):
                    r.__dict__.pop(k, None)

    def state(dag):
        Optimizer._add_dummy_source_sink_nodes(dag)  # pylint: disable=protected-access
        try:
            graph = dag.get_graph()
            topo = [t for t in nx.topological_sort(graph)
                    if not opt_lib._is_dummy(t)]  # pylint: disable=protected-access
            return Optimizer._state_problem(graph, topo, True, [], True)  # pylint: disable=protected-access
        finally:
            Optimizer._remove_dummy_source_sink_nodes(dag)  # pylint: disable=protected-access

    def measure_chain(store, chain, n_rows):
        """Device-resident loop, roofline loop, end-to-end loops of one chain
        DAG on the active catalog."""
        n_candidates = n_rows * len(chain['tasks'])
        dag, tasks = workloads.build_dag(chain)
        builder = state(dag).builder
        if args.scan_mode != 'auto':
            store.set_scan_mode(args.scan_mode, local_rank)
        engine.solve_timed(builder, args.warmup, True, local_rank)
        barrier()
        t0 = time.perf_counter()
        sol, iter_ms, _ = engine.solve_timed(builder, args.steps, True,
                                             local_rank)
        wall = time.perf_counter() - t0
        device_ms = max_over_ranks(float(np.sum(iter_ms)))
        stats = sol.stats
        out = {
            'ms_per_step': device_ms / args.steps,
            'value': world * n_candidates / (device_ms / args.steps / 1e3),
            'launches_per_step': int(stats.total_launches),
            'scan_form': int(stats.scan_form),
            'wall_check_ms_per_step': 1e3 * wall / args.steps,
            'n_candidates': n_candidates,
        }
        assert sol.dag[0]['status'] == 0
        # the scan kernel on its own (separate launches, events around it)
        store.set_scan_mode('fast-split', local_rank)
        engine.solve_timed(builder, args.warmup, True, local_rank)
        s2, _, scan_ms = engine.solve_timed(builder, args.steps, True,
                                            local_rank)
        store.set_scan_mode('auto', local_rank)
        st2 = s2.stats
        out['split'] = {
            'scan_kernel_ms': float(np.mean(scan_ms)),
            'scan_kernel_ms_min': float(np.min(scan_ms)),
            'phases_ms': {'scan': float(st2.scan_ms),
                          'place': float(st2.expand_ms),
                          'solve': float(st2.solve_ms)},
            'scan_form': int(st2.scan_form),
            'pass_rows': int(st2.scan_passes_rows),
            'layout_rows': int(st2.reserved_),
        }
        barrier()
        # ---- end to end through the public API, host buffers in and out,
        # nothing replayed from an earlier call
        for _ in range(args.warmup):
            drop_memos(tasks)
            Optimizer.optimize(dag, quiet=True)
        barrier()
        cold = []
        t_e2e0 = time.perf_counter()
        for _ in range(args.steps):
            drop_memos(tasks)
            t1 = time.perf_counter()
            Optimizer.optimize(dag, quiet=True)
            cold.append(time.perf_counter() - t1)
        e2e_total = max_over_ranks(time.perf_counter() - t_e2e0)
        barrier()
        while len(cold) < 60:
            drop_memos(tasks)
            t1 = time.perf_counter()
            Optimizer.optimize(dag, quiet=True)
            cold.append(time.perf_counter() - t1)
        warm = []
        for _ in range(100):
            t1 = time.perf_counter()
            Optimizer.optimize(dag, quiet=True)
            warm.append(time.perf_counter() - t1)
        # a new request with the same content: new Task / Resources objects,
        # the content-keyed statement memo of the catalog store kept
        fresh = []
        for _ in range(60):
            fdag, _ = workloads.build_dag(chain)
            t1 = time.perf_counter()
            Optimizer.optimize(fdag, quiet=True)
            fresh.append(time.perf_counter() - t1)
        packed = builder.pack()
        out.update({
            'e2e': {
                'value': world * n_candidates * args.steps / e2e_total,
                'unit': 'candidates/s',
                'h2d_bytes_per_step': packed.h2d_bytes(),
                'd2h_bytes_per_step': sol.d2h_bytes(),
                'ms_per_step': 1e3 * e2e_total / args.steps,
                'what': ('Optimizer.optimize(dag), host-side request memos '
                         'dropped before every call'),
            },
            'optimize_cold_p50_ms': 1e3 * statistics.median(cold),
            'optimize_cold_p90_ms': 1e3 * sorted(cold)[int(0.9 * len(cold))],
            'optimize_warm_p50_ms': 1e3 * statistics.median(warm),
            'optimize_fresh_request_p50_ms': 1e3 * statistics.median(fresh),
            'optimize_calls': len(cold) + len(warm),
            'plan': [workloads.res_record(t.best_resources) for t in tasks],
        })
        return out

    def roofline_of(split, traffic_key=None):
        peak, peak_src = measured_peaks()
        ms = split['scan_kernel_ms']
        rows = split['pass_rows']
        alg40 = rows * 40          # 40 B per (row, pass)
        layout = split['layout_rows'] * 10
        traffic = committed_traffic(traffic_key) if traffic_key else None
        gbs = lambda b: b / (ms / 1e3) / 1e9 if ms else 0.0  # noqa: E731
        return {
            'kernel': 'scan2_kernel', 'bound': 'hbm', 'unit': 'GB/s',
            'achieved': gbs(alg40), 'peak': peak, 'peak_source': peak_src,
            'frac': gbs(alg40) / peak,
            'algorithmic_bytes_per_launch': alg40,
            'bytes_per_row': 40,
            'bytes_per_row_note': (
                'SURVEY.md section 8d figure (the reference columns a row '
                'contributes to the predicate); rows x minimum passes '
                '(ceil(queries / 32) per cloud)'),
            'rows_streamed_per_launch': rows,
            'layout_bytes_per_launch': layout,
            'layout_note': ('what the kernel has to move: 10 B per row '
                            '(u32 rank + 3 u16 class codes) x one pass per '
                            'query group of <= 32 queries and price column'),
            'layout_frac': gbs(layout) / peak,
            'kernel_ms': ms,
            'traffic': traffic,
            'dram_frac': (gbs(traffic) / peak) if traffic else None,
            'traffic_source': ('ncu dram__bytes_read.sum + '
                               'dram__bytes_write.sum per launch, '
                               'profiles/round2_traffic.json'),
        }

    # ------------------------------------------------------------ headline
    store = load_store(spec)
    n_rows = store.n_real_rows
    with ClockSampler(local_rank) as clocks:
        clocks.sample()
        head = measure_chain(store, scenario, n_rows)
        clocks.sample()
    split = head.pop('split')
    line = {
        'metric': METRIC, 'value': head['value'], 'unit': 'candidates/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': head['ms_per_step'], 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {
            'workload': desc, 'catalog': spec, 'catalog_rows': n_rows,
            'catalog_rows_note': ('the generator is asked for n_rows and '
                                  'keeps whole instance types'),
            'tasks': n_tasks, 'candidates_per_step': head['n_candidates'],
            'l2': 'flushed before every step (192 MB write)',
            'parallelism': ('one replica of the DAG and of the catalog per '
                            'GPU, no collective; the DAG-sharded batch is '
                            'under scaling_cfg5'),
            'kernels': {3: 'scan2 + place + solve (separate launches)',
                        4: 'step_kernel (one cooperative launch)'}.get(
                            head['scan_form'], 'round-1 kernels'),
        },
        'e2e': head['e2e'],
        'gpu_launches': head['launches_per_step'] * args.steps,
        'launches_per_step': head['launches_per_step'],
        'optimize_cold_p50_ms': head['optimize_cold_p50_ms'],
        'optimize_cold_p90_ms': head['optimize_cold_p90_ms'],
        'optimize_warm_p50_ms': head['optimize_warm_p50_ms'],
        'optimize_fresh_request_p50_ms': head['optimize_fresh_request_p50_ms'],
        'optimize_calls': head['optimize_calls'],
        'roofline': roofline_of(split, traffic_key=f'{args.workload}_scan2'),
        'phases_ms_separate_launches': split['phases_ms'],
        'clocks': clocks.summary(),
        'wall_check_ms_per_step': head['wall_check_ms_per_step'],
    }
    line['roofline']['step_frac'] = (
        line['roofline']['algorithmic_bytes_per_launch'] /
        (head['ms_per_step'] / 1e3) / 1e9 / line['roofline']['peak'])
    plan = head['plan']
    line['plan'] = [p['instance_type'] for p in plan]

    # parity of the timed workload against the reference's record
    golden_path = os.path.join(_REPO, 'tests', 'golden', 'cfg4_1m.json')
    if args.workload == 'cfg4' and os.path.exists(golden_path):
        with open(golden_path, encoding='utf-8') as f:
            rec = next(r for r in json.load(f)['records']
                       if r['name'] == scenario['name'])
        keys = ('cloud', 'instance_type', 'region', 'zone')
        line['plan_matches_reference'] = (
            [[p[k] for k in keys] for p in plan] ==
            [[p[k] for k in keys] for p in rec['plan']])
        line['reference_record'] = 'tests/golden/cfg4_1m.json:chain32'

    extras = not args.no_extras
    store5 = None
    # ---------------------------------------------------------- cfg5 (all N)
    if extras:
        store5 = store if args.workload == 'cfg2' else load_store(
            workloads.CATALOGS['cfg2'])
        rows5 = store5.n_real_rows
        scs = workloads.cfg5_scenarios(args.cfg5_dags)
        dags = [workloads.build_dag(sc)[0] for sc in scs]
        sharding.optimize_shard(dags[:200 * world], rank, world, local_rank,
                                return_exceptions=True)
        best, out = None, None
        for _ in range(3):
            for d in dags:
                drop_memos(d.tasks)
            barrier()
            t0 = time.perf_counter()
            out = sharding.optimize_shard(dags, rank, world, local_rank,
                                          return_exceptions=True)
            dt = max_over_ranks(time.perf_counter() - t0)
            best = dt if best is None else min(best, dt)
        mine = sharding.shard(dags, rank, world)
        records = []
        for d, res in zip(mine, out):
            if isinstance(res, Exception):
                records.append(None)
            else:
                r = d.tasks[0].best_resources
                records.append([str(r.cloud).lower(), r.instance_type,
                                r.region, r.zone])
        gathered = sharding.gather_on_root(records, len(dags), dist)
        if rank == 0:
            info = {
                'workload': (f'cfg5: {len(dags)} single-task DAGs, cfg2 '
                             f'catalog ({rows5} rows), DAG i -> rank i mod N'),
                'scaling': 'strong', 'n_gpus': world, 'seconds': best,
                'dags_per_s': len(dags) / best,
                'value': len(dags) * rows5 / best, 'unit': 'candidates/s',
                'what': ('Optimizer.optimize_batch per rank (fresh requests: '
                         'memos dropped), best of 3, max over ranks; end to '
                         'end, host buffers'),
                'feasible': sum(r is not None for r in gathered),
            }
            g5 = os.path.join(_REPO, 'tests', 'golden', 'cfg5_50k.json')
            if os.path.exists(g5) and len(dags) <= 10000:
                with open(g5, encoding='utf-8') as f:
                    recs = json.load(f)['records'][:len(dags)]
                bad = sum(
                    (got is None) != ('error' in rec) or
                    (got is not None and got != rec['plan'])
                    for got, rec in zip(gathered, recs))
                info['plans_matching_reference'] = len(dags) - bad
                info['reference_record'] = 'tests/golden/cfg5_50k.json'
            line['scaling_cfg5'] = info

    # ------------------------------------------------------- N = 1 extras
    if extras and world == 1:
        if args.workload == 'cfg4':
            # cfg2: the optimize() latency configuration (store5 is active)
            sc2 = workloads.chain_scenario(8)
            lat = measure_chain(store5, sc2, store5.n_real_rows)
            sp2 = lat.pop('split')
            line['latency_cfg2'] = {
                'workload': 'cfg2: 8-task chain DAG, ~50k-row catalog',
                'ms_per_step': lat['ms_per_step'], 'value': lat['value'],
                'unit': 'candidates/s', 'e2e': lat['e2e'],
                'optimize_cold_p50_ms': lat['optimize_cold_p50_ms'],
                'optimize_cold_p90_ms': lat['optimize_cold_p90_ms'],
                'optimize_warm_p50_ms': lat['optimize_warm_p50_ms'],
                'optimize_fresh_request_p50_ms':
                    lat['optimize_fresh_request_p50_ms'],
                'roofline': roofline_of(sp2, traffic_key='cfg2_scan2'),
                'phases_ms_separate_launches': sp2['phases_ms'],
            }
            sky.catalog.set_store(store, local_rank)
        # HBM stress: 8 copies of every row, nothing pruned
        big = store.replicated(8)
        sky.catalog.set_store(big, local_rank)
        big.handle(local_rank)
        dag, tasks = workloads.build_dag(scenario)
        builder = state(dag).builder
        big.set_scan_mode('fast-split-noprune', local_rank)
        engine.solve_timed(builder, args.warmup, True, local_rank)
        s3, _, scan_ms = engine.solve_timed(builder, args.steps, True,
                                            local_rank)
        st3 = s3.stats
        sp3 = {'scan_kernel_ms': float(np.mean(scan_ms)),
               'pass_rows': int(st3.scan_passes_rows),
               'layout_rows': int(st3.reserved_)}
        stress = roofline_of(sp3, traffic_key='stress_scan2')
        drop_memos(tasks)
        Optimizer.optimize(dag, quiet=True)
        stress.update({
            'workload': (f'{desc.split(":")[0]} chain on the catalog with '
                         f'every row repeated 8 times ({big.n_real_rows} '
                         'rows), scan2_kernel without zone map or bound: '
                         'every row of every pass is streamed and scored'),
            'catalog_rows': big.n_real_rows,
            'candidates_per_s': big.n_real_rows * n_tasks /
                                (sp3['scan_kernel_ms'] / 1e3),
            'plan_equals_base_catalog': [
                workloads.res_record(t.best_resources)['instance_type']
                for t in tasks] == line['plan'],
        })
        line['hbm_stress'] = stress
        big.close()
        sky.catalog.set_store(store, local_rank)
        # CPU baseline: the reference on a bounded sample of this workload
        ref_args = argparse.Namespace(steps=1, warmup=0)
        ref = run_reference_arm(ref_args, spec, scenario, args.workload,
                                args.cpu_baseline_budget)
        line['cpu_baseline'] = ref['cpu_baseline']
        if ref['plan'] is not None:
            # the sample is a shorter chain with its own optimum: compare it
            # with OUR plan of that same sub-chain
            sdag, stasks = workloads.build_dag(
                sub_chain(scenario, ref['sample_tasks']))
            Optimizer.optimize(sdag, quiet=True)
            line['cpu_baseline']['sample_plan_equals_ours'] = (
                [workloads.res_record(t.best_resources) for t in stasks] ==
                ref['plan'])
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
