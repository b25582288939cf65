"""bench.py -- candidates scored / second of the placement-optimizer hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W]
                    [--workload cfg2|cfg4] [--impl ours|reference]

One *step* = one pass of the hot path over one DAG: every task's constraint
vector is scored against every catalog row (filter + argmin), the winners are
expanded to region/zone candidates and costed, and the chain DP picks the
plan. Workloads (BASELINE.json `configs`):

  cfg4  32-task chain DAG, synthetic 1M-row catalog -- the configuration the
        throughput / HBM-roofline target is quoted on (BASELINE.md section 4,
        item 4; BASELINE.json configs[3])                             [default]
  cfg2  8-task chain DAG, synthetic multi-cloud catalog (~50k rows) -- the
        optimize() p50 latency configuration (configs[1]); measured in the
        same run and reported under "latency_cfg2"

`value`  = candidates / s with everything resident in HBM (CUDA events around
           the kernels; L2 is flushed by writing 192 MB before every step);
`e2e`    = the same metric through the public API, Optimizer.optimize(dag):
           host Python, one H2D of the problem, one D2H of the plan per step;
`roofline` is for the dominant kernel (scan_kernel): algorithmic bytes of one
           launch / its event-timed duration vs MEASURED_PEAKS.json;
`cpu_baseline` times the pandas oracle (oracle/, a port of the reference's
           algorithm) on this box's host cores on a bounded sample.

Multi-GPU (torchrun, one rank per GPU): independent DAGs per GPU, catalog
replicated, no collective on the data path; weak scaling.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

_REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _REPO)

WORKLOADS = {
    'cfg2': {
        'catalog': {'seed': 1, 'n_rows': 50000,
                    'clouds': ['aws', 'gcp', 'azure', 'lambda']},
        'tasks': 8,
        'desc': '8-task chain DAG, synthetic multi-cloud catalog (~50k rows)',
    },
    'cfg4': {
        'catalog': {'seed': 3, 'n_rows': 1000000,
                    'clouds': ['aws', 'gcp', 'azure', 'lambda']},
        'tasks': 32,
        'desc': '32-task chain DAG, synthetic 1M-row catalog',
    },
}


def chain_scenario(n_tasks: int):
    """The cfg2 constraint set, cycled with varying thresholds (cfg4)."""
    from skypilot_b200 import workloads  # pylint: disable=import-outside-toplevel
    return workloads.chain_scenario(n_tasks)


def measured_peaks():
    path = os.path.join(_REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path, encoding='utf-8') as f:
            peaks = json.load(f)
        return float(peaks['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def measured_traffic(workload_name: str, scan_form: int = 0):
    """dram__bytes_read.sum + dram__bytes_write.sum of one scan launch, from
    the committed ncu captures (profiles/round1_traffic.json); None if no
    capture covers this workload with the scan kernel that ran."""
    path = os.path.join(_REPO, 'profiles', 'round1_traffic.json')
    try:
        with open(path, encoding='utf-8') as f:
            table = json.load(f)
        entry = table[workload_name]
        form = {'scan_queue_kernel': 2, 'scan_kernel': 0}.get(
            entry.get('kernel', 'scan_kernel'))
        if form != scan_form:
            entry = table[f'{workload_name}_one_tile_per_block']
            if scan_form != 0:
                return None
        return entry['traffic']
    except (OSError, KeyError, ValueError):
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        query = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
                 'clocks_event_reasons.hw_thermal_slowdown,'
                 'clocks_event_reasons.sw_thermal_slowdown,'
                 'clocks_event_reasons.sw_power_cap')
        while not self._stop.is_set():
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
            self._stop.wait(0.01)

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


def bench_reference(args, workload, scenario, n_candidates):
    """`--impl reference`: the CPU arm. /root/reference is Python and does not
    travel to the GPU box, so this times oracle/ -- the pandas port of the
    reference's algorithm (kind "port"), single-threaded like the reference's
    catalog path -- on a bounded sample of the same workload."""
    from oracle import optimizer_oracle as oo  # pylint: disable=import-outside-toplevel
    spec = workload['catalog']
    sample = scenario
    sample_tasks = len(scenario['tasks'])
    if args.workload == 'cfg4':
        # one task of the 32 per step keeps the run within minutes
        sample = chain_scenario(32)
        sample['tasks'] = sample['tasks'][:2]
        sample['edges'] = [[0, 1]]
        sample_tasks = 2
    oo.catalog_for(spec)
    for _ in range(max(1, min(args.warmup, 2))):
        oo.run_scenario(spec, sample)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        rec = oo.run_scenario(spec, sample)
        times.append(time.perf_counter() - t0)
        assert 'error' not in rec, rec
    ms = 1e3 * sum(times) / len(times)
    cands = n_candidates * sample_tasks / len(scenario['tasks'])
    value = cands / (ms / 1e3)
    return {
        'metric': 'candidate (task,instance) placements scored/sec',
        'value': value, 'unit': 'candidates/s', 'impl': 'reference',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms, 'p50_ms': 1e3 * statistics.median(times),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': f'{args.workload}: {workload["desc"]}',
                   'catalog': spec},
        'cpu_baseline': {
            'value': value, 'unit': 'candidates/s',
            'cores': 1, 'kind': 'port',
            'threads': ('single-threaded pandas, like the reference except for '
                        'its ThreadPool fan-out across clouds, which is slower '
                        'under the GIL (measured 1.73 s vs 1.54 s per cfg2 step; '
                        'SKYOPT_ORACLE_THREADS=1 enables it)'),
            'sample': (f'{sample_tasks} of {len(scenario["tasks"])} tasks per '
                       f'step, {args.steps} steps, full catalog')
        },
        'e2e': {'value': value, 'unit': 'candidates/s',
                'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--gpus', type=int, default=1)
    parser.add_argument('--steps', type=int, default=30)
    parser.add_argument('--warmup', type=int, default=5)
    parser.add_argument('--workload', default='cfg4', choices=list(WORKLOADS))
    parser.add_argument('--impl', default='ours',
                        choices=['ours', 'reference'])
    parser.add_argument('--cpu-baseline-steps', type=int, default=5)
    parser.add_argument('--no-latency', action='store_true',
                        help='skip the extra cfg2 (optimize() p50) run')
    parser.add_argument('--scan-mode', default='auto',
                        choices=['auto', 'tile', 'stream', 'stream3', 'queue', 'fast',
                                 'fast-noprune'],
                        help='scan kernel variant (tuning / tests)')
    args = parser.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    workload = WORKLOADS[args.workload]
    scenario = chain_scenario(workload['tasks'])

    from skypilot_b200 import synth  # pylint: disable=import-outside-toplevel
    if args.impl == 'reference':
        if rank != 0:
            return
        frames = synth.make_catalogs(**workload['catalog'])
        n_rows = synth.total_rows(frames)
        line = bench_reference(args, workload, scenario,
                               n_rows * workload['tasks'])
        print(json.dumps(line), flush=True)
        return

    dist = None
    if world > 1:
        import torch  # pylint: disable=import-outside-toplevel
        import torch.distributed as dist  # pylint: disable=import-outside-toplevel
        torch.cuda.set_device(local_rank)
        dist.init_process_group('nccl', device_id=torch.device(
            'cuda', local_rank))

    import numpy as np  # pylint: disable=import-outside-toplevel
    import skypilot_b200 as sky  # pylint: disable=import-outside-toplevel
    from skypilot_b200 import engine  # pylint: disable=import-outside-toplevel
    from skypilot_b200 import optimizer as opt_lib  # pylint: disable=import-outside-toplevel
    from skypilot_b200 import workloads as runner  # pylint: disable=import-outside-toplevel
    import networkx as nx  # pylint: disable=import-outside-toplevel

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        from skypilot_b200 import sharding  # pylint: disable=import-outside-toplevel
        return sharding.max_over_ranks(x, dist, device='cuda')

    def measure(workload_name):
        workload = WORKLOADS[workload_name]
        scenario = chain_scenario(workload['tasks'])
        frames = synth.make_catalogs(**workload['catalog'])
        n_rows = synth.total_rows(frames)
        store = sky.catalog.load_frames(frames, device=local_rank)
        store.handle(local_rank)
        if args.scan_mode != 'auto':
            store.set_scan_mode(args.scan_mode, local_rank)
        n_tasks = workload['tasks']
        n_candidates = n_rows * n_tasks

        dag, tasks = runner.build_dag(scenario)
        Optimizer = opt_lib.Optimizer

        # ---- device-resident arm: the problem is uploaded once, kernels re-run
        Optimizer._add_dummy_source_sink_nodes(dag)  # pylint: disable=protected-access
        try:
            graph = dag.get_graph()
            topo = [t for t in nx.topological_sort(graph)
                    if not opt_lib._is_dummy(t)]  # pylint: disable=protected-access
            problem = Optimizer._state_problem(graph, topo, True, [], True)  # pylint: disable=protected-access
        finally:
            Optimizer._remove_dummy_source_sink_nodes(dag)  # pylint: disable=protected-access
        engine.solve_timed(problem.builder, args.warmup, True, local_rank)
        barrier()
        with ClockSampler(local_rank) as clocks:
            t0 = time.perf_counter()
            sol, iter_ms, scan_ms = engine.solve_timed(problem.builder, args.steps,
                                                       True, local_rank)
            wall_resident = time.perf_counter() - t0
            device_ms = float(np.sum(iter_ms))
            device_ms = max_over_ranks(device_ms)
            barrier()

            # ---- end to end through the public API, host buffers in and out
            for _ in range(args.warmup):
                Optimizer.optimize(dag, quiet=True)
            barrier()
            e2e_times = []
            t_e2e0 = time.perf_counter()
            for _ in range(args.steps):
                t1 = time.perf_counter()
                Optimizer.optimize(dag, quiet=True)
                e2e_times.append(time.perf_counter() - t1)
            e2e_total = time.perf_counter() - t_e2e0
            e2e_total = max_over_ranks(e2e_total)
            barrier()
            # the latency percentiles are taken over at least 100 calls
            # (SURVEY.md section 8d item 2); the e2e rate above is over
            # exactly `steps`
            while len(e2e_times) < 100:
                t1 = time.perf_counter()
                Optimizer.optimize(dag, quiet=True)
                e2e_times.append(time.perf_counter() - t1)
            # Same call with the host-side request memos dropped before every
            # call ("request cache cleared each call").
            cold_times = []
            for _ in range(args.steps):
                sky.catalog.clear_request_level_cache()
                for t in tasks:
                    for r in t.resources:
                        r.__dict__.pop('_request_key', None)
                        r.__dict__.pop('_validated_store', None)
                        r.__dict__.pop('_plan_templates', None)
                t1 = time.perf_counter()
                Optimizer.optimize(dag, quiet=True)
                cold_times.append(time.perf_counter() - t1)
        assert sol.dag[0]['status'] == 0
        plan = [runner.res_record(t.best_resources) for t in tasks]

        ms_per_step = device_ms / args.steps
        value = world * n_candidates / (ms_per_step / 1e3)
        e2e_value = world * n_candidates * args.steps / e2e_total
        packed = problem.builder.pack()
        stats = sol.stats
        scan_kernel_ms = float(np.mean(scan_ms))
        row_bytes = store.row_bytes()
        scan_bytes = int(stats.scan_passes_rows) * row_bytes
        peak, peak_src = measured_peaks()
        achieved = scan_bytes / (scan_kernel_ms / 1e3) / 1e9 if scan_kernel_ms else 0

        scan_form = min(3, max(0, int(getattr(stats, 'scan_form', 0))))

        line = {
            'metric': 'candidate (task,instance) placements scored/sec',
            'value': value, 'unit': 'candidates/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_per_step, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic',
            'config': {
                'workload': f'{workload_name}: {workload["desc"]}',
                'catalog': workload['catalog'], 'catalog_rows': n_rows,
                'tasks': n_tasks, 'candidates_per_step': n_candidates,
                'l2': 'flushed before every step (192 MB write)',
                'parallelism': ('one DAG stream per GPU, catalog replicated, '
                                'no collective'),
            },
            'optimize_calls': len(e2e_times),
            'optimize_p50_ms': 1e3 * statistics.median(e2e_times),
            'optimize_p90_ms': 1e3 * sorted(e2e_times)[int(0.9 *
                                                           len(e2e_times))],
            'optimize_cold_p50_ms': 1e3 * statistics.median(cold_times),
            'e2e': {
                'value': e2e_value, 'unit': 'candidates/s',
                'h2d_bytes_per_step': packed.h2d_bytes(),
                'd2h_bytes_per_step': sol.d2h_bytes(),
                'ms_per_step': 1e3 * e2e_total / args.steps,
            },
            'gpu_launches': int(stats.total_launches) * args.steps,
            'phases_ms': {
                'scan_kernel': scan_kernel_ms,
                'scan_total': float(stats.scan_ms),
                'expand': float(stats.expand_ms),
                'solve': float(stats.solve_ms),
            },
            'roofline': {
                'kernel': ('scan_kernel', 'scan_stream_kernel',
                           'scan_queue_kernel', 'scan2_kernel')[scan_form],
                'bound': 'hbm', 'achieved': achieved,
                'peak': peak, 'peak_source': peak_src, 'unit': 'GB/s',
                'frac': achieved / peak if peak else None,
                'algorithmic_bytes_per_launch': scan_bytes,
                'bytes_per_row': row_bytes,
                'rows_streamed_per_launch': int(stats.scan_passes_rows),
                'queries_fused_per_pass': 32,
                'traffic': measured_traffic(workload_name, scan_form),
                'traffic_source': 'ncu, per launch: profiles/round1_traffic.json',
            },
            'clocks': clocks.summary(),
            'wall_check_ms_per_step': 1e3 * wall_resident / args.steps,
            'plan': [p['instance_type'] for p in plan],
        }

        return line, scenario, n_candidates

    line, scenario, n_candidates = measure(args.workload)
    workload = WORKLOADS[args.workload]
    if args.workload != 'cfg2' and not args.no_latency:
        # the latency configuration (BASELINE.json configs[1]) next to the
        # throughput / roofline configuration
        lat, _, _ = measure('cfg2')
        line['latency_cfg2'] = {
            k: lat[k] for k in ('value', 'unit', 'ms_per_step', 'e2e',
                                'roofline', 'phases_ms', 'config',
                                'optimize_p50_ms', 'optimize_p90_ms',
                                'optimize_cold_p50_ms')
        }
    if rank == 0 and world == 1:
        # ---- CPU baseline: the pandas oracle on a bounded sample
        ref_args = argparse.Namespace(**vars(args))
        ref_args.steps = args.cpu_baseline_steps
        ref_args.warmup = 1
        ref = bench_reference(ref_args, workload, scenario, n_candidates)
        line['cpu_baseline'] = ref['cpu_baseline']
        line['cpu_baseline']['ms_per_step'] = ref['ms_per_step']
        # the oracle's plan must be ours (cheap cross-check, cfg2 only)
        if args.workload == 'cfg2':
            from oracle import optimizer_oracle as oo  # pylint: disable=import-outside-toplevel
            want = oo.run_scenario(workload['catalog'], scenario)
            line['plan_matches_oracle'] = (
                [p['instance_type'] for p in want['plan']] == line['plan'])
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
