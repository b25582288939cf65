"""Wall-clock split of one cold Optimizer.optimize(dag) (GPU box): each named
function is wrapped with a perf_counter pair; nested sections are reported
inclusive.

    python tools/e2e_sections.py cfg2|cfg4
"""
import collections
import statistics
import sys
import time
sys.path.insert(0, '.')
import skypilot_b200 as sky  # noqa: E402
from skypilot_b200 import engine, synth, workloads  # noqa: E402
from skypilot_b200 import optimizer as opt_lib  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
w = workloads.WORKLOADS[name]
sky.catalog.load_frames(synth.make_catalogs(**w['catalog']))
sky.check.set_enabled_clouds(sky.check.ALL_CATALOG_CLOUDS)
scenario = workloads.chain_scenario(w['tasks'])
dag, tasks = workloads.build_dag(scenario)
acc = collections.defaultdict(list)
cur = collections.defaultdict(float)


def wrap(owner, attr, label, static=False):
    fn = getattr(owner, attr)

    def inner(*a, **k):
        t = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            cur[label] += time.perf_counter() - t
    setattr(owner, attr, staticmethod(inner) if static else inner)


O = opt_lib.Optimizer
wrap(O, '_state_problem', 'state', True)
wrap(O, '_add_dummy_source_sink_nodes', 'dummy nodes', True)
wrap(O, '_remove_dummy_source_sink_nodes', 'dummy nodes', True)
wrap(engine, 'solve', 'engine.solve (pack + native call + unpack)')
wrap(engine.ProblemBuilder, 'pack', '  pack')
wrap(opt_lib._Problem, 'launchable', 'launchable x tasks')
wrap(opt_lib.nx, 'topological_sort', 'topological_sort')
native = engine._native.load()
orig = native.skyopt_optimize


def timed_native(*a):
    t = time.perf_counter()
    r = orig(*a)
    cur['  native skyopt_optimize'] += time.perf_counter() - t
    return r


engine._native.load().skyopt_optimize = timed_native


def drop():
    sky.catalog.clear_request_level_cache()
    for t in tasks:
        t.__dict__.pop('_stated', None)
        for r in t.resources:
            for k in ('_request_key', '_validated_store', '_plan_templates'):
                r.__dict__.pop(k, None)


for i in range(260):
    drop()
    cur.clear()
    t0 = time.perf_counter()
    sky.optimize(dag, quiet=True)
    cur['TOTAL'] = time.perf_counter() - t0
    if i >= 60:
        for k, v in cur.items():
            acc[k].append(v)
print(name, 'cold optimize, p50 over 200 calls (ms)')
for k in sorted(acc, key=lambda k: -statistics.median(acc[k])):
    print(f'  {k:48s} {1e3 * statistics.median(acc[k]):7.3f}')
