"""Wall-clock split of Optimizer.optimize_batch on the cfg5 batch (GPU box).

    python tools/batch_sections.py [n_dags]
"""
import collections
import sys
import time
sys.path.insert(0, '.')
import skypilot_b200 as sky  # noqa: E402
from skypilot_b200 import engine, synth, workloads  # noqa: E402
from skypilot_b200 import optimizer as opt_lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
w = workloads.WORKLOADS['cfg2']
sky.catalog.load_frames(synth.make_catalogs(**w['catalog']))
sky.check.set_enabled_clouds(sky.check.ALL_CATALOG_CLOUDS)
scs = workloads.cfg5_scenarios(n)
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
wrap(O, '_state_problem', 'state (all DAGs)', True)
wrap(O, '_resolve_ordered_resources', 'resolve ordered', True)
wrap(opt_lib, '_check_specified_clouds', 'check specified clouds')
wrap(engine, 'solve', 'engine.solve (pack + native + unpack)')
wrap(engine.ProblemBuilder, 'pack', '  pack')
wrap(opt_lib._Problem, 'launchable', 'launchable (all tasks)')
for rep in range(3):
    dags = [workloads.build_dag(sc)[0] for sc in scs]
    sky.catalog.clear_request_level_cache()
    cur.clear()
    t0 = time.perf_counter()
    out = sky.optimize_batch(dags, return_exceptions=True)
    total = time.perf_counter() - t0
    print(f'rep {rep}: {n} DAGs in {total:.3f} s')
    for k, v in sorted(cur.items(), key=lambda kv: -kv[1]):
        print(f'    {k:44s} {v:7.3f} s')
