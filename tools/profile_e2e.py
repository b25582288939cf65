"""cProfile of Optimizer.optimize(dag) on a bench workload (GPU box), in the
three regimes bench.py reports: same objects (warm), fresh request objects
with the catalog-level statement cache kept, and everything dropped (cold).

    python tools/profile_e2e.py cfg2|cfg4 [warm|fresh|cold]
"""
import cProfile
import pstats
import sys
import time
sys.path.insert(0, '.')
import skypilot_b200 as sky  # noqa: E402
from skypilot_b200 import synth, workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
mode = sys.argv[2] if len(sys.argv) > 2 else 'cold'
w = workloads.WORKLOADS[name]
sky.catalog.load_frames(synth.make_catalogs(**w['catalog']))
sky.check.set_enabled_clouds(sky.check.ALL_CATALOG_CLOUDS)
scenario = workloads.chain_scenario(w['tasks'])
dag, tasks = workloads.build_dag(scenario)


def one():
    global dag, tasks
    if mode == 'fresh':
        dag, tasks = workloads.build_dag(scenario)
    elif mode == 'cold':
        sky.catalog.clear_request_level_cache()
        dag, tasks = workloads.build_dag(scenario)
    sky.optimize(dag, quiet=True)


for _ in range(20):
    one()
times = []
for _ in range(200):
    t = time.perf_counter()
    one()
    times.append(time.perf_counter() - t)
times.sort()
print(f'{name} {mode}: optimize p50 ms', times[100] * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    one()
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(22)
