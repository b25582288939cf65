"""cProfile of Optimizer.optimize(dag) on a bench workload (GPU box)."""
import cProfile
import pstats
import sys
import time
sys.path.insert(0, '.')
from skypilot_b200 import workloads as bench  # noqa: E402
import skypilot_b200 as sky  # noqa: E402
from skypilot_b200 import synth  # noqa: E402
from tests import scenario_runner as runner  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
w = bench.WORKLOADS[name]
sky.catalog.load_frames(synth.make_catalogs(**w['catalog']))
sky.check.set_enabled_clouds(sky.check.ALL_CATALOG_CLOUDS)
dag, tasks = runner.build_dag(bench.chain_scenario(w['tasks']))
for _ in range(20):
    sky.optimize(dag, quiet=True)
t = time.perf_counter()
for _ in range(200):
    sky.optimize(dag, quiet=True)
print('optimize ms', (time.perf_counter() - t) / 200 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    sky.optimize(dag, quiet=True)
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(25)
