"""A short, fixed launch sequence for ncu: the cfg4 (or cfg2) chain in one
kernel mode, `iters` device-resident steps with the L2 flush in between.

    python tools/ncu_target.py cfg4 fast-split 6          # scan2 / place / solve launches
    python tools/ncu_target.py cfg4 auto 6                # step_kernel
    python tools/ncu_target.py stress fast-split-noprune 4   # 8x rows, nothing pruned
"""
import sys
sys.path.insert(0, '.')
import networkx as nx  # noqa: E402
import skypilot_b200 as sky  # noqa: E402
from skypilot_b200 import engine, synth, workloads  # noqa: E402
from skypilot_b200 import optimizer as opt_lib  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'cfg4'
mode = sys.argv[2] if len(sys.argv) > 2 else 'auto'
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 6
base = 'cfg4' if name == 'stress' else name
w = workloads.WORKLOADS[base]
store = sky.catalog.load_frames(synth.make_catalogs(**w['catalog']))
sky.check.set_enabled_clouds(sky.check.ALL_CATALOG_CLOUDS)
if name == 'stress':
    store = store.replicated(8)
    sky.catalog.set_store(store)
store.handle(0)
store.set_scan_mode(mode)
dag, tasks = workloads.build_dag(workloads.chain_scenario(w['tasks']))
O = opt_lib.Optimizer
O._add_dummy_source_sink_nodes(dag)
graph = dag.get_graph()
topo = [t for t in nx.topological_sort(graph) if not opt_lib._is_dummy(t)]
problem = O._state_problem(graph, topo, True, [], True)
O._remove_dummy_source_sink_nodes(dag)
sol, iter_ms, scan_ms = engine.solve_timed(problem.builder, iters, True, 0)
print(name, mode, 'rows', store.n_real_rows, 'step ms', [round(float(x), 4) for x in iter_ms],
      'scan ms', [round(float(x), 4) for x in scan_ms], 'form', int(sol.stats.scan_form))
