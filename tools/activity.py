"""CPU-side estimate of the scan's work: how many (128-row chunk, query) pairs
survive the zone-map test, per cloud, with and without vCPU / memory ranges
in the summary. Profiling aid (no GPU needed)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import networkx as nx
from skypilot_b200 import workloads as bench
from skypilot_b200 import synth, _native, engine
from skypilot_b200.catalog.store import CatalogStore
import skypilot_b200 as sky
from skypilot_b200 import optimizer as opt_lib
from skypilot_b200.optimizer import Optimizer
from tests import scenario_runner as runner

name = sys.argv[1] if len(sys.argv) > 1 else 'cfg4'
w = bench.WORKLOADS[name]
scenario = bench.chain_scenario(w['tasks'])
frames = synth.make_catalogs(**w['catalog'])
store = CatalogStore.from_frames(frames)
sky.catalog.set_store(store, 0)
sky.check.set_enabled_clouds(sky.check.ALL_CATALOG_CLOUDS)
dag, tasks = runner.build_dag(scenario)
Optimizer._add_dummy_source_sink_nodes(dag)
graph = dag.get_graph()
topo = [t for t in nx.topological_sort(graph) if not opt_lib._is_dummy(t)]
problem = Optimizer._state_problem(graph, topo, True, [], True)
packed = problem.builder.pack()
q = packed.queries[:packed.n_queries]
sets = engine.set_table(store).reshape(-1, _native.ACC_SET_WORDS)
cols = store.columns
zm = cols['zone_map']
zr = _native.ZONE_ROWS
nz = len(zm)
fl = cols['flags'].reshape(nz, zr)
valid = (fl & _native.F_VALID) != 0
vc = cols['vcpus'].reshape(nz, zr); mm = cols['mem'].reshape(nz, zr)
with np.errstate(all='ignore'):
    vmin = np.nanmin(np.where(valid, vc, np.nan), axis=1); vmax = np.nanmax(np.where(valid, vc, np.nan), axis=1)
    mmin = np.nanmin(np.where(valid, mm, np.nan), axis=1); mmax = np.nanmax(np.where(valid, mm, np.nan), axis=1)
offs = store.cloud_row_offsets if hasattr(store, 'cloud_row_offsets') else cols['cloud_row_offsets']
tot_old = tot_new = 0
for c, t in enumerate(store.clouds):
    z0, z1 = offs[c] // zr, offs[c + 1] // zr
    qs = q[q['cloud'] == c]
    old = new = 0
    for rec in qs:
        rq = (int(rec['flags_require']) | _native.F_VALID) & 0xFF
        ok = (zm['flags_or'][z0:z1] & rq) == rq
        if rec['qflags'] & _native.Q_ACC:
            s = sets[int(rec['acc_set'])]
            lo = np.bitwise_or.reduce(s[0::2]); hi = np.bitwise_or.reduce(s[1::2])
            ok &= ((zm['sig_lo'][z0:z1] & lo) | (zm['sig_hi'][z0:z1] & hi)) != 0
        old += int(ok.sum())
        ok2 = ok.copy()
        if rec['cpus_op']:
            lo_ = rec['cpus']; hi_ = np.inf if rec['cpus_op'] == _native.OP_GE else rec['cpus']
            ok2 &= (vmax[z0:z1] >= lo_) & (vmin[z0:z1] <= hi_)
        if rec['mem_op'] == _native.OP_RATIO:
            ok2 &= mmax[z0:z1] >= vmin[z0:z1] * rec['mem']
        elif rec['mem_op']:
            lo_ = rec['mem']; hi_ = np.inf if rec['mem_op'] == _native.OP_GE else rec['mem']
            ok2 &= (mmax[z0:z1] >= lo_) & (mmin[z0:z1] <= hi_)
        new += int(ok2.sum())
    print(t.name, 'chunks', z1 - z0, 'queries', len(qs), 'active pairs old', old, 'new', new,
          'per chunk old %.2f new %.2f' % (old / max(z1 - z0, 1), new / max(z1 - z0, 1)))
    tot_old += old; tot_new += new
print('total', tot_old, tot_new)

if len(sys.argv) > 2:
    cname = sys.argv[2]
    c = store.cloud_index[cname]
    z0, z1 = offs[c] // zr, offs[c + 1] // zr
    for i, rec in enumerate(q):
        if rec['cloud'] != c: continue
        rq = (int(rec['flags_require']) | _native.F_VALID) & 0xFF
        ok = (zm['flags_or'][z0:z1] & rq) == rq
        n_flag = int(ok.sum())
        if rec['qflags'] & _native.Q_ACC:
            s = sets[int(rec['acc_set'])]
            lo = np.bitwise_or.reduce(s[0::2]); hi = np.bitwise_or.reduce(s[1::2])
            ok &= ((zm['sig_lo'][z0:z1] & lo) | (zm['sig_hi'][z0:z1] & hi)) != 0
        print(i, 'qflags', rec['qflags'], 'req %#x' % rq, 'group', rec['group'], 'cpus', rec['cpus_op'], rec['cpus'], 'mem', rec['mem_op'], rec['mem'],
              'pcol', rec['price_col'], 'flagpass', n_flag, 'active', int(ok.sum()), 'keys', int(sum(bin(int(x)).count('1') for x in sets[int(rec['acc_set'])])) if rec['qflags'] & 1 else '-')
