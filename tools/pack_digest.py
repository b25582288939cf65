"""Digest of everything the device is given for the reference-pinned suites:
the catalog columns and, per scenario, the packed problem (queries, slots,
tasks, accelerator sets, parents, tariffs, blocked entries, DAG records).

The device result is a pure function of these bytes, so a host-side change is
proven result-neutral without a GPU by running this at two commits and
comparing the JSON files:

    python tools/pack_digest.py /tmp/new.json
    git worktree add /tmp/old <commit> && (cd /tmp/old && python tools/pack_digest.py /tmp/old.json)

(Scenarios whose task holds a Python *set* of requests iterate in address
order and may differ between processes.)
"""
import sys, json, hashlib
sys.path.insert(0, '.')
import networkx as nx
import numpy as np
from skypilot_b200 import optimizer as opt_lib
from tests import scenario_runner as runner, scenarios
out = {}
for catalog in ['multi6k', 'three4k', 'aws50k', 'multi50k', 'gpuclouds', 'ibm5k', 'hyperprime']:
    spec = scenarios.CATALOGS[catalog]
    store = runner.activate_catalog(spec)
    h = hashlib.md5()
    for k in sorted(store.columns):
        h.update(k.encode()); h.update(np.ascontiguousarray(store.columns[k]).tobytes())
    out[f'{catalog}:columns'] = h.hexdigest()
    for sc in scenarios.SUITES[catalog]():
        try:
            dag, tasks = runner.build_dag(sc)
            blocked = runner.blocked_list(sc)
            O = opt_lib.Optimizer
            O._add_dummy_source_sink_nodes(dag)
            g = dag.get_graph()
            topo = [t for t in nx.topological_sort(g) if not opt_lib._is_dummy(t)]
            mc = sc.get('minimize', 'cost') == 'cost'
            p = O._state_problem(g, topo, mc, blocked or [], dag.is_chain()).builder.pack()
            hh = hashlib.md5()
            for arr, n in ((p.queries, p.n_queries), (p.slots, p.n_slots), (p.tasks, p.n_tasks)):
                hh.update(arr[:n].tobytes())
            hh.update(np.asarray(p.acc_sets).tobytes())
            for name in ('parents', 'tariffs', 'blocked', 'dags'):
                a = getattr(p, name, None)
                if a is not None:
                    hh.update(np.asarray(a).tobytes())
            out[f'{catalog}:{sc["name"]}'] = hh.hexdigest()
        except Exception as e:
            out[f'{catalog}:{sc["name"]}'] = 'EXC ' + type(e).__name__
json.dump(out, open(sys.argv[1], 'w'), indent=0, sort_keys=True)
print(len(out))
