"""Failover loop: OptimizerSession.optimize(blocked) (candidate sets resident,
re-mask + re-solve) vs a full Optimizer.optimize(dag, blocked) per iteration
(SURVEY.md section 8f rank 3), cfg2 catalog and chain. One JSON line."""
import json
import statistics
import sys
import time

sys.path.insert(0, '.')
from skypilot_b200 import workloads as bench  # noqa: E402
import skypilot_b200 as sky  # noqa: E402
from skypilot_b200 import synth  # noqa: E402
from skypilot_b200.utils import registry  # noqa: E402
from tests import scenario_runner as runner  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
w = bench.WORKLOADS[name]
sky.catalog.load_frames(synth.make_catalogs(**w['catalog']))
sky.check.set_enabled_clouds(sky.check.ALL_CATALOG_CLOUDS)
scenario = bench.chain_scenario(w['tasks'])


def wildcard(rec):
    kw = dict(cloud=registry.CLOUD_REGISTRY.from_str(rec['cloud']),
              instance_type=rec['instance_type'], region=rec['region'])
    if rec['zone'] is not None:
        kw['zone'] = rec['zone']
    r = sky.Resources(**kw)
    r._use_spot_specified = False
    return r


def loop(use_session, rounds=12):
    dag, tasks = runner.build_dag(scenario)
    blocked, times, plans = [], [], []
    session = sky.Optimizer.session(dag) if use_session else None
    for i in range(rounds):
        t0 = time.perf_counter()
        if session is not None:
            session.optimize(blocked)
        else:
            sky.Optimizer.optimize(dag, blocked_resources=blocked, quiet=True)
        times.append(time.perf_counter() - t0)
        plan = [runner.res_record(t.best_resources) for t in tasks]
        plans.append(plan)
        blocked.append(wildcard(plan[i % len(plan)]))
    if session is not None:
        session.close()
    return times, plans


loop(False, 3)
full_t, full_p = loop(False)
sess_t, sess_p = loop(True)
print(json.dumps({
    'workload': name, 'rounds': len(full_t), 'same_plans': full_p == sess_p,
    'full_optimize_p50_ms': 1e3 * statistics.median(full_t[1:]),
    'session_first_ms': 1e3 * sess_t[0],
    'session_resolve_p50_ms': 1e3 * statistics.median(sess_t[1:]),
}))
