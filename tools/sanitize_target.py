"""Workload for compute-sanitizer: a dozen multi6k scenarios (chains, a
diamond, blocked lists, spot, GCP accelerators) in one kernel mode, plus the
catalog API scan.

    compute-sanitizer --tool memcheck python tools/sanitize_target.py fast
"""
import sys
sys.path.insert(0, '.')
import skypilot_b200 as sky  # noqa: E402
from tests import scenario_runner as runner  # noqa: E402
from tests import scenarios  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else 'auto'
payload = runner.load_golden('multi6k')
records = {r['name']: r for r in payload['records']}
store = runner.activate_catalog(payload['catalog'])
store.set_scan_mode(mode)
names = ['acc_V100', 'acc_A100x8', 'cpu_default', 'mem_4x', 'spot_t4x4',
         'chain3_mixed', 'cfg2_chain8', 'cfg3_diamond', 'diamond_time',
         'chain2_egress_big', 'acc_tpu-v3-8', 'cpu_96p']
suite = {s['name']: s for s in scenarios.basic_scenarios()}
bad = 0
for n in names:
    if n not in suite:
        continue
    got = runner.run_scenario(suite[n])
    diffs = runner.compare(records[n], got)
    bad += bool(diffs)
    print(n, 'ok' if not diffs else diffs[:2])
print('mode', mode, 'mismatches', bad)
sys.exit(1 if bad else 0)
