#!/bin/bash
for c in multi6k three4k; do
python tests/level2_child.py $c /tmp/l2_$c.json > /dev/null 2>&1
python - <<EOF
import json, sys
sys.path.insert(0, '.')
from tests import scenario_runner as runner, scenarios
got = json.load(open('/tmp/l2_$c.json'))
golden = {r['name']: r for r in runner.load_golden('$c')['records']}
suite = {s['name']: s for s in scenarios.SUITES['$c']()}
bad = 0
for rec in got['records']:
    sc = suite[rec['name']]
    unordered = any(t.get('resources_kind') == 'set' for t in sc['tasks'])
    diffs = runner.compare(golden[rec['name']], rec, unordered_candidates=unordered)
    if diffs:
        bad += 1
        print('$c', rec['name'], str(diffs[:1])[:400])
        tb = rec.get('error', {}).get('traceback', '')
        if tb: print('   ', tb[-500:].replace('\n', '\n    '))
print('$c', len(got['records']), 'records', bad, 'differ', 'calls', got['catalog_calls'])
EOF
done
