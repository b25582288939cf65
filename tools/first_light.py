"""GPU bring-up helper: run scenarios of one catalog against the golden
fixture and print every difference (used under compute-sanitizer)."""
import json
import sys
import time
import traceback

sys.path.insert(0, '.')
from tests import scenario_runner as runner  # noqa: E402
from tests import scenarios  # noqa: E402


def main():
    catalog = sys.argv[1] if len(sys.argv) > 1 else 'multi6k'
    limit = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    only = sys.argv[3].split(',') if len(sys.argv) > 3 else None
    payload = runner.load_golden(catalog)
    records = {r['name']: r for r in payload['records']}
    runner.activate_catalog(payload['catalog'])
    n_ok = n_bad = 0
    t0 = time.time()
    for sc in scenarios.SUITES[catalog]()[:limit]:
        if only and sc['name'] not in only:
            continue
        try:
            got = runner.run_scenario(sc)
            unordered = any(
                t.get('resources_kind') == 'set' for t in sc['tasks'])
            diffs = runner.compare(records[sc['name']], got, unordered)
        except Exception:  # pylint: disable=broad-except
            diffs = ['EXCEPTION ' + traceback.format_exc()[-1500:]]
        if diffs:
            n_bad += 1
            print(f'[FAIL] {sc["name"]}')
            for d in diffs[:6]:
                print('    ', d)
        else:
            n_ok += 1
            print(f'[ ok ] {sc["name"]}')
    print(f'{n_ok} ok, {n_bad} failed in {time.time() - t0:.1f}s')


if __name__ == '__main__':
    main()
