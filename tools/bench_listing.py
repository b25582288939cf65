"""list_accelerators: device path vs the CPU restatement on the cfg2 catalog
(SURVEY.md section 8f rank 2). Prints one JSON line."""
import json
import statistics
import sys
import time

sys.path.insert(0, '.')
from skypilot_b200 import workloads as bench  # noqa: E402
import skypilot_b200 as sky  # noqa: E402
from oracle import listing_oracle as lo  # noqa: E402
from skypilot_b200 import synth  # noqa: E402

frames = synth.make_catalogs(**bench.WORKLOADS['cfg2']['catalog'])
store = sky.catalog.load_frames(frames)
sky.check.set_enabled_clouds(sky.check.ALL_CATALOG_CLOUDS)
clouds = [t.name for t in store.clouds]
CASES = [dict(), dict(all_regions=True, name_filter='A100'),
         dict(region_filter='us-', quantity_filter=8), dict(require_price=False)]
out = {'catalog_rows': int(store.n_rows), 'cases': []}
for kw in CASES:
    kw = dict(kw, clouds=clouds)
    for _ in range(3):
        sky.catalog.list_accelerators(**kw)
    ts = []
    for _ in range(20):
        t0 = time.perf_counter()
        got = sky.catalog.list_accelerators(**kw)
        ts.append(time.perf_counter() - t0)
    t0 = time.perf_counter()
    want = lo.list_accelerators(frames, **kw)
    cpu = time.perf_counter() - t0
    same = {k: len(v) for k, v in got.items()} == {k: len(v) for k, v in want.items()}
    out['cases'].append({'kwargs': {k: v for k, v in kw.items() if k != 'clouds'},
                         'gpu_p50_ms': 1e3 * statistics.median(ts),
                         'cpu_oracle_ms': 1e3 * cpu, 'same_shape': same,
                         'entries': sum(len(v) for v in got.values())})
print(json.dumps(out))
