"""Summarise a SKYOPT_TIMELINE dump (per-block ns timestamps of the scan)."""
import sys
import numpy as np
raw = np.fromfile(sys.argv[1], dtype=np.uint64)
trace = raw[-128:]
a = raw[:-128].reshape(-1, 8)
t0 = a[:, 0].min()
start, staged, scored, end = [(a[:, i] - t0).astype(np.int64) for i in range(4)]
sm = a[:, 7].astype(int)
print('blocks', len(a), 'span_us', (end.max()) / 1e3)
print('block life us: mean %.2f p50 %.2f p90 %.2f max %.2f' % (
    (end - start).mean() / 1e3, np.median(end - start) / 1e3,
    np.percentile(end - start, 90) / 1e3, (end - start).max() / 1e3))
print('stage us mean %.2f, score us mean %.2f p90 %.2f max %.2f, finish us mean %.2f' % (
    (staged - start).mean() / 1e3, (scored - staged).mean() / 1e3,
    np.percentile(scored - staged, 90) / 1e3, (scored - staged).max() / 1e3,
    (end - scored).mean() / 1e3))
print('first start us: p50 %.2f p90 %.2f max %.2f' % (
    np.median(start) / 1e3, np.percentile(start, 90) / 1e3, start.max() / 1e3))
for q in (0, 25, 50, 75, 90, 99, 100):
    print(' start pct', q, '%.2f us' % (np.percentile(start, q) / 1e3), ' end pct %.2f us' % (np.percentile(end, q) / 1e3))
per_sm = {}
for s_, b, e in zip(sm, start, end):
    per_sm.setdefault(s_, []).append((b, e))
busy = [sum(e - b for b, e in v) / 1e3 for v in per_sm.values()]
print('SMs', len(per_sm), 'blocks/SM min %d max %d' % (min(len(v) for v in per_sm.values()), max(len(v) for v in per_sm.values())), 'sum block-time per SM us: mean %.1f max %.1f' % (np.mean(busy), np.max(busy)))

if trace[1]:
    n = 0
    while n < 30 and trace[4 * n + 1]:
        n += 1
    end = int(trace[126])
    print('trace: %d iterations (cycles), active at exit popc %d' % (n, int(trace[127])))
    for i in range(n):
        q, c0, c1, c2 = [int(v) for v in trace[4 * i:4 * i + 4]]
        nxt = int(trace[4 * (i + 1) + 1]) if i + 1 < n else end
        print('  query %3d  stage1 %5d  stage2 %5d  reduce %5d  total %5d' % (q, c1 - c0, c2 - c1, nxt - c2, nxt - c0))
