"""Per-group summary of a SKYOPT_TIMELINE dump of the queue-form scan kernel."""
import sys
import numpy as np
raw = np.fromfile(sys.argv[1], dtype=np.uint64)
a = raw[:-128].reshape(-1, 8)
t0 = a[:, 0].min()
start, staged, scored, end = [(a[:, i] - t0).astype(np.int64) / 1e3 for i in range(4)]
tested = (a[:, 6] - t0).astype(np.int64) / 1e3
items = (a[:, 4] & 0xFFFFFFFF).astype(int)
pairs = (a[:, 4] >> 32).astype(int)
grp = (a[:, 5] & 0xFFFFFFFF).astype(int)
nq = (a[:, 5] >> 32).astype(int)
print('blocks', len(a), 'span', end.max(), 'items', items.sum(), 'pairs', pairs.sum())
print('stage %.2f test %.2f pop-loop(w0) %.2f wait+finish %.2f' % (
    (staged - start).mean(), (tested - staged).mean(), (scored - tested).mean(),
    (end - scored).mean()))
print('group  nq blocks items pairs  life_mean life_max  us/pair(block life*8/pairs)')
for g in sorted(set(grp)):
    m = grp == g
    life = end[m] - start[m]
    print('%5d %3d %6d %5d %5d   %7.2f %7.2f   %6.2f' % (
        g, nq[m][0], m.sum(), items[m].sum(), pairs[m].sum(), life.mean(), life.max(),
        (life.sum() * 8 / max(1, pairs[m].sum()))))
# the slowest blocks
order = np.argsort(-(end - start))[:12]
for b in order:
    print('block %4d grp %5d items %3d pairs %4d life %.2f (stage %.2f test %.2f loop %.2f tail %.2f)' % (
        b, grp[b], items[b], pairs[b], end[b] - start[b], staged[b] - start[b],
        tested[b] - staged[b], scored[b] - tested[b], end[b] - scored[b]))
