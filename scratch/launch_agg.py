"""Aggregate an ncu launch list (csv) of scratch/train_time.py: the second training step, by kernel."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]; kn = h.index('Kernel Name'); mv = h.index('Metric Value'); mu = h.index('Metric Unit'); gs = h.index('Grid Size')
L = []
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    try: v = float(r[mv].replace(',', ''))
    except ValueError: continue
    if r[mu] == 'ns': v /= 1e3
    elif r[mu] == 'ms': v *= 1e3
    L.append((re.sub(r'\(.*', '', r[kn])[:60], r[gs], v))
idx = [i for i, (n, g, v) in enumerate(L) if 'multi_tensor' in n or 'adam_step' in n]
# two training steps were captured (1 warm-up + 1 timed): the second one starts after the first half of the Adam launches
S = L[idx[len(idx) // 2 - 1] + 1:idx[-1] + 1]
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0
for n, g, v in S: agg[n][0] += 1; agg[n][1] += v; tot += v
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:24]:
    print('%-62s %5d %10.1f us %5.1f%%' % (k, n, t, 100 * t / tot))
print('total %.1f us over %d launches' % (tot, len(S)))
if len(sys.argv) > 2:
    for i, (n, g, v) in enumerate(S):
        if v > float(sys.argv[2]): print(i, n, g, '%.0f' % v)
