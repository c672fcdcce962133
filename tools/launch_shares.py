"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel over ONE steady-state denoising step:
the launches from the last-but-one knn launch (inclusive) to the last knn launch (exclusive)."""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
names = [r[4] for r in rows]
knn = [i for i, n in enumerate(names) if n.startswith('knn')]
lo, hi = (knn[-2], knn[-1]) if len(knn) >= 2 else (0, len(rows))
agg = collections.OrderedDict()
for r in rows[lo:hi]:
    m = re.match(r'(void )?([A-Za-z0-9_]+)(<[^>]*>)?', r[4])
    key = (m.group(2) + (m.group(3) or '')).replace(',', ';').replace('(int)', '')
    t = float(r[14]) / 1e3
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(a[1] for a in agg.values())
print('kernel,launches,total_us,avg_us,share')
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%s,%d,%.1f,%.1f,%.4f' % (k, n, t, t / n, t / tot))
print('TOTAL,%d,%.1f,,1.0' % (sum(a[0] for a in agg.values()), tot))
