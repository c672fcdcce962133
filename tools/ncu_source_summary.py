"""Summarise an `ncu --page source --csv` export: stall reasons, hottest SASS instructions, instruction mix."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
H = rows[1]
si, ii, srci = H.index('Warp Stall Sampling (All Samples)'), H.index('Instructions Executed'), H.index('Source')
stalls = [(h, H.index(h)) for h in H if h.startswith('stall')]
data = [r for r in rows[2:] if len(r) > max(si, ii) and (r[si] or '0').isdigit() and (r[ii] or '0').isdigit()]
tot = sum(int(r[si] or 0) for r in data)
toti = sum(int(r[ii] or 0) for r in data)
print('total samples', tot, ' total warp instructions', toti, ' SASS rows', len(data))
agg = collections.Counter()
for r in data:
    for h, i in stalls:
        try:
            agg[h] += int(r[i] or 0)
        except ValueError:
            pass
s = sum(agg.values())
print('stall reasons:', ', '.join('%s %.1f%%' % (h, 100.0 * v / max(1, s)) for h, v in agg.most_common(8)))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for r in sorted(data, key=lambda r: -int(r[si] or 0))[:n]:
    st = sorted([(int(r[i] or 0), h) for h, i in stalls], reverse=True)[:2]
    print(r[si].rjust(7), r[ii].rjust(10), r[srci][:100], st)
mix = collections.Counter()
for r in data:
    o = [x for x in r[srci].split() if not x.startswith('@')]
    if o:
        mix[o[0].split('.')[0]] += int(r[ii] or 0)
print('instruction mix:', mix.most_common(24))
