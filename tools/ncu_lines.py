"""Join an `ncu --page source --csv --print-source sass` export with the line table of the profiled cubin (nvdisasm -g) and
aggregate warp-stall samples / executed instructions per CUDA source line.

    cuobjdump -xelf all targetdiff_b200/csrc/build/edge_mlp_v4.o          # -> edge_mlp_v4.sm_100a.cubin
    python tools/ncu_lines.py gpurun_out/<tag>_v4_source.csv edge_mlp_v4.sm_100a.cubin 'edge_mlp_v4_kernelILi128' [top]
"""
import collections
import csv
import re
import subprocess
import sys

csv_path, cubin, kernel_pat = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout.splitlines()
line_of, cur, on = {}, None, False
for ln in dis:
    if ln.startswith('//-----') and '.text.' in ln:
        on = kernel_pat in ln
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/', ln)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(csv_path)))
H = rows[1]
ai, si, ii = H.index('Address'), H.index('Warp Stall Sampling (All Samples)'), H.index('Instructions Executed')
data = [r for r in rows[2:] if len(r) > ii and r[ai].startswith('0x')]
# the export may hold several launches back to back: keep the FIRST block of monotonically increasing addresses
base = int(data[0][ai], 16)
block = []
for r in data:
    a = int(r[ai], 16)
    if block and a <= block[-1][0]:
        break
    block.append((a, r))
agg = collections.defaultdict(lambda: [0, 0, 0])
for a, r in block:
    key = line_of.get(a - base)
    if key is None:
        continue
    agg[key][0] += int(r[si] or 0)
    agg[key][1] += int(r[ii] or 0)
    agg[key][2] += 1
tot = sum(v[0] for v in agg.values())
toti = sum(v[1] for v in agg.values())
src = {}
for f in set(k[0] for k in agg):
    try:
        src[f] = open('targetdiff_b200/csrc/' + f).read().splitlines()
    except OSError:
        src[f] = []
print('samples %d, warp instructions %d, SASS rows %d (first launch in the export)' % (tot, toti, len(block)))
print('%7s %6s %12s %5s  %s' % ('samples', 'share', 'warp instr', 'sass', 'line'))
for (f, n), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = src[f][n - 1].strip()[:110] if 0 < n <= len(src[f]) else ''
    print('%7d %5.1f%% %12d %5d  %s:%d  %s' % (v[0], 100.0 * v[0] / max(1, tot), v[1], v[2], f, n, text))
