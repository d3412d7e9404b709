"""Joins an ncu source-page CSV (SASS view) with nvdisasm -g line info: stall samples per source line.
usage: stall_by_line.py src.csv lines.txt kernel_symbol_substr"""
import csv, re, sys, collections
src, lines, sym = sys.argv[1:4]
# address -> (file,line) from nvdisasm
amap = {}; cur = None; infn = False
for l in open(lines):
    if l.startswith('.text.'):
        infn = sym in l
        continue
    if not infn: continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,6})\*/', l)
    if m: amap[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src)))
hdr = rows[1]
ia, isamp, iex = hdr.index('Address'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
base = None
per = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
for r in rows[2:]:
    if len(r) < len(hdr): continue
    a = int(r[ia], 16)
    if base is None: base = a
    key = amap.get(a - base, ('?', 0))
    per[key]['samples'] += int(r[isamp] or 0)
    per[key]['inst'] += int(r[iex] or 0)
    for i in stall_cols:
        v = int(r[i] or 0)
        per[key][hdr[i]] += v; tot[hdr[i]] += v
T = sum(v['samples'] for v in per.values())
print('total samples', T, dict(tot.most_common(8)))
for key, v in sorted(per.items(), key=lambda kv: -kv[1]['samples'])[:int(sys.argv[4]) if len(sys.argv) > 4 else 60]:
    top = ', '.join('%s %.1f%%' % (k[6:], 100.0 * c / T) for k, c in v.most_common(5) if k.startswith('stall_') and c)
    print('%-18s %5d  %5.1f%%  inst %9d  %s' % (key[0], key[1], 100.0 * v['samples'] / T, v['inst'], top))
