"""dynamic instruction counts + stall samples aggregated by OUTERMOST source line of the kernel file (inline chains
resolved with nvdisasm -gi).  usage: inst_by_region.py src.csv inl.txt kernel_symbol_substr file.cu min_line path/to/file.cu"""
import csv, re, sys, collections
src, inl, sym, fname, lo = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5])
amap = {}; ctx = []; infn = False
for l in open(inl):
    if l.startswith('.text.'):
        infn = sym in l; continue
    if not infn: continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)(.*)', l)
    if m:
        e = (m.group(1).split('/')[-1], int(m.group(2)))
        if 'inlined at' in l: ctx.append(e)
        else: ctx = [e]
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,6})\*/', l)
    if m:
        outer = [c for c in ctx if c[0] == fname and c[1] >= lo]
        amap[int(m.group(1), 16)] = outer[-1][1] if outer else -1
rows = list(csv.reader(open(src))); hdr = rows[1]
ia, isamp, iex = hdr.index('Address'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
base = None; per = collections.defaultdict(collections.Counter)
for r in rows[2:]:
    if len(r) < len(hdr): continue
    a = int(r[ia], 16)
    if base is None: base = a
    k = amap.get(a - base, -2)
    per[k]['samples'] += int(r[isamp] or 0); per[k]['inst'] += int(r[iex] or 0); per[k]['static'] += 1
    for i in stall_cols: per[k][hdr[i]] += int(r[i] or 0)
T = sum(v['samples'] for v in per.values()); I = sum(v['inst'] for v in per.values())
lines = open(sys.argv[6]).read().splitlines()
print('total samples %d, inst %d' % (T, I))
for k in sorted(per):
    v = per[k]
    if v['samples'] < 0.006 * T and v['inst'] < 0.006 * I: continue
    top = ', '.join('%s %.1f' % (n[6:], 100.0 * c / T) for n, c in v.most_common(6) if n.startswith('stall_') and c > 0.003 * T)
    print('%4d smp %5.1f%% inst %5.1f%% static %5d | %-58s | %s' % (k, 100.0 * v['samples'] / T, 100.0 * v['inst'] / I, v['static'], (lines[k - 1].strip()[:58] if k > 0 else ''), top))
