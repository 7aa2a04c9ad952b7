"""Correlate an ncu SASS source page (ncu -i X.ncu-rep --page source --csv) with source lines via nvdisasm -g.
usage: ncu_by_line.py src.csv dis.txt kernel_mangled_substring source_file"""
import csv, re, sys
from collections import defaultdict
src_csv, dis, kname, srcfile = sys.argv[1:5]
# 1) offset -> line (innermost file line of our source)
off2line = {}
cur = None
infunc = False
for l in open(dis):
    if l.startswith(".text."):
        infunc = kname in l
        continue
    if not infunc: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);', l)
    if m and cur:
        off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
base = None
agg = defaultdict(lambda: [0, 0, 0, 0])
tot = [0, 0, 0, 0]
for r in rows[2:]:
    if len(r) < len(hdr) - 5: continue
    a = int(r[0], 16)
    if base is None: base = a
    f, ln = off2line.get(a - base, ('?', 0))
    key = ln if f.endswith(srcfile) else -1
    v = [int(r[ci['# Samples']] or 0), int(r[ci['Instructions Executed']] or 0), int(r[ci['stall_barrier']] or 0), int(r[ci['stall_long_sb']] or 0)]
    for i in range(4): agg[key][i] += v[i]; tot[i] += v[i]
src = open(sys.argv[5] if len(sys.argv) > 5 else srcfile).read().split('\n')
funcs = []
for i, l in enumerate(src, 1):
    m = re.match(r'^(?:RBPE_DEV|__global__|__host__ __device__ inline)\s+.*?(\w+)\(', l)
    if m: funcs.append((i, m.group(1)))
def fn(line):
    name = '?'
    for i, n in funcs:
        if i <= line: name = n
    return name
fa = defaultdict(lambda: [0, 0, 0, 0])
for k, v in agg.items():
    for i in range(4): fa[fn(k) if k > 0 else 'other-file'][i] += v[i]
print("total samples %d, warp-instructions %d" % (tot[0], tot[1]))
print("%-16s %8s %8s %10s %10s" % ("function", "samples%", "inst%", "barrier%", "long_sb%"))
for k, v in sorted(fa.items(), key=lambda kv: -kv[1][0]):
    print("%-16s %8.2f %8.2f %10.2f %10.2f" % (k, 100 * v[0] / tot[0], 100 * v[1] / tot[1], 100 * v[2] / tot[0], 100 * v[3] / tot[0]))
print("--- top lines by samples")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:30]:
    print("%4d %6.2f%% inst %6.2f%% barrier %6.2f%% | %s" % (k, 100 * v[0] / tot[0], 100 * v[1] / tot[1], 100 * v[2] / tot[0], src[k - 1].strip()[:100] if k > 0 else ''))
