"""Summarise `ncu --page source --csv` output per source line / per region of hrp_env.cu."""
import collections
import csv
import sys

path = sys.argv[1]
srcfile = sys.argv[2] if len(sys.argv) > 2 else "highway-rope-ppo_b200/csrc/hrp_env.cu"
rows = list(csv.reader(open(path)))
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == 'Line No']
h = rows[hdr_idx[0]]
ie, ss = h.index('Instructions Executed'), h.index('# Samples')
names = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
stalls = {n: h.index(n) for n in names}
end = hdr_idx[1] if len(hdr_idx) > 1 else len(rows)
per = collections.defaultdict(lambda: [0, 0, collections.Counter()])
src = open(srcfile).read().split('\n')
tot_i = tot_s = 0
allst = collections.Counter()
for r in rows[hdr_idx[0] + 1:end]:
    if len(r) <= ie:
        continue
    try:
        ln = int(r[0]); n = int(r[ie] or 0); s = int(r[ss] or 0)
    except ValueError:
        continue
    per[ln][0] += n; per[ln][1] += s
    for k, i in stalls.items():
        try:
            v = int(r[i] or 0)
        except ValueError:
            v = 0
        per[ln][2][k] += v; allst[k] += v
    tot_i += n; tot_s += s
print("total warp-inst", tot_i, "samples", tot_s)
print("stall mix:", ', '.join(f"{k[6:]} {100*v/max(1,sum(allst.values())):.1f}%" for k, v in allst.most_common(10)))
import re
marks = []
for i, line in enumerate(src, 1):
    m = re.match(r'^(?:__device__|__global__|static|template).*?(\w+)\s*\(', line)
    if m and not line.startswith(' '):
        marks.append((i, m.group(1)))
def region(ln):
    name = 'pre'
    for l, n in marks:
        if ln >= l:
            name = n
    return name
reg = collections.defaultdict(lambda: [0, 0])
for ln, (n, s, st) in per.items():
    reg[region(ln)][0] += n; reg[region(ln)][1] += s
for k, (n, s) in sorted(reg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:24s} inst {100*n/tot_i:5.1f}%  samples {100*s/tot_s:5.1f}%")
print()
for ln, (n, s, st) in sorted(per.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 40]:
    top = ', '.join(f"{k[6:]}:{v}" for k, v in st.most_common(3))
    print(f"{ln:4d} inst {100*n/tot_i:4.1f}% smp {100*s/tot_s:4.1f}% [{top}] {src[ln-1].strip()[:100]}")
