"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and one minibatch step."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hi]
kn, mv = h.index('Kernel Name'), h.index('Metric Value')
seq = [(r[kn], float(r[mv].replace(',', ''))) for r in rows[hi + 2:] if len(r) > mv and r[mv]]


def short(n):
    n = n.replace('<unnamed>::', '').replace('void ', '')
    m = re.match(r'([\w:]+(<[^(]*>)?)', n)
    return (m.group(1) if m else n)[:70]


tot, cnt = collections.Counter(), collections.Counter()
for n, t in seq:
    tot[short(n)] += t
    cnt[short(n)] += 1
print(f"{len(seq)} launches, {sum(tot.values()) / 1e3:.1f} us of device time (cold-cache, serialised)")
for k, v in tot.most_common(30):
    print(f"{v / 1e3:10.1f} us {cnt[k]:5d} x {v / cnt[k] / 1e3:8.2f}  {k}")
if len(sys.argv) > 2:  # print the launches between two consecutive launches of the named kernel
    idx = [i for i, (n, _) in enumerate(seq) if sys.argv[2] in n]
    a, b = idx[-3], idx[-2]
    print(f"--- one period of {sys.argv[2]} ({sum(t for _, t in seq[a:b]) / 1e3:.1f} us)")
    for n, t in seq[a:b]:
        print(f"  {t / 1e3:8.2f} us {short(n)}")
