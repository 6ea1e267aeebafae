"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in data:
    if len(r) <= vi:
        continue
    name = r[ki].split('(')[0]
    name = name.replace('<unnamed>::', '').replace('void ', '')[-70:]
    v = float(r[vi].replace(',', ''))
    v = v / 1e3 if r[ui] == 'ns' else v * 1e3 if r[ui] == 'ms' else v
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{t:10.1f} us {n:5d} x {t/n:8.2f} us  {100*t/tot:5.1f}%  {k}")
