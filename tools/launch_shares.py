"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: share / count / mean per kernel."""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    name = re.sub(r'\(.*', '', r[ki])
    v = float(r[vi].replace(',', ''))
    v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1] / tot * 100:6.2f}%  n={v[0]:4d}  avg={v[1] / v[0]:9.1f} us  {k[:100]}")
print(f"total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches")
