"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/launches_rNN_summary.txt"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r'\(anonymous namespace\)::|<unnamed>::|void ', '', r[ki]).split('(')[0]
    tot[name][0] += 1
    tot[name][1] += float(r[vi].replace(',', '')) / 1e6
s = sum(v[1] for v in tot.values())
print(f"launches {sum(v[0] for v in tot.values())}  total {s:.2f} ms (ncu: serialised, cold cache - compare SHARES)")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:64]:64s} n={v[0]:4d} {v[1]:8.3f} ms {100 * v[1] / s:5.1f}%")
