#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (second half of the
launches = the timed step when the script ran one warm-up step and one timed step)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mv = h.index("Kernel Name"), h.index("Metric Value")
data = [r for r in rows[hi + 1:] if len(r) > mv]
if len(sys.argv) < 3 or sys.argv[2] != "all":
    data = data[len(data) // 2:]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in data:
    name = r[kn].split("(")[0][:70]
    agg[name][0] += 1
    agg[name][1] += float(r[mv].replace(",", ""))
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1] / 1e6:9.3f} ms {v[0]:5d} launches {100 * v[1] / tot:5.1f}%  {k}")
print(f"total {tot / 1e6:.2f} ms")
