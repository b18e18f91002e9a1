"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): time and launch count per kernel.
  python scripts/launch_summary.py gpurun_out/launches.csv [top=30] [passes=1]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
passes = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
ki, vi = rows[hdr].index("Kernel Name"), rows[hdr].index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hdr + 2:]:
  if len(r) <= vi:
    continue
  try:
    v = float(r[vi].replace(",", ""))
  except ValueError:
    continue
  a = agg.setdefault(r[ki][:90], [0, 0.0])
  a[0] += 1
  a[1] += v
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
  print(f"{a[1] / 1e6 / passes:9.3f} ms  x{a[0] / passes:6.1f}  {k}")
print(f"{tot / 1e6 / passes:9.3f} ms total, {sum(a[0] for a in agg.values()) / passes:.0f} launches")
