"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, average duration and share per kernel.
usage: ncu_launch_summary.py <launches.csv> > summary.txt"""
import csv, re, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, agg = 0.0, {}
for r in rows[1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").split("::")[-1]
    us = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
    tot += us
print("ncu --metrics gpu__time_duration.sum --clock-control none, bench.py --steps 6 --warmup 3 --no-graph, the 6 timed steps "
      "(cold caches, serialised: shares, not absolutes)")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-28s launches %2d  avg %8.2f us  share of the captured launches %.3f" % (name, n, us / n, us / tot))
