"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel name.
usage: python tools/launch_summary.py launches.csv [name-width]"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
width = int(sys.argv[2]) if len(sys.argv) > 2 else 64
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] in ("ns", "nsecond") else v * 1e3 if r[ui] in ("ms", "msecond") else v
    agg.setdefault(r[ki][:width], []).append(v)
total = sum(sum(v) for v in agg.values())
print(f"{'kernel':{width}s} {'n':>5} {'mean_us':>9} {'max_us':>9} {'sum_us':>10} share")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:{width}s} {len(v):5d} {sum(v) / len(v):9.1f} {max(v):9.1f} {sum(v):10.1f} {sum(v) / total:5.3f}")
print(f"total {total:.1f} us over {sum(len(v) for v in agg.values())} launches")
