"""Pivot an ncu multi-metric launch list (csv) into one line per launch.
usage: python tools/pass_metrics.py pass_metrics.csv [min_us]"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 8]
min_us = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
hdr = rows[0]
ci = {h: i for i, h in enumerate(hdr)}
launches = OrderedDict()
for r in rows[1:]:
    key = r[ci["ID"]]
    d = launches.setdefault(key, {"name": r[ci["Kernel Name"]], "grid": r[ci["Grid Size"]]})
    v = float(r[ci["Metric Value"]].replace(",", ""))
    unit = r[ci["Metric Unit"]]
    name = r[ci["Metric Name"]]
    if name == "gpu__time_duration.sum":
        v = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
    if name.startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    d[name] = v
print(f"{'id':>4} {'kernel':46} {'grid':>14} {'us':>8} {'dramMB':>8} {'GB/s':>7} {'issue%':>6} {'warps%':>6} {'tensor%':>7} {'Minst':>7}")
for k, d in launches.items():
    us = d.get("gpu__time_duration.sum", 0.0)
    if us < min_us:
        continue
    mb = (d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)) / 1e6
    print(f"{k:>4} {d['name'][:46]:46} {d['grid']:>14} {us:8.1f} {mb:8.1f} {mb / us * 1e3 if us else 0:7.0f} "
          f"{d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0):6.1f} "
          f"{d.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0):6.1f} "
          f"{d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):7.1f} "
          f"{d.get('smsp__inst_executed.sum', 0) / 1e6:7.1f}")
