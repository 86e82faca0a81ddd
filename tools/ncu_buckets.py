"""Where the samples of the first kernel of an .ncu-rep fall: totals per stall reason and per block of B SASS
instructions (only blocks above 2 % are printed).  usage: python tools/ncu_buckets.py report.ncu-rep [B]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 100
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
for k in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
          "launch__registers_per_thread"):
    if k in rr[0]:
        i = rr[0].index(k)
        print(k, rr[1][i], rr[2][i])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
start = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hdr = rows[start[0]]
end = start[1] - 1 if len(start) > 1 else len(rows)
data = [r for r in rows[start[0] + 1:end] if len(r) == len(hdr)]
ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ci["# Samples"]]) for r in data)
print("instructions", len(data), "samples", tot)
agg = {h: sum(int(r[ci[h]]) for r in data) for h in stalls}
print("  ".join(f"{h[6:]} {100 * v / tot:.1f}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for b in range(0, len(data), B):
    ch = data[b:b + B]
    s = sum(int(r[ci["# Samples"]]) for r in ch)
    if s < 0.02 * tot:
        continue
    ex = sum(int(r[ci["Instructions Executed"]]) for r in ch)
    top = max(ch, key=lambda r: int(r[ci["# Samples"]]))
    ag = {h: sum(int(r[ci[h]]) for r in ch) for h in stalls}
    t2 = ", ".join(f"{h[6:]} {v}" for h, v in sorted(ag.items(), key=lambda kv: -kv[1])[:3])
    print(f"{b:5d} {100 * s / tot:5.1f}%  exec {ex / 1e6:7.1f}M  top {top[ci['Source']].strip()[:44]:44s} {top[ci['# Samples']]:>6}  [{t2}]")
