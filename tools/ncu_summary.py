"""Summarise an .ncu-rep here (no GPU needed): headline metrics + hottest SASS lines + per-region instruction counts.
usage: python tools/ncu_summary.py report.ncu-rep [kernel-regex] [n_top]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__inst_executed_pipe_uniform.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
for k in want:
    if k in hdr:
        i = hdr.index(k)
        print(f"{k} | " + " | ".join(r[i] for r in rows[1:]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# first kernel only
start = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
if start:
    hdr = rows[start[0]]
    end = start[1] - 1 if len(start) > 1 else len(rows)
    data = [r for r in rows[start[0] + 1:end] if len(r) == len(hdr)]
    ci = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[ci["# Samples"]]) for r in data)
    print(f"\n# hottest SASS lines of the first kernel (total samples {tot})")
    for r in sorted(data, key=lambda r: -int(r[ci["# Samples"]]))[:ntop]:
        s = sorted(((h, int(r[ci[h]])) for h in stalls if int(r[ci[h]]) > 0), key=lambda kv: -kv[1])[:2]
        print(r[ci["Address"]][-5:], r[ci["# Samples"]].rjust(7), r[ci["Instructions Executed"]].rjust(10),
              r[ci["Source"]].strip()[:76].ljust(76), s)
