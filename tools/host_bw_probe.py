"""Concurrent page-locked host<->device copy probe, one rank per GPU (run under torchrun):
bytes/s per GPU alone and with all ranks copying at once, with the rank's host buffers placed by the OS (unbound) and
on the GPU's own NUMA node (aru_bind_host_to_device).  Names the limit of the end-to-end (host buffer) path at 8 GPUs.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
      tools/host_bw_probe.py > profiles/r02_host_bw_probe_8gpu.json
"""
import ctypes
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import __graft_entry__ as entry

MB = 512


def node_of(arr):
    """NUMA node of the first page of a numpy array (move_pages with a null node list only queries)."""
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        page = ctypes.c_void_p(arr.ctypes.data & ~4095)
        status = ctypes.c_int(-1)
        rc = libc.syscall(279, 0, 1, ctypes.byref(page), None, ctypes.byref(status), 0)   # SYS_move_pages (x86-64)
        return status.value if rc == 0 else None
    except Exception:
        return None


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    entry.build()
    from aru_b200.engine import PINNED, bind_host_to_device

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    dev = torch.empty(MB << 20, dtype=torch.uint8, device="cuda")
    dev2 = torch.empty(MB << 20, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def rate(host_a, host_b, what, reps=6):
        ta, tb = torch.from_numpy(host_a), torch.from_numpy(host_b)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(reps):
            if what in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    dev.copy_(ta, non_blocking=True)
            if what in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    tb.copy_(dev2, non_blocking=True)
        s1.synchronize(); s2.synchronize()
        dt = time.perf_counter() - t0
        nbytes = reps * (MB << 20) * (2 if what == "both" else 1)
        return nbytes / dt / 1e9

    result = {"rank": rank, "gpu": local, "modes": {}}
    for mode in ("unbound", "bound"):
        node = -1
        if mode == "bound":
            node = bind_host_to_device(local)
        # fresh page-locked blocks in this mode (not recycled ones)
        a = PINNED.empty(((MB << 20) + (1 + rank) * 4096 + (0 if mode == "unbound" else 1 << 16),), np.uint8)[:MB << 20]
        b = PINNED.empty(((MB << 20) + (1 + rank) * 4096 + (1 << 17) + (0 if mode == "unbound" else 1 << 16),), np.uint8)[:MB << 20]
        a[::4096] = 1; b[::4096] = 1
        m = {"gpu_numa_node": node, "cpus_allowed": len(os.sched_getaffinity(0)), "buffer_node": node_of(a)}
        rate(a, b, "both", 2)
        for what in ("h2d", "d2h", "both"):
            solo = None
            for r in range(world):          # ranks take turns
                barrier()
                if r == rank:
                    solo = rate(a, b, what)
            barrier()
            together = rate(a, b, what)     # all ranks at once
            barrier()
            m[what] = {"alone_GBps": round(solo, 2), "all_ranks_GBps": round(together, 2)}
        result["modes"][mode] = m
        del a, b
    gathered = [None] * world
    if world > 1:
        dist.all_gather_object(gathered, result)
    else:
        gathered = [result]
    if rank == 0:
        def sh(cmd):
            try:
                return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout
            except Exception as ex:  # noqa
                return str(ex)
        summary = {}
        for mode in ("unbound", "bound"):
            for what in ("h2d", "d2h", "both"):
                summary[f"{mode}_{what}_sum_all_ranks_GBps"] = round(sum(g["modes"][mode][what]["all_ranks_GBps"] for g in gathered), 1)
                summary[f"{mode}_{what}_mean_alone_GBps"] = round(sum(g["modes"][mode][what]["alone_GBps"] for g in gathered) / world, 1)
        print(json.dumps({"world": world, "MB_per_copy": MB, "summary": summary, "ranks": gathered,
                          "topo": sh("nvidia-smi topo -m"), "lscpu": sh("lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)'"),
                          "meminfo": sh("grep -E 'MemTotal|MemFree' /proc/meminfo")}, indent=1))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
