#!/usr/bin/env python
"""bench.py - pages/s of the ARU-Net separator forward pass on B200 (BASELINE.json metric).

Workload (BASELINE.json configs[2], SURVEY.md section 8d "Config 3"): the separator net (synthetic frozen
GraphDef, seed 0, S=5 / A=3, 1 043 839 parameters) on a batch of 64 synthetic 1856x1344 grayscale pages per
GPU per step.  Pages are independent, so N GPUs run N page shards with no collective (weak scaling).

  value      device-resident pages/s: inputs already in HBM, CUDA-event timed on the launch stream
  e2e        the same through the reference-facing call (Engine.forward == get_net_output for a batch):
             pinned host float32 pages in, float32 probability maps out, copies inside the timed region
  roofline   the dominant kernel class (tcgen05 conv) against the measured peaks (MEASURED_PEAKS.json)
  cpu_baseline / --impl reference
             the CPU oracle (PyTorch/oneDNN fp32 interpreter of the same GraphDef - the stand-in for the
             reference's TF1 CPU session, which cannot be installed here) on a bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, PAGES_PER_STEP = 1856, 1344, 64
NET = "separator"
METRIC = "pages/sec ARU-Net separator fwd"
UNIT = "pages/s"
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {k: float(p[k]) for k in FALLBACK_PEAKS}, "measured"
    except Exception:
        return dict(FALLBACK_PEAKS), "fallback"


def make_pages(n, h, w):
    """n synthetic pages [n,h,w] float32 in [0,1]: 4 generated layouts, the rest shifted copies."""
    from aru_b200.synth import synth_page
    base = [synth_page(h, w, seed=s).astype(np.float32) / np.float32(255.0) for s in range(min(n, 4))]
    pages = np.empty((n, h, w), np.float32)
    for i in range(n):
        pages[i] = np.roll(base[i % len(base)], shift=(7 * (i // len(base)), 13 * (i // len(base))), axis=(0, 1))
    return pages


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU oracle on a bounded sample
# ---------------------------------------------------------------------------------------------------
def cpu_pages_per_s(steps, warmup, pages_per_sample=1):
    import torch
    from aru_b200.synth import synth_pb
    from oracle.aru_oracle import Oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    orc = Oracle(synth_pb(NET))
    x = make_pages(pages_per_sample, H, W)
    for _ in range(warmup):
        orc.run(x)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.run(x)
    dt = time.perf_counter() - t0
    return steps * pages_per_sample / dt, dt / steps, cores, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the driver's --steps / --warmup are honoured (one page of the workload per step: ~0.7 s on 16-32 host threads);
    # only a run that would not end within a few minutes is cut, and then it says so
    steps, warmup = max(1, min(args.steps, 200)), max(0, min(args.warmup, 50))
    pps, s_per_step, cores, threads = cpu_pages_per_s(steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": pps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"separator net (synthetic .pb, S=5 A=3), {PAGES_PER_STEP} pages of {H}x{W} per GPU per step",
                   "sample": f"1 page of {H}x{W} per step" + ("" if (steps, warmup) == (args.steps, args.warmup) else
                                                                f" (asked for {args.steps}/{args.warmup} steps/warm-up, capped)")},
        "cpu_baseline": {"value": pps, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{steps} x 1 page of {H}x{W}, PyTorch-CPU fp32 oracle of the same GraphDef "
                                   f"(TF1 is not installable here), {threads} threads on {cores} host cores"},
        "e2e": {"value": pps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# roofline of the kernel classes of one plan
# ---------------------------------------------------------------------------------------------------
def kernel_of_chain_is_block(op, prog, prof):
    """True when the launch that performs `op` is a block_mma launch (its conv1 pre-activation stays in shared memory)."""
    ops = list(prog.ops)
    i = ops.index(op)
    buf = op.out.buf
    for j in range(i + 1, len(ops)):
        if ops[j].inp.buf == buf:
            k = prof[j][1]
            if k == "pair_fused":
                buf = ops[j].out.buf
                continue
            return k.startswith("block_mma")
    return False


def class_rooflines(eng, n, h, w, peaks, iters=3):
    """Per kernel class: algorithmic FLOPs / bytes (every layer reads its input and writes its
    output once at the storage dtype) and the CUDA-event time of its launches on the engine stream."""
    from aru_b200 import program as P
    prog = eng.program
    eng.plan(n, h, w)
    prof = eng.profile_ops(iters)
    dims = {}

    def hw(buf):
        if buf not in dims:
            import ctypes
            hh, ww, cc = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
            eng.lib.aru_buffer_dims(eng.handle, buf, ctypes.byref(hh), ctypes.byref(ww), ctypes.byref(cc))
            dims[buf] = (hh.value, ww.value)
        return dims[buf]

    classes, shapes = {}, {}
    writer = {}      # buffer -> class entries of the launch that wrote it (a fused pool's output is its producer's work)
    pending = {}     # buffer -> (flops, bytes) of a convolution performed by its consumer's launch (conv_band2.cu)
    for op, (_, kernel, ms) in zip(prog.ops, prof):
        flops = bytes_ = 0.0
        if kernel == "pool_fused":
            oh, ow = hw(op.out.buf)
            for ent in writer.get(op.inp.buf, ()):
                ent["bytes"] += n * oh * ow * op.inp.ch * 2
            continue
        if kernel == "pair_fused":      # performed by a later op's launch (conv_band2.cu, block_mma.cu): carry its work
            oh, ow = hw(op.out.buf)
            ih, iw = hw(op.inp.buf)
            f0, b0 = pending.pop(op.inp.buf, (0.0, None))
            if b0 is None:              # first op of the chain: its input is what the launch reads
                b0 = n * ih * iw * op.inp.ch * (4 if op.inp.ch == 1 else 2)
            pending[op.out.buf] = (f0 + 2.0 * n * oh * ow * op.ksize * op.ksize * op.inp.ch * op.out.ch,
                                   b0 + (n * oh * ow * op.out.ch * 2 if op.out_pre.buf >= 0 and not kernel_of_chain_is_block(op, prog, prof) else 0))
            continue
        if kernel == "head_fused":      # the attention combine, performed by the classifier's launch (combine_head.cu)
            b = 0.0
            for a, d in zip(op.att, op.det):
                ah, aw = hw(a.buf)
                dh, dw = hw(d.buf)
                b += n * (ah * aw * 4 + dh * dw * d.ch * 2)
            pending[op.out.buf] = (0.0, b)
            continue
        if op.kind in (P.OP_CONV, P.OP_DECONV):
            oh, ow = hw(op.out.buf)
            ih, iw = hw(op.inp.buf)
            px_mac = ih * iw if op.kind == P.OP_DECONV else oh * ow
            flops = 2.0 * n * px_mac * op.ksize * op.ksize * op.inp.ch * op.out.ch
            in_b = 4 if op.inp.ch == 1 else 2
            out_b = 4 if kernel in ("conv_small", "conv_tc_head", "conv_band_head") else 2
            bytes_ = n * (ih * iw * op.inp.ch * in_b + oh * ow * op.out.ch * out_b)
            if kernel == "conv_stem_pool":     # only the pooled tensor is stored (added by the pool_fused op that follows)
                bytes_ = n * ih * iw * 4
            if kernel == "conv_stem_pre":      # only the pre-activation is stored (the block that follows starts from it)
                bytes_ = n * ih * iw * 4
            if op.res.buf >= 0:
                bytes_ += n * oh * ow * op.out.ch * 2
            if op.out_pre.buf >= 0:
                bytes_ += n * oh * ow * op.out.ch * 2
            if kernel.startswith(("conv_band2", "block_mma", "combine_head")) and op.inp.buf in pending:
                f0, b0 = pending.pop(op.inp.buf)
                flops += f0
                bytes_ += b0 - n * ih * iw * op.inp.ch * in_b     # the intermediates never leave the SM
                if kernel.startswith("block_mma") and op.res.buf >= 0:
                    bytes_ -= n * oh * ow * op.out.ch * 2         # the residual is the block's own x0 (computed or already read)
        elif op.kind in (P.OP_MAXPOOL, P.OP_AVGPOOL):
            oh, ow = hw(op.out.buf)
            ih, iw = hw(op.inp.buf)
            b = 4 if op.inp.ch == 1 else 2
            bytes_ = n * (ih * iw + oh * ow) * op.inp.ch * b
        elif op.kind == P.OP_COMBINE:
            oh, ow = hw(op.out.buf)
            bytes_ = n * oh * ow * op.out.ch * 2
            for a, d in zip(op.att, op.det):
                ah, aw = hw(a.buf)
                dh, dw = hw(d.buf)
                bytes_ += n * (ah * aw * 4 + dh * dw * d.ch * 2)
        c = classes.setdefault(kernel, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
        c["ms"] += ms; c["flops"] += flops; c["bytes"] += bytes_; c["launches"] += 1
        writer[op.out.buf] = [c]
        if op.kind in (P.OP_CONV, P.OP_DECONV):
            key = (f"{kernel} {op.ksize}x{op.ksize} {op.inp.ch}->{op.out.ch} {oh}x{ow} n{n}"
                   + ("+res" if op.res.buf >= 0 else "") + ("+pre" if op.out_pre.buf >= 0 else ""))
            sgrp = shapes.setdefault(key, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
            sgrp["ms"] += ms; sgrp["flops"] += flops; sgrp["bytes"] += bytes_; sgrp["launches"] += 1
            writer[op.out.buf].append(sgrp)
    total_ms = sum(c["ms"] for c in classes.values())
    out = []
    for k, c in sorted(classes.items(), key=lambda kv: -kv[1]["ms"]):
        t = c["ms"] * 1e-3
        t_hbm = c["bytes"] / (peaks["hbm_gbs"] * 1e9)
        t_tc = c["flops"] / (peaks["bf16_tflops"] * 1e12)
        bound = "tensor" if t_tc > t_hbm else "hbm"
        achieved = (c["flops"] / t / 1e12) if bound == "tensor" else (c["bytes"] / t / 1e9)
        peak = peaks["bf16_tflops"] if bound == "tensor" else peaks["hbm_gbs"]
        out.append({"kernel": k, "launches": c["launches"], "ms": round(c["ms"], 4), "share": round(c["ms"] / total_ms, 4),
                    "bound": bound, "achieved": round(achieved, 2), "peak": peak, "unit": "TFLOP/s" if bound == "tensor" else "GB/s",
                    "frac": round(achieved / peak, 4), "tflops": round(c["flops"] / t / 1e12, 2),
                    "gbs": round(c["bytes"] / t / 1e9, 1)})
    # the dominant kernel = the launch shape with the largest share of the pass (same kernel, same tensor sizes)
    key, g = max(shapes.items(), key=lambda kv: kv[1]["ms"])
    t = g["ms"] * 1e-3 / g["launches"]
    b, f = g["bytes"] / g["launches"], g["flops"] / g["launches"]
    bound = "tensor" if f / (peaks["bf16_tflops"] * 1e12) > b / (peaks["hbm_gbs"] * 1e9) else "hbm"
    achieved = f / t / 1e12 if bound == "tensor" else b / t / 1e9
    peak = peaks["bf16_tflops"] if bound == "tensor" else peaks["hbm_gbs"]
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            traffic = json.load(fh).get(key, {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    extra = {}
    if key.startswith(("block_mma", "combine_head")):
        # these launches keep their intermediates in shared memory and run on the warp-level tensor path (mma.sync); its
        # measured dense peak on this part is 557 TFLOP/s (tools/hmma_rate_bench.cu, profiles/r02l_hmma_rate.txt), a
        # third of the tcgen05 / cuBLAS figure in MEASURED_PEAKS.json that `peak` must quote
        extra = {"pipe": "mma.sync m16n8k16 (operands in registers; DESIGN.md 4.6)", "mma_sync_peak_tflops": 557.0,
                 "frac_of_mma_sync_peak": round(f / t / 557.0e12, 4)}
    dominant = {"kernel": key, "bound": bound, **extra, "achieved": round(achieved, 2), "peak": peak,
                "unit": "TFLOP/s" if bound == "tensor" else "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                "launches_per_pass": g["launches"], "us_per_launch": round(t * 1e6, 2),
                "algorithmic_bytes_per_launch": int(b), "flops_per_launch": int(f),
                "share_of_step": round(g["ms"] / total_ms, 4)}
    return out, total_ms, dominant


# ---------------------------------------------------------------------------------------------------
# main arm
# ---------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import __graft_entry__ as entry
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the B200 engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    entry.build()
    from aru_b200.engine import Engine, pinned_empty
    from aru_b200.synth import synth_pb
    from aru_b200.graphdef import aru_conv_macs

    peaks, peaks_src = load_peaks()
    n = args.pages
    from aru_b200.engine import bind_host_to_device
    # each rank's page-locked buffers must be local to its GPU's NUMA node (Engine() binds too; done first here because
    # nothing may be page-locked before it)
    numa_node = bind_host_to_device(local) if os.environ.get("ARU_NUMA_BIND", "1") != "0" else -1
    eng = Engine(synth_pb(NET), device=local)
    if args.micro_batch:
        from aru_b200.engine import OPT_MICRO_BATCH
        eng.set_option(OPT_MICRO_BATCH, args.micro_batch)
    C = eng.n_class
    pages = make_pages(n, H, W)
    x_host = pinned_empty((n, H, W), np.float32)
    x_host[...] = pages
    x_dev = torch.from_numpy(pages).cuda()
    y_dev = torch.empty((n, H, W, C), dtype=torch.float32, device="cuda")
    stream = torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput (value) ----
    def step_dev():
        eng.forward_device(x_dev.data_ptr(), n, H, W, out_ptr=y_dev.data_ptr(), stream=stream.cuda_stream)

    for _ in range(args.warmup):
        step_dev()
    stream.synchronize()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_dev()
    e1.record(stream)
    stream.synchronize()
    dev_s = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    barrier()
    eng.sync()

    # ---- end to end through the public call (e2e) ----
    def pipelined(call, *a, **kw):
        """args.steps calls through the public API, two in flight (Engine.submit / wait): every step's inputs are copied
        from pinned host memory and every step's result is read on the host inside the timed region; the next step's
        copy-in overlaps the previous step's tail.  Returns (seconds, last result, checksum)."""
        warm = [eng.submit(call, *a, **kw) for _ in range(max(2, min(args.warmup, 3)))]   # also fills the pinned-block pool
        for t, _ in warm:
            eng.wait(t)
        del warm
        barrier()
        acc, pending = 0.0, None
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ticket, res = eng.submit(call, *a, **kw)
            if pending is not None:
                eng.wait(pending[0])
                acc += touch(pending[1])
            pending = (ticket, res)
        eng.wait(pending[0])
        acc += touch(pending[1])
        last = pending[1]
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        barrier()
        return dt, last, acc

    def touch(res):
        arr = next(iter(res.values())) if isinstance(res, dict) else (res[0] if isinstance(res, tuple) else res)
        return float(arr.reshape(-1)[0]) + float(arr.reshape(-1)[-1])

    # host float32 in -> host float32 [n,H,W,C] out, copies included
    e2e_s, y_host, checksum = pipelined(eng.forward, x_host)

    # ---- the contract-minimal call: uint8 pages in (as cv2.imread / scale_image give them), channel 0 of
    # np.array(net_output * 255, dtype=np.uint8) out - all any caller of get_net_output keeps of the float map
    # (separator_net_post_processor.py:33,147, heading_net_post_processor.py:209,287): 1 B/px up, 1 B/px down ----
    p_host = pinned_empty((n, H, W), np.uint8)
    p_host[...] = np.rint(pages * 255.0).astype(np.uint8)
    u8_s, r_u8, c2 = pipelined(eng.separator_pages, p_host, want_u8=True, want_separators=False, u8_channels=1)
    u8_host = r_u8["u8"]
    checksum += c2

    # ---- end to end one level up: uint8 pages in, the two separator masks out (SURVEY.md 8 f1+f2) ----
    # = SeparatorNetPostProcessor.run up to the polygon step; only 1 B/px goes up and 2 B/px come down
    pages_s, r_pages, c3 = pipelined(eng.separator_pages, p_host)
    checksum += c3

    # ---- what the unmodified caller does: one pageable float64 page per get_net_output call (helper.py:31,56-72) ----
    n_single = min(n, 8)
    singles = [pages[i].astype(np.float64) for i in range(n_single)]
    eng.forward(singles[0]); eng.forward(singles[0])
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for pg in singles:
            checksum += float(eng.forward(pg)[0, 0, 0, 0])
    torch.cuda.synchronize()
    single_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-class roofline (rank 0) ----
    launches = eng.launches_per_forward
    mb = None
    line = None
    if rank == 0:
        import ctypes
        mb_pages = min(n, args.micro_batch) if args.micro_batch else min(n, max(1, min((80 << 20) // (H * W), 32)))
        classes, prof_ms, dom = class_rooflines(eng, mb_pages, H, W, peaks)
        n_mb = -(-n // mb_pages)
        gflop_page = 2.0 * aru_conv_macs(H, W) / 1e9
        total_pages = world * n * args.steps
        value = total_pages / dev_s
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp16 operands, fp32 accumulate", "data": "synthetic",
            "config": {"workload": f"separator net (synthetic .pb, S=5 A=3), {n} pages of {H}x{W} per GPU per step "
                                   f"(BASELINE configs[2])",
                       "pages_per_step_per_gpu": n, "micro_batch": mb_pages, "gflop_per_page": round(gflop_page, 1),
                       "l2": "inputs larger than L2 (per-layer tensors are 0.2-2.6 GB per micro-batch); no flush needed",
                       "sharding": "page-wise, one process per GPU, no collective"},
            "e2e": {"value": total_pages / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(x_host.nbytes),
                    "d2h_bytes_per_step": int(y_host.nbytes), "ms_per_step": e2e_s / args.steps * 1e3,
                    "call": "Engine.forward == get_net_output for a batch: float32 pages in, float32 probability maps out; two steps in "
                            "flight (Engine.submit / wait), every step's copies and host read inside the timed region"},
            "e2e_uint8": {"value": total_pages / u8_s, "unit": UNIT, "h2d_bytes_per_step": int(p_host.nbytes),
                          "d2h_bytes_per_step": int(u8_host.nbytes), "ms_per_step": u8_s / args.steps * 1e3,
                          "call": "Engine.separator_pages(want_u8=True, u8_channels=1): uint8 pages in; colour step, net and "
                                  "np.array(p * 255, uint8)[..., 0] on the device; the uint8 channel-0 map out - the part of "
                                  "get_net_output's result its callers keep (sep:33,147, head:209,287)"},
            "e2e_masks": {"value": total_pages / pages_s, "unit": UNIT, "h2d_bytes_per_step": int(p_host.nbytes),
                          "d2h_bytes_per_step": int(r_pages["horizontal"].nbytes + r_pages["vertical"].nbytes),
                          "ms_per_step": pages_s / args.steps * 1e3,
                          "call": "Engine.separator_pages == SeparatorNetPostProcessor.run up to the polygon step: uint8 "
                                  "pages in; colour step, net, uint8, threshold, component filter and openings on the "
                                  "device; horizontal + vertical uint8 masks out"},
            "e2e_pageable": {"value": world * n_single * args.steps / single_s, "unit": UNIT,
                             "ms_per_page": single_s / (n_single * args.steps) * 1e3,
                             "call": "one pageable float64 [H,W] page per Engine.forward call (what get_net_output receives "
                                     "from load_and_scale_image, helper.py:31): conversion, page-locked staging and one "
                                     "single-page pass per call"},
            "gpu_launches": int(launches * n_mb * args.steps),
            "achieved_tflops": round(value / world * gflop_page / 1e3, 2),
            "roofline": dict(dom, peaks=peaks_src,
                             note="dominant launch shape of one micro-batch pass; CUDA events on the engine stream, mean over "
                                  "its launches; traffic = ncu dram read+write bytes of the same launch (profiles/ncu_traffic.json)"),
            "roofline_by_class": classes,
            "clocks": clocks,
            "numa": {"node": numa_node, "cpus": len(os.sched_getaffinity(0)),
                     "bound": os.environ.get("ARU_NUMA_BIND", "1") != "0"},
        }
        # bounded CPU baseline on the host cores of this box
        if world == 1 and not args.no_cpu:
            pps, s_per, cores, threads = cpu_pages_per_s(2, 1)
            line["cpu_baseline"] = {"value": pps, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"2 x 1 page of {H}x{W}, PyTorch-CPU fp32 oracle of the same GraphDef, "
                                              f"{threads} threads on {cores} host cores ({s_per:.2f} s/page)"}
        else:
            line["cpu_baseline"] = None
        del ctypes, mb
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pages", type=int, default=PAGES_PER_STEP, help="pages per GPU per step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--micro-batch", type=int, default=0, help="pages per pass through the net (0 = the engine's choice)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
