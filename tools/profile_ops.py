"""Per-op CUDA-event profile of one micro-batch pass (engine stream, plain launches).
usage: python tools/profile_ops.py [net] [n] [h] [w] [iters]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402

entry.build()
from aru_b200 import program as P  # noqa: E402
from aru_b200.engine import Engine  # noqa: E402
from aru_b200.synth import synth_pb  # noqa: E402

net = sys.argv[1] if len(sys.argv) > 1 else "separator"
n, h, w = (int(v) for v in (sys.argv[2:5] if len(sys.argv) > 4 else (16, 1856, 1344)))
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 3
eng = Engine(synth_pb(net), device=0)
eng.plan(n, h, w)
prof = eng.profile_ops(iters)


def hw(buf):
    hh, ww, cc = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    eng.lib.aru_buffer_dims(eng.handle, buf, ctypes.byref(hh), ctypes.byref(ww), ctypes.byref(cc))
    return hh.value, ww.value


print(f"# {net} n={n} {h}x{w}; ms = mean of {iters} launches")
print(f"{'op':4} {'kernel':14} {'ks':2} {'cin':>4} {'cout':>4} {'HxW':>11} {'ms':>8} {'GB/s':>8} {'TF/s':>7} res pre  name")
tot = 0.0
for i, (op, (name, kernel, ms)) in enumerate(zip(eng.program.ops, prof)):
    tot += ms
    if op.kind == P.OP_COMBINE:
        print(f"{i:4d} {kernel:14} {'':2} {'':>4} {op.out.ch:4d} {'':>11} {ms:8.4f}")
        continue
    oh, ow = hw(op.out.buf)
    ih, iw = hw(op.inp.buf)
    flops = bytes_ = 0.0
    if op.kind in (P.OP_CONV, P.OP_DECONV):
        px = ih * iw if op.kind == P.OP_DECONV else oh * ow
        flops = 2.0 * n * px * op.ksize ** 2 * op.inp.ch * op.out.ch
    ib = 4 if op.inp.ch == 1 else 2
    ob = 4 if kernel in ("conv_small", "conv_tc_head", "conv_band_head") else 2
    bytes_ = n * (ih * iw * op.inp.ch * ib + oh * ow * op.out.ch * ob)
    if kernel == "conv_stem_pool":     # only the pooled tensor is stored
        bytes_ = n * (ih * iw * 4 + -(-oh // 2) * -(-ow // 2) * op.out.ch * 2)
    if op.res.buf >= 0:
        bytes_ += n * oh * ow * op.out.ch * 2
    if op.out_pre.buf >= 0:
        bytes_ += n * oh * ow * op.out.ch * 2
    print(f"{i:4d} {kernel:14} {op.ksize:2d} {op.inp.ch:4d} {op.out.ch:4d} {oh:5d}x{ow:<5d} {ms:8.4f} {bytes_ / ms / 1e6:8.1f} "
          f"{flops / ms / 1e9:7.2f} {int(op.res.buf >= 0):3d} {int(op.out_pre.buf >= 0):3d}  {name[-48:]}")
print(f"# total {tot:.3f} ms per pass of {n} pages -> {n / tot * 1e3:.1f} pages/s")
