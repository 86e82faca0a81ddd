"""GPU diagnostic (test infrastructure: it uses the CPU oracle, so it lives under tests/): per-layer comparison of the tcgen05 conv path against the CUDA-core path and the oracle.
Prints one line per lowered op; for mismatching ops a summary of where the errors are."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402

entry.build()
from aru_b200.engine import Engine, EngineError, OPT_CONV_PATH, OPT_USE_GRAPH  # noqa: E402
from aru_b200.synth import synth_pb, synth_page, page_to_net_input  # noqa: E402
from oracle.aru_oracle import Oracle  # noqa: E402


def diag(net, h, w, n=1, max_report=4):
    print(f"==== {net} {n}x{h}x{w}", flush=True)
    pb = synth_pb(net)
    x = np.stack([page_to_net_input(synth_page(h, w, 100 + i)) for i in range(n)])
    eng = Engine(pb, device=0)
    eng.set_option(OPT_USE_GRAPH, 0)
    vals = {}
    for path in (1, 0):
        eng.set_option(OPT_CONV_PATH, path)
        t = time.time()
        try:
            out = eng.forward(x).copy()
        except EngineError as ex:
            print(f"  path {path}: forward FAILED: {ex}", flush=True)
            eng.close()
            return False
        vals[path] = {"out": out, "dt": time.time() - t}
        nodes = {}
        for node, v in eng.program.tensor_of_node.items():
            try:
                nodes[node] = eng.read_node(node, page=n - 1)
            except EngineError as ex:
                print("  read_node failed", node, ex)
        vals[path]["nodes"] = nodes
        vals[path]["kernels"] = {o.name: k for (o, (_, k, _)) in zip(eng.program.ops, eng.profile_ops(1))}
    ref = Oracle(pb).run(x)
    for path in (1, 0):
        d = np.abs(vals[path]["out"] - ref)
        print(f"  path {path}: vs oracle max {d.max():.3e} mean {d.mean():.3e}  ({vals[path]['dt']:.2f}s)", flush=True)
    bad = 0
    for op in eng.program.ops:
        k = vals[0]["kernels"].get(op.name)
        if not ((k or "").startswith("conv_tc") or k in ("deconv_tc", "conv_band", "conv_band_head")):
            continue
        # the op's output node: find by view
        for node, v in eng.program.tensor_of_node.items():
            if (v.buf, v.ch_off, v.ch) == (op.out.buf, op.out.ch_off, op.out.ch):
                a, b = vals[0]["nodes"].get(node), vals[1]["nodes"].get(node)
                if a is None or b is None:
                    continue
                d = np.abs(a - b)
                scale = max(1.0, float(np.abs(b).max()))
                ok = d.max() <= 1e-2 * scale
                if not ok:
                    bad += 1
                    if bad <= max_report:
                        ys, xs, cs = np.nonzero(d > 1e-2 * scale)
                        print(f"  MISMATCH {op.name} ks={op.ksize} cin={op.inp.ch} cout={op.out.ch} shape={a.shape} "
                              f"max {d.max():.3e} scale {scale:.2f} nbad {len(ys)}/{d.size}")
                        print(f"     rows {ys.min()}..{ys.max()} cols {xs.min()}..{xs.max()} chans {sorted(set(cs.tolist()))[:16]}")
                        print(f"     tc[0,:4,0]={a[0,:4,0]} direct={b[0,:4,0]}  tc nonfinite={np.isnan(a).sum()} tc zeros={np.mean(a==0):.3f} direct zeros={np.mean(b==0):.3f}")
                break
    from collections import Counter
    print("  kernels:", dict(Counter(vals[0]["kernels"].values())))
    print(f"  conv_tc ops mismatching direct path: {bad}", flush=True)
    eng.close()
    return bad == 0


if __name__ == "__main__":
    ok = True
    for net, h, w, n in [("tiny", 32, 32, 1), ("tiny", 45, 39, 2), ("tiny", 20, 300, 1), ("separator", 128, 96, 1),
                         ("separator", 150, 113, 2), ("separator", 40, 1250, 1), ("separator", 300, 420, 3)]:
        ok = diag(net, h, w, n) and ok
    print("TC_DIAG", "PASS" if ok else "FAIL")
