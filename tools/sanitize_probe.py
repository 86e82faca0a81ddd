"""Small passes through every tcgen05 kernel family for compute-sanitizer (racecheck / synccheck / memcheck):
  compute-sanitizer --tool racecheck python tools/sanitize_probe.py
conv_tc (position-major), conv_band (row-banded, incl. residual / pooled / head launches) and conv_band2 (fused pairs) on
the tiny and the separator nets, plus the integer pre / post-processing calls."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import __graft_entry__ as entry

entry.build()
from aru_b200.engine import Engine, OPT_CONV_PATH, OPT_FUSE_PAIRS, OPT_USE_GRAPH  # noqa: E402
from aru_b200.synth import page_to_net_input, synth_page, synth_pb  # noqa: E402

for net, h, w in (("tiny", 40, 37), ("separator", 96, 80)):
    eng = Engine(synth_pb(net), device=0)
    eng.set_option(OPT_USE_GRAPH, 0)
    x = np.stack([page_to_net_input(synth_page(h, w, s)) for s in range(2)]).astype(np.float32)
    outs = {}
    for path in (0, 2, 3):
        eng.set_option(OPT_CONV_PATH, path)
        outs[path] = eng.forward(x).copy()
    eng.set_option(OPT_CONV_PATH, 0)
    eng.set_option(OPT_FUSE_PAIRS, 1)
    outs["fused"] = eng.forward(x).copy()
    kernels = sorted({k for _, k, _ in eng.profile_ops(1)})
    eng.set_option(OPT_FUSE_PAIRS, 0)
    pages = np.stack([synth_page(max(h, 64), max(w, 112), s) for s in range(2)])
    r = eng.separator_pages(pages, want_u8=True, want_mask=True)
    d = eng.swt_distance(pages)
    print(net, h, w, "max spread between conv paths", max(float(np.abs(outs[0] - v).max()) for v in outs.values()),
          "kernels (fused pass):", kernels, "swt max", int(d.max()), flush=True)
    eng.close()
print("sanitize probe done")
