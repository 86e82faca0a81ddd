"""Two fresh engines, same pages: where do the outputs differ?  usage: python tools/determinism_probe.py [n] [h] [w]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as entry
entry.build()
from aru_b200.engine import Engine, OPT_FUSE_PAIRS, OPT_USE_GRAPH
from aru_b200.synth import synth_pb, synth_page, page_to_net_input

n, h, w = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (2, 1856, 1344)))
x = np.stack([page_to_net_input(synth_page(h, w, s)) for s in range(n)]).astype(np.float32)
for fuse in (0, 1):
    outs = []
    for rep in range(3):
        eng = Engine(synth_pb("separator"), device=0)
        eng.set_option(OPT_FUSE_PAIRS, fuse)
        a = eng.forward(x).copy()
        b = eng.forward(x).copy()
        print(f"fuse={fuse} engine {rep}: same engine twice equal: {np.array_equal(a, b)}")
        outs.append(a)
        eng.close()
    for rep in (1, 2):
        d = np.abs(outs[0] - outs[rep])[..., 0]
        nz = np.argwhere(d > 0)
        print(f"fuse={fuse}: engines 0 vs {rep}: differing px {len(nz)} of {d.size}, max {d.max():.3e}")
        if len(nz):
            print("   pages", np.unique(nz[:, 0]), "rows", nz[:, 1].min(), nz[:, 1].max(), "cols", nz[:, 2].min(), nz[:, 2].max())
            print("   first", nz[:5].tolist())
