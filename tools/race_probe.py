"""Find the first tensor that differs between two passes of the same engine (fused conv pairs on).
usage: python tools/race_probe.py [n] [h] [w]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as entry
entry.build()
from aru_b200.engine import Engine, OPT_FUSE_PAIRS, OPT_KEEP_ALL, OPT_USE_GRAPH
from aru_b200.synth import synth_pb, synth_page, page_to_net_input

n, h, w = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (2, 1856, 1344)))
x = np.stack([page_to_net_input(synth_page(h, w, s)) for s in range(n)]).astype(np.float32)
eng = Engine(synth_pb("separator"), device=0)
eng.set_option(OPT_FUSE_PAIRS, 1)
eng.set_option(OPT_KEEP_ALL, int(os.environ.get("KEEP", "1")))
eng.set_option(OPT_USE_GRAPH, 0)
prof_names = None
runs = []
for rep in range(2):
    out = eng.forward(x).copy()
    if prof_names is None:
        prof_names = [k for _, k, _ in eng.profile_ops(1)]
        out = eng.forward(x).copy()
    bufs = {}
    for node in eng.program.tensor_of_node:
        try:
            bufs[node] = eng.read_node(node, page=n - 1).copy()
        except Exception as ex:  # noqa
            bufs[node] = None
    runs.append((out, bufs))
print("final equal:", np.array_equal(runs[0][0], runs[1][0]))
names = list(eng.program.tensor_of_node)
bad = 0
for node in names:
    a, b = runs[0][1][node], runs[1][1][node]
    if a is None or b is None:
        continue
    if not np.array_equal(a, b):
        d = np.abs(a - b)
        nz = np.argwhere(d.max(-1) > 0)
        print(f"DIFF {node[-60:]} shape {a.shape} px {len(nz)} max {d.max():.3e} rows {nz[:,0].min()}..{nz[:,0].max()} cols {nz[:,1].min()}..{nz[:,1].max()}")
        print("     first", nz[:6].tolist())
        bad += 1
        if bad >= 6:
            break
print("kernels:", sorted(set(prof_names)))
