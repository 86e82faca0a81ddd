"""Runs Engine.separator_pages on one micro-batch (for the ncu launch list of the pre / post-processing kernels) and
prints host-side timings of the calls.  usage: python tools/post_probe.py [n] [h] [w] [iters]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402

entry.build()
from aru_b200.engine import Engine, pinned_empty  # noqa: E402
from aru_b200.synth import synth_page, synth_pb, synth_separator_mask  # noqa: E402

n, h, w = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (16, 1856, 1344)))
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
eng = Engine(synth_pb("separator"), device=0)
pages = pinned_empty((n, h, w), np.uint8)
for i in range(n):
    pages[i] = synth_page(h, w, seed=i % 4)
for label, kw in (("net only (uint8 in, no outputs copied)", dict(want_separators=False, want_mask=True)),
                  ("net + post_process (masks out)", dict(want_separators=True))):
    eng.separator_pages(pages, **kw)
    t0 = time.perf_counter()
    for _ in range(iters):
        r = eng.separator_pages(pages, **kw)
    dt = (time.perf_counter() - t0) / iters
    print(f"{label}: {dt * 1e3:.2f} ms per {n} pages -> {n / dt:.1f} pages/s")
print("separator pixels: horizontal", int((r["horizontal"] > 0).sum()), "vertical", int((r["vertical"] > 0).sum()))
# worst-case-ish masks for the component labelling: noisy separator masks
masks = np.stack([synth_separator_mask(h, w, seed=s, noise=0.05) for s in range(min(n, 4))])
eng.separator_post(masks)
t0 = time.perf_counter()
eng.separator_post(masks)
print(f"separator_post on {len(masks)} noisy masks (host in/out, allocs included): {(time.perf_counter() - t0) * 1e3:.2f} ms")
