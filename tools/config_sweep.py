"""Times the BASELINE.json configurations other than the benchmark one (configs[0], [1], [3] per GPU, [4]) on one GPU
and prints one JSON line per case.  Wall-clock around the public calls (host buffers in, host buffers out).
usage: python tools/config_sweep.py [--quick]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402

entry.build()
from aru_b200.engine import Engine, pinned_empty  # noqa: E402
from aru_b200.synth import synth_page, synth_pb  # noqa: E402

quick = "--quick" in sys.argv


def pages_u8(n, h, w):
    base = [synth_page(h, w, seed=s) for s in range(min(n, 4))]
    out = pinned_empty((n, h, w), np.uint8)
    for i in range(n):
        out[i] = np.roll(base[i % len(base)], (5 * (i // len(base)), 11 * (i // len(base))), axis=(0, 1))
    return out


def timed(fn, reps):
    # steady state of a caller that keeps the previous result while the next call runs: two warm-up calls create both
    # sets of pinned result buffers (page-locking a fresh buffer costs ~0.25 ms per MB and would dominate otherwise)
    r = fn()
    r = fn()  # noqa: F841
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    return (time.perf_counter() - t0) / reps, r


def report(**kw):
    print(json.dumps(kw), flush=True)


engines = {net: Engine(synth_pb(net), device=0) for net in ("separator", "heading")}

# configs[0] / configs[1]: one 1024x768 page through get_net_output (float64 gray/255 in, float32 map out), and through
# the page-level calls (uint8 in, masks / text-line sums out)
page = pages_u8(1, 1024, 768)
x64 = page[0] / 255.0
boxes = [(0, 40 + 30 * i, 60 + 30 * i, 50, 700) for i in range(30)]
for net in ("separator", "heading"):
    eng = engines[net]
    dt, out = timed(lambda: eng.forward(x64), 20)
    report(config=f"configs[{0 if net == 'separator' else 1}]", net=net, page="1024x768", call="get_net_output (Engine.forward)",
           ms_per_page=round(dt * 1e3, 3), pages_per_s=round(1 / dt, 1), out_shape=list(out.shape))
    if net == "separator":
        dt, r = timed(lambda: eng.separator_pages(page), 20)
        report(config="configs[0]", net=net, page="1024x768", call="Engine.separator_pages (uint8 in, horizontal+vertical out)",
               ms_per_page=round(dt * 1e3, 3), pages_per_s=round(1 / dt, 1))
    else:
        dt, r = timed(lambda: eng.heading_pages(page, boxes), 20)
        report(config="configs[1]", net=net, page="1024x768", call="Engine.heading_pages (uint8 in, 30 text-line sums out)",
               ms_per_page=round(dt * 1e3, 3), pages_per_s=round(1 / dt, 1))

# configs[3], one GPU's share: 1024x768 pages through both nets
n = 64 if quick else 256
pages = pages_u8(n, 1024, 768)
xf = pinned_empty((n, 1024, 768), np.float32)
xf[...] = pages.astype(np.float32) / np.float32(255)
tot = 0.0
for net in ("separator", "heading"):
    eng = engines[net]
    dt, out = timed(lambda: eng.forward(xf), 3)
    tot += dt
    report(config="configs[3] (per GPU)", net=net, page="1024x768", pages=n, call="Engine.forward (float32 in / float32 out)",
           ms_per_step=round(dt * 1e3, 2), pages_per_s=round(n / dt, 1))
dts, _ = timed(lambda: engines["separator"].separator_pages(pages), 3)
dth, _ = timed(lambda: engines["heading"].heading_pages(pages, [(i % n, 100, 140, 50, 700) for i in range(20 * n)]), 3)
report(config="configs[3] (per GPU)", net="separator+heading", page="1024x768", pages=n,
       call="separator_pages + heading_pages (uint8 in; masks / text-line sums out)",
       ms_per_step=round((dts + dth) * 1e3, 2), pages_per_s_both_nets=round(n / (dts + dth), 1),
       separator_pages_per_s=round(n / dts, 1), heading_pages_per_s=round(n / dth, 1),
       float_path_pages_per_s_both_nets=round(n / tot, 1))
del pages, xf

# the CLI default size (odd at every level) and configs[4]: broadsheet pages
for (h, w, n) in ((1500, 1125, 32 if quick else 64), (6000, 4500, 4 if quick else 8)):
    eng = engines["separator"]
    pages = pages_u8(n, h, w)
    dt, r = timed(lambda: eng.separator_pages(pages, want_u8=True), 3)
    u8 = r["u8"]
    ok = bool(np.all(u8[..., 0].astype(np.int32) + u8[..., 1].astype(np.int32) >= 253))   # softmax pair sums to ~255
    report(config="configs[4]" if h == 6000 else "CLI default 1500x1125", net="separator", page=f"{h}x{w}", pages=n,
           call="Engine.separator_pages (uint8 in; uint8 map + masks out)", ms_per_page=round(dt / n * 1e3, 2),
           pages_per_s=round(n / dt, 2), mpix_per_s=round(n * h * w / dt / 1e6, 1), softmax_pair_sums_ok=ok,
           separator_px=int((r["horizontal"] > 0).sum() + (r["vertical"] > 0).sum()))
    del pages, r, u8
