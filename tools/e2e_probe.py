import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
import __graft_entry__ as entry
entry.build()
import torch
from aru_b200.engine import Engine, pinned_empty, OPT_MICRO_BATCH
from aru_b200.synth import synth_page, synth_pb
eng = Engine(synth_pb("separator"), device=0)
n, h, w = 256, 1024, 768
pages = pinned_empty((n, h, w), np.uint8)
for i in range(n): pages[i] = synth_page(h, w, seed=i % 4)
xf = pinned_empty((n, h, w), np.float32); xf[...] = pages.astype(np.float32) / np.float32(255)
def T(label, fn, reps=2):
    fn(); t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    dt = (time.perf_counter() - t0) / reps
    print(f"{label}: {dt*1e3:.1f} ms -> {n/dt:.0f} pages/s", flush=True); return r
x_dev = torch.from_numpy(xf).cuda(); y_dev = torch.empty((n, h, w, 2), dtype=torch.float32, device="cuda")
def dev():
    eng.forward_device(x_dev.data_ptr(), n, h, w, out_ptr=y_dev.data_ptr()); eng.sync()
T("device-resident forward_device", dev)
T("forward float (prob out)", lambda: eng.forward(xf))
T("forward float, u8 out only", lambda: eng.forward(xf, want_prob=False, want_u8=True))
T("separator_pages mask only", lambda: eng.separator_pages(pages, want_separators=False, want_mask=True))
r = T("separator_pages full", lambda: eng.separator_pages(pages, want_mask=True))
print("mask fg frac", float((r["mask"] > 0).mean()), "hor", int((r["horizontal"]>0).sum()), "ver", int((r["vertical"]>0).sum()))
m = np.ascontiguousarray(r["mask"][:32])
t0 = time.perf_counter(); eng.separator_post(m); print("separator_post 32 masks (host in/out):", (time.perf_counter()-t0)*1e3, "ms")
for mb in (8, 16, 32):
    eng.set_option(OPT_MICRO_BATCH, mb)
    T(f"forward float mb={mb}", lambda: eng.forward(xf))
