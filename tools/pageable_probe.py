"""Where one pageable float64 page per call (what get_net_output receives, helper.py:31,56-72) spends its time.
usage: python tools/pageable_probe.py [h] [w]"""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import __graft_entry__ as entry

entry.build()
import torch  # noqa: E402
from aru_b200.engine import Engine, pinned_empty  # noqa: E402
from aru_b200.synth import page_to_net_input, synth_page, synth_pb  # noqa: E402

h, w = (int(v) for v in (sys.argv[1:3] if len(sys.argv) > 2 else (1856, 1344)))
eng = Engine(synth_pb("separator"), device=0)
page64 = page_to_net_input(synth_page(h, w, 0)).astype(np.float64)
page32 = pinned_empty((1, h, w), np.float32)
page32[0] = page64


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


x_dev = torch.from_numpy(page32).cuda()
y_dev = torch.empty((1, h, w, eng.n_class), dtype=torch.float32, device="cuda")
dst = pinned_empty((1, h, w), np.float32)
print(f"{h}x{w}:")
print(f"  numpy float64 -> pinned float32           {timed(lambda: np.copyto(dst, page64[None], casting='unsafe')):7.3f} ms")
print(f"  aru_f64_to_f32 -> pinned float32          {timed(lambda: eng.lib.aru_f64_to_f32(ctypes.c_void_p(page64.ctypes.data), ctypes.c_void_p(dst.ctypes.data), page64.size, 0)):7.3f} ms")
print(f"  pinned_empty of the result                {timed(lambda: pinned_empty((1, h, w, eng.n_class), np.float32)):7.3f} ms")
print(f"  device-resident pass (1 page)             {timed(lambda: (eng.forward_device(x_dev.data_ptr(), 1, h, w, out_ptr=y_dev.data_ptr()), eng.sync())):7.3f} ms")
print(f"  forward(pinned float32 page)              {timed(lambda: eng.forward(page32)):7.3f} ms")
print(f"  forward(pinned float32 page), uint8 ch 0  {timed(lambda: eng.forward(page32, want_u8=True, want_prob=False, u8_channels=1)):7.3f} ms")
def by_hand():
    xin = pinned_empty((1, h, w), np.float32)
    eng.lib.aru_f64_to_f32(ctypes.c_void_p(page64.ctypes.data), ctypes.c_void_p(xin.ctypes.data), page64.size, 0)
    return eng.forward(xin)


print(f"  pinned_empty + aru_f64_to_f32 + forward   {timed(by_hand):7.3f} ms")
print(f"  forward(pageable float64 page)            {timed(lambda: eng.forward(page64)):7.3f} ms")
