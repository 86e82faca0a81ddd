"""Does a second, independent pass stream on the same GPU raise the device-resident throughput?  Two engines (own
arenas, own streams) fed alternately against one engine, same pages.  usage: python tools/dual_engine_probe.py [n] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import __graft_entry__ as entry

entry.build()
import torch  # noqa: E402
from aru_b200.engine import Engine  # noqa: E402
from aru_b200.synth import page_to_net_input, synth_page, synth_pb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
H, W = 1856, 1344
pages = np.stack([page_to_net_input(synth_page(H, W, s % 4)) for s in range(n)]).astype(np.float32)
x = torch.from_numpy(pages).cuda()


def run(n_eng):
    engs = [Engine(synth_pb("separator"), device=0) for _ in range(n_eng)]
    streams = [torch.cuda.Stream() for _ in range(n_eng)]
    ys = [torch.empty((n, H, W, 2), dtype=torch.float32, device="cuda") for _ in range(n_eng)]
    for _ in range(3):
        for e, s, y in zip(engs, streams, ys):
            e.forward_device(x.data_ptr(), n, H, W, out_ptr=y.data_ptr(), stream=s.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        for e, s, y in zip(engs, streams, ys):
            e.forward_device(x.data_ptr(), n, H, W, out_ptr=y.data_ptr(), stream=s.cuda_stream)
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    same = all(torch.equal(ys[0], y) for y in ys[1:])
    for e in engs:
        e.close()
    return steps * n_eng * n / ms * 1e3, same


for k in (1, 2, 1, 2):
    pps, same = run(k)
    print(f"{k} engine(s), {n} pages per pass: {pps:8.1f} pages/s  (outputs equal: {same})")
