"""Cost of meeting a new page shape: get_net_output on single pages whose widths differ (real scans scaled to a fixed
height do).  usage: python tools/shape_churn_probe.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402

entry.build()
from aru_b200.engine import Engine  # noqa: E402
from aru_b200.synth import synth_page, synth_pb  # noqa: E402

eng = Engine(synth_pb("separator"), device=0)
widths = [1125, 1100, 1163, 1125, 1088, 1100, 1201, 1163, 1125, 1142, 1088, 1201]
for w in widths:
    x = synth_page(1500, w, seed=w) / 255.0
    t0 = time.perf_counter()
    y = eng.forward(x)
    t1 = time.perf_counter()
    y = eng.forward(x)
    t2 = time.perf_counter()
    print(f"1500x{w}: first call {1e3 * (t1 - t0):8.1f} ms, second call {1e3 * (t2 - t1):6.2f} ms", flush=True)
