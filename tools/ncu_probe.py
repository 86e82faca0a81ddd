"""One eager pass of every op of a planned micro-batch between cudaProfilerStart/Stop, for
`ncu --profile-from-start off -k regex:<kernel> --launch-skip <i> -c 1` captures of a specific launch (launches appear
in program order, see tools/profile_ops.py for the op list).  usage: python tools/ncu_probe.py [net] [n] [h] [w]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402

entry.build()
import torch  # noqa: E402
from aru_b200.engine import Engine  # noqa: E402
from aru_b200.synth import synth_pb  # noqa: E402

net = sys.argv[1] if len(sys.argv) > 1 else "separator"
n, h, w = (int(v) for v in (sys.argv[2:5] if len(sys.argv) > 4 else (16, 1856, 1344)))
eng = Engine(synth_pb(net), device=0)
eng.plan(n, h, w)
eng.profile_ops(1)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.profile_ops(1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
