import os, sys
sys.path.insert(0, os.getcwd())
import __graft_entry__ as entry
entry.build()
from aru_b200.engine import Engine, OPT_USE_GRAPH
from aru_b200.synth import synth_pb
eng = Engine(synth_pb("separator"), device=0)
eng.set_option(OPT_USE_GRAPH, 0)
import sys
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
eng.plan(n, 1856, 1344)
