"""Error table of the engine against the CPU oracle over nets / shapes (BASELINE tolerances: max-abs 2e-2, mean-abs 1e-3,
mask agreement >= 99.9 %).  usage: [ARU_B200_LIB=.../libaru_b200_bf16.so] python tools/parity_table.py [--assert] [--big]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)) + "/..")
import numpy as np

import __graft_entry__ as entry

entry.build()
from aru_b200.engine import LIB_PATH, Engine  # noqa: E402
from aru_b200.synth import page_to_net_input, synth_page, synth_pb  # noqa: E402
from oracle.aru_oracle import Oracle  # noqa: E402

CASES = [("separator", 128, 96), ("separator", 150, 113), ("separator", 257, 130), ("separator", 40, 1250),
         ("heading", 90, 68), ("ru", 77, 101), ("aru_s6a5", 129, 97), ("tiny", 9, 200),
         ("separator", 1024, 768), ("heading", 1024, 768)]
if "--big" in sys.argv:
    CASES += [("separator", 1500, 1125), ("separator", 1856, 1344)]


def mask(p):
    return (p[..., 0] * 255).astype(np.uint8) > 12.75


rows, ok = [], True
engines, oracles = {}, {}
for net, h, w in CASES:
    if net not in engines:
        engines[net], oracles[net] = Engine(synth_pb(net), device=0), Oracle(synth_pb(net))
    x = page_to_net_input(synth_page(h, w, seed=h * 1000 + w))
    ref = oracles[net].run(x)[0]
    got = engines[net].forward(x)[0]
    d = np.abs(got - ref)
    u8g, u8r = (got * 255).astype(np.uint8), (ref * 255).astype(np.uint8)
    du = np.abs(u8g.astype(int) - u8r.astype(int))
    row = {"net": net, "h": h, "w": w, "max_abs": float(d.max()), "mean_abs": float(d.mean()),
           "mask_agree": float((mask(got) == mask(ref)).mean()), "u8_equal": float((du == 0).mean()),
           "u8_within_1": float((du <= 1).mean())}
    row["within_tolerance"] = bool(row["max_abs"] <= 2e-2 and row["mean_abs"] <= 1e-3 and row["mask_agree"] >= 0.999)
    ok &= row["within_tolerance"]
    rows.append(row)
    print(f"{net:10s} {h:5d}x{w:<5d} max|d| {row['max_abs']:.3e} mean|d| {row['mean_abs']:.3e} mask {row['mask_agree']:.5f} "
          f"u8 equal {row['u8_equal']:.4f} within 1 LSB {row['u8_within_1']:.4f} {'ok' if row['within_tolerance'] else 'OUT OF TOLERANCE'}")
print(json.dumps({"library": os.path.basename(LIB_PATH), "all_within_tolerance": ok, "rows": rows}))
if "--assert" in sys.argv and not ok:
    raise SystemExit(1)
