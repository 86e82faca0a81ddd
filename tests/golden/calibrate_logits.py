"""Offline calibration of the synthetic nets' classifier gain/bias (run once, results pasted into
aru_b200/synth.py NETS).  Uses the CPU oracle, hence lives under tests/.

Goal: with random weights the 2-class softmax saturates; rescale the 4x4
classifier so that the class-0 logit margin d = l0 - l1 has std 2 and a chosen quantile of the
pixels sits exactly on the consumer's threshold (separator: p0 = 0.05 at the 95 % quantile,
separator_net_post_processor.py:147-149; heading: p0 = 0.4 at the 80 % quantile,
run_net_post_processing.py:15-23).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from aru_b200.synth import NETS, synth_pb, synth_page, page_to_net_input  # noqa: E402
from oracle.aru_oracle import Oracle  # noqa: E402

TARGETS = {"separator": (0.95, 0.05), "heading": (0.80, 0.4), "ru": (0.80, 0.4), "aru_s6a5": (0.95, 0.05)}

if __name__ == "__main__":
    x = page_to_net_input(synth_page(1024, 768, 0))
    for name, (q, p) in TARGETS.items():
        o = Oracle(synth_pb(name, logit_gain=1.0, logit_bias=(0.1, 0.1)))
        lg = o.run(x, fetch="aru_net/logit/logits")[0]
        d = (lg[..., 0] - lg[..., 1]).astype(np.float64)
        g = 2.0 / d.std()
        db = np.log(p / (1 - p)) - g * np.quantile(d, q)
        print(f'"{name}": logit_gain={g:.6f}, logit_bias=({db:.6f}, 0.0)')
        o2 = Oracle(synth_pb(name, logit_gain=round(g, 6), logit_bias=(round(db, 6), 0.0)))
        pr = o2.run(x)[0][..., 0]
        print("   check: frac(p0 > thr) =", float((pr > p).mean()), " p0 quantiles", np.quantile(pr, [.05, .5, .95]))
