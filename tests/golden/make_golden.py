"""Generates tests/golden/*.npz - input pages and fp32 oracle outputs for the synthetic nets.

The reference ships no golden vectors for this path and its TF1 run cannot be
executed here, so the fixtures pin the CPU oracle (oracle/aru_oracle.py) *after* it has been
cross-checked against OpenCV's independent TF importer on the same GraphDef bytes (asserted below).
Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from aru_b200.synth import synth_pb, synth_page, page_to_net_input  # noqa: E402
from oracle.aru_oracle import Oracle  # noqa: E402
from oracle.cv2_oracle import run_cv2  # noqa: E402

CASES = [  # (net, H, W, page seed)
    ("tiny", 37, 29, 1),
    ("tiny", 64, 48, 2),
    ("tiny_sigmoid", 33, 40, 3),
    ("ru", 50, 35, 4),
    ("separator", 96, 80, 5),
    ("separator", 101, 77, 6),
    ("heading", 75, 57, 7),
    ("aru_s6a5", 70, 66, 8),
]

if __name__ == "__main__":
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for net, h, w, seed in CASES:
        pb = synth_pb(net)
        page = synth_page(h, w, seed)
        x = page_to_net_input(page)
        y = Oracle(pb).run(x)[0]
        y_cv = run_cv2(pb, x)
        d = float(np.abs(y - y_cv).max())
        assert d < 2e-5, (net, h, w, d)
        name = f"{net}_{h}x{w}.npz"
        np.savez_compressed(os.path.join(out_dir, name), page=page, prob=y.astype(np.float32),
                            pb_sha256=hashlib.sha256(pb).hexdigest(), cv2_max_abs_diff=d)
        print(name, y.shape, "oracle-vs-cv2", d, "frac>0.05", float((y[..., 0] > 0.05).mean()))
