"""Synthetic inputs for tests and benchmarks.

The reference ships no sample images and no frozen graphs, so both are synthesised:
  * pages: uint8 grayscale, background N(225, 6^2) clipped, text-like dark runs in a column
    layout and 1-3 px wide vertical / horizontal rules, ``numpy.default_rng(seed=page_index)``;
  * graphs: ``graphdef.build_aru_graphdef`` with fixed seeds; ``SEPARATOR`` / ``HEADING`` name the
    two configurations that play ``separator_detection_net.pb`` / ``heading_detection_net.pb``.
The classifier gain/bias were calibrated offline with ``tests/golden/calibrate_logits.py`` so
that about 5 % of the pixels of a synthetic page exceed the reference's 0.05 separator threshold
(separator_net_post_processor.py:147-149) instead of the mask being trivially all-on.
"""
from __future__ import annotations

import numpy as np

from .graphdef import build_aru_graphdef

# name -> kwargs of build_aru_graphdef (logit_gain/logit_bias from calibrate_logits.py)
NETS = {
    "separator": dict(graph="ARU", scale_space_num=5, num_scales_att=3, n_class=2, seed=0,
                      logit_gain=0.773170, logit_bias=(-15.335476, 0.0)),
    "heading": dict(graph="ARU", scale_space_num=5, num_scales_att=3, n_class=2, seed=1,
                    logit_gain=1.112696, logit_bias=(4.220346, 0.0)),
    "ru": dict(graph="RU", scale_space_num=5, n_class=2, seed=2, logit_gain=0.592449, logit_bias=(-5.465143, 0.0)),
    "aru_s6a5": dict(graph="ARU", scale_space_num=6, num_scales_att=5, n_class=2, seed=3,
                     logit_gain=1.072904, logit_bias=(-19.107500, 0.0)),
    # small nets for fast CPU tests / per-layer parity
    "tiny": dict(graph="ARU", scale_space_num=3, num_scales_att=2, n_class=2, seed=4,
                 logit_gain=0.5, logit_bias=(-3.0, 0.0)),
    "tiny_sigmoid": dict(graph="RU", scale_space_num=2, n_class=1, seed=5, output="sigmoid",
                         logit_gain=0.5, logit_bias=(-1.0,)),
    "tiny_aru_sigmoid": dict(graph="ARU", scale_space_num=3, num_scales_att=2, n_class=1, seed=6, output="sigmoid",
                             logit_gain=0.5, logit_bias=(-1.0,)),
}


def synth_pb(name: str = "separator", **overrides) -> bytes:
    kw = dict(NETS[name])
    kw.update(overrides)
    return build_aru_graphdef(**kw)


def synth_page(h: int, w: int, seed: int = 0) -> np.ndarray:
    """uint8 [h, w] newspaper-like page."""
    rng = np.random.default_rng(seed)
    img = np.clip(rng.normal(225.0, 6.0, size=(h, w)), 0, 255)
    if min(h, w) < 8:                                       # too small for a layout: noise only
        return img.astype(np.uint8)
    n_cols = int(rng.integers(2, 6))
    margin = max(1, min(max(4, w // 40), (min(h, w) - 1) // 2))
    n_cols = max(1, min(n_cols, (w - margin) // (margin + 8)))
    col_w = max(4, (w - margin * (n_cols + 1)) // n_cols)
    line_h = max(3, h // 170)
    for c in range(n_cols):
        x0 = margin + c * (col_w + margin)
        y = margin + int(rng.integers(0, 3 * line_h))
        while y + line_h < h - margin:
            if rng.random() < 0.04:                        # paragraph gap
                y += int(rng.integers(2, 6)) * line_h
                continue
            x = x0
            x_end = x0 + col_w - (int(rng.integers(0, max(1, col_w // 2))) if rng.random() < 0.15 else 0)
            th = int(rng.integers(max(2, line_h - 2), line_h + 1))
            while x < x_end:
                wl = int(rng.integers(2, max(3, col_w // 6)))
                x1 = min(x + wl, x_end)
                x1 = min(x1, w)
                if x1 > x:
                    img[y:y + th, x:x1] = rng.normal(60.0, 25.0, size=(min(th, h - y), x1 - x)).clip(0, 255)
                x = x1 + int(rng.integers(1, 4))
            y += line_h + int(rng.integers(1, 3))
        if c < n_cols - 1 and rng.random() < 0.8:           # vertical rule between columns
            xr = x0 + col_w + margin // 2
            t = int(rng.integers(1, 4))
            y0, y1 = sorted(int(v) for v in rng.integers(margin, h - margin, size=2))
            if y1 - y0 > h // 8 and xr < w:
                img[y0:y1, xr:xr + t] = rng.normal(40.0, 10.0, size=(y1 - y0, min(t, w - xr))).clip(0, 255)
    for _ in range(int(rng.integers(1, 5))):                # horizontal rules
        yr = int(rng.integers(margin, h - margin))
        x0, x1 = sorted(int(v) for v in rng.integers(margin, w - margin, size=2))
        t = int(rng.integers(1, 4))
        if x1 - x0 > w // 8:
            img[yr:yr + t, x0:x1] = rng.normal(40.0, 10.0, size=(min(t, h - yr), x1 - x0)).clip(0, 255)
    return img.astype(np.uint8)


def page_to_net_input(page_u8: np.ndarray) -> np.ndarray:
    """What the reference feeds the net: gray / 255.0 as float64 (net_post_processing_helper.py:31)."""
    return page_u8 / 255.0


def synth_separator_mask(h: int, w: int, seed: int = 0, noise: float = 0.02) -> np.ndarray:
    """uint8 [h, w] in {0,255}: what a thresholded separator map looks like - long 1-6 px wide vertical and horizontal
    rules (some broken, some crossing, some touching the borders), diagonal strokes, text-sized blobs around the
    100-pixel component-size limit and salt noise.  Exercises every branch of the reference's post_process
    (separator_net_post_processor.py:25-99): components below / above the size limit, runs shorter / longer than the
    structuring elements, horizontal-vertical crossings removed by the subtract, border-touching runs."""
    rng = np.random.default_rng(seed)
    m = rng.random((h, w)) < noise
    for _ in range(int(rng.integers(3, 9))):                       # vertical rules
        x = int(rng.integers(0, w))
        y0, y1 = sorted(int(v) for v in rng.integers(-h // 8, h + h // 8, size=2))
        t = int(rng.integers(1, 7))
        m[max(y0, 0):max(y1, 0), x:x + t] = True
        if rng.random() < 0.4 and y1 - y0 > 8:                     # a gap
            g = int(rng.integers(max(y0, 0), max(y1, 1)))
            m[g:g + int(rng.integers(1, 6)), x:x + t] = False
    for _ in range(int(rng.integers(3, 9))):                       # horizontal rules
        y = int(rng.integers(0, h))
        x0, x1 = sorted(int(v) for v in rng.integers(-w // 8, w + w // 8, size=2))
        t = int(rng.integers(1, 7))
        m[y:y + t, max(x0, 0):max(x1, 0)] = True
    for _ in range(int(rng.integers(1, 5))):                       # diagonal strokes (8-connectivity only)
        y, x = int(rng.integers(0, h)), int(rng.integers(0, w))
        dx = 1 if rng.random() < 0.5 else -1
        for i in range(int(rng.integers(10, max(11, min(h, w))))):
            yy, xx = y + i, x + dx * i
            if 0 <= yy < h and 0 <= xx < w:
                m[yy, xx] = True
    for _ in range(int(rng.integers(5, 30))):                      # blobs around the size limit
        bh, bw = int(rng.integers(2, 16)), int(rng.integers(2, 40))
        y, x = int(rng.integers(0, max(1, h - bh))), int(rng.integers(0, max(1, w - bw)))
        m[y:y + bh, x:x + bw] |= rng.random((min(bh, h - y), min(bw, w - x))) < 0.8
    return np.where(m, 255, 0).astype(np.uint8)
