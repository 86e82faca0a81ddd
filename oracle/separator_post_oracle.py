"""CPU oracle (TEST INFRASTRUCTURE - never imported by the product path) for the integer steps either side
of the ARU-Net forward pass.  numpy / scipy restatement of

  * ``load_and_scale_image``'s colour step: ``cv2.cvtColor(image, cv2.COLOR_BGR2GRAY) / 255.0``
    (net_post_processing_helper.py:28-33) -> ``bgr2gray_u8`` / ``u8_to_net_input``;
  * ``np.array(net_output * 255, dtype=np.uint8)`` and ``apply_threshold`` (separator_net_post_processor.py:147-149,
    net_post_processing_helper.py:75-78) -> ``quantize_u8`` / ``apply_threshold``;
  * ``RegionNetPostProcessor.apply_cc_analysis`` (region_net_post_processor_base.py:230-251) -> ``cc_size_filter``;
  * ``SeparatorNetPostProcessor.post_process`` (separator_net_post_processor.py:25-99) -> ``separator_post_process``;
  * ``HeadingNetPostProcessor.post_process`` / ``get_net_prob_for_text_line`` (heading_net_post_processor.py:203-209,
    247-270) -> ``heading_post_process`` / ``box_sum_u8`` / ``net_prob_for_box``.

The arithmetic of the reference lives in OpenCV (``cv2.connectedComponentsWithStats``, ``cv2.morphologyEx``,
``cv2.subtract``, ``cv2.cvtColor``); the OpenCV rules restated here:

  * 8-bit BGR -> gray is fixed point: ``(B*3735 + G*19235 + R*9798 + 2^14) >> 15``;
  * ``cv2.erode`` / ``cv2.dilate`` with a rectangular element of size (kw, kh) and the default anchor (kw//2, kh//2):
    ``dst(x) = min / max over j in [0,k) of src(x + j - k//2)``; taps outside the image are ignored (the default
    border value is +inf for erode, -inf for dilate).  Erode and dilate use the SAME offsets, so for an even k
    ``MORPH_OPEN`` is the true opening shifted by one pixel towards +x / +y - reproduced exactly here;
  * ``cv2.subtract`` on uint8 saturates at 0.

PINNED: unlike the network forward pass (whose arithmetic needs TensorFlow 1.x) this part of the reference RUNS in
the build container.  ``tests/golden/make_post_golden.py`` imports the reference's own ``SeparatorNetPostProcessor``
(with stub modules for its unrelated imports), runs ``post_process`` and ``cv2.cvtColor`` on seeded inputs, asserts
this restatement reproduces them bit for bit and commits the vectors as ``tests/golden/post_*.npz``.
"""
from __future__ import annotations

import numpy as np


# ---- colour step / net input (helper.py:28-33) ---------------------------------------------------------------------
def bgr2gray_u8(bgr: np.ndarray) -> np.ndarray:
    """uint8 [...,3] (B, G, R) -> uint8 [...], OpenCV's 8-bit fixed-point luma."""
    b = bgr[..., 0].astype(np.int64)
    g = bgr[..., 1].astype(np.int64)
    r = bgr[..., 2].astype(np.int64)
    return ((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15).astype(np.uint8)


def u8_to_net_input(gray: np.ndarray) -> np.ndarray:
    """``gray / 255.0`` (float64, helper.py:31) as the float32 the graph's placeholder receives."""
    return (gray / 255.0).astype(np.float32)


# ---- integer forms of the probability map (sep:147-149, helper.py:75-78) ------------------------------------------
def quantize_u8(prob: np.ndarray) -> np.ndarray:
    return np.array(prob * 255, dtype=np.uint8)


def apply_threshold(net_output: np.ndarray, threshold: float) -> np.ndarray:
    if net_output.dtype == np.uint8:
        threshold = threshold * 255
    return np.array((net_output > threshold) * 255, dtype=np.uint8)


# ---- connected-component size filter (base:230-251) ----------------------------------------------------------------
def cc_min_size(n_pixels: int) -> int:
    """``int(net_output.size * threshold)`` with ``threshold = 1 / net_output.size * 100`` (sep:37, base:244):
    100, or 99 where the float product rounds below 100."""
    return int(n_pixels * (1 / n_pixels * 100))


def cc_size_filter(mask: np.ndarray, min_size: int) -> np.ndarray:
    """Keep the 8-connected components of the non-zero pixels whose area is >= min_size; output {0, 255}."""
    from scipy import ndimage
    lab, n = ndimage.label(mask != 0, structure=np.ones((3, 3), dtype=bool))
    area = np.bincount(lab.ravel(), minlength=n + 1)
    keep = area >= min_size
    keep[0] = False
    return np.where(keep[lab], 255, 0).astype(mask.dtype)


# ---- rectangular morphology with OpenCV's anchor / border rules -----------------------------------------------------
def _slide(img: np.ndarray, k: int, axis: int, op: str) -> np.ndarray:
    n = img.shape[axis]
    a = k // 2
    out = None
    fill = 255 if op == "min" else 0
    for j in range(k):
        s = j - a                       # dst(x) takes src(x + s)
        sh = np.full_like(img, fill)
        src = [slice(None)] * img.ndim
        dst = [slice(None)] * img.ndim
        if s >= 0:
            if s >= n:
                continue
            src[axis] = slice(s, n)
            dst[axis] = slice(0, n - s)
        else:
            if -s >= n:
                continue
            src[axis] = slice(0, n + s)
            dst[axis] = slice(-s, n)
        sh[tuple(dst)] = img[tuple(src)]
        out = sh if out is None else (np.minimum(out, sh) if op == "min" else np.maximum(out, sh))
    return out


def erode_rect(img: np.ndarray, kw: int, kh: int) -> np.ndarray:
    out = img
    if kw > 1:
        out = _slide(out, kw, 1, "min")
    if kh > 1:
        out = _slide(out, kh, 0, "min")
    return out


def dilate_rect(img: np.ndarray, kw: int, kh: int) -> np.ndarray:
    out = img
    if kw > 1:
        out = _slide(out, kw, 1, "max")
    if kh > 1:
        out = _slide(out, kh, 0, "max")
    return out


def open_rect(img: np.ndarray, kw: int, kh: int) -> np.ndarray:
    """``cv2.morphologyEx(img, cv2.MORPH_OPEN, cv2.getStructuringElement(cv2.MORPH_RECT, (kw, kh)))``."""
    return dilate_rect(erode_rect(img, kw, kh), kw, kh)


# ---- SeparatorNetPostProcessor.post_process (sep:25-99) -------------------------------------------------------------
def separator_kernel_sizes(h: int, w: int):
    """(horizontal opening width, vertical opening height, clean-up width), sep:71,76,85."""
    return int(15 * w / 1000), int(30 * h / 1500), int(10 * w / 1000)


def separator_post_process(mask: np.ndarray):
    """mask: uint8 [H,W] or [H,W,C] thresholded net output ({0,255}; channel 0 is used, sep:33).
    Returns (horizontal, vertical) uint8 {0,255} masks."""
    if mask.ndim == 3:
        mask = mask[:, :, 0]
    h, w = mask.shape
    kh1, kv, kh2 = separator_kernel_sizes(h, w)
    if min(kh1, kv, kh2) < 1:
        raise ValueError(f"page {h}x{w} too small: OpenCV rejects an empty structuring element")
    post = cc_size_filter(mask, cc_min_size(mask.size))
    horizontal = open_rect(post, kh1, 1)
    vertical = open_rect(post, 1, kv)
    horizontal = np.where(horizontal > vertical, horizontal - np.minimum(horizontal, vertical), 0).astype(np.uint8)
    horizontal = open_rect(horizontal, kh2, 1)
    return horizontal, vertical


# ---- HeadingNetPostProcessor: network feature of a text line (head:203-209, 247-270) --------------------------------
def heading_post_process(net_output_u8: np.ndarray) -> np.ndarray:
    """``net_output[:, :, 0] / 255`` (head:209)."""
    return net_output_u8[:, :, 0] / 255


def box_sum_u8(net_output_u8: np.ndarray, ya: int, yb: int, xa: int, xb: int) -> int:
    """Exact integer sum of channel 0 over ``[ya:yb, xa:xb]`` with numpy slice semantics (what the device returns)."""
    ch0 = net_output_u8[:, :, 0] if net_output_u8.ndim == 3 else net_output_u8
    return int(ch0[ya:yb, xa:xb].astype(np.int64).sum())


def net_prob_for_box(net_output_u8: np.ndarray, x: int, y: int, width: int, height: int) -> float:
    """``get_net_prob_for_text_line`` for a bounding box (x, y, width, height) (head:262-270):
    ``np.sum((u8/255)[ya:yb, xa:xb]) / (width * height)``, evaluated as integer sum / 255 / area (equal to the reference's
    float64 pairwise sum to ~1e-13 relative; the integer sum itself is exact)."""
    s = box_sum_u8(net_output_u8, y, y + height, x, x + width)
    return s / 255 / (width * height)
