"""CPU oracle (TEST INFRASTRUCTURE - never imported by the product path) for ``scale_image``
(net_post_processing_helper.py:14-25) in the shrinking case: ``cv2.resize(image, None, fx=sc, fy=sc,
interpolation=cv2.INTER_AREA)`` for uint8 images, restated operation by operation in numpy.

The arithmetic lives in OpenCV (the reference's dependency, opencv-python 4.x); the rules restated here:
  * destination size = (cvRound(H*sc), cvRound(W*sc)), rounding half to even; ``scale = 1 / sc`` (a double);
  * general scale: per destination index a list of (source index, float32 weight) entries - a partial first cell when
    ``ceil(fsx1) - fsx1 > 1e-3``, whole cells with weight ``1 / cellWidth``, a partial last cell when
    ``fsx2 - floor(fsx2) > 1e-3`` - computed in double and rounded to float32; a source row is reduced horizontally
    first (``buf += S * alpha``, float32, entries in order, multiply and add rounded separately), rows are then combined
    (``sum = beta_0 * buf_0``, ``sum += beta_j * buf_j``); the result is rounded half to even and saturated;
  * integer scale on both axes: integer box sums; ``(sum + 2) >> 2`` for 2x2, else ``round(float32(sum) * float32(1/area))``.

PINNED: ``tests/golden/make_post_golden.py`` calls the reference's own ``scale_image`` (which calls cv2.resize) on
seeded images and asserts this restatement reproduces it bit for bit (``tests/golden/post_resize.npz``)."""
from __future__ import annotations

import math

import numpy as np


def scaled_size(h: int, w: int, sc: float):
    return int(round(h * sc)), int(round(w * sc))      # Python's round is half-to-even like cvRound


def area_table(ssize: int, dsize: int, scale: float):
    tab = []
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1, sx2 = math.ceil(fsx1), math.floor(fsx2)
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab.append((dx, sx1 - 1, np.float32((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            tab.append((dx, sx, np.float32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            tab.append((dx, sx2, np.float32(min(min(fsx2 - sx2, 1.0), cell) / cell)))
    return tab


def resize_area(img: np.ndarray, sc: float) -> np.ndarray:
    """uint8 [H,W] or [H,W,C] -> cv2.resize(img, None, fx=sc, fy=sc, interpolation=cv2.INTER_AREA), sc < 1."""
    if not 0.0 < sc < 1.0:
        raise ValueError("INTER_AREA is the reference's choice for sc < 1 only")
    sh, sw = img.shape[:2]
    dh, dw = scaled_size(sh, sw, sc)
    cn = img.shape[2] if img.ndim == 3 else 1
    src = img.reshape(sh, sw, cn)
    scale = 1.0 / sc
    iscale = int(round(scale))
    if abs(scale - iscale) < 2.220446049250313e-16:
        if dh * iscale > sh or dw * iscale > sw:
            raise NotImplementedError("partial last cell of the integer-scale path")
        s = src[:dh * iscale, :dw * iscale].astype(np.int64).reshape(dh, iscale, dw, iscale, cn).sum(axis=(1, 3))
        if iscale == 2:
            out = (s + 2) >> 2
        else:
            out = np.rint(s.astype(np.float32) * (np.float32(1.0) / np.float32(iscale * iscale)))
        res = np.clip(out, 0, 255).astype(np.uint8)
    else:
        srcf = src.astype(np.float32)
        buf = np.zeros((sh, dw, cn), np.float32)
        for dx, sx, a in area_table(sw, dw, scale):
            buf[:, dx] = buf[:, dx] + srcf[:, sx] * a          # float32 multiply, then float32 add
        out = np.zeros((dh, dw, cn), np.float32)
        started = np.zeros(dh, bool)
        for dy, sy, b in area_table(sh, dh, scale):
            if started[dy]:
                out[dy] = out[dy] + buf[sy] * b
            else:
                out[dy] = buf[sy] * b
                started[dy] = True
        res = np.clip(np.rint(out), 0, 255).astype(np.uint8)
    return res.reshape((dh, dw) + ((cn,) if img.ndim == 3 else ()))


# ---- enlarging: cv2.resize(INTER_CUBIC) (scale_image with sc > 1, helper.py:21-23) ------------------------------------
# OpenCV's 8-bit cubic path in its scalar form: fx = float32((dx + 0.5) / sc - 0.5), taps floor(fx) - 1 .. + 2 clamped to
# the image, weights with A = -0.75 in float32 stored as cvRound(w * 2048), horizontal pass, vertical pass,
# (sum + 2^21) >> 22, saturation.  NOT pinned bit for bit: OpenCV's SIMD builds evaluate the vertical pass in float and
# differ from this by one grey level on a few percent of the pixels (tests assert |restatement - cv2| <= 1).
def cubic_table(ssize: int, dsize: int, scale: float):
    idx = np.zeros((dsize, 4), np.int64)
    coef = np.zeros((dsize, 4), np.int64)
    A = np.float32(-0.75)
    one = np.float32(1.0)
    for d in range(dsize):
        fx = np.float32((d + 0.5) * scale - 0.5)
        sx = int(math.floor(float(fx)))
        fx = np.float32(fx - np.float32(sx))
        c0 = ((A * (fx + one) - np.float32(5) * A) * (fx + one) + np.float32(8) * A) * (fx + one) - np.float32(4) * A
        c1 = ((A + np.float32(2)) * fx - (A + np.float32(3))) * fx * fx + one
        c2 = ((A + np.float32(2)) * (one - fx) - (A + np.float32(3))) * (one - fx) * (one - fx) + one
        c3 = one - c0 - c1 - c2
        for k, c in enumerate((c0, c1, c2, c3)):
            idx[d, k] = min(max(sx - 1 + k, 0), ssize - 1)
            coef[d, k] = int(np.rint(np.float32(c) * np.float32(2048)))
    return idx, coef


def resize_cubic(img: np.ndarray, sc: float) -> np.ndarray:
    h, w = img.shape[:2]
    dh, dw = scaled_size(h, w, sc)
    xi, xc = cubic_table(w, dw, 1.0 / sc)
    yi, yc = cubic_table(h, dh, 1.0 / sc)
    src = img.astype(np.int64)
    if src.ndim == 2:
        src = src[:, :, None]
    hor = sum(src[:, xi[:, k], :] * xc[:, k][None, :, None] for k in range(4))
    ver = sum(hor[yi[:, k], :, :] * yc[:, k][:, None, None] for k in range(4))
    out = np.clip((ver + (1 << 21)) >> 22, 0, 255).astype(np.uint8)
    return out if img.ndim == 3 else out[:, :, 0]
