"""CPU ORACLE (test infrastructure, never shipped): numpy restatement of the SWT feature image
``StrokeWidthDistanceTransform.distance_transform`` (python_util/image_processing/swt_dist_trafo.py:18-29), which
``HeadingNetPostProcessor`` computes twice per page at full resolution (heading_net_post_processor.py:86,297):

    gray uint8 -> 255 - gray (dark on bright) -> cv2.GaussianBlur 5x5, sigma 0 -> cv2.threshold(OTSU) ->
    cv2.distanceTransform(DIST_L2, DIST_MASK_PRECISE) -> .astype(np.uint8)

PINNED: tests/golden/make_post_golden.py runs the reference's own class (cv2 is available in the build container) on
seeded pages and asserts this module reproduces it bit for bit before tests/golden/post_swt_*.npz are written.

OpenCV rules restated here:
  * GaussianBlur on 8-bit images with ksize 5 and sigma <= 0 uses the fixed kernel [1, 4, 6, 4, 1] / 16 in fixed-point
    arithmetic, BORDER_REFLECT_101, one rounding at the end: (sum_ij k_i k_j p_ij + 128) >> 8;
  * Otsu: getThreshVal_Otsu_8u - 256-bin histogram, double arithmetic, first maximum of the between-class variance,
    bins whose class probability is < FLT_EPSILON skipped; THRESH_BINARY keeps pixels > threshold;
  * DIST_MASK_PRECISE: exact Euclidean distance to the nearest zero pixel (squared distance as an integer, float32 sqrt);
  * astype(uint8) truncates toward zero (distances >= 256 wrap modulo 256 like numpy's C cast).
"""
from __future__ import annotations

import numpy as np

FLT_EPSILON = 1.1920928955078125e-07


def gaussian5_u8(img: np.ndarray) -> np.ndarray:
    k = np.array([1, 4, 6, 4, 1], np.int64)
    p = np.pad(img.astype(np.int64), 2, mode="reflect")          # numpy "reflect" == BORDER_REFLECT_101
    h, w = img.shape
    tmp = sum(k[j] * p[:, j:j + w] for j in range(5))
    out = sum(k[i] * tmp[i:i + h, :] for i in range(5))
    return ((out + 128) >> 8).astype(np.uint8)


def otsu_threshold(img: np.ndarray) -> int:
    hist = np.bincount(img.ravel(), minlength=256).astype(np.float64)
    scale = 1.0 / img.size
    mu = float((np.arange(256) * hist).sum()) * scale
    mu1 = q1 = 0.0
    max_sigma, max_val = 0.0, 0
    for i in range(256):
        p_i = hist[i] * scale
        mu1 *= q1
        q1 += p_i
        q2 = 1.0 - q1
        if min(q1, q2) < FLT_EPSILON or max(q1, q2) > 1.0 - FLT_EPSILON:
            continue
        mu1 = (mu1 + i * p_i) / q1
        mu2 = (mu - q1 * mu1) / q2
        sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2)
        if sigma > max_sigma:
            max_sigma, max_val = sigma, i
    return max_val


def squared_edt(binary: np.ndarray) -> np.ndarray:
    """Exact squared Euclidean distance of every non-zero pixel to the nearest zero pixel (int64); pixels of an image
    without any zero pixel get a huge value."""
    h, w = binary.shape
    inf = (h + w) * 2
    g = np.where(binary != 0, inf, 0).astype(np.int64)          # vertical distance to the nearest zero in the column
    for y in range(1, h):
        g[y] = np.minimum(g[y], g[y - 1] + 1)
    for y in range(h - 2, -1, -1):
        g[y] = np.minimum(g[y], g[y + 1] + 1)
    g2 = g * g
    out = np.empty((h, w), np.int64)
    xs = np.arange(w)
    dx2 = (xs[:, None] - xs[None, :]) ** 2                        # [x, x']
    for y in range(h):
        out[y] = (dx2 + g2[y][None, :]).min(axis=1)
    return out


def swt_distance_transform(gray: np.ndarray, dark_on_bright: bool = True) -> np.ndarray:
    img = np.asarray(gray, np.uint8)
    if dark_on_bright:
        img = (255 - img.astype(np.int32)).astype(np.uint8)      # uint8 "-image + 255"
    blur = gaussian5_u8(img)
    thr = otsu_threshold(blur)
    binary = np.where(blur > thr, 255, 0).astype(np.uint8)
    if not (binary == 0).any():
        # no zero pixel: OpenCV's distances are infinite and the uint8 cast of the reference turns them into 0
        return np.zeros(img.shape, np.uint8), thr, blur
    d = np.sqrt(squared_edt(binary).astype(np.float32)).astype(np.float32)
    return (d.astype(np.int64) & 0xFF).astype(np.uint8), thr, blur
