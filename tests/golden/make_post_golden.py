"""Generates tests/golden/post_*.npz - golden vectors of the integer steps around the net, produced by the REFERENCE
ITSELF: ``/root/reference`` is imported (tests/golden/_reference_import.py stubs its unrelated missing imports) and

  * ``SeparatorNetPostProcessor.post_process`` (separator_net_post_processor.py:25-99),
  * ``RegionNetPostProcessor.apply_cc_analysis`` (region_net_post_processor_base.py:230-251),
  * ``apply_threshold`` (net_post_processing_helper.py:75-78) and the colour step ``cv2.cvtColor(BGR2GRAY) / 255.0``
    (helper.py:31)

run on seeded inputs.  Before a fixture is written the numpy restatement in oracle/separator_post_oracle.py must
reproduce the reference's output bit for bit - that is what pins the oracle for this part of the path.
Masks are stored bit-packed (np.packbits).  Run from the repo root:  python tests/golden/make_post_golden.py
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from _reference_import import reference_separator_post_processor  # noqa: E402
from aru_b200.synth import synth_separator_mask  # noqa: E402
from oracle import separator_post_oracle as O  # noqa: E402

CASES = [  # (name, H, W, seed, noise)
    ("a", 150, 220, 1, 0.02),
    ("odd", 101, 177, 2, 0.05),
    ("wide", 64, 333, 3, 0.01),      # W not a multiple of 4 / 32, even structuring elements
    ("tall", 400, 100, 4, 0.03),     # smallest legal width (elements 1x1)
    ("dense", 128, 256, 5, 0.35),    # one huge noisy component
    ("cfg1", 1024, 768, 6, 0.01),    # BASELINE config 1 size: min_size 99, elements 11x1 / 1x20 / 7x1
]

if __name__ == "__main__":
    pp = reference_separator_post_processor()
    # reference helper functions (the module imports tensorflow at the top, stubbed)
    from article_separation.image_segmentation.net_post_processing.net_post_processing_helper import \
        apply_threshold as ref_apply_threshold
    for name, h, w, seed, noise in CASES:
        mask = synth_separator_mask(h, w, seed, noise)
        ref = pp.post_process(mask[:, :, None])
        ref_cc = pp.apply_cc_analysis(mask, 1 / mask.size * 100)
        hor, ver = O.separator_post_process(mask)
        assert np.array_equal(O.cc_size_filter(mask, O.cc_min_size(mask.size)), ref_cc), name
        assert np.array_equal(hor, ref["horizontal"]) and np.array_equal(ver, ref["vertical"]), name
        np.savez_compressed(os.path.join(HERE, f"post_{name}_{h}x{w}.npz"), shape=np.array([h, w]),
                            mask=np.packbits(mask > 0), cc=np.packbits(ref_cc > 0),
                            horizontal=np.packbits(ref["horizontal"] > 0), vertical=np.packbits(ref["vertical"] > 0),
                            min_size=int(mask.size * (1 / mask.size * 100)))
        print(f"post_{name}_{h}x{w}: fg {int((mask > 0).sum())} cc {int((ref_cc > 0).sum())} "
              f"hor {int((ref['horizontal'] > 0).sum())} ver {int((ref['vertical'] > 0).sum())}")

    # colour step + threshold
    rng = np.random.default_rng(7)
    bgr = rng.integers(0, 256, size=(96, 130, 3), dtype=np.uint8)
    gray = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
    grey_f = gray / 255.0                                   # helper.py:31
    assert np.array_equal(O.bgr2gray_u8(bgr), gray)
    assert np.array_equal(O.u8_to_net_input(gray), grey_f.astype(np.float32))
    u8 = rng.integers(0, 256, size=(64, 64), dtype=np.uint8)
    thr_cases = {}
    for thr in (0.05, 0.2, 0.4, 0.5, 1 / 255, 0.0):
        t = ref_apply_threshold(u8, thr)
        assert np.array_equal(O.apply_threshold(u8, thr), t)
        thr_cases[f"thr_{thr!r}"] = np.packbits(t > 0)
    np.savez_compressed(os.path.join(HERE, "post_colour_threshold.npz"), bgr=bgr, gray=gray,
                        net_input=grey_f.astype(np.float32), u8=u8, **thr_cases)
    print("post_colour_threshold: ok")

    # heading mode: network feature of a text line (heading_net_post_processor.py:247-270), by the reference itself
    from article_separation.image_segmentation.net_post_processing.heading_net_post_processor import \
        HeadingNetPostProcessor
    from python_util.geometry.polygon import Polygon

    class _SurrP:
        def __init__(self, xs, ys):
            self.xs, self.ys = xs, ys

        def to_polygon(self):
            return Polygon(list(self.xs), list(self.ys), len(self.xs))

    class _TextLine:
        def __init__(self, xs, ys):
            self.surr_p = _SurrP(xs, ys)

    hp = object.__new__(HeadingNetPostProcessor)
    rng = np.random.default_rng(11)
    u8map = rng.integers(0, 256, size=(180, 260, 2), dtype=np.uint8)
    post = hp.post_process(u8map)
    sc = 0.5                                        # image is twice the size of the net output
    polys, bboxes, probs = [], [], []
    for i in range(40):
        cx, cy = int(rng.integers(0, 520)), int(rng.integers(0, 360))
        w2, h2 = int(rng.integers(4, 200)), int(rng.integers(4, 60))
        xs = [cx - w2, cx + w2, cx + w2 + int(rng.integers(-3, 4)), cx - w2]
        ys = [cy - h2, cy - h2 + int(rng.integers(-3, 4)), cy + h2, cy + h2]
        if i % 7 == 0:                              # boxes hanging over the right / bottom border
            xs = [x + 300 for x in xs]
        xs = [max(0, x) for x in xs]                # PAGE coordinates are non-negative
        ys = [max(0, y) for y in ys]
        tl = _TextLine(xs, ys)
        prob = hp.get_net_prob_for_text_line(post, tl, sc)
        poly = tl.surr_p.to_polygon()
        poly.rescale(sc)
        bb = poly.get_bounding_box()
        if bb.width * bb.height == 0 or not np.isfinite(prob):
            continue
        mine = O.net_prob_for_box(u8map, bb.x, bb.y, bb.width, bb.height)
        assert abs(mine - prob) <= 1e-12 * max(1.0, abs(prob)), (i, mine, prob)
        polys.append(xs + ys)
        bboxes.append([bb.x, bb.y, bb.width, bb.height])
        probs.append(prob)
    np.savez_compressed(os.path.join(HERE, "post_heading_lines.npz"), u8=u8map, scale=sc, polygons=np.array(polys),
                        bboxes=np.array(bboxes), probs=np.array(probs, np.float64))
    print("post_heading_lines:", len(probs), "text lines")

    # scale_image (helper.py:14-25), shrinking case, by the reference itself (it calls cv2.resize INTER_AREA)
    from article_separation.image_segmentation.net_post_processing.net_post_processing_helper import \
        scale_image as ref_scale_image
    from oracle import resize_oracle as R
    rng = np.random.default_rng(21)
    cases = {}
    for name, (sh, sw, ch, fixed_height) in {"general_bgr": (173, 131, 3, 96), "general_gray": (211, 160, 1, 150),
                                             "half_bgr": (192, 130, 3, 96), "third_bgr": (180, 123, 3, 60),
                                             "slight_bgr": (101, 97, 3, 100)}.items():
        img = rng.integers(0, 256, size=(sh, sw, ch) if ch == 3 else (sh, sw), dtype=np.uint8)
        scaled, sc = ref_scale_image(img, fixed_height=fixed_height, scaling_factor=1.0)
        assert sc < 1.0
        mine = R.resize_area(img, sc)
        assert mine.shape == scaled.shape and np.array_equal(mine, scaled), name
        cases[name + "_src"] = img
        cases[name + "_dst"] = scaled
        cases[name + "_sc"] = np.float64(sc)
    np.savez_compressed(os.path.join(HERE, "post_resize.npz"), **cases)
    print("post_resize:", len(cases) // 3, "cases")

    # SWT feature image (python_util/image_processing/swt_dist_trafo.py:18-29) by the reference's own class: it reads the
    # file itself, so every page goes through a temporary PNG
    import tempfile
    from python_util.image_processing.swt_dist_trafo import StrokeWidthDistanceTransform
    from aru_b200.synth import synth_page
    from oracle import swt_oracle as S
    swt = StrokeWidthDistanceTransform(dark_on_bright=True)
    rng = np.random.default_rng(31)
    pages = {"page": synth_page(300, 220, 1), "odd": synth_page(257, 131, 2),
             "noise": rng.integers(0, 256, size=(90, 140), dtype=np.uint8)}
    blob = synth_page(400, 300, 3)
    blob[50:350, 40:260] = 20                         # a photograph-sized dark block: distances up to 110
    pages["blob"] = blob
    wrap = np.full((700, 600), 230, np.uint8)
    wrap[20:680, 10:590] = 10                         # distances beyond 255: the uint8 cast wraps
    pages["wrap"] = wrap
    pages["flat"] = np.full((40, 50), 17, np.uint8)   # no pixel at or below the Otsu threshold: all zeros
    out = {}
    for name, gray in pages.items():
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "page.png")
            cv2.imwrite(path, gray)
            with np.errstate(invalid="ignore"):
                ref = swt.distance_transform(path)
        mine, thr, _ = S.swt_distance_transform(gray)
        assert ref.dtype == np.uint8 and np.array_equal(mine, ref), name
        out[name + "_gray"], out[name + "_dt"], out[name + "_thr"] = gray, ref, np.int32(thr)
        print(f"post_swt {name}: {gray.shape} otsu {thr} max distance {int(ref.max())}")
    np.savez_compressed(os.path.join(HERE, "post_swt.npz"), **out)
