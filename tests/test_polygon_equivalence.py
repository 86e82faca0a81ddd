"""Output equivalence one level above the masks (BASELINE north_star: identical separator polylines / heading flags):
the ENGINE's uint8 maps for 100 synthetic pages (tests/golden/equiv/engine_r02.npz, written on a B200 by
tools/equivalence_gpu.py) against the fp32 oracle's maps of the same pages, both pushed through the reference's own code
where polygon extraction starts:

    apply_threshold (helper.py:75-78) -> SeparatorNetPostProcessor.post_process (sep:25-99) -> to_polygons (sep:99-115,
    apply_contour_detection2 base:186-197, rasterio.features.shapes replaced by tests/shims/rasterio_features_shim.py) ->
    rescale_polygons (base:253-268)

and, for the heading net, get_net_prob_for_text_line's value per text line (head:247-270) -> the heading decision of
HeadingNetPostProcessor.to_page_xml (head:153-177, restated below: it is entangled with lxml page objects) with the CLI's
weights / thresholds (run_net_post_processing.py:15-23).

What is asserted is what was MEASURED, and why it is not 99 %: the synthetic nets have random weights, their maps are
smooth around the decision boundary (~1.1 % of all pixels sit within one uint8 step of the 13/255 cut), and the engine's
16-bit activations move ~12 pixels per 49 152-pixel page across it - always pixels whose oracle value is within two steps of the cut.
Scaling the classifier ("sharp" nets, 4x the logit margin) does not change that count (the error scales with the gain);
only a trained net, whose features are far from the boundary, does.  So: identical thresholded masks on 0 % of pages,
identical polygon sets on 55 - 63 %, >= 99.97 % identical mask pixels, identical heading flags on 97 - 98 % of pages
(2 - 3 of 3 696 text lines, each within 1e-3 of the 0.4 threshold)."""
import os
import sys
import types
from collections import Counter

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(0, os.path.join(HERE, "golden"))

from aru_b200.synth import page_to_net_input, synth_page  # noqa: E402
from oracle import separator_post_oracle as post_oracle  # noqa: E402
from oracle.aru_oracle import Oracle  # noqa: E402

FIXTURE = os.path.join(HERE, "golden", "equiv", "engine_r02.npz")
HAVE_REFERENCE = os.path.isdir("/root/reference/article_separation")
SC = 0.5      # the page was shrunk by 0.5 for the net; polygons go back to image coordinates with 1 / sc


@pytest.fixture(scope="module")
def fx():
    from equivalence_gpu import H, N, SEED0, W, net_pb
    z = np.load(FIXTURE)
    assert int(z["n"]) == N and tuple(z["shape"]) == (H, W) and int(z["pages_seed0"]) == SEED0
    pages = [synth_page(H, W, SEED0 + i) for i in range(N)]
    return z, pages, net_pb


@pytest.fixture(scope="module")
def ref_sep():
    """(apply_threshold, post_process, to_polygons, rescale_polygons): the reference's real functions in the build
    container (rasterio -> shim, its other missing imports -> inert stubs), restatements elsewhere."""
    import rasterio_features_shim as shim
    if HAVE_REFERENCE:
        import _reference_import as R
        ras, feat = types.ModuleType("rasterio"), types.ModuleType("rasterio.features")
        feat.shapes = shim.shapes
        ras.features = feat
        saved = {k: sys.modules.get(k) for k in ("rasterio", "rasterio.features")}
        sys.modules["rasterio"], sys.modules["rasterio.features"] = ras, feat
        pp = R.reference_separator_post_processor()
        base = sys.modules["article_separation.image_segmentation.net_post_processing.region_net_post_processor_base"]
        base.rasterio = ras                      # the module may have been imported earlier with the inert stub
        from article_separation.image_segmentation.net_post_processing.net_post_processing_helper import apply_threshold
        yield apply_threshold, pp.post_process, pp.to_polygons, pp.rescale_polygons, True
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        return

    def post_process(mask):
        hor, ver = post_oracle.separator_post_process(mask[:, :, 0])
        return {"horizontal": hor, "vertical": ver}

    def to_polygons(img, kind):
        return {"SeparatorRegion_" + kind: [p[0]["coordinates"] for p in shim.shapes(img, connectivity=8) if p[1] == 255]}

    def rescale(d, s):
        return {k: [[[(int(x * s), int(y * s)) for x, y in ring] for ring in poly] for poly in v] for k, v in d.items()}

    yield (lambda u8, t: post_oracle.apply_threshold(u8, t)), post_process, to_polygons, rescale, False


def test_shim_rings():
    from rasterio_features_shim import shapes
    m = np.zeros((6, 7), np.uint8)
    m[1, 1] = m[2, 2] = 255                                    # a diagonal pair: one region under 8-connectivity
    got8 = [g["coordinates"] for g, v in shapes(m, connectivity=8) if v == 255]
    got4 = [g["coordinates"] for g, v in shapes(m, connectivity=4) if v == 255]
    assert len(got8) == 1 and len(got4) == 2
    assert got8[0][0][0] == got8[0][0][-1] and (2.0, 2.0) in got8[0][0]
    ring = np.zeros((7, 7), np.uint8)
    ring[1:6, 1:6] = 255
    ring[3, 3] = 0
    (coords,) = [g["coordinates"] for g, v in shapes(ring, connectivity=8) if v == 255]
    assert len(coords) == 2 and coords[0] == [(1.0, 1.0), (6.0, 1.0), (6.0, 6.0), (1.0, 6.0), (1.0, 1.0)]
    assert sorted(coords[1][:-1]) == [(3.0, 3.0), (3.0, 4.0), (4.0, 3.0), (4.0, 4.0)]
    # areas: the shoelace area of exterior minus holes equals the pixel count of every region
    rng = np.random.default_rng(0)
    img = np.where(rng.random((40, 50)) < 0.45, 255, 0).astype(np.uint8)
    total = 0.0
    for g, v in shapes(img, connectivity=8):
        if v != 255:
            continue
        a = [abs(sum(p[0] * q[1] - q[0] * p[1] for p, q in zip(r[:-1], r[1:]))) / 2 for r in g["coordinates"]]
        total += a[0] - sum(a[1:])
    assert total == float((img == 255).sum())


@pytest.mark.parametrize("net", ["separator", "separator_sharp"])
def test_separator_polygons_engine_vs_oracle(fx, ref_sep, net):
    z, pages, net_pb = fx
    apply_threshold, post_process, to_polygons, rescale_polygons, real = ref_sep
    orc = Oracle(net_pb(net))
    n = len(pages)
    shape = pages[0].shape

    def polygons(u8):
        post = post_process(apply_threshold(u8[:, :, None].copy(), 0.05))
        d = {}
        for kind, img in post.items():
            d.update(to_polygons(img, kind))
        return rescale_polygons(d, 1 / SC), post

    same_poly = flips = px = 0
    for i in range(n):
        ref_u8 = (orc.run(page_to_net_input(pages[i]))[0][..., 0] * 255).astype(np.uint8)
        eng_u8 = z[net + "_u8"][i]
        # BASELINE's tolerance 2e-2 is 5 uint8 steps.  The calibrated net stays within 2; the "sharp" variant multiplies
        # the classifier (and with it every upstream rounding error) by 4 and reaches 6 steps at isolated pixels
        steps = 2 if net == "separator" else 7
        assert np.abs(ref_u8.astype(int) - eng_u8.astype(int)).max() <= steps
        flipped = (ref_u8 >= 13) != (eng_u8 >= 13)
        assert (np.abs(ref_u8[flipped].astype(int) - 12.5) <= steps).all()                   # only pixels on the cut move
        flips += int(flipped.sum())
        px += flipped.size
        pr, _ = polygons(ref_u8)
        pe, post_e = polygons(eng_u8)
        same_poly += pr == pe
        # the device post-processing of the engine's own mask (from the GPU run) == the reference's post_process of it
        hor = np.unpackbits(z[net + "_hor"])[:n * shape[0] * shape[1]].reshape((n,) + shape)[i] * 255
        ver = np.unpackbits(z[net + "_ver"])[:n * shape[0] * shape[1]].reshape((n,) + shape)[i] * 255
        assert np.array_equal(post_e["horizontal"], hor) and np.array_equal(post_e["vertical"], ver), i
    print(f"{net}: identical polygon sets on {same_poly}/{n} pages, {flips / n:.1f} flipped mask pixels per page "
          f"({1 - flips / px:.5%} identical), reference code: {real}")
    assert 1 - flips / px >= 0.9995
    assert same_poly >= 0.45 * n          # measured 55 / 63 of 100 (see the module docstring for why not 99)


def _scale(v, lo, hi):
    return v if hi - lo == 0 else (v - lo) / (hi - lo)          # scale_to_new_interval, head:48-60


def heading_flags(net_conf, stroke_width, text_height):
    """The decision of HeadingNetPostProcessor.to_page_xml (head:103-177) for one page, CLI settings
    (run_net_post_processing.py:15-23): weights net 0.8 / stroke width 0 / text height 0.2, early-outs at 1.0 / 1.0 / 0.9 /
    0.9, heading if the confidence exceeds 0.4."""
    sw_mode = Counter(stroke_width).most_common(1)[0][0]
    th_mode = Counter(text_height).most_common(1)[0][0]
    swd = [s - sw_mode for s in stroke_width]
    thd = [t - th_mode for t in text_height]
    out = []
    for c, s, t in zip(net_conf, swd, thd):
        sc, tc = _scale(s, min(swd), max(swd)), _scale(t, min(thd), max(thd))
        if sc >= 1.0 or tc >= 0.9 or (sc + tc) / 2 >= 0.9 or c >= 1.0:
            conf = 1.0
        else:
            conf = 0.8 * c + 0.0 * sc + 0.2 * tc
        out.append(bool(conf > 0.4))
    return out


@pytest.mark.parametrize("net", ["heading", "heading_sharp"])
def test_heading_flags_engine_vs_oracle(fx, net):
    z, pages, net_pb = fx
    orc = Oracle(net_pb(net))
    boxes = z["boxes"]
    same = lines = differing = 0
    for i, page in enumerate(pages):
        ref_u8 = (orc.run(page_to_net_input(page))[0][..., 0] * 255).astype(np.uint8)
        sel = boxes[:, 0] == i
        bx = boxes[sel]
        area = (bx[:, 2] - bx[:, 1]) * (bx[:, 4] - bx[:, 3])
        # get_net_prob_for_text_line (head:247-270): mean of net_output / 255 over the text line's bounding box
        ref_conf = np.array([ref_u8[b[1]:b[2], b[3]:b[4]].astype(np.uint64).sum() for b in bx]) / 255 / area
        eng_conf = z[net + "_sums"][sel] / 255 / area
        assert np.abs(ref_conf - eng_conf).max() <= 2e-3
        rng = np.random.default_rng(7000 + i)                 # the SWT features are the same on both sides
        sw = [int(v) for v in rng.integers(2, 7, len(bx))]
        th = [int(v) for v in rng.integers(8, 15, len(bx))]
        fr, fe = heading_flags(ref_conf, sw, th), heading_flags(eng_conf, sw, th)
        same += fr == fe
        lines += len(bx)
        differing += sum(a != b for a, b in zip(fr, fe))
    print(f"{net}: identical heading flags on {same}/{len(pages)} pages, {differing} of {lines} text lines differ")
    assert same >= 0.95 * len(pages) and differing <= 0.002 * lines
