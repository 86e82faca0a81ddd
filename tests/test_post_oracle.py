"""CPU tests of oracle/separator_post_oracle.py against the golden vectors the REFERENCE ITSELF produced
(tests/golden/make_post_golden.py ran the reference's SeparatorNetPostProcessor.post_process / apply_cc_analysis /
apply_threshold and cv2.cvtColor in the build container) and against OpenCV where it is importable."""
import glob
import os

import numpy as np
import pytest

from oracle import separator_post_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
POST_FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "post_*_[0-9]*x[0-9]*.npz")))


def load_post_fixture(path):
    z = np.load(path)
    h, w = (int(v) for v in z["shape"])
    unpack = lambda k: np.unpackbits(z[k])[:h * w].reshape(h, w) * np.uint8(255)  # noqa: E731
    return {"h": h, "w": w, "mask": unpack("mask"), "cc": unpack("cc"), "horizontal": unpack("horizontal"),
            "vertical": unpack("vertical"), "min_size": int(z["min_size"])}


def test_fixtures_present():
    assert len(POST_FIXTURES) >= 6


@pytest.mark.parametrize("path", POST_FIXTURES, ids=[os.path.basename(p)[:-4] for p in POST_FIXTURES])
def test_oracle_reproduces_reference_post_process(path):
    f = load_post_fixture(path)
    assert O.cc_min_size(f["h"] * f["w"]) == f["min_size"]
    assert np.array_equal(O.cc_size_filter(f["mask"], f["min_size"]), f["cc"])
    hor, ver = O.separator_post_process(f["mask"])
    assert np.array_equal(hor, f["horizontal"])
    assert np.array_equal(ver, f["vertical"])
    hor3, ver3 = O.separator_post_process(np.stack([f["mask"], 255 - f["mask"]], axis=-1))   # HWC: channel 0 only
    assert np.array_equal(hor3, hor) and np.array_equal(ver3, ver)


def test_min_size_quirk():
    # int(size * (1 / size * 100)) is 99 for some sizes, e.g. BASELINE config 1
    assert O.cc_min_size(1024 * 768) == 99
    assert O.cc_min_size(1856 * 1344) == 100
    assert O.cc_min_size(1500 * 1125) == 100


def test_kernel_sizes():
    assert O.separator_kernel_sizes(1024, 768) == (11, 20, 7)
    assert O.separator_kernel_sizes(1500, 1125) == (16, 30, 11)
    assert O.separator_kernel_sizes(1856, 1344) == (20, 37, 13)
    with pytest.raises(ValueError):
        O.separator_post_process(np.zeros((49, 200), np.uint8))      # 1 x 0 element: OpenCV raises in the reference


def test_colour_step_and_threshold_fixture():
    z = np.load(os.path.join(GOLDEN, "post_colour_threshold.npz"))
    assert np.array_equal(O.bgr2gray_u8(z["bgr"]), z["gray"])
    assert np.array_equal(O.u8_to_net_input(z["gray"]), z["net_input"])
    u8 = z["u8"]
    for key in z.files:
        if key.startswith("thr_"):
            thr = float(key[4:])
            want = np.unpackbits(z[key])[:u8.size].reshape(u8.shape) * np.uint8(255)
            assert np.array_equal(O.apply_threshold(u8, thr), want), key


def test_u8_to_float_is_a_plain_float32_division():
    # float32(u8 / 255.0 computed in float64) == float32(u8) / float32(255) for every uint8: the device kernel divides in fp32
    u = np.arange(256, dtype=np.uint8)
    assert np.array_equal((u / 255.0).astype(np.float32), u.astype(np.float32) / np.float32(255))


def test_quantize_truncates():
    assert O.quantize_u8(np.array([1.0], np.float32))[0] == 255
    assert O.quantize_u8(np.array([0.9999], np.float32))[0] == 254


def test_morphology_against_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for _ in range(60):
        h, w = int(rng.integers(1, 50)), int(rng.integers(1, 70))
        img = ((rng.random((h, w)) < rng.uniform(0.4, 0.95)) * 255).astype(np.uint8)
        k = int(rng.integers(1, 40))
        kw, kh = (k, 1) if rng.random() < 0.5 else (1, k)
        el = cv2.getStructuringElement(cv2.MORPH_RECT, (kw, kh))
        assert np.array_equal(cv2.erode(img, el), O.erode_rect(img, kw, kh))
        assert np.array_equal(cv2.dilate(img, el), O.dilate_rect(img, kw, kh))
        assert np.array_equal(cv2.morphologyEx(img, cv2.MORPH_OPEN, el), O.open_rect(img, kw, kh))


def test_cc_filter_against_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    for _ in range(20):
        h, w = int(rng.integers(5, 90)), int(rng.integers(5, 120))
        img = ((rng.random((h, w)) < rng.uniform(0.2, 0.6)) * 255).astype(np.uint8)
        min_size = int(rng.integers(1, 60))
        n, lab, stats, _ = cv2.connectedComponentsWithStats(img, connectivity=8)
        want = np.zeros_like(img)
        for i in range(1, n):
            if stats[i, -1] >= min_size:
                want[lab == i] = 255
        assert np.array_equal(O.cc_size_filter(img, min_size), want)


def test_heading_box_feature_fixture():
    """get_net_prob_for_text_line values computed by the reference itself (40 text lines, scaling factor 0.5)."""
    z = np.load(os.path.join(GOLDEN, "post_heading_lines.npz"))
    u8 = z["u8"]
    assert len(z["probs"]) >= 30
    for (x, y, w, h), prob in zip(z["bboxes"], z["probs"]):
        mine = O.net_prob_for_box(u8, int(x), int(y), int(w), int(h))
        assert abs(mine - prob) <= 1e-12 * max(1.0, abs(prob))
    # numpy slice semantics of the box: clipped at the page border, negative start counts from the end
    assert O.box_sum_u8(u8, 170, 400, 250, 999) == int(u8[170:, 250:, 0].astype(np.int64).sum())
    assert O.box_sum_u8(u8, -5, 180, 0, 3) == int(u8[175:180, 0:3, 0].astype(np.int64).sum())
    assert O.box_sum_u8(u8, 50, 40, 0, 10) == 0


def test_resize_oracle_reproduces_reference_scale_image():
    """oracle/resize_oracle.py against the outputs of the reference's own scale_image (cv2.resize INTER_AREA)."""
    from oracle import resize_oracle as R
    z = np.load(os.path.join(GOLDEN, "post_resize.npz"))
    names = sorted(k[:-4] for k in z.files if k.endswith("_src"))
    assert len(names) >= 5
    for name in names:
        got = R.resize_area(z[name + "_src"], float(z[name + "_sc"]))
        assert np.array_equal(got, z[name + "_dst"]), name


def test_resize_oracle_against_opencv():
    cv2 = pytest.importorskip("cv2")
    from oracle import resize_oracle as R
    rng = np.random.default_rng(4)
    for sh, sw, sc in ((97, 131, 0.73), (200, 150, 0.375), (120, 90, 0.5), (150, 93, 1 / 3), (64, 64, 0.25), (333, 77, 0.9017)):
        img = rng.integers(0, 256, size=(sh, sw, 3), dtype=np.uint8)
        assert np.array_equal(R.resize_area(img, sc), cv2.resize(img, None, fx=sc, fy=sc, interpolation=cv2.INTER_AREA)), (sh, sw, sc)


def _swt_cv2(gray):
    """The reference's distance_transform after its cv2.imread (swt_dist_trafo.py:19-24), same OpenCV calls."""
    import cv2
    image = -gray + 255
    # OpenCV 4.13's multi-threaded GaussianBlur is not deterministic on images of a few rows (8 x 40 with 8 threads gives
    # two different results run to run); single-threaded it is, and that is what the restatement reproduces
    threads = cv2.getNumThreads()
    cv2.setNumThreads(1)
    try:
        blur = cv2.GaussianBlur(image, (5, 5), 0)
    finally:
        cv2.setNumThreads(threads)
    _, image_t = cv2.threshold(blur, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
    with np.errstate(invalid="ignore"):
        return cv2.distanceTransform(image_t, cv2.DIST_L2, cv2.DIST_MASK_PRECISE).astype(np.uint8)


def test_swt_oracle_reproduces_reference_fixture():
    """tests/golden/post_swt.npz was written by the reference's own StrokeWidthDistanceTransform.distance_transform."""
    from oracle import swt_oracle as S
    z = np.load(os.path.join(GOLDEN, "post_swt.npz"))
    names = sorted(k[:-5] for k in z.files if k.endswith("_gray"))
    assert len(names) >= 6
    for name in names:
        dt, thr, _ = S.swt_distance_transform(z[name + "_gray"])
        assert thr == int(z[name + "_thr"]) and np.array_equal(dt, z[name + "_dt"]), name


def test_swt_oracle_against_opencv():
    from aru_b200.synth import synth_page
    from oracle import swt_oracle as S
    rng = np.random.default_rng(5)
    for gray in (synth_page(120, 97, 8), rng.integers(0, 256, size=(33, 61), dtype=np.uint8),
                 np.where(rng.random((70, 90)) < 0.5, 200, 30).astype(np.uint8), synth_page(64, 5, 2)):
        assert np.array_equal(S.swt_distance_transform(gray)[0], _swt_cv2(gray))


def test_cubic_restatement_is_within_one_grey_level_of_opencv():
    """scale_image for sc > 1 (helper.py:21-23) calls cv2.resize(INTER_CUBIC).  The fixed-point restatement the device
    kernel follows reproduces OpenCV's destination size exactly and its values to one grey level (OpenCV's SIMD vertical
    pass works in float) - the stated tolerance of the enlarging path."""
    import cv2
    from oracle import resize_oracle as R
    rng = np.random.default_rng(3)
    for shape, sc in (((120, 90, 3), 1.125), ((97, 131), 1.5), ((64, 80, 3), 2.0), ((50, 60), 3.3), ((31, 17, 3), 900 / 800)):
        img = rng.integers(0, 256, size=shape, dtype=np.uint8)
        ref = cv2.resize(img, None, fx=sc, fy=sc, interpolation=cv2.INTER_CUBIC)
        mine = R.resize_cubic(img, sc)
        assert mine.shape == ref.shape
        d = np.abs(mine.astype(int) - ref.astype(int))
        assert d.max() <= 1 and (d > 0).mean() < 0.15, (shape, sc, int(d.max()), float((d > 0).mean()))
