"""GPU parity of the integer pre / post-processing kernels (csrc/post.cu) through
the C ABI: bit-exact against the golden vectors produced by the reference's own post_process, against the numpy
oracle on seeded masks (edge cases: widths that are not multiples of 4 / 32, even and odd structuring elements,
border-touching runs, empty and full masks, smallest legal page), and at the benchmark page size."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
POST_FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "post_*x*.npz")))


@pytest.fixture(scope="module")
def eng(built_lib):
    from aru_b200.engine import Engine
    from aru_b200.synth import synth_pb
    e = Engine(synth_pb("separator"), device=0)
    yield e
    e.close()


def _load(path):
    from test_post_oracle import load_post_fixture
    return load_post_fixture(path)


@pytest.mark.parametrize("path", POST_FIXTURES, ids=[os.path.basename(p)[:-4] for p in POST_FIXTURES])
def test_separator_post_matches_reference_golden(eng, path):
    f = _load(path)
    hor, ver = eng.separator_post(f["mask"])
    assert np.array_equal(ver, f["vertical"])
    assert np.array_equal(hor, f["horizontal"])


def test_separator_post_batch_of_different_masks(eng):
    from aru_b200.synth import synth_separator_mask
    from oracle import separator_post_oracle as O
    masks = np.stack([synth_separator_mask(203, 311, seed=s, noise=0.01 * (s + 1)) for s in range(5)])
    hor, ver = eng.separator_post(masks)
    for i in range(len(masks)):
        h_ref, v_ref = O.separator_post_process(masks[i])
        assert np.array_equal(hor[i], h_ref) and np.array_equal(ver[i], v_ref), i


@pytest.mark.parametrize("h,w", [(50, 100), (57, 129), (300, 128), (333, 1000), (1500, 1125)])
def test_separator_post_shapes(eng, h, w):
    from aru_b200.synth import synth_separator_mask
    from oracle import separator_post_oracle as O
    m = synth_separator_mask(h, w, seed=h + w)
    hor, ver = eng.separator_post(m)
    h_ref, v_ref = O.separator_post_process(m)
    assert np.array_equal(hor, h_ref) and np.array_equal(ver, v_ref)


def test_separator_post_degenerate_masks(eng):
    from oracle import separator_post_oracle as O
    h, w = 120, 260
    cases = [np.zeros((h, w), np.uint8), np.full((h, w), 255, np.uint8)]
    chk = np.zeros((h, w), np.uint8)
    chk[::2, ::2] = 255
    chk[1::2, 1::2] = 255                       # checkerboard: one 8-connected component, no 4-connected neighbours
    cases.append(chk)
    snake = np.zeros((h, w), np.uint8)          # serpentine: long union-find chains
    snake[::4, :] = 255
    snake[2::8, -1] = 255
    snake[1::8, -1] = 255
    snake[3::8, -1] = 255
    snake[5::8, 0] = 255
    snake[6::8, 0] = 255
    snake[7::8, 0] = 255
    cases.append(snake)
    lines = np.zeros((h, w), np.uint8)          # long thin vertical / diagonal lines
    lines[:, 7] = 255
    lines[:, w - 1] = 255
    for i in range(min(h, w)):
        lines[i, i] = 255
    cases.append(lines)
    for i, m in enumerate(cases):
        hor, ver = eng.separator_post(m)
        h_ref, v_ref = O.separator_post_process(m)
        assert np.array_equal(hor, h_ref) and np.array_equal(ver, v_ref), i


def test_separator_post_rejects_pages_opencv_rejects(eng):
    from aru_b200.engine import EngineError
    with pytest.raises(EngineError):
        eng.separator_post(np.zeros((49, 200), np.uint8))
    with pytest.raises(EngineError):
        eng.separator_post(np.zeros((200, 66), np.uint8))


def test_open_rect_against_oracle(eng):
    from oracle import separator_post_oracle as O
    rng = np.random.default_rng(3)
    for t in range(40):
        h, w = int(rng.integers(1, 90)), int(rng.integers(1, 150))
        img = ((rng.random((h, w)) < rng.uniform(0.5, 0.97)) * 255).astype(np.uint8)
        k = int(rng.integers(1, 140 if t % 5 == 0 else 45))
        kw, kh = (k, 1) if t % 2 == 0 else (1, k)
        got = eng.open_rect(img, kw, kh)
        assert np.array_equal(got, O.open_rect(img, kw, kh)), (h, w, kw, kh)


def test_pages_to_input_matches_reference_colour_step(eng):
    from oracle import separator_post_oracle as O
    z = np.load(os.path.join(GOLDEN, "post_colour_threshold.npz"))
    got = eng.pages_to_input(z["bgr"])
    assert got.dtype == np.float32 and np.array_equal(got[0], z["net_input"])
    got = eng.pages_to_input(z["gray"])
    assert np.array_equal(got[0], z["net_input"])
    ramp = np.arange(256, dtype=np.uint8).reshape(16, 16)
    assert np.array_equal(eng.pages_to_input(ramp)[0], O.u8_to_net_input(ramp))
    rng = np.random.default_rng(5)
    bgr = rng.integers(0, 256, size=(3, 40, 52, 3), dtype=np.uint8)
    assert np.array_equal(eng.pages_to_input(bgr), O.u8_to_net_input(O.bgr2gray_u8(bgr)))


@pytest.mark.parametrize("channels", [1, 3])
def test_separator_pages_is_run_up_to_polygons(eng, channels):
    """aru_separator_pages == colour step -> net -> uint8 -> threshold -> post_process, every integer step bit-exact
    given the engine's own probability map (the map itself is checked against the oracle in test_gpu_parity.py)."""
    from aru_b200.synth import synth_page
    from oracle import separator_post_oracle as O
    n, h, w = 5, 256, 192
    gray = np.stack([synth_page(h, w, seed=10 + i) for i in range(n)])
    pages = gray if channels == 1 else np.stack([gray, gray, gray], axis=-1)   # B=G=R: luma == gray
    for thr in (0.05, 0.2):
        r = eng.separator_pages(pages, threshold=thr, want_prob=True, want_u8=True, want_mask=True)
        prob = eng.forward(O.u8_to_net_input(gray))
        assert np.array_equal(r["prob"], prob)                        # same kernels, same input bits
        assert np.array_equal(r["u8"], O.quantize_u8(prob))
        assert np.array_equal(r["mask"], O.apply_threshold(r["u8"][..., 0], thr))
        for i in range(n):
            h_ref, v_ref = O.separator_post_process(r["mask"][i])
            assert np.array_equal(r["horizontal"][i], h_ref) and np.array_equal(r["vertical"][i], v_ref)


def test_separator_pages_benchmark_size_properties(eng):
    """BASELINE config 3 page size (1856x1344), micro-batched: oracle comparison on one page, and the size-independent
    properties on all: idempotence of the (odd) openings, horizontal AND vertical == 0, containment in the mask."""
    from aru_b200.synth import synth_page
    from oracle import separator_post_oracle as O
    n, h, w = 6, 1856, 1344
    pages = np.stack([synth_page(h, w, seed=100 + i) for i in range(n)])
    r = eng.separator_pages(pages, want_mask=True)
    hor, ver, mask = r["horizontal"], r["vertical"], r["mask"]
    assert set(np.unique(hor)) <= {0, 255} and set(np.unique(ver)) <= {0, 255}
    assert not np.any(hor & ver)
    k1, kv, k2 = O.separator_kernel_sizes(h, w)
    assert not np.any(ver & ~mask)
    # the 20-wide (even) element makes OpenCV's opening the true opening shifted by one pixel to the right, so the
    # horizontal mask may leave the thresholded mask by exactly that pixel
    right = np.zeros_like(mask)
    right[:, :, 1:] = mask[:, :, :-1]
    assert k1 % 2 == 0 and not np.any(hor & ~(mask | right))
    assert kv % 2 == 1 and k2 % 2 == 1
    assert np.array_equal(eng.open_rect(ver, 1, kv), ver)             # opening is idempotent
    assert np.array_equal(eng.open_rect(hor, k2, 1), hor)
    h_ref, v_ref = O.separator_post_process(mask[0])
    assert np.array_equal(hor[0], h_ref) and np.array_equal(ver[0], v_ref)
    h2, v2 = eng.separator_post(np.ascontiguousarray(mask[:2]))
    assert np.array_equal(h2, hor[:2]) and np.array_equal(v2, ver[:2])
