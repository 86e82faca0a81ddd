"""The CPU oracle: TF kernel semantics as known-answer tests, agreement with OpenCV's TF importer,
and the committed golden fixtures (no GPU)."""
import glob
import hashlib
import os

import numpy as np
import pytest
import torch

from aru_b200.graphdef import GraphBuilder
from aru_b200.synth import synth_pb, synth_page, page_to_net_input
from oracle.aru_oracle import Oracle, _conv2d_same, _conv2d_transpose_same
from oracle.cv2_oracle import run_cv2

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_conv4x4_same_pads_one_before_two_after():
    # an impulse at (y,x) through an all-ones 4x4 filter covers rows y-2..y+1 of the output (pad 1 top, 2 bottom)
    x = torch.zeros(1, 1, 6, 6)
    x[0, 0, 3, 3] = 1.0
    y = _conv2d_same(x, torch.ones(4, 4, 1, 1))[0, 0].numpy()
    ys, xs = np.nonzero(y)
    assert ys.min() == 1 and ys.max() == 4 and xs.min() == 1 and xs.max() == 4


def test_pools_are_ceil_mode_with_valid_divisor():
    b = GraphBuilder()
    x = b.placeholder("inImg", [None, None, None, 1])
    b.identity("output", b.pool("p", "AvgPool", x), unique=False)
    img = np.arange(15, dtype=np.float32).reshape(3, 5)
    out = Oracle(b.serialize()).run(img)[0, ..., 0]
    assert out.shape == (2, 3)
    assert out[0, 0] == pytest.approx((0 + 1 + 5 + 6) / 4) and out[0, 2] == pytest.approx((4 + 9) / 2)
    assert out[1, 2] == pytest.approx(14.0)
    b = GraphBuilder()
    x = b.placeholder("inImg", [None, None, None, 1])
    b.identity("output", b.pool("p", "MaxPool", x), unique=False)
    out = Oracle(b.serialize()).run(-img)[0, ..., 0]
    assert out[1, 2] == -14.0 and out[0, 0] == 0.0


@pytest.mark.parametrize("h_out", [6, 7])
def test_conv2d_transpose_crop_offset(h_out):
    # k=3, s=2, SAME: full result has size 2*Hin+1 and is cropped at 0 (even) / 1 (odd)
    h_in = -(-h_out // 2)
    x = torch.zeros(1, 1, h_in, h_in)
    x[0, 0, 1, 1] = 1.0
    w = torch.arange(9, dtype=torch.float32).reshape(3, 3, 1, 1)
    y = _conv2d_transpose_same(x, w, (h_out, h_out), 2)[0, 0].numpy()
    off = 0 if h_out % 2 == 0 else 1
    assert y.shape == (h_out, h_out)
    for ky in range(3):
        for kx in range(3):
            oy, ox = 2 + ky - off, 2 + kx - off
            if oy < h_out and ox < h_out:
                assert y[oy, ox] == ky * 3 + kx


def test_upsample_simple_sums_channels():
    # layers.py:716-720: ones filter [up,up,C,C] -> every output channel is the sum over input channels
    b = GraphBuilder()
    x = b.placeholder("inImg", [None, None, None, 1])
    w = b.variable("w", np.stack([np.full((3, 3, 1), 1.0), np.full((3, 3, 1), 2.0)], -1).astype(np.float32))
    bi = b.variable("b", np.zeros((2,), np.float32))
    c = b.bias_add("ba", b.conv2d("c", x, w), bi)
    shp = b.const("shp", np.array([1, 7, 5, 2], np.int32))
    ones = b.const("ones", np.ones((2, 2, 2, 2), np.float32), splat=True)
    b.identity("output", b.conv2d_transpose("up", shp, ones, b.pool("p", "MaxPool", c), 2), unique=False)
    img = np.random.default_rng(0).random((7, 5)).astype(np.float32)
    out = Oracle(b.serialize()).run(img)[0]
    assert out.shape == (7, 5, 2)
    np.testing.assert_allclose(out[..., 0], out[..., 1])
    conv = _conv2d_same(torch.from_numpy(img)[None, None], torch.ones(3, 3, 1, 1))[0, 0]
    pooled = torch.nn.functional.max_pool2d(conv[None, None], 2, 2, ceil_mode=True)[0, 0].numpy() * 3.0
    # crop offset (4*2 - 7)//2 = 0 rows, (3*2 - 5)//2 = 0 cols
    np.testing.assert_allclose(out[..., 0], np.repeat(np.repeat(pooled, 2, 0), 2, 1)[:7, :5], rtol=1e-5)


@pytest.mark.parametrize("net,h,w", [
    ("tiny", 40, 32), ("tiny", 45, 39), ("ru", 48, 33), ("tiny_sigmoid", 21, 30),
    # the graphs the engine is benchmarked and parity-tested on, at BASELINE sizes: configs[0] / [1] (1024x768, separator
    # and heading nets), the CLI default size (odd at every pyramid level) and the upstream S = 6 / A = 5 topology
    ("separator", 1024, 768), ("heading", 1024, 768), ("separator", 1500, 1125), ("aru_s6a5", 257, 193),
])
def test_oracle_matches_opencv_importer(net, h, w):
    pb = synth_pb(net)
    x = page_to_net_input(synth_page(h, w, 11))
    a = Oracle(pb).run(x)[0]
    b = run_cv2(pb, x)
    assert a.shape == b.shape == (h, w, a.shape[-1])
    assert np.abs(a - b).max() < 2e-5


def test_oracle_intermediate_matches_opencv():
    pb = synth_pb("tiny")
    x = page_to_net_input(synth_page(45, 39, 3))
    name = "aru_net/featMapG/unet_up_0/deconv/activation"
    a = Oracle(pb).run(x, fetch=name)[0]
    b = run_cv2(pb, x, fetch=name)
    assert np.abs(a - b).max() < 2e-5


def test_fp64_oracle_agrees_with_fp32():
    pb = synth_pb("tiny")
    x = page_to_net_input(synth_page(40, 32, 5))
    a = Oracle(pb).run(x)[0]
    b = Oracle(pb, dtype=torch.float64).run(x)[0]
    assert np.abs(a - b).max() < 1e-5


@pytest.mark.parametrize("path", sorted(p for p in glob.glob(os.path.join(GOLDEN, "*.npz")) if not os.path.basename(p).startswith("post_")), ids=os.path.basename)
def test_oracle_reproduces_golden(path):
    g = np.load(path)
    net = os.path.basename(path).rsplit("_", 1)[0]
    pb = synth_pb(net)
    assert hashlib.sha256(pb).hexdigest() == str(g["pb_sha256"]), "synthetic graph writer is no longer deterministic"
    got = Oracle(pb).run(page_to_net_input(g["page"]))[0]
    assert np.abs(got - g["prob"]).max() < 1e-5


def test_batch_equals_per_page():
    pb = synth_pb("tiny")
    xs = np.stack([page_to_net_input(synth_page(33, 27, s)) for s in range(3)])
    o = Oracle(pb)
    full = o.run(xs)
    for i in range(3):
        assert np.abs(full[i] - o.run(xs[i])[0]).max() < 1e-6
