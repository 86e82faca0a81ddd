"""GPU parity tests proper: the CUDA engine (through the C ABI) against the CPU oracle, the committed
golden fixtures, and size-independent properties at the benchmark sizes.

Tolerances are BASELINE.json's: probability maps max-abs <= 2e-2, mean-abs <= 1e-3 (16-bit operands,
fp32 accumulate), >= 99.9 % pixel agreement of the binarised mask (u8 > 0.05*255, i.e. u8 >= 13;
separator_net_post_processor.py:147-149).  The uint8 / mask kernels are integer work and are checked
bit-exactly against numpy applied to the engine's own float output.
"""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MAX_ABS, MEAN_ABS, MASK_AGREE = 2e-2, 1e-3, 0.999
PATHS = [(0, "tcgen05"), (1, "cuda-core"), (2, "tcgen05-position-major"), (3, "tcgen05-row-banded")]


def _mask(p):
    return (p[..., 0] * 255).astype(np.uint8) > 12.75


def _check(got, ref, what=""):
    assert got.shape == ref.shape and got.dtype == np.float32, what
    assert np.isfinite(got).all(), what
    d = np.abs(got - ref)
    agree = float((_mask(got) == _mask(ref)).mean())
    assert d.max() <= MAX_ABS, f"{what}: max-abs {d.max():.3e}"
    assert d.mean() <= MEAN_ABS, f"{what}: mean-abs {d.mean():.3e}"
    assert agree >= MASK_AGREE, f"{what}: mask agreement {agree:.5f}"


@pytest.fixture(scope="module")
def engines(built_lib):
    from aru_b200.engine import Engine
    from aru_b200.synth import synth_pb
    cache = {}

    def get(net):
        if net not in cache:
            cache[net] = Engine(synth_pb(net), device=0)
        return cache[net]

    yield get
    for e in cache.values():
        e.close()


@pytest.fixture(scope="module")
def oracles():
    from aru_b200.synth import synth_pb
    from oracle.aru_oracle import Oracle
    cache = {}

    def get(net):
        if net not in cache:
            cache[net] = Oracle(synth_pb(net))
        return cache[net]

    return get


def _native_loaded():
    with open("/proc/self/maps") as f:
        return "libaru_b200.so" in f.read()


def test_native_library_is_what_runs(engines):
    engines("tiny")
    assert _native_loaded()


@pytest.mark.parametrize("path,pname", PATHS)
@pytest.mark.parametrize("fixture", sorted(p for p in glob.glob(os.path.join(GOLDEN, "*.npz")) if not os.path.basename(p).startswith("post_")), ids=os.path.basename)
def test_golden_fixtures(engines, fixture, path, pname):
    from aru_b200.engine import OPT_CONV_PATH
    from aru_b200.synth import page_to_net_input
    g = np.load(fixture)
    net = os.path.basename(fixture).rsplit("_", 1)[0]
    eng = engines(net)
    eng.set_option(OPT_CONV_PATH, path)
    got = eng.forward(page_to_net_input(g["page"]))[0]
    _check(got, g["prob"], f"{os.path.basename(fixture)} [{pname}]")


@pytest.mark.parametrize("path,pname", PATHS)
@pytest.mark.parametrize("net,h,w", [
    ("separator", 128, 96),      # /32 in both dims
    ("separator", 150, 113),     # odd at every level (the CLI default 1500x1125, scaled down)
    ("separator", 257, 130),     # W+2 spans several 128-position tiles; odd H
    ("separator", 40, 1250),     # wide page: the 128-channel layers take the split-K / weight-streaming launches
    ("heading", 90, 68),         # heading default height 900, scaled down
    ("ru", 77, 101),
    ("aru_s6a5", 129, 97),
    ("tiny", 8, 8), ("tiny", 9, 200), ("tiny", 200, 9), ("tiny", 1, 1),
])
def test_parity_with_oracle(engines, oracles, net, h, w, path, pname):
    from aru_b200.engine import OPT_CONV_PATH
    from aru_b200.synth import synth_page, page_to_net_input
    x = page_to_net_input(synth_page(h, w, seed=h * 1000 + w))
    ref = oracles(net).run(x)[0]
    eng = engines(net)
    eng.set_option(OPT_CONV_PATH, path)
    _check(eng.forward(x)[0], ref, f"{net} {h}x{w} [{pname}]")


@pytest.mark.parametrize("path,pname", PATHS)
def test_per_layer_parity(engines, oracles, path, pname):
    """Every surviving GraphDef node of the tiny net against the oracle's value of the same node."""
    from aru_b200.engine import OPT_CONV_PATH, OPT_KEEP_ALL
    from aru_b200.synth import synth_page, page_to_net_input
    eng, orc = engines("tiny"), oracles("tiny")
    eng.set_option(OPT_CONV_PATH, path)
    eng.set_option(OPT_KEEP_ALL, 1)     # also store tensors that only a fused pool reads
    x = page_to_net_input(synth_page(45, 39, seed=3))
    eng.forward(x)
    checked = 0
    for node in eng.program.tensor_of_node:
        if orc.ir[node].op in ("Placeholder", "ConcatV2"):
            continue
        ref = orc.run(x, fetch=node)[0]
        got = eng.read_node(node)
        scale = max(1.0, float(np.abs(ref).max()))
        assert got.shape == ref.shape, node
        assert np.abs(got - ref).max() <= 2e-2 * scale, f"{node} [{pname}]: {np.abs(got - ref).max():.3e} (scale {scale:.2f})"
        checked += 1
    assert checked > 40
    eng.set_option(OPT_KEEP_ALL, 0)


@pytest.mark.parametrize("net,n,h,w", [
    ("separator", 1, 257, 130),     # odd sizes, one strip at every level
    ("separator", 2, 150, 613),     # several strips (W > 126) and a batch: strip halo columns, page boundaries
    ("separator", 3, 420, 300),     # more row tiles than TMEM buffers, several CTAs per strip segment
    ("heading", 1, 90, 68),
    ("aru_s6a5", 1, 129, 97),
    ("ru", 1, 333, 101),
    ("tiny", 1, 9, 200), ("tiny", 1, 1, 1),
])
def test_fused_conv_pairs_against_one_launch_per_layer(engines, oracles, net, n, h, w):
    """conv_band2.cu (two chained 3x3 convolutions per launch, intermediate in a shared-memory row FIFO) against the
    same layers run one launch each (ARU_OPT_FUSE_PAIRS 0): same 16-bit operands and fp32 accumulation; the vertical
    tap pairs group rows with the other parity, so the two agree to fp32 summation order (<= 5e-3 on the probabilities,
    like the other kernel paths) - and each of them is within BASELINE's tolerance of the oracle."""
    from aru_b200.engine import OPT_CONV_PATH, OPT_FUSE_PAIRS
    from aru_b200.synth import synth_page, page_to_net_input
    eng = engines(net)
    eng.set_option(OPT_CONV_PATH, 0)
    x = np.stack([page_to_net_input(synth_page(h, w, seed=7 * h + w + i)) for i in range(n)]).astype(np.float32)
    try:
        eng.set_option(OPT_FUSE_PAIRS, 1)
        fused = eng.forward(x).copy()
        names = [k for _, k, _ in eng.profile_ops(1)]
        if net != "tiny" or min(h, w) > 1:
            assert any(k.startswith("conv_band2") for k in names), names
        eng.set_option(OPT_FUSE_PAIRS, 0)
        plain = eng.forward(x).copy()
        assert not any(k.startswith("conv_band2") for _, k, _ in eng.profile_ops(1))
    finally:
        eng.set_option(OPT_FUSE_PAIRS, 0)
    assert np.abs(fused - plain).max() <= 5e-3, float(np.abs(fused - plain).max())
    ref = oracles(net).run(x)
    for i in range(n):
        _check(fused[i], ref[i], f"{net} {h}x{w} page {i} [fused pairs]")
    # a page gives the same bits alone and inside a batch (tiles are partitioned differently over the CTAs), run to run
    eng.set_option(OPT_FUSE_PAIRS, 1)
    try:
        if n > 1:
            assert np.array_equal(eng.forward(x[n - 1])[0], fused[n - 1])
        assert np.array_equal(eng.forward(x), fused)
    finally:
        eng.set_option(OPT_FUSE_PAIRS, 0)


@pytest.mark.parametrize("net,n,h,w", [
    ("separator", 1, 257, 130),     # odd sizes, one strip at every level
    ("separator", 2, 150, 613),     # several strips (W > 168) and a batch: strip halo columns, page boundaries
    ("separator", 3, 420, 300),     # several row segments per page and strip
    ("separator", 1, 37, 1250),     # wide and flat: more pipeline fill than rows
    ("heading", 1, 90, 68),
    ("aru_s6a5", 1, 129, 97),
    ("ru", 1, 333, 101),
    ("tiny", 1, 9, 200), ("tiny", 1, 1, 1), ("tiny", 2, 3, 5),
])
def test_fused_residual_blocks_against_one_launch_per_layer(engines, oracles, net, n, h, w):
    """block_mma.cu (a whole residual block of the 8-channel levels per launch, intermediates in shared memory, warp-level
    tensor cores) against the same layers run one launch each (ARU_OPT_FUSE_BLOCKS 0): same 16-bit operands and rounding
    points (every intermediate is rounded to 16 bits where the per-layer path stores it), fp32 accumulation in another
    order and the residual added before the one rounding instead of after it - and each path is within BASELINE's
    tolerance of the oracle."""
    from aru_b200.engine import OPT_CONV_PATH, OPT_FUSE_BLOCKS
    from aru_b200.synth import synth_page, page_to_net_input
    eng = engines(net)
    eng.set_option(OPT_CONV_PATH, 0)
    x = np.stack([page_to_net_input(synth_page(h, w, seed=11 * h + w + i)) for i in range(n)]).astype(np.float32)
    try:
        eng.set_option(OPT_FUSE_BLOCKS, 1)
        fused = eng.forward(x).copy()
        names = [k for _, k, _ in eng.profile_ops(1)]
        assert any(k.startswith("block_mma") for k in names), names
        again = eng.forward(x).copy()
        alone = eng.forward(x[n - 1])[0].copy()
        eng.set_option(OPT_FUSE_BLOCKS, 0)
        plain = eng.forward(x).copy()
        assert not any(k.startswith("block_mma") for _, k, _ in eng.profile_ops(1))
    finally:
        eng.set_option(OPT_FUSE_BLOCKS, 1)
    assert np.abs(fused - plain).max() <= 5e-3, float(np.abs(fused - plain).max())
    ref = oracles(net).run(x)
    for i in range(n):
        _check(fused[i], ref[i], f"{net} {h}x{w} page {i} [fused blocks]")
    # run to run, and a page alone against the same page inside a batch (row segments are cut differently)
    assert np.array_equal(again, fused)
    assert np.array_equal(alone, fused[n - 1])


@pytest.mark.parametrize("net,n,h,w", [
    ("separator", 1, 257, 130),     # odd sizes: partial tiles in both directions, odd crop offsets of the upsampled maps
    ("separator", 2, 150, 613),     # several column tiles and a batch
    ("separator", 1, 29, 131),      # one row tile shorter than a tile, 3 columns into the second column tile
    ("heading", 2, 90, 68),
    ("aru_s6a5", 1, 129, 97),       # five attention scales
    ("tiny", 3, 9, 200),            # two scales, pages shorter than a tile
    ("tiny_aru_sigmoid", 1, 48, 40),   # one class, sigmoid head
])
def test_fused_attention_tail_and_classifier_against_two_launches(built_lib, oracles, monkeypatch, net, n, h, w):
    """combine_head.cu (attention combine + 4x4 classifier in one launch, the combined map in shared memory) against
    k_combine followed by the stand-alone head: the intermediate is the same bit for bit (same expression, same rounding),
    the classifier's fp32 sums are accumulated in another order -> probabilities within 5e-5; and against the oracle."""
    from aru_b200.engine import Engine
    from aru_b200.synth import synth_pb, synth_page, page_to_net_input
    x = np.stack([page_to_net_input(synth_page(h, w, seed=13 * h + w + i)) for i in range(n)]).astype(np.float32)
    outs = {}
    for fuse in ("1", "0"):
        monkeypatch.setenv("ARU_FUSE_HEAD", fuse)
        eng = Engine(synth_pb(net), device=0)
        outs[fuse] = eng.forward(x).copy()
        names = [k for _, k, _ in eng.profile_ops(1)]
        assert ("combine_head" in names) == (fuse == "1"), names
        assert ("combine" in names) == (fuse == "0"), names
        if fuse == "1":
            assert np.array_equal(eng.forward(x), outs[fuse])
            assert np.array_equal(eng.forward(x[n - 1])[0], outs[fuse][n - 1])
        eng.close()
    assert np.abs(outs["1"] - outs["0"]).max() <= 5e-5, float(np.abs(outs["1"] - outs["0"]).max())
    ref = oracles(net).run(x)
    for i in range(n):
        _check(outs["1"][i], ref[i], f"{net} {h}x{w} page {i} [fused head]")


@pytest.mark.parametrize("fuse", [0, 1])
def test_two_fresh_engines_give_identical_bits(built_lib, fuse):
    """Kernel selection is a fixed rule (no plan-time timing): two engines, two plans, same pages -> same bits, at the
    benchmark size (many tiles per CTA: the steady state of the kernels' pipelines), with and without fused conv pairs."""
    from aru_b200.engine import Engine, OPT_FUSE_PAIRS
    from aru_b200.synth import synth_pb, synth_page, page_to_net_input
    x = np.stack([page_to_net_input(synth_page(1856, 1344, s)) for s in range(2)]).astype(np.float32)
    outs = []
    for _ in range(2):
        eng = Engine(synth_pb("separator"), device=0)
        eng.set_option(OPT_FUSE_PAIRS, fuse)
        outs.append(eng.forward(x).copy())
        outs.append(eng.forward(x).copy())
        eng.close()
    for o in outs[1:]:
        assert np.array_equal(outs[0], o)


@pytest.mark.parametrize("net,n,h,w", [("separator", 3, 150, 213), ("aru_s6a5", 2, 129, 97), ("heading", 1, 257, 130),
                                        ("tiny", 2, 9, 200), ("ru", 1, 64, 48)])
def test_branch_streams_give_the_same_bits(built_lib, net, n, h, w):
    """ARU_OPT_BRANCH_STREAMS: the pyramid scales and the attention CNNs of a pass on streams of their own (events on the
    edges that cross streams, fork / join inside the captured graph) against everything on one stream: the same launches,
    so the same bits - eagerly (first pass), from the captured graph (later passes), through the host-buffer pipeline with
    micro-batches and through the device-resident call."""
    import torch
    from aru_b200.engine import Engine, OPT_BRANCH_STREAMS, OPT_MICRO_BATCH
    from aru_b200.synth import synth_pb, synth_page, page_to_net_input
    x = np.stack([page_to_net_input(synth_page(h, w, seed=17 * h + w + i)) for i in range(n)]).astype(np.float32)
    outs = {}
    for branch in (1, 0):
        eng = Engine(synth_pb(net), device=0)
        eng.set_option(OPT_BRANCH_STREAMS, branch)
        runs = [eng.forward(x).copy() for _ in range(4)]          # eager pass, capture, two replays
        eng.set_option(OPT_MICRO_BATCH, 2)
        runs.append(eng.forward(x).copy())
        eng.set_option(OPT_MICRO_BATCH, 0)
        x_dev = torch.from_numpy(x).cuda()
        y_dev = torch.zeros((n, h, w, eng.n_class), dtype=torch.float32, device="cuda")
        for _ in range(3):
            eng.forward_device(x_dev.data_ptr(), n, h, w, out_ptr=y_dev.data_ptr())
        eng.sync()
        runs.append(y_dev.cpu().numpy())
        for r in runs[1:]:
            assert np.array_equal(runs[0], r), (net, branch)
        outs[branch] = runs[0]
        eng.close()
    assert np.array_equal(outs[0], outs[1])


def test_device_resident_call_binds_caller_buffers_in_place(built_lib):
    """aru_forward_device on caller-owned device buffers (torch tensors): the first layers read the caller's input and
    the classifier writes the caller's output, no staging copies, one captured graph per (in, out) binding.  Same bits
    as the host-buffer call; more bindings than slots (the least recently used one is recycled), repeated calls (eager
    pass, then the graph), a misaligned buffer (falls back to the staging copies), uint8 / mask outputs, micro-batches."""
    import torch
    from aru_b200.engine import Engine, OPT_MICRO_BATCH
    from aru_b200.synth import synth_pb, synth_page, page_to_net_input
    n, h, w = 5, 96, 80
    eng = Engine(synth_pb("separator"), device=0)
    xs = np.stack([page_to_net_input(synth_page(h, w, s)) for s in range(n)]).astype(np.float32)
    want, want_u8, want_mask = eng.forward(xs, want_u8=True, want_mask=True, threshold=0.05)
    want, want_u8, want_mask = want.copy(), want_u8.copy(), want_mask.copy()
    x_dev = torch.from_numpy(xs).cuda()
    for mb in (0, 2):
        eng.set_option(OPT_MICRO_BATCH, mb)
        outs = [torch.zeros((n, h, w, 2), dtype=torch.float32, device="cuda") for _ in range(6)]   # 6 bindings, 4 slots
        for rep in range(3):
            for o in outs:
                o.zero_()
                eng.forward_device(x_dev.data_ptr(), n, h, w, out_ptr=o.data_ptr())
                eng.sync()
                assert np.array_equal(o.cpu().numpy(), want), (mb, rep)
        # the input must be untouched, uint8 / mask outputs come from the bound output
        assert np.array_equal(x_dev.cpu().numpy(), xs)
        u8 = torch.zeros((n, h, w, 2), dtype=torch.uint8, device="cuda")
        mask = torch.zeros((n, h, w), dtype=torch.uint8, device="cuda")
        eng.forward_device(x_dev.data_ptr(), n, h, w, out_ptr=outs[0].data_ptr(), u8_ptr=u8.data_ptr(), mask_ptr=mask.data_ptr())
        eng.sync()
        assert np.array_equal(u8.cpu().numpy(), want_u8) and np.array_equal(mask.cpu().numpy(), want_mask)
        # without a float32 result (staging path) and with a misaligned input (4 bytes off a 16-byte boundary)
        u8.zero_()
        eng.forward_device(x_dev.data_ptr(), n, h, w, u8_ptr=u8.data_ptr())
        eng.sync()
        assert np.array_equal(u8.cpu().numpy(), want_u8)
        flat = torch.zeros(n * h * w + 1, dtype=torch.float32, device="cuda")
        flat[1:] = x_dev.reshape(-1)
        outs[1].zero_()
        eng.forward_device(flat.data_ptr() + 4, n, h, w, out_ptr=outs[1].data_ptr())
        eng.sync()
        assert np.array_equal(outs[1].cpu().numpy(), want)
    eng.close()


def test_batch_and_micro_batch_consistency(engines):
    """forward([p0..p4]) == forward(pi) per page, bit-exact, for any micro-batch split (pages are independent)."""
    from aru_b200.engine import OPT_CONV_PATH, OPT_MICRO_BATCH
    from aru_b200.synth import synth_page, page_to_net_input
    eng = engines("separator")
    eng.set_option(OPT_CONV_PATH, 0)
    xs = np.stack([page_to_net_input(synth_page(96, 80, s)) for s in range(5)])
    eng.set_option(OPT_MICRO_BATCH, 0)
    full = eng.forward(xs).copy()
    for i in range(5):
        assert np.array_equal(full[i], eng.forward(xs[i])[0]), i
    eng.set_option(OPT_MICRO_BATCH, 2)       # 2 + 2 + 1: exercises the pipelined double buffering and a tail plan
    assert np.array_equal(full, eng.forward(xs))
    eng.set_option(OPT_MICRO_BATCH, 0)


def test_uint8_and_mask_outputs_are_bit_exact(engines):
    from aru_b200.synth import synth_page, page_to_net_input
    eng = engines("separator")
    xs = np.stack([page_to_net_input(synth_page(101, 77, s)) for s in range(3)])
    prob, u8, mask = eng.forward(xs, want_u8=True, want_mask=True, threshold=0.05)
    ref_u8 = np.array(prob * 255, dtype=np.uint8)                          # separator_net_post_processor.py:147
    ref_mask = np.array((ref_u8[..., 0] > 0.05 * 255) * 255, dtype=np.uint8)  # helper.py:75-78
    assert np.array_equal(u8, ref_u8) and np.array_equal(mask, ref_mask)
    assert 0 < (mask > 0).mean() < 1


def test_uint8_channel_subset_output(engines):
    """ARU_OPT_U8_CHANNELS: the uint8 output restricted to the leading channel equals channel 0 of the full map, for the
    float-page and the uint8-page calls, and the option does not leak into later calls."""
    from aru_b200.synth import synth_page
    eng = engines("separator")
    pages = np.stack([synth_page(90, 133, s) for s in range(5)])
    x = (pages / 255.0).astype(np.float32)
    _, full = eng.forward(x, want_u8=True)
    only0 = eng.forward(x, want_u8=True, want_prob=False, u8_channels=1)
    assert only0.shape == (5, 90, 133, 1) and np.array_equal(only0[..., 0], full[..., 0])
    r1 = eng.separator_pages(pages, want_u8=True, want_mask=True, want_separators=False, u8_channels=1)
    r2 = eng.separator_pages(pages, want_u8=True, want_mask=True, want_separators=False)
    assert r2["u8"].shape[-1] == eng.n_class and np.array_equal(r1["u8"][..., 0], r2["u8"][..., 0])
    assert np.array_equal(r1["mask"], r2["mask"])
    sums, _, u8 = eng.heading_pages(pages, [(0, 3, 40, 5, 90), (4, 0, 90, 0, 133)], want_u8=True)
    assert u8.shape[-1] == eng.n_class
    assert int(sums[1]) == int(u8[4, :, :, 0].astype(np.uint64).sum())


def test_bf16_build_is_measured_against_the_oracle(built_lib):
    """The same sources built with -DARU_USE_BF16 (bf16 storage and tensor-core operands, fp32 accumulate: the dtype
    BASELINE's north_star names) run through the oracle parity cases in a fresh interpreter (ARU_B200_LIB selects the
    library).  Measured (profiles/r02f_parity_bf16.txt): with 8 mantissa bits per stored activation the 100+ layer nets
    land at max-abs 1.2e-2 .. 4.5e-2 and 99.4 - 99.9 % mask agreement - OUTSIDE north_star's own 2e-2 / 99.9 % bounds on
    most shapes, while the default fp16 storage (11 bits, saturating at 65504) stays below 6e-3 / above 99.93 %
    (profiles/r02f_parity_fp16.txt).  That is why fp16 is the product build; this test pins the bf16 numbers to the
    looser envelope they actually reach so that a regression of the variant is still caught."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "citlab-article-separation-new_b200", "libaru_b200_bf16.so")
    assert os.path.exists(lib), "build() must produce the bf16 variant"
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "parity_table.py")],
                       env=dict(os.environ, ARU_B200_LIB=lib), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    table = json.loads(r.stdout.strip().splitlines()[-1])
    assert table["library"] == "libaru_b200_bf16.so" and len(table["rows"]) >= 10
    for row in table["rows"]:
        assert row["max_abs"] <= 8e-2 and row["mean_abs"] <= 6e-3 and row["mask_agree"] >= 0.99, row


def test_float64_input_and_4d_input(engines):
    from aru_b200.net_boundary import GraphHandle, get_net_output
    from aru_b200.synth import synth_pb, synth_page, page_to_net_input
    x = page_to_net_input(synth_page(64, 48, 1))
    assert x.dtype == np.float64
    h = GraphHandle(synth_pb("tiny"))
    a = get_net_output(x, h, gpu_device="0")
    b = get_net_output(x[None, :, :, None].astype(np.float32), h, gpu_device="")
    assert a.shape == (64, 48, 2) and a.dtype == np.float32 and a.flags.c_contiguous
    assert np.array_equal(a, b)


def test_tensor_core_and_cuda_core_paths_agree(engines):
    from aru_b200.engine import OPT_CONV_PATH
    from aru_b200.synth import synth_page, page_to_net_input
    eng = engines("separator")
    x = page_to_net_input(synth_page(300, 200, 9))
    eng.set_option(OPT_CONV_PATH, 0)
    a = eng.forward(x).copy()
    names = {k for _, k, _ in eng.profile_ops(1)}
    assert any(k.startswith("conv_tc") or k.startswith("conv_band") for k in names), names
    eng.set_option(OPT_CONV_PATH, 3)
    c = eng.forward(x).copy()
    assert any(k.startswith("conv_band") for _, k, _ in eng.profile_ops(1))
    assert np.abs(a - c).max() <= 5e-3
    eng.set_option(OPT_CONV_PATH, 1)
    b = eng.forward(x).copy()
    assert not any(k.startswith("conv_tc") or k.startswith("conv_band") for _, k, _ in eng.profile_ops(1))
    eng.set_option(OPT_CONV_PATH, 0)
    assert np.abs(a - b).max() <= 5e-3          # same 16-bit operands, different fp32 summation order


def test_full_size_page_against_oracle(engines, oracles):
    """BASELINE config 1: one 1024x768 page, separator net (the oracle needs a few seconds)."""
    from aru_b200.synth import synth_page, page_to_net_input
    x = page_to_net_input(synth_page(1024, 768, 0))
    _check(engines("separator").forward(x)[0], oracles("separator").run(x)[0], "separator 1024x768")


def test_full_size_page_heading_net_against_oracle(engines, oracles):
    """BASELINE config 2: the same 1024x768 page through the heading net."""
    from aru_b200.synth import synth_page, page_to_net_input
    x = page_to_net_input(synth_page(1024, 768, 0))
    _check(engines("heading").forward(x)[0], oracles("heading").run(x)[0], "heading 1024x768")


def test_benchmark_shape_page_against_oracle(engines, oracles):
    """BASELINE config 3 shape: one 1856x1344 page, separator net, against the oracle (and inside a batch of three: the
    page must give the same bits as alone)."""
    from aru_b200.synth import synth_page, page_to_net_input
    eng = engines("separator")
    xs = np.stack([page_to_net_input(synth_page(1856, 1344, s)) for s in (5, 6, 7)]).astype(np.float32)
    got = eng.forward(xs).copy()
    _check(got[1], oracles("separator").run(xs[1])[0], "separator 1856x1344")
    assert np.array_equal(eng.forward(xs[1])[0], got[1])


def test_cli_default_size_page_against_oracle(engines, oracles):
    """The CLI's default net input (fixed height 1500 -> 1500x1125 for a 4:3 scan, run_net_post_processing.py:51-57): odd at
    every pyramid level, so every ceil-mode pool, transposed-conv crop and upsample offset is exercised at full size."""
    from aru_b200.synth import synth_page, page_to_net_input
    x = page_to_net_input(synth_page(1500, 1125, 3))
    _check(engines("separator").forward(x)[0], oracles("separator").run(x)[0], "separator 1500x1125")
    x = page_to_net_input(synth_page(900, 675, 4))       # the heading net's default height
    _check(engines("heading").forward(x)[0], oracles("heading").run(x)[0], "heading 900x675")


def test_broadsheet_page_against_oracle(engines, oracles):
    """BASELINE config 5: one 6000x4500 page through the full multi-scale pyramid (27 Mpx: 64-bit index math, an 11.6 GB
    arena, one page per pass) against the oracle; the oracle needs ~20 s and ~25 GB of host memory."""
    from aru_b200.synth import synth_page, page_to_net_input
    x = page_to_net_input(synth_page(6000, 4500, 1)).astype(np.float32)
    got = engines("separator").forward(x)[0]
    ref = oracles("separator").run(x)[0]
    _check(got, ref, "separator 6000x4500")


def test_benchmark_size_properties(engines):
    """BASELINE config 3 shape (1856x1344): batch independence and determinism at full size."""
    from aru_b200.synth import synth_page, page_to_net_input
    eng = engines("separator")
    xs = np.stack([page_to_net_input(synth_page(1856, 1344, s)) for s in range(3)]).astype(np.float32)
    a = eng.forward(xs).copy()
    b = eng.forward(xs[::-1].copy())
    assert np.array_equal(a, b[::-1])             # page order does not matter
    assert np.isfinite(a).all() and 0.0 <= a.min() and a.max() <= 1.0
    np.testing.assert_allclose(a.sum(-1), 1.0, atol=1e-5)   # softmax over classes
    frac = float(_mask(a).mean())
    assert 0.001 < frac < 0.5, frac


def test_cli_dump_matches_the_engine(tmp_path, engines):
    """run_net_post_processing --dump_dir on two ranks' worth of pages == Engine.forward on the same scaled pages."""
    import cv2
    from aru_b200 import net_boundary
    from aru_b200.run_net_post_processing import main
    from aru_b200.synth import synth_pb, synth_page
    pb = tmp_path / "sep.pb"
    pb.write_bytes(synth_pb("separator"))
    paths = []
    for i, (h, w) in enumerate([(120, 90), (150, 100), (120, 90), (90, 130), (200, 64)]):
        p = tmp_path / f"page_{i}.png"
        cv2.imwrite(str(p), synth_page(h, w, 40 + i))
        paths.append(str(p))
    lst = tmp_path / "pages.lst"
    lst.write_text("\n".join(paths) + "\n")
    out = tmp_path / "out"
    base = ["--path_to_image_list", str(lst), "--path_to_pb", str(pb), "--mode", "separator", "--fixed_height", "96",
            "--dump_dir", str(out)]
    env = dict(os.environ)
    try:
        for r in range(2):                                     # two ranks, run back to back on the one GPU
            os.environ.update(RANK=str(r), WORLD_SIZE="2", LOCAL_RANK="0")
            assert main(base) == 0
    finally:
        os.environ.clear()
        os.environ.update(env)
    import json
    seen = []
    for r in range(2):
        seen += list(json.load(open(out / f"manifest_rank{r}.json"))["pages"])
    assert sorted(seen) == sorted(paths)
    eng = engines("separator")
    for p in paths:
        _, grey, _ = net_boundary.load_and_scale_image(p, 96, 1.0)
        assert grey.shape[0] == 96
        _, u8, mask = eng.forward(grey, want_u8=True, want_mask=True, threshold=0.05)
        stem = os.path.splitext(os.path.basename(p))[0]
        assert np.array_equal(cv2.imread(str(out / f"{stem}_prob.png"), cv2.IMREAD_UNCHANGED), u8[0, :, :, 0])
        assert np.array_equal(cv2.imread(str(out / f"{stem}_mask.png"), cv2.IMREAD_UNCHANGED), mask[0])


def test_plan_cache_eviction_and_lazy_graphs(engines):
    """More page shapes than the plan cache holds (8), each met one to three times: a plan's first pass runs eagerly, later
    passes replay the CUDA graph captured on the second, evicted shapes are rebuilt - the results must not depend on any
    of it."""
    from aru_b200.synth import synth_page, page_to_net_input
    eng = engines("tiny")
    pages = {w: page_to_net_input(synth_page(48, w, seed=w)).astype(np.float32) for w in range(40, 52)}
    first = {w: eng.forward(x)[0].copy() for w, x in pages.items()}          # 12 shapes: the early ones get evicted
    for w, x in pages.items():                                               # rebuilt (eager), then graph, then graph
        for _ in range(3):
            assert np.array_equal(eng.forward(x)[0], first[w]), w


@pytest.mark.parametrize("n", [9, 13, 16, 23, 37])
def test_micro_batch_schedules_cover_every_page_once(engines, n):
    """The host-buffer call splits n pages into a head, full passes and a (ramped) tail; whatever the split, page i of the
    batch must equal page i run alone - for the float32 outputs (ramped tail) and for the uint8 path (plain edges)."""
    from aru_b200.engine import OPT_CONV_PATH, OPT_MICRO_BATCH
    from aru_b200.synth import synth_page
    eng = engines("separator")
    eng.set_option(OPT_CONV_PATH, 0)
    pages = np.stack([synth_page(64, 112, seed=500 + i) for i in range(n)])
    x = (pages / 255.0).astype(np.float32)
    eng.set_option(OPT_MICRO_BATCH, 0)
    alone = np.stack([eng.forward(x[i])[0] for i in range(n)])
    try:
        for mb in (8, 16):
            eng.set_option(OPT_MICRO_BATCH, mb)
            got = eng.forward(x)
            assert np.array_equal(got, alone), (n, mb)
            r = eng.separator_pages(pages, want_u8=True, want_separators=False)
            assert np.array_equal(r["u8"], (alone * 255).astype(np.uint8)), (n, mb)
    finally:
        eng.set_option(OPT_MICRO_BATCH, 0)


def test_async_calls_overlap_and_give_the_same_results(engines):
    """Engine.submit / wait (ARU_OPT_ASYNC): several host-buffer calls in flight give the bits of the synchronous calls,
    whatever the order of the waits, for the float and the uint8-page entry points."""
    from aru_b200.synth import synth_page
    eng = engines("separator")
    pages = [np.stack([synth_page(120, 112, 40 * k + i) for i in range(5)]) for k in range(4)]
    xs = [(p / 255.0).astype(np.float32) for p in pages]
    want = [eng.forward(x).copy() for x in xs]
    want_pages = [eng.separator_pages(p, want_u8=True) for p in pages]
    tickets = [eng.submit(eng.forward, x) for x in xs]
    for k in (2, 0, 3, 1):
        eng.wait(tickets[k][0])
        assert np.array_equal(tickets[k][1], want[k]), k
    tickets = [eng.submit(eng.separator_pages, p, want_u8=True) for p in pages]
    for k, (t, r) in enumerate(tickets):
        eng.wait(t)
        for key in ("u8", "horizontal", "vertical"):
            assert np.array_equal(r[key], want_pages[k][key]), (k, key)
    # a float64 pageable input is staged by the call: the caller's array may change right after submit
    x64 = xs[0].astype(np.float64)
    t, r = eng.submit(eng.forward, x64)
    x64[...] = 0
    eng.wait(t)
    assert np.array_equal(r, want[0])
    assert np.array_equal(eng.forward(xs[1]), want[1])          # synchronous calls still complete before returning


def test_range_check_reports_saturating_activations(built_lib):
    """fp16 stores saturate at 65504 (bf16 build: only inf / nan count).  A graph whose second layer exceeds that must
    raise the engine's warning on its first pass; the calibrated nets must not."""
    from aru_b200.engine import Engine
    from aru_b200.graphdef import GraphBuilder
    from aru_b200.synth import synth_pb, synth_page, page_to_net_input
    x = page_to_net_input(synth_page(40, 48, 1)).astype(np.float32)
    eng = Engine(synth_pb("tiny"), device=0)
    eng.forward(x)
    assert eng.last_warning == ""
    eng.close()
    b = GraphBuilder()
    inp = b.placeholder("inImg", [None, None, None, 1])
    rng = np.random.default_rng(0)
    h = inp
    for i, (cin, cout, scale) in enumerate(((1, 8, 300.0), (8, 8, 300.0), (8, 8, 1e-6))):
        w = b.variable(f"w{i}", (np.abs(rng.normal(0, 1, size=(3, 3, cin, cout))) * scale).astype(np.float32))
        bi = b.variable(f"b{i}", np.zeros((cout,), np.float32))
        h = b.relu(f"a{i}", b.bias_add(f"p{i}", b.conv2d(f"c{i}", h, w), bi))
    w = b.variable("wl", rng.normal(0, 1e-3, size=(4, 4, 8, 2)).astype(np.float32))
    bi = b.variable("bl", np.zeros((2,), np.float32))
    b.softmax("output", b.bias_add("logits", b.conv2d("cl", h, w), bi), unique=False)
    eng = Engine(b.serialize(), device=0)
    out = eng.forward(x)
    assert np.isfinite(out).all()                         # stores saturate instead of producing inf
    if os.path.basename(os.environ.get("ARU_B200_LIB", "")) != "libaru_b200_bf16.so":
        assert "storage limit" in eng.last_warning and "bf16" in eng.last_warning, eng.last_warning
    eng.close()
