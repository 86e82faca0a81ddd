"""The drop-in for the reference's plot_net_output.py.

CPU part: the helper functions against the reference's OWN functions (imported from /root/reference under the stubs of
tests/golden/_reference_import.py; build container only) and the vectorised arg-max statistics against a literal
restatement of the reference's per-pixel loop (plot_net_output.py:212-225).  GPU part: the whole call on synthetic pages
with ground truth - files written, printed / returned accuracy equal to a recomputation from Engine.forward."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_module():
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import _reference_import as R
    if R.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, R.REFERENCE_ROOT)
    if not any(isinstance(f, R._Finder) for f in sys.meta_path):
        sys.meta_path.insert(0, R._Finder())
    import importlib
    return importlib.import_module("article_separation.plot_net_output")


@pytest.mark.skipif(not os.path.isdir("/root/reference/article_separation"), reason="the reference tree only exists in the build container")
def test_helpers_equal_the_reference_functions():
    ref = _reference_module()
    from aru_b200 import plot_net_output as mine
    rng = np.random.default_rng(3)
    hyp = (rng.random((40, 31)) > 0.5).astype(np.float32)
    gt = rng.choice([0, 255, 128], size=(40, 31)).astype(np.uint8) / 255
    assert mine.compute_accuracy(hyp, gt) == ref.compute_accuracy(hyp, gt)
    for dtype in (np.uint32, np.uint8):
        img = rng.integers(0, 256, size=(40, 31, 3)).astype(dtype)
        mask = rng.choice([0, 255, 7], size=(40, 31)).astype(np.uint8)
        a = mine.apply_mask(img.copy(), mask, (10, 200, 255), alpha=0.3)
        b = ref.apply_mask(img.copy(), mask, (10, 200, 255), alpha=0.3)
        assert a.dtype == b.dtype and np.array_equal(a, b)
        a = mine.plot_image_with_net_output(img.copy(), mask)
        b = ref.plot_image_with_net_output(img.copy(), mask)
        assert np.array_equal(a, b)
    gray = rng.integers(0, 256, size=(64, 48)).astype(np.uint8)
    assert np.array_equal(mine.plot_connected_components(gray), ref.plot_connected_components(gray))


def test_argmax_statistics_equal_the_reference_loop():
    from aru_b200.plot_net_output import argmax_one_hot
    rng = np.random.default_rng(5)
    for n_class, dtype in ((2, np.float32), (3, np.float32), (1, np.float32), (2, np.int32)):
        out_img = rng.random((1, 9, 7, n_class)).astype(np.float32)
        out_img[0, 0, 0, :] = 0.5                     # a tie: the first class wins
        if dtype == np.int32:
            out_img = np.array(out_img > 0.6, np.int32)
        # plot_net_output.py:212-225, literally
        values = np.argmax(out_img, axis=3)
        want = np.zeros_like(out_img)
        counts = {"class_" + str(i): 0 for i in range(n_class)}
        for i in range(out_img.shape[0]):
            for j in range(out_img.shape[1]):
                for k in range(out_img.shape[2]):
                    want[i, j, k, values[i, j, k]] = 1
                    counts["class_" + str(values[i, j, k])] += 1
        got, got_counts = argmax_one_hot(out_img)
        assert got.dtype == want.dtype and np.array_equal(got, want) and got_counts == counts


def test_scaling_factor_follows_the_reference_branches():
    from aru_b200.plot_net_output import _scaling_factor

    def reference(img_height, rescale, fixed_height):   # plot_net_output.py:168-174
        scaling_factor = None
        if fixed_height and rescale and rescale != 1:
            scaling_factor = rescale * fixed_height / img_height
        elif fixed_height:
            scaling_factor = fixed_height / img_height
        elif rescale:
            scaling_factor = rescale
        return scaling_factor

    for h in (1000, 3333):
        for rescale in (None, 0, 1, 1.0, 0.5, 2.0):
            for fixed in (None, 0, 900, 1250):
                assert _scaling_factor(h, rescale, fixed) == reference(h, rescale, fixed), (h, rescale, fixed)


@pytest.mark.gpu
def test_plot_net_output_end_to_end(built_lib, tmp_path, capsys):
    import cv2
    from aru_b200 import plot_net_output as mine
    from aru_b200.engine import Engine
    from aru_b200.synth import synth_pb, synth_page
    pb = tmp_path / "net.pb"
    pb.write_bytes(synth_pb("separator"))
    (tmp_path / "C2").mkdir()
    (tmp_path / "out").mkdir()
    paths = []
    for s in range(2):
        page = synth_page(400, 300, seed=s)
        path = tmp_path / f"page{s}.png"
        cv2.imwrite(str(path), cv2.cvtColor(page, cv2.COLOR_GRAY2BGR))
        gt0 = ((page < 128) * 255).astype(np.uint8)
        cv2.imwrite(str(tmp_path / "C2" / f"page{s}_GT0.png"), gt0)
        cv2.imwrite(str(tmp_path / "C2" / f"page{s}_GT1.png"), 255 - gt0)
        paths.append(str(path))
    lst = tmp_path / "pages.lst"
    lst.write_text("\n".join(paths) + "\n")
    acc = mine.plot_net_output(str(pb), str(lst), save_folder=str(tmp_path / "out"), rescale=0.5, plot_with_img=True,
                               calculate_accuracy=True)
    printed = capsys.readouterr().out
    assert printed.count("\nAccuracy = ") == 2 and printed.count("Overall Accuracy = ") == 1 and "Percentage of pixels in class_1" in printed
    # recomputation from the engine's own output of the same scaled pages
    eng = Engine(synth_pb("separator"), device=0)
    for s, path in enumerate(paths):
        bgr = cv2.imread(path)
        gray = cv2.resize(cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY), None, fx=0.5, fy=0.5, interpolation=cv2.INTER_AREA)
        prob = eng.forward(gray / 255.0)
        winners = np.argmax(prob, axis=3)[0]
        want = 0.0
        for cl in range(2):
            gt = cv2.resize(cv2.cvtColor(cv2.imread(str(tmp_path / "C2" / f"page{s}_GT{cl}.png")), cv2.COLOR_BGR2GRAY), None, fx=0.5, fy=0.5)
            want += np.sum((winners == cl).astype(np.float32) == gt / 255) / gt.size
        assert acc[s] == pytest.approx(want / 2, abs=1e-12)
        for cl in range(2):
            out = cv2.imread(str(tmp_path / "out" / f"page{s}_OUT{cl}.png"))
            assert out is not None and out.shape == (200, 150, 3)
    eng.close()
    # the single-channel maps (plot_with_img=False) are written as they are
    mine.plot_net_output(str(pb), str(lst), save_folder=str(tmp_path / "out"), rescale=1.0, calculate_accuracy=False)
    assert cv2.imread(str(tmp_path / "out" / "page0_OUT0.png"), cv2.IMREAD_UNCHANGED).shape == (400, 300)
