"""CPU tests of the drop-in boundary module (net_boundary.py): the helper functions restated from the reference's
net_post_processing_helper.py behave like it, and install() rebinds the names the reference's post-processors imported.
No GPU: nothing here creates an engine."""
import sys
import types

import numpy as np
import pytest

from aru_b200 import net_boundary

PKG = "article_separation.image_segmentation.net_post_processing"


def test_scaling_factor_matches_reference_values():
    # values produced by the reference's python_util.image_processing.image_stats.get_scaling_factor (:10-20)
    f = net_boundary._scaling_factor
    assert f(3000, 2000, 1.0, fixed_height=1500) == 0.5
    assert f(3000, 2000, 0.5, fixed_height=1500) == 0.25
    assert f(3000, 2000, 0.05, fixed_height=1500) == 0.5          # scaling factors <= 0.1 are ignored
    assert f(3000, 2000, None, fixed_height=1500) == 0.5
    assert f(3000, 2000, 1.0, fixed_width=1000) == 0.5
    assert f(3000, 2000, 0.7) == 0.7
    assert f(3000, 2000, None) is None
    assert f(800, 700, 1.0, fixed_height=900) == 1.125            # enlarging: INTER_CUBIC in the reference


def test_apply_threshold_and_paths(tmp_path):
    u8 = np.array([[0, 12, 13, 255]], np.uint8)
    assert net_boundary.apply_threshold(u8, 0.05).tolist() == [[0, 0, 255, 255]]          # u8 > 12.75
    p = np.array([[0.04, 0.05, 0.0501]], np.float32)
    assert net_boundary.apply_threshold(p, 0.05).tolist() == [[0, 0, 255]]
    lst = tmp_path / "images.lst"
    lst.write_text("a.png\nb c.jpg  \n")
    assert net_boundary.load_image_paths(str(lst)) == ["a.png", "b c.jpg"]


def test_scale_image_uses_area_when_shrinking():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(120, 90, 3), dtype=np.uint8)
    out, sc = net_boundary.scale_image(img, fixed_height=60, scaling_factor=1.0)
    assert sc == 0.5 and np.array_equal(out, cv2.resize(img, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_AREA))
    out, sc = net_boundary.scale_image(img, fixed_height=180, scaling_factor=1.0)
    assert sc == 1.5 and np.array_equal(out, cv2.resize(img, None, fx=1.5, fy=1.5, interpolation=cv2.INTER_CUBIC))
    out, sc = net_boundary.scale_image(img, fixed_height=120, scaling_factor=1.0)
    assert sc == 1.0 and out is img


def test_resolve_device(monkeypatch):
    monkeypatch.delenv("ARU_B200_DEVICE", raising=False)
    monkeypatch.delenv("LOCAL_RANK", raising=False)
    assert net_boundary.resolve_device("") == 0 and net_boundary.resolve_device(None) == 0
    assert net_boundary.resolve_device("2,3") == 2
    monkeypatch.setenv("LOCAL_RANK", "5")
    assert net_boundary.resolve_device("") == 5
    monkeypatch.setenv("ARU_B200_DEVICE", "1")
    assert net_boundary.resolve_device("") == 1


def test_graph_handle_pickles_without_engines(tmp_path):
    import pickle
    pb = tmp_path / "x.pb"
    pb.write_bytes(b"not a graph, never parsed here")
    g = net_boundary.load_graph(str(pb))
    g._engines[(1, 0)] = object()
    g2 = pickle.loads(pickle.dumps(g))
    assert g2.pb_bytes == g.pb_bytes and g2._engines == {}


def test_install_rebinds_the_reference_modules(monkeypatch):
    """The post-processors import the helper functions by name (separator_net_post_processor.py:7-8 ...): install() must
    replace the module under its dotted name and rebind the names in modules that were imported earlier."""
    sentinel = object()
    fake = {}
    for name in ("separator_net_post_processor", "heading_net_post_processor", "region_net_post_processor_base",
                 "text_block_net_post_processor"):
        m = types.ModuleType(f"{PKG}.{name}")
        for fn in ("load_graph", "get_net_output", "load_image_paths", "load_and_scale_image", "apply_threshold"):
            setattr(m, fn, sentinel)
        fake[name] = m
        monkeypatch.setitem(sys.modules, m.__name__, m)
    parent = types.ModuleType(PKG)
    monkeypatch.setitem(sys.modules, PKG, parent)
    monkeypatch.delitem(sys.modules, net_boundary.REFERENCE_MODULE, raising=False)

    class RegionNetPostProcessor:
        def apply_cc_analysis(self, net_output, threshold):
            raise AssertionError("the CPU version must have been replaced")

    class SeparatorNetPostProcessor(RegionNetPostProcessor):
        def post_process(self, net_output):
            raise AssertionError("the CPU version must have been replaced")

    fake["region_net_post_processor_base"].RegionNetPostProcessor = RegionNetPostProcessor
    fake["separator_net_post_processor"].SeparatorNetPostProcessor = SeparatorNetPostProcessor

    me = net_boundary.install()
    assert sys.modules[net_boundary.REFERENCE_MODULE] is me is net_boundary
    assert parent.net_post_processing_helper is net_boundary
    for m in fake.values():
        assert m.load_graph is net_boundary.load_graph and m.get_net_output is net_boundary.get_net_output
        assert m.apply_threshold is net_boundary.apply_threshold
    assert RegionNetPostProcessor.apply_cc_analysis is net_boundary.apply_cc_analysis
    assert SeparatorNetPostProcessor.post_process is net_boundary.separator_post_process
    import importlib
    assert importlib.import_module(net_boundary.REFERENCE_MODULE) is net_boundary


def test_textline_boxes_match_reference_polygon_code(tmp_path):
    """textline_box against the bounding boxes the reference's Polygon.rescale / get_bounding_box produced for the golden
    text lines (tests/golden/post_heading_lines.npz), and the stdlib PAGE-XML reader."""
    import os
    from aru_b200 import page_textlines as T
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "post_heading_lines.npz"))
    sc = float(z["scale"])
    for poly, bbox in zip(z["polygons"], z["bboxes"]):
        pts = list(zip(poly[:4].tolist(), poly[4:].tolist()))
        assert T.textline_box(pts, sc) == tuple(int(v) for v in bbox)
    xml = tmp_path / "page" / "scan_7.xml"
    xml.parent.mkdir()
    xml.write_text('<?xml version="1.0"?><PcGts xmlns="http://schema.primaresearch.org/PAGE/gts/pagecontent/2013-07-15">'
                   '<Page imageFilename="scan_7.png" imageWidth="100" imageHeight="80"><TextRegion id="r1">'
                   '<TextLine id="l1"><Coords points="10,20 90,22 95,60 12,58"/></TextLine>'
                   '<TextLine id="l2"/></TextRegion></Page></PcGts>')
    assert T.page_path_for_image(str(tmp_path / "scan_7.png")) == str(xml)
    lines = T.read_textlines(str(xml))
    assert lines == [("l1", [(10, 20), (90, 22), (95, 60), (12, 58)]), ("l2", None)]
    assert T.textline_box(lines[0][1], 0.5) == (5, 10, 43, 21)      # the values the reference printed for this polygon
    assert abs(T.net_prob(255 * 43 * 21, 43, 21) - 1.0) < 1e-15


def test_pinned_pool_size_classes():
    from aru_b200.engine import _PinnedPool
    b = _PinnedPool._bucket
    assert b(1) == 64 << 10 and b(64 << 10) == 64 << 10 and b((64 << 10) + 1) == 128 << 10
    for n in (1 << 20, (1 << 20) + 1, 13_200_000, 13_956_000, 1_610_612_736, 3 * (1 << 30) + 5):
        c = b(n)
        assert c >= n and c <= n * 1.125 + (64 << 10)
    # page widths 1088..1201 at height 1500 (float32, 2 classes) fall into a handful of classes, not one per width
    classes = {b(1500 * w * 2 * 4) for w in range(1088, 1202)}
    assert len(classes) <= 3
