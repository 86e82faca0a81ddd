"""install() against the REAL reference classes (build container only: /root/reference is absent on the GPU box).

The reference's post-processors are imported from /root/reference under the inert stubs of
tests/golden/_reference_import.py (tensorflow, rasterio, shapely, lxml ... are missing here and unused by what is checked),
each scenario in a fresh interpreter because the imports are process-global.  Both orders the CLI can meet are covered:
classes imported before install() (rebinding of names held by the modules) and install() first, import later,
patch_separator_post_processor() after the import (run_net_post_processing.run_rank)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/article_separation"),
                                reason="the reference tree only exists in the build container")

PRELUDE = f"""
import sys
sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests', 'golden')!r})
import _reference_import as R
sys.path.insert(0, R.REFERENCE_ROOT)
sys.meta_path.insert(0, R._Finder())
from aru_b200 import net_boundary as nb
PKG = 'article_separation.image_segmentation.net_post_processing.'
"""

CHECK = """
import importlib
sep = importlib.import_module(PKG + 'separator_net_post_processor')
head = importlib.import_module(PKG + 'heading_net_post_processor')
base = importlib.import_module(PKG + 'region_net_post_processor_base')
tb = importlib.import_module(PKG + 'text_block_net_post_processor')
# every helper name a module imported (sep:7-8, head:6-7, base:10-11) is ours now
for m, names in ((sep, ('load_and_scale_image', 'get_net_output', 'apply_threshold')),
                 (head, ('load_and_scale_image', 'get_net_output')),
                 (base, ('load_image_paths', 'load_and_scale_image', 'load_graph', 'get_net_output', 'apply_threshold'))):
    for n in names:
        assert getattr(m, n) is getattr(nb, n), (m.__name__, n)
assert base.RegionNetPostProcessor.apply_cc_analysis is nb.apply_cc_analysis
assert sep.SeparatorNetPostProcessor.post_process is nb.separator_post_process
# subclasses resolve the rebound methods; the heading / text-block classes keep their own post_process
assert tb.TextBlockNetPostProcessor.apply_cc_analysis is nb.apply_cc_analysis
assert head.HeadingNetPostProcessor.apply_cc_analysis is nb.apply_cc_analysis
assert head.HeadingNetPostProcessor.get_swt_features_image is nb.heading_swt_features_image
assert importlib.import_module(nb.REFERENCE_MODULE) is nb
# the constructor of the real class goes through our load_graph and keeps our handle
import tempfile, os
with tempfile.NamedTemporaryFile(suffix='.pb', delete=False) as f:
    f.write(b'frozen graph bytes (parsed lazily, in the worker that owns the GPU)')
pp = sep.SeparatorNetPostProcessor(['a.png'], f.name, 1500, 1.0, 0.05, gpu_devices='')
assert isinstance(pp.pb_graph, nb.GraphHandle) and pp.pb_graph.path == f.name
os.unlink(f.name)
print('OK')
"""


def _run(body):
    r = subprocess.run([sys.executable, "-c", PRELUDE + body + CHECK], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr


def test_classes_imported_before_install():
    _run("""
import importlib
for n in ('separator_net_post_processor', 'heading_net_post_processor', 'text_block_net_post_processor'):
    importlib.import_module(PKG + n)
assert nb.install() is nb
""")


def test_install_before_the_classes_are_imported():
    # the order run_net_post_processing.run_rank uses: install(), import the class, then bind the device post-processing
    _run("""
assert nb.install() is nb
import importlib
importlib.import_module(PKG + 'separator_net_post_processor')
importlib.import_module(PKG + 'heading_net_post_processor')
assert nb.patch_separator_post_processor() is True
""")
