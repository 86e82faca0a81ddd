"""Import the reference's own post-processors in the build container (golden-vector generation only).

``/root/reference`` is importable here but several of its module-level imports (tensorflow 1.x, rasterio, shapely,
lxml, matplotlib, cssutils ...) are absent.  None of them is touched by the functions the fixtures are generated from
(``SeparatorNetPostProcessor.post_process`` / ``apply_cc_analysis`` run on numpy + cv2 only), so they are replaced by
inert stub modules.  Never used at test or run time: the GPU box has no /root/reference.
"""
import importlib.abc
import importlib.machinery
import sys
import types
from unittest import mock

REFERENCE_ROOT = "/root/reference"
STUB_ROOTS = ("tensorflow", "rasterio", "shapely", "lxml", "matplotlib", "skimage", "networkx", "cssutils", "jpype",
              "Levenshtein", "editdistance", "colour")


class _Stub(types.ModuleType):
    def __getattr__(self, key):
        if key.startswith("__"):
            raise AttributeError(key)
        m = mock.MagicMock(name=self.__name__ + "." + key)
        setattr(self, key, m)
        return m


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in STUB_ROOTS:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _Stub(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def reference_separator_post_processor():
    """An instance of the reference's SeparatorNetPostProcessor without running its __init__ (which loads a .pb)."""
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if not any(isinstance(f, _Finder) for f in sys.meta_path):
        sys.meta_path.insert(0, _Finder())
    from article_separation.image_segmentation.net_post_processing.separator_net_post_processor import \
        SeparatorNetPostProcessor
    return object.__new__(SeparatorNetPostProcessor)
