"""Import alias for the product package.

The product lives in ``citlab-article-separation-new_b200/`` (a directory name that is not a
valid Python identifier); this stub makes it importable as ``aru_b200`` by pointing the
package search path at that directory.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "citlab-article-separation-new_b200")
__path__.insert(0, _PKG_DIR)  # noqa: F821  (submodules resolve to files in the product dir)
PACKAGE_DIR = _PKG_DIR
