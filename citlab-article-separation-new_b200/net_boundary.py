"""Drop-in for ``article_separation/image_segmentation/net_post_processing/net_post_processing_helper.py``.

Same names, argument meaning and return contract as the reference module, with the TensorFlow 1
session replaced by the B200 engine:

  * ``load_graph(path_to_pb)``                    reference :36-53  (GraphDef parse + import)
  * ``get_net_output(image, pb_graph, gpu_device)`` reference :56-72  (``sess.run('output:0', {'inImg:0': x})[0]``)
  * ``load_image_paths`` / ``scale_image`` / ``load_and_scale_image`` / ``apply_threshold``
                                                  reference :8-33, :75-78 (host-side helpers the callers import
                                                  from the same module; restated so the module can stand in whole)

``install()`` registers this module under the reference's dotted name, so the unmodified
``SeparatorNetPostProcessor`` / ``HeadingNetPostProcessor`` / ``RegionNetPostProcessor`` (which import the
functions *by name*, separator_net_post_processor.py:7-8, heading_net_post_processor.py:6-7,
region_net_post_processor_base.py:10-11) bind to the B200 engine.  See INTEGRATION.md.

Differences to the reference, all deliberate:
  * ``gpu_device == ''`` / ``None`` means "CPU" in the reference (it sets CUDA_VISIBLE_DEVICES=-1
    process-wide); there is no CPU path here, so it means "the GPU assigned to this rank"
    (``ARU_B200_DEVICE`` / ``LOCAL_RANK`` / 0) and the environment is never mutated;
  * the engine is created lazily in the calling process (CUDA must be initialised after
    ``ProcessPoolExecutor`` forked the worker) and kept for the life of the handle instead of building
    a session per page.
"""
from __future__ import annotations

import os
import sys
from typing import Optional

import numpy as np

from .engine import Engine

REFERENCE_MODULE = "article_separation.image_segmentation.net_post_processing.net_post_processing_helper"


class GraphHandle:
    """What ``load_graph`` returns in place of a ``tf.Graph``: the frozen GraphDef bytes plus a
    per-(process, device) engine cache."""

    def __init__(self, pb_bytes: bytes, path: Optional[str] = None):
        self.pb_bytes = pb_bytes
        self.path = path
        self._engines = {}

    def engine(self, device: int) -> Engine:
        key = (os.getpid(), device)
        eng = self._engines.get(key)
        if eng is None:
            eng = Engine(self.pb_bytes, device=device)   # raises KeyError when inImg:0 / output:0 are absent
            self._engines[key] = eng
        return eng

    def __getstate__(self):  # engines never cross process boundaries
        return {"pb_bytes": self.pb_bytes, "path": self.path, "_engines": {}}


def resolve_device(gpu_device) -> int:
    """Reference semantics: comma-separated CUDA ordinals, first one is used (helper.py:61-62)."""
    if gpu_device is None or str(gpu_device).strip() == "":
        for var in ("ARU_B200_DEVICE", "LOCAL_RANK"):
            if os.environ.get(var, "").strip():
                return int(os.environ[var])
        return 0
    return int(str(gpu_device).split(",")[0])


def load_image_paths(image_list):
    with open(image_list) as f:
        return [line.rstrip() for line in f.readlines()]


def _scaling_factor(image_height, image_width, scaling_factor, fixed_height=None, fixed_width=None):
    # python_util/image_processing/image_stats.py:10-20
    if fixed_height is not None and scaling_factor is not None and 0.1 < scaling_factor:
        return scaling_factor * fixed_height / image_height
    if fixed_width is not None and scaling_factor is not None and 0.1 < scaling_factor:
        return scaling_factor * fixed_width / image_width
    if fixed_height:
        return fixed_height / image_height
    if fixed_width:
        return fixed_width / image_width
    if scaling_factor:
        return scaling_factor
    return None


def scale_image(image, fixed_height=None, scaling_factor=1.0):
    import cv2
    image_height, image_width = image.shape[:2]
    sc = _scaling_factor(image_height, image_width, scaling_factor, fixed_height=fixed_height)
    if sc < 1.0:
        image = cv2.resize(image, None, fx=sc, fy=sc, interpolation=cv2.INTER_AREA)
    elif sc > 1.0:
        image = cv2.resize(image, None, fx=sc, fy=sc, interpolation=cv2.INTER_CUBIC)
    return image, sc


def load_and_scale_image(path_to_image, fixed_height, scaling_factor):
    import cv2
    image = cv2.imread(path_to_image)
    image, sc = scale_image(image, fixed_height, scaling_factor)
    image_grey = cv2.cvtColor(image, cv2.COLOR_BGR2GRAY) / 255.0
    return image, image_grey, sc


def load_graph(path_to_pb) -> GraphHandle:
    with open(path_to_pb, "rb") as f:
        return GraphHandle(f.read(), str(path_to_pb))


def get_net_output(image, pb_graph: GraphHandle, gpu_device="0"):
    """image: [H,W] (or [1,H,W,1]) float in [0,1]  ->  new float32 [H,W,C] array, channel 0 = class of interest."""
    image = np.asarray(image)
    if image.ndim == 4:
        if image.shape[0] != 1:
            raise ValueError("get_net_output takes one page; use Engine.forward for batches")
    elif image.ndim != 2:
        raise ValueError(f"expected a [H,W] or [1,H,W,1] image, got shape {image.shape}")
    eng = pb_graph.engine(resolve_device(gpu_device))
    return eng.forward(image)[0]


def get_net_output_batch(images, pb_graph: GraphHandle, gpu_device="0", **kw):
    """Batched extension: [N,H,W] -> float32 [N,H,W,C] (micro-batched and copy/compute-overlapped by the engine)."""
    return pb_graph.engine(resolve_device(gpu_device)).forward(images, **kw)


def apply_threshold(net_output, threshold):
    if net_output.dtype == np.uint8:
        threshold *= 255
    return np.array((net_output > threshold) * 255, dtype=np.uint8)


def separator_post_process(self, net_output):
    """Drop-in for ``SeparatorNetPostProcessor.post_process`` (separator_net_post_processor.py:25-99): the component
    size filter, the three rectangular openings and the subtract run on the GPU (csrc/post.cu, bit-exact against the
    reference).  ``self`` is the post-processor (its ``pb_graph`` is the ``GraphHandle`` from ``load_graph``);
    ``net_output`` is the thresholded uint8 [H,W,C] map.  Returns {"horizontal": mask, "vertical": mask}."""
    eng = self.pb_graph.engine(resolve_device(getattr(self, "gpu_devices", "")))
    horizontal, vertical = eng.separator_post(np.ascontiguousarray(net_output[:, :, 0]))
    return {"horizontal": horizontal, "vertical": vertical}


def apply_cc_analysis(self, net_output, threshold):
    """Drop-in for ``RegionNetPostProcessor.apply_cc_analysis`` (region_net_post_processor_base.py:230-251), used by the
    separator and the text-block post-processors (text_block_net_post_processor.py:22): connected components on the GPU.
    ``min_size = int(net_output.size * threshold)`` is evaluated here exactly as the reference does (Python floats)."""
    eng = self.pb_graph.engine(resolve_device(getattr(self, "gpu_devices", "")))
    min_size = int(net_output.size * threshold)
    return eng.cc_size_filter(np.ascontiguousarray(net_output), min_size).astype(net_output.dtype, copy=False)


def heading_swt_features_image(self, image_path):
    """Drop-in for ``HeadingNetPostProcessor.get_swt_features_image`` (heading_net_post_processor.py:215-220), i.e.
    ``StrokeWidthDistanceTransform.distance_transform`` (swt_dist_trafo.py:18-24), which the heading post-processor runs
    twice per page on the full-resolution scan (head:86,297): the decode stays ``cv2.imread(path, IMREAD_GRAYSCALE)`` as in
    the reference; inversion, 5x5 Gaussian, Otsu threshold and the exact Euclidean distance transform run on the GPU
    (csrc/swt.cu, bit-exact against the reference's function)."""
    import cv2
    image = cv2.imread(image_path, cv2.IMREAD_GRAYSCALE)
    eng = self.pb_graph.engine(resolve_device(getattr(self, "gpu_devices", "")))
    return eng.swt_distance(image, dark_on_bright=getattr(self.SWT, "_dark_on_bright", True))


def separator_pages(images_bgr, pb_graph: GraphHandle, threshold=0.05, gpu_device="0", **kw):
    """One call for ``SeparatorNetPostProcessor.run`` up to the polygon step (sep:141-151) on uint8 pages as
    ``cv2.imread`` / ``scale_image`` return them ([N,H,W,3] BGR, or [N,H,W] gray): see ``Engine.separator_pages``."""
    return pb_graph.engine(resolve_device(gpu_device)).separator_pages(images_bgr, threshold=threshold, **kw)


def separator_images(images_bgr, sc, pb_graph: GraphHandle, threshold=0.05, gpu_device="0", **kw):
    """``scale_image`` (helper.py:14-25, shrinking case) + ``separator_pages`` in one device call: unscaled uint8 images of
    one size in, results at the scaled size out; see ``Engine.separator_images``."""
    return pb_graph.engine(resolve_device(gpu_device)).separator_images(images_bgr, sc, threshold=threshold, **kw)


def heading_pages(images_bgr, boxes, pb_graph: GraphHandle, gpu_device="0", **kw):
    """``HeadingNetPostProcessor.run`` up to ``get_net_prob_for_text_line`` (head:247-291): see ``Engine.heading_pages``."""
    return pb_graph.engine(resolve_device(gpu_device)).heading_pages(images_bgr, boxes, **kw)


def install(patch_post_process: bool = True):
    """Make ``import article_separation....net_post_processing_helper`` resolve to this module.  With
    ``patch_post_process`` the reference's ``SeparatorNetPostProcessor.post_process`` (if that module is already imported,
    or as soon as ``patch_separator_post_processor()`` is called after importing it) runs on the GPU as well."""
    me = sys.modules[__name__]
    sys.modules[REFERENCE_MODULE] = me
    parent = sys.modules.get(REFERENCE_MODULE.rsplit(".", 1)[0])
    if parent is not None:
        setattr(parent, "net_post_processing_helper", me)
    # modules that were imported before install() hold the TF functions by name: rebind them
    for name in ("separator_net_post_processor", "heading_net_post_processor", "region_net_post_processor_base",
                 "text_block_net_post_processor"):
        mod = sys.modules.get(REFERENCE_MODULE.rsplit(".", 1)[0] + "." + name)
        if mod is None:
            continue
        for fn in ("load_graph", "get_net_output", "load_image_paths", "load_and_scale_image", "scale_image",
                   "apply_threshold"):
            if hasattr(mod, fn):
                setattr(mod, fn, getattr(me, fn))
    if patch_post_process:
        patch_separator_post_processor()
    return me


def patch_separator_post_processor() -> bool:
    """Rebind ``SeparatorNetPostProcessor.post_process``, ``RegionNetPostProcessor.apply_cc_analysis`` and
    ``HeadingNetPostProcessor.get_swt_features_image`` to the GPU versions; returns False when the reference classes have
    not been imported (call again after importing them)."""
    pkg = REFERENCE_MODULE.rsplit(".", 1)[0]
    base = sys.modules.get(pkg + ".region_net_post_processor_base")
    base_cls = getattr(base, "RegionNetPostProcessor", None) if base is not None else None
    if base_cls is not None:
        base_cls.apply_cc_analysis = apply_cc_analysis      # separator, text-block and any other region post-processor
    head = sys.modules.get(pkg + ".heading_net_post_processor")
    head_cls = getattr(head, "HeadingNetPostProcessor", None) if head is not None else None
    if head_cls is not None:
        head_cls.get_swt_features_image = heading_swt_features_image
    mod = sys.modules.get(pkg + ".separator_net_post_processor")
    cls = getattr(mod, "SeparatorNetPostProcessor", None) if mod is not None else None
    if cls is None:
        return base_cls is not None
    cls.post_process = separator_post_process
    return True
