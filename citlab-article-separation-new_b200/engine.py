"""ctypes binding of libaru_b200.so (include/aru_b200.h) - the B200 ARU-Net forward engine.

This is the host side of the drop-in boundary: ``Engine`` is what replaces the ``tf.Graph`` +
``tf.Session`` pair of net_post_processing_helper.py:36-72.  There is no CPU path: if the shared
library is missing or no sm_100 GPU is present, construction raises.
"""
from __future__ import annotations

import ctypes
import os
import threading
import weakref
from typing import Optional, Tuple, Union

import numpy as np

from .graphdef import parse_graphdef
from .program import CGraphDesc, Program, lower_graph

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# ARU_B200_LIB selects another build of the same library (the bf16 variant the parity suite also runs)
LIB_PATH = os.environ.get("ARU_B200_LIB") or os.path.join(_PKG_DIR, "libaru_b200.so")

ARU_OK, ARU_EINVAL, ARU_ECUDA, ARU_ENOMEM, ARU_EUNSUP, ARU_ENODEV = 0, 1, 2, 3, 4, 5
OPT_CONV_PATH, OPT_USE_GRAPH, OPT_MICRO_BATCH, OPT_KEEP_ALL, OPT_FUSE_PAIRS, OPT_U8_CHANNELS, OPT_ASYNC = 1, 2, 3, 4, 5, 6, 7
OPT_FUSE_BLOCKS = 8
OPT_BRANCH_STREAMS = 9

# every symbol include/aru_b200.h declares (checked by tests/test_cabi.py)
EXPORTS = [
    "aru_abi_version", "aru_device_count", "aru_create", "aru_destroy", "aru_set_option", "aru_num_classes",
    "aru_plan", "aru_forward", "aru_forward_device", "aru_sync", "aru_launches_per_forward", "aru_read_buffer",
    "aru_buffer_dims", "aru_profile_ops", "aru_op_kernel_name", "aru_last_error", "aru_host_alloc", "aru_host_free",
    "aru_separator_pages", "aru_separator_post", "aru_open_rect", "aru_pages_to_input", "aru_heading_pages",
    "aru_box_sums", "aru_cc_filter", "aru_scaled_size", "aru_scale_pages", "aru_separator_images", "aru_heading_images",
    "aru_bind_host_to_device", "aru_swt_distance", "aru_last_ticket", "aru_wait", "aru_last_warning",
    "aru_f64_to_f32",
]

_lib = None
_lib_lock = threading.Lock()


class EngineError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"aru_b200 error {code}: {message}")
        self.code = code


def load_library() -> ctypes.CDLL:
    """dlopen the engine; fails loudly when it has not been built (``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with __graft_entry__.build(); "
                               "the B200 engine has no CPU fallback")
        lib = ctypes.CDLL(LIB_PATH)
        vp, i32, i64, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t
        fp = ctypes.POINTER(ctypes.c_float)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        lib.aru_abi_version.restype = i32
        lib.aru_device_count.restype = i32
        lib.aru_create.argtypes = [ctypes.POINTER(CGraphDesc), i32, ctypes.POINTER(vp)]
        lib.aru_create.restype = i32
        lib.aru_destroy.argtypes = [vp]
        lib.aru_destroy.restype = None
        lib.aru_set_option.argtypes = [vp, i32, i64]
        lib.aru_num_classes.argtypes = [vp]
        lib.aru_plan.argtypes = [vp, i32, i32, i32]
        lib.aru_forward.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, ctypes.c_float]
        lib.aru_forward_device.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, ctypes.c_float, vp]
        lib.aru_separator_pages.argtypes = [vp, vp, i32, i32, i32, i32, ctypes.c_double, vp, vp, vp, vp, vp]
        lib.aru_separator_post.argtypes = [vp, vp, i32, i32, i32, vp, vp]
        lib.aru_open_rect.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp]
        lib.aru_pages_to_input.argtypes = [vp, vp, i32, i32, i32, i32, vp]
        lib.aru_heading_pages.argtypes = [vp, vp, i32, i32, i32, i32, vp, i32, vp, vp]
        lib.aru_box_sums.argtypes = [vp, vp, i32, i32, i32, i32, vp, i32, vp]
        lib.aru_cc_filter.argtypes = [vp, vp, i32, i32, i32, i32, vp]
        lib.aru_scaled_size.argtypes = [i32, i32, ctypes.c_double, ctypes.POINTER(i32), ctypes.POINTER(i32)]
        lib.aru_scale_pages.argtypes = [vp, vp, i32, i32, i32, i32, ctypes.c_double, vp]
        lib.aru_separator_images.argtypes = [vp, vp, i32, i32, i32, i32, ctypes.c_double, ctypes.c_double, vp, vp, vp, vp, vp]
        lib.aru_heading_images.argtypes = [vp, vp, i32, i32, i32, i32, ctypes.c_double, vp, i32, vp, vp]
        lib.aru_sync.argtypes = [vp]
        lib.aru_launches_per_forward.argtypes = [vp]
        lib.aru_read_buffer.argtypes = [vp, i32, i32, fp, sz]
        lib.aru_buffer_dims.argtypes = [vp, i32, ctypes.POINTER(i32), ctypes.POINTER(i32), ctypes.POINTER(i32)]
        lib.aru_profile_ops.argtypes = [vp, i32, fp, i32]
        lib.aru_op_kernel_name.argtypes = [vp, i32]
        lib.aru_op_kernel_name.restype = ctypes.c_char_p
        lib.aru_last_error.argtypes = [vp]
        lib.aru_last_error.restype = ctypes.c_char_p
        lib.aru_host_alloc.argtypes = [ctypes.POINTER(vp), sz]
        lib.aru_host_free.argtypes = [vp]
        lib.aru_host_free.restype = None
        lib.aru_bind_host_to_device.argtypes = [i32, ctypes.POINTER(i32)]
        lib.aru_f64_to_f32.argtypes = [vp, vp, ctypes.c_longlong, i32]
        lib.aru_swt_distance.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp]
        lib.aru_last_ticket.argtypes = [vp]
        lib.aru_last_ticket.restype = ctypes.c_uint64
        lib.aru_wait.argtypes = [vp, ctypes.c_uint64]
        lib.aru_last_warning.argtypes = [vp]
        lib.aru_last_warning.restype = ctypes.c_char_p
        del u8p
        _lib = lib
        return lib


# ------------------------------------------------------------------------------------------------
# pinned host arrays
# ------------------------------------------------------------------------------------------------
class _PinnedPool:
    """Recycles page-locked host blocks; numpy arrays handed out return their block on garbage collection."""

    def __init__(self, max_cached_bytes: int = 8 << 30):
        self.free = {}
        self.cached = 0
        self.max_cached = max_cached_bytes
        self.lock = threading.Lock()

    def _release(self, ptr: int, nbytes: int):
        with self.lock:
            if self.cached + nbytes <= self.max_cached:
                self.free.setdefault(nbytes, []).append(ptr)
                self.cached += nbytes
                return
        load_library().aru_host_free(ctypes.c_void_p(ptr))

    @staticmethod
    def _bucket(nbytes: int) -> int:
        """Size class of a block: multiples of 64 KB up to 1 MB, then eighths of a power of two (<= 12.5 % slack).
        Page-locking costs ~0.25 ms per MB, and pages scaled to a fixed height differ slightly in width from scan to
        scan - exact-size blocks would never be reused in such a stream."""
        if nbytes <= (1 << 20):
            return -(-nbytes // (64 << 10)) * (64 << 10)
        step = (1 << (nbytes.bit_length() - 1)) >> 3
        return -(-nbytes // step) * step

    def empty(self, shape, dtype) -> np.ndarray:
        dtype = np.dtype(dtype)
        nbytes = self._bucket(max(int(np.prod(shape)) * dtype.itemsize, 1))
        ptr = None
        with self.lock:
            lst = self.free.get(nbytes)
            if lst:
                ptr = lst.pop()
                self.cached -= nbytes
        if ptr is None:
            p = ctypes.c_void_p()
            rc = load_library().aru_host_alloc(ctypes.byref(p), nbytes)
            if rc != ARU_OK:
                raise EngineError(rc, "pinned host allocation failed")
            ptr = p.value
        buf = (ctypes.c_byte * nbytes).from_address(ptr)
        weakref.finalize(buf, self._release, ptr, nbytes)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


PINNED = _PinnedPool()


def pinned_empty(shape, dtype=np.float32) -> np.ndarray:
    return PINNED.empty(shape, dtype)


def bind_host_to_device(device: int) -> int:
    """Pin this thread to the CPUs next to CUDA device ``device`` and prefer its NUMA node for new (page-locked) memory;
    returns the node (-1 when unknown / not applied).  ``Engine()`` does this itself unless ``ARU_NUMA_BIND=0``; call it
    first when pinned buffers are allocated before the engine exists."""
    node = ctypes.c_int(-1)
    rc = load_library().aru_bind_host_to_device(int(device), ctypes.byref(node))
    return node.value if rc == ARU_OK else -1


# ------------------------------------------------------------------------------------------------
# engine
# ------------------------------------------------------------------------------------------------
class Engine:
    """One frozen graph on one GPU.  ``forward`` = the reference's ``sess.run(out, {x: image})``."""

    def __init__(self, pb: Union[bytes, str, os.PathLike], device: int = 0, in_name: str = "inImg",
                 out_name: str = "output"):
        if not isinstance(pb, (bytes, bytearray)):
            with open(pb, "rb") as f:
                pb = f.read()
        self.program: Program = lower_graph(parse_graphdef(bytes(pb)), in_name, out_name)
        self.lib = load_library()
        if self.lib.aru_abi_version() != 1:
            raise RuntimeError("libaru_b200.so ABI version mismatch")
        desc, keep = self.program.desc()
        handle = ctypes.c_void_p()
        rc = self.lib.aru_create(ctypes.byref(desc), int(device), ctypes.byref(handle))
        del keep
        if rc != ARU_OK:
            raise EngineError(rc, (self.lib.aru_last_error(None) or b"").decode())
        self.handle = handle
        self.device = int(device)
        self.n_class = self.lib.aru_num_classes(handle)
        self._in_flight = {}
        self._staged, self._collect = [], False
        self._finalizer = weakref.finalize(self, self.lib.aru_destroy, handle)

    # -- helpers ---------------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != ARU_OK:
            raise EngineError(rc, (self.lib.aru_last_error(self.handle) or b"").decode())

    def close(self):
        self._finalizer()

    @property
    def last_warning(self) -> str:
        """Result of the dynamic-range check of the engine's first pass ('' = every stored activation was in range)."""
        return (self.lib.aru_last_warning(self.handle) or b"").decode()

    def set_option(self, option: int, value: int):
        self._check(self.lib.aru_set_option(self.handle, option, int(value)))

    def plan(self, n: int, h: int, w: int):
        self._check(self.lib.aru_plan(self.handle, n, h, w))

    @property
    def launches_per_forward(self) -> int:
        return self.lib.aru_launches_per_forward(self.handle)

    # -- forward ---------------------------------------------------------------------------------
    @staticmethod
    def _as_batch(images: np.ndarray) -> np.ndarray:
        x = np.asarray(images)
        if x.ndim == 2:
            x = x[None]
        elif x.ndim == 4:
            if x.shape[-1] != 1:
                raise ValueError(f"expected a 1-channel image batch, got shape {x.shape}")
            x = x[..., 0]
        elif x.ndim != 3:
            raise ValueError(f"expected [H,W], [N,H,W] or [N,H,W,1], got shape {x.shape}")
        return x

    def _u8_channels(self, k: int) -> int:
        """Select how many leading channels the uint8 outputs hold (0 = all); returns the resulting channel count."""
        k = int(k)
        if k < 0 or k > self.n_class:
            raise ValueError(f"u8_channels must be 0..{self.n_class}")
        self.set_option(OPT_U8_CHANNELS, k)
        return k if 0 < k < self.n_class else self.n_class

    def forward(self, images: np.ndarray, want_u8: bool = False, want_mask: bool = False, threshold: float = 0.05,
                want_prob: bool = True, u8_channels: int = 0):
        """images: float [N,H,W] in [0,1] (any float dtype; pinned float32 is zero-copy).
        Returns prob float32 [N,H,W,C] (pinned), plus uint8 / mask arrays when requested.  ``u8_channels=1`` returns only
        channel 0 of the uint8 map ([N,H,W,1]) - the one every consumer reads (sep:33, head:209)."""
        x = self._as_batch(images)
        n, h, w = x.shape
        if not (x.dtype == np.float32 and x.flags.c_contiguous):
            xin = pinned_empty((n, h, w), np.float32)
            if x.dtype == np.float64 and x.flags.c_contiguous:   # what get_net_output's callers pass (helper.py:31)
                self._check(self.lib.aru_f64_to_f32(ctypes.c_void_p(x.ctypes.data), ctypes.c_void_p(xin.ctypes.data), x.size, 0))
            else:
                np.copyto(xin, x, casting="unsafe")
            x = xin
        if self._collect:
            self._staged.append(x)
        c = self.n_class
        prob = pinned_empty((n, h, w, c), np.float32) if want_prob else None
        cu = self._u8_channels(u8_channels)
        u8 = pinned_empty((n, h, w, cu), np.uint8) if want_u8 else None
        mask = pinned_empty((n, h, w), np.uint8) if want_mask else None
        ptr = lambda a: ctypes.c_void_p(a.ctypes.data) if a is not None else None  # noqa: E731
        self._check(self.lib.aru_forward(self.handle, ptr(x), n, h, w, ptr(prob), ptr(u8), ptr(mask),
                                         ctypes.c_float(threshold)))
        outs = tuple(a for a in (prob, u8, mask) if a is not None)
        return outs[0] if len(outs) == 1 else outs

    def forward_device(self, in_ptr: int, n: int, h: int, w: int, out_ptr: int = 0, u8_ptr: int = 0, mask_ptr: int = 0,
                       threshold: float = 0.05, stream: int = 0):
        """Device-resident variant (raw device pointers, e.g. torch ``tensor.data_ptr()``); only enqueues."""
        vp = lambda v: ctypes.c_void_p(v) if v else None  # noqa: E731
        self._check(self.lib.aru_forward_device(self.handle, vp(in_ptr), n, h, w, vp(out_ptr), vp(u8_ptr), vp(mask_ptr),
                                                ctypes.c_float(threshold), vp(stream)))

    def sync(self):
        self._check(self.lib.aru_sync(self.handle))

    # -- asynchronous host-buffer calls ----------------------------------------------------------------------------
    def submit(self, call, *args, **kw):
        """Run ``call`` (``self.forward``, ``self.separator_pages`` or ``self.separator_images``) without waiting for it:
        returns ``(ticket, result)``; the result arrays are filled when ``wait(ticket)`` returns.  Inputs that are not
        already pinned float32 / uint8 arrays are staged, so the caller's arrays can be reused at once; a pinned input
        must stay untouched until ``wait``.  Up to 8 calls may be in flight; consecutive calls overlap (the copy-in of
        the next runs under the tail of the previous)."""
        self._staged, self._collect = [], True
        self.set_option(OPT_ASYNC, 1)
        try:
            res = call(*args, **kw)
        finally:
            self._collect = False
            self.set_option(OPT_ASYNC, 0)              # calls in flight stay in flight
        ticket = int(self.lib.aru_last_ticket(self.handle))
        self._in_flight[ticket] = (res, args, self._staged)   # keeps every buffer of the call alive until wait()
        self._staged = []
        return ticket, res

    def wait(self, ticket: int):
        self._check(self.lib.aru_wait(self.handle, int(ticket)))
        self._in_flight.pop(int(ticket), None)

    # -- integer pre / post-processing on the device ------------------------
    @staticmethod
    def _as_pages(pages: np.ndarray):
        x = np.asarray(pages)
        if x.dtype != np.uint8:
            raise ValueError(f"pages must be uint8 (as cv2.imread returns them), got {x.dtype}")
        if x.ndim == 2:
            x = x[None]
        if x.ndim == 3 and x.shape[-1] == 3 and x.shape[0] != 3:   # one [H,W,3] BGR page
            x = x[None]
        if x.ndim == 3:
            n, h, w = x.shape
            ch = 1
        elif x.ndim == 4 and x.shape[-1] in (1, 3):
            n, h, w, ch = x.shape
        else:
            raise ValueError(f"expected uint8 [H,W], [N,H,W], [H,W,3] or [N,H,W,{{1,3}}], got shape {x.shape}")
        return x, n, h, w, ch

    def separator_pages(self, pages: np.ndarray, threshold: float = 0.05, want_prob: bool = False,
                        want_u8: bool = False, want_mask: bool = False, want_separators: bool = True,
                        u8_channels: int = 0) -> dict:
        """One iteration of ``SeparatorNetPostProcessor.run`` up to the polygon step for a batch of uint8 pages
        (gray ``[N,H,W]`` or BGR ``[N,H,W,3]``): colour step, net, ``uint8(p*255)``, threshold, ``post_process``
        (separator_net_post_processor.py:141-151).  Returns a dict with the requested arrays (pinned host memory):
        ``prob`` float32 [N,H,W,C], ``u8`` uint8 [N,H,W,C], ``mask`` uint8 [N,H,W], ``horizontal`` / ``vertical``
        uint8 [N,H,W] in {0,255}."""
        x, n, h, w, ch = self._as_pages(pages)
        if not x.flags.c_contiguous:
            xin = pinned_empty(x.shape, np.uint8)
            np.copyto(xin, x)
            x = xin
        if self._collect:
            self._staged.append(x)
        c = self.n_class
        res = {}
        if want_prob:
            res["prob"] = pinned_empty((n, h, w, c), np.float32)
        cu = self._u8_channels(u8_channels)
        if want_u8:
            res["u8"] = pinned_empty((n, h, w, cu), np.uint8)
        if want_mask:
            res["mask"] = pinned_empty((n, h, w), np.uint8)
        if want_separators:
            res["horizontal"] = pinned_empty((n, h, w), np.uint8)
            res["vertical"] = pinned_empty((n, h, w), np.uint8)
        ptr = lambda k: ctypes.c_void_p(res[k].ctypes.data) if k in res else None  # noqa: E731
        self._check(self.lib.aru_separator_pages(self.handle, ctypes.c_void_p(x.ctypes.data), ch, n, h, w,
                                                 ctypes.c_double(threshold), ptr("prob"), ptr("u8"), ptr("mask"),
                                                 ptr("horizontal"), ptr("vertical")))
        return res

    def scaled_size(self, h: int, w: int, sc: float):
        """(cvRound(h*sc), cvRound(w*sc)): the size ``cv2.resize(image, None, fx=sc, fy=sc)`` produces."""
        hh, ww = ctypes.c_int(), ctypes.c_int()
        if self.lib.aru_scaled_size(h, w, ctypes.c_double(sc), ctypes.byref(hh), ctypes.byref(ww)) != ARU_OK:
            raise ValueError(f"bad size / scale {h}x{w} * {sc}")
        return hh.value, ww.value

    def scale_pages(self, pages: np.ndarray, sc: float) -> np.ndarray:
        """``scale_image`` (helper.py:14-25) of uint8 pages [N,H,W] / [N,H,W,3] (or one page) on the device: sc < 1 is
        ``cv2.resize(..., INTER_AREA)`` bit-exact against OpenCV; sc > 1 is ``INTER_CUBIC`` within one grey level of
        OpenCV (its scalar fixed-point form)."""
        x, n, h, w, ch = self._as_pages(pages)
        x = np.ascontiguousarray(x)
        dh, dw = self.scaled_size(h, w, sc)
        out = np.empty((n, dh, dw) + ((ch,) if x.ndim == 4 else ()), np.uint8)
        self._check(self.lib.aru_scale_pages(self.handle, ctypes.c_void_p(x.ctypes.data), ch, n, h, w, ctypes.c_double(sc),
                                             ctypes.c_void_p(out.ctypes.data)))
        return out

    def separator_images(self, images: np.ndarray, sc: float, threshold: float = 0.05, want_prob: bool = False,
                         want_u8: bool = False, want_mask: bool = False, want_separators: bool = True,
                         u8_channels: int = 0) -> dict:
        """``load_and_scale_image`` after the decode + ``SeparatorNetPostProcessor.run`` up to the polygon step: the
        unscaled uint8 images (as ``cv2.imread`` returns them, all of one size) go up, are shrunk by ``sc`` on the device
        (INTER_AREA; ``sc == 1`` is a no-op as in the reference; ``sc > 1`` is INTER_CUBIC within one grey level), and the
        results come back at the scaled size.  Same dict as ``separator_pages``."""
        cu = self._u8_channels(u8_channels)
        x, n, h, w, ch = self._as_pages(images)
        x = np.ascontiguousarray(x)
        if self._collect:
            self._staged.append(x)
        dh, dw = (h, w) if sc == 1.0 else self.scaled_size(h, w, sc)
        c = self.n_class
        res = {}
        if want_prob:
            res["prob"] = pinned_empty((n, dh, dw, c), np.float32)
        if want_u8:
            res["u8"] = pinned_empty((n, dh, dw, cu), np.uint8)
        if want_mask:
            res["mask"] = pinned_empty((n, dh, dw), np.uint8)
        if want_separators:
            res["horizontal"] = pinned_empty((n, dh, dw), np.uint8)
            res["vertical"] = pinned_empty((n, dh, dw), np.uint8)
        ptr = lambda k: ctypes.c_void_p(res[k].ctypes.data) if k in res else None  # noqa: E731
        self._check(self.lib.aru_separator_images(self.handle, ctypes.c_void_p(x.ctypes.data), ch, n, h, w,
                                                  ctypes.c_double(sc), ctypes.c_double(threshold), ptr("prob"), ptr("u8"),
                                                  ptr("mask"), ptr("horizontal"), ptr("vertical")))
        return res

    def separator_post(self, mask: np.ndarray):
        """``SeparatorNetPostProcessor.post_process`` (sep:25-99) on thresholded masks uint8 [N,H,W] (or [H,W]):
        returns (horizontal, vertical) uint8 {0,255} arrays of the same shape."""
        m = np.ascontiguousarray(mask, dtype=np.uint8)
        single = m.ndim == 2
        if single:
            m = m[None]
        n, h, w = m.shape
        hor, ver = np.empty_like(m), np.empty_like(m)
        self._check(self.lib.aru_separator_post(self.handle, ctypes.c_void_p(m.ctypes.data), n, h, w,
                                                ctypes.c_void_p(hor.ctypes.data), ctypes.c_void_p(ver.ctypes.data)))
        return (hor[0], ver[0]) if single else (hor, ver)

    @staticmethod
    def clip_boxes(boxes, h: int, w: int) -> np.ndarray:
        """[(page, ya, yb, xa, xb)] with the reference's numpy-slice semantics (``net_output[ya:yb, xa:xb]``,
        heading_net_post_processor.py:266: negative indices count from the end, ranges are clipped, an inverted range is
        empty) -> int32 [n_boxes, 5] half-open boxes inside the page."""
        out = np.zeros((len(boxes), 5), np.int32)
        for i, (pg, ya, yb, xa, xb) in enumerate(boxes):
            y0, y1, _ = slice(int(ya), int(yb)).indices(h)
            x0, x1, _ = slice(int(xa), int(xb)).indices(w)
            out[i] = (int(pg), y0, max(y0, y1), x0, max(x0, x1))
        return out

    def heading_pages(self, pages: np.ndarray, boxes, want_u8: bool = False):
        """One iteration of ``HeadingNetPostProcessor.run`` up to the network feature of every text line
        (heading_net_post_processor.py:280-291, 247-270): colour step, net, ``uint8(p*255)`` and the sum of channel 0 over
        ``net_output[ya:yb, xa:xb]`` for every box ``(page, ya, yb, xa, xb)`` - all on the device.
        Returns (sums uint64 [n_boxes], clipped boxes int32 [n_boxes,5][, u8]); the reference's
        ``get_net_prob_for_text_line`` value is ``sums / 255 / (bounding_box.width * bounding_box.height)``."""
        self._u8_channels(0)
        x, n, h, w, ch = self._as_pages(pages)
        x = np.ascontiguousarray(x)
        bx = self.clip_boxes(boxes, h, w)
        sums = np.zeros(len(bx), np.uint64)
        u8 = pinned_empty((n, h, w, self.n_class), np.uint8) if want_u8 else None
        self._check(self.lib.aru_heading_pages(self.handle, ctypes.c_void_p(x.ctypes.data), ch, n, h, w,
                                               ctypes.c_void_p(bx.ctypes.data), len(bx), ctypes.c_void_p(sums.ctypes.data),
                                               ctypes.c_void_p(u8.ctypes.data) if u8 is not None else None))
        return (sums, bx, u8) if want_u8 else (sums, bx)

    def heading_images(self, images: np.ndarray, sc: float, boxes, want_u8: bool = False):
        """``heading_pages`` with ``scale_image`` in front, on the device: unscaled uint8 images in;
        ``boxes`` are in the coordinates of the scaled page (the reference rescales the text-line polygons by the same
        factor, heading_net_post_processor.py:262-263)."""
        self._u8_channels(0)
        x, n, h, w, ch = self._as_pages(images)
        x = np.ascontiguousarray(x)
        dh, dw = (h, w) if sc == 1.0 else self.scaled_size(h, w, sc)
        bx = self.clip_boxes(boxes, dh, dw)
        sums = np.zeros(len(bx), np.uint64)
        u8 = pinned_empty((n, dh, dw, self.n_class), np.uint8) if want_u8 else None
        self._check(self.lib.aru_heading_images(self.handle, ctypes.c_void_p(x.ctypes.data), ch, n, h, w,
                                                ctypes.c_double(sc), ctypes.c_void_p(bx.ctypes.data), len(bx),
                                                ctypes.c_void_p(sums.ctypes.data),
                                                ctypes.c_void_p(u8.ctypes.data) if u8 is not None else None))
        return (sums, bx, u8) if want_u8 else (sums, bx)

    def swt_distance(self, gray: np.ndarray, dark_on_bright: bool = True, return_thresholds: bool = False):
        """``StrokeWidthDistanceTransform.distance_transform`` (swt_dist_trafo.py:18-29) of uint8 gray pages ``[H,W]`` /
        ``[N,H,W]`` on the device: invert, 5x5 Gaussian, Otsu, exact Euclidean distance transform, uint8 truncation -
        bit-exact against the reference (which decodes the file itself: pass ``cv2.imread(path, cv2.IMREAD_GRAYSCALE)``)."""
        x = np.asarray(gray)
        if x.dtype != np.uint8 or x.ndim not in (2, 3):
            raise ValueError(f"expected uint8 [H,W] or [N,H,W] gray pages, got {x.dtype} {x.shape}")
        single = x.ndim == 2
        x = np.ascontiguousarray(x[None] if single else x)
        n, h, w = x.shape
        out = np.empty((n, h, w), np.uint8)
        thr = np.zeros(n, np.int32)
        self._check(self.lib.aru_swt_distance(self.handle, ctypes.c_void_p(x.ctypes.data), n, h, w, int(bool(dark_on_bright)),
                                              ctypes.c_void_p(out.ctypes.data), ctypes.c_void_p(thr.ctypes.data)))
        res = out[0] if single else out
        return (res, thr) if return_thresholds else res

    def box_sums(self, u8: np.ndarray, boxes) -> np.ndarray:
        """Sum of channel 0 of a host uint8 map [N,H,W,C] (or [H,W,C] / [H,W]) over numpy-slice boxes."""
        m = np.ascontiguousarray(u8, dtype=np.uint8)
        if m.ndim == 2:
            m = m[None, :, :, None]
        elif m.ndim == 3:
            m = m[None]
        n, h, w, c = m.shape
        bx = self.clip_boxes(boxes, h, w)
        sums = np.zeros(len(bx), np.uint64)
        self._check(self.lib.aru_box_sums(self.handle, ctypes.c_void_p(m.ctypes.data), n, h, w, c,
                                          ctypes.c_void_p(bx.ctypes.data), len(bx), ctypes.c_void_p(sums.ctypes.data)))
        return sums

    def cc_size_filter(self, mask: np.ndarray, min_size: int) -> np.ndarray:
        """``RegionNetPostProcessor.apply_cc_analysis`` (region_net_post_processor_base.py:230-251): keep the 8-connected
        components of the non-zero pixels of uint8 masks [N,H,W] (or [H,W]) with area >= min_size; output {0,255}."""
        m = np.ascontiguousarray(mask, dtype=np.uint8)
        single = m.ndim == 2
        if single:
            m = m[None]
        n, h, w = m.shape
        out = np.empty_like(m)
        self._check(self.lib.aru_cc_filter(self.handle, ctypes.c_void_p(m.ctypes.data), n, h, w, int(min_size),
                                           ctypes.c_void_p(out.ctypes.data)))
        return out[0] if single else out

    def open_rect(self, mask: np.ndarray, kw: int, kh: int) -> np.ndarray:
        """``cv2.morphologyEx(mask, MORPH_OPEN, RECT(kw, kh))`` for binary masks and kw == 1 or kh == 1."""
        m = np.ascontiguousarray(mask, dtype=np.uint8)
        single = m.ndim == 2
        if single:
            m = m[None]
        n, h, w = m.shape
        out = np.empty_like(m)
        self._check(self.lib.aru_open_rect(self.handle, ctypes.c_void_p(m.ctypes.data), n, h, w, int(kw), int(kh),
                                           ctypes.c_void_p(out.ctypes.data)))
        return out[0] if single else out

    def pages_to_input(self, pages: np.ndarray) -> np.ndarray:
        """``cv2.cvtColor(image, COLOR_BGR2GRAY) / 255.0`` (helper.py:31) as the float32 [N,H,W] the net receives."""
        x, n, h, w, ch = self._as_pages(pages)
        x = np.ascontiguousarray(x)
        out = np.empty((n, h, w), np.float32)
        self._check(self.lib.aru_pages_to_input(self.handle, ctypes.c_void_p(x.ctypes.data), ch, n, h, w,
                                                ctypes.c_void_p(out.ctypes.data)))
        return out

    # -- introspection -----------------------------------------------------------------------------
    def read_buffer(self, buf: int, page: int = 0) -> np.ndarray:
        h, w, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        self._check(self.lib.aru_buffer_dims(self.handle, buf, ctypes.byref(h), ctypes.byref(w), ctypes.byref(c)))
        out = np.empty((h.value, w.value, c.value), np.float32)
        self._check(self.lib.aru_read_buffer(self.handle, buf, page, out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                             out.size))
        return out

    def read_node(self, node_name: str, page: int = 0) -> np.ndarray:
        """Value of a GraphDef node that survived lowering (per-layer parity tests)."""
        v = self.program.tensor_of_node[node_name]
        full = self.read_buffer(v.buf, page)
        return full[..., v.ch_off:v.ch_off + v.ch]

    def profile_ops(self, iters: int = 5):
        """[(op name, kernel label, ms)] for the current plan (CUDA events, plain launches)."""
        n = len(self.program.ops)
        ms = (ctypes.c_float * n)()
        self._check(self.lib.aru_profile_ops(self.handle, iters, ms, n))
        return [(self.program.ops[i].name, (self.lib.aru_op_kernel_name(self.handle, i) or b"").decode(), float(ms[i]))
                for i in range(n)]
