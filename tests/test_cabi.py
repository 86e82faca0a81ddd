"""The C-ABI library: loads, exports every symbol include/aru_b200.h declares, and refuses to run
without a GPU (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "aru_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(aru_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_list_agree():
    from aru_b200 import engine
    assert _declared_functions() == sorted(engine.EXPORTS)


def test_library_exports_every_declared_symbol(built_lib):
    for name in _declared_functions():
        assert hasattr(built_lib, name), name
    assert built_lib.aru_abi_version() == 1


def test_struct_layouts_match_header(built_lib):
    from aru_b200.program import CBuffer, CGraphDesc, COp, CView
    assert ctypes.sizeof(CView) == 12 and ctypes.sizeof(CBuffer) == 8
    assert ctypes.sizeof(COp) == 6 * 4 + 2 * 8 + 4 * 12 + 16 * 12 + 16 * 4
    assert ctypes.sizeof(CGraphDesc) == 4 + 4 + 4 + 4 + 8 + 3 * 8


def test_sass_contains_blackwell_instructions():
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    import __graft_entry__ as entry
    entry.build()
    sass = subprocess.run([cuobjdump, "-sass", entry.LIB], capture_output=True, text=True, check=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UBLKCP", "UTCBAR"):   # tcgen05.mma / tcgen05.ld / cp.async.bulk / tcgen05.commit
        assert mnemonic in sass, mnemonic


def test_create_without_gpu_fails_with_enodev(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from aru_b200.engine import ARU_ENODEV, Engine, EngineError
    from aru_b200.synth import synth_pb
    with pytest.raises(EngineError) as ei:
        Engine(synth_pb("tiny"))
    assert ei.value.code == ARU_ENODEV and "no CPU fallback" in str(ei.value)


def test_missing_library_is_a_loud_error(monkeypatch):
    from aru_b200 import engine
    monkeypatch.setattr(engine, "_lib", None)
    monkeypatch.setattr(engine, "LIB_PATH", "/nonexistent/libaru_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        engine.load_library()


def test_f64_to_f32_staging_conversion_equals_numpy(built_lib):
    """Host helper (no GPU involved): the float64 page get_net_output receives -> the float32 staging buffer, on several
    threads; the same bits as numpy's astype (round to nearest even, overflow to inf, NaN kept)."""
    import numpy as np
    rng = np.random.default_rng(0)
    for count, threads in ((0, 0), (1, 0), (65535, 0), (1024 * 768, 0), (1856 * 1344 + 3, 5), (400000, 1), (3000000, 64)):
        src = rng.random(count) if count else np.zeros(0)
        if count > 10:
            src[:6] = [1e39, -1e39, np.nan, 1e-46, 0.1, 1.0 + 2.0 ** -24]
        dst = np.full(count, -7.0, dtype=np.float32)
        rc = built_lib.aru_f64_to_f32(ctypes.c_void_p(src.ctypes.data), ctypes.c_void_p(dst.ctypes.data), count, threads)
        assert rc == 0
        with np.errstate(over="ignore"):
            want = src.astype(np.float32)
        assert np.array_equal(dst, want, equal_nan=True), (count, threads)
    assert built_lib.aru_f64_to_f32(None, None, 5, 0) != 0
