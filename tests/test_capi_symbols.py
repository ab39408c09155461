"""CPU: the C-ABI shared library builds for sm_100a, loads without a GPU, and exports every symbol that
include/rmem_b200.h declares (no compute calls here)."""
import os
import re

from rmem_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "rmem_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rmem_[a-z0-9_]+)\s*\(", src)))


def test_header_binding_and_library_agree():
    lib = _capi.load()
    hdr = header_symbols()
    assert sorted(_capi.SYMBOLS) == hdr, f"binding list and header differ: {set(_capi.SYMBOLS) ^ set(hdr)}"
    for s in hdr:
        assert hasattr(lib, s), f"{s} not exported by {_capi.LIB_PATH}"
    assert lib.rmem_version() >= 100
    assert lib.rmem_operand_dtype() in (b"fp16", b"bf16")


def test_sass_is_blackwell_native():
    """The attention kernel must be tcgen05 + TMA, not a recompiled mma.sync path."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        import pytest
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", _capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass, "no tcgen05.mma in the built library"
    assert "UTMALDG" in sass, "no TMA loads in the built library"
    assert "LDTM" in sass, "no tcgen05.ld in the built library"


def test_errors_are_reported_not_thrown():
    import ctypes as C
    lib = _capi.load()
    nbytes = C.c_size_t()
    cfg = _capi.EngineConfig(0, 480, 854, 1, 7, 1, 0, 5, 0, 0, 0, 0)   # not 16k+1 -> must be refused with a message
    rc = lib.rmem_engine_arena_bytes(C.byref(cfg), C.byref(nbytes))
    assert rc != 0 and b"16k+1" in lib.rmem_last_error()
    cfg = _capi.EngineConfig(0, 481, 849, 1, 7, 1, 0, 5, 0, 0, 0, 1)   # GRU_MEMORY is refused
    assert lib.rmem_engine_arena_bytes(C.byref(cfg), C.byref(nbytes)) != 0 and b"GRU_MEMORY" in lib.rmem_last_error()
    cfg = _capi.EngineConfig(0, 481, 849, 1, 7, 1, 0, 5, 0, 0, 0, 0)
    assert lib.rmem_engine_arena_bytes(C.byref(cfg), C.byref(nbytes)) == 0
    assert 100e6 < nbytes.value < 4e9
