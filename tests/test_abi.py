"""CPU tests: the C-ABI library builds for sm_100a, loads, and exports every symbol include/*.h declares.
No compute calls are made (there is no GPU here)."""
import ctypes
import os
import re
import subprocess

import pytest

from dualdiffusion_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dualdiffusion_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"DD_API\s+[\w\s\*]+?\b(dd_\w+)\s*\(", text)))


def test_header_declares_api():
    syms = declared_symbols()
    assert "dd_mpconv_forward" in syms and "dd_attention" in syms and "dd_last_error" in syms
    assert len(syms) >= 20


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in the header but not exported"


def test_binding_covers_header():
    assert set(_lib.EXPORTED_SYMBOLS) == set(declared_symbols())
    lib = _lib.load()
    assert lib.dd_abi_version() == 1


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", build.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass      # tcgen05.mma
    assert "UTMALDG" in sass      # cp.async.bulk.tensor (TMA)
    assert "LDTM" in sass         # tcgen05.ld


def test_no_cpu_fallback_raises():
    import torch
    from dualdiffusion_b200 import ops
    with pytest.raises(RuntimeError):
        ops.weight_prep(torch.randn(4, 4, 1, 1))
