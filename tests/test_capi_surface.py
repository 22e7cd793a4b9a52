"""CPU checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and exports every
symbol include/suhmo_gpu.h declares.  No compute calls (no GPU here)."""
import os
import re

import pytest

from suhmo_b200 import build, capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "suhmo_gpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(sg_[A-Za-z0-9_]+)\s*\(", hdr)))


def test_library_builds_and_loads():
    so = build.build()
    assert os.path.exists(so)
    L = capi.lib()
    assert L.sg_version() >= 100


def test_every_declared_symbol_is_exported():
    L = capi.lib()
    syms = declared_symbols()
    assert len(syms) > 70
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing


def test_python_signatures_cover_the_header():
    syms = set(declared_symbols()) - {"sg_last_error", "sg_version"}
    assert syms == set(capi.SIGNATURES), (syms ^ set(capi.SIGNATURES))


def test_no_cpu_fallback_without_device():
    """Without a CUDA device context creation must fail loudly, not fall back to a CPU path."""
    import ctypes as C
    L = capi.lib()
    h = C.c_void_p()
    st = L.sg_ctx_create(C.byref(h), 0, 0, 1, None)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        assert st == capi.OK
        L.sg_ctx_destroy(h)
    else:
        assert st == capi.ERR_CUDA
        assert b"no CPU fallback" in L.sg_last_error()


def test_product_package_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under suhmo_b200/ may import, link or call it."""
    pkg = os.path.join(ROOT, "suhmo_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                code = "\n".join(l for l in txt.splitlines() if not l.strip().startswith(("#", "//", "*", "/*")))
                assert "suhmo_oracle" not in code and "from oracle" not in code and "import oracle" not in code, os.path.join(dp, f)
