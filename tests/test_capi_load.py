"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and
exports every symbol include/miniamr_b200.h declares; argument validation that
needs no device.  No compute calls here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from miniamr_b200 import build, capi
    build.build()
    return capi.load_library()


def header_functions():
    src = open(os.path.join(ROOT, "include", "miniamr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mamr_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    names = header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/miniamr_b200.h but not exported"


def test_binding_lists_every_symbol(lib):
    from miniamr_b200 import capi
    assert sorted(capi.EXPORTS) == header_functions()


def test_abi_version(lib):
    assert lib.mamr_abi_version() == 1


def test_create_rejects_bad_parameters_without_device(lib):
    from miniamr_b200 import capi
    h = C.c_void_p()
    bad = [dict(nx=3), dict(ny=0), dict(num_vars=0), dict(max_blocks=0), dict(stencil=0),
           dict(stencil=13), dict(code=3), dict(rank=2, num_ranks=2),
           # the one message-mode combination whose results differ from --code 0 in the reference
           dict(code=1, permute=1, stencil=27), dict(code=2, permute=1, stencil=0, num_vars=8)]
    for kw in bad:
        base = dict(nx=4, ny=4, nz=4, num_vars=2, comm_vars=0, max_blocks=8, stencil=7, code=0,
                    permute=0, device=-1, rank=0, num_ranks=1)
        base.update(kw)
        p = capi.Params(**base)
        rc = lib.mamr_create(C.byref(p), C.byref(h))
        assert rc in (2, 3), (kw, rc)          # MAMR_EINVAL / MAMR_EUNSUPPORTED
        assert lib.mamr_last_error()


def test_no_cpu_fallback(lib):
    """Without a CUDA device creation must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from miniamr_b200 import capi
    with pytest.raises(capi.MamrError, match="no CPU fallback|CUDA"):
        capi.DeviceMesh(4, 4, 4, 2, 8)
