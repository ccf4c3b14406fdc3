"""No-GPU checks of the drop-in boundary: the C-ABI library builds, loads, exports every symbol that
include/*.h declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import re

import pytest

from util import ROOT


def declared_symbols(headers):
    names = []
    for h in headers:
        txt = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / h).read_text(), flags=re.S)
        names += re.findall(r"\b((?:flowamok|futhark)_[a-z0-9_]+)\s*\(", txt)
    return sorted(set(names))


def test_every_header_belongs_to_a_library():
    assert sorted(p.name for p in (ROOT / "include").glob("*.h")) == ["flowamok.h", "flowamok_futhark.h", "flowamok_lys.h"]


def test_library_exports_every_declared_symbol():
    import flowamok_b200.capi as capi
    L = capi.lib()
    syms = declared_symbols(["flowamok.h", "flowamok_futhark.h"])
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/*.h but not exported"


@pytest.mark.parametrize("name", ["explorer", "designer"])
def test_lys_libraries_export_every_declared_symbol(name):
    """include/flowamok_lys.h is the surface of BOTH frontends' libraries (same names, like the generated ones)"""
    import flowamok_b200.lys as lys
    L = lys.load(name)
    syms = declared_symbols(["flowamok_lys.h"])
    assert sorted(syms) == sorted(lys.EXPORTS)
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/flowamok_lys.h but not exported by libflowamok_{name}.so"


@pytest.mark.parametrize("name", ["explorer", "designer"])
def test_lys_frontend_without_device(name):
    """text_format needs no device; init does and says so (no CPU path behind the lys surface either)"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import flowamok_b200.lys as lys
    L = lys.load(name)
    cfg = L.futhark_context_config_new()
    ctx = L.futhark_context_new(cfg)
    L.futhark_context_config_free(cfg)
    arr = C.c_void_p()
    assert L.futhark_entry_text_format(ctx, C.byref(arr)) == 0
    n = L.futhark_shape_u8_1d(ctx, arr)[0]
    buf = C.create_string_buffer(n)
    assert L.futhark_values_u8_1d(ctx, arr, C.cast(buf, C.c_void_p)) == 0
    assert buf.raw.startswith(b"FPS: %d\n")
    L.futhark_free_u8_1d(ctx, arr)
    st = C.c_void_p()
    assert L.futhark_entry_init(ctx, C.byref(st), 1, 600, 800) != 0 and not st.value
    msg = L.futhark_context_get_error(ctx)
    assert msg and b"CUDA" in C.string_at(msg)
    L.futhark_context_free(ctx)


def test_version_string():
    import flowamok_b200.capi as capi
    assert b"flowamok_b200" in capi.lib().flowamok_version()


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import flowamok_b200 as fa
    with pytest.raises(fa.FlowamokError, match="no CUDA device"):
        fa.Context()


def test_product_never_touches_the_oracle():
    for p in (ROOT / "flowamok_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h", ".cpp", ".inc"):
            txt = p.read_text()
            assert not re.search(r"flowamok_oracle|from\s+oracle|import\s+oracle|oracle/|liboracle", txt), p
