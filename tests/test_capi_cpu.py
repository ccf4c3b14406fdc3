"""No-GPU checks of the drop-in boundary: the C-ABI library builds, loads, exports every symbol that
include/*.h declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import re

import pytest

from util import ROOT


def declared_symbols():
    names = []
    for h in sorted((ROOT / "include").glob("*.h")):
        txt = re.sub(r"/\*.*?\*/", "", h.read_text(), flags=re.S)
        names += re.findall(r"\b((?:flowamok|futhark)_[a-z0-9_]+)\s*\(", txt)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    import flowamok_b200.capi as capi
    L = capi.lib()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/*.h but not exported"


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
        if p.suffix in (".py", ".cu", ".cuh", ".h", ".cpp"):
            txt = p.read_text()
            assert not re.search(r"flowamok_oracle|from\s+oracle|import\s+oracle|oracle/|liboracle", txt), p
