"""In-tree build of the CUDA library (nvcc, sm_100a only).  The .so is git-ignored but travels to the
GPU box with the gpurun snapshot."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
SO = PKG / "libflowamok_b200.so"
SOURCES = [CSRC / "fa_api.cu"]
DEPS = [CSRC / f for f in ("fa_api.cu", "fa_kernels.cuh", "fa_stream.h", "fa_simt.h", "fa_tiles.h", "fa_bits.h",
                           "fa_futhark.inc")] + [PKG.parent / "include" / "flowamok.h"]

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared", "-ldl"]


# the lys frontends: host code above the C ABI (g++, no CUDA), one library per program like the generated ones
LYS = PKG / "lys"
LYS_SO = {"explorer": PKG / "libflowamok_explorer.so", "designer": PKG / "libflowamok_designer.so"}
LYS_DEPS = [LYS / "fa_lys.h", LYS / "fa_lys_api.inc", CSRC / "fa_bits.h", PKG.parent / "include" / "flowamok.h",
            PKG.parent / "include" / "flowamok_lys.h"]
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wall", "-Wl,-Bsymbolic", "-Wl,--exclude-libs,ALL"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not SO.exists():
        return True
    t = SO.stat().st_mtime
    return any(d.exists() and d.stat().st_mtime > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return SO
    cmd = [nvcc_path(), *NVCC_FLAGS, "-o", str(SO), *map(str, SOURCES)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return SO


def lys_needs_build(name: str) -> bool:
    so = LYS_SO[name]
    if not so.exists():
        return True
    t = so.stat().st_mtime
    return any(d.exists() and d.stat().st_mtime > t for d in LYS_DEPS + [LYS / f"{name}_lys.cpp"])


def build_lys(name: str, force: bool = False) -> Path:
    """libflowamok_<name>.so: links against libflowamok_b200.so next to it ($ORIGIN rpath)"""
    so = LYS_SO[name]
    build()
    if not force and not lys_needs_build(name):
        return so
    cxx = os.environ.get("CXX") or shutil.which("g++") or "g++"
    cmd = [cxx, *CXX_FLAGS, "-o", str(so), str(LYS / f"{name}_lys.cpp"), f"-L{PKG}", "-lflowamok_b200", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"{cxx} failed:\n" + r.stdout + r.stderr)
    return so


if __name__ == "__main__":
    for _n in LYS_SO:
        print(build_lys(_n, force=True))
    print(build(force=True, verbose=True))
