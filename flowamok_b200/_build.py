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


if __name__ == "__main__":
    print(build(force=True, verbose=True))
