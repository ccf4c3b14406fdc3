"""A lys-style C `main` (tests/lys_main/lys_main.c: the calls lys' C driver makes, minus SDL) compiled as C against
include/flowamok_lys.h and linked against libflowamok_explorer.so / libflowamok_designer.so -- the drop-in claim of
INTEGRATION.md 1b at the level of the linker.  Without a GPU: it compiles, links and refuses to run; with one: every
frame it renders equals the oracle-side model's."""
import shutil
import subprocess

import numpy as np
import pytest

from util import ROOT

SRC = ROOT / "tests" / "lys_main" / "lys_main.c"


def build_main(name, tmp_path):
    from flowamok_b200 import _build
    _build.build_lys(name)
    exe = tmp_path / f"lys_main_{name}"
    cmd = [shutil.which("gcc") or "gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-O1", f"-I{ROOT / 'include'}", str(SRC), "-o", str(exe),
           f"-L{_build.PKG}", f"-lflowamok_{name}", "-lflowamok_b200", f"-Wl,-rpath,{_build.PKG}"]
    if name == "designer":
        cmd.insert(1, "-DFLOWAMOK_LYS_DESIGNER")
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


@pytest.mark.parametrize("name", ["explorer", "designer"])
def test_c_main_links_and_has_no_cpu_path(name, tmp_path):
    import torch
    exe = build_main(name, tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present: the gpu test runs it")
    r = subprocess.run([str(exe), "2"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr


def checksum(frames):
    """sum = sum * 1099511628211 + pixel over all pixels of all frames, mod 2^64 (as lys_main.c)"""
    P = np.uint64(1099511628211)
    n = frames[0].size
    with np.errstate(over="ignore"):
        pw = np.cumprod(np.concatenate([[np.uint64(1)], np.full(n, P, np.uint64)]))   # P^0 .. P^n
        s = np.uint64(0)
        for f in frames:
            h = np.sum(f.reshape(-1).astype(np.uint64) * pw[n - 1::-1], dtype=np.uint64)
            s = s * pw[n] + h
    return int(s)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["explorer", "designer"])
def test_c_main_frames_equal_the_model(name, tmp_path):
    import lys_model as M
    exe = build_main(name, tmp_path)
    n_frames = 40
    r = subprocess.run([str(exe), str(n_frames)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    if name == "explorer":
        m = M.ExplorerModel(42, 600, 800)
        for k in (M.SDLK_RIGHT,) * 4 + (M.SDLK_a,):
            m.keydown(k)
    else:
        m = M.DesignerModel(42, 600, 800)
        for ev in ((1, 105, 255), (1, 105, 455), (0, 0, 0)):
            m.mouse(*ev)
    frames = []
    td = np.float32(np.float32(1.0) / np.float32(30.0)) + np.float32(1e-4)
    for _ in range(n_frames):
        m.event_step(td)
        frames.append(m.render())
    text = ",".join(str(v) for v in m.text_content(30.0))
    want = f"{name} frames={n_frames} checksum={checksum(frames):016x} text=({text}) colour=ffffffff grab=0"
    assert r.stdout.startswith(want), (r.stdout, want)
