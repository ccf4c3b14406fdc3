"""The loop of INTEGRATION.md section 1 as a C program (tests/readme_main/readme_main.c) against include/flowamok_futhark.h,
linked against libflowamok_b200.so.  Without a GPU: it compiles as C, links and refuses to run; with one: its frames are
those of README.fut executed from the reference source (pixel structure from tests/golden/lys_traces.npz, the grey / mix
arithmetic of athas/matte through the oracle's render)."""
import shutil
import subprocess

import numpy as np
import pytest

from util import O, ROOT

SRC = ROOT / "tests" / "readme_main" / "readme_main.c"


def build_main(tmp_path):
    from flowamok_b200 import _build
    _build.build()
    exe = tmp_path / "readme_main"
    cmd = [shutil.which("gcc") or "gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-O1", f"-I{ROOT / 'include'}", str(SRC), "-o", str(exe),
           f"-L{_build.PKG}", "-lflowamok_b200", f"-Wl,-rpath,{_build.PKG}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_host_program_links_and_has_no_cpu_path(tmp_path):
    import torch
    exe = build_main(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present: the gpu test runs it")
    r = subprocess.run([str(exe), "2"], capture_output=True, text=True)
    assert r.returncode == 1 and "futhark_entry_init" in r.stderr


@pytest.mark.gpu
def test_c_host_program_frames(tmp_path):
    from test_lys_main import checksum
    from test_lys_traces import T, check_frame
    exe = build_main(tmp_path)
    n_frames = int(T["readme/meta"][4])
    r = subprocess.run([str(exe), str(n_frames)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    got = [line.split()[2] for line in r.stdout.splitlines()]
    oa = O.Anim(8, 30, 30, 10, 123)
    for k in range(n_frames):
        frame = oa.step(render=True)
        check_frame(k, frame)                                   # the oracle's frame has the reference's pixel structure ...
        assert got[k] == f"{checksum([frame]):016x}", f"frame {k}"   # ... and the C program printed exactly that frame
