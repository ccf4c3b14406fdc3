"""Shared helpers for the test-suite: random road networks, the oracle driver, the CPU emulation
of the CUDA tile programs (tests/emul) and field-by-field comparison."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from oracle import oracle as O  # noqa: E402  (tests are allowed to use the oracle)

PCG_INC = ((0xda3e39cb94b95bdb << 1) | 1) & 0xFFFFFFFFFFFFFFFF
STATE_FIELDS = ("dy", "dx", "color", "pref", "rng_state")


def random_grid(seed, gh, gw, n_lines=None, car_p=0.35, weird_p=0.03, offroad_p=0.01):
    """Random road network + cars as a dict of numpy arrays in the reference's field layout.

    Lines are drawn like utils.fut add_line_* (so crossings become corners); cars mostly head along
    the road under them, a few head anywhere (including diagonally and against the flow) and a few
    sit on cells without road -- every state the public API can construct (utils.fut:35)."""
    r = np.random.default_rng(seed)
    uy = np.zeros((gh, gw), np.int8)
    ux = np.zeros((gh, gw), np.int8)
    if n_lines is None:
        n_lines = max(2, (gh + gw) // 4)
    for _ in range(n_lines):
        flow = r.choice([-1, 1])
        if r.random() < 0.5 and gw > 1:
            y = r.integers(0, gh)
            a, b = sorted(r.integers(0, gw, 2).tolist())
            ux[y, a:b + 1] = flow
        elif gh > 1:
            x = r.integers(0, gw)
            a, b = sorted(r.integers(0, gh, 2).tolist())
            uy[a:b + 1, x] = flow
    road = (uy != 0) | (ux != 0)
    dy = np.zeros((gh, gw), np.int8)
    dx = np.zeros((gh, gw), np.int8)
    car = road & (r.random((gh, gw)) < car_p)
    along_y = (uy != 0) & ((ux == 0) | (r.random((gh, gw)) < 0.5))
    dy[car & along_y] = uy[car & along_y]
    dx[car & ~along_y] = ux[car & ~along_y]
    weird = (r.random((gh, gw)) < weird_p) & road
    dy[weird] = r.integers(-1, 2, (gh, gw))[weird]
    dx[weird] = r.integers(-1, 2, (gh, gw))[weird]
    off = (r.random((gh, gw)) < offroad_p) & ~road
    dy[off] = r.integers(-1, 2, (gh, gw))[off]
    dx[off] = r.integers(-1, 2, (gh, gw))[off]
    color = r.integers(0, 2**32, (gh, gw), dtype=np.uint64).astype(np.uint32)
    pref = (r.random((gh, gw)) < 0.5).astype(np.uint8)
    rng_state = r.integers(0, 2**63, (gh, gw), dtype=np.uint64) * np.uint64(2) + r.integers(0, 2, (gh, gw), dtype=np.uint64)
    return dict(uy=uy, ux=ux, dy=dy, dx=dx, color=color, pref=pref, rng_state=rng_state,
                rng_inc=np.full((gh, gw), PCG_INC, np.uint64))


def oracle_grid(arrs):
    gh, gw = arrs["uy"].shape
    g = O.Grid(gh, gw)
    g.set_arrays(uy=arrs["uy"], ux=arrs["ux"], dy=arrs["dy"], dx=arrs["dx"], color=arrs["color"],
                 pref=arrs["pref"], rng_state=arrs["rng_state"], rng_inc=arrs["rng_inc"])
    return g


def state_of(g: O.Grid):
    a = g.arrays()
    return {k: a[k] for k in STATE_FIELDS}


def assert_state_equal(a, b, what=""):
    for k in STATE_FIELDS:
        if not np.array_equal(a[k], b[k]):
            bad = np.argwhere(a[k] != b[k])
            raise AssertionError(f"{what}: field {k} differs at {len(bad)} cells, first {bad[:5].tolist()}")


def _splitmix64(x):
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def state_digest(st, row_lo=0, row_hi=None):
    """numpy restatement of flowamok_grid_digest (digest_cell in csrc/fa_kernels.cuh) over rows [row_lo, row_hi) of
    whole-grid state arrays: (sum, xor) mod 2^64."""
    gh, gw = st["dy"].shape
    row_hi = gh if row_hi is None else row_hi
    sl = slice(row_lo, row_hi)
    dy, dx = st["dy"][sl].astype(np.int64), st["dx"][sl].astype(np.int64)
    bits = ((dy != 0) * 1 + (dy < 0) * 2 + (dx != 0) * 4 + (dx < 0) * 8 + (st["pref"][sl] != 0) * 16).astype(np.uint64)
    idx = (np.arange(row_lo, row_hi, dtype=np.uint64)[:, None] * np.uint64(gw) + np.arange(gw, dtype=np.uint64)[None, :])
    with np.errstate(over="ignore"):
        v = _splitmix64(_splitmix64(idx * np.uint64(0x9E3779B97F4A7C15) + ((bits << np.uint64(32)) | st["color"][sl].astype(np.uint64)))
                        ^ st["rng_state"][sl].astype(np.uint64))
        return int(np.add.reduce(v.ravel(), dtype=np.uint64)), int(np.bitwise_xor.reduce(v.ravel()))


# ---------------------------------------------------------------------------------------------
# CPU emulation of the strip programs
# ---------------------------------------------------------------------------------------------
_EMUL_SRC = ROOT / "tests" / "emul" / "emul.cpp"
_EMUL_SO = ROOT / "tests" / "emul" / "_build" / "libemul.so"
_emul = None


def emul_lib():
    global _emul
    if _emul is not None:
        return _emul
    deps = [_EMUL_SRC, ROOT / "tests/emul/simt_host.h"] + [ROOT / "flowamok_b200/csrc" / f
                                                         for f in ("fa_tiles.h", "fa_bits.h", "fa_stream.h", "fa_simt.h")]
    # FLOWAMOK_EMUL_SANITIZE=1: the kernel programs under AddressSanitizer + UBSan (the stand-in for compute-sanitizer's
    # memcheck, which is closed on the GPU pool).  Run as
    #   FLOWAMOK_EMUL_SANITIZE=1 LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 pytest tests/test_emul.py tests/test_bands_cpu.py
    san = os.environ.get("FLOWAMOK_EMUL_SANITIZE") == "1"
    so = _EMUL_SO.with_name("libemul_asan.so") if san else _EMUL_SO
    if not so.exists() or any(d.stat().st_mtime > so.stat().st_mtime for d in deps):
        so.parent.mkdir(parents=True, exist_ok=True)
        cxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
        flags = (["-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-fno-sanitize-recover=undefined"]
                 if san else ["-O2"])
        subprocess.run([cxx, *flags, "-std=c++17", "-fPIC", "-shared", "-o", str(so), str(_EMUL_SRC)], check=True)
    L = C.CDLL(str(so))
    L.emul_new.restype = C.c_void_p
    L.emul_new.argtypes = [C.c_int, C.c_int]
    L.emul_free.argtypes = [C.c_void_p]
    L.emul_upload.argtypes = [C.c_void_p] + [C.c_void_p] * 7 + [C.c_uint64]
    L.emul_download.argtypes = [C.c_void_p] + [C.c_void_p] * 5
    L.emul_step.restype = C.c_int64
    L.emul_step.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int64] + [C.c_void_p] * 7
    _emul = L
    return L


class EmulGrid:
    def __init__(self, arrs):
        self.L = emul_lib()
        self.gh, self.gw = arrs["uy"].shape
        self.p = self.L.emul_new(self.gh, self.gw)
        c = {k: np.ascontiguousarray(v) for k, v in arrs.items()}
        self.L.emul_upload(self.p, c["uy"].ctypes.data, c["ux"].ctypes.data, c["dy"].ctypes.data,
                           c["dx"].ctypes.data, c["color"].ctypes.data, c["pref"].ctypes.data,
                           c["rng_state"].ctypes.data, int(c["rng_inc"].flat[0]) if c["rng_inc"].size else 0)
        self.stats = np.zeros(2, np.uint32)

    def __del__(self):
        try:
            self.L.emul_free(self.p)
        except Exception:
            pass

    def step(self, variant=0, nthreads=7, chooser=O.CHOOSE_RANDOM, tape_bit=0, rings=None):
        if rings is None:
            return int(self.L.emul_step(self.p, variant, nthreads, chooser, tape_bit, -1, None, None, None,
                                        None, None, None, self.stats.ctypes.data))
        off, idx, cy, cx, grant = rings
        ry = np.ascontiguousarray(idx // self.gw, np.int64)
        rx = np.ascontiguousarray(idx % self.gw, np.int64)
        return int(self.L.emul_step(self.p, variant, nthreads, chooser, tape_bit, len(off) - 1, off.ctypes.data,
                                    ry.ctypes.data, rx.ctypes.data, cy.ctypes.data, cx.ctypes.data,
                                    grant.ctypes.data, self.stats.ctypes.data))

    def state(self):
        n = (self.gh, self.gw)
        dy, dx = np.zeros(n, np.int8), np.zeros(n, np.int8)
        color, pref, rs = np.zeros(n, np.uint32), np.zeros(n, np.uint8), np.zeros(n, np.uint64)
        self.L.emul_download(self.p, dy.ctypes.data, dx.ctypes.data, color.ctypes.data, pref.ctypes.data,
                             rs.ctypes.data)
        return dict(dy=dy, dx=dx, color=color, pref=pref, rng_state=rs)
