"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE -- see flowamok_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package flowamok_b200 never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "libflowamok_oracle.so"


def build(force: bool = False) -> Path:
    src = [_HERE / "flowamok_oracle.c", _HERE / "flowamok_oracle.h"]
    if force or not _SO.exists() or any(s.stat().st_mtime > _SO.stat().st_mtime for s in src):
        subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
    return _SO


class _Rng(C.Structure):
    _fields_ = [("state", C.c_uint64), ("inc", C.c_uint64)]


class _Grid(C.Structure):
    _fields_ = [("gh", C.c_int64), ("gw", C.c_int64),
                ("rng_state", C.POINTER(C.c_uint64)), ("rng_inc", C.POINTER(C.c_uint64)),
                ("uy", C.POINTER(C.c_int8)), ("ux", C.POINTER(C.c_int8)),
                ("dy", C.POINTER(C.c_int8)), ("dx", C.POINTER(C.c_int8)),
                ("color", C.POINTER(C.c_uint32)),
                ("calc", C.POINTER(C.c_uint8)), ("my", C.POINTER(C.c_uint8)),
                ("mx", C.POINTER(C.c_uint8)), ("pref", C.POINTER(C.c_uint8))]


class _Leak(C.Structure):
    _fields_ = [("y_next", C.c_int64), ("x_next", C.c_int64), ("y_src", C.c_int64), ("x_src", C.c_int64),
                ("rng", _Rng), ("uy", C.c_int8), ("ux", C.c_int8), ("dy", C.c_int8), ("dx", C.c_int8),
                ("color", C.c_uint32), ("calc", C.c_uint8), ("my", C.c_uint8), ("mx", C.c_uint8),
                ("pref", C.c_uint8)]


class _Chooser(C.Structure):
    _fields_ = [("mode", C.c_int), ("tape", C.POINTER(C.c_uint8)), ("tape_len", C.c_int64),
                ("tick", C.c_int64)]


CHOOSE_RANDOM, CHOOSE_Y, CHOOSE_X, CHOOSE_TAPE = 0, 1, 2, 3
FIELDS = ("rng_state", "rng_inc", "uy", "ux", "dy", "dx", "color", "calc", "my", "mx", "pref")
_DT = dict(rng_state=np.uint64, rng_inc=np.uint64, uy=np.int8, ux=np.int8, dy=np.int8, dx=np.int8,
           color=np.uint32, calc=np.uint8, my=np.uint8, mx=np.uint8, pref=np.uint8)

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(str(_SO))
    GP = C.POINTER(_Grid)
    L.fo_grid_alloc.restype = GP
    L.fo_grid_alloc.argtypes = [C.c_int64, C.c_int64]
    L.fo_grid_free.argtypes = [GP]
    L.fo_grid_clone.restype = GP
    L.fo_grid_clone.argtypes = [GP]
    L.fo_create_grid.restype = _Rng
    L.fo_create_grid.argtypes = [GP, _Rng]
    L.fo_rng_from_seed.restype = _Rng
    L.fo_rng_from_seed.argtypes = [C.POINTER(C.c_int32), C.c_int64]
    L.fo_split_rng.argtypes = [C.c_int64, _Rng, C.POINTER(_Rng)]
    L.fo_join_rng.restype = _Rng
    L.fo_join_rng.argtypes = [C.POINTER(_Rng), C.c_int64]
    L.fo_pcg32_rand.restype = C.c_uint32
    L.fo_pcg32_rand.argtypes = [C.POINTER(_Rng)]
    L.fo_dist_i8_rand.restype = C.c_int8
    L.fo_dist_i8_rand.argtypes = [C.c_int8, C.c_int8, C.POINTER(_Rng)]
    L.fo_random_color.restype = C.c_uint32
    L.fo_random_color.argtypes = [C.POINTER(_Rng)]
    L.fo_add_line_vertical.argtypes = [GP, C.c_int64, C.c_int64, C.c_int64, C.c_int8]
    L.fo_add_line_horizontal.argtypes = [GP, C.c_int64, C.c_int64, C.c_int64, C.c_int8]
    L.fo_add_cell.argtypes = [GP, C.c_int64, C.c_int64, C.c_uint32, C.c_int8, C.c_int8]
    L.fo_update_pass.argtypes = [GP]
    L.fo_move_until_fixpoint.restype = C.c_int64
    L.fo_move_until_fixpoint.argtypes = [GP]
    L.fo_resolve_cycles.restype = C.c_int
    L.fo_resolve_cycles.argtypes = [GP, C.c_void_p, C.c_int64]
    LPP = C.POINTER(C.POINTER(_Leak))
    L.fo_move_and_reset_cells.restype = C.c_int64
    L.fo_move_and_reset_cells.argtypes = [GP, C.POINTER(_Chooser), LPP]
    L.fo_step_quick.restype = C.c_int64
    L.fo_step_quick.argtypes = [GP, C.POINTER(_Chooser), LPP]
    L.fo_step_perfect.restype = C.c_int64
    L.fo_step_perfect.argtypes = [GP, C.POINTER(_Chooser), C.c_void_p, C.c_int64, LPP]
    L.fo_find_cycles.restype = C.c_int64
    L.fo_find_cycles.argtypes = [GP, C.POINTER(C.c_void_p), C.c_int64]
    L.fo_resolve_cycles_sparse.argtypes = [GP, C.c_int64] + [C.c_void_p] * 5
    L.fo_step_perfect_sparse.restype = C.c_int64
    L.fo_step_perfect_sparse.argtypes = [GP, C.POINTER(_Chooser), C.c_int64] + [C.c_void_p] * 5 + [LPP]
    L.fo_generate_synthetic.restype = C.c_int64
    L.fo_generate_synthetic.argtypes = [GP, C.c_uint64, C.c_uint32, C.c_uint32] + [C.POINTER(C.c_void_p)] * 5
    L.fo_scenario_name.restype = C.c_char_p
    L.fo_scenario_name.argtypes = [C.c_int]
    L.fo_scenario_init.argtypes = [C.c_int, GP]
    L.fo_scenario_step.restype = C.c_int
    L.fo_scenario_step.argtypes = [C.c_int, GP, C.c_int64, C.POINTER(_Rng)]
    L.fo_render.argtypes = [GP, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]
    L.fo_set_threads.argtypes = [C.c_int]
    L.fo_get_max_threads.restype = C.c_int
    _lib = L
    _lib._libc = C.CDLL(None)
    _lib._libc.free.argtypes = [C.c_void_p]
    return _lib


class Rng:
    """pcg32 state (cpprandom restatement, unpinned)."""

    def __init__(self, state=0, inc=0):
        self.c = _Rng(state, inc)

    @classmethod
    def from_seed(cls, seeds):
        a = (C.c_int32 * len(seeds))(*seeds)
        r = lib().fo_rng_from_seed(a, len(seeds))
        return cls(r.state, r.inc)

    @property
    def state(self):
        return int(self.c.state)

    @property
    def inc(self):
        return int(self.c.inc)

    def rand(self):
        return int(lib().fo_pcg32_rand(C.byref(self.c)))

    def dist_i8(self, lo, hi):
        return int(lib().fo_dist_i8_rand(lo, hi, C.byref(self.c)))

    def split(self, n):
        out = (_Rng * n)()
        lib().fo_split_rng(n, self.c, out)
        return [Rng(o.state, o.inc) for o in out]


class Chooser:
    def __init__(self, mode=CHOOSE_RANDOM, tape=None):
        self.c = _Chooser()
        self.c.mode = mode
        self._tape = None
        if tape is not None:
            self._tape = np.ascontiguousarray(tape, dtype=np.uint8)
            self.c.tape = self._tape.ctypes.data_as(C.POINTER(C.c_uint8))
            self.c.tape_len = len(self._tape)
        self.c.tick = 0


class Grid:
    """Owning wrapper; field arrays are exposed as zero-copy numpy views (re-fetched on access
    because the C side swaps buffers on every pass)."""

    def __init__(self, gh, gw, _ptr=None):
        self.L = lib()
        self.p = _ptr if _ptr is not None else self.L.fo_grid_alloc(gh, gw)
        self.gh, self.gw = int(gh), int(gw)

    def __del__(self):
        try:
            if self.p:
                self.L.fo_grid_free(self.p)
                self.p = None
        except Exception:
            pass

    def clone(self):
        return Grid(self.gh, self.gw, self.L.fo_grid_clone(self.p))

    def field(self, name):
        ptr = getattr(self.p.contents, name)
        n = self.gh * self.gw
        if n == 0:
            return np.zeros((self.gh, self.gw), _DT[name])
        a = np.ctypeslib.as_array(ptr, shape=(n,))
        return a.reshape(self.gh, self.gw)

    def arrays(self, copy=True):
        return {f: (self.field(f).copy() if copy else self.field(f)) for f in FIELDS}

    def set_arrays(self, **kw):
        for k, v in kw.items():
            self.field(k)[...] = np.asarray(v, dtype=_DT[k]).reshape(self.gh, self.gw)

    # --- construction (src/utils.fut) ---
    def create(self, rng: Rng) -> Rng:
        r = self.L.fo_create_grid(self.p, rng.c)
        return Rng(r.state, r.inc)

    def add_line_vertical(self, y0, y1, x, flow):
        self.L.fo_add_line_vertical(self.p, y0, y1, x, flow)

    def add_line_horizontal(self, y, x0, x1, flow):
        self.L.fo_add_line_horizontal(self.p, y, x0, x1, flow)

    def add_cell(self, y, x, color, dy, dx):
        self.L.fo_add_cell(self.p, y, x, color & 0xFFFFFFFF, dy, dx)

    # --- steppers ---
    def update_pass(self):
        self.L.fo_update_pass(self.p)

    def move_until_fixpoint(self):
        return int(self.L.fo_move_until_fixpoint(self.p))

    def resolve_cycles(self, masks):
        m = np.ascontiguousarray(masks, dtype=np.int8)
        n = 0 if m.size == 0 else m.shape[0]
        return int(self.L.fo_resolve_cycles(self.p, m.ctypes.data, n))

    def _leaks(self, lp, n, want):
        if not want:
            return None
        out = []
        for i in range(n):
            l = lp[i]
            out.append(dict(y_next=l.y_next, x_next=l.x_next, y_src=l.y_src, x_src=l.x_src,
                            rng_state=l.rng.state, rng_inc=l.rng.inc, uy=l.uy, ux=l.ux, dy=l.dy, dx=l.dx,
                            color=l.color, calc=l.calc, my=l.my, mx=l.mx, pref=l.pref))
        self.L._libc.free(C.cast(lp, C.c_void_p))
        return out

    def move_and_reset(self, chooser: Chooser, leaks=False):
        lp = C.POINTER(_Leak)()
        n = int(self.L.fo_move_and_reset_cells(self.p, C.byref(chooser.c), C.byref(lp) if leaks else None))
        chooser.c.tick += 1
        return n, self._leaks(lp, n, leaks)

    def step_quick(self, chooser: Chooser, leaks=False):
        lp = C.POINTER(_Leak)()
        n = int(self.L.fo_step_quick(self.p, C.byref(chooser.c), C.byref(lp) if leaks else None))
        chooser.c.tick += 1
        return n, self._leaks(lp, n, leaks)

    def step_perfect(self, chooser: Chooser, masks, leaks=False):
        m = np.ascontiguousarray(masks, dtype=np.int8)
        nm = 0 if m.size == 0 else m.shape[0]
        lp = C.POINTER(_Leak)()
        n = int(self.L.fo_step_perfect(self.p, C.byref(chooser.c), m.ctypes.data, nm,
                                       C.byref(lp) if leaks else None))
        chooser.c.tick += 1
        return n, self._leaks(lp, n, leaks)

    def step_perfect_sparse(self, chooser: Chooser, rings, leaks=False):
        off, idx, cy, cx, grant = rings
        lp = C.POINTER(_Leak)()
        n = int(self.L.fo_step_perfect_sparse(self.p, C.byref(chooser.c), len(off) - 1, off.ctypes.data,
                                              idx.ctypes.data, cy.ctypes.data, cx.ctypes.data,
                                              grant.ctypes.data, C.byref(lp) if leaks else None))
        chooser.c.tick += 1
        return n, self._leaks(lp, n, leaks)

    def resolve_cycles_sparse(self, rings):
        off, idx, cy, cx, grant = rings
        self.L.fo_resolve_cycles_sparse(self.p, len(off) - 1, off.ctypes.data, idx.ctypes.data,
                                        cy.ctypes.data, cx.ctypes.data, grant.ctypes.data)

    def find_cycles(self, max_paths=0):
        out = C.c_void_p()
        n = int(self.L.fo_find_cycles(self.p, C.byref(out), max_paths))
        if n < 0:
            raise RuntimeError("find_cycles: path budget exceeded")
        sz = n * self.gh * self.gw * 2
        if n == 0:
            if out.value:
                self.L._libc.free(out)
            return np.zeros((0, self.gh, self.gw, 2), np.int8)
        buf = (C.c_int8 * sz).from_address(out.value)
        m = np.frombuffer(buf, dtype=np.int8).reshape(n, self.gh, self.gw, 2).copy()
        self.L._libc.free(out)
        return m

    def rings_from_masks(self, masks):
        """dense masks -> sparse (off, idx, cy, cx, grant) as numpy arrays (same rule as the C helper)."""
        m = np.ascontiguousarray(masks, dtype=np.int8)
        n = 0 if m.size == 0 else m.shape[0]
        uy, ux = self.field("uy"), self.field("ux")
        off = [0]
        idx, cy, cx, grant = [], [], [], []
        for k in range(n):
            ys, xs = np.nonzero((m[k, :, :, 0] != 0) | (m[k, :, :, 1] != 0))
            for y, x in zip(ys.tolist(), xs.tolist()):
                idx.append(y * self.gw + x)
                cy.append(int(m[k, y, x, 0]))
                cx.append(int(m[k, y, x, 1]))
                if uy[y, x] != 0 and ux[y, x] != 0:
                    yp = y - int(uy[y, x])
                    isy = 0 <= yp < self.gh and m[k, yp, x, 0] == uy[y, x] and m[k, yp, x, 1] == 0
                    grant.append(1 if isy else 2)
                else:
                    grant.append(1 if uy[y, x] != 0 else 2)
            off.append(len(idx))
        return (np.asarray(off, np.int64), np.asarray(idx, np.int64), np.asarray(cy, np.int8),
                np.asarray(cx, np.int8), np.asarray(grant, np.uint8))

    def generate_synthetic(self, seed=1, density16=6554, ring_full16=2):
        """fills the grid with the synthetic network; returns the sparse ring arrays"""
        ptrs = [C.c_void_p() for _ in range(5)]
        n = int(self.L.fo_generate_synthetic(self.p, seed, density16, ring_full16, *[C.byref(p) for p in ptrs]))
        tot = n * 48
        def take(p, dt, cnt):
            a = np.frombuffer((C.c_char * (cnt * np.dtype(dt).itemsize)).from_address(p.value), dtype=dt).copy()
            self.L._libc.free(p)
            return a
        return (take(ptrs[0], np.int64, n + 1), take(ptrs[1], np.int64, max(tot, 1))[:tot],
                take(ptrs[2], np.int8, max(tot, 1))[:tot], take(ptrs[3], np.int8, max(tot, 1))[:tot],
                take(ptrs[4], np.uint8, max(tot, 1))[:tot])

    # --- scenarios / render ---
    def scenario_init(self, sid):
        self.L.fo_scenario_init(sid, self.p)

    def scenario_step(self, sid, steps, rng: Rng) -> bool:
        return bool(self.L.fo_scenario_step(sid, self.p, steps, C.byref(rng.c)))

    def render(self, scale):
        h, w = self.gh * scale, self.gw * scale
        px = np.zeros((h, w), np.uint32)
        self.L.fo_render(self.p, h, w, scale, px.ctypes.data)
        return px


def scenario_name(sid):
    return lib().fo_scenario_name(sid).decode()


def set_threads(n):
    lib().fo_set_threads(n)


def max_threads():
    return int(lib().fo_get_max_threads())


# ---- reference drivers restated: animation.fut:14-32 and explorer/explorer.fut:84-130 ----

class Anim:
    """mk_anim(scenario) of animation.fut: init gh gw scale seed; step -> (pixels, state)."""

    def __init__(self, sid, gh, gw, scale, seed, chooser: Chooser | None = None):
        self.sid, self.scale = sid, scale
        self.grid = Grid(gh, gw)
        self.rng = self.grid.create(Rng.from_seed([seed]))          # animation.fut:15
        self.grid.scenario_init(sid)                                 # :16
        self.cycles = self.grid.find_cycles()                        # :17
        self.steps = 0                                               # never incremented (:29-31)
        self.chooser = chooser or Chooser(CHOOSE_RANDOM)

    def step(self, render=True):
        px = self.grid.render(self.scale) if render else None       # :22 render BEFORE the tick
        self.grid.step_perfect(self.chooser, self.cycles)            # :24
        if self.grid.scenario_step(self.sid, self.steps, self.rng):  # :25
            self.cycles = self.grid.find_cycles()                    # :26-28
        return px


class Explorer:
    """explorer/explorer.fut init (:89-101) and step s n (:109-130) for one scenario id."""
    GH, GW = 108, 192                                                # :48-49

    def __init__(self, seed, sid, perfect=True, chooser: Chooser | None = None):
        rng = Rng.from_seed([np.int32(np.uint32(seed)).item()])      # :90 i32.u32 seed
        rngs = rng.split(14)                                         # :91
        self.grids = []
        outs = []
        for k in range(14):                                          # :92 map2 init_grid
            g = Grid(self.GH, self.GW)
            outs.append(g.create(rngs[k]))
            g.scenario_init(k)
            self.grids.append(g)
        arr = (_Rng * 14)(*[o.c for o in outs])
        j = lib().fo_join_rng(arr, 14)                               # :94
        self.rng = Rng(j.state, j.inc)
        self.sid, self.perfect = sid, perfect
        self.grid = self.grids[sid]
        self.cycles = self.grid.find_cycles() if perfect else np.zeros((0, self.GH, self.GW, 2), np.int8)
        self.step_cur = 0
        self.n_cell_leaks = 0
        self.chooser = chooser or Chooser(CHOOSE_RANDOM)

    def step(self, n_steps=1):
        leaks = 0
        for _ in range(n_steps):                                     # :115
            if self.perfect:
                n, _l = self.grid.step_perfect(self.chooser, self.cycles)   # :118-119
            else:
                n, _l = self.grid.step_quick(self.chooser)                  # :120
            rec = self.grid.scenario_step(self.sid, self.step_cur, self.rng)  # :121
            if rec and self.perfect:
                self.cycles = self.grid.find_cycles()                # :122-124
            self.step_cur += 1
            leaks += n                                               # :125
        self.n_cell_leaks = leaks
        return leaks
