"""flowamok_b200 -- B200-native per-tick grid update of nqpz/flowamok.

Host-side mirror of the reference's operator interface for the hot path:

    reference (Futhark)                                   here
    ----------------------------------------------------  -------------------------------------------
    create_grid gh gw () rng           utils.fut:14       Context.create_grid(gh, gw, seed) -> Grid
    add_line_vertical/_horizontal      utils.fut:19,27    Grid.add_line_vertical / add_line_horizontal
    add_cell                           utils.fut:35       Grid.add_cell
    stepper_perfect.find_cycles        steppers.fut:233   Grid.find_cycles()
    stepper_quick.step                 steppers.fut:177   stepper_quick.step(grid, ...)
    stepper_perfect.step               steppers.fut:302   stepper_perfect.step(grid, ...)
    choose_direction_random            flowamok.fut:10    chooser=CHOOSE_RANDOM (default)
    render                             flowamok.fut:16    Grid.render(scale)

Everything executes in the CUDA library behind include/flowamok.h; this module only marshals
arguments.  Grids are opaque device values, consumed and replaced by `step` like the reference's
unique arrays.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .capi import FlowamokError, lib
from . import _build as _buildmod

CHOOSE_RANDOM, CHOOSE_Y, CHOOSE_X, CHOOSE_TAPE = 0, 1, 2, 3
PCG_INC = ((0xda3e39cb94b95bdb << 1) | 1) & 0xFFFFFFFFFFFFFFFF

__all__ = ["Context", "Grid", "stepper_quick", "stepper_perfect", "FlowamokError", "build",
           "CHOOSE_RANDOM", "CHOOSE_Y", "CHOOSE_X", "CHOOSE_TAPE"]


def build(force: bool = False):
    return _buildmod.build(force=force)


def _ptr(a):
    return None if a is None else a.ctypes.data


class Context:
    """flowamok_context: one CUDA device + stream (the analogue of futhark_context)."""

    def __init__(self, device: int = -1, stream: int | None = None):
        self.L = lib()
        self.h = C.c_void_p()
        self.stream_ptr = int(stream) if stream else -1      # -1: the library's private stream
        rc = self.L.flowamok_context_new(C.byref(self.h), device, C.c_void_p(stream) if stream else None)
        if rc:
            msg = self._err()
            self.L.flowamok_context_free(self.h)
            self.h = None
            raise FlowamokError(msg or "flowamok_context_new failed")

    def _err(self):
        p = self.L.flowamok_context_get_error(self.h)
        if not p:
            return ""
        s = C.string_at(p).decode()
        self.L._libc.free(p)
        return s

    def check(self, rc):
        if rc:
            raise FlowamokError(self._err() or f"error {rc}")

    def sync(self):
        self.check(self.L.flowamok_context_sync(self.h))

    @property
    def device(self):
        return self.L.flowamok_context_device(self.h)

    @classmethod
    def _borrowed(cls, handle) -> "Context":
        """a view of a flowamok_context somebody else owns (flowamok_lys_context); never freed from here"""
        c = cls.__new__(cls)
        c.L, c.h, c.stream_ptr, c._owned = lib(), C.c_void_p(handle), -1, False
        return c

    def close(self):
        if self.h and getattr(self, "_owned", True):
            self.L.flowamok_context_free(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- constructors of grid values ---
    def grid(self, gh, gw) -> "Grid":
        return Grid(self, gh, gw)

    def create_grid(self, gh, gw, seed) -> "Grid":
        """create_grid gh gw () (rnge.rng_from_seed [seed])  (utils.fut:14-17, animation.fut:15)"""
        g = Grid(self, gh, gw)
        g.create(seed)
        return g

    def grid_from_arrays(self, **arrs) -> "Grid":
        gh, gw = arrs["uy"].shape
        g = Grid(self, gh, gw)
        g.upload(**arrs)
        return g


class Grid:
    """Opaque device value `[gh][gw](cell ())` (model.fut:19-24) plus its cycle_checks."""

    def __init__(self, ctx: Context, gh, gw):
        self.ctx, self.L = ctx, ctx.L
        self.gh, self.gw = int(gh), int(gw)
        self.h = C.c_void_p()
        ctx.check(self.L.flowamok_grid_new(ctx.h, C.byref(self.h), self.gh, self.gw))
        self.joined_rng = None

    @classmethod
    def _borrowed(cls, ctx: Context, handle) -> "Grid":
        """a view of a grid value somebody else owns (flowamok_lys_state_grid); never freed from here"""
        g = cls.__new__(cls)
        g.ctx, g.L, g.h, g.joined_rng, g._owned = ctx, ctx.L, C.c_void_p(handle), None, False
        gh, gw = C.c_int64(), C.c_int64()
        ctx.check(g.L.flowamok_grid_shape(ctx.h, g.h, C.byref(gh), C.byref(gw)))
        g.gh, g.gw = gh.value, gw.value
        return g

    def free(self):
        if self.h and self.ctx.h and getattr(self, "_owned", True):
            self.L.flowamok_grid_free(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    # --- utils.fut ---
    def create(self, seed, want_joined=False):
        js, ji = C.c_uint64(), C.c_uint64()
        self.ctx.check(self.L.flowamok_create_grid(self.ctx.h, self.h, int(np.int32(seed)),
                                                   C.byref(js) if want_joined else None,
                                                   C.byref(ji) if want_joined else None))
        if want_joined:
            self.joined_rng = (js.value, ji.value)

    def add_line_vertical(self, y_start, y_end, x, flow):
        self.ctx.check(self.L.flowamok_add_line_vertical(self.ctx.h, self.h, y_start, y_end, x, flow))

    def add_line_horizontal(self, y, x_start, x_end, flow):
        self.ctx.check(self.L.flowamok_add_line_horizontal(self.ctx.h, self.h, y, x_start, x_end, flow))

    def add_cell(self, y, x, color, dy, dx):
        self.ctx.check(self.L.flowamok_add_cell(self.ctx.h, self.h, y, x, color & 0xFFFFFFFF, dy, dx))

    # --- host <-> device in the reference's record-of-arrays layout ---
    def upload(self, uy=None, ux=None, dy=None, dx=None, color=None, pref=None, rng_state=None,
               rng_inc=PCG_INC, **_ignored):
        def prep(a, dt):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=dt)
            assert a.shape == (self.gh, self.gw), (a.shape, (self.gh, self.gw))
            return a
        arrs = [prep(uy, np.int8), prep(ux, np.int8), prep(dy, np.int8), prep(dx, np.int8),
                prep(color, np.uint32), prep(pref, np.uint8), prep(rng_state, np.uint64)]
        inc = int(np.asarray(rng_inc).flat[0]) if rng_inc is not None else PCG_INC
        self.ctx.check(self.L.flowamok_grid_upload(self.ctx.h, self.h, *[_ptr(a) for a in arrs], inc))

    def download(self, fields=("uy", "ux", "dy", "dx", "color", "pref", "rng_state")):
        dts = dict(uy=np.int8, ux=np.int8, dy=np.int8, dx=np.int8, color=np.uint32, pref=np.uint8,
                   rng_state=np.uint64)
        out = {k: np.zeros((self.gh, self.gw), dts[k]) for k in fields}
        inc = C.c_uint64()
        args = [_ptr(out.get(k)) for k in ("uy", "ux", "dy", "dx", "color", "pref", "rng_state")]
        self.ctx.check(self.L.flowamok_grid_download(self.ctx.h, self.h, *args, C.byref(inc)))
        out["rng_inc"] = inc.value
        return out

    def rng_corners(self, idx=None, state=None):
        """(flat index, pcg32 state) of the cells with a road on both axes -- the only ones whose state can change.
        idx / state: preallocated (e.g. pinned) arrays to fill; otherwise fresh ones are returned."""
        n = C.c_int64()
        if idx is None:
            self.ctx.check(self.L.flowamok_grid_rng_corners(self.ctx.h, self.h, 0, None, None, C.byref(n)))
            idx, state = np.zeros(n.value, np.int64), np.zeros(n.value, np.uint64)
        self.ctx.check(self.L.flowamok_grid_rng_corners(self.ctx.h, self.h, len(idx), _ptr(idx), _ptr(state), C.byref(n)))
        return idx[:n.value], state[:n.value]

    def rng_scatter(self, idx, state):
        self.ctx.check(self.L.flowamok_grid_rng_scatter(self.ctx.h, self.h, len(idx), _ptr(idx), _ptr(state)))

    # --- chooser / cycles ---
    def set_chooser(self, mode, tape=None):
        t = None if tape is None else np.ascontiguousarray(tape, np.uint8)
        self.ctx.check(self.L.flowamok_set_chooser(self.ctx.h, self.h, mode, _ptr(t), 0 if t is None else len(t)))

    def set_cycles(self, off, y, x, cy, cx, grant):
        a = [np.ascontiguousarray(off, np.int64), np.ascontiguousarray(y, np.int64),
             np.ascontiguousarray(x, np.int64), np.ascontiguousarray(cy, np.int8),
             np.ascontiguousarray(cx, np.int8), np.ascontiguousarray(grant, np.uint8)]
        self.ctx.check(self.L.flowamok_set_cycles(self.ctx.h, self.h, len(a[0]) - 1, *[_ptr(v) for v in a]))

    def find_cycles(self, max_paths=0) -> int:
        """stepper_perfect.find_cycles (steppers.fut:233-300); the result becomes this grid's cycle_checks."""
        n = C.c_int64()
        self.ctx.check(self.L.flowamok_find_cycles(self.ctx.h, self.h, max_paths, C.byref(n)))
        return n.value

    def cycle_count(self):
        n, c = C.c_int64(), C.c_int64()
        self.ctx.check(self.L.flowamok_cycle_count(self.ctx.h, self.h, C.byref(n), C.byref(c)))
        return n.value, c.value

    def cycles_dense(self):
        n, _ = self.cycle_count()
        m = np.zeros((n, self.gh, self.gw, 2), np.int8)
        if n:
            self.ctx.check(self.L.flowamok_get_cycles_dense(self.ctx.h, self.h, m.ctypes.data))
        return m

    # --- stepping ---
    def step(self, perfect: bool, n_ticks=1, leaks=True):
        nl = C.c_int64()
        f = self.L.flowamok_step_perfect if perfect else self.L.flowamok_step_quick
        self.ctx.check(f(self.ctx.h, self.h, n_ticks, C.byref(nl) if leaks else None))
        return nl.value if leaks else None

    def step_profile(self, perfect: bool, n_ticks=10):
        ms = (C.c_double * 4)()
        self.ctx.check(self.L.flowamok_step_profile(self.ctx.h, self.h, int(perfect), n_ticks, ms))
        return dict(k_u_local=ms[0], k_u_iter=ms[1], k_rings=ms[2], k_move=ms[3])

    def step_timeline(self, perfect: bool, n_ticks=10):
        ms = (C.c_double * 5)()
        self.ctx.check(self.L.flowamok_step_timeline(self.ctx.h, self.h, int(perfect), n_ticks, ms))
        return dict(local=ms[0], fork_to_fixpoint_end=ms[1], fork_to_move_end=ms[2], fixup=ms[3], tick=ms[4])

    def step_leaks(self, perfect: bool, cap=1 << 20):
        buf = np.zeros((cap, 8), np.int64)
        n = C.c_int64()
        self.ctx.check(self.L.flowamok_step_leaks(self.ctx.h, self.h, int(perfect), cap, buf.ctypes.data, C.byref(n)))
        return n.value, buf[:min(n.value, cap)]

    def step_leaks_cells(self, perfect: bool, cap=1 << 20):
        """one tick; the leaks with the whole source cell: rows of 14 int64 (include/flowamok.h)"""
        buf = np.zeros((cap, 14), np.int64)
        n = C.c_int64()
        self.ctx.check(self.L.flowamok_step_leaks_cells(self.ctx.h, self.h, int(perfect), cap, buf.ctypes.data, C.byref(n)))
        return n.value, buf[:min(n.value, cap)]

    def render(self, scale):
        px = np.zeros((self.gh * scale, self.gw * scale), np.uint32)
        self.ctx.check(self.L.flowamok_render(self.ctx.h, self.h, scale, px.ctypes.data))
        return px

    def stats(self):
        t, p, s = C.c_int64(), C.c_int64(), C.c_int64()
        self.ctx.check(self.L.flowamok_stats(self.ctx.h, self.h, C.byref(t), C.byref(p), C.byref(s)))
        return dict(ticks=t.value, passes=p.value, sweeps=s.value)

    def stats_hist(self):
        """ticks by number of fixpoint iterations (index 31: 31 and more), since the grid was created"""
        h = (C.c_int64 * 32)()
        self.ctx.check(self.L.flowamok_stats_hist(self.ctx.h, self.h, h))
        return list(h)

    def digest(self):
        """(sum, xor) mod 2^64 of a position-dependent hash of the owned cells' state; bands' digests combine"""
        a, b = C.c_uint64(), C.c_uint64()
        self.ctx.check(self.L.flowamok_grid_digest(self.ctx.h, self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def generate_synthetic(self, seed=1, density16=6554, ring_full16=2):
        self.ctx.check(self.L.flowamok_generate_synthetic(self.ctx.h, self.h, seed, density16, ring_full16))


class _Stepper:
    def __init__(self, perfect):
        self.perfect = perfect

    def step(self, grid: Grid, n_ticks=1, leaks=True):
        """(cells, cell_leaks) = step gh gw choose_direction [cycle_checks] () cells -- the grid value is
        replaced in place; returns length cell_leaks summed over n_ticks (explorer.fut:125)."""
        return grid.step(self.perfect, n_ticks, leaks)


stepper_quick = _Stepper(False)     # src/steppers.fut:176-186
stepper_perfect = _Stepper(True)    # src/steppers.fut:188-313
