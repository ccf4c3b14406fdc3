"""ctypes binding of include/flowamok_lys.h: the entry points a lys C `main` calls, for explorer/explorer.fut and
designer/designer.fut (libflowamok_explorer.so / libflowamok_designer.so, host code over the operator-level C ABI).

    p = Program("explorer", device=0)
    s = p.init(seed, 1080, 1920)          # futhark_entry_init
    s = p.key(s, SDLK_SPACE)              # futhark_entry_key (keydown); the old state is freed like lys does
    frame = p.render(s)                   # [h][w] argb

Nothing here computes; a missing library or device is an error.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _build
from .capi import FlowamokError

SDLK_SPACE, SDLK_1, SDLK_2, SDLK_a, SDLK_p, SDLK_r = 32, ord("1"), ord("2"), ord("a"), ord("p"), ord("r")
SDLK_RIGHT, SDLK_LEFT, SDLK_DOWN, SDLK_UP = 0x4000004F, 0x40000050, 0x40000051, 0x40000052

P, I32, I64, U32, U64, F32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_float

EXPORTS = ["futhark_context_config_new", "futhark_context_config_free", "futhark_context_config_set_debugging",
           "futhark_context_config_set_profiling", "futhark_context_config_set_logging", "futhark_context_config_set_device",
           "futhark_context_new", "futhark_context_free", "futhark_context_sync", "futhark_context_get_error",
           "futhark_context_config_set_tuning_param", "futhark_get_tuning_param_count", "futhark_get_tuning_param_name",
           "futhark_get_tuning_param_class", "futhark_context_config_set_cache_file", "futhark_context_set_logging_file",
           "futhark_context_pause_profiling", "futhark_context_unpause_profiling", "futhark_context_clear_caches",
           "futhark_context_report",
           "futhark_entry_init", "futhark_entry_resize", "futhark_entry_key", "futhark_entry_mouse", "futhark_entry_wheel",
           "futhark_entry_step", "futhark_entry_render", "futhark_entry_text_format", "futhark_entry_text_content",
           "futhark_entry_text_colour", "futhark_entry_grab_mouse", "futhark_free_opaque_state", "futhark_free_u32_2d",
           "futhark_values_u32_2d", "futhark_shape_u32_2d", "futhark_free_u8_1d", "futhark_values_u8_1d", "futhark_shape_u8_1d",
           "flowamok_lys_context", "flowamok_lys_state_grid", "flowamok_lys_state_fields"]

EXPLORER_FIELDS = ["h", "w", "scale", "auto", "steps_auto_per_second", "scenario_id", "steps", "perfect_simulation",
                   "n_cycles", "n_cell_leaks", "rng_state", "rng_inc"]
DESIGNER_FIELDS = ["h", "w", "gh", "gw", "scale", "running", "steps_per_second", "n_steps_total", "n_leaks", "cursor_some",
                   "cursor_y", "cursor_x", "rng_state", "rng_inc"]


def load(name: str) -> C.CDLL:
    """build (if needed) and load libflowamok_<name>.so with the signatures of include/flowamok_lys.h"""
    L = C.CDLL(str(_build.build_lys(name)))
    st, ctx = P, P
    sig = {
        "futhark_context_config_new": (P, []),
        "futhark_context_config_free": (None, [P]),
        "futhark_context_config_set_device": (None, [P, C.c_char_p]),
        "futhark_context_new": (P, [P]),
        "futhark_context_free": (None, [P]),
        "futhark_context_sync": (C.c_int, [P]),
        "futhark_context_get_error": (P, [P]),
        "futhark_entry_init": (C.c_int, [ctx, C.POINTER(P), U32, I64, I64]),
        "futhark_entry_resize": (C.c_int, [ctx, C.POINTER(P), I64, I64, st]),
        "futhark_entry_key": (C.c_int, [ctx, C.POINTER(P), I32, I32, st]),
        "futhark_entry_mouse": (C.c_int, [ctx, C.POINTER(P), I32, I32, I32, st]),
        "futhark_entry_wheel": (C.c_int, [ctx, C.POINTER(P), I32, I32, st]),
        "futhark_entry_step": (C.c_int, [ctx, C.POINTER(P), F32, st]),
        "futhark_entry_render": (C.c_int, [ctx, C.POINTER(P), st]),
        "futhark_entry_text_format": (C.c_int, [ctx, C.POINTER(P)]),
        "futhark_entry_text_colour": (C.c_int, [ctx, C.POINTER(U32), st]),
        "futhark_entry_grab_mouse": (C.c_int, [ctx, C.POINTER(C.c_bool)]),
        "futhark_free_opaque_state": (C.c_int, [ctx, st]),
        "futhark_free_u32_2d": (C.c_int, [ctx, P]),
        "futhark_values_u32_2d": (C.c_int, [ctx, P, P]),
        "futhark_shape_u32_2d": (C.POINTER(I64), [ctx, P]),
        "futhark_free_u8_1d": (C.c_int, [ctx, P]),
        "futhark_values_u8_1d": (C.c_int, [ctx, P, P]),
        "futhark_shape_u8_1d": (C.POINTER(I64), [ctx, P]),
        "flowamok_lys_context": (P, [ctx]),
        "flowamok_lys_state_grid": (P, [st]),
        "flowamok_lys_state_fields": (C.c_int, [st, C.POINTER(U64), C.c_int, C.POINTER(F32)]),
    }
    if name == "explorer":
        sig["futhark_entry_text_content"] = (C.c_int, [ctx, C.POINTER(I32), C.POINTER(I32), C.POINTER(I32), C.POINTER(I32),
                                                       C.POINTER(I64), C.POINTER(I32), C.POINTER(I64), C.POINTER(I64), F32, st])
    else:
        sig["futhark_entry_text_content"] = (C.c_int, [ctx, C.POINTER(I32), C.POINTER(I32), C.POINTER(I32), F32, st])
    for fn, (res, args) in sig.items():
        f = getattr(L, fn)
        f.restype, f.argtypes = res, args
    return L


class Program:
    """One lys program (explorer or designer) driven the way lys' C main drives the generated library."""

    def __init__(self, name: str, device: int = 0):
        assert name in ("explorer", "designer")
        self.name, self.L = name, load(name)
        cfg = self.L.futhark_context_config_new()
        self.L.futhark_context_config_set_device(cfg, f"#{device}".encode())
        self.ctx = self.L.futhark_context_new(cfg)
        self.L.futhark_context_config_free(cfg)
        if not self.L.flowamok_lys_context(self.ctx):
            msg = self.error()
            self.L.futhark_context_free(self.ctx)
            self.ctx = None
            raise FlowamokError(msg or "no device context")

    def close(self):
        if self.ctx:
            self.L.futhark_context_free(self.ctx)
            self.ctx = None

    def error(self) -> str:
        m = self.L.futhark_context_get_error(self.ctx)
        if not m:
            return ""
        s = C.string_at(m).decode(errors="replace")
        C.CDLL(None).free(C.c_void_p(m))
        return s

    def check(self, rc):
        if rc != 0:
            raise FlowamokError(self.error() or f"lys entry failed ({rc})")

    def _next(self, fn, old, *args, keep=False):
        out = P()
        self.check(fn(self.ctx, C.byref(out), *args, old))
        if not keep:
            self.free(old)
        return out

    # -- the entry points ---------------------------------------------------------------------------------------
    def init(self, seed, h, w):
        out = P()
        self.check(self.L.futhark_entry_init(self.ctx, C.byref(out), seed & 0xFFFFFFFF, h, w))
        return out

    def free(self, s):
        self.L.futhark_free_opaque_state(self.ctx, s)

    def resize(self, s, h, w, **kw):
        return self._next(self.L.futhark_entry_resize, s, h, w, **kw)

    def key(self, s, key, up=False, **kw):
        return self._next(self.L.futhark_entry_key, s, 1 if up else 0, key, **kw)

    def mouse(self, s, buttons, x, y, **kw):
        return self._next(self.L.futhark_entry_mouse, s, buttons, x, y, **kw)

    def wheel(self, s, dx, dy, **kw):
        return self._next(self.L.futhark_entry_wheel, s, dx, dy, **kw)

    def step(self, s, td, **kw):
        return self._next(self.L.futhark_entry_step, s, F32(td), **kw)

    def render(self, s) -> np.ndarray:
        arr = P()
        self.check(self.L.futhark_entry_render(self.ctx, C.byref(arr), s))
        shp = self.L.futhark_shape_u32_2d(self.ctx, arr)
        out = np.empty((shp[0], shp[1]), np.uint32)
        self.check(self.L.futhark_values_u32_2d(self.ctx, arr, out.ctypes.data))
        self.check(self.L.futhark_context_sync(self.ctx))
        self.L.futhark_free_u32_2d(self.ctx, arr)
        return out

    def text_format(self) -> str:
        arr = P()
        self.check(self.L.futhark_entry_text_format(self.ctx, C.byref(arr)))
        n = self.L.futhark_shape_u8_1d(self.ctx, arr)[0]
        buf = np.empty(n, np.uint8)
        self.check(self.L.futhark_values_u8_1d(self.ctx, arr, buf.ctypes.data))
        self.L.futhark_free_u8_1d(self.ctx, arr)
        return buf.tobytes().decode()

    def text_content(self, s, fps: float) -> tuple:
        if self.name == "explorer":
            o = [I32(), I32(), I32(), I32(), I64(), I32(), I64(), I64()]
        else:
            o = [I32(), I32(), I32()]
        self.check(self.L.futhark_entry_text_content(self.ctx, *[C.byref(v) for v in o], F32(fps), s))
        return tuple(v.value for v in o)

    def text_colour(self, s) -> int:
        v = U32()
        self.check(self.L.futhark_entry_text_colour(self.ctx, C.byref(v), s))
        return v.value

    def grab_mouse(self) -> bool:
        v = C.c_bool(True)
        self.check(self.L.futhark_entry_grab_mouse(self.ctx, C.byref(v)))
        return v.value

    # -- extensions ----------------------------------------------------------------------------------------------
    def fields(self, s) -> dict:
        names = EXPLORER_FIELDS if self.name == "explorer" else DESIGNER_FIELDS
        buf = (U64 * 16)()
        tu = F32()
        n = self.L.flowamok_lys_state_fields(s, buf, 16, C.byref(tu))
        assert n == len(names)
        d = {k: int(buf[i]) for i, k in enumerate(names)}
        for k in d:                                           # i64 fields travel as bit patterns
            if k not in ("rng_state", "rng_inc") and d[k] >= 1 << 63:
                d[k] -= 1 << 64
        d["time_unused"] = np.float32(tu.value)
        return d

    def grid(self, s, ctx):
        """the grid a state shows, as a borrowed flowamok_b200.Grid of `ctx` (a Context wrapping flowamok_lys_context)"""
        from . import Grid
        h = self.L.flowamok_lys_state_grid(s)
        return Grid._borrowed(ctx, h)

    def operator_context(self):
        from . import Context
        return Context._borrowed(self.L.flowamok_lys_context(self.ctx))
