"""ctypes binding of include/flowamok.h -- the same stub a reference-side maintainer would write
(see INTEGRATION.md).  Nothing here computes: every call goes to libflowamok_b200.so, and a missing
library or a missing CUDA device is a hard error (there is no CPU path)."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

from . import _build

_lib = None

I64, I32, U64, U32, I8, U8 = C.c_int64, C.c_int32, C.c_uint64, C.c_uint32, C.c_int8, C.c_uint8
P = C.c_void_p


class FlowamokError(RuntimeError):
    pass


def lib():
    """Load the in-tree CUDA library (building it first if sources are newer and nvcc exists)."""
    global _lib
    if _lib is not None:
        return _lib
    so = _build.SO
    override = os.environ.get("FLOWAMOK_B200_LIB")   # kernel-tuning builds (tools/tune_variants.py); never a CPU path
    if override:
        so = Path(override)
    elif _build.needs_build():
        try:
            _build.build()
        except Exception as e:  # pragma: no cover
            if not so.exists():
                raise FlowamokError(f"libflowamok_b200.so is missing and could not be built: {e}") from e
    L = C.CDLL(str(so))
    sig = {
        "flowamok_context_new": (C.c_int, [C.POINTER(P), C.c_int, P]),
        "flowamok_context_free": (C.c_int, [P]),
        "flowamok_context_sync": (C.c_int, [P]),
        "flowamok_context_get_error": (P, [P]),
        "flowamok_context_device": (C.c_int, [P]),
        "flowamok_set_jam_threshold": (C.c_int, [P, C.c_double]),
        "flowamok_jam_ticks": (C.c_int, [P, P, C.POINTER(I64)]),
        "flowamok_grid_new": (C.c_int, [P, C.POINTER(P), I64, I64]),
        "flowamok_grid_free": (C.c_int, [P, P]),
        "flowamok_grid_recycle": (C.c_int, [P, P]),
        "flowamok_grid_shape": (C.c_int, [P, P, C.POINTER(I64), C.POINTER(I64)]),
        "flowamok_create_grid": (C.c_int, [P, P, I32, C.POINTER(U64), C.POINTER(U64)]),
        "flowamok_grid_upload": (C.c_int, [P, P, P, P, P, P, P, P, P, U64]),
        "flowamok_grid_download": (C.c_int, [P, P, P, P, P, P, P, P, P, C.POINTER(U64)]),
        "flowamok_grid_rng_corners": (C.c_int, [P, P, I64, P, P, C.POINTER(I64)]),
        "flowamok_grid_rng_scatter": (C.c_int, [P, P, I64, P, P]),
        "flowamok_add_line_vertical": (C.c_int, [P, P, I64, I64, I64, I8]),
        "flowamok_add_line_horizontal": (C.c_int, [P, P, I64, I64, I64, I8]),
        "flowamok_add_cell": (C.c_int, [P, P, I64, I64, U32, I8, I8]),
        "flowamok_set_chooser": (C.c_int, [P, P, C.c_int, P, I64]),
        "flowamok_set_cycles": (C.c_int, [P, P, I64, P, P, P, P, P, P]),
        "flowamok_find_cycles": (C.c_int, [P, P, I64, C.POINTER(I64)]),
        "flowamok_cycle_count": (C.c_int, [P, P, C.POINTER(I64), C.POINTER(I64)]),
        "flowamok_get_cycles_dense": (C.c_int, [P, P, P]),
        "flowamok_step_quick": (C.c_int, [P, P, I64, C.POINTER(I64)]),
        "flowamok_step_perfect": (C.c_int, [P, P, I64, C.POINTER(I64)]),
        "flowamok_step_profile": (C.c_int, [P, P, C.c_int, I64, C.POINTER(C.c_double)]),
        "flowamok_step_timeline": (C.c_int, [P, P, C.c_int, I64, C.POINTER(C.c_double)]),
        "flowamok_step_leaks": (C.c_int, [P, P, C.c_int, I64, P, C.POINTER(I64)]),
        "flowamok_step_leaks_cells": (C.c_int, [P, P, C.c_int, I64, P, C.POINTER(I64)]),
        "flowamok_render": (C.c_int, [P, P, I64, P]),
        "flowamok_render_device": (C.c_int, [P, P, I64, C.POINTER(P)]),
        "flowamok_grid_clone": (C.c_int, [P, P, C.POINTER(P)]),
        "flowamok_stats": (C.c_int, [P, P, C.POINTER(I64), C.POINTER(I64), C.POINTER(I64)]),
        "flowamok_generate_synthetic": (C.c_int, [P, P, U64, U32, U32]),
        "flowamok_grid_digest": (C.c_int, [P, P, C.POINTER(U64), C.POINTER(U64)]),
        "flowamok_stats_hist": (C.c_int, [P, P, P]),
        "flowamok_band_p2p_stats": (C.c_int, [P, P, C.POINTER(I64), C.POINTER(I64)]),
        "flowamok_band_new": (C.c_int, [P, C.POINTER(P), I64, I64, I64, I64]),
        "flowamok_band_msg_bytes": (C.c_int, [P, P, C.POINTER(I64), C.POINTER(I64)]),
        "flowamok_band_pack_state": (C.c_int, [P, P, C.c_int, P]),
        "flowamok_band_unpack_state": (C.c_int, [P, P, C.c_int, P]),
        "flowamok_band_u_begin": (C.c_int, [P, P]),
        "flowamok_band_u_init": (C.c_int, [P, P]),
        "flowamok_band_u_iter": (C.c_int, [P, P]),
        "flowamok_band_pack_u": (C.c_int, [P, P, C.c_int, P]),
        "flowamok_band_merge_u": (C.c_int, [P, P, C.c_int, P, C.c_int, P]),
        "flowamok_band_woken": (C.c_int, [P, P, P]),
        "flowamok_band_u_resume": (C.c_int, [P, P]),
        "flowamok_band_rings": (C.c_int, [P, P]),
        "flowamok_band_move": (C.c_int, [P, P]),
        "flowamok_band_leaks": (C.c_int, [P, P, C.POINTER(I64)]),
        "flowamok_band_comm_id": (C.c_int, [P]),
        "flowamok_band_comm_init": (C.c_int, [P, P, C.c_int, C.c_int]),
        "flowamok_band_step": (C.c_int, [P, P, C.c_int, I64, C.POINTER(I64)]),
        "flowamok_grid_upload_cars": (C.c_int, [P, P, P, P, P, P, P]),
        "flowamok_band_set_cuts": (C.c_int, [P, P, C.c_int, P]),
        "flowamok_band_p2p_open": (C.c_int, [P, P, C.c_int, C.c_int, P, C.POINTER(P)]),
        "flowamok_band_p2p_connect": (C.c_int, [P, P, P, P, C.c_int]),
        "flowamok_band_p2p_close": (C.c_int, [P, P]),
        "flowamok_band_step_profile": (C.c_int, [P, P, C.c_int, I64, C.POINTER(C.c_double)]),
        "flowamok_band_ring_words": (C.c_int, [P, P, C.POINTER(I64)]),
        "flowamok_band_ring_verdicts": (C.c_int, [P, P, P]),
        "flowamok_band_rings_apply": (C.c_int, [P, P, P]),
        "flowamok_scenario_init": (C.c_int, [P, P, C.c_int]),
        "flowamok_explorer_step": (C.c_int, [P, P, C.c_int, C.c_int, I64, C.POINTER(I64), C.POINTER(U64), C.POINTER(U64), C.POINTER(I64)]),
        "flowamok_designer_step": (C.c_int, [P, P, I64, C.POINTER(I64), C.POINTER(U64), C.POINTER(U64), C.POINTER(I32)]),
        "flowamok_version": (C.c_char_p, []),
        # Futhark-generated-library surface (include/flowamok_futhark.h)
        "futhark_context_config_new": (P, []),
        "futhark_context_config_free": (None, [P]),
        "futhark_context_config_set_debugging": (None, [P, C.c_int]),
        "futhark_context_config_set_profiling": (None, [P, C.c_int]),
        "futhark_context_config_set_logging": (None, [P, C.c_int]),
        "futhark_context_config_set_device": (None, [P, C.c_char_p]),
        "flowamok_futhark_set_scenario": (None, [P, C.c_int]),
        "futhark_context_new": (P, [P]),
        "futhark_context_free": (None, [P]),
        "futhark_context_sync": (C.c_int, [P]),
        "futhark_context_get_error": (P, [P]),
        "futhark_entry_init": (C.c_int, [P, C.POINTER(P), I64, I64, I64, I32]),
        "futhark_entry_step": (C.c_int, [P, C.POINTER(P), C.POINTER(P), P]),
        "futhark_free_opaque_state": (C.c_int, [P, P]),
        "futhark_new_u32_2d": (P, [P, P, I64, I64]),
        "futhark_free_u32_2d": (C.c_int, [P, P]),
        "futhark_values_u32_2d": (C.c_int, [P, P, P]),
        "futhark_shape_u32_2d": (C.POINTER(I64), [P, P]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    L._libc = C.CDLL(None)
    L._libc.free.argtypes = [P]
    _lib = L
    return L
