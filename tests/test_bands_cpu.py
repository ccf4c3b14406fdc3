"""world_size-2/3 gloo tests of the multi-GPU row-band path on the CPU: the product's BandRunner
(flowamok_b200/bands.py) drives bands backed by the CPU emulation of the CUDA strip programs (tests/emul), each
rank in its own process; the gathered result must equal the oracle's single-grid run bit for bit."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import synth_ref
from util import O, STATE_FIELDS, emul_lib, oracle_grid, random_grid, state_of


class EmulBand:
    """Same interface as flowamok_b200.bands.DeviceBand, on host memory."""

    def __init__(self, arrs, row_begin, row_end, variant, chooser=O.CHOOSE_RANDOM, tape=None, rings=None):
        self.L = emul_lib()
        L = self.L
        L.emul_band_new.restype = C.c_void_p
        L.emul_band_new.argtypes = [C.c_int] * 5
        for f, at in (("emul_band_msg_bytes", [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
                      ("emul_band_pack_state", [C.c_void_p, C.c_int, C.c_void_p]),
                      ("emul_band_unpack_state", [C.c_void_p, C.c_int, C.c_void_p]),
                      ("emul_band_u_init", [C.c_void_p, C.c_int, C.c_uint32]),
                      ("emul_band_pack_u", [C.c_void_p, C.c_int, C.c_void_p]),
                      ("emul_band_merge_u", [C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
                      ("emul_band_u_resume", [C.c_void_p]), ("emul_band_rings", [C.c_void_p]),
                      ("emul_band_ring_words", [C.c_void_p]), ("emul_band_ring_verdicts", [C.c_void_p, C.c_void_p]),
                      ("emul_band_rings_apply", [C.c_void_p, C.c_void_p]),
                      ("emul_band_move", [C.c_void_p, C.c_int, C.c_uint32]),
                      ("emul_set_rings", [C.c_void_p, C.c_int64] + [C.c_void_p] * 6)):
            getattr(L, f).argtypes = at
        L.emul_band_merge_u.restype = C.c_int64
        L.emul_band_ring_words.restype = C.c_int64
        L.emul_band_move.restype = C.c_int64
        self.gh, self.gw = arrs["uy"].shape
        self.p = L.emul_band_new(self.gh, self.gw, row_begin, row_end, variant)
        c = {k: np.ascontiguousarray(v) for k, v in arrs.items()}
        L.emul_upload(self.p, c["uy"].ctypes.data, c["ux"].ctypes.data, c["dy"].ctypes.data, c["dx"].ctypes.data,
                      c["color"].ctypes.data, c["pref"].ctypes.data, c["rng_state"].ctypes.data,
                      int(c["rng_inc"].flat[0]))
        sb, ub = C.c_int64(), C.c_int64()
        L.emul_band_msg_bytes(self.p, C.byref(sb), C.byref(ub))
        self.state_bytes, self.u_bytes = sb.value, ub.value
        self.chooser, self.tape, self.tick_no, self.woken_n, self.leaks_n = chooser, tape, 0, 0, 0
        if rings is not None:      # the GLOBAL list: every band judges / marks the cells of its own rows
            off, idx, cy, cx, grant = rings
            a = [np.ascontiguousarray(off, np.int64), np.ascontiguousarray(idx // self.gw), np.ascontiguousarray(idx % self.gw),
                 np.ascontiguousarray(cy), np.ascontiguousarray(cx), np.ascontiguousarray(grant)]
            self._rings = a
            L.emul_set_rings(self.p, len(off) - 1, *[v.ctypes.data for v in a])

    def _tb(self):
        return int(self.tape[self.tick_no % len(self.tape)]) if self.tape is not None else 0

    def buffer(self, n):
        return torch.zeros(n, dtype=torch.uint8)

    def counter(self):
        return torch.zeros(1, dtype=torch.int32)

    def pack_state(self, side, buf):
        self.L.emul_band_pack_state(self.p, side, buf.data_ptr())

    def unpack_state(self, side, buf):
        self.L.emul_band_unpack_state(self.p, side, buf.data_ptr())

    def u_init(self):
        self.L.emul_band_u_init(self.p, self.chooser, self._tb())

    def u_iter(self):
        self.L.emul_band_u_resume(self.p)

    def pack_u(self, side, buf):
        self.L.emul_band_pack_u(self.p, side, buf.data_ptr())

    def merge_u(self, side, buf, wake):
        self.woken_n = int(self.L.emul_band_merge_u(self.p, side, buf.data_ptr(), int(wake)))

    def woken(self, counter):
        counter[0] = self.woken_n
        self.woken_n = 0

    def u_resume(self):
        self.L.emul_band_u_resume(self.p)

    def rings(self):
        self.L.emul_band_rings(self.p)

    def ring_words(self):
        return int(self.L.emul_band_ring_words(self.p))

    def ring_buffer(self):
        return torch.zeros(max(1, self.ring_words()), dtype=torch.int32)

    def ring_verdicts(self, buf):
        self.L.emul_band_ring_verdicts(self.p, buf.data_ptr())

    def rings_apply(self, buf):
        self.L.emul_band_rings_apply(self.p, buf.data_ptr())

    def move(self):
        self.leaks_n += int(self.L.emul_band_move(self.p, self.chooser, self._tb()))
        self.tick_no += 1

    def owned_state(self, out):
        n = (self.gh, self.gw)
        self.L.emul_download(self.p, out["dy"].ctypes.data, out["dx"].ctypes.data, out["color"].ctypes.data,
                             out["pref"].ctypes.data, out["rng_state"].ctypes.data)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, ret):
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    for q in (str(root), str(root / "tests")):
        if q not in sys.path:
            sys.path.insert(0, q)
    from flowamok_b200.bands import BandRunner, split_rows
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        arrs, rings, perfect, ticks, variant, align = case()
        gh, gw = arrs["uy"].shape
        lo, hi = split_rows(gh, world, align)[rank]
        band = EmulBand(arrs, lo, hi, variant, rings=rings if perfect else None)
        run = BandRunner(band, rank, world)
        run.step(perfect, ticks)
        out = {k: np.zeros((gh, gw), dt) for k, dt in (("dy", np.int8), ("dx", np.int8), ("color", np.uint32),
                                                         ("pref", np.uint8), ("rng_state", np.uint64))}
        band.owned_state(out)
        ret[rank] = (lo, hi, {k: v[lo:hi].copy() for k, v in out.items()}, band.leaks_n, run.rounds)
    finally:
        dist.destroy_process_group()


def _case_jam():
    arrs = random_grid(77, 66, 90, car_p=0.85, weird_p=0.02, offroad_p=0.01)
    return arrs, None, False, 6, 2, 1


def _case_random():
    arrs = random_grid(5, 48, 70, car_p=0.4)
    return arrs, None, False, 8, 1, 1


def _case_synth_perfect():
    arrs, rings = synth_ref.generate(96, 150, seed=3)
    return arrs, rings, True, 10, 0, 16


def _case_synth_perfect_straddle():
    """cuts that are NOT multiples of 16: rings straddle the cuts (per-ring pass bits AND-ed over the bands)"""
    arrs, rings = synth_ref.generate(100, 150, seed=3, ring_full16=12)
    return arrs, rings, True, 10, 0, 1


@pytest.mark.parametrize("world,case", [(2, _case_random), (3, _case_jam), (2, _case_synth_perfect),
                                        (2, _case_synth_perfect_straddle), (3, _case_synth_perfect_straddle)])
def test_bands_equal_single_grid(world, case):
    arrs, rings, perfect, ticks, variant, align = case()
    og, ch = oracle_grid(arrs), O.Chooser(O.CHOOSE_RANDOM)
    leaks = 0
    for _ in range(ticks):
        n, _l = og.step_perfect_sparse(ch, rings) if perfect else og.step_quick(ch)
        leaks += n
    want = state_of(og)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), case, ret), nprocs=world, join=True)
    assert len(ret) == world
    total_leaks, rounds = 0, 0
    for r in range(world):
        lo, hi, st, nl, nr = ret[r]
        total_leaks += nl
        rounds = max(rounds, nr)
        for k in STATE_FIELDS:
            assert np.array_equal(st[k], want[k][lo:hi]), f"rank {r} rows {lo}:{hi} field {k}"
    assert total_leaks == leaks
    assert rounds >= ticks            # at least one convergence round per tick


def test_band_whose_rows_plus_halo_rows_equal_the_grid_height():
    """regression (found by tests/fuzz_bands.py): a band of gh - 4 rows holds gh - 4 + 2 * 2 = gh rows with its halo --
    it must still be treated as a band (ghost rows, boundary rows on the frontier), not as the whole grid.  In-process
    driver (LocalBands), bands on the CPU emulation, against the oracle."""
    from flowamok_b200.bands import LocalBands
    gh, gw = 54, 82
    arrs = random_grid(238, gh, gw, n_lines=44, car_p=0.8, weird_p=0.05, offroad_p=0.02)
    for variant in (0, 2):
        bands = [EmulBand(arrs, 0, gh - 4, variant), EmulBand(arrs, gh - 4, gh, variant)]
        run = LocalBands(bands)
        og, ch = oracle_grid(arrs), O.Chooser(O.CHOOSE_RANDOM)
        for t in range(6):
            og.step_quick(ch)
            run.tick(False)
            want = state_of(og)
            got = {k: np.zeros((gh, gw), want[k].dtype) for k in STATE_FIELDS}
            for b in bands:
                b.owned_state(got)
            for k in STATE_FIELDS:
                assert np.array_equal(want[k], got[k]), f"variant {variant} tick {t} field {k}"
