"""CPU checks of the product's bit-plane programs (flowamok_b200/csrc/fa_stream.h and fa_tiles.h, compiled for the
host by tests/emul; the warp programs run with their 32 lanes as coroutines) against the oracle: same code the CUDA
kernels run, every tick, full state.  The variants change the strip width / height, the number of Jacobi levels, the
k_move tile shape, the arrival-list capacity and the frontier visiting order."""
import numpy as np
import pytest

import synth_ref
from util import O, EmulGrid, assert_state_equal, oracle_grid, random_grid, state_of

SHAPES = [(5, 7), (17, 31), (30, 32), (33, 33), (64, 65), (70, 100), (1, 40), (40, 1), (97, 130)]


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_quick_random_grids(variant):
    for seed, (gh, gw) in enumerate(SHAPES):
        arrs = random_grid(100 + seed, gh, gw)
        og, eg, ch = oracle_grid(arrs), EmulGrid(arrs), O.Chooser(O.CHOOSE_RANDOM)
        for t in range(10):
            nl, _ = og.step_quick(ch)
            assert nl == eg.step(variant=variant)
            assert_state_equal(state_of(og), eg.state(), f"{gh}x{gw} variant {variant} tick {t}")


def test_wide_grid_crosses_strip_and_tile_borders():
    """production strip shape on a grid wider than two strips (30 plane words each) and than one k_move tile: the
    halo lanes, the tile-wide colour lists and the per-strip frontier staging all see a neighbour"""
    arrs = random_grid(21, 40, 2300, n_lines=500, car_p=0.5)
    og, eg, ch = oracle_grid(arrs), EmulGrid(arrs), O.Chooser(O.CHOOSE_RANDOM)
    for t in range(5):
        nl, _ = og.step_quick(ch)
        assert nl == eg.step(variant=0)
        assert_state_equal(state_of(og), eg.state(), f"tick {t}")


def test_dense_jam_needs_many_passes():
    arrs = random_grid(7, 64, 100, car_p=0.9, weird_p=0.0, offroad_p=0.0)
    og, eg, ch = oracle_grid(arrs), EmulGrid(arrs), O.Chooser(O.CHOOSE_RANDOM)
    ref_passes = og.clone().move_until_fixpoint()
    assert ref_passes > 10
    for t in range(6):
        nl, _ = og.step_quick(ch)
        assert nl == eg.step(variant=1)
        assert_state_equal(state_of(og), eg.state(), f"tick {t}")
    assert eg.stats[0] > 6 * 10     # long queues need many worklist iterations


@pytest.mark.parametrize("chooser", [O.CHOOSE_Y, O.CHOOSE_X, O.CHOOSE_TAPE])
def test_choosers(chooser):
    arrs = random_grid(11, 40, 70, car_p=0.5)
    tape = np.array([1, 0, 0, 1, 1, 0, 1], np.uint8)
    og, eg = oracle_grid(arrs), EmulGrid(arrs)
    ch = O.Chooser(chooser, tape if chooser == O.CHOOSE_TAPE else None)
    for t in range(12):
        nl, _ = og.step_quick(ch)
        assert nl == eg.step(variant=2, chooser=chooser, tape_bit=int(tape[t % len(tape)]))
        assert_state_equal(state_of(og), eg.state(), f"tick {t}")


@pytest.mark.parametrize("sid", list(range(14)))
def test_perfect_scenarios(sid):
    """BASELINE config 2 shape: 108x192 (explorer.fut:48-49), perfect mode, scenario.step edits applied."""
    ex = O.Explorer(123, sid, perfect=True)
    rings = ex.grid.rings_from_masks(ex.cycles)
    eg = EmulGrid(ex.grid.arrays())
    for t in range(40):
        before = ex.grid.arrays()
        nl, _ = ex.grid.step_perfect(ex.chooser, ex.cycles)
        ne = eg.step(variant=0, rings=rings)
        assert nl == ne
        assert_state_equal(state_of(ex.grid), eg.state(), f"sid {sid} tick {t}")
        rec = ex.grid.scenario_step(sid, ex.step_cur, ex.rng)
        ex.step_cur += 1
        after = ex.grid.arrays()
        if rec:
            ex.cycles = ex.grid.find_cycles()
        if rec or any(not np.array_equal(before[k], after[k]) for k in ("uy", "ux")) or True:
            eg = EmulGrid(after)                # host-side scenario edits: re-upload (cheap at 108x192)
            rings = ex.grid.rings_from_masks(ex.cycles)


def test_synthetic_network_quick_and_perfect():
    arrs, rings = synth_ref.generate(100, 150, seed=2)
    for perfect in (False, True):
        og, eg, ch = oracle_grid(arrs), EmulGrid(arrs), O.Chooser(O.CHOOSE_RANDOM)
        for t in range(25):
            if perfect:
                nl, _ = og.step_perfect_sparse(ch, rings)
                ne = eg.step(variant=0, rings=rings)
            else:
                nl, _ = og.step_quick(ch)
                ne = eg.step(variant=0)
            assert nl == ne
            assert_state_equal(state_of(og), eg.state(), f"perfect={perfect} tick {t}")
