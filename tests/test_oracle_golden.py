"""Pins the CPU oracle against the reference's own artefacts:
 * the README GIF (config 1, 400 frames) replayed with a tape-driven chooser,
 * hand-derivable scenario facts,
 * the 398-bit RNG known answer recovered from the GIF (expected to fail: cpprandom is not vendored)."""
import hashlib

import numpy as np
import pytest

from util import O, ROOT

TAPE = ("0111111001011111111101100111110101100100011010011001100101000011101010011110001011000000001100011100"
        "1000101011010101101101110111010000101101101001001000010011010001100101110011100110110110110100000101"
        "0110101100110001000010001100011011010011100000111011100011110001001011011000110010000111101010010011"
        "00111100111110010100111010010101111101001010111000001101101001111110111101011100100000111011000001")
CLASS_OF = {0xFFFF0000: 2, 0xFF0000FF: 3, 0xFFFFFF00: 4, 0xFFFF00FF: 5, 0xFF7C301C: 6}


def classes(arrs):
    occ = (arrs["dy"] != 0) | (arrs["dx"] != 0)
    road = (arrs["uy"] != 0) | (arrs["ux"] != 0)
    out = np.where(road, 1, 0).astype(np.uint8)
    for c, k in CLASS_OF.items():
        out[occ & (arrs["color"] == c)] = k
    return out


def test_tape_is_the_surveyed_stream():
    assert len(TAPE) == 398
    assert hashlib.sha256(TAPE.encode()).hexdigest() == "7369b5e92bc3b60ff4e69c9423d61b6f0f12c67e82d86bb30274b3122e2b8e46"


def test_readme_gif_replay_oracle():
    """BASELINE config 1: init 30 30 10 123, 400 steps of mk_anim overlapping_tight_cycles (README.fut:11-15)."""
    gold = np.load(ROOT / "tests/golden/readme_gif_classes.npz")["frames"]
    tape = np.array([int(c) for c in TAPE] + [0, 0], np.uint8)
    a = O.Anim(8, 30, 30, 10, 123, O.Chooser(O.CHOOSE_TAPE, tape))
    assert a.cycles.shape[0] == 2
    for k in range(400):
        assert np.array_equal(classes(a.grid.arrays()), gold[k]), f"frame {k}"
        a.step(render=False)


@pytest.mark.xfail(reason="diku-dk/cpprandom 1.3.1 is not vendored in the reference tree; the recalled pcg32 "
                          "split/seed arithmetic does not reproduce the GIF's stream: rng parity unpinned",
                   strict=False)
def test_rng_known_answer_from_gif():
    rng = O.Rng.from_seed([123]).split(900)[14 * 30 + 14]
    got = "".join(str(rng.dist_i8(0, 1)) for _ in range(398))
    assert got == TAPE


def _grid(gh=40, gw=40):
    g = O.Grid(gh, gw)
    g.create(O.Rng.from_seed([1]))
    return g


def test_empty_grid_two_passes_nothing_changes():
    g = _grid()
    before = g.arrays()
    assert g.clone().move_until_fixpoint() == 2
    g.step_quick(O.Chooser(O.CHOOSE_Y))
    after = g.arrays()
    for k in before:
        assert np.array_equal(before[k], after[k]), k


def test_single_car_moves_and_leaves():
    g = _grid()
    g.add_line_horizontal(5, 2, 39, 1)
    g.add_cell(5, 36, 0xFF123456, 0, 1)
    ch = O.Chooser(O.CHOOSE_Y)
    assert g.clone().move_until_fixpoint() == 3          # lone car on open road: L + 2 passes
    for x in (37, 38, 39):
        n, _ = g.step_quick(ch)
        assert n == 0 and g.field("dx")[5, x] == 1 and g.field("color")[5, x] == 0xFF123456
        assert g.field("dx")[5, x - 1] == 0 and g.field("color")[5, x - 1] == 0xFF123456   # stale colour stays
    n, leaks = g.step_quick(ch, leaks=True)              # drives off the grid: no leak record (:74, :117)
    assert n == 0 and leaks == [] and not g.field("dx").any()


def test_leak_into_non_road():
    g = _grid()
    g.add_line_horizontal(5, 2, 10, 1)
    g.add_cell(5, 10, 0xFFABCDEF, 0, 1)
    n, leaks = g.step_quick(O.Chooser(O.CHOOSE_Y), leaks=True)
    assert n == 1 and (leaks[0]["y_next"], leaks[0]["x_next"], leaks[0]["y_src"], leaks[0]["x_src"]) == (5, 11, 5, 10)
    assert leaks[0]["color"] == 0xFFABCDEF and leaks[0]["dx"] == 1
    assert not g.field("dx").any()


def test_queue_advances_together_and_passes():
    g = _grid()
    g.add_line_horizontal(7, 0, 39, 1)
    L = 6
    for i in range(L):
        g.add_cell(7, 10 + i, 0xFF000000 + i, 0, 1)
    assert g.clone().move_until_fixpoint() == L + 2
    g.step_quick(O.Chooser(O.CHOOSE_Y))
    assert list(np.nonzero(g.field("dx")[7])[0]) == list(range(11, 11 + L))


def test_merge_alternates_starting_with_y():
    g = _grid()
    g.add_line_horizontal(20, 0, 39, 1)
    g.add_line_vertical(0, 20, 20, 1)
    ch = O.Chooser(O.CHOOSE_X)
    came_from = []
    for t in range(8):
        g.add_cell(20, 19, 0xFF0000AA, 0, 1)   # x predecessor always occupied
        g.add_cell(19, 20, 0xFF00BB00, 1, 0)   # y predecessor always occupied
        g.add_cell(20, 20, 0xFFFFFFFF, 0, 0)   # keep the merge cell empty
        g.step_quick(ch)
        came_from.append({0xFF0000AA: "x", 0xFF00BB00: "y"}[int(g.field("color")[20, 20])])
    assert came_from == ["y", "x"] * 4


def test_full_ring_frozen_in_quick_rotates_in_perfect():
    for sid, n_cycles in ((6, 1), (7, 2), (8, 2)):
        g = O.Grid(30, 30)
        g.create(O.Rng.from_seed([3]))
        g.scenario_init(sid)
        masks = g.find_cycles()
        assert masks.shape[0] == n_cycles
        q = g.clone()
        before = q.arrays()
        for _ in range(5):
            q.step_quick(O.Chooser(O.CHOOSE_Y))
        assert np.array_equal(before["dy"], q.field("dy")) and np.array_equal(before["dx"], q.field("dx"))
    g = O.Grid(30, 30)
    g.create(O.Rng.from_seed([3]))
    g.scenario_init(6)
    masks = g.find_cycles()
    start = g.arrays()
    for t in range(8):
        g.step_perfect(O.Chooser(O.CHOOSE_Y), masks)
        if t < 7:
            assert not np.array_equal(start["color"] * ((start["dy"] != 0) | (start["dx"] != 0)),
                                      g.field("color") * ((g.field("dy") != 0) | (g.field("dx") != 0)))
    end = g.arrays()                               # period 8: colours and headings back in place
    for k in ("dy", "dx", "color"):
        assert np.array_equal(start[k], end[k]), k


def test_hits_an_edge_is_stuck_forever():
    g = O.Grid(30, 30)
    g.create(O.Rng.from_seed([3]))
    g.scenario_init(9)
    for _ in range(12):
        g.step_quick(O.Chooser(O.CHOOSE_Y))
    assert g.field("dx")[10, 15] == 1 and int((g.field("dx") != 0).sum()) == 1


def test_sparse_rings_equal_dense_masks():
    from util import random_grid, oracle_grid, state_of, assert_state_equal
    for sid in (6, 7, 8, 12):
        g = O.Grid(60, 60)
        g.create(O.Rng.from_seed([5]))
        g.scenario_init(sid)
        masks = g.find_cycles()
        rings = g.rings_from_masks(masks)
        a, b = g.clone(), g.clone()
        ca, cb = O.Chooser(O.CHOOSE_RANDOM), O.Chooser(O.CHOOSE_RANDOM)
        for t in range(30):
            na, _ = a.step_perfect(ca, masks)
            nb, _ = b.step_perfect_sparse(cb, rings)
            assert na == nb
            assert_state_equal(state_of(a), state_of(b), f"sid {sid} tick {t}")


def test_synthetic_rings_match_find_cycles():
    """The analytic ring list of the synthetic generator equals stepper_perfect.find_cycles on that grid."""
    import synth_ref
    arrs, rings = synth_ref.generate(50, 60, seed=4)
    from util import oracle_grid
    g = oracle_grid(arrs)
    masks = g.find_cycles()
    got = g.rings_from_masks(masks)

    def as_set(r):
        off, idx, cy, cx, grant = r
        return {frozenset(zip(idx[off[k]:off[k + 1]].tolist(), cy[off[k]:off[k + 1]].tolist(),
                              cx[off[k]:off[k + 1]].tolist(), grant[off[k]:off[k + 1]].tolist()))
                for k in range(len(off) - 1)}
    assert as_set(got) == as_set(rings) and len(rings[0]) - 1 > 0
