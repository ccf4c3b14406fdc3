"""Golden vectors made by EXECUTING the reference's own source text (src/steppers.fut & co. under tests/golden/futhark_mini.py,
generator tests/golden/make_stepper_fixture.py): every field of every cell after every tick, the ordered cell_leak list,
find_cycles -- quick and perfect mode, 40+ small networks (queues, saturated merges, forks, leaks off the road and off the
grid, off-road cars, diagonal headings, overlapping rings).  The oracle is held against them on the CPU, the CUDA path
(through the C ABI) on the GPU box.  The fork DRAWS use the repository's restated pcg32 (cpprandom is not in the reference
tree); which cell draws, when, and where the new state is stored is the reference's."""
import numpy as np
import pytest

from util import O, PCG_INC, ROOT, STATE_FIELDS, assert_state_equal, oracle_grid, state_of

V = np.load(ROOT / "tests" / "golden" / "stepper_vectors.npz")
NAMES = [str(n) for n in V["names"]]


def case(name):
    gh, gw, perfect, mode, ticks = (int(v) for v in V[f"{name}/meta"])
    arrs = {k: V[f"{name}/t0/{k}"] for k in ("uy", "ux", "dy", "dx", "color", "pref", "rng_state")}
    arrs["rng_inc"] = np.full((gh, gw), PCG_INC, np.uint64)
    return gh, gw, bool(perfect), mode, ticks, arrs, V[f"{name}/tape"]


def expected_states(name, arrs, ticks):
    """the state after every tick (the rng plane is stored as the cells that changed)"""
    rng = arrs["rng_state"].copy()
    for t in range(1, ticks + 1):
        rng = rng.copy()
        rng.reshape(-1)[V[f"{name}/t{t}/rng_idx"]] = V[f"{name}/t{t}/rng_val"]
        st = {k: V[f"{name}/t{t}/{k}"] for k in ("dy", "dx", "color", "pref")}
        st.update(uy=arrs["uy"], ux=arrs["ux"], rng_state=rng)
        yield t, {k: st[k] for k in STATE_FIELDS}, V[f"{name}/t{t}/leaks"]


def test_vector_set_is_what_the_generator_describes():
    assert len(NAMES) >= 40
    assert sum(int(V[f"{n}/meta"][2]) for n in NAMES) >= 15                       # perfect-mode cases
    assert sum(len(V[f"{n}/t{t}/leaks"]) for n in NAMES for t in range(1, int(V[f"{n}/meta"][4]) + 1)) > 2000
    assert sum(len(V[f"{n}/t{t}/rng_idx"]) for n in NAMES for t in range(1, int(V[f"{n}/meta"][4]) + 1)) > 300   # fork draws
    assert max(len(V[f"{n}/cycles"]) for n in NAMES if f"{n}/cycles" in V) >= 7


@pytest.mark.parametrize("name", NAMES)
def test_oracle_equals_the_reference_source(name):
    gh, gw, perfect, mode, ticks, arrs, tape = case(name)
    og = oracle_grid(arrs)
    ch = O.Chooser(mode, tape if mode == O.CHOOSE_TAPE else None)
    cycles = None
    if perfect:
        cycles = og.find_cycles()
        assert np.array_equal(cycles, V[f"{name}/cycles"]), "find_cycles (steppers.fut:233-300): masks, order included"
    for t, want, leaks in expected_states(name, arrs, ticks):
        n, got = og.step_perfect(ch, cycles, leaks=True) if perfect else og.step_quick(ch, leaks=True)
        assert_state_equal(want, state_of(og), f"{name} tick {t}")
        rows = [[l["y_next"], l["x_next"], l["uy"], l["ux"], l["dy"], l["dx"], l["color"], l["calc"], l["my"], l["mx"], l["pref"],
                 int(np.uint64(l["rng_state"]).astype(np.int64))] for l in got]
        assert n == len(leaks) and rows == leaks.tolist(), f"{name} tick {t}: cell_leak list"


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_equals_the_reference_source(ctx, name):
    import flowamok_b200 as fa
    gh, gw, perfect, mode, ticks, arrs, tape = case(name)
    g = ctx.grid_from_arrays(**arrs)
    g.set_chooser(mode, tape if mode == fa.CHOOSE_TAPE else None)
    if perfect:
        assert g.find_cycles() == len(V[f"{name}/cycles"])
        assert np.array_equal(g.cycles_dense(), V[f"{name}/cycles"]), "find_cycles: masks, order included"
    for t, want, leaks in expected_states(name, arrs, ticks):
        n, rec = g.step_leaks_cells(perfect)
        d = g.download()
        assert_state_equal(want, {k: d[k] for k in STATE_FIELDS}, f"{name} tick {t}")
        assert n == len(leaks) and np.delete(rec, [2, 3], axis=1).tolist() == leaks.tolist(), f"{name} tick {t}: cell_leak list"
    g.free()
