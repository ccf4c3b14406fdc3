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
    assert len(NAMES) >= 50
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


# ---- render (flowamok.fut:16-64) and the point edits (utils.fut:19-38), executed from the reference source the same way ----
def check_render(name, px):
    """px: a rendered window.  The vector says, per pixel, whether the reference takes a plain colour (and which), argb.gray 0.3,
    or argb.mix 0.5 grey 0.5 colour (athas/matte is not in the reference tree: grey and mix must be FUNCTIONS of their
    arguments -- one grey everywhere, one result per mixed colour -- and geometry must agree exactly)."""
    kind, base = V[f"{name}/kind"], V[f"{name}/base"]
    assert px.shape == kind.shape
    plain = kind == 0
    assert np.array_equal(px[plain], base[plain]), f"{name}: plain pixels"
    greys = np.unique(px[kind == 1])
    assert len(greys) <= 1, f"{name}: grey pixels differ among themselves"
    for b in np.unique(base[kind == 2]):
        assert len(np.unique(px[(kind == 2) & (base == b)])) == 1, f"{name}: mix of colour {b:#x} is not a function of it"
    if len(greys):
        assert greys[0] not in (0xFFFFFFFF, 0xFF000000)


def render_case(i):
    name = f"render/{i}"
    gh, gw, scale, h, w = (int(v) for v in V[f"{name}/meta"])
    arrs = {k: V[f"{name}/{k}"] for k in ("uy", "ux", "dy", "dx", "color")}
    arrs["pref"] = np.zeros((gh, gw), np.uint8)
    arrs["rng_state"] = np.zeros((gh, gw), np.uint64)
    arrs["rng_inc"] = np.full((gh, gw), PCG_INC, np.uint64)
    return name, gh, gw, scale, h, w, arrs


@pytest.mark.parametrize("i", range(int(V["render/n"])))
def test_oracle_render_equals_the_reference_source(i):
    name, gh, gw, scale, h, w, arrs = render_case(i)
    og = oracle_grid(arrs)
    px = np.zeros((h, w), np.uint32)
    O.lib().fo_render(og.p, h, w, scale, px.ctypes.data)
    check_render(name, px)


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(int(V["render/n"])))
def test_cuda_render_equals_the_reference_source(ctx, i):
    name, gh, gw, scale, h, w, arrs = render_case(i)
    g = ctx.grid_from_arrays(**arrs)
    img = g.render(scale)                                  # [gh*scale][gw*scale]; the window cuts or pads it with black
    px = np.full((h, w), 0xFF000000, np.uint32)
    ch, cw = min(h, img.shape[0]), min(w, img.shape[1])
    px[:ch, :cw] = img[:ch, :cw]
    check_render(name, px)
    g.free()


def apply_edits(obj, edits):
    for e in edits.tolist():
        if e[0] == 0:
            obj.add_line_vertical(e[1], e[2], e[3], e[4])
        elif e[0] == 1:
            obj.add_line_horizontal(e[1], e[2], e[3], e[4])
        else:
            obj.add_cell(e[1], e[2], e[3], e[4], e[5])


@pytest.mark.parametrize("i", range(int(V["edits/n"])))
def test_oracle_edits_equal_the_reference_source(i):
    gh, gw = (int(v) for v in V[f"edits/{i}/meta"])
    og = O.Grid(gh, gw)
    og.create(O.Rng.from_seed([1]))
    apply_edits(og, V[f"edits/{i}/edits"])
    a = og.arrays()
    for k in ("uy", "ux", "dy", "dx", "color"):
        assert np.array_equal(a[k], V[f"edits/{i}/{k}"]), k


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(int(V["edits/n"])))
def test_cuda_edits_equal_the_reference_source(ctx, i):
    gh, gw = (int(v) for v in V[f"edits/{i}/meta"])
    g = ctx.create_grid(gh, gw, 1)
    apply_edits(g, V[f"edits/{i}/edits"])
    d = g.download()
    for k in ("uy", "ux", "dy", "dx", "color"):
        assert np.array_equal(d[k], V[f"edits/{i}/{k}"]), k
    g.free()


# ---- the kernel programs themselves (fa_stream.h / fa_tiles.h compiled for the host, warp lanes as coroutines) ----
@pytest.mark.parametrize("variant", [0, 1, 2])
def test_emulated_kernel_programs_equal_the_reference_source(variant):
    """no GPU needed: the code the CUDA kernels run (strip / tile programs, speculative move + fix-up, sparse ring resolution),
    in three shapes of strips, tiles and Jacobi levels, against the vectors -- state and leak COUNT (the emulation does not
    build the leak list)"""
    from util import EmulGrid
    for name in NAMES:
        gh, gw, perfect, mode, ticks, arrs, tape = case(name)
        eg = EmulGrid(arrs)
        rings = None
        if perfect:
            rings = oracle_grid(arrs).rings_from_masks(V[f"{name}/cycles"])      # dense masks -> the sparse list the kernels take
        for t, want, leaks in expected_states(name, arrs, min(ticks, 40)):
            n = eg.step(variant=variant, chooser=mode, tape_bit=int(tape[(t - 1) % len(tape)]), rings=rings)
            assert n == len(leaks), f"{name} variant {variant} tick {t}: leak count"
            assert_state_equal(want, eg.state(), f"{name} variant {variant} tick {t}")
