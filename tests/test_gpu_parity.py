"""Parity tests proper: the CUDA path, called through the C ABI (flowamok_b200 -> libflowamok_b200.so),
against the CPU oracle on the same inputs -- bit-exact, full state, every tick -- plus the committed
golden fixture and size-independent properties at BASELINE.json's full sizes."""
import hashlib

import numpy as np
import pytest

import synth_ref
from util import O, ROOT, STATE_FIELDS, assert_state_equal, oracle_grid, random_grid, state_digest, state_of

pytestmark = pytest.mark.gpu


def dev_state(g):
    d = g.download()
    return {k: d[k] for k in STATE_FIELDS}


def rings_to_api(gw, rings):
    off, idx, cy, cx, grant = rings
    return off, idx // gw, idx % gw, cy, cx, grant


SHAPES = [(5, 7), (17, 31), (30, 32), (33, 33), (64, 65), (70, 100), (1, 40), (40, 1), (97, 130), (200, 1100),
          (130, 2100)]


def test_upload_download_roundtrip(ctx):
    for seed, (gh, gw) in enumerate(SHAPES):
        arrs = random_grid(seed, gh, gw)
        g = ctx.grid_from_arrays(**arrs)
        d = g.download()
        for k in ("uy", "ux", "dy", "dx", "color", "pref", "rng_state"):
            assert np.array_equal(d[k], arrs[k]), (k, gh, gw)


def test_quick_random_grids_every_tick(ctx):
    for seed, (gh, gw) in enumerate(SHAPES):
        arrs = random_grid(300 + seed, gh, gw)
        og, ch = oracle_grid(arrs), O.Chooser(O.CHOOSE_RANDOM)
        g = ctx.grid_from_arrays(**arrs)
        for t in range(12):
            nl, _ = og.step_quick(ch)
            assert g.step(False, 1) == nl, f"leaks {gh}x{gw} tick {t}"
            assert_state_equal(state_of(og), dev_state(g), f"{gh}x{gw} tick {t}")


def test_dense_jam_multi_pass(ctx):
    """long queues crossing strip borders: k_u_local leaves them to k_u_iter (many frontier iterations)"""
    arrs = random_grid(7, 300, 2200, n_lines=400, car_p=0.93, weird_p=0.0, offroad_p=0.0)
    og, ch = oracle_grid(arrs), O.Chooser(O.CHOOSE_RANDOM)
    g = ctx.grid_from_arrays(**arrs)
    g.stats()
    for t in range(6):
        nl, _ = og.step_quick(ch)
        assert g.step(False, 1) == nl
        assert_state_equal(state_of(og), dev_state(g), f"tick {t}")
    st = g.stats()
    assert st["ticks"] == 6 and st["passes"] > 6, st


def test_jam_mode_switch_is_exact():
    """A jammed network (many visits per fixpoint iteration) makes the tick switch from "small k_u_iter beside a speculative
    k_move" to "full-size k_u_iter, then k_move" a few ticks later (jam detection reads device counters without synchronising)
    and back when the threshold is raised: every tick of both schedules must equal the oracle.  The threshold is lowered so
    that a grid the oracle steps in seconds trips it."""
    import flowamok_b200 as fa
    c = fa.Context()
    gh, gw = 1024, 4096
    arrs, rings = synth_ref.generate(gh, gw, seed=5, density16=60000)
    g = c.grid_from_arrays(**arrs)
    og, ch = oracle_grid(arrs), O.Chooser(O.CHOOSE_RANDOM)
    g.stats()
    assert c.L.flowamok_set_jam_threshold(c.h, 500.0) == 0
    for t in range(16):
        if t == 10:
            assert c.L.flowamok_set_jam_threshold(c.h, 1e9) == 0     # back to the speculative order
        nl, _ = og.step_quick(ch)
        assert g.step(False, 1) == nl, f"tick {t}"
        if t in (0, 1, 5, 9, 10, 12, 15):
            assert_state_equal(state_of(og), dev_state(g), f"tick {t}")
    st = g.stats()
    assert st["sweeps"] / max(1, st["passes"]) > 500, st      # the workload really was above the lowered threshold
    import ctypes as C
    n = C.c_int64()
    assert c.L.flowamok_jam_ticks(c.h, g.h, C.byref(n)) == 0
    assert 3 <= n.value <= 10, n.value                         # both orders ran


def test_dense_free_flow_overflows_the_arrival_list(ctx):
    """every row an avenue, every other cell a car: half of all cells receive a car each tick, far more than the
    shared-memory arrival list of a k_move tile holds (the direct pass from the descriptors runs); the grid is wider
    than one k_move tile and than one strip, and taller than one strip"""
    gh, gw = 150, 4200
    r = np.random.default_rng(5)
    uy = np.zeros((gh, gw), np.int8)
    ux = np.ones((gh, gw), np.int8)
    ux[1::2] = -1                                   # alternate rows flow the other way
    dy = np.zeros((gh, gw), np.int8)
    dx = np.zeros((gh, gw), np.int8)
    cars = np.zeros((gh, gw), bool)
    cars[:, ::2] = True
    dx[cars] = ux[cars]
    arrs = dict(uy=uy, ux=ux, dy=dy, dx=dx, color=r.integers(0, 2**32, (gh, gw), dtype=np.uint64).astype(np.uint32),
                pref=np.zeros((gh, gw), np.uint8), rng_state=r.integers(0, 2**63, (gh, gw), dtype=np.uint64),
                rng_inc=np.full((gh, gw), 1, np.uint64))
    og, ch = oracle_grid(arrs), O.Chooser(O.CHOOSE_RANDOM)
    g = ctx.grid_from_arrays(**arrs)
    for t in range(5):
        nl, _ = og.step_quick(ch)
        assert g.step(False, 1) == nl, f"leaks tick {t}"
        assert_state_equal(state_of(og), dev_state(g), f"tick {t}")


@pytest.mark.parametrize("mode", [O.CHOOSE_Y, O.CHOOSE_X, O.CHOOSE_TAPE])
def test_choosers(ctx, mode):
    arrs = random_grid(11, 40, 70, car_p=0.5)
    tape = np.array([1, 0, 0, 1, 1, 0, 1], np.uint8)
    og = oracle_grid(arrs)
    ch = O.Chooser(mode, tape if mode == O.CHOOSE_TAPE else None)
    g = ctx.grid_from_arrays(**arrs)
    g.set_chooser(mode, tape if mode == O.CHOOSE_TAPE else None)
    for t in range(14):
        nl, _ = og.step_quick(ch)
        assert g.step(False, 1) == nl
        assert_state_equal(state_of(og), dev_state(g), f"tick {t}")


def test_readme_gif_replay_cuda(ctx):
    """BASELINE config 1 through the C ABI only (no oracle): create_grid 30 30 seed 123, scenario 8 via the
    point-edit API, find_cycles, 400 perfect ticks with the tape chooser; every frame vs the GIF fixture."""
    import flowamok_b200 as fa
    from flowamok_b200 import scenarios as S
    from test_oracle_golden import TAPE, classes
    gold = np.load(ROOT / "tests/golden/readme_gif_classes.npz")["frames"]
    g = ctx.create_grid(30, 30, 123)
    S.init(8, g)
    assert g.find_cycles() == 2
    g.set_chooser(fa.CHOOSE_TAPE, np.array([int(c) for c in TAPE] + [0, 0], np.uint8))
    for k in range(400):
        assert np.array_equal(classes(g.download()), gold[k]), f"frame {k}"
        fa.stepper_perfect.step(g, 1)


@pytest.mark.parametrize("sid", list(range(14)))
def test_explorer_scenarios_perfect(ctx, sid):
    """BASELINE config 2: every explorer scenario, 108x192, perfect mode, stepping exactly as
    explorer.fut:113-125 (device ticks + host scenario.step edits through the point-edit API)."""
    from flowamok_b200 import scenarios as S
    ox = O.Explorer(123, sid, perfect=True)
    dx = S.Explorer(ctx, 123, sid, perfect=True)
    assert_state_equal(state_of(ox.grid), dev_state(dx.grid), f"sid {sid} init")
    assert np.array_equal(dx.grid.cycles_dense(), ox.cycles), "find_cycles differs from the oracle"
    n_ticks = 1000                                # BASELINE config 2: 1000 ticks each
    for t in range(n_ticks):
        lo = ox.step(1)
        ld = dx.step(1)
        assert lo == ld, f"sid {sid} tick {t} leaks"
        if t < 60 or t % 25 == 0 or t == n_ticks - 1:
            assert_state_equal(state_of(ox.grid), dev_state(dx.grid), f"sid {sid} tick {t}")
    assert np.array_equal(dx.grid.cycles_dense(), ox.cycles)


@pytest.mark.parametrize("sid,perfect", [(1, True), (5, False), (10, True), (11, True), (4, False)])
def test_explorer_step_c_entry(ctx, sid, perfect):
    """explorer/explorer.fut:109-130 `step s n` as ONE C entry (flowamok_explorer_step: ticks + scenario.step + cycle
    recomputation + leak sum inside the library), in batches of 7 steps, against the oracle's Explorer."""
    import ctypes as C
    from flowamok_b200 import scenarios as S
    ox = O.Explorer(123, sid, perfect=perfect)
    dx = S.Explorer(ctx, 123, sid, perfect=perfect)          # init only; the stepping below is the C entry
    step_cur, st, inc, nl = C.c_int64(0), C.c_uint64(dx.rng.state), C.c_uint64(dx.rng.inc), C.c_int64()
    for batch in range(12):
        want = ox.step(7)
        ctx.check(ctx.L.flowamok_explorer_step(ctx.h, dx.grid.h, sid, int(perfect), 7, C.byref(step_cur), C.byref(st), C.byref(inc),
                                               C.byref(nl)))
        assert nl.value == want and step_cur.value == ox.step_cur, f"sid {sid} batch {batch}"
        assert (st.value, inc.value) == (ox.rng.state, ox.rng.inc)
        assert_state_equal(state_of(ox.grid), dev_state(dx.grid), f"sid {sid} batch {batch}")


def test_designer_step_c_entry(ctx):
    """designer/designer.fut:59-75 `step s n`: a car of random colour enters at (25, 0) every 10th step, stepper_quick.step;
    the C entry against the same loop spelled out with the oracle."""
    import ctypes as C
    gh, gw = 60, 80
    og = O.Grid(gh, gw)
    og.create(O.Rng.from_seed([9]))
    g = ctx.create_grid(gh, gw, 9)
    for obj in (og, g):
        obj.add_line_horizontal(25, 0, 60, 1)
        obj.add_line_vertical(25, 50, 60, 1)
    rng = O.Rng.from_seed([77])
    total, st, inc, nl = C.c_int64(0), C.c_uint64(rng.state), C.c_uint64(rng.inc), C.c_int32()
    ch = O.Chooser(O.CHOOSE_RANDOM)
    n_total = 0
    for batch in range(6):
        leaks = 0
        for _ in range(13):
            if n_total % 10 == 0:
                og.add_cell(25, 0, int(O.lib().fo_random_color(C.byref(rng.c))), 0, 1)
            leaks += og.step_quick(ch)[0]
            n_total += 1
        ctx.check(ctx.L.flowamok_designer_step(ctx.h, g.h, 13, C.byref(total), C.byref(st), C.byref(inc), C.byref(nl)))
        assert (nl.value, total.value) == (leaks, n_total)
        assert (st.value, inc.value) == (rng.state, rng.inc)
        assert_state_equal(state_of(og), dev_state(g), f"batch {batch}")


def test_find_cycles_matches_oracle(ctx):
    import synth_ref
    arrs, rings = synth_ref.generate(50, 60, seed=4)
    og = oracle_grid(arrs)
    g = ctx.grid_from_arrays(**arrs)
    n = g.find_cycles()
    masks = og.find_cycles()
    assert n == masks.shape[0] > 0
    assert np.array_equal(g.cycles_dense(), masks)


def test_point_edits_and_create_grid(ctx):
    og = O.Grid(50, 70)
    joined = og.create(O.Rng.from_seed([-77]))
    g = ctx.grid(50, 70)
    g.create(-77, want_joined=True)
    assert g.joined_rng == (joined.state, joined.inc)
    r = np.random.default_rng(5)
    for _ in range(60):
        kind = r.integers(0, 3)
        flow = int(r.choice([-1, 1]))
        if kind == 0:
            a, b = sorted(r.integers(0, 50, 2).tolist()); x = int(r.integers(0, 70))
            og.add_line_vertical(a, b, x, flow); g.add_line_vertical(a, b, x, flow)
        elif kind == 1:
            a, b = sorted(r.integers(0, 70, 2).tolist()); y = int(r.integers(0, 50))
            og.add_line_horizontal(y, a, b, flow); g.add_line_horizontal(y, a, b, flow)
        else:
            y, x = int(r.integers(0, 50)), int(r.integers(0, 70))
            c, dy, dx = int(r.integers(0, 2**32)), int(r.integers(-1, 2)), int(r.integers(-1, 2))
            og.add_cell(y, x, c, dy, dx); g.add_cell(y, x, c, dy, dx)
    d = g.download()
    a = og.arrays()
    for k in ("uy", "ux", "dy", "dx", "color", "pref", "rng_state"):
        assert np.array_equal(d[k], a[k]), k
    assert d["rng_inc"] == int(a["rng_inc"][0, 0])


def test_leak_list_order_and_content(ctx):
    arrs = random_grid(21, 80, 120, car_p=0.5, weird_p=0.05, offroad_p=0.03)
    og, ch = oracle_grid(arrs), O.Chooser(O.CHOOSE_RANDOM)
    g = ctx.grid_from_arrays(**arrs)
    total = 0
    for t in range(6):
        nl, leaks = og.step_quick(ch, leaks=True)
        n, rec = g.step_leaks(False)
        assert n == nl
        exp = [[l["y_next"], l["x_next"], l["y_src"], l["x_src"], l["dy"], l["dx"], l["color"], l["pref"]] for l in leaks]
        assert rec.tolist() == exp, f"tick {t}"
        total += n
        assert_state_equal(state_of(og), dev_state(g), f"tick {t}")
    assert total > 0


def test_leak_list_whole_source_cell(ctx):
    """cell_leak = (y_next, x_next, cell) (model.fut:26): the whole source cell as move_cell saw it (steppers.fut:122) --
    roads, heading, colour, calculated / direction / preference after the fixpoint, and the rng state BEFORE the tick, also
    when the leaking cell drew in the same tick."""
    total = drew = 0
    for seed in (21, 22, 23):
        arrs = random_grid(seed, 90, 130, n_lines=90, car_p=0.7, weird_p=0.08, offroad_p=0.04)
        og, ch = oracle_grid(arrs), O.Chooser(O.CHOOSE_RANDOM)
        g = ctx.grid_from_arrays(**arrs)
        for t in range(8):
            before = og.arrays()["rng_state"].copy()
            nl, leaks = og.step_quick(ch, leaks=True)
            n, rec = g.step_leaks_cells(False)
            assert n == nl
            exp = [[l["y_next"], l["x_next"], l["y_src"], l["x_src"], l["uy"], l["ux"], l["dy"], l["dx"], l["color"], l["calc"],
                    l["my"], l["mx"], l["pref"], int(np.uint64(l["rng_state"]).astype(np.int64))] for l in leaks]
            assert rec.tolist() == exp, f"seed {seed} tick {t}"
            after = og.arrays()["rng_state"]
            drew += sum(before[l["y_src"], l["x_src"]] != after[l["y_src"], l["x_src"]] for l in leaks)
            total += n
            assert_state_equal(state_of(og), dev_state(g), f"seed {seed} tick {t}")
    assert total > 20


def test_render_matches_oracle(ctx):
    for scale in (1, 4, 5, 10):
        arrs = random_grid(31, 30, 45, car_p=0.5)
        og = oracle_grid(arrs)
        g = ctx.grid_from_arrays(**arrs)
        assert np.array_equal(g.render(scale), og.render(scale)), scale


@pytest.mark.parametrize("perfect", [False, True])
def test_synthetic_network_vs_oracle(ctx, perfect):
    """BASELINE configs 3/4 generator at a size the oracle finishes in seconds; device generator vs the
    numpy restatement, then every tick vs the oracle."""
    gh, gw = 600, 1300
    arrs, rings = synth_ref.generate(gh, gw, seed=1)
    g = ctx.grid(gh, gw)
    g.generate_synthetic(seed=1)
    d = g.download()
    for k in ("uy", "ux", "dy", "dx", "color", "pref", "rng_state"):
        assert np.array_equal(d[k], arrs[k]), f"generator field {k}"
    assert g.cycle_count()[0] == len(rings[0]) - 1
    og, ch = oracle_grid(arrs), O.Chooser(O.CHOOSE_RANDOM)
    for t in range(40):
        if perfect:
            nl, _ = og.step_perfect_sparse(ch, rings)
        else:
            nl, _ = og.step_quick(ch)
        assert g.step(perfect, 1) == nl
        if t < 10 or t % 10 == 9:
            assert_state_equal(state_of(og), dev_state(g), f"perfect={perfect} tick {t}")


@pytest.mark.parametrize("perfect", [False, True])
def test_1000_ticks_vs_oracle(ctx, perfect):
    """north_star: bit-identical grids after 1000 ticks -- the synthetic network at a size the oracle steps 1000 times
    in seconds, whole state compared at ticks 100 and 1000 (leak counts on the way), batched stepping on the device"""
    gh, gw = 272, 1100
    arrs, rings = synth_ref.generate(gh, gw, seed=3)
    g = ctx.grid(gh, gw)
    g.generate_synthetic(seed=3)
    og, ch = oracle_grid(arrs), O.Chooser(O.CHOOSE_RANDOM)
    done = 0
    for stop in (100, 1000):
        leaks = 0
        for _ in range(stop - done):
            nl, _ = og.step_perfect_sparse(ch, rings) if perfect else og.step_quick(ch)
            leaks += nl
        assert g.step(perfect, stop - done) == leaks, f"leaks up to tick {stop}"
        done = stop
        assert_state_equal(state_of(og), dev_state(g), f"perfect={perfect} tick {stop}")


def _plane_digest(g):
    d = g.download(fields=("dy", "dx", "color", "pref"))
    h = hashlib.sha256()
    for k in ("dy", "dx", "color", "pref"):
        h.update(d[k].tobytes())
    occ = (d["dy"] != 0) | (d["dx"] != 0)
    return h.hexdigest(), int(occ.sum())


def test_full_size_properties_4096(ctx):
    """BASELINE config 3 (4096^2, 10 % density): determinism, car conservation (cars only disappear by
    leaving the grid: the lattice has no leaks), batch == single-tick stepping."""
    a = ctx.grid(4096, 4096)
    a.generate_synthetic(seed=1)
    b = ctx.grid(4096, 4096)
    b.generate_synthetic(seed=1)
    _, cars0 = _plane_digest(a)
    la = a.step(True, 200)
    lb = sum(b.step(True, 1) for _ in range(200))
    da, cars_a = _plane_digest(a)
    db, cars_b = _plane_digest(b)
    assert da == db and la == lb == 0
    assert cars_a <= cars0 and cars0 - cars_a < cars0 * 0.2
    st = a.stats()
    assert st["ticks"] == 200 and st["passes"] >= 200


def test_band_of_16384_vs_oracle(ctx):
    """BASELINE config 4 generator on a 512 x 16384 band: 1000 ticks vs the oracle,
    whole state at ticks 30, 100 and 1000, leak counts on the way."""
    gh, gw = 512, 16384
    arrs, rings = synth_ref.generate(gh, gw, seed=1)
    g = ctx.grid(gh, gw)
    g.generate_synthetic(seed=1)
    og, ch = oracle_grid(arrs), O.Chooser(O.CHOOSE_RANDOM)
    done = 0
    for stop in (30, 100, 1000):
        leaks = 0
        for _ in range(stop - done):
            leaks += og.step_quick(ch)[0]
        assert g.step(False, stop - done) == leaks, f"leaks up to tick {stop}"
        done = stop
        assert_state_equal(state_of(og), dev_state(g), f"band tick {stop}")


def test_futhark_entry_points_readme_program(ctx):
    """The surface `futhark cuda --library README.fut` would generate (include/flowamok_futhark.h):
    init 30 30 10 123 then 60 steps; every returned frame and the final grid vs the oracle's mk_anim."""
    import ctypes as C
    import flowamok_b200.capi as capi
    L = capi.lib()
    cfg = L.futhark_context_config_new()
    fctx = L.futhark_context_new(cfg)
    assert not L.futhark_context_get_error(fctx)
    st = C.c_void_p()
    assert L.futhark_entry_init(fctx, C.byref(st), 30, 30, 10, 123) == 0
    oa = O.Anim(8, 30, 30, 10, 123)
    for k in range(60):
        px, st2 = C.c_void_p(), C.c_void_p()
        assert L.futhark_entry_step(fctx, C.byref(px), C.byref(st2), st) == 0
        shp = L.futhark_shape_u32_2d(fctx, px)
        assert (shp[0], shp[1]) == (300, 300)
        frame = np.zeros((300, 300), np.uint32)
        assert L.futhark_values_u32_2d(fctx, px, frame.ctypes.data) == 0
        assert L.futhark_context_sync(fctx) == 0
        assert np.array_equal(frame, oa.step(render=True)), f"frame {k}"
        L.futhark_free_u32_2d(fctx, px)
        L.futhark_free_opaque_state(fctx, st)      # the input state is ours to free; the step did not consume it
        st = st2
    L.futhark_free_opaque_state(fctx, st)
    L.futhark_context_free(fctx)
    L.futhark_context_config_free(cfg)


@pytest.mark.parametrize("sid", range(14))
def test_futhark_entry_points_all_scenarios(ctx, sid):
    """The scenario table of csrc/fa_futhark.inc (third transcription of explorer/scenarios.fut) against the oracle's,
    which tests/test_scenario_tables.py pins to the mechanically derived fixture: mk_anim(scenario sid) on the
    explorer's 108 x 192 grid, every frame of the first 12 steps (init + scenario.step edits included)."""
    import ctypes as C
    import flowamok_b200.capi as capi
    L = capi.lib()
    cfg = L.futhark_context_config_new()
    L.flowamok_futhark_set_scenario(cfg, sid)
    fctx = L.futhark_context_new(cfg)
    assert not L.futhark_context_get_error(fctx)
    st = C.c_void_p()
    assert L.futhark_entry_init(fctx, C.byref(st), 108, 192, 2, 123) == 0
    oa = O.Anim(sid, 108, 192, 2, 123)
    for k in range(12):
        px, st2 = C.c_void_p(), C.c_void_p()
        assert L.futhark_entry_step(fctx, C.byref(px), C.byref(st2), st) == 0
        frame = np.zeros((216, 384), np.uint32)
        assert L.futhark_values_u32_2d(fctx, px, frame.ctypes.data) == 0
        assert L.futhark_context_sync(fctx) == 0
        assert np.array_equal(frame, oa.step(render=True)), f"sid {sid} frame {k}"
        L.futhark_free_u32_2d(fctx, px)
        L.futhark_free_opaque_state(fctx, st)
        st = st2
    L.futhark_free_opaque_state(fctx, st)
    L.futhark_context_free(fctx)
    L.futhark_context_config_free(cfg)


def test_grid_clone_is_independent(ctx):
    import ctypes as C
    arrs = random_grid(5, 60, 90, car_p=0.5)
    g = ctx.grid_from_arrays(**arrs)
    h = C.c_void_p()
    ctx.check(ctx.L.flowamok_grid_clone(ctx.h, g.h, C.byref(h)))
    g.step(False, 3)
    import flowamok_b200 as fa
    c = fa.Grid.__new__(fa.Grid)
    c.ctx, c.L, c.gh, c.gw, c.h, c.joined_rng = ctx, ctx.L, 60, 90, h, None
    d = c.download()
    for k in ("uy", "ux", "dy", "dx", "color", "pref", "rng_state"):
        assert np.array_equal(d[k], arrs[k]), k
    og, ch = oracle_grid(arrs), O.Chooser(O.CHOOSE_RANDOM)
    for t in range(3):
        og.step_quick(ch)
        c.step(False, 1)
    assert_state_equal(state_of(og), dev_state(c), "clone after 3 ticks")
    assert_state_equal(state_of(og), dev_state(g), "original after 3 ticks")


def test_host_driven_band_protocol_with_rings_across_the_cuts(ctx):
    """The phase-by-phase protocol (flowamok_band_rings / _ring_verdicts / _rings_apply, driven by LocalBands) with cuts that
    run THROUGH ring tiles: every band gets the GLOBAL cycle list after flowamok_band_set_cuts."""
    from flowamok_b200.bands import DeviceBand, LocalBands
    gh, gw = 200, 700
    arrs, rings = synth_ref.generate(gh, gw, seed=8, ring_full16=9)
    off, idx, cy, cx, grant = rings
    cuts = [0, 55, 121, 200]
    bands = []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        b = DeviceBand(ctx, gh, gw, lo, hi)
        b.set_cuts(cuts)
        b.grid.upload(**arrs)
        b.grid.set_cycles(off, idx // gw, idx % gw, cy, cx, grant)
        bands.append(b)
    assert bands[0].ring_words() > 0
    run = LocalBands(bands)
    og, ch = oracle_grid(arrs), O.Chooser(O.CHOOSE_RANDOM)
    for t in range(9):
        og.step_perfect_sparse(ch, rings)
        run.tick(True)
        if t % 4 == 0:
            want = state_of(og)
            assert_state_equal(want, _gather_bands(bands, gh, gw, want), f"tick {t}")


@pytest.mark.parametrize("n_bands,perfect", [(2, False), (3, False), (4, True)])
def test_row_bands_on_one_gpu_equal_single_grid(ctx, n_bands, perfect):
    """BASELINE config 5 shape at a size the oracle finishes in seconds: the grid split into row bands (the
    flowamok_band_* entry points, messages handed over in-process) must equal the oracle's single grid and the
    single-grid CUDA run, every few ticks."""
    from flowamok_b200.bands import DeviceBand, LocalBands, split_rows
    if perfect:
        gh, gw = 256, 1300
        arrs, rings = synth_ref.generate(gh, gw, seed=2)
    else:
        gh, gw = 150, 2100
        arrs, rings = random_grid(9, gh, gw, n_lines=300, car_p=0.8, weird_p=0.02, offroad_p=0.01), None
    bands = []
    for lo, hi in split_rows(gh, n_bands, 16 if perfect else 1):
        b = DeviceBand(ctx, gh, gw, lo, hi)
        b.grid.upload(**arrs)
        if perfect:
            off, idx, cy, cx, grant = rings
            ys = idx // gw
            keep = [k for k in range(len(off) - 1) if lo <= ys[off[k]:off[k + 1]].min() and ys[off[k]:off[k + 1]].max() < hi]
            sel = np.concatenate([np.arange(off[k], off[k + 1]) for k in keep]) if keep else np.zeros(0, np.int64)
            o2 = np.concatenate([[0], np.cumsum([off[k + 1] - off[k] for k in keep])]).astype(np.int64)
            b.grid.set_cycles(o2, ys[sel], idx[sel] % gw, cy[sel], cx[sel], grant[sel])
        bands.append(b)
    run = LocalBands(bands)
    og, ch = oracle_grid(arrs), O.Chooser(O.CHOOSE_RANDOM)
    leaks = 0
    for t in range(12):
        n, _l = og.step_perfect_sparse(ch, rings) if perfect else og.step_quick(ch)
        leaks += n
        run.tick(perfect)
        if t % 3 == 2:
            want = state_of(og)
            got = {k: np.zeros((gh, gw), want[k].dtype) for k in STATE_FIELDS}
            for b in bands:
                d = b.grid.download(fields=STATE_FIELDS)
                lo, hi = b.rows
                for k in STATE_FIELDS:
                    got[k][lo:hi] = d[k][lo:hi]
            assert_state_equal(want, got, f"{n_bands} bands tick {t}")
    assert sum(b.leaks() for b in bands) == leaks
    assert run.rounds >= 12


def _gather_bands(bands, gh, gw, like):
    got = {k: np.zeros((gh, gw), like[k].dtype) for k in STATE_FIELDS}
    for b in bands:
        d = b.grid.download(fields=STATE_FIELDS)
        lo, hi = b.rows
        for k in STATE_FIELDS:
            got[k][lo:hi] = d[k][lo:hi]
    return got


P2P_CASES = [
    # (bands' cuts, perfect, how the cycles are installed)
    ([0, 70, 150], False, None),                 # quick mode, arbitrary cuts, random network (diagonal / off-road cars)
    ([0, 41, 99, 150], False, None),
    ([0, 64, 128, 192, 256], True, "synth"),     # cuts on multiples of 16: no ring is cut
    ([0, 71, 135, 256], True, "synth"),          # cuts THROUGH the ring tiles: verdict bits AND-ed over the bands (k_band_xring)
    ([0, 130, 256], True, "list"),               # the same through flowamok_set_cycles with the global list
    ([0, 66, 256], True, "synth"),               # a cut on a ring's first row: touches, does not straddle
]


@pytest.mark.parametrize("cuts,perfect,cycles", P2P_CASES)
def test_p2p_bands_on_one_gpu_equal_oracle(cuts, perfect, cycles):
    """The multi-GPU transport itself on ONE device: every band has its own context / stream, the peers' mailboxes are
    plain device pointers, and the bands' kernels (k_band_xstate, the persistent k_u_iter with the convergence rounds
    and the termination vote inside, k_band_xring, k_band_xu) wait for each other on the device exactly as they do
    across NVLink.  Whole state vs the oracle every few ticks."""
    import flowamok_b200 as fa
    from flowamok_b200.bands import DeviceBand, LocalBandsP2P
    if perfect:
        gh, gw = 256, 1300
        arrs, rings = synth_ref.generate(gh, gw, seed=2, ring_full16=6)
    else:
        gh, gw = 150, 2100
        arrs, rings = random_grid(9, gh, gw, n_lines=300, car_p=0.8, weird_p=0.02, offroad_p=0.01), None
    ctxs, bands = [], []
    for r in range(len(cuts) - 1):
        c = fa.Context()                      # its own stream
        b = DeviceBand(c, gh, gw, cuts[r], cuts[r + 1])
        b.set_cuts(cuts)
        if cycles == "synth":
            b.grid.generate_synthetic(seed=2, ring_full16=6)
        else:
            b.grid.upload(**arrs)
            if cycles == "list":
                off, idx, cy, cx, grant = rings
                b.grid.set_cycles(off, idx // gw, idx % gw, cy, cx, grant)      # the GLOBAL list on every band
        ctxs.append(c); bands.append(b)
    run = LocalBandsP2P(bands)
    og, ch = oracle_grid(arrs), O.Chooser(O.CHOOSE_RANDOM)
    leaks = 0
    for t in range(12):
        n, _l = og.step_perfect_sparse(ch, rings) if perfect else og.step_quick(ch)
        leaks += n
        run.step(perfect, 1)
        if t % 4 == 3:
            want = state_of(og)
            assert_state_equal(want, _gather_bands(bands, gh, gw, want), f"cuts {cuts} tick {t}")
    assert sum(b.leaks() for b in bands) == leaks
    assert run.rounds >= 12
    want = state_of(og)
    ds = [b.grid.digest() for b in bands]
    s = sum(d[0] for d in ds) & (2**64 - 1)
    x = 0
    for d in ds:
        x ^= d[1]
    assert (s, x) == state_digest(want), "the bands' device digests must combine to the whole grid's"
    for b in bands:
        b.p2p_close()
    for b in bands:
        b.grid.free()


def test_sparse_rng_transfer(ctx):
    """flowamok_grid_rng_corners / _rng_scatter: only cells with a road on both axes can draw (steppers.fut:101-103), so their
    (index, state) pairs are all of underlying.rng a host has to move: a grid rebuilt from the dense fields WITHOUT the rng
    plane plus the corner list continues bit-identically."""
    arrs = random_grid(17, 120, 300, car_p=0.6)
    g = ctx.grid_from_arrays(**arrs)
    og, ch = oracle_grid(arrs), O.Chooser(O.CHOOSE_RANDOM)
    for _ in range(5):
        og.step_quick(ch)
    g.step(False, 5)
    idx, st = g.rng_corners()
    corners = (arrs["uy"] != 0) & (arrs["ux"] != 0)
    assert np.array_equal(idx, np.flatnonzero(corners.ravel()))
    d = g.download()
    assert np.array_equal(st, d["rng_state"].ravel()[idx])
    changed = d["rng_state"] != arrs["rng_state"]
    assert changed.any() and not (changed & ~corners).any()          # draws happened, and only on corner cells
    h = ctx.grid(120, 300)
    dense = {k: d[k] for k in ("uy", "ux", "dy", "dx", "color", "pref")}
    h.upload(**dense, rng_state=None)                                 # rng plane stays as it is (zeros in a fresh grid)
    h.rng_scatter(idx, st)
    for _ in range(5):
        og.step_quick(ch)
    h.step(False, 5)
    want, got = state_of(og), dev_state(h)
    for k in ("dy", "dx", "color", "pref"):
        assert np.array_equal(want[k], got[k]), k
    assert np.array_equal(want["rng_state"].ravel()[idx], h.rng_corners()[1])


def test_device_digest_matches_numpy(ctx):
    arrs = random_grid(3, 90, 200, car_p=0.5)
    g = ctx.grid_from_arrays(**arrs)
    g.step(False, 3)
    assert g.digest() == state_digest(dev_state(g))


def _oracle_synth(gh, gw, seed=1):
    og = O.Grid(gh, gw)
    rings = og.generate_synthetic(seed=seed)
    return og, rings


def test_band_whose_rows_plus_halo_rows_equal_the_grid_height(ctx):
    """regression (found by tests/fuzz_bands.py on the CPU emulation): a band of gh - 4 rows holds exactly gh rows with
    its halo and must still run as a band"""
    from flowamok_b200.bands import DeviceBand, LocalBands
    gh, gw = 54, 82
    arrs = random_grid(238, gh, gw, n_lines=44, car_p=0.8, weird_p=0.05, offroad_p=0.02)
    bands = []
    for lo, hi in ((0, gh - 4), (gh - 4, gh)):
        b = DeviceBand(ctx, gh, gw, lo, hi)
        b.grid.upload(**arrs)
        bands.append(b)
    run = LocalBands(bands)
    og, ch = oracle_grid(arrs), O.Chooser(O.CHOOSE_RANDOM)
    for t in range(6):
        og.step_quick(ch)
        run.tick(False)
    want = state_of(og)
    got = {k: np.zeros((gh, gw), want[k].dtype) for k in STATE_FIELDS}
    for b in bands:
        d = b.grid.download(fields=STATE_FIELDS)
        lo, hi = b.rows
        for k in STATE_FIELDS:
            got[k][lo:hi] = d[k][lo:hi]
    assert_state_equal(want, got, "bands after 6 ticks")


@pytest.mark.parametrize("perfect", [False, True])
def test_config3_4096_vs_oracle_1000_ticks(ctx, perfect):
    """BASELINE config 3 (4096^2, seed 1, 10 % density, 1000 ticks): quick and perfect vs the oracle, whole state at
    ticks 1, 2, 3, 10, 100 and 1000, leak counts in between."""
    og, rings = _oracle_synth(4096, 4096)
    g = ctx.grid(4096, 4096)
    g.generate_synthetic(seed=1)
    ch = O.Chooser(O.CHOOSE_RANDOM)
    done = 0
    for stop in (1, 2, 3, 10, 100, 1000):
        leaks = 0
        for _ in range(stop - done):
            leaks += (og.step_perfect_sparse(ch, rings) if perfect else og.step_quick(ch))[0]
        assert g.step(perfect, stop - done) == leaks, f"leaks up to tick {stop}"
        done = stop
        assert_state_equal(state_of(og), dev_state(g), f"4096^2 perfect={perfect} tick {stop}")
    g.free()


def test_config4_16384_vs_oracle_first_ticks(ctx):
    """BASELINE config 4 (the roofline headline, 16384^2 quick): the first 10 ticks of the full grid vs the oracle
    (BASELINE.md section 3 row 4), whole state at ticks 1, 3 and 10."""
    og, _rings = _oracle_synth(16384, 16384)
    g = ctx.grid(16384, 16384)
    g.generate_synthetic(seed=1)
    ch = O.Chooser(O.CHOOSE_RANDOM)
    for t in range(1, 11):
        nl, _ = og.step_quick(ch)
        assert g.step(False, 1) == nl
        if t in (1, 3, 10):
            assert_state_equal(state_of(og), dev_state(g), f"16384^2 tick {t}")
    del og
    g.free()
