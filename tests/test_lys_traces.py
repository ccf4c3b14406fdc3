"""Golden traces of the lys frontends made by EXECUTING explorer/explorer.fut and designer/designer.fut themselves under
tests/golden/futhark_mini.py (generator: tests/golden/make_lys_fixture.py): every scalar field of the state, text_content and
the grid the state shows, after lys.init and after every event of a scripted session.

CPU suite: the oracle-side models (tests/lys_model.py, which the event-by-event GPU tests of test_lys_frontends.py use as
their checker) reproduce the traces.  GPU suite: libflowamok_explorer.so / libflowamok_designer.so do, through the entry points
of include/flowamok_lys.h."""
import numpy as np
import pytest

import lys_model as M
from util import ROOT, STATE_FIELDS, state_digest, state_of

T = np.load(ROOT / "tests" / "golden" / "lys_traces.npz")
KINDS = ["key", "keyup", "step", "mouse", "wheel", "resize"]
EXPLORER_SCALARS = ["h", "w", "scale", "auto", "steps_auto_per_second", "scenario_id", "steps", "perfect_simulation", "n_cycles",
                    "n_cell_leaks", "rng_state", "time_unused"]
DESIGNER_SCALARS = ["h", "w", "gh", "gw", "scale", "running", "steps_per_second", "n_steps_total", "n_leaks", "cursor_some", "cursor_y",
                    "cursor_x", "rng_state", "time_unused"]


def events(prog):
    out = []
    for row in T[f"{prog}/events"]:
        kind = KINDS[int(row[0])]
        if kind == "step":
            out.append((kind, np.float32(row[1])))
        elif kind in ("key", "keyup"):
            out.append((kind, int(row[1])))
        elif kind == "mouse":
            out.append((kind, int(row[1]), int(row[2]), int(row[3])))
        else:
            out.append((kind, int(row[1]), int(row[2])))
    return out


def want_scalars(prog, i):
    names = EXPLORER_SCALARS if prog == "explorer" else DESIGNER_SCALARS
    d = dict(zip(names, (int(v) for v in T[f"{prog}/{i}/scalars"])))
    d["rng_state"] &= (1 << 64) - 1
    d["time_unused"] = np.uint32(d["time_unused"]).view(np.float32)
    return d


def check(prog, i, fields, state, text, what):
    want = want_scalars(prog, i)
    for k, v in want.items():
        got = fields[k]
        if k == "time_unused":
            assert np.float32(got).tobytes() == np.float32(v).tobytes(), f"{what}: time_unused {got!r}, the reference has {v!r}"
        elif k in ("cursor_y", "cursor_x") and not want["cursor_some"]:
            continue
        else:
            assert (got or 0) == v, f"{what}: {k} = {got}, the reference has {v}"
    for k in ("dy", "dx", "color", "pref"):
        assert np.array_equal(state[k], T[f"{prog}/{i}/{k}"]), f"{what}: grid field {k}"
    assert state_digest(state) == tuple(int(v) for v in T[f"{prog}/{i}/digest"]), f"{what}: digest of the grid (rng states included)"
    assert tuple(text) == tuple(int(v) for v in T[f"{prog}/{i}/text"]), f"{what}: text_content"


def model_apply(m, ev):
    if ev[0] == "key":
        m.keydown(ev[1])
    elif ev[0] == "step":
        m.event_step(ev[1])
    elif ev[0] == "mouse":
        m.mouse(*ev[1:])
    elif ev[0] == "resize":
        m.resize(*ev[1:])


@pytest.mark.parametrize("prog", ["explorer", "designer"])
def test_models_reproduce_the_reference_source_traces(prog):
    seed, h, w = (int(v) for v in T[f"{prog}/meta"])
    m = (M.ExplorerModel if prog == "explorer" else M.DesignerModel)(seed, h, w)
    if prog == "explorer":
        for k in range(14):                                    # lys.init builds all fourteen grids (explorer.fut:89-101)
            a = m.grids[k].arrays()
            assert np.array_equal(np.stack([a["uy"], a["ux"]]), T[f"explorer/init/{k}/roads"]), f"scenario {k}: roads"
            assert np.array_equal(np.stack([a["dy"], a["dx"]]), T[f"explorer/init/{k}/cars"]), f"scenario {k}: cars"
            assert state_digest(state_of(m.grids[k])) == tuple(int(v) for v in T[f"explorer/init/{k}/digest"])
        assert M.explorer_text_format() == bytes(T["explorer/text_format"]).decode()
    check(prog, 0, m.fields(), state_of(m.grid), m.text_content(59.94), "init")
    for i, ev in enumerate(events(prog), 1):
        model_apply(m, ev)
        check(prog, i, m.fields(), state_of(m.grid), m.text_content(59.94), f"event {i} {ev}")
        a = m.grid.arrays()
        assert np.array_equal(np.stack([a["uy"], a["ux"]]), T[f"{prog}/{i}/roads"]), f"event {i} {ev}: roads"


@pytest.mark.gpu
@pytest.mark.parametrize("prog", ["explorer", "designer"])
def test_libraries_reproduce_the_reference_source_traces(prog):
    import flowamok_b200.lys as lys
    seed, h, w = (int(v) for v in T[f"{prog}/meta"])
    p = lys.Program(prog, device=0)
    octx = p.operator_context()
    assert p.text_format() == bytes(T[f"{prog}/text_format"]).decode()
    s = p.init(seed, h, w)

    def look(i, what):
        g = p.grid(s, octx)
        d = g.download()
        check(prog, i, p.fields(s), {k: d[k] for k in STATE_FIELDS}, p.text_content(s, 59.94), what)
        assert np.array_equal(np.stack([d["uy"], d["ux"]]), T[f"{prog}/{i}/roads"]), f"{what}: roads"
    look(0, "init")
    if prog == "explorer":                                     # the other thirteen grids of lys.init: walk through them and back
        for k in range(1, 14):
            s = p.key(s, lys.SDLK_RIGHT)
            d = p.grid(s, octx).download()
            assert np.array_equal(np.stack([d["uy"], d["ux"]]), T[f"explorer/init/{k}/roads"]), f"scenario {k}: roads"
            assert state_digest({f: d[f] for f in STATE_FIELDS}) == tuple(int(v) for v in T[f"explorer/init/{k}/digest"])
        s = p.key(s, lys.SDLK_RIGHT)
        assert p.fields(s)["scenario_id"] == 0
    for i, ev in enumerate(events(prog), 1):
        if ev[0] == "key":
            s = p.key(s, ev[1])
        elif ev[0] == "keyup":
            s = p.key(s, ev[1], up=True)
        elif ev[0] == "step":
            s = p.step(s, float(ev[1]))
        elif ev[0] == "mouse":
            s = p.mouse(s, *ev[1:])
        elif ev[0] == "wheel":
            s = p.wheel(s, *ev[1:])
        else:
            s = p.resize(s, *ev[1:])
        look(i, f"event {i} {ev}")
    p.free(s)
    p.close()


# ---- README.fut: init 30 30 10 123 and step (animation.fut:14-32), the program of the README GIF ----
def check_frame(k, px):
    kind, base = T[f"readme/{k}/kind"], T[f"readme/{k}/base"]
    plain = kind == 0
    assert np.array_equal(px[plain], base[plain]), f"frame {k}: plain pixels"
    assert len(np.unique(px[kind == 1])) <= 1, f"frame {k}: grey"
    for b in np.unique(base[kind == 2]):
        assert len(np.unique(px[(kind == 2) & (base == b)])) == 1, f"frame {k}: mix of {b:#x}"


def test_oracle_anim_reproduces_the_readme_program():
    from util import O
    gh, gw, scale, seed, n_frames, n_cycles = (int(v) for v in T["readme/meta"])
    oa = O.Anim(8, gh, gw, scale, seed)
    assert len(oa.cycles) == n_cycles
    for k in range(n_frames):
        check_frame(k, oa.step(render=True))
        st = state_of(oa.grid)
        for f in ("dy", "dx", "color", "pref"):
            assert np.array_equal(st[f], T[f"readme/{k}/{f}"]), f"after step {k}: {f}"
        assert state_digest(st) == tuple(int(v) for v in T[f"readme/{k}/digest"]), f"after step {k}: digest (rng states included)"
    assert (np.int64(np.uint64(oa.rng.state)), oa.steps) == tuple(T["readme/rng"])


@pytest.mark.gpu
def test_generated_library_surface_reproduces_the_readme_program():
    """futhark_entry_init / futhark_entry_step of libflowamok_b200.so (include/flowamok_futhark.h): every frame"""
    import ctypes as C
    import flowamok_b200.capi as capi
    L = capi.lib()
    gh, gw, scale, seed, n_frames, _n = (int(v) for v in T["readme/meta"])
    cfg = L.futhark_context_config_new()
    fctx = L.futhark_context_new(cfg)
    st = C.c_void_p()
    assert L.futhark_entry_init(fctx, C.byref(st), gh, gw, scale, seed) == 0
    for k in range(n_frames):
        px, st2 = C.c_void_p(), C.c_void_p()
        assert L.futhark_entry_step(fctx, C.byref(px), C.byref(st2), st) == 0
        frame = np.zeros((gh * scale, gw * scale), np.uint32)
        assert L.futhark_values_u32_2d(fctx, px, frame.ctypes.data) == 0
        assert L.futhark_context_sync(fctx) == 0
        check_frame(k, frame)
        L.futhark_free_u32_2d(fctx, px)
        L.futhark_free_opaque_state(fctx, st)
        st = st2
    L.futhark_free_opaque_state(fctx, st)
    L.futhark_context_free(fctx)
    L.futhark_context_config_free(cfg)


# ---- the scenario hooks of explorer/scenarios.fut through explorer.fut's `scenario` module: init + 120 step calls each ----
HOOK_SNAPSHOTS = (0, 9, 29, 59, 119)


@pytest.mark.parametrize("sid", range(14))
def test_oracle_scenario_hooks_equal_the_reference_source(sid):
    from util import O
    gh, gw, n_steps, n_scen = (int(v) for v in T["hooks/meta"])
    assert n_scen == 14 and O.scenario_name(sid) == bytes(T[f"hooks/{sid}/name"]).decode()
    og = O.Grid(gh, gw)
    og.create(O.Rng.from_seed([1]))
    og.scenario_init(sid)
    rng = O.Rng.from_seed([1000 + sid])
    for steps in range(n_steps):
        rec = og.scenario_step(sid, steps, rng)
        assert int(rec) == int(T[f"hooks/{sid}/recompute"][steps]), f"scenario {sid} step {steps}: recompute_cycles"
        assert np.int64(np.uint64(rng.state)) == T[f"hooks/{sid}/rng"][steps], f"scenario {sid} step {steps}: rng"
        if steps in HOOK_SNAPSHOTS:
            a = og.arrays()
            for f in ("uy", "ux", "dy", "dx", "color"):
                assert np.array_equal(a[f], T[f"hooks/{sid}/{steps}/{f}"]), f"scenario {sid} step {steps}: {f}"


@pytest.mark.gpu
@pytest.mark.parametrize("sid", range(14))
def test_product_scenario_hooks_equal_the_reference_source(ctx, sid):
    """flowamok_b200.scenarios (the host-side hooks over the point-edit entries of the C ABI) on a device grid"""
    from flowamok_b200 import scenarios as S
    gh, gw, n_steps, _n = (int(v) for v in T["hooks/meta"])
    g = ctx.create_grid(gh, gw, 1)
    S.init(sid, g)
    rng = S.Pcg32.from_seed([1000 + sid])
    for steps in range(n_steps):
        rec = S.step(sid, g, steps, rng)
        assert int(bool(rec)) == int(T[f"hooks/{sid}/recompute"][steps]), f"scenario {sid} step {steps}: recompute_cycles"
        assert np.int64(np.uint64(rng.state)) == T[f"hooks/{sid}/rng"][steps], f"scenario {sid} step {steps}: rng"
        if steps in HOOK_SNAPSHOTS:
            d = g.download()
            for f in ("uy", "ux", "dy", "dx", "color"):
                assert np.array_equal(d[f], T[f"hooks/{sid}/{steps}/{f}"]), f"scenario {sid} step {steps}: {f}"
    g.free()
