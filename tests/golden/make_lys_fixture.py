#!/usr/bin/env python
"""Golden traces of the two lys frontends from the reference's own source text: explorer/explorer.fut (with
explorer_scenario_helper.fut and scenarios.fut) and designer/designer.fut EXECUTED by tests/golden/futhark_mini.py --
`lys.init`, then a scripted sequence of `lys.event` (#keydown, #step td, #mouse) and `lys.resize` calls, with every scalar field
of the state, `text_content` and the grid the state shows recorded after every event.

tests/test_lys_traces.py holds the oracle-side models (tests/lys_model.py) and, on the GPU box, libflowamok_explorer.so /
libflowamok_designer.so against these traces.  This pins the state machines, the scenario hooks (which scenario injects what,
when; cycle recomputation) and the order in which the frontends consume their rng; the rng ARITHMETIC (cpprandom: rng_from_seed,
split_rng, join_rng, dist_i8, dist_f32 behind random_color) is the repository's restatement, not in the reference tree.

The reference tree does not travel to the GPU box: the traces are committed (tests/golden/lys_traces.npz, derived data only).
Run again only if the reference changes:  python tests/golden/make_lys_fixture.py   (about ten minutes: the explorer steps
108 x 192 grids under an interpreter)
"""
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path[:0] = [str(HERE), str(HERE.parent), str(HERE.parent.parent)]
import futhark_mini as F  # noqa: E402
from make_stepper_fixture import PCG_INC, from_cells  # noqa: E402
from util import O, state_digest  # noqa: E402

REF = Path("/root/reference")
K = dict(SPACE=32, K1=49, K2=50, a=97, p=112, r=114, x=120, RIGHT=0x4000004F, LEFT=0x40000050, DOWN=0x40000051, UP=0x40000052)


def stubs():
    """what the sources import from outside the reference tree: cpprandom (restated, through the oracle's C functions), matte's
    colour constants (payload), lys' key codes (SDL2)"""
    def dist_i8_rand(bounds, rng):
        r = O.Rng(rng, PCG_INC)
        v = r.dist_i8(bounds[0], bounds[1])
        return (r.state, v)

    def split(n, rng):
        return [x.state for x in O.Rng(rng, PCG_INC).split(n)]

    def join(rs):
        arr = (O._Rng * len(rs))(*[O.Rng(s, PCG_INC).c for s in rs])
        return int(O.lib().fo_join_rng(arr, len(rs)).state)

    def random_color(rng):
        r = O.Rng(rng, PCG_INC)
        col = int(O.lib().fo_random_color(C.byref(r.c)))
        return (r.state, col)

    def from_seed(xs):
        return O.Rng.from_seed([int(np.int32(x)) for x in xs]).state
    rnge = F.Module({"rng_from_seed": F.builtin("rng_from_seed", 1, from_seed), "split_rng": F.builtin("split_rng", 2, split),
                     "join_rng": F.builtin("join_rng", 1, join)})
    random_stub = {"rnge": rnge, "dist_i8": F.Module({"rand": F.builtin("rand", 2, dist_i8_rand)}), "dist_f32": F.Module({}),
                   "random_color": F.builtin("random_color", 1, random_color)}
    from flowamok_b200 import scenarios as S
    argb = F.Module({"white": S.WHITE, "black": S.BLACK, "red": S.RED, "green": S.GREEN, "blue": S.BLUE, "yellow": S.YELLOW,
                     "orange": S.ORANGE, "magenta": S.MAGENTA, "violet": S.VIOLET, "brown": S.BROWN,
                     "gray": F.builtin("gray", 1, lambda v: ("gray", v)),
                     "mix": F.builtin("mix", 4, lambda a, c1, b, c2: ("mix", a, c1, b, c2))})
    keys = {"SDLK_SPACE": K["SPACE"], "SDLK_1": K["K1"], "SDLK_2": K["K2"], "SDLK_a": K["a"], "SDLK_p": K["p"], "SDLK_r": K["r"],
            "SDLK_RIGHT": K["RIGHT"], "SDLK_LEFT": K["LEFT"], "SDLK_DOWN": K["DOWN"], "SDLK_UP": K["UP"]}
    return {"matte/colour": {"argb": argb}, "lys/lys": {"argb": argb, **keys}, "random": random_stub}


def apply_event(lys, s, ev):
    kind = ev[0]
    if kind == "key":
        return F.call(lys["event"], ("#keydown", {"key": ev[1]}), s)
    if kind == "keyup":
        return F.call(lys["event"], ("#keyup", {"key": ev[1]}), s)
    if kind == "step":
        return F.call(lys["event"], ("#step", np.float32(ev[1])), s)
    if kind == "mouse":
        return F.call(lys["event"], ("#mouse", {"buttons": ev[1], "x": ev[2], "y": ev[3]}), s)
    if kind == "wheel":
        return F.call(lys["event"], ("#wheel", {"dx": ev[1], "dy": ev[2]}), s)
    if kind == "resize":
        return F.call(lys["resize"], ev[1], ev[2], s)
    raise ValueError(kind)


EVENT_KINDS = ["key", "keyup", "step", "mouse", "wheel", "resize"]


def encode_events(events):
    """[kind, a, b, c] as float64 (the frame times are f32 values)"""
    out = np.zeros((len(events), 4))
    for i, ev in enumerate(events):
        out[i, 0] = EVENT_KINDS.index(ev[0])
        out[i, 1:1 + len(ev) - 1] = [float(np.float32(v)) if ev[0] == "step" else v for v in ev[1:]]
    return out


def record(data, key, i, scalars, grid_cells):
    st = from_cells(grid_cells)
    data[f"{key}/{i}/scalars"] = np.array(scalars, np.int64)
    for k in ("dy", "dx", "color", "pref"):
        data[f"{key}/{i}/{k}"] = st[k]
    data[f"{key}/{i}/roads"] = np.stack([st["uy"], st["ux"]])
    data[f"{key}/{i}/digest"] = np.array(state_digest(st), np.uint64)


def f32_bits(v):
    return int(np.float32(v).view(np.uint32))


def explorer_session(data):
    lys = F.Program(stubs()).load(REF / "explorer" / "explorer.fut")["lys"].names
    seed, h, w = 2024, 540, 960
    ev = []
    def visit(sid_steps):      # RIGHT until the scenario, then: one step by SPACE, auto on, frames, auto off
        n_right, frames = sid_steps
        ev.extend([("key", K["RIGHT"])] * n_right)
        ev.append(("key", K["SPACE"]))
        ev.append(("key", K["a"]))
        ev.extend(("step", td) for td in frames)
        ev.append(("key", K["a"]))
    visit((1, (0.016, 0.05, 0.02)))                                   # scenario 1, tight queues: a car every step
    ev += [("key", K["UP"]), ("key", K["UP"]), ("key", K["DOWN"])]
    visit((3, (0.04, 0.07)))                                          # 4, close crossing: random colours, fork draws
    visit((1, (0.1,)))                                                # 5, crossroads: dist_i8 draws with four ranges
    visit((3, (0.034, 0.034)))                                        # 8, overlapping tight cycles, perfect
    ev += [("key", K["p"]), ("key", K["SPACE"]), ("key", K["p"])]     # quick: the full rings stand still; perfect again
    visit((2, (0.07, 0.1)))                                           # 10, adding lines: roads appear, cycles recomputed
    ev += [("key", K["p"]), ("key", K["SPACE"]), ("key", K["SPACE"]), ("key", K["p"])]   # roads added in quick mode, then `p` recomputes
    visit((1, (0.2,)))                                                # 11, multi spill
    ev += [("keyup", K["a"]), ("mouse", 1, 10, 10), ("wheel", 0, 1), ("key", K["x"]), ("step", 1.0)]   # all ignored
    ev += [("key", K["K1"]), ("key", K["K2"]), ("key", K["K2"]), ("resize", 333, 777)]
    ev += [("key", K["r"]), ("key", K["SPACE"]), ("key", K["LEFT"]), ("key", K["SPACE"])]
    ev += [("key", K["LEFT"])] * 11                                   # over the wrap: 10 -> 13
    ev += [("key", K["SPACE"])]
    data["explorer/meta"] = np.array([seed, h, w], np.int64)
    data["explorer/events"] = encode_events(ev)

    def scalars(s):
        sid = s["scenario_id"]
        return [s["h"], s["w"], s["scale"], int(s["auto"]), s["steps_auto_per_second"], sid, s["steps"][sid],
                int(s["perfect_simulation"]), len(s["cycle_checks"]), s["n_cell_leaks"],
                s["rng"] - (1 << 64) if s["rng"] >= 1 << 63 else s["rng"], f32_bits(s["time_unused"])]
    t0 = time.time()
    s = F.call(lys["init"], seed, h, w)
    for k in range(14):                                               # init: all fourteen grids
        st = from_cells(s["grids"][k])
        data[f"explorer/init/{k}/digest"] = np.array(state_digest(st), np.uint64)
        data[f"explorer/init/{k}/roads"] = np.stack([st["uy"], st["ux"]])
        data[f"explorer/init/{k}/cars"] = np.stack([st["dy"], st["dx"]])
    record(data, "explorer", 0, scalars(s), s["grids"][s["scenario_id"]])
    data["explorer/0/text"] = np.array(F.call(lys["text_content"], np.float32(59.94), s), np.int64)
    print(f"explorer init {time.time() - t0:.1f} s", flush=True)
    for i, e in enumerate(ev, 1):
        t0 = time.time()
        s = apply_event(lys, s, e)
        record(data, "explorer", i, scalars(s), s["grids"][s["scenario_id"]])
        data[f"explorer/{i}/text"] = np.array(F.call(lys["text_content"], np.float32(59.94), s), np.int64)
        print(f"explorer event {i:3d} {str(e):28s} scenario {s['scenario_id']:2d} steps {s['steps'][s['scenario_id']]:3d} "
              f"cycles {len(s['cycle_checks'])} leaks {s['n_cell_leaks']}  {time.time() - t0:5.1f} s", flush=True)
    data["explorer/text_format"] = np.frombuffer(F.call(lys["text_format"], ()).encode(), np.uint8)


def designer_session(data):
    lys = F.Program(stubs()).load(REF / "designer" / "designer.fut")["lys"].names
    seed, h, w = 99, 400, 500
    ev = [("step", td) for td in (0.016, 0.05, 0.4, 0.033)]
    ev += [("mouse", 1, 105, 255), ("mouse", 1, 305, 355), ("mouse", 0, 0, 0),          # an L from the entry road down to row 35
           ("mouse", 1, 205, 55), ("mouse", 1, 205, 385), ("mouse", 1, 45, 385),        # column 20 crossing both roads, west on row 38
           ("mouse", 1, 9999, 9999), ("mouse", 1, 45, 355), ("mouse", 2, 45, 355)]
    ev += [("step", td) for td in (0.5, 0.5, 0.7, 0.2, 1.0)]
    ev += [("key", K["SPACE"]), ("step", 0.5), ("key", K["SPACE"]), ("step", 2.0)]
    ev += [("key", K["r"]), ("step", 0.6), ("wheel", 1, 1), ("keyup", K["r"]), ("resize", 400, 500), ("step", 0.3)]
    data["designer/meta"] = np.array([seed, h, w], np.int64)
    data["designer/events"] = encode_events(ev)

    def scalars(s):
        cur = s["cursor_prev"]
        some = cur[0] == "#some"
        return [s["h"], s["w"], s["gh"], s["gw"], s["scale"], int(s["running"]), s["steps_per_second"], s["n_steps_total"], s["n_leaks"],
                int(some), cur[1]["y"] if some else 0, cur[1]["x"] if some else 0,
                s["rng"] - (1 << 64) if s["rng"] >= 1 << 63 else s["rng"], f32_bits(s["time_unused"])]
    s = F.call(lys["init"], seed, h, w)
    record(data, "designer", 0, scalars(s), s["grid"])
    data["designer/0/text"] = np.array(F.call(lys["text_content"], np.float32(59.94), s), np.int64)
    for i, e in enumerate(ev, 1):
        t0 = time.time()
        s = apply_event(lys, s, e)
        record(data, "designer", i, scalars(s), s["grid"])
        data[f"designer/{i}/text"] = np.array(F.call(lys["text_content"], np.float32(59.94), s), np.int64)
        print(f"designer event {i:3d} {str(e):28s} steps {s['n_steps_total']:4d} leaks {s['n_leaks']}  {time.time() - t0:5.1f} s", flush=True)
    data["designer/text_format"] = np.frombuffer(F.call(lys["text_format"], ()).encode(), np.uint8)


def readme_session(data):
    """README.fut: `init 30 30 10 123` and `step` (animation.fut:14-32), what `futhark literate` runs 400 times for the GIF"""
    ref = F.Program(stubs()).load(REF / "README.fut")
    n_frames = 40
    s = F.call(ref["init"], 30, 30, 10, 123)
    data["readme/meta"] = np.array([30, 30, 10, 123, n_frames, len(s["cycles"])], np.int64)
    for k in range(n_frames):
        t0 = time.time()
        px, s = F.call(ref["step"], s)                      # the frame shows the grid BEFORE the tick (animation.fut:22-24)
        h, w = len(px), len(px[0])
        kind, base = np.zeros((h, w), np.uint8), np.zeros((h, w), np.uint32)
        for y in range(h):
            for x in range(w):
                v = px[y][x]
                if isinstance(v, int):
                    base[y, x] = v
                elif v[0] == "gray":
                    kind[y, x] = 1
                else:
                    kind[y, x], base[y, x] = 2, v[4]
        data[f"readme/{k}/kind"], data[f"readme/{k}/base"] = kind, base
        st = from_cells(s["grid"])
        for f in ("dy", "dx", "color", "pref"):
            data[f"readme/{k}/{f}"] = st[f]
        data[f"readme/{k}/digest"] = np.array(state_digest(st), np.uint64)
        print(f"readme frame {k:3d}  {time.time() - t0:5.1f} s", flush=True)
    data["readme/rng"] = np.array([s["rng"] - (1 << 64) if s["rng"] >= 1 << 63 else s["rng"], s["steps"]], np.int64)


def hooks_session(data):
    """scenario.init and 120 consecutive scenario.step calls (steps = 0 .. 119) of every scenario of explorer/scenarios.fut, through
    explorer.fut's own `scenario` module (explorer.fut:4-44), without ticks in between: which hook injects what, when, what it
    draws from the rng, and when it asks for new cycle_checks"""
    from make_stepper_fixture import to_cells
    scenario = F.Program(stubs()).load(REF / "explorer" / "explorer.fut")["scenario"].names
    gh, gw, n_steps = 108, 192, 120
    z = np.zeros((gh, gw), np.int8)
    empty = dict(uy=z, ux=z, dy=z, dx=z, color=np.full((gh, gw), 0xFFFFFFFF, np.uint32), pref=np.zeros((gh, gw), np.uint8),
                 rng_state=np.zeros((gh, gw), np.uint64))
    data["hooks/meta"] = np.array([gh, gw, n_steps, scenario["n_scenarios"]], np.int64)
    for sid in range(scenario["n_scenarios"]):
        t0 = time.time()
        grid = F.call(scenario["init"], sid, to_cells(empty))
        rng = O.Rng.from_seed([1000 + sid]).state
        recs, rngs = [], []
        for steps in range(n_steps):
            rng, rec, grid = F.call(scenario["step"], sid, grid, steps, rng)
            recs.append(int(rec))
            rngs.append(rng - (1 << 64) if rng >= 1 << 63 else rng)
            if steps in (0, 9, 29, 59, n_steps - 1):
                st = from_cells(grid)
                for f in ("uy", "ux", "dy", "dx", "color"):
                    data[f"hooks/{sid}/{steps}/{f}"] = st[f]
        data[f"hooks/{sid}/recompute"] = np.array(recs, np.uint8)
        data[f"hooks/{sid}/rng"] = np.array(rngs, np.int64)
        data[f"hooks/{sid}/name"] = np.frombuffer(F.call(scenario["name"], sid).encode(), np.uint8)
        print(f"hooks scenario {sid:2d} {F.call(scenario['name'], sid):28s} recompute at {[i for i, r in enumerate(recs) if r]}  "
              f"{time.time() - t0:5.1f} s", flush=True)


def main():
    data = {}
    t0 = time.time()
    if len(sys.argv) > 1:                                    # python make_lys_fixture.py readme: only that session, the rest is kept
        data = dict(np.load(HERE / "lys_traces.npz"))
        {"readme": readme_session, "designer": designer_session, "explorer": explorer_session, "hooks": hooks_session}[sys.argv[1]](data)
    else:
        readme_session(data)
        hooks_session(data)
        designer_session(data)
        explorer_session(data)
    out = HERE / "lys_traces.npz"
    np.savez_compressed(out, **data)
    print(f"wrote {out} ({out.stat().st_size} bytes) in {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
