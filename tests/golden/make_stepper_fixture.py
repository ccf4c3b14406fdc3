#!/usr/bin/env python
"""Golden vectors from the reference's own source text: /root/reference/src/steppers.fut (with helpers.fut, option.fut,
reduce1.fut, utils.fut) EXECUTED by tests/golden/futhark_mini.py -- there is no Futhark compiler in the image -- on small
grids, tick by tick, in quick and in perfect mode (find_cycles included), with every field of every cell and the ordered
cell_leak list recorded.  tests/test_stepper_vectors.py holds the oracle (and, on the GPU box, the CUDA path) against them.

What this pins that the README GIF does not: merge alternation (next_preference_flow_x), queues, leaks off the road and
off the grid, cars on cells without road, diagonal headings, the order and content of the leak list, find_cycles on
networks with several overlapping rings, and WHICH cell's rng a fork draw consumes and updates (steppers.fut:101-112).
What it cannot pin: the arithmetic of the draw itself (diku-dk/cpprandom is not in the reference tree): choose_direction
is a parameter of `step` (steppers.fut:180, :305), so the vectors are made with deterministic choosers and with the
repository's restated pcg32 (fork decisions then depend on the restatement, the plumbing around them does not).

The reference tree does not travel to the GPU box, so the vectors are committed (tests/golden/stepper_vectors.npz, derived
data only); run again only if the reference changes:  python tests/golden/make_stepper_fixture.py
"""
import sys
import time
import zlib
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path[:0] = [str(HERE), str(HERE.parent), str(HERE.parent.parent)]
import futhark_mini as F  # noqa: E402
from util import O, random_grid  # noqa: E402  (test infrastructure: the oracle's restated pcg32 for the RANDOM chooser)

REF = Path("/root/reference/src")
CHOOSE_RANDOM, CHOOSE_Y, CHOOSE_X, CHOOSE_TAPE = 0, 1, 2, 3
PCG_INC = ((0xda3e39cb94b95bdb << 1) | 1) & (2**64 - 1)


def load_reference():
    """flowamok.fut (which opens src/steppers, src/utils, ...) under the interpreter.  Not in the reference tree and therefore
    stubbed: athas/matte (argb.gray / argb.mix return symbolic values, so that `render` yields WHICH pixels are grey or mixed
    and from which colour), diku-dk/lys, diku-dk/cpprandom (dist_i8.rand is the repository's restated pcg32 draw)."""
    unavailable = F.builtin("unavailable", 1, lambda *_: (_ for _ in ()).throw(F.FutharkError("not in the reference tree")))

    def dist_i8_rand(bounds, rng):
        r = O.Rng(rng, PCG_INC)
        v = r.dist_i8(bounds[0], bounds[1])
        return (r.state, v)
    rnge = F.Module({"rng_from_seed": F.builtin("rng_from_seed", 1, lambda xs: 0), "split_rng": unavailable, "join_rng": unavailable})
    random_stub = {"rnge": rnge, "dist_i8": F.Module({"rand": F.builtin("dist_i8.rand", 2, dist_i8_rand)}), "dist_f32": F.Module({}),
                   "random_color": unavailable}
    argb = F.Module({"white": 0xFFFFFFFF, "black": 0xFF000000, "gray": F.builtin("gray", 1, lambda v: ("gray", v)),
                     "mix": F.builtin("mix", 4, lambda a, c1, b, c2: ("mix", a, c1, b, c2))})
    stubs = {"matte/colour": {"argb": argb}, "lys/lys": {}, "random": random_stub}
    return F.Program(stubs).load(REF.parent / "flowamok.fut")


def to_cells(a):
    gh, gw = a["uy"].shape
    return [[{"underlying": {"rng": int(a["rng_state"][y, x]), "direction": {"y": int(a["uy"][y, x]), "x": int(a["ux"][y, x])}},
              "direction": {"y": int(a["dy"][y, x]), "x": int(a["dx"][y, x])},
              "color": int(a["color"][y, x]),
              "can_be_moved_to_from": {"calculated": False, "direction": {"y": False, "x": False},
                                       "next_preference_flow_x": bool(a["pref"][y, x])},
              "aux": ()} for x in range(gw)] for y in range(gh)]


def from_cells(cells):
    gh, gw = len(cells), len(cells[0])
    out = {k: np.zeros((gh, gw), dt) for k, dt in (("uy", np.int8), ("ux", np.int8), ("dy", np.int8), ("dx", np.int8),
                                                   ("color", np.uint32), ("pref", np.uint8), ("rng_state", np.uint64),
                                                   ("transient", np.uint8))}
    for y in range(gh):
        for x in range(gw):
            c = cells[y][x]
            out["uy"][y, x], out["ux"][y, x] = c["underlying"]["direction"]["y"], c["underlying"]["direction"]["x"]
            out["dy"][y, x], out["dx"][y, x] = c["direction"]["y"], c["direction"]["x"]
            out["color"][y, x] = c["color"]
            m = c["can_be_moved_to_from"]
            out["pref"][y, x] = m["next_preference_flow_x"]
            out["rng_state"][y, x] = c["underlying"]["rng"]
            out["transient"][y, x] = m["calculated"] | (m["direction"]["y"] << 1) | (m["direction"]["x"] << 2)
    return out


def leak_rows(leaks):
    """cell_leak = (y_next, x_next, source cell): 12 integers per leak, in list order"""
    rows = []
    for y_next, x_next, c in leaks:
        m = c["can_be_moved_to_from"]
        rows.append([y_next, x_next, c["underlying"]["direction"]["y"], c["underlying"]["direction"]["x"], c["direction"]["y"],
                     c["direction"]["x"], c["color"], int(m["calculated"]), int(m["direction"]["y"]), int(m["direction"]["x"]),
                     int(m["next_preference_flow_x"]), c["underlying"]["rng"] - (1 << 64) if c["underlying"]["rng"] >= 1 << 63
                     else c["underlying"]["rng"]])
    return np.array(rows, np.int64).reshape(len(rows), 12)


class Chooser:
    """choose_direction y x cell -> (rng, direction flow)  (steppers.fut:60, flowamok.fut:10-14 for RANDOM)"""

    def __init__(self, mode, tape=None, reference_random=None):
        self.mode, self.tape, self.tick = mode, tape, 0
        # RANDOM: the reference's own choose_direction_random (flowamok.fut:10-14); only dist_i8.rand underneath is restated
        self.fn = reference_random if mode == CHOOSE_RANDOM else F.builtin("choose_direction", 3, self.choose)

    def choose(self, _y, _x, cell):
        u, rng = cell["underlying"]["direction"], cell["underlying"]["rng"]
        if self.mode == CHOOSE_Y:
            choice = 0
        elif self.mode == CHOOSE_X:
            choice = 1
        else:
            choice = int(self.tape[self.tick % len(self.tape)])
        return (rng, {"y": u["y"], "x": 0} if choice == 0 else {"y": 0, "x": u["x"]})


def grid_from_lines(gh, gw, lines, cars, seed):
    """roads by add_line_* (utils.fut:19-33 semantics), cars as (y, x, dy, dx); colours / preferences / rng states seeded"""
    r = np.random.default_rng(seed)
    a = random_grid(seed, gh, gw, n_lines=0, car_p=0.0, weird_p=0.0, offroad_p=0.0)
    for k in ("uy", "ux", "dy", "dx"):
        a[k][...] = 0
    for kind, p, q0, q1, flow in lines:
        if kind == "h":
            a["ux"][p, q0:q1 + 1] = flow
        else:
            a["uy"][q0:q1 + 1, p] = flow
    for y, x, dy, dx in cars:
        a["dy"][y, x], a["dx"][y, x] = dy, dx
    a["pref"][...] = (r.random((gh, gw)) < 0.5)
    return a


def ring(y0, x0, n, sense):
    """a one-way ring with corners (y0,x0) .. (y0+n, x0+n), like scenarios.fut:117-120; sense +1 = clockwise"""
    y1, x1 = y0 + n, x0 + n
    if sense > 0:
        return [("h", y0, x0, x1, 1), ("v", x1, y0, y1, 1), ("h", y1, x0, x1, -1), ("v", x0, y0, y1, -1)]
    return [("h", y0, x0, x1, -1), ("v", x0, y0, y1, 1), ("h", y1, x0, x1, 1), ("v", x1, y0, y1, -1)]


def fill_roads(a, p, seed):
    """a car heading along its road on a fraction p of the road cells"""
    r = np.random.default_rng(seed)
    road = (a["uy"] != 0) | (a["ux"] != 0)
    car = road & (r.random(road.shape) < p)
    along_y = (a["uy"] != 0) & ((a["ux"] == 0) | (r.random(road.shape) < 0.5))
    a["dy"][car & along_y] = a["uy"][car & along_y]
    a["dx"][car & ~along_y] = a["ux"][car & ~along_y]
    return a


def cp(a):
    return {k: v.copy() for k, v in a.items()}


PALETTE = np.array([0xFFFF0000, 0xFF00FF00, 0xFF0000FF, 0xFFFFFF00, 0xFFFF00FF, 0xFF7C301C, 0xFF00FFFF, 0xFF808080, 0xFFFFFFFF,
                    0xFF123456, 0xFFABCDEF, 0xFF0F0F0F, 0x80FEDCBA], np.uint32)


def palette(a):
    """colours from a small palette (the vectors compress); stale colours on empty cells stay distinguishable"""
    a["color"] = PALETTE[a["color"] % np.uint32(len(PALETTE))]
    return a


def cases():
    out = []
    # quick mode on random networks: queues, merges, forks, leaks, off-road cars, diagonal headings
    r = np.random.default_rng(2024)
    shapes = [(9, 13), (12, 16), (7, 33), (20, 9), (14, 14), (16, 20), (1, 24), (18, 1), (10, 35), (24, 31), (30, 40), (33, 33),
              (5, 64), (40, 12), (2, 2), (3, 70), (26, 26), (17, 45), (31, 32), (32, 33),
              (1, 1), (4, 32), (4, 33), (3, 64), (3, 65), (2, 31), (3, 96), (3, 97), (6, 129)]      # plane-word boundaries of the CUDA layout
    for i, (gh, gw) in enumerate(shapes):
        nl = int(r.integers(2, 3 + (gh + gw) // 2))
        car_p, wp, op = float(r.choice([0.3, 0.6, 0.9, 1.0])), float(r.choice([0.0, 0.03, 0.15])), float(r.choice([0.0, 0.01, 0.08]))
        mode = int(r.choice([CHOOSE_RANDOM, CHOOSE_RANDOM, CHOOSE_TAPE, CHOOSE_Y, CHOOSE_X]))
        out.append(dict(name=f"quick_random_{i}", arrs=random_grid(100 + i, gh, gw, n_lines=nl, car_p=car_p, weird_p=wp, offroad_p=op),
                        perfect=False, mode=mode, ticks=12))
    # a saturated merge: two full queues into one road (merge alternation, steppers.fut:41-50)
    a = grid_from_lines(12, 18, [("h", 6, 0, 17, 1), ("v", 9, 0, 6, 1)], [], 7)
    out.append(dict(name="quick_merge_saturated", arrs=fill_roads(a, 1.0, 7), perfect=False, mode=CHOOSE_RANDOM, ticks=14))
    # roads that end in nothing and at the grid edge: several leaks per tick, some into the same cell (scenarios.fut:215-233)
    a = grid_from_lines(11, 11, [("v", 5, 0, 4, 1), ("h", 5, 0, 4, 1), ("v", 5, 6, 10, -1), ("h", 5, 6, 10, -1), ("h", 0, 0, 10, 1),
                                 ("v", 0, 0, 10, -1)], [], 8)
    out.append(dict(name="quick_leaks", arrs=fill_roads(a, 0.7, 8), perfect=False, mode=CHOOSE_RANDOM, ticks=10))
    # perfect mode: full rings (frozen in quick mode, rotating here), overlapping rings with a fork and a merge, a ring with
    # a feeder road, two rings of opposite sense
    a = grid_from_lines(10, 10, ring(2, 2, 3, 1), [], 9)
    out.append(dict(name="perfect_small_tight_cycle", arrs=fill_roads(a, 1.0, 9), perfect=True, mode=CHOOSE_RANDOM, ticks=8))
    out.append(dict(name="quick_small_tight_cycle", arrs=cp(a), perfect=False, mode=CHOOSE_RANDOM, ticks=4))
    a = grid_from_lines(16, 18, ring(2, 2, 8, -1) + [("h", 6, 2, 14, 1), ("v", 14, 6, 10, 1), ("h", 10, 2, 14, -1)], [], 10)
    out.append(dict(name="perfect_overlapping_cycles", arrs=fill_roads(a, 1.0, 10), perfect=True, mode=CHOOSE_RANDOM, ticks=16))
    out.append(dict(name="perfect_overlapping_cycles_tape", arrs=fill_roads(cp(a), 1.0, 11),
                    perfect=True, mode=CHOOSE_TAPE, ticks=12))
    a = grid_from_lines(14, 20, ring(1, 1, 4, 1) + ring(7, 9, 5, -1) + [("h", 3, 5, 18, 1)], [], 12)
    out.append(dict(name="perfect_two_rings_and_a_road", arrs=fill_roads(a, 0.9, 12), perfect=True, mode=CHOOSE_RANDOM, ticks=12))
    a = grid_from_lines(13, 15, ring(3, 3, 6, 1) + [("v", 6, 0, 11, 1), ("h", 6, 0, 13, 1)], [], 13)
    out.append(dict(name="perfect_ring_crossed_by_avenues", arrs=fill_roads(a, 0.85, 13), perfect=True, mode=CHOOSE_RANDOM, ticks=12))
    # random networks in perfect mode.  NB: the reference's find_cycles walks a road from a corner without a bounds check
    # (steppers.fut:262-265, :292-295), so a one-way road that runs from a corner off the grid aborts the Futhark program with an
    # index error; such networks are not valid inputs and are skipped here (the seeds below are the first that are valid
    # and have at least one ring).
    for i, seed in enumerate(PERFECT_RANDOM_SEEDS):
        out.append(dict(name=f"perfect_random_{i}", arrs=random_grid(seed, 12, 14, n_lines=7, car_p=0.9, weird_p=0.02, offroad_p=0.01),
                        perfect=True, mode=[CHOOSE_RANDOM, CHOOSE_TAPE, CHOOSE_RANDOM, CHOOSE_Y][i % 4], ticks=12))
    # long runs: alternation patterns and ring rotation over a hundred ticks and more
    out.append(dict(name="long_quick_random", arrs=random_grid(4242, 16, 22, n_lines=11, car_p=0.8, weird_p=0.02, offroad_p=0.01),
                    perfect=False, mode=CHOOSE_RANDOM, ticks=150))
    a = grid_from_lines(16, 18, ring(2, 2, 8, -1) + [("h", 6, 2, 14, 1), ("v", 14, 6, 10, 1), ("h", 10, 2, 14, -1)], [], 21)
    out.append(dict(name="long_perfect_overlapping_cycles", arrs=fill_roads(a, 1.0, 21), perfect=True, mode=CHOOSE_RANDOM, ticks=120))
    a = grid_from_lines(12, 18, [("h", 6, 0, 16, 1), ("v", 9, 0, 6, 1), ("v", 4, 6, 11, 1)], [], 22)
    out.append(dict(name="long_quick_merge_and_fork", arrs=fill_roads(a, 1.0, 22), perfect=False, mode=CHOOSE_RANDOM, ticks=100))
    for i, seed in enumerate(RING_NETWORK_SEEDS):
        out.append(dict(name=f"perfect_ring_network_{i}", arrs=ring_network(seed), perfect=True,
                        mode=[CHOOSE_RANDOM, CHOOSE_RANDOM, CHOOSE_TAPE, CHOOSE_X][i % 4], ticks=14))
        if i % 3 == 0:      # the same network in quick mode: full rings stay frozen (steppers.fut:177-185)
            out.append(dict(name=f"quick_ring_network_{i}", arrs=ring_network(seed), perfect=False, mode=CHOOSE_RANDOM, ticks=8))
    return out


PERFECT_RANDOM_SEEDS = []
RING_NETWORK_SEEDS = []


def ring_network(seed, gh=22, gw=26):
    """rings of random size, sense and position (they may overlap and share road cells, like scenarios.fut:151-177) plus a
    few inner one-way roads; everything stays one cell away from the border"""
    r = np.random.default_rng(seed)
    lines = []
    for _ in range(int(r.integers(2, 5))):
        n = int(r.integers(2, 9))
        y0, x0 = int(r.integers(1, gh - n - 1)), int(r.integers(1, gw - n - 1))
        lines += ring(y0, x0, n, int(r.choice([-1, 1])))
    for _ in range(int(r.integers(0, 4))):
        if r.random() < 0.5:
            a, b = sorted(r.integers(1, gw - 1, 2).tolist())
            lines.append(("h", int(r.integers(1, gh - 1)), a, b, int(r.choice([-1, 1]))))
        else:
            a, b = sorted(r.integers(1, gh - 1, 2).tolist())
            lines.append(("v", int(r.integers(1, gw - 1)), a, b, int(r.choice([-1, 1]))))
    a = grid_from_lines(gh, gw, lines, [], seed)
    return fill_roads(a, float(r.choice([0.7, 0.9, 1.0, 1.0])), seed)


def find_ring_network_seeds(find_cycles, want=10):
    seeds, seed = [], 500
    while len(seeds) < want and seed < 1500:
        try:
            n = len(F.call(find_cycles, to_cells(ring_network(seed))))
            if 2 <= n <= 16:
                seeds.append(seed)
        except F.FutharkError:
            pass
        seed += 1
    return seeds


def find_perfect_seeds(find_cycles, want=8):
    seeds, seed = [], 300
    while len(seeds) < want and seed < 2000:
        a = random_grid(seed, 12, 14, n_lines=7, car_p=0.9, weird_p=0.02, offroad_p=0.01)
        try:
            n = len(F.call(find_cycles, to_cells(a)))
            if 1 <= n <= 12:
                seeds.append(seed)
        except F.FutharkError:
            pass
        seed += 1
    return seeds


def render_vectors(ref, data):
    """render h w gh gw scale grid (flowamok.fut:16-64) with symbolic grey / mix: per pixel its kind (0 a plain colour, 1 grey,
    2 grey mixed with a car's colour) and the colour involved"""
    n = 0
    for seed, (gh, gw), scale, (dh, dw) in [(1, (6, 9), 1, (0, 0)), (2, (6, 9), 2, (1, -1)), (3, (5, 8), 3, (0, 0)), (4, (7, 7), 4, (3, -2)),
                                            (5, (6, 10), 5, (0, 0)), (6, (4, 6), 8, (-5, 7)), (7, (5, 5), 10, (0, 0)), (8, (3, 4), 13, (2, 2)),
                                            (9, (8, 11), 7, (-9, -20))]:
        a = palette(random_grid(900 + seed, gh, gw, n_lines=6, car_p=0.5, weird_p=0.1, offroad_p=0.05))
        h, w = gh * scale + dh, gw * scale + dw
        px = F.call(ref["render"], h, w, gh, gw, scale, to_cells(a))
        kind, base = np.zeros((h, w), np.uint8), np.zeros((h, w), np.uint32)
        for y in range(h):
            for x in range(w):
                v = px[y][x]
                if isinstance(v, int):
                    base[y, x] = v
                elif v[0] == "gray":
                    assert v == ("gray", 0.3)
                    kind[y, x] = 1
                else:
                    assert v[:4] == ("mix", 0.5, ("gray", 0.3), 0.5) and isinstance(v[4], int)
                    kind[y, x], base[y, x] = 2, v[4]
        key = f"render/{n}"
        data[f"{key}/meta"] = np.array([gh, gw, scale, h, w], np.int64)
        for k in ("uy", "ux", "dy", "dx", "color"):
            data[f"{key}/{k}"] = a[k]
        data[f"{key}/kind"], data[f"{key}/base"] = kind, base
        print(f"{key}: {gh} x {gw} cells, scale {scale}, window {h} x {w}: {int((kind == 1).sum())} grey, {int((kind == 2).sum())} mixed pixels")
        n += 1
    data["render/n"] = np.array(n)


def edit_vectors(ref, data):
    """add_line_vertical / add_line_horizontal / add_cell (utils.fut:19-38) applied in random order to an empty grid"""
    n = 0
    for seed, (gh, gw) in [(1, (12, 15)), (2, (9, 31)), (3, (20, 20))]:
        r = np.random.default_rng(700 + seed)
        a = palette(random_grid(seed, gh, gw, n_lines=0, car_p=0.0, weird_p=0.0, offroad_p=0.0))
        for k in ("uy", "ux", "dy", "dx"):
            a[k][...] = 0
        a["color"][...] = 0xFFFFFFFF
        cells, edits = to_cells(a), []
        for _ in range(40):
            kind = int(r.integers(0, 3))
            if kind == 0:
                y0, y1 = sorted(r.integers(0, gh, 2).tolist())
                x, flow = int(r.integers(0, gw)), int(r.choice([-1, 1]))
                cells = F.call(ref["add_line_vertical"], y0, y1, x, flow, cells)
                edits.append([0, y0, y1, x, flow, 0, 0])
            elif kind == 1:
                x0, x1 = sorted(r.integers(0, gw, 2).tolist())
                y, flow = int(r.integers(0, gh)), int(r.choice([-1, 1]))
                cells = F.call(ref["add_line_horizontal"], y, x0, x1, flow, cells)
                edits.append([1, y, x0, x1, flow, 0, 0])
            else:
                y, x, col = int(r.integers(0, gh)), int(r.integers(0, gw)), int(PALETTE[int(r.integers(0, len(PALETTE)))])
                dy, dx = int(r.integers(-1, 2)), int(r.integers(-1, 2))
                cells = F.call(ref["add_cell"], y, x, col, {"y": dy, "x": dx}, cells)
                edits.append([2, y, x, col, dy, dx, 0])
        st = from_cells(cells)
        key = f"edits/{n}"
        data[f"{key}/meta"] = np.array([gh, gw], np.int64)
        data[f"{key}/edits"] = np.array(edits, np.int64)
        for k in ("uy", "ux", "dy", "dx", "color"):
            data[f"{key}/{k}"] = st[k]
        n += 1
    data["edits/n"] = np.array(n)


def main():
    ref = load_reference()
    quick, perfect = ref["stepper_quick"].names["step"], ref["stepper_perfect"].names["step"]
    find_cycles = ref["stepper_perfect"].names["find_cycles"]
    data, names = {}, []
    t00 = time.time()
    PERFECT_RANDOM_SEEDS[:] = find_perfect_seeds(find_cycles)
    RING_NETWORK_SEEDS[:] = find_ring_network_seeds(find_cycles)
    print("perfect-mode random networks: seeds", PERFECT_RANDOM_SEEDS, "ring networks: seeds", RING_NETWORK_SEEDS, flush=True)
    for c in cases():
        t0 = time.time()
        a = palette(c["arrs"])
        gh, gw = a["uy"].shape
        tape = np.random.default_rng(zlib.crc32(c["name"].encode())).integers(0, 2, 7).astype(np.uint8)
        ch = Chooser(c["mode"], tape, ref["choose_direction_random"])
        cells = to_cells(a)
        key = c["name"]
        names.append(key)
        data[f"{key}/meta"] = np.array([gh, gw, int(c["perfect"]), c["mode"], c["ticks"]], np.int64)
        data[f"{key}/tape"] = tape
        for k in ("uy", "ux", "dy", "dx", "color", "pref", "rng_state"):
            data[f"{key}/t0/{k}"] = a[k]
        cycles = []
        prev_rng = a["rng_state"]
        if c["perfect"]:
            cycles = F.call(find_cycles, cells)
            dense = np.zeros((len(cycles), gh, gw, 2), np.int8)
            for n, grid in enumerate(cycles):
                for y in range(gh):
                    for x in range(gw):
                        dense[n, y, x] = (grid[y][x]["y"], grid[y][x]["x"])
            data[f"{key}/cycles"] = dense
        for t in range(1, c["ticks"] + 1):
            ch.tick = t - 1
            if c["perfect"]:
                cells, leaks = F.call(perfect, gh, gw, ch.fn, cycles, (), cells)
            else:
                cells, leaks = F.call(quick, gh, gw, ch.fn, (), cells)
            st = from_cells(cells)
            assert not st["transient"].any(), "calculated / direction must be false between ticks (steppers.fut:140-145)"
            for k in ("dy", "dx", "color", "pref"):
                data[f"{key}/t{t}/{k}"] = st[k]
            changed = np.flatnonzero(st["rng_state"].reshape(-1) != prev_rng.reshape(-1))      # rng states: the cells that drew
            data[f"{key}/t{t}/rng_idx"] = changed.astype(np.int64)
            data[f"{key}/t{t}/rng_val"] = st["rng_state"].reshape(-1)[changed]
            prev_rng = st["rng_state"]
            assert np.array_equal(st["uy"], a["uy"]) and np.array_equal(st["ux"], a["ux"])
            data[f"{key}/t{t}/leaks"] = leak_rows(leaks)
        n_leaks = sum(len(data[f"{key}/t{t}/leaks"]) for t in range(1, c["ticks"] + 1))
        print(f"{key:36s} {gh:3d} x {gw:3d}  ticks {c['ticks']:3d}  cycles {len(cycles):3d}  leaks {n_leaks:4d}  {time.time() - t0:6.1f} s", flush=True)
    data["names"] = np.array(names)
    render_vectors(ref, data)
    edit_vectors(ref, data)
    np.savez_compressed(HERE / "stepper_vectors.npz", **data)
    print(f"wrote {HERE / 'stepper_vectors.npz'} ({(HERE / 'stepper_vectors.npz').stat().st_size} bytes) in {time.time() - t00:.0f} s")


if __name__ == "__main__":
    main()
