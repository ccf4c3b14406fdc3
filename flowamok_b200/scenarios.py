"""Host-side scenario layer (explorer/scenarios.fut:5-274) over the device grid's point-edit API.

This is the "next" row f2 of SURVEY.md 8(f): between ticks a scenario injects cars or adds roads with
add_cell / add_line_* (utils.fut:19-38).  The scenario-level RNG (model.fut:33) and random_color
(random.fut:9-13) run on the host; the arithmetic of diku-dk/cpprandom and athas/matte is restated from
the published algorithms (not vendored in the reference: unpinned).
"""
from __future__ import annotations

import numpy as np

M64 = (1 << 64) - 1

WHITE, BLACK = 0xFFFFFFFF, 0xFF000000
RED, GREEN, BLUE = 0xFFFF0000, 0xFF00FF00, 0xFF0000FF
YELLOW = 0xFFFFFF00
ORANGE = 0xFFFFFF00   # matte: add yellow red saturates
MAGENTA = 0xFFFF00FF
VIOLET = 0xFFFF00FF   # matte: add magenta blue saturates
BROWN = 0xFF7C301C


class Pcg32:
    """rnge = pcg32 (random.fut:4): state/inc u64, XSH-RR output from the old state."""

    def __init__(self, state=0, inc=0):
        self.state, self.inc = state & M64, inc & M64

    @classmethod
    def from_seed(cls, seeds):
        r = cls(0, ((0xda3e39cb94b95bdb << 1) | 1) & M64)
        r.rand()
        for s in seeds:
            r.state = (r.state + (int(np.int32(s)) & M64)) & M64
        r.rand()
        return r

    def rand(self):
        old = self.state
        self.state = (old * 6364136223846793005 + (self.inc | 1)) & M64
        xs = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
        rot = old >> 59
        return ((xs >> rot) | (xs << ((-rot) & 31))) & 0xFFFFFFFF

    def dist_i8(self, lo, hi):
        rng_ = (hi - lo + 1) & 0xFFFFFFFF
        secure_max = 0xFFFFFFFF - 0xFFFFFFFF % rng_
        x = self.rand()
        while x >= secure_max:
            x = self.rand()
        return lo + x // (secure_max // rng_)

    def dist_f32(self, lo, hi):
        x = np.float32(self.rand())
        xf = x / np.float32(0xFFFFFFFF)
        return np.float32(lo) + xf * (np.float32(hi) - np.float32(lo))


def _hash32(x):
    x &= 0xFFFFFFFF
    x = (((x >> 16) ^ x) * 0x45d9f3b) & 0xFFFFFFFF
    x = (((x >> 16) ^ x) * 0x45d9f3b) & 0xFFFFFFFF
    return (x >> 16) ^ x


def split_rng(n, r: Pcg32):
    out = []
    for i in range(n):
        h = _hash32(i)
        if h & 0x80000000:
            h |= 0xFFFFFFFF00000000     # u64.i32 sign-extends
        c = Pcg32(r.state ^ h, r.inc)
        c.rand()
        out.append(c)
    return out


def join_rng(rs):
    s, inc = 1, 0
    for r in rs:
        s = (s * r.state) & M64
        inc |= r.inc
    return Pcg32(s, inc)


def from_rgba(r, g, b, a):
    ch = lambda v: int(np.float32(min(np.float32(1), max(np.float32(0), np.float32(v)))) * np.float32(255))
    return (ch(a) << 24) | (ch(r) << 16) | (ch(g) << 8) | ch(b)


def random_color(rng: Pcg32):
    r = rng.dist_f32(0.2, 0.8)
    g = rng.dist_f32(0.2, 0.8)
    b = rng.dist_f32(0.2, 0.8)
    return from_rgba(r, g, b, 1.0)


NAMES = ["single fork", "tight queues", "basic cycle", "crossing", "close crossing", "crossroads",
         "small tight cycle", "independent tight cycles", "overlapping tight cycles", "hits an edge",
         "adding lines", "multi spill", "many cycles (n = 2)", "t cross"]
N_SCENARIOS = len(NAMES)


def _small_tight_cycle(g):
    g.add_line_horizontal(11, 11, 13, 1); g.add_line_vertical(11, 13, 13, 1)
    g.add_line_horizontal(13, 11, 13, -1); g.add_line_vertical(11, 13, 11, -1)
    for i in range(2): g.add_cell(11, 11 + i, RED, 0, 1)
    for i in range(2): g.add_cell(11 + i, 13, BLUE, 1, 0)
    for i in range(2): g.add_cell(13, 12 + i, ORANGE, 0, -1)
    for i in range(2): g.add_cell(12 + i, 11, MAGENTA, -1, 0)


def init(sid, g):
    """scenario.init (scenarios.fut), sid as in explorer/explorer.fut:9-22; g: any object with the
    add_line_* / add_cell methods of flowamok_b200.Grid."""
    H, V, A = g.add_line_horizontal, g.add_line_vertical, g.add_cell
    if sid == 0:
        H(10, 10, 20, 1); V(10, 40, 20, 1); H(30, 5, 25, -1)
    elif sid == 1:
        H(40, 10, 20, 1); V(10, 40, 20, -1)
    elif sid == 2:
        H(10, 10, 30, 1); V(10, 30, 30, 1); H(30, 10, 30, -1); V(10, 30, 10, -1)
        A(10, 20, RED, 0, 1); A(20, 30, BLUE, 1, 0); A(30, 20, BROWN, 0, -1); A(20, 10, MAGENTA, -1, 0)
    elif sid in (3, 4):
        H(20, 10, 30, 1); V(10, 30, 20, 1)
    elif sid == 5:
        H(29, 1, 58, -1); H(30, 1, 58, 1); V(1, 58, 29, 1); V(1, 58, 30, -1)
    elif sid == 6:
        _small_tight_cycle(g)
    elif sid == 7:
        _small_tight_cycle(g)
        H(10, 10, 14, -1); V(10, 14, 14, -1); H(14, 10, 14, 1); V(10, 14, 10, 1)
        for i in range(4): A(10, 11 + i, RED, 0, -1)
        for i in range(4): A(11 + i, 14, BLUE, -1, 0)
        for i in range(4): A(14, 10 + i, ORANGE, 0, 1)
        for i in range(4): A(10 + i, 10, MAGENTA, 1, 0)
    elif sid == 8:
        H(10, 10, 14, -1); V(10, 24, 14, -1); H(24, 10, 14, 1); V(10, 24, 10, 1)
        H(14, 14, 24, 1); V(14, 20, 24, 1); H(20, 14, 24, -1)
        for i in range(4): A(10, 11 + i, RED, 0, -1)
        for i in range(14): A(11 + i, 14, BLUE, -1, 0)
        for i in range(4): A(24, 10 + i, ORANGE, 0, 1)
        for i in range(14): A(10 + i, 10, MAGENTA, 1, 0)
        for i in range(9): A(14, 15 + i, BROWN, 0, 1)
        for i in range(6): A(14 + i, 24, VIOLET, 1, 0)
        for i in range(10): A(20, 15 + i, BROWN, 0, -1)
    elif sid == 9:
        H(10, 5, 15, 1); V(5, 15, 16, -1); A(10, 10, BROWN, 0, 1)
    elif sid == 10:
        H(1, 1, 6, 1)
    elif sid == 11:
        V(1, 9, 10, 1); H(10, 1, 9, 1); V(11, 19, 10, -1); H(10, 11, 19, -1)
    elif sid == 12:
        n, yo, xo = 2, 19, 10
        for y in range(n):
            for x in range(n):
                y1, x1 = y + 1, x + 1
                H(yo + 10 * y, xo + 10 * x, xo + 10 * x1 - (x == n - 1), 1)
                V(yo + 10 * y, yo + 10 * y1 - (y == n - 1), xo + 10 * x1 - 1, 1)
                H(yo + 10 * y1 - 1, xo + 10 * x - 1 + (x == 0), xo + 10 * x1 - 1, -1)
                V(yo + 10 * y - 1 + (y == 0), yo + 10 * y1 - 1, xo + 10 * x, -1)
        A(yo, xo, GREEN, 0, 1)
    elif sid == 13:
        y = 25
        H(y, 0, 10, 1); H(y - 1, 10, 30, 1); H(y + 1, 10, 30, 1); V(y, y + 1, 10, 1); V(y - 1, y, 10, -1)
        A(y, 0, MAGENTA, 0, 1)


def step(sid, g, steps, rng: Pcg32) -> bool:
    """scenario.step grid steps rng -> recompute_cycles (rng advanced in place)."""
    A = g.add_cell
    if sid == 0:
        if steps % 5 == 0:
            A(10, 10, random_color(rng), 0, 1)
    elif sid == 1:
        A(40, 10, random_color(rng), 0, 1)
    elif sid == 3:
        if steps % 5 == 0:
            A(20, 10, RED, 0, 1); A(10, 20, BLUE, 1, 0)
    elif sid == 4:
        A(20, 10, random_color(rng), 0, 1)
        A(10, 20, random_color(rng), 1, 0)
    elif sid == 5:
        east = rng.dist_i8(0, 1); south = rng.dist_i8(0, 4); north = rng.dist_i8(0, 9); west = rng.dist_i8(0, 19)
        if north == 0: A(1, 29, BROWN, 1, 0)
        if east == 0: A(29, 58, GREEN, 0, -1)
        if south == 0: A(58, 30, VIOLET, -1, 0)
        if west == 0: A(30, 1, YELLOW, 0, 1)
    elif sid == 10:
        s = steps + 10
        rec = False
        if s % 5 == 0 and s // 5 < 20:
            k = (s // 10) * 5 + 1
            if (s // 5) % 2 == 1:
                g.add_line_horizontal(k, k, k + 5, 1)
            else:
                g.add_line_vertical(k - 5, k, k, 1)
            rec = True
        if rng.dist_i8(0, 3) == 0:
            A(1, 1, random_color(rng), 0, 1)
        return rec
    elif sid == 11:
        if steps % 5 == 0:
            A(1, 10, RED, 1, 0); A(10, 1, BLUE, 0, 1); A(19, 10, GREEN, -1, 0); A(10, 19, YELLOW, 0, -1)
    return False


def split_states(n, r: Pcg32):
    """vectorised split_rng: the n per-cell states of create_grid (utils.fut:15); inc is shared"""
    i = np.arange(n, dtype=np.uint64).astype(np.uint32)
    with np.errstate(over="ignore"):
        x = ((i >> np.uint32(16)) ^ i) * np.uint32(0x45d9f3b)
        x = ((x >> np.uint32(16)) ^ x) * np.uint32(0x45d9f3b)
        x = (x >> np.uint32(16)) ^ x
        h = x.astype(np.int32).astype(np.int64).astype(np.uint64)      # u64.i32 sign-extends
        st = np.uint64(r.state) ^ h
        return st * np.uint64(6364136223846793005) + np.uint64(r.inc | 1)


def join_states(states, inc):
    with np.errstate(over="ignore"):
        return Pcg32(int(np.multiply.reduce(states.astype(np.uint64))), inc)


class Explorer:
    """explorer/explorer.fut init (:89-101) + step s n (:109-130) for one scenario id on the device."""
    GH, GW = 108, 192                                                   # explorer.fut:48-49

    def __init__(self, ctx, seed, sid, perfect=True):
        gh, gw = self.GH, self.GW
        rng = Pcg32.from_seed([int(np.int32(np.uint32(seed)))])          # :90
        rngs = split_rng(N_SCENARIOS, rng)                               # :91
        per_grid = [split_states(gh * gw, r) for r in rngs]              # :92 create_grid of every scenario
        self.rng = join_rng([join_states(st, r.inc) for st, r in zip(per_grid, rngs)])   # :94
        self.sid, self.perfect = sid, perfect
        z8 = np.zeros((gh, gw), np.int8)
        self.grid = ctx.grid(gh, gw)
        self.grid.upload(uy=z8, ux=z8, dy=z8, dx=z8, color=np.full((gh, gw), WHITE, np.uint32),
                         pref=np.zeros((gh, gw), np.uint8), rng_state=per_grid[sid].reshape(gh, gw),
                         rng_inc=rngs[sid].inc)
        init(sid, self.grid)
        self.step_cur = 0
        self.n_cycles = self.grid.find_cycles() if perfect else 0        # :95-97

    def step(self, n_steps=1):
        leaks = 0
        for _ in range(n_steps):                                         # :113-125
            leaks += self.grid.step(self.perfect, 1)
            rec = step(self.sid, self.grid, self.step_cur, self.rng)
            if rec and self.perfect:
                self.n_cycles = self.grid.find_cycles()
            self.step_cur += 1
        return leaks
