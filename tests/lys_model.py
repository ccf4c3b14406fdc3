"""TEST INFRASTRUCTURE: the two lys frontends (explorer/explorer.fut, designer/designer.fut) restated on the CPU oracle,
as the checker for libflowamok_explorer.so / libflowamok_designer.so (flowamok_b200/lys/).  Written from the .fut
sources, independently of the C++; line numbers refer to /root/reference/explorer/explorer.fut and
/root/reference/designer/designer.fut.  Mutates in place (the tests compare one linear history)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from oracle import oracle as O

SDLK_SPACE, SDLK_1, SDLK_2, SDLK_a, SDLK_p, SDLK_r = 32, ord("1"), ord("2"), ord("a"), ord("p"), ord("r")
SDLK_RIGHT, SDLK_LEFT, SDLK_DOWN, SDLK_UP = 0x4000004F, 0x40000050, 0x40000051, 0x40000052
f32 = np.float32


class FutharkError(Exception):
    """what the Futhark program would abort with (index out of bounds, shape coercion)"""


def join_rng(rngs):
    arr = (O._Rng * len(rngs))(*[r.c for r in rngs])
    j = O.lib().fo_join_rng(arr, len(rngs))
    return O.Rng(j.state, j.inc)


def steps_due(td, time_unused, per_second):
    """explorer.fut:136-139 / designer.fut:80-83 in f32"""
    time_total = f32(td) + f32(time_unused)
    steps_new = time_total * f32(per_second)
    n = int(steps_new)                                     # i64.f32 truncates
    return n, f32(time_total - f32(n) / f32(per_second))


class ExplorerModel:
    GH, GW, N = 108, 192, 14                               # :48-49, :9-22

    def init_grid(self, rng, sid):                         # :85-88
        g = O.Grid(self.GH, self.GW)
        rng = g.create(rng)
        g.scenario_init(sid)
        return rng, g

    def __init__(self, seed, h, w):                        # :89-101
        rng = O.Rng.from_seed([int(np.int32(np.uint32(seed)))])
        outs, self.grids = [], []
        for k, r in enumerate(rng.split(self.N)):
            r2, g = self.init_grid(r, k)
            outs.append(r2)
            self.grids.append(g)
        self.perfect = True
        self.rng = join_rng(outs)
        self.sid = 0
        self.cycles = self.grids[0].find_cycles()
        self.h, self.w, self.scale, self.auto, self.time_unused = h, w, 10, False, f32(0)
        self.sps = 30
        self.steps = [0] * self.N
        self.n_cell_leaks = 0
        self.chooser = O.Chooser(O.CHOOSE_RANDOM)

    @property
    def grid(self):
        return self.grids[self.sid]

    def step_n(self, n):                                   # :109-130
        g, step_cur, leaks = self.grid, self.steps[self.sid], 0
        for _ in range(n):
            if self.perfect:
                k, _l = g.step_perfect(self.chooser, self.cycles)
            else:
                k, _l = g.step_quick(self.chooser)
            rec = g.scenario_step(self.sid, step_cur, self.rng)
            if rec and self.perfect:
                self.cycles = g.find_cycles()
            step_cur += 1
            leaks += k
        self.steps[self.sid] = step_cur
        self.n_cell_leaks = leaks

    def event_step(self, td):                              # :134-144
        if not self.auto:
            return
        n, unused = steps_due(td, self.time_unused, self.sps)
        if n > 0:
            self.step_n(n)
        self.time_unused = unused

    def keydown(self, key):                                # :145-198
        if key == SDLK_SPACE:
            self.step_n(1)
            self.auto = False
        elif key == SDLK_LEFT:
            self.sid = (self.sid - 1) % self.N
            if self.perfect:
                self.cycles = self.grid.find_cycles()
            self.n_cell_leaks = 0
        elif key == SDLK_RIGHT:
            self.sid = (self.sid + 1) % self.N
            if self.perfect:
                self.cycles = self.grid.find_cycles()
        elif key == SDLK_a:
            self.auto = not self.auto
        elif key == SDLK_p:
            self.perfect = not self.perfect
            if self.perfect:
                self.cycles = self.grid.find_cycles()
        elif key == SDLK_UP:
            self.sps += 1
        elif key == SDLK_DOWN:
            self.sps = max(1, self.sps - 1)
        elif key == SDLK_r:
            self.rng, g = self.init_grid(self.rng, self.sid)
            self.grids[self.sid] = g
            if self.perfect:
                self.cycles = g.find_cycles()
            self.steps[self.sid] = 0
            self.auto = False
        elif key == SDLK_1:
            self.scale += 1
        elif key == SDLK_2:
            self.scale = max(1, self.scale - 1)

    def mouse(self, buttons, x, y):                        # :199-200 `case _ -> s`
        pass

    def resize(self, h, w):                                # :105-107
        self.h, self.w = h, w

    def render(self):                                      # :202-203
        px = np.zeros((self.h, self.w), np.uint32)
        O.lib().fo_render(self.grid.p, self.h, self.w, self.scale, px.ctypes.data)
        return px

    def text_content(self, fps):                           # :77-81
        return (int(f32(fps)), self.sid, int(self.auto), self.sps, self.steps[self.sid], int(self.perfect), len(self.cycles),
                self.n_cell_leaks)

    def fields(self):
        return dict(h=self.h, w=self.w, scale=self.scale, auto=int(self.auto), steps_auto_per_second=self.sps,
                    scenario_id=self.sid, steps=self.steps[self.sid], perfect_simulation=int(self.perfect),
                    n_cycles=len(self.cycles), n_cell_leaks=self.n_cell_leaks, rng_state=self.rng.state, rng_inc=self.rng.inc,
                    time_unused=self.time_unused)


def explorer_text_format():                                # :70-75
    names = "|".join(O.scenario_name(i) for i in range(14))
    return ("FPS: %d\nScenario: %[" + names + "]\nAuto: %[no|yes]\nSteps per second: %d\nSteps: %ld\nPerfect: %[false|true]\n"
            "Cycles detected: %ld\nCell leaks this step: %ld")


class DesignerModel:
    def init_grid(self, gh, gw, rng):                      # :32-36
        if gh <= 35 or gw <= 10:
            raise FutharkError("index out of bounds")
        g = O.Grid(gh, gw)
        rng = g.create(rng)
        g.add_line_horizontal(25, 0, 10, 1)
        g.add_line_horizontal(35, 0, 10, -1)
        return rng, g

    def __init__(self, seed, h, w):                        # :38-47
        self.scale = 10
        self.h, self.w, self.gh, self.gw = h, w, h // self.scale, w // self.scale
        self.rng, self.grid = self.init_grid(self.gh, self.gw, O.Rng.from_seed([int(np.int32(np.uint32(seed)))]))
        self.time_unused, self.running, self.sps = f32(0), True, 30
        self.n_steps_total, self.n_leaks, self.cursor_prev = 0, 0, None
        self.chooser = O.Chooser(O.CHOOSE_RANDOM)

    def coerce(self):                                      # `s.grid :> [gh][gw]cell`
        if (self.grid.gh, self.grid.gw) != (self.gh, self.gw):
            raise FutharkError("shape coercion")

    def resize(self, h, w):                                # :49-55
        self.h, self.w, self.gh, self.gw = h, w, h // self.scale, w // self.scale

    def step_n(self, n):                                   # :59-75
        self.coerce()
        leaks = 0
        for _ in range(n):
            if self.n_steps_total % 10 == 0:
                col = int(O.lib().fo_random_color(C.byref(self.rng.c)))
                self.grid.add_cell(25, 0, col, 0, 1)
            leaks += self.grid.step_quick(self.chooser)[0]
            self.n_steps_total += 1
        self.n_leaks = int(np.int32(leaks))

    def event_step(self, td):                              # :77-86
        if not self.running:
            return
        n, unused = steps_due(td, self.time_unused, self.sps)
        if n > 0:
            self.step_n(n)
        self.time_unused = unused

    def keydown(self, key):                                # :88-95
        if key == SDLK_SPACE:
            self.running = not self.running
        elif key == SDLK_r:
            self.rng, self.grid = self.init_grid(self.gh, self.gw, self.rng)

    def add_lines(self, p0, p1):                           # :97-110
        (gy0, gx0), (gy1, gx1) = p0, p1
        g = self.grid

        def inside(y, x):
            return 0 <= y < g.gh and 0 <= x < g.gw
        ylo, yhi, xlo, xhi = min(gy0, gy1), max(gy0, gy1), min(gx0, gx1), max(gx0, gx1)
        if gx0 != gx1:
            if not (inside(gy0, xlo) and inside(gy0, xhi)):
                raise FutharkError("index out of bounds")
            g.add_line_horizontal(gy0, xlo, xhi, int(np.sign(gx1 - gx0)))
        if gy0 != gy1:
            if not (inside(ylo, gx1) and inside(yhi, gx1)):
                raise FutharkError("index out of bounds")
            g.add_line_vertical(ylo, yhi, gx1, int(np.sign(gy1 - gy0)))

    def mouse(self, buttons, x, y):                        # :112-125
        if buttons == 0b001:
            if 0 <= y < self.h and 0 <= x < self.w:
                cur = (y // self.scale, x // self.scale)
                if self.cursor_prev is not None:
                    self.add_lines(self.cursor_prev, cur)
                self.cursor_prev = cur
        else:
            self.cursor_prev = None

    def render(self):                                      # :133-135
        self.coerce()
        px = np.zeros((self.h, self.w), np.uint32)
        O.lib().fo_render(self.grid.p, self.h, self.w, self.scale, px.ctypes.data)
        return px

    def text_content(self, fps):                           # :27-28
        return (int(f32(fps)), self.sps, self.n_leaks)

    def fields(self):
        cy, cx = self.cursor_prev if self.cursor_prev is not None else (None, None)
        return dict(h=self.h, w=self.w, gh=self.gh, gw=self.gw, scale=self.scale, running=int(self.running),
                    steps_per_second=self.sps, n_steps_total=self.n_steps_total, n_leaks=self.n_leaks,
                    cursor_some=int(self.cursor_prev is not None), cursor_y=cy, cursor_x=cx, rng_state=self.rng.state,
                    rng_inc=self.rng.inc, time_unused=self.time_unused)
