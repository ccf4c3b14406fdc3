"""the lys frontends (explorer/explorer.fut, designer/designer.fut) behind the entry points a lys C main
calls (include/flowamok_lys.h, libflowamok_explorer.so / libflowamok_designer.so), against the same state machines
restated on the CPU oracle (tests/lys_model.py): every scalar field of the state, the full grid and the rendered
window after every event of scripted sessions."""
import numpy as np
import pytest

import lys_model as M
from util import STATE_FIELDS, assert_state_equal, state_of

pytestmark = pytest.mark.gpu


def dev_state(g):
    d = g.download()
    return {k: d[k] for k in STATE_FIELDS}


class Session:
    """a program and its model driven by the same events"""

    def __init__(self, name, seed, h, w):
        import flowamok_b200.lys as lys
        self.p = lys.Program(name, device=0)
        self.octx = self.p.operator_context()
        self.s = self.p.init(seed, h, w)
        self.m = (M.ExplorerModel if name == "explorer" else M.DesignerModel)(seed, h, w)
        self.n_events = 0

    def close(self):
        self.p.free(self.s)
        self.p.close()

    def check(self, what, render=False):
        want, got = self.m.fields(), self.p.fields(self.s)
        for k, v in want.items():
            if v is None:
                continue
            if k == "time_unused":
                assert np.float32(v).tobytes() == np.float32(got[k]).tobytes(), f"{what}: time_unused {v!r} != {got[k]!r}"
            else:
                assert got[k] == v, f"{what}: {k} = {got[k]}, the reference has {v}"
        assert_state_equal(state_of(self.m.grid), dev_state(self.p.grid(self.s, self.octx)), what)
        assert self.p.text_content(self.s, 59.94) == self.m.text_content(59.94), what
        if render:
            assert np.array_equal(self.p.render(self.s), self.m.render()), f"{what}: render"

    def do(self, kind, *args, render=False):
        self.n_events += 1
        what = f"event {self.n_events} {kind}{args}"
        if kind == "key":
            self.s = self.p.key(self.s, args[0])
            self.m.keydown(args[0])
        elif kind == "keyup":
            self.s = self.p.key(self.s, args[0], up=True)
        elif kind == "step":
            self.s = self.p.step(self.s, args[0])
            self.m.event_step(args[0])
        elif kind == "resize":
            self.s = self.p.resize(self.s, *args)
            self.m.resize(*args)
        elif kind == "mouse":
            self.s = self.p.mouse(self.s, *args)
            self.m.mouse(*args)
        elif kind == "wheel":
            self.s = self.p.wheel(self.s, *args)
        self.check(what, render=render)


def test_explorer_text_and_constants():
    import flowamok_b200.lys as lys
    p = lys.Program("explorer")
    assert p.text_format() == M.explorer_text_format()
    assert p.grab_mouse() is False
    s = p.init(7, 300, 400)
    assert p.text_colour(s) == 0xFFFFFFFF
    p.free(s)
    p.close()


def test_explorer_init_every_scenario():
    """init (:89-101) builds all fourteen grids; walk through them with RIGHT (and back over the wrap with LEFT)"""
    se = Session("explorer", 123, 1080, 1920)
    se.check("init", render=True)
    for _ in range(14):
        se.do("key", M.SDLK_RIGHT)
    assert se.m.sid == 0
    se.do("key", M.SDLK_LEFT, render=True)
    assert se.m.sid == 13
    se.close()


def test_explorer_session():
    """keys, auto stepping with fractional frame times, perfect/quick, reset, scale and resize; scenarios with random
    injection (4, 5), road edits + cycle recomputation (10), full rings (8), leaks (11)"""
    se = Session("explorer", 2024, 540, 960)
    K = M
    for sid in (1, 4, 5, 8, 10, 11):
        while se.m.sid != sid:
            se.do("key", K.SDLK_RIGHT)
        se.do("key", K.SDLK_SPACE, render=True)
        se.do("key", K.SDLK_a)
        for td in (0.016, 0.05, 0.1, 0.033, 0.2, 0.001):
            se.do("step", td)
        se.do("key", K.SDLK_UP)
        se.do("key", K.SDLK_UP)
        se.do("step", 0.25, render=True)
        se.do("key", K.SDLK_p)                     # quick
        se.do("step", 0.3)
        se.do("step", 0.15)
        se.do("key", K.SDLK_p)                     # perfect again: cycles recomputed
        se.do("step", 0.4)
        for _ in range(3):
            se.do("key", K.SDLK_DOWN)
        se.do("keyup", K.SDLK_a)                   # ignored
        se.do("mouse", 1, 10, 10)                  # ignored
        se.do("wheel", 0, 1)                       # ignored
        se.do("key", K.SDLK_a)                     # auto off
        se.do("step", 1.0)                         # nothing happens
    se.do("key", K.SDLK_1, render=True)            # scale 11: the window cuts the grid
    se.do("key", K.SDLK_2)
    se.do("key", K.SDLK_2)
    se.do("key", K.SDLK_2, render=True)            # scale 8: black margin
    se.do("resize", 333, 777, render=True)
    se.do("key", K.SDLK_r, render=True)            # the scenario starts over with grids of the state's rng
    se.do("key", K.SDLK_SPACE)
    se.do("key", K.SDLK_LEFT)
    se.do("key", K.SDLK_SPACE, render=True)
    se.do("key", ord("x"))                         # unbound
    for _ in range(12):
        se.do("key", K.SDLK_2)                     # scale stops at 1
    assert se.m.scale == 1
    se.do("resize", 120, 200, render=True)
    se.close()


def test_explorer_states_are_values():
    """an entry leaves its input state valid: step the SAME state twice -> the same result, and the old state still renders"""
    import flowamok_b200.lys as lys
    p = lys.Program("explorer")
    octx = p.operator_context()
    s0 = p.init(5, 216, 384)
    for _ in range(4):
        s0 = p.key(s0, M.SDLK_RIGHT)               # scenario 4: cars of random colour every step
    before = dev_state(p.grid(s0, octx))
    a = p.key(s0, M.SDLK_SPACE, keep=True)
    b = p.key(s0, M.SDLK_SPACE, keep=True)
    assert_state_equal(dev_state(p.grid(a, octx)), dev_state(p.grid(b, octx)), "same input, same output")
    assert p.fields(a) == p.fields(b)
    assert_state_equal(before, dev_state(p.grid(s0, octx)), "the input state is untouched")
    assert p.fields(s0)["steps"] == 0 and p.fields(a)["steps"] == 1
    for s in (a, b, s0):
        p.free(s)
    p.close()


def test_designer_session():
    se = Session("designer", 99, 600, 800)
    import flowamok_b200.lys as lys
    assert se.p.text_format() == "FPS: %d\nSteps per second: %d\nLeaks: %d"
    se.check("init", render=True)
    for td in (0.016, 0.05, 0.4, 0.033):
        se.do("step", td)
    # draw an L, then a second stroke that crosses the entry road: a fork and a merge
    se.do("mouse", 1, 105, 255)                    # first point: (25, 10)
    se.do("mouse", 1, 405, 455, render=True)       # (25,10) -> (45,40): row 25 east, column 40 south
    se.do("mouse", 0, 0, 0)                        # button released
    se.do("mouse", 1, 305, 105)                    # (10, 30)
    se.do("mouse", 1, 305, 505)                    # straight down column 30 to row 50
    se.do("mouse", 1, 55, 505)                     # west along row 50
    se.do("mouse", 1, 9999, 9999)                  # outside the window: ignored, the stroke goes on
    se.do("mouse", 1, 55, 355)                     # north along column 5
    se.do("mouse", 2, 55, 355)                     # another button: stroke ends
    for td in (0.5, 0.5, 0.7, 0.2, 1.0):
        se.do("step", td, render=(td == 1.0))
    se.do("key", M.SDLK_SPACE)                     # pause
    se.do("step", 0.5)
    se.do("key", M.SDLK_SPACE)
    se.do("step", 2.0, render=True)
    assert se.m.n_steps_total > 150
    se.do("key", M.SDLK_r, render=True)            # a fresh grid from the state's rng; n_steps_total goes on
    se.do("step", 0.6)
    se.do("wheel", 1, 1)
    se.do("keyup", M.SDLK_r)
    se.close()


def test_designer_errors_like_the_futhark_program():
    """what aborts the Futhark program is an error here, and the input state stays usable"""
    import flowamok_b200.lys as lys
    from flowamok_b200 import FlowamokError
    p = lys.Program("designer")
    with pytest.raises(FlowamokError):             # init_grid draws roads on rows 25 and 35: a 30-row grid is out of bounds
        p.init(1, 300, 800)
    s = p.init(1, 600, 800)
    s2 = p.resize(s, 610, 805, keep=True)          # gh, gw still 61, 80?  610/10 = 61: the grid keeps 60 rows (FIXME in the reference)
    with pytest.raises(FlowamokError, match="shape"):
        p.render(s2)
    with pytest.raises(FlowamokError, match="shape"):
        p.step(s2, 0.5, keep=True)
    s3 = p.mouse(s2, 1, 5, 5, keep=True)
    with pytest.raises(FlowamokError):             # (0,0) -> row 60 of a 60-row grid
        p.mouse(s3, 1, 5, 605, keep=True)
    s4 = p.key(s2, M.SDLK_r, keep=True)            # `r` makes a grid of the new size
    assert p.render(s4).shape == (610, 805)
    assert p.render(s).shape == (600, 800)         # the first state is still a value
    for x in (s, s2, s3, s4):
        p.free(x)
    p.close()
