"""The scenario tables exist three times (oracle C, flowamok_b200/scenarios.py, csrc/fa_futhark.inc), written by one
author.  tests/golden/scenario_inits.npz is derived MECHANICALLY from /root/reference/explorer/scenarios.fut by
tests/golden/make_scenario_fixture.py (a small interpreter for the init bodies); here the oracle's table and the Python
product table are compared with it plane by plane.  The third table (fa_futhark.inc) is compared with the oracle on
the GPU (test_gpu_parity.py::test_futhark_entry_points_all_scenarios)."""
import numpy as np
import pytest

from util import O, ROOT

from flowamok_b200 import scenarios as S

GOLD = np.load(ROOT / "tests/golden/scenario_inits.npz")
COLOUR_OF = {"white": S.WHITE, "red": S.RED, "green": S.GREEN, "blue": S.BLUE, "yellow": S.YELLOW, "orange": S.ORANGE,
             "magenta": S.MAGENTA, "violet": S.VIOLET, "brown": S.BROWN, "black": S.BLACK}
GH, GW = 108, 192


def gold_planes(sid):
    names = [str(c) for c in GOLD["colours"]]
    lut = np.array([COLOUR_OF[n] for n in names], np.uint32)
    p = {k: GOLD[f"s{sid:02d}_{k}"] for k in ("uy", "ux", "dy", "dx")}
    p["color"] = lut[GOLD[f"s{sid:02d}_col"]]
    return p


class Recorder:
    """stands in for flowamok_b200.Grid: the point-edit API of utils.fut:19-38 on numpy planes"""

    def __init__(self):
        self.uy = np.zeros((GH, GW), np.int8); self.ux = np.zeros((GH, GW), np.int8)
        self.dy = np.zeros((GH, GW), np.int8); self.dx = np.zeros((GH, GW), np.int8)
        self.color = np.full((GH, GW), S.WHITE, np.uint32)

    def add_line_vertical(self, y0, y1, x, flow):
        self.uy[y0:y1 + 1, x] = flow

    def add_line_horizontal(self, y, x0, x1, flow):
        self.ux[y, x0:x1 + 1] = flow

    def add_cell(self, y, x, color, dy, dx):
        self.color[y, x] = color; self.dy[y, x] = dy; self.dx[y, x] = dx


@pytest.mark.parametrize("sid", range(14))
def test_oracle_scenario_init_matches_scenarios_fut(sid):
    g = O.Grid(GH, GW)
    g.create(O.Rng.from_seed([1]))
    g.scenario_init(sid)
    a, want = g.arrays(), gold_planes(sid)
    for k in ("uy", "ux", "dy", "dx", "color"):
        assert np.array_equal(a[k], want[k]), (sid, k)


@pytest.mark.parametrize("sid", range(14))
def test_product_scenario_init_matches_scenarios_fut(sid):
    r = Recorder()
    S.init(sid, r)
    want = gold_planes(sid)
    for k in ("uy", "ux", "dy", "dx", "color"):
        assert np.array_equal(getattr(r, k), want[k]), (sid, k)


def test_scenario_names_and_count():
    assert S.N_SCENARIOS == 14
    for sid in range(14):
        assert O.lib().fo_scenario_name(sid).decode() == S.NAMES[sid]
