#!/usr/bin/env python
"""Mechanical transcription check of the scenario tables.

Reads /root/reference/explorer/scenarios.fut, evaluates every scenario's `init` body with a tiny interpreter for the
subset of Futhark those bodies use (let-chains of add_line_* / add_cell calls, `loop grid for v < n do ...`, integer
expressions, `i64.bool`), on the explorer's 108 x 192 grid (explorer.fut:48-49), and writes the resulting road /
heading / colour-NAME planes to tests/golden/scenario_inits.npz (derived data only; no reference source is copied).
tests/test_scenario_tables.py compares the oracle's and the product's scenario tables with these planes.

The reference tree does not travel to the GPU box, so the fixture is committed; run this script again only if the
reference changes:  python tests/golden/make_scenario_fixture.py
"""
import re
import sys
from pathlib import Path

import numpy as np

REF = Path("/root/reference/explorer")
OUT = Path(__file__).resolve().parent
GH, GW = 108, 192
# explorer.fut:9-22
ORDER = ["single_fork", "tight_queues", "basic_cycle", "crossing", "close_crossing", "crossroads", "small_tight_cycle",
         "independent_tight_cycles", "overlapping_tight_cycles", "hits_an_edge", "adding_lines", "multi_spill",
         "mk_many_cycles", "t_cross"]
COLOURS = ["white", "red", "green", "blue", "yellow", "orange", "magenta", "violet", "brown", "black"]


def atoms(s):
    """split a Futhark application's arguments at top-level whitespace"""
    out, depth, cur = [], 0, ""
    for ch in s.strip():
        if ch in "({":
            depth += 1
        if ch in ")}":
            depth -= 1
        if ch.isspace() and depth == 0:
            if cur:
                out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur:
        out.append(cur)
    return out


def py(e):
    e = e.replace("'", "_p")
    e = re.sub(r"i64\.bool\s*\(([^()]*)\)", r"int(\1)", e)
    e = e.replace("n -1", "n - 1")
    return e


class Rec:
    def __init__(self):
        self.uy = np.zeros((GH, GW), np.int8); self.ux = np.zeros((GH, GW), np.int8)
        self.dy = np.zeros((GH, GW), np.int8); self.dx = np.zeros((GH, GW), np.int8)
        self.col = np.zeros((GH, GW), np.uint8)     # index into COLOURS; 0 = white (create_cell, utils.fut:10)

    def add_line_horizontal(self, y, x0, x1, flow):   # utils.fut:27-33
        for x in range(x0, x1 + 1):
            self.ux[y, x] = flow

    def add_line_vertical(self, y0, y1, x, flow):     # utils.fut:19-25
        for y in range(y0, y1 + 1):
            self.uy[y, x] = flow

    def add_cell(self, y, x, colour, d):              # utils.fut:35-38
        self.col[y, x] = COLOURS.index(colour)
        self.dy[y, x], self.dx[y, x] = d


def run_init(body, env, rec=None, sub=None):
    """body: the text between `def init ... =` and `def step`; sub(name, rec) runs another module's init on rec."""
    rec = rec or Rec()
    lines = [ln.strip() for ln in body.splitlines() if ln.strip()]
    src, ind = [], 0

    def emit(s):
        src.append("    " * ind + s)

    def call(txt):
        txt = txt.strip()
        assert txt.endswith(" grid"), txt
        a = atoms(txt[:-5])
        f, args = a[0], a[1:]
        if f.endswith(".init") and not args:          # e.g. `small_tight_cycle.init grid`
            return f"sub({f[:-5]!r}, rec)"
        if f == "add_cell":
            m = re.fullmatch(r"\{y=\s*(-?\s*\d+),\s*x=\s*(-?\s*\d+)\}", args[3])
            assert m and args[2].startswith("argb."), txt
            d = (int(m.group(1).replace(" ", "")), int(m.group(2).replace(" ", "")))
            return f"rec.add_cell({py(args[0])}, {py(args[1])}, {args[2][5:]!r}, {d})"
        assert f in ("add_line_horizontal", "add_line_vertical"), txt
        return f"rec.{f}({', '.join(py(x) for x in args)})"

    for ln in lines:
        ln = re.sub(r"^do\s+", "", ln)
        if ln == "in grid":
            ind = 0
            continue
        m = re.fullmatch(r"(?:let grid =\s*)?loop grid for (\w+) < (\w+)(?:\s+do\s+(.*))?", ln)
        if m:
            emit(f"for {m.group(1)} in range({m.group(2)}):")
            if m.group(3):
                ind += 1; emit(call(m.group(3))); ind -= 1
            else:
                ind += 1
            continue
        if ln == "let grid =":
            continue
        m = re.fullmatch(r"let grid = (.*)", ln)
        if m:
            emit(call(m.group(1)))
            continue
        m = re.fullmatch(r"let (\(?[\w', ]+\)?) = (.*)", ln)
        if m:
            emit(f"{py(m.group(1))} = {py(m.group(2))}")
            continue
        raise SystemExit(f"unparsed line: {ln!r}")
    exec("\n".join(src), {"rec": rec, "sub": sub, **env})
    return rec, src


def main():
    text = (REF / "scenarios.fut").read_text()
    mods = {}
    for m in re.finditer(r"^module (\w+)[^\n]*= \{\n(.*?)^\}", text, re.S | re.M):
        mods[m.group(1)] = m.group(2)
    planes, steps, programs = {}, {}, {}

    def init_of(name):
        return re.search(r"def init [^\n]*=\n(.*?)\n\s*def step", mods[name], re.S).group(1)

    def sub(name, rec):
        run_init(init_of(name), {}, rec, sub)

    for sid, name in enumerate(ORDER):
        body = mods[name]
        env = {"n": 2} if name == "mk_many_cycles" else {}   # explorer.fut:21 `mk_many_cycles { def n = 2i64 }`
        rec, src = run_init(init_of(name), env, None, sub)
        for k in ("uy", "ux", "dy", "dx", "col"):
            planes[f"s{sid:02d}_{k}"] = getattr(rec, k)
        programs[name] = src
    np.savez_compressed(OUT / "scenario_inits.npz", colours=np.array(COLOURS), **planes)
    print("wrote", OUT / "scenario_inits.npz",
          {n: int((planes[f's{i:02d}_uy'] != 0).sum() + (planes[f's{i:02d}_ux'] != 0).sum()) for i, n in enumerate(ORDER)})


if __name__ == "__main__":
    sys.exit(main())
