#!/usr/bin/env python
"""BASELINE config 2 from the reference's own source text: every scenario of the explorer (explorer/explorer.fut:89-130: `lys.init`
with seed 123, then `lys.step s 1` again and again, perfect simulation) EXECUTED by tests/golden/futhark_mini.py for the first
N_TICKS ticks -- the interpreter needs about four seconds per tick of a 108 x 192 grid, so not the 1000 ticks of the config.
Recorded per tick: the digest of the grid (tests/util.state_digest: headings, colours, preferences, rng states), the cell
leaks of the tick, the explorer's rng, the number of cycle_checks; the full grid at ticks 1, 10, 50, 100, 150, 200.

  python tests/golden/make_explorer_runs.py 0 1 2      # these scenario ids -> tests/golden/_parts/explorer_run_<sid>.npz
  python tests/golden/make_explorer_runs.py merge      # -> tests/golden/explorer_runs.npz (committed; derived data only)
"""
import sys
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path[:0] = [str(HERE), str(HERE.parent), str(HERE.parent.parent)]
import futhark_mini as F  # noqa: E402
from make_lys_fixture import K, REF, stubs  # noqa: E402
from make_stepper_fixture import from_cells  # noqa: E402
from util import state_digest  # noqa: E402

N_TICKS, SEED = 200, 123
PARTS = HERE / "_parts"


def run(sid):
    lys = F.Program(stubs()).load(REF / "explorer" / "explorer.fut")["lys"].names
    s = F.call(lys["init"], SEED, 1080, 1920)
    for _ in range(sid):
        s = F.call(lys["event"], ("#keydown", {"key": K["RIGHT"]}), s)
    assert s["scenario_id"] == sid and s["perfect_simulation"]
    data = {"meta": np.array([sid, N_TICKS, SEED, len(s["cycle_checks"])], np.int64)}
    rows = []
    for t in range(1, N_TICKS + 1):
        t0 = time.time()
        s = F.call(lys["step"], s, 1)
        st = from_cells(s["grids"][sid])
        d = state_digest(st)
        rng = s["rng"] - (1 << 64) if s["rng"] >= 1 << 63 else s["rng"]
        rows.append([np.uint64(d[0]).astype(np.int64), np.uint64(d[1]).astype(np.int64), s["n_cell_leaks"], rng, len(s["cycle_checks"]),
                     s["steps"][sid]])
        if t in (1, 10, 50, 100, 150, 200):
            for f in ("uy", "ux", "dy", "dx", "color", "pref"):
                data[f"t{t}/{f}"] = st[f]
        print(f"scenario {sid:2d} tick {t:3d}: leaks {s['n_cell_leaks']} cycles {len(s['cycle_checks'])}  {time.time() - t0:4.1f} s", flush=True)
    data["ticks"] = np.array(rows, np.int64)
    PARTS.mkdir(exist_ok=True)
    np.savez_compressed(PARTS / f"explorer_run_{sid}.npz", **data)


def merge():
    out = {}
    for sid in range(14):
        part = np.load(PARTS / f"explorer_run_{sid}.npz")
        for k in part.files:
            out[f"{sid}/{k}"] = part[k]
    np.savez_compressed(HERE / "explorer_runs.npz", **out)
    print("wrote", HERE / "explorer_runs.npz", (HERE / "explorer_runs.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    if sys.argv[1:] == ["merge"]:
        merge()
    else:
        for a in sys.argv[1:]:
            run(int(a))
