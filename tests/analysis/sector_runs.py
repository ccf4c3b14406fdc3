#!/usr/bin/env python
"""How long are the runs of adjacent 32-byte colour sectors that k_move's phase B1 copies?  (The answer owed to the
question whether bulk-async copies, cp.async.bulk / TMA, could move the colour payload: a bulk copy needs >= 16 contiguous
bytes, and pays an shared-memory round trip + mbarrier per transfer, so it only has a chance on LONG runs.)
Oracle run of the synthetic network (statistics do not depend on the grid size); a sector is copied iff a car arrived in
it in this tick or the previous one.  Analysis script under tests/ because it runs the ORACLE (test infrastructure).
python tests/analysis/sector_runs.py [size] [ticks]"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from oracle import oracle as O

S = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
T = int(sys.argv[2]) if len(sys.argv) > 2 else 40
g = O.Grid(S, S)
g.generate_synthetic(seed=1)
ch = O.Chooser(O.CHOOSE_RANDOM)


def snap():
    a = g.arrays()
    return ((a["dy"] != 0) | (a["dx"] != 0)).copy(), a["color"].copy()


prev_arr = np.zeros((S, S), bool)
occ, col = snap()
for t in range(T):
    g.step_quick(ch)
    occ2, col2 = snap()
    arr = occ2 & ~(occ & (col == col2))           # a car that was not here (with this colour) a tick ago
    touched = (arr | prev_arr).reshape(S, S // 8, 8).any(axis=2)   # 32-byte sectors = 8 cells
    prev_arr, occ, col = arr, occ2, col2
pad = np.zeros((S, 1), bool)
t2 = np.concatenate([pad, touched, pad], axis=1).astype(np.int8)
d = np.diff(t2, axis=1)
runs = (np.argwhere(d == -1)[:, 1] - np.argwhere(d == 1)[:, 1])
tot = runs.sum()
print(f"{S}x{S}, tick {T}: {touched.mean() * 100:.1f} % of the colour sectors are copied; arrivals in {arr.mean() * 100:.2f} % of the cells")
for lo, hi in ((1, 1), (2, 3), (4, 7), (8, 15), (16, 10**9)):
    m = (runs >= lo) & (runs <= hi)
    print(f"  runs of {lo}{'' if hi == lo else ('-' + str(hi) if hi < 10**9 else '+')} sectors: {m.sum() / len(runs) * 100:5.1f} % of the runs, "
          f"{runs[m].sum() / tot * 100:5.1f} % of the copied sectors")
