#!/usr/bin/env python
"""Opt-in fuzzer (needs /root/reference, so it runs here only): the oracle against the reference's own steppers.fut executed by
tests/golden/futhark_mini.py on random small networks -- every field of every cell after every tick, the ordered cell_leak
list, find_cycles (order included) -- beyond the fixed cases of tests/golden/stepper_vectors.npz.
usage: python tests/golden/fuzz_oracle_vs_reference_source.py [seconds] [seed]"""
import sys
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path[:0] = [str(HERE), str(HERE.parent), str(HERE.parent.parent)]
import futhark_mini as F  # noqa: E402
from make_stepper_fixture import (CHOOSE_RANDOM, CHOOSE_TAPE, CHOOSE_X, CHOOSE_Y, Chooser, from_cells, leak_rows, load_reference,  # noqa: E402
                                  palette, ring_network, to_cells)
from util import O, oracle_grid, random_grid  # noqa: E402


def one(ref, seed, tally):
    r = np.random.default_rng(seed)
    perfect = r.random() < 0.35
    if perfect and r.random() < 0.7:
        a = ring_network(seed, int(r.integers(12, 24)), int(r.integers(12, 28)))
    else:
        gh, gw = int(r.integers(1, 26)), int(r.integers(1, 40))
        a = random_grid(seed, gh, gw, n_lines=int(r.integers(1, 3 + (gh + gw) // 2)), car_p=float(r.choice([0.1, 0.4, 0.7, 0.95, 1.0])),
                        weird_p=float(r.choice([0.0, 0.03, 0.2])), offroad_p=float(r.choice([0.0, 0.02, 0.1])))
    a = palette(a)
    gh, gw = a["uy"].shape
    mode = int(r.choice([CHOOSE_RANDOM, CHOOSE_RANDOM, CHOOSE_Y, CHOOSE_X, CHOOSE_TAPE]))
    tape = r.integers(0, 2, int(r.integers(1, 9))).astype(np.uint8)
    cells = to_cells(a)
    cycles = []
    og = oracle_grid(a)
    if perfect:
        try:
            cycles = F.call(ref["stepper_perfect"].names["find_cycles"], cells)
        except F.FutharkError:
            tally["invalid for the reference (find_cycles walks off the grid)"] += 1
            return
        if len(cycles) > 16:
            tally["too many cycles"] += 1
            return
        dense = np.array([[[(c["y"], c["x"]) for c in row] for row in g] for g in cycles], np.int8).reshape(len(cycles), gh, gw, 2)
        masks = og.find_cycles()
        assert np.array_equal(masks, dense), f"seed {seed}: find_cycles"
    ch_ref = Chooser(mode, tape, ref["choose_direction_random"])
    ch_orc = O.Chooser(mode, tape if mode == O.CHOOSE_TAPE else None)
    for t in range(int(r.integers(3, 13))):
        ch_ref.tick = t
        if perfect:
            cells, leaks = F.call(ref["stepper_perfect"].names["step"], gh, gw, ch_ref.fn, cycles, (), cells)
            n, got = og.step_perfect(ch_orc, masks, leaks=True)
        else:
            cells, leaks = F.call(ref["stepper_quick"].names["step"], gh, gw, ch_ref.fn, (), cells)
            n, got = og.step_quick(ch_orc, leaks=True)
        st, oa = from_cells(cells), og.arrays()
        for k in ("dy", "dx", "color", "pref", "rng_state"):
            assert np.array_equal(st[k], oa[k]), f"seed {seed} tick {t}: {k}"
        rows = [[l["y_next"], l["x_next"], l["uy"], l["ux"], l["dy"], l["dx"], l["color"], l["calc"], l["my"], l["mx"], l["pref"],
                 int(np.uint64(l["rng_state"]).astype(np.int64))] for l in got]
        assert rows == leak_rows(leaks).tolist(), f"seed {seed} tick {t}: cell_leak list"
        tally["ticks"] += 1
        tally["leaks"] += n
    tally["perfect cases" if perfect else "quick cases"] += 1


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    ref = load_reference()
    from collections import Counter
    tally, t0, n = Counter(), time.time(), 0
    while time.time() - t0 < budget:
        one(ref, seed + n, tally)
        n += 1
    print(f"oracle == reference source: {n} cases from seed {seed} in {time.time() - t0:.0f} s: {dict(tally)}")


if __name__ == "__main__":
    main()
