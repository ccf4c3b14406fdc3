#!/usr/bin/env python
"""The README GIF against the reference's OWN source text (not the oracle): README.fut's `init 30 30 10 123` and 399 ticks of
`stepper_perfect.step` (animation.fut:24) executed by tests/golden/futhark_mini.py, with the fork decisions taken from the
398-bit tape recovered from the GIF (tests/test_oracle_golden.py) in place of the unavailable cpprandom draw -- every frame must
classify like the GIF's (tests/golden/readme_gif_classes.npz).  It closes the loop GIF <-> steppers.fut <-> oracle without the
oracle in the middle.  Needs /root/reference, so it runs here only:  python tests/golden/check_gif_with_reference_source.py"""
import sys
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path[:0] = [str(HERE), str(HERE.parent), str(HERE.parent.parent)]
import futhark_mini as F  # noqa: E402
from make_lys_fixture import REF, stubs  # noqa: E402
from make_stepper_fixture import CHOOSE_TAPE, Chooser, from_cells  # noqa: E402
from test_oracle_golden import TAPE, classes  # noqa: E402


def main():
    ref = F.Program(stubs()).load(REF / "README.fut")
    flowamok = F.Program(stubs()).load(REF / "flowamok.fut")
    step = flowamok["stepper_perfect"].names["step"]
    gold = np.load(HERE / "readme_gif_classes.npz")["frames"]
    s = F.call(ref["init"], 30, 30, 10, 123)
    cells, cycles = s["grid"], s["cycles"]
    ch = Chooser(CHOOSE_TAPE, np.array([int(c) for c in TAPE] + [0, 0], np.uint8))
    t0 = time.time()
    for k in range(400):
        st = from_cells(cells)
        assert np.array_equal(classes(st), gold[k]), f"frame {k} differs from the GIF"
        ch.tick = k
        cells, _leaks = F.call(step, 30, 30, ch.fn, cycles, (), cells)
    print(f"README.fut init + 399 ticks of steppers.fut under the interpreter, fork decisions from the GIF's tape: "
          f"400 of 400 frames classify like the GIF ({len(cycles)} cycles, {time.time() - t0:.0f} s)")


if __name__ == "__main__":
    main()
