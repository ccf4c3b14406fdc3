"""Opt-in fuzzer (not collected by pytest): random road networks, densities, shapes and emulation variants, the CPU
emulation of the kernel programs against the oracle, every tick.  usage: python tests/fuzz_emul.py [seconds] [seed]"""
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 1)[0])
from util import O, EmulGrid, assert_state_equal, oracle_grid, random_grid, state_of  # noqa: E402


def one(seed):
    r = np.random.default_rng(seed)
    gh, gw = int(r.integers(1, 90)), int(r.integers(1, 260))
    if r.random() < 0.2:
        gw = int(r.integers(900, 2200))
        gh = int(r.integers(2, 30))
    n_lines = int(r.integers(1, 3 + (gh + gw) // 2))
    arrs = random_grid(seed, gh, gw, n_lines=n_lines, car_p=float(r.choice([0.05, 0.3, 0.6, 0.95])),
                       weird_p=float(r.choice([0.0, 0.03, 0.2])), offroad_p=float(r.choice([0.0, 0.01, 0.1])))
    variant = int(r.integers(0, 3))
    # find_cycles returns dense [n][gh][gw] masks and enumerates every cycle of the road graph: small networks only
    perfect = r.random() < 0.4 and gh * gw <= 6000 and n_lines <= 10
    mode = int(r.choice([O.CHOOSE_RANDOM, O.CHOOSE_RANDOM, O.CHOOSE_Y, O.CHOOSE_X, O.CHOOSE_TAPE]))
    tape = r.integers(0, 2, int(r.integers(1, 9))).astype(np.uint8)
    og, eg, ch = oracle_grid(arrs), EmulGrid(arrs), O.Chooser(mode, tape if mode == O.CHOOSE_TAPE else None)
    rings = None
    if perfect:
        cyc = og.find_cycles()
        if len(cyc) > 64:
            return
        rings = og.rings_from_masks(cyc)
    for t in range(int(r.integers(3, 14))):
        if perfect:
            nl, _ = og.step_perfect(ch, cyc)
            ne = eg.step(variant=variant, rings=rings, chooser=mode, tape_bit=int(tape[t % len(tape)]))
        else:
            nl, _ = og.step_quick(ch)
            ne = eg.step(variant=variant, chooser=mode, tape_bit=int(tape[t % len(tape)]))
        assert nl == ne, (seed, t, "leaks", nl, ne)
        assert_state_equal(state_of(og), eg.state(), f"seed {seed} {gh}x{gw} variant {variant} perfect {perfect} tick {t}")


if __name__ == "__main__":
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    t0, n = time.time(), 0
    while time.time() - t0 < budget:
        one(seed + n)
        n += 1
    print(f"fuzz ok: {n} cases from seed {seed} in {time.time() - t0:.0f} s")
