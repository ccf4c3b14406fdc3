"""Opt-in fuzzer (not collected by pytest): the row-band protocol (flowamok_b200.bands.LocalBands, the product's own
driver) over bands backed by the CPU emulation of the kernel programs, random networks / cuts / variants, against the
oracle's single grid on every tick.  usage: python tests/fuzz_bands.py [seconds] [seed]"""
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 1)[0])
from util import O, ROOT, STATE_FIELDS, oracle_grid, random_grid, state_of  # noqa: E402
from test_bands_cpu import EmulBand  # noqa: E402
from flowamok_b200.bands import LocalBands  # noqa: E402


def one_synth(seed, r):
    """perfect mode on the synthetic network: cuts on multiples of 16 (no ring is cut) or ANYWHERE (rings straddle the cuts:
    every band judges its part, the verdict bits are AND-ed over the bands, then the ring is marked)"""
    import synth_ref
    nb16 = int(r.integers(3, 8))
    gh, gw = 16 * nb16 + int(r.integers(0, 16)), int(r.integers(40, 400))
    n_bands = int(r.integers(2, min(4, nb16) + 1))
    if r.random() < 0.5:
        cuts = sorted((16 * r.choice(np.arange(1, nb16), n_bands - 1, replace=False)).tolist())
    else:
        while True:
            cuts = sorted(r.choice(np.arange(2, gh - 2), n_bands - 1, replace=False).tolist())
            if all(b - a >= 2 for a, b in zip([0] + cuts, cuts + [gh])):
                break
    rows = list(zip([0] + cuts, cuts + [gh]))
    arrs, rings = synth_ref.generate(gh, gw, seed=seed, ring_full16=int(r.integers(0, 17)))
    variant = int(r.integers(0, 3))
    bands = [EmulBand(arrs, lo, hi, variant, rings=rings) for lo, hi in rows]
    run = LocalBands(bands)
    og, ch = oracle_grid(arrs), O.Chooser(O.CHOOSE_RANDOM)
    for t in range(int(r.integers(3, 9))):
        og.step_perfect_sparse(ch, rings)
        run.tick(True)
        want = state_of(og)
        got = {k: np.zeros((gh, gw), want[k].dtype) for k in STATE_FIELDS}
        for b in bands:
            b.owned_state(got)
        for k in STATE_FIELDS:
            if not np.array_equal(want[k], got[k]):
                bad = np.argwhere(want[k] != got[k])
                raise AssertionError(f"synth seed {seed} {gh}x{gw} rows {rows} variant {variant} tick {t}: {k} differs at {bad[:5].tolist()}")
    return True


def one(seed):
    r = np.random.default_rng(seed)
    if r.random() < 0.3:
        return one_synth(seed, r)
    gh, gw = int(r.integers(6, 80)), int(r.integers(1, 200))
    if r.random() < 0.15:
        gw = int(r.integers(900, 1400))
    n_bands = int(r.integers(2, 5))
    if gh < 2 * n_bands:
        n_bands = 2
    cuts = sorted(r.choice(np.arange(2, gh - 1), n_bands - 1, replace=False).tolist()) if gh > 4 else [gh // 2]
    rows = list(zip([0] + cuts, cuts + [gh]))
    if any(hi - lo < 2 for lo, hi in rows):
        return False
    arrs = random_grid(seed, gh, gw, n_lines=int(r.integers(1, 3 + (gh + gw) // 2)), car_p=float(r.choice([0.1, 0.4, 0.8, 0.97])),
                       weird_p=float(r.choice([0.0, 0.05, 0.3])), offroad_p=float(r.choice([0.0, 0.02])))
    variant = int(r.integers(0, 3))
    bands = [EmulBand(arrs, lo, hi, variant) for lo, hi in rows]
    run = LocalBands(bands)
    og, ch = oracle_grid(arrs), O.Chooser(O.CHOOSE_RANDOM)
    leaks = 0
    for t in range(int(r.integers(3, 10))):
        nl, _ = og.step_quick(ch)
        leaks += nl
        run.tick(False)
        want = state_of(og)
        got = {k: np.zeros((gh, gw), want[k].dtype) for k in STATE_FIELDS}
        for b in bands:
            b.owned_state(got)
        for k in STATE_FIELDS:
            if not np.array_equal(want[k], got[k]):
                bad = np.argwhere(want[k] != got[k])
                raise AssertionError(f"seed {seed} {gh}x{gw} rows {rows} variant {variant} tick {t}: {k} differs at {bad[:5].tolist()}")
        assert sum(b.leaks_n for b in bands) == leaks, (seed, t)
    return True


if __name__ == "__main__":
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    t0, n, k = time.time(), 0, 0
    while time.time() - t0 < budget:
        k += bool(one(seed + n))
        n += 1
    print(f"band fuzz ok: {k} cases from seed {seed} in {time.time() - t0:.0f} s")
