"""BASELINE config 2 against the reference's own source text: every explorer scenario, perfect simulation, stepped exactly as
explorer/explorer.fut:109-130 does -- EXECUTED from explorer.fut by tests/golden/futhark_mini.py for the first 200 ticks
(tests/golden/make_explorer_runs.py).  Per tick: the digest of the whole grid (rng states included), the cell leaks, the
explorer's rng and the number of cycle_checks; the full grid at ticks 1, 10, 50, 100, 150, 200.  The oracle on the CPU, the
CUDA path on the GPU box (where the 1000-tick runs of tests/test_gpu_parity.py continue from there against the oracle)."""
import numpy as np
import pytest

from util import O, ROOT, STATE_FIELDS, state_digest, state_of

R = np.load(ROOT / "tests" / "golden" / "explorer_runs.npz")


def rows(sid):
    sid_, n_ticks, seed, n_cycles0 = (int(v) for v in R[f"{sid}/meta"])
    assert sid_ == sid
    return n_ticks, seed, n_cycles0, R[f"{sid}/ticks"]


@pytest.mark.parametrize("sid", range(14))
def test_oracle_explorer_run_equals_the_reference_source(sid):
    n_ticks, seed, n_cycles0, ticks = rows(sid)
    ox = O.Explorer(seed, sid, perfect=True)
    assert len(ox.cycles) == n_cycles0
    for t in range(1, n_ticks + 1):
        leaks = ox.step(1)
        st = state_of(ox.grid)
        d = state_digest(st)
        want = ticks[t - 1]
        got = [np.uint64(d[0]).astype(np.int64), np.uint64(d[1]).astype(np.int64), leaks, np.uint64(ox.rng.state).astype(np.int64),
               len(ox.cycles), ox.step_cur]
        assert [int(v) for v in got] == [int(v) for v in want], f"scenario {sid} tick {t}: digest / leaks / rng / cycles / steps"
        if f"{sid}/t{t}/dy" in R:
            a = ox.grid.arrays()
            for f in ("uy", "ux", "dy", "dx", "color", "pref"):
                assert np.array_equal(a[f], R[f"{sid}/t{t}/{f}"]), f"scenario {sid} tick {t}: {f}"


@pytest.mark.gpu
@pytest.mark.parametrize("sid", range(14))
def test_cuda_explorer_run_equals_the_reference_source(ctx, sid):
    from flowamok_b200 import scenarios as S
    n_ticks, seed, n_cycles0, ticks = rows(sid)
    dx = S.Explorer(ctx, seed, sid, perfect=True)
    assert dx.n_cycles == n_cycles0
    for t in range(1, n_ticks + 1):
        leaks = dx.step(1)
        d = dx.grid.digest()
        want = ticks[t - 1]
        got = [np.uint64(d[0]).astype(np.int64), np.uint64(d[1]).astype(np.int64), leaks, np.uint64(dx.rng.state).astype(np.int64),
               dx.n_cycles, dx.step_cur]
        assert [int(v) for v in got] == [int(v) for v in want], f"scenario {sid} tick {t}: digest / leaks / rng / cycles / steps"
        if f"{sid}/t{t}/dy" in R:
            a = dx.grid.download()
            for f in ("uy", "ux", "dy", "dx", "color", "pref"):
                assert np.array_equal(a[f], R[f"{sid}/t{t}/{f}"]), f"scenario {sid} tick {t}: {f}"
    dx.grid.free()
