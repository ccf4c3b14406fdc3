"""Row bands on N GPUs, one process per GPU (run under torchrun; tests/test_multi_gpu.py spawns it from pytest -m gpu):

   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py

BASELINE config 5's correctness criterion, for every cut pattern in CHECK_CUTS:
  (a) the synthetic CHECK_SIZE^2 grid (default 65536) stepped CHECK_TICKS ticks in perfect and quick mode as N row bands over
      the P2P transport == the single-grid run of the same network on rank 0's GPU, by device digest (the planes never
      leave the GPUs: 65536^2 does not fit the host) and leak count;
  (b) a CHECK_ORACLE_SIZE^2 instance (default 2048) with the same cut pattern == the CPU oracle, whole state.
Cut patterns: "aligned" = band boundaries on multiples of 16 (no ring is cut), "straddle" = boundaries shifted by 7 rows
(every ring tile row next to a cut straddles it: per-ring verdict bits are AND-ed over the bands).
CHECK_TRANSPORT=nccl runs the NCCL send/recv transport instead (aligned cuts only)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import flowamok_b200 as fa  # noqa: E402
from flowamok_b200.bands import BandRunner, DeviceBand, split_rows  # noqa: E402

M64 = (1 << 64) - 1


def cuts_for(S, world, pattern):
    rows = split_rows(S, world, 16)
    cuts = [r[0] for r in rows] + [S]
    if pattern == "straddle":
        cuts = [0] + [c + 7 for c in cuts[1:-1]] + [S]
    return cuts


def make_band(ctx, S, cuts, rank, world, seed, transport, ring_full16):
    band = DeviceBand(ctx, S, S, cuts[rank], cuts[rank + 1])
    band.set_cuts(cuts)
    band.grid.generate_synthetic(seed=seed, ring_full16=ring_full16)
    if transport == "p2p":
        band.p2p_init(rank, world)
    else:
        band.comm_init(rank, world)
    return band


def close_band(band, transport):
    if transport == "p2p":
        band.p2p_close()
    dist.barrier()          # nobody maps this band's mailbox any more
    band.grid.free()


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    S, T = int(os.environ.get("CHECK_SIZE", "65536")), int(os.environ.get("CHECK_TICKS", "50"))
    S2, T2 = int(os.environ.get("CHECK_ORACLE_SIZE", "2048")), int(os.environ.get("CHECK_ORACLE_TICKS", "40"))
    transport = os.environ.get("CHECK_TRANSPORT", "p2p")
    patterns = os.environ.get("CHECK_CUTS", "aligned,straddle" if transport == "p2p" else "aligned").split(",")
    stream = torch.cuda.Stream()
    ok = True
    with torch.cuda.stream(stream):
        ctx = fa.Context(lr, stream.cuda_stream)
        for pattern in patterns:
            for perfect in (True, False):
                # ---- (a) full size: bands vs the single grid, by digest
                cuts = cuts_for(S, world, pattern)
                band = make_band(ctx, S, cuts, rank, world, 5, transport, 4)
                run = BandRunner(band, rank, world)
                dist.barrier()
                run.step(perfect, T)
                s, x = band.grid.digest()
                t = torch.tensor([s - (1 << 64) if s >= (1 << 63) else s, band.leaks()], device="cuda", dtype=torch.int64)
                dist.all_reduce(t)                         # int64 addition wraps: the sum of the bands' digests mod 2^64
                xs = [None] * world
                dist.all_gather_object(xs, x)
                rounds = run.rounds
                close_band(band, transport)
                if rank == 0:
                    g = ctx.grid(S, S)
                    g.generate_synthetic(seed=5, ring_full16=4)
                    l1 = g.step(perfect, T)
                    s1, x1 = g.digest()
                    g.free()
                    xa = 0
                    for v in xs:
                        xa ^= v
                    same = (int(t[0].item()) & M64) == s1 and xa == x1 and int(t[1].item()) == l1
                    ok &= same
                    print(f"[{pattern:8s} perfect={int(perfect)}] {S}x{S}, {T} ticks, {world} bands over {transport}: digest "
                          f"{'identical to' if same else 'DIFFERENT from'} the single-grid run; leaks {int(t[1].item())} vs {l1}; "
                          f"convergence rounds/tick {rounds / T:.2f}", flush=True)
                # ---- (b) down-scaled instance with the same cut pattern vs the CPU oracle
                cuts2 = cuts_for(S2, world, pattern)
                band = make_band(ctx, S2, cuts2, rank, world, 6, transport, 4)
                run = BandRunner(band, rank, world)
                dist.barrier()
                run.step(perfect, T2)
                d = band.grid.download(fields=("dy", "dx", "color", "pref", "rng_state"))
                lo, hi = cuts2[rank], cuts2[rank + 1]
                mine = {k: d[k][lo:hi].copy() for k in ("dy", "dx", "color", "pref", "rng_state")}
                parts = [None] * world
                dist.all_gather_object(parts, (lo, hi, mine, band.leaks()))
                close_band(band, transport)
                if rank == 0:
                    from oracle import oracle as O
                    og = O.Grid(S2, S2)
                    rings = og.generate_synthetic(seed=6, ring_full16=4)
                    ch = O.Chooser(O.CHOOSE_RANDOM)
                    leaks = 0
                    for _ in range(T2):
                        leaks += (og.step_perfect_sparse(ch, rings) if perfect else og.step_quick(ch))[0]
                    a = og.arrays()
                    same = sum(p[3] for p in parts) == leaks
                    for lo_r, hi_r, st, _l in parts:
                        for k in ("dy", "dx", "color", "pref", "rng_state"):
                            same &= bool(np.array_equal(st[k], a[k][lo_r:hi_r]))
                    ok &= same
                    print(f"[{pattern:8s} perfect={int(perfect)}] {S2}x{S2}, {T2} ticks, cuts {cuts2[1:-1]}: "
                          f"{'identical to' if same else 'DIFFERENT from'} the CPU oracle (whole state + {leaks} leaks)", flush=True)
    if rank == 0:
        print("MULTI_GPU_CHECK", "OK" if ok else "FAILED", f"({world} GPUs, transport {transport}, cuts {patterns})", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    return 0 if int(flag.item()) else 1


if __name__ == "__main__":
    sys.exit(main())
