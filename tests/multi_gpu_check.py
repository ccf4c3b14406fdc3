"""Run under torchrun on N GPUs (not collected by pytest):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py
Row bands over NCCL vs the single-grid run of the same synthetic network (rank 0 holds both): bit-equality of the
gathered state after T ticks, in quick and perfect mode (BASELINE config 5's correctness criterion)."""
import hashlib
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import flowamok_b200 as fa  # noqa: E402
from flowamok_b200.bands import BandRunner, DeviceBand, split_rows  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    S, T = int(os.environ.get("CHECK_SIZE", "4096")), int(os.environ.get("CHECK_TICKS", "60"))
    stream = torch.cuda.Stream()
    ok = True
    with torch.cuda.stream(stream):
        ctx = fa.Context(lr, stream.cuda_stream)
        for perfect in (False, True):
            lo, hi = split_rows(S, world, 16)[rank]
            band = DeviceBand(ctx, S, S, lo, hi)
            band.grid.generate_synthetic(seed=5)
            if os.environ.get("CHECK_NATIVE", "1") == "1":
                band.comm_init(rank, world)          # ticks driven inside the library over its own NCCL communicator
            run = BandRunner(band, rank, world)
            run.step(perfect, T)
            d = band.grid.download(fields=("dy", "dx", "color", "pref", "rng_state"))
            h = hashlib.sha256()
            for k in ("dy", "dx", "color", "pref", "rng_state"):
                h.update(d[k][lo:hi].tobytes())
            leaks = torch.tensor([band.leaks()], device="cuda", dtype=torch.int64)
            dist.all_reduce(leaks)
            digests = [None] * world
            dist.all_gather_object(digests, (lo, hi, h.hexdigest()))
            if rank == 0:
                g = ctx.grid(S, S)
                g.generate_synthetic(seed=5)
                l1 = g.step(perfect, T)
                f = g.download(fields=("dy", "dx", "color", "pref", "rng_state"))
                for lo_r, hi_r, dig in digests:
                    h = hashlib.sha256()
                    for k in ("dy", "dx", "color", "pref", "rng_state"):
                        h.update(f[k][lo_r:hi_r].tobytes())
                    same = h.hexdigest() == dig
                    ok &= same
                    print(f"perfect={perfect} rows {lo_r}:{hi_r} {'identical' if same else 'DIFFERENT'}")
                ok &= int(leaks.item()) == l1
                print(f"perfect={perfect}: leaks {int(leaks.item())} vs {l1}; convergence rounds/tick {run.rounds / T:.2f}")
                g.free()
            band.grid.free()
    if rank == 0:
        print("MULTI_GPU_CHECK", "OK" if ok else "FAILED", f"({world} GPUs, {S}x{S}, {T} ticks)")
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
