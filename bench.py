#!/usr/bin/env python
"""bench.py -- flowamok per-tick grid update on B200 (BASELINE.json metric: cell-updates/sec, GCUPS).

  python bench.py --gpus N --steps K --warmup W            our CUDA path (one process per GPU)
  python bench.py --impl reference --gpus N --steps K ...   the reference's CPU algorithm (oracle port,
                                                            all host threads, bounded sample), rank 0 only

A "step" is one tick (stepper_quick.step, steppers.fut:177-185) over the whole grid.  At N=1 the workload is
BASELINE config 4: the synthetic road network at 16384 x 16384, seed 1, 10 % of road cells
occupied, quick mode.  Prints ONE JSON line (see the keys below).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

B_ALG = 12.0  # algorithmic bytes per cell-update (SURVEY 8d / BASELINE.md): r+w of a u16 control word and a u32 colour

RNG_PARITY = ("unverified: the fork stream of the reference's README GIF is not reproduced (diku-dk/cpprandom 1.3.1 is not "
              "vendored; DESIGN.md section 4); CUDA == oracle bit for bit; everything but the draw arithmetic is pinned by the GIF and by "
              "vectors made by executing the reference's own steppers.fut (tests/golden/futhark_mini.py)")


def csrc_hash():
    """sha256 over the kernel sources: the key under which profiles/*_traffic.json records ncu's DRAM bytes"""
    import hashlib
    h = hashlib.sha256()
    for f in sorted((ROOT / "flowamok_b200" / "csrc").iterdir()):
        if f.suffix in (".h", ".cuh", ".cu", ".inc"):
            h.update(f.name.encode())
            h.update(f.read_bytes())
    return h.hexdigest()[:16]


def ncu_traffic(workload):
    """DRAM bytes per launch and kernel (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture),
    written by profiles/summarize.py --traffic.  None unless it was captured on THESE kernel sources for THIS workload:
    a number from other code is not reported."""
    best = None
    for f in sorted((ROOT / "profiles").glob("*_traffic.json")):
        try:
            d = json.loads(f.read_text())
        except Exception:
            continue
        if d.get("csrc_sha256") == csrc_hash() and d.get("workload") == workload:
            best = d
    return best


def bind_near_gpu(index):
    """Best effort: run (and first-touch the pinned host buffers) on the CPUs NVML calls ideal for this GPU, so that the
    host <-> device copies of the e2e leg do not cross sockets.  Returns the previous affinity (restored before the CPU legs)."""
    try:
        old = os.sched_getaffinity(0)
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        ideal = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        use = ideal & old
        if use and use != old:
            os.sched_setaffinity(0, use)
        return old
    except Exception:
        return None


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (profiling recipe's clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        """t0, t1: wall-clock bounds of the timed region; the sampler itself is started earlier (nvidia-smi needs a moment to
        come up), and only the rows sampled inside the region -- or, for a very short region, the first one after it began --
        are used."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [r for _t, r in self.rows]
        if t0 is not None:
            inside = [r for t, r in self.rows if t0 <= t <= t1 + 0.12]
            after = [r for t, r in self.rows if t >= t0]
            rows = inside or after[:1] or rows[-1:]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_run(size_w, rows, ticks, warm, threads=None, perfect=False):
    """The oracle port (reference layout: record of arrays, one full-grid Jacobi pass per fixpoint iteration,
    count-based termination exactly as steppers.fut:160-163) on a band of the same synthetic network."""
    from oracle import oracle as O
    # all host cores, explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers
    O.set_threads(threads or len(os.sched_getaffinity(0)) or os.cpu_count() or 1)
    g = O.Grid(rows, size_w)
    rings = g.generate_synthetic(seed=1)
    ch = O.Chooser(O.CHOOSE_RANDOM)

    def tick():
        if perfect:
            g.step_perfect_sparse(ch, rings)
        else:
            g.step_quick(ch)
    for _ in range(warm):
        tick()
    t0 = time.perf_counter()
    for _ in range(ticks):
        tick()
    dt = time.perf_counter() - t0
    return rows * size_w * ticks / dt / 1e9, dt, O.max_threads()


def run_reference(args, rank, emit):
    if rank != 0:
        return 0
    size = args.size
    # calibrate on a thin band, then size the band so that (steps + warmup) ticks take about a minute
    perfect = args.mode == "perfect"
    g1, dt1, thr = cpu_oracle_run(size, 64, 2, 1, perfect=perfect)
    per_cell_tick = dt1 / (64 * size * 2)
    budget = 60.0
    rows = int(budget / (per_cell_tick * size * max(1, args.steps + args.warmup)))
    rows = max(32, min(size, (rows // 16) * 16))
    gcups, dt, thr = cpu_oracle_run(size, rows, args.steps, args.warmup, perfect=perfect)
    sample = f"{rows}x{size} band of the synthetic-{size} network, {args.steps} ticks, {args.mode} mode"
    line = {"impl": "reference", "metric": "cell-updates/sec", "value": gcups, "unit": "GCUPS", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"synthetic-{size}x{size}-{args.mode} (SURVEY App. E, seed 1, 10% of road cells occupied)",
                       "grid": [size, size], "mode": args.mode, "sample": sample},
            "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": thr, "kind": "port", "sample": sample,
                             "nproc": os.cpu_count()},
            "rng_parity": RNG_PARITY,
            "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference is pure Futhark; no Futhark compiler in the image, so this is the C/OpenMP oracle port "
                    "in the reference's layout and schedule (BASELINE.md plan B)"}
    emit(line)
    return 0


def main():
    # ONE JSON line on stdout: everything else that libraries print there (NCCL's version banner, ...) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=0, help="grid edge; default 16384 on 1 GPU, 65536 on several")
    ap.add_argument("--mode", default="", choices=["", "quick", "perfect"])
    ap.add_argument("--transport", default="p2p", choices=["p2p", "nccl"],
                    help="row bands: peer-mapped mailboxes + device-side flags (default) or NCCL send/recv + host round trip")
    ap.add_argument("--cuts", default="aligned", choices=["aligned", "straddle"],
                    help="row bands: boundaries on multiples of 16, or shifted by 7 rows (rings straddle the cuts)")
    ap.add_argument("--density16", type=int, default=6554, help="car probability per road cell in 1/65536 (default: 10 %%)")
    ap.add_argument("--no-one-gpu", action="store_true", help="several GPUs: skip timing the same grid on rank 0 alone")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # BASELINE.json: config 4 (16384^2, quick) is the 1-GPU headline; config 5 (65536^2 row bands, perfect) the multi-GPU one
    if not args.size:
        args.size = 16384 if world == 1 else 65536
    if not args.mode:
        args.mode = "quick" if world == 1 else "perfect"

    if args.impl == "reference":
        return run_reference(args, rank, emit)

    import ctypes as C
    import numpy as np
    import torch
    import flowamok_b200 as fa
    from flowamok_b200.bands import BandRunner, DeviceBand, split_rows

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (flowamok_b200 has no CPU path)")
    torch.cuda.set_device(local_rank)
    old_affinity = bind_near_gpu(local_rank)
    sampler = ClockSampler(local_rank)   # started NOW: nvidia-smi takes a few hundred ms to deliver its first row; only the rows
    sampler.start()                      # sampled during the timed region are used (ClockSampler.stop)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.Stream()
    perfect = args.mode == "perfect"
    S = args.size
    cells = S * S
    one_gpu = None
    with torch.cuda.stream(stream):
        ctx = fa.Context(local_rank, stream.cuda_stream)
        if world > 1 and rank == 0 and not args.no_one_gpu:
            # the SAME grid on one GPU (efficiency = T(1 GPU) / (N * T(N GPUs)) on 65536^2; 82 GB fit)
            g1 = ctx.grid(S, S)
            g1.generate_synthetic(seed=1, density16=args.density16, ring_full16=2)
            k1 = max(5, min(args.steps, 50))
            g1.step(perfect, args.warmup, leaks=False)
            ctx.sync()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            g1.step(perfect, k1, leaks=False)
            a1.record(stream)
            ctx.sync()
            ms1 = a0.elapsed_time(a1) / k1
            one_gpu = {"value": cells / (ms1 * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": ms1, "steps": k1,
                       "what": f"the same {S}x{S} {args.mode} grid stepped on rank 0's GPU alone, before the banded run"}
            g1.free()
            del g1
        if world == 1:
            g = ctx.grid(S, S)
            rows = (0, S)
            runner = None
            cuts = [0, S]
        else:
            cuts = [r[0] for r in split_rows(S, world, 16)] + [S]
            if args.cuts == "straddle":
                cuts = [0] + [c + 7 for c in cuts[1:-1]] + [S]
            rows = (cuts[rank], cuts[rank + 1])
            band = DeviceBand(ctx, S, S, rows[0], rows[1])
            g = band.grid
            band.set_cuts(cuts)
        g.generate_synthetic(seed=1, density16=args.density16, ring_full16=2)
        if world > 1:
            if args.transport == "p2p":
                band.p2p_init(rank, world)                # peer-mapped mailboxes: ticks run natively, no NCCL, no host sync
            else:
                band.comm_init(rank, world)               # the library's own NCCL communicator
            runner = BandRunner(band, rank, world)
        n_rings, _ = g.cycle_count()

        def run_ticks(n, leaks=False):
            if runner is None:
                return g.step(perfect, n, leaks=leaks)
            runner.step(perfect, n)
            return band.leaks() if leaks else None

        def barrier():
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()

        def maxr(x):
            if dist is None:
                return x
            t = torch.tensor([x], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        run_ticks(args.warmup)
        ctx.sync()
        g.stats()
        rounds0 = runner.rounds if runner else 0
        use_p2p = runner is not None and args.transport == "p2p"
        p2p0 = band.p2p_stats() if use_p2p else None

        # ---- device-resident throughput: K ticks, inputs already in HBM, CUDA events on the launching stream
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        wall0 = time.time()
        e0.record(stream)
        run_ticks(args.steps)
        e1.record(stream)
        barrier()
        wall1 = time.time()
        ms = maxr(e0.elapsed_time(e1))
        clocks = sampler.stop(wall0, wall1)
        st = g.stats()
        rounds = (runner.rounds - rounds0) if runner else 0
        p2p1 = band.p2p_stats() if use_p2p else None
        value = cells * args.steps / (ms * 1e-3) / 1e9

        # ---- per-kernel device time (events between the launches, same stream); single-grid path only
        prof, prof_ticks = None, min(50, args.steps)
        if runner is None:
            prof = g.step_profile(perfect, prof_ticks)
        elif use_p2p:
            prof = band.step_profile(perfect, prof_ticks)      # this rank's band; every band steps the same ticks

        # ---- end to end through the C ABI with HOST buffers: upload + K ticks + leak count + download
        e2e = None
        if not args.no_e2e:
            # one grid: the whole value (roads included); row bands: the car state of the owned rows (roads, cycles
            # and halo rows stay resident: flowamok_grid_upload_cars)
            # The per-cell RNG states travel sparsely: only cells with a road on both axes can ever draw (steppers.fut:101-103),
            # so the host keeps (flat index, state) of those "corner" cells -- 16 bytes per corner instead of 8 per cell.
            fields_up = ("uy", "ux", "dy", "dx", "color", "pref") if world == 1 else ("dy", "dx", "color", "pref")
            fields_dn = ("dy", "dx", "color", "pref")
            host = g.download(fields=tuple(dict.fromkeys(fields_up + fields_dn)))   # whole-grid shaped; only my rows are filled
            lo, hi = rows
            # Host buffers are page-locked.  One grid: torch's pinned allocator.  Row bands: the rows this rank owns are
            # registered in place (cudaHostRegister: no second copy of 30 GB of car state) when the node's memory allows
            # it -- locked pages of all ranks together stay below 60 % of the memory this job may use; pageable otherwise.
            pinned = world == 1
            registered = []
            cudart = torch.cuda.cudart()
            def node_memory_bytes():
                tot = 0
                for line in open("/proc/meminfo"):
                    if line.startswith("MemTotal:"):
                        tot = int(line.split()[1]) * 1024
                try:
                    lim = open("/sys/fs/cgroup/memory.max").read().strip()
                    if lim != "max":
                        tot = min(tot, int(lim))
                except OSError:
                    pass
                return tot
            lock_ok = world > 1 and (hi - lo) * S * 7 * world <= 0.6 * node_memory_bytes()
            def hostbuf(a):
                a = np.ascontiguousarray(a)
                if pinned:
                    return torch.from_numpy(a).pin_memory().numpy()
                if lock_ok and a.nbytes and int(cudart.cudaHostRegister(a.ctypes.data, a.nbytes, 0)) == 0:
                    registered.append(a)
                return a
            up = {k: hostbuf(host[k][lo:hi]) for k in fields_up}
            n_bufs = len(up) + 2
            del host
            dn = {k: up[k] for k in fields_dn}               # the result lands in the buffers the input came from
            ci, cs = g.rng_corners()
            ci, cs = hostbuf(ci), hostbuf(cs)
            h2d = sum(up[k].nbytes for k in fields_dn) + ci.nbytes + cs.nbytes      # per timed pass: the car state
            d2h = sum(v.nbytes for v in dn.values()) + ci.nbytes + cs.nbytes + 8

            def band_ptr(a):                                 # the ABI takes whole-grid base pointers
                return a.ctypes.data - lo * a.strides[0]
            legs = {}
            def e2e_pass(n_ticks, first=False):              # every call below is synchronous: the legs add up to the pass
                ta = time.perf_counter()
                if world == 1 and first:                     # the whole value once: roads are static within a run
                    ctx.check(ctx.L.flowamok_grid_upload(ctx.h, g.h, *[band_ptr(up[k]) for k in fields_up], None, fa.PCG_INC))
                elif world == 1:                             # afterwards what a tick can change: the car state
                    ctx.check(ctx.L.flowamok_grid_upload_cars(ctx.h, g.h, *[band_ptr(up[k]) for k in fields_dn], None))
                else:
                    band.upload_cars(*[band_ptr(up[k]) for k in fields_up], None)
                g.rng_scatter(ci, cs)
                tb = time.perf_counter()
                n_leaks = run_ticks(n_ticks, leaks=True)
                tc = time.perf_counter()
                inc = C.c_uint64()
                ctx.check(ctx.L.flowamok_grid_download(ctx.h, g.h, None, None, *[band_ptr(dn[k]) for k in fields_dn], None, C.byref(inc)))
                g.rng_corners(ci, cs)
                td = time.perf_counter()
                legs.update(upload_s=tb - ta, ticks_s=tc - tb, download_s=td - tc)
                return n_leaks

            def pcie_probe():                                # what a plain pinned copy of 1 GiB gets on this box, each way
                n = 1 << 30
                hbuf, dbuf = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8, device="cuda")
                out = {}
                for name, (dst, src) in (("h2d_gbs", (dbuf, hbuf)), ("d2h_gbs", (hbuf, dbuf))):
                    dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
                    t = time.perf_counter()
                    dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
                    out[name] = n / (time.perf_counter() - t) / 1e9
                return out

            my_h2d, my_d2h = h2d, d2h
            probe = pcie_probe() if rank == 0 else None
            e2e_pass(1, first=True)                          # untimed: the whole grid value goes up once (roads included); staging buffers, first touch
            barrier()
            t0 = time.perf_counter()
            leaks = e2e_pass(args.steps)
            barrier()
            dt = maxr(time.perf_counter() - t0)
            locked = 1.0 if pinned or len(registered) == n_bufs else 0.0
            if dist is not None:
                t = torch.tensor([h2d, d2h, int(leaks), locked], device="cuda", dtype=torch.float64)
                dist.all_reduce(t)
                h2d, d2h, leaks, locked = (float(x) for x in t.tolist())
            host_memory = "pinned" if locked == world else "pageable" if locked == 0 else f"pinned on {int(locked)} of {world} ranks"
            torch.cuda.synchronize()
            for a in registered:
                cudart.cudaHostUnregister(a.ctypes.data)
            e2e = {"value": cells * args.steps / dt / 1e9, "unit": "GCUPS",
                   "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
                   "seconds": dt, "leaks": int(leaks), "host_memory": host_memory,
                   "rng_corner_cells": int(len(ci)),
                   # rank 0's legs of the timed pass, and the box's own pinned-copy bandwidth (one rank copying alone)
                   "legs": {**legs, "upload_gbs": my_h2d / legs["upload_s"] / 1e9, "download_gbs": my_d2h / legs["download_s"] / 1e9},
                   "pcie_probe": probe,
                   "what": ("flowamok_grid_upload_cars (host record-of-arrays: headings, colours, preferences; the roads, static within a "
                            "run, went up once with the untimed flowamok_grid_upload)" if world == 1 else
                            "flowamok_grid_upload_cars (car state of every band's rows)") +
                           " + flowamok_grid_rng_scatter (states of the corner cells, the only ones that can draw) + steps ticks through "
                           "the C ABI + leak count + flowamok_grid_download + flowamok_grid_rng_corners, wall clock, max over ranks"}
            del up, dn, ci, cs

    peak, peak_src = measured_peak()
    per_gpu_gcups = value / world
    achieved = per_gpu_gcups * B_ALG  # GB/s of algorithmic traffic per GPU
    kernels = "k_u_local + k_u_iter" + (" + k_rings" if perfect else "") + " + k_move"
    wl = f"synthetic-{S}x{S}-{args.mode}" + ("" if args.density16 == 6554 else f"-density16={args.density16}")
    tr = ncu_traffic(wl) if world == 1 else None            # ncu's DRAM bytes per launch, only if captured on these sources
    ms_tick = ms / args.steps
    traffic = sum(tr["dram_bytes"].values()) if tr else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes": cells / world * B_ALG,
                "what": f"whole tick ({kernels}): ALGORITHMIC bytes (12 B per cell-update, SURVEY 8d) / tick time, per GPU.  "
                        "The tick moves fewer DRAM bytes than that (lazy colour buffer, bit planes): `physical` is what "
                        "the DRAM actually did",
                "physical": ({"dram_bytes": traffic, "gbs": traffic / (ms_tick * 1e-3) / 1e9,
                              "frac": traffic / (ms_tick * 1e-3) / 1e9 / peak,
                              "source": tr["source"], "csrc_sha256": tr["csrc_sha256"]} if tr else None)}
    if prof:
        tot = sum(prof.values())
        dom = max(prof, key=prof.get)
        dom_ms = prof[dom] / prof_ticks
        dom_bytes = tr["dram_bytes"].get(dom) if tr else None
        roofline.update({"kernel_ms_per_tick": {k: v / prof_ticks for k, v in prof.items()},
                         "dominant_kernel": dom,
                         "dominant_share": max(prof.values()) / tot if tot else None,
                         # the dominant kernel against the DRAM bytes ncu measured for it (never above 1)
                         "dominant": {"kernel": dom, "ms": dom_ms, "traffic": dom_bytes,
                                      "dram_gbs": dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_bytes else None,
                                      "frac": dom_bytes / (dom_ms * 1e-3) / 1e9 / peak if dom_bytes else None}})
    # launches of OUR kernels inside the timed region: k_u_local, [k_rings], k_u_iter || k_move, k_move_fix per tick;
    # row bands over P2P add k_band_xstate and k_band_xu (+ k_band_xring and k_rings_mark when rings straddle the cuts)
    launches_per_tick = 4 + (1 if perfect and n_rings else 0)
    if runner is not None and args.transport == "p2p":
        launches_per_tick += 2 + (2 if perfect and args.cuts == "straddle" else 0)
    gpu_launches = launches_per_tick * args.steps
    if runner is not None and args.transport == "nccl":   # kernels one after the other + pack / merge / resume per swap and round
        gpu_launches = (3 + (1 if perfect and n_rings else 0) + 4) * args.steps + rounds * 3
    hist = g.stats_hist()

    def jam_ticks():
        n = C.c_int64()
        ctx.check(ctx.L.flowamok_jam_ticks(ctx.h, g.h, C.byref(n)))
        return n.value
    if runner is None:
        par = "1 grid"
    elif args.transport == "p2p":
        par = (f"{world} row bands (cuts {args.cuts}: {cuts[1:-1]}), per-tick halo rows by direct stores into the neighbours' "
               "peer-mapped mailboxes over NVLink + device-side flags; convergence rounds and termination vote inside k_u_iter; "
               "no NCCL call, no host synchronisation per tick")
    else:
        par = f"{world} row bands, per-tick halo rows over NCCL send/recv + 1 all-reduced convergence word (host round trip)"
    line = {"metric": "cell-updates/sec", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak" if world == 1 else "strong",
            "vs_baseline": None, "dtype": "u32 bit-planes (32 cells per word) + u32 colour payload",
            "data": "synthetic",
            "config": {"workload": f"{wl} (SURVEY App. E, seed 1, {args.density16 / 655.36:.0f}% of road cells occupied)",
                       "grid": [S, S], "mode": args.mode, "rings_on_rank0": int(n_rings),
                       "parallelism": par,
                       "l2": f"planes read+written per tick {cells / world * 7.7 / 1e9:.2f} GB per GPU >> 126 MB L2: inputs larger than L2, no flush"},
            "fixpoint": {"iterations_per_tick": st["passes"] / max(1, st["ticks"]),
                         "worklist_entries_relaxed_per_tick": st["sweeps"] / max(1, st["ticks"]),
                         "iterations_histogram": {"what": "ticks by number of frontier iterations of k_u_iter on rank 0 "
                                                          "(all ticks since the grid was made; bucket 31 = 31 and more)",
                                                  "ticks": {str(k): v for k, v in enumerate(hist) if v}},
                         "ticks_in_jam_order": jam_ticks(),
                         "band_convergence_rounds_per_tick": rounds / args.steps if runner else None,
                         "band_wait_ms_per_tick_rank0": ((p2p1["wait_ns"] - p2p0["wait_ns"]) / 1e6 / args.steps) if p2p1 else None},
            "roofline": roofline, "clocks": clocks, "gpu_launches": int(gpu_launches), "e2e": e2e,
            "rng_parity": RNG_PARITY, "nproc": os.cpu_count()}
    if one_gpu:
        line["one_gpu_same_workload"] = one_gpu
        line["efficiency_same_workload"] = value / (world * one_gpu["value"])
    if old_affinity:
        try:
            os.sched_setaffinity(0, old_affinity)     # the CPU baseline uses every core
        except Exception:
            pass
    if rank == 0 and not args.no_cpu:
        try:
            g1, dt1, thr = cpu_oracle_run(S, 64, 2, 1, perfect=perfect)
            rows_c = int(30.0 / (dt1 / (64 * S * 2) * S * 6))
            rows_c = max(32, min(S, (rows_c // 16) * 16))
            gc, dt, thr = cpu_oracle_run(S, rows_c, 5, 1, perfect=perfect)
            line["cpu_baseline"] = {"value": gc, "unit": "GCUPS", "cores": thr, "kind": "port", "nproc": os.cpu_count(),
                                    "sample": f"{rows_c}x{S} band of the same network, 5 ticks, {args.mode} mode, {dt:.1f} s"}
        except Exception as e:  # the oracle is only a reported baseline
            line["cpu_baseline"] = {"value": None, "unit": "GCUPS", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
    if rank == 0:
        emit(line)
    if dist is not None:
        if use_p2p:
            band.p2p_close()      # unmap the peers' mailboxes, then make sure everybody has before anything is freed
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
