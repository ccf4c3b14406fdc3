#!/usr/bin/env python
"""Kernel-tuning helper: build libflowamok_b200 variants (different -D settings) here, bench them on the GPU box.

  python tools/tune_variants.py build  NAME=-DFA_M_MINB=4,-DFA_U_MINB=5 ...   (no GPU needed)
  python tools/tune_variants.py bench [--steps K] [ENV=VAL ...]               (on the GPU box; one line per variant)
"""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
VDIR = ROOT / "flowamok_b200" / "_variants"
sys.path.insert(0, str(ROOT))


def build(specs):
    from flowamok_b200 import _build
    VDIR.mkdir(exist_ok=True)
    for spec in specs:
        name, _, flags = spec.partition("=")
        out = VDIR / f"lib_{name}.so"
        cmd = [_build.nvcc_path(), *_build.NVCC_FLAGS, "-Xptxas=-v", *[f for f in flags.split(",") if f], "-o", str(out),
               *map(str, _build.SOURCES)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            print(name, "FAILED", r.stderr[-2000:])
            continue
        regs = {}
        cur = None
        for line in r.stderr.splitlines():
            if "Compiling entry function" in line:
                cur = line.split("'")[1]
            elif "Used" in line and cur:
                for k in ("k_u_localILb0", "k_moveILi4", "k_moveILi6", "k_moveILi8", "k_u_iter"):
                    if k in cur:
                        regs[k] = line.split("Used")[1].split(",")[0].strip().split()[0]
            elif "spill" in line and cur and any(k in cur for k in ("k_u_local", "k_move")) and "0 bytes spill stores" not in line:
                regs[cur[:20] + "_spill"] = line.strip()
        print(name, regs)


def bench(args):
    steps = "60"
    env_extra = {}
    for a in args:
        if a.startswith("--steps="):
            steps = a.split("=")[1]
        elif "=" in a:
            k, v = a.split("=", 1)
            env_extra[k] = v
    for so in sorted(VDIR.glob("lib_*.so")):
        env = dict(os.environ, FLOWAMOK_B200_LIB=str(so), **env_extra)
        r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", steps, "--warmup", "5", "--no-cpu", "--no-e2e"],
                           capture_output=True, text=True, env=env)
        try:
            j = json.loads(r.stdout.strip().splitlines()[-1])
            k = j["roofline"]["kernel_ms_per_tick"]
            print(f"{so.name:28s} {env_extra} GCUPS {j['value']:7.1f} ms {j['ms_per_step']:.4f} "
                  + " ".join(f"{a}={b:.4f}" for a, b in k.items())
                  + f" iters {j['fixpoint']['iterations_per_tick']:.1f} visits {j['fixpoint']['worklist_entries_relaxed_per_tick']:.0f}", flush=True)
        except Exception as e:
            print(so.name, "FAILED", e, r.stderr[-800:], flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    else:
        bench(sys.argv[2:])
