#!/usr/bin/env python
"""Time line of the production tick (speculative move on three streams) at BASELINE config 4: python tools/timeline.py [size] [mode]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import flowamok_b200 as fa

S = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
perfect = len(sys.argv) > 2 and sys.argv[2] == "perfect"
ctx = fa.Context(0)
g = ctx.grid(S, S)
g.generate_synthetic(seed=1)
g.step(perfect, 10, leaks=False)
ctx.sync()
n = 50
t = g.step_timeline(perfect, n)
print({k: round(v / n * 1e3, 1) for k, v in t.items()}, "us per tick")
p = g.step_profile(perfect, n)
print({k: round(v / n * 1e3, 1) for k, v in p.items()}, "us per tick, kernels one after the other")
