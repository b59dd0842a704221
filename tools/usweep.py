"""bench.py's uniform sweep (configs[4]) alone: FV strip kernel and flux-based diffusion on a uniform 2D level-L mesh."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import samurai_b200 as sb
sb.initialize(0)
for L in [int(x) for x in (sys.argv[1:] or ["13"])]:
    n, s, g, d = bench.uniform_sweep(sb, torch, L)
    print(f"level {L}: fv {16*n/s/1e9:.0f} GB/s ({1e3*s:.3f} ms)   diffusion {24*n/d/1e9:.0f} GB/s ({1e3*d:.3f} ms)")
