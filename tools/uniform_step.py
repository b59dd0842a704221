"""Whole step (MRadaptation: ghost update + detail + criteria + keep propagation; update_ghost_mr; upwind FV; swap) on a
UNIFORM level-L mesh with min_level = L-1: the shape where HBM bandwidth bounds the step (BASELINE.json configs[4]).
epsilon < 0 makes every detail significant, so nothing coarsens and the mesh stays uniform while all the MR work is done.
usage: python tools/uniform_step.py [dim] [L] [iters]      (SMR_WF_TRACE=1 prints the per-phase times)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import samurai_b200 as sb

dim = int(sys.argv[1]) if len(sys.argv) > 1 else 2
L = int(sys.argv[2]) if len(sys.argv) > 2 else 13
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
sb.initialize(0)
if os.environ.get("UNFUSED"):
    sb.set_fused(False)
if os.environ.get("UNFUSED") or os.environ.get("PROFILE"):
    sb.profile_enable(True)
r = bench.uniform_full_step(sb, torch, dim, L, iters)
print(r)
if os.environ.get("UNFUSED") or os.environ.get("PROFILE"):
    prof = sb.profile_get()
    for k, v in prof.items():
        if v[0]:
            print(f"  {k:12s} launches {v[0]:4d}  us/launch {1e6 * v[1] / v[0]:9.1f}  cells/launch {v[2] / v[0]:.0f}")
