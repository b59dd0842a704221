"""Dump the leaves of the adapted advection_2d mesh (bench workload) for host-side profiling on the CPU box."""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, samurai_b200 as sb
import argparse
ap = argparse.ArgumentParser(); ap.add_argument("--max-level", type=int, default=14); ap.add_argument("--steps", type=int, default=3); ap.add_argument("--dim", type=int, default=2)
a = ap.parse_args()
class A: pass
args = A(); args.min_level = 4; args.max_level = a.max_level; args.eps = 2e-4; args.dim = a.dim
sb.initialize(0)
sim = bench.Sim(sb, args)
sim.adapt(sim.mra)
for _ in range(a.steps):
    sim.step()
lv, iv = [], []
for l in range(a.max_level + 1):
    x = sim.mesh.intervals(sb.CELLS, l)
    lv.append(np.full(x.size, l, np.int32)); iv.append(x)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"mesh_L{a.max_level}.npz" if a.dim == 2 else f"mesh_{a.dim}d_L{a.max_level}.npz"), levels=np.concatenate(lv), intervals=np.concatenate(iv))
print("leaves", sim.mesh.nb_cells(), "intervals", sum(len(x) for x in iv), "nproc", os.cpu_count())
