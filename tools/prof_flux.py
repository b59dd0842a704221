"""Flux-based diffusion on an adapted 2D mesh (levels 4..L, disc initial condition) for ncu captures of FluxGenOp and for
its algorithmic bandwidth: per apply 8 B zero fill + 8 B read + 8 B write per leaf (the neighbours are re-reads)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import samurai_b200 as sb

ap = argparse.ArgumentParser()
ap.add_argument("--level", type=int, default=13)
ap.add_argument("--eps", type=float, default=1e-5)
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
sb.initialize(0)
cfg = sb.mesh_config(2, 1).min_level(4).max_level(a.level).max_stencil_size(2).disable_minimal_ghost_width()
mesh = sb.MRMesh.make_mesh([0.0, 0.0], [1.0, 1.0], cfg)
u = sb.make_scalar_field("u", mesh); u.resize(); u.init_ball([0.3, 0.3], 0.2)
sb.make_bc(u, sb.DIRICHLET, 0.0)
sb.make_MRAdapt(u)(sb.mra_config().epsilon(a.eps))
diff = sb.make_diffusion_order2([1.0, 1.0])
rhs = diff(u)
sb.profile_enable(True)
for _ in range(a.iters):
    diff.apply(rhs, u)
n, s, c = sb.profile_get()["fv"]
print(f"flux diffusion, adapted 2D levels 4..{a.level}, eps {a.eps}: {mesh.nb_cells()} leaves, {1e6*s/n:.1f} us/launch, "
      f"{16*c/s/1e9:.1f} GB/s algorithmic (gather kernel alone: read u + write rhs)")
