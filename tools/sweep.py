"""Large uniform-mesh kernels in isolation (BASELINE.json configs[4] shapes) for ncu captures and bandwidth numbers:
the upwind FV sweep on a uniform 2D level-L mesh and one harten iteration from the uniform mesh (projection, detail,
criteria, maximum, update_fields at full size)."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import samurai_b200 as sb

ap = argparse.ArgumentParser()
ap.add_argument("--level", type=int, default=13)
ap.add_argument("--dim", type=int, default=2)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--adapt", action="store_true")
a = ap.parse_args()
sb.initialize(0)
dim = a.dim
if not a.adapt:
    cfg = sb.mesh_config(dim, 1).min_level(a.level).max_level(a.level).max_stencil_size(2).disable_minimal_ghost_width()
    mesh = sb.MRMesh.make_mesh([0.0] * dim, [1.0] * dim, cfg)
    u = sb.make_scalar_field("u", mesh); u.resize(); u.fill(0.0); u.init_ball([0.3] * dim, 0.2)
    sb.make_bc(u, sb.DIRICHLET, 0.0)
    v = sb.make_scalar_field("v", mesh); v.resize()
    sb.update_ghost_mr(u)
    dt = 0.5 * mesh.min_cell_length()
    sb.profile_enable(True)
    for _ in range(a.iters):
        sb.upwind_step(v, u, [1.0] * dim, dt)
        sb.swap(u, v)
    n, s, c = sb.profile_get()["fv"]
    print(f"fv uniform dim {dim} level {a.level}: {c/n:.0f} cells, {1e6*s/n:.1f} us/launch, {16*c/s/1e9:.1f} GB/s algorithmic")
else:
    cfg = sb.mesh_config(dim, 1).min_level(2).max_level(a.level).max_stencil_size(2).disable_minimal_ghost_width()
    mesh = sb.MRMesh.make_mesh([0.0] * dim, [1.0] * dim, cfg)
    u = sb.make_scalar_field("u", mesh); u.resize(); u.fill(0.0); u.init_ball([0.3] * dim, 0.2)
    sb.make_bc(u, sb.DIRICHLET, 0.0)
    ad = sb.make_MRAdapt(u)
    sb.profile_enable(True)
    t0 = time.perf_counter()
    ad.iteration(sb.mra_config().epsilon(2e-4), 0)
    sb.synchronize()
    print(f"one harten iteration from uniform level {a.level}: {time.perf_counter()-t0:.3f} s wall")
    nchild = 1 << dim
    bytes_per = {"fv": 16, "projection": 8 * (nchild + 1), "prediction": 8 * (1 + 1 / nchild), "detail": 8 * (1 + 2 * nchild),
                 "criteria": 8 * (nchild + 1) + 2 * nchild, "maximum": 2 * nchild + 2, "bc": 24, "copy": 16, "keep": 1, "init": 8}
    for k, (n, s, c) in sb.profile_get().items():
        if n:
            print(f"  {k:11s} launches {n:3d}  cells {c:12d}  total {1e3*s:8.3f} ms  {bytes_per[k]*c/s/1e9:8.1f} GB/s algorithmic")
