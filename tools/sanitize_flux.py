"""compute-sanitizer target (no torch import, so the process attaches quickly): the flux-scheme kernels added in round 2 -- WENO5 linear /
non-linear (scalar, vector) on an adapted periodic mesh with ghost width 3, vector Burgers upwind on an adapted Dirichlet mesh, the
two-cell schemes through periodic boundaries.  usage: compute-sanitizer --tool memcheck --launch-timeout 300 python tools/sanitize_flux.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import samurai_b200 as sb

assert sb.initialize(0)
for dim, lo, hi in ((1, 2, 8), (2, 1, 6), (3, 1, 4)):
    cfg = sb.mesh_config(dim, 1).min_level(lo).max_level(hi).periodic([True] * dim).max_stencil_size(6)
    mesh = sb.MRMesh.make_mesh([-1.0] * dim, [1.0] * dim, cfg)
    u = sb.make_scalar_field("u", mesh)
    u.resize()
    u.init_ball([-0.8] * dim, 0.35)
    sb.make_MRAdapt(u)(sb.mra_config())
    for scheme in (sb.make_convection_weno5([1.0, -1.0, 0.5][:dim]), sb.make_convection_weno5(), sb.make_convection_upwind(),
                   sb.make_diffusion_order2([1.0] * dim), sb.make_convection_upwind([1.0, -1.0, 0.5][:dim])):
        r = scheme(u)
        assert np.all(np.isfinite(r.download()))
        r.destroy()
    if dim > 1:
        v = sb.make_vector_field("v", mesh, dim)
        v.resize()
        for c in v.components:
            c.init_ball([-0.8] * dim, 0.35)
        for scheme in (sb.make_convection_weno5(), sb.make_convection_upwind()):
            r = scheme(v)
            assert np.all(np.isfinite(r.download()))
            r.destroy()
        v.destroy()
    print("dim", dim, "leaves", mesh.nb_cells(), "OK", flush=True)
    u.destroy()
    mesh.destroy()
# vector Burgers on a Dirichlet mesh
mesh = sb.MRMesh.make_mesh([0.0, 0.0], [1.0, 1.0], sb.mesh_config(2, 1).min_level(2).max_level(7))
v = sb.make_vector_field("v", mesh, 2)
v.resize()
for c in v.components:
    c.init_ball([0.3, 0.3], 0.2)
sb.make_bc(v, sb.DIRICHLET, 0.0, 0.0)
sb.make_MRAdapt(v)(sb.mra_config())
r = sb.make_convection_upwind()(v)
assert np.all(np.isfinite(r.download()))
print("vector Burgers OK", flush=True)
