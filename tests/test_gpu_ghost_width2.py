"""Ghost width 2 -- the library's default (max_stencil_radius >= 2 unless disable_minimal_ghost_width(), mesh_config.hpp:388-393) -- at
non-periodic boundaries on the GPU: second ghost layer by polynomial extrapolation (bc/apply_field_bc.hpp:499-563), two-layer corner
block (:313-466), contiguous-boundary graduation rule (graduation.hpp:372-455), graduation width 2."""
import os

import numpy as np
import pytest

import parity_utils as pu
from parity_utils import sb, so

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim,min_level,max_level,steps", [(1, 2, 8, 12), (2, 2, 6, 8), (3, 1, 4, 3)])
def test_ghost_width_2_advection_matches_oracle(gpu, dim, min_level, max_level, steps):
    """the advection loop with the DEFAULT mesh_config (no disable_minimal_ghost_width): every step all sub-meshes and storage offsets
    bit-exact, tags and details of every harten iteration, every ghost the oracle defines after the ghost update, leaves after the step"""
    r = pu.run_advection_parity(dim=dim, min_level=min_level, max_level=max_level, pred_radius=1, steps=steps, msr=2, trace_tags=True,
                                a=[1.0] * dim, cfl=0.5 if dim < 3 else 0.25)
    assert r["max_rel_err"] <= pu.REL_TOL


def test_mra_burgers_hat_reproduces_reference_golden(gpu):
    """demos/FiniteVolume/burgers_mra.cpp --nfiles=1 --min-level=2 --max-level=9 --init-sol=hat --mr-eps=1e-5 on the GPU
    (tests/test_demo_finite_volume.py:191-207): 1D, box [-2, 3], max_stencil_radius(2).graduation_width(2), Dirichlet<1>(0), regularity 0,
    `unp1 = u - dt * scheme(u)` with scheme = 0.5 * make_convection_upwind<Field>() (the non-linear flux-based scheme),
    against the reference's own test_finite_volume_demo_mra_burgers_hat.h5 (tests/golden/mra_burgers_hat.npz)."""
    g = np.load(os.path.join(pu.ROOT, "tests", "golden", "mra_burgers_hat.npz"))

    def hat(x):
        out = np.zeros_like(x)
        m1 = (x > -1) & (x < 0)
        out[m1] = (1.0 / (0.0 - -1.0)) * (x[m1] - -1.0)
        m2 = (x >= 0) & (x < 1)
        out[m2] = (-1.0 / (1.0 - 0.0)) * (x[m2] - 0.0) + 1.0
        return out

    cfg = sb.mesh_config(1, 1).min_level(2).max_level(9).max_stencil_radius(2).graduation_width(2)
    pmesh = sb.MRMesh.make_mesh([-2.0], [3.0], cfg)
    ocfg = so.MeshConfig(dim=1, min_level=2, max_level=9, pred_radius=1, max_stencil_radius=2, graduation_width=2, origin=(-2.0,), scaling=5.0)
    om = so.Mesh.uniform(ocfg)
    pu.assert_same_mesh(pmesh, om)
    f0 = np.zeros(om.nref)
    f0[om.index(9, om.cells[9])] = hat(om.cell_centers(9, om.cells[9])[:, 0])
    u = sb.make_scalar_field("u", pmesh)
    u.resize()
    u.upload(f0)
    unp1 = sb.make_scalar_field("unp1", pmesh)
    sb.make_bc(u, sb.DIRICHLET, 0.0)
    sb.make_bc(unp1, sb.DIRICHLET, 0.0)
    scheme = 0.5 * sb.make_convection_upwind()
    adapt = sb.make_MRAdapt(u)
    mra = sb.mra_config().epsilon(1e-5).regularity(0.0)
    dt = 0.95 * pmesh.cell_length(9)
    Tf, t, nt = 0.1, 0.0, 0
    adapt(mra)
    while t != Tf:
        t += dt
        if t > Tf:
            dt += Tf - t
            t = Tf
        adapt(mra)
        unp1.resize()
        sb.lincomb(unp1, 1.0, u, -dt, scheme(u))
        sb.swap(u, unp1)
        nt += 1
    assert nt == 11
    lv, idx, off = pmesh.cell_table(sb.CELLS)
    assert np.array_equal(lv, g["level"].astype(np.int64)) and np.array_equal(idx[:, :1], g["idx"].astype(np.int64)), "mesh differs from the reference golden"
    got = u.download()[off]
    assert np.max(np.abs(got - g["u"])) <= 1e-15
