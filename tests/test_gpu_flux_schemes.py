"""Flux-based linear homogeneous schemes (make_convection_upwind, make_diffusion_order2) on uniform-level meshes:
bit-exact against the oracle's literal restatement of the reference's scatter loops, plus the reference's analytic
checks (tests/test_fv_operators.cpp:87-226: diffusion exact on quadratics, convection exact on linear fields, zero on
constants) and the explicit heat step of demos/FiniteVolume/heat.cpp."""
import numpy as np
import pytest

import parity_utils as pu

sb, so = pu.sb, pu.so
pytestmark = pytest.mark.gpu


def _setup(dim, L, fn):
    pmesh = sb.MRMesh.make_mesh([0.0] * dim, [1.0] * dim, pu.product_cfg(dim, L, L, 1))
    omesh = so.Mesh.uniform(pu.oracle_cfg(dim, L, L, 1))
    x = omesh.cell_centers(L, omesh.ref[L])
    ou = np.zeros(omesh.nref)
    ou[omesh.index(L, omesh.ref[L])] = fn(x)
    u = sb.make_scalar_field("u", pmesh)
    u.resize()
    u.upload(ou)
    return pmesh, omesh, u, ou


@pytest.mark.parametrize("dim,L", [(1, 6), (2, 5), (2, 7), (3, 4)])
def test_schemes_match_oracle_bitwise(gpu, dim, L):
    rng = np.random.default_rng(7)
    pmesh, omesh, u, ou = _setup(dim, L, lambda x: np.sin(3 * x[:, 0]) + 0.1 * rng.standard_normal(x.shape[0]))
    leaves = omesh.index(L, omesh.cells[L])
    for scheme, coeffs in ((sb.make_convection_upwind([1.0, -0.5, 0.25][:dim]), so.convection_upwind_coeffs([1.0, -0.5, 0.25][:dim])),
                           (sb.make_diffusion_order2([1.0, 2.0, 0.5][:dim]), so.diffusion_order2_coeffs([1.0, 2.0, 0.5][:dim]))):
        for bc_kind, bc_name, val in ((sb.DIRICHLET, "dirichlet", 0.3), (sb.NEUMANN, "neumann", -0.2)):
            sb.make_bc(u, bc_kind, val)
            u.upload(ou)  # clears ghosts_updated
            rhs = scheme(u)
            og = ou.copy()
            so.update_ghost_mr(omesh, og, so.Bc(bc_name, val))
            ref = so.flux_linhom_apply(omesh, og, coeffs)
            got = rhs.download()
            assert np.array_equal(got[leaves], ref[leaves]), f"{scheme.name} {bc_name}: max diff {np.max(np.abs(got[leaves] - ref[leaves])):.3e}"
            # u - dt * S(u)
            unp1 = sb.make_scalar_field("unp1", pmesh)
            sb.lincomb(unp1, 1.0, u, -0.01, rhs)
            assert np.array_equal(unp1.download()[leaves], (og - 0.01 * ref)[leaves])
            rhs.destroy()
            unp1.destroy()
    u.destroy()
    pmesh.destroy()


def test_analytic_exactness(gpu):
    """diffusion of a quadratic, convection of a linear field: exact on interior cells (test_fv_operators.cpp:87-226)."""
    dim, L = 2, 6
    pmesh, omesh, u, ou = _setup(dim, L, lambda x: x[:, 0] ** 2 + 2 * x[:, 1] ** 2 + 3 * x[:, 0])
    sb.make_bc(u, sb.DIRICHLET, 0.0)
    inner = omesh.index(L, so.box_cells([1, 1], [(1 << L) - 1] * 2))
    lap = sb.make_diffusion_order2([1.0, 1.0])(u).download()
    assert np.max(np.abs(lap[inner] + 6.0)) < 1e-9
    conv = sb.make_convection_upwind([1.0, 1.0])(u)
    u2 = sb.make_scalar_field("lin", pmesh)
    x = omesh.cell_centers(L, omesh.ref[L])
    lin = np.zeros(omesh.nref)
    lin[omesh.index(L, omesh.ref[L])] = 2 * x[:, 0] - 3 * x[:, 1]
    u2.resize()
    u2.upload(lin)
    sb.make_bc(u2, sb.DIRICHLET, 0.0)
    c2 = sb.make_convection_upwind([1.0, 0.5])(u2).download()
    assert np.max(np.abs(c2[inner] - 0.5)) < 1e-10


def test_adapted_mesh_is_rejected_not_approximated(gpu):
    pmesh = sb.MRMesh.make_mesh([0.0, 0.0], [1.0, 1.0], pu.product_cfg(2, 2, 6, 1))
    u = sb.make_scalar_field("u", pmesh)
    u.resize()
    u.init_ball([0.3, 0.3], 0.2)
    sb.make_bc(u, sb.DIRICHLET, 0.0)
    sb.make_MRAdapt(u)(sb.mra_config().epsilon(2e-4))
    with pytest.raises(ValueError, match="uniform-level"):
        sb.make_diffusion_order2([1.0, 1.0])(u)
