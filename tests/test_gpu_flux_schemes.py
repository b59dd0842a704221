"""Flux-based schemes (make_convection_upwind linear and non-linear, make_diffusion_order2) on uniform-level and adapted
meshes: bit-exact against the oracle's literal restatement of the reference's scatter loops, plus the reference's analytic
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


def _adapted(dim, lo, hi, bc=("dirichlet", 0.0)):
    """Same adapted mesh on both sides (disc initial condition, one MRadaptation), leaves perturbed with seeded noise."""
    pmesh = sb.MRMesh.make_mesh([0.0] * dim, [1.0] * dim, pu.product_cfg(dim, lo, hi, 1))
    u = sb.make_scalar_field("u", pmesh)
    u.resize()
    u.init_ball([0.3] * dim, 0.2)
    sb.make_bc(u, sb.DIRICHLET if bc[0] == "dirichlet" else sb.NEUMANN, bc[1])
    sb.make_MRAdapt(u)(sb.mra_config().epsilon(2e-4))
    omesh = so.Mesh.uniform(pu.oracle_cfg(dim, lo, hi, 1))
    obc = so.Bc(*bc)
    omesh, ou = so.adapt(omesh, so.init_disc(omesh, [0.3] * dim, 0.2), obc, eps=2e-4)
    pu.assert_same_mesh(pmesh, omesh)
    rng = np.random.default_rng(11)
    leaves = np.concatenate([omesh.index(l, omesh.cells[l]) for l in omesh.leaf_levels()])
    ou[leaves] += 0.1 * rng.standard_normal(leaves.size)
    u.upload(ou)
    return pmesh, omesh, u, ou, obc, leaves


@pytest.mark.parametrize("dim,lo,hi", [(1, 2, 7), (2, 2, 6), (2, 3, 7), (3, 2, 5)])
@pytest.mark.parametrize("bc", [("dirichlet", 0.3), ("neumann", -0.2)])
def test_schemes_across_level_jumps_bitwise(gpu, dim, lo, hi, bc):
    """a8-a10 on adapted meshes: linear homogeneous and non-linear flux schemes, every leaf bit-identical to the oracle's
    literal restatement of the reference's scatter loops (same-level, both jump orientations, boundary)."""
    pmesh, omesh, u, ou, obc, leaves = _adapted(dim, lo, hi, bc)
    assert len(omesh.leaf_levels()) > 1
    og = ou.copy()
    so.update_ghost_mr(omesh, og, obc)
    vel, K = [1.0, -0.5, 0.25][:dim], [1.0, 2.0, 0.5][:dim]
    cases = [(sb.make_convection_upwind(vel), so.flux_linhom_apply(omesh, og, so.convection_upwind_coeffs(vel))),
             (sb.make_diffusion_order2(K), so.flux_linhom_apply(omesh, og, so.diffusion_order2_coeffs(K))),
             (sb.make_convection_upwind(), so.flux_nonlin_apply(omesh, og, so.burgers_upwind_flux())),
             (0.5 * sb.make_convection_upwind(), so.flux_nonlin_apply(omesh, og, so.burgers_upwind_flux(0.5)))]
    for scheme, ref in cases:
        rhs = scheme(u)
        got = rhs.download()
        bad = np.flatnonzero(got[leaves] != ref[leaves])
        assert bad.size == 0, f"{scheme.name}: {bad.size} of {leaves.size} leaves differ, max {np.max(np.abs(got[leaves] - ref[leaves])):.3e}"
        rhs.destroy()
    u.destroy()
    pmesh.destroy()


def test_general_kernel_equals_strip_kernel_on_uniform_mesh(gpu, monkeypatch):
    """the uniform-level fast path (strip kernel) and the general gather kernel give the same bits"""
    import subprocess, sys, os
    code = ("import numpy as np, sys; sys.path.insert(0, 'tests'); import parity_utils as pu; sb = pu.sb\n"
            "assert sb.initialize(0)\n"
            "m = sb.MRMesh.make_mesh([0.,0.],[1.,1.], pu.product_cfg(2, 6, 6, 1)); u = sb.make_scalar_field('u', m); u.resize()\n"
            "rng = np.random.default_rng(3); u.upload(rng.standard_normal(m.nb_cells(sb.REFERENCE))); sb.make_bc(u, sb.DIRICHLET, 0.1)\n"
            "np.save(sys.argv[1], sb.make_diffusion_order2([1., 2.])(u).download())\n")
    outs = []
    for env_extra in ({}, {"SMR_FLUX_GENERAL": "1"}):
        path = f"/tmp/flux_uniform_{len(outs)}.npy"
        subprocess.run([sys.executable, "-c", code, path], check=True, cwd=pu.ROOT, env={**os.environ, **env_extra})
        outs.append(np.load(path))
    omesh = so.Mesh.uniform(pu.oracle_cfg(2, 6, 6, 1))
    leaves = omesh.index(6, omesh.cells[6])
    assert np.array_equal(outs[0][leaves], outs[1][leaves])


def test_heat_demo_reproduces_reference_golden(gpu):
    """demos/FiniteVolume/heat.cpp --explicit --init-sol=dirac --Tf=0.1 --min-level=3 --max-level=6 on the GPU against the
    reference's own golden file (tests/golden/heat_explicit.npz <- test_finite_volume_demo_heat_explicit.h5):
    mesh identical, field within the reference's tolerance (rel 1e-14 / abs 1e-7; we require 1e-15 absolute)."""
    import os
    dim = 2
    cfg = pu.product_cfg(dim, 3, 6, 1)
    pmesh = sb.MRMesh.make_mesh([-4.0, -4.0], [4.0, 4.0], cfg)
    ocfg = so.MeshConfig(dim=2, min_level=3, max_level=6, pred_radius=1, origin=(-4.0, -4.0), scaling=8.0)
    om = so.Mesh.uniform(ocfg)
    t = 1e-2
    f0 = np.zeros(om.nref)
    f0[om.index(6, om.cells[6])] = so.heat_exact(om.cell_centers(6, om.cells[6]), t)
    u = sb.make_scalar_field("u", pmesh)
    u.resize()
    u.upload(f0)
    unp1 = sb.make_scalar_field("unp1", pmesh)
    sb.make_bc(u, sb.NEUMANN, 0.0)
    sb.make_bc(unp1, sb.NEUMANN, 0.0)
    diff = sb.make_diffusion_order2([1.0, 1.0])
    dx = pmesh.cell_length(6)
    dt = 0.95 * (dx * dx) / (pow(2, dim) * 1.0)
    adapt = sb.make_MRAdapt(u)
    mra = sb.mra_config()
    adapt(mra)
    Tf, nt = 0.1, 0
    while t != Tf:
        t += dt
        if t > Tf:
            dt += Tf - t
            t = Tf
        adapt(mra)
        unp1.resize()
        sb.lincomb(unp1, 1.0, u, -dt, diff(u))
        sb.swap(u, unp1)
        nt += 1
    assert nt == 25
    g = np.load(os.path.join(pu.ROOT, "tests", "golden", "heat_explicit.npz"))
    lv, idx, off = pmesh.cell_table(sb.CELLS)
    assert np.array_equal(lv, g["level"].astype(np.int64)) and np.array_equal(idx[:, :2], g["idx"].astype(np.int64)), "mesh differs"
    got = u.download()[off]
    assert np.max(np.abs(got - g["u"])) <= 1e-15


def _two_level_mesh_2d():
    """reference tests/test_fv_operators.cpp:965-993: coarse left half (level 3), fine right half (level 4), jump at x = 1/2"""
    nc, nf = 8, 16
    levels = [3] * nc + [4] * nf
    ivl = np.zeros(nc + nf, dtype=sb.INTERVAL_DTYPE)
    for y in range(nc):
        ivl[y] = (y, 0, 0, nc // 2, 0)
    for y in range(nf):
        ivl[nc + y] = (y, 0, nf // 2, nf, 0)
    cfg = sb.mesh_config(2, 1).min_level(3).max_level(4).max_stencil_size(2).disable_minimal_ghost_width()
    return sb.MRMesh.from_intervals([0.0, 0.0], [1.0, 1.0], cfg, levels, ivl)


def test_level_jump_constant_conservation(gpu):
    """fv_operators.level_jump_constant_conservation (:1016-1052): every operator vanishes on a constant field, everywhere,
    including the level-jump interface and the boundary (the Dirichlet value matches the constant)."""
    mesh = _two_level_mesh_2d()
    lv, co, off = mesh.cell_table(sb.CELLS)
    assert set(np.unique(lv)) == {3, 4}
    u = sb.make_scalar_field("s", mesh)
    u.resize()
    u.fill(2.5)
    sb.make_bc(u, sb.DIRICHLET, 2.5)
    for scheme in (sb.make_diffusion_order2([1.0, 1.0]), sb.make_convection_upwind([1.0, 2.0]), sb.make_convection_upwind([-1.0, -2.0])):
        r = scheme(u).download()
        assert np.max(np.abs(r[off])) < 1e-10, scheme.name


def test_level_jump_linear_exactness_of_diffusion(gpu):
    """fv_operators.level_jump_linear_exactness (:1068-1128), diffusion part: the central flux stays exact across the jump
    because projection and prediction reproduce a linear field.  The reference imposes the linear field on the boundary
    with a function-valued Dirichlet (not on the device path): here the BC is a constant, so the cells whose ghost
    stencils reach the boundary (two coarse cells from it) are left out instead of only the boundary-touching ones."""
    a, b, c = 3.0, -2.0, 1.0
    mesh = _two_level_mesh_2d()
    lv, co, off = mesh.cell_table(sb.CELLS)
    h = np.array([mesh.cell_length(int(l)) for l in lv])
    x = (co + 0.5) * h[:, None]
    host = np.zeros(mesh.nb_cells(sb.REFERENCE))
    host[off] = a * x[:, 0] + b * x[:, 1] + c
    u = sb.make_scalar_field("u", mesh)
    u.resize()
    u.upload(host)
    sb.make_bc(u, sb.DIRICHLET, 0.0)
    r = sb.make_diffusion_order2([1.0, 1.0])(u).download()
    inner = np.all((x > 0.25) & (x < 0.75), axis=1)
    near_jump = inner & (np.abs(x[:, 0] - 0.5) < 0.13)
    assert near_jump.sum() > 8 and set(np.unique(lv[near_jump])) == {3, 4}
    assert np.max(np.abs(r[off][inner])) < 1e-9


@pytest.mark.parametrize("dim,lo,hi,kind", [(1, 2, 8, "burgers"), (2, 2, 6, "burgers"), (2, 2, 6, "diffusion"), (3, 1, 4, "convection")])
def test_flux_scheme_time_loop_matches_oracle(gpu, dim, lo, hi, kind):
    """`MRadaptation; unp1 = u - dt * scheme(u); swap` for several steps (the loop of demos/FiniteVolume/burgers_mra.cpp:142-160 and
    heat.cpp:196-222) on the GPU and on the oracle: meshes identical at every step, leaves bit-identical."""
    pmesh = sb.MRMesh.make_mesh([0.0] * dim, [1.0] * dim, pu.product_cfg(dim, lo, hi, 1))
    omesh = so.Mesh.uniform(pu.oracle_cfg(dim, lo, hi, 1))
    bc = so.Bc("dirichlet", 0.0)
    ou = so.init_disc(omesh, [0.3] * dim, 0.2)
    u = sb.make_scalar_field("u", pmesh)
    u.resize()
    u.upload(ou)
    unp1 = sb.make_scalar_field("unp1", pmesh)
    sb.make_bc(u, sb.DIRICHLET, 0.0)
    sb.make_bc(unp1, sb.DIRICHLET, 0.0)
    vel, K = [1.0, 0.5, -0.25][:dim], [1.0] * dim
    if kind == "burgers":
        scheme, dt = 0.5 * sb.make_convection_upwind(), 0.5 * pmesh.min_cell_length()
        apply_o = lambda m, f: so.flux_nonlin_apply(m, f, so.burgers_upwind_flux(0.5))
    elif kind == "diffusion":
        scheme, dt = sb.make_diffusion_order2(K), 0.2 * pmesh.min_cell_length() ** 2
        apply_o = lambda m, f: so.flux_linhom_apply(m, f, so.diffusion_order2_coeffs(K))
    else:
        scheme, dt = sb.make_convection_upwind(vel), 0.25 * pmesh.min_cell_length()
        apply_o = lambda m, f: so.flux_linhom_apply(m, f, so.convection_upwind_coeffs(vel))
    adapt = sb.make_MRAdapt(u)
    mra = sb.mra_config().epsilon(2e-4)
    for step in range(4):
        adapt(mra)
        omesh, ou = so.adapt(omesh, ou, bc, 2e-4, 1.0)
        pu.assert_same_mesh(pmesh, omesh)
        unp1.resize()
        sb.lincomb(unp1, 1.0, u, -dt, scheme(u))
        sb.swap(u, unp1)
        so.update_ghost_mr(omesh, ou, bc)
        ou = so.lincomb_leaves(omesh, 1.0, ou, -dt, apply_o(omesh, ou))
        _, _, leaf = omesh.leaf_table()
        got = u.download()
        assert np.array_equal(got[leaf], ou[leaf]), f"step {step}: max diff {np.max(np.abs(got[leaf] - ou[leaf])):.3e}"
    assert len(omesh.leaf_levels()) > 1


@pytest.mark.parametrize("dim,lo,hi", [(2, 2, 6), (2, 3, 7), (3, 2, 5)])
@pytest.mark.parametrize("bc", [("dirichlet", 0.3), ("neumann", -0.2)])
def test_vector_convection_upwind_across_level_jumps_bitwise(gpu, dim, lo, hi, bc):
    """a9, vector form: make_convection_upwind<VectorField>() with n_comp == dim (flux u(d) * u upwinded by the mean of component d,
    operators/convection_nonlin.hpp:24-76) on adapted meshes: every component of every leaf bit-identical to the oracle's restatement
    of the reference's scatter loops (same-level, both jump orientations, boundary), with and without a scalar factor."""
    pmesh, omesh, u0, ou0, obc, leaves = _adapted(dim, lo, hi, bc)
    assert len(omesh.leaf_levels()) > 1
    rng = np.random.default_rng(23)
    v = sb.make_vector_field("v", pmesh, dim)
    v.resize()
    comps = []
    for c in range(dim):
        oc = ou0.copy()
        oc[leaves] += 0.3 * rng.standard_normal(leaves.size) - 0.2 * c  # both signs of the upwinding velocity occur
        comps.append(oc)
    v.upload(np.stack(comps, axis=1))
    sb.make_bc(v, sb.DIRICHLET if bc[0] == "dirichlet" else sb.NEUMANN, *([bc[1]] * dim))
    og = [c.copy() for c in comps]
    for c in og:
        so.update_ghost_mr(omesh, c, obc)
    for scale in (1.0, 0.5):
        scheme = sb.make_convection_upwind() if scale == 1.0 else scale * sb.make_convection_upwind()
        want = so.flux_nonlin_apply(omesh, og, so.burgers_upwind_flux_vector(dim, scale))
        rhs = scheme(v)
        got = rhs.download()
        for c in range(dim):
            bad = np.flatnonzero(got[leaves, c] != want[c][leaves])
            assert bad.size == 0, (f"scale {scale} component {c}: {bad.size} of {leaves.size} leaves differ, "
                                   f"max {np.max(np.abs(got[leaves, c] - want[c][leaves])):.3e}")
        rhs.destroy()
    v.destroy()
    u0.destroy()
    pmesh.destroy()


@pytest.mark.parametrize("dim,lo,hi,msr", [(1, 2, 7, 1), (2, 2, 6, 1), (2, 2, 6, 2), (3, 2, 4, 1)])
def test_schemes_on_periodic_meshes_bitwise(gpu, dim, lo, hi, msr):
    """a10, periodic variants (interface.hpp:83-92 same level, :179-189 and :280-290 level jumps): the two-cell flux schemes on a fully
    periodic adapted mesh whose refined region touches the boundary, so that same-level interfaces and level jumps go through it.
    Linear homogeneous (upwind convection, diffusion), non-linear scalar (Burgers) and vector (u(d) * u): every leaf bit-identical to
    the oracle's restatement of the reference's scatter loops (no boundary interfaces in a periodic direction)."""
    periodic = (True,) * dim
    pmesh = sb.MRMesh.make_mesh([0.0] * dim, [1.0] * dim, pu.product_cfg(dim, lo, hi, 1, periodic, msr))
    u = sb.make_scalar_field("u", pmesh)
    u.resize()
    u.init_ball([0.1] * dim, 0.2)  # crosses the lower boundaries: its periodic images matter
    sb.make_MRAdapt(u)(sb.mra_config().epsilon(2e-4))
    ocfg = pu.oracle_cfg(dim, lo, hi, 1, periodic, msr)
    obc = so.Bc("neumann", 0.0)
    omesh = so.Mesh.uniform(ocfg)
    omesh, ou = so.adapt(omesh, so.init_disc(omesh, [0.1] * dim, 0.2), obc, eps=2e-4)
    pu.assert_same_mesh(pmesh, omesh)
    assert len(omesh.leaf_levels()) > 1
    rng = np.random.default_rng(31)
    leaves = np.concatenate([omesh.index(l, omesh.cells[l]) for l in omesh.leaf_levels()])
    ou[leaves] += 0.2 * rng.standard_normal(leaves.size) - 0.05
    u.upload(ou)
    og = ou.copy()
    so.update_ghost_mr(omesh, og, obc)
    vel, K = [1.0, -0.5, 0.25][:dim], [1.0, 2.0, 0.5][:dim]
    cases = [(sb.make_convection_upwind(vel), so.flux_linhom_apply(omesh, og, so.convection_upwind_coeffs(vel))),
             (sb.make_convection_upwind([-v for v in vel]), so.flux_linhom_apply(omesh, og, so.convection_upwind_coeffs([-v for v in vel]))),
             (sb.make_diffusion_order2(K), so.flux_linhom_apply(omesh, og, so.diffusion_order2_coeffs(K))),
             (sb.make_convection_upwind(), so.flux_nonlin_apply(omesh, og, so.burgers_upwind_flux())),
             (0.5 * sb.make_convection_upwind(), so.flux_nonlin_apply(omesh, og, so.burgers_upwind_flux(0.5)))]
    for scheme, ref in cases:
        rhs = scheme(u)
        got = rhs.download()
        bad = np.flatnonzero(got[leaves] != ref[leaves])
        assert bad.size == 0, f"{scheme.name}: {bad.size} of {leaves.size} leaves differ, max {np.max(np.abs(got[leaves] - ref[leaves])):.3e}"
        rhs.destroy()
    if dim > 1:
        v = sb.make_vector_field("v", pmesh, dim)
        v.resize()
        comps = []
        for c in range(dim):
            oc = ou.copy()
            oc[leaves] += 0.3 * rng.standard_normal(leaves.size) - 0.2 * c
            comps.append(oc)
        v.upload(np.stack(comps, axis=1))
        ogv = [c.copy() for c in comps]
        for c in ogv:
            so.update_ghost_mr(omesh, c, obc)
        want = so.flux_nonlin_apply(omesh, ogv, so.burgers_upwind_flux_vector(dim))
        rhsv = sb.make_convection_upwind()(v)
        got = rhsv.download()
        for c in range(dim):
            bad = np.flatnonzero(got[leaves, c] != want[c][leaves])
            assert bad.size == 0, f"vector component {c}: {bad.size} of {leaves.size} leaves differ"
        rhsv.destroy()
        v.destroy()
    u.destroy()
    pmesh.destroy()
