"""WENO5 linear convection + TVD-RK3 on the GPU: make_convection_weno5 (operators/convection_lin.hpp:95-178,
weno_impl.hpp:26-63) is a NON-LINEAR flux scheme with the six-cell line stencil {-2 .. 3}; the mesh needs max_stencil_size(6) (ghost
width 3) and, without Dirichlet<3>, is fully periodic as in demos/FiniteVolume/linear_convection.cpp.  Checked against the oracle
(pinned on the reference's linear_convection_explicit golden, tests/test_oracle_golden.py) and against that golden directly."""
import os

import numpy as np
import pytest

import parity_utils as pu

sb, so = pu.sb, pu.so
pytestmark = pytest.mark.gpu


def _cfgs(dim, lo, hi, box=(-1.0, 1.0)):
    pcfg = sb.mesh_config(dim, 1).min_level(lo).max_level(hi).periodic([True] * dim).max_stencil_size(6)
    ocfg = so.MeshConfig(dim=dim, min_level=lo, max_level=hi, pred_radius=1, max_stencil_radius=3, graduation_width=1,
                         origin=(box[0],) * dim, scaling=box[1] - box[0], periodic=(True,) * dim)
    return pcfg, ocfg


def _leaves(omesh):
    return np.concatenate([omesh.index(l, omesh.cells[l]) for l in omesh.leaf_levels()])


@pytest.mark.parametrize("dim,lo,hi,vel", [(1, 2, 8, [1.0]), (1, 2, 8, [-0.7]), (2, 1, 6, [1.0, -1.0]), (2, 2, 6, [-0.5, 2.0]),
                                           (3, 1, 4, [1.0, -1.0, 0.5])])
def test_weno5_scheme_on_adapted_periodic_mesh(gpu, dim, lo, hi, vel):
    """rhs = make_convection_weno5(velocity)(u) on an adapted fully periodic mesh (level jumps, interfaces through the periodic
    boundary, both velocity signs): every leaf against the oracle's restatement of the reference's scatter loops.  Same operations
    in the same order: bit-equal expected, 1e-12 relative required (north_star)."""
    pcfg, ocfg = _cfgs(dim, lo, hi)
    pmesh = sb.MRMesh.make_mesh([-1.0] * dim, [1.0] * dim, pcfg)
    omesh = so.Mesh.uniform(ocfg)
    c = omesh.cell_centers(hi, omesh.cells[hi])
    # a box that touches the periodic boundary on the low side: leaves of different levels face each other through the boundary
    inside = np.all((c >= -1.0) & (c <= -0.45), axis=1)
    f0 = np.zeros(omesh.nref)
    f0[omesh.index(hi, omesh.cells[hi])] = np.where(inside, 1.0, 0.0)
    u = sb.make_scalar_field("u", pmesh)
    u.resize()
    u.upload(f0)
    sb.make_MRAdapt(u)(sb.mra_config())
    bc = so.Bc("neumann", 0.0)
    omesh, ou = so.adapt(omesh, f0, bc, 1e-4, 1.0)
    pu.assert_same_mesh(pmesh, omesh)
    assert len(omesh.leaf_levels()) > 1
    leaves = _leaves(omesh)
    rng = np.random.default_rng(5)
    ou[leaves] += 0.1 * rng.standard_normal(leaves.size)
    u.upload(ou)
    og = ou.copy()
    so.update_ghost_mr(omesh, og, bc)
    ref = so.flux_nonlin_apply(omesh, og, so.weno5_flux(vel), so.WENO5_OFFSETS)
    for scale in (1.0, 0.5):
        scheme = sb.make_convection_weno5(vel)
        if scale != 1.0:
            scheme = scale * scheme
        rhs = scheme(u)
        got = rhs.download()
        want = ref if scale == 1.0 else so.flux_nonlin_apply(omesh, og, [(lambda *s, f=f: f(*s) * scale) for f in so.weno5_flux(vel)], so.WENO5_OFFSETS)
        err = np.max(np.abs(got[leaves] - want[leaves])) / max(1.0, np.max(np.abs(want[leaves])))
        assert err <= pu.REL_TOL, f"scale {scale}: max rel err {err:.3e}, {np.count_nonzero(got[leaves] != want[leaves])} of {leaves.size} leaves differ"
        rhs.destroy()
    u.destroy()
    pmesh.destroy()


def test_linear_convection_demo_matches_oracle_and_reference_golden(gpu):
    """demos/FiniteVolume/linear_convection.cpp --min-level=1 --max-level=6 --Tf=0.1 (explicit): MRadaptation + TVD-RK3 with WENO5 every
    step.  After every step the mesh is identical to the oracle's and the leaves agree within 1e-12; the final state reproduces the
    reference's own test_finite_volume_demo_linear_convection_explicit.h5 (mesh identical, 1e-13)."""
    dim, lo, hi = 2, 1, 6
    pcfg, ocfg = _cfgs(dim, lo, hi)
    states = []
    r = so.run_linear_convection(ocfg, Tf=0.1, on_step=lambda nt, m, f: states.append((m, f.copy())))
    assert r["steps"] == 7
    pmesh = sb.MRMesh.make_mesh([-1.0] * dim, [1.0] * dim, pcfg)
    om = so.Mesh.uniform(ocfg)
    c = om.cell_centers(hi, om.cells[hi])
    f0 = np.zeros(om.nref)
    f0[om.index(hi, om.cells[hi])] = np.where((c[:, 0] >= -0.8) & (c[:, 0] <= -0.3) & (c[:, 1] >= 0.3) & (c[:, 1] <= 0.8), 1.0, 0.0)
    u = sb.make_scalar_field("u", pmesh)
    u.resize()
    u.upload(f0)
    unp1, u1, u2, tmp = (sb.make_scalar_field(n, pmesh) for n in ("unp1", "u1", "u2", "tmp"))
    vel = [1.0, -1.0]
    conv = sb.make_convection_weno5(vel)
    dt = 0.95 * pmesh.cell_length(hi) / 2.0
    adapt = sb.make_MRAdapt(u)
    mra = sb.mra_config()
    adapt(mra)
    pu.assert_same_mesh(pmesh, r["init"][0])
    t, Tf, nt, worst = 0.0, 0.1, 0, 0.0
    while t != Tf:
        t += dt
        if t > Tf:
            dt += Tf - t
            t = Tf
        adapt(mra)
        for f in (unp1, u1, u2, tmp):
            f.resize()
        c0 = conv(u)
        sb.lincomb(u1, 1.0, u, -dt, c0)                    # u1 = u - dt * conv(u)
        c1 = conv(u1)
        sb.lincomb(tmp, 1.0, u1, -dt, c1)
        sb.lincomb(u2, 3. / 4, u, 1. / 4, tmp)             # u2 = 3/4 u + 1/4 (u1 - dt * conv(u1))
        c2 = conv(u2)
        sb.lincomb(tmp, 1.0, u2, -dt, c2)
        sb.lincomb(unp1, 1. / 3, u, 2. / 3, tmp)           # unp1 = 1/3 u + 2/3 (u2 - dt * conv(u2))
        for f in (c0, c1, c2):
            f.destroy()
        sb.swap(u, unp1)
        omesh, ou = states[nt]
        pu.assert_same_mesh(pmesh, omesh)
        leaves = _leaves(omesh)
        got = u.download()
        worst = max(worst, float(np.max(np.abs(got[leaves] - ou[leaves]))))
        nt += 1
    assert nt == 7
    assert worst <= pu.REL_TOL, f"max abs difference to the oracle over the run: {worst:.3e}"
    g = np.load(os.path.join(pu.ROOT, "tests", "golden", "linear_convection_explicit.npz"))
    lv, idx, off = pmesh.cell_table(sb.CELLS)
    assert np.array_equal(lv, g["level"].astype(np.int64)) and np.array_equal(idx[:, :2], g["idx"].astype(np.int64)), "mesh differs from the golden"
    assert np.max(np.abs(u.download()[off] - g["u"])) <= 1e-13


@pytest.mark.parametrize("dim,lo,hi,n_comp", [(1, 2, 8, 1), (2, 1, 6, 1), (2, 1, 6, 2), (3, 1, 4, 3)])
def test_nonlinear_weno5_burgers_steps_match_oracle(gpu, dim, lo, hi, n_comp):
    """make_convection_weno5<Field>() (Burgers form, scalar and vector: operators/convection_nonlin.hpp:162-233) in
    `unp1 = u - dt * conv(u)` with MRadaptation at every step on a fully periodic mesh: meshes identical to the oracle's at every step,
    leaves within 1e-12 (bit-equal expected)."""
    pcfg, ocfg = _cfgs(dim, lo, hi)
    pmesh = sb.MRMesh.make_mesh([-1.0] * dim, [1.0] * dim, pcfg)
    om = so.Mesh.uniform(ocfg)
    c = om.cell_centers(hi, om.cells[hi])
    ix = om.index(hi, om.cells[hi])
    r = np.max(np.abs(c + 0.4), axis=1)
    comps = []
    for k in range(n_comp):
        f = np.zeros(om.nref)
        f[ix] = np.where(r < 0.5, (1.0 - 2.0 * r) * (1.0 if k % 2 == 0 else -0.7), 0.0)  # hats of both signs, touching the periodic boundary
        comps.append(f)
    bcs = [so.Bc("neumann", 0.0)] * n_comp
    if n_comp == 1:
        u, unp1 = sb.make_scalar_field("u", pmesh), sb.make_scalar_field("unp1", pmesh)
        u.resize()
        u.upload(comps[0])
    else:
        u, unp1 = sb.make_vector_field("u", pmesh, n_comp), sb.make_vector_field("unp1", pmesh, n_comp)
        u.resize()
        u.upload(np.stack(comps, axis=1))
    conv = sb.make_convection_weno5()
    flux = so.weno5_flux_nonlinear(dim, n_comp)
    dt = 0.3 * pmesh.cell_length(hi)
    adapt = sb.make_MRAdapt(u)
    mra = sb.mra_config()
    adapt(mra)
    om, comps = so.adapt_fields(om, comps, bcs, 1e-4, 1.0)
    pu.assert_same_mesh(pmesh, om)
    worst = 0.0
    for step in range(6):
        adapt(mra)
        om, comps = so.adapt_fields(om, comps, bcs, 1e-4, 1.0)
        pu.assert_same_mesh(pmesh, om)
        unp1.resize()
        rhs = conv(u)
        for a, b, cc in zip(sb._scalars(unp1), sb._scalars(u), sb._scalars(rhs)):
            sb.lincomb(a, 1.0, b, -dt, cc)
        rhs.destroy()
        sb.swap(u, unp1)
        for f, bc in zip(comps, bcs):
            so.update_ghost_mr(om, f, bc)
        orhs = so.flux_nonlin_apply(om, comps if n_comp > 1 else comps[0], flux, so.WENO5_OFFSETS)
        orhs = orhs if n_comp > 1 else [orhs]
        leaves = _leaves(om)
        new = []
        for f, rh in zip(comps, orhs):
            o = np.full(om.nref, np.nan)
            o[leaves] = f[leaves] - dt * rh[leaves]
            new.append(o)
        comps = new
        got = u.download()
        got = got if n_comp > 1 else got[:, None]
        for k in range(n_comp):
            worst = max(worst, float(np.max(np.abs(got[leaves, k] - comps[k][leaves]))))
    assert len(om.leaf_levels()) > 1
    assert worst <= pu.REL_TOL, f"max abs difference to the oracle: {worst:.3e}"
