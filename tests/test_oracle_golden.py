"""Pins the oracle (oracle/samurai_oracle.py) on the reference's OWN golden files.

tests/golden/advection_2d_pred_{0,1}{,_init}.npz are the datasets of
/root/reference/tests/reference/finite_volume/test_finite_volume_advection_2d_*.h5 (converted by
tests/golden/make_golden.py): the outputs of samurai's finite-volume-advection-2d demo run with --Tf 0.01 that samurai's
pytest suite compares against (tests/test_demo_finite_volume.py:55-72, rel 1e-14 / abs 1e-7).  The full pipeline is
exercised: uniform level-10 mesh -> MRadaptation -> 21 x (MRadaptation, update_ghost_mr, upwind) for prediction radius
0 and 1.  Mesh: exact.  Field: the reference's own tolerance (we observe <= 4.5e-16 absolute).
"""
import os

import numpy as np
import pytest

import samurai_oracle as so

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _check(mesh, u, name):
    g = np.load(os.path.join(GOLD, name))
    lv, co, ix = mesh.leaf_table()
    assert lv.size == g["level"].size, f"{name}: {lv.size} leaves vs golden {g['level'].size}"
    assert np.array_equal(lv, g["level"].astype(np.int64)), f"{name}: levels differ"
    assert np.array_equal(co, g["idx"].astype(np.int64)), f"{name}: cell indices differ (mesh not identical)"
    ref = g["u"]
    got = u[ix]
    # reference tolerance: pytest.approx(rel=1e-14, abs=1e-7) (tests/conftest.py:121-122); we require far tighter
    assert np.max(np.abs(got - ref)) <= 1e-14, f"{name}: max abs diff {np.max(np.abs(got - ref)):.3e}"


@pytest.mark.parametrize("pred", [0, 1])
def test_oracle_matches_reference_golden(pred):
    cfg = so.MeshConfig(dim=2, min_level=4, max_level=10, pred_radius=pred)
    res = so.run_advection(cfg, Tf=0.01, eps=2e-4)
    assert res["steps"] == 21
    _check(*res["init"], f"advection_2d_pred_{pred}_init.npz")
    _check(*res["final"], f"advection_2d_pred_{pred}.npz")


def test_projection_prediction_exactness():
    """reference tests/test_projection_prediction_roundtrip.cpp: projection is the exact mean of the children and the
    order-1 prediction reproduces polynomials of per-axis degree <= 2 exactly (up to rounding)."""
    cfg = so.MeshConfig(dim=2, min_level=2, max_level=5, pred_radius=1)
    # two-level mesh: left half at level 4, right half at level 5
    c4 = so.box_cells([0, 0], [8, 16])
    c5 = so.box_cells([16, 0], [32, 32])
    mesh = so.Mesh(cfg, {4: c4, 5: c5})
    f = np.zeros(mesh.nref)

    def poly(level, keys):
        x = mesh.cell_centers(level, keys)
        h = cfg.cell_length(level)
        # cell average of 1 + 2x + 3y + x^2 over the cell (exact): x^2 average = xc^2 + h^2/12
        return 1 + 2 * x[:, 0] + 3 * x[:, 1] + x[:, 0] ** 2 + h * h / 12 + 0.5 * x[:, 0] * x[:, 1]

    for l in range(mesh.nlev):
        if mesh.ref[l].size:
            f[mesh.index(l, mesh.ref[l])] = poly(l, mesh.ref[l])
    exact = f.copy()
    # projection of level 5 onto its parents
    parents = so.coarsen(c5, 1, 2)
    f[mesh.index(4, parents)] = -1
    so.projection(mesh, f, 4, parents)
    assert np.max(np.abs(f[mesh.index(4, parents)] - exact[mesh.index(4, parents)])) < 1e-13
    # prediction of interior level-5 cells from level 4 (parents with a full neighbourhood in the reference mesh)
    inner = so.box_cells([18, 2], [30, 30])
    pred = so.predict_values(mesh, exact, mesh, 5, inner)
    assert np.max(np.abs(pred - exact[mesh.index(5, inner)])) < 1e-12


def test_upwind_constant_and_linear():
    """reference tests/test_fv_operators.cpp:87-226: upwind convection of a constant is 0, of a linear field is a.grad."""
    cfg = so.MeshConfig(dim=2, min_level=5, max_level=5)
    mesh = so.Mesh.uniform(cfg)
    u = np.zeros(mesh.nref)
    x = mesh.cell_centers(5, mesh.ref[5])
    u[mesh.index(5, mesh.ref[5])] = 2 * x[:, 0] - 3 * x[:, 1] + 1
    a = [1.0, 0.5]
    out = so.fv_step(mesh, u, a, dt=1.0)
    leaves = mesh.index(5, mesh.cells[5])
    # unp1 = u - dt * (a . grad u) = u - (2*1 - 3*0.5)
    assert np.max(np.abs(out[leaves] - (u[leaves] - 0.5))) < 1e-11


def test_oracle_matches_reference_heat_golden():
    """Pins the flux-based path (a8/a10: make_diffusion_order2 across level jumps, Neumann(0), adaptation every step)
    on the reference's own golden: demos/FiniteVolume/heat.cpp --explicit --init-sol=dirac --Tf=0.1 --min-level=3
    --max-level=6 (tests/test_demo_finite_volume.py:99-116) -> test_finite_volume_demo_heat_explicit.h5."""
    cfg = so.MeshConfig(dim=2, min_level=3, max_level=6, pred_radius=1, origin=(-4.0, -4.0), scaling=8.0)
    res = so.run_heat(cfg)
    assert res["steps"] == 25
    mesh, u = res["final"]
    g = np.load(os.path.join(GOLD, "heat_explicit.npz"))
    lv, co, ix = mesh.leaf_table()
    assert np.array_equal(lv, g["level"].astype(np.int64)) and np.array_equal(co, g["idx"].astype(np.int64)), "mesh not identical"
    assert np.max(np.abs(u[ix] - g["u"])) <= 1e-15  # observed 2.2e-19 (max |u| = 0.77)


@pytest.mark.parametrize("dim,lo,hi", [(1, 2, 6), (2, 2, 6), (3, 2, 5)])
def test_flux_schemes_on_adapted_meshes(dim, lo, hi):
    """reference tests/test_fv_operators.cpp:953-1128 (two-level mesh): every flux operator returns 0 on a constant field
    including across level jumps; the conservative form telescopes (sum of h^dim * out = boundary fluxes only)."""
    cfg = so.MeshConfig(dim=dim, min_level=lo, max_level=hi, pred_radius=1)
    mesh = so.Mesh.uniform(cfg)
    bc = so.Bc("dirichlet", 0.0)
    mesh, u = so.adapt(mesh, so.init_disc(mesh, [0.3] * dim, 0.2), bc, eps=2e-4)
    assert len(mesh.leaf_levels()) > 1
    const = np.full(mesh.nref, 1.5)
    leaves = np.concatenate([mesh.index(l, mesh.cells[l]) for l in mesh.leaf_levels()])
    vol = np.concatenate([np.full(mesh.cells[l].size, cfg.cell_length(l) ** dim) for l in mesh.leaf_levels()])
    so.update_ghost_mr(mesh, u, bc)
    for out in (so.flux_linhom_apply(mesh, const, so.convection_upwind_coeffs([1.0, -0.5, 0.25][:dim])),
                so.flux_linhom_apply(mesh, const, so.diffusion_order2_coeffs([1.0, 2.0, 0.5][:dim]))):
        assert np.max(np.abs(out[leaves])) == 0.0
    # the disc does not touch the boundary: u = 0 there, so all boundary fluxes vanish and the interior telescopes
    for out in (so.flux_linhom_apply(mesh, u, so.diffusion_order2_coeffs([1.0] * dim)),
                so.flux_nonlin_apply(mesh, u, so.burgers_upwind_flux(0.5))):
        assert abs(np.sum(out[leaves] * vol)) < 1e-12 * max(1.0, np.max(np.abs(out[leaves])))


def test_oracle_matches_reference_mra_burgers_hat_golden():
    """demos/FiniteVolume/burgers_mra.cpp with the reference test's arguments (tests/test_demo_finite_volume.py:191-207:
    --nfiles=1 --min-level=2 --max-level=9 --init-sol=hat --mr-eps=1e-5): 1D, box [-2, 3], max_stencil_radius 2 (the library's
    default ghost width: further-ghost polynomial extrapolation, contiguous-boundary graduation rule), graduation width 2,
    Dirichlet<1>(0), regularity 0, `unp1 = u - dt * scheme(u)` with scheme = 0.5 * make_convection_upwind<Field>() (the NON-LINEAR
    flux-based scheme), cfl 0.95, Tf 0.1.  Against the reference's own test_finite_volume_demo_mra_burgers_hat.h5
    (tests/golden/mra_burgers_hat.npz): mesh identical, values within 1e-15."""
    g = np.load(os.path.join(GOLD, "mra_burgers_hat.npz"))

    def hat(x):  # hat_exact_solution(x, 0), burgers_mra.cpp:17-47
        out = np.zeros_like(x)
        m1 = (x > -1) & (x < 0)
        out[m1] = (1.0 / (0.0 - -1.0)) * (x[m1] - -1.0)
        m2 = (x >= 0) & (x < 1)
        out[m2] = (-1.0 / (1.0 - 0.0)) * (x[m2] - 0.0) + 1.0
        return out

    cfg = so.MeshConfig(dim=1, min_level=2, max_level=9, pred_radius=1, max_stencil_radius=2, graduation_width=2, origin=(-2.0,), scaling=5.0)
    bc = so.Bc("dirichlet", 0.0)
    mesh = so.Mesh.uniform(cfg)
    L = cfg.max_level
    u = np.zeros(mesh.nref)
    u[mesh.index(L, mesh.cells[L])] = hat(mesh.cell_centers(L, mesh.cells[L])[:, 0])
    eps, reg = 1e-5, 0.0
    dt = 0.95 * cfg.cell_length(L)
    Tf = 0.1
    mesh, u = so.adapt(mesh, u, bc, eps, reg)
    flux = so.burgers_upwind_flux(0.5)
    t, nt = 0.0, 0
    while t != Tf:
        t += dt
        if t > Tf:
            dt += Tf - t
            t = Tf
        mesh, u = so.adapt(mesh, u, bc, eps, reg)
        so.update_ghost_mr(mesh, u, bc)
        rhs = so.flux_nonlin_apply(mesh, u, flux)
        unp1 = np.full(mesh.nref, np.nan)
        for l in mesh.leaf_levels():
            i = mesh.index(l, mesh.cells[l])
            unp1[i] = u[i] - dt * rhs[i]
        u = unp1
        nt += 1
    assert nt == 11
    lv, co, ix = mesh.leaf_table()
    assert lv.size == g["level"].size and np.array_equal(lv, g["level"]) and np.array_equal(co, g["idx"]), "mesh differs from the reference golden"
    assert np.max(np.abs(u[ix] - g["u"])) <= 1e-15


def test_oracle_matches_reference_linear_convection_weno5_golden():
    """demos/FiniteVolume/linear_convection.cpp with the reference test's arguments (tests/test_demo_finite_volume.py:280-298, explicit:
    --min-level=1 --max-level=6 --Tf=0.1): 2D box [-1, 1]^2 periodic in both directions, max_stencil_size(6) (ghost width 3),
    `make_convection_weno5` (NON-LINEAR flux scheme with a six-cell line stencil, level jumps through predicted fine
    ghosts, interfaces through the periodic boundary), TVD-RK3 field expressions, default mra_config.  Against the reference's own
    test_finite_volume_demo_linear_convection_explicit.h5 (tests/golden/linear_convection_explicit.npz): mesh identical, values
    within 1e-13 (the reference's own tolerance is rel 1e-14 / abs 1e-7, tests/conftest.py:121-122)."""
    g = np.load(os.path.join(GOLD, "linear_convection_explicit.npz"))
    cfg = so.MeshConfig(dim=2, min_level=1, max_level=6, pred_radius=1, max_stencil_radius=3, graduation_width=1, origin=(-1.0, -1.0),
                        scaling=2.0, periodic=(True, True))
    r = so.run_linear_convection(cfg, Tf=0.1)
    assert r["steps"] == 7
    mesh, u = r["final"]
    lv, co, ix = mesh.leaf_table()
    assert lv.size == g["level"].size and np.array_equal(lv, g["level"]) and np.array_equal(co, g["idx"]), "mesh differs from the reference golden"
    assert np.max(np.abs(u[ix] - g["u"])) <= 1e-13
