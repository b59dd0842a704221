"""GPU parity tests proper: the CUDA path, through the C ABI, against the oracle on the same inputs.
Bar (BASELINE.json north_star): meshes (cells, ghosts, storage offsets) and tags bit-exact at every step; fields
within 1e-12 relative in fp64."""
import numpy as np
import pytest

import parity_utils as pu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim,lmin,lmax,pred,steps", [(2, 2, 6, 1, 4), (2, 2, 6, 0, 4), (2, 4, 8, 1, 3), (3, 1, 4, 1, 3), (3, 2, 5, 0, 2)])
def test_advection_loop_matches_oracle(gpu, dim, lmin, lmax, pred, steps):
    pu.run_advection_parity(dim=dim, min_level=lmin, max_level=lmax, pred_radius=pred, steps=steps)


def test_burgers_loop_matches_oracle(gpu):
    pu.run_advection_parity(dim=2, min_level=2, max_level=7, pred_radius=1, steps=4, scheme="burgers")


def test_relative_detail_matches_oracle(gpu):
    """mra_config().relative_detail(true) (mr/rel_detail.hpp:73-112): details normalised by max_leaves |u|; with an
    amplitude of 37.5 the absolute and relative criteria give different meshes, so the path is really exercised."""
    pu.run_advection_parity(dim=2, min_level=2, max_level=7, pred_radius=1, steps=3, relative_detail=True, amplitude=37.5)


@pytest.mark.parametrize("dim,lmin,lmax,pred", [(1, 2, 8, 1), (2, 2, 7, 1), (2, 3, 7, 0), (3, 1, 4, 1)])
def test_per_sweep_launches_match_oracle_too(gpu, dim, lmin, lmax, pred):
    """the default path runs every level wavefront as one cooperative launch (smr_set_fused); the one-launch-per-sweep
    path must give the same meshes and fields (it is what the multi-GPU runs and the per-family profile use)."""
    pu.sb.set_fused(False)
    try:
        pu.run_advection_parity(dim=dim, min_level=lmin, max_level=lmax, pred_radius=pred, steps=3)
    finally:
        pu.sb.set_fused(True)


@pytest.mark.parametrize("refine_boundary", [False, True])
def test_fused_and_per_sweep_paths_are_bit_identical(gpu, refine_boundary):
    """same run twice, fused and per-sweep: every reference cell (leaves AND ghosts) and every tag byte identical
    (also with `--refine-boundary`, whose keep phase sits between the criteria and the keep propagation in both paths)"""
    sb = pu.sb
    outs = []
    for fused in (True, False):
        sb.set_fused(fused)
        cfg = pu.product_cfg(2, 3, 9, 1)
        if refine_boundary:
            cfg = cfg.refine_boundary()
        mesh = sb.MRMesh.make_mesh([0.0, 0.0], [1.0, 1.0], cfg)
        u = sb.make_scalar_field("u", mesh)
        u.resize()
        u.init_ball([0.3, 0.3], 0.2)
        sb.make_bc(u, sb.NEUMANN, 0.25)
        unp1 = sb.make_scalar_field("unp1", mesh)
        adapt = sb.make_MRAdapt(u)
        mra = sb.mra_config().epsilon(1e-4)
        adapt(mra)
        for _ in range(5):
            adapt(mra)
            sb.update_ghost_mr(u)
            unp1.resize()
            sb.upwind_step(unp1, u, [1.0, 0.5], 0.5 * mesh.min_cell_length())
            sb.swap(u, unp1)
        adapt(mra)
        sb.update_ghost_mr(u)
        outs.append((mesh.cell_table(sb.REFERENCE), u.download()))
        u.destroy()
        unp1.destroy()
        mesh.destroy()
    sb.set_fused(True)
    (ta, fa), (tb, fb) = outs
    assert all(np.array_equal(x, y) for x, y in zip(ta, tb))
    leaves_and_ghosts = ta[2]
    assert np.array_equal(fa[leaves_and_ghosts], fb[leaves_and_ghosts])


@pytest.mark.parametrize("dim,lmin,lmax", [(2, 2, 7), (3, 1, 4)])
def test_two_fields_adapted_together_match_oracle(gpu, dim, lmin, lmax):
    """make_MRAdapt(u, v): one mesh driven by the details of both fields (coarsen only where every field allows it, refine
    where any asks for it: mr/criteria.hpp loops over the components), different boundary conditions per field."""
    sb, so = pu.sb, pu.so
    ocfg = pu.oracle_cfg(dim, lmin, lmax, 1)
    omesh = so.Mesh.uniform(ocfg)
    ou = so.init_disc(omesh, [0.3] * dim, 0.2)
    ov = 0.5 * so.init_disc(omesh, [0.7] * dim, 0.15)
    bcs = [so.Bc("dirichlet", 0.0), so.Bc("neumann", 0.0)]
    pmesh = sb.MRMesh.make_mesh([0.0] * dim, [1.0] * dim, pu.product_cfg(dim, lmin, lmax, 1))
    u = sb.make_scalar_field("u", pmesh)
    v = sb.make_scalar_field("v", pmesh)
    for f, host in ((u, ou), (v, ov)):
        f.resize()
        f.upload(host)
    sb.make_bc(u, sb.DIRICHLET, 0.0)
    sb.make_bc(v, sb.NEUMANN, 0.0)
    unp1 = sb.make_scalar_field("unp1", pmesh)
    vnp1 = sb.make_scalar_field("vnp1", pmesh)
    adapt = sb.make_MRAdapt(u, v)
    mra = sb.mra_config().epsilon(2e-4)
    dt = (0.5 if dim == 2 else 0.25) * pmesh.min_cell_length()
    a, b = [1.0] * dim, [-1.0] * dim
    only_u = so.adapt(so.Mesh.uniform(ocfg), so.init_disc(omesh, [0.3] * dim, 0.2), bcs[0], 2e-4, 1.0)[0]
    for step in range(3):
        adapt(mra)
        omesh, (ou, ov) = so.adapt_fields(omesh, [ou, ov], bcs, 2e-4, 1.0)
        pu.assert_same_mesh(pmesh, omesh)
        if step == 0:
            assert omesh.nb_cells() > only_u.nb_cells(), "the second field must really influence the mesh"
        for f, nf, of, vel, bc in ((u, unp1, ou, a, bcs[0]), (v, vnp1, ov, b, bcs[1])):
            sb.update_ghost_mr(f)
            so.update_ghost_mr(omesh, of, bc)
            nf.resize()
            sb.upwind_step(nf, f, vel, dt)
        ou, ov = so.fv_step(omesh, ou, a, dt), so.fv_step(omesh, ov, b, dt)
        sb.swap(u, unp1)
        sb.swap(v, vnp1)
        _, _, leaf = omesh.leaf_table()
        pu.assert_fields_close(u.download()[leaf], ou[leaf], f"u step {step}")
        pu.assert_fields_close(v.download()[leaf], ov[leaf], f"v step {step}")


def test_vector_field_soa_components_match_oracle(gpu):
    """make_vector_field<double, 2>: two SoA device arrays ghost-updated, adapted (make_MRAdapt(vec): the criteria loop over
    the components, mr/criteria.hpp:33-85) and advanced together; host layout AoS [cell][comp] as in the reference."""
    sb, so = pu.sb, pu.so
    dim, lmin, lmax = 2, 2, 7
    omesh = so.Mesh.uniform(pu.oracle_cfg(dim, lmin, lmax, 1))
    o0 = so.init_disc(omesh, [0.3, 0.3], 0.2)
    o1 = -2.0 * so.init_disc(omesh, [0.6, 0.4], 0.1)
    bcs = [so.Bc("dirichlet", 0.0), so.Bc("dirichlet", 0.0)]
    pmesh = sb.MRMesh.make_mesh([0.0] * dim, [1.0] * dim, pu.product_cfg(dim, lmin, lmax, 1))
    vec = sb.make_vector_field("v", pmesh, 2)
    vnp1 = sb.make_vector_field("vnp1", pmesh, 2)
    vec.resize()
    vec.upload(np.stack([o0, o1], axis=1))
    sb.make_bc(vec, sb.DIRICHLET, 0.0, 0.0)
    adapt = sb.make_MRAdapt(vec)
    mra = sb.mra_config().epsilon(2e-4)
    a, dt = [1.0, 0.5], 0.5 * pmesh.min_cell_length()
    for step in range(3):
        adapt(mra)
        omesh, (o0, o1) = so.adapt_fields(omesh, [o0, o1], bcs, 2e-4, 1.0)
        pu.assert_same_mesh(pmesh, omesh)
        sb.update_ghost_mr(vec)
        for of in (o0, o1):
            so.update_ghost_mr(omesh, of, bcs[0])
        vnp1.resize()
        sb.upwind_step(vnp1, vec, a, dt)
        sb.swap(vec, vnp1)
        o0, o1 = so.fv_step(omesh, o0, a, dt), so.fv_step(omesh, o1, a, dt)
        _, _, leaf = omesh.leaf_table()
        got = vec.download()
        assert got.shape == (omesh.nref, 2)
        pu.assert_fields_close(got[leaf, 0], o0[leaf], f"component 0 step {step}")
        pu.assert_fields_close(got[leaf, 1], o1[leaf], f"component 1 step {step}")


@pytest.mark.parametrize("dim,min_level,max_level,msr,steps", [(1, 2, 8, 1, 6), (2, 2, 6, 1, 5), (2, 2, 6, 2, 5), (3, 1, 4, 1, 3)])
def test_refine_boundary_matches_oracle(gpu, dim, min_level, max_level, msr, steps):
    """`--refine-boundary` (arguments.hpp:67): keep_boundary_refined after the criteria of every harten iteration (mr/adapt.hpp:245-274,
    340-345) keeps the max_level leaves within max_stencil_radius cells of the boundary.  Tags and details of every harten iteration,
    all sub-meshes and the leaves at every step against the oracle."""
    r = pu.run_advection_parity(dim=dim, min_level=min_level, max_level=max_level, pred_radius=1, steps=steps, msr=msr, trace_tags=True,
                                refine_boundary=True, cfl=0.5 if dim < 3 else 0.25)
    assert r["max_rel_err"] <= pu.REL_TOL
    # the boundary stays refined: more leaves than without the flag
    r0 = pu.run_advection_parity(dim=dim, min_level=min_level, max_level=max_level, pred_radius=1, steps=steps, msr=msr,
                                 cfl=0.5 if dim < 3 else 0.25)
    assert r["leaves"] > r0["leaves"]
