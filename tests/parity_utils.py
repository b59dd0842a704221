"""Shared helpers for the parity tests: product (C ABI via samurai_b200) vs oracle (oracle/samurai_oracle.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import samurai_b200 as sb  # noqa: E402
import samurai_oracle as so  # noqa: E402

REL_TOL = 1e-12  # BASELINE.json north_star: fields within 1e-12 relative in fp64


def oracle_cfg(dim, min_level, max_level, pred_radius):
    return so.MeshConfig(dim=dim, min_level=min_level, max_level=max_level, pred_radius=pred_radius)


def product_cfg(dim, min_level, max_level, pred_radius):
    return (sb.mesh_config(dim, pred_radius).min_level(min_level).max_level(max_level).max_stencil_size(2).disable_minimal_ghost_width())


def oracle_sub(omesh, mesh_id):
    return {sb.CELLS: omesh.cells, sb.CELLS_AND_GHOSTS: omesh.cag, sb.PROJ_CELLS: omesh.proj, sb.UNION_CELLS: omesh.union,
            sb.REFERENCE: omesh.ref}[mesh_id]


def product_keys(pmesh, mesh_id, level):
    """(sorted packed keys, offsets) of one product sub-mesh level, in the oracle's key encoding."""
    iv = pmesh.intervals(mesh_id, level)
    if iv.size == 0:
        return so.EMPTY, np.zeros(0, np.int64)
    n = (iv["end"] - iv["start"]).astype(np.int64)
    rep = np.repeat(np.arange(iv.size), n)
    k = np.arange(int(n.sum())) - np.repeat(np.cumsum(n) - n, n)
    cols = [iv["start"][rep] + k, iv["y"][rep], iv["z"][rep]][: pmesh.dim]
    keys = so.pack(np.stack(cols, axis=1))
    return keys, iv["offset"][rep] + k


def assert_same_mesh(pmesh, omesh, ids=(sb.CELLS, sb.CELLS_AND_GHOSTS, sb.PROJ_CELLS, sb.REFERENCE), check_offsets=True):
    """Interval lists / cell sets bit-exact, and storage offsets identical (renumbering parity)."""
    for mesh_id in ids:
        osub = oracle_sub(omesh, mesh_id)
        for level in range(omesh.nlev):
            keys, off = product_keys(pmesh, mesh_id, level)
            assert np.all(np.diff(keys) > 0), f"product mesh {mesh_id} level {level} not sorted/unique"
            okeys = osub[level]
            assert keys.size == okeys.size and np.array_equal(keys, okeys), (
                f"mesh id {mesh_id} level {level}: product {keys.size} cells vs oracle {okeys.size}")
            if check_offsets and keys.size:
                assert np.array_equal(off, omesh.index(level, keys)), f"storage offsets differ: mesh id {mesh_id} level {level}"
    assert pmesh.nb_cells(sb.REFERENCE) == omesh.nref


def max_rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0)))


def assert_fields_close(a, b, what, tol=REL_TOL, finite_only=False):
    a = np.asarray(a)
    b = np.asarray(b)
    if finite_only:
        m = np.isfinite(b)
        a, b = a[m], b[m]
    assert np.all(np.isfinite(a)), f"{what}: non-finite values in the product result"
    err = max_rel_err(a, b)
    assert err <= tol, f"{what}: max relative error {err:.3e} > {tol:.1e}"
    return err


def run_advection_parity(dim=2, min_level=2, max_level=6, pred_radius=1, steps=3, eps=2e-4, regularity=1.0, device=0, verbose=False,
                         scheme="upwind", relative_detail=False, amplitude=1.0):
    """demos/FiniteVolume/advection_2d.cpp time loop on the GPU, checked against the oracle at every step:
    meshes bit-identical (cells + all ghosts + storage offsets), fields within 1e-12 relative."""
    if not sb.initialize(device):
        raise sb.SamuraiError("a CUDA device is required")
    ocfg = oracle_cfg(dim, min_level, max_level, pred_radius)
    bc = so.Bc("dirichlet", 0.0)
    omesh = so.Mesh.uniform(ocfg)
    center, radius = [0.3] * dim, 0.2
    ou = so.init_disc(omesh, center, radius) * amplitude

    pmesh = sb.MRMesh.make_mesh([0.0] * dim, [1.0] * dim, product_cfg(dim, min_level, max_level, pred_radius))
    assert_same_mesh(pmesh, omesh)
    u = sb.make_scalar_field("u", pmesh)
    u.resize()
    u.upload(ou)
    sb.make_bc(u, sb.DIRICHLET, 0.0)
    unp1 = sb.make_scalar_field("unp1", pmesh)
    adapt = sb.make_MRAdapt(u)
    mcfg = sb.mra_config().epsilon(eps).regularity(regularity).relative_detail(relative_detail)
    a = [1.0] * dim
    dt = 0.5 * pmesh.min_cell_length() if dim == 2 else 0.25 * pmesh.min_cell_length()

    def check(tag):
        assert_same_mesh(pmesh, omesh)
        _, _, leaf_idx = omesh.leaf_table()
        pu = u.download()
        err = assert_fields_close(pu[leaf_idx], ou[leaf_idx], tag)
        if verbose:
            print(f"{tag}: leaves {omesh.nb_cells()} ref {omesh.nref} max rel err {err:.2e}")

    adapt(mcfg)
    omesh, ou = so.adapt(omesh, ou, bc, eps, regularity, relative_detail=relative_detail)
    check("initial adaptation")
    for it in range(steps):
        adapt(mcfg)
        omesh, ou = so.adapt(omesh, ou, bc, eps, regularity, relative_detail=relative_detail)
        sb.update_ghost_mr(u)
        so.update_ghost_mr(omesh, ou, bc)
        # after the ghost update every reference cell the oracle defines must agree
        pu = u.download()
        assert_fields_close(pu, ou, f"step {it} ghosts", finite_only=True)
        unp1.resize()
        if scheme == "upwind":
            sb.upwind_step(unp1, u, a, dt)
        else:
            sb.upwind_scalar_burgers_step(unp1, u, a, dt)
        ou = so.fv_step(omesh, ou, a, dt, scheme=scheme)
        sb.swap(u, unp1)
        check(f"step {it}")
    u.destroy()
    unp1.destroy()
    pmesh.destroy()
