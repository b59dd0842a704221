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


def oracle_cfg(dim, min_level, max_level, pred_radius, periodic=None, msr=1):
    return so.MeshConfig(dim=dim, min_level=min_level, max_level=max_level, pred_radius=pred_radius, periodic=periodic, max_stencil_radius=msr)


def product_cfg(dim, min_level, max_level, pred_radius, periodic=None, msr=1):
    cfg = sb.mesh_config(dim, pred_radius).min_level(min_level).max_level(max_level)
    if msr == 1:
        cfg = cfg.max_stencil_size(2).disable_minimal_ghost_width()
    if periodic is not None:
        cfg = cfg.periodic(list(periodic))
    return cfg


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


def init_square(omesh, half=0.1):
    """README.md:111-118: u = 1 where every |center_d - 0.5| <= 0.5 * 0.2."""
    u = np.zeros(omesh.nref)
    for l in omesh.leaf_levels():
        k = omesh.cells[l]
        c = omesh.cell_centers(l, k)
        u[omesh.index(l, k)] = np.where(np.all(np.abs(c - 0.5) <= half, axis=1), 1.0, 0.0)
    return u


def adapt_both(adapt, mcfg, pmesh, omesh, ou, bc, eps, regularity, relative_detail=False, trace_tags=False, what="", refine_boundary=False):
    """One MRadaptation call on both sides.  With trace_tags the product runs iteration by iteration (smr_adapt_iteration)
    and the tag array (uint8, every reference cell) and the detail array (fp64, every reference cell) of EVERY harten
    iteration are compared with the oracle's: tags bit-exact (north_star), details within REL_TOL (observed: bit-exact)."""
    if not trace_tags:
        adapt(mcfg)
        return so.adapt(omesh, ou, bc, eps, regularity, relative_detail=relative_detail, refine_boundary=refine_boundary)
    trace = []
    omesh2, ou2 = so.adapt(omesh, ou, bc, eps, regularity, trace=trace, relative_detail=relative_detail, refine_boundary=refine_boundary)
    cfg = omesh.cfg
    n_ite = 0
    for ite in range(cfg.max_level - cfg.min_level):
        unchanged = adapt.iteration(mcfg, ite)
        n_ite += 1
        assert ite < len(trace), f"{what}: product ran more harten iterations than the oracle ({len(trace)})"
        ref = trace[ite]
        tags = adapt.last_tags()
        assert tags.size == ref["tag"].size, f"{what} ite {ite}: tag array size {tags.size} vs {ref['tag'].size}"
        bad = np.flatnonzero(tags != ref["tag"])
        assert bad.size == 0, f"{what} ite {ite}: {bad.size} tag bytes differ (first at {bad[:5]}: {tags[bad[:5]]} vs {ref['tag'][bad[:5]]})"
        det = adapt.last_detail()
        assert det.size == ref["detail"].size
        assert_fields_close(det, ref["detail"], f"{what} ite {ite} detail")
        if unchanged:
            break
    assert n_ite == len(trace), f"{what}: {n_ite} harten iterations vs {len(trace)} in the oracle"
    return omesh2, ou2


def run_advection_parity(dim=2, min_level=2, max_level=6, pred_radius=1, steps=3, eps=2e-4, regularity=1.0, device=0, verbose=False,
                         scheme="upwind", relative_detail=False, amplitude=1.0, init="disc", trace_tags=False, a=None, cfl=None,
                         discs=None, check_ghosts=True, periodic=None, msr=1, refine_boundary=False):
    """demos/FiniteVolume/advection_2d.cpp time loop on the GPU, checked against the oracle at every step:
    meshes bit-identical (cells + all ghosts + storage offsets), fields within 1e-12 relative; with trace_tags also the
    tag and detail arrays of every harten iteration.  `discs`: [(center, radius, value)] initial condition
    (scalar_burgers_2d.cpp:20-50); init="square": README.md:111-118."""
    if not sb.initialize(device):
        raise sb.SamuraiError("a CUDA device is required")
    ocfg = oracle_cfg(dim, min_level, max_level, pred_radius, periodic, msr)
    bc = so.Bc("dirichlet", 0.0)
    omesh = so.Mesh.uniform(ocfg)
    if discs is not None:
        ou = np.zeros(omesh.nref)
        for center, radius, value in discs:
            m = so.init_disc(omesh, center, radius)
            ou = np.where(m != 0, value, ou)
    elif init == "square":
        ou = init_square(omesh) * amplitude
    else:
        ou = so.init_disc(omesh, [0.3] * dim, 0.2) * amplitude

    pcfg = product_cfg(dim, min_level, max_level, pred_radius, periodic, msr)
    if refine_boundary:
        pcfg = pcfg.refine_boundary()
    pmesh = sb.MRMesh.make_mesh([0.0] * dim, [1.0] * dim, pcfg)
    assert_same_mesh(pmesh, omesh)
    u = sb.make_scalar_field("u", pmesh)
    u.resize()
    u.upload(ou)
    if periodic is None or not all(periodic):
        sb.make_bc(u, sb.DIRICHLET, 0.0)
    unp1 = sb.make_scalar_field("unp1", pmesh)
    adapt = sb.make_MRAdapt(u)
    mcfg = sb.mra_config().epsilon(eps).regularity(regularity).relative_detail(relative_detail)
    a = [1.0] * dim if a is None else list(a)
    if cfl is None:
        cfl = 0.5 if dim == 2 else 0.25
    dt = cfl * pmesh.min_cell_length()
    worst = 0.0

    def check(tag):
        nonlocal worst
        assert_same_mesh(pmesh, omesh)
        _, _, leaf_idx = omesh.leaf_table()
        pu = u.download()
        err = assert_fields_close(pu[leaf_idx], ou[leaf_idx], tag)
        worst = max(worst, err)
        if verbose:
            print(f"{tag}: leaves {omesh.nb_cells()} ref {omesh.nref} max rel err {err:.2e}")

    omesh, ou = adapt_both(adapt, mcfg, pmesh, omesh, ou, bc, eps, regularity, relative_detail, trace_tags and not relative_detail, "initial adaptation",
                           refine_boundary)
    check("initial adaptation")
    for it in range(steps):
        omesh, ou = adapt_both(adapt, mcfg, pmesh, omesh, ou, bc, eps, regularity, relative_detail, trace_tags and not relative_detail, f"step {it}",
                               refine_boundary)
        sb.update_ghost_mr(u)
        so.update_ghost_mr(omesh, ou, bc)
        if check_ghosts:
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
    leaves = omesh.nb_cells()
    u.destroy()
    unp1.destroy()
    pmesh.destroy()
    return dict(leaves=leaves, max_rel_err=worst)
