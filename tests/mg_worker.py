"""Multi-rank worker (run under torchrun). Modes:
  gpu   : one process per GPU; the advection loop on N GPUs must equal the oracle (== the 1-GPU result) bit for bit:
          meshes identical on every rank, leaf values gathered from their owners identical to the oracle's.
  host  : gloo, no GPU; the slab partition derived independently by every rank is the same and is a partition.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    mode = sys.argv[1]
    dim = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    lmin, lmax = (2, 7) if dim == 2 else (1, 5)
    if len(sys.argv) > 4:
        lmin, lmax = int(sys.argv[3]), int(sys.argv[4])
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    if mode == "gpu":
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo")
    import parity_utils as pu
    sb, so = pu.sb, pu.so

    def gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    have = sb.initialize_multi(rank, world, device=local if mode == "gpu" else -1, pool_bytes=1 << 30)
    assert have == (mode == "gpu")
    pmesh = sb.MRMesh.make_mesh([0.0] * dim, [1.0] * dim, pu.product_cfg(dim, lmin, lmax, 1))
    ocfg = pu.oracle_cfg(dim, lmin, lmax, 1)
    bc = so.Bc("dirichlet", 0.0)
    omesh = so.Mesh.uniform(ocfg)
    ou = so.init_disc(omesh, [0.3] * dim, 0.2)

    def check_partition():
        owners = sb.mg_leaf_owners(pmesh)
        allo = gather(owners)
        for o in allo:
            assert np.array_equal(o, allo[0]), "ranks disagree on the partition"
        assert owners.min() >= 0 and owners.max() < world
        counts = np.bincount(owners, minlength=world)
        return owners, counts

    if mode == "host":
        owners, counts = check_partition()
        assert counts.sum() == pmesh.nb_cells()
        # uniform mesh: slabs balanced to within one row of cells
        assert counts.max() - counts.min() <= (1 << (lmax * (dim - 1))), counts
        # drive the mesh with the oracle's tags: still identical and a partition on every rank
        omesh2, ou2 = so.adapt(omesh, ou, bc, 2e-4, 1.0, trace=(trace := []))
        for t in trace:
            pmesh.update_from_tags(t["tag"])
        pu.assert_same_mesh(pmesh, omesh2)
        owners, counts = check_partition()
        assert counts.sum() == omesh2.nb_cells()
        if rank == 0:
            print("host partition OK", counts.tolist())
        dist.barrier()
        dist.destroy_process_group()
        return

    u = sb.make_scalar_field("u", pmesh)
    u.resize()
    u.upload(ou)
    sb.make_bc(u, sb.DIRICHLET, 0.0)
    unp1 = sb.make_scalar_field("unp1", pmesh)
    adapt = sb.make_MRAdapt(u)
    mcfg = sb.mra_config().epsilon(2e-4)
    a = [1.0] * dim
    dt = (0.5 if dim == 2 else 0.25) * pmesh.min_cell_length()

    def check(tag):
        pu.assert_same_mesh(pmesh, omesh)
        owners, counts = check_partition()
        _, _, leaf_idx = omesh.leaf_table()
        mine = u.download()[leaf_idx]
        parts = gather((owners == rank, mine[owners == rank]))
        full = np.empty(leaf_idx.size)
        for m, v in parts:
            full[m] = v
        err = pu.max_rel_err(full, ou[leaf_idx])
        assert err == 0.0, f"{tag}: N-GPU result differs from the oracle, max rel err {err:.3e}"
        # after a broadcast every rank holds the complete field
        sb.mg_broadcast(u)
        allv = u.download()[leaf_idx]
        assert np.array_equal(allv, ou[leaf_idx]), f"{tag}: broadcast copy incomplete on rank {rank}"
        if rank == 0:
            print(f"{tag}: leaves {omesh.nb_cells()} per-rank {counts.tolist()} OK", flush=True)

    adapt(mcfg)
    omesh, ou = so.adapt(omesh, ou, bc, 2e-4, 1.0)
    check("initial adaptation")
    sb.mg_rebalance(u)
    for it in range(4):
        adapt(mcfg)
        omesh, ou = so.adapt(omesh, ou, bc, 2e-4, 1.0)
        sb.update_ghost_mr(u)
        so.update_ghost_mr(omesh, ou, bc)
        unp1.resize()
        sb.upwind_step(unp1, u, a, dt)
        ou = so.fv_step(omesh, ou, a, dt)
        sb.swap(u, unp1)
        check(f"step {it}")
        if it == 1:
            sb.mg_rebalance(u)
    # flux-based schemes on the adapted mesh (level jumps across the slab cuts): each rank computes its slab's leaves, every owner's
    # values are bit-equal to the oracle's
    owners, counts = check_partition()
    _, _, leaf_idx = omesh.leaf_table()
    og = ou.copy()
    so.update_ghost_mr(omesh, og, bc)
    vel, K = [1.0, -0.5, 0.25][:dim], [1.0, 2.0, 0.5][:dim]
    for scheme, want in ((sb.make_diffusion_order2(K), so.flux_linhom_apply(omesh, og, so.diffusion_order2_coeffs(K))),
                         (sb.make_convection_upwind(vel), so.flux_linhom_apply(omesh, og, so.convection_upwind_coeffs(vel))),
                         (0.5 * sb.make_convection_upwind(), so.flux_nonlin_apply(omesh, og, so.burgers_upwind_flux(0.5)))):
        rhs = scheme(u)
        mine = rhs.download()[leaf_idx]
        parts = gather((owners == rank, mine[owners == rank]))
        full = np.empty(leaf_idx.size)
        for m, v in parts:
            full[m] = v
        bad = np.count_nonzero(full != want[leaf_idx])
        assert bad == 0, f"{scheme.name}: {bad} of {leaf_idx.size} owner leaves differ from the oracle on {world} GPUs"
        rhs.destroy()
    if rank == 0:
        print("flux schemes on", world, "GPUs OK", flush=True)
    # WENO5 (six-cell stencils, ghost width 3) on a fully periodic adapted mesh across the slab cuts, incl. the interfaces through the
    # periodic boundary between the first and the last slab
    if dim < 3:
        wl, wh = (2, 8) if dim == 1 else (1, 6)
        pcfg = sb.mesh_config(dim, 1).min_level(wl).max_level(wh).periodic([True] * dim).max_stencil_size(6)
        wcfg = so.MeshConfig(dim=dim, min_level=wl, max_level=wh, pred_radius=1, max_stencil_radius=3, graduation_width=1,
                             origin=(-1.0,) * dim, scaling=2.0, periodic=(True,) * dim)
        wmesh = sb.MRMesh.make_mesh([-1.0] * dim, [1.0] * dim, pcfg)
        wom = so.Mesh.uniform(wcfg)
        c = wom.cell_centers(wh, wom.cells[wh])
        f0 = np.zeros(wom.nref)
        f0[wom.index(wh, wom.cells[wh])] = np.where(np.all((c >= -1.0) & (c <= -0.45), axis=1), 1.0, 0.0)
        wu = sb.make_scalar_field("wu", wmesh)
        wu.resize()
        wu.upload(f0)
        sb.make_MRAdapt(wu)(sb.mra_config())
        wbc = so.Bc("neumann", 0.0)
        wom, wou = so.adapt(wom, f0, wbc, 1e-4, 1.0)
        pu.assert_same_mesh(wmesh, wom)
        wowners = sb.mg_leaf_owners(wmesh)
        _, _, wleaf = wom.leaf_table()
        so.update_ghost_mr(wom, wou, wbc)
        vel = [1.0, -1.0][:dim]
        want = so.flux_nonlin_apply(wom, wou, so.weno5_flux(vel), so.WENO5_OFFSETS)
        rhs = sb.make_convection_weno5(vel)(wu)
        mine = rhs.download()[wleaf]
        parts = gather((wowners == rank, mine[wowners == rank]))
        full = np.empty(wleaf.size)
        for m, v in parts:
            full[m] = v
        bad = np.count_nonzero(full != want[wleaf])
        assert bad == 0, f"WENO5: {bad} of {wleaf.size} owner leaves differ from the oracle on {world} GPUs"
        if rank == 0:
            print("WENO5 on", world, "GPUs OK", flush=True)
        rhs.destroy()
        # the advection loop on the same periodic mesh (ghost width 3): periodic ghost copies, tag OR through the boundary and
        # graduation across it, with the first and the last slab as neighbours
        wunp1 = sb.make_scalar_field("wunp1", wmesh)
        wadapt = sb.make_MRAdapt(wu)
        wdt = 0.5 * wmesh.min_cell_length()
        wa = [1.0, -1.0][:dim]
        wou = wou.copy()
        for it in range(3):
            wadapt(sb.mra_config())
            wom, wou = so.adapt(wom, wou, wbc, 1e-4, 1.0)
            pu.assert_same_mesh(wmesh, wom)
            sb.update_ghost_mr(wu)
            so.update_ghost_mr(wom, wou, wbc)
            wunp1.resize()
            sb.upwind_step(wunp1, wu, wa, wdt)
            wou = so.fv_step(wom, wou, wa, wdt)
            sb.swap(wu, wunp1)
            wowners = sb.mg_leaf_owners(wmesh)
            _, _, wleaf = wom.leaf_table()
            mine = wu.download()[wleaf]
            parts = gather((wowners == rank, mine[wowners == rank]))
            full = np.empty(wleaf.size)
            for m, v in parts:
                full[m] = v
            bad = np.count_nonzero(full != wou[wleaf])
            assert bad == 0, f"periodic advection step {it}: {bad} of {wleaf.size} owner leaves differ from the oracle on {world} GPUs"
        if rank == 0:
            print("periodic advection on", world, "GPUs OK", flush=True)
    st = sb.stats()
    if rank == 0:
        print("multi-GPU parity OK; launches", st["kernel_launches"], flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
