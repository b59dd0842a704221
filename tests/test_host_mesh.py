"""Host half of the product (C ABI, no GPU needed) against the oracle: sub-meshes and storage numbering bit-exact,
update_cell_array_from_tag + make_graduation + mesh rebuild driven by the oracle's tag arrays."""
import numpy as np
import pytest

import parity_utils as pu

sb, so = pu.sb, pu.so


@pytest.fixture(scope="module", autouse=True)
def _init(lib):
    lib.initialize(-1)


@pytest.mark.parametrize("dim,lmin,lmax,pred", [(1, 1, 6, 1), (2, 2, 6, 0), (2, 2, 6, 1), (2, 4, 8, 1), (3, 1, 4, 0), (3, 1, 4, 1)])
def test_uniform_mesh_matches_oracle(dim, lmin, lmax, pred):
    pm = sb.MRMesh.make_mesh([0.0] * dim, [1.0] * dim, pu.product_cfg(dim, lmin, lmax, pred))
    om = so.Mesh.uniform(pu.oracle_cfg(dim, lmin, lmax, pred))
    pu.assert_same_mesh(pm, om)
    assert pm.nb_cells(sb.CELLS) == (1 << (dim * lmax))
    pm.destroy()


def test_single_level_mesh_has_no_projection_cells():
    pm = sb.MRMesh.make_mesh([0.0, 0.0], [1.0, 1.0], pu.product_cfg(2, 5, 5, 1))
    om = so.Mesh.uniform(pu.oracle_cfg(2, 5, 5, 1))
    pu.assert_same_mesh(pm, om)
    assert pm.nb_cells(sb.PROJ_CELLS) == 0
    assert pm.nb_cells(sb.REFERENCE) == (32 + 2) ** 2
    pm.destroy()


@pytest.mark.parametrize("dim,lmin,lmax,pred,steps", [(2, 2, 7, 1, 4), (2, 2, 7, 0, 4), (3, 1, 5, 1, 3)])
def test_adaptation_host_path_matches_oracle(dim, lmin, lmax, pred, steps):
    """Every harten iteration of the oracle's advection run: feed its tag array to the product's host path and require
    the same graduated mesh (all sub-meshes + offsets), including the fixed-point detection."""
    ocfg = pu.oracle_cfg(dim, lmin, lmax, pred)
    bc = so.Bc()
    om = so.Mesh.uniform(ocfg)
    ou = so.init_disc(om, [0.3] * dim, 0.2)
    pm = sb.MRMesh.make_mesh([0.0] * dim, [1.0] * dim, pu.product_cfg(dim, lmin, lmax, pred))
    n_checked = 0
    for _ in range(steps):
        trace = []
        om2, ou2 = so.adapt(om, ou, bc, 2e-4, 1.0, trace=trace)
        for k, t in enumerate(trace):
            pu.assert_same_mesh(pm, t["mesh"])
            gen = pm.generation()
            unchanged = pm.update_from_tags(t["tag"])
            last = k == len(trace) - 1
            if unchanged:
                assert last and pm.generation() == gen
            n_checked += 1
        pu.assert_same_mesh(pm, om2)
        om, ou = om2, ou2
        so.update_ghost_mr(om, ou, bc)
        ou = so.fv_step(om, ou, [1.0] * dim, 0.5 * ocfg.cell_length(lmax))
    assert n_checked >= steps
    pm.destroy()


def test_from_intervals_roundtrip_and_get_index():
    ocfg = pu.oracle_cfg(2, 2, 6, 1)
    om = so.Mesh.uniform(ocfg)
    ou = so.init_disc(om, [0.3, 0.3], 0.2)
    om, ou = so.adapt(om, ou, so.Bc(), 2e-4, 1.0)
    lv, co, ix = om.leaf_table()
    iv = np.zeros(lv.size, dtype=sb.INTERVAL_DTYPE)
    iv["start"], iv["end"], iv["y"] = co[:, 0], co[:, 0] + 1, co[:, 1]
    pm = sb.MRMesh.from_intervals([0, 0], [1, 1], pu.product_cfg(2, 2, 6, 1), lv.astype(np.int32), iv)
    pu.assert_same_mesh(pm, om)
    # get_index == oracle index for a few leaves; missing cells raise like LevelCellArray::get_interval
    for k in (0, lv.size // 2, lv.size - 1):
        assert pm.get_index(int(lv[k]), int(co[k, 0]), int(co[k, 1])) == ix[k]
    with pytest.raises(IndexError):
        pm.get_index(6, 1000, 1000)
    # leaf table order == for_each_cell order
    plv, pco, poff = pm.cell_table(sb.CELLS)
    assert np.array_equal(plv, lv) and np.array_equal(pco, co) and np.array_equal(poff, ix)
    pm.destroy()


def test_invalid_configs_raise():
    with pytest.raises(ValueError):
        sb.MRMesh.make_mesh([0, 0], [1, 1], sb.mesh_config(2, 1).min_level(5).max_level(3).disable_minimal_ghost_width())
    with pytest.raises(ValueError):  # ghost width 3 (WENO5 stencils) is not implemented: must fail loudly, not silently differ
        sb.MRMesh.make_mesh([0, 0], [1, 1], sb.mesh_config(2, 1).min_level(2).max_level(4).max_stencil_radius(3))


@pytest.mark.parametrize("dim,lmin,lmax", [(1, 2, 8), (2, 2, 6), (3, 1, 4)])
def test_default_ghost_width_2_mesh_matches_oracle(dim, lmin, lmax):
    """the library's default mesh_config (max_stencil_radius 2, mesh_config.hpp:388-393) on the host: sub-meshes, storage offsets and the
    graduation with the contiguous-boundary rule (graduation.hpp:372-455), driven by the oracle's tag arrays"""
    ocfg = pu.oracle_cfg(dim, lmin, lmax, 1, None, 2)
    om = so.Mesh.uniform(ocfg)
    pm = sb.MRMesh.make_mesh([0.0] * dim, [1.0] * dim, pu.product_cfg(dim, lmin, lmax, 1, None, 2))
    pu.assert_same_mesh(pm, om)
    trace = []
    om2, _ = so.adapt(om, so.init_disc(om, [0.3] * dim, 0.2), so.Bc("dirichlet", 0.0), 2e-4, 1.0, trace=trace)
    for t in trace:
        pm.update_from_tags(t["tag"])
    pu.assert_same_mesh(pm, om2)
    assert len(om2.leaf_levels()) > 1


@pytest.mark.parametrize("dim,lmin,lmax", [(1, 2, 7), (2, 2, 6), (3, 1, 4)])
def test_flux_face_classification_matches_interface_sets(dim, lmin, lmax):
    """Host half of the flux-based schemes (csrc/batches.hpp: flux_items) against the reference's interface sets restated
    by the oracle (interface.hpp:35-306, boundary.hpp:6-33): the records tile the leaves exactly, and for every cell and
    every face the neighbour kind is: same-level leaf / coarser leaf / finer leaves / domain boundary."""
    ocfg = pu.oracle_cfg(dim, lmin, lmax, 1)
    om = so.Mesh.uniform(ocfg)
    om, _ = so.adapt(om, so.init_disc(om, [0.3] * dim, 0.2), so.Bc(), 2e-4, 1.0)
    assert len(om.leaf_levels()) > 1
    levels, ivl = [], []
    for l in om.leaf_levels():
        c = so.unpack(om.cells[l], dim)
        full = np.zeros((c.shape[0], 3), dtype=np.int64)
        full[:, :dim] = c
        for x, y, z in full:  # one single-cell interval per leaf: the product merges them
            levels.append(l)
            ivl.append((y, z, x, x + 1, 0))
    pm = sb.MRMesh.from_intervals([0.0] * dim, [1.0] * dim, pu.product_cfg(dim, lmin, lmax, 1), levels, np.array(ivl, dtype=sb.INTERVAL_DTYPE))
    pu.assert_same_mesh(pm, om)
    rec = pm.debug_flux_records()
    seen = {l: [] for l in om.leaf_levels()}
    SAME, COARSE, FINE, BDRY = 0, 1, 2, 3
    for level, x, y, z, n, kinds in rec:
        xs = np.arange(x, x + n)
        coords = np.stack([xs, np.full(n, y), np.full(n, z)], axis=1)[:, :dim]
        keys = so.pack(coords)
        seen[int(level)].append(keys)
        for d in range(dim):
            for plus in (0, 1):
                e = [0] * dim
                e[d] = 1 if plus else -1
                nb = so.translate(keys, e)
                is_same = np.isin(nb, om.cells[level])
                is_bdry = ~om.in_domain(level, nb)
                is_coarse = np.isin(so.pack(so.unpack(nb, dim) >> 1), om.cells[level - 1]) & ~is_bdry if level > 0 else np.zeros(n, bool)
                is_fine = np.isin(so.pack(so.unpack(nb, dim) << 1), om.cells[level + 1]) & ~is_bdry
                truth = np.select([is_same, is_bdry, is_coarse, is_fine], [SAME, BDRY, COARSE, FINE], default=-1)
                assert np.all(truth >= 0)
                kind = (int(kinds) >> (2 * (2 * d + plus))) & 3
                if d == 0:
                    # x faces: interior faces of a record are same-level; the record's kind applies to its end cell
                    inner = truth[1:] if not plus else truth[:-1]
                    assert np.all(inner == SAME)
                    assert kind == (truth[-1] if plus else truth[0])
                else:
                    assert np.all(truth == kind), f"level {level} row ({y},{z}) x {x}+{n} face d={d} plus={plus}: {truth} vs {kind}"
    for l in om.leaf_levels():
        got = np.sort(np.concatenate(seen[l]))
        assert np.array_equal(got, om.cells[l]), f"level {l}: the records do not tile the leaves"
    pm.destroy()
