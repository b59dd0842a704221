"""Host half of the WENO5 path, no GPU: ghost width 3 on fully periodic meshes (sub-meshes and storage
numbering bit-exact against the oracle) and the records of the six-cell-stencil flux batch, evaluated on the host with the kernel's own
per-cell function (smr_debug_fluxw_apply, a testing aid), against the oracle's restatement of the reference's scatter loops."""
import numpy as np
import pytest

import parity_utils as pu

sb, so = pu.sb, pu.so


@pytest.fixture(scope="module", autouse=True)
def _init(lib):
    lib.initialize(-1)


def adapted_periodic(dim, lo, hi):
    """The same adapted, fully periodic, ghost-width-3 mesh on both sides: the oracle adapts, the product's host path follows its tags."""
    pcfg = sb.mesh_config(dim, 1).min_level(lo).max_level(hi).periodic([True] * dim).max_stencil_size(6)
    ocfg = so.MeshConfig(dim=dim, min_level=lo, max_level=hi, pred_radius=1, max_stencil_radius=3, graduation_width=1,
                         origin=(-1.0,) * dim, scaling=2.0, periodic=(True,) * dim)
    pm = sb.MRMesh.make_mesh([-1.0] * dim, [1.0] * dim, pcfg)
    om = so.Mesh.uniform(ocfg)
    pu.assert_same_mesh(pm, om)
    c = om.cell_centers(hi, om.cells[hi])
    inside = np.all((c >= -1.0) & (c <= -0.45), axis=1)  # touches the periodic boundary: level jumps through it
    f0 = np.zeros(om.nref)
    f0[om.index(hi, om.cells[hi])] = np.where(inside, 1.0, 0.0)
    bc = so.Bc("neumann", 0.0)
    trace = []
    om2, ou = so.adapt(om, f0, bc, 1e-4, 1.0, trace=trace)
    for t in trace:
        pu.assert_same_mesh(pm, t["mesh"])
        pm.update_from_tags(t["tag"])
    pu.assert_same_mesh(pm, om2)
    return pm, om2, ou, bc


@pytest.mark.parametrize("dim,lo,hi,vel", [(1, 2, 8, [1.0]), (1, 2, 8, [-0.7]), (2, 1, 6, [1.0, -1.0]), (2, 2, 6, [-0.5, 2.0]),
                                           (3, 1, 4, [1.0, -1.0, 0.5])])
def test_weno5_records_match_oracle(dim, lo, hi, vel):
    pm, om, ou, bc = adapted_periodic(dim, lo, hi)
    assert len(om.leaf_levels()) > 1
    leaves = np.concatenate([om.index(l, om.cells[l]) for l in om.leaf_levels()])
    rng = np.random.default_rng(5)
    ou[leaves] += 0.1 * rng.standard_normal(leaves.size)
    so.update_ghost_mr(om, ou, bc)
    want = so.flux_nonlin_apply(om, ou, so.weno5_flux(vel), so.WENO5_OFFSETS)
    got = pm.debug_fluxw_apply(ou, vel)
    bad = np.flatnonzero(got[leaves] != want[leaves])
    assert bad.size == 0, f"{bad.size} of {leaves.size} leaves differ, max {np.max(np.abs(got[leaves] - want[leaves])):.3e}"
    pm.destroy()


@pytest.mark.parametrize("dim,lo,hi,n_comp", [(1, 2, 8, 1), (2, 1, 6, 1), (2, 1, 6, 2), (3, 1, 4, 3)])
def test_nonlinear_weno5_records_match_oracle(dim, lo, hi, n_comp):
    """make_convection_weno5<Field>() (operators/convection_nonlin.hpp:162-233): f = u * u for a scalar field, u(d) * u for a vector
    field with n_comp == dim, upwinded by the mean next to the interface; with and without a scalar factor."""
    pm, om, ou, bc = adapted_periodic(dim, lo, hi)
    leaves = np.concatenate([om.index(l, om.cells[l]) for l in om.leaf_levels()])
    rng = np.random.default_rng(7)
    comps = []
    for c in range(n_comp):
        oc = ou.copy()
        oc[leaves] += 0.4 * rng.standard_normal(leaves.size) - 0.3 * c  # both upwinding signs occur
        so.update_ghost_mr(om, oc, bc)
        comps.append(oc)
    arg = comps if n_comp > 1 else comps[0]
    for scale in (1.0, 0.5):
        want = so.flux_nonlin_apply(om, arg, so.weno5_flux_nonlinear(dim, n_comp, scale), so.WENO5_OFFSETS)
        got = pm.debug_fluxw_apply(arg, None, scale)
        for c in range(n_comp):
            g, w = (got[c], want[c]) if n_comp > 1 else (got, want)
            bad = np.flatnonzero(g[leaves] != w[leaves])
            assert bad.size == 0, f"scale {scale} component {c}: {bad.size} of {leaves.size} leaves differ, max {np.max(np.abs(g[leaves] - w[leaves])):.3e}"
    pm.destroy()
