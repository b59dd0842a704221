"""Regenerate tests/golden/*.npz from the reference's own golden HDF5 files.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
Source: /root/reference/tests/reference/finite_volume/test_finite_volume_advection_2d_*.h5, the files
samurai's pytest suite compares the advection_2d demo against (tests/test_demo_finite_volume.py:55-72,
run with --Tf 0.01; tolerance rel 1e-14 / abs 1e-7, tests/conftest.py:121-122).
Each .npz holds, per leaf cell in for_each_cell order: level (int8), idx (int32 [N,2]), u (float64).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import h5mini  # noqa: E402

REF = "/root/reference/tests/reference/finite_volume/"
BASE = "test_finite_volume_advection_2d_finite-volume-advection-2d-0.01_"

if __name__ == "__main__":
    for name in ("pred_0_init", "pred_0", "pred_1_init", "pred_1"):
        level, idx, fields = h5mini.read_samurai_mesh(REF + BASE + name + ".h5")
        out = os.path.join(HERE, "advection_2d_" + name + ".npz")
        np.savez_compressed(out, level=level.astype(np.int8), idx=idx.astype(np.int32), u=fields["u"])
        print(out, len(level), os.path.getsize(out))
    # heat.cpp --explicit --init-sol=dirac --Tf=0.1 --min-level=3 --max-level=6 on [-4, 4]^2 (tests/test_demo_finite_volume.py:99-116):
    # flux-based diffusion (make_diffusion_order2) across level jumps, Neumann(0), MR adaptation every step
    h5 = h5mini.H5File(REF + "test_finite_volume_demo_heat_explicit.h5")
    pts = h5.read("/mesh/points")
    conn = h5.read("/mesh/connectivity").reshape(-1, 4).astype(np.int64)
    lo = pts[conn].min(axis=1)[:, :2]
    level = h5.read("/mesh/fields/level").astype(np.int64)
    length = 8.0 / (1 << level)
    idx = np.rint((lo + 4.0) / length[:, None]).astype(np.int64)
    assert np.allclose(pts[conn].max(axis=1)[:, 0] - lo[:, 0], length)
    out = os.path.join(HERE, "heat_explicit.npz")
    np.savez_compressed(out, level=level.astype(np.int8), idx=idx.astype(np.int32), u=h5.read("/mesh/fields/u"))
    print(out, len(level), os.path.getsize(out))

    # burgers_mra.cpp --nfiles=1 --min-level=2 --max-level=9 --init-sol=hat --mr-eps=1e-5 (tests/test_demo_finite_volume.py:191-207): 1D,
    # box [-2, 3], max_stencil_radius 2 (ghost width 2), graduation width 2, Dirichlet<1>(0), regularity 0, scheme = 0.5 * make_convection_upwind
    # (non-linear), cfl 0.95, Tf 0.1; the file without the max-level-flux option
    h5 = h5mini.H5File(REF + "test_finite_volume_demo_mra_burgers_hat.h5")
    pts = h5.read("/mesh/points")
    conn = h5.read("/mesh/connectivity").reshape(-1, 2).astype(np.int64)
    lo = pts[conn].min(axis=1)[:, 0]
    level = h5.read("/mesh/fields/level").astype(np.int64)
    length = 5.0 / (1 << level)
    idx = np.rint((lo + 2.0) / length).astype(np.int64)
    assert np.allclose(pts[conn].max(axis=1)[:, 0] - lo, length)
    out = os.path.join(HERE, "mra_burgers_hat.npz")
    np.savez_compressed(out, level=level.astype(np.int8), idx=idx.astype(np.int32)[:, None], u=h5.read("/mesh/fields/u"))
    print(out, len(level), os.path.getsize(out))

    # linear_convection.cpp --nfiles=1 --min-level=1 --max-level=6 --Tf=0.1 (tests/test_demo_finite_volume.py:280-298, explicit): 2D,
    # box [-1, 1]^2 periodic in both directions, max_stencil_size(6) (ghost width 3), make_convection_weno5 (NON-LINEAR flux scheme,
    # six-cell line stencil), velocity (1, -1), TVD-RK3, cfl 0.95, default mra_config (eps 1e-4, regularity 1)
    h5 = h5mini.H5File(REF + "test_finite_volume_demo_linear_convection_explicit.h5")
    pts = h5.read("/mesh/points")
    conn = h5.read("/mesh/connectivity").reshape(-1, 4).astype(np.int64)
    lo = pts[conn].min(axis=1)[:, :2]
    level = h5.read("/mesh/fields/level").astype(np.int64)
    length = 2.0 / (1 << level)
    idx = np.rint((lo + 1.0) / length[:, None]).astype(np.int64)
    assert np.allclose(pts[conn].max(axis=1)[:, 0] - lo[:, 0], length)
    out = os.path.join(HERE, "linear_convection_explicit.npz")
    np.savez_compressed(out, level=level.astype(np.int8), idx=idx.astype(np.int32), u=h5.read("/mesh/fields/u"))
    print(out, len(level), os.path.getsize(out))
