"""The compiled CPU path (oracle/cpu_path.cpp, the CPU arm of bench.py) against the numpy oracle: meshes identical and
every reference cell BIT-identical after every step -- the numpy oracle is the one pinned on the reference's golden files."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import cpu_path  # noqa: E402
import samurai_oracle as so  # noqa: E402


def oracle_leaves(omesh, u):
    lv, co, ix = omesh.leaf_table()
    return lv, co, u[ix]


def cpu_leaf_table(sim, dim):
    iv, vals = sim.leaves()
    n = (iv[:, 4] - iv[:, 3]).astype(np.int64)
    rep = np.repeat(np.arange(iv.shape[0]), n)
    k = np.arange(int(n.sum())) - np.repeat(np.cumsum(n) - n, n)
    cols = [iv[rep, 3] + k, iv[rep, 1], iv[rep, 2]][:dim]
    return iv[rep, 0], np.stack(cols, axis=1), vals


@pytest.mark.parametrize("dim,min_level,max_level,radius,steps", [(2, 2, 6, 1, 6), (2, 3, 7, 0, 4), (3, 2, 4, 1, 3), (1, 2, 8, 1, 5)])
def test_cpu_path_is_bit_identical_to_the_numpy_oracle(dim, min_level, max_level, radius, steps):
    eps = 2e-4
    cfg = so.MeshConfig(dim=dim, min_level=min_level, max_level=max_level, pred_radius=radius)
    bc = so.Bc("dirichlet", 0.0)
    omesh = so.Mesh.uniform(cfg)
    ou = so.init_disc(omesh, [0.3] * dim, 0.2)
    sim = cpu_path.CpuSim(dim, min_level, max_level, radius, eps=eps, regularity=1.0)
    sim.init_ball([0.3] * dim, 0.2)
    omesh, ou = so.adapt(omesh, ou, bc, eps, 1.0)
    sim.adapt()
    a = [1.0] * dim
    dt = 0.5 * cfg.cell_length(max_level) / dim
    for step in range(steps):
        omesh, ou = so.adapt(omesh, ou, bc, eps, 1.0)
        so.update_ghost_mr(omesh, ou, bc)
        ou = so.fv_step(omesh, ou, a, dt)
        assert sim.steps(1, a, dt) == omesh.nb_cells()
        lv, co, vals = cpu_leaf_table(sim, dim)
        olv, oco, ovals = oracle_leaves(omesh, ou)
        assert np.array_equal(lv, olv) and np.array_equal(co, oco), f"step {step}: leaves differ"
        assert sim.nb_cells(True) == omesh.nref
        assert np.array_equal(vals, ovals), f"step {step}: max abs diff {np.max(np.abs(vals - ovals))}"
    # ghosts too: every reference cell the oracle defines
    omesh, ou = so.adapt(omesh, ou, bc, eps, 1.0)
    so.update_ghost_mr(omesh, ou, bc)
    sim.adapt()
    sim.update_ghost()
    f = sim.field()
    m = np.isfinite(ou)
    assert np.array_equal(f[m], ou[m])


def test_cpu_path_uses_all_cores():
    assert cpu_path.CpuSim.threads() >= 1
