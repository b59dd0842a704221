"""Periodic meshes in the oracle, checked the way the reference checks them (tests/test_periodic.cpp:44-108): on a mesh periodic in
every direction the field is shifted by one finest cell per step along the diagonal, `unp1(level, i, index) = u(level, i - 1, index - 1)`,
with MRadaptation before every step; after one full period (2 / dx steps on [-1, 1]^dim) the field must be back on its initial state.
Default mesh_config: max_stencil_radius 2 (mesh_config.hpp:388-393), so this also exercises the ghost-width-2 mesh construction."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import samurai_oracle as so  # noqa: E402


def init(mesh):
    cfg, dim = mesh.cfg, mesh.cfg.dim
    dx = cfg.cell_length(cfg.max_level)
    radius = np.floor(0.2 / dx) * dx
    u = np.zeros(mesh.nref)
    for l in mesh.leaf_levels():
        k = mesh.cells[l]
        c = mesh.cell_centers(l, k)
        u[mesh.index(l, k)] = np.where(np.all(np.abs(c) <= radius, axis=1), 1.0, 0.0)
    return u


def shift_step(mesh, u):
    dim = mesh.cfg.dim
    unp1 = np.full(mesh.nref, np.nan)
    for l in mesh.leaf_levels():
        k = mesh.cells[l]
        unp1[mesh.index(l, k)] = u[mesh.index(l, so.translate(k, [-1] * dim))]
    return unp1


@pytest.mark.parametrize("dim,min_level,max_level,msr", [(1, 3, 6, 2), (2, 3, 6, 2), (2, 2, 5, 1), (3, 2, 4, 2)])
def test_one_period_of_diagonal_shifts_returns_to_the_initial_state(dim, min_level, max_level, msr):
    cfg = so.MeshConfig(dim=dim, min_level=min_level, max_level=max_level, pred_radius=1, max_stencil_radius=msr, origin=(-1.0,) * dim, scaling=2.0,
                        periodic=(True,) * dim)
    bc = so.Bc("dirichlet", 0.0)  # never used: no boundary
    mesh = so.Mesh.uniform(cfg)
    u = init(mesh)
    mesh, u = so.adapt(mesh, u, bc, 1e-4, 1.0)
    n_steps = 1 << max_level  # Tf / dt = (2 / dx) / 1
    seen_levels = set()
    for _ in range(n_steps):
        mesh, u = so.adapt(mesh, u, bc, 1e-4, 1.0)
        so.update_ghost_mr(mesh, u, bc)
        u = shift_step(mesh, u)
        seen_levels.update(mesh.leaf_levels())
    assert len(seen_levels) > 1, "the mesh never adapted"
    u0 = init(mesh)
    for l in mesh.leaf_levels():
        i = mesh.index(l, mesh.cells[l])
        assert np.array_equal(u[i], u0[i]), f"level {l}: {np.count_nonzero(u[i] != u0[i])} leaves differ after one period"
