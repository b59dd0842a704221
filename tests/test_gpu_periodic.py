"""Periodic meshes on the GPU (update_ghost_periodic, periodic ghost cells in the mesh construction,
update_tag_periodic, graduation across the periodic boundary) against the oracle, whose periodic path is checked with the
reference's own test property in tests/test_periodic_oracle.py."""
import numpy as np
import pytest

import parity_utils as pu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim,min_level,max_level,periodic,msr,steps", [
    (2, 2, 6, (True, True), 1, 40),     # the disc crosses the upper-right corner and re-enters
    (2, 2, 6, (True, True), 2, 12),     # library default ghost width (mesh_config.hpp:388-393)
    # Not here: periodic in one direction only.  The reference algorithm leaves the ghosts at the corner between a periodic and a
    # non-periodic boundary unfilled when their mirror is absent from the mesh (no corner extrapolation in periodic directions,
    # update_outer_ghost.hpp:352-366; no periodic pair, update_periodic.hpp:113-124), and the detail of the neighbouring cells reads
    # them: the result depends on stale memory in the reference itself, so there is nothing bit-exact to compare.
    (1, 2, 8, (True,), 2, 30),
    (3, 2, 4, (True, True, True), 1, 6),
])
def test_periodic_advection_matches_oracle(gpu, dim, min_level, max_level, periodic, msr, steps):
    """every step: all sub-meshes and storage offsets bit-exact, tags and details of every harten iteration, ghosts after the ghost
    update and leaves after the upwind step within 1e-12 (observed: bit-equal)."""
    r = pu.run_advection_parity(dim=dim, min_level=min_level, max_level=max_level, pred_radius=1, steps=steps, periodic=periodic, msr=msr,
                                trace_tags=True, a=[1.0] * dim, cfl=0.5 if dim < 3 else 0.25)
    assert r["max_rel_err"] <= pu.REL_TOL


def test_reference_periodic_test_through_dropin_headers(gpu):
    """tests/cpp/periodic_shift.cpp = the body of the reference's tests/test_periodic.cpp (1D, 2D, 3D; default mesh_config, i.e. ghost
    width 2; the shift is a user lambda over u(level, i - 1, index - 1)) compiled against include/samurai: after one period of diagonal
    shifts with MRadaptation at every step the field is back on its initial state."""
    import os
    import subprocess

    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "build", "demos", "periodic-shift")
    if not os.path.exists(exe):
        pytest.skip("periodic-shift not built")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0 and "periodic OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("back on the initial state") == 3
