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
