"""The reference's own demo programs, compiled UNCHANGED from /root/reference/demos/FiniteVolume/*.cpp against the drop-in
headers in include/samurai (done by __graft_entry__.build() in the build container; the binaries travel in build/demos),
run on the GPU and compared with the reference's golden datasets."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMOS = os.path.join(ROOT, "build", "demos")
GOLD = os.path.join(ROOT, "tests", "golden")
pytestmark = pytest.mark.gpu


def _load_csv(path):
    data = np.loadtxt(path, delimiter=",", skiprows=1)
    return data


def test_reference_advection_2d_demo_unchanged(gpu, tmp_path):
    exe = os.path.join(DEMOS, "finite-volume-advection-2d")
    if not os.path.exists(exe):
        pytest.skip("demo binary not built (needs /root/reference at build time)")
    # same command as the reference's regression test (tests/test_demo_finite_volume.py:55-72)
    r = subprocess.run([exe, "--path", str(tmp_path), "--filename", "adv2d", "--Tf", "0.01"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "iteration 20" in r.stdout
    for pred in (0, 1):
        for suffix, gold in (("_init", f"advection_2d_pred_{pred}_init.npz"), ("", f"advection_2d_pred_{pred}.npz")):
            got = _load_csv(tmp_path / f"adv2d_pred_{pred}{suffix}.csv")
            g = np.load(os.path.join(GOLD, gold))
            assert got.shape[0] == g["level"].size, f"{gold}: {got.shape[0]} cells vs {g['level'].size}"
            assert np.array_equal(got[:, 0].astype(np.int64), g["level"].astype(np.int64))
            assert np.array_equal(got[:, 1:3].astype(np.int64), g["idx"].astype(np.int64)), f"{gold}: mesh differs"
            assert np.max(np.abs(got[:, 3] - g["u"])) <= 1e-14, f"{gold}: max abs diff {np.max(np.abs(got[:, 3] - g['u'])):.3e}"


@pytest.mark.parametrize("demo,args,ncol", [("finite-volume-advection-3d", ["--Tf", "0.01", "--max-level", "6"], 5),
                                            ("finite-volume-scalar-burgers-2d", ["--Tf", "0.002", "--max-level", "8"], 5)])
def test_other_reference_demos_run(gpu, tmp_path, demo, args, ncol):
    exe = os.path.join(DEMOS, demo)
    if not os.path.exists(exe):
        pytest.skip("demo binary not built (needs /root/reference at build time)")
    r = subprocess.run([exe, "--path", str(tmp_path), "--filename", "out"] + args, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    files = sorted(p for p in os.listdir(tmp_path) if p.endswith(".csv") and "restart" not in p)
    assert files
    data = _load_csv(tmp_path / files[-1])
    assert data.shape[0] > 100 and np.all(np.isfinite(data))
    u = data[:, ncol - 1] if "burgers" not in demo else data[:, 3]
    assert u.min() > -1.5 and u.max() < 1.5


def test_cpp_heat_explicit_matches_reference_golden(gpu, tmp_path):
    """tests/cpp/heat_explicit.cpp (the explicit branch of demos/FiniteVolume/heat.cpp written against the drop-in headers:
    make_diffusion_order2, `unp1 = u - dt * diff(u)`, Neumann<1>, MRadaptation every step) against the reference's own
    golden file test_finite_volume_demo_heat_explicit.h5 (tests/golden/heat_explicit.npz)."""
    exe = os.path.join(DEMOS, "heat-explicit")
    if not os.path.exists(exe):
        pytest.skip("heat-explicit not built")
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "steps 25" in r.stdout
    got = _load_csv(tmp_path / "heat_explicit.csv")
    g = np.load(os.path.join(GOLD, "heat_explicit.npz"))
    assert got.shape[0] == g["level"].size
    assert np.array_equal(got[:, 0].astype(np.int64), g["level"].astype(np.int64))
    assert np.array_equal(got[:, 1:3].astype(np.int64), g["idx"].astype(np.int64)), "mesh differs"
    assert np.max(np.abs(got[:, 3] - g["u"])) <= 1e-15


def test_reference_advection_1d_demo_unchanged_matches_oracle(gpu, tmp_path):
    """demos/FiniteVolume/advection_1d.cpp compiled unchanged (scalar velocity, `make_bc<Dirichlet<1>>(u, 0.)->on(left, right)`,
    `mesh_config().periodic(false)`), run with the reference test's `--Tf 0.1` (tests/test_demo_finite_volume.py:21-52) and
    compared with the oracle's run of the same loop: box [-2, 2], levels 6-12, eps 2e-4, cfl 0.95, a = 1."""
    import sys

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import samurai_oracle as so

    exe = os.path.join(DEMOS, "finite-volume-advection-1d")
    if not os.path.exists(exe):
        pytest.skip("demo binary not built (needs /root/reference at build time)")
    r = subprocess.run([exe, "--path", str(tmp_path), "--filename", "adv1d", "--Tf", "0.1"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert os.path.exists(tmp_path / "adv1d.csv"), os.listdir(tmp_path)
    got = _load_csv(tmp_path / "adv1d.csv")
    cfg = so.MeshConfig(dim=1, min_level=6, max_level=12, pred_radius=1, origin=(-2.0,), scaling=4.0)
    res = so.run_advection(cfg, Tf=0.1, eps=2e-4, a=[1.0], cfl=0.95, center=[0.0], radius=0.2)
    mesh, u = res["final"]
    lv, co, ix = mesh.leaf_table()
    assert got.shape[0] == lv.size, f"{got.shape[0]} cells vs oracle {lv.size}"
    assert np.array_equal(got[:, 0].astype(np.int64), lv) and np.array_equal(got[:, 1].astype(np.int64), co[:, 0]), "mesh differs"
    assert np.max(np.abs(got[:, 2] - u[ix])) <= 1e-13  # columns: level, i, u, level
