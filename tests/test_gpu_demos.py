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


def _load_h5(path, dim=2, field="u"):
    """The demos' samurai::save output (reference layout: /mesh/points, /mesh/connectivity, /mesh/fields/*), read with the pure-Python
    HDF5 reader that also parses the reference's golden files.  Returns one row per cell: level, indices..., field value."""
    import sys

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import h5mini

    h5 = h5mini.H5File(str(path))
    pts = h5.read("/mesh/points")
    conn = h5.read("/mesh/connectivity").astype(np.int64)
    corners = pts[conn]
    lo = corners.min(axis=1)[:, :dim]
    hi = corners.max(axis=1)[:, :dim]
    level = h5.read("/mesh/fields/level").astype(np.int64) if "level" in h5.listdir("/mesh/fields") else None
    length = hi[:, 0] - lo[:, 0]
    return lo, length, level, h5.read("/mesh/fields/" + field)


def _cells_h5(path, dim, origin, scaling, field="u"):
    lo, length, level, u = _load_h5(path, dim, field)
    if level is None:
        level = np.rint(np.log2(scaling / length)).astype(np.int64)
    idx = np.rint((lo - np.asarray(origin)[None, :dim]) / length[:, None]).astype(np.int64)
    return np.concatenate([level[:, None].astype(np.float64), idx.astype(np.float64), u[:, None]], axis=1)


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
            got = _cells_h5(tmp_path / f"adv2d_pred_{pred}{suffix}.h5", 2, [0.0, 0.0], 1.0)
            assert os.path.exists(tmp_path / f"adv2d_pred_{pred}{suffix}.xdmf")
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
    files = sorted(p for p in os.listdir(tmp_path) if p.endswith(".h5") and "restart" not in p)
    assert files
    dim = 3 if "3d" in demo else 2
    data = _cells_h5(tmp_path / files[-1], dim, [0.0] * dim, 1.0)
    assert data.shape[0] > 100 and np.all(np.isfinite(data))
    u = data[:, -1]
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
    got = _cells_h5(tmp_path / "heat_explicit.h5", 2, [-4.0, -4.0], 8.0)
    g = np.load(os.path.join(GOLD, "heat_explicit.npz"))
    assert got.shape[0] == g["level"].size
    assert np.array_equal(got[:, 0].astype(np.int64), g["level"].astype(np.int64))
    assert np.array_equal(got[:, 1:3].astype(np.int64), g["idx"].astype(np.int64)), "mesh differs"
    assert np.max(np.abs(got[:, 3] - g["u"])) <= 1e-15


def test_cpp_linear_convection_weno5_matches_reference_golden(gpu, tmp_path):
    """tests/cpp/linear_convection_explicit.cpp (the explicit mode of demos/FiniteVolume/linear_convection.cpp written with the same
    API calls: fully periodic box, max_stencil_size(6), make_convection_weno5, TVD-RK3 as field expressions, MRadaptation every step) with
    the reference test's arguments (tests/test_demo_finite_volume.py:280-298) against the reference's own golden file
    test_finite_volume_demo_linear_convection_explicit.h5 (tests/golden/linear_convection_explicit.npz)."""
    exe = os.path.join(DEMOS, "linear-convection-explicit")
    if not os.path.exists(exe):
        pytest.skip("linear-convection-explicit not built")
    r = subprocess.run([exe, "--path", str(tmp_path), "--filename", "lc", "--nfiles=1", "--min-level=1", "--max-level=6", "--Tf=0.1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "steps 7" in r.stdout
    got = _cells_h5(tmp_path / "lc.h5", 2, [-1.0, -1.0], 2.0)
    g = np.load(os.path.join(GOLD, "linear_convection_explicit.npz"))
    assert got.shape[0] == g["level"].size
    assert np.array_equal(got[:, 0].astype(np.int64), g["level"].astype(np.int64))
    assert np.array_equal(got[:, 1:3].astype(np.int64), g["idx"].astype(np.int64)), "mesh differs"
    assert np.max(np.abs(got[:, 3] - g["u"])) <= 1e-13


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
    assert os.path.exists(tmp_path / "adv1d.h5"), os.listdir(tmp_path)
    got = _cells_h5(tmp_path / "adv1d.h5", 1, [-2.0], 4.0)
    cfg = so.MeshConfig(dim=1, min_level=6, max_level=12, pred_radius=1, origin=(-2.0,), scaling=4.0)
    res = so.run_advection(cfg, Tf=0.1, eps=2e-4, a=[1.0], cfl=0.95, center=[0.0], radius=0.2)
    mesh, u = res["final"]
    lv, co, ix = mesh.leaf_table()
    assert got.shape[0] == lv.size, f"{got.shape[0]} cells vs oracle {lv.size}"
    assert np.array_equal(got[:, 0].astype(np.int64), lv) and np.array_equal(got[:, 1].astype(np.int64), co[:, 0]), "mesh differs"
    assert np.max(np.abs(got[:, 2] - u[ix])) <= 1e-13  # columns: level, i, u, level


def test_readme_example_user_lambda_matches_oracle(gpu):
    """The reference's README example (README.md:84-155) assembled statement for statement in tests/cpp/readme_advection.cpp:
    legacy constructor MRMesh<Config>(box, 2, 8), make_field<double, 1>, u[cell], MRadaptation(1e-4, 2), and the upwind scheme as
    the USER lambda over u(level, i, j) row views (the drop-in's host path, field/access_base.hpp:69-103).  Compared with the
    oracle driven through the same loop, the stencil evaluated in the lambda's operation order: mesh identical, values bit-equal."""
    import sys

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import samurai_oracle as so
    from parity_utils import init_square

    exe = os.path.join(DEMOS, "readme-advection")
    if not os.path.exists(exe):
        pytest.skip("readme-advection not built")
    steps = 50
    r = subprocess.run([exe, str(steps)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = r.stdout.strip().splitlines()
    k = next(i for i, ln in enumerate(lines) if ln.startswith("leaves "))
    got = np.array([[float(x) for x in ln.split()] for ln in lines[k + 1:]])
    cfg = so.MeshConfig(dim=2, min_level=2, max_level=8, pred_radius=1)
    bc = so.Bc("dirichlet", 0.0)
    mesh = so.Mesh.uniform(cfg)
    u = init_square(mesh)
    dt = 0.5 * cfg.cell_length(8)
    for _ in range(steps):
        mesh, u = so.adapt(mesh, u, bc, 1e-4, 2.0)
        so.update_ghost_mr(mesh, u, bc)
        unp1 = np.full(mesh.nref, np.nan)
        for l in mesh.leaf_levels():
            keys = mesh.cells[l]
            ic = mesh.index(l, keys)
            uc = u[ic]
            uxm = u[mesh.index(l, so.translate(keys, [-1, 0]))]
            uym = u[mesh.index(l, so.translate(keys, [0, -1]))]
            dx = cfg.cell_length(l)
            unp1[ic] = uc - dt / dx * (uc - uxm + uc - uym)
        u = unp1
    lv, co, ix = mesh.leaf_table()
    assert got.shape[0] == lv.size, f"{got.shape[0]} leaves vs oracle {lv.size}"
    assert np.array_equal(got[:, 0].astype(np.int64), lv) and np.array_equal(got[:, 1:3].astype(np.int64), co), "mesh differs"
    assert np.array_equal(got[:, 3], u[ix]), f"max abs diff {np.max(np.abs(got[:, 3] - u[ix])):.3e}"


def test_rk3_general_field_expressions_match_oracle(gpu):
    """tests/cpp/rk3_expressions.cpp: SSP-RK3 stages written as general field expressions (`3./4 * u + 1./4 * (u1 - dt * diff(u1))`,
    field/field_expression.hpp:62-141, burgers.cpp:262-269), evaluated on the device node by node.  The oracle evaluates the same
    trees element-wise in numpy: mesh identical, leaf values bit-equal."""
    import sys

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import samurai_oracle as so

    exe = os.path.join(DEMOS, "rk3-expressions")
    if not os.path.exists(exe):
        pytest.skip("rk3-expressions not built")
    steps = 10
    r = subprocess.run([exe, str(steps)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = r.stdout.strip().splitlines()
    k = next(i for i, ln in enumerate(lines) if ln.startswith("leaves "))
    got = np.array([[float(x) for x in ln.split()] for ln in lines[k + 1:]])

    K, cfl = 1.0, 0.5
    cfg = so.MeshConfig(dim=2, min_level=3, max_level=6, pred_radius=1, origin=(-4.0, -4.0), scaling=8.0)
    bc = so.Bc("neumann", 0.0)
    mesh = so.Mesh.uniform(cfg)
    u = np.zeros(mesh.nref)
    L = cfg.max_level
    u[mesh.index(L, mesh.cells[L])] = so.heat_exact(mesh.cell_centers(L, mesh.cells[L]), 1e-2, K)
    dx = cfg.cell_length(L)
    dt = cfl * (dx * dx) / (pow(2, 2) * K)
    coeffs = so.diffusion_order2_coeffs([K, K])
    mesh, u = so.adapt(mesh, u, bc, 1e-4, 1.0)

    def S(v):
        so.update_ghost_mr(mesh, v, bc)
        return so.flux_linhom_apply(mesh, v, coeffs)

    def leaves_expr(fn):
        out = np.full(mesh.nref, np.nan)
        for l in mesh.leaf_levels():
            i = mesh.index(l, mesh.cells[l])
            out[i] = fn(i)
        return out

    for _ in range(steps):
        mesh, u = so.adapt(mesh, u, bc, 1e-4, 1.0)
        r0 = S(u)
        u1 = leaves_expr(lambda i: u[i] - dt * r0[i])
        r1 = S(u1)
        u2 = leaves_expr(lambda i: 3. / 4 * u[i] + 1. / 4 * (u1[i] - dt * r1[i]))
        r2 = S(u2)
        u = leaves_expr(lambda i: 1. / 3 * u[i] + 2. / 3 * (u2[i] - dt * r2[i]))
    lv, co, ix = mesh.leaf_table()
    assert got.shape[0] == lv.size, f"{got.shape[0]} leaves vs oracle {lv.size}"
    assert np.array_equal(got[:, 0].astype(np.int64), lv) and np.array_equal(got[:, 1:3].astype(np.int64), co), "mesh differs"
    assert np.array_equal(got[:, 3], u[ix]), f"max abs diff {np.max(np.abs(got[:, 3] - u[ix])):.3e}"


def test_cpp_vector_field_matches_oracle(gpu):
    """tests/cpp/vector_advection.cpp: make_vector_field<double, 2> (field/vector_field.hpp:234) through the drop-in headers --
    u[cell][c], one Dirichlet value per component, both components adapted together by make_MRAdapt(u) (one tag array, criteria over
    all components: mr/operators.hpp:623-677), `unp1 = u - dt * upwind(a, u)`, std::swap(u.array(), unp1.array()).  Against the
    oracle's adapt_fields + per-component upwind: mesh identical, both components bit-equal."""
    import sys

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import samurai_oracle as so

    exe = os.path.join(DEMOS, "vector-advection")
    if not os.path.exists(exe):
        pytest.skip("vector-advection not built")
    steps = 10
    r = subprocess.run([exe, str(steps)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = r.stdout.strip().splitlines()
    k = next(i for i, ln in enumerate(lines) if ln.startswith("leaves "))
    got = np.array([[float(x) for x in ln.split()] for ln in lines[k + 1:]])

    cfg = so.MeshConfig(dim=2, min_level=2, max_level=7, pred_radius=1)
    bcs = [so.Bc("dirichlet", 0.0), so.Bc("dirichlet", 0.0)]
    mesh = so.Mesh.uniform(cfg)
    L = cfg.max_level
    c = mesh.cell_centers(L, mesh.cells[L])
    ix = mesh.index(L, mesh.cells[L])
    u0, u1 = np.zeros(mesh.nref), np.zeros(mesh.nref)
    u0[ix] = np.where((c[:, 0] - 0.3) * (c[:, 0] - 0.3) + (c[:, 1] - 0.3) * (c[:, 1] - 0.3) <= 0.2 * 0.2, 1.0, 0.0)
    u1[ix] = np.where((c[:, 0] - 0.6) * (c[:, 0] - 0.6) + (c[:, 1] - 0.5) * (c[:, 1] - 0.5) <= 0.15 * 0.15, 2.0, 0.0)
    fields = [u0, u1]
    dt = 0.5 * cfg.cell_length(L)
    mesh, fields = so.adapt_fields(mesh, fields, bcs, 2e-4, 1.0)
    for _ in range(steps):
        mesh, fields = so.adapt_fields(mesh, fields, bcs, 2e-4, 1.0)
        for f, bc in zip(fields, bcs):
            so.update_ghost_mr(mesh, f, bc)
        fields = [so.fv_step(mesh, f, [1.0, 1.0], dt) for f in fields]
    lv, co, ix = mesh.leaf_table()
    assert got.shape[0] == lv.size, f"{got.shape[0]} leaves vs oracle {lv.size}"
    assert np.array_equal(got[:, 0].astype(np.int64), lv) and np.array_equal(got[:, 1:3].astype(np.int64), co), "mesh differs"
    for comp in range(2):
        assert np.array_equal(got[:, 3 + comp], fields[comp][ix]), f"component {comp}: max abs diff {np.max(np.abs(got[:, 3 + comp] - fields[comp][ix])):.3e}"


def test_cpp_vector_burgers_upwind_matches_oracle(gpu):
    """tests/cpp/vector_burgers.cpp: `make_convection_upwind<VectorField>()` (vector form: flux u(d) * u upwinded by the
    mean of component d, operators/convection_nonlin.hpp:24-76) in `unp1 = u - dt * conv(u)` through the drop-in headers, both
    components adapted together.  Against the oracle's adapt_fields + flux_nonlin_apply on the component list: mesh identical,
    both components bit-equal."""
    import sys

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import samurai_oracle as so

    exe = os.path.join(DEMOS, "vector-burgers")
    if not os.path.exists(exe):
        pytest.skip("vector-burgers not built")
    steps = 10
    r = subprocess.run([exe, str(steps)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = r.stdout.strip().splitlines()
    k = next(i for i, ln in enumerate(lines) if ln.startswith("leaves "))
    got = np.array([[float(x) for x in ln.split()] for ln in lines[k + 1:]])

    cfg = so.MeshConfig(dim=2, min_level=2, max_level=6, pred_radius=1, origin=(-1.0, -1.0), scaling=2.0)
    bcs = [so.Bc("dirichlet", 0.0), so.Bc("dirichlet", 0.0)]
    mesh = so.Mesh.uniform(cfg)
    L = cfg.max_level
    c = mesh.cell_centers(L, mesh.cells[L])
    ix = mesh.index(L, mesh.cells[L])
    u0, u1 = np.zeros(mesh.nref), np.zeros(mesh.nref)
    r0 = np.maximum(np.abs(c[:, 0]), np.abs(c[:, 1]))
    r1 = np.maximum(np.abs(c[:, 0] - 0.25), np.abs(c[:, 1] + 0.25))
    u0[ix] = np.where(r0 < 0.5, 1.0 - 2.0 * r0, 0.0)
    u1[ix] = np.where(r1 < 0.4, -(1.0 - 2.5 * r1), 0.0)
    fields = [u0, u1]
    dt = 0.4 * cfg.cell_length(L)
    flux = so.burgers_upwind_flux_vector(2)
    mesh, fields = so.adapt_fields(mesh, fields, bcs, 1e-3, 1.0)
    for _ in range(steps):
        mesh, fields = so.adapt_fields(mesh, fields, bcs, 1e-3, 1.0)
        for f, bc in zip(fields, bcs):
            so.update_ghost_mr(mesh, f, bc)
        rhs = so.flux_nonlin_apply(mesh, fields, flux)
        new = []
        for f, rh in zip(fields, rhs):
            out = np.full(mesh.nref, np.nan)
            for l in mesh.leaf_levels():
                i = mesh.index(l, mesh.cells[l])
                out[i] = f[i] - dt * rh[i]
            new.append(out)
        fields = new
    lv, co, ix = mesh.leaf_table()
    assert len(mesh.leaf_levels()) > 1
    assert got.shape[0] == lv.size, f"{got.shape[0]} leaves vs oracle {lv.size}"
    assert np.array_equal(got[:, 0].astype(np.int64), lv) and np.array_equal(got[:, 1:3].astype(np.int64), co), "mesh differs"
    for comp in range(2):
        assert np.array_equal(got[:, 3 + comp], fields[comp][ix]), f"component {comp}: max abs diff {np.max(np.abs(got[:, 3 + comp] - fields[comp][ix])):.3e}"


def test_reference_demo_restart_roundtrip(gpu, tmp_path):
    """samurai::dump / samurai::load through the unchanged advection_2d demo (`--restart-file`, advection_2d.cpp:57,105-113):
    20 steps in one run equal 10 steps + checkpoint + 10 steps from the checkpoint, bit for bit (mesh and field)."""
    exe = os.path.join(DEMOS, "finite-volume-advection-2d")
    if not os.path.exists(exe):
        pytest.skip("demo binary not built (needs /root/reference at build time)")
    dt = 0.5 * 2.0 ** -10

    def run(name, steps, restart=None):
        cmd = [exe, "--path", str(tmp_path), "--filename", name, "--Tf", repr(steps * dt)]
        if restart:
            cmd += ["--restart-file", str(tmp_path / restart)]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        assert f"iteration {steps - 1}:" in r.stdout and f"iteration {steps}:" not in r.stdout
        return _cells_h5(tmp_path / f"{name}_pred_1.h5", 2, [0.0, 0.0], 1.0)

    full = run("full", 20)
    run("half", 10)
    assert os.path.exists(tmp_path / "half_pred_1_restart.h5")
    cont = run("cont", 10, restart="half_pred_1_restart")
    assert full.shape == cont.shape and np.array_equal(full, cont)
