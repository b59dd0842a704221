"""The drop-in `<samurai/...>` header set: the reference's own demo sources compile UNCHANGED against include/ and link
with libsamurai_b200.so (needs /root/reference, i.e. the build container; the GPU box runs the prebuilt binaries in
tests/test_gpu_demos.py)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/demos/FiniteVolume"


@pytest.mark.parametrize("demo", ["advection_1d", "advection_2d", "advection_3d", "scalar_burgers_2d"])
def test_reference_demo_compiles_unchanged(lib, tmp_path, demo):
    src = os.path.join(REF, demo + ".cpp")
    if not os.path.exists(src):
        pytest.skip("/root/reference is not available here")
    out = tmp_path / demo
    cmd = ["g++", "-std=c++20", "-O1", "-I" + os.path.join(ROOT, "include"), "-o", str(out), src, "-L" + os.path.join(ROOT, "samurai_b200"),
           "-lsamurai_b200", "-Wl,-rpath," + os.path.join(ROOT, "samurai_b200")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-4000:]
    # without a GPU the program must fail loudly at samurai::initialize (no CPU fallback), not compute on the host
    import torch

    if not torch.cuda.is_available():
        run = subprocess.run([str(out), "--Tf", "0.001"], capture_output=True, text=True, timeout=120)
        assert run.returncode != 0
        assert "no CUDA device" in (run.stderr + run.stdout)


def test_fmt_and_cli_shims(tmp_path):
    code = r'''
#include <samurai/samurai.hpp>
#include <array>
#include <cassert>
int main(int argc, char** argv) {
    assert(fmt::format("{}_pred_{}", "FV", 1) == "FV_pred_1");
    assert(fmt::format("t = {}, dt = {}", 0.5, 0.00048828125) == "t = 0.5, dt = 0.00048828125");
    double Tf = .1; std::array<double, 2> a{{1, 1}}; xt::xtensor_fixed<double, xt::xshape<2>> c = {0., 0.};
    samurai::app.add_option("--Tf", Tf, "Final time")->capture_default_str()->group("g");
    samurai::app.add_option("--velocity", a, "v")->capture_default_str();
    samurai::app.add_option("--min-corner", c, "c");
    SAMURAI_PARSE(argc, argv);
    assert(Tf == 0.01 && a[0] == 2 && a[1] == 3 && c[1] == -1);
    return 0;
}'''
    src = tmp_path / "t.cpp"
    src.write_text(code)
    exe = tmp_path / "t"
    r = subprocess.run(["g++", "-std=c++20", "-I" + os.path.join(ROOT, "include"), "-o", str(exe), str(src), "-L" + os.path.join(ROOT, "samurai_b200"),
                        "-lsamurai_b200", "-Wl,-rpath," + os.path.join(ROOT, "samurai_b200")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([str(exe), "--Tf", "0.01", "--velocity", "2", "3", "--min-corner", "0", "-1"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr


def test_scheme_objects_and_readme_spellings_compile(lib, tmp_path):
    """<samurai/schemes/fv.hpp>: make_convection_upwind (linear and non-linear), make_diffusion_order2, scalar * scheme,
    rhs = scheme(u), scheme.apply(out, in), unp1 = u - dt * scheme(u); README spellings make_field<double, 1> and
    Box({..}, {..}); plus tests/cpp/heat_explicit.cpp (the explicit branch of demos/FiniteVolume/heat.cpp)."""
    code = r'''
#include <samurai/mr/adapt.hpp>
#include <samurai/mr/mesh.hpp>
#include <samurai/samurai.hpp>
#include <samurai/schemes/fv.hpp>
int main(int argc, char** argv) {
    samurai::initialize("schemes", argc, argv);
    constexpr std::size_t dim = 2;
    const samurai::Box<double, dim> box({0., 0.}, {1., 1.});
    auto cfg  = samurai::mesh_config<dim>().min_level(2).max_level(5).max_stencil_size(2).disable_minimal_ghost_width();
    auto mesh = samurai::mra::make_mesh(box, cfg);
    auto u    = samurai::make_field<double, 1>("u", mesh);
    auto unp1 = samurai::make_scalar_field<double>("unp1", mesh);
    auto rhs  = samurai::make_scalar_field<double>("rhs", mesh);
    samurai::make_bc<samurai::Dirichlet<1>>(u, 0.);
    samurai::VelocityVector<dim> velocity;
    velocity.fill(1);
    samurai::DiffCoeff<dim> K;
    K.fill(0.5);
    auto conv    = samurai::make_convection_upwind<decltype(u)>(velocity);
    auto burgers = 0.5 * samurai::make_convection_upwind<decltype(u)>();
    auto diff    = samurai::make_diffusion_order2<decltype(u)>(K);
    double dt    = 1e-3;
    rhs          = conv(u);
    diff.apply(rhs, u);
    unp1 = u - dt * burgers(u);
    unp1 = u - dt * diff(u);
    samurai::finalize();
    return 0;
}'''
    src = tmp_path / "schemes.cpp"
    src.write_text(code)
    link = ["-L" + os.path.join(ROOT, "samurai_b200"), "-lsamurai_b200", "-Wl,-rpath," + os.path.join(ROOT, "samurai_b200")]
    for s, exe in ((src, tmp_path / "schemes"), (os.path.join(ROOT, "tests", "cpp", "heat_explicit.cpp"), tmp_path / "heat")):
        r = subprocess.run(["g++", "-std=c++20", "-O1", "-I" + os.path.join(ROOT, "include"), "-o", str(exe), str(s)] + link,
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-4000:]
