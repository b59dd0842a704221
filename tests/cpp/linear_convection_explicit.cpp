// WENO5 + TVD-RK3 on a fully periodic adapted mesh through the drop-in <samurai/...> headers.
//
// What the reference's demos/FiniteVolume/linear_convection.cpp does in its explicit mode (the demo itself also builds a PETSc solver
// for --implicit, so it cannot be compiled here): periodic box [-1, 1]^2, six-cell stencils (ghost width 3), a unit square of ones
// transported with velocity (1, -1), mesh adaptation before every step, three RK stages written as field expressions.
// tests/test_gpu_demos.py runs it with --min-level=1 --max-level=6 --Tf=0.1 and compares the saved file with the reference's own
// golden test_finite_volume_demo_linear_convection_explicit.h5.
#include <samurai/io/hdf5.hpp>
#include <samurai/mr/adapt.hpp>
#include <samurai/mr/mesh.hpp>
#include <samurai/samurai.hpp>
#include <samurai/schemes/fv.hpp>

#include <cmath>
#include <filesystem>
#include <iostream>

int main(int argc, char* argv[])
{
    auto& cli = samurai::initialize("WENO5 linear convection, explicit RK3", argc, argv);
    constexpr std::size_t dim = 2;

    double t_end = 3, cfl = 0.95;
    std::filesystem::path out_dir = std::filesystem::current_path();
    std::string out_name          = "linear_convection_2D";
    cli.add_option("--Tf", t_end, "final time");
    cli.add_option("--cfl", cfl, "CFL number");
    cli.add_option("--path", out_dir, "output directory");
    cli.add_option("--filename", out_name, "output file name");
    cli.allow_extras();
    SAMURAI_PARSE(argc, argv);

    samurai::Box<double, dim> box({-1., -1.}, {1., 1.});
    auto cfg  = samurai::mesh_config<dim>().min_level(1).max_level(4).periodic(true).max_stencil_size(6);
    auto mesh = samurai::mra::make_mesh(box, cfg);

    auto in_square = [](const auto& x)
    {
        return (x(0) >= -0.8 && x(0) <= -0.3 && x(1) >= 0.3 && x(1) <= 0.8) ? 1. : 0.;
    };
    auto u      = samurai::make_scalar_field<double>("u", mesh, in_square);
    auto stage1 = samurai::make_scalar_field<double>("u1", mesh);
    auto stage2 = samurai::make_scalar_field<double>("u2", mesh);
    auto next   = samurai::make_scalar_field<double>("unp1", mesh);

    samurai::VelocityVector<dim> a;
    a(0)      = 1;
    a(1)      = -1;
    auto conv = samurai::make_convection_weno5<decltype(u)>(a);

    double dt = cfl * mesh.min_cell_length() / (std::abs(a(0)) + std::abs(a(1)));

    auto adapt = samurai::make_MRAdapt(u);
    auto mra   = samurai::mra_config();
    adapt(mra);

    std::size_t n_steps = 0;
    for (double t = 0; t != t_end; ++n_steps)
    {
        t += dt;
        if (t > t_end) // last step lands on t_end exactly
        {
            dt += t_end - t;
            t = t_end;
        }
        adapt(mra);
        stage1.resize();
        stage2.resize();
        next.resize();
        stage1 = u - dt * conv(u);
        stage2 = 3. / 4 * u + 1. / 4 * (stage1 - dt * conv(stage1));
        next   = 1. / 3 * u + 2. / 3 * (stage2 - dt * conv(stage2));
        samurai::swap(u, next);
    }
    std::cout << "steps " << n_steps << std::endl;
    samurai::save(out_dir, out_name, mesh, u);
    samurai::finalize();
    return 0;
}
