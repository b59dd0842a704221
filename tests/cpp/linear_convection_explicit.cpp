// Linear convection with WENO5 and TVD-RK3 on a fully periodic adapted mesh through the drop-in <samurai/...> headers: the explicit
// branch of the reference's demos/FiniteVolume/linear_convection.cpp:84-222 with its statements kept as they are (the demo itself also
// instantiates a PETSc solver for --implicit, which is out of scope, so it cannot be compiled unchanged).  tests/test_gpu_demos.py runs it
// with the reference test's arguments (--min-level=1 --max-level=6 --Tf=0.1, tests/test_demo_finite_volume.py:280-298) and compares the
// saved HDF5 file with the reference's golden test_finite_volume_demo_linear_convection_explicit.h5.
#include <samurai/io/hdf5.hpp>
#include <samurai/mr/adapt.hpp>
#include <samurai/mr/mesh.hpp>
#include <samurai/samurai.hpp>
#include <samurai/schemes/fv.hpp>

#include <filesystem>
namespace fs = std::filesystem;

template <class Field>
void save(const fs::path& path, const std::string& filename, const Field& u, const std::string& suffix = "")
{
    auto mesh   = u.mesh();
    auto level_ = samurai::make_scalar_field<std::size_t>("level", mesh);

    if (!fs::exists(path))
    {
        fs::create_directory(path);
    }

    samurai::for_each_cell(mesh,
                           [&](const auto& cell)
                           {
                               level_[cell] = cell.level;
                           });

    samurai::save(path, fmt::format("{}{}", filename, suffix), mesh, u, level_);
    samurai::dump(path, fmt::format("{}_restart{}", filename, suffix), mesh, u);
}

int main(int argc, char* argv[])
{
    auto& app = samurai::initialize("Finite volume example for the linear convection equation", argc, argv);

    static constexpr std::size_t dim = 2;
    using Box                        = samurai::Box<double, dim>;
    using point_t                    = typename Box::point_t;

    double left_box  = -1;
    double right_box = 1;
    double Tf        = 3;
    double dt        = 0;
    double cfl       = 0.95;
    double t         = 0.;
    fs::path path        = fs::current_path();
    std::string filename = "linear_convection_" + std::to_string(dim) + "D";

    app.add_option("--Tf", Tf, "Final time")->capture_default_str()->group("Simulation parameters");
    app.add_option("--dt", dt, "Time step")->capture_default_str()->group("Simulation parameters");
    app.add_option("--cfl", cfl, "The CFL")->capture_default_str()->group("Simulation parameters");
    app.add_option("--path", path, "Output path")->capture_default_str()->group("Output");
    app.add_option("--filename", filename, "File name prefix")->capture_default_str()->group("Output");
    app.allow_extras();
    SAMURAI_PARSE(argc, argv);

    point_t box_corner1, box_corner2;
    box_corner1.fill(left_box);
    box_corner2.fill(right_box);
    Box box(box_corner1, box_corner2);
    auto config = samurai::mesh_config<dim>().min_level(1).max_level(dim == 1 ? 6 : 4).periodic(true).max_stencil_size(6);
    auto mesh   = samurai::mra::make_mesh(box, config);
    // Initial solution
    auto u = samurai::make_scalar_field<double>("u",
                                                mesh,
                                                [](const auto& coords)
                                                {
                                                    const auto& x = coords(0);
                                                    const auto& y = coords(1);
                                                    return (x >= -0.8 && x <= -0.3 && y >= 0.3 && y <= 0.8) ? 1. : 0.;
                                                });

    auto unp1 = samurai::make_scalar_field<double>("unp1", mesh);
    // Intermediary fields for the RK3 scheme
    auto u1 = samurai::make_scalar_field<double>("u1", mesh);
    auto u2 = samurai::make_scalar_field<double>("u2", mesh);

    // Convection operator
    samurai::VelocityVector<dim> velocity;
    velocity.fill(1);
    if constexpr (dim == 2)
    {
        velocity(1) = -1;
    }
    auto conv = samurai::make_convection_weno5<decltype(u)>(velocity);

    if (dt == 0)
    {
        double dx             = mesh.min_cell_length();
        double sum_velocities = 0;
        for (std::size_t d = 0; d < dim; ++d)
        {
            sum_velocities += std::abs(velocity(d));
        }
        dt = cfl * dx / sum_velocities;
    }

    auto MRadaptation = samurai::make_MRAdapt(u);
    auto mra_config   = samurai::mra_config();
    MRadaptation(mra_config);

    std::size_t nt = 0;
    while (t != Tf)
    {
        // Move to next timestep
        t += dt;
        if (t > Tf)
        {
            dt += Tf - t;
            t = Tf;
        }
        ++nt;

        // Mesh adaptation
        MRadaptation(mra_config);
        unp1.resize();
        u1.resize();
        u2.resize();

        // TVD-RK3 (SSPRK3)
        u1   = u - dt * conv(u);
        u2   = 3. / 4 * u + 1. / 4 * (u1 - dt * conv(u1));
        unp1 = 1. / 3 * u + 2. / 3 * (u2 - dt * conv(u2));

        // u <-- unp1
        samurai::swap(u, unp1);
    }
    std::cout << "steps " << nt << std::endl;
    save(path, filename, u);
    samurai::finalize();
    return 0;
}
