// Explicit heat equation on an adapted mesh through the drop-in <samurai/...> headers: the explicit branch of the reference's
// demos/FiniteVolume/heat.cpp:112-236 (`--explicit --init-sol=dirac`; the demo itself also instantiates a PETSc solver, which
// is out of scope, so it cannot be compiled unchanged).  Writes the final leaves as CSV; tests/test_gpu_demos.py compares
// them with the reference's golden file test_finite_volume_demo_heat_explicit.h5.
#include <samurai/mr/adapt.hpp>
#include <samurai/mr/mesh.hpp>
#include <samurai/samurai.hpp>
#include <samurai/schemes/fv.hpp>
#include <samurai/io/hdf5.hpp>

#include <cmath>
#include <filesystem>

int main(int argc, char* argv[])
{
    samurai::initialize("explicit heat, flux-based diffusion", argc, argv);
    constexpr std::size_t dim = 2;
    using Box                 = samurai::Box<double, dim>;
    std::filesystem::path path = argc > 1 ? argv[1] : ".";
    const std::size_t min_level = 3, max_level = 6;
    const double K = 1, cfl = 0.95, Tf = 0.1;

    typename Box::point_t box_corner1, box_corner2;
    box_corner1.fill(-4.);
    box_corner2.fill(4.);
    Box box(box_corner1, box_corner2);
    auto config = samurai::mesh_config<dim>().min_level(min_level).max_level(max_level).max_stencil_size(2).disable_minimal_ghost_width();
    auto mesh   = samurai::mra::make_mesh(box, config);
    auto u      = samurai::make_scalar_field<double>("u", mesh);
    auto unp1   = samurai::make_scalar_field<double>("unp1", mesh);
    u.resize();
    double t = 1e-2;
    samurai::for_each_cell(mesh,
                           [&](auto& cell)
                           {
                               double r = 1;
                               for (std::size_t d = 0; d < dim; ++d)
                               {
                                   r *= 1 / (2 * std::sqrt(M_PI * K * t)) * std::exp(-cell.center(d) * cell.center(d) / (4 * K * t));
                               }
                               u[cell] = r;
                           });
    samurai::make_bc<samurai::Neumann<1>>(u, 0.);
    samurai::make_bc<samurai::Neumann<1>>(unp1, 0.);

    samurai::DiffCoeff<dim> Kd;
    Kd.fill(K);
    auto diff = samurai::make_diffusion_order2<decltype(u)>(Kd);

    const double dx = mesh.min_cell_length();
    double dt       = cfl * (dx * dx) / (std::pow(2, dim) * K);
    auto MRadaptation = samurai::make_MRAdapt(u);
    auto mra_config   = samurai::mra_config();
    MRadaptation(mra_config);
    std::size_t nt = 0;
    while (t != Tf)
    {
        t += dt;
        if (t > Tf)
        {
            dt += Tf - t;
            t = Tf;
        }
        MRadaptation(mra_config);
        unp1.resize();
        unp1 = u - dt * diff(u);
        samurai::swap(u, unp1);
        ++nt;
    }
    std::cout << "steps " << nt << std::endl;
    samurai::save(path, "heat_explicit", mesh, u);
    samurai::finalize();
    return 0;
}
