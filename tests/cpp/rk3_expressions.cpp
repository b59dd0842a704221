// SSP-RK3 time stepping written with general field expressions, as the reference's demos do
// (demos/FiniteVolume/burgers.cpp:262-269, linear_convection.cpp): every stage is a tree of +, -, scalar * over fields and
// scheme(u) terms, evaluated on the device by the drop-in headers (include/samurai/b200_api.hpp, namespace fx).
// Heat equation, explicit diffusion (make_diffusion_order2), adapted mesh; leaves printed for the comparison with the oracle.
#include <samurai/mr/adapt.hpp>
#include <samurai/mr/mesh.hpp>
#include <samurai/samurai.hpp>
#include <samurai/schemes/fv.hpp>

#include <cmath>
#include <cstdio>

int main(int argc, char* argv[])
{
    samurai::initialize("rk3 with field expressions", argc, argv);
    constexpr std::size_t dim = 2;
    using Box                 = samurai::Box<double, dim>;
    const std::size_t n_steps = argc > 1 ? static_cast<std::size_t>(std::atoi(argv[1])) : 10;
    const std::size_t min_level = 3, max_level = 6;
    const double K = 1, cfl = 0.5;
    {
        typename Box::point_t box_corner1, box_corner2;
        box_corner1.fill(-4.);
        box_corner2.fill(4.);
        Box box(box_corner1, box_corner2);
        auto config = samurai::mesh_config<dim>().min_level(min_level).max_level(max_level).max_stencil_size(2).disable_minimal_ghost_width();
        auto mesh   = samurai::mra::make_mesh(box, config);
        auto u      = samurai::make_scalar_field<double>("u", mesh);
        auto u1     = samurai::make_scalar_field<double>("u1", mesh);
        auto u2     = samurai::make_scalar_field<double>("u2", mesh);
        auto unp1   = samurai::make_scalar_field<double>("unp1", mesh);
        u.resize();
        const double t0 = 1e-2;
        samurai::for_each_cell(mesh,
                               [&](auto& cell)
                               {
                                   double r = 1;
                                   for (std::size_t d = 0; d < dim; ++d)
                                   {
                                       r *= 1 / (2 * std::sqrt(M_PI * K * t0)) * std::exp(-cell.center(d) * cell.center(d) / (4 * K * t0));
                                   }
                                   u[cell] = r;
                               });
        samurai::make_bc<samurai::Neumann<1>>(u, 0.);
        samurai::make_bc<samurai::Neumann<1>>(u1, 0.);
        samurai::make_bc<samurai::Neumann<1>>(u2, 0.);
        samurai::make_bc<samurai::Neumann<1>>(unp1, 0.);

        auto diff = samurai::make_diffusion_order2<decltype(u)>(K);

        const double dx = mesh.min_cell_length();
        const double dt = cfl * (dx * dx) / (std::pow(2, dim) * K);
        auto MRadaptation = samurai::make_MRAdapt(u);
        auto mra_config   = samurai::mra_config();
        MRadaptation(mra_config);
        for (std::size_t nt = 0; nt < n_steps; ++nt)
        {
            MRadaptation(mra_config);
            u1.resize();
            u2.resize();
            unp1.resize();
            // TVD-RK3 (SSPRK3), burgers.cpp:262-269
            u1   = u - dt * diff(u);
            u2   = 3. / 4 * u + 1. / 4 * (u1 - dt * diff(u1));
            unp1 = 1. / 3 * u + 2. / 3 * (u2 - dt * diff(u2));
            samurai::swap(u, unp1);
        }
        std::printf("leaves %zu\n", mesh.nb_cells());
        samurai::for_each_cell(mesh, [&](const auto& cell)
        {
            std::printf("%zu %d %d %.17g\n", cell.level, cell.indices[0], cell.indices[1], u[cell]);
        });
    }
    samurai::finalize();
    return 0;
}
