// Two-component vector field through the drop-in headers: make_vector_field<double, 2>, u[cell][c], one boundary value per
// component, all components adapted together (make_MRAdapt(u)), shared ghost update, `unp1 = u - dt * upwind(a, u)`.
// Leaves are printed for the comparison with the oracle (tests/test_gpu_demos.py).
#include <samurai/mr/adapt.hpp>
#include <samurai/mr/mesh.hpp>
#include <samurai/samurai.hpp>
#include <samurai/stencil_field.hpp>

#include <cstdio>

int main(int argc, char* argv[])
{
    samurai::initialize("vector field advection", argc, argv);
    constexpr std::size_t dim = 2;
    const std::size_t n_steps = argc > 1 ? static_cast<std::size_t>(std::atoi(argv[1])) : 10;
    {
        samurai::Box<double, dim> box({0., 0.}, {1., 1.});
        auto config = samurai::mesh_config<dim>().min_level(2).max_level(7).max_stencil_size(2).disable_minimal_ghost_width();
        auto mesh   = samurai::mra::make_mesh(box, config);
        auto u      = samurai::make_vector_field<double, 2>("u", mesh);
        auto unp1   = samurai::make_vector_field<double, 2>("unp1", mesh);
        u.fill(0.);
        samurai::for_each_cell(mesh,
                               [&](auto& cell)
                               {
                                   const auto c   = cell.center();
                                   const double a = (c[0] - 0.3) * (c[0] - 0.3) + (c[1] - 0.3) * (c[1] - 0.3);
                                   const double b = (c[0] - 0.6) * (c[0] - 0.6) + (c[1] - 0.5) * (c[1] - 0.5);
                                   u[cell][0]     = a <= 0.2 * 0.2 ? 1. : 0.;
                                   u[cell][1]     = b <= 0.15 * 0.15 ? 2. : 0.;
                               });
        samurai::make_bc<samurai::Dirichlet<1>>(u, 0., 0.);
        samurai::make_bc<samurai::Dirichlet<1>>(unp1, 0., 0.);

        xt::xtensor_fixed<double, xt::xshape<dim>> a{1., 1.};
        const double dt   = 0.5 * mesh.min_cell_length();
        auto MRadaptation = samurai::make_MRAdapt(u);
        auto mra_config   = samurai::mra_config().epsilon(2e-4);
        MRadaptation(mra_config);
        for (std::size_t nt = 0; nt < n_steps; ++nt)
        {
            MRadaptation(mra_config);
            samurai::update_ghost_mr(u);
            unp1.resize();
            unp1 = u - dt * samurai::upwind(a, u);
            std::swap(u.array(), unp1.array());
        }
        std::printf("leaves %zu\n", mesh.nb_cells());
        samurai::for_each_cell(mesh, [&](const auto& cell)
        {
            std::printf("%zu %d %d %.17g %.17g\n", cell.level, cell.indices[0], cell.indices[1], u[cell][0], u[cell][1]);
        });
    }
    samurai::finalize();
    return 0;
}
