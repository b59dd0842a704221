// 2D vector Burgers equation with the flux-based upwind convection operator through the drop-in headers:
// `conv = make_convection_upwind<decltype(u)>()` on a two-component vector field (non-linear flux scheme, flux u(d) * u:
// schemes/fv/operators/convection_nonlin.hpp:24-76), `unp1 = u - dt * conv(u)` as in demos/FiniteVolume/burgers.cpp:262 (forward
// Euler line), both components adapted together.  Leaves are printed for the comparison with the oracle (tests/test_gpu_demos.py).
#include <samurai/mr/adapt.hpp>
#include <samurai/mr/mesh.hpp>
#include <samurai/samurai.hpp>
#include <samurai/schemes/fv.hpp>

#include <cstdio>

int main(int argc, char* argv[])
{
    samurai::initialize("vector Burgers, upwind flux scheme", argc, argv);
    constexpr std::size_t dim = 2;
    const std::size_t n_steps = argc > 1 ? static_cast<std::size_t>(std::atoi(argv[1])) : 10;
    {
        samurai::Box<double, dim> box({-1., -1.}, {1., 1.});
        auto config = samurai::mesh_config<dim>().min_level(2).max_level(6).max_stencil_size(2).disable_minimal_ghost_width();
        auto mesh   = samurai::mra::make_mesh(box, config);
        auto u      = samurai::make_vector_field<double, 2>("u", mesh);
        auto unp1   = samurai::make_vector_field<double, 2>("unp1", mesh);
        u.fill(0.);
        // "hat" of burgers.cpp:131-148 in the first component, a shifted negative one in the second: both upwinding signs occur
        samurai::for_each_cell(mesh,
                               [&](auto& cell)
                               {
                                   const auto c    = cell.center();
                                   const double r0 = std::max(std::abs(c[0]), std::abs(c[1]));
                                   const double r1 = std::max(std::abs(c[0] - 0.25), std::abs(c[1] + 0.25));
                                   u[cell][0]      = r0 < 0.5 ? 1. - 2. * r0 : 0.;
                                   u[cell][1]      = r1 < 0.4 ? -(1. - 2.5 * r1) : 0.;
                               });
        samurai::make_bc<samurai::Dirichlet<1>>(u, 0., 0.);
        samurai::make_bc<samurai::Dirichlet<1>>(unp1, 0., 0.);

        auto conv = samurai::make_convection_upwind<decltype(u)>();

        const double dt   = 0.4 * mesh.min_cell_length();
        auto MRadaptation = samurai::make_MRAdapt(u);
        auto mra_config   = samurai::mra_config().epsilon(1e-3);
        MRadaptation(mra_config);
        for (std::size_t nt = 0; nt < n_steps; ++nt)
        {
            MRadaptation(mra_config);
            unp1.resize();
            unp1 = u - dt * conv(u);
            samurai::swap(u, unp1);
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
