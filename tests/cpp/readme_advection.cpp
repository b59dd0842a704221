// The reference's README example (README.md:84-155), statement for statement, assembled into a program: the upwind scheme
// is the USER lambda over u(level, i, j) views, i.e. the host path of the drop-in (mesh adaptation and the ghost update run
// on the device).  Prints the leaves (level, i, j, value) so the test can compare with the oracle.
#include <samurai/samurai.hpp>
#include <samurai/mr/adapt.hpp>
#include <samurai/mr/mesh.hpp>
#include <samurai/field.hpp>
#include <samurai/bc.hpp>
#include <samurai/algorithm.hpp>

#include <cstdio>

int main(int argc, char* argv[])
{
    samurai::initialize(argc, argv);
    const std::size_t n_steps = argc > 1 ? static_cast<std::size_t>(std::atoi(argv[1])) : 50;
    {
        constexpr std::size_t dim = 2;
        using Config = samurai::MRConfig<dim>;
        std::size_t min_level = 2, max_level = 8;

        const samurai::Box<double, dim> box({0., 0.}, {1., 1.});
        samurai::MRMesh<Config> mesh(box, min_level, max_level);

        auto u = samurai::make_field<double, 1>("u", mesh);
        samurai::make_bc<samurai::Dirichlet<1>>(u, 0.);

        samurai::for_each_cell(mesh, [&](const auto& cell)
        {
            double length = 0.2;
            if (xt::all(xt::abs(cell.center() - 0.5) <= 0.5*length))
            {
                u[cell] = 1;
            }
        });

        auto MRadaptation = samurai::make_MRAdapt(u);

        double dx = mesh.cell_length(max_level);
        double dt = 0.5*dx;
        auto unp1 = samurai::make_field<double, 1>("u", mesh);

        // Time loop
        for (std::size_t nite = 0; nite < n_steps; ++nite)
        {
            // adapt u
            MRadaptation(1e-4, 2);

            // update the ghosts used by the upwind scheme
            samurai::update_ghost_mr(u);

            // upwind scheme
            samurai::for_each_interval(mesh, [&](std::size_t level, const auto& i, const auto& index)
            {
                double dx = mesh.cell_length(level);
                auto j = index[0];

                unp1(level, i, j) = u(level, i, j) - dt / dx * (u(level, i, j) - u(level, i - 1, j)
                                                              + u(level, i, j) - u(level, i, j - 1));
            });

            std::swap(unp1.array(), u.array());
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
