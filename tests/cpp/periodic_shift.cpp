// The reference's periodic test (tests/test_periodic.cpp:44-108) against the drop-in headers, body unchanged except that gtest's
// ASSERT_EQ is spelled out: on a mesh periodic in every direction (default mesh_config: ghost width 2) the field is shifted by one
// finest cell per step along the diagonal by a USER lambda over u(level, i - 1, index - 1) views, with MRadaptation before every
// step; after one full period it must be back on its initial state.
#include <xtensor/containers/xfixed.hpp>

#include <samurai/algorithm.hpp>
#include <samurai/field.hpp>
#include <samurai/mr/adapt.hpp>
#include <samurai/mr/mesh.hpp>
#include <samurai/samurai.hpp>

#include <cmath>
#include <cstdio>

template <class Mesh>
auto init(Mesh& mesh)
{
    double dx = mesh.cell_length(mesh.max_level());
    auto u    = samurai::make_scalar_field<double>("u", mesh);
    u.fill(0.);

    samurai::for_each_cell(mesh,
                           [&](auto& cell)
                           {
                               auto center   = cell.center();
                               double radius = std::floor(.2 / dx) * dx;

                               if (xt::all(xt::abs(center) <= radius))
                               {
                                   u[cell] = 1;
                               }
                           });

    return u;
}

template <std::size_t dim>
bool periodic_test()
{
    using namespace samurai;
    xt::xtensor_fixed<double, xt::xshape<dim>> min_corner, max_corner;
    min_corner.fill(-1);
    max_corner.fill(1);

    Box<double, dim> box(min_corner, max_corner);
    auto mesh_cfg   = samurai::mesh_config<dim>().min_level(3).max_level(dim == 3 ? 5 : 6).periodic(true).graduation_width(1);
    auto mesh       = samurai::mra::make_mesh(box, mesh_cfg);
    using mesh_id_t = typename decltype(mesh)::mesh_id_t;

    double dt = 1;
    double Tf = 2 / mesh.cell_length(mesh.max_level());
    double t  = 0.;

    auto u    = init(mesh);
    auto unp1 = make_scalar_field<double>("unp1", mesh);
    unp1.fill(0);

    auto MRadaptation = make_MRAdapt(u);
    auto mra_config   = samurai::mra_config();
    MRadaptation(mra_config);

    std::size_t min_leaves = mesh.nb_cells();
    while (t != Tf)
    {
        MRadaptation(mra_config);

        t += dt;
        if (t > Tf)
        {
            dt += Tf - t;
            t = Tf;
        }

        update_ghost_mr(u);
        unp1.resize();
        for_each_interval(mesh[mesh_id_t::cells],
                          [&](std::size_t level, auto& i, auto& index)
                          {
                              if constexpr (dim == 1)
                              {
                                  unp1(level, i) = u(level, i - 1);
                              }
                              else
                              {
                                  unp1(level, i, index) = u(level, i - 1, index - 1);
                              }
                          });

        std::swap(u.array(), unp1.array());
        min_leaves = std::min(min_leaves, mesh.nb_cells());
    }
    auto u_init = init(mesh);

    bool same = true;
    std::size_t n = 0;
    samurai::for_each_cell(mesh,
                           [&](auto& cell)
                           {
                               same = same && u[cell] == u_init[cell];
                               ++n;
                           });
    std::printf("dim %zu: %zu leaves (uniform would be %zu), %s\n", dim, n, std::size_t(1) << (dim * mesh.max_level()), same ? "back on the initial state" : "DIFFERENT");
    return same && n < (std::size_t(1) << (dim * mesh.max_level()));
}

int main(int argc, char* argv[])
{
    samurai::initialize(argc, argv);
    bool ok = periodic_test<1>();
    ok      = periodic_test<2>() && ok;
    ok      = periodic_test<3>() && ok;
    samurai::finalize();
    std::printf("%s\n", ok ? "periodic OK" : "periodic FAILED");
    return ok ? 0 : 1;
}
