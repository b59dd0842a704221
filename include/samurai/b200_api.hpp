// samurai's public C++ API for the per-time-step hot path, bound to libsamurai_b200.so (include/samurai_b200.h).
//
// Source-compatible with what the reference's FV demos use (demos/FiniteVolume/advection_2d.cpp, advection_3d.cpp,
// scalar_burgers_2d.cpp compile unchanged against this include directory): samurai::initialize/finalize/app/SAMURAI_PARSE,
// Box, mesh_config, mra::make_mesh / make_empty_mesh, MRMesh, make_scalar_field, for_each_cell, make_bc<Dirichlet<1>>,
// make_MRAdapt, mra_config, update_ghost_mr, upwind, upwind_scalar_burgers, `unp1 = u - dt * op(a, u)`, save/dump.
// Fields are device resident; host accessors (u[cell]) work on a mirror with dirty tracking.  Expressions other than the
// recognised FV step forms do not compile: nothing silently runs on the CPU.
// Reference interfaces replaced: see the file:line notes on every declaration (paths relative to include/samurai/).
#pragma once
#include "../samurai_b200.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <fstream>
#include <functional>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "b200_h5.hpp"
#include <fmt/format.h>
#include <xtensor/containers/xfixed.hpp>

// ---------------------------------------------------------------------------------------------------------------------
// command line (the subset of CLI11 the demos touch through samurai::app)
// ---------------------------------------------------------------------------------------------------------------------
namespace CLI
{
    class Option
    {
      public:

        Option* capture_default_str()
        {
            return this;
        }

        Option* group(const std::string&)
        {
            return this;
        }

        std::string name, description;
        std::function<std::size_t(const std::vector<std::string>&, std::size_t)> parse; // consumes values, returns count
    };

    class App
    {
      public:

        template <class T>
        Option* add_option(const std::string& name, T& var, const std::string& desc = "")
        {
            auto it = m_options.find(name);
            if (it == m_options.end())
            {
                it = m_options.emplace(name, std::make_unique<Option>()).first;
            }
            Option* o      = it->second.get();
            o->name        = name;
            o->description = desc;
            o->parse       = [&var](const std::vector<std::string>& args, std::size_t i) -> std::size_t
            {
                return read(var, args, i);
            };
            return o;
        }

        Option* add_flag(const std::string& name, bool& var, const std::string& desc = "")
        {
            Option* o = add_option(name, var, desc);
            o->parse  = [&var](const std::vector<std::string>&, std::size_t) -> std::size_t
            {
                var = true;
                return 0;
            };
            return o;
        }

        void parse(int argc, char** argv)
        {
            std::vector<std::string> args(argv + 1, argv + argc);
            for (std::size_t i = 0; i < args.size();)
            {
                std::string key = args[i], inline_val;
                auto eq         = key.find('=');
                if (eq != std::string::npos)
                {
                    inline_val = key.substr(eq + 1);
                    key        = key.substr(0, eq);
                }
                if (key == "-h" || key == "--help")
                {
                    for (auto& kv : m_options)
                    {
                        std::cout << "  " << kv.first << "  " << kv.second->description << "\n";
                    }
                    std::exit(0);
                }
                auto it = m_options.find(key);
                if (it == m_options.end())
                {
                    if (m_allow_extras) // CLI::App::allow_extras(): unknown arguments are left alone (PETSc options in the demos)
                    {
                        ++i;
                        continue;
                    }
                    throw std::invalid_argument("The following argument was not expected: " + key);
                }
                if (!inline_val.empty())
                {
                    std::vector<std::string> one{inline_val};
                    it->second->parse(one, 0);
                    ++i;
                }
                else
                {
                    i += 1 + it->second->parse(args, i + 1);
                }
            }
        }

        App* allow_extras(bool allow = true)
        {
            m_allow_extras = allow;
            return this;
        }

        std::string description;

      private:

        template <class T>
        static void read_one(T& v, const std::string& s)
        {
            if constexpr (std::is_same_v<T, std::string>)
            {
                v = s;
            }
            else if constexpr (std::is_same_v<T, std::filesystem::path>)
            {
                v = std::filesystem::path(s);
            }
            else if constexpr (std::is_same_v<T, bool>)
            {
                v = (s == "1" || s == "true" || s == "on");
            }
            else if constexpr (std::is_floating_point_v<T>)
            {
                v = static_cast<T>(std::stod(s));
            }
            else
            {
                v = static_cast<T>(std::stoll(s));
            }
        }

        template <class T>
        static std::size_t read(T& var, const std::vector<std::string>& args, std::size_t i)
        {
            if (i >= args.size())
            {
                throw std::invalid_argument("missing value for an option");
            }
            read_one(var, args[i]);
            return 1;
        }

        template <class T, std::size_t N>
        static std::size_t read(std::array<T, N>& var, const std::vector<std::string>& args, std::size_t i)
        {
            std::size_t n = 0;
            while (n < N && i + n < args.size() && !(args[i + n].size() > 1 && args[i + n][0] == '-' && args[i + n][1] == '-'))
            {
                read_one(var[n], args[i + n]);
                ++n;
            }
            return n;
        }

        template <class T, std::size_t N>
        static std::size_t read(xt::xtensor_fixed<T, xt::xshape<N>>& var, const std::vector<std::string>& args, std::size_t i)
        {
            std::size_t n = 0;
            while (n < N && i + n < args.size() && !(args[i + n].size() > 1 && args[i + n][0] == '-' && args[i + n][1] == '-'))
            {
                read_one(var[n], args[i + n]);
                ++n;
            }
            return n;
        }

        std::map<std::string, std::unique_ptr<Option>> m_options;
        bool m_allow_extras = false;
    };
}

namespace samurai
{
    namespace fs = std::filesystem;

    namespace b200
    {
        inline void check(int rc)
        {
            if (rc == SMR_OK)
            {
                return;
            }
            const std::string msg = smr_last_error();
            if (rc == SMR_ERR_INVALID)
            {
                throw std::invalid_argument(msg);
            }
            if (rc == SMR_ERR_OUT_OF_RANGE)
            {
                throw std::out_of_range(msg);
            }
            throw std::runtime_error(msg);
        }
    }

    // ---- arguments.hpp:11-89 -------------------------------------------------------------------------------------------
    namespace args
    {
        inline std::size_t min_level        = std::numeric_limits<std::size_t>::max();
        inline std::size_t max_level        = std::numeric_limits<std::size_t>::max();
        inline std::size_t start_level      = std::numeric_limits<std::size_t>::max();
        inline std::size_t graduation_width = std::numeric_limits<std::size_t>::max();
        inline int max_stencil_radius       = std::numeric_limits<int>::max();
        inline double epsilon               = std::numeric_limits<double>::infinity();
        inline double regularity            = std::numeric_limits<double>::infinity();
        inline bool timers                  = false;
        inline bool rel_detail              = false;
        inline bool refine_boundary         = false;
    }

    // ---- samurai.hpp:22-114 --------------------------------------------------------------------------------------------
    inline CLI::App app;

    inline CLI::App& initialize(const std::string& description, int& /*argc*/, char**& /*argv*/)
    {
        app.description = description;
        app.add_option("--min-level", args::min_level, "The minimum level of the mesh");
        app.add_option("--max-level", args::max_level, "The maximum level of the mesh");
        app.add_option("--start-level", args::start_level, "Start level of AMR");
        app.add_option("--graduation-width", args::graduation_width, "The graduation width of the mesh");
        app.add_option("--max-stencil-radius", args::max_stencil_radius, "The maximum number of neighbour in each direction");
        app.add_option("--mr-eps", args::epsilon, "The epsilon used by the multiresolution to adapt the mesh");
        app.add_option("--mr-reg", args::regularity, "The regularity criteria used by the multiresolution to adapt the mesh");
        app.add_flag("--timers", args::timers, "Print timers at the end of the program");
        app.add_flag("--mr-rel-detail", args::rel_detail, "Use relative detail instead of absolute detail");
        app.add_flag("--refine-boundary", args::refine_boundary, "Keep the boundary refined at max_level");
        int device = 0;
        if (const char* e = std::getenv("SAMURAI_B200_DEVICE"))
        {
            device = std::atoi(e);
        }
        b200::check(smr_init(device));
        return app;
    }

    inline CLI::App& initialize(int& argc, char**& argv)
    {
        return initialize("SAMURAI", argc, argv);
    }

    inline void finalize()
    {
        if (args::timers)
        {
            smr_stats st;
            smr_stats_get(&st);
            std::cout << "samurai_b200 timers: device " << st.device_seconds << " s, host mesh " << st.host_mesh_seconds << " s, host batches "
                      << st.host_batch_seconds << " s, kernel launches " << st.kernel_launches << std::endl;
        }
        b200::check(smr_finalize());
    }

#define SAMURAI_PARSE(argc, argv)                \
    try                                          \
    {                                            \
        samurai::app.parse(argc, argv);          \
    }                                            \
    catch (const std::exception& e)              \
    {                                            \
        std::cerr << e.what() << std::endl;      \
        return 1;                                \
    }

    // ---- box.hpp -------------------------------------------------------------------------------------------------------
    template <class value_t, std::size_t dim_>
    class Box
    {
      public:

        static constexpr std::size_t dim = dim_;
        using point_t                    = xt::xtensor_fixed<value_t, xt::xshape<dim>>;

        Box() = default;

        template <class P1, class P2>
        Box(const P1& min_corner, const P2& max_corner)
        {
            for (std::size_t d = 0; d < dim; ++d)
            {
                m_min[d] = min_corner[d];
                m_max[d] = max_corner[d];
            }
        }

        // `Box<double, 2> box({0., 0.}, {1., 1.})` (README.md:96, demos/FiniteVolume/burgers_mra.cpp:265)
        Box(std::initializer_list<value_t> min_corner, std::initializer_list<value_t> max_corner)
        {
            if (min_corner.size() != dim || max_corner.size() != dim)
            {
                throw std::invalid_argument("Box corners must have `dim` coordinates");
            }
            std::size_t d = 0;
            for (auto v : min_corner)
            {
                m_min[d++] = v;
            }
            d = 0;
            for (auto v : max_corner)
            {
                m_max[d++] = v;
            }
        }

        const point_t& min_corner() const
        {
            return m_min;
        }

        const point_t& max_corner() const
        {
            return m_max;
        }

        point_t length() const
        {
            point_t l;
            for (std::size_t d = 0; d < dim; ++d)
            {
                l[d] = m_max[d] - m_min[d];
            }
            return l;
        }

      private:

        point_t m_min, m_max;
    };

    // ---- mesh_config.hpp:20-432 ----------------------------------------------------------------------------------------
    struct default_interval_t
    {
    };

    template <std::size_t dim_, int prediction_stencil_radius_ = 1, std::size_t max_refinement_level_ = 20, class interval_t_ = default_interval_t>
    class mesh_config
    {
      public:

        static constexpr std::size_t dim                  = dim_;
        static constexpr int prediction_stencil_radius    = prediction_stencil_radius_;
        static constexpr std::size_t max_refinement_level = max_refinement_level_;

        auto& max_stencil_radius(int r)
        {
            m_max_stencil_radius = r;
            return *this;
        }

        int max_stencil_radius() const
        {
            return m_max_stencil_radius;
        }

        auto& max_stencil_size(int s)
        {
            m_max_stencil_radius = s / 2 + (s % 2);
            return *this;
        }

        auto& graduation_width(std::size_t w)
        {
            m_graduation_width = w;
            return *this;
        }

        std::size_t graduation_width() const
        {
            return m_graduation_width;
        }

        int ghost_width() const
        {
            return m_ghost_width;
        }

        auto& min_level(std::size_t l)
        {
            m_min_level = l;
            return *this;
        }

        std::size_t min_level() const
        {
            return m_min_level;
        }

        auto& max_level(std::size_t l)
        {
            m_max_level = l;
            return *this;
        }

        std::size_t max_level() const
        {
            return m_max_level;
        }

        auto& start_level(std::size_t l)
        {
            m_start_level = l;
            return *this;
        }

        std::size_t& start_level()
        {
            return m_start_level;
        }

        std::size_t start_level() const
        {
            return m_start_level;
        }

        auto& approx_box_tol(double t)
        {
            m_approx_box_tol = t;
            return *this;
        }

        auto& scaling_factor(double s)
        {
            m_scaling_factor = s;
            return *this;
        }

        double scaling_factor() const
        {
            return m_scaling_factor;
        }

        auto& disable_args_parse()
        {
            m_disable_args_parse = true;
            return *this;
        }

        auto& disable_minimal_ghost_width()
        {
            m_disable_minimal_ghost_width = true;
            return *this;
        }

        // mesh_config.hpp:224-253.  Periodic meshes are not on the device path yet: the flags are kept
        // so that demos which expose `--periodic` compile, and a periodic mesh is refused when it is built.
        auto& periodic(bool p)
        {
            m_periodic.fill(p);
            return *this;
        }

        auto& periodic(const std::array<bool, dim_>& p)
        {
            m_periodic = p;
            return *this;
        }

        const std::array<bool, dim_>& periodic() const
        {
            return m_periodic;
        }

        bool periodic(std::size_t d) const
        {
            return m_periodic[d];
        }

        void parse_args() // mesh_config.hpp:358-396
        {
            if (!m_disable_args_parse)
            {
                if (args::max_stencil_radius != std::numeric_limits<int>::max())
                {
                    m_max_stencil_radius = args::max_stencil_radius;
                }
                if (args::graduation_width != std::numeric_limits<std::size_t>::max())
                {
                    m_graduation_width = args::graduation_width;
                }
                if (args::min_level != std::numeric_limits<std::size_t>::max())
                {
                    m_min_level = args::min_level;
                }
                if (args::max_level != std::numeric_limits<std::size_t>::max())
                {
                    m_max_level = args::max_level;
                }
                if (args::start_level != std::numeric_limits<std::size_t>::max())
                {
                    m_start_level = args::start_level;
                }
                if (m_max_level < m_min_level)
                {
                    throw std::invalid_argument("Max level must be greater than min level.");
                }
            }
            if (!m_disable_minimal_ghost_width)
            {
                m_max_stencil_radius = std::max(m_max_stencil_radius, 2);
            }
            m_ghost_width = std::max(m_max_stencil_radius, static_cast<int>(prediction_stencil_radius));
        }

      private:

        int m_max_stencil_radius       = 1;
        std::size_t m_graduation_width = 1;
        int m_ghost_width              = 1;
        std::size_t m_min_level        = 0;
        std::size_t m_max_level        = 6;
        std::size_t m_start_level      = 6;
        double m_approx_box_tol        = 0.05;
        double m_scaling_factor        = 0;
        bool m_disable_args_parse          = false;
        bool m_disable_minimal_ghost_width = false;
        std::array<bool, dim_> m_periodic{};
    };

    // ---- mr/mesh.hpp:25-34 ---------------------------------------------------------------------------------------------
    enum class MRMeshId
    {
        cells            = 0,
        cells_and_ghosts = 1,
        proj_cells       = 2,
        union_cells      = 3,
        reference        = 4,
        count            = 5,
        all_cells        = reference
    };

    // ---- interval.hpp:50-64, cell.hpp:32-77 ------------------------------------------------------------------------------
    struct Interval
    {
        int start = 0, end = 0, step = 1;
        long long index = 0; // storage index: value of cell x lives at index + x

        std::size_t size() const
        {
            return static_cast<std::size_t>(end - start);
        }
    };

    // interval.hpp:262-300: translation of an interval (`i - 1`, `i + 1` in stencil expressions)
    inline Interval operator+(Interval i, int s)
    {
        i.start += s;
        i.end += s;
        i.index -= s; // the same storage cells are no longer addressed: callers re-resolve the row (ScalarField::operator())
        return i;
    }

    inline Interval operator-(Interval i, int s)
    {
        return i + (-s);
    }

    // ---- host-side row expressions -----------------------------------------------------------------------------------------
    // `u(level, i, j)` (field/access_base.hpp:69-103) is a view of one x-interval of the field; the reference combines such views
    // with xtensor expression templates inside user lambdas (README.md:144-151).  This is the HOST path of the drop-in: the views
    // address the field's host mirror (downloaded once after the device last wrote the field, uploaded before the device next reads
    // it).  The device kernels never go through it.
    namespace rowx
    {
        template <class E>
        struct expr
        {
            const E& self() const
            {
                return static_cast<const E&>(*this);
            }
        };

        struct scalar : expr<scalar>
        {
            double v;

            explicit scalar(double x)
                : v(x)
            {
            }

            double operator[](std::size_t) const
            {
                return v;
            }

            std::size_t size() const
            {
                return 0;
            }
        };

        template <class A, class B, class Op>
        struct binary : expr<binary<A, B, Op>>
        {
            A a;
            B b;

            binary(const A& a_, const B& b_)
                : a(a_)
                , b(b_)
            {
            }

            double operator[](std::size_t k) const
            {
                return Op::apply(a[k], b[k]);
            }

            std::size_t size() const
            {
                return std::max(a.size(), b.size());
            }
        };

        struct add
        {
            static double apply(double x, double y)
            {
                return x + y;
            }
        };

        struct sub
        {
            static double apply(double x, double y)
            {
                return x - y;
            }
        };

        struct mul
        {
            static double apply(double x, double y)
            {
                return x * y;
            }
        };

        struct div
        {
            static double apply(double x, double y)
            {
                return x / y;
            }
        };

        struct view : expr<view>
        {
            double* p     = nullptr;
            std::size_t n = 0;

            view(double* p_, std::size_t n_)
                : p(p_)
                , n(n_)
            {
            }

            view(const view&) = default;

            double operator[](std::size_t k) const
            {
                return p[k];
            }

            double& operator[](std::size_t k)
            {
                return p[k];
            }

            std::size_t size() const
            {
                return n;
            }

            template <class E>
            view& operator=(const expr<E>& e)
            {
                const E& x = e.self();
                for (std::size_t k = 0; k < n; ++k)
                {
                    p[k] = x[k];
                }
                return *this;
            }

            view& operator=(const view& o)
            {
                for (std::size_t k = 0; k < n; ++k)
                {
                    p[k] = o.p[k];
                }
                return *this;
            }

            view& operator=(double v)
            {
                for (std::size_t k = 0; k < n; ++k)
                {
                    p[k] = v;
                }
                return *this;
            }

            template <class E>
            view& operator+=(const expr<E>& e)
            {
                const E& x = e.self();
                for (std::size_t k = 0; k < n; ++k)
                {
                    p[k] += x[k];
                }
                return *this;
            }

            template <class E>
            view& operator-=(const expr<E>& e)
            {
                const E& x = e.self();
                for (std::size_t k = 0; k < n; ++k)
                {
                    p[k] -= x[k];
                }
                return *this;
            }
        };

#define SMR_ROWX_OP(SYM, NAME)                                                             \
    template <class A, class B>                                                            \
    auto operator SYM(const expr<A>& a, const expr<B>& b)                                  \
    {                                                                                      \
        return binary<A, B, NAME>(a.self(), b.self());                                     \
    }                                                                                      \
    template <class A, class S, class = std::enable_if_t<std::is_arithmetic_v<S>>>         \
    auto operator SYM(const expr<A>& a, S b)                                               \
    {                                                                                      \
        return binary<A, scalar, NAME>(a.self(), scalar(static_cast<double>(b)));          \
    }                                                                                      \
    template <class B, class S, class = std::enable_if_t<std::is_arithmetic_v<S>>>         \
    auto operator SYM(S a, const expr<B>& b)                                               \
    {                                                                                      \
        return binary<scalar, B, NAME>(scalar(static_cast<double>(a)), b.self());          \
    }
        SMR_ROWX_OP(+, add)
        SMR_ROWX_OP(-, sub)
        SMR_ROWX_OP(*, mul)
        SMR_ROWX_OP(/, div)
#undef SMR_ROWX_OP

        template <class A>
        auto operator-(const expr<A>& a)
        {
            return binary<scalar, A, sub>(scalar(0.0), a.self());
        }
    } // namespace rowx

    template <std::size_t dim_>
    struct Cell
    {
        static constexpr std::size_t dim = dim_;
        using coords_t                   = xt::xtensor_fixed<double, xt::xshape<dim>>;

        std::size_t level = 0;
        xt::xtensor_fixed<int, xt::xshape<dim>> indices;
        long long index = 0;
        double length   = 0;
        coords_t origin_point;

        coords_t center() const // cell.hpp:131-134
        {
            coords_t c;
            for (std::size_t d = 0; d < dim; ++d)
            {
                c[d] = origin_point[d] + length * (indices[d] + 0.5);
            }
            return c;
        }

        double center(std::size_t d) const
        {
            return origin_point[d] + length * (indices[d] + 0.5);
        }

        coords_t corner() const
        {
            coords_t c;
            for (std::size_t d = 0; d < dim; ++d)
            {
                c[d] = origin_point[d] + length * indices[d];
            }
            return c;
        }
    };

    // ---- mr/mesh.hpp:74-122, mesh.hpp ------------------------------------------------------------------------------------
    template <class Config>
    class MRMesh
    {
      public:

        using config_t                   = Config;
        using mesh_id_t                  = MRMeshId;
        using cell_t                     = Cell<Config::dim>;
        static constexpr std::size_t dim = Config::dim;

        MRMesh() = default;

        MRMesh(const Box<double, dim>& b, const Config& cfg)
            : m_cfg(cfg)
        {
            smr_mesh_config c{};
            for (std::size_t d = 0; d < dim; ++d)
            {
                c.periodic[d] = cfg.periodic(d) ? 1 : 0;
            }
            c.dim                = static_cast<int32_t>(dim);
            c.min_level          = static_cast<int32_t>(cfg.min_level());
            c.max_level          = static_cast<int32_t>(cfg.max_level());
            c.pred_radius        = Config::prediction_stencil_radius;
            c.max_stencil_radius = cfg.max_stencil_radius();
            c.refine_boundary    = args::refine_boundary ? 1 : 0;
            c.graduation_width   = static_cast<int32_t>(cfg.graduation_width());
            // approximate_box (box.hpp:280-360) for boxes whose edge lengths are integer multiples of the smallest one
            const auto len = b.length();
            double s       = cfg.scaling_factor();
            if (s <= 0)
            {
                s = len[0];
                for (std::size_t d = 1; d < dim; ++d)
                {
                    s = std::min(s, len[d]);
                }
            }
            for (std::size_t d = 0; d < 3; ++d)
            {
                if (d < dim)
                {
                    const double n = len[d] / s;
                    if (std::abs(n - std::round(n)) > 1e-12)
                    {
                        throw std::invalid_argument("box edge lengths must be integer multiples of the cell length at level 0");
                    }
                    c.n_cells0[d] = static_cast<int32_t>(std::lround(n));
                    c.origin[d]   = b.min_corner()[d];
                }
                else
                {
                    c.n_cells0[d] = 1;
                    c.origin[d]   = 0;
                }
            }
            c.scaling_factor = s;
            m_c              = c;
            smr_mesh_t h     = 0;
            b200::check(smr_mesh_create_uniform(&c, static_cast<int>(cfg.start_level()), &h));
            m_owner = std::make_shared<Owner>();
            m_owner->h = h;
        }

        // `samurai::MRMesh<Config> mesh(box, min_level, max_level);` -- the constructor the reference's README still shows
        // (README.md:96-100); stencil width 1, no minimal ghost width, starts uniform at max_level
        MRMesh(const Box<double, dim>& b, std::size_t min_level, std::size_t max_level)
            : MRMesh(b, legacy_config(min_level, max_level))
        {
        }

        static Config legacy_config(std::size_t min_level, std::size_t max_level)
        {
            Config cfg;
            cfg.min_level(min_level).max_level(max_level).max_stencil_radius(1).disable_minimal_ghost_width();
            cfg.parse_args();
            cfg.start_level() = cfg.max_level();
            return cfg;
        }

        // samurai::load(): replace this mesh by the one whose leaves are given as rows (level, y, z, start, end); `c` comes from
        // the checkpoint (the object may be an empty mesh made by make_empty_mesh)
        void rebuild_from_leaves(const smr_mesh_config& c, const std::vector<int64_t>& ivl5)
        {
            const std::size_t n = ivl5.size() / 5;
            std::vector<int32_t> levels(n);
            std::vector<smr_interval> ivl(n);
            for (std::size_t i = 0; i < n; ++i)
            {
                levels[i] = static_cast<int32_t>(ivl5[5 * i]);
                ivl[i]    = smr_interval{static_cast<int32_t>(ivl5[5 * i + 1]), static_cast<int32_t>(ivl5[5 * i + 2]), static_cast<int32_t>(ivl5[5 * i + 3]),
                                      static_cast<int32_t>(ivl5[5 * i + 4]), 0};
            }
            smr_mesh_t h = 0;
            b200::check(smr_mesh_create_from_intervals(&c, levels.data(), ivl.data(), static_cast<int64_t>(n), &h));
            m_c = c;
            m_cfg.min_level(static_cast<std::size_t>(c.min_level)).max_level(static_cast<std::size_t>(c.max_level));
            m_owner    = std::make_shared<Owner>();
            m_owner->h = h;
        }

        // Copies share the library mesh (the demos copy a mesh to iterate over it, scalar_burgers_2d.cpp:23);
        // `mesh = samurai::mra::make_mesh(box, config);` (advection_2d.cpp:107) rebinds this object, and the fields that
        // point at it follow.
        MRMesh(const MRMesh&)                = default;
        MRMesh& operator=(const MRMesh&)     = default;
        MRMesh(MRMesh&&) noexcept            = default;
        MRMesh& operator=(MRMesh&&) noexcept = default;

        smr_mesh_t handle() const
        {
            return m_owner ? m_owner->h : 0;
        }

        // mesh[mesh_id_t::cells]: one of the five sub-meshes as an iterable (mesh.hpp:560-580): for_each_interval / for_each_cell
        struct SubMesh
        {
            static constexpr std::size_t dim = Config::dim;
            using cell_t                     = Cell<Config::dim>;
            const MRMesh* mesh;
            MRMeshId id;

            std::size_t max_level() const
            {
                return mesh->max_level() + 2;
            }

            auto intervals(MRMeshId, std::size_t level) const
            {
                return mesh->intervals(id, level);
            }

            std::size_t nb_cells() const
            {
                return mesh->nb_cells(id);
            }

            double cell_length(std::size_t level) const
            {
                return mesh->cell_length(level);
            }

            const smr_mesh_config& c_config() const
            {
                return mesh->c_config();
            }
        };

        SubMesh operator[](MRMeshId id) const
        {
            return SubMesh{this, id};
        }

        std::size_t min_level() const
        {
            return m_cfg.min_level();
        }

        std::size_t max_level() const
        {
            return m_cfg.max_level();
        }

        double scaling_factor() const
        {
            return m_c.scaling_factor;
        }

        double cell_length(std::size_t level) const
        {
            return m_c.scaling_factor / static_cast<double>(1 << level);
        }

        double min_cell_length() const
        {
            return cell_length(max_level());
        }

        std::size_t nb_cells(mesh_id_t id = mesh_id_t::cells) const
        {
            int64_t n = 0;
            b200::check(smr_mesh_nb_cells(handle(), static_cast<int>(id), -1, &n));
            return static_cast<std::size_t>(n);
        }

        std::size_t nb_cells(std::size_t level, mesh_id_t id = mesh_id_t::cells) const
        {
            int64_t n = 0;
            b200::check(smr_mesh_nb_cells(handle(), static_cast<int>(id), static_cast<int>(level), &n));
            return static_cast<std::size_t>(n);
        }

        std::vector<smr_interval> intervals(mesh_id_t id, std::size_t level) const
        {
            int64_t n = 0;
            b200::check(smr_mesh_nb_intervals(handle(), static_cast<int>(id), static_cast<int>(level), &n));
            std::vector<smr_interval> v(static_cast<std::size_t>(n));
            if (n)
            {
                b200::check(smr_mesh_get_intervals(handle(), static_cast<int>(id), static_cast<int>(level), v.data()));
            }
            return v;
        }

        const smr_mesh_config& c_config() const
        {
            return m_c;
        }

      private:

        struct Owner
        {
            smr_mesh_t h = 0;

            ~Owner()
            {
                // refused while fields are still attached; whatever is left is reclaimed by smr_finalize()
                if (h)
                {
                    smr_mesh_destroy(h);
                }
            }
        };

        std::shared_ptr<Owner> m_owner;
        Config m_cfg;
        smr_mesh_config m_c{};
    };

    // README.md:90: `using Config = samurai::MRConfig<dim>;`
    template <std::size_t dim_, std::size_t max_stencil_width_ = 1, std::size_t graduation_width_ = 1, std::size_t prediction_order_ = 1>
    using MRConfig = mesh_config<dim_, static_cast<int>(prediction_order_)>;

    namespace mra
    {
        template <class mesh_config_t>
        auto make_empty_mesh(const mesh_config_t&)
        {
            return MRMesh<mesh_config_t>();
        }

        template <class mesh_config_t>
        auto make_mesh(const Box<double, mesh_config_t::dim>& b, const mesh_config_t& cfg) // mr/mesh.hpp:510-518
        {
            auto mesh_cfg = cfg;
            mesh_cfg.parse_args();
            mesh_cfg.start_level() = mesh_cfg.max_level();
            return MRMesh<mesh_config_t>(b, mesh_cfg);
        }
    }

    // ---- algorithm.hpp:53-363 ------------------------------------------------------------------------------------------
    template <class Mesh, class Func>
    void for_each_interval(const Mesh& mesh, Func&& f) // f(level, interval, index)  (leaves)
    {
        for (std::size_t level = 0; level <= mesh.max_level(); ++level)
        {
            for (const auto& iv : mesh.intervals(MRMeshId::cells, level))
            {
                Interval i{iv.start, iv.end, 1, static_cast<long long>(iv.offset) - iv.start};
                xt::xtensor_fixed<int, xt::xshape<(Mesh::dim > 1 ? Mesh::dim - 1 : 1)>> index;
                index[0] = iv.y;
                if constexpr (Mesh::dim > 2)
                {
                    index[1] = iv.z;
                }
                f(level, i, index);
            }
        }
    }

    template <class Mesh, class Func>
    void for_each_cell(const Mesh& mesh, Func&& f)
    {
        typename Mesh::cell_t cell;
        for (std::size_t d = 0; d < Mesh::dim; ++d)
        {
            cell.origin_point[d] = mesh.c_config().origin[d];
        }
        for (std::size_t level = 0; level <= mesh.max_level(); ++level)
        {
            cell.level  = level;
            cell.length = mesh.cell_length(level);
            for (const auto& iv : mesh.intervals(MRMeshId::cells, level))
            {
                if constexpr (Mesh::dim > 1)
                {
                    cell.indices[1] = iv.y;
                }
                if constexpr (Mesh::dim > 2)
                {
                    cell.indices[2] = iv.z;
                }
                for (int x = iv.start; x < iv.end; ++x)
                {
                    cell.indices[0] = x;
                    cell.index      = iv.offset + (x - iv.start);
                    f(cell);
                }
            }
        }
    }

    // ---- field/scalar_field.hpp ------------------------------------------------------------------------------------------
    // Storage = what `u.array()` returns: std::swap(u.array(), unp1.array()) exchanges the device buffers (advection_2d.cpp:146)
    struct FieldStorage
    {
        smr_field_t handle = 0;
        std::vector<double> host; // mirror for u[cell]
        bool host_valid  = false; // mirror equals the device copy
        bool host_dirty  = false; // mirror was written and not uploaded yet
    };

    struct Dirichlet1Tag
    {
    };

    template <std::size_t order = 1>
    struct Dirichlet
    {
        static constexpr int type = SMR_BCTYPE_DIRICHLET;
        static_assert(order == 1, "only Dirichlet<1> is implemented on the device path");
    };

    template <std::size_t order = 1>
    struct Neumann
    {
        static constexpr int type = SMR_BCTYPE_NEUMANN;
        static_assert(order == 1, "only Neumann<1> is implemented on the device path");
    };

    template <class A, class Field>
    struct upwind_expr
    {
        A a;
        const Field* u;
        bool burgers;
    };

    template <class A, class Field>
    struct scaled_upwind_expr
    {
        double dt;
        upwind_expr<A, Field> op;
    };

    template <class A, class Field>
    struct fv_step_expr
    {
        const Field* u;
        scaled_upwind_expr<A, Field> rhs;
    };

    // ---- flux-based schemes (schemes/fv/FV_scheme.hpp:202-238, flux_based/flux_based_scheme.hpp) ---------------------------
    template <class Field>
    class FluxBasedScheme;

    template <class Field>
    struct scheme_expr // scheme(u)
    {
        FluxBasedScheme<Field> scheme;
        Field* u;
    };

    template <class Field>
    struct scaled_scheme_expr // dt * scheme(u)
    {
        double factor;
        scheme_expr<Field> e;
    };

    template <class Field>
    struct scheme_step_expr // v - dt * scheme(u)
    {
        const Field* v;
        scaled_scheme_expr<Field> rhs;
    };

    // general field expressions (field/field_expression.hpp:62-141), evaluated on the device: see namespace fx below
    namespace fx
    {
        template <class E>
        struct node
        {
            const E& self() const
            {
                return static_cast<const E&>(*this);
            }
        };
    }

    template <class mesh_t_, class value_t = double>
    class ScalarField
    {
      public:

        using mesh_t                     = mesh_t_;
        using cell_t                     = typename mesh_t::cell_t;
        static constexpr std::size_t dim = mesh_t::dim;
        static constexpr bool on_device  = std::is_same_v<value_t, double>;

        ScalarField(std::string name, mesh_t& mesh)
            : m_name(std::move(name))
            , p_mesh(&mesh)
        {
        }

        // component of a vector field: same object, storage owned by the caller
        ScalarField(std::string name, mesh_t& mesh, FieldStorage* external)
            : m_name(std::move(name))
            , p_mesh(&mesh)
            , p_ext(external)
        {
        }

        FieldStorage& st()
        {
            return p_ext ? *p_ext : m_storage;
        }

        const FieldStorage& st() const
        {
            return p_ext ? *p_ext : m_storage;
        }

        ScalarField(const ScalarField&)            = delete;
        ScalarField& operator=(const ScalarField&) = delete;

        ScalarField(ScalarField&& o) noexcept
            : m_name(std::move(o.m_name))
            , p_mesh(o.p_mesh)
            , m_storage(std::move(o.m_storage))
            , m_plain(std::move(o.m_plain))
            , m_bc_type(o.m_bc_type)
            , m_bc_value(o.m_bc_value)
        {
            o.m_storage.handle = 0;
        }

        ~ScalarField()
        {
            if (st().handle)
            {
                smr_field_destroy(st().handle);
            }
        }

        const std::string& name() const
        {
            return m_name;
        }

        mesh_t& mesh()
        {
            return *p_mesh;
        }

        const mesh_t& mesh() const
        {
            return *p_mesh;
        }

        FieldStorage& array()
        {
            return st();
        }

        std::size_t size() const
        {
            return p_mesh->nb_cells(MRMeshId::reference);
        }

        void resize() // field/access_base.hpp:105-115
        {
            if constexpr (on_device)
            {
                push_host();
                b200::check(smr_field_resize(handle()));
                st().host_valid = false;
            }
            else
            {
                m_plain.resize(size());
            }
        }

        void fill(value_t v)
        {
            if constexpr (on_device)
            {
                b200::check(smr_field_resize(handle()));
                b200::check(smr_field_fill(handle(), v));
                st().host_valid = false;
                st().host_dirty = false;
            }
            else
            {
                m_plain.assign(size(), v);
            }
        }

        // host access: u[cell]
        value_t& operator[](const cell_t& cell)
        {
            if constexpr (on_device)
            {
                pull_host();
                st().host_dirty = true;
                return st().host[static_cast<std::size_t>(cell.index)];
            }
            else
            {
                if (m_plain.size() != size())
                {
                    m_plain.resize(size());
                }
                return m_plain[static_cast<std::size_t>(cell.index)];
            }
        }

        value_t operator[](const cell_t& cell) const
        {
            if constexpr (on_device)
            {
                const_cast<ScalarField*>(this)->pull_host();
                return st().host[static_cast<std::size_t>(cell.index)];
            }
            else
            {
                return m_plain[static_cast<std::size_t>(cell.index)];
            }
        }

        // host access: u(level, i), u(level, i, j), u(level, i, j, k), u(level, i, index)  (field/access_base.hpp:69-103).
        // The interval may be translated (`i - 1`), the row too (`j - 1`): the storage offset is looked up in the reference
        // sub-mesh (ghosts included), an absent row throws like the reference's debug build does.
        rowx::view operator()(std::size_t level, const Interval& i, int j = 0, int k = 0)
        {
            static_assert(on_device, "u(level, i, ...) needs a double field");
            pull_host();
            st().host_dirty = true;
            return rowx::view(st().host.data() + row_offset(level, i, j, k), i.size());
        }

        rowx::view operator()(std::size_t level, const Interval& i, int j = 0, int k = 0) const
        {
            static_assert(on_device, "u(level, i, ...) needs a double field");
            auto* self = const_cast<ScalarField*>(this);
            self->pull_host();
            return rowx::view(self->st().host.data() + row_offset(level, i, j, k), i.size());
        }

        template <std::size_t N>
        rowx::view operator()(std::size_t level, const Interval& i, const xt::xtensor_fixed<int, xt::xshape<N>>& index)
        {
            return (*this)(level, i, dim > 1 ? index[0] : 0, (dim > 2 && N > 1) ? index[N > 1 ? 1 : 0] : 0);
        }

        template <std::size_t N>
        rowx::view operator()(std::size_t level, const Interval& i, const xt::xtensor_fixed<int, xt::xshape<N>>& index) const
        {
            return (*this)(level, i, dim > 1 ? index[0] : 0, (dim > 2 && N > 1) ? index[N > 1 ? 1 : 0] : 0);
        }

        // device handle with pending host writes flushed and the boundary condition attached
        smr_field_t device() const
        {
            auto* self = const_cast<ScalarField*>(this);
            self->push_host();
            if (m_bc_type >= 0)
            {
                b200::check(smr_field_set_bc(self->handle(), m_bc_type, m_bc_value));
            }
            return self->handle();
        }

        void device_written()
        {
            st().host_valid = false;
            st().host_dirty = false;
        }

        void attach_bc(int type, double value)
        {
            m_bc_type  = type;
            m_bc_value = value;
        }

        // `unp1 = u - dt * upwind(a, u)`  (field/field_base.hpp:230-242 + stencil_field.hpp)
        template <class A>
        ScalarField& operator=(const fv_step_expr<A, ScalarField>& e)
        {
            static_assert(on_device, "FV expressions need a double field");
            if (e.u != e.rhs.op.u)
            {
                throw std::invalid_argument("only `u - dt * op(a, u)` with the same field u is recognised on the device path");
            }
            double a[3] = {0, 0, 0};
            if constexpr (std::is_arithmetic_v<A>) // 1D: `upwind(a, u)` with a scalar velocity (advection_1d.cpp:148)
            {
                a[0] = e.rhs.op.a;
            }
            else
            {
                for (std::size_t d = 0; d < dim; ++d)
                {
                    a[d] = e.rhs.op.a[d];
                }
            }
            const smr_field_t in = e.u->device();
            b200::check(smr_field_resize(handle()));
            if (e.rhs.op.burgers)
            {
                b200::check(smr_fv_upwind_burgers(handle(), in, a, e.rhs.dt));
            }
            else
            {
                b200::check(smr_fv_upwind(handle(), in, a, e.rhs.dt));
            }
            device_written();
            return *this;
        }

        // `rhs = scheme(u)` (explicit application, out.fill(0) first: schemes/fv/explicit_FV_scheme.hpp)
        ScalarField& operator=(const scheme_expr<ScalarField>& e)
        {
            static_assert(on_device, "FV schemes need a double field");
            e.scheme.apply(*this, *e.u);
            return *this;
        }

        // `unp1 = v - dt * scheme(u)` (field expression over the leaves, field/field_base.hpp:230-242)
        ScalarField& operator=(const scheme_step_expr<ScalarField>& e)
        {
            static_assert(on_device, "FV schemes need a double field");
            ScalarField tmp(e.rhs.e.scheme.name() + "(" + e.rhs.e.u->name() + ")", *p_mesh);
            e.rhs.e.scheme.apply(tmp, *e.rhs.e.u);
            b200::check(smr_field_resize(handle()));
            b200::check(smr_field_lincomb(handle(), 1.0, e.v->device(), -e.rhs.factor, tmp.device()));
            device_written();
            return *this;
        }

        // `unp1 = 1./3 * u + 2./3 * (u2 - dt * conv(u2))` and friends: any tree of +, -, scalar * over fields and scheme(u)
        template <class E>
        ScalarField& operator=(const fx::node<E>& e)
        {
            static_assert(on_device, "field expressions need a double field");
            e.self().assign_to(*this);
            return *this;
        }

        std::vector<double> leaf_values() const
        {
            const_cast<ScalarField*>(this)->pull_host();
            return st().host;
        }

      private:

        std::size_t row_offset(std::size_t level, const Interval& i, int j, int k) const
        {
            int64_t first = -1, last = -1;
            if (smr_mesh_get_index(p_mesh->handle(), static_cast<int>(level), i.start, j, k, &first) != SMR_OK || first < 0
                || smr_mesh_get_index(p_mesh->handle(), static_cast<int>(level), i.end - 1, j, k, &last) != SMR_OK
                || last - first != static_cast<int64_t>(i.size()) - 1)
            {
                throw std::out_of_range("field '" + m_name + "': interval [" + std::to_string(i.start) + ", " + std::to_string(i.end) + ") of row ("
                                        + std::to_string(j) + ", " + std::to_string(k) + ") at level " + std::to_string(level)
                                        + " is not in the reference mesh");
            }
            return static_cast<std::size_t>(first);
        }

        smr_field_t handle()
        {
            if (!st().handle)
            {
                if (!p_mesh->handle())
                {
                    throw std::runtime_error("field '" + m_name + "' used before its mesh was built");
                }
                b200::check(smr_field_create(p_mesh->handle(), m_name.c_str(), &st().handle));
            }
            return st().handle;
        }

        void pull_host()
        {
            const std::size_t n = size();
            if (!st().host_valid || st().host.size() != n)
            {
                int64_t fs_ = 0;
                b200::check(smr_field_resize(handle()));
                b200::check(smr_field_size(handle(), &fs_));
                st().host.resize(n);
                b200::check(smr_field_download(handle(), st().host.data(), static_cast<int64_t>(n)));
                st().host_valid = true;
                st().host_dirty = false;
            }
        }

        void push_host()
        {
            if (st().host_dirty && st().host.size() != size())
            {
                // the mirror was taken on a previous mesh (a field that was only read through u(level, i, ...) or that is about to be
                // resized): nothing of it addresses the current numbering
                st().host_dirty = false;
                st().host_valid = false;
            }
            if (st().host_dirty)
            {
                b200::check(smr_field_upload(handle(), st().host.data(), static_cast<int64_t>(st().host.size())));
                b200::check(smr_synchronize());
                st().host_dirty = false;
                st().host_valid = true;
            }
        }

        std::string m_name;
        mesh_t* p_mesh;
        FieldStorage m_storage;
        FieldStorage* p_ext = nullptr; // component of a VectorField: the storage lives in the vector field (so that swapping its array() swaps every component)
        std::vector<value_t> m_plain; // non-double fields are host-only (e.g. the `level` field the demos save)
        int m_bc_type     = -1;
        double m_bc_value = 0;
    };

    // ---- field/vector_field.hpp:234 ------------------------------------------------------------------------------------------
    // A vector field is stored as n_comp scalar components (SoA: one device array per component, as north_star asks), each
    // with the reference's storage numbering.  The host view the user code sees is the reference's: `u[cell]` is a small
    // vector (u[cell][c], u[cell] = {..}), `u(c, level, i, j)` the row of one component.  All components are adapted together
    // (one tag array, the criteria takes the max over the components: mr/operators.hpp:623-677, mr/criteria.hpp) and share the
    // ghost update.
    template <class mesh_t_, class value_t, std::size_t n_comp_>
    class VectorField
    {
      public:

        using mesh_t                        = mesh_t_;
        using cell_t                        = typename mesh_t::cell_t;
        using component_t                   = ScalarField<mesh_t, value_t>;
        static constexpr std::size_t dim    = mesh_t::dim;
        static constexpr std::size_t n_comp = n_comp_;

        VectorField(std::string name, mesh_t& mesh)
            : m_name(std::move(name))
            , p_mesh(&mesh)
        {
            for (std::size_t c = 0; c < n_comp; ++c)
            {
                m_comp[c] = std::make_unique<component_t>(m_name + "_" + std::to_string(c), mesh, &m_storage.s[c]);
            }
        }

        VectorField(const VectorField&)            = delete; // the components point into m_storage
        VectorField& operator=(const VectorField&) = delete;

        const std::string& name() const
        {
            return m_name;
        }

        mesh_t& mesh()
        {
            return *p_mesh;
        }

        const mesh_t& mesh() const
        {
            return *p_mesh;
        }

        component_t& component(std::size_t c)
        {
            return *m_comp[c];
        }

        const component_t& component(std::size_t c) const
        {
            return *m_comp[c];
        }

        void resize()
        {
            for (auto& f : m_comp)
            {
                f->resize();
            }
        }

        void fill(value_t v)
        {
            for (auto& f : m_comp)
            {
                f->fill(v);
            }
        }

        // what `u.array()` returns: std::swap(u.array(), unp1.array()) exchanges every component's device buffer and host mirror
        struct Storage
        {
            std::array<FieldStorage, n_comp> s;
        };

        Storage& array()
        {
            return m_storage;
        }

        // u[cell]: proxy over the n_comp values of one cell
        struct CellRef
        {
            VectorField* f;
            const cell_t* cell;

            value_t& operator[](std::size_t c)
            {
                return (*f->m_comp[c])[*cell];
            }

            value_t& operator()(std::size_t c)
            {
                return (*f->m_comp[c])[*cell];
            }

            CellRef& operator=(value_t v)
            {
                for (std::size_t c = 0; c < n_comp; ++c)
                {
                    (*f->m_comp[c])[*cell] = v;
                }
                return *this;
            }

            CellRef& operator=(const xt::xtensor_fixed<value_t, xt::xshape<n_comp>>& v)
            {
                for (std::size_t c = 0; c < n_comp; ++c)
                {
                    (*f->m_comp[c])[*cell] = v[c];
                }
                return *this;
            }

            CellRef& operator=(std::initializer_list<value_t> v)
            {
                std::size_t c = 0;
                for (const value_t& x : v)
                {
                    if (c < n_comp)
                    {
                        (*f->m_comp[c++])[*cell] = x;
                    }
                }
                return *this;
            }

            operator xt::xtensor_fixed<value_t, xt::xshape<n_comp>>() const
            {
                xt::xtensor_fixed<value_t, xt::xshape<n_comp>> r;
                for (std::size_t c = 0; c < n_comp; ++c)
                {
                    r[c] = (*f->m_comp[c])[*cell];
                }
                return r;
            }
        };

        CellRef operator[](const cell_t& cell)
        {
            return CellRef{this, &cell};
        }

        xt::xtensor_fixed<value_t, xt::xshape<n_comp>> operator[](const cell_t& cell) const
        {
            xt::xtensor_fixed<value_t, xt::xshape<n_comp>> r;
            for (std::size_t c = 0; c < n_comp; ++c)
            {
                r[c] = (*m_comp[c])[cell];
            }
            return r;
        }

        // u(c, level, i, j...): one component's row (field/access_base.hpp:117-160)
        template <class... Index>
        auto operator()(std::size_t c, std::size_t level, const Interval& i, const Index&... index)
        {
            return (*m_comp[c])(level, i, index...);
        }

        template <class... Index>
        auto operator()(std::size_t c, std::size_t level, const Interval& i, const Index&... index) const
        {
            return static_cast<const component_t&>(*m_comp[c])(level, i, index...);
        }

        void attach_bc(int type, const std::array<double, n_comp>& values)
        {
            for (std::size_t c = 0; c < n_comp; ++c)
            {
                m_comp[c]->attach_bc(type, values[c]);
            }
        }

        // `unp1 = u - dt * upwind(a, u)`: every component is transported by the same velocity (stencil_field.hpp:83-173)
        template <class A>
        VectorField& operator=(const fv_step_expr<A, VectorField>& e)
        {
            for (std::size_t c = 0; c < n_comp; ++c)
            {
                const component_t& uc = e.u->component(c);
                *m_comp[c] = fv_step_expr<A, component_t>{&uc, scaled_upwind_expr<A, component_t>{e.rhs.dt, upwind_expr<A, component_t>{e.rhs.op.a, &uc, e.rhs.op.burgers}}};
            }
            return *this;
        }

        // `rhs = scheme(u)` and `unp1 = v - dt * scheme(u)` for vector fields (make_convection_upwind<VectorField>() couples the components)
        VectorField& operator=(const scheme_expr<VectorField>& e)
        {
            e.scheme.apply(*this, *e.u);
            return *this;
        }

        VectorField& operator=(const scheme_step_expr<VectorField>& e)
        {
            VectorField tmp(e.rhs.e.scheme.name() + "(" + e.rhs.e.u->name() + ")", *p_mesh);
            e.rhs.e.scheme.apply(tmp, *e.rhs.e.u);
            for (std::size_t c = 0; c < n_comp; ++c)
            {
                b200::check(smr_field_resize(m_comp[c]->device()));
                b200::check(smr_field_lincomb(m_comp[c]->device(), 1.0, e.v->component(c).device(), -e.rhs.factor, tmp.component(c).device()));
                m_comp[c]->device_written();
            }
            return *this;
        }

      private:

        std::string m_name;
        mesh_t* p_mesh;
        Storage m_storage; // declared before the components: destroyed after them
        std::array<std::unique_ptr<component_t>, n_comp> m_comp;
    };

    template <class value_t, std::size_t n_comp, class mesh_t>
    auto make_vector_field(const std::string& name, mesh_t& mesh) // field/vector_field.hpp:234-300
    {
        return VectorField<mesh_t, value_t, n_comp>(name, mesh);
    }

    template <std::size_t n_comp, class mesh_t>
    auto make_vector_field(const std::string& name, mesh_t& mesh)
    {
        return VectorField<mesh_t, double, n_comp>(name, mesh);
    }

    template <class mesh_t, class T, std::size_t n>
    void swap(VectorField<mesh_t, T, n>& a, VectorField<mesh_t, T, n>& b)
    {
        std::swap(a.array(), b.array());
    }

    namespace b200
    {
        // every scalar component behind a field argument, in order
        template <class mesh_t, class T, class F>
        void for_each_component(ScalarField<mesh_t, T>& f, F&& fn)
        {
            fn(f);
        }

        template <class mesh_t, class T, std::size_t n, class F>
        void for_each_component(VectorField<mesh_t, T, n>& f, F&& fn)
        {
            for (std::size_t c = 0; c < n; ++c)
            {
                fn(f.component(c));
            }
        }
    }

    template <class value_t, class mesh_t>
    auto make_scalar_field(const std::string& name, mesh_t& mesh) // field/scalar_field.hpp:171-215
    {
        return ScalarField<mesh_t, value_t>(name, mesh);
    }

    template <class value_t, class mesh_t>
    auto make_scalar_field(const std::string& name, mesh_t& mesh, value_t init_value) // field/scalar_field.hpp:178-185
    {
        auto field = ScalarField<mesh_t, value_t>(name, mesh);
        field.fill(init_value);
        return field;
    }

    // initial value from a function of the cell centre (field/scalar_field.hpp:201-215)
    template <class value_t, class mesh_t, class Func>
        requires std::is_invocable_v<Func, typename Cell<mesh_t::dim>::coords_t>
    auto make_scalar_field(const std::string& name, mesh_t& mesh, Func&& f)
    {
        auto field = ScalarField<mesh_t, value_t>(name, mesh);
        field.fill(0);
        for_each_cell(mesh,
                      [&](const auto& cell)
                      {
                          field[cell] = f(cell.center());
                      });
        return field;
    }

    // `make_field<T, n>` only survives in the reference's README (README.md:104,132); kept as an alias of the scalar field
    template <class value_t, std::size_t n_comp, class mesh_t>
    auto make_field(const std::string& name, mesh_t& mesh)
    {
        static_assert(n_comp == 1, "the device path stores scalar fields (one SoA array per component)");
        return ScalarField<mesh_t, value_t>(name, mesh);
    }

    // field/swap.hpp:14-18
    template <class mesh_t, class T>
    void swap(ScalarField<mesh_t, T>& a, ScalarField<mesh_t, T>& b)
    {
        std::swap(a.array(), b.array());
    }

    // ---- bc/bc.hpp:751-815 -----------------------------------------------------------------------------------------------
    // make_bc returns a Bc*; `->on(directions...)` restricts it to boundary regions (bc/bc.hpp:326-363, 490-560).  The device
    // path applies one constant condition on the whole boundary: `on()` is accepted when the directions cover it.
    template <std::size_t dim>
    class BcHandle
    {
      public:

        BcHandle* operator->()
        {
            return this;
        }

        template <class... Directions>
        BcHandle& on(const Directions&... directions)
        {
            std::array<bool, 2 * dim> seen{};
            auto mark = [&](const auto& dir)
            {
                std::size_t axis = dim;
                int sign         = 0;
                for (std::size_t d = 0; d < dim; ++d)
                {
                    if (dir[d] != 0)
                    {
                        if (axis != dim)
                        {
                            throw std::invalid_argument("make_bc(...)->on(): only Cartesian directions are supported on the device path");
                        }
                        axis = d;
                        sign = dir[d] > 0 ? 1 : 0;
                    }
                }
                if (axis != dim)
                {
                    seen[2 * axis + static_cast<std::size_t>(sign)] = true;
                }
            };
            (mark(directions), ...);
            for (bool b : seen)
            {
                if (!b)
                {
                    throw std::invalid_argument("make_bc(...)->on(): the device path applies one condition on the whole boundary; list every direction");
                }
            }
            return *this;
        }
    };

    template <class BcType, class Field>
    auto make_bc(Field& u, double value)
    {
        if constexpr (requires { Field::n_comp; })
        {
            std::array<double, Field::n_comp> v;
            v.fill(value);
            u.attach_bc(BcType::type, v);
        }
        else
        {
            u.attach_bc(BcType::type, value);
        }
        return BcHandle<Field::dim>();
    }

    // make_bc<Dirichlet<1>>(u, v0, v1, ...): one constant per component of a vector field (bc/bc.hpp:751-815)
    template <class BcType, class Field, class... Values>
        requires(sizeof...(Values) >= 1 && requires { Field::n_comp; })
    auto make_bc(Field& u, double v0, Values... vs)
    {
        static_assert(sizeof...(Values) + 1 == Field::n_comp, "one boundary value per component");
        u.attach_bc(BcType::type, std::array<double, Field::n_comp>{v0, static_cast<double>(vs)...});
        return BcHandle<Field::dim>();
    }

    // ---- stencil_field.hpp:175-179, 245-249 ------------------------------------------------------------------------------
    template <class A, class Field>
    auto upwind(const A& a, const Field& u)
    {
        return upwind_expr<A, Field>{a, &u, false};
    }

    template <class A, class Field>
    auto upwind_scalar_burgers(const A& k, const Field& u)
    {
        return upwind_expr<A, Field>{k, &u, true};
    }

    template <class A, class Field>
    auto operator*(double dt, const upwind_expr<A, Field>& op)
    {
        return scaled_upwind_expr<A, Field>{dt, op};
    }

    template <class A, class mesh_t>
    auto operator-(const ScalarField<mesh_t, double>& u, const scaled_upwind_expr<A, ScalarField<mesh_t, double>>& rhs)
    {
        return fv_step_expr<A, ScalarField<mesh_t, double>>{&u, rhs};
    }

    template <class A, class mesh_t, std::size_t n>
    auto operator-(const VectorField<mesh_t, double, n>& u, const scaled_upwind_expr<A, VectorField<mesh_t, double, n>>& rhs)
    {
        return fv_step_expr<A, VectorField<mesh_t, double, n>>{&u, rhs};
    }

    // ---- schemes/fv/operators/{convection_lin,convection_nonlin,diffusion}.hpp ------------------------------------------------
    template <std::size_t dim>
    using VelocityVector = xt::xtensor_fixed<double, xt::xshape<dim>>; // operators/convection_lin.hpp:9
    template <std::size_t dim>
    using DiffCoeff = xt::xtensor_fixed<double, xt::xshape<dim>>; // operators/diffusion.hpp:111

    template <class Field>
    class FluxBasedScheme
    {
      public:

        using field_t = Field;

        FluxBasedScheme(int kind, const double* params, std::string name)
            : m_kind(kind)
            , m_name(std::move(name))
        {
            for (std::size_t d = 0; d < Field::dim; ++d)
            {
                m_params[d] = params[d];
            }
        }

        const std::string& name() const
        {
            return m_name;
        }

        void set_name(const std::string& n)
        {
            m_name = n;
        }

        // scheme(u): usable as `rhs = scheme(u)` and inside `v - dt * scheme(u)`
        scheme_expr<Field> operator()(Field& u) const
        {
            return {*this, &u};
        }

        // scheme.apply(out, in) (schemes/fv/FV_scheme.hpp:212-238): ghosts of `in` are updated if needed, out.fill(0), then the fluxes
        void apply(Field& out, Field& in) const
        {
            if constexpr (requires { Field::n_comp; }) // VectorField: SoA component fields on one mesh
            {
                constexpr std::size_t n = Field::n_comp;
                if (m_kind == SMR_SCHEME_CONVECTION_UPWIND_NONLINEAR || m_kind == SMR_SCHEME_CONVECTION_WENO5_NONLINEAR) // flux u(d) * u couples the components
                {
                    smr_field_t oh[n], ih[n];
                    for (std::size_t c = 0; c < n; ++c)
                    {
                        oh[c] = out.component(c).device();
                        ih[c] = in.component(c).device();
                    }
                    b200::check(smr_scheme_apply_vector(oh, ih, static_cast<int>(n), m_kind, m_params, m_scale));
                }
                else // the linear schemes act on every component separately
                {
                    for (std::size_t c = 0; c < n; ++c)
                    {
                        b200::check(smr_scheme_apply(out.component(c).device(), in.component(c).device(), m_kind, m_params, m_scale));
                    }
                }
                for (std::size_t c = 0; c < n; ++c)
                {
                    out.component(c).device_written();
                    in.component(c).device_written();
                }
            }
            else
            {
                b200::check(smr_scheme_apply(out.device(), in.device(), m_kind, m_params, m_scale));
                out.device_written();
                in.device_written(); // its ghosts may just have been updated on the device
            }
        }

        FluxBasedScheme scaled(double s) const // flux_based/algebraic_operators.hpp:7-82
        {
            FluxBasedScheme r(*this);
            r.m_scale *= s;
            r.m_name = std::to_string(s) + " * " + m_name;
            return r;
        }

      private:

        int m_kind;
        double m_params[3] = {0, 0, 0};
        double m_scale     = 1.0;
        std::string m_name;
    };

    template <class Field>
    auto make_convection_upwind(const VelocityVector<Field::dim>& velocity) // operators/convection_lin.hpp:15-89
    {
        double v[3] = {0, 0, 0};
        for (std::size_t d = 0; d < Field::dim; ++d)
        {
            v[d] = velocity(d);
        }
        return FluxBasedScheme<Field>(SMR_SCHEME_CONVECTION_UPWIND, v, "convection");
    }

    template <class Field>
    auto make_convection_weno5(const VelocityVector<Field::dim>& velocity) // operators/convection_lin.hpp:95-178 (WENO5, Jiang & Shu)
    {
        double v[3] = {0, 0, 0};
        for (std::size_t d = 0; d < Field::dim; ++d)
        {
            v[d] = velocity(d);
        }
        return FluxBasedScheme<Field>(SMR_SCHEME_CONVECTION_WENO5, v, "convection");
    }

    template <class Field>
    auto make_convection_weno5() // operators/convection_nonlin.hpp:162-233 (f = u * u or u(d) * u, WENO5)
    {
        const double v[3] = {0, 0, 0};
        return FluxBasedScheme<Field>(SMR_SCHEME_CONVECTION_WENO5_NONLINEAR, v, "convection(u)");
    }

    template <class Field>
    auto make_convection_upwind() // operators/convection_nonlin.hpp:24-76 (scalar field: Burgers)
    {
        const double v[3] = {0, 0, 0};
        return FluxBasedScheme<Field>(SMR_SCHEME_CONVECTION_UPWIND_NONLINEAR, v, "convection(u)");
    }

    template <class Field>
    auto make_diffusion_order2(const DiffCoeff<Field::dim>& K) // operators/diffusion.hpp:123-175
    {
        double k[3] = {0, 0, 0};
        for (std::size_t d = 0; d < Field::dim; ++d)
        {
            k[d] = K(d);
        }
        return FluxBasedScheme<Field>(SMR_SCHEME_DIFFUSION_ORDER2, k, "diffusion");
    }

    template <class Field>
    auto make_diffusion_order2(double k = 1.0) // operators/diffusion.hpp:236-252
    {
        DiffCoeff<Field::dim> K;
        K.fill(k);
        return make_diffusion_order2<Field>(K);
    }

    template <class Field>
    auto operator*(double s, const FluxBasedScheme<Field>& scheme)
    {
        return scheme.scaled(s);
    }

    template <class Field>
    auto operator*(double dt, const scheme_expr<Field>& e)
    {
        return scaled_scheme_expr<Field>{dt, e};
    }

    template <class mesh_t, std::size_t n>
    auto operator-(const VectorField<mesh_t, double, n>& v, const scaled_scheme_expr<VectorField<mesh_t, double, n>>& rhs)
    {
        return scheme_step_expr<VectorField<mesh_t, double, n>>{&v, rhs};
    }

    template <class mesh_t>
    auto operator-(const ScalarField<mesh_t, double>& v, const scaled_scheme_expr<ScalarField<mesh_t, double>>& rhs)
    {
        return scheme_step_expr<ScalarField<mesh_t, double>>{&v, rhs};
    }

    // ---- general field expressions (field/field_expression.hpp:62-141, field/field_base.hpp:230-242) -----------------------
    // The reference evaluates `unp1 = 1./3 * u + 2./3 * (u2 - dt * conv(u2))` lazily, cell by cell over the leaves.  Here the
    // tree is walked once per assignment and every `a * X + b * Y` node becomes one LinCombOp launch over the leaves
    // (smr_field_lincomb: a * x + b * y, no FMA contraction), so each cell sees the same operations in the same order as the
    // reference's expression template.  scheme(u) leaves are applied into temporaries first (explicit_FV_scheme.hpp), like
    // the reference's `make_field_operator_function` does.  Temporaries come from a small per-mesh pool.
    namespace fx
    {
        template <class Field>
        struct pool
        {
            static std::vector<std::unique_ptr<Field>>& fields(typename Field::mesh_t* mesh)
            {
                static std::map<typename Field::mesh_t*, std::vector<std::unique_ptr<Field>>> m;
                return m[mesh];
            }

            static int& used()
            {
                static int n = 0;
                return n;
            }

            static Field& take(typename Field::mesh_t& mesh)
            {
                auto& v = fields(&mesh);
                if (static_cast<int>(v.size()) <= used())
                {
                    v.push_back(std::make_unique<Field>("tmp" + std::to_string(v.size()), mesh));
                }
                Field& f = *v[static_cast<std::size_t>(used()++)];
                f.resize();
                return f;
            }
        };

        template <class Field>
        struct term // value = c * f
        {
            double c;
            Field* f;
        };

        template <class Field>
        struct leaf : node<leaf<Field>>
        {
            using field_t = Field;
            Field* f;

            term<Field> eval() const
            {
                return {1.0, f};
            }

            void assign_to(Field& out) const
            {
                if (&out != f)
                {
                    b200::check(smr_field_resize(out.device()));
                    b200::check(smr_field_lincomb(out.device(), 1.0, f->device(), 0.0, f->device()));
                    out.device_written();
                }
            }
        };

        template <class Field>
        struct scheme_node : node<scheme_node<Field>> // scheme(u)
        {
            using field_t = Field;
            FluxBasedScheme<Field> scheme;
            Field* u;

            term<Field> eval() const
            {
                Field& tmp = pool<Field>::take(u->mesh());
                scheme.apply(tmp, *u);
                return {1.0, &tmp};
            }

            void assign_to(Field& out) const
            {
                scheme.apply(out, *u);
            }
        };

        template <class E>
        struct scaled : node<scaled<E>>
        {
            using field_t = typename E::field_t;
            double c;
            E e;

            term<field_t> eval() const
            {
                term<field_t> t = e.eval();
                if (t.c != 1.0) // c * (t.c * f): materialise the inner product so the roundings stay where the expression has them
                {
                    field_t& tmp = pool<field_t>::take(t.f->mesh());
                    b200::check(smr_field_lincomb(tmp.device(), t.c, t.f->device(), 0.0, t.f->device()));
                    tmp.device_written();
                    t = {1.0, &tmp};
                }
                return {c, t.f};
            }

            void assign_to(field_t& out) const
            {
                const int mark = pool<field_t>::used();
                term<field_t> t = eval();
                b200::check(smr_field_resize(out.device()));
                b200::check(smr_field_lincomb(out.device(), t.c, t.f->device(), 0.0, t.f->device()));
                out.device_written();
                pool<field_t>::used() = mark;
            }
        };

        template <class A, class B>
        struct sum : node<sum<A, B>> // a + sign * b
        {
            using field_t = typename A::field_t;
            A a;
            B b;
            double sign;

            void into(field_t& out) const
            {
                term<field_t> ta = a.eval();
                term<field_t> tb = b.eval();
                b200::check(smr_field_resize(out.device()));
                b200::check(smr_field_lincomb(out.device(), ta.c, ta.f->device(), sign * tb.c, tb.f->device()));
                out.device_written();
            }

            term<field_t> eval() const
            {
                term<field_t> ta = a.eval();
                term<field_t> tb = b.eval();
                field_t& tmp     = pool<field_t>::take(ta.f->mesh());
                b200::check(smr_field_lincomb(tmp.device(), ta.c, ta.f->device(), sign * tb.c, tb.f->device()));
                tmp.device_written();
                return {1.0, &tmp};
            }

            void assign_to(field_t& out) const
            {
                const int mark = pool<field_t>::used();
                into(out);
                pool<field_t>::used() = mark;
            }
        };

        // operands: fields, scheme(u), dt * scheme(u), v - dt * scheme(u), and nodes
        template <class mesh_t>
        auto as_node(const ScalarField<mesh_t, double>& f)
        {
            return leaf<ScalarField<mesh_t, double>>{{}, const_cast<ScalarField<mesh_t, double>*>(&f)};
        }

        template <class Field>
        auto as_node(const scheme_expr<Field>& e)
        {
            return scheme_node<Field>{{}, e.scheme, e.u};
        }

        template <class Field>
        auto as_node(const scaled_scheme_expr<Field>& e)
        {
            return scaled<scheme_node<Field>>{{}, e.factor, as_node(e.e)};
        }

        template <class Field>
        auto as_node(const scheme_step_expr<Field>& e)
        {
            return sum<leaf<Field>, scaled<scheme_node<Field>>>{{}, as_node(*e.v), as_node(e.rhs), -1.0};
        }

        template <class E>
        const E& as_node(const node<E>& e)
        {
            return e.self();
        }

        template <class T, class = void>
        struct is_operand : std::false_type
        {
        };

        template <class mesh_t>
        struct is_operand<ScalarField<mesh_t, double>> : std::true_type
        {
        };

        template <class Field>
        struct is_operand<scheme_expr<Field>> : std::true_type
        {
        };

        template <class Field>
        struct is_operand<scaled_scheme_expr<Field>> : std::true_type
        {
        };

        template <class Field>
        struct is_operand<scheme_step_expr<Field>> : std::true_type
        {
        };

        template <class T>
        struct is_operand<T, std::enable_if_t<std::is_base_of_v<node<T>, T>>> : std::true_type
        {
        };

        template <class T>
        using node_of = std::decay_t<decltype(as_node(std::declval<const T&>()))>;
    } // namespace fx

    template <class A, class B, class = std::enable_if_t<fx::is_operand<A>::value && fx::is_operand<B>::value>>
    auto operator+(const A& a, const B& b)
    {
        return fx::sum<fx::node_of<A>, fx::node_of<B>>{{}, fx::as_node(a), fx::as_node(b), 1.0};
    }

    template <class A, class B, class = std::enable_if_t<fx::is_operand<A>::value && fx::is_operand<B>::value>, class = void>
    auto operator-(const A& a, const B& b)
    {
        return fx::sum<fx::node_of<A>, fx::node_of<B>>{{}, fx::as_node(a), fx::as_node(b), -1.0};
    }

    template <class S, class B, class = std::enable_if_t<std::is_arithmetic_v<S> && fx::is_operand<B>::value>, class = void, class = void>
    auto operator*(S c, const B& b)
    {
        return fx::scaled<fx::node_of<B>>{{}, static_cast<double>(c), fx::as_node(b)};
    }

    // ---- algorithm/update_ghost_mr.hpp:260-270 ---------------------------------------------------------------------------
    template <class Field, class... Fields>
    void update_ghost_mr(Field& u, Fields&... others)
    {
        b200::for_each_component(u,
                                 [](auto& f)
                                 {
                                     b200::check(smr_update_ghost_mr(f.device()));
                                     f.device_written();
                                 });
        if constexpr (sizeof...(Fields) > 0)
        {
            update_ghost_mr(others...);
        }
    }

    // ---- mr/config.hpp:10-68 -----------------------------------------------------------------------------------------------
    class mra_config
    {
      public:

        mra_config& epsilon(double e)
        {
            m_eps = e;
            return *this;
        }

        double epsilon() const
        {
            return m_eps;
        }

        mra_config& regularity(double r)
        {
            m_reg = r;
            return *this;
        }

        double regularity() const
        {
            return m_reg;
        }

        mra_config& relative_detail(bool b)
        {
            m_rel = b;
            return *this;
        }

        bool relative_detail() const
        {
            return m_rel;
        }

        void parse_args()
        {
            if (std::isfinite(args::epsilon))
            {
                m_eps = args::epsilon;
            }
            if (std::isfinite(args::regularity))
            {
                m_reg = args::regularity;
            }
            if (args::rel_detail) // mr/config.hpp:57-60
            {
                m_rel = true;
            }
        }

      private:

        double m_eps = 1e-4, m_reg = 1.;
        bool m_rel   = false;
    };

    // ---- mr/adapt.hpp:391-397 ----------------------------------------------------------------------------------------------
    template <class... Fields>
    class Adapt
    {
      public:

        explicit Adapt(Fields&... fields)
            : m_fields{&fields...}
        {
        }

        void operator()(mra_config& cfg)
        {
            cfg.parse_args();
            call(cfg.epsilon(), cfg.regularity(), cfg.relative_detail(), std::index_sequence_for<Fields...>{});
        }

        void operator()(double eps, double regularity)
        {
            call(eps, regularity, false, std::index_sequence_for<Fields...>{});
        }

      private:

        template <std::size_t... I>
        void call(double eps, double reg, bool rel, std::index_sequence<I...>)
        {
            // scalar fields and the components of vector fields, in argument order: one detail array each, one tag array
            std::vector<smr_field_t> h;
            (b200::for_each_component(*std::get<I>(m_fields), [&](auto& f) { h.push_back(f.device()); }), ...);
            int it = 0;
            b200::check(smr_adapt_ex(h.data(), static_cast<int>(h.size()), eps, reg, rel ? 1 : 0, &it));
            (b200::for_each_component(*std::get<I>(m_fields), [](auto& f) { f.device_written(); }), ...);
        }

        std::tuple<Fields*...> m_fields;
    };

    template <class... Fields>
    auto make_MRAdapt(Fields&... fields)
    {
        return Adapt<Fields...>(fields...);
    }

    // ---- io/hdf5.hpp:680-950, io/restart.hpp --------------------------------------------------------------------------
    // save() writes the reference's layout: <name>.h5 with /mesh/points [P,3] f64, /mesh/connectivity [N, 2^dim] u64 and one
    // /mesh/fields/<field> [N] dataset per scalar field / vector component (cells in for_each_cell order, points numbered
    // in order of first appearance like extract_coords_and_connectivity, io/hdf5.hpp:85-145), plus the <name>.xdmf companion.
    // No HDF5 library exists here: the file structures are written directly (b200_h5.hpp).
    namespace detail
    {
        template <class Mesh>
        void write_header(std::ostream& os, const Mesh&)
        {
            os << "level";
            for (std::size_t d = 0; d < Mesh::dim; ++d)
            {
                os << "," << "ijk"[d];
            }
        }

        inline const char* element_type(std::size_t dim) // io/hdf5.hpp:44-58
        {
            return dim == 1 ? "Polyline" : (dim == 2 ? "Quadrilateral" : "Hexahedron");
        }

        // per-cell values of one scalar field, in for_each_cell order
        template <class Mesh, class mesh_t, class T>
        void add_field_datasets(b200::h5::Writer& w, const Mesh& mesh, const ScalarField<mesh_t, T>& f, std::vector<std::string>& names, const std::string& prefix)
        {
            std::vector<uint64_t> shape{static_cast<uint64_t>(mesh.nb_cells())};
            if constexpr (std::is_floating_point_v<T>)
            {
                std::vector<double> v;
                v.reserve(shape[0]);
                for_each_cell(mesh, [&](const auto& cell) { v.push_back(static_cast<double>(f[cell])); });
                w.add(prefix + "/fields/" + f.name(), b200::h5::Type::f64, shape, v.data());
            }
            else if constexpr (std::is_signed_v<T>)
            {
                std::vector<int64_t> v;
                v.reserve(shape[0]);
                for_each_cell(mesh, [&](const auto& cell) { v.push_back(static_cast<int64_t>(f[cell])); });
                w.add(prefix + "/fields/" + f.name(), b200::h5::Type::i64, shape, v.data());
            }
            else
            {
                std::vector<uint64_t> v;
                v.reserve(shape[0]);
                for_each_cell(mesh, [&](const auto& cell) { v.push_back(static_cast<uint64_t>(f[cell])); });
                w.add(prefix + "/fields/" + f.name(), b200::h5::Type::u64, shape, v.data());
            }
            names.push_back(f.name());
        }

        template <class Mesh, class mesh_t, class T, std::size_t n>
        void add_field_datasets(b200::h5::Writer& w, const Mesh& mesh, const VectorField<mesh_t, T, n>& f, std::vector<std::string>& names, const std::string& prefix)
        {
            for (std::size_t c = 0; c < n; ++c) // io/hdf5.hpp:871-881: <name>_<component>
            {
                std::vector<double> v;
                v.reserve(mesh.nb_cells());
                for_each_cell(mesh, [&](const auto& cell) { v.push_back(static_cast<double>(f.component(c)[cell])); });
                const std::string name = f.name() + "_" + std::to_string(c);
                w.add(prefix + "/fields/" + name, b200::h5::Type::f64, {static_cast<uint64_t>(mesh.nb_cells())}, v.data());
                names.push_back(name);
            }
        }
    }

    template <class T>
    concept mesh_like = requires(const T& m) { m.c_config(); };

    template <class Mesh, class... Fields>
        requires mesh_like<Mesh>
    void save(const fs::path& path, const std::string& filename, const Mesh& mesh, const Fields&... fields)
    {
        constexpr std::size_t dim = Mesh::dim;
        fs::create_directories(path);
        const std::size_t n_cells = mesh.nb_cells();
        const std::size_t per_cell = std::size_t(1) << dim;
        // io/hdf5.hpp:60-82: corner order of a segment / quadrilateral / hexahedron
        static const int element[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
        std::map<std::array<double, dim>, uint64_t> points_id;
        std::vector<uint64_t> connectivity;
        connectivity.reserve(n_cells * per_cell);
        for_each_cell(mesh,
                      [&](const auto& cell)
                      {
                          const auto start = cell.corner();
                          for (std::size_t i = 0; i < per_cell; ++i)
                          {
                              std::array<double, dim> a;
                              for (std::size_t d = 0; d < dim; ++d)
                              {
                                  a[d] = start[d] + cell.length * element[dim == 1 ? (i == 0 ? 0 : 1) : i][d];
                              }
                              auto it = points_id.find(a);
                              if (it == points_id.end())
                              {
                                  it = points_id.emplace(a, static_cast<uint64_t>(points_id.size())).first;
                              }
                              connectivity.push_back(it->second);
                          }
                      });
        std::vector<double> coords(points_id.size() * 3, 0.0);
        for (const auto& kv : points_id)
        {
            for (std::size_t d = 0; d < dim; ++d)
            {
                coords[kv.second * 3 + d] = kv.first[d];
            }
        }
        b200::h5::Writer w;
        w.add("/mesh/connectivity", b200::h5::Type::u64, {static_cast<uint64_t>(n_cells), static_cast<uint64_t>(per_cell)}, connectivity.data());
        w.add("/mesh/points", b200::h5::Type::f64, {static_cast<uint64_t>(points_id.size()), 3}, coords.data());
        std::vector<std::string> names;
        (detail::add_field_datasets(w, mesh, fields, names, "/mesh"), ...);
        w.write((path / (filename + ".h5")).string());

        std::ofstream x(path / (filename + ".xdmf"));
        x << "<?xml version=\"1.0\"?>\n<Xdmf>\n    <Domain>\n        <Grid Name=\"mesh\">\n";
        x << "            <Topology TopologyType=\"" << detail::element_type(dim) << "\" NumberOfElements=\"" << n_cells << "\">\n";
        x << "                <DataItem Dimensions=\"" << n_cells * per_cell << "\" Format=\"HDF\">" << filename << ".h5:/mesh/connectivity</DataItem>\n";
        x << "            </Topology>\n            <Geometry GeometryType=\"XYZ\">\n";
        x << "                <DataItem Dimensions=\"" << points_id.size() * 3 << "\" Format=\"HDF\">" << filename << ".h5:/mesh/points</DataItem>\n";
        x << "            </Geometry>\n";
        for (const std::string& name : names)
        {
            x << "            <Attribute Name=\"" << name << "\" Center=\"Cell\">\n";
            x << "                <DataItem Dimensions=\"" << n_cells << "\" Format=\"HDF\" Precision=\"8\">" << filename << ".h5:/mesh/fields/" << name
              << "</DataItem>\n            </Attribute>\n";
        }
        x << "        </Grid>\n    </Domain>\n</Xdmf>\n";

        if (std::getenv("SAMURAI_B200_CSV") != nullptr) // plain-text copy of the same table (level, indices, one column per field)
        {
            std::ofstream os(path / (filename + ".csv"));
            os.precision(17);
            detail::write_header(os, mesh);
            for (const std::string& name : names)
            {
                os << "," << name;
            }
            os << "\n";
            for_each_cell(mesh,
                          [&](const auto& cell)
                          {
                              os << cell.level;
                              for (std::size_t d = 0; d < Mesh::dim; ++d)
                              {
                                  os << "," << cell.indices[d];
                              }
                              auto put = [&](const auto& f)
                              {
                                  if constexpr (requires { f.component(0); })
                                  {
                                      for (std::size_t c = 0; c < std::decay_t<decltype(f)>::n_comp; ++c)
                                      {
                                          os << "," << f.component(c)[cell];
                                      }
                                  }
                                  else
                                  {
                                      os << "," << f[cell];
                                  }
                              };
                              (put(fields), ...);
                              os << "\n";
                          });
        }
    }

    // Checkpoints.  dump() writes <name>.h5 holding the leaves and the leaf values: /n_process, /mesh/{dim, min_level, max_level,
    // origin_point, scaling_factor} as in the reference's restart files (io/restart.hpp:96-176), the leaf intervals as
    // /mesh/b200/intervals [n, 5] (level, y, z, start, end) and /fields/<name>/data with the leaf values in for_each_cell order.
    // NOT interchangeable with the reference's restart files: those store samurai's per-dimension interval arrays as an HDF5
    // compound type (io/restart.hpp:30-47), which this writer does not produce; load() says so when it is given one.
    template <class Mesh, class... Fields>
        requires mesh_like<Mesh>
    void dump(const fs::path& path, const std::string& filename, const Mesh& mesh, const Fields&... fields)
    {
        fs::create_directories(path);
        b200::h5::Writer w;
        const auto& c = mesh.c_config();
        w.add_scalar("/n_process", uint64_t(1));
        w.add_scalar("/mesh/dim", static_cast<uint64_t>(Mesh::dim));
        w.add_scalar("/mesh/min_level", static_cast<uint64_t>(mesh.min_level()));
        w.add_scalar("/mesh/max_level", static_cast<uint64_t>(mesh.max_level()));
        w.add("/mesh/origin_point", b200::h5::Type::f64, {static_cast<uint64_t>(Mesh::dim)}, c.origin);
        w.add_scalar("/mesh/scaling_factor", static_cast<double>(c.scaling_factor));
        std::vector<int64_t> ivl;
        for (std::size_t level = 0; level <= mesh.max_level(); ++level)
        {
            for (const auto& iv : mesh.intervals(MRMeshId::cells, level))
            {
                ivl.insert(ivl.end(), {static_cast<int64_t>(level), static_cast<int64_t>(iv.y), static_cast<int64_t>(iv.z), static_cast<int64_t>(iv.start),
                                       static_cast<int64_t>(iv.end)});
            }
        }
        w.add("/mesh/b200/intervals", b200::h5::Type::i64, {static_cast<uint64_t>(ivl.size() / 5), 5}, ivl.data());
        const int64_t cfgv[8] = {c.min_level, c.max_level, c.pred_radius, c.max_stencil_radius, c.graduation_width, c.n_cells0[0], c.n_cells0[1], c.n_cells0[2]};
        w.add("/mesh/b200/config", b200::h5::Type::i64, {8}, cfgv);
        auto put = [&](const auto& f)
        {
            std::vector<std::string> names;
            b200::h5::Writer tmp; // names only
            if constexpr (requires { f.component(0); })
            {
                constexpr std::size_t n = std::decay_t<decltype(f)>::n_comp;
                w.add_scalar("/fields/" + f.name() + "/n_comp", static_cast<uint64_t>(n));
                std::vector<double> v;
                for_each_cell(mesh,
                              [&](const auto& cell)
                              {
                                  for (std::size_t k = 0; k < n; ++k)
                                  {
                                      v.push_back(f.component(k)[cell]);
                                  }
                              });
                w.add("/fields/" + f.name() + "/data", b200::h5::Type::f64, {static_cast<uint64_t>(v.size())}, v.data());
            }
            else
            {
                w.add_scalar("/fields/" + f.name() + "/n_comp", uint64_t(1));
                std::vector<double> v;
                for_each_cell(mesh, [&](const auto& cell) { v.push_back(static_cast<double>(f[cell])); });
                w.add("/fields/" + f.name() + "/data", b200::h5::Type::f64, {static_cast<uint64_t>(v.size())}, v.data());
            }
        };
        (put(fields), ...);
        w.write((path / (filename + ".h5")).string());
    }

    template <class Mesh, class... Fields>
        requires mesh_like<Mesh>
    void dump(const std::string& filename, const Mesh& mesh, const Fields&... fields)
    {
        dump(fs::current_path(), filename, mesh, fields...);
    }

    // load(): rebuilds the mesh from the checkpoint's leaves and fills the fields (io/restart.hpp:383-470)
    template <class Mesh, class... Fields>
    void load(const fs::path& file, Mesh& mesh, Fields&... fields)
    {
        fs::path p = file;
        if (p.extension() != ".h5")
        {
            p += ".h5";
        }
        b200::h5::Reader r(p.string());
        if (!r.exists("/mesh/b200/intervals"))
        {
            throw std::runtime_error("'" + p.string() + "' is not a samurai_b200 checkpoint (the reference's restart files store compound-type interval arrays, "
                                                          "which this reader does not decode)");
        }
        if (r.template read<uint64_t>("/mesh/dim").at(0) != Mesh::dim)
        {
            throw std::runtime_error("checkpoint dimension differs from the mesh's");
        }
        std::vector<uint64_t> shape;
        const std::vector<int64_t> ivl  = r.template read<int64_t>("/mesh/b200/intervals", &shape);
        const std::vector<int64_t> cfgv = r.template read<int64_t>("/mesh/b200/config");
        const std::vector<double> org   = r.template read<double>("/mesh/origin_point");
        smr_mesh_config c{};
        c.dim                = static_cast<int32_t>(Mesh::dim);
        c.min_level          = static_cast<int32_t>(cfgv.at(0));
        c.max_level          = static_cast<int32_t>(cfgv.at(1));
        c.pred_radius        = static_cast<int32_t>(cfgv.at(2));
        c.max_stencil_radius = static_cast<int32_t>(cfgv.at(3));
        c.graduation_width   = static_cast<int32_t>(cfgv.at(4));
        for (std::size_t d = 0; d < 3; ++d)
        {
            c.n_cells0[d] = static_cast<int32_t>(cfgv.at(5 + d));
            c.origin[d]   = d < Mesh::dim ? org.at(d) : 0.0;
        }
        c.scaling_factor = r.template read<double>("/mesh/scaling_factor").at(0);
        mesh.rebuild_from_leaves(c, ivl);
        auto get = [&](auto& f)
        {
            const std::vector<double> v = r.template read<double>("/fields/" + f.name() + "/data");
            std::size_t k               = 0;
            f.resize();
            if constexpr (requires { f.component(0); })
            {
                constexpr std::size_t n = std::decay_t<decltype(f)>::n_comp;
                if (v.size() != mesh.nb_cells() * n)
                {
                    throw std::runtime_error("checkpoint field '" + f.name() + "' does not match the mesh");
                }
                for_each_cell(mesh,
                              [&](const auto& cell)
                              {
                                  for (std::size_t c = 0; c < n; ++c)
                                  {
                                      f[cell][c] = v[k++];
                                  }
                              });
            }
            else
            {
                if (v.size() != mesh.nb_cells())
                {
                    throw std::runtime_error("checkpoint field '" + f.name() + "' does not match the mesh");
                }
                for_each_cell(mesh, [&](const auto& cell) { f[cell] = v[k++]; });
            }
        };
        (get(fields), ...);
    }

    template <class Mesh, class... Fields>
    void load(const std::string& file, Mesh& mesh, Fields&... fields)
    {
        load(fs::path(file), mesh, fields...);
    }
} // namespace samurai
