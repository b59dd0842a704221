// Host-side multiresolution mesh: the five sub-meshes of samurai's MRMesh, built with eager interval algebra,
// plus the two host stages of the adaptation that north_star keeps on the CPU (new cells from tags, graduation).
//
// Behavioural spec (reference, include/samurai/):
//   mesh.hpp:326-341,426-439   constructors / finalize_mesh
//   mesh.hpp:1231-1262         construct_union
//   mr/mesh.hpp:222-455        MRMesh::update_sub_mesh_impl (cells_and_ghosts, reference, proj_cells)
//   mesh.hpp:894-911 + cell_array.hpp:484-493   renumbering / update_index (storage offsets)
//   algorithm/graduation.hpp:743-842  update_cell_array_from_tag
//   algorithm/graduation.hpp:245-330,573-726  make_graduation (max_stencil_radius == 1: no boundary rule)
// Scope: serial (one subdomain == domain), non-periodic box domains.
#pragma once
#ifdef _OPENMP
#include <omp.h>
#endif
#include <cstdio>
#include "intervals.hpp"

#include <array>
#include <cstring>
#include <string>

namespace smr
{
    enum MeshId
    {
        CELLS            = 0,
        CELLS_AND_GHOSTS = 1,
        PROJ_CELLS       = 2,
        UNION_CELLS      = 3,
        REFERENCE        = 4
    };

    constexpr uint8_t TAG_KEEP    = 1; // reference cell_flag.hpp:11-17
    constexpr uint8_t TAG_COARSEN = 2;
    constexpr uint8_t TAG_REFINE  = 4;

    struct MeshConfig
    {
        int dim                = 2;
        int min_level          = 0;
        int max_level          = 6;
        int pred_radius        = 1; // mesh_config<dim, prediction_stencil_radius>
        int max_stencil_radius = 1;
        int graduation_width   = 1;
        int n0[3]              = {1, 1, 1}; // domain size in level-0 cells
        double origin[3]       = {0, 0, 0};
        double scaling         = 1.0;
        bool periodic[3]       = {false, false, false}; // mesh_config::periodic(d)
        bool refine_boundary   = false;                 // args::refine_boundary: keep_boundary_refined (mr/adapt.hpp:245-274)

        bool any_periodic() const
        {
            return periodic[0] || periodic[1] || periodic[2];
        }

        bool all_periodic() const
        {
            for (int d = 0; d < dim; ++d)
            {
                if (!periodic[d])
                {
                    return false;
                }
            }
            return true;
        }

        // translations that map the domain onto its periodic images at `level` (mesh.hpp:45-81 get_periodic_directions):
        // every combination of {-N_d, 0, +N_d} over the periodic dimensions except the null one
        std::vector<std::array<int, 3>> periodic_directions(int level) const
        {
            std::vector<std::array<int, 3>> out;
            if (!any_periodic())
            {
                return out;
            }
            for (int cz = -1; cz <= 1; ++cz)
            {
                for (int cy = -1; cy <= 1; ++cy)
                {
                    for (int cx = -1; cx <= 1; ++cx)
                    {
                        const int c[3] = {cx, cy, cz};
                        bool ok = (cx | cy | cz) != 0;
                        for (int d = 0; d < 3; ++d)
                        {
                            if (c[d] != 0 && (d >= dim || !periodic[d]))
                            {
                                ok = false;
                            }
                        }
                        if (ok)
                        {
                            out.push_back({cx * (n0[0] << level), cy * (n0[1] << level), cz * (n0[2] << level)});
                        }
                    }
                }
            }
            return out;
        }

        int ghost_width() const
        {
            return std::max(max_stencil_radius, pred_radius); // mesh_config.hpp:395
        }

        double cell_length(int level) const
        {
            return scaling / static_cast<double>(1 << level); // cell.hpp:16-20
        }
    };

    using CellArray = std::vector<LevelSet>; // indexed by level

    struct Mesh
    {
        MeshConfig cfg;
        int nlev = 0;
        CellArray cells, cag, proj, uni, ref;
        std::vector<int64_t> level_start; // first storage offset of each level (size nlev+1)
        int64_t nref      = 0;
        int64_t nleaves   = 0;
        uint64_t generation = 0;

        const CellArray& sub(int id) const
        {
            switch (id)
            {
                case CELLS:
                    return cells;
                case CELLS_AND_GHOSTS:
                    return cag;
                case PROJ_CELLS:
                    return proj;
                case UNION_CELLS:
                    return uni;
                default:
                    return ref;
            }
        }

        void domain_box(int level, int grow, int lo[3], int hi[3]) const
        {
            for (int d = 0; d < 3; ++d)
            {
                lo[d] = -grow;
                hi[d] = (cfg.n0[d] << level) + grow;
            }
        }

        LevelSet in_domain(const LevelSet& s, int level, int grow = 0) const
        {
            int lo[3], hi[3];
            domain_box(level, grow, lo, hi);
            return clip_box(s, cfg.dim, lo, hi);
        }

        LevelSet outside_domain(const LevelSet& s, int level) const
        {
            int lo[3], hi[3];
            domain_box(level, 0, lo, hi);
            return minus_box(s, cfg.dim, lo, hi);
        }

        bool cell_in_domain(int level, int x, int y, int z) const
        {
            const int c[3] = {x, y, z};
            for (int d = 0; d < cfg.dim; ++d)
            {
                if (c[d] < 0 || c[d] >= (cfg.n0[d] << level))
                {
                    return false;
                }
            }
            return true;
        }

        int min_leaf_level() const
        {
            for (int l = 0; l < nlev; ++l)
            {
                if (!cells[l].empty())
                {
                    return l;
                }
            }
            return nlev;
        }

        int max_leaf_level() const
        {
            for (int l = nlev - 1; l >= 0; --l)
            {
                if (!cells[l].empty())
                {
                    return l;
                }
            }
            return -1;
        }

        static int levels_for(const MeshConfig& c)
        {
            return c.max_level + 3;
        }

        void init_uniform(const MeshConfig& c, int level)
        {
            cfg  = c;
            nlev = levels_for(c);
            cells.assign(nlev, LevelSet());
            int lo[3] = {0, 0, 0}, hi[3];
            for (int d = 0; d < 3; ++d)
            {
                hi[d] = c.n0[d] << level;
            }
            cells[level] = make_box(c.dim, lo, hi);
            build();
        }

        void init_from_cells(const MeshConfig& c, CellArray&& ca)
        {
            cfg  = c;
            nlev = levels_for(c);
            cells = std::move(ca);
            cells.resize(nlev);
            build();
        }

        // construct_union + update_sub_mesh_impl + renumbering
        void build()
        {
            const int dim = cfg.dim, L = cfg.max_level;
            const int msr = cfg.max_stencil_radius, pr = cfg.pred_radius;
            for (auto& c : cells)
            {
                c.off.clear();
            }
            uni.assign(nlev, LevelSet());
            cag.assign(nlev, LevelSet());
            proj.assign(nlev, LevelSet());
            const bool multi = cfg.max_level != cfg.min_level;
#ifdef SMR_PLAN_TIMING
            const double tb0 = omp_get_wtime();
            std::vector<double> tlev(nlev + 1, 0.0);
#endif
            std::vector<LevelSet> add1(nlev), add2(nlev); // prediction ghosts sent one / two levels down (mr/mesh.hpp:329-359)
            // Tasks: every level cut into row chunks (intervals.hpp, "intra-level parallelism") plus the union pyramid, a
            // serial cascade that runs next to them.  A chunk yields partial cells_and_ghosts / prediction-ghost sets.
            struct Task
            {
                int level;
                size_t r0, r1;
                LevelSet cag, add1, add2;
                LevelSet per0, per1, per2; // periodic images of the three sets above (mr/mesh.hpp:276-331, 366-380)
            };
            std::vector<Task> tasks;
            std::vector<std::pair<int, int>> level_tasks(nlev, {0, 0});
            const bool periodic = cfg.any_periodic();
            // union over the periodic directions of translate(s, direction(level)) clipped to the domain grown by `grow`
            auto images = [&](const LevelSet& s, int level, int grow)
            {
                LevelSet acc;
                if (s.empty() || level < 0)
                {
                    return acc;
                }
                for (const auto& dv : cfg.periodic_directions(level))
                {
                    LevelSet t = in_domain(translate(s, dv[0], dv[1], dv[2]), level, grow);
                    if (!t.empty())
                    {
                        acc = acc.empty() ? std::move(t) : set_union(acc, t);
                    }
                }
                return acc;
            };
            for (int l = 0; l < nlev; ++l)
            {
                level_tasks[l].first = static_cast<int>(tasks.size());
                if (!cells[l].empty())
                {
                    const std::vector<size_t> cut = chunk_rows(cells[l], 3000, 16);
                    for (size_t c = 0; c + 1 < cut.size(); ++c)
                    {
                        tasks.push_back(Task{l, cut[c], cut[c + 1], {}, {}, {}, {}, {}, {}});
                    }
                }
                level_tasks[l].second = static_cast<int>(tasks.size());
            }
            const int ntasks = static_cast<int>(tasks.size());
#pragma omp parallel for schedule(dynamic, 1)
            for (int t = ntasks; t >= 0; --t)
            {
#ifdef SMR_PLAN_TIMING
                const double tl0 = omp_get_wtime();
#endif
                if (t == ntasks)
                {
                    for (int l = L; l >= 1; --l)
                    {
                        uni[l - 1] = coarsen(set_union(cells[l], uni[l]), 1, dim);
                    }
#ifdef SMR_PLAN_TIMING
                    tlev[nlev] = omp_get_wtime() - tl0;
#endif
                    continue;
                }
                Task& tk    = tasks[t];
                const int l = tk.level;
                const bool whole = tk.r0 == 0 && tk.r1 == cells[l].rows();
                LevelSet part_storage;
                if (!whole)
                {
                    part_storage = slice_rows(cells[l], tk.r0, tk.r1);
                }
                const LevelSet& part = whole ? cells[l] : part_storage;
                tk.cag               = expand(part, msr, dim);
                if (periodic)
                {
                    // ghost cells of the periodic images: expand(translate(cells, d), msr) inside the domain grown by msr
                    tk.per0 = images(tk.cag, l, msr);
                }
                if (multi && l >= 1)
                {
                    LevelSet below = expand(coarsen(tk.cag, 1, dim), pr, dim);
                    if (periodic)
                    {
                        tk.per1 = images(below, l - 1, 0); // prediction ghosts of the images, inside the domain only
                    }
                    tk.add1 = in_domain(below, l - 1, pr);
                    if (l - 1 > 0)
                    {
                        tk.add2 = expand(coarsen(part, 2, dim), pr, dim);
                        if (periodic)
                        {
                            tk.per2 = images(tk.add2, l - 2, 0);
                        }
                    }
                }
#ifdef SMR_PLAN_TIMING
#pragma omp atomic
                tlev[l] += omp_get_wtime() - tl0;
#endif
            }
            std::vector<LevelSet> per0(nlev), per1(nlev), per2(nlev);
#pragma omp parallel for schedule(dynamic, 1)
            for (int t = 6 * nlev - 1; t >= 0; --t)
            {
                const int l = t / 6, what = t % 6;
                if (what >= 3 && !periodic)
                {
                    continue;
                }
                auto member = [what](Task& tk) -> LevelSet&
                {
                    switch (what)
                    {
                        case 0:
                            return tk.cag;
                        case 1:
                            return tk.add1;
                        case 2:
                            return tk.add2;
                        case 3:
                            return tk.per0;
                        case 4:
                            return tk.per1;
                        default:
                            return tk.per2;
                    }
                };
                LevelSet& dst = what == 0 ? cag[l] : (what == 1 ? add1[l] : (what == 2 ? add2[l] : (what == 3 ? per0[l] : (what == 4 ? per1[l] : per2[l]))));
                std::vector<const LevelSet*> parts;
                for (int k = level_tasks[l].first; k < level_tasks[l].second; ++k)
                {
                    parts.push_back(&member(tasks[static_cast<size_t>(k)]));
                }
                if (parts.size() == 1)
                {
                    dst = std::move(member(tasks[static_cast<size_t>(level_tasks[l].first)]));
                }
                else if (!parts.empty())
                {
                    dst = union_all(parts);
                }
            }
#ifdef SMR_PLAN_TIMING
            const double tb1 = omp_get_wtime();
#endif
            ref.assign(nlev, LevelSet());
#pragma omp parallel for schedule(dynamic, 1)
            for (int l = nlev - 1; l >= 0; --l)
            {
                LevelSet r = cag[l];
                if (l + 1 < nlev && !add1[l + 1].empty())
                {
                    r = set_union(r, add1[l + 1]);
                }
                if (l + 2 < nlev && !add2[l + 2].empty())
                {
                    r = set_union(r, add2[l + 2]);
                }
                if (periodic)
                {
                    if (!per0[l].empty())
                    {
                        r = set_union(r, per0[l]);
                    }
                    if (l + 1 < nlev && !per1[l + 1].empty())
                    {
                        r = set_union(r, per1[l + 1]);
                    }
                    if (l + 2 < nlev && !per2[l + 2].empty())
                    {
                        r = set_union(r, per2[l + 2]);
                    }
                }
                ref[l] = std::move(r);
            }
#ifdef SMR_PLAN_TIMING
            const double tb2 = omp_get_wtime();
#endif
            if (multi)
            {
                int l = 0;
                while (l < nlev && ref[l].empty())
                {
                    ++l;
                }
                auto max_ref_level = [&]()
                {
                    int m = -1;
                    for (int k = 0; k < nlev; ++k)
                    {
                        if (!ref[k].empty())
                        {
                            m = k;
                        }
                    }
                    return m;
                };
                while (l < nlev - 1 && l <= max_ref_level())
                {
                    if (!ref[l].empty())
                    {
                        proj[l] = par_inter(ref[l], uni[l]);
                        if (!proj[l].empty())
                        {
                            ref[l + 1] = par_union(ref[l + 1], par_rows(proj[l],
                                                                        [dim](const LevelSet& part)
                                                                        {
                                                                            return refine(part, 1, dim);
                                                                        }));
                        }
                    }
                    ++l;
                }
            }
#ifdef SMR_PLAN_TIMING
            const double tb3 = omp_get_wtime();
#endif
            // storage numbering: level ascending, then rows (z, y), then x
            level_start.assign(nlev + 1, 0);
            int64_t counter = 0;
            for (int l = 0; l < nlev; ++l)
            {
                level_start[l] = counter;
                LevelSet& r    = ref[l];
                r.off.resize(r.xs.size());
                for (size_t i = 0; i < r.xs.size(); ++i)
                {
                    r.off[i] = counter;
                    counter += r.xe[i] - r.xs[i];
                }
            }
            level_start[nlev] = counter;
            nref              = counter;
            nleaves           = 0;
#pragma omp parallel for schedule(dynamic, 1)
            for (int l = nlev - 1; l >= 0; --l)
            {
                locate(cells[l], ref[l]);
                locate(cag[l], ref[l]);
                locate(proj[l], ref[l]);
            }
            for (int l = 0; l < nlev; ++l)
            {
                nleaves += cells[l].n_cells();
            }
#ifdef SMR_PLAN_TIMING
            std::printf("  mesh build: cag/add/union %.2f ms (", (tb1 - tb0) * 1e3);
            for (int t = 0; t <= nlev; ++t)
            {
                if (tlev[t] > 1e-4)
                {
                    std::printf(" %d:%.2f", t, tlev[t] * 1e3);
                }
            }
            std::printf(" ), ref union %.2f ms, proj cascade %.2f ms, numbering+locate %.2f ms\n", (tb2 - tb1) * 1e3, (tb3 - tb2) * 1e3,
                        (omp_get_wtime() - tb3) * 1e3);
#endif
            ++generation;
        }
    };

    inline bool same_cells(const CellArray& a, const CellArray& b)
    {
        const size_t n = std::max(a.size(), b.size());
        static const LevelSet empty;
        for (size_t l = 0; l < n; ++l)
        {
            const LevelSet& x = l < a.size() ? a[l] : empty;
            const LevelSet& y = l < b.size() ? b[l] : empty;
            if (!x.same_cells(y))
            {
                return false;
            }
        }
        return true;
    }

    // update_cell_array_from_tag: tags (reference-sized, indexed by storage offset) -> new leaf sets
    // update_cell_array_from_tag: tags (reference-sized, indexed by storage offset) -> new leaf sets.
    // `*unchanged` (optional) is set when no leaf is refined or coarsened: the result then equals m.cells.
    inline CellArray cells_from_tags(const Mesh& m, const uint8_t* tag, bool* unchanged = nullptr)
    {
        const int dim = m.cfg.dim;
        CellArray out(m.nlev);
        // scan tasks: row chunks of every level; each collects the cells it removes from its level, the children it adds
        // one level up and the parents it adds one level down
        struct Task
        {
            int level;
            size_t r0, r1;
            SetBuilder rem, up, down;
        };
        std::vector<Task> tasks;
        std::vector<std::pair<int, int>> level_tasks(m.nlev, {0, 0});
        for (int l = 0; l < m.nlev; ++l)
        {
            level_tasks[l].first = static_cast<int>(tasks.size());
            if (!m.cells[l].empty())
            {
                const std::vector<size_t> cut = chunk_rows(m.cells[l], 2000, 16);
                for (size_t c = 0; c + 1 < cut.size(); ++c)
                {
                    tasks.push_back(Task{l, cut[c], cut[c + 1], {}, {}, {}});
                }
            }
            level_tasks[l].second = static_cast<int>(tasks.size());
        }
#pragma omp parallel for schedule(dynamic, 1)
        for (int t = static_cast<int>(tasks.size()) - 1; t >= 0; --t)
        {
            Task& tk          = tasks[static_cast<size_t>(t)];
            const int l       = tk.level;
            const LevelSet& c = m.cells[l];
            for (size_t r = tk.r0; r < tk.r1; ++r)
            {
                const int y = key_y(c.key[r]), z = key_z(c.key[r]);
                const bool yz_even = (dim < 2 || (y & 1) == 0) && (dim < 3 || (z & 1) == 0);
                for (int q = c.ptr[r]; q < c.ptr[r + 1]; ++q)
                {
                    const uint8_t* t = tag + c.off[q];
                    const int s = c.xs[q], e = c.xe[q];
                    // run-length scan of the interval: 0 keep, 1 refine, 2 coarsen
                    int run_start = s, run_kind = -1;
                    auto flush    = [&](int x_end)
                    {
                        if (run_kind == 1)
                        {
                            tk.rem.add(c.key[r], run_start, x_end);
                            for (int cz = 0; cz < (dim > 2 ? 2 : 1); ++cz)
                            {
                                for (int cy = 0; cy < (dim > 1 ? 2 : 1); ++cy)
                                {
                                    tk.up.add(mk_key(dim > 1 ? 2 * y + cy : 0, dim > 2 ? 2 * z + cz : 0), 2 * run_start, 2 * x_end);
                                }
                            }
                        }
                        else if (run_kind == 2)
                        {
                            tk.rem.add(c.key[r], run_start, x_end);
                            if (yz_even)
                            {
                                // parent added once, through the even child (graduation.hpp:806-809)
                                const int ps = (run_start + 1) >> 1; // first even x >= run_start, halved
                                const int pe = ((x_end - 1) >> 1) + 1;
                                tk.down.add(mk_key(y >> 1, z >> 1), ps, pe);
                            }
                        }
                    };
                    for (int x = s; x < e; ++x)
                    {
                        const uint8_t tv = t[x - s];
                        int kind         = 0;
                        if ((tv & TAG_REFINE) && l < m.cfg.max_level)
                        {
                            kind = 1;
                        }
                        else if ((tv & TAG_COARSEN) && !(tv & TAG_KEEP) && l > m.cfg.min_level)
                        {
                            kind = 2;
                        }
                        if (kind != run_kind)
                        {
                            flush(x);
                            run_start = x;
                            run_kind  = kind;
                        }
                    }
                    flush(e);
                }
            }
        }
        bool any = false;
        for (const Task& tk : tasks)
        {
            any = any || !tk.rem.empty();
        }
        if (unchanged != nullptr)
        {
            *unchanged = !any;
        }
        if (!any && unchanged != nullptr)
        {
            return out; // the caller asked for the flag: it copies m.cells itself if it still needs them
        }
        if (!any)
        {
            for (int l = 0; l < m.nlev; ++l)
            {
                out[l] = m.cells[l];
                out[l].off.clear();
            }
            return out;
        }
        // per level: the cells added (from the scans of level-1 and level+1) and removed
        std::vector<LevelSet> add(m.nlev), rem(m.nlev);
#pragma omp parallel for schedule(dynamic, 1)
        for (int k = 2 * m.nlev - 1; k >= 0; --k)
        {
            const int l = k >> 1;
            SetBuilder sb;
            if (k & 1)
            {
                for (int t = level_tasks[l].first; t < level_tasks[l].second; ++t)
                {
                    sb.v.insert(sb.v.end(), tasks[static_cast<size_t>(t)].rem.v.begin(), tasks[static_cast<size_t>(t)].rem.v.end());
                }
                rem[l] = sb.build();
            }
            else
            {
                if (l > 0)
                {
                    for (int t = level_tasks[l - 1].first; t < level_tasks[l - 1].second; ++t)
                    {
                        sb.v.insert(sb.v.end(), tasks[static_cast<size_t>(t)].up.v.begin(), tasks[static_cast<size_t>(t)].up.v.end());
                    }
                }
                if (l + 1 < m.nlev)
                {
                    for (int t = level_tasks[l + 1].first; t < level_tasks[l + 1].second; ++t)
                    {
                        sb.v.insert(sb.v.end(), tasks[static_cast<size_t>(t)].down.v.begin(), tasks[static_cast<size_t>(t)].down.v.end());
                    }
                }
                add[l] = sb.build();
            }
        }
        // new level = (old ∪ added) \ removed, by key-range chunks of the old level
        struct Piece
        {
            int level;
            int64_t klo, khi;
            bool first, last;
            LevelSet res;
        };
        std::vector<Piece> pieces;
        std::vector<std::pair<int, int>> level_pieces(m.nlev, {0, 0});
        for (int l = 0; l < m.nlev; ++l)
        {
            level_pieces[l].first = static_cast<int>(pieces.size());
            if (l >= m.cfg.min_level && l <= m.cfg.max_level)
            {
                const LevelSet& c = m.cells[l];
                if (c.empty())
                {
                    pieces.push_back(Piece{l, 0, 0, true, true, {}});
                }
                else
                {
                    const std::vector<size_t> cut = chunk_rows(c, 3000, 16);
                    for (size_t k = 0; k + 1 < cut.size(); ++k)
                    {
                        pieces.push_back(Piece{l, c.key[cut[k]], k + 2 < cut.size() ? c.key[cut[k + 1]] : 0, k == 0, k + 2 == cut.size(), {}});
                    }
                }
            }
            level_pieces[l].second = static_cast<int>(pieces.size());
        }
        auto key_slice = [](const LevelSet& s, const Piece& pc)
        {
            const size_t r0 = pc.first ? 0 : static_cast<size_t>(std::lower_bound(s.key.begin(), s.key.end(), pc.klo) - s.key.begin());
            const size_t r1 = pc.last ? s.rows() : static_cast<size_t>(std::lower_bound(s.key.begin(), s.key.end(), pc.khi) - s.key.begin());
            return slice_rows(s, r0, r1);
        };
#pragma omp parallel for schedule(dynamic, 1)
        for (int k = static_cast<int>(pieces.size()) - 1; k >= 0; --k)
        {
            Piece& pc   = pieces[static_cast<size_t>(k)];
            const int l = pc.level;
            LevelSet s  = key_slice(m.cells[l], pc);
            if (!add[l].empty())
            {
                s = set_union(s, key_slice(add[l], pc));
            }
            if (!rem[l].empty())
            {
                s = set_diff(s, key_slice(rem[l], pc));
            }
            pc.res = std::move(s);
        }
#pragma omp parallel for schedule(dynamic, 1)
        for (int l = m.cfg.max_level; l >= m.cfg.min_level; --l)
        {
            std::vector<const LevelSet*> parts;
            for (int k = level_pieces[l].first; k < level_pieces[l].second; ++k)
            {
                parts.push_back(&pieces[static_cast<size_t>(k)].res);
            }
            out[l] = union_all(parts);
        }
        return out;
    }

    // make_graduation fixed point (grad width w; max_stencil_radius == 1 so no contiguous-boundary rule)
    inline int make_graduation(const MeshConfig& cfg, CellArray& ca)
    {
        const int dim = cfg.dim, w = cfg.graduation_width;
        const int nlev = static_cast<int>(ca.size());
        int nit        = 0;
        while (true)
        {
            int lo = nlev, hi = -1;
            for (int l = 0; l < nlev; ++l)
            {
                if (!ca[l].empty())
                {
                    lo = std::min(lo, l);
                    hi = std::max(hi, l);
                }
            }
            if (hi < 0)
            {
                return nit;
            }
            std::vector<LevelSet> out(nlev);
            bool any = false;
            // every row chunk of every fine level contributes independently to the coarser levels it overlaps
            struct Task
            {
                int fine;
                size_t r0, r1;
                std::vector<LevelSet> contrib; // indexed by coarse level
            };
            std::vector<Task> tasks;
            for (int fine = hi; fine > lo + 1; --fine)
            {
                if (ca[fine].empty())
                {
                    continue;
                }
                const std::vector<size_t> cut = chunk_rows(ca[fine], 3000, 16);
                for (size_t c = 0; c + 1 < cut.size(); ++c)
                {
                    tasks.push_back(Task{fine, cut[c], cut[c + 1], {}});
                }
            }
#pragma omp parallel for schedule(dynamic, 1)
            for (int t = 0; t < static_cast<int>(tasks.size()); ++t)
            {
                Task& tk       = tasks[static_cast<size_t>(t)];
                const int fine = tk.fine;
                tk.contrib.assign(static_cast<size_t>(nlev), LevelSet());
                // expand(X, 2w).on(fine-2) == expand(X.on(fine-1), w).on(fine-2): 2w is even and coarse cells are aligned on
                // multiples of 4, so the halo test x in [4c-2w, 4c+3+2w] is exactly (x>>1) in [2c-w, 2c+1+w]; coarsening
                // first halves the rows the expansion has to merge
                LevelSet p = coarsen(expand(coarsen(slice_rows(ca[fine], tk.r0, tk.r1), 1, dim), w, dim), 1, dim);
                if (cfg.any_periodic() && !p.empty())
                {
                    // the periodic images of the fine cells graduate the other side of the domain too (graduation.hpp:309-320):
                    // translate(fine, N(fine)) expanded and coarsened twice == translate(p, N(fine - 2)), N(fine) being a multiple of 4
                    LevelSet all = p;
                    for (const auto& dv : cfg.periodic_directions(fine - 2))
                    {
                        all = set_union(all, translate(p, dv[0], dv[1], dv[2]));
                    }
                    p = std::move(all);
                }
                for (int cl = fine - 2;; --cl)
                {
                    if (!p.empty() && !ca[cl].empty())
                    {
                        tk.contrib[static_cast<size_t>(cl)] = set_inter(p, ca[cl]);
                    }
                    if (cl == lo || p.empty())
                    {
                        break;
                    }
                    p = coarsen(p, 1, dim);
                }
            }
#pragma omp parallel for schedule(dynamic, 1) reduction(|| : any)
            for (int cl = lo; cl <= hi; ++cl)
            {
                std::vector<const LevelSet*> parts;
                for (const Task& tk : tasks)
                {
                    if (!tk.contrib.empty() && !tk.contrib[static_cast<size_t>(cl)].empty())
                    {
                        parts.push_back(&tk.contrib[static_cast<size_t>(cl)]);
                    }
                }
                if (!parts.empty())
                {
                    out[cl] = union_all(parts);
                    any     = true;
                }
            }
            // max_stencil_radius 2: the cell two steps inward of a boundary leaf must not lie in a coarser leaf, so that the stencil of
            // the coarser level never reaches outside the domain (list_interval_to_refine_for_contiguous_boundary_cells,
            // graduation.hpp:372-455: n_contiguous_boundary_cells = max(2, 2 * (2 - 2)) = 2; its second part needs radius > 2)
            if (cfg.max_stencil_radius == 2)
            {
                for (int k = 0; k < dim; ++k)
                {
                    if (cfg.periodic[k])
                    {
                        continue;
                    }
                    for (int sgn = 1; sgn >= -1; sgn -= 2)
                    {
                        int dv[3] = {0, 0, 0};
                        dv[k]     = sgn;
                        for (int level = hi; level > lo; --level)
                        {
                            if (ca[level].empty() || ca[level - 1].empty())
                            {
                                continue;
                            }
                            int blo[3], bhi[3];
                            for (int a = 0; a < 3; ++a)
                            {
                                blo[a] = -dv[a]; // translate(domain, -direction)
                                bhi[a] = (cfg.n0[a] << level) - dv[a];
                            }
                            const LevelSet bdry = minus_box(ca[level], dim, blo, bhi);
                            if (bdry.empty())
                            {
                                continue;
                            }
                            LevelSet r = set_inter(coarsen(translate(bdry, -2 * dv[0], -2 * dv[1], -2 * dv[2]), 1, dim), ca[level - 1]);
                            if (!r.empty())
                            {
                                out[level - 1] = out[level - 1].empty() ? std::move(r) : set_union(out[level - 1], r);
                                any            = true;
                            }
                        }
                    }
                }
            }
            if (!any)
            {
                return nit;
            }
            ++nit;
            CellArray nca(nlev);
            bool changed = false;
#pragma omp parallel for schedule(dynamic, 1) reduction(|| : changed)
            for (int l = nlev - 1; l >= 0; --l)
            {
                const bool add = l > 0 && !out[l - 1].empty();
                const bool rem = !out[l].empty();
                if (!add && !rem)
                {
                    nca[l] = std::move(ca[l]);
                    continue;
                }
                LevelSet s = add ? set_union(ca[l], refine(out[l - 1], 1, dim)) : ca[l];
                if (rem)
                {
                    s = set_diff(s, out[l]);
                }
                if (!s.same_cells(ca[l]))
                {
                    changed = true;
                }
                nca[l] = std::move(s);
            }
            ca.swap(nca);
            if (!changed)
            {
                return nit;
            }
        }
    }
} // namespace smr
