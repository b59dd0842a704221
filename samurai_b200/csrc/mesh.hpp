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
            std::vector<LevelSet> add1(nlev), add2(nlev); // prediction ghosts sent one / two levels down (mr/mesh.hpp:329-359)
            // levels are independent here; the union pyramid (a serial cascade) runs as one more task next to them
#pragma omp parallel for schedule(dynamic, 1)
            for (int t = nlev; t >= 0; --t)
            {
                if (t == nlev)
                {
                    for (int l = L; l >= 1; --l)
                    {
                        uni[l - 1] = coarsen(set_union(cells[l], uni[l]), 1, dim);
                    }
                    continue;
                }
                const int l = t;
                cag[l]      = expand(cells[l], msr, dim);
                if (multi && l >= 1 && !cells[l].empty())
                {
                    add1[l] = in_domain(expand(coarsen(cag[l], 1, dim), pr, dim), l - 1, pr);
                    if (l - 1 > 0)
                    {
                        add2[l] = expand(coarsen(cells[l], 2, dim), pr, dim);
                    }
                }
            }
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
                ref[l] = std::move(r);
            }
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
                        proj[l] = set_inter(ref[l], uni[l]);
                        if (!proj[l].empty())
                        {
                            ref[l + 1] = set_union(ref[l + 1], refine(proj[l], 1, dim));
                        }
                    }
                    ++l;
                }
            }
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
    inline CellArray cells_from_tags(const Mesh& m, const uint8_t* tag)
    {
        const int dim = m.cfg.dim;
        CellArray out(m.nlev);
        // add[l][k]: cells created at level l by the scan of level l+1 (k = 0, parents) or l-1 (k = 1, children)
        std::vector<std::array<SetBuilder, 2>> add(m.nlev + 1);
        std::vector<SetBuilder> rem(m.nlev);
#pragma omp parallel for schedule(dynamic, 1)
        for (int l = m.nlev - 1; l >= 0; --l)
        {
            const LevelSet& c = m.cells[l];
            for (size_t r = 0; r < c.rows(); ++r)
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
                            rem[l].add(c.key[r], run_start, x_end);
                            for (int cz = 0; cz < (dim > 2 ? 2 : 1); ++cz)
                            {
                                for (int cy = 0; cy < (dim > 1 ? 2 : 1); ++cy)
                                {
                                    add[l + 1][1].add(mk_key(dim > 1 ? 2 * y + cy : 0, dim > 2 ? 2 * z + cz : 0), 2 * run_start, 2 * x_end);
                                }
                            }
                        }
                        else if (run_kind == 2)
                        {
                            rem[l].add(c.key[r], run_start, x_end);
                            if (yz_even)
                            {
                                // parent added once, through the even child (graduation.hpp:806-809)
                                const int ps = (run_start + 1) >> 1; // first even x >= run_start, halved
                                const int pe = ((x_end - 1) >> 1) + 1;
                                add[l - 1][0].add(mk_key(y >> 1, z >> 1), ps, pe);
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
#pragma omp parallel for schedule(dynamic, 1)
        for (int l = m.cfg.max_level; l >= m.cfg.min_level; --l)
        {
            LevelSet s = m.cells[l];
            s.off.clear();
            for (int k = 0; k < 2; ++k)
            {
                if (!add[l][k].empty())
                {
                    s = set_union(s, add[l][k].build());
                }
            }
            if (!rem[l].empty())
            {
                s = set_diff(s, rem[l].build());
            }
            out[l] = std::move(s);
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
            // every fine level contributes independently to the coarser levels it overlaps
            std::vector<std::vector<LevelSet>> contrib(nlev, std::vector<LevelSet>(nlev));
#pragma omp parallel for schedule(dynamic, 1)
            for (int fine = hi; fine > lo + 1; --fine)
            {
                if (ca[fine].empty())
                {
                    continue;
                }
                // expand(X, 2w).on(fine-2) == expand(X.on(fine-1), w).on(fine-2): 2w is even and coarse cells are aligned on
                // multiples of 4, so the halo test x in [4c-2w, 4c+3+2w] is exactly (x>>1) in [2c-w, 2c+1+w]; coarsening
                // first halves the rows the expansion has to merge
                LevelSet p = coarsen(expand(coarsen(ca[fine], 1, dim), w, dim), 1, dim);
                for (int cl = fine - 2;; --cl)
                {
                    if (!p.empty())
                    {
                        contrib[fine][cl] = set_inter(p, ca[cl]);
                    }
                    if (cl == lo || p.empty())
                    {
                        break;
                    }
                    p = coarsen(p, 1, dim);
                }
            }
            for (int cl = lo; cl <= hi; ++cl)
            {
                for (int fine = hi; fine > cl + 1; --fine)
                {
                    if (!contrib[fine][cl].empty())
                    {
                        out[cl] = set_union(out[cl], contrib[fine][cl]);
                        any     = true;
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
            for (int l = 0; l < nlev; ++l)
            {
                LevelSet s = ca[l];
                if (l > 0 && !out[l - 1].empty())
                {
                    s = set_union(s, refine(out[l - 1], 1, dim));
                }
                if (!out[l].empty())
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
