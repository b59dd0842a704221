// Host traversal of the set-algebra subsets of one mesh into flat, coalesced work batches ("cached index batches"
// in BASELINE.json's north_star).  Runs once per mesh; its time is reported separately from the device time.
//
// Which subsets (reference, include/samurai/):
//   ghost update wavefront      algorithm/update_ghost_mr.hpp:194-237
//   outer ghosts / BC           algorithm/update_outer_ghost.hpp:20-432, bc/apply_field_bc.hpp:53-101,315-466
//   detail / tag / keep sets    mr/adapt.hpp:310-357
//   leaves for FV expressions   field/field_base.hpp:230-242 (for_each_interval over mesh[cells])
//   field transfer              algorithm/update_fields.hpp:27-54
#pragma once
#include "items.h"
#include "mesh.hpp"

#include <cstring>
#include <string>

namespace smr
{
    enum BatchKind
    {
        B_FV = 0,
        B_PROJ,
        B_PRED,
        B_DETAIL,
        B_TAG,
        B_COPY,
        B_BC
    };

    struct Batch
    {
        int kind        = -1;
        int n_items     = 0;
        int n_ctas      = 0;
        int64_t n_cells = 0;
        // byte offsets into the arena
        int64_t items = -1, prefix = -1, cta_first = -1, aux = -1;
        int level     = -1;

        bool empty() const
        {
            return n_items == 0;
        }
    };

    struct Arena
    {
        std::vector<uint8_t> bytes;

        template <class T>
        int64_t push(const T* p, size_t n)
        {
            size_t off = (bytes.size() + 15) & ~size_t(15);
            bytes.resize(off + n * sizeof(T));
            if (n)
            {
                std::memcpy(bytes.data() + off, p, n * sizeof(T));
            }
            return static_cast<int64_t>(off);
        }

        template <class T>
        int64_t push(const std::vector<T>& v)
        {
            return push(v.data(), v.size());
        }
    };

    template <class Item>
    inline Batch finish_batch(Arena& arena, int kind, const std::vector<Item>& items, int level = -1)
    {
        Batch b;
        b.kind    = kind;
        b.level   = level;
        b.n_items = static_cast<int>(items.size());
        if (items.empty())
        {
            return b;
        }
        std::vector<int64_t> prefix(items.size() + 1);
        int64_t acc = 0;
        for (size_t i = 0; i < items.size(); ++i)
        {
            prefix[i] = acc;
            acc += items[i].n;
        }
        prefix[items.size()] = acc;
        b.n_cells            = acc;
        b.n_ctas             = static_cast<int>((acc + SMR_CTA_CELLS - 1) / SMR_CTA_CELLS);
        std::vector<int32_t> first(b.n_ctas + 1);
        size_t it = 0;
        for (int c = 0; c < b.n_ctas; ++c)
        {
            const int64_t g = static_cast<int64_t>(c) * SMR_CTA_CELLS;
            while (prefix[it + 1] <= g)
            {
                ++it;
            }
            first[c] = static_cast<int32_t>(it);
        }
        first[b.n_ctas] = b.n_items - 1;
        b.items         = arena.push(items);
        b.prefix        = arena.push(prefix);
        b.cta_first     = arena.push(first);
        return b;
    }

    inline Batch finish_bc_batch(Arena& arena, const std::vector<smr_item_bc>& items, const std::vector<int64_t>& srcs, int level)
    {
        Batch b;
        b.kind    = B_BC;
        b.level   = level;
        b.n_items = static_cast<int>(items.size());
        if (items.empty())
        {
            return b;
        }
        b.n_cells = b.n_items;
        b.n_ctas  = (b.n_items + SMR_CTA_THREADS - 1) / SMR_CTA_THREADS;
        b.items   = arena.push(items);
        b.aux     = arena.push(srcs);
        return b;
    }

    [[noreturn]] inline void missing(const char* what, int level, int x, int y, int z)
    {
        throw std::out_of_range(std::string("interval not found in the reference mesh (") + what + ") at level " + std::to_string(level)
                                + ", i = " + std::to_string(x) + ", index = " + std::to_string(y) + " " + std::to_string(z));
    }

    inline int64_t need(const LevelSet& ref, const char* what, int level, int y, int z, int x, int x_last)
    {
        const int64_t o = ref.offset_of(mk_key(y, z), x, x_last);
        if (o < 0)
        {
            missing(what, level, x, y, z);
        }
        return o;
    }

    // ---------------------------------------------------------------------------------------------------------
    // per-interval item builders
    // ---------------------------------------------------------------------------------------------------------
    inline void fv_items(const Mesh& m, std::vector<smr_item_fv>& out)
    {
        const int dim = m.cfg.dim;
        for (int l = 0; l < m.nlev; ++l)
        {
            const LevelSet& c   = m.cells[l];
            const LevelSet& ref = m.ref[l];
            for (size_t r = 0; r < c.rows(); ++r)
            {
                const int y = key_y(c.key[r]), z = key_z(c.key[r]);
                for (int q = c.ptr[r]; q < c.ptr[r + 1]; ++q)
                {
                    const int s = c.xs[q], e = c.xe[q];
                    smr_item_fv it;
                    it.c  = need(ref, "fv x", l, y, z, s - 1, e) + 1;
                    it.ym = it.yp = it.zm = it.zp = it.c;
                    if (dim > 1)
                    {
                        it.ym = need(ref, "fv y-1", l, y - 1, z, s, e - 1);
                        it.yp = need(ref, "fv y+1", l, y + 1, z, s, e - 1);
                    }
                    if (dim > 2)
                    {
                        it.zm = need(ref, "fv z-1", l, y, z - 1, s, e - 1);
                        it.zp = need(ref, "fv z+1", l, y, z + 1, s, e - 1);
                    }
                    it.n     = e - s;
                    it.level = l;
                    it.x     = s;
                    it.y     = y;
                    it.z     = z;
                    it.pad   = 0;
                    out.push_back(it);
                }
            }
        }
    }

    // coarse set `cs` at level lc (dst offsets from dst_ref) <- children rows in src_ref (level lc+1)
    inline void proj_items(int dim, const LevelSet& cs, int lc, const LevelSet& dst_ref, const LevelSet& src_ref, std::vector<smr_item_proj>& out)
    {
        for (size_t r = 0; r < cs.rows(); ++r)
        {
            const int y = key_y(cs.key[r]), z = key_z(cs.key[r]);
            for (int q = cs.ptr[r]; q < cs.ptr[r + 1]; ++q)
            {
                const int s = cs.xs[q], e = cs.xe[q];
                smr_item_proj it;
                it.dst = need(dst_ref, "projection dst", lc, y, z, s, e - 1);
                for (int cz = 0; cz < 2; ++cz)
                {
                    for (int cy = 0; cy < 2; ++cy)
                    {
                        const bool used = (dim > 1 || cy == 0) && (dim > 2 || cz == 0);
                        it.src[cy + 2 * cz] = used ? need(src_ref, "projection src", lc + 1, dim > 1 ? 2 * y + cy : 0, dim > 2 ? 2 * z + cz : 0, 2 * s, 2 * e - 1)
                                                   : 0;
                    }
                }
                it.n   = e - s;
                it.pad = 0;
                out.push_back(it);
            }
        }
    }

    // fine interval [s, e) at level lf, row (y, z): predicted from src_ref (level lf-1)
    inline void pred_item(int dim, int radius, int lf, int y, int z, int s, int e, int64_t dst, const LevelSet& src_ref, std::vector<smr_item_pred>& out)
    {
        smr_item_pred it;
        it.dst = dst;
        it.n   = e - s;
        it.par = (s & 1) | ((dim > 1 ? (y & 1) : 0) << 1) | ((dim > 2 ? (z & 1) : 0) << 2);
        const int yc = y >> 1, zc = z >> 1, sc = s >> 1, ec = (e - 1) >> 1;
        const int ry_ = dim > 1 ? radius : 0, rz_ = dim > 2 ? radius : 0;
        for (int k = 0; k < 9; ++k)
        {
            it.src[k] = 0;
        }
        for (int rz = -rz_; rz <= rz_; ++rz)
        {
            for (int ry = -ry_; ry <= ry_; ++ry)
            {
                it.src[(ry + 1) + 3 * (rz + 1)] = need(src_ref, "prediction src", lf - 1, yc + ry, zc + rz, sc - radius, ec + radius) + radius;
            }
        }
        out.push_back(it);
    }

    inline void detail_items(const Mesh& m, int level, const LevelSet& cs, std::vector<smr_item_detail>& out)
    {
        const int dim = m.cfg.dim, radius = m.cfg.pred_radius;
        const LevelSet& rc = m.ref[level];
        const LevelSet& rf = m.ref[level + 1];
        const int ry_ = dim > 1 ? radius : 0, rz_ = dim > 2 ? radius : 0;
        for (size_t r = 0; r < cs.rows(); ++r)
        {
            const int y = key_y(cs.key[r]), z = key_z(cs.key[r]);
            for (int q = cs.ptr[r]; q < cs.ptr[r + 1]; ++q)
            {
                const int s = cs.xs[q], e = cs.xe[q];
                smr_item_detail it;
                std::memset(&it, 0, sizeof(it));
                for (int rz = -rz_; rz <= rz_; ++rz)
                {
                    for (int ry = -ry_; ry <= ry_; ++ry)
                    {
                        it.coarse[(ry + 1) + 3 * (rz + 1)] = need(rc, "detail coarse", level, y + ry, z + rz, s - radius, e - 1 + radius) + radius;
                    }
                }
                for (int cz = 0; cz < (dim > 2 ? 2 : 1); ++cz)
                {
                    for (int cy = 0; cy < (dim > 1 ? 2 : 1); ++cy)
                    {
                        it.fine[cy + 2 * cz] = need(rf, "detail fine", level + 1, dim > 1 ? 2 * y + cy : 0, dim > 2 ? 2 * z + cz : 0, 2 * s, 2 * e - 1);
                    }
                }
                it.n = e - s;
                out.push_back(it);
            }
        }
    }

    inline void tag_items(const Mesh& m, int fine_level, const LevelSet& cs, std::vector<smr_item_tag>& out)
    {
        const int dim = m.cfg.dim;
        const LevelSet& rc = m.ref[fine_level - 1];
        const LevelSet& rf = m.ref[fine_level];
        for (size_t r = 0; r < cs.rows(); ++r)
        {
            const int y = key_y(cs.key[r]), z = key_z(cs.key[r]);
            for (int q = cs.ptr[r]; q < cs.ptr[r + 1]; ++q)
            {
                const int s = cs.xs[q], e = cs.xe[q];
                smr_item_tag it;
                std::memset(&it, 0, sizeof(it));
                it.coarse = need(rc, "tag coarse", fine_level - 1, y, z, s, e - 1);
                for (int cz = 0; cz < (dim > 2 ? 2 : 1); ++cz)
                {
                    for (int cy = 0; cy < (dim > 1 ? 2 : 1); ++cy)
                    {
                        it.fine[cy + 2 * cz] = need(rf, "tag fine", fine_level, dim > 1 ? 2 * y + cy : 0, dim > 2 ? 2 * z + cz : 0, 2 * s, 2 * e - 1);
                    }
                }
                it.n     = e - s;
                it.level = fine_level;
                out.push_back(it);
            }
        }
    }

    // ---------------------------------------------------------------------------------------------------------
    // directions (reference stencil.hpp:299-357)
    // ---------------------------------------------------------------------------------------------------------
    struct Dir
    {
        int v[3];
    };

    inline std::vector<Dir> cartesian_directions(int dim)
    {
        std::vector<Dir> out;
        for (int d = 0; d < dim; ++d)
        {
            for (int sgn : {1, -1})
            {
                Dir x{{0, 0, 0}};
                x.v[d] = sgn;
                out.push_back(x);
            }
        }
        return out;
    }

    inline std::vector<Dir> diagonal_directions(int dim)
    {
        std::vector<Dir> out;
        const int zr = dim > 2 ? 1 : 0, yr = dim > 1 ? 1 : 0;
        for (int z = -zr; z <= zr; ++z)
        {
            for (int y = -yr; y <= yr; ++y)
            {
                for (int x = -1; x <= 1; ++x)
                {
                    if (std::abs(x) + std::abs(y) + std::abs(z) > 1)
                    {
                        out.push_back(Dir{{x, y, z}});
                    }
                }
            }
        }
        return out;
    }

    // ---------------------------------------------------------------------------------------------------------
    // the per-mesh plan
    // ---------------------------------------------------------------------------------------------------------
    struct GhostPhase
    {
        Batch bc1;  // corner extrapolation + project_bc + apply_field_bc         (level)
        Batch bc2;  // project_corner_below + predict_bc                           (level)
        Batch proj; // projection level -> level-1
    };

    struct MeshPlan
    {
        Arena arena;
        Batch fv;                      // all leaves
        std::vector<GhostPhase> down;  // indexed by level (top-down sweep uses L..0)
        std::vector<Batch> pred;       // indexed by level (bottom-up sweep 1..L)
        std::vector<Batch> detail;     // indexed by coarse level
        std::vector<Batch> tag;        // indexed by fine level
        double build_seconds = 0;
    };

    class BcBuilder
    {
      public:

        std::vector<smr_item_bc> items;
        std::vector<int64_t> srcs;

        void copy(int64_t dst, int64_t src)
        {
            items.push_back({dst, 0.0, SMR_BC_COPY, 1, static_cast<int64_t>(srcs.size())});
            srcs.push_back(src);
        }

        void value(int64_t dst, int64_t src, double coef)
        {
            items.push_back({dst, coef, SMR_BC_VALUE, 1, static_cast<int64_t>(srcs.size())});
            srcs.push_back(src);
        }

        void begin_avg(int64_t dst)
        {
            items.push_back({dst, 0.0, SMR_BC_AVG, 0, static_cast<int64_t>(srcs.size())});
        }

        void add_src(int64_t src)
        {
            srcs.push_back(src);
            items.back().n_src++;
        }
    };

    // the corner cells of the (box) domain for a diagonal direction, `.on(level)` (mesh.hpp:914-997)
    inline LevelSet corner_cells(const Mesh& m, int level, const Dir& d)
    {
        int lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
        for (int k = 0; k < m.cfg.dim; ++k)
        {
            const int n = m.cfg.n0[k] << level;
            if (d.v[k] > 0)
            {
                lo[k] = n - 1;
                hi[k] = n;
            }
            else if (d.v[k] < 0)
            {
                lo[k] = 0;
                hi[k] = 1;
            }
            else
            {
                lo[k] = 0;
                hi[k] = n;
            }
        }
        return make_box(m.cfg.dim, lo, hi);
    }

    // leaves at `level` whose neighbour in direction d lies outside the domain (boundary.hpp:6-33)
    inline LevelSet boundary_leaves(const Mesh& m, int level, const Dir& d)
    {
        int lo[3], hi[3];
        m.domain_box(level, 0, lo, hi);
        for (int k = 0; k < 3; ++k)
        {
            lo[k] -= d.v[k]; // translate(domain, -d)
            hi[k] -= d.v[k];
        }
        return minus_box(m.cells[level], m.cfg.dim, lo, hi);
    }

    template <class F>
    inline void for_each_cell(const LevelSet& s, F&& f)
    {
        for (size_t r = 0; r < s.rows(); ++r)
        {
            const int y = key_y(s.key[r]), z = key_z(s.key[r]);
            for (int q = s.ptr[r]; q < s.ptr[r + 1]; ++q)
            {
                for (int x = s.xs[q]; x < s.xe[q]; ++x)
                {
                    f(x, y, z);
                }
            }
        }
    }

    inline void build_ghost_phase(const Mesh& m, int level, Arena& arena, GhostPhase& ph)
    {
        const MeshConfig& cfg = m.cfg;
        const int dim = cfg.dim, L = cfg.max_level, lmin = cfg.min_level;
        BcBuilder g1, g2;
        const LevelSet& ref = m.ref[level];
        if (ref.empty() && (level == 0 || m.ref[level - 1].empty()))
        {
            return;
        }
        if (dim > 1 && level >= lmin && level <= L)
        {
            for (const Dir& d : diagonal_directions(dim))
            {
                const LevelSet corner = corner_cells(m, level, d);
                // update_outer_corners_by_polynomial_extrapolation, ghost width 1: u[c + d] = u[c]
                LevelSet cc = set_inter(m.cells[level], corner);
                for_each_cell(cc,
                              [&](int x, int y, int z)
                              {
                                  g1.copy(need(ref, "corner ghost", level, y + d.v[1], z + d.v[2], x + d.v[0], x + d.v[0]),
                                          need(ref, "corner cell", level, y, z, x, x));
                              });
                // project_corner_below
                if (level > 0)
                {
                    const LevelSet fine_outer = set_inter(translate(corner, d.v[0], d.v[1], d.v[2]), ref);
                    for (int dl = 1; dl <= 2; ++dl)
                    {
                        const int pl = level - dl;
                        LevelSet ghosts = set_inter(coarsen(fine_outer, dl, dim), m.ref[pl]);
                        const int add   = (1 << dl) - 1;
                        for_each_cell(ghosts,
                                      [&](int x, int y, int z)
                                      {
                                          const int cx = (x << dl) + (d.v[0] == -1 ? add : 0);
                                          const int cy = dim > 1 ? (y << dl) + (d.v[1] == -1 ? add : 0) : 0;
                                          const int cz = dim > 2 ? (z << dl) + (d.v[2] == -1 ? add : 0) : 0;
                                          const int64_t src = ref.offset_of(mk_key(cy, cz), cx, cx);
                                          if (src >= 0)
                                          {
                                              g2.copy(need(m.ref[pl], "corner below", pl, y, z, x, x), src);
                                          }
                                      });
                        if (pl == 0)
                        {
                            break;
                        }
                    }
                }
            }
        }
        for (const Dir& d : cartesian_directions(dim))
        {
            if (level < L)
            {
                // project_bc, layer 1
                LevelSet ghosts = set_inter(m.outside_domain(translate(m.uni[level], d.v[0], d.v[1], d.v[2]), level), ref);
                for_each_cell(ghosts,
                              [&](int x, int y, int z)
                              {
                                  g1.begin_avg(need(ref, "project_bc ghost", level, y, z, x, x));
                                  for (int dl = 1; dl <= 2; ++dl)
                                  {
                                      const LevelSet& rf = m.ref[level + dl];
                                      const int n        = 1 << dl;
                                      for (int cz = 0; cz < (dim > 2 ? n : 1); ++cz)
                                      {
                                          for (int cy = 0; cy < (dim > 1 ? n : 1); ++cy)
                                          {
                                              for (int cx = 0; cx < n; ++cx)
                                              {
                                                  const int64_t o = rf.offset_of(mk_key(dim > 1 ? (y << dl) + cy : 0, dim > 2 ? (z << dl) + cz : 0),
                                                                                 (x << dl) + cx,
                                                                                 (x << dl) + cx);
                                                  if (o >= 0)
                                                  {
                                                      g1.add_src(o);
                                                  }
                                              }
                                          }
                                      }
                                      if (g1.items.back().n_src > 0)
                                      {
                                          break;
                                      }
                                  }
                              });
            }
            LevelSet bl;
            if (level >= lmin)
            {
                bl = boundary_leaves(m, level, d);
                const double dx = cfg.cell_length(level);
                for_each_cell(bl,
                              [&](int x, int y, int z)
                              {
                                  g1.value(need(ref, "bc ghost", level, y + d.v[1], z + d.v[2], x + d.v[0], x + d.v[0]),
                                           need(ref, "bc cell", level, y, z, x, x),
                                           dx);
                              });
            }
            if (level >= lmin && level < L && !bl.empty())
            {
                // predict_bc(level + 1): children of the BC ghosts present in reference[level+1]
                LevelSet fine = set_inter(refine(translate(bl, d.v[0], d.v[1], d.v[2]), 1, dim), m.ref[level + 1]);
                locate(fine, m.ref[level + 1]);
                for (size_t r = 0; r < fine.rows(); ++r)
                {
                    const int y = key_y(fine.key[r]), z = key_z(fine.key[r]);
                    for (int q = fine.ptr[r]; q < fine.ptr[r + 1]; ++q)
                    {
                        for (int x = fine.xs[q]; x < fine.xe[q]; ++x)
                        {
                            g2.copy(fine.off[q] + (x - fine.xs[q]), need(ref, "predict_bc parent", level, y >> 1, z >> 1, x >> 1, x >> 1));
                        }
                    }
                }
            }
        }
        ph.bc1 = finish_bc_batch(arena, g1.items, g1.srcs, level);
        ph.bc2 = finish_bc_batch(arena, g2.items, g2.srcs, level);
        if (level > 0)
        {
            LevelSet ps = set_inter(coarsen(ref, 1, dim), m.proj[level - 1]);
            std::vector<smr_item_proj> items;
            proj_items(dim, ps, level - 1, m.ref[level - 1], ref, items);
            ph.proj = finish_batch(arena, B_PROJ, items, level);
        }
    }

    inline LevelSet prediction_set(const Mesh& m, int level)
    {
        const int dim = m.cfg.dim;
        if (level > m.cfg.max_level || m.ref[level].empty())
        {
            return LevelSet();
        }
        LevelSet pg = m.in_domain(set_diff(m.ref[level], set_union(m.cells[level], m.proj[level])), level);
        if (pg.empty())
        {
            return pg;
        }
        return set_inter(pg, refine(m.ref[level - 1], 1, dim));
    }

    inline LevelSet detail_set(const Mesh& m, int level)
    {
        const int dim = m.cfg.dim;
        LevelSet below = coarsen(m.cells[level + 1], 1, dim);
        if (level + 2 < m.nlev)
        {
            below = set_union(below, coarsen(m.cells[level + 2], 2, dim));
        }
        return set_inter(m.ref[level], below);
    }

    inline LevelSet tag_set(const Mesh& m, int fine_level)
    {
        return set_inter(m.ref[fine_level - 1], coarsen(m.cells[fine_level], 1, m.cfg.dim));
    }

    inline void build_plan(const Mesh& m, MeshPlan& plan)
    {
        const MeshConfig& cfg = m.cfg;
        const int dim = cfg.dim, L = cfg.max_level, lmin = cfg.min_level;
        plan.arena.bytes.clear();
        {
            std::vector<smr_item_fv> items;
            fv_items(m, items);
            plan.fv = finish_batch(plan.arena, B_FV, items);
        }
        plan.down.assign(m.nlev, GhostPhase());
        plan.pred.assign(m.nlev, Batch());
        plan.detail.assign(m.nlev, Batch());
        plan.tag.assign(m.nlev, Batch());
        for (int level = L; level >= 0; --level)
        {
            build_ghost_phase(m, level, plan.arena, plan.down[level]);
        }
        for (int level = 1; level <= L; ++level)
        {
            LevelSet ps = prediction_set(m, level);
            if (ps.empty())
            {
                continue;
            }
            locate(ps, m.ref[level]);
            std::vector<smr_item_pred> items;
            for (size_t r = 0; r < ps.rows(); ++r)
            {
                const int y = key_y(ps.key[r]), z = key_z(ps.key[r]);
                for (int q = ps.ptr[r]; q < ps.ptr[r + 1]; ++q)
                {
                    pred_item(dim, cfg.pred_radius, level, y, z, ps.xs[q], ps.xe[q], ps.off[q], m.ref[level - 1], items);
                }
            }
            plan.pred[level] = finish_batch(plan.arena, B_PRED, items, level);
        }
        if (lmin != L)
        {
            for (int level = std::max(lmin - 1, 0); level < L; ++level)
            {
                std::vector<smr_item_detail> items;
                detail_items(m, level, detail_set(m, level), items);
                plan.detail[level] = finish_batch(plan.arena, B_DETAIL, items, level);
            }
            for (int level = std::max(lmin, 1); level <= L; ++level)
            {
                std::vector<smr_item_tag> items;
                tag_items(m, level, tag_set(m, level), items);
                plan.tag[level] = finish_batch(plan.arena, B_TAG, items, level);
            }
        }
    }

    // old mesh -> new mesh field transfer (update_fields): copy, projection, prediction batches
    struct TransferPlan
    {
        Arena arena;
        Batch copy, proj, pred;
    };

    inline void build_transfer(const Mesh& old_m, const Mesh& new_m, TransferPlan& tp)
    {
        const MeshConfig& cfg = old_m.cfg;
        const int dim = cfg.dim;
        tp.arena.bytes.clear();
        std::vector<smr_item_copy> copies;
        std::vector<smr_item_proj> projs;
        std::vector<smr_item_pred> preds;
        for (int l = cfg.min_level; l <= cfg.max_level; ++l)
        {
            LevelSet s = set_inter(old_m.ref[l], new_m.cells[l]);
            if (s.empty())
            {
                continue;
            }
            LevelSet so = s;
            locate(s, new_m.ref[l]);
            locate(so, old_m.ref[l]);
            for (size_t q = 0; q < s.xs.size(); ++q)
            {
                copies.push_back({s.off[q], so.off[q], s.xe[q] - s.xs[q], 0});
            }
        }
        for (int l = cfg.min_level + 1; l <= cfg.max_level; ++l)
        {
            LevelSet sc = set_inter(coarsen(old_m.cells[l], 1, dim), new_m.cells[l - 1]);
            proj_items(dim, sc, l - 1, new_m.ref[l - 1], old_m.ref[l], projs);
            LevelSet sr = set_inter(coarsen(new_m.cells[l], 1, dim), old_m.cells[l - 1]);
            for (size_t r = 0; r < sr.rows(); ++r)
            {
                const int y = key_y(sr.key[r]), z = key_z(sr.key[r]);
                for (int q = sr.ptr[r]; q < sr.ptr[r + 1]; ++q)
                {
                    const int s = sr.xs[q], e = sr.xe[q];
                    for (int cz = 0; cz < (dim > 2 ? 2 : 1); ++cz)
                    {
                        for (int cy = 0; cy < (dim > 1 ? 2 : 1); ++cy)
                        {
                            const int fy = dim > 1 ? 2 * y + cy : 0, fz = dim > 2 ? 2 * z + cz : 0;
                            const int64_t dst = need(new_m.ref[l], "update_fields prediction dst", l, fy, fz, 2 * s, 2 * e - 1);
                            pred_item(dim, cfg.pred_radius, l, fy, fz, 2 * s, 2 * e, dst, old_m.ref[l - 1], preds);
                        }
                    }
                }
            }
        }
        tp.copy = finish_batch(tp.arena, B_COPY, copies);
        tp.proj = finish_batch(tp.arena, B_PROJ, projs);
        tp.pred = finish_batch(tp.arena, B_PRED, preds);
    }
} // namespace smr
