// TEST / BENCH INFRASTRUCTURE -- never linked into the product (libsamurai_b200.so has no CPU floating-point path).
//
// Compiled, all-cores CPU execution of samurai's per-time-step hot path: the CPU arm of bench.py (`--impl reference`,
// `cpu_baseline`) and a second checker next to the numpy oracle (oracle/samurai_oracle.py).  The reference library
// cannot be built in this image (xtensor, HighFive/HDF5, CLI11, pugixml, fmt are absent: SURVEY.md section 8c), so this
// is a *port*, never "samurai": the floating-point operators below are restated from the reference files they cite,
// in the reference's operation order, and are checked bit for bit against the numpy oracle, which itself is pinned
// on the reference's golden HDF5 files (tests/test_cpu_path.py, tests/test_oracle_golden.py).
//
// The integer side (interval set algebra, graduation, sub-meshes, traversal of the subsets into seeds) is the product's
// own host code (samurai_b200/csrc/{intervals,mesh,batches}.hpp, C++17 + OpenMP): it is host-only in the product too.
// What the product runs as CUDA kernels -- record derivation (derive.cuh) and every fp operator (kernels.cuh) -- is
// written here as OpenMP loops over the same seeds, which is the structure of the reference's own loops
// (for_each_interval over a subset, one xtensor expression per interval).
//
// Build (see oracle/Makefile): parity   g++ -O3 -ffp-contract=off -fopenmp
//                               speed    g++ -O3 -march=native -ffp-contract=off -fopenmp
// (-ffp-contract=off in both: the reference's x86-64 build does not contract to FMA and tags hinge on |d| > eps.)
#include "../samurai_b200/csrc/batches.hpp"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <omp.h>
#include <string>

using namespace smr;

namespace
{
    struct Sim
    {
        MeshConfig cfg;
        Mesh mesh;
        MeshPlan plan;
        TransferPlan tr;
        PlanFilter flt; // identity: one process owns the whole mesh
        bool plan_ready = false;
        bool graduated  = false;
        std::vector<double> u, unp1, detail;
        std::vector<uint8_t> tag;
        int bc_type     = SMR_BCTYPE_DIRICHLET;
        double bc_value = 0.0;
        double eps = 2e-4, regularity = 1.0;
        std::string error;
        // seconds per stage, for the report
        double t_mesh = 0, t_batches = 0, t_fp = 0;
    };

    double now()
    {
        return omp_get_wtime();
    }

    int64_t need(const LevelSet& ref, int level, int y, int z, int x, int x_last)
    {
        const int64_t o = ref.offset_of(mk_key(y, z), x, x_last);
        if (o < 0)
        {
            missing("cpu path", level, x, y, z);
        }
        return o;
    }

    struct SeedView
    {
        const smr_seed* s = nullptr;
        const int64_t* prefix = nullptr;
        int n = 0;
    };

    SeedView seeds_of(const Arena& a, const Batch& b)
    {
        SeedView v;
        if (!b.empty())
        {
            v.s      = reinterpret_cast<const smr_seed*>(a.p + b.seeds);
            v.prefix = reinterpret_cast<const int64_t*>(a.p + b.prefix);
            v.n      = b.n_items;
        }
        return v;
    }

    // number of leading records whose output cells lie below `limit` (limits fall on record boundaries)
    int records_below(const SeedView& v, int64_t limit)
    {
        if (limit < 0)
        {
            return v.n;
        }
        int k = 0;
        while (k < v.n && v.prefix[k] < limit)
        {
            ++k;
        }
        return k;
    }

    template <class F>
    void par_seeds(const SeedView& v, int count, F&& f)
    {
#pragma omp parallel for schedule(dynamic, 64)
        for (int i = 0; i < count; ++i)
        {
            f(v.s[i]);
        }
    }

    // ---- floating-point operators, restated from the reference ---------------------------------------------------------

    // upwind_op::flux, stencil_field.hpp:30-54: .5 * a * (ul + ur) + .5 * |a| * (ul - ur)
    inline double upwind_flux(double ha, double haa, double ul, double ur)
    {
        return ha * (ul + ur) + haa * (ul - ur);
    }

    // unp1 = u - dt * upwind(a, u)   (stencil_field.hpp:83-173 right_flux - left_flux per direction, / dx;
    // field/field_base.hpp:230-242 assigns the expression interval by interval)
    void fv_upwind(Sim& S, const double* a, double dt)
    {
        const int dim = S.cfg.dim;
        double ha[3], haa[3];
        for (int d = 0; d < 3; ++d)
        {
            ha[d]  = .5 * (d < dim ? a[d] : 0.0);
            haa[d] = .5 * std::abs(d < dim ? a[d] : 0.0);
        }
        const double* u = S.u.data();
        double* out     = S.unp1.data();
        const SeedView v = seeds_of(S.plan.arena, S.plan.fv);
        par_seeds(v, v.n,
                  [&](const smr_seed& sd)
                  {
                      const int l = sd.level & 0xff, s = sd.xs, e = sd.xs + sd.n, y = sd.y, z = sd.z;
                      const LevelSet& ref = S.mesh.ref[l];
                      const int64_t c  = need(ref, l, y, z, s - 1, e) + 1;
                      const int64_t ym = dim > 1 ? need(ref, l, y - 1, z, s, e - 1) : c;
                      const int64_t yp = dim > 1 ? need(ref, l, y + 1, z, s, e - 1) : c;
                      const int64_t zm = dim > 2 ? need(ref, l, y, z - 1, s, e - 1) : c;
                      const int64_t zp = dim > 2 ? need(ref, l, y, z + 1, s, e - 1) : c;
                      const double dx  = S.cfg.cell_length(l);
                      for (int k = 0; k < sd.n; ++k)
                      {
                          const double uc = u[c + k];
                          double acc      = -upwind_flux(ha[0], haa[0], u[c + k - 1], uc) + upwind_flux(ha[0], haa[0], uc, u[c + k + 1]);
                          if (dim > 1)
                          {
                              acc = (acc + -upwind_flux(ha[1], haa[1], u[ym + k], uc)) + upwind_flux(ha[1], haa[1], uc, u[yp + k]);
                          }
                          if (dim > 2)
                          {
                              acc = (acc + -upwind_flux(ha[2], haa[2], u[zm + k], uc)) + upwind_flux(ha[2], haa[2], uc, u[zp + k]);
                          }
                          out[c + k] = uc - dt * (acc / dx);
                      }
                  });
    }

    // projection_op_, numeric/projection.hpp:22-64: mean of the 2^dim children, rows summed in (y, z) order
    void projection(const Mesh& dst_mesh, const Mesh& src_mesh, int dim, const SeedView& v, const double* src, double* dst)
    {
        par_seeds(v, v.n,
                  [&](const smr_seed& sd)
                  {
                      const int l = sd.level & 0xff, s = sd.xs, e = sd.xs + sd.n, y = sd.y, z = sd.z;
                      const int64_t d = need(dst_mesh.ref[l], l, y, z, s, e - 1);
                      int64_t r[4]    = {0, 0, 0, 0};
                      const int nr    = 1 << (dim - 1);
                      for (int cz = 0; cz < (dim > 2 ? 2 : 1); ++cz)
                      {
                          for (int cy = 0; cy < (dim > 1 ? 2 : 1); ++cy)
                          {
                              r[cy + 2 * cz] = need(src_mesh.ref[l + 1], l + 1, dim > 1 ? 2 * y + cy : 0, dim > 2 ? 2 * z + cz : 0, 2 * s, 2 * e - 1);
                          }
                      }
                      for (int k = 0; k < sd.n; ++k)
                      {
                          double sum = 0.0;
                          for (int q = 0; q < nr; ++q)
                          {
                              const double* c = src + r[q] + 2 * k;
                              sum += c[0] + c[1];
                          }
                          dst[d + k] = sum * (1.0 / static_cast<double>(1 << dim));
                      }
                  });
    }

    // interp_coeffs<3>(sign): {sign/8, 1, -sign/8}, sign = +1 for even cells, -1 for odd (numeric/prediction.hpp:31-35, 304-306)
    inline double interp1(int parity, int k)
    {
        const double s = parity ? -0.125 : 0.125;
        return k == 0 ? s : (k == 1 ? 1.0 : -s);
    }

    // prediction_op, numeric/prediction.hpp:259-361 (ghost form) and :107-257 (update_fields form): tensor product of the
    // 1D interpolation weights over the 3^dim coarse neighbours, accumulated z-outer, y, x-inner
    void prediction(const Mesh& dst_mesh, const Mesh& src_mesh, int dim, int radius, const SeedView& v, const double* src, double* dst)
    {
        par_seeds(v, v.n,
                  [&](const smr_seed& sd)
                  {
                      const int l = sd.level & 0xff, s = sd.xs, e = sd.xs + sd.n, y = sd.y, z = sd.z;
                      const int64_t d = need(dst_mesh.ref[l], l, y, z, s, e - 1);
                      const int ry_ = dim > 1 ? radius : 0, rz_ = dim > 2 ? radius : 0;
                      const int sc = s >> 1, ec = (e - 1) >> 1;
                      int64_t rows[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
                      for (int rz = -rz_; rz <= rz_; ++rz)
                      {
                          for (int ry = -ry_; ry <= ry_; ++ry)
                          {
                              rows[(ry + 1) + 3 * (rz + 1)] = need(src_mesh.ref[l - 1], l - 1, (y >> 1) + ry, (z >> 1) + rz, sc - radius, ec + radius) + radius;
                          }
                      }
                      const int py = dim > 1 ? (y & 1) : 0, pz = dim > 2 ? (z & 1) : 0;
                      for (int k = 0; k < sd.n; ++k)
                      {
                          const int ic = ((s & 1) + k) >> 1;
                          if (radius == 0)
                          {
                              dst[d + k] = src[rows[4] + ic];
                              continue;
                          }
                          const int px = (s + k) & 1;
                          double val   = 0.0;
                          for (int rz = (dim > 2 ? 0 : 1); rz < (dim > 2 ? 3 : 2); ++rz)
                          {
                              for (int ry = (dim > 1 ? 0 : 1); ry < (dim > 1 ? 3 : 2); ++ry)
                              {
                                  const double* row = src + rows[ry + 3 * rz] + ic - 1;
                                  for (int rx = 0; rx < 3; ++rx)
                                  {
                                      double coeff = 1.0 * interp1(px, rx);
                                      if (dim > 1)
                                      {
                                          coeff *= interp1(py, ry);
                                      }
                                      if (dim > 2)
                                      {
                                          coeff *= interp1(pz, rz);
                                      }
                                      val += row[rx] * coeff;
                                  }
                              }
                          }
                          dst[d + k] = val;
                      }
                  });
    }

    // compute_detail_op, mr/operators.hpp:146-175 (radius 0), :226-358 (2D), :360-533 (3D): detail(child) = f(child) - prediction,
    // the prediction terms subtracted one by one in the reference's loop order
    void compute_detail(Sim& S, int count)
    {
        const int dim = S.cfg.dim, radius = S.cfg.pred_radius;
        const double* f = S.u.data();
        double* det     = S.detail.data();
        const SeedView v = seeds_of(S.plan.arena, S.plan.detail);
        par_seeds(v, count,
                  [&](const smr_seed& sd)
                  {
                      const int l = sd.level & 0xff, s = sd.xs, e = sd.xs + sd.n, y = sd.y, z = sd.z;
                      const int ry_ = dim > 1 ? radius : 0, rz_ = dim > 2 ? radius : 0;
                      const int nr = 1 << (dim - 1);
                      int64_t coarse[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, fine[4] = {0, 0, 0, 0};
                      for (int rz = -rz_; rz <= rz_; ++rz)
                      {
                          for (int ry = -ry_; ry <= ry_; ++ry)
                          {
                              coarse[(ry + 1) + 3 * (rz + 1)] = need(S.mesh.ref[l], l, y + ry, z + rz, s - radius, e - 1 + radius) + radius;
                          }
                      }
                      for (int cz = 0; cz < (dim > 2 ? 2 : 1); ++cz)
                      {
                          for (int cy = 0; cy < (dim > 1 ? 2 : 1); ++cy)
                          {
                              fine[cy + 2 * cz] = need(S.mesh.ref[l + 1], l + 1, dim > 1 ? 2 * y + cy : 0, dim > 2 ? 2 * z + cz : 0, 2 * s, 2 * e - 1);
                          }
                      }
                      for (int k = 0; k < sd.n; ++k)
                      {
                          double d[4][2];
                          for (int r = 0; r < nr; ++r)
                          {
                              d[r][0] = f[fine[r] + 2 * k];
                              d[r][1] = f[fine[r] + 2 * k + 1];
                          }
                          if (radius == 0)
                          {
                              const double c = f[coarse[4] + k];
                              for (int r = 0; r < nr; ++r)
                              {
                                  d[r][0] -= c;
                                  d[r][1] -= c;
                              }
                          }
                          else
                          {
                              for (int rz = (dim > 2 ? 0 : 1); rz < (dim > 2 ? 3 : 2); ++rz)
                              {
                                  for (int ry = (dim > 1 ? 0 : 1); ry < (dim > 1 ? 3 : 2); ++ry)
                                  {
                                      const double* row = f + coarse[ry + 3 * rz] + k - 1;
                                      for (int rx = 0; rx < 3; ++rx)
                                      {
                                          const double c = row[rx];
                                          for (int r = 0; r < nr; ++r)
                                          {
                                              const int py = r & 1, pz = r >> 1;
                                              for (int px = 0; px < 2; ++px)
                                              {
                                                  double coeff = interp1(px, rx);
                                                  if (dim > 1)
                                                  {
                                                      coeff *= interp1(py, ry);
                                                  }
                                                  if (dim > 2)
                                                  {
                                                      coeff *= interp1(pz, rz);
                                                  }
                                                  d[r][px] -= coeff * c;
                                              }
                                          }
                                      }
                                  }
                              }
                          }
                          for (int r = 0; r < nr; ++r)
                          {
                              det[fine[r] + 2 * k]     = d[r][0];
                              det[fine[r] + 2 * k + 1] = d[r][1];
                          }
                      }
                  });
    }

    struct TagEps
    {
        double eps[SMR_MAX_LEVELS], fine_eps[SMR_MAX_LEVELS], coarse_eps[SMR_MAX_LEVELS];
    };

    void tag_rows(const Sim& S, const smr_seed& sd, int64_t& coarse, int64_t fine[4])
    {
        const int dim = S.cfg.dim;
        const int l = sd.level & 0xff, s = sd.xs, e = sd.xs + sd.n, y = sd.y, z = sd.z;
        coarse = need(S.mesh.ref[l], l, y, z, s, e - 1);
        for (int cz = 0; cz < (dim > 2 ? 2 : 1); ++cz)
        {
            for (int cy = 0; cy < (dim > 1 ? 2 : 1); ++cy)
            {
                fine[cy + 2 * cz] = need(S.mesh.ref[l + 1], l + 1, dim > 1 ? 2 * y + cy : 0, dim > 2 ? 2 * z + cz : 0, 2 * s, 2 * e - 1);
            }
        }
    }

    // mr_criteria_op, mr/criteria.hpp:19-113: coarsen a sibling group when the parent's detail is below eps_l * 2^regularity
    // and every child's is below eps_l; refine a child whose detail exceeds 2^(regularity + dim) * eps_l
    void criteria(Sim& S, const TagEps& te, int count)
    {
        const int dim = S.cfg.dim;
        const int nr  = 1 << (dim - 1);
        const double* d = S.detail.data();
        uint8_t* tag    = S.tag.data();
        const SeedView v = seeds_of(S.plan.arena, S.plan.tag_all);
        par_seeds(v, count,
                  [&](const smr_seed& sd)
                  {
                      const int fl = (sd.level & 0xff) + 1;
                      int64_t coarse, fine[4] = {0, 0, 0, 0};
                      tag_rows(S, sd, coarse, fine);
                      for (int k = 0; k < sd.n; ++k)
                      {
                          bool coarsen_ok = fl > S.cfg.min_level;
                          bool refine[4][2] = {};
                          if (std::abs(d[coarse + k]) > te.coarse_eps[fl])
                          {
                              coarsen_ok = false;
                          }
                          for (int r = 0; r < nr; ++r)
                          {
                              for (int x = 0; x < 2; ++x)
                              {
                                  const double a = std::abs(d[fine[r] + 2 * k + x]);
                                  if (a > te.eps[fl])
                                  {
                                      coarsen_ok = false;
                                  }
                                  if (a > te.fine_eps[fl])
                                  {
                                      refine[r][x] = true;
                                  }
                              }
                          }
                          for (int r = 0; r < nr; ++r)
                          {
                              for (int x = 0; x < 2; ++x)
                              {
                                  uint8_t* t = tag + fine[r] + 2 * k + x;
                                  if (coarsen_ok)
                                  {
                                      *t = TAG_COARSEN;
                                  }
                                  if (fl < S.cfg.max_level && refine[r][x])
                                  {
                                      *t = static_cast<uint8_t>(*t | TAG_REFINE);
                                  }
                              }
                          }
                      }
                  });
    }

    // maximum_op (keep propagation), mr/operators.hpp:29-89, one fine level at a time from max_level down
    void maximum(Sim& S, int fine_level)
    {
        const int dim = S.cfg.dim;
        const int nr  = 1 << (dim - 1);
        uint8_t* tag  = S.tag.data();
        const SeedView v = seeds_of(S.plan.arena, S.plan.tag[fine_level]);
        par_seeds(v, v.n,
                  [&](const smr_seed& sd)
                  {
                      int64_t coarse, fine[4] = {0, 0, 0, 0};
                      tag_rows(S, sd, coarse, fine);
                      for (int k = 0; k < sd.n; ++k)
                      {
                          uint8_t any = 0, all = 0xff;
                          for (int r = 0; r < nr; ++r)
                          {
                              for (int x = 0; x < 2; ++x)
                              {
                                  const uint8_t t = tag[fine[r] + 2 * k + x];
                                  any |= t;
                                  all &= t;
                              }
                          }
                          if (any & TAG_KEEP)
                          {
                              for (int r = 0; r < nr; ++r)
                              {
                                  tag[fine[r] + 2 * k] |= TAG_KEEP;
                                  tag[fine[r] + 2 * k + 1] |= TAG_KEEP;
                              }
                              tag[coarse + k] |= TAG_KEEP;
                          }
                          else if (all & TAG_COARSEN)
                          {
                              tag[coarse + k] |= TAG_KEEP;
                          }
                          else
                          {
                              for (int r = 0; r < nr; ++r)
                              {
                                  tag[fine[r] + 2 * k] &= static_cast<uint8_t>(~TAG_COARSEN);
                                  tag[fine[r] + 2 * k + 1] &= static_cast<uint8_t>(~TAG_COARSEN);
                              }
                          }
                      }
                  });
    }

    // boundary ghosts of one level (update_outer_ghost.hpp:36-131,210-335; bc/dirichlet.hpp:29: 2 v - u; bc/neumann.hpp:27-28:
    // dx v + u); the records of a phase never read what the same phase writes (batches.hpp: build_ghost_phase)
    void boundary(Sim& S, const Batch& b, double* f)
    {
        if (b.empty())
        {
            return;
        }
        const smr_item_bc* items = reinterpret_cast<const smr_item_bc*>(S.plan.arena.p + b.items);
        const int64_t* srcs      = reinterpret_cast<const int64_t*>(S.plan.arena.p + b.aux);
#pragma omp parallel for schedule(static) if (b.n_items > 2048)
        for (int i = 0; i < b.n_items; ++i)
        {
            const smr_item_bc& it = items[i];
            const int64_t* s      = srcs + it.src_first;
            const int kind        = it.kind & 0xff;
            if (kind == SMR_BC_COPY)
            {
                f[it.dst] = f[s[0]];
            }
            else if (kind == SMR_BC_VALUE)
            {
                f[it.dst] = S.bc_type == SMR_BCTYPE_DIRICHLET ? 2 * S.bc_value - f[s[0]] : it.coef * S.bc_value + f[s[0]];
            }
            else if (kind == SMR_BC_EXTRAP4) // bc/polynomial_extrapolation.hpp:67-70
            {
                f[it.dst] = f[s[0]] - f[s[1]] * 3.0 + f[s[2]] * 3.0;
            }
            else
            {
                double sum = 0.0;
                for (int j = 0; j < it.n_src; ++j)
                {
                    sum += f[s[j]];
                }
                if (it.n_src > 0)
                {
                    sum /= it.n_src;
                }
                f[it.dst] = sum;
            }
        }
    }

    void ensure_plan(Sim& S)
    {
        if (!S.plan_ready)
        {
            const double t0 = now();
            build_plan(S.mesh, S.plan, S.flt);
            S.t_batches += now() - t0;
            S.plan_ready = true;
        }
    }

    // update_ghost_mr, algorithm/update_ghost_mr.hpp:194-237: top-down outer ghosts + projection, bottom-up prediction
    void update_ghost(Sim& S)
    {
        ensure_plan(S);
        const double t0 = now();
        const int dim = S.cfg.dim;
        double* f     = S.u.data();
        for (int level = S.cfg.max_level; level >= 0; --level)
        {
            const GhostPhase& ph = S.plan.down[level];
            boundary(S, ph.bc, f);
            projection(S.mesh, S.mesh, dim, seeds_of(S.plan.arena, ph.proj), f, f);
            boundary(S, ph.bc2, f); // second ghost layer (ghost width 2): reads the first layer
        }
        for (int level = 1; level <= S.cfg.max_level; ++level)
        {
            prediction(S.mesh, S.mesh, dim, S.cfg.pred_radius, seeds_of(S.plan.arena, S.plan.pred[level]), f, f);
        }
        S.t_fp += now() - t0;
    }

    // one iteration of harten (mr/adapt.hpp:277-389); true when the mesh did not change
    bool harten(Sim& S, int ite)
    {
        const MeshConfig& cfg = S.cfg;
        const int dim = cfg.dim, L = cfg.max_level, lmin = cfg.min_level;
        ensure_plan(S);
        const int64_t n = S.mesh.nref;
        double t0       = now();
        S.detail.assign(static_cast<size_t>(n), 0.0); // mr/adapt.hpp:165-168
        S.tag.assign(static_cast<size_t>(n), 0);
        {
            uint8_t* tag = S.tag.data();
            const SeedView v = seeds_of(S.plan.arena, S.plan.fv); // tag[leaf] = keep, mr/adapt.hpp:286-290
            par_seeds(v, v.n,
                      [&](const smr_seed& sd)
                      {
                          const int l = sd.level & 0xff;
                          std::memset(tag + need(S.mesh.ref[l], l, sd.y, sd.z, sd.xs, sd.xs + sd.n - 1), TAG_KEEP, static_cast<size_t>(sd.n));
                      });
        }
        S.t_fp += now() - t0;
        update_ghost(S);
        t0 = now();
        TagEps te;
        for (int l = 0; l < SMR_MAX_LEVELS; ++l)
        {
            const int exponent = dim * (L - l);
            if (l > L || exponent >= 31)
            {
                te.eps[l] = te.fine_eps[l] = te.coarse_eps[l] = 0;
                continue;
            }
            const double eps_l = S.eps / (1 << exponent);            // mr/adapt.hpp:328-329
            te.eps[l]          = eps_l;
            te.fine_eps[l]     = std::pow(2.0, S.regularity + dim) * eps_l; // mr/criteria.hpp:31
            te.coarse_eps[l]   = te.fine_eps[l] / (1 << dim);        // mr/criteria.hpp:32
        }
        const int64_t detail_limit   = S.plan.detail_cum[static_cast<size_t>(std::max(std::min(L - ite, S.mesh.nlev), 0))];
        const int64_t criteria_limit = (L - ite) >= 0 ? S.plan.tag_cum[static_cast<size_t>(L - ite)] : 0;
        compute_detail(S, records_below(seeds_of(S.plan.arena, S.plan.detail), detail_limit));
        criteria(S, te, records_below(seeds_of(S.plan.arena, S.plan.tag_all), criteria_limit));
        for (int level = L; level >= 1; --level)
        {
            maximum(S, level);
        }
        S.t_fp += now() - t0;

        // tags -> leaves, graduation, fixed-point test (mr/adapt.hpp:360-379)
        t0 = now();
        bool no_tag_changes = false;
        CellArray ca        = cells_from_tags(S.mesh, S.tag.data(), &no_tag_changes);
        if (no_tag_changes && S.graduated)
        {
            S.t_mesh += now() - t0;
            return true;
        }
        if (no_tag_changes)
        {
            ca = S.mesh.cells;
            for (LevelSet& s : ca)
            {
                s.off.clear();
            }
        }
        bool same = same_cells(ca, S.mesh.cells);
        if (!same || !S.graduated)
        {
            make_graduation(cfg, ca);
            same = same_cells(ca, S.mesh.cells);
        }
        S.graduated = true;
        if (same)
        {
            S.t_mesh += now() - t0;
            return true;
        }
        Mesh new_mesh;
        new_mesh.generation = S.mesh.generation;
        new_mesh.init_from_cells(cfg, std::move(ca));
        S.t_mesh += now() - t0;

        // update_fields (algorithm/update_fields.hpp:27-54,101-127): copy, project, predict into the new numbering
        t0 = now();
        build_transfer(S.mesh, new_mesh, S.tr, S.flt);
        S.t_batches += now() - t0;
        t0 = now();
        std::vector<double> nu(static_cast<size_t>(new_mesh.nref), 0.0);
        {
            const SeedView v = seeds_of(S.tr.arena, S.tr.copy);
            const double* src = S.u.data();
            double* dst       = nu.data();
            par_seeds(v, v.n,
                      [&](const smr_seed& sd)
                      {
                          const int l = sd.level & 0xff;
                          const int64_t d = need(new_mesh.ref[l], l, sd.y, sd.z, sd.xs, sd.xs + sd.n - 1);
                          const int64_t s = need(S.mesh.ref[l], l, sd.y, sd.z, sd.xs, sd.xs + sd.n - 1);
                          std::memcpy(dst + d, src + s, static_cast<size_t>(sd.n) * sizeof(double));
                      });
        }
        projection(new_mesh, S.mesh, dim, seeds_of(S.tr.arena, S.tr.proj), S.u.data(), nu.data());
        prediction(new_mesh, S.mesh, dim, cfg.pred_radius, seeds_of(S.tr.arena, S.tr.pred), S.u.data(), nu.data());
        S.u.swap(nu);
        S.mesh       = std::move(new_mesh);
        S.plan_ready = false;
        S.t_fp += now() - t0;
        (void)lmin;
        return false;
    }

    void adapt(Sim& S)
    {
        if (S.cfg.min_level == S.cfg.max_level)
        {
            return;
        }
        for (int ite = 0; ite < S.cfg.max_level - S.cfg.min_level; ++ite) // mr/adapt.hpp:162
        {
            if (harten(S, ite))
            {
                break;
            }
        }
    }

    // u[cell] = inside where |center - c|^2 <= r^2 (demos/FiniteVolume/advection_2d.cpp:23-45), leaves only
    void init_ball(Sim& S, const double* center, double radius, double inside, double outside)
    {
        ensure_plan(S);
        S.u.assign(static_cast<size_t>(S.mesh.nref), 0.0);
        const int dim = S.cfg.dim;
        const SeedView v = seeds_of(S.plan.arena, S.plan.fv);
        double* u = S.u.data();
        par_seeds(v, v.n,
                  [&](const smr_seed& sd)
                  {
                      const int l = sd.level & 0xff;
                      const int64_t c = need(S.mesh.ref[l], l, sd.y, sd.z, sd.xs, sd.xs + sd.n - 1);
                      const double length = S.cfg.scaling / static_cast<double>(1 << l);
                      for (int k = 0; k < sd.n; ++k)
                      {
                          const int idx[3] = {sd.xs + k, sd.y, sd.z};
                          double r2 = 0.0;
                          for (int d = 0; d < dim; ++d)
                          {
                              const double cc = S.cfg.origin[d] + length * (idx[d] + 0.5);
                              const double t  = (cc - center[d]) * (cc - center[d]);
                              r2 = d == 0 ? t : r2 + t;
                          }
                          u[c + k] = r2 <= radius * radius ? inside : outside;
                      }
                  });
    }

    template <class F>
    int guarded(Sim* S, F&& f)
    {
        try
        {
            f();
            return 0;
        }
        catch (const std::exception& e)
        {
            if (S)
            {
                S->error = e.what();
            }
            std::fprintf(stderr, "cpu_path: %s\n", e.what());
            return 1;
        }
    }
} // namespace

extern "C"
{
    void* cpu_sim_create(int dim, int min_level, int max_level, int pred_radius, int start_level, double eps, double regularity, int bc_type, double bc_value)
    {
        auto S = std::make_unique<Sim>();
        S->cfg.dim         = dim;
        S->cfg.min_level   = min_level;
        S->cfg.max_level   = max_level;
        S->cfg.pred_radius = pred_radius;
        S->eps             = eps;
        S->regularity      = regularity;
        S->bc_type         = bc_type;
        S->bc_value        = bc_value;
        if (guarded(S.get(), [&] { S->mesh.init_uniform(S->cfg, start_level); }))
        {
            return nullptr;
        }
        S->u.assign(static_cast<size_t>(S->mesh.nref), 0.0);
        return S.release();
    }

    // start from given leaves: `n` x-intervals (level, y, z, xs, xe) and the leaf values in for_each_cell order
    void* cpu_sim_create_from_leaves(int dim, int min_level, int max_level, int pred_radius, double eps, double regularity, int bc_type, double bc_value,
                                     const int32_t* ivl5, int64_t n, const double* leaf_values)
    {
        auto S = std::make_unique<Sim>();
        S->cfg.dim         = dim;
        S->cfg.min_level   = min_level;
        S->cfg.max_level   = max_level;
        S->cfg.pred_radius = pred_radius;
        S->eps             = eps;
        S->regularity      = regularity;
        S->bc_type         = bc_type;
        S->bc_value        = bc_value;
        const int rc = guarded(S.get(),
                               [&]
                               {
                                   const int nlev = Mesh::levels_for(S->cfg);
                                   std::vector<SetBuilder> b(static_cast<size_t>(nlev));
                                   for (int64_t i = 0; i < n; ++i)
                                   {
                                       b[static_cast<size_t>(ivl5[5 * i])].add(mk_key(ivl5[5 * i + 1], ivl5[5 * i + 2]), ivl5[5 * i + 3], ivl5[5 * i + 4]);
                                   }
                                   CellArray ca(static_cast<size_t>(nlev));
                                   for (int l = 0; l < nlev; ++l)
                                   {
                                       ca[static_cast<size_t>(l)] = b[static_cast<size_t>(l)].build();
                                   }
                                   S->mesh.init_from_cells(S->cfg, std::move(ca));
                                   S->u.assign(static_cast<size_t>(S->mesh.nref), 0.0);
                                   int64_t k = 0;
                                   for (int l = 0; l < S->mesh.nlev; ++l)
                                   {
                                       const LevelSet& c = S->mesh.cells[static_cast<size_t>(l)];
                                       for (size_t q = 0; q < c.n_intervals(); ++q)
                                       {
                                           for (int x = 0; x < c.xe[q] - c.xs[q]; ++x)
                                           {
                                               S->u[static_cast<size_t>(c.off[q] + x)] = leaf_values[k++];
                                           }
                                       }
                                   }
                               });
        return rc ? nullptr : S.release();
    }

    void cpu_sim_destroy(void* h)
    {
        delete static_cast<Sim*>(h);
    }

    int cpu_sim_threads(void)
    {
        return omp_get_max_threads();
    }

    // the OpenMP runtime is shared with whatever else the process loaded (torchrun pins OMP_NUM_THREADS=1): let the caller
    // give this arm all host cores and restore the previous setting afterwards
    int cpu_sim_set_threads(int n)
    {
        const int before = omp_get_max_threads();
        if (n > 0)
        {
            omp_set_num_threads(n);
        }
        return before;
    }

    int cpu_sim_init_ball(void* h, const double* center, double radius, double inside, double outside)
    {
        Sim* S = static_cast<Sim*>(h);
        return guarded(S, [&] { init_ball(*S, center, radius, inside, outside); });
    }

    int cpu_sim_adapt(void* h)
    {
        Sim* S = static_cast<Sim*>(h);
        return guarded(S, [&] { adapt(*S); });
    }

    int cpu_sim_update_ghost(void* h)
    {
        Sim* S = static_cast<Sim*>(h);
        return guarded(S, [&] { update_ghost(*S); });
    }

    // n_steps of the demo loop: MRadaptation -> update_ghost_mr -> unp1 = u - dt * upwind(a, u) -> swap
    // (demos/FiniteVolume/advection_2d.cpp:129-153).  Returns the number of cell updates through *cell_updates.
    int cpu_sim_steps(void* h, int n_steps, const double* a, double dt, int64_t* cell_updates)
    {
        Sim* S = static_cast<Sim*>(h);
        return guarded(S,
                       [&]
                       {
                           int64_t total = 0;
                           for (int i = 0; i < n_steps; ++i)
                           {
                               adapt(*S);
                               update_ghost(*S);
                               const double t0 = now();
                               S->unp1.assign(static_cast<size_t>(S->mesh.nref), 0.0);
                               fv_upwind(*S, a, dt);
                               S->u.swap(S->unp1);
                               S->t_fp += now() - t0;
                               total += S->mesh.nleaves;
                           }
                           if (cell_updates)
                           {
                               *cell_updates = total;
                           }
                       });
    }

    int64_t cpu_sim_nb_cells(void* h, int reference)
    {
        Sim* S = static_cast<Sim*>(h);
        return reference ? S->mesh.nref : S->mesh.nleaves;
    }

    int64_t cpu_sim_nb_leaf_intervals(void* h)
    {
        Sim* S    = static_cast<Sim*>(h);
        int64_t n = 0;
        for (const LevelSet& c : S->mesh.cells)
        {
            n += static_cast<int64_t>(c.n_intervals());
        }
        return n;
    }

    // leaves as (level, y, z, xs, xe) rows in for_each_cell order, and the leaf values in the same order
    int cpu_sim_get_leaves(void* h, int32_t* ivl5, double* leaf_values)
    {
        Sim* S = static_cast<Sim*>(h);
        return guarded(S,
                       [&]
                       {
                           int64_t i = 0, k = 0;
                           for (int l = 0; l < S->mesh.nlev; ++l)
                           {
                               const LevelSet& c = S->mesh.cells[static_cast<size_t>(l)];
                               for (size_t r = 0; r < c.rows(); ++r)
                               {
                                   for (int q = c.ptr[r]; q < c.ptr[r + 1]; ++q)
                                   {
                                       if (ivl5)
                                       {
                                           ivl5[5 * i]     = l;
                                           ivl5[5 * i + 1] = key_y(c.key[r]);
                                           ivl5[5 * i + 2] = key_z(c.key[r]);
                                           ivl5[5 * i + 3] = c.xs[static_cast<size_t>(q)];
                                           ivl5[5 * i + 4] = c.xe[static_cast<size_t>(q)];
                                       }
                                       ++i;
                                       if (leaf_values)
                                       {
                                           for (int x = 0; x < c.xe[static_cast<size_t>(q)] - c.xs[static_cast<size_t>(q)]; ++x)
                                           {
                                               leaf_values[k++] = S->u[static_cast<size_t>(c.off[static_cast<size_t>(q)] + x)];
                                           }
                                       }
                                   }
                               }
                           }
                       });
    }

    // the whole reference-sized field (storage numbering of the reference mesh: cell_array.hpp:484-493) and the tags /
    // details of the last harten iteration, for comparisons with the numpy oracle
    int cpu_sim_get_field(void* h, double* out)
    {
        Sim* S = static_cast<Sim*>(h);
        std::memcpy(out, S->u.data(), S->u.size() * sizeof(double));
        return 0;
    }

    void cpu_sim_times(void* h, double* mesh_s, double* batch_s, double* fp_s, int reset)
    {
        Sim* S   = static_cast<Sim*>(h);
        *mesh_s  = S->t_mesh;
        *batch_s = S->t_batches;
        *fp_s    = S->t_fp;
        if (reset)
        {
            S->t_mesh = S->t_batches = S->t_fp = 0;
        }
    }
}
