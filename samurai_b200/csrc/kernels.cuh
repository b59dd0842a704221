// sm_100a kernels for samurai's per-time-step hot path.  All fp64, HBM-bound streaming/gather work: no tensor cores.
// Compiled with -fmad=false: the x86-64 reference build has no FMA contraction, and tags are decided by
// |detail| > eps comparisons, so the operation order below follows the reference expression by expression.
//
// Execution model: one thread per output cell.  A batch is a list of x-interval records with an exclusive prefix sum
// of their lengths; CTA b owns output cells [b*SMR_CTA_CELLS, (b+1)*SMR_CTA_CELLS), stages the slice of the prefix
// array it needs in shared memory and binary-searches it per cell.  Consecutive threads map to consecutive cells of
// an interval, so every row access is a coalesced 8-byte-per-lane stream; long intervals (uniform meshes) degenerate
// to one record per CTA.
#pragma once
#include "items.h"

#include <cuda_runtime.h>

namespace smr
{
    template <class Item>
    struct BatchView
    {
        const Item* items;
        const int64_t* prefix;
        const int32_t* cta_first;
        int64_t n_cells;
    };

    // body of one CTA of a batch: output cells [cta * SMR_CTA_CELLS, (cta + 1) * SMR_CTA_CELLS) ∩ [0, n_cells)
    template <class Item, class Op>
    __device__ __forceinline__ void run_batch_cta(const BatchView<Item>& b, const Op& op, int cta)
    {
        __shared__ int64_t s_prefix[SMR_CTA_CELLS + 2];
        const int first = b.cta_first[cta];
        const int nloc  = b.cta_first[cta + 1] - first + 1;
        for (int i = threadIdx.x; i <= nloc; i += SMR_CTA_THREADS)
        {
            s_prefix[i] = b.prefix[first + i];
        }
        __syncthreads();
        const int64_t base = static_cast<int64_t>(cta) * SMR_CTA_CELLS;
#pragma unroll
        for (int k = 0; k < SMR_CELLS_PER_THREAD; ++k)
        {
            const int64_t g = base + threadIdx.x + k * SMR_CTA_THREADS;
            if (g < b.n_cells)
            {
                int lo = 0, hi = nloc - 1;
                while (lo < hi)
                {
                    const int mid = (lo + hi + 1) >> 1;
                    if (s_prefix[mid] <= g)
                    {
                        lo = mid;
                    }
                    else
                    {
                        hi = mid - 1;
                    }
                }
                op(b.items[first + lo], static_cast<int>(g - s_prefix[lo]));
            }
        }
    }

    template <class Item, class Op>
    __global__ void __launch_bounds__(SMR_CTA_THREADS) batch_kernel(BatchView<Item> b, Op op)
    {
        run_batch_cta(b, op, blockIdx.x);
    }

    // ------------------------------------------------------------------------------------------------------------
    // FV field expressions: unp1 = u - dt * upwind(a, u)      (stencil_field.hpp:30-54, 92-97; field_base.hpp:230-242)
    // ------------------------------------------------------------------------------------------------------------
    struct FvParams
    {
        double half_a[3];     // .5 * a[d]
        double half_abs_a[3]; // .5 * |a[d]|
        double a[3];
        double dt;
        double dx[SMR_MAX_LEVELS];
    };

    __device__ __forceinline__ double upwind_flux(double ha, double haa, double ul, double ur)
    {
        return ha * (ul + ur) + haa * (ul - ur);
    }

    // upwind_scalar_burgers_op::flux (stencil_field.hpp:193-217)
    __device__ __forceinline__ double burgers_flux(double a, double ul, double ur)
    {
        const bool mask1 = (a * ul) < (a * ur);
        const bool mask2 = (ul * ur) > 0.0;
        const double mn  = fmin(fabs(ul), fabs(ur));
        const double mx  = fmax(fabs(ul), fabs(ur));
        double out       = 0.0;
        if (mask1 && mask2)
        {
            out = .5 * mn * mn;
        }
        if (!mask1)
        {
            out = .5 * mx * mx;
        }
        return out;
    }

    template <int DIM, bool BURGERS>
    struct FvOp
    {
        const double* __restrict__ u;
        double* __restrict__ out;
        FvParams p;

        __device__ __forceinline__ double flux(int d, double ul, double ur) const
        {
            if (BURGERS)
            {
                return burgers_flux(p.a[d], ul, ur);
            }
            return upwind_flux(p.half_a[d], p.half_abs_a[d], ul, ur);
        }

        __device__ __forceinline__ void operator()(const smr_item_fv& it, int k) const
        {
            const double* c = u + it.c + k;
            const double uc = c[0];
            double acc      = -flux(0, c[-1], uc) + flux(0, uc, c[1]);
            if (DIM > 1)
            {
                acc = (acc + -flux(1, u[it.ym + k], uc)) + flux(1, uc, u[it.yp + k]);
            }
            if (DIM > 2)
            {
                acc = (acc + -flux(2, u[it.zm + k], uc)) + flux(2, uc, u[it.zp + k]);
            }
            out[it.c + k] = uc - p.dt * (acc / p.dx[it.level]);
        }
    };

    // ------------------------------------------------------------------------------------------------------------
    // projection (numeric/projection.hpp:22-64)
    // ------------------------------------------------------------------------------------------------------------
    template <int DIM>
    struct ProjOp
    {
        const double* src; // may alias dst (ghost update projects inside one field)
        double* dst;

        __device__ __forceinline__ void operator()(const smr_item_proj& it, int k) const
        {
            double sum = 0.0;
#pragma unroll
            for (int r = 0; r < (1 << (DIM - 1)); ++r)
            {
                const double* s = src + it.src[r] + 2 * k;
                sum += s[0] + s[1];
            }
            dst[it.dst + k] = sum * (1.0 / static_cast<double>(1 << DIM));
        }
    };

    // ------------------------------------------------------------------------------------------------------------
    // prediction (numeric/prediction.hpp:259-361 ghost form, :107-257 update_fields form)
    // ------------------------------------------------------------------------------------------------------------
    __device__ __forceinline__ double interp1(int parity, int k)
    {
        // interp_coeffs<3>(+1) = {1/8, 1, -1/8} for even cells, (-1) for odd (prediction.hpp:31-35, 304-306)
        const double s = parity ? -0.125 : 0.125;
        return k == 0 ? s : (k == 1 ? 1.0 : -s);
    }

    template <int DIM, int RADIUS>
    struct PredOp
    {
        const double* src; // may alias dst (ghost update predicts inside one field)
        double* dst;

        __device__ __forceinline__ void operator()(const smr_item_pred& it, int k) const
        {
            const int ic = ((it.par & 1) + k) >> 1;
            if (RADIUS == 0)
            {
                dst[it.dst + k] = src[it.src[4] + ic];
                return;
            }
            const int px = (it.par ^ k) & 1;
            const int py = (it.par >> 1) & 1;
            const int pz = (it.par >> 2) & 1;
            double val   = 0.0;
#pragma unroll
            for (int rz = (DIM > 2 ? 0 : 1); rz < (DIM > 2 ? 3 : 2); ++rz)
            {
#pragma unroll
                for (int ry = (DIM > 1 ? 0 : 1); ry < (DIM > 1 ? 3 : 2); ++ry)
                {
                    const double* row = src + it.src[ry + 3 * rz] + ic - 1;
#pragma unroll
                    for (int rx = 0; rx < 3; ++rx)
                    {
                        double coeff = 1.0 * interp1(px, rx);
                        if (DIM > 1)
                        {
                            coeff *= interp1(py, ry);
                        }
                        if (DIM > 2)
                        {
                            coeff *= interp1(pz, rz);
                        }
                        val += row[rx] * coeff;
                    }
                }
            }
            dst[it.dst + k] = val;
        }
    };

    // ------------------------------------------------------------------------------------------------------------
    // detail (mr/operators.hpp:146-175, 226-358, 360-533)
    // ------------------------------------------------------------------------------------------------------------
    template <int DIM, int RADIUS>
    struct DetailOp
    {
        const double* __restrict__ f;
        double* __restrict__ detail;

        __device__ __forceinline__ void operator()(const smr_item_detail& it, int k) const
        {
            constexpr int NR = 1 << (DIM - 1);
            double d[NR][2];
#pragma unroll
            for (int r = 0; r < NR; ++r)
            {
                const double* c = f + it.fine[r] + 2 * k;
                d[r][0]         = c[0];
                d[r][1]         = c[1];
            }
            if (RADIUS == 0)
            {
                const double s = f[it.coarse[4] + k];
#pragma unroll
                for (int r = 0; r < NR; ++r)
                {
                    d[r][0] -= s;
                    d[r][1] -= s;
                }
            }
            else
            {
#pragma unroll
                for (int rz = (DIM > 2 ? 0 : 1); rz < (DIM > 2 ? 3 : 2); ++rz)
                {
#pragma unroll
                    for (int ry = (DIM > 1 ? 0 : 1); ry < (DIM > 1 ? 3 : 2); ++ry)
                    {
                        const double* row = f + it.coarse[ry + 3 * rz] + k - 1;
#pragma unroll
                        for (int rx = 0; rx < 3; ++rx)
                        {
                            const double s = row[rx];
#pragma unroll
                            for (int r = 0; r < NR; ++r)
                            {
                                const int py = r & 1, pz = r >> 1;
#pragma unroll
                                for (int px = 0; px < 2; ++px)
                                {
                                    double coeff = interp1(px, rx);
                                    if (DIM > 1)
                                    {
                                        coeff *= interp1(py, ry);
                                    }
                                    if (DIM > 2)
                                    {
                                        coeff *= interp1(pz, rz);
                                    }
                                    d[r][px] -= coeff * s;
                                }
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < NR; ++r)
            {
                double* o = detail + it.fine[r] + 2 * k;
                o[0]      = d[r][0];
                o[1]      = d[r][1];
            }
        }
    };

    // ------------------------------------------------------------------------------------------------------------
    // tagging criteria (mr/criteria.hpp:19-113) and keep propagation (mr/operators.hpp:29-89)
    // ------------------------------------------------------------------------------------------------------------
    struct TagParams
    {
        double eps[SMR_MAX_LEVELS];        // eps_l, indexed by the children's level
        double fine_eps[SMR_MAX_LEVELS];   // 2^(regularity+dim) * eps_l
        double coarse_eps[SMR_MAX_LEVELS]; // fine_eps / 2^dim
        int min_level, max_level;
    };

    template <int DIM>
    struct CriteriaOp
    {
        const double* __restrict__ detail; // ncomp arrays of `stride` entries (one per adapted field)
        uint8_t* __restrict__ tag;
        TagParams p;
        int ncomp;
        int64_t stride;

        __device__ __forceinline__ void operator()(const smr_item_tag& it, int k) const
        {
            constexpr int NR = 1 << (DIM - 1);
            const int fl     = it.level;
            bool coarsen_ok  = fl > p.min_level;
            bool refine[NR][2];
#pragma unroll
            for (int r = 0; r < NR; ++r)
            {
                refine[r][0] = refine[r][1] = false;
            }
            for (int c = 0; c < ncomp; ++c)
            {
                const double* d = detail + c * stride;
                if (fabs(d[it.coarse + k]) > p.coarse_eps[fl])
                {
                    coarsen_ok = false;
                }
#pragma unroll
                for (int r = 0; r < NR; ++r)
                {
#pragma unroll
                    for (int x = 0; x < 2; ++x)
                    {
                        const double v = fabs(d[it.fine[r] + 2 * k + x]);
                        if (v > p.eps[fl])
                        {
                            coarsen_ok = false;
                        }
                        if (v > p.fine_eps[fl])
                        {
                            refine[r][x] = true;
                        }
                    }
                }
            }
            if (coarsen_ok)
            {
#pragma unroll
                for (int r = 0; r < NR; ++r)
                {
                    tag[it.fine[r] + 2 * k]     = 2;
                    tag[it.fine[r] + 2 * k + 1] = 2;
                }
            }
            if (fl < p.max_level)
            {
#pragma unroll
                for (int r = 0; r < NR; ++r)
                {
#pragma unroll
                    for (int x = 0; x < 2; ++x)
                    {
                        if (refine[r][x])
                        {
                            tag[it.fine[r] + 2 * k + x] |= 4;
                        }
                    }
                }
            }
        }
    };

    template <int DIM>
    struct MaximumOp
    {
        uint8_t* __restrict__ tag;

        __device__ __forceinline__ void operator()(const smr_item_tag& it, int k) const
        {
            constexpr int NR = 1 << (DIM - 1);
            uint8_t any = 0, all = 0xff;
#pragma unroll
            for (int r = 0; r < NR; ++r)
            {
                const uint8_t a = tag[it.fine[r] + 2 * k], b = tag[it.fine[r] + 2 * k + 1];
                any |= a | b;
                all &= a & b;
            }
            if (any & 1)
            {
#pragma unroll
                for (int r = 0; r < NR; ++r)
                {
                    tag[it.fine[r] + 2 * k] |= 1;
                    tag[it.fine[r] + 2 * k + 1] |= 1;
                }
                tag[it.coarse + k] |= 1;
            }
            else if (all & 2)
            {
                tag[it.coarse + k] |= 1;
            }
            else
            {
#pragma unroll
                for (int r = 0; r < NR; ++r)
                {
                    tag[it.fine[r] + 2 * k] &= static_cast<uint8_t>(~2);
                    tag[it.fine[r] + 2 * k + 1] &= static_cast<uint8_t>(~2);
                }
            }
        }
    };

    // tag[leaf] = keep (mr/adapt.hpp:286-290), driven by the FV leaf batch
    struct KeepLeavesOp
    {
        uint8_t* __restrict__ tag;

        __device__ __forceinline__ void operator()(const smr_item_fv& it, int k) const
        {
            tag[it.c + k] = 1;
        }
    };

    // u[cell] = (|center - c|^2 <= r^2) ? inside : outside   -- the demos' initial condition
    // (demos/FiniteVolume/advection_2d.cpp:23-45, advection_3d.cpp:32-45, scalar_burgers_2d.cpp:20-50)
    template <int DIM>
    struct InitBallOp
    {
        double* __restrict__ u;
        double origin[3];
        double scaling;
        double center[3];
        double radius;
        double inside;
        bool overwrite_outside;
        double outside;

        __device__ __forceinline__ void operator()(const smr_item_fv& it, int k) const
        {
            const double length = scaling / static_cast<double>(1 << it.level); // cell.hpp:16-20
            const int idx[3]    = {it.x + k, it.y, it.z};
            double r2           = 0.0;
#pragma unroll
            for (int d = 0; d < DIM; ++d)
            {
                const double c = origin[d] + length * (idx[d] + 0.5); // Cell::center, cell.hpp:131-134
                const double t = (c - center[d]) * (c - center[d]);
                r2             = d == 0 ? t : r2 + t;
            }
            if (r2 <= radius * radius)
            {
                u[it.c + k] = inside;
            }
            else if (overwrite_outside)
            {
                u[it.c + k] = outside;
            }
        }
    };

    struct CopyOp
    {
        const double* __restrict__ src;
        double* __restrict__ dst;

        __device__ __forceinline__ void operator()(const smr_item_copy& it, int k) const
        {
            dst[it.dst + k] = src[it.src + k];
        }
    };

    // ------------------------------------------------------------------------------------------------------------
    // boundary ghosts: one thread per ghost cell (update_outer_ghost.hpp, bc/dirichlet.hpp:29, bc/neumann.hpp:27-28)
    // ------------------------------------------------------------------------------------------------------------
    struct BcView
    {
        const smr_item_bc* items;
        const int64_t* srcs;
        int n;
        int bc_type;
        double bc_value;
    };

    __device__ __forceinline__ void run_bc(const BcView& v, double* __restrict__ f, int i)
    {
        if (i >= v.n)
        {
            return;
        }
        const smr_item_bc it = v.items[i];
        const int64_t* s     = v.srcs + it.src_first;
        if (it.kind == SMR_BC_COPY)
        {
            f[it.dst] = f[s[0]];
        }
        else if (it.kind == SMR_BC_VALUE)
        {
            if (v.bc_type == SMR_BCTYPE_DIRICHLET)
            {
                f[it.dst] = 2 * v.bc_value - f[s[0]];
            }
            else
            {
                f[it.dst] = it.coef * v.bc_value + f[s[0]];
            }
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

    // One level of the top-down ghost sweep in a single launch: CTAs [0, bc_ctas) fill the outside ghosts of the level,
    // the remaining CTAs run the projection level -> level-1.  The two halves never touch the same cells (batches.hpp).
    template <int DIM>
    __global__ void __launch_bounds__(SMR_CTA_THREADS) ghost_phase_kernel(BcView bc, int bc_ctas, BatchView<smr_item_proj> pv, double* __restrict__ f)
    {
        if (static_cast<int>(blockIdx.x) < bc_ctas)
        {
            run_bc(bc, f, blockIdx.x * SMR_CTA_THREADS + threadIdx.x);
        }
        else
        {
            run_batch_cta(pv, ProjOp<DIM>{f, f}, static_cast<int>(blockIdx.x) - bc_ctas);
        }
    }
} // namespace smr
