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

#ifndef STRIP_UPT
#define STRIP_UPT 4
#endif
#ifndef STRIP_MIN_BLOCKS
#define STRIP_MIN_BLOCKS 6
#endif
// resident CTAs per SM the streaming operators are compiled for (register cap = 65536 / (256 * blocks)); measured on the
// uniform level-13 step (profiles/r02_summary.md)
#ifndef SMR_MB_DETAIL
#define SMR_MB_DETAIL 6
#endif
#ifndef SMR_MB_DETAIL3D
#define SMR_MB_DETAIL3D 4 /* 64 registers, no spills: 759 -> 550 us at level 9 (2: 128 registers, 24 % occupancy, latency bound; 5 and 6 spill) */
#endif
#ifndef SMR_MB_CRIT3D
#define SMR_MB_CRIT3D 6 /* 3D criteria: 249 (4) / 239 (6) / 273 us (8 CTAs per SM) at level 9 */
#endif
#ifndef SMR_MB_MAX3D
#define SMR_MB_MAX3D 8 /* 3D keep propagation: 211 (4) / 158 (6) / 141 us (8) */
#endif
#ifndef SMR_MB_CRITERIA
#define SMR_MB_CRITERIA 6
#endif
#ifndef SMR_MB_MAXIMUM
#define SMR_MB_MAXIMUM 6
#endif
#ifndef SMR_MB_COPY
#define SMR_MB_COPY 6
#endif
#ifndef SMR_MB_PROJ
#define SMR_MB_PROJ 6
#endif

#include <cooperative_groups.h>
#include <cuda_runtime.h>

namespace smr
{
    // ------------------------------------------------------------------------------------------------------------
    // multi-GPU: every buffer kernels write lives in a per-rank pool mapped into all peers (CUDA IPC) at the same
    // offsets, so the address of a cell in peer p's copy is the local address + delta[p].  A record's mask says which
    // peers need its outputs (their slab, widened by the stencil reach, contains the record): the producing thread
    // stores there directly over NVLink, fused into the kernel -- no pack / send / unpack pass.
    // ------------------------------------------------------------------------------------------------------------
    struct PeerTable
    {
        long long delta[SMR_MAX_RANKS]; // peer pool base - local pool base, bytes
        unsigned long long* flags;      // local barrier flags, one per rank
        unsigned long long* error;      // set when a barrier times out
        int rank, world;
    };

    __constant__ PeerTable g_peers;

    template <class T>
    __device__ __forceinline__ void mstore(T* addr, T v, unsigned mask)
    {
        *addr = v;
        if (mask)
        {
#pragma unroll
            for (int p = 0; p < SMR_MAX_RANKS; ++p)
            {
                if ((mask >> p) & 1u)
                {
                    *reinterpret_cast<T*>(reinterpret_cast<char*>(addr) + g_peers.delta[p]) = v;
                }
            }
        }
    }

#ifdef SMR_MISC_KERNELS // non-template kernels: compiled by capi.cu only
    // all ranks have finished everything queued before this kernel (and their peer stores have landed) once it returns
    __global__ void mg_barrier_kernel(unsigned long long epoch)
    {
        const int t = threadIdx.x;
        if (t < g_peers.world && t != g_peers.rank)
        {
            __threadfence_system();
            volatile unsigned long long* remote = reinterpret_cast<volatile unsigned long long*>(
                reinterpret_cast<char*>(g_peers.flags) + g_peers.delta[t]);
            remote[g_peers.rank] = epoch;
            __threadfence_system();
            volatile unsigned long long* mine = g_peers.flags;
            long long spins                   = 0;
            while (mine[t] < epoch)
            {
                if (++spins > 400000000LL) // ~ a second: a peer died; fail loudly instead of hanging the GPU
                {
                    *g_peers.error = epoch;
                    break;
                }
            }
        }
    }
#endif

    // Field pointers of an op: __restrict__ (read-only data path, loads hoisted above stores) when the op runs as its own
    // launch; plain when it runs inside the fused wavefront kernel, where other phases of the same launch write the arrays
    template <class T, bool R>
    struct PtrOf
    {
        using type = T* __restrict__;
    };
    template <class T>
    struct PtrOf<T, false>
    {
        using type = T*;
    };

    template <class Item>
    struct BatchView
    {
        const Item* items;
        const int64_t* prefix;
        const int32_t* cta_first;
        int64_t n_cells;
    };

    // ops may offer span(it, k0, count): the whole CTA processes `count` consecutive units of ONE record cooperatively
    // (byte-granular ops on long records: vector accesses instead of one byte per thread)
    template <class Op, class = void>
    struct HasSpan
    {
        static constexpr bool value = false;
    };
    template <class Op>
    struct HasSpan<Op, decltype(void(Op::has_span))>
    {
        static constexpr bool value = Op::has_span;
    };

    // body of one CTA of a batch: output units [cta * U, (cta + 1) * U) ∩ [0, n_cells), U = 256 * Op::units_per_thread
    // `s_prefix`: shared staging for SMR_CTA_THREADS * Op::units_per_thread + 2 entries, owned by the calling kernel
    template <class Item, class Op>
    __device__ __forceinline__ void run_batch_cta(const BatchView<Item>& b, const Op& op, int cta, int32_t* s_prefix)
    {
        constexpr int UPT   = Op::units_per_thread;
        constexpr int UNITS = SMR_CTA_THREADS * UPT;
        const int first    = b.cta_first[cta];
        const int nloc     = b.cta_first[cta + 1] - first + 1;
        const int64_t base = static_cast<int64_t>(cta) * UNITS;
        const int64_t left = b.n_cells - base; // output units from `base` to the end of the batch
        if (nloc == 1)
        {
            // the whole CTA lies inside one record (long intervals: uniform or nearly uniform levels): no staging,
            // no search, the record is read once and every access is base pointer + small immediate
            const Item it = b.items[first];
            if constexpr (HasSpan<Op>::value)
            {
                if (op.span(it, static_cast<int>(base - b.prefix[first]), static_cast<int>(left >= UNITS ? UNITS : left)))
                {
                    return;
                }
            }
            const int k0  = static_cast<int>(base - b.prefix[first]) + threadIdx.x;
            if (left >= UNITS)
            {
                if constexpr (Op::two_phase)
                {
                    // __restrict__ on functor members does not let the compiler move loads above earlier stores, so ops
                    // that are worth it expose load()/finish(): all loads first, then the arithmetic and the stores
                    typename Op::Vals v[UPT];
#pragma unroll
                    for (int k = 0; k < UPT; ++k)
                    {
                        v[k] = op.load(it, k0 + k * SMR_CTA_THREADS);
                    }
#pragma unroll
                    for (int k = 0; k < UPT; ++k)
                    {
                        op.finish(it, k0 + k * SMR_CTA_THREADS, v[k]);
                    }
                }
                else if constexpr (Op::warp_uniform)
                {
                    // all 32 lanes of every warp are active, in the same record, on consecutive cells: an op may trade its
                    // i-1 / i+1 loads for warp shuffles.  Tried for the strip kernels (profiles/r01_summary.md section 8): slower
                    // than the L1 loads at 6 CTAs/SM (4.5 vs 5.0 TB/s), so no op sets warp_uniform today.
#pragma unroll
                    for (int k = 0; k < UPT; ++k)
                    {
                        op.warp(it, k0 + k * SMR_CTA_THREADS);
                    }
                }
                else
                {
#pragma unroll
                    for (int k = 0; k < UPT; ++k)
                    {
                        op(it, k0 + k * SMR_CTA_THREADS);
                    }
                }
                return;
            }
#pragma unroll
            for (int k = 0; k < UPT; ++k)
            {
                if (static_cast<int64_t>(threadIdx.x) + k * SMR_CTA_THREADS < left)
                {
                    op(it, k0 + k * SMR_CTA_THREADS);
                }
            }
            return;
        }
        // prefix relative to the CTA's first output unit: fits 32 bits (a record holds < 2^31 cells)
        for (int i = threadIdx.x; i <= nloc; i += SMR_CTA_THREADS)
        {
            const int64_t rel = b.prefix[first + i] - base;
            s_prefix[i]       = rel > 2 * UNITS ? 2 * UNITS : static_cast<int32_t>(rel < -2147483647LL ? -2147483647LL : rel);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < UPT; ++k)
        {
            const int g = threadIdx.x + k * SMR_CTA_THREADS;
            if (g < left)
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
                op(b.items[first + lo], g - s_prefix[lo]);
            }
        }
    }

    // Quarter of a chunk: output units [cta * U + sub * 256, + 256) with ONE unit per thread.  Used by the fused wavefront
    // kernel, whose phases are latency bound (a few chunks, hundreds of idle CTAs): four CTAs share a chunk instead of one
    // thread walking four units back to back.
    template <class Item, class Op>
    __device__ __forceinline__ void run_batch_sub(const BatchView<Item>& b, const Op& op, int cta, int sub, int32_t* s_prefix)
    {
        constexpr int UNITS = SMR_CTA_THREADS * Op::units_per_thread;
        const int first     = b.cta_first[cta];
        const int nloc      = b.cta_first[cta + 1] - first + 1;
        const int64_t base  = static_cast<int64_t>(cta) * UNITS;
        const int64_t left  = b.n_cells - base;
        const int g         = sub * SMR_CTA_THREADS + threadIdx.x;
        if (nloc == 1)
        {
            if (g < left)
            {
                op(b.items[first], static_cast<int>(base - b.prefix[first]) + g);
            }
            return;
        }
        for (int i = threadIdx.x; i <= nloc; i += SMR_CTA_THREADS)
        {
            const int64_t rel = b.prefix[first + i] - base;
            s_prefix[i]       = rel > 2 * UNITS ? 2 * UNITS : static_cast<int32_t>(rel < -2147483647LL ? -2147483647LL : rel);
        }
        __syncthreads();
        if (g < left)
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
            op(b.items[first + lo], g - s_prefix[lo]);
        }
    }

    // One warp per record: byte-granular operators over the leaf intervals (keep tags, change flag) when they run as their
    // own launch.  A 1024-cell chunk is 1 KB of tags, so the chunked kernel spends its time on the per-CTA index chain
    // (measured 140-190 us for 67 MB on the uniform level-13 mesh); a warp streams its whole record with 16-byte accesses.
    template <class Item, class Op>
    __global__ void __launch_bounds__(SMR_CTA_THREADS) record_kernel(const Item* __restrict__ items, int n_items, Op op)
    {
        const int r = static_cast<int>(blockIdx.x) * (SMR_CTA_THREADS / 32) + static_cast<int>(threadIdx.x >> 5);
        if (r < n_items)
        {
            op.record(items[r], static_cast<int>(threadIdx.x & 31));
        }
    }

    template <class Item, class Op>
    __global__ void __launch_bounds__(SMR_CTA_THREADS, Op::min_blocks) batch_kernel(BatchView<Item> b, Op op)
    {
        __shared__ int32_t s_prefix[SMR_CTA_THREADS * Op::units_per_thread + 2];
        run_batch_cta(b, op, blockIdx.x, s_prefix);
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
        double inv_dx[SMR_MAX_LEVELS]; // 1/dx, used when dx is a power of two (x / dx == x * (1/dx) bit for bit)
        int exact_inv;
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

        static constexpr bool two_phase = true;
        static constexpr bool warp_uniform = false;
        static constexpr int min_blocks = 1;
        static constexpr int units_per_thread = SMR_CELLS_PER_THREAD;

        struct Vals
        {
            double c, xm, xp, ym, yp, zm, zp;
        };

        __device__ __forceinline__ Vals load(const smr_item_fv& it, int k) const
        {
            Vals v;
            const double* c = u + it.c + k;
            v.c             = c[0];
            v.xm            = c[-1];
            v.xp            = c[1];
            v.ym = v.yp = v.zm = v.zp = 0.0;
            if (DIM > 1)
            {
                v.ym = u[it.ym + k];
                v.yp = u[it.yp + k];
            }
            if (DIM > 2)
            {
                v.zm = u[it.zm + k];
                v.zp = u[it.zp + k];
            }
            return v;
        }

        __device__ __forceinline__ void finish(const smr_item_fv& it, int k, const Vals& v) const
        {
            const double uc = v.c;
            double acc      = -flux(0, v.xm, uc) + flux(0, uc, v.xp);
            if (DIM > 1)
            {
                acc = (acc + -flux(1, v.ym, uc)) + flux(1, uc, v.yp);
            }
            if (DIM > 2)
            {
                acc = (acc + -flux(2, v.zm, uc)) + flux(2, uc, v.zp);
            }
            const double div = p.exact_inv ? acc * p.inv_dx[it.level] : acc / p.dx[it.level];
            mstore(out + it.c + k, uc - p.dt * div, static_cast<unsigned>(it.mask));
        }

        __device__ __forceinline__ void operator()(const smr_item_fv& it, int k) const
        {
            finish(it, k, load(it, k));
        }
    };

    // Strip form of the same expression: the thread owns column k of SMR_STRIP_ROWS consecutive rows and keeps the column in
    // registers, so a row is read from L2 once (as itself) instead of three times (as itself and as the j-1 / j+1
    // neighbour of the rows around it).  Per cell the arithmetic is exactly FvOp's.
    template <int DIM, bool BURGERS>
    struct FvStripOp
    {
        static constexpr bool two_phase = false;
        static constexpr bool warp_uniform = false;
        static constexpr int min_blocks = STRIP_MIN_BLOCKS;
        static constexpr int units_per_thread = STRIP_UPT;

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

        __device__ __forceinline__ void operator()(const smr_item_fvstrip& it, int k) const
        {
            constexpr int R = SMR_STRIP_ROWS;
            // the column (the only loads that go to L2/DRAM) is fetched up front; the i-1 / i+1 values are L1 hits on lines
            // the neighbouring lanes just brought in and are read row by row to keep the register footprint small
            double v[R + 2];
#pragma unroll
            for (int r = 0; r < R + 2; ++r)
            {
                v[r] = u[it.row[r] + k];
            }
            const double inv = p.inv_dx[it.level], dx = p.dx[it.level];
#pragma unroll
            for (int r = 0; r < R; ++r)
            {
                const double* c = u + it.row[r + 1] + k;
                const double uc = v[r + 1];
                double acc      = -flux(0, c[-1], uc) + flux(0, uc, c[1]);
                acc             = (acc + -flux(1, v[r], uc)) + flux(1, uc, v[r + 2]);
                if (DIM > 2)
                {
                    acc = (acc + -flux(2, u[it.zm[r] + k], uc)) + flux(2, uc, u[it.zp[r] + k]);
                }
                const double div       = p.exact_inv ? acc * inv : acc / dx;
                mstore(out + it.row[r + 1] + k, uc - p.dt * div, static_cast<unsigned>(it.mask));
            }
        }
    };

    // ------------------------------------------------------------------------------------------------------------
    // Flux-based linear homogeneous schemes on a uniform-level mesh: out = S(u) for stencil-2 conservative fluxes
    // (make_convection_upwind, make_diffusion_order2).  The reference scatters `out[left] += lc[c]*u[st_c]`,
    // `out[right] += rc[c]*u[st_c]` interface interval by interface interval, then the boundary interfaces
    // (flux_based/explicit_flux_based_scheme__lin_hom.hpp:39-119, 233-319); a cell therefore receives, per direction,
    // the four products {lc0*u_c, lc1*u_+, rc0*u_-, rc1*u_c} in an order that depends only on whether it touches the low
    // or the high boundary and on the direction (x: both sides come from ONE interface interval; y/z: the lower row's
    // interval is visited before the cell's own).  This gather adds them in exactly that order.
    // ------------------------------------------------------------------------------------------------------------
    struct FluxParams
    {
        double lc[3][2]; // left-cell coefficients factor * flux_coeffs, per direction
        int n[3];        // domain size in cells at the mesh level
    };

    template <int DIM>
    struct FluxLinHomOp
    {
        static constexpr bool two_phase = false;
        static constexpr bool warp_uniform = false;
        static constexpr int min_blocks = 1;
        static constexpr int units_per_thread = SMR_CELLS_PER_THREAD;

        const double* __restrict__ u;
        double* __restrict__ out;
        FluxParams p;

        __device__ __forceinline__ void operator()(const smr_item_fv& it, int k) const
        {
            const double* c = u + it.c + k;
            const double uc = c[0];
            const int idx[3] = {it.x + k, it.y, it.z};
            double acc = 0.0;
#pragma unroll
            for (int d = 0; d < DIM; ++d)
            {
                const double um = d == 0 ? c[-1] : (d == 1 ? u[it.ym + k] : u[it.zm + k]);
                const double up = d == 0 ? c[1] : (d == 1 ? u[it.yp + k] : u[it.zp + k]);
                const double tLc = p.lc[d][0] * uc, tLp = p.lc[d][1] * up;
                const double tRm = (-p.lc[d][0]) * um, tRc = (-p.lc[d][1]) * uc;
                const bool low = idx[d] == 0, high = idx[d] == p.n[d] - 1;
                if (low)
                {
                    acc = (((acc + tLc) + tLp) + tRm) + tRc;
                }
                else if (d == 0 && !high)
                {
                    acc = (((acc + tLc) + tRm) + tLp) + tRc;
                }
                else
                {
                    acc = (((acc + tRm) + tRc) + tLc) + tLp;
                }
            }
            mstore(out + it.c + k, acc, static_cast<unsigned>(it.mask));
        }
    };

    // strip form (see FvStripOp): the column of SMR_STRIP_ROWS rows lives in registers
    template <int DIM>
    struct FluxLinHomStripOp
    {
        static constexpr bool two_phase = false;
        static constexpr bool warp_uniform = false;
        static constexpr int min_blocks = STRIP_MIN_BLOCKS;
        static constexpr int units_per_thread = STRIP_UPT;

        const double* __restrict__ u;
        double* __restrict__ out;
        FluxParams p;

        __device__ __forceinline__ double dir_terms(double acc, int d, int idx, double um, double uc, double up) const
        {
            const double tLc = p.lc[d][0] * uc, tLp = p.lc[d][1] * up;
            const double tRm = (-p.lc[d][0]) * um, tRc = (-p.lc[d][1]) * uc;
            if (idx == 0)
            {
                return (((acc + tLc) + tLp) + tRm) + tRc;
            }
            if (d == 0 && idx != p.n[0] - 1)
            {
                return (((acc + tLc) + tRm) + tLp) + tRc;
            }
            return (((acc + tRm) + tRc) + tLc) + tLp;
        }

        __device__ __forceinline__ void operator()(const smr_item_fvstrip& it, int k) const
        {
            constexpr int R = SMR_STRIP_ROWS;
            double v[R + 2];
#pragma unroll
            for (int r = 0; r < R + 2; ++r)
            {
                v[r] = u[it.row[r] + k];
            }
#pragma unroll
            for (int r = 0; r < R; ++r)
            {
                const double* c = u + it.row[r + 1] + k;
                const double uc = v[r + 1];
                double acc      = dir_terms(0.0, 0, it.x + k, c[-1], uc, c[1]);
                acc             = dir_terms(acc, 1, it.y + r, v[r], uc, v[r + 2]);
                if (DIM > 2)
                {
                    acc = dir_terms(acc, 2, it.z, u[it.zm[r] + k], uc, u[it.zp[r] + k]);
                }
                mstore(out + it.row[r + 1] + k, acc, static_cast<unsigned>(it.mask));
            }
        }
    };

    // Flux-based schemes on multi-level meshes, gather form.  The reference scatters, per direction: same-level
    // interfaces (all levels), then per level the jumps "coarse | fine" (A) and "fine | coarse" (B), then the boundary
    // interfaces in direction and opposite direction (flux_based_scheme__lin_hom.hpp:74-229, __nonlin.hpp:407-520).
    // For one cell that fixes the order of the terms of each side by the kind of neighbour across the face:
    //            minus side  plus side
    //   same        1           1        (x direction, linear: interleaved c = 0 left/right, c = 1 left/right; else minus first)
    //   coarse      2 (A)       2.5 (B)  cell = fine side, stencil = its own-level ghost inside the coarse leaf
    //   fine        3.5 (B)     3 (A)    cell = coarse side, one contribution per level+1 interface interval
    //   boundary    4.5         4
    // `tab` = [2][SMR_MAX_LEVELS][3][2] doubles: same-level left-cell coefficients h_factor(h,h)*coeffs(h) of the cell's
    // level, then the coarse-side jump coefficients h_factor(h_{l+1},h_l)*coeffs(h_{l+1}); right-cell coefficients are
    // their negatives.  Non-linear schemes keep the two factors in [.][l][0][0].
    template <int DIM, int NONLIN>
    struct FluxGenOp
    {
        static constexpr bool two_phase = false;
        static constexpr bool warp_uniform = false;
        static constexpr int min_blocks = 1;
        static constexpr int units_per_thread = SMR_CELLS_PER_THREAD;

        const double* __restrict__ u;
        double* __restrict__ out;
        const int64_t* __restrict__ aux;
        const double* __restrict__ tab;
        double scale;

        // scale * make_convection_upwind<Field>() on a scalar field (operators/convection_nonlin.hpp:63-76,
        // flux_based/algebraic_operators.hpp:38-46)
        __device__ __forceinline__ double nflux(double ul, double ur) const
        {
            const double v = 0.5 * (ul + ur);
            const double f = v >= 0 ? ul * ul : ur * ur;
            return scale != 1 ? f * scale : f;
        }

        __device__ __forceinline__ double side(double acc, const smr_item_flux& it, int k, int d, int plus, int kind, const double* c,
                                               const double* cs, const double* cj) const
        {
            const double uc = c[0];
            const double sg = plus ? 1.0 : -1.0; // right-cell coefficients / fluxes[1] are exact negations
            if (kind != SMR_FACE_FINE)
            {
                const double un = d == 0 ? c[plus ? 1 : -1] : u[it.nb[2 * (d - 1) + plus] + k];
                if (NONLIN)
                {
                    const double f = plus ? nflux(uc, un) : nflux(un, uc);
                    return acc + (sg * f) * cs[0];
                }
                const double s0 = plus ? uc : un, s1 = plus ? un : uc; // stencil {left cell, right cell}
                return (acc + (sg * cs[2 * d]) * s0) + (sg * cs[2 * d + 1]) * s1;
            }
            const int64_t* fx = aux + it.fine + (2 * d + plus) * 4;
            if (d == 0)
            {
#pragma unroll
                for (int r = 0; r < (1 << (DIM - 1)); ++r)
                {
                    const double s0 = u[fx[r]], s1 = u[fx[r] + 1];
                    if (NONLIN)
                    {
                        acc = acc + (sg * nflux(s0, s1)) * cj[0];
                    }
                    else
                    {
                        acc = acc + (sg * cj[0]) * s0;
                        acc = acc + (sg * cj[1]) * s1;
                    }
                }
                return acc;
            }
#pragma unroll
            for (int b = 0; b < (DIM > 2 ? 2 : 1); ++b)
            {
                const double* r0 = u + fx[2 * b] + 2 * k;
                const double* r1 = u + fx[2 * b + 1] + 2 * k;
                if (NONLIN)
                {
                    acc = acc + (sg * nflux(r0[0], r1[0])) * cj[0];
                    acc = acc + (sg * nflux(r0[1], r1[1])) * cj[0];
                }
                else
                {
                    const double a0 = sg * cj[2 * d], a1 = sg * cj[2 * d + 1];
                    acc = acc + (a0 * r0[0] + a0 * r0[1]);
                    acc = acc + (a1 * r1[0] + a1 * r1[1]);
                }
            }
            return acc;
        }

        __device__ __forceinline__ void operator()(const smr_item_flux& it, int k) const
        {
            const double* c  = u + it.c + k;
            const double* cs = tab + it.level * 6;
            const double* cj = tab + (SMR_MAX_LEVELS + it.level) * 6;
            double acc       = 0.0;
#pragma unroll
            for (int d = 0; d < DIM; ++d)
            {
                int km = (it.kinds >> (4 * d)) & 3, kp = (it.kinds >> (4 * d + 2)) & 3;
                if (d == 0)
                {
                    km = k != 0 ? SMR_FACE_SAME : km;
                    kp = k != it.n - 1 ? SMR_FACE_SAME : kp;
                }
                // same-level interfaces through a periodic boundary come after the regular ones (interface.hpp:83-92): first the one
                // seen from the cell left of the boundary (its plus face), then the one seen from the cell right of it (its minus face)
                bool m_thr = ((it.kinds >> (SMR_FLUXW_SWAP_SHIFT + d)) & 1) != 0, p_thr = ((it.kinds >> (SMR_FLUX_PLUS_THROUGH_SHIFT + d)) & 1) != 0;
                if (d == 0)
                {
                    m_thr = m_thr && k == 0;
                    p_thr = p_thr && k == it.n - 1;
                }
                if (!NONLIN && d == 0 && km == SMR_FACE_SAME && kp == SMR_FACE_SAME && !m_thr && !p_thr)
                {
                    // one interface interval holds both x-interfaces of the cell: for c = 0, 1 the left-cell loop runs
                    // before the right-cell loop (explicit_flux_based_scheme__lin_hom.hpp:63-76)
                    const double tLc = cs[0] * c[0], tLp = cs[1] * c[1];
                    const double tRm = (-cs[0]) * c[-1], tRc = (-cs[1]) * c[0];
                    acc = (((acc + tLc) + tRm) + tLp) + tRc;
                    continue;
                }
                // pass numbers x2: minus {2, 4, 7, 9}, plus {2, 5, 6, 8}; through the periodic boundary: plus 3, minus 3 after it
                const int pm = km == SMR_FACE_SAME ? (m_thr ? 3 : 2) : (km == SMR_FACE_COARSE ? 4 : (km == SMR_FACE_FINE ? 7 : 9));
                const int pp = kp == SMR_FACE_SAME ? (p_thr ? 3 : 2) : (kp == SMR_FACE_COARSE ? 5 : (kp == SMR_FACE_FINE ? 6 : 8));
                if (pm <= pp && !(m_thr && p_thr))
                {
                    acc = side(acc, it, k, d, 0, km, c, cs, cj);
                    acc = side(acc, it, k, d, 1, kp, c, cs, cj);
                }
                else
                {
                    acc = side(acc, it, k, d, 1, kp, c, cs, cj);
                    acc = side(acc, it, k, d, 0, km, c, cs, cj);
                }
            }
            mstore(out + it.c + k, acc, static_cast<unsigned>(it.mask));
        }
    };

    // scale * make_convection_upwind<VectorField>() with n_comp == dim (operators/convection_nonlin.hpp:24-76): in direction d the
    // interface flux is the whole vector  v >= 0 ? uL[d] * uL : uR[d] * uR,  v = .5 * (uL[d] + uR[d]).  Same records and term order
    // as FluxGenOp<DIM, 1>; the components are separate SoA arrays on the same mesh, so one record serves all of them.
    template <int DIM>
    struct FluxVecOp
    {
        static constexpr bool two_phase = false;
        static constexpr bool warp_uniform = false;
        static constexpr int min_blocks = 1;
        static constexpr int units_per_thread = SMR_CELLS_PER_THREAD;

        const double* u[3];
        double* out[3];
        const int64_t* __restrict__ aux;
        const double* __restrict__ tab;
        double scale;

        // acc[c] += (sg * flux_c(left cell at offset l, right cell at offset r)) * coef
        __device__ __forceinline__ void add(double* acc, int d, int64_t l, int64_t r, double sg, double coef) const
        {
            const double uld = u[d][l], urd = u[d][r];
            const bool up    = 0.5 * (uld + urd) >= 0;
#pragma unroll
            for (int c = 0; c < DIM; ++c)
            {
                double f = up ? uld * u[c][l] : urd * u[c][r];
                f        = scale != 1 ? f * scale : f;
                acc[c]   = acc[c] + (sg * f) * coef;
            }
        }

        __device__ __forceinline__ void side(double* acc, const smr_item_flux& it, int k, int d, int plus, int kind, double cs, double cj) const
        {
            const double sg = plus ? 1.0 : -1.0;
            const int64_t i = it.c + k;
            if (kind != SMR_FACE_FINE)
            {
                const int64_t n = d == 0 ? i + (plus ? 1 : -1) : it.nb[2 * (d - 1) + plus] + k;
                add(acc, d, plus ? i : n, plus ? n : i, sg, cs);
                return;
            }
            const int64_t* fx = aux + it.fine + (2 * d + plus) * 4;
            if (d == 0)
            {
#pragma unroll
                for (int r = 0; r < (1 << (DIM - 1)); ++r)
                {
                    add(acc, d, fx[r], fx[r] + 1, sg, cj);
                }
                return;
            }
#pragma unroll
            for (int b = 0; b < (DIM > 2 ? 2 : 1); ++b)
            {
                const int64_t r0 = fx[2 * b] + 2 * k, r1 = fx[2 * b + 1] + 2 * k;
                add(acc, d, r0, r1, sg, cj);
                add(acc, d, r0 + 1, r1 + 1, sg, cj);
            }
        }

        __device__ __forceinline__ void operator()(const smr_item_flux& it, int k) const
        {
            const double cs = tab[it.level * 6];
            const double cj = tab[(SMR_MAX_LEVELS + it.level) * 6];
            double acc[DIM];
#pragma unroll
            for (int c = 0; c < DIM; ++c)
            {
                acc[c] = 0.0;
            }
#pragma unroll
            for (int d = 0; d < DIM; ++d)
            {
                int km = (it.kinds >> (4 * d)) & 3, kp = (it.kinds >> (4 * d + 2)) & 3;
                if (d == 0)
                {
                    km = k != 0 ? SMR_FACE_SAME : km;
                    kp = k != it.n - 1 ? SMR_FACE_SAME : kp;
                }
                bool m_thr = ((it.kinds >> (SMR_FLUXW_SWAP_SHIFT + d)) & 1) != 0, p_thr = ((it.kinds >> (SMR_FLUX_PLUS_THROUGH_SHIFT + d)) & 1) != 0;
                if (d == 0)
                {
                    m_thr = m_thr && k == 0;
                    p_thr = p_thr && k == it.n - 1;
                }
                // pass numbers x2 as in FluxGenOp: minus {2, 4, 7, 9}, plus {2, 5, 6, 8}; through the periodic boundary: plus 3, then minus
                const int pm = km == SMR_FACE_SAME ? (m_thr ? 3 : 2) : (km == SMR_FACE_COARSE ? 4 : (km == SMR_FACE_FINE ? 7 : 9));
                const int pp = kp == SMR_FACE_SAME ? (p_thr ? 3 : 2) : (kp == SMR_FACE_COARSE ? 5 : (kp == SMR_FACE_FINE ? 6 : 8));
                if (pm <= pp && !(m_thr && p_thr))
                {
                    side(acc, it, k, d, 0, km, cs, cj);
                    side(acc, it, k, d, 1, kp, cs, cj);
                }
                else
                {
                    side(acc, it, k, d, 1, kp, cs, cj);
                    side(acc, it, k, d, 0, km, cs, cj);
                }
            }
#pragma unroll
            for (int c = 0; c < DIM; ++c)
            {
                mstore(out[c] + it.c + k, acc[c], static_cast<unsigned>(it.mask));
            }
        }
    };

    // WENO5 convection (Jiang & Shu, weno_impl.hpp:26-63) as non-linear flux schemes with the line stencil {-2 .. 3}, gather form as
    // above, on fully periodic meshes (a face on the periodic boundary reads the periodic ghosts; a same-level minus face through the
    // boundary comes after the plus face: swap bit, items.h).
    //   NONLIN = 0, NC = 1: make_convection_weno5<Field>(velocity)   (operators/convection_lin.hpp:95-178): f = velocity[d] * u
    //   NONLIN = 1:         make_convection_weno5<Field>()           (operators/convection_nonlin.hpp:162-233): f = u * u (NC = 1) or
    //                       u(d) * u (NC = DIM, SoA components), upwinded by the mean of the two cells next to the interface
    // `tab`: [0][l][0] = h_factor(h_l, h_l), [1][l][0] = h_factor(h_{l+1}, h_l).
    template <int DIM, int NONLIN = 0, int NC = 1>
    struct FluxWenoOp
    {
        static constexpr bool two_phase = false;
        static constexpr bool warp_uniform = false;
        static constexpr int min_blocks = 1;
        static constexpr int units_per_thread = SMR_CELLS_PER_THREAD;

        const double* u[3];
        double* out[3];
        const int64_t* __restrict__ aux;
        const double* __restrict__ tab;
        double vel[3];
        double scale; // `scale * scheme` (flux_based/algebraic_operators.hpp:38-46)

        // compute_weno5_flux of the five flux values; pow(x, 2) as x * x
        static __host__ __device__ __forceinline__ double weno5(double f0, double f1, double f2, double f3, double f4)
        {
            const double q0 = 1. / 3 * f0 - 7. / 6 * f1 + 11. / 6 * f2;
            const double q1 = -1. / 6 * f1 + 5. / 6 * f2 + 1. / 3 * f3;
            const double q2 = 1. / 3 * f2 + 5. / 6 * f3 - 1. / 6 * f4;
            const double a0 = f0 - 2 * f1 + f2, b0 = f0 - 4 * f1 + 3 * f2;
            const double a1 = f1 - 2 * f2 + f3, b1 = f1 - f3;
            const double a2 = f2 - 2 * f3 + f4, b2 = 3 * f2 - 4 * f3 + f4;
            const double IS0 = 13. / 12 * (a0 * a0) + 1. / 4 * (b0 * b0);
            const double IS1 = 13. / 12 * (a1 * a1) + 1. / 4 * (b1 * b1);
            const double IS2 = 13. / 12 * (a2 * a2) + 1. / 4 * (b2 * b2);
            const double eps = 1e-6;
            const double e0 = eps + IS0, e1 = eps + IS1, e2 = eps + IS2;
            const double al0 = 0.1 / (e0 * e0);
            const double al1 = 0.6 / (e1 * e1);
            const double al2 = 0.3 / (e2 * e2);
            const double sa  = al0 + al1 + al2;
            return (al0 / sa) * q0 + (al1 / sa) * q1 + (al2 / sa) * q2;
        }

        // acc[c] += (sg * flux_c) * coef for the interface whose six stencil cells sit at the storage offsets i[0..5]
        __host__ __device__ __forceinline__ void add(double* acc, int d, const int64_t* i, double sg, double coef) const
        {
            if (NONLIN == 0)
            {
                const double v  = vel[d];
                const double* p = u[0];
                double f        = v >= 0 ? weno5(p[i[0]] * v, p[i[1]] * v, p[i[2]] * v, p[i[3]] * v, p[i[4]] * v)
                                         : weno5(p[i[5]] * v, p[i[4]] * v, p[i[3]] * v, p[i[2]] * v, p[i[1]] * v);
                f               = scale != 1 ? f * scale : f;
                acc[0]          = acc[0] + (sg * f) * coef;
                return;
            }
            const double* pd = u[NC == 1 ? 0 : d];
            double ud[6];
#pragma unroll
            for (int k = 0; k < 6; ++k)
            {
                ud[k] = pd[i[k]];
            }
            const bool up = 0.5 * (ud[2] + ud[3]) >= 0;
#pragma unroll
            for (int c = 0; c < NC; ++c)
            {
                const double* pc = u[c];
                double g[6];
#pragma unroll
                for (int k = 0; k < 6; ++k)
                {
                    g[k] = ud[k] * pc[i[k]];
                }
                double f = up ? weno5(g[0], g[1], g[2], g[3], g[4]) : weno5(g[5], g[4], g[3], g[2], g[1]);
                f        = scale != 1 ? f * scale : f;
                acc[c]   = acc[c] + (sg * f) * coef;
            }
        }

        // storage offset of the own-level cell o steps along d from cell k of the record
        __host__ __device__ __forceinline__ int64_t off(const smr_item_fluxw& it, int k, int d, int o) const
        {
            if (d == 0 || o == 0)
            {
                return it.c + k + (d == 0 ? o : 0);
            }
            return it.nb[6 * (d - 1) + (o < 0 ? o + 3 : o + 2)] + k;
        }

        __host__ __device__ __forceinline__ void side(double* acc, const smr_item_fluxw& it, int k, int d, int plus, int kind, double cs,
                                                      double cj) const
        {
            const double sg = plus ? 1.0 : -1.0; // fluxes[1] = -fluxes[0] (flux_definition.hpp:125-136)
            int64_t i[6];
            if (kind != SMR_FACE_FINE)
            {
                const int o = plus ? 0 : -1; // stencil origin = the cell left of the interface
#pragma unroll
                for (int s = 0; s < 6; ++s)
                {
                    i[s] = off(it, k, d, o - 2 + s);
                }
                add(acc, d, i, sg, cs);
                return;
            }
            if (d == 0)
            {
                const int64_t* fx = aux + it.fine + plus * 4;
#pragma unroll
                for (int r = 0; r < (1 << (DIM - 1)); ++r)
                {
#pragma unroll
                    for (int s = 0; s < 6; ++s)
                    {
                        i[s] = fx[r] - 2 + s;
                    }
                    add(acc, d, i, sg, cj);
                }
                return;
            }
            const int64_t* fx = aux + it.fine + 8 + ((2 * d + plus) - 2) * 12;
#pragma unroll
            for (int b = 0; b < (DIM > 2 ? 2 : 1); ++b)
            {
                const int64_t* rows = fx + 6 * b;
#pragma unroll
                for (int x = 0; x < 2; ++x)
                {
#pragma unroll
                    for (int s = 0; s < 6; ++s)
                    {
                        i[s] = rows[s] + 2 * k + x;
                    }
                    add(acc, d, i, sg, cj);
                }
            }
        }

        // the values of cell k of the record (host-callable: smr_debug_fluxw_apply evaluates the records without a device in the CPU tests)
        __host__ __device__ __forceinline__ void compute(const smr_item_fluxw& it, int k, double* acc) const
        {
            const double cs = tab[it.level * 6];
            const double cj = tab[(SMR_MAX_LEVELS + it.level) * 6];
#pragma unroll
            for (int c = 0; c < NC; ++c)
            {
                acc[c] = 0.0;
            }
#pragma unroll
            for (int d = 0; d < DIM; ++d)
            {
                int km = (it.kinds >> (4 * d)) & 3, kp = (it.kinds >> (4 * d + 2)) & 3;
                bool swap = ((it.kinds >> (SMR_FLUXW_SWAP_SHIFT + d)) & 1) != 0;
                if (d == 0)
                {
                    swap = swap && k == 0;
                    km   = k != 0 ? SMR_FACE_SAME : km;
                    kp   = k != it.n - 1 ? SMR_FACE_SAME : kp;
                }
                // pass numbers x2 as in FluxGenOp; a same-level minus face through the periodic boundary: 3
                const int pm = km == SMR_FACE_SAME ? (swap ? 3 : 2) : (km == SMR_FACE_COARSE ? 4 : 7);
                const int pp = kp == SMR_FACE_SAME ? 2 : (kp == SMR_FACE_COARSE ? 5 : 6);
                if (pm <= pp)
                {
                    side(acc, it, k, d, 0, km, cs, cj);
                    side(acc, it, k, d, 1, kp, cs, cj);
                }
                else
                {
                    side(acc, it, k, d, 1, kp, cs, cj);
                    side(acc, it, k, d, 0, km, cs, cj);
                }
            }
        }

        __device__ __forceinline__ void operator()(const smr_item_fluxw& it, int k) const
        {
            double acc[NC];
            compute(it, k, acc);
#pragma unroll
            for (int c = 0; c < NC; ++c)
            {
                mstore(out[c] + it.c + k, acc[c], static_cast<unsigned>(it.mask));
            }
        }
    };

    // out = a * x + b * y on the leaves (the field-expression tail `u - dt * S(u)` is a = 1, b = -dt)
    struct LinCombOp
    {
        static constexpr bool two_phase = false;
        static constexpr bool warp_uniform = false;
        static constexpr int min_blocks = 1;
        static constexpr int units_per_thread = SMR_CELLS_PER_THREAD;

        const double* x;
        const double* y;
        double* out;
        double a, b;
        bool a_is_one;

        __device__ __forceinline__ void operator()(const smr_item_fv& it, int k) const
        {
            const int64_t i = it.c + k;
            const double v  = a_is_one ? x[i] + b * y[i] : a * x[i] + b * y[i];
            mstore(out + i, v, static_cast<unsigned>(it.mask));
        }
    };

    // ------------------------------------------------------------------------------------------------------------
    // projection (numeric/projection.hpp:22-64)
    // ------------------------------------------------------------------------------------------------------------
    template <int DIM>
    struct ProjOp
    {
        static constexpr bool two_phase = false;
        static constexpr bool warp_uniform = false;
        static constexpr int min_blocks = SMR_MB_PROJ;
        static constexpr int units_per_thread = SMR_CELLS_PER_THREAD;

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
            mstore(dst + it.dst + k, sum * (1.0 / static_cast<double>(1 << DIM)), static_cast<unsigned>(it.mask));
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
        static constexpr bool two_phase = false;
        static constexpr bool warp_uniform = false;
        static constexpr int min_blocks = 1;
        static constexpr int units_per_thread = SMR_CELLS_PER_THREAD;

        const double* src; // may alias dst (ghost update predicts inside one field)
        double* dst;

        __device__ __forceinline__ void operator()(const smr_item_pred& it, int k) const
        {
            const int ic = ((it.par & 1) + k) >> 1;
            if (RADIUS == 0)
            {
                mstore(dst + it.dst + k, src[it.src[4] + ic], static_cast<unsigned>(it.par) >> 8);
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
            mstore(dst + it.dst + k, val, static_cast<unsigned>(it.par) >> 8);
        }
    };

    // ------------------------------------------------------------------------------------------------------------
    // detail (mr/operators.hpp:146-175, 226-358, 360-533)
    // ------------------------------------------------------------------------------------------------------------
    template <int DIM, int RADIUS, bool RESTRICT = true>
    struct DetailOp
    {
        static constexpr bool two_phase = false;
        static constexpr bool warp_uniform = false;
        static constexpr int min_blocks = DIM > 2 ? SMR_MB_DETAIL3D : SMR_MB_DETAIL; // 3D radius 1 holds 16 accumulators + 27 coarse values (measured slower at level 9: a one-child-row-at-a-time variant with 64 registers, 933 vs 759 us; one unit per thread with four CTAs per chunk, 1029 us)
        static constexpr int units_per_thread = SMR_CELLS_PER_THREAD;

        typename PtrOf<const double, RESTRICT>::type f;
        typename PtrOf<double, RESTRICT>::type detail;

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
                mstore(o, d[r][0], static_cast<unsigned>(it.mask));
                mstore(o + 1, d[r][1], static_cast<unsigned>(it.mask));
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

    template <int DIM, bool RESTRICT = true>
    struct CriteriaOp
    {
        static constexpr bool two_phase = true; // all detail loads of a thread's units are issued before its first tag store
        static constexpr bool warp_uniform = false;
        static constexpr int min_blocks = DIM > 2 ? SMR_MB_CRIT3D : SMR_MB_CRITERIA;
        static constexpr int units_per_thread = SMR_CELLS_PER_THREAD;

        typename PtrOf<const double, RESTRICT>::type detail; // ncomp arrays of `stride` entries (one per adapted field)
        typename PtrOf<uint8_t, RESTRICT>::type tag;
        TagParams p;
        int ncomp;
        int64_t stride;

        struct Vals
        {
            unsigned bits; // bit 0: the sibling group may coarsen; bit 1 + 2 r + x: child (r, x) must refine
        };

        __device__ __forceinline__ Vals load(const smr_item_tag& it, int k) const
        {
            constexpr int NR = 1 << (DIM - 1);
            const int fl    = it.level & 0xff;
            bool coarsen_ok = fl > p.min_level;
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
            Vals out{coarsen_ok ? 1u : 0u};
#pragma unroll
            for (int r = 0; r < NR; ++r)
            {
#pragma unroll
                for (int x = 0; x < 2; ++x)
                {
                    out.bits |= refine[r][x] ? (2u << (2 * r + x)) : 0u;
                }
            }
            return out;
        }

        __device__ __forceinline__ void operator()(const smr_item_tag& it, int k) const
        {
            finish(it, k, load(it, k));
        }

        __device__ __forceinline__ void finish(const smr_item_tag& it, int k, const Vals& v) const
        {
            constexpr int NR = 1 << (DIM - 1);
            const int fl          = it.level & 0xff;
            const unsigned mask   = static_cast<unsigned>(it.level) >> 8;
            const bool coarsen_ok = (v.bits & 1u) != 0u;
            if (coarsen_ok)
            {
#pragma unroll
                for (int r = 0; r < NR; ++r)
                {
                    mstore(tag + it.fine[r] + 2 * k, static_cast<uint8_t>(2), mask);
                    mstore(tag + it.fine[r] + 2 * k + 1, static_cast<uint8_t>(2), mask);
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
                        if (v.bits & (2u << (2 * r + x)))
                        {
                            uint8_t* t = tag + it.fine[r] + 2 * k + x;
                            // the coarsen store above (if any) went to the same peers, so the local value is theirs too
                            mstore(t, static_cast<uint8_t>((coarsen_ok ? 2 : *t) | 4), mask);
                        }
                    }
                }
            }
        }
    };

    template <int DIM, bool RESTRICT = true>
    struct MaximumOp
    {
        static constexpr bool two_phase = true; // the tag loads of a thread's units are issued before its first store
        static constexpr bool warp_uniform = false;
        static constexpr int min_blocks = DIM > 2 ? SMR_MB_MAX3D : SMR_MB_MAXIMUM;
        static constexpr int units_per_thread = SMR_CELLS_PER_THREAD;

        typename PtrOf<uint8_t, RESTRICT>::type tag;

        // Every tag this unit touches (its 2^dim children and their parent) is written by this unit alone during the sweep of
        // one fine level, so the values may be read up front and a store that would not change the byte may be dropped.
        struct Vals
        {
            uint8_t t[1 << (DIM - 1)][2];
            uint8_t tc;
        };

        __device__ __forceinline__ Vals load(const smr_item_tag& it, int k) const
        {
            constexpr int NR = 1 << (DIM - 1);
            Vals v;
#pragma unroll
            for (int r = 0; r < NR; ++r)
            {
                v.t[r][0] = tag[it.fine[r] + 2 * k];
                v.t[r][1] = tag[it.fine[r] + 2 * k + 1];
            }
            v.tc = tag[it.coarse + k];
            return v;
        }

        __device__ __forceinline__ void put(uint8_t* where, uint8_t old_value, uint8_t new_value, unsigned mask) const
        {
            if (new_value != old_value || mask != 0u)
            {
                mstore(where, new_value, mask);
            }
        }

        __device__ __forceinline__ void finish(const smr_item_tag& it, int k, const Vals& v) const
        {
            constexpr int NR    = 1 << (DIM - 1);
            const unsigned mask = static_cast<unsigned>(it.level) >> 8;
            uint8_t any = 0, all = 0xff;
#pragma unroll
            for (int r = 0; r < NR; ++r)
            {
                any |= v.t[r][0] | v.t[r][1];
                all &= v.t[r][0] & v.t[r][1];
            }
            if (any & 1)
            {
#pragma unroll
                for (int r = 0; r < NR; ++r)
                {
                    put(tag + it.fine[r] + 2 * k, v.t[r][0], static_cast<uint8_t>(v.t[r][0] | 1), mask);
                    put(tag + it.fine[r] + 2 * k + 1, v.t[r][1], static_cast<uint8_t>(v.t[r][1] | 1), mask);
                }
                put(tag + it.coarse + k, v.tc, static_cast<uint8_t>(v.tc | 1), mask);
            }
            else if (all & 2)
            {
                put(tag + it.coarse + k, v.tc, static_cast<uint8_t>(v.tc | 1), mask);
            }
            else
            {
#pragma unroll
                for (int r = 0; r < NR; ++r)
                {
                    put(tag + it.fine[r] + 2 * k, v.t[r][0], static_cast<uint8_t>(v.t[r][0] & ~2), mask);
                    put(tag + it.fine[r] + 2 * k + 1, v.t[r][1], static_cast<uint8_t>(v.t[r][1] & ~2), mask);
                }
            }
        }

        __device__ __forceinline__ void operator()(const smr_item_tag& it, int k) const
        {
            finish(it, k, load(it, k));
        }
    };

    // ------------------------------------------------------------------------------------------------------------
    // relative detail (mr/rel_detail.hpp:73-112): max over the leaves of |f| (max is order independent, so the parallel
    // reduction is bit-identical to the reference's loop), then detail *= 1/max over the whole array.
    // Non-negative doubles order like their bit patterns, so the reduction is an integer atomicMax.
    // ------------------------------------------------------------------------------------------------------------
    struct AbsMaxOp
    {
        static constexpr bool two_phase = false;
        static constexpr bool warp_uniform = false;
        static constexpr int min_blocks = 1;
        static constexpr int units_per_thread = SMR_CELLS_PER_THREAD;

        const double* __restrict__ f;
        unsigned long long* slot; // this rank's slot

        __device__ __forceinline__ void operator()(const smr_item_fv& it, int k) const
        {
            const unsigned long long v = static_cast<unsigned long long>(__double_as_longlong(fabs(f[it.c + k])));
            if (v > *reinterpret_cast<volatile unsigned long long*>(slot))
            {
                atomicMax(slot, v);
            }
        }
    };

#ifdef SMR_MISC_KERNELS
    // publish this rank's slot into every peer's table (multi-GPU), one thread per peer
    __global__ void publish_slot_kernel(unsigned long long* slots)
    {
        const int p = threadIdx.x;
        if (p < g_peers.world && p != g_peers.rank)
        {
            unsigned long long* remote = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(slots) + g_peers.delta[p]);
            remote[g_peers.rank]       = slots[g_peers.rank];
        }
    }

    __global__ void __launch_bounds__(SMR_CTA_THREADS) scale_detail_kernel(double* __restrict__ detail, int64_t n, const unsigned long long* slots, int world)
    {
        unsigned long long mbits = static_cast<unsigned long long>(__double_as_longlong(2.2250738585072014e-308)); // DBL_MIN
        for (int r = 0; r < world; ++r)
        {
            mbits = slots[r] > mbits ? slots[r] : mbits;
        }
        double m = __longlong_as_double(static_cast<long long>(mbits));
        if (m < 2.220446049250313e-16) // DBL_EPSILON
        {
            m = 1.0;
        }
        const double inv = 1. / m;
        for (int64_t i = static_cast<int64_t>(blockIdx.x) * SMR_CTA_THREADS + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * SMR_CTA_THREADS)
        {
            detail[i] *= inv;
        }
    }
#endif

    // tag[leaf] = keep (mr/adapt.hpp:286-290), driven by the FV leaf batch
    template <bool RESTRICT = true>
    struct KeepLeavesOpT
    {
        static constexpr bool two_phase = false;
        static constexpr bool warp_uniform = false;
        static constexpr int min_blocks = 1;
        static constexpr int units_per_thread = SMR_CELLS_PER_THREAD;

        typename PtrOf<uint8_t, RESTRICT>::type tag;
        unsigned mask_all; // tags are replicated on every rank

        __device__ __forceinline__ void operator()(const smr_item_fv& it, int k) const
        {
            mstore(tag + it.c + k, static_cast<uint8_t>(1), mask_all);
        }

        // long leaf intervals: a memset of `count` bytes with 16-byte stores (one byte store per thread was measured at
        // 0.1 TB/s on the uniform level-13 mesh: partial-sector writes)
        static constexpr bool has_span = true;
        __device__ __forceinline__ bool span(const smr_item_fv& it, int k0, int count) const
        {
            if (mask_all != 0u)
            {
                return false;
            }
            uint8_t* p     = tag + it.c + k0;
            const int head = min(count, static_cast<int>((16u - (static_cast<unsigned>(reinterpret_cast<uintptr_t>(p)) & 15u)) & 15u));
            const int nvec = (count - head) >> 4;
            const int tail = count - head - (nvec << 4);
            const int t    = threadIdx.x;
            if (t < head)
            {
                p[t] = 1;
            }
            uint4* v = reinterpret_cast<uint4*>(p + head);
            for (int i = t; i < nvec; i += SMR_CTA_THREADS)
            {
                v[i] = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
            }
            if (t < tail)
            {
                p[head + (nvec << 4) + t] = 1;
            }
            return true;
        }

        // the whole record by one warp (record_kernel)
        __device__ __forceinline__ void record(const smr_item_fv& it, int lane) const
        {
            uint8_t* p     = tag + it.c;
            const int head = min(it.n, static_cast<int>((16u - (static_cast<unsigned>(reinterpret_cast<uintptr_t>(p)) & 15u)) & 15u));
            const int nvec = (it.n - head) >> 4;
            const int tail = it.n - head - (nvec << 4);
            if (lane < head)
            {
                p[lane] = 1;
            }
            uint4* v = reinterpret_cast<uint4*>(p + head);
            for (int i = lane; i < nvec; i += 32)
            {
                v[i] = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
            }
            if (lane < tail)
            {
                p[head + (nvec << 4) + lane] = 1;
            }
        }
    };

    using KeepLeavesOp = KeepLeavesOpT<true>;

    // does update_cell_array_from_tag change anything? (algorithm/graduation.hpp:743-842: a leaf is replaced when it is
    // tagged refine below max_level, or coarsen without keep above min_level).  Lets the host skip its scan of the tags
    // on the iteration that finds the mesh converged.
    struct TagsChangeOp
    {
        static constexpr bool two_phase = false;
        static constexpr bool warp_uniform = false;
        static constexpr int min_blocks = 1;
        static constexpr int units_per_thread = SMR_CELLS_PER_THREAD;

        const uint8_t* tag;
        unsigned* flag;
        int min_level, max_level;
        unsigned mask_all = 0; // multi-GPU: every rank scans its own leaves and raises the flag on all ranks

        __device__ __forceinline__ void operator()(const smr_item_fv& it, int k) const
        {
            const uint8_t t = tag[it.c + k];
            if (((t & 4) && it.level < max_level) || ((t & 2) && !(t & 1) && it.level > min_level))
            {
                mstore(flag, 1u, mask_all); // benign race: every writer stores the same value
            }
        }

        // long leaf intervals: 16 tags per load, the two conditions evaluated on packed bytes
        static constexpr bool has_span = true;
        __device__ __forceinline__ bool span(const smr_item_fv& it, int k0, int count) const
        {
            const uint8_t* p = tag + it.c + k0;
            const int head   = min(count, static_cast<int>((16u - (static_cast<unsigned>(reinterpret_cast<uintptr_t>(p)) & 15u)) & 15u));
            const int nvec   = (count - head) >> 4;
            const int tail   = count - head - (nvec << 4);
            const int t      = threadIdx.x;
            const unsigned refine_on  = it.level < max_level ? 0x04040404u : 0u;
            const unsigned coarsen_on = it.level > min_level ? 0x02020202u : 0u;
            unsigned hit = 0;
            auto test = [&](unsigned w) { hit |= (w & refine_on) | (w & coarsen_on & ~((w & 0x01010101u) << 1)); };
            if (t < head)
            {
                test(p[t]);
            }
            const uint4* v = reinterpret_cast<const uint4*>(p + head);
            for (int i = t; i < nvec; i += SMR_CTA_THREADS)
            {
                const uint4 w = v[i];
                test(w.x);
                test(w.y);
                test(w.z);
                test(w.w);
            }
            if (t < tail)
            {
                test(p[head + (nvec << 4) + t]);
            }
            if (hit != 0u)
            {
                mstore(flag, 1u, mask_all);
            }
            return true;
        }

        // the whole record by one warp (record_kernel)
        __device__ __forceinline__ void record(const smr_item_fv& it, int lane) const
        {
            const uint8_t* p = tag + it.c;
            const int head   = min(it.n, static_cast<int>((16u - (static_cast<unsigned>(reinterpret_cast<uintptr_t>(p)) & 15u)) & 15u));
            const int nvec   = (it.n - head) >> 4;
            const int tail   = it.n - head - (nvec << 4);
            const unsigned refine_on  = it.level < max_level ? 0x04040404u : 0u;
            const unsigned coarsen_on = it.level > min_level ? 0x02020202u : 0u;
            unsigned hit = 0;
            auto test = [&](unsigned w) { hit |= (w & refine_on) | (w & coarsen_on & ~((w & 0x01010101u) << 1)); };
            if (lane < head)
            {
                test(p[lane]);
            }
            const uint4* v = reinterpret_cast<const uint4*>(p + head);
#pragma unroll 4
            for (int i = lane; i < nvec; i += 32)
            {
                const uint4 w = v[i];
                test(w.x);
                test(w.y);
                test(w.z);
                test(w.w);
            }
            if (lane < tail)
            {
                test(p[head + (nvec << 4) + lane]);
            }
            if (hit != 0u)
            {
                mstore(flag, 1u, mask_all);
            }
        }
    };

    // u[cell] = (|center - c|^2 <= r^2) ? inside : outside   -- the demos' initial condition
    // (demos/FiniteVolume/advection_2d.cpp:23-45, advection_3d.cpp:32-45, scalar_burgers_2d.cpp:20-50)
    template <int DIM>
    struct InitBallOp
    {
        static constexpr bool two_phase = false;
        static constexpr bool warp_uniform = false;
        static constexpr int min_blocks = 1;
        static constexpr int units_per_thread = SMR_CELLS_PER_THREAD;

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
                mstore(u + it.c + k, inside, static_cast<unsigned>(it.mask));
            }
            else if (overwrite_outside)
            {
                mstore(u + it.c + k, outside, static_cast<unsigned>(it.mask));
            }
        }
    };

    template <bool RESTRICT = true>
    struct CopyOpT
    {
        static constexpr bool two_phase = false;
        static constexpr bool warp_uniform = false;
        static constexpr int min_blocks = SMR_MB_COPY;
        static constexpr int units_per_thread = SMR_CELLS_PER_THREAD;

        typename PtrOf<const double, RESTRICT>::type src;
        typename PtrOf<double, RESTRICT>::type dst;

        __device__ __forceinline__ void operator()(const smr_item_copy& it, int k) const
        {
            mstore(dst + it.dst + k, src[it.src + k], static_cast<unsigned>(it.mask));
        }
    };

    using CopyOp = CopyOpT<true>;

    // update_tag_periodic (algorithm/update_periodic.hpp:211-300): a periodic ghost and the cell it mirrors end up with the OR
    // of their tags; driven by the same records as the periodic ghost copy of the fields (dst = ghost, src = mirrored cell)
    struct TagOrOp
    {
        static constexpr bool two_phase = false;
        static constexpr bool warp_uniform = false;
        static constexpr int min_blocks = 1;
        static constexpr int units_per_thread = SMR_CELLS_PER_THREAD;

        uint8_t* tag;
        unsigned mask_all;

        __device__ __forceinline__ void operator()(const smr_item_copy& it, int k) const
        {
            const uint8_t v = static_cast<uint8_t>(tag[it.dst + k] | tag[it.src + k]);
            mstore(tag + it.dst + k, v, mask_all);
            mstore(tag + it.src + k, v, mask_all);
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

    __device__ __forceinline__ void run_bc(const BcView& v, double* f, int i)
    {
        if (i >= v.n)
        {
            return;
        }
        const smr_item_bc it = v.items[i];
        const int64_t* s     = v.srcs + it.src_first;
        const int kind       = it.kind & 0xff;
        const unsigned mask  = static_cast<unsigned>(it.kind) >> 8;
        if (kind == SMR_BC_COPY)
        {
            mstore(f + it.dst, f[s[0]], mask);
        }
        else if (kind == SMR_BC_VALUE)
        {
            if (v.bc_type == SMR_BCTYPE_DIRICHLET)
            {
                mstore(f + it.dst, 2 * v.bc_value - f[s[0]], mask);
            }
            else
            {
                mstore(f + it.dst, it.coef * v.bc_value + f[s[0]], mask);
            }
        }
        else if (kind == SMR_BC_EXTRAP4)
        {
            // bc/polynomial_extrapolation.hpp:67-70: u[cells[0]] - u[cells[1]] * 3.0 + u[cells[2]] * 3.0
            mstore(f + it.dst, f[s[0]] - f[s[1]] * 3.0 + f[s[2]] * 3.0, mask);
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
            mstore(f + it.dst, sum, mask);
        }
    }

    // One level of the top-down ghost sweep in a single launch: CTAs [0, bc_ctas) fill the outside ghosts of the level,
    // the remaining CTAs run the projection level -> level-1.  The two halves never touch the same cells (batches.hpp).
    template <int DIM>
    __global__ void __launch_bounds__(SMR_CTA_THREADS) ghost_phase_kernel(BcView bc, int bc_ctas, BatchView<smr_item_proj> pv, double* __restrict__ f)
    {
        __shared__ int32_t s_prefix[SMR_CTA_CELLS + 2];
        if (static_cast<int>(blockIdx.x) < bc_ctas)
        {
            run_bc(bc, f, blockIdx.x * SMR_CTA_THREADS + threadIdx.x);
        }
        else
        {
            run_batch_cta(pv, ProjOp<DIM>{f, f}, static_cast<int>(blockIdx.x) - bc_ctas, s_prefix);
        }
    }

    // ------------------------------------------------------------------------------------------------------------
    // Fused level wavefront.  An adapted mesh of ~10^6 cells spreads a ghost update over ~2 (L+1) dependent sweeps of a
    // few thousand cells each, and a harten iteration adds detail, criteria and L keep-propagation sweeps: as separate
    // launches the step is bound by launch latency (profiles/r01_summary.md).  Here ONE cooperative launch walks a
    // table of phases; a phase is a set of independent jobs (each job = one batch), the CTAs of the grid share the
    // CTA-chunks of all jobs of the phase, and a grid-wide barrier separates phases.  Per-cell arithmetic is the ops'
    // own operator(): results are bit-identical to the launch-per-sweep path.
    // ------------------------------------------------------------------------------------------------------------
    enum
    {
        WF_BC = 0,
        WF_PROJ,
        WF_PRED,
        WF_DETAIL,
        WF_CRITERIA,
        WF_MAXIMUM,
        WF_KEEP,
        WF_ZERO_DETAIL,
        WF_ZERO_TAG,
        WF_COPY,
        WF_TAGS_CHANGE,
        WF_TAG_OR
    };

    struct WfJob
    {
        int32_t op;
        int32_t n_ctas;
        int64_t items, prefix, cta_first, aux; // byte offsets into the arena
        int64_t n_cells;                       // output units (bc: records; zero: bytes)
        int32_t field;                         // index into WfArgs::dst / src (detail: component)
        int32_t pad;                           // 1: work items are full chunks, 0: quarter chunks
    };

    struct WfPhase
    {
        int32_t first_job, n_jobs, total_ctas, pad;
    };

    #define SMR_WF_MAX_FIELDS 8
    #define SMR_WF_ERROR_WORD 16 /* word index (from the barrier counter) of the launch error word */
    #define SMR_WF_ZERO_BYTES 16384 /* bytes cleared by one CTA-chunk of a WF_ZERO_* job */

    struct WfArgs
    {
        const char* arena;
        const WfPhase* phases;
        const WfJob* jobs;
        int n_phases;
        int n_jobs;
        int ncomp;
        const double* src[SMR_WF_MAX_FIELDS]; // == dst for the ghost update, the old field for update_fields
        double* dst[SMR_WF_MAX_FIELDS];
        int bc_type[SMR_WF_MAX_FIELDS];
        double bc_value[SMR_WF_MAX_FIELDS];
        double* detail;
        uint8_t* tag;
        unsigned* change_flag; // set by WF_TAGS_CHANGE (lives behind the tags, cleared by WF_ZERO_TAG)
        int64_t n; // reference cells (detail stride)
        unsigned mask_all;
        unsigned* barrier;      // grid barrier counter (monotonic across launches)
        unsigned barrier_base;  // its value when this launch starts
        // multi-GPU: every phase boundary is also a barrier across the ranks (all run the same phase list).  CTA 0 exchanges
        // flags with the peers once the local grid has arrived, then releases the local grid through counter[1].
        int mg_world;                      // 1: single GPU
        unsigned release_base;             // value of counter[1] when this launch starts
        unsigned long long mg_epoch_base;  // flag value of this launch's k-th barrier is mg_epoch_base + k
        unsigned long long* trace; // optional: CTA 0 stamps %globaltimer at the start of every phase and at the end
        TagParams tp;
    };

    __device__ __forceinline__ unsigned long long wf_now()
    {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        return t;
    }

    // Grid-wide barrier of the cooperative launch (all CTAs co-resident): one atomic arrival per CTA on a monotonic
    // counter, thread 0 spins until the whole grid has arrived.  Lighter than cooperative_groups' grid.sync().
    // Returns false when the wait timed out: the caller raises the launch's error word (counter[SMR_WF_ERROR_WORD]) and
    // every CTA leaves the kernel instead of running on unsynchronised.
    __device__ __forceinline__ bool wf_grid_barrier(unsigned* counter, unsigned target)
    {
        __shared__ int s_ok;
        __syncthreads();
        if (threadIdx.x == 0)
        {
            __threadfence();
            atomicAdd(counter, 1u);
            // bounded: a launch whose CTAs were not co-resident (it is launched cooperatively, so this cannot happen) or a
            // host/device disagreement on the barrier count must not hang the GPU
            for (long long spins = 0; static_cast<int>(*reinterpret_cast<volatile unsigned*>(counter) - target) < 0; ++spins)
            {
                if (spins > (1LL << 28) || *reinterpret_cast<volatile unsigned*>(counter + SMR_WF_ERROR_WORD) != 0u)
                {
                    atomicExch(counter + SMR_WF_ERROR_WORD, target | 1u); // sticky: the host throws at its next synchronisation
                    break;
                }
            }
            s_ok = *reinterpret_cast<volatile unsigned*>(counter + SMR_WF_ERROR_WORD) == 0u;
            __threadfence();
        }
        __syncthreads();
        return s_ok != 0;
    }

    // Phase boundary across GPUs: local arrival as above; CTA 0 then tells every peer "my rank has finished barrier k" by storing
    // the epoch into the peer's flag table over NVLink, waits for the same word from every peer, and releases the local grid.
    // Thread 0 of every CTA releases at device scope after the CTA barrier and CTA 0 fences at system scope before sending the flags, so
    // the halo values the grid stored into the peers' pools are visible there before the flag is.  All waits are bounded: a rank that stopped raises the error words instead of hanging the GPUs.
    __device__ __forceinline__ bool wf_mg_barrier(const WfArgs& a, unsigned k)
    {
        __shared__ int s_ok_mg;
        __syncthreads();
        if (threadIdx.x == 0)
        {
            // Fences are cumulative (PTX memory model): the CTA barrier orders every thread's peer stores before this device-scope
            // release, CTA 0 acquires all arrivals and issues ONE system-scope fence before it sends the flags, so the stores of the
            // whole grid are visible to the peers before the flag is.  A system-scope fence costs ~4 us on this path (measured: five of
            // them in sequence made a phase boundary 21 us); the barrier now has two on its critical path, both in CTA 0.
            __threadfence();
            unsigned* counter     = a.barrier;
            const unsigned target = a.barrier_base + k * gridDim.x;
            const unsigned rel    = a.release_base + k;
            volatile unsigned* vc = reinterpret_cast<volatile unsigned*>(counter);
            atomicAdd(counter, 1u);
            bool ok = true;
            if (blockIdx.x == 0)
            {
                for (long long spins = 0; static_cast<int>(vc[0] - target) < 0; ++spins)
                {
                    if (spins > (1LL << 28) || vc[SMR_WF_ERROR_WORD] != 0u)
                    {
                        ok = false;
                        break;
                    }
                }
                const unsigned long long epoch = a.mg_epoch_base + k;
                __threadfence_system();
                for (int p = 0; ok && p < g_peers.world; ++p)
                {
                    if (p != g_peers.rank)
                    {
                        volatile unsigned long long* remote = reinterpret_cast<volatile unsigned long long*>(
                            reinterpret_cast<char*>(g_peers.flags) + g_peers.delta[p]);
                        remote[g_peers.rank] = epoch;
                    }
                }
                volatile unsigned long long* mine = g_peers.flags;
                for (int p = 0; ok && p < g_peers.world; ++p)
                {
                    if (p == g_peers.rank)
                    {
                        continue;
                    }
                    for (long long spins = 0; mine[p] < epoch; ++spins)
                    {
                        if (spins > 400000000LL) // ~ a second
                        {
                            *g_peers.error = epoch;
                            ok             = false;
                            break;
                        }
                    }
                }
                if (!ok)
                {
                    atomicExch(counter + SMR_WF_ERROR_WORD, target | 1u);
                }
                __threadfence_system(); // acquire: the peers' halo stores (made visible before their flags) are visible to this grid
                atomicExch(counter + 1, rel); // release the local grid (also on failure: the others then see the error word)
            }
            else
            {
                for (long long spins = 0; static_cast<int>(vc[1] - rel) < 0; ++spins)
                {
                    if (spins > (1LL << 30) || vc[SMR_WF_ERROR_WORD] != 0u)
                    {
                        atomicExch(counter + SMR_WF_ERROR_WORD, target | 1u);
                        break;
                    }
                }
            }
            s_ok_mg = vc[SMR_WF_ERROR_WORD] == 0u;
            __threadfence(); // the halo values live in this GPU's memory: a device-scope acquire after CTA 0's system-scope one
        }
        __syncthreads();
        return s_ok_mg != 0;
    }

    template <class Item>
    __device__ __forceinline__ BatchView<Item> wf_view(const char* arena, const WfJob& jb)
    {
        return BatchView<Item>{reinterpret_cast<const Item*>(arena + jb.items), reinterpret_cast<const int64_t*>(arena + jb.prefix),
                               reinterpret_cast<const int32_t*>(arena + jb.cta_first), jb.n_cells};
    }

    __device__ __forceinline__ void wf_zero(void* base, int64_t bytes, int chunk)
    {
        const int64_t lo = static_cast<int64_t>(chunk) * SMR_WF_ZERO_BYTES;
        const int64_t hi = lo + SMR_WF_ZERO_BYTES < bytes ? lo + SMR_WF_ZERO_BYTES : bytes;
        char* p          = static_cast<char*>(base);
        // buffers are 256-byte aligned and chunks are multiples of 16 bytes
        for (int64_t o = lo + 16 * static_cast<int64_t>(threadIdx.x); o + 16 <= hi; o += 16 * SMR_CTA_THREADS)
        {
            *reinterpret_cast<uint4*>(p + o) = make_uint4(0, 0, 0, 0);
        }
        const int64_t tail = lo + ((hi - lo) & ~int64_t(15));
        if (tail + static_cast<int64_t>(threadIdx.x) < hi && threadIdx.x < 16)
        {
            p[tail + threadIdx.x] = 0;
        }
    }

    // one work item of a batch job: a full chunk (`wide`: phases with far more work than CTAs, four units per thread keep
    // more loads in flight) or a quarter chunk (latency-bound phases: one unit per thread, four CTAs share the chunk)
    template <class Item, class Op>
    __device__ __forceinline__ void wf_item(const BatchView<Item>& b, const Op& op, int local, bool wide, int32_t* s_prefix)
    {
        if (wide)
        {
            run_batch_cta(b, op, local, s_prefix);
        }
        else
        {
            run_batch_sub(b, op, local >> 2, local & 3, s_prefix);
        }
    }

    template <int DIM, int RADIUS>
    __device__ __forceinline__ void wf_chunk(const WfArgs& a, const WfJob& jb, int local, int32_t* s_prefix)
    {
        switch (jb.op)
        {
            case WF_BC:
            {
                const BcView bc{reinterpret_cast<const smr_item_bc*>(a.arena + jb.items), reinterpret_cast<const int64_t*>(a.arena + jb.aux),
                                static_cast<int>(jb.n_cells), a.bc_type[jb.field], a.bc_value[jb.field]};
                run_bc(bc, a.dst[jb.field], local * SMR_CTA_THREADS + threadIdx.x);
                break;
            }
            case WF_PROJ:
                wf_item(wf_view<smr_item_proj>(a.arena, jb), ProjOp<DIM>{a.src[jb.field], a.dst[jb.field]}, local, jb.pad != 0, s_prefix);
                break;
            case WF_PRED:
                wf_item(wf_view<smr_item_pred>(a.arena, jb), PredOp<DIM, RADIUS>{a.src[jb.field], a.dst[jb.field]}, local, jb.pad != 0, s_prefix);
                break;
            case WF_DETAIL:
                wf_item(wf_view<smr_item_detail>(a.arena, jb), DetailOp<DIM, RADIUS, false>{a.dst[jb.field], a.detail + jb.field * a.n}, local, jb.pad != 0, s_prefix);
                break;
            case WF_CRITERIA:
                wf_item(wf_view<smr_item_tag>(a.arena, jb), CriteriaOp<DIM, false>{a.detail, a.tag, a.tp, a.ncomp, a.n}, local, jb.pad != 0, s_prefix);
                break;
            case WF_MAXIMUM:
                wf_item(wf_view<smr_item_tag>(a.arena, jb), MaximumOp<DIM, false>{a.tag}, local, jb.pad != 0, s_prefix);
                break;
            case WF_KEEP:
                wf_item(wf_view<smr_item_fv>(a.arena, jb), KeepLeavesOpT<false>{a.tag, a.mask_all}, local, jb.pad != 0, s_prefix);
                break;
            case WF_ZERO_DETAIL:
                wf_zero(a.detail, jb.n_cells, local);
                break;
            case WF_ZERO_TAG:
                wf_zero(a.tag, jb.n_cells, local);
                break;
            case WF_TAGS_CHANGE:
                wf_item(wf_view<smr_item_fv>(a.arena, jb), TagsChangeOp{a.tag, a.change_flag, a.tp.min_level, a.tp.max_level, a.mask_all}, local, jb.pad != 0, s_prefix);
                break;
            case WF_TAG_OR:
                wf_item(wf_view<smr_item_copy>(a.arena, jb), TagOrOp{a.tag, a.mask_all}, local, jb.pad != 0, s_prefix);
                break;
            default: // WF_COPY
                wf_item(wf_view<smr_item_copy>(a.arena, jb), CopyOpT<false>{a.src[jb.field], a.dst[jb.field]}, local, jb.pad != 0, s_prefix);
                break;
        }
    }

    __device__ __forceinline__ void wf_prefetch_range(const char* base, int64_t lo, int64_t hi)
    {
        lo &= ~int64_t(127);
        hi = hi - lo > 32768 ? lo + 32768 : hi; // the head of a long record list is enough to start the chunk
        for (int64_t o = lo + 128 * static_cast<int64_t>(threadIdx.x); o < hi; o += 128 * SMR_CTA_THREADS)
        {
            asm volatile("prefetch.global.L1 [%0];" ::"l"(base + o));
        }
    }

    // Pull the index data (CTA table, prefix slice, records) of chunk `w` of phase `ph` into L1 while the CTA is about to
    // wait at the barrier: after the barrier the chunk starts with cache hits instead of a chain of four L2 round trips.
    // Only the read-only arena is touched, never field data (which the phases in flight are still writing).
    __device__ __forceinline__ void wf_prefetch(const WfArgs& a, const WfJob* jobs, const WfPhase& ph, int w)
    {
        if (w >= ph.total_ctas)
        {
            return;
        }
        int j = ph.first_job, local = w;
        while (local >= jobs[j].n_ctas)
        {
            local -= jobs[j].n_ctas;
            ++j;
        }
        const WfJob& jb = jobs[j];
        int isz;
        switch (jb.op)
        {
            case WF_BC:
                wf_prefetch_range(a.arena + jb.items, static_cast<int64_t>(local) * SMR_CTA_THREADS * sizeof(smr_item_bc),
                                  static_cast<int64_t>(local + 1) * SMR_CTA_THREADS * sizeof(smr_item_bc));
                return;
            case WF_PROJ:
                isz = sizeof(smr_item_proj);
                break;
            case WF_PRED:
                isz = sizeof(smr_item_pred);
                break;
            case WF_DETAIL:
                isz = sizeof(smr_item_detail);
                break;
            case WF_CRITERIA:
            case WF_MAXIMUM:
                isz = sizeof(smr_item_tag);
                break;
            case WF_KEEP:
            case WF_TAGS_CHANGE:
                isz = sizeof(smr_item_fv);
                break;
            case WF_COPY:
            case WF_TAG_OR:
                isz = sizeof(smr_item_copy);
                break;
            default:
                return;
        }
        const int32_t* cf = reinterpret_cast<const int32_t*>(a.arena + jb.cta_first);
        const int chunk = jb.pad != 0 ? local : (local >> 2);
        const int first = cf[chunk], last = cf[chunk + 1];
        wf_prefetch_range(a.arena + jb.prefix, static_cast<int64_t>(first) * 8, static_cast<int64_t>(last + 2) * 8);
        wf_prefetch_range(a.arena + jb.items, static_cast<int64_t>(first) * isz, static_cast<int64_t>(last + 1) * isz);
    }

    // Dynamic shared memory: the phase and job tables (loaded once).  A phase flagged `serial` (a handful of chunks: the
    // coarse levels) is run by CTA 0 alone and needs no grid barrier before the next serial phase, only a CTA barrier.
    template <int DIM, int RADIUS>
    __global__ void __launch_bounds__(SMR_CTA_THREADS, 2) wavefront_kernel(WfArgs a)
    {
        __shared__ int32_t s_prefix[SMR_CTA_CELLS + 2];
        extern __shared__ __align__(16) unsigned char s_tab[];
        unsigned barriers = 0;
        {
            // phases and jobs are contiguous in the staging slot
            const int words        = (a.n_phases * static_cast<int>(sizeof(WfPhase)) + a.n_jobs * static_cast<int>(sizeof(WfJob))) / 4;
            const uint32_t* src    = reinterpret_cast<const uint32_t*>(a.phases);
            uint32_t* dst          = reinterpret_cast<uint32_t*>(s_tab);
            for (int i = threadIdx.x; i < words; i += SMR_CTA_THREADS)
            {
                dst[i] = src[i];
            }
            __syncthreads();
        }
        const WfPhase* phases = reinterpret_cast<const WfPhase*>(s_tab);
        const WfJob* jobs     = reinterpret_cast<const WfJob*>(s_tab + a.n_phases * sizeof(WfPhase));
        for (int p = 0; p < a.n_phases; ++p)
        {
            const WfPhase ph  = phases[p];
            const bool serial = ph.pad != 0;
            if (a.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0)
            {
                a.trace[p] = wf_now();
            }
            if (!serial || blockIdx.x == 0)
            {
                for (int w = serial ? 0 : blockIdx.x; w < ph.total_ctas; w += serial ? 1 : gridDim.x)
                {
                    int j = ph.first_job, local = w;
                    while (local >= jobs[j].n_ctas)
                    {
                        local -= jobs[j].n_ctas;
                        ++j;
                    }
                    wf_chunk<DIM, RADIUS>(a, jobs[j], local, s_prefix);
                    __syncthreads(); // s_prefix is reused by the next chunk
                }
            }
            if (p + 1 < a.n_phases)
            {
                const WfPhase nx = phases[p + 1];
                if (nx.pad != 0)
                {
                    if (blockIdx.x == 0)
                    {
                        for (int w = 0; w < nx.total_ctas; ++w)
                        {
                            wf_prefetch(a, jobs, nx, w);
                        }
                    }
                }
                else
                {
                    wf_prefetch(a, jobs, nx, blockIdx.x);
                }
                if (serial && phases[p + 1].pad != 0)
                {
                    __threadfence(); // CTA 0 runs the next phase too: its own writes, ordered by the CTA barrier above
                    __syncthreads();
                }
                else
                {
                    ++barriers;
                    if (!(a.mg_world > 1 ? wf_mg_barrier(a, barriers) : wf_grid_barrier(a.barrier, a.barrier_base + barriers * gridDim.x)))
                    {
                        return; // barrier timed out: error word raised, results are invalid and the host will throw
                    }
                }
            }
        }
        if (a.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0)
        {
            a.trace[a.n_phases] = wf_now();
        }
    }
} // namespace smr
