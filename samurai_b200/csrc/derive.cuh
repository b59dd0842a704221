// Device-side derivation of the index records (items.h) from seeds and the reference sub-mesh CSR.
//
// What it replaces on the reference side: every access `field(level, interval, index)` resolves the storage offset of
// a row through LevelCellArray::find / get_interval (level_cell_array.hpp, cell_array.hpp:484-493 for the numbering).
// Round 1 did these lookups on the host for every record of every subset after every adaptation (2.4 ms per step on the
// max_level-14 workload and 15 MB of records over PCIe).  Now the host ships the CSR itself (16 B per reference interval)
// plus a 24 B seed per record, and one launch of derive_kernel writes all records of a plan straight into HBM: one thread
// per record, binary search of the row key then of the interval, all lookups of a record issued independently.
#pragma once
#include "items.h"

#include <cuda_runtime.h>

namespace smr
{
    struct DeriveArgs
    {
        const char* arena;          // seeds, job table (uploaded) and records (device-only region)
        char* arena_out;            // same buffer, writable
        const smr_derive_job* jobs; // n_jobs entries (inside the arena)
        int n_jobs;
        int dim, radius;
        const char* csr_dst; // CSR buffer the destination / own-level offsets come from
        const char* csr_src; // CSR buffer of the source mesh (== csr_dst except for the field transfer old -> new)
        smr_csr_table tab_dst, tab_src;
        unsigned* error; // [0] = sticky flag (1 + job kind), [1..4] = level, x, y, z of the first failing lookup
    };

#ifdef SMR_DERIVE_KERNEL // the kernel itself is compiled by k_derive.cu only; capi.cu just fills DeriveArgs
    __device__ __forceinline__ int64_t dv_key(int y, int z)
    {
        return (static_cast<int64_t>(z + (1 << 24)) << 32) | static_cast<uint32_t>(y + (1 << 24));
    }

    // storage offset of (x, row (y, z)) at one level when [x, x_last] lies inside one interval, else -1
    __device__ __forceinline__ int64_t csr_find(const char* base, const smr_csr_level& L, int y, int z, int x, int x_last)
    {
        if (L.rows == 0)
        {
            return -1;
        }
        const int64_t* __restrict__ key = reinterpret_cast<const int64_t*>(base + L.key);
        const int64_t k = dv_key(y, z);
        int lo = 0, hi = L.rows;
        while (lo < hi)
        {
            const int mid = (lo + hi) >> 1;
            if (key[mid] < k)
            {
                lo = mid + 1;
            }
            else
            {
                hi = mid;
            }
        }
        if (lo >= L.rows || key[lo] != k)
        {
            return -1;
        }
        const int32_t* __restrict__ ptr = reinterpret_cast<const int32_t*>(base + L.ptr);
        const int32_t* __restrict__ xs  = reinterpret_cast<const int32_t*>(base + L.xs);
        const int q0 = ptr[lo];
        int a = q0, b = ptr[lo + 1];
        while (a < b)
        {
            const int m = (a + b) >> 1;
            if (xs[m] <= x)
            {
                a = m + 1;
            }
            else
            {
                b = m;
            }
        }
        const int i = a - 1;
        if (i < q0)
        {
            return -1;
        }
        const int32_t* __restrict__ xe = reinterpret_cast<const int32_t*>(base + L.xe);
        if (x_last >= xe[i])
        {
            return -1;
        }
        return reinterpret_cast<const int64_t*>(base + L.off)[i] + (x - xs[i]);
    }

    struct DeriveCtx
    {
        const DeriveArgs& a;
        int kind;
        bool failed = false;
        int f_level = 0, f_x = 0, f_y = 0, f_z = 0;

        __device__ __forceinline__ int64_t need(bool src, int level, int y, int z, int x, int x_last)
        {
            int64_t o = -1;
            if (level >= 0 && level < SMR_MAX_LEVELS)
            {
                o = src ? csr_find(a.csr_src, a.tab_src.lv[level], y, z, x, x_last) : csr_find(a.csr_dst, a.tab_dst.lv[level], y, z, x, x_last);
            }
            if (o < 0 && !failed)
            {
                failed  = true;
                f_level = level;
                f_x     = x;
                f_y     = y;
                f_z     = z;
            }
            return o;
        }

        __device__ __forceinline__ void report()
        {
            if (failed && atomicCAS(a.error, 0u, static_cast<unsigned>(1 + kind)) == 0u)
            {
                a.error[1] = static_cast<unsigned>(f_level);
                a.error[2] = static_cast<unsigned>(f_x);
                a.error[3] = static_cast<unsigned>(f_y);
                a.error[4] = static_cast<unsigned>(f_z);
            }
        }
    };

    __global__ void __launch_bounds__(SMR_CTA_THREADS) derive_kernel(DeriveArgs a)
    {
        // block -> job: the table is small (a few jobs per level), first_block is non-decreasing
        __shared__ int s_job;
        if (threadIdx.x == 0)
        {
            int lo = 0, hi = a.n_jobs - 1;
            while (lo < hi)
            {
                const int mid = (lo + hi + 1) >> 1;
                if (a.jobs[mid].first_block <= static_cast<int>(blockIdx.x))
                {
                    lo = mid;
                }
                else
                {
                    hi = mid - 1;
                }
            }
            s_job = lo;
        }
        __syncthreads();
        const smr_derive_job jb = a.jobs[s_job];
        const int i = (static_cast<int>(blockIdx.x) - jb.first_block) * SMR_CTA_THREADS + static_cast<int>(threadIdx.x);
        if (i >= jb.n)
        {
            return;
        }
        const smr_seed sd  = reinterpret_cast<const smr_seed*>(a.arena + jb.seeds)[i];
        const int level    = sd.level & 0xff;
        const int mask     = (sd.level >> 8) & 0xff;
        const int dim      = a.dim;
        const int s        = sd.xs;
        const int e        = sd.xs + sd.n;
        const int y = sd.y, z = sd.z;
        const int ny = dim > 1 ? 2 : 1, nz = dim > 2 ? 2 : 1;
        DeriveCtx c{a, jb.kind};
        switch (jb.kind)
        {
            case SMR_DERIVE_FV:
            {
                smr_item_fv it;
                it.c  = c.need(false, level, y, z, s - 1, e) + 1;
                it.ym = it.yp = it.zm = it.zp = it.c;
                if (dim > 1)
                {
                    it.ym = c.need(false, level, y - 1, z, s, e - 1);
                    it.yp = c.need(false, level, y + 1, z, s, e - 1);
                }
                if (dim > 2)
                {
                    it.zm = c.need(false, level, y, z - 1, s, e - 1);
                    it.zp = c.need(false, level, y, z + 1, s, e - 1);
                }
                it.n     = sd.n;
                it.level = level;
                it.x     = s;
                it.y     = y;
                it.z     = z;
                it.mask  = mask;
                reinterpret_cast<smr_item_fv*>(a.arena_out + jb.items)[i] = it;
                break;
            }
            case SMR_DERIVE_FVSTRIP:
            {
                constexpr int R = SMR_STRIP_ROWS;
                smr_item_fvstrip it;
                it.row[0]     = c.need(false, level, y - 1, z, s, e - 1);
                it.row[R + 1] = c.need(false, level, y + R, z, s, e - 1);
#pragma unroll
                for (int r = 0; r < R; ++r)
                {
                    it.row[r + 1] = c.need(false, level, y + r, z, s - 1, e) + 1;
                    it.zm[r] = it.zp[r] = 0;
                    if (dim > 2)
                    {
                        it.zm[r] = c.need(false, level, y + r, z - 1, s, e - 1);
                        it.zp[r] = c.need(false, level, y + r, z + 1, s, e - 1);
                    }
                }
                it.n     = sd.n;
                it.level = level;
                it.mask  = mask;
                it.x     = s;
                it.y     = y;
                it.z     = z;
                reinterpret_cast<smr_item_fvstrip*>(a.arena_out + jb.items)[i] = it;
                break;
            }
            case SMR_DERIVE_PROJ:
            {
                smr_item_proj it;
                it.dst = c.need(false, level, y, z, s, e - 1);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                {
                    it.src[k] = 0;
                }
                for (int cz = 0; cz < nz; ++cz)
                {
                    for (int cy = 0; cy < ny; ++cy)
                    {
                        it.src[cy + 2 * cz] = c.need(true, level + 1, dim > 1 ? 2 * y + cy : 0, dim > 2 ? 2 * z + cz : 0, 2 * s, 2 * e - 1);
                    }
                }
                it.n    = sd.n;
                it.mask = mask;
                reinterpret_cast<smr_item_proj*>(a.arena_out + jb.items)[i] = it;
                break;
            }
            case SMR_DERIVE_PRED:
            {
                smr_item_pred it;
                const int radius = a.radius;
                const int ry_ = dim > 1 ? radius : 0, rz_ = dim > 2 ? radius : 0;
                it.dst = c.need(false, level, y, z, s, e - 1);
                it.n   = sd.n;
                it.par = (s & 1) | ((dim > 1 ? (y & 1) : 0) << 1) | ((dim > 2 ? (z & 1) : 0) << 2) | (mask << 8);
                const int sc = s >> 1, ec = (e - 1) >> 1;
#pragma unroll
                for (int k = 0; k < 9; ++k)
                {
                    it.src[k] = 0;
                }
                for (int rz = -rz_; rz <= rz_; ++rz)
                {
                    for (int ry = -ry_; ry <= ry_; ++ry)
                    {
                        it.src[(ry + 1) + 3 * (rz + 1)] = c.need(true, level - 1, (y >> 1) + ry, (z >> 1) + rz, sc - radius, ec + radius) + radius;
                    }
                }
                reinterpret_cast<smr_item_pred*>(a.arena_out + jb.items)[i] = it;
                break;
            }
            case SMR_DERIVE_DETAIL:
            {
                smr_item_detail it;
                const int radius = a.radius;
                const int ry_ = dim > 1 ? radius : 0, rz_ = dim > 2 ? radius : 0;
#pragma unroll
                for (int k = 0; k < 9; ++k)
                {
                    it.coarse[k] = 0;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k)
                {
                    it.fine[k] = 0;
                }
                for (int rz = -rz_; rz <= rz_; ++rz)
                {
                    for (int ry = -ry_; ry <= ry_; ++ry)
                    {
                        it.coarse[(ry + 1) + 3 * (rz + 1)] = c.need(false, level, y + ry, z + rz, s - radius, e - 1 + radius) + radius;
                    }
                }
                for (int cz = 0; cz < nz; ++cz)
                {
                    for (int cy = 0; cy < ny; ++cy)
                    {
                        it.fine[cy + 2 * cz] = c.need(false, level + 1, dim > 1 ? 2 * y + cy : 0, dim > 2 ? 2 * z + cz : 0, 2 * s, 2 * e - 1);
                    }
                }
                it.n    = sd.n;
                it.mask = mask;
                reinterpret_cast<smr_item_detail*>(a.arena_out + jb.items)[i] = it;
                break;
            }
            case SMR_DERIVE_TAG:
            {
                smr_item_tag it;
                it.coarse = c.need(false, level, y, z, s, e - 1);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                {
                    it.fine[k] = 0;
                }
                for (int cz = 0; cz < nz; ++cz)
                {
                    for (int cy = 0; cy < ny; ++cy)
                    {
                        it.fine[cy + 2 * cz] = c.need(false, level + 1, dim > 1 ? 2 * y + cy : 0, dim > 2 ? 2 * z + cz : 0, 2 * s, 2 * e - 1);
                    }
                }
                it.n     = sd.n;
                it.level = (level + 1) | (mask << 8);
                reinterpret_cast<smr_item_tag*>(a.arena_out + jb.items)[i] = it;
                break;
            }
            default: // SMR_DERIVE_COPY
            {
                smr_item_copy it;
                it.dst  = c.need(false, level, y, z, s, e - 1);
                it.src  = c.need(true, level, y, z, s, e - 1);
                it.n    = sd.n;
                it.mask = mask;
                reinterpret_cast<smr_item_copy*>(a.arena_out + jb.items)[i] = it;
                break;
            }
        }
        c.report();
    }
#endif
} // namespace smr
