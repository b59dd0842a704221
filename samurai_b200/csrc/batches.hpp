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

#include <cstdio>
#include <functional>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <unordered_map>
#ifdef _OPENMP
#include <omp.h>
#endif

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
        int cta_units   = SMR_CTA_CELLS; // output units per CTA (256 x the kernel's units per thread)
        // byte offsets into the arena
        int64_t items = -1, prefix = -1, cta_first = -1, aux = -1;
        int64_t seeds = -1; // device-derived batches: the uploaded seeds the records come from
        int level     = -1;

        bool empty() const
        {
            return n_items == 0;
        }
    };

    // Host staging buffer for all batches of one plan.  Allocation goes through alloc_fn/free_fn so the C ABI can make
    // it pinned memory (one cudaMemcpyAsync uploads the whole plan without an intermediate copy).
    struct Arena
    {
        uint8_t* p  = nullptr;
        size_t cap  = 0;
        size_t size = 0;

        static inline void* (*alloc_fn)(size_t) = nullptr;
        static inline size_t min_pinned_cap    = size_t(48) << 20;
        size_t min_cap = 0; // this arena's own floor (0: min_pinned_cap when pinned)
        static inline void (*free_fn)(void*)    = nullptr;

        Arena()                        = default;
        Arena(const Arena&)            = delete;
        Arena& operator=(const Arena&) = delete;

        ~Arena()
        {
            release();
        }

        void release()
        {
            if (p)
            {
                if (free_fn)
                {
                    free_fn(p);
                }
                else
                {
                    std::free(p);
                }
            }
            p   = nullptr;
            cap = 0;
        }

        // Device-only region: records the device derives itself (derive.cuh) are laid out behind the uploaded bytes; the
        // host buffer never holds them.  take_dev() hands out offsets relative to the region, close_layout() turns them
        // into arena offsets once the size of the uploaded part is known.
        size_t dev_size = 0;
        size_t dev_base = 0;
        std::vector<int64_t*> dev_fixups;

        void clear()
        {
            size     = 0;
            dev_size = 0;
            dev_base = 0;
            dev_fixups.clear();
        }

        void take_dev(size_t bytes, int64_t* where)
        {
            dev_size = (dev_size + 15) & ~size_t(15);
            *where   = static_cast<int64_t>(dev_size);
            dev_size += bytes;
            dev_fixups.push_back(where);
        }

        void close_layout()
        {
            dev_base = (size + 255) & ~size_t(255);
            for (int64_t* w : dev_fixups)
            {
                *w += static_cast<int64_t>(dev_base);
            }
            dev_fixups.clear();
        }

        // bytes the device buffer needs (uploaded part + derived records)
        size_t device_bytes() const
        {
            return dev_size ? dev_base + dev_size : size;
        }

        // layout pass: reserve `bytes` (16-byte aligned) and return their offset
        size_t take(size_t bytes)
        {
            size             = (size + 15) & ~size_t(15);
            const size_t off = size;
            size += bytes;
            return off;
        }

        // after the layout pass: make sure the memory exists
        void commit()
        {
            if (size > cap)
            {
                release();
                // pinned when the C ABI installs alloc_fn: pinning pages (and mapping them for every GPU of the box) was
                // measured at ~300 ms per regrowth, so start generously and double
                cap = std::max<size_t>(2 * size + 4096, alloc_fn ? (min_cap ? min_cap : min_pinned_cap) : 0);
                p   = static_cast<uint8_t*>(alloc_fn ? alloc_fn(cap) : std::malloc(cap));
                if (!p)
                {
                    throw std::bad_alloc();
                }
            }
        }
    };

    // A batch waiting to be laid out and written: the concatenation of `parts` (per-level record vectors).
    template <class Item>
    struct Pending
    {
        Batch* out = nullptr;
        int kind   = -1;
        int level  = -1;
        std::vector<const std::vector<Item>*> parts;
        std::vector<int64_t>* cum = nullptr; // optional per-group cumulative output-cell counts
        bool inclusive            = false;
        int cta_units             = SMR_CTA_CELLS;
        std::vector<int> group;   // group (e.g. level) of every part, non-decreasing; empty: every part is its own group
        int n_groups = 0;
        std::vector<size_t> part_item;  // filled by layout_batch: first record of every part ...
        std::vector<int64_t> part_unit; // ... and the output units before it
    };

    template <class Item>
    inline void layout_batch(Pending<Item>& pd, Arena& arena)
    {
        Batch& b  = *pd.out;
        b         = Batch();
        b.kind    = pd.kind;
        b.level   = pd.level;
        size_t n  = 0;
        int64_t c = 0;
        const int ngroups = pd.n_groups > 0 ? pd.n_groups : static_cast<int>(pd.parts.size());
        if (pd.cum)
        {
            pd.cum->assign(static_cast<size_t>(ngroups) + 1, 0);
        }
        std::vector<int64_t> per_group(static_cast<size_t>(ngroups), 0);
        pd.part_item.assign(pd.parts.size(), 0);
        pd.part_unit.assign(pd.parts.size(), 0);
        for (size_t k = 0; k < pd.parts.size(); ++k)
        {
            pd.part_item[k] = n;
            pd.part_unit[k] = c;
            int64_t cp = 0;
            for (const Item& it : *pd.parts[k])
            {
                cp += it.n;
            }
            c += cp;
            per_group[static_cast<size_t>(pd.n_groups > 0 ? pd.group[k] : static_cast<int>(k))] += cp;
            n += pd.parts[k]->size();
        }
        if (pd.cum)
        {
            int64_t acc = 0;
            for (int gi = 0; gi < ngroups; ++gi)
            {
                if (!pd.inclusive)
                {
                    (*pd.cum)[static_cast<size_t>(gi)] = acc;
                }
                acc += per_group[static_cast<size_t>(gi)];
                if (pd.inclusive)
                {
                    (*pd.cum)[static_cast<size_t>(gi)] = acc;
                }
            }
            (*pd.cum)[static_cast<size_t>(ngroups)] = acc;
        }
        b.n_items = static_cast<int>(n);
        if (n == 0)
        {
            return;
        }
        b.n_cells   = c;
        b.cta_units = pd.cta_units;
        b.n_ctas    = static_cast<int>((c + pd.cta_units - 1) / pd.cta_units);
        b.items     = static_cast<int64_t>(arena.take(n * sizeof(Item)));
        b.prefix    = static_cast<int64_t>(arena.take((n + 1) * sizeof(int64_t)));
        b.cta_first = static_cast<int64_t>(arena.take((static_cast<size_t>(b.n_ctas) + 1) * sizeof(int32_t)));
    }

    // records and prefix entries of ONE part (parts of a batch can be filled concurrently)
    template <class Item>
    inline void fill_part(const Pending<Item>& pd, Arena& arena, size_t k)
    {
        const Batch& b = *pd.out;
        if (b.n_items == 0 || pd.parts[k]->empty())
        {
            return;
        }
        Item* items     = reinterpret_cast<Item*>(arena.p + b.items);
        int64_t* prefix = reinterpret_cast<int64_t*>(arena.p + b.prefix);
        size_t i        = pd.part_item[k];
        int64_t acc     = pd.part_unit[k];
        std::memcpy(items + i, pd.parts[k]->data(), pd.parts[k]->size() * sizeof(Item));
        for (const Item& it : *pd.parts[k])
        {
            prefix[i++] = acc;
            acc += it.n;
        }
    }

    // closing prefix entry and the per-CTA table, once every part is in place
    template <class Item>
    inline void finish_batch(const Pending<Item>& pd, Arena& arena)
    {
        const Batch& b = *pd.out;
        if (b.n_items == 0)
        {
            return;
        }
        int64_t* prefix = reinterpret_cast<int64_t*>(arena.p + b.prefix);
        int32_t* first  = reinterpret_cast<int32_t*>(arena.p + b.cta_first);
        prefix[b.n_items] = b.n_cells;
        size_t it = 0;
        for (int c = 0; c < b.n_ctas; ++c)
        {
            const int64_t g = static_cast<int64_t>(c) * b.cta_units;
            while (prefix[it + 1] <= g)
            {
                ++it;
            }
            first[c] = static_cast<int32_t>(it);
        }
        first[b.n_ctas] = b.n_items - 1;
    }

    template <class Item>
    inline void fill_batch(const Pending<Item>& pd, Arena& arena)
    {
        const Batch& b = *pd.out;
        if (b.n_items == 0)
        {
            return;
        }
        Item* items     = reinterpret_cast<Item*>(arena.p + b.items);
        int64_t* prefix = reinterpret_cast<int64_t*>(arena.p + b.prefix);
        int32_t* first  = reinterpret_cast<int32_t*>(arena.p + b.cta_first);
        size_t i        = 0;
        int64_t acc     = 0;
        for (const auto* part : pd.parts)
        {
            if (!part->empty())
            {
                std::memcpy(items + i, part->data(), part->size() * sizeof(Item));
            }
            for (const Item& it : *part)
            {
                prefix[i++] = acc;
                acc += it.n;
            }
        }
        prefix[i] = acc;
        size_t it = 0;
        for (int c = 0; c < b.n_ctas; ++c)
        {
            const int64_t g = static_cast<int64_t>(c) * b.cta_units;
            while (prefix[it + 1] <= g)
            {
                ++it;
            }
            first[c] = static_cast<int32_t>(it);
        }
        first[b.n_ctas] = b.n_items - 1;
    }

    // ---------------------------------------------------------------------------------------------------------
    // Batches whose records the DEVICE derives (derive.cuh): the host lays out seeds + prefix + per-CTA table in the
    // uploaded part of the arena and reserves the records in the device-only part.
    // ---------------------------------------------------------------------------------------------------------
    struct DeriveList
    {
        std::vector<smr_derive_job> jobs;
        int64_t jobs_off = -1; // uploaded copy of `jobs`
        int blocks       = 0;  // CTAs of the derive launch

        void clear()
        {
            jobs.clear();
            jobs_off = -1;
            blocks   = 0;
        }
    };

    struct PendingSeeds
    {
        Batch* out       = nullptr;
        int kind         = -1; // B_*
        int derive       = -1; // SMR_DERIVE_*
        int level        = -1;
        size_t item_size = 0;
        std::vector<const std::vector<smr_seed>*> parts;
        std::vector<int64_t>* cum = nullptr; // optional per-group cumulative output-cell counts
        bool inclusive            = false;
        int cta_units             = SMR_CTA_CELLS;
        std::vector<int> group; // group (e.g. level) of every part, non-decreasing; empty: every part is its own group
        int n_groups = 0;
        // records shared with another batch (the per-level tag batches are slices of tag_all): no seeds, no derive job
        const PendingSeeds* alias = nullptr;
        size_t alias_part0        = 0;
        std::vector<size_t> part_item;
        std::vector<int64_t> part_unit;
        int64_t seeds = -1;
    };

    // output units of seed lists whose sum is already known (computed by the parallel tasks that produced them): layout_seeds runs
    // serially over ~40 batches and would otherwise walk every seed again (0.2 ms per plan on the 2D max_level-14 mesh)
    using SeedSums = std::unordered_map<const std::vector<smr_seed>*, int64_t>;

    inline void layout_seeds(PendingSeeds& pd, Arena& arena, DeriveList& dl, const SeedSums* sums = nullptr)
    {
        Batch& b  = *pd.out;
        b         = Batch();
        b.kind    = pd.kind;
        b.level   = pd.level;
        size_t n  = 0;
        int64_t c = 0;
        const int ngroups = pd.n_groups > 0 ? pd.n_groups : static_cast<int>(pd.parts.size());
        if (pd.cum)
        {
            pd.cum->assign(static_cast<size_t>(ngroups) + 1, 0);
        }
        std::vector<int64_t> per_group(static_cast<size_t>(ngroups), 0);
        pd.part_item.assign(pd.parts.size(), 0);
        pd.part_unit.assign(pd.parts.size(), 0);
        for (size_t k = 0; k < pd.parts.size(); ++k)
        {
            pd.part_item[k] = n;
            pd.part_unit[k] = c;
            int64_t cp = 0;
            const auto known = sums != nullptr ? sums->find(pd.parts[k]) : SeedSums::const_iterator();
            if (sums != nullptr && known != sums->end())
            {
                cp = known->second;
            }
            else
            {
                for (const smr_seed& sd : *pd.parts[k])
                {
                    cp += sd.n;
                }
            }
            c += cp;
            per_group[static_cast<size_t>(pd.n_groups > 0 ? pd.group[k] : static_cast<int>(k))] += cp;
            n += pd.parts[k]->size();
        }
        if (pd.cum)
        {
            int64_t acc = 0;
            for (int gi = 0; gi < ngroups; ++gi)
            {
                if (!pd.inclusive)
                {
                    (*pd.cum)[static_cast<size_t>(gi)] = acc;
                }
                acc += per_group[static_cast<size_t>(gi)];
                if (pd.inclusive)
                {
                    (*pd.cum)[static_cast<size_t>(gi)] = acc;
                }
            }
            (*pd.cum)[static_cast<size_t>(ngroups)] = acc;
        }
        b.n_items = static_cast<int>(n);
        if (n == 0)
        {
            return;
        }
        b.n_cells   = c;
        b.cta_units = pd.cta_units;
        b.n_ctas    = static_cast<int>((c + pd.cta_units - 1) / pd.cta_units);
        b.prefix    = static_cast<int64_t>(arena.take((n + 1) * sizeof(int64_t)));
        b.cta_first = static_cast<int64_t>(arena.take((static_cast<size_t>(b.n_ctas) + 1) * sizeof(int32_t)));
        if (pd.alias != nullptr)
        {
            // still relative to the device-only region: fixed up together with the aliased batch
            b.items = pd.alias->out->items + static_cast<int64_t>(pd.alias->part_item[pd.alias_part0] * pd.item_size);
            b.seeds = pd.alias->seeds + static_cast<int64_t>(pd.alias->part_item[pd.alias_part0] * sizeof(smr_seed));
            arena.dev_fixups.push_back(&b.items);
            return;
        }
        pd.seeds = static_cast<int64_t>(arena.take(n * sizeof(smr_seed)));
        b.seeds  = pd.seeds;
        arena.take_dev(n * pd.item_size, &b.items);
        smr_derive_job jb{};
        jb.kind        = pd.derive;
        jb.n           = static_cast<int32_t>(n);
        jb.first_block = dl.blocks;
        jb.seeds       = pd.seeds;
        jb.items       = b.items; // relative: close_derive adds the base
        dl.jobs.push_back(jb);
        dl.blocks += static_cast<int>((n + SMR_CTA_THREADS - 1) / SMR_CTA_THREADS);
    }

    // after every layout_seeds / layout_bc of the arena: place the job table, fix the device-only offsets, get the memory
    inline void close_derive(Arena& arena, DeriveList& dl)
    {
        dl.jobs_off = static_cast<int64_t>(arena.take(std::max<size_t>(dl.jobs.size(), 1) * sizeof(smr_derive_job)));
        arena.close_layout();
        for (smr_derive_job& jb : dl.jobs)
        {
            jb.items += static_cast<int64_t>(arena.dev_base);
        }
        arena.commit();
        if (!dl.jobs.empty())
        {
            std::memcpy(arena.p + dl.jobs_off, dl.jobs.data(), dl.jobs.size() * sizeof(smr_derive_job));
        }
    }

    // seeds and prefix entries of ONE part (parts of a batch can be filled concurrently)
    inline void fill_seeds_part(const PendingSeeds& pd, Arena& arena, size_t k)
    {
        const Batch& b = *pd.out;
        if (b.n_items == 0 || pd.parts[k]->empty())
        {
            return;
        }
        int64_t* prefix = reinterpret_cast<int64_t*>(arena.p + b.prefix);
        size_t i        = pd.part_item[k];
        int64_t acc     = pd.part_unit[k];
        if (pd.alias == nullptr)
        {
            std::memcpy(reinterpret_cast<smr_seed*>(arena.p + pd.seeds) + i, pd.parts[k]->data(), pd.parts[k]->size() * sizeof(smr_seed));
        }
        for (const smr_seed& sd : *pd.parts[k])
        {
            prefix[i++] = acc;
            acc += sd.n;
        }
    }

    // closing prefix entry and the per-CTA table, once every part is in place
    inline void finish_seeds(const PendingSeeds& pd, Arena& arena)
    {
        const Batch& b = *pd.out;
        if (b.n_items == 0)
        {
            return;
        }
        int64_t* prefix = reinterpret_cast<int64_t*>(arena.p + b.prefix);
        int32_t* first  = reinterpret_cast<int32_t*>(arena.p + b.cta_first);
        prefix[b.n_items] = b.n_cells;
        size_t it = 0;
        for (int c = 0; c < b.n_ctas; ++c)
        {
            const int64_t g = static_cast<int64_t>(c) * b.cta_units;
            while (prefix[it + 1] <= g)
            {
                ++it;
            }
            first[c] = static_cast<int32_t>(it);
        }
        first[b.n_ctas] = b.n_items - 1;
    }

    inline void fill_seeds(const PendingSeeds& pd, Arena& arena)
    {
        for (size_t k = 0; k < pd.parts.size(); ++k)
        {
            fill_seeds_part(pd, arena, k);
        }
        finish_seeds(pd, arena);
    }

    inline PendingSeeds pending_seeds(Batch* out, int kind, int derive, int level, size_t item_size, int cta_units = SMR_CTA_CELLS)
    {
        PendingSeeds pd;
        pd.out       = out;
        pd.kind      = kind;
        pd.derive    = derive;
        pd.level     = level;
        pd.item_size = item_size;
        pd.cta_units = cta_units;
        return pd;
    }

    struct PendingBc
    {
        Batch* out = nullptr;
        int level  = -1;
        const std::vector<smr_item_bc>* items = nullptr;
        const std::vector<int64_t>* srcs      = nullptr;
    };

    inline void layout_bc(PendingBc& pd, Arena& arena)
    {
        Batch& b  = *pd.out;
        b         = Batch();
        b.kind    = B_BC;
        b.level   = pd.level;
        b.n_items = static_cast<int>(pd.items->size());
        if (b.n_items == 0)
        {
            return;
        }
        b.n_cells = b.n_items;
        b.n_ctas  = (b.n_items + SMR_CTA_THREADS - 1) / SMR_CTA_THREADS;
        b.items   = static_cast<int64_t>(arena.take(pd.items->size() * sizeof(smr_item_bc)));
        b.aux     = static_cast<int64_t>(arena.take(pd.srcs->size() * sizeof(int64_t)));
    }

    inline void fill_bc(const PendingBc& pd, Arena& arena)
    {
        const Batch& b = *pd.out;
        if (b.n_items == 0)
        {
            return;
        }
        std::memcpy(arena.p + b.items, pd.items->data(), pd.items->size() * sizeof(smr_item_bc));
        if (!pd.srcs->empty())
        {
            std::memcpy(arena.p + b.aux, pd.srcs->data(), pd.srcs->size() * sizeof(int64_t));
        }
    }

    // Multi-GPU ownership: the domain is cut into `world` slabs along the last axis (y in 2D, z in 3D) at positions that
    // balance the leaf count; a record belongs to the rank whose slab contains the centre of its output row, and its
    // outputs are also stored (over NVLink, by the producing thread) on every peer whose slab comes within `margin`
    // cells of the row.  All ranks hold the same global mesh, so every rank derives the same cuts with no negotiation.
    // Replaces the reference's MPI subdomains + interface sets (mesh.hpp:1032-1408, update_ghost_mr.hpp:63-187).
    struct PlanFilter
    {
        int rank = 0, world = 1;
        int dim = 2, L = 0;
        int margin = 8; // cells of the record's level: stencil reach (<= 3) + strip height (4) + slack
        std::vector<int64_t> cut2; // world + 1 cut positions in half cells of level L along the slab axis
        bool periodic_axis = false; // the slab axis is periodic: the first and the last slab are neighbours through the boundary

        bool active() const
        {
            return world > 1;
        }

        unsigned mask_all() const
        {
            return active() ? (((1u << world) - 1u) & ~(1u << rank)) : 0u;
        }

        int axis_coord(int y, int z) const
        {
            return dim > 2 ? z : y;
        }

        // centre of row j (level l) in half cells of level L, clamped into the domain
        int64_t center2(int level, int j) const
        {
            const int sh = L - level;
            int64_t c    = 2 * static_cast<int64_t>(j) + 1;
            c            = sh >= 0 ? (c << sh) : (c >> (-sh));
            return std::min(std::max<int64_t>(c, 1), cut2.back() - 1);
        }

        int owner(int level, int j) const
        {
            if (!active())
            {
                return 0;
            }
            const int64_t c = center2(level, j);
            int r           = 0;
            while (r + 1 < world && cut2[r + 1] <= c)
            {
                ++r;
            }
            return r;
        }

        bool owns(int level, int y, int z) const
        {
            return !active() || owner(level, axis_coord(y, z)) == rank;
        }

        // peers whose slab intersects rows [j - margin, j + rows + margin) of level l
        unsigned mask(int level, int y, int z, int rows = 1) const
        {
            if (!active())
            {
                return 0;
            }
            const int j      = axis_coord(y, z);
            const int sh     = L - level;
            int64_t lo       = 2 * (static_cast<int64_t>(j) - margin);
            int64_t hi       = 2 * (static_cast<int64_t>(j) + rows + margin);
            lo               = sh >= 0 ? (lo << sh) : (lo >> (-sh));
            hi               = sh >= 0 ? (hi << sh) : ((hi >> (-sh)) + 1);
            unsigned m       = 0;
            // a periodic slab axis: the rows near one boundary are the sources of the periodic ghosts beyond the other one
            // (algorithm/update_periodic.hpp:34-125) and of the stencils that reach through it, so the window also counts shifted
            // by one period in both directions
            const int64_t period = cut2.back();
            for (int r = 0; r < world; ++r)
            {
                if (r == rank)
                {
                    continue;
                }
                bool hit = cut2[r] < hi && cut2[r + 1] > lo;
                if (periodic_axis)
                {
                    hit = hit || (cut2[r] < hi + period && cut2[r + 1] > lo + period) || (cut2[r] < hi - period && cut2[r + 1] > lo - period);
                }
                if (hit)
                {
                    m |= 1u << r;
                }
            }
            return m;
        }

        // leaf-balanced cuts for mesh `m` (uniform weight per leaf, like load_balancing/weight.hpp:21)
        void compute_cuts(const Mesh& m)
        {
            dim = m.cfg.dim;
            L   = m.cfg.max_level;
            const int axis   = dim > 2 ? 2 : 1;
            periodic_axis    = dim > 1 && m.cfg.periodic[axis];
            const int64_t nL = static_cast<int64_t>(m.cfg.n0[axis]) << L;
            cut2.assign(world + 1, 0);
            cut2[world] = 2 * nL;
            if (!active())
            {
                return;
            }
            std::vector<int64_t> w(static_cast<size_t>(nL), 0);
            int64_t total = 0;
            for (int l = 0; l < m.nlev; ++l)
            {
                const LevelSet& c = m.cells[l];
                const int sh      = L - l;
                for (size_t r = 0; r < c.rows(); ++r)
                {
                    const int j = axis_coord(key_y(c.key[r]), key_z(c.key[r]));
                    int64_t n   = 0;
                    for (int q = c.ptr[r]; q < c.ptr[r + 1]; ++q)
                    {
                        n += c.xe[q] - c.xs[q];
                    }
                    const int64_t bin = std::min(std::max<int64_t>(((2 * static_cast<int64_t>(j) + 1) << sh) >> 1, 0), nL - 1);
                    w[static_cast<size_t>(bin)] += n;
                    total += n;
                }
            }
            int64_t acc = 0;
            int k       = 1;
            for (int64_t b = 0; b < nL && k < world; ++b)
            {
                acc += w[static_cast<size_t>(b)];
                while (k < world && acc * world >= total * k)
                {
                    cut2[k++] = 2 * (b + 1);
                }
            }
            for (; k < world; ++k)
            {
                cut2[k] = 2 * nL;
            }
        }
    };

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

    // probe-based variant: rows are visited in increasing key order and x in increasing order inside a row
    inline int64_t need(Probe& p, const char* what, int level, int y, int z, int x, int x_last)
    {
        const int64_t o = p.offset(x, x_last);
        if (o < 0)
        {
            missing(what, level, x, y, z);
        }
        return o;
    }

    // ---------------------------------------------------------------------------------------------------------
    // per-interval seed builders: the storage offsets are looked up on the device (derive.cuh)
    // ---------------------------------------------------------------------------------------------------------
    inline smr_seed mk_seed(int level, int y, int z, int s, int e, unsigned mask)
    {
        return smr_seed{s, e - s, y, z, level | static_cast<int>(mask << 8), 0};
    }

    // one seed per x-interval of rows [row_begin, row_end) of `cs` (a set at `level`) whose row this rank owns
    inline void set_seeds(const LevelSet& cs, int level, const PlanFilter& flt, bool to_all, std::vector<smr_seed>& out, size_t row_begin = 0,
                          size_t row_end = ~size_t(0))
    {
        row_end = std::min(row_end, cs.rows());
        if (row_begin >= row_end)
        {
            return;
        }
        out.reserve(out.size() + static_cast<size_t>(cs.ptr[row_end] - cs.ptr[row_begin]));
        for (size_t r = row_begin; r < row_end; ++r)
        {
            const int y = key_y(cs.key[r]), z = key_z(cs.key[r]);
            if (!flt.owns(level, y, z))
            {
                continue;
            }
            const unsigned mask = to_all ? flt.mask_all() : flt.mask(level, y, z);
            for (int q = cs.ptr[r]; q < cs.ptr[r + 1]; ++q)
            {
                out.push_back(mk_seed(level, y, z, cs.xs[q], cs.xe[q], mask));
            }
        }
    }

    inline void fv_seeds(const Mesh& m, int l, const PlanFilter& flt, std::vector<smr_seed>& out, size_t row_begin = 0, size_t row_end = ~size_t(0))
    {
        set_seeds(m.cells[l], l, flt, false, out, row_begin, row_end);
    }

    // Leaves of level l split for the FV kernels: strips of SMR_STRIP_ROWS consecutive rows sharing an x-range, and the
    // single-row remainder.  Every leaf cell lands in exactly one of the two lists.
    inline void fv_split_seeds(const Mesh& m, int l, const PlanFilter& flt, std::vector<smr_seed>& strips, std::vector<smr_seed>& singles,
                               size_t row_begin = 0, size_t row_end = ~size_t(0))
    {
        constexpr int R = SMR_STRIP_ROWS;
        const int dim   = m.cfg.dim;
        const LevelSet& c = m.cells[l];
        if (dim < 2)
        {
            fv_seeds(m, l, flt, singles, row_begin, row_end);
            return;
        }
        std::vector<std::pair<int, int>> common, tmp;
        auto single = [&](int y, int z, int s, int e)
        {
            if (flt.owns(l, y, z))
            {
                singles.push_back(mk_seed(l, y, z, s, e, flt.mask(l, y, z)));
            }
        };
        size_t r0 = row_begin;
        const size_t nrows = std::min(row_end, c.rows());
        while (r0 < nrows)
        {
            const int y0 = key_y(c.key[r0]), z0 = key_z(c.key[r0]);
            bool group = r0 + R <= nrows;
            for (int r = 1; group && r < R; ++r)
            {
                group = c.key[r0 + r] == mk_key(y0 + r, z0);
            }
            if (!group)
            {
                for (int q = c.ptr[r0]; q < c.ptr[r0 + 1]; ++q)
                {
                    single(y0, z0, c.xs[q], c.xe[q]);
                }
                ++r0;
                continue;
            }
            // x-ranges common to the R rows
            common.clear();
            for (int q = c.ptr[r0]; q < c.ptr[r0 + 1]; ++q)
            {
                common.emplace_back(c.xs[q], c.xe[q]);
            }
            for (int r = 1; r < R && !common.empty(); ++r)
            {
                tmp.clear();
                size_t i = 0;
                int q    = c.ptr[r0 + r];
                const int qe = c.ptr[r0 + r + 1];
                while (i < common.size() && q < qe)
                {
                    const int s = std::max(common[i].first, c.xs[q]), e = std::min(common[i].second, c.xe[q]);
                    if (s < e)
                    {
                        tmp.emplace_back(s, e);
                    }
                    if (common[i].second < c.xe[q])
                    {
                        ++i;
                    }
                    else
                    {
                        ++q;
                    }
                }
                common.swap(tmp);
            }
            // strips shorter than 8 columns are not worth a record: leave them to the single-row path
            tmp.clear();
            for (auto& iv : common)
            {
                if (iv.second - iv.first >= 8)
                {
                    tmp.push_back(iv);
                }
            }
            common.swap(tmp);
            if (flt.owns(l, y0, z0))
            {
                const unsigned mask = flt.dim > 2 ? flt.mask(l, y0, z0) : flt.mask(l, y0, z0, R);
                for (auto& iv : common)
                {
                    strips.push_back(mk_seed(l, y0, z0, iv.first, iv.second, mask));
                }
            }
            else
            {
                // another rank runs the strips of this group (ownership is decided on the first row)
            }
            // remainder of every row
            for (int r = 0; r < R; ++r)
            {
                size_t i = 0;
                for (int q = c.ptr[r0 + r]; q < c.ptr[r0 + r + 1]; ++q)
                {
                    int s = c.xs[q];
                    const int e = c.xe[q];
                    while (i < common.size() && common[i].second <= s)
                    {
                        ++i;
                    }
                    size_t j = i;
                    while (s < e)
                    {
                        if (j >= common.size() || common[j].first >= e)
                        {
                            single(y0 + r, z0, s, e);
                            break;
                        }
                        if (common[j].first > s)
                        {
                            single(y0 + r, z0, s, common[j].first);
                        }
                        s = std::max(s, common[j].second);
                        ++j;
                    }
                }
            }
            r0 += R;
        }
    }

    // ---------------------------------------------------------------------------------------------------------
    // flux-based schemes on multi-level meshes: the reference enumerates interfaces (interface.hpp:35-306, 440-509) and
    // scatters both sides' contributions; here every leaf gathers its own contributions, so the host classifies what
    // lies across each face of each leaf (same-level leaf / coarser leaf / finer leaves / boundary) and cuts the leaf
    // intervals where a transverse classification changes.
    // ---------------------------------------------------------------------------------------------------------
    inline void flux_items(const Mesh& m, int l, const PlanFilter& flt, std::vector<smr_item_flux>& out, std::vector<int64_t>& aux,
                           size_t row_begin = 0, size_t row_end = ~size_t(0))
    {
        const int dim        = m.cfg.dim;
        const LevelSet& c    = m.cells[l];
        const LevelSet& ref  = m.ref[l];
        const LevelSet* cc   = l > 0 ? &m.cells[l - 1] : nullptr;
        const LevelSet* cf   = l + 1 < m.nlev ? &m.cells[l + 1] : nullptr;
        const LevelSet* rf   = l + 1 < m.nlev ? &m.ref[l + 1] : nullptr;
        const int nfaces     = 2 * dim;
        int nl[3];
        for (int d = 0; d < 3; ++d)
        {
            nl[d] = m.cfg.n0[d] << l;
        }
        // a face on a periodic boundary is classified by the leaf at the wrapped position (interface.hpp:83-92, 179-189, 280-290) and
        // reads the periodic ghosts at the unwrapped one; the other boundaries keep their boundary interfaces
        auto wrap = [&](int d, int v) { return ((v % nl[d]) + nl[d]) % nl[d]; };
        struct Seg
        {
            int a, b, kind;
        };
        std::vector<Seg> segs[6];
        std::vector<int> cuts;
        auto has = [](const LevelSet* s, int y, int z, int x)
        {
            return s != nullptr && s->contains(mk_key(y, z), x);
        };
        row_end = std::min(row_end, c.rows());
        out.reserve(out.size() + static_cast<size_t>(c.ptr[row_end] - c.ptr[row_begin]));
        for (size_t r = row_begin; r < row_end; ++r)
        {
            const int y = key_y(c.key[r]), z = key_z(c.key[r]);
            if (!flt.owns(l, y, z))
            {
                continue;
            }
            const int mask = static_cast<int>(flt.mask(l, y, z));
            const int py = dim > 1 ? (y >> 1) : 0, pz = dim > 2 ? (z >> 1) : 0; // parent row
            auto cy_ = [&](int a) { return dim > 1 ? 2 * y + a : 0; };             // child rows
            auto cz_ = [&](int a) { return dim > 2 ? 2 * z + a : 0; };
            for (int q = c.ptr[r]; q < c.ptr[r + 1]; ++q)
            {
                const int s = c.xs[q], e = c.xe[q];
                cuts.clear();
                cuts.push_back(s);
                cuts.push_back(e);
                for (int f = 2; f < nfaces; ++f)
                {
                    segs[f].clear();
                    const int d = f >> 1, sgn = (f & 1) ? 1 : -1;
                    int yy = y + (d == 1 ? sgn : 0), zz = z + (d == 2 ? sgn : 0);
                    const int j        = d == 1 ? yy : zz;
                    const bool through = j < 0 || j >= nl[d];
                    if (through && !m.cfg.periodic[d])
                    {
                        segs[f].push_back({s, e, SMR_FACE_BDRY});
                        continue;
                    }
                    if (through)
                    {
                        (d == 1 ? yy : zz) = wrap(d, j);
                    }
                    const int rS = c.find_row(mk_key(yy, zz));
                    const int rC = cc ? cc->find_row(mk_key(dim > 1 ? (yy >> 1) : 0, dim > 2 ? (zz >> 1) : 0)) : -1;
                    // the child row of the neighbour position that touches this leaf
                    const int fy = d == 1 ? 2 * yy + (sgn < 0 ? 1 : 0) : 2 * yy;
                    const int fz = dim > 2 ? (d == 2 ? 2 * zz + (sgn < 0 ? 1 : 0) : 2 * zz) : 0;
                    const int rF = cf ? cf->find_row(mk_key(fy, fz)) : -1;
                    int pos = s;
                    while (pos < e)
                    {
                        int i, end, kind;
                        if (rS >= 0 && (i = c.find_ivl(rS, pos)) >= 0)
                        {
                            end  = std::min(e, c.xe[i]);
                            kind = SMR_FACE_SAME | (through ? 4 : 0); // bit 2 (segments only): through the periodic boundary
                        }
                        else if (rC >= 0 && (i = cc->find_ivl(rC, pos >> 1)) >= 0)
                        {
                            end  = std::min(e, 2 * cc->xe[i]);
                            kind = SMR_FACE_COARSE;
                        }
                        else if (rF >= 0 && (i = cf->find_ivl(rF, 2 * pos)) >= 0)
                        {
                            end  = std::min(e, (cf->xe[i] + 1) >> 1);
                            kind = SMR_FACE_FINE;
                        }
                        else
                        {
                            missing("flux: neighbour leaf", l, pos, yy, zz);
                        }
                        segs[f].push_back({pos, end, kind});
                        cuts.push_back(end);
                        pos = end;
                    }
                }
                std::sort(cuts.begin(), cuts.end());
                cuts.erase(std::unique(cuts.begin(), cuts.end()), cuts.end());
                size_t cur[6] = {0, 0, 0, 0, 0, 0};
                for (size_t ci = 0; ci + 1 < cuts.size(); ++ci)
                {
                    const int a = cuts[ci], b = cuts[ci + 1];
                    int kinds  = 0;
                    bool any_fine = false;
                    // x faces: the neighbour cells a - 1 and b (wrapped through a periodic boundary)
                    int kx[2];
                    for (int side = 0; side < 2; ++side)
                    {
                        if (side == 0 ? a > s : b < e)
                        {
                            kx[side] = SMR_FACE_SAME;
                            continue;
                        }
                        const int xn       = side == 0 ? a - 1 : b;
                        const bool through = xn < 0 || xn >= nl[0];
                        if (through && !m.cfg.periodic[0])
                        {
                            kx[side] = SMR_FACE_BDRY;
                            continue;
                        }
                        const int xw = through ? wrap(0, xn) : xn;
                        if (through && has(&c, y, z, xw))
                        {
                            kx[side] = SMR_FACE_SAME;
                            kinds |= 1 << ((side == 0 ? SMR_FLUXW_SWAP_SHIFT : SMR_FLUX_PLUS_THROUGH_SHIFT) + 0);
                        }
                        else if (has(cc, py, pz, xw >> 1))
                        {
                            kx[side] = SMR_FACE_COARSE;
                        }
                        else if (has(cf, cy_(0), cz_(0), side == 0 ? 2 * xw + 1 : 2 * xw))
                        {
                            kx[side] = SMR_FACE_FINE;
                        }
                        else
                        {
                            missing("flux: x neighbour leaf", l, xn, y, z);
                        }
                    }
                    kinds |= kx[0] | (kx[1] << 2);
                    any_fine = kx[0] == SMR_FACE_FINE || kx[1] == SMR_FACE_FINE;
                    for (int f = 2; f < nfaces; ++f)
                    {
                        while (segs[f][cur[f]].b <= a)
                        {
                            ++cur[f];
                        }
                        const int kind = segs[f][cur[f]].kind & 3;
                        kinds |= kind << (2 * f);
                        any_fine = any_fine || kind == SMR_FACE_FINE;
                        if (segs[f][cur[f]].kind & 4) // same-level interface through the periodic boundary
                        {
                            kinds |= 1 << (((f & 1) ? SMR_FLUX_PLUS_THROUGH_SHIFT : SMR_FLUXW_SWAP_SHIFT) + (f >> 1));
                        }
                    }
                    smr_item_flux it;
                    it.c = ref.offset_of(c.key[r], a - 1, b);
                    if (it.c < 0)
                    {
                        missing("flux x", l, a - 1, y, z);
                    }
                    it.c += 1;
                    for (int k = 0; k < 4; ++k)
                    {
                        it.nb[k] = it.c;
                    }
                    for (int f = 2; f < nfaces; ++f)
                    {
                        const int d = f >> 1, sgn = (f & 1) ? 1 : -1;
                        it.nb[f - 2] = need(ref, "flux transverse row", l, y + (d == 1 ? sgn : 0), z + (d == 2 ? sgn : 0), a, b - 1);
                    }
                    it.fine = 0;
                    if (any_fine)
                    {
                        it.fine = static_cast<int64_t>(aux.size());
                        aux.resize(aux.size() + SMR_FLUX_AUX_SLOTS, 0);
                        int64_t* fx = aux.data() + it.fine;
                        const int nr = 1 << (dim - 1);
                        for (int side = 0; side < 2; ++side)
                        {
                            if (kx[side] == SMR_FACE_FINE)
                            {
                                // stencil cells at level+1: x-: {fine 2a-1, ghost 2a}; x+: {ghost 2b-1, fine 2b}
                                const int x0 = side == 0 ? 2 * a - 1 : 2 * b - 1;
                                for (int rr = 0; rr < nr; ++rr)
                                {
                                    fx[side * 4 + rr] = need(*rf, "flux fine x rows", l + 1, cy_(rr & 1), cz_(rr >> 1), x0, x0 + 1);
                                }
                            }
                        }
                        for (int f = 2; f < nfaces; ++f)
                        {
                            if (((kinds >> (2 * f)) & 3) != SMR_FACE_FINE)
                            {
                                continue;
                            }
                            const int d = f >> 1, sgn = (f & 1) ? 1 : -1;
                            const int base = d == 1 ? 2 * y : 2 * z;
                            const int fine_row  = sgn < 0 ? base - 1 : base + 2;
                            const int ghost_row = sgn < 0 ? base : base + 1;
                            const int st_row[2] = {sgn < 0 ? fine_row : ghost_row, sgn < 0 ? ghost_row : fine_row};
                            const int nb_other  = dim > 2 ? 2 : 1;
                            for (int bb = 0; bb < nb_other; ++bb)
                            {
                                for (int st = 0; st < 2; ++st)
                                {
                                    const int ry = d == 1 ? st_row[st] : cy_(bb);
                                    const int rz = d == 2 ? st_row[st] : cz_(bb);
                                    fx[f * 4 + 2 * bb + st] = need(*rf, "flux fine transverse rows", l + 1, ry, rz, 2 * a, 2 * b - 1);
                                }
                            }
                        }
                    }
                    it.n     = b - a;
                    it.level = l;
                    it.kinds = kinds;
                    it.mask  = mask;
                    out.push_back(it);
                }
            }
        }
    }

    // Six-cell line stencil {-2 .. 3} (WENO5, operators/convection_lin.hpp:95-178) on a fully periodic mesh: the same face
    // classification, with the neighbour across a periodic boundary looked up at the wrapped position (interface.hpp:83-92 same level,
    // :179-189 and :280-290 level jumps) while the stencil values are the periodic ghosts at the unwrapped one.
    inline void fluxw_items(const Mesh& m, int l, const PlanFilter& flt, std::vector<smr_item_fluxw>& out, std::vector<int64_t>& aux)
    {
        const int dim       = m.cfg.dim;
        const LevelSet& c   = m.cells[l];
        const LevelSet& ref = m.ref[l];
        const LevelSet* cc  = l > 0 ? &m.cells[l - 1] : nullptr;
        const LevelSet* cf  = l + 1 < m.nlev ? &m.cells[l + 1] : nullptr;
        const LevelSet* rf  = l + 1 < m.nlev ? &m.ref[l + 1] : nullptr;
        const int nfaces    = 2 * dim;
        int nl[3];
        for (int d = 0; d < 3; ++d)
        {
            nl[d] = m.cfg.n0[d] << l;
        }
        auto wrap = [&](int d, int v) { return ((v % nl[d]) + nl[d]) % nl[d]; };
        struct Seg
        {
            int a, b, kind;
        };
        std::vector<Seg> segs[6];
        std::vector<int> cuts;
        auto has = [](const LevelSet* s, int y, int z, int x)
        {
            return s != nullptr && s->contains(mk_key(y, z), x);
        };
        out.reserve(out.size() + static_cast<size_t>(c.ptr[c.rows()]));
        for (size_t r = 0; r < c.rows(); ++r)
        {
            const int y = key_y(c.key[r]), z = key_z(c.key[r]);
            if (!flt.owns(l, y, z))
            {
                continue;
            }
            const int mask = static_cast<int>(flt.mask(l, y, z));
            const int py = dim > 1 ? (y >> 1) : 0, pz = dim > 2 ? (z >> 1) : 0; // parent row
            auto cy_ = [&](int a) { return dim > 1 ? 2 * y + a : 0; };             // child rows
            auto cz_ = [&](int a) { return dim > 2 ? 2 * z + a : 0; };
            for (int q = c.ptr[r]; q < c.ptr[r + 1]; ++q)
            {
                const int s = c.xs[q], e = c.xe[q];
                cuts.clear();
                cuts.push_back(s);
                cuts.push_back(e);
                int swap_bits = 0;
                for (int f = 2; f < nfaces; ++f)
                {
                    segs[f].clear();
                    const int d = f >> 1, sgn = (f & 1) ? 1 : -1;
                    int yy = y + (d == 1 ? sgn : 0), zz = z + (d == 2 ? sgn : 0);
                    const int j        = d == 1 ? yy : zz;
                    const bool through = j < 0 || j >= nl[d];
                    if (through)
                    {
                        (d == 1 ? yy : zz) = wrap(d, j);
                    }
                    const int rS = c.find_row(mk_key(yy, zz));
                    const int rC = cc ? cc->find_row(mk_key(dim > 1 ? (yy >> 1) : 0, dim > 2 ? (zz >> 1) : 0)) : -1;
                    // the child row of the neighbour position that touches this leaf
                    const int fy = d == 1 ? 2 * yy + (sgn < 0 ? 1 : 0) : 2 * yy;
                    const int fz = dim > 2 ? (d == 2 ? 2 * zz + (sgn < 0 ? 1 : 0) : 2 * zz) : 0;
                    const int rF = cf ? cf->find_row(mk_key(fy, fz)) : -1;
                    int pos = s;
                    while (pos < e)
                    {
                        int i, end, kind;
                        if (rS >= 0 && (i = c.find_ivl(rS, pos)) >= 0)
                        {
                            end  = std::min(e, c.xe[i]);
                            kind = SMR_FACE_SAME;
                            if (through && sgn < 0)
                            {
                                swap_bits |= 1 << (SMR_FLUXW_SWAP_SHIFT + d); // whole interval or none: see the cut below
                            }
                        }
                        else if (rC >= 0 && (i = cc->find_ivl(rC, pos >> 1)) >= 0)
                        {
                            end  = std::min(e, 2 * cc->xe[i]);
                            kind = SMR_FACE_COARSE;
                        }
                        else if (rF >= 0 && (i = cf->find_ivl(rF, 2 * pos)) >= 0)
                        {
                            end  = std::min(e, (cf->xe[i] + 1) >> 1);
                            kind = SMR_FACE_FINE;
                        }
                        else
                        {
                            missing("wide flux: neighbour leaf", l, pos, yy, zz);
                        }
                        segs[f].push_back({pos, end, kind});
                        cuts.push_back(end);
                        pos = end;
                    }
                }
                std::sort(cuts.begin(), cuts.end());
                cuts.erase(std::unique(cuts.begin(), cuts.end()), cuts.end());
                size_t cur[6] = {0, 0, 0, 0, 0, 0};
                for (size_t ci = 0; ci + 1 < cuts.size(); ++ci)
                {
                    const int a = cuts[ci], b = cuts[ci + 1];
                    int kinds     = 0;
                    bool any_fine = false;
                    // x faces: the neighbour cells a - 1 and b, wrapped through the periodic boundary
                    int kx[2];
                    for (int side = 0; side < 2; ++side)
                    {
                        const bool inside = side == 0 ? a > s : b < e;
                        if (inside)
                        {
                            kx[side] = SMR_FACE_SAME;
                            continue;
                        }
                        const int xn       = side == 0 ? a - 1 : b;
                        const bool through = xn < 0 || xn >= nl[0];
                        const int xw       = wrap(0, xn);
                        if (through && has(&c, y, z, xw))
                        {
                            kx[side] = SMR_FACE_SAME;
                            if (side == 0)
                            {
                                kinds |= 1 << SMR_FLUXW_SWAP_SHIFT;
                            }
                        }
                        else if (has(cc, py, pz, xw >> 1))
                        {
                            kx[side] = SMR_FACE_COARSE;
                        }
                        else if (has(cf, cy_(0), cz_(0), side == 0 ? 2 * xw + 1 : 2 * xw))
                        {
                            kx[side] = SMR_FACE_FINE;
                        }
                        else
                        {
                            missing("wide flux: x neighbour leaf", l, xn, y, z);
                        }
                    }
                    kinds |= kx[0] | (kx[1] << 2);
                    any_fine = kx[0] == SMR_FACE_FINE || kx[1] == SMR_FACE_FINE;
                    for (int f = 2; f < nfaces; ++f)
                    {
                        while (segs[f][cur[f]].b <= a)
                        {
                            ++cur[f];
                        }
                        const int kind = segs[f][cur[f]].kind;
                        kinds |= kind << (2 * f);
                        any_fine = any_fine || kind == SMR_FACE_FINE;
                        if (kind == SMR_FACE_SAME && !(f & 1))
                        {
                            kinds |= swap_bits & (1 << (SMR_FLUXW_SWAP_SHIFT + (f >> 1)));
                        }
                    }
                    smr_item_fluxw it;
                    it.c = need(ref, "wide flux x", l, y, z, a - 3, b + 2) + 3;
                    for (int k = 0; k < 12; ++k)
                    {
                        it.nb[k] = it.c;
                    }
                    for (int d = 1; d < dim; ++d)
                    {
                        for (int o = -3; o <= 3; ++o)
                        {
                            if (o != 0)
                            {
                                it.nb[6 * (d - 1) + (o < 0 ? o + 3 : o + 2)] = need(ref, "wide flux transverse row", l, y + (d == 1 ? o : 0),
                                                                                    z + (d == 2 ? o : 0), a, b - 1);
                            }
                        }
                    }
                    it.fine = 0;
                    if (any_fine)
                    {
                        it.fine = static_cast<int64_t>(aux.size());
                        aux.resize(aux.size() + SMR_FLUXW_AUX_SLOTS, 0);
                        int64_t* fx  = aux.data() + it.fine;
                        const int nr = 1 << (dim - 1);
                        for (int side = 0; side < 2; ++side)
                        {
                            if (kx[side] == SMR_FACE_FINE)
                            {
                                // stencil origin at level+1 (unwrapped): x-: the fine leaf 2a-1 (ghost 2a); x+: the ghost 2b-1 (fine leaf 2b)
                                const int x0 = side == 0 ? 2 * a - 1 : 2 * b - 1;
                                for (int rr = 0; rr < nr; ++rr)
                                {
                                    fx[side * 4 + rr] = need(*rf, "wide flux fine x rows", l + 1, cy_(rr & 1), cz_(rr >> 1), x0 - 2, x0 + 3) + 2;
                                }
                            }
                        }
                        for (int f = 2; f < nfaces; ++f)
                        {
                            if (((kinds >> (2 * f)) & 3) != SMR_FACE_FINE)
                            {
                                continue;
                            }
                            const int d = f >> 1, sgn = (f & 1) ? 1 : -1;
                            const int base     = d == 1 ? 2 * y : 2 * z;
                            const int origin   = sgn < 0 ? base - 1 : base + 1; // minus: the fine leaf row; plus: the ghost row
                            const int nb_other = dim > 2 ? 2 : 1;
                            for (int bb = 0; bb < nb_other; ++bb)
                            {
                                for (int st = 0; st < 6; ++st)
                                {
                                    const int row = origin - 2 + st;
                                    const int ry  = d == 1 ? row : cy_(bb);
                                    const int rz  = d == 2 ? row : cz_(bb);
                                    fx[8 + ((f - 2) * 2 + bb) * 6 + st] = need(*rf, "wide flux fine transverse rows", l + 1, ry, rz, 2 * a, 2 * b - 1);
                                }
                            }
                        }
                    }
                    it.n     = b - a;
                    it.level = l;
                    it.kinds = kinds;
                    it.mask  = mask;
                    out.push_back(it);
                }
            }
        }
    }

    // batches for the flux-based schemes on multi-level meshes, built on first use for a mesh (flux schemes only)
    struct FluxPlan
    {
        Arena arena;
        Batch items;
        bool ready = false;
    };

    template <class Item, class ItemsFn>
    inline void build_flux_plan_t(const Mesh& m, FluxPlan& plan, ItemsFn&& items_fn)
    {
        const int nlev = m.nlev;
        std::vector<std::vector<Item>> items(nlev);
        std::vector<std::vector<int64_t>> aux(nlev);
        std::string error;
#pragma omp parallel for schedule(dynamic, 1)
        for (int l = nlev - 1; l >= 0; --l)
        {
            try
            {
                if (!m.cells[l].empty())
                {
                    items_fn(l, items[l], aux[l]);
                }
            }
            catch (const std::exception& e)
            {
#pragma omp critical
                error = e.what();
            }
        }
        if (!error.empty())
        {
            throw std::out_of_range(error);
        }
        // per-level aux indices -> indices into the concatenated aux array
        int64_t base = 0;
        for (int l = 0; l < nlev; ++l)
        {
            if (base != 0)
            {
                for (Item& it : items[l])
                {
                    it.fine += base;
                }
            }
            base += static_cast<int64_t>(aux[l].size());
        }
        plan.arena.clear();
        Pending<Item> pd{&plan.items, B_FV, -1, {}, nullptr, false};
        for (int l = 0; l < nlev; ++l)
        {
            pd.parts.push_back(&items[l]);
        }
        layout_batch(pd, plan.arena);
        plan.items.aux = static_cast<int64_t>(plan.arena.take(static_cast<size_t>(std::max<int64_t>(base, 1)) * sizeof(int64_t)));
        plan.arena.commit();
        fill_batch(pd, plan.arena);
        int64_t* dst = reinterpret_cast<int64_t*>(plan.arena.p + plan.items.aux);
        for (int l = 0; l < nlev; ++l)
        {
            if (!aux[l].empty())
            {
                std::memcpy(dst, aux[l].data(), aux[l].size() * sizeof(int64_t));
                dst += aux[l].size();
            }
        }
        plan.ready = true;
    }

    inline void build_flux_plan(const Mesh& m, FluxPlan& plan, const PlanFilter& flt = PlanFilter())
    {
        build_flux_plan_t<smr_item_flux>(m, plan, [&](int l, std::vector<smr_item_flux>& it, std::vector<int64_t>& aux) { flux_items(m, l, flt, it, aux); });
    }

    // six-cell line stencils (WENO5) on fully periodic meshes
    inline void build_fluxw_plan(const Mesh& m, FluxPlan& plan, const PlanFilter& flt = PlanFilter())
    {
        build_flux_plan_t<smr_item_fluxw>(m, plan, [&](int l, std::vector<smr_item_fluxw>& it, std::vector<int64_t>& aux) { fluxw_items(m, l, flt, it, aux); });
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
    // Top-down sweep of update_ghost_mr, one fused launch per level (update_ghost_mr.hpp:204-222).  The reference
    // runs, per level: corner extrapolation, project_corner_below, then per direction project_bc / apply_field_bc /
    // predict_bc, then the projection to level-1.  Dependencies allow regrouping without changing any value:
    //   * apply_field_bc reads leaves only; predict_bc(level+1) copies the ghost apply_field_bc(level) just wrote, so
    //     both become direct "value" records sourced from the leaf (same arithmetic, same result);
    //   * project_bc(level) reads outside ghosts of level+1/+2 written by earlier (finer) phases;
    //   * project_corner_below(level+1) only feeds project_corner_below(level) and the corner ghost it writes at
    //     `level` is overwritten by the corner extrapolation of `level` when that exists, so its records are folded
    //     into phase `level` with the overwritten ones dropped;
    //   * the projection chain never reads an outside ghost, so it shares the launch.
    struct GhostPhase
    {
        Batch bc;     // extrapolated corners(level) + corner copies from level+1 + project_bc(level) + BC values(level, level+1 children)
        Batch bc2;    // ghost width 2: second ghost layer by polynomial extrapolation and the rest of the corner block; reads what `bc` wrote
        Batch proj;   // projection level -> level-1
        Batch per[3]; // periodic ghosts of the level, one batch per periodic dimension (update_ghost_periodic: the dimensions are
                      // processed one after the other, the later ones copy ghosts the earlier ones filled)
    };

    struct MeshPlan
    {
        Arena arena;
        Batch fv;                     // all leaves, level ascending (initialisation, keep tags)
        Batch keep_bdry;              // `--refine-boundary`: the max_level leaves within max_stencil_radius cells of the domain boundary
        Batch fv_strip, fv_single;    // the same leaves split for the FV kernels (strips of rows + remainder)
        int64_t fv_strip_cells = 0;
        std::vector<GhostPhase> down; // indexed by level (top-down sweep uses L..0)
        std::vector<Batch> pred;      // indexed by level (bottom-up sweep 1..L)
        Batch detail;                 // all coarse levels, ascending
        std::vector<int64_t> detail_cum; // detail_cum[k] = output cells of detail records with coarse level < k
        Batch tag_all;                // criteria: all fine levels, ascending
        std::vector<int64_t> tag_cum; // tag_cum[k] = output cells of tag records with fine level <= k
        std::vector<Batch> tag;       // per fine level, for the sequential keep propagation (records shared with tag_all)
        DeriveList derive;            // what the device has to derive after the upload (all batches above except the bc ones)
        double build_seconds = 0;
    };

    class BcBuilder
    {
      public:

        std::vector<smr_item_bc> items;
        std::vector<int64_t> srcs;
        const PlanFilter* flt = nullptr;
        int cur_mask          = 0;

        // select the row of the ghost cells about to be emitted; false when another rank owns it
        bool target(int level, int y, int z)
        {
            if (flt == nullptr || !flt->active())
            {
                cur_mask = 0;
                return true;
            }
            if (!flt->owns(level, y, z))
            {
                return false;
            }
            cur_mask = static_cast<int>(flt->mask(level, y, z) << 8);
            return true;
        }

        void copy(int64_t dst, int64_t src)
        {
            items.push_back({dst, 0.0, SMR_BC_COPY | cur_mask, 1, static_cast<int64_t>(srcs.size())});
            srcs.push_back(src);
        }

        void value(int64_t dst, int64_t src, double coef)
        {
            items.push_back({dst, coef, SMR_BC_VALUE | cur_mask, 1, static_cast<int64_t>(srcs.size())});
            srcs.push_back(src);
        }

        void extrap4(int64_t dst, int64_t s0, int64_t s1, int64_t s2)
        {
            items.push_back({dst, 0.0, SMR_BC_EXTRAP4 | cur_mask, 3, static_cast<int64_t>(srcs.size())});
            srcs.push_back(s0);
            srcs.push_back(s1);
            srcs.push_back(s2);
        }

        void begin_avg(int64_t dst)
        {
            items.push_back({dst, 0.0, SMR_BC_AVG | cur_mask, 0, static_cast<int64_t>(srcs.size())});
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

    struct PhaseItems
    {
        BcBuilder bc;
        BcBuilder bc2;
        std::vector<smr_seed> proj; // seeds at the coarse level
        std::vector<smr_item_copy> per[3];
    };

    // update_ghost_periodic(level) for dimension d (algorithm/update_periodic.hpp:34-125): the ghosts within ghost_width beyond the
    // upper boundary copy the cells just inside the lower one and vice versa, wherever both exist in the reference sub-mesh.  In the
    // other dimensions the slabs span the domain grown by ghost_width, so that corner ghosts are reached dimension after dimension.
    inline void periodic_items(const Mesh& m, int level, int d, const PlanFilter& flt, std::vector<smr_item_copy>& out)
    {
        const MeshConfig& cfg = m.cfg;
        const LevelSet& ref   = m.ref[level];
        if (ref.empty() || level > cfg.max_level)
        {
            return;
        }
        const int dim = cfg.dim, gw = cfg.ghost_width();
        const int N = cfg.n0[d] << level;
        int lo[3], hi[3];
        for (int side = 0; side < 2; ++side)
        {
            m.domain_box(level, gw, lo, hi);
            // side 0: ghosts in [N, N + gw) <- cells in [0, gw) ; side 1: ghosts in [-gw, 0) <- cells in [N - gw, N)
            const int shift = side == 0 ? N : -N;
            lo[d] = side == 0 ? 0 : N - gw;
            hi[d] = side == 0 ? gw : N;
            const LevelSet src = clip_box(ref, dim, lo, hi);
            lo[d] += shift;
            hi[d] += shift;
            const LevelSet dst = clip_box(ref, dim, lo, hi);
            if (src.empty() || dst.empty())
            {
                continue;
            }
            int sv[3] = {0, 0, 0};
            sv[d]     = shift;
            const LevelSet set = set_inter(translate(src, sv[0], sv[1], sv[2]), dst);
            for (size_t r = 0; r < set.rows(); ++r)
            {
                const int y = key_y(set.key[r]), z = key_z(set.key[r]);
                if (!flt.owns(level, y, z))
                {
                    continue;
                }
                const int mask = static_cast<int>(flt.mask(level, y, z));
                for (int q = set.ptr[r]; q < set.ptr[r + 1]; ++q)
                {
                    const int s = set.xs[q], e = set.xe[q];
                    smr_item_copy it;
                    it.dst  = need(ref, "periodic ghost", level, y, z, s, e - 1);
                    it.src  = need(ref, "periodic source", level, y - sv[1], z - sv[2], s - sv[0], e - 1 - sv[0]);
                    it.n    = e - s;
                    it.mask = mask;
                    out.push_back(it);
                }
            }
        }
    }

    // project_corner_below(src_level) (update_outer_ghost.hpp:267-336): copies of the corner ghost of src_level into
    // the corner ghosts one (dl = 1) and two (dl = 2) levels below, where the corner-most child exists
    template <class F>
    inline void corner_below(const Mesh& m, int src_level, const Dir& d, F&& emit)
    {
        const int dim = m.cfg.dim;
        if (src_level <= 0)
        {
            return;
        }
        const LevelSet& ref       = m.ref[src_level];
        const LevelSet corner     = corner_cells(m, src_level, d);
        const LevelSet fine_outer = set_inter(translate(corner, d.v[0], d.v[1], d.v[2]), ref);
        for (int dl = 1; dl <= 2; ++dl)
        {
            const int pl = src_level - dl;
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
                                  emit(pl, y, z, need(m.ref[pl], "corner below", pl, y, z, x, x), src);
                              }
                          });
            if (pl == 0)
            {
                break;
            }
        }
    }

    // Ghost width 2, corner leaf (x, y, z) of diagonal direction d (bc/apply_field_bc.hpp:313-466): the second diagonal ghost by the
    // 4-point extrapolation along the diagonal, then the off-diagonal ghosts of the corner block copied from the diagonal ghost of
    // their layer.  The copies of layer 2 repeat the extrapolation (same operands, same result) instead of reading the diagonal ghost,
    // which the same phase writes.
    inline void corner_block_width2(const Mesh& m, int level, const Dir& d, int x, int y, int z, const PlanFilter& flt, BcBuilder& g)
    {
        const int dim       = m.cfg.dim;
        const LevelSet& ref = m.ref[level];
        g.flt               = &flt;
        auto off = [&](int cx, int cy, int cz) { return ref.offset_of(mk_key(cy, cz), cx, cx); };
        const int64_t sm = off(x - d.v[0], y - d.v[1], z - d.v[2]);
        const int64_t s0 = off(x, y, z);
        const int64_t s1 = off(x + d.v[0], y + d.v[1], z + d.v[2]);
        const int64_t s2 = off(x + 2 * d.v[0], y + 2 * d.v[1], z + 2 * d.v[2]);
        const bool layer2 = s2 >= 0; // apply_extrapolation_bc_cells<4>: only where the farthest ghost exists
        if (layer2 && (sm < 0 || s0 < 0 || s1 < 0))
        {
            missing("corner extrapolation stencil", level, x, y, z);
        }
        if (layer2 && g.target(level, y + 2 * d.v[1], z + 2 * d.v[2]))
        {
            g.extrap4(s2, sm, s0, s1);
        }
        int nz[3], n_nz = 0;
        for (int k = 0; k < dim; ++k)
        {
            if (d.v[k] != 0)
            {
                nz[n_nz++] = k;
            }
        }
        if (n_nz < 2)
        {
            return;
        }
        const int combos = n_nz == 2 ? 2 : 4; // ghost_width^(n_nz - 1)
        for (int k = 1; k <= 2; ++k)
        {
            if (k == 2 && !layer2)
            {
                continue;
            }
            const int sx = x + k * d.v[0], sy = y + k * d.v[1], sz = z + k * d.v[2];
            for (int combo = 0; combo < combos; ++combo)
            {
                int delta[3] = {0, 0, 0};
                int tmp      = combo;
                for (int p = 1; p < n_nz; ++p)
                {
                    const int gp = tmp % 2;
                    tmp /= 2;
                    delta[nz[p]] += (gp - (k - 1)) * d.v[nz[p]];
                }
                if (delta[0] == 0 && delta[1] == 0 && delta[2] == 0)
                {
                    continue;
                }
                const int tx = sx + delta[0], ty = sy + delta[1], tz = sz + delta[2];
                const int64_t dst = off(tx, ty, tz);
                if (dst < 0 || !g.target(level, ty, tz))
                {
                    continue;
                }
                if (k == 1)
                {
                    g.copy(dst, s1);
                }
                else
                {
                    g.extrap4(dst, sm, s0, s1);
                }
            }
        }
    }

    inline void build_ghost_phase(const Mesh& m, int level, const PlanFilter& flt, PhaseItems& out)
    {
        const MeshConfig& cfg = m.cfg;
        const int dim = cfg.dim, L = cfg.max_level, lmin = cfg.min_level;
        BcBuilder& g          = out.bc;
        g.flt                 = &flt;
        const LevelSet& ref   = m.ref[level];
        const bool have_below = level > 0 && !m.ref[level - 1].empty();
        if (ref.empty() && !have_below)
        {
            return;
        }
        auto touches_periodic = [&](const Dir& d)
        {
            for (int k = 0; k < dim; ++k)
            {
                if (d.v[k] != 0 && cfg.periodic[k])
                {
                    return true;
                }
            }
            return false;
        };
        for (int k = 0; k < dim; ++k)
        {
            if (cfg.periodic[k])
            {
                periodic_items(m, level, k, flt, out.per[k]);
            }
        }
        std::vector<int64_t> extrap_dst;
        if (dim > 1)
        {
            for (const Dir& d : diagonal_directions(dim))
            {
                if (touches_periodic(d))
                {
                    continue; // a periodic direction has no boundary, hence no corner ghost (update_outer_ghost.hpp:352-366)
                }
                if (level >= lmin && level <= L && !ref.empty())
                {
                    // update_outer_corners_by_polynomial_extrapolation, ghost width 1: u[c + d] = u[c]
                    LevelSet cc = set_inter(m.cells[level], corner_cells(m, level, d));
                    for_each_cell(cc,
                                  [&](int x, int y, int z)
                                  {
                                      const int64_t dst = need(ref, "corner ghost", level, y + d.v[1], z + d.v[2], x + d.v[0], x + d.v[0]);
                                      extrap_dst.push_back(dst);
                                      if (g.target(level, y + d.v[1], z + d.v[2]))
                                      {
                                          g.copy(dst, need(ref, "corner cell", level, y, z, x, x));
                                      }
                                      if (cfg.ghost_width() == 2)
                                      {
                                          corner_block_width2(m, level, d, x, y, z, flt, out.bc2);
                                      }
                                  });
                }
            }
            std::sort(extrap_dst.begin(), extrap_dst.end());
            for (const Dir& d : diagonal_directions(dim))
            {
                if (touches_periodic(d))
                {
                    continue;
                }
                // records of project_corner_below(level + 1): written in this phase, after phase level+1 produced their source
                if (level + 1 >= lmin && level + 1 <= L)
                {
                    corner_below(m,
                                 level + 1,
                                 d,
                                 [&](int pl, int y, int z, int64_t dst, int64_t src)
                                 {
                                     if (pl == level && std::binary_search(extrap_dst.begin(), extrap_dst.end(), dst))
                                     {
                                         return; // overwritten by the extrapolation of this level before anything reads it
                                     }
                                     if (g.target(pl, y, z))
                                     {
                                         g.copy(dst, src);
                                     }
                                 });
                }
            }
        }
        for (const Dir& d : cartesian_directions(dim))
        {
            if (touches_periodic(d))
            {
                continue; // update_outer_ghost.hpp:372-375
            }
            if (level < L && !ref.empty())
            {
                // project_bc, layer 1
                // only the boundary layer of the union can leave the domain: clip before translating (the levels are large)
                int blo[3], bhi[3];
                m.domain_box(level, 0, blo, bhi);
                for (int k = 0; k < 3; ++k)
                {
                    blo[k] -= d.v[k];
                    bhi[k] -= d.v[k];
                }
                LevelSet ghosts = set_inter(translate(minus_box(m.uni[level], dim, blo, bhi), d.v[0], d.v[1], d.v[2]), ref);
                locate(ghosts, ref);
                for (size_t r = 0; r < ghosts.rows(); ++r)
                {
                    const int y = key_y(ghosts.key[r]), z = key_z(ghosts.key[r]);
                    if (!g.target(level, y, z))
                    {
                        continue;
                    }
                    for (int q = ghosts.ptr[r]; q < ghosts.ptr[r + 1]; ++q)
                    {
                        for (int x = ghosts.xs[q]; x < ghosts.xe[q]; ++x)
                        {
                            g.begin_avg(ghosts.off[q] + (x - ghosts.xs[q]));
                            for (int dl = 1; dl <= 2; ++dl)
                            {
                                const LevelSet& rf = m.ref[level + dl];
                                const int n        = 1 << dl;
                                for (int cz = 0; cz < (dim > 2 ? n : 1); ++cz)
                                {
                                    for (int cy = 0; cy < (dim > 1 ? n : 1); ++cy)
                                    {
                                        const int row = rf.find_row(mk_key(dim > 1 ? (y << dl) + cy : 0, dim > 2 ? (z << dl) + cz : 0));
                                        if (row < 0)
                                        {
                                            continue;
                                        }
                                        for (int cx = 0; cx < n; ++cx)
                                        {
                                            const int i = rf.find_ivl(row, (x << dl) + cx);
                                            if (i >= 0)
                                            {
                                                g.add_src(rf.off[i] + ((x << dl) + cx - rf.xs[i]));
                                            }
                                        }
                                    }
                                }
                                if (g.items.back().n_src > 0)
                                {
                                    break;
                                }
                            }
                        }
                    }
                }
            }
            if (level >= lmin && !ref.empty())
            {
                // apply_field_bc(level) and, folded in, predict_bc(level + 1)
                LevelSet bl     = boundary_leaves(m, level, d);
                const double dx = cfg.cell_length(level);
                locate(bl, ref);
                LevelSet gh = translate(bl, d.v[0], d.v[1], d.v[2]);
                locate(gh, ref);
                for (size_t r = 0; r < gh.rows(); ++r)
                {
                    if (!g.target(level, key_y(gh.key[r]), key_z(gh.key[r])))
                    {
                        continue;
                    }
                    for (int q = gh.ptr[r]; q < gh.ptr[r + 1]; ++q)
                    {
                        for (int k = 0; k < gh.xe[q] - gh.xs[q]; ++k)
                        {
                            g.value(gh.off[q] + k, bl.off[q] + k, dx);
                        }
                    }
                }
                if (level < L && !bl.empty())
                {
                    LevelSet fine = set_inter(refine(gh, 1, dim), m.ref[level + 1]);
                    locate(fine, m.ref[level + 1]);
                    Probe pl(bl);
                    for (size_t r = 0; r < fine.rows(); ++r)
                    {
                        const int y = key_y(fine.key[r]), z = key_z(fine.key[r]);
                        if (!g.target(level + 1, y, z))
                        {
                            continue;
                        }
                        // the leaf behind the parent ghost: parent - d
                        pl.seek(mk_key((y >> 1) - d.v[1], (z >> 1) - d.v[2]));
                        for (int q = fine.ptr[r]; q < fine.ptr[r + 1]; ++q)
                        {
                            for (int x = fine.xs[q]; x < fine.xe[q]; ++x)
                            {
                                const int lx = (x >> 1) - d.v[0];
                                g.value(fine.off[q] + (x - fine.xs[q]), need(pl, "predict_bc leaf", level, (y >> 1) - d.v[1], (z >> 1) - d.v[2], lx, lx), dx);
                            }
                        }
                    }
                }
            }
        }
        // the boundary condition fills one ghost layer, the second one is extrapolated (update_outer_ghost.hpp:412-423,
        // bc/apply_field_bc.hpp:499-563); these records read the first layer and therefore run one phase later (GhostPhase::bc2)
        if (cfg.ghost_width() == 2 && level >= lmin && !ref.empty())
        {
            out.bc2.flt = &flt;
            for (const Dir& d : cartesian_directions(dim))
            {
                if (touches_periodic(d))
                {
                    continue;
                }
                const LevelSet has1 = translate(ref, -d.v[0], -d.v[1], -d.v[2]);
                const LevelSet has2 = translate(ref, -2 * d.v[0], -2 * d.v[1], -2 * d.v[2]);
                auto emit = [&](const LevelSet& centres)
                {
                    for_each_cell(centres,
                                  [&](int x, int y, int z)
                                  {
                                      const int gx = x + 2 * d.v[0], gy = y + 2 * d.v[1], gz = z + 2 * d.v[2];
                                      if (!out.bc2.target(level, gy, gz))
                                      {
                                          return;
                                      }
                                      out.bc2.extrap4(need(ref, "second ghost", level, gy, gz, gx, gx),
                                                      need(ref, "extrapolation stencil -1", level, y - d.v[1], z - d.v[2], x - d.v[0], x - d.v[0]),
                                                      need(ref, "extrapolation stencil 0", level, y, z, x, x),
                                                      need(ref, "extrapolation stencil +1", level, y + d.v[1], z + d.v[2], x + d.v[0], x + d.v[0]));
                                  });
                };
                // 1. beyond the boundary leaves (apply_extrapolation_bc_cells<4>)
                emit(set_inter(boundary_leaves(m, level, d), has2));
                // 2. beyond the cells of the boundary layer that lie under finer leaves (apply_extrapolation_bc_ghosts<4>)
                int lo[3], hi[3];
                m.domain_box(level, 0, lo, hi);
                for (int k = 0; k < 3; ++k)
                {
                    lo[k] -= d.v[k];
                    hi[k] -= d.v[k];
                }
                LevelSet layer = minus_box(m.in_domain(ref, level), dim, lo, hi); // domain \ translate(domain, -d)
                LevelSet cand  = set_inter(set_inter(set_inter(layer, has2), has1), m.uni[level]);
                emit(set_diff(cand, m.cells[level]));
            }
        }
        if (level > 0 && !ref.empty())
        {
            // projection targets: proj_cells[level - 1] ∩ coarsen(reference[level]) (update_ghost_mr.hpp:219) == proj_cells[level - 1]:
            // the mesh construction adds the children of every projection cell to the reference (mr/mesh.hpp:415-452)
            set_seeds(m.proj[level - 1], level - 1, flt, false, out.proj);
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

    // the rows [row_begin, row_end) of ref[level] of prediction_set(m, level): every operation of prediction_set is row-local (the
    // parent test needs the coarse rows (y >> 1, z >> 1) only), so the largest levels are cut into row chunks that run as separate
    // tasks of build_plan instead of one task that was its critical path
    inline LevelSet prediction_set_rows(const Mesh& m, int level, size_t row_begin, size_t row_end)
    {
        const int dim       = m.cfg.dim;
        const LevelSet& ref = m.ref[level];
        if (level > m.cfg.max_level || row_begin >= row_end)
        {
            return LevelSet();
        }
        const int64_t klo = ref.key[row_begin];
        const bool last   = row_end >= ref.rows();
        const int64_t khi = last ? 0 : ref.key[row_end];
        auto rows_of = [&](const LevelSet& s)
        {
            const size_t r0 = static_cast<size_t>(std::lower_bound(s.key.begin(), s.key.end(), klo) - s.key.begin());
            const size_t r1 = last ? s.rows() : static_cast<size_t>(std::lower_bound(s.key.begin(), s.key.end(), khi) - s.key.begin());
            return slice_rows(s, r0, r1);
        };
        LevelSet pg = m.in_domain(set_diff(slice_rows(ref, row_begin, std::min(row_end, ref.rows())), set_union(rows_of(m.cells[level]), rows_of(m.proj[level]))), level);
        if (pg.empty())
        {
            return pg;
        }
        // parent rows of the chunk: a contiguous key range of the coarse level (rows are sorted by (z, y))
        const LevelSet& coarse = m.ref[level - 1];
        size_t c0 = 0, c1 = coarse.rows();
        if (dim == 2)
        {
            const int y0 = key_y(pg.key.front()) >> 1, y1 = key_y(pg.key.back()) >> 1;
            c0 = static_cast<size_t>(std::partition_point(coarse.key.begin(), coarse.key.end(), [y0](int64_t k) { return key_y(k) < y0; }) - coarse.key.begin());
            c1 = static_cast<size_t>(std::partition_point(coarse.key.begin(), coarse.key.end(), [y1](int64_t k) { return key_y(k) <= y1; }) - coarse.key.begin());
        }
        else if (dim == 3)
        {
            const int z0 = key_z(pg.key.front()) >> 1, z1 = key_z(pg.key.back()) >> 1;
            c0 = static_cast<size_t>(std::partition_point(coarse.key.begin(), coarse.key.end(), [z0](int64_t k) { return key_z(k) < z0; }) - coarse.key.begin());
            c1 = static_cast<size_t>(std::partition_point(coarse.key.begin(), coarse.key.end(), [z1](int64_t k) { return key_z(k) <= z1; }) - coarse.key.begin());
        }
        return set_inter(pg, refine(slice_rows(coarse, c0, c1), 1, dim));
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

    inline void build_plan(const Mesh& m, MeshPlan& plan, const PlanFilter& flt = PlanFilter())
    {
        const MeshConfig& cfg = m.cfg;
        const int dim = cfg.dim, L = cfg.max_level, lmin = cfg.min_level;
        const int nlev = m.nlev;
        plan.arena.clear();
        // Two passes of independent tasks.  Pass A, one task per (kind, level): the set algebra of the tag / detail /
        // prediction subsets and the whole ghost phase of a level.  Pass B, one task per (kind, level, row chunk): the
        // traversal of a subset (or of the leaves) into records.  Chunks keep the finest levels, which hold most of the
        // cells, from being the critical path.
        std::vector<PhaseItems> phases(nlev);
        std::vector<LevelSet> tagset(nlev), detailset(nlev), predset(nlev);
        std::string error;
#ifdef SMR_PLAN_TIMING
        const double tt0 = omp_get_wtime();
#endif
        // prediction sets of the big levels: row chunks as extra tasks (task index >= 4 * nlev)
        struct PredPart
        {
            int level;
            size_t r0, r1;
            LevelSet res;
        };
        std::vector<PredPart> pred_parts;
        std::vector<char> pred_split(static_cast<size_t>(nlev), 0);
        for (int level = 1; level <= L && level < nlev; ++level)
        {
            if (m.ref[level].n_intervals() >= 8000)
            {
                const std::vector<size_t> cut = chunk_rows(m.ref[level], 4000, 8);
                if (cut.size() > 2)
                {
                    pred_split[static_cast<size_t>(level)] = 1;
                    for (size_t c = 0; c + 1 < cut.size(); ++c)
                    {
                        pred_parts.push_back(PredPart{level, cut[c], cut[c + 1], {}});
                    }
                }
            }
        }
        const int n_extra = static_cast<int>(pred_parts.size());
#pragma omp parallel for schedule(dynamic, 1)
        for (int t = 4 * nlev + n_extra - 1; t >= 0; --t)
        {
            if (t >= 4 * nlev)
            {
                PredPart& pp = pred_parts[static_cast<size_t>(t - 4 * nlev)];
                try
                {
                    pp.res = prediction_set_rows(m, pp.level, pp.r0, pp.r1);
                }
                catch (const std::exception& e)
                {
#pragma omp critical
                    error = e.what();
                }
                continue;
            }
            const int kind = t / nlev, level = t % nlev;
            try
            {
                switch (kind)
                {
                    case 3:
                        if (level <= L)
                        {
                            build_ghost_phase(m, level, flt, phases[level]);
                        }
                        break;
                    case 2:
                        if (level >= 1 && level <= L && !pred_split[static_cast<size_t>(level)])
                        {
                            predset[level] = prediction_set(m, level);
                        }
                        break;
                    case 1:
                        if (lmin != L && level >= std::max(lmin - 1, 0) && level < L)
                        {
                            detailset[level] = detail_set(m, level);
                        }
                        break;
                    default:
                        if (lmin != L && level >= std::max(lmin, 1) && level <= L)
                        {
                            tagset[level] = tag_set(m, level);
                        }
                        break;
                }
            }
            catch (const std::exception& e)
            {
#pragma omp critical
                error = e.what();
            }
        }
        if (!error.empty())
        {
            throw std::out_of_range(error);
        }
        for (int level = 1; level < nlev; ++level)
        {
            if (pred_split[static_cast<size_t>(level)])
            {
                std::vector<const LevelSet*> parts;
                for (const PredPart& pp : pred_parts)
                {
                    if (pp.level == level)
                    {
                        parts.push_back(&pp.res);
                    }
                }
                predset[level] = union_all(parts); // disjoint, ascending key ranges: bulk copies
#ifdef SMR_CHECK_SPLIT
                if (!predset[level].same_cells(prediction_set(m, level)))
                {
                    throw std::logic_error("prediction_set_rows differs from prediction_set");
                }
#endif
            }
        }
        struct Chunk
        {
            int kind, level;
            size_t r0, r1;
            std::vector<smr_seed> seeds, strips; // strips: kind 4 only (seeds = the single-row remainder)
        };
        std::vector<Chunk> chunks;
        // chunk lists per kind, levels ascending, so that the concatenation below keeps level order
        for (int kind = 0; kind < 5; ++kind)
        {
            for (int level = 0; level < nlev; ++level)
            {
                const LevelSet& src = kind == 0 ? tagset[level] : (kind == 1 ? detailset[level] : (kind == 2 ? predset[level] : m.cells[level]));
                if (src.empty())
                {
                    continue;
                }
                const std::vector<size_t> cut = chunk_rows(src, 2500, 16);
                for (size_t c = 0; c + 1 < cut.size(); ++c)
                {
                    chunks.push_back(Chunk{kind, level, cut[c], cut[c + 1], {}, {}});
                }
            }
        }
#ifdef SMR_PLAN_TIMING
        const double tt05 = omp_get_wtime();
#endif
#pragma omp parallel for schedule(dynamic, 1)
        for (int t = static_cast<int>(chunks.size()) - 1; t >= 0; --t)
        {
            Chunk& ck = chunks[static_cast<size_t>(t)];
            try
            {
                switch (ck.kind)
                {
                    case 4:
                        fv_split_seeds(m, ck.level, flt, ck.strips, ck.seeds, ck.r0, ck.r1);
                        break;
                    case 3:
                        fv_seeds(m, ck.level, flt, ck.seeds, ck.r0, ck.r1);
                        break;
                    case 2:
                        set_seeds(predset[ck.level], ck.level, flt, false, ck.seeds, ck.r0, ck.r1);
                        break;
                    case 1:
                        set_seeds(detailset[ck.level], ck.level, flt, false, ck.seeds, ck.r0, ck.r1);
                        break;
                    default:
                        // tagset[l] lives at the coarse level l - 1; tags are replicated on every rank
                        set_seeds(tagset[ck.level], ck.level - 1, flt, true, ck.seeds, ck.r0, ck.r1);
                        break;
                }
            }
            catch (const std::exception& e)
            {
#pragma omp critical
                error = e.what();
            }
        }
        if (!error.empty())
        {
            throw std::out_of_range(error);
        }
#ifdef SMR_PLAN_TIMING
        const double tt1 = omp_get_wtime();
        std::printf("  build_plan: sets+phases %.2f ms, %zu chunks %.2f ms\n", (tt05 - tt0) * 1e3, chunks.size(), (tt1 - tt05) * 1e3);
#endif
        // units of every chunk's seed lists, summed in parallel
        std::vector<int64_t> chunk_units(2 * chunks.size(), 0);
#pragma omp parallel for schedule(static)
        for (int t = 0; t < static_cast<int>(chunks.size()); ++t)
        {
            int64_t a = 0, b = 0;
            for (const smr_seed& sd : chunks[static_cast<size_t>(t)].seeds)
            {
                a += sd.n;
            }
            for (const smr_seed& sd : chunks[static_cast<size_t>(t)].strips)
            {
                b += sd.n;
            }
            chunk_units[2 * static_cast<size_t>(t)]     = a;
            chunk_units[2 * static_cast<size_t>(t) + 1] = b;
        }
        SeedSums seed_sums;
        seed_sums.reserve(2 * chunks.size());
        for (size_t t = 0; t < chunks.size(); ++t)
        {
            seed_sums.emplace(&chunks[t].seeds, chunk_units[2 * t]);
            seed_sums.emplace(&chunks[t].strips, chunk_units[2 * t + 1]);
        }
        plan.down.assign(nlev, GhostPhase());
        plan.pred.assign(nlev, Batch());
        plan.tag.assign(nlev, Batch());
        // layout (serial, cheap) then fill (parallel) straight into the staging arena
        plan.derive.clear();
        PendingSeeds p_fv        = pending_seeds(&plan.fv, B_FV, SMR_DERIVE_FV, -1, sizeof(smr_item_fv));
        // keep_boundary_refined (mr/adapt.hpp:245-274): cells[max_level] minus translate(domain, -w * direction) for every Cartesian
        // direction (boundary.hpp:6-22) = the leaves outside the domain shrunk by w on every side; tag = keep for all of them
        std::vector<smr_seed> bdry_seeds;
        PendingSeeds p_bdry = pending_seeds(&plan.keep_bdry, B_FV, SMR_DERIVE_FV, -1, sizeof(smr_item_fv));
        if (m.cfg.refine_boundary && m.cfg.max_level < nlev && !m.cells[m.cfg.max_level].empty())
        {
            const int Lb = m.cfg.max_level, w = m.cfg.max_stencil_radius;
            int blo[3], bhi[3];
            m.domain_box(Lb, 0, blo, bhi);
            for (int k = 0; k < m.cfg.dim; ++k)
            {
                blo[k] += w;
                bhi[k] -= w;
            }
            set_seeds(minus_box(m.cells[Lb], m.cfg.dim, blo, bhi), Lb, flt, false, bdry_seeds);
        }
        p_bdry.parts.push_back(&bdry_seeds);
        PendingSeeds p_fv_single = pending_seeds(&plan.fv_single, B_FV, SMR_DERIVE_FV, -1, sizeof(smr_item_fv));
        PendingSeeds p_fv_strip  = pending_seeds(&plan.fv_strip, B_FV, SMR_DERIVE_FVSTRIP, -1, sizeof(smr_item_fvstrip), SMR_CTA_THREADS * STRIP_UPT);
        PendingSeeds p_detail    = pending_seeds(&plan.detail, B_DETAIL, SMR_DERIVE_DETAIL, -1, sizeof(smr_item_detail));
        PendingSeeds p_tag_all   = pending_seeds(&plan.tag_all, B_TAG, SMR_DERIVE_TAG, -1, sizeof(smr_item_tag));
        p_detail.cum             = &plan.detail_cum;
        p_tag_all.cum            = &plan.tag_cum;
        p_tag_all.inclusive      = true;
        p_detail.n_groups = p_tag_all.n_groups = nlev; // cumulative counts are indexed by level
        std::vector<PendingSeeds> p_proj(nlev), p_pred(nlev), p_tag(nlev);
        std::vector<PendingBc> p_bc(nlev), p_bc2(nlev);
        std::vector<Pending<smr_item_copy>> p_per(static_cast<size_t>(3 * nlev));
        for (int l = 0; l < nlev; ++l)
        {
            for (int k = 0; k < 3; ++k)
            {
                p_per[static_cast<size_t>(3 * l + k)] = Pending<smr_item_copy>{&plan.down[l].per[k], B_COPY, l, {&phases[l].per[k]}, nullptr, false};
            }
            p_proj[l] = pending_seeds(&plan.down[l].proj, B_PROJ, SMR_DERIVE_PROJ, l, sizeof(smr_item_proj));
            p_proj[l].parts.push_back(&phases[l].proj);
            p_pred[l] = pending_seeds(&plan.pred[l], B_PRED, SMR_DERIVE_PRED, l, sizeof(smr_item_pred));
            p_tag[l]  = pending_seeds(&plan.tag[l], B_TAG, SMR_DERIVE_TAG, l, sizeof(smr_item_tag));
            p_bc[l]   = PendingBc{&plan.down[l].bc, l, &phases[l].bc.items, &phases[l].bc.srcs};
            p_bc2[l]  = PendingBc{&plan.down[l].bc2, l, &phases[l].bc2.items, &phases[l].bc2.srcs};
        }
        for (const Chunk& ck : chunks)
        {
            switch (ck.kind)
            {
                case 4:
                    p_fv_single.parts.push_back(&ck.seeds);
                    p_fv_strip.parts.push_back(&ck.strips);
                    break;
                case 3:
                    p_fv.parts.push_back(&ck.seeds);
                    break;
                case 2:
                    p_pred[ck.level].parts.push_back(&ck.seeds);
                    break;
                case 1:
                    p_detail.parts.push_back(&ck.seeds);
                    p_detail.group.push_back(ck.level);
                    break;
                default:
                    if (p_tag[ck.level].parts.empty())
                    {
                        p_tag[ck.level].alias       = &p_tag_all;
                        p_tag[ck.level].alias_part0 = p_tag_all.parts.size();
                    }
                    p_tag_all.parts.push_back(&ck.seeds);
                    p_tag_all.group.push_back(ck.level);
                    p_tag[ck.level].parts.push_back(&ck.seeds);
                    break;
            }
        }
        layout_seeds(p_fv, plan.arena, plan.derive, &seed_sums);
        layout_seeds(p_bdry, plan.arena, plan.derive, &seed_sums);
        layout_seeds(p_fv_single, plan.arena, plan.derive, &seed_sums);
        layout_seeds(p_fv_strip, plan.arena, plan.derive, &seed_sums);
        plan.fv_strip_cells = plan.fv_strip.n_cells * SMR_STRIP_ROWS;
        layout_seeds(p_detail, plan.arena, plan.derive, &seed_sums);
        layout_seeds(p_tag_all, plan.arena, plan.derive, &seed_sums);
        for (int l = 0; l < nlev; ++l)
        {
            layout_seeds(p_proj[l], plan.arena, plan.derive, &seed_sums);
            layout_seeds(p_pred[l], plan.arena, plan.derive, &seed_sums);
            layout_seeds(p_tag[l], plan.arena, plan.derive, &seed_sums);
            layout_bc(p_bc[l], plan.arena);
            layout_bc(p_bc2[l], plan.arena);
            for (int k = 0; k < 3; ++k)
            {
                layout_batch(p_per[static_cast<size_t>(3 * l + k)], plan.arena);
            }
        }
        close_derive(plan.arena, plan.derive);
        {
            // every (batch, part) pair is an independent copy; the per-CTA tables follow once the parts are in
            std::vector<std::function<void()>> jobs, finish;
            auto add = [&](PendingSeeds& pd)
            {
                for (size_t k = 0; k < pd.parts.size(); ++k)
                {
                    if (!pd.parts[k]->empty())
                    {
                        jobs.push_back([&pd, &plan, k] { fill_seeds_part(pd, plan.arena, k); });
                    }
                }
                finish.push_back([&pd, &plan] { finish_seeds(pd, plan.arena); });
            };
            add(p_fv_single);
            add(p_fv_strip);
            add(p_fv);
            add(p_bdry);
            add(p_detail);
            add(p_tag_all);
            for (int l = 0; l < nlev; ++l)
            {
                add(p_proj[l]);
                add(p_pred[l]);
                add(p_tag[l]);
                PendingBc* pb = &p_bc[l];
                jobs.push_back([pb, &plan] { fill_bc(*pb, plan.arena); });
                PendingBc* pb2 = &p_bc2[l];
                jobs.push_back([pb2, &plan] { fill_bc(*pb2, plan.arena); });
                for (int k = 0; k < 3; ++k)
                {
                    Pending<smr_item_copy>* pp = &p_per[static_cast<size_t>(3 * l + k)];
                    if (!phases[l].per[k].empty())
                    {
                        jobs.push_back([pp, &plan] { fill_batch(*pp, plan.arena); });
                    }
                }
            }
#pragma omp parallel for schedule(dynamic, 1)
            for (int t = 0; t < static_cast<int>(jobs.size()); ++t)
            {
                jobs[static_cast<size_t>(t)]();
            }
#pragma omp parallel for schedule(dynamic, 1)
            for (int t = 0; t < static_cast<int>(finish.size()); ++t)
            {
                finish[static_cast<size_t>(t)]();
            }
        }
#ifdef SMR_PLAN_TIMING
        std::printf("  build_plan: tasks %.2f ms, serial finish %.2f ms\n", (tt1 - tt0) * 1e3, (omp_get_wtime() - tt1) * 1e3);
#endif
    }

    // old mesh -> new mesh field transfer (update_fields): copy, projection, prediction batches
    struct TransferPlan
    {
        Arena arena;
        Batch copy, proj, pred;
        DeriveList derive; // copy: dst in the new mesh, src in the old one; proj / pred likewise (empty for build_broadcast)
    };

    inline void build_transfer(const Mesh& old_m, const Mesh& new_m, TransferPlan& tp, const PlanFilter& flt = PlanFilter())
    {
        const MeshConfig& cfg = old_m.cfg;
        const int dim = cfg.dim;
        const int nlev = old_m.nlev;
        tp.arena.clear();
        // tasks: per level the projection + prediction sets (small: only where the mesh changed) and, per row chunk of the
        // new leaves, the copy records (the bulk)
        struct Task
        {
            int kind, level;
            size_t r0, r1;
            std::vector<smr_seed> copies, projs, preds;
        };
        std::vector<Task> tasks;
        for (int l = cfg.min_level; l <= cfg.max_level && l < nlev; ++l)
        {
            if (!new_m.cells[l].empty() && !old_m.ref[l].empty())
            {
                const std::vector<size_t> cut = chunk_rows(new_m.cells[l], 3000, 16);
                for (size_t c = 0; c + 1 < cut.size(); ++c)
                {
                    tasks.push_back(Task{0, l, cut[c], cut[c + 1], {}, {}, {}});
                }
            }
            if (l > cfg.min_level)
            {
                tasks.push_back(Task{1, l, 0, 0, {}, {}, {}});
            }
        }
        std::string error;
#pragma omp parallel for schedule(dynamic, 1)
        for (int t = static_cast<int>(tasks.size()) - 1; t >= 0; --t)
        {
            Task& tk    = tasks[static_cast<size_t>(t)];
            const int l = tk.level;
            try
            {
                if (tk.kind == 0)
                {
                    LevelSet s = set_inter(old_m.ref[l], slice_rows(new_m.cells[l], tk.r0, tk.r1));
                    set_seeds(s, l, flt, false, tk.copies);
                }
                else
                {
                    LevelSet sc = set_inter(coarsen(old_m.cells[l], 1, dim), new_m.cells[l - 1]);
                    set_seeds(sc, l - 1, flt, false, tk.projs);
                    // set_refine = (new cells[l] ∩ old cells[l-1]).on(l-1); every coarse cell fills all its children
                    LevelSet sr = set_inter(coarsen(new_m.cells[l], 1, dim), old_m.cells[l - 1]);
                    if (!sr.empty())
                    {
                        set_seeds(refine(sr, 1, dim), l, flt, false, tk.preds);
                    }
                }
            }
            catch (const std::exception& e)
            {
#pragma omp critical
                error = e.what();
            }
        }
        if (!error.empty())
        {
            throw std::out_of_range(error);
        }
        tp.derive.clear();
        PendingSeeds p_copy = pending_seeds(&tp.copy, B_COPY, SMR_DERIVE_COPY, -1, sizeof(smr_item_copy));
        PendingSeeds p_proj = pending_seeds(&tp.proj, B_PROJ, SMR_DERIVE_PROJ, -1, sizeof(smr_item_proj));
        PendingSeeds p_pred = pending_seeds(&tp.pred, B_PRED, SMR_DERIVE_PRED, -1, sizeof(smr_item_pred));
        for (const Task& tk : tasks)
        {
            if (tk.kind == 0)
            {
                p_copy.parts.push_back(&tk.copies);
            }
            else
            {
                p_proj.parts.push_back(&tk.projs);
                p_pred.parts.push_back(&tk.preds);
            }
        }
        layout_seeds(p_copy, tp.arena, tp.derive);
        layout_seeds(p_proj, tp.arena, tp.derive);
        layout_seeds(p_pred, tp.arena, tp.derive);
        close_derive(tp.arena, tp.derive);
#pragma omp parallel sections
        {
#pragma omp section
            fill_seeds(p_copy, tp.arena);
#pragma omp section
            fill_seeds(p_proj, tp.arena);
#pragma omp section
            fill_seeds(p_pred, tp.arena);
        }
    }

    // Records that re-store every reference cell this rank owns to ALL peers (CopyOp with src == dst): after it every
    // rank holds the complete field, which is what a re-cut of the slabs (and a host download) needs.
    inline void build_broadcast(const Mesh& m, const PlanFilter& flt, TransferPlan& tp)
    {
        tp.arena.clear();
        tp.derive.clear();
        std::vector<std::vector<smr_item_copy>> copies(1);
        const int mask = static_cast<int>(flt.mask_all());
        for (int l = 0; l < m.nlev; ++l)
        {
            const LevelSet& r = m.ref[l];
            for (size_t row = 0; row < r.rows(); ++row)
            {
                if (!flt.owns(l, key_y(r.key[row]), key_z(r.key[row])))
                {
                    continue;
                }
                for (int q = r.ptr[row]; q < r.ptr[row + 1]; ++q)
                {
                    copies[0].push_back({r.off[q], r.off[q], r.xe[q] - r.xs[q], mask});
                }
            }
        }
        Pending<smr_item_copy> p_copy{&tp.copy, B_COPY, -1, {&copies[0]}, nullptr, false};
        tp.proj = Batch();
        tp.pred = Batch();
        layout_batch(p_copy, tp.arena);
        tp.arena.commit();
        fill_batch(p_copy, tp.arena);
    }

    // The reference sub-mesh as the device sees it (items.h: smr_csr_table): per level the CSR arrays of Mesh::ref, back to
    // back in one staging buffer.  16 B per interval + 12 B per row: 1-2 MB for a 10^6-cell adapted mesh.
    struct CsrImage
    {
        Arena arena;
        smr_csr_table tab;

        CsrImage()
        {
            arena.min_cap = size_t(8) << 20;
        }
    };

    inline void build_csr(const Mesh& m, CsrImage& img)
    {
        img.arena.clear();
        std::memset(&img.tab, 0, sizeof(img.tab));
        if (m.nlev > SMR_MAX_LEVELS)
        {
            throw std::invalid_argument("more levels than SMR_MAX_LEVELS");
        }
        for (int l = 0; l < m.nlev; ++l)
        {
            const LevelSet& r = m.ref[l];
            if (r.empty())
            {
                continue;
            }
            smr_csr_level& lv = img.tab.lv[l];
            lv.rows = static_cast<int32_t>(r.rows());
            lv.key  = static_cast<int64_t>(img.arena.take(r.rows() * sizeof(int64_t)));
            lv.ptr  = static_cast<int64_t>(img.arena.take((r.rows() + 1) * sizeof(int32_t)));
            lv.xs   = static_cast<int64_t>(img.arena.take(r.n_intervals() * sizeof(int32_t)));
            lv.xe   = static_cast<int64_t>(img.arena.take(r.n_intervals() * sizeof(int32_t)));
            lv.off  = static_cast<int64_t>(img.arena.take(r.n_intervals() * sizeof(int64_t)));
        }
        img.arena.commit();
#pragma omp parallel for schedule(dynamic, 1)
        for (int l = m.nlev - 1; l >= 0; --l)
        {
            const LevelSet& r       = m.ref[l];
            const smr_csr_level& lv = img.tab.lv[l];
            if (lv.rows == 0)
            {
                continue;
            }
            std::memcpy(img.arena.p + lv.key, r.key.data(), r.rows() * sizeof(int64_t));
            std::memcpy(img.arena.p + lv.ptr, r.ptr.data(), (r.rows() + 1) * sizeof(int32_t));
            std::memcpy(img.arena.p + lv.xs, r.xs.data(), r.n_intervals() * sizeof(int32_t));
            std::memcpy(img.arena.p + lv.xe, r.xe.data(), r.n_intervals() * sizeof(int32_t));
            std::memcpy(img.arena.p + lv.off, r.off.data(), r.n_intervals() * sizeof(int64_t));
        }
    }
} // namespace smr
