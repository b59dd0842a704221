// C ABI of samurai_b200 (include/samurai_b200.h): host mesh/plan management + kernel launches on one CUDA stream.
// No CPU fallback: every compute entry point requires a device and fails loudly otherwise.
#include "../../include/samurai_b200.h"
#include "batches.hpp"
#define SMR_MISC_KERNELS
#include "launch.hpp"

#include <chrono>
#include <cmath>
#include <memory>
#include <string>
#include <unordered_map>

namespace smr
{
    struct CudaError : std::runtime_error
    {
        using std::runtime_error::runtime_error;
    };

#define SMR_CUDA(call)                                                                                                    \
    do                                                                                                                    \
    {                                                                                                                     \
        cudaError_t e_ = (call);                                                                                          \
        if (e_ != cudaSuccess)                                                                                            \
        {                                                                                                                 \
            throw CudaError(std::string(#call) + ": " + cudaGetErrorString(e_));                                         \
        }                                                                                                                 \
    } while (0)

    // Per-rank device pool for every buffer peers write into (fields, detail, tags).  All ranks issue the same sequence of
    // allocations with the same sizes, so a buffer sits at the same pool offset on every rank and one address delta per
    // peer reaches all of them (kernels.cuh: PeerTable).  First fit, deterministic.
    struct Pool
    {
        char* base  = nullptr;
        size_t size = 0;
        std::vector<std::pair<size_t, size_t>> free_list; // (offset, bytes), sorted by offset

        void init(char* b, size_t n, size_t reserved)
        {
            base = b;
            size = n;
            free_list.assign(1, {reserved, n - reserved});
        }

        void* alloc(size_t n)
        {
            n = (n + 255) & ~size_t(255);
            for (size_t i = 0; i < free_list.size(); ++i)
            {
                if (free_list[i].second >= n)
                {
                    const size_t off = free_list[i].first;
                    free_list[i].first += n;
                    free_list[i].second -= n;
                    if (free_list[i].second == 0)
                    {
                        free_list.erase(free_list.begin() + static_cast<std::ptrdiff_t>(i));
                    }
                    return base + off;
                }
            }
            return nullptr;
        }

        void release(void* p, size_t n)
        {
            n                = (n + 255) & ~size_t(255);
            const size_t off = static_cast<size_t>(static_cast<char*>(p) - base);
            auto it          = std::lower_bound(free_list.begin(), free_list.end(), std::make_pair(off, size_t(0)));
            it               = free_list.insert(it, {off, n});
            if (it + 1 != free_list.end() && it->first + it->second == (it + 1)->first)
            {
                it->second += (it + 1)->second;
                free_list.erase(it + 1);
            }
            if (it != free_list.begin() && (it - 1)->first + (it - 1)->second == it->first)
            {
                (it - 1)->second += it->second;
                free_list.erase(it);
            }
        }
    };

    static Pool g_pool;
    static bool g_pool_on = false;

    struct DevBuf
    {
        void* p     = nullptr;
        size_t cap  = 0;
        bool shared = false; // peers store into it: must live in the pool when running multi-GPU
        bool pooled = false;
        size_t min_cap = 0; // first allocation at least this big (index-batch buffers: a regrowth is a cudaFree + cudaMalloc stall)

        void ensure(size_t bytes)
        {
            if (bytes > cap)
            {
                release();
                // every regrowth is a cudaFree + cudaMalloc (a device synchronisation and up to milliseconds): grow
                // geometrically, generously outside the fixed multi-GPU pool
                // (measured: smr_field_resize spent 90-490 ms in cudaFree + cudaMalloc when a 5-12 MB field buffer regrew).
                // Outside the pool: at least 64 MB (or min_cap) and double, so steady-state runs never reallocate.
                size_t want = (shared && g_pool_on) ? bytes + bytes / 4 + 4096 : std::max(2 * bytes + 4096, std::max(min_cap, size_t(64) << 20));
                if (shared && g_pool_on)
                {
                    want = (want + 255) & ~size_t(255);
                    p    = g_pool.alloc(want);
                    if (!p)
                    {
                        throw CudaError("multi-GPU pool exhausted: pass a larger pool_bytes to smr_mg_init");
                    }
                    pooled = true;
                }
                else
                {
                    SMR_CUDA(cudaMalloc(&p, want));
                    pooled = false;
                }
                cap = want;
            }
        }

        void release()
        {
            if (p)
            {
                if (pooled)
                {
                    g_pool.release(p, cap);
                }
                else
                {
                    cudaFree(p);
                }
            }
            p   = nullptr;
            cap = 0;
        }

        ~DevBuf()
        {
            release();
        }

        DevBuf()                         = default;
        DevBuf(const DevBuf&)            = delete;
        DevBuf& operator=(const DevBuf&) = delete;

        void swap(DevBuf& o)
        {
            std::swap(p, o.p);
            std::swap(cap, o.cap);
            std::swap(pooled, o.pooled);
        }
    };

    struct PinnedBuf
    {
        void* p    = nullptr;
        size_t cap = 0;

        void ensure(size_t bytes)
        {
            if (bytes > cap)
            {
                if (p)
                {
                    cudaFreeHost(p);
                }
                size_t want = std::max<size_t>(2 * bytes + 4096, size_t(8) << 20); // pinning is slow (~300 ms per regrowth measured): regrow rarely
                SMR_CUDA(cudaHostAlloc(&p, want, cudaHostAllocDefault));
                cap = want;
            }
        }

        ~PinnedBuf()
        {
            if (p)
            {
                cudaFreeHost(p);
            }
        }
    };

    struct FieldObj;

    struct MeshObj
    {
        Mesh mesh;
        MeshPlan plan;
        bool plan_ready = false;
        double plan_seconds = 0;
        DevBuf d_arena;
        FluxPlan flux;  // flux-based schemes on multi-level meshes, built on first use
        FluxPlan fluxw; // the same for six-cell line stencils (WENO5)
        DevBuf d_flux, d_fluxw, d_fluxtab;

        void invalidate_plans();
        DevBuf d_detail, d_tag, d_relmax;
        PlanFilter filter; // multi-GPU slab ownership for this mesh (identity when world == 1)

        MeshObj()
        {
            d_arena.min_cap = size_t(48) << 20;
            d_csr.min_cap = d_csr_old.min_cap = size_t(8) << 20;
            d_detail.shared = true;
            d_tag.shared    = true;
            d_relmax.shared = true;
        }
        // the reference sub-mesh on the device (items.h: smr_csr_table), read by derive_kernel; the previous mesh's copy is
        // kept for the field transfer old -> new
        CsrImage h_csr[2];
        cudaEvent_t h_csr_done[2] = {nullptr, nullptr};
        int h_csr_next  = 0;
        DevBuf d_csr, d_csr_old;
        smr_csr_table csr_tab{}, csr_tab_old{};
        bool csr_ready = false;

        void retire_csr() // the mesh is about to be replaced
        {
            d_csr.swap(d_csr_old);
            csr_tab_old = csr_tab;
            csr_ready   = false;
        }

        ~MeshObj()
        {
            for (cudaEvent_t e : h_csr_done)
            {
                if (e != nullptr)
                {
                    cudaEventDestroy(e);
                }
            }
        }

        PinnedBuf h_tag;
        bool h_tag_valid    = false; // h_tag holds the tags of the last harten iteration (else they are still in d_tag only)
        int64_t last_size   = 0;
        int last_ncomp      = 0;
        bool graduated      = false; // true once the leaves are known to be a fixed point of make_graduation
        std::vector<FieldObj*> fields;
    };

    struct FieldObj
    {
        MeshObj* mesh = nullptr;
        std::string name;
        DevBuf data;
        DevBuf spare; // second buffer reused by the field transfer so adaptation never calls cudaMalloc in steady state
        int64_t n    = 0;
        int bc_type  = -1;
        double bc_value = 0;
        bool ghosts_valid = false; // Field::ghosts_updated() (field/field_base.hpp:264-274)

        FieldObj()
        {
            data.shared  = true;
            spare.shared = true;
        }
    };

    struct Ctx
    {
        bool device         = false;
        int dev             = -1;
        cudaStream_t stream = nullptr;
        std::string err;
        smr_stats stats{};
        uint64_t next_id = 1;
        bool profile     = false;
        double prof_seconds[SMR_FAM_COUNT]  = {};
        uint64_t prof_launches[SMR_FAM_COUNT] = {};
        uint64_t prof_cells[SMR_FAM_COUNT]  = {};
        cudaEvent_t prof_a = nullptr, prof_b = nullptr;
        std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending; // device-time sections not yet resolved
        std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pool;
        DevBuf d_transfer;
        TransferPlan transfer; // reused so its pinned arena is allocated once
        uint64_t prof_bytes[SMR_FAM_COUNT] = {};
        // fused wavefront: phase / job tables go through a ring of pinned staging slots
        bool fuse          = true;
        int wf_grid        = 0;
        void* wf_host      = nullptr;
        void* wf_dev       = nullptr;
        unsigned* wf_barrier   = nullptr;
        unsigned wf_barrier_at = 0;
        unsigned wf_release_at = 0; // counter[1]: multi-GPU release word (kernels.cuh: wf_mg_barrier)
        unsigned* wf_error_host = nullptr; // pinned copy of the wavefront error word (kernels.cuh: SMR_WF_ERROR_WORD)
        int wf_next        = 0;
        cudaEvent_t wf_done[16] = {};
        unsigned* derive_error      = nullptr; // device: sticky error words of derive_kernel (derive.cuh)
        unsigned* derive_error_host = nullptr; // pinned copy
        // multi-GPU
        int mg_rank = 0, mg_world = 1;
        void* mg_pool            = nullptr;
        size_t mg_pool_bytes     = 0;
        unsigned long long epoch = 0;
        void* mg_peer_base[SMR_MAX_RANKS] = {};
        bool mg_connected        = false;
        std::unordered_map<uint64_t, std::unique_ptr<MeshObj>> meshes;
        std::unordered_map<uint64_t, std::unique_ptr<FieldObj>> fields;
    };

    static Ctx g;

    // every kernel translation unit holds its own copy of the __constant__ peer table (kernels.cuh: g_peers)
    static std::vector<PeerSetter>& peer_setters()
    {
        static std::vector<PeerSetter> v;
        return v;
    }

    void register_peer_setter(PeerSetter f)
    {
        peer_setters().push_back(f);
    }

    cudaError_t set_peer_table(const PeerTable& t)
    {
        cudaError_t e = cudaMemcpyToSymbol(g_peers, &t, sizeof(t)); // this unit's copy (mg_barrier_kernel, publish_slot_kernel)
        for (PeerSetter f : peer_setters())
        {
            if (e == cudaSuccess)
            {
                e = f(t);
            }
        }
        return e;
    }

    static double now()
    {
        return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    }

    static void require_device()
    {
        if (!g.device)
        {
            throw CudaError("no CUDA device: samurai_b200 has no CPU fallback for compute entry points (call smr_init(device >= 0))");
        }
    }

    // ---- device-time sections: event pairs around every contiguous stretch of device work -----------------------
    struct Section
    {
        std::pair<cudaEvent_t, cudaEvent_t> ev{nullptr, nullptr};

        Section()
        {
            if (!g.device)
            {
                return;
            }
            if (g.pool.empty())
            {
                SMR_CUDA(cudaEventCreate(&ev.first));
                SMR_CUDA(cudaEventCreate(&ev.second));
            }
            else
            {
                ev = g.pool.back();
                g.pool.pop_back();
            }
            SMR_CUDA(cudaEventRecord(ev.first, g.stream));
        }

        void close()
        {
            if (ev.first)
            {
                cudaEventRecord(ev.second, g.stream);
                g.pending.push_back(ev);
                ev.first = nullptr;
                if (g.pending.size() > 64)
                {
                    retire_finished();
                }
            }
        }

        // a program that never polls the statistics must not accumulate events: fold the sections the device has already
        // finished into the total and recycle their events (no synchronisation)
        static void retire_finished()
        {
            size_t keep = 0;
            for (size_t i = 0; i < g.pending.size(); ++i)
            {
                auto& p = g.pending[i];
                float ms = 0;
                if (cudaEventQuery(p.second) == cudaSuccess && cudaEventElapsedTime(&ms, p.first, p.second) == cudaSuccess)
                {
                    g.stats.device_seconds += ms * 1e-3;
                    g.pool.push_back(p);
                }
                else
                {
                    g.pending[keep++] = p;
                }
            }
            g.pending.resize(keep);
            (void) cudaGetLastError(); // cudaErrorNotReady from the queries is not an error
        }

        ~Section()
        {
            close();
        }
    };

    static void resolve_sections()
    {
        if (!g.device || g.pending.empty())
        {
            return;
        }
        cudaStreamSynchronize(g.stream);
        for (auto& p : g.pending)
        {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, p.first, p.second) == cudaSuccess)
            {
                g.stats.device_seconds += ms * 1e-3;
            }
            g.pool.push_back(p);
        }
        g.pending.clear();
    }

    static void prof_begin()
    {
        if (g.profile)
        {
            if (!g.prof_a)
            {
                SMR_CUDA(cudaEventCreate(&g.prof_a));
                SMR_CUDA(cudaEventCreate(&g.prof_b));
            }
            SMR_CUDA(cudaEventRecord(g.prof_a, g.stream));
        }
    }

    static void prof_end(int fam, int64_t cells)
    {
        if (g.profile)
        {
            SMR_CUDA(cudaEventRecord(g.prof_b, g.stream));
            SMR_CUDA(cudaEventSynchronize(g.prof_b));
            float ms = 0;
            SMR_CUDA(cudaEventElapsedTime(&ms, g.prof_a, g.prof_b));
            g.prof_seconds[fam] += ms * 1e-3;
            g.prof_launches[fam] += 1;
            g.prof_cells[fam] += static_cast<uint64_t>(cells);
        }
    }

    // cross-GPU phase barrier: queued after every launch whose outputs a peer may read (and after local memsets/copies of
    // buffers a peer may write next)
    static void mg_barrier()
    {
        if (g.mg_world > 1 && g.mg_connected)
        {
            mg_barrier_kernel<<<1, 32, 0, g.stream>>>(++g.epoch);
            SMR_CUDA(cudaGetLastError());
            ++g.stats.kernel_launches;
        }
    }

    static void mg_check_error()
    {
        if (g.mg_world > 1 && g.mg_connected)
        {
            unsigned long long err = 0;
            SMR_CUDA(cudaMemcpy(&err, static_cast<char*>(g.mg_pool) + 1024, sizeof(err), cudaMemcpyDeviceToHost));
            if (err != 0)
            {
                throw CudaError("multi-GPU barrier timed out at epoch " + std::to_string(err) + " (a peer stopped)");
            }
        }
    }

    static MeshObj& get_mesh(smr_mesh_t h)
    {
        auto it = g.meshes.find(h);
        if (it == g.meshes.end())
        {
            throw std::invalid_argument("invalid mesh handle");
        }
        return *it->second;
    }

    static FieldObj& get_field(smr_field_t h)
    {
        auto it = g.fields.find(h);
        if (it == g.fields.end())
        {
            throw std::invalid_argument("invalid field handle");
        }
        return *it->second;
    }

    static void config_to_mesh_cfg(const smr_mesh_config* c, MeshConfig& out)
    {
        if (!c)
        {
            throw std::invalid_argument("null mesh config");
        }
        if (c->dim < 1 || c->dim > 3)
        {
            throw std::invalid_argument("dim must be 1, 2 or 3");
        }
        if (c->max_level < c->min_level)
        {
            throw std::invalid_argument("Max level must be greater than min level."); // mesh_config.hpp:383-386
        }
        if (c->max_level + 3 > SMR_MAX_LEVELS)
        {
            throw std::invalid_argument("max_level too large (max_refinement_level is 20, samurai_config.hpp:50)");
        }
        bool all_periodic = true;
        for (int d = 0; d < c->dim; ++d)
        {
            all_periodic = all_periodic && c->periodic[d] != 0;
        }
        // a fully periodic mesh has no boundary: any ghost width up to 3 (max_stencil_size(6), the WENO5 stencil)
        if (c->max_stencil_radius < 1 || c->max_stencil_radius > (all_periodic ? 3 : 2))
        {
            // ghost width 2 (the library default, mesh_config.hpp:388-393) is built for boundary conditions that fill one layer:
            // further-ghost extrapolation (bc/apply_field_bc.hpp:499-563), two-layer corner block (:313-466), contiguous-boundary
            // graduation rule (graduation.hpp:372-455).  Wider stencils (WENO5: radius 3) are not.
            throw std::invalid_argument("max_stencil_radius must be 1 or 2 (up to 3 on a fully periodic mesh)");
        }
        if (c->pred_radius < 0 || c->pred_radius > 1)
        {
            throw std::invalid_argument("prediction_stencil_radius must be 0 or 1 (update_outer_ghost.hpp:341)");
        }
        out.dim                = c->dim;
        out.min_level          = c->min_level;
        out.max_level          = c->max_level;
        out.pred_radius        = c->pred_radius;
        out.max_stencil_radius = c->max_stencil_radius;
        out.graduation_width   = c->graduation_width;
        out.refine_boundary    = c->refine_boundary != 0;
        for (int d = 0; d < 3; ++d)
        {
            out.n0[d]     = d < c->dim ? c->n_cells0[d] : 1;
            out.origin[d] = d < c->dim ? c->origin[d] : 0.0;
            if (out.n0[d] < 1)
            {
                throw std::invalid_argument("n_cells0 must be >= 1");
            }
        }
        out.scaling = c->scaling_factor;
        for (int d = 0; d < 3; ++d)
        {
            out.periodic[d] = d < c->dim && c->periodic[d] != 0;
        }
    }

    // ---------------------------------------------------------------------------------------------------------
    // plan upload and launches
    // ---------------------------------------------------------------------------------------------------------
    static void* pinned_alloc(size_t n)
    {
        void* p = nullptr;
        if (cudaHostAlloc(&p, n, cudaHostAllocDefault) != cudaSuccess)
        {
            return nullptr;
        }
        return p;
    }

    static void pinned_free(void* p)
    {
        cudaFreeHost(p);
    }

    // the arena is pinned host memory (Arena::alloc_fn): one async copy moves the whole plan
    static void upload_arena(const Arena& a, DevBuf& d)
    {
        if (a.size == 0)
        {
            return;
        }
        d.ensure(a.device_bytes());
        SMR_CUDA(cudaMemcpyAsync(d.p, a.p, a.size, cudaMemcpyHostToDevice, g.stream));
        g.stats.h2d_bytes += a.size;
    }

    // upload the reference sub-mesh CSR of the current mesh (once per mesh)
    static void ensure_csr(MeshObj& mo)
    {
        if (mo.csr_ready)
        {
            return;
        }
        const int k   = mo.h_csr_next;
        mo.h_csr_next = k ^ 1;
        if (mo.h_csr_done[k] == nullptr)
        {
            SMR_CUDA(cudaEventCreateWithFlags(&mo.h_csr_done[k], cudaEventDisableTiming));
        }
        SMR_CUDA(cudaEventSynchronize(mo.h_csr_done[k])); // the copy that last read this staging buffer has finished
        const double t0 = now();
        build_csr(mo.mesh, mo.h_csr[k]);
        const double dt = now() - t0;
        g.stats.host_batch_seconds += dt;
        g.stats.host_stage_seconds[5] += dt;
        upload_arena(mo.h_csr[k].arena, mo.d_csr);
        SMR_CUDA(cudaEventRecord(mo.h_csr_done[k], g.stream));
        mo.csr_tab   = mo.h_csr[k].tab;
        mo.csr_ready = true;
    }

    // one launch turns the seeds of an uploaded arena into records (derive.cuh).  `from_old`: sources are looked up in the
    // previous mesh (field transfer), destinations always in the current one.
    static void run_derive(DevBuf& d_arena, const DeriveList& dl, MeshObj& mo, bool from_old)
    {
        if (dl.jobs.empty() || dl.blocks == 0)
        {
            return;
        }
        if (g.derive_error == nullptr)
        {
            SMR_CUDA(cudaMalloc(reinterpret_cast<void**>(&g.derive_error), 64));
            SMR_CUDA(cudaMemsetAsync(g.derive_error, 0, 64, g.stream));
            SMR_CUDA(cudaMallocHost(reinterpret_cast<void**>(&g.derive_error_host), 64));
            std::memset(g.derive_error_host, 0, 64);
        }
        DeriveArgs a;
        a.arena     = static_cast<const char*>(d_arena.p);
        a.arena_out = static_cast<char*>(d_arena.p);
        a.jobs      = reinterpret_cast<const smr_derive_job*>(a.arena + dl.jobs_off);
        a.n_jobs    = static_cast<int>(dl.jobs.size());
        a.dim       = mo.mesh.cfg.dim;
        a.radius    = mo.mesh.cfg.pred_radius;
        a.csr_dst   = static_cast<const char*>(mo.d_csr.p);
        a.tab_dst   = mo.csr_tab;
        a.csr_src   = static_cast<const char*>(from_old ? mo.d_csr_old.p : mo.d_csr.p);
        a.tab_src   = from_old ? mo.csr_tab_old : mo.csr_tab;
        a.error     = g.derive_error;
        SMR_CUDA(launch_derive(dl.blocks, g.stream, a));
        ++g.stats.kernel_launches;
    }

    // host half of ensure_plan: slab cuts + traversal of the mesh into index batches (no CUDA calls except the pinned arena)
    static void update_filter(MeshObj& mo)
    {
        if (mo.filter.world != g.mg_world || mo.filter.cut2.empty())
        {
            mo.filter.rank  = g.mg_rank;
            mo.filter.world = g.mg_world;
            mo.filter.compute_cuts(mo.mesh);
        }
        else
        {
            // keep the cuts across adaptations (data only lives near its owner): re-cut with smr_mg_rebalance
            mo.filter.dim = mo.mesh.cfg.dim;
            mo.filter.L   = mo.mesh.cfg.max_level;
        }
    }

    static void build_plan_host(MeshObj& mo)
    {
        const double t0 = now();
        build_plan(mo.mesh, mo.plan, mo.filter);
        mo.plan_seconds = now() - t0;
    }

    void MeshObj::invalidate_plans()
    {
        plan_ready = false;
        flux.ready  = false;
        fluxw.ready = false;
    }

    static void ensure_plan(MeshObj& mo)
    {
        if (mo.plan_ready)
        {
            return;
        }
        update_filter(mo);
        build_plan_host(mo);
        g.stats.host_batch_seconds += mo.plan_seconds;
        g.stats.host_stage_seconds[5] += mo.plan_seconds;
        // the previous arena may still be in use by queued kernels: stream-ordered, so a sync is needed before reuse
        const double tw = now();
        SMR_CUDA(cudaStreamSynchronize(g.stream));
        g.stats.host_stage_seconds[6] += now() - tw;
        ensure_csr(mo);
        upload_arena(mo.plan.arena, mo.d_arena);
        run_derive(mo.d_arena, mo.plan.derive, mo, false);
        mo.plan_ready = true;
    }

    template <class Item>
    static BatchView<Item> view_of(const void* arena, const Batch& b, int64_t n_cells)
    {
        const char* base = static_cast<const char*>(arena);
        return BatchView<Item>{reinterpret_cast<const Item*>(base + b.items),
                               reinterpret_cast<const int64_t*>(base + b.prefix),
                               reinterpret_cast<const int32_t*>(base + b.cta_first),
                               n_cells};
    }

    // launch the first `limit` output cells of a batch (limit < 0: all of it)
    template <class Item, class Op>
    static void launch(int fam, const void* arena, const Batch& b, const Op& op, int64_t limit = -1)
    {
        const int64_t n_cells = limit < 0 ? b.n_cells : std::min(limit, b.n_cells);
        if (b.empty() || n_cells <= 0)
        {
            mg_barrier(); // every rank queues the same number of barriers whatever its share of the records
            return;
        }
        prof_begin();
        if (b.cta_units != SMR_CTA_THREADS * Op::units_per_thread)
        {
            throw std::logic_error("batch laid out for a different CTA size than its kernel");
        }
        const int n_ctas = static_cast<int>((n_cells + b.cta_units - 1) / b.cta_units);
        SMR_CUDA((launch_batch<Item, Op>(n_ctas, g.stream, view_of<Item>(arena, b, n_cells), op)));
        ++g.stats.kernel_launches;
        prof_end(fam, n_cells);
        mg_barrier();
    }

    static void launch_ghost_phase(int dim, const void* arena, const GhostPhase& ph, double* f, int bc_type, double bc_value)
    {
        if (ph.bc.empty() && ph.proj.empty())
        {
            mg_barrier();
            return;
        }
        const char* base = static_cast<const char*>(arena);
        BcView bc{nullptr, nullptr, 0, bc_type, bc_value};
        if (!ph.bc.empty())
        {
            bc.items = reinterpret_cast<const smr_item_bc*>(base + ph.bc.items);
            bc.srcs  = reinterpret_cast<const int64_t*>(base + ph.bc.aux);
            bc.n     = ph.bc.n_items;
        }
        BatchView<smr_item_proj> pv{nullptr, nullptr, nullptr, 0};
        if (!ph.proj.empty())
        {
            pv = view_of<smr_item_proj>(arena, ph.proj, ph.proj.n_cells);
        }
        const int grid = ph.bc.n_ctas + ph.proj.n_ctas;
        prof_begin();
        SMR_CUDA(launch_ghost_phase_kernel(dim, grid, g.stream, bc, ph.bc.n_ctas, pv, f));
        ++g.stats.kernel_launches;
        prof_end(ph.proj.n_cells >= ph.bc.n_items ? SMR_FAM_PROJ : SMR_FAM_BC, ph.proj.n_cells + ph.bc.n_items);
        mg_barrier();
    }

    template <template <int> class OpT, class Item, class... Args>
    static void launch_dim(int fam, int dim, const void* arena, const Batch& b, int64_t limit, Args... args)
    {
        switch (dim)
        {
            case 1:
                launch<Item>(fam, arena, b, OpT<1>{args...}, limit);
                break;
            case 2:
                launch<Item>(fam, arena, b, OpT<2>{args...}, limit);
                break;
            default:
                launch<Item>(fam, arena, b, OpT<3>{args...}, limit);
                break;
        }
    }

    template <int D>
    using PredOp0 = PredOp<D, 0>;
    template <int D>
    using PredOp1 = PredOp<D, 1>;
    template <int D>
    using DetailOp0 = DetailOp<D, 0>;
    template <int D>
    using DetailOp1 = DetailOp<D, 1>;
    template <int D>
    using MaximumOpR = MaximumOp<D, true>;
    template <int D>
    using UpwindOp = FvOp<D, false>;
    template <int D>
    using BurgersOp = FvOp<D, true>;
    template <int D>
    using UpwindStripOp = FvStripOp<D, false>;
    template <int D>
    using BurgersStripOp = FvStripOp<D, true>;

    static void launch_pred(int dim, int radius, const void* arena, const Batch& b, const double* src, double* dst)
    {
        if (radius == 0)
        {
            launch_dim<PredOp0, smr_item_pred>(SMR_FAM_PRED, dim, arena, b, -1, src, dst);
        }
        else
        {
            launch_dim<PredOp1, smr_item_pred>(SMR_FAM_PRED, dim, arena, b, -1, src, dst);
        }
    }

    // ---- fused level wavefront (kernels.cuh: wavefront_kernel) -------------------------------------------------------
    constexpr int WF_SLOTS      = 16;
    constexpr size_t WF_SLOT_BYTES = 16384; // phase + job tables of one launch (also the kernel's dynamic shared memory bound)
    constexpr int WF_WIDE_FACTOR   = 8;     // phases above this many quarter chunks per CTA of the grid use full chunks
    constexpr int WF_SERIAL_CTAS   = 1;     // phases of at most this many chunks are run by CTA 0 alone

    struct WfBuilder
    {
        std::vector<WfPhase> phases;
        std::vector<WfJob> jobs;
        int64_t units = 0;
        uint64_t bytes = 0;
        int dim       = 2;
        bool open     = false;
        bool keep_empty = false; // multi-GPU: phases without local work still count (every rank crosses the same barriers)

        void begin_phase()
        {
            phases.push_back(WfPhase{static_cast<int32_t>(jobs.size()), 0, 0, 0});
            open = true;
        }

        void end_phase()
        {
            if (open && phases.back().n_jobs == 0 && !keep_empty)
            {
                phases.pop_back();
            }
            open = false;
        }

        // algorithmic bytes per output unit
        uint64_t unit_bytes(int op) const
        {
            const int c = 1 << dim;
            switch (op)
            {
                case WF_BC:
                    return 24;
                case WF_PROJ:
                    return 8 * (c + 1);
                case WF_PRED:
                    return 8 + 8 / c + (8 % c ? 1 : 0);
                case WF_DETAIL:
                    return 8 * (1 + 2 * c);
                case WF_CRITERIA:
                    return 8 * (c + 1) + 2 * c;
                case WF_MAXIMUM:
                    return 2 * c + 2;
                case WF_KEEP:
                case WF_TAGS_CHANGE:
                    return 1;
                case WF_TAG_OR:
                    return 4;
                case WF_COPY:
                    return 16;
                default:
                    return 1; // zero fill: per byte
            }
        }

        void add(int op, const Batch& b, int field, int64_t limit = -1)
        {
            if (b.empty())
            {
                return;
            }
            WfJob j{};
            j.op = op;
            if (op == WF_BC)
            {
                j.n_cells = b.n_items;
                j.n_ctas  = (b.n_items + SMR_CTA_THREADS - 1) / SMR_CTA_THREADS;
            }
            else
            {
                j.n_cells = limit < 0 ? b.n_cells : std::min(limit, b.n_cells);
                if (j.n_cells <= 0)
                {
                    return;
                }
                if (b.cta_units != SMR_CTA_CELLS)
                {
                    throw std::logic_error("wavefront job laid out for a different CTA size");
                }
                // work items of the kernel are quarter chunks (kernels.cuh: run_batch_sub)
                j.n_ctas = static_cast<int32_t>((j.n_cells + SMR_CTA_THREADS - 1) / SMR_CTA_THREADS);
            }
            j.items     = b.items;
            j.prefix    = b.prefix;
            j.cta_first = b.cta_first;
            j.aux       = op == WF_BC ? b.aux : static_cast<int64_t>(b.n_items); // bc: source offsets; batch jobs: number of records
            j.field     = field;
            push(j);
        }

        void add_zero(int op, int64_t nbytes)
        {
            if (nbytes <= 0)
            {
                return;
            }
            WfJob j{};
            j.op      = op;
            j.n_cells = nbytes;
            j.n_ctas  = static_cast<int32_t>((nbytes + SMR_WF_ZERO_BYTES - 1) / SMR_WF_ZERO_BYTES);
            push(j);
        }

        void push(const WfJob& j)
        {
            jobs.push_back(j);
            phases.back().n_jobs += 1;
            phases.back().total_ctas += j.n_ctas;
            units += j.n_cells;
            bytes += unit_bytes(j.op) * static_cast<uint64_t>(j.n_cells);
        }
    };

    // multi-GPU: every rank walks the SAME phase list (empty phases are kept, nothing is run serially or split off), the phase
    // barrier exchanges flags with the peers (kernels.cuh: wf_mg_barrier); SMR_WF_MG=0 falls back to one launch per sweep
    static bool wf_multi()
    {
        return g.mg_world > 1 && g.mg_connected;
    }

    static bool wf_enabled()
    {
        static const bool mg_off = std::getenv("SMR_WF_MG") != nullptr && std::getenv("SMR_WF_MG")[0] == '0';
        return g.fuse && (g.mg_world == 1 || (g.mg_connected && !mg_off));
    }

    template <int DIM, int RADIUS>
    static void wf_launch_t(WfArgs& a, int grid, size_t smem)
    {
        SMR_CUDA((wf_launch_inst<DIM, RADIUS>(a, grid, smem, g.stream)));
    }

    template <int DIM, int RADIUS>
    static int wf_occupancy()
    {
        const int per_sm = wf_occupancy_inst<DIM, RADIUS>(WF_SLOT_BYTES);
        if (per_sm < 0)
        {
            throw CudaError("cudaOccupancyMaxActiveBlocksPerMultiprocessor(wavefront_kernel) failed");
        }
        return per_sm;
    }

    // queue the read-back of the wavefront error word; wf_check_error() after the next stream synchronisation throws if a
    // grid barrier timed out (the launch then left its outputs incomplete)
    static void wf_fetch_error()
    {
        if (g.derive_error != nullptr)
        {
            SMR_CUDA(cudaMemcpyAsync(g.derive_error_host, g.derive_error, 32, cudaMemcpyDeviceToHost, g.stream));
        }
        if (g.wf_barrier != nullptr)
        {
            SMR_CUDA(cudaMemcpyAsync(g.wf_error_host, g.wf_barrier + SMR_WF_ERROR_WORD, sizeof(unsigned), cudaMemcpyDeviceToHost, g.stream));
        }
    }

    static void wf_check_error()
    {
        if (g.derive_error_host != nullptr && g.derive_error_host[0] != 0u)
        {
            static const char* kinds[] = {"fv", "fv strip", "projection", "prediction", "detail", "tag", "copy"};
            const unsigned* e  = g.derive_error_host;
            const unsigned k   = e[0] - 1;
            std::string msg    = std::string("interval not found in the reference mesh (") + (k < 7 ? kinds[k] : "?") + " record) at level "
                              + std::to_string(static_cast<int>(e[1])) + ", i = " + std::to_string(static_cast<int>(e[2]))
                              + ", index = " + std::to_string(static_cast<int>(e[3])) + " " + std::to_string(static_cast<int>(e[4]));
            g.derive_error_host[0] = 0;
            cudaMemsetAsync(g.derive_error, 0, 64, g.stream);
            throw std::out_of_range(msg);
        }
        if (g.wf_error_host != nullptr && *g.wf_error_host != 0u)
        {
            const unsigned at = *g.wf_error_host;
            *g.wf_error_host  = 0;
            cudaMemsetAsync(g.wf_barrier, 0, 256, g.stream);
            g.wf_barrier_at = 0;
            g.wf_release_at = 0;
            throw CudaError("fused wavefront: grid barrier timed out (target " + std::to_string(at) + "); results of that launch are invalid");
        }
    }

    static void wf_init()
    {
        if (g.wf_host == nullptr)
        {
            SMR_CUDA(cudaMallocHost(&g.wf_host, WF_SLOTS * WF_SLOT_BYTES));
            SMR_CUDA(cudaMalloc(&g.wf_dev, WF_SLOTS * WF_SLOT_BYTES));
            SMR_CUDA(cudaMalloc(&g.wf_barrier, 256));
            SMR_CUDA(cudaMemset(g.wf_barrier, 0, 256));
            SMR_CUDA(cudaMallocHost(reinterpret_cast<void**>(&g.wf_error_host), 64));
            *g.wf_error_host = 0;
            for (int i = 0; i < WF_SLOTS; ++i)
            {
                SMR_CUDA(cudaEventCreateWithFlags(&g.wf_done[i], cudaEventDisableTiming));
            }
            cudaDeviceProp prop;
            SMR_CUDA(cudaGetDeviceProperties(&prop, g.dev));
            int per_sm = 2;
            // the same occupancy holds for every instantiation we launch at this grid size: take the smallest
            per_sm = std::min(per_sm, std::min(std::min(wf_occupancy<1, 1>(), wf_occupancy<2, 1>()), wf_occupancy<3, 1>()));
            per_sm = std::min(per_sm, std::min(std::min(wf_occupancy<1, 0>(), wf_occupancy<2, 0>()), wf_occupancy<3, 0>()));
            if (per_sm < 1)
            {
                throw CudaError("wavefront kernel does not fit on an SM");
            }
            g.wf_grid = prop.multiProcessorCount * per_sm;
        }
    }

    // One job of a throughput-bound phase as its own launch at full occupancy.  The cooperative kernel keeps 2 CTAs of 256
    // threads per SM (it is sized for the latency-bound level sweeps of adapted meshes); a phase with hundreds of chunks per
    // CTA (uniform and near-uniform levels) streams 2-3x faster from a plain launch of the same operator
    // (profiles/r02_summary.md: uniform level-13 step 3.0 -> 1.3 ms).  Per-cell arithmetic is the same functor.
    static void wf_job_standalone(const WfArgs& a, const WfJob& jb, const void* arena, int dim, int radius)
    {
        const char* base = static_cast<const char*>(arena);
        auto view = [&](auto tag_item)
        {
            using Item = decltype(tag_item);
            return BatchView<Item>{reinterpret_cast<const Item*>(base + jb.items), reinterpret_cast<const int64_t*>(base + jb.prefix),
                                   reinterpret_cast<const int32_t*>(base + jb.cta_first), jb.n_cells};
        };
        const int n_ctas = static_cast<int>((jb.n_cells + SMR_CTA_CELLS - 1) / SMR_CTA_CELLS);
        int fam          = SMR_FAM_WAVEFRONT;
        prof_begin();
        switch (jb.op)
        {
            case WF_BC:
            {
                BcView bc{reinterpret_cast<const smr_item_bc*>(base + jb.items), reinterpret_cast<const int64_t*>(base + jb.aux), static_cast<int>(jb.n_cells),
                          a.bc_type[jb.field], a.bc_value[jb.field]};
                const int ctas = static_cast<int>((jb.n_cells + SMR_CTA_THREADS - 1) / SMR_CTA_THREADS);
                SMR_CUDA(launch_ghost_phase_kernel(dim, ctas, g.stream, bc, ctas, BatchView<smr_item_proj>{nullptr, nullptr, nullptr, 0}, a.dst[jb.field]));
                fam = SMR_FAM_BC;
                break;
            }
            case WF_PROJ:
                fam = SMR_FAM_PROJ;
                switch (dim)
                {
                    case 1:
                        SMR_CUDA((launch_batch<smr_item_proj, ProjOp<1>>(n_ctas, g.stream, view(smr_item_proj{}), ProjOp<1>{a.src[jb.field], a.dst[jb.field]})));
                        break;
                    case 2:
                        SMR_CUDA((launch_batch<smr_item_proj, ProjOp<2>>(n_ctas, g.stream, view(smr_item_proj{}), ProjOp<2>{a.src[jb.field], a.dst[jb.field]})));
                        break;
                    default:
                        SMR_CUDA((launch_batch<smr_item_proj, ProjOp<3>>(n_ctas, g.stream, view(smr_item_proj{}), ProjOp<3>{a.src[jb.field], a.dst[jb.field]})));
                        break;
                }
                break;
            case WF_PRED:
                fam = SMR_FAM_PRED;
#define SMR_WF_PRED_CASE(D, R)                                                                                                                    \
    SMR_CUDA((launch_batch<smr_item_pred, PredOp<D, R>>(n_ctas, g.stream, view(smr_item_pred{}), PredOp<D, R>{a.src[jb.field], a.dst[jb.field]})))
                switch (dim * 2 + (radius ? 1 : 0))
                {
                    case 2:
                        SMR_WF_PRED_CASE(1, 0);
                        break;
                    case 3:
                        SMR_WF_PRED_CASE(1, 1);
                        break;
                    case 4:
                        SMR_WF_PRED_CASE(2, 0);
                        break;
                    case 5:
                        SMR_WF_PRED_CASE(2, 1);
                        break;
                    case 6:
                        SMR_WF_PRED_CASE(3, 0);
                        break;
                    default:
                        SMR_WF_PRED_CASE(3, 1);
                        break;
                }
#undef SMR_WF_PRED_CASE
                break;
            case WF_DETAIL:
                fam = SMR_FAM_DETAIL;
#define SMR_WF_DETAIL_CASE(D, R)                                                                                          \
    SMR_CUDA((launch_batch<smr_item_detail, DetailOp<D, R>>(n_ctas, g.stream, view(smr_item_detail{}), \
                                                             DetailOp<D, R>{a.dst[jb.field], a.detail + jb.field * a.n})))
                switch (dim * 2 + (radius ? 1 : 0))
                {
                    case 2:
                        SMR_WF_DETAIL_CASE(1, 0);
                        break;
                    case 3:
                        SMR_WF_DETAIL_CASE(1, 1);
                        break;
                    case 4:
                        SMR_WF_DETAIL_CASE(2, 0);
                        break;
                    case 5:
                        SMR_WF_DETAIL_CASE(2, 1);
                        break;
                    case 6:
                        SMR_WF_DETAIL_CASE(3, 0);
                        break;
                    default:
                        SMR_WF_DETAIL_CASE(3, 1);
                        break;
                }
#undef SMR_WF_DETAIL_CASE
                break;
            case WF_CRITERIA:
                fam = SMR_FAM_CRITERIA;
                switch (dim)
                {
                    case 1:
                        SMR_CUDA((launch_batch<smr_item_tag, CriteriaOp<1>>(n_ctas, g.stream, view(smr_item_tag{}), CriteriaOp<1>{a.detail, a.tag, a.tp, a.ncomp, a.n})));
                        break;
                    case 2:
                        SMR_CUDA((launch_batch<smr_item_tag, CriteriaOp<2>>(n_ctas, g.stream, view(smr_item_tag{}), CriteriaOp<2>{a.detail, a.tag, a.tp, a.ncomp, a.n})));
                        break;
                    default:
                        SMR_CUDA((launch_batch<smr_item_tag, CriteriaOp<3>>(n_ctas, g.stream, view(smr_item_tag{}), CriteriaOp<3>{a.detail, a.tag, a.tp, a.ncomp, a.n})));
                        break;
                }
                break;
            case WF_MAXIMUM:
                fam = SMR_FAM_MAXIMUM;
                switch (dim)
                {
                    case 1:
                        SMR_CUDA((launch_batch<smr_item_tag, MaximumOp<1, true>>(n_ctas, g.stream, view(smr_item_tag{}), MaximumOp<1, true>{a.tag})));
                        break;
                    case 2:
                        SMR_CUDA((launch_batch<smr_item_tag, MaximumOp<2, true>>(n_ctas, g.stream, view(smr_item_tag{}), MaximumOp<2, true>{a.tag})));
                        break;
                    default:
                        SMR_CUDA((launch_batch<smr_item_tag, MaximumOp<3, true>>(n_ctas, g.stream, view(smr_item_tag{}), MaximumOp<3, true>{a.tag})));
                        break;
                }
                break;
            case WF_KEEP:
                fam = SMR_FAM_KEEP;
                SMR_CUDA((launch_records<smr_item_fv, KeepLeavesOp>(g.stream, reinterpret_cast<const smr_item_fv*>(base + jb.items), static_cast<int>(jb.aux),
                                                                    KeepLeavesOp{a.tag, a.mask_all})));
                break;
            case WF_TAGS_CHANGE:
                fam = SMR_FAM_KEEP;
                SMR_CUDA((launch_records<smr_item_fv, TagsChangeOp>(g.stream, reinterpret_cast<const smr_item_fv*>(base + jb.items), static_cast<int>(jb.aux),
                                                                    TagsChangeOp{a.tag, a.change_flag, a.tp.min_level, a.tp.max_level})));
                break;
            case WF_ZERO_DETAIL:
                fam = SMR_FAM_INIT;
                SMR_CUDA(cudaMemsetAsync(a.detail, 0, static_cast<size_t>(jb.n_cells), g.stream));
                break;
            case WF_ZERO_TAG:
                fam = SMR_FAM_INIT;
                SMR_CUDA(cudaMemsetAsync(a.tag, 0, static_cast<size_t>(jb.n_cells), g.stream));
                break;
            case WF_TAG_OR:
                fam = SMR_FAM_MAXIMUM;
                SMR_CUDA((launch_batch<smr_item_copy, TagOrOp>(n_ctas, g.stream, view(smr_item_copy{}), TagOrOp{a.tag, a.mask_all})));
                break;
            default: // WF_COPY
                fam = SMR_FAM_COPY;
                SMR_CUDA((launch_batch<smr_item_copy, CopyOp>(n_ctas, g.stream, view(smr_item_copy{}), CopyOp{a.src[jb.field], a.dst[jb.field]})));
                break;
        }
        ++g.stats.kernel_launches;
        prof_end(fam, jb.n_cells);
    }

    static void wf_run_fused(WfBuilder& wb, WfArgs& a, const void* arena, int dim, int radius);

    // run the phases of `wb` (offsets relative to `arena`): runs of latency-bound phases go into one cooperative launch each,
    // throughput-bound phases (far more chunks than the cooperative grid has CTAs) are launched job by job at full occupancy
    static void wf_run(WfBuilder& wb, WfArgs& a, const void* arena, int dim, int radius)
    {
        wb.end_phase();
        if (wb.phases.empty())
        {
            return;
        }
        wf_init();
        static const bool no_split_env = std::getenv("SMR_WF_NO_SPLIT") != nullptr;
        const bool no_split            = no_split_env || wf_multi();
        const int wide_at = WF_WIDE_FACTOR * g.wf_grid;
        size_t p = 0;
        while (p < wb.phases.size())
        {
            if (!no_split && wb.phases[p].total_ctas > wide_at)
            {
                const WfPhase& ph = wb.phases[p];
                for (int j = ph.first_job; j < ph.first_job + ph.n_jobs; ++j)
                {
                    wf_job_standalone(a, wb.jobs[static_cast<size_t>(j)], arena, dim, radius);
                }
                ++p;
                continue;
            }
            size_t q = p;
            while (q < wb.phases.size() && (no_split || wb.phases[q].total_ctas <= wide_at))
            {
                ++q;
            }
            if (p == 0 && q == wb.phases.size())
            {
                wf_run_fused(wb, a, arena, dim, radius); // the usual case: nothing to split
                return;
            }
            WfBuilder sub;
            sub.dim = wb.dim;
            for (size_t k = p; k < q; ++k)
            {
                const WfPhase& ph = wb.phases[k];
                sub.begin_phase();
                for (int j = ph.first_job; j < ph.first_job + ph.n_jobs; ++j)
                {
                    sub.push(wb.jobs[static_cast<size_t>(j)]);
                }
                sub.end_phase();
            }
            wf_run_fused(sub, a, arena, dim, radius);
            p = q;
        }
    }

    static void wf_run_fused(WfBuilder& wb, WfArgs& a, const void* arena, int dim, int radius)
    {
        for (WfPhase& ph : wb.phases)
        {
            // a phase with far more quarter chunks than CTAs is throughput bound: switch its batch jobs to full chunks
            // (only reached with SMR_WF_NO_SPLIT: wf_run launches such phases on their own)
            if (g.wf_grid > 0 && ph.total_ctas > WF_WIDE_FACTOR * g.wf_grid)
            {
                int total = 0;
                for (int j = ph.first_job; j < ph.first_job + ph.n_jobs; ++j)
                {
                    WfJob& jb = wb.jobs[static_cast<size_t>(j)];
                    if (jb.op != WF_BC && jb.op != WF_ZERO_DETAIL && jb.op != WF_ZERO_TAG)
                    {
                        jb.pad    = 1;
                        jb.n_ctas = static_cast<int32_t>((jb.n_cells + SMR_CTA_CELLS - 1) / SMR_CTA_CELLS);
                    }
                    total += jb.n_ctas;
                }
                ph.total_ctas = total;
            }
            static const bool no_serial = std::getenv("SMR_WF_NO_SERIAL") != nullptr;
            ph.pad = (!no_serial && !wf_multi() && ph.total_ctas <= WF_SERIAL_CTAS) ? 1 : 0;
        }
        static const bool trace = std::getenv("SMR_WF_TRACE") != nullptr;
        if (trace)
        {
            std::fprintf(stderr, "wavefront: %zu phases, chunks per phase:", wb.phases.size());
            for (const WfPhase& ph : wb.phases)
            {
                std::fprintf(stderr, " %d%s", ph.total_ctas, ph.pad ? "s" : "");
            }
            std::fprintf(stderr, "\n");
        }
        const size_t pbytes = wb.phases.size() * sizeof(WfPhase), jbytes = wb.jobs.size() * sizeof(WfJob);
        if (pbytes + jbytes > WF_SLOT_BYTES)
        {
            throw std::logic_error("wavefront table exceeds its staging slot");
        }
        const int slot = g.wf_next;
        g.wf_next      = (g.wf_next + 1) % WF_SLOTS;
        SMR_CUDA(cudaEventSynchronize(g.wf_done[slot])); // the launch that last used this slot has consumed it
        char* h = static_cast<char*>(g.wf_host) + slot * WF_SLOT_BYTES;
        char* d = static_cast<char*>(g.wf_dev) + slot * WF_SLOT_BYTES;
        std::memcpy(h, wb.phases.data(), pbytes);
        std::memcpy(h + pbytes, wb.jobs.data(), jbytes);
        SMR_CUDA(cudaMemcpyAsync(d, h, pbytes + jbytes, cudaMemcpyHostToDevice, g.stream));
        a.arena    = static_cast<const char*>(arena);
        a.phases   = reinterpret_cast<const WfPhase*>(d);
        a.jobs     = reinterpret_cast<const WfJob*>(d + pbytes);
        a.n_phases = static_cast<int>(wb.phases.size());
        a.n_jobs   = static_cast<int>(wb.jobs.size());
        int widest = 1;
        for (const WfPhase& ph : wb.phases)
        {
            widest = std::max(widest, ph.total_ctas);
        }
        const int grid = std::min(g.wf_grid, widest);
        // the kernel crosses one grid barrier per transition between phases that is not serial -> serial
        unsigned n_barriers = 0;
        for (size_t i = 0; i + 1 < wb.phases.size(); ++i)
        {
            n_barriers += (wb.phases[i].pad != 0 && wb.phases[i + 1].pad != 0) ? 0u : 1u;
        }
        a.barrier      = g.wf_barrier;
        a.barrier_base = g.wf_barrier_at;
        a.mg_world     = wf_multi() ? g.mg_world : 1;
        a.release_base = g.wf_release_at;
        a.mg_epoch_base = g.epoch;
        static void* d_trace = nullptr;
        if (trace)
        {
            if (d_trace == nullptr)
            {
                SMR_CUDA(cudaMalloc(&d_trace, 4096));
            }
            a.trace = static_cast<unsigned long long*>(d_trace);
        }
        prof_begin();
        switch (dim * 2 + (radius ? 1 : 0))
        {
            case 2:
                wf_launch_t<1, 0>(a, grid, pbytes + jbytes);
                break;
            case 3:
                wf_launch_t<1, 1>(a, grid, pbytes + jbytes);
                break;
            case 4:
                wf_launch_t<2, 0>(a, grid, pbytes + jbytes);
                break;
            case 5:
                wf_launch_t<2, 1>(a, grid, pbytes + jbytes);
                break;
            case 6:
                wf_launch_t<3, 0>(a, grid, pbytes + jbytes);
                break;
            default:
                wf_launch_t<3, 1>(a, grid, pbytes + jbytes);
                break;
        }
        // only a launch that was accepted advances the expected counter value (wraps modulo 2^32 like the device counter)
        g.wf_barrier_at += n_barriers * static_cast<unsigned>(grid);
        if (wf_multi())
        {
            g.wf_release_at += n_barriers;
            g.epoch += n_barriers; // the kernel used epochs g.epoch + 1 ... g.epoch + n_barriers on every rank
        }
        SMR_CUDA(cudaEventRecord(g.wf_done[slot], g.stream));
        ++g.stats.kernel_launches;
        if (g.profile)
        {
            g.prof_bytes[SMR_FAM_WAVEFRONT] += wb.bytes;
        }
        prof_end(SMR_FAM_WAVEFRONT, wb.units);
        mg_barrier(); // multi-GPU: the last phase's peer stores have landed everywhere before anything else is queued
        if (trace)
        {
            std::vector<unsigned long long> t(wb.phases.size() + 1);
            SMR_CUDA(cudaStreamSynchronize(g.stream));
            SMR_CUDA(cudaMemcpy(t.data(), d_trace, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            std::fprintf(stderr, "wavefront: grid %d, total %.1f us, per phase (us):", grid, (t.back() - t.front()) * 1e-3);
            for (size_t i = 0; i + 1 < t.size(); ++i)
            {
                std::fprintf(stderr, " %.1f", (t[i + 1] - t[i]) * 1e-3);
            }
            std::fprintf(stderr, "\n");
        }
    }

    // phases of update_ghost_mr for `fields` (algorithm/update_ghost_mr.hpp:194-237): top-down ghost phases, then the
    // bottom-up prediction; `extra(i)` lets the caller slip independent jobs into the i-th phase
    template <class Extra>
    static void wf_add_ghost_phases(WfBuilder& wb, MeshObj& mo, int nfields, Extra&& extra)
    {
        const MeshConfig& cfg = mo.mesh.cfg;
        int index             = 0;
        const bool periodic = cfg.any_periodic();
        // update_ghost_periodic(level): one phase per periodic dimension (the later dimensions copy ghosts of the earlier ones)
        auto periodic_phases = [&](int level)
        {
            for (int k = 0; k < cfg.dim; ++k)
            {
                const Batch& b = mo.plan.down[level].per[k];
                if (cfg.periodic[k] && (!b.empty() || wb.keep_empty))
                {
                    wb.begin_phase();
                    for (int f = 0; f < nfields; ++f)
                    {
                        wb.add(WF_COPY, b, f);
                    }
                    wb.end_phase();
                }
            }
        };
        for (int level = cfg.max_level; level >= 0; --level)
        {
            const GhostPhase& ph = mo.plan.down[level];
            // algorithm/update_ghost_mr.hpp:204-222: periodic ghosts, outer ghosts, periodic ghosts again, projection.  Both passes
            // are needed even without outer ghosts: a corner ghost whose mirror along the last dimension is absent from the mesh is
            // filled along the first one, from a ghost the first pass has just written (measured: a single pass leaves such corners
            // stale).  The projection only touches cells inside the domain, so it shares the phase of the outer ghosts.
            if (periodic)
            {
                periodic_phases(level);
            }
            if (!(ph.bc.empty() && ph.proj.empty()) || wb.keep_empty)
            {
                wb.begin_phase();
                extra(index++);
                for (int f = 0; f < nfields; ++f)
                {
                    wb.add(WF_BC, ph.bc, f);
                    wb.add(WF_PROJ, ph.proj, f);
                }
                wb.end_phase();
            }
            if (cfg.ghost_width() == 2 && !cfg.all_periodic() && (!ph.bc2.empty() || wb.keep_empty))
            {
                // second ghost layer: extrapolated from the first one, which the phase above wrote
                wb.begin_phase();
                for (int f = 0; f < nfields; ++f)
                {
                    wb.add(WF_BC, ph.bc2, f);
                }
                wb.end_phase();
            }
            if (periodic)
            {
                periodic_phases(level);
            }
        }
        for (int level = 1; level <= cfg.max_level; ++level)
        {
            if (!mo.plan.pred[level].empty() || wb.keep_empty)
            {
                wb.begin_phase();
                extra(index++);
                for (int f = 0; f < nfields; ++f)
                {
                    wb.add(WF_PRED, mo.plan.pred[level], f);
                }
                wb.end_phase();
            }
            if (periodic)
            {
                // also without prediction ghosts at this level: predict_bc(level) wrote boundary ghosts of this level after its
                // top-down periodic passes (update_outer_ghost.hpp:396-400), and the mirrored corners must see them
                periodic_phases(level);
            }
        }
        // phases the caller still owes its extras to
        for (; index < 2; ++index)
        {
            wb.begin_phase();
            extra(index);
            wb.end_phase();
        }
    }

    static void wf_set_fields(WfArgs& a, const std::vector<FieldObj*>& fields)
    {
        if (fields.size() > SMR_WF_MAX_FIELDS)
        {
            throw std::invalid_argument("too many fields for one fused launch");
        }
        for (size_t i = 0; i < fields.size(); ++i)
        {
            if (fields[i]->bc_type < 0 && !fields[i]->mesh->mesh.cfg.all_periodic())
            {
                throw std::invalid_argument("field '" + fields[i]->name + "' has no boundary condition attached (make_bc)");
            }
            a.dst[i]      = static_cast<double*>(fields[i]->data.p);
            a.src[i]      = a.dst[i];
            a.bc_type[i]  = fields[i]->bc_type;
            a.bc_value[i] = fields[i]->bc_value;
        }
    }

    static void check_field_ready(FieldObj& f)
    {
        if (f.n != f.mesh->mesh.nref || f.data.p == nullptr)
        {
            throw std::invalid_argument("field '" + f.name + "' is not sized for its mesh: call smr_field_resize() after the mesh changed");
        }
    }

    static void do_update_ghost(FieldObj& f)
    {
        MeshObj& mo = *f.mesh;
        check_field_ready(f);
        if (f.bc_type < 0 && !mo.mesh.cfg.all_periodic())
        {
            throw std::invalid_argument("field '" + f.name + "' has no boundary condition attached (make_bc)");
        }
        ensure_plan(mo);
        const MeshConfig& cfg = mo.mesh.cfg;
        double* u             = static_cast<double*>(f.data.p);
        const void* arena     = mo.d_arena.p;
        if (wf_enabled())
        {
            WfBuilder wb;
            wb.dim = cfg.dim;
            wb.keep_empty = wf_multi();
            WfArgs a{};
            std::vector<FieldObj*> one{&f};
            wf_set_fields(a, one);
            wf_add_ghost_phases(wb, mo, 1, [](int) {});
            wf_run(wb, a, arena, cfg.dim, cfg.pred_radius);
            f.ghosts_valid = true;
            return;
        }
        auto periodic_copies = [&](int level)
        {
            for (int k = 0; k < cfg.dim; ++k)
            {
                if (cfg.periodic[k])
                {
                    launch<smr_item_copy>(SMR_FAM_COPY, arena, mo.plan.down[level].per[k], CopyOp{u, u});
                }
            }
        };
        for (int level = cfg.max_level; level >= 0; --level)
        {
            periodic_copies(level);
            launch_ghost_phase(cfg.dim, arena, mo.plan.down[level], u, f.bc_type, f.bc_value);
            if (cfg.ghost_width() == 2 && !cfg.all_periodic())
            {
                GhostPhase second;
                second.bc = mo.plan.down[level].bc2;
                launch_ghost_phase(cfg.dim, arena, second, u, f.bc_type, f.bc_value);
            }
            periodic_copies(level);
        }
        for (int level = 1; level <= cfg.max_level; ++level)
        {
            launch_pred(cfg.dim, cfg.pred_radius, arena, mo.plan.pred[level], u, u);
            periodic_copies(level);
        }
        f.ghosts_valid = true;
    }

    static void do_fv(FieldObj& out, FieldObj& in, const double* a, double dt, bool burgers)
    {
        require_device();
        if (out.mesh != in.mesh)
        {
            throw std::invalid_argument("fields live on different meshes");
        }
        if (&out == &in)
        {
            throw std::invalid_argument("output field must differ from input field");
        }
        MeshObj& mo = *in.mesh;
        check_field_ready(in);
        check_field_ready(out);
        ensure_plan(mo);
        Section sec;
        const MeshConfig& cfg = mo.mesh.cfg;
        if (burgers && cfg.dim != 2)
        {
            throw std::invalid_argument("upwind_scalar_burgers is only defined in 2D (stencil_field.hpp:219-243)");
        }
        FvParams p;
        for (int d = 0; d < 3; ++d)
        {
            const double ad = d < cfg.dim ? a[d] : 0.0;
            p.a[d]          = ad;
            p.half_a[d]     = .5 * ad;
            p.half_abs_a[d] = .5 * std::abs(ad);
        }
        p.dt = dt;
        // dx = scaling / 2^level: when `scaling` is a power of two so is dx, and dividing by it equals multiplying by
        // its (exact) inverse for every finite operand, which spares the fp64 division sequence
        int exp2      = 0;
        p.exact_inv   = (std::frexp(cfg.scaling, &exp2) == 0.5) ? 1 : 0;
        for (int l = 0; l < SMR_MAX_LEVELS; ++l)
        {
            p.dx[l]     = cfg.cell_length(l);
            p.inv_dx[l] = 1.0 / p.dx[l];
        }
        const double* u = static_cast<const double*>(in.data.p);
        double* o       = static_cast<double*>(out.data.p);
        out.ghosts_valid = false;
        if (g.profile && cfg.dim > 1)
        {
            g.prof_cells[SMR_FAM_FV] += static_cast<uint64_t>(mo.plan.fv_strip.n_cells) * (SMR_STRIP_ROWS - 1); // a strip unit = R cells
        }
        // strips of rows first (the bulk on uniform / smooth regions), then the single-row remainder
        if (burgers)
        {
            if (cfg.dim > 1)
            {
                launch_dim<BurgersStripOp, smr_item_fvstrip>(SMR_FAM_FV, cfg.dim, mo.d_arena.p, mo.plan.fv_strip, -1, u, o, p);
            }
            launch_dim<BurgersOp, smr_item_fv>(SMR_FAM_FV, cfg.dim, mo.d_arena.p, mo.plan.fv_single, -1, u, o, p);
        }
        else
        {
            if (cfg.dim > 1)
            {
                launch_dim<UpwindStripOp, smr_item_fvstrip>(SMR_FAM_FV, cfg.dim, mo.d_arena.p, mo.plan.fv_strip, -1, u, o, p);
            }
            launch_dim<UpwindOp, smr_item_fv>(SMR_FAM_FV, cfg.dim, mo.d_arena.p, mo.plan.fv_single, -1, u, o, p);
        }
    }

    // compute_relative_detail (mr/rel_detail.hpp:73-112), one component per adapted field
    static void relative_detail_pass(MeshObj& mo, std::vector<FieldObj*>& fields, double* detail, int64_t n)
    {
        const int ncomp   = static_cast<int>(fields.size());
        const void* arena = mo.d_arena.p;
        mo.d_relmax.ensure(sizeof(unsigned long long) * SMR_MAX_RANKS * 8);
        unsigned long long* slots = static_cast<unsigned long long*>(mo.d_relmax.p);
        mg_barrier();
        SMR_CUDA(cudaMemsetAsync(slots, 0, sizeof(unsigned long long) * SMR_MAX_RANKS * 8, g.stream));
        mg_barrier();
        for (int c = 0; c < ncomp; ++c)
        {
            launch<smr_item_fv>(SMR_FAM_DETAIL, arena, mo.plan.fv, AbsMaxOp{static_cast<const double*>(fields[c]->data.p), slots + c * SMR_MAX_RANKS + g.mg_rank});
            if (g.mg_world > 1 && g.mg_connected)
            {
                publish_slot_kernel<<<1, 32, 0, g.stream>>>(slots + c * SMR_MAX_RANKS);
                SMR_CUDA(cudaGetLastError());
                ++g.stats.kernel_launches;
                mg_barrier();
            }
            const int grid = static_cast<int>(std::min<int64_t>((n + SMR_CTA_THREADS - 1) / SMR_CTA_THREADS, 148 * 16));
            scale_detail_kernel<<<grid, SMR_CTA_THREADS, 0, g.stream>>>(detail + c * n, n, slots + c * SMR_MAX_RANKS, g.mg_world);
            SMR_CUDA(cudaGetLastError());
            ++g.stats.kernel_launches;
            mg_barrier();
        }
    }

    // one harten iteration; returns true when the mesh is unchanged
    static bool do_harten(std::vector<FieldObj*>& fields, double eps, double regularity, int ite, bool relative_detail = false)
    {
        require_device();
        MeshObj& mo           = *fields[0]->mesh;
        const MeshConfig& cfg = mo.mesh.cfg;
        const int dim = cfg.dim, L = cfg.max_level, lmin = cfg.min_level;
        const int ncomp = static_cast<int>(fields.size());
        for (auto* f : fields)
        {
            if (f->mesh != &mo)
            {
                throw std::invalid_argument("all adapted fields must live on the same mesh");
            }
            check_field_ready(*f);
        }
        ensure_plan(mo);
        Section sec1;
        const int64_t n = mo.mesh.nref;
        mo.d_detail.ensure(static_cast<size_t>(n) * sizeof(double) * ncomp);
        const int64_t flag_at = (n + 15) & ~int64_t(15); // change flag behind the tags, copied back with them
        mo.d_tag.ensure(static_cast<size_t>(flag_at) + 16);
        TagParams tp;
        tp.min_level = lmin;
        tp.max_level = L;
        for (int l = 0; l < SMR_MAX_LEVELS; ++l)
        {
            const int exponent = dim * (L - l);
            if (l > L || exponent >= 31)
            {
                tp.eps[l] = tp.fine_eps[l] = tp.coarse_eps[l] = 0;
                continue;
            }
            const double eps_l = eps / (1 << exponent);          // mr/adapt.hpp:328-329
            const double reg   = regularity + dim;               // mr/adapt.hpp:331
            tp.eps[l]          = eps_l;
            tp.fine_eps[l]     = std::pow(2.0, reg) * eps_l;     // mr/criteria.hpp:31
            tp.coarse_eps[l]   = tp.fine_eps[l] / (1 << dim);    // mr/criteria.hpp:32
        }
        if (dim * (L - lmin) >= 31)
        {
            throw std::invalid_argument("dim*(max_level-min_level) >= 31 overflows the reference's `1 << exponent` (mr/adapt.hpp:328)");
        }
        uint8_t* tag      = static_cast<uint8_t*>(mo.d_tag.p);
        double* detail    = static_cast<double*>(mo.d_detail.p);
        const void* arena = mo.d_arena.p;
        const int64_t detail_limit   = mo.plan.detail_cum[std::max(std::min(L - ite, mo.mesh.nlev), 0)];
        const int64_t criteria_limit = (L - ite) >= 0 ? mo.plan.tag_cum[L - ite] : 0;
        bool fused_flag              = false;
        if (wf_enabled())
        {
            // the whole device side of the iteration in one (two with relative detail) cooperative launch:
            // zero fill | keep tags | ghost wavefront of every field | detail | criteria | keep propagation
            WfArgs a{};
            wf_set_fields(a, fields);
            a.detail   = detail;
            a.tag      = tag;
            a.change_flag = reinterpret_cast<unsigned*>(tag + flag_at);
            a.n        = n;
            a.ncomp    = ncomp;
            a.mask_all = mo.filter.mask_all();
            a.tp       = tp;
            WfBuilder wb;
            wb.dim = dim;
            wb.keep_empty = wf_multi();
            wf_add_ghost_phases(wb, mo, ncomp,
                                [&](int index)
                                {
                                    if (index == 0)
                                    {
                                        wb.add_zero(WF_ZERO_DETAIL, static_cast<int64_t>(n) * static_cast<int64_t>(sizeof(double)) * ncomp);
                                        wb.add_zero(WF_ZERO_TAG, flag_at + 16);
                                    }
                                    else if (index == 1)
                                    {
                                        wb.add(WF_KEEP, mo.plan.fv, 0);
                                    }
                                });
            for (auto* f : fields)
            {
                f->ghosts_valid = true;
            }
            wb.begin_phase();
            for (int c = 0; c < ncomp; ++c)
            {
                wb.add(WF_DETAIL, mo.plan.detail, c, detail_limit);
            }
            wb.end_phase();
            if (relative_detail)
            {
                wf_run(wb, a, arena, dim, cfg.pred_radius);
                wb = WfBuilder();
                wb.dim = dim;
            wb.keep_empty = wf_multi();
                relative_detail_pass(mo, fields, detail, n);
            }
            wb.begin_phase();
            wb.add(WF_CRITERIA, mo.plan.tag_all, 0, criteria_limit);
            wb.end_phase();
            if (cfg.refine_boundary && (!mo.plan.keep_bdry.empty() || wb.keep_empty)) // keep_boundary_refined, mr/adapt.hpp:340-345
            {
                wb.begin_phase();
                wb.add(WF_KEEP, mo.plan.keep_bdry, 0);
                wb.end_phase();
            }
            for (int level = L; level >= 1; --level)
            {
                for (int k = 0; k < dim; ++k) // update_tag_periodic(level), mr/adapt.hpp:353
                {
                    if (cfg.periodic[k] && (!mo.plan.down[level].per[k].empty() || wb.keep_empty))
                    {
                        wb.begin_phase();
                        wb.add(WF_TAG_OR, mo.plan.down[level].per[k], 0);
                        wb.end_phase();
                    }
                }
                wb.begin_phase();
                wb.add(WF_MAXIMUM, mo.plan.tag[level], 0);
                wb.end_phase();
            }
            wb.begin_phase();
            wb.add(WF_TAGS_CHANGE, mo.plan.fv, 0);
            wb.end_phase();
            wf_run(wb, a, arena, dim, cfg.pred_radius);
            fused_flag = true;
        }
        else
        {
            SMR_CUDA(cudaMemsetAsync(mo.d_detail.p, 0, static_cast<size_t>(n) * sizeof(double) * ncomp, g.stream));
            SMR_CUDA(cudaMemsetAsync(mo.d_tag.p, 0, static_cast<size_t>(n), g.stream));
            mg_barrier(); // local memsets done everywhere before any peer stores a tag
            launch<smr_item_fv>(SMR_FAM_KEEP, arena, mo.plan.fv, KeepLeavesOp{tag, mo.filter.mask_all()});
            for (auto* f : fields)
            {
                do_update_ghost(*f);
            }
            {
                // detail for every coarse level < L - ite in one launch (levels are independent: mr/adapt.hpp:310-317)
                const int64_t limit = mo.plan.detail_cum[std::max(std::min(L - ite, mo.mesh.nlev), 0)];
                for (int c = 0; c < ncomp; ++c)
                {
                    const double* u = static_cast<const double*>(fields[c]->data.p);
                    if (cfg.pred_radius == 0)
                    {
                        launch_dim<DetailOp0, smr_item_detail>(SMR_FAM_DETAIL, dim, arena, mo.plan.detail, limit, u, detail + c * n);
                    }
                    else
                    {
                        launch_dim<DetailOp1, smr_item_detail>(SMR_FAM_DETAIL, dim, arena, mo.plan.detail, limit, u, detail + c * n);
                    }
                }
            }
            if (relative_detail)
            {
                relative_detail_pass(mo, fields, detail, n);
            }
            {
                // criteria for every fine level <= L - ite in one launch (disjoint tag writes, detail is read-only)
                const int64_t limit = (L - ite) >= 0 ? mo.plan.tag_cum[L - ite] : 0;
                switch (dim)
                {
                    case 1:
                        launch<smr_item_tag>(SMR_FAM_CRITERIA, arena, mo.plan.tag_all, CriteriaOp<1>{detail, tag, tp, ncomp, n}, limit);
                        break;
                    case 2:
                        launch<smr_item_tag>(SMR_FAM_CRITERIA, arena, mo.plan.tag_all, CriteriaOp<2>{detail, tag, tp, ncomp, n}, limit);
                        break;
                    default:
                        launch<smr_item_tag>(SMR_FAM_CRITERIA, arena, mo.plan.tag_all, CriteriaOp<3>{detail, tag, tp, ncomp, n}, limit);
                        break;
                }
            }
            if (cfg.refine_boundary) // keep_boundary_refined, mr/adapt.hpp:340-345
            {
                launch<smr_item_fv>(SMR_FAM_KEEP, arena, mo.plan.keep_bdry, KeepLeavesOp{tag, mo.filter.mask_all()});
            }
            for (int level = L; level >= 1; --level)
            {
                for (int k = 0; k < dim; ++k)
                {
                    if (cfg.periodic[k])
                    {
                        launch<smr_item_copy>(SMR_FAM_MAXIMUM, arena, mo.plan.down[level].per[k], TagOrOp{tag, mo.filter.mask_all()});
                    }
                }
                launch_dim<MaximumOpR, smr_item_tag>(SMR_FAM_MAXIMUM, dim, arena, mo.plan.tag[level], -1, tag);
            }
        }
        mo.h_tag.ensure(static_cast<size_t>(flag_at) + 16);
        // Tags go back to the host for the graduation (1 B per reference cell).  The fused launch leaves a "some leaf changes"
        // flag behind the tags: it is fetched first, and when it is clear on a graduated mesh the mesh is a fixed point and the
        // tag array itself stays on the device (smr_adapt_last_tags fetches it on demand).
        bool flag_clear = false;
        if (fused_flag && mo.graduated)
        {
            SMR_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(mo.h_tag.p) + flag_at, tag + flag_at, 16, cudaMemcpyDeviceToHost, g.stream));
            sec1.close();
            wf_fetch_error();
            SMR_CUDA(cudaStreamSynchronize(g.stream));
            wf_check_error();
            g.stats.d2h_bytes += 16;
            flag_clear = *reinterpret_cast<const unsigned*>(static_cast<const uint8_t*>(mo.h_tag.p) + flag_at) == 0u;
        }
        mo.h_tag_valid = !flag_clear;
        if (!flag_clear)
        {
            Section sec_tags;
            SMR_CUDA(cudaMemcpyAsync(mo.h_tag.p, tag, static_cast<size_t>(fused_flag ? flag_at + 16 : n), cudaMemcpyDeviceToHost, g.stream));
            sec_tags.close();
            sec1.close();
            wf_fetch_error();
            SMR_CUDA(cudaStreamSynchronize(g.stream));
            wf_check_error();
            g.stats.d2h_bytes += static_cast<uint64_t>(n);
        }
        mo.last_size  = n;
        mo.last_ncomp = ncomp;

        // host: new leaves from tags, graduation, fixed-point test (mr/adapt.hpp:360-379)
        ++g.stats.harten_iterations;
        double t0    = now();
        double ts    = t0;
        auto stage   = [&](int k)
        {
            const double t = now();
            g.stats.host_stage_seconds[k] += t - ts;
            ts = t;
        };
        bool no_tag_changes = false;
        CellArray ca;
        if (flag_clear)
        {
            no_tag_changes = true; // the device already looked at every leaf tag
        }
        else
        {
            ca = cells_from_tags(mo.mesh, static_cast<const uint8_t*>(mo.h_tag.p), &no_tag_changes);
        }
        stage(0);
        if (no_tag_changes && mo.graduated)
        {
            // no leaf refined or coarsened, and the leaves are a fixed point of make_graduation: mesh == new_mesh
            g.stats.host_mesh_seconds += now() - t0;
            return true;
        }
        if (no_tag_changes)
        {
            ca = mo.mesh.cells;
            for (LevelSet& s : ca)
            {
                s.off.clear();
            }
        }
        // every mesh held here came out of make_graduation (or is uniform), and make_graduation leaves a graduated
        // cell array untouched, so it only has to run when the tags changed something
        bool same = same_cells(ca, mo.mesh.cells);
        stage(1);
        if (!same || !mo.graduated)
        {
            make_graduation(cfg, ca);
            stage(2);
            same = same_cells(ca, mo.mesh.cells);
            stage(1);
        }
        mo.graduated = true;
        if (same)
        {
            g.stats.host_mesh_seconds += now() - t0;
            return true;
        }
        auto new_mesh = std::make_unique<Mesh>();
        new_mesh->generation = mo.mesh.generation;
        new_mesh->init_from_cells(cfg, std::move(ca));
        stage(3);
        g.stats.host_mesh_seconds += now() - t0;
        ++g.stats.mesh_rebuilds;

        // the new mesh takes its place (building its batches on a second host thread behind the transfer was tried: two OpenMP
        // teams on the 16-core box made the step slower, 6.2 -> 10.8 ms, also with the cores split between the teams)
        const int64_t nn = new_mesh->nref;
        Mesh old_mesh    = std::move(mo.mesh);
        mo.mesh          = std::move(*new_mesh);
        new_mesh.reset();
        mo.invalidate_plans();
        mo.retire_csr();
        // update_fields (algorithm/update_fields.hpp:27-54,101-127)
        t0 = now();
        TransferPlan& tpn = g.transfer;
        SMR_CUDA(cudaStreamSynchronize(g.stream)); // the previous transfer upload must have left the staging arena
        ts = now();
        g.stats.host_stage_seconds[6] += ts - t0;
        build_transfer(old_mesh, mo.mesh, tpn, mo.filter);
        stage(4);
        g.stats.host_batch_seconds += now() - t0;
        DevBuf& d_tr = g.d_transfer;
        d_tr.min_cap = size_t(16) << 20;
        for (auto* f : fields)
        {
            f->spare.ensure(static_cast<size_t>(nn) * sizeof(double));
        }
        Section sec2;
        ensure_csr(mo);
        upload_arena(tpn.arena, d_tr);
        run_derive(d_tr, tpn.derive, mo, true);
        if (wf_enabled() && fields.size() <= SMR_WF_MAX_FIELDS)
        {
            // copy / projection / prediction of every field: independent jobs of one phase, one launch
            WfArgs a{};
            WfBuilder wb;
            wb.dim = dim;
            wb.keep_empty = wf_multi();
            wb.begin_phase();
            for (size_t i = 0; i < fields.size(); ++i)
            {
                SMR_CUDA(cudaMemsetAsync(fields[i]->spare.p, 0, static_cast<size_t>(nn) * sizeof(double), g.stream));
                a.src[i] = static_cast<const double*>(fields[i]->data.p);
                a.dst[i] = static_cast<double*>(fields[i]->spare.p);
                wb.add(WF_COPY, tpn.copy, static_cast<int>(i));
                wb.add(WF_PROJ, tpn.proj, static_cast<int>(i));
                wb.add(WF_PRED, tpn.pred, static_cast<int>(i));
            }
            mg_barrier(); // multi-GPU: every rank has zeroed its new buffers before any peer stores halo values into them
            wf_run(wb, a, d_tr.p, dim, cfg.pred_radius);
        }
        else
        {
            for (auto* f : fields)
            {
                DevBuf* nb = &f->spare;
                SMR_CUDA(cudaMemsetAsync(nb->p, 0, static_cast<size_t>(nn) * sizeof(double), g.stream));
                mg_barrier();
                const double* src = static_cast<const double*>(f->data.p);
                double* dst       = static_cast<double*>(nb->p);
                launch<smr_item_copy>(SMR_FAM_COPY, d_tr.p, tpn.copy, CopyOp{src, dst});
                launch_dim<ProjOp, smr_item_proj>(SMR_FAM_PROJ, dim, d_tr.p, tpn.proj, -1, src, dst);
                launch_pred(dim, cfg.pred_radius, d_tr.p, tpn.pred, src, dst);
            }
        }
        sec2.close();
        for (size_t i = 0; i < fields.size(); ++i)
        {
            fields[i]->data.swap(fields[i]->spare);
            fields[i]->n            = nn;
            fields[i]->ghosts_valid = false;
        }
        const double tm = now();
        old_mesh        = Mesh();
        g.stats.host_stage_seconds[7] += now() - tm;
        return false;
    }

    // flux coefficients {left, right} of direction d for cells of size h
    static void scheme_coeffs(int kind, const double* params, double scale, int d, double h, double fc[2])
    {
        if (kind == SMR_SCHEME_CONVECTION_UPWIND) // operators/convection_lin.hpp:31-72
        {
            const double v = params[d];
            fc[0]          = v >= 0 ? v : 0.0;
            fc[1]          = v >= 0 ? 0.0 : v;
        }
        else // operators/diffusion.hpp:143-170
        {
            fc[0] = -1 / h;
            fc[1] = 1 / h;
            fc[0] *= -params[d];
            fc[1] *= -params[d];
        }
        if (scale != 1) // scalar * scheme: flux_based/algebraic_operators.hpp:20-28
        {
            fc[0] *= scale;
            fc[1] *= scale;
        }
    }

    static double h_factor(int dim, double h_face, double h_cell) // flux_based_scheme__lin_hom.hpp:62-67
    {
        return std::pow(h_face, dim - 1) / std::pow(h_cell, dim);
    }

    template <int NONLIN>
    static void launch_flux_general(int dim, MeshObj& mo, const double* u, double* o, double scale)
    {
        const int64_t* aux = reinterpret_cast<const int64_t*>(static_cast<const char*>(mo.d_flux.p) + mo.flux.items.aux);
        const double* tab  = static_cast<const double*>(mo.d_fluxtab.p);
        switch (dim)
        {
            case 1:
                launch<smr_item_flux>(SMR_FAM_FV, mo.d_flux.p, mo.flux.items, FluxGenOp<1, NONLIN>{u, o, aux, tab, scale});
                break;
            case 2:
                launch<smr_item_flux>(SMR_FAM_FV, mo.d_flux.p, mo.flux.items, FluxGenOp<2, NONLIN>{u, o, aux, tab, scale});
                break;
            default:
                launch<smr_item_flux>(SMR_FAM_FV, mo.d_flux.p, mo.flux.items, FluxGenOp<3, NONLIN>{u, o, aux, tab, scale});
                break;
        }
    }

    static void wide_tab(const MeshConfig& cfg, std::vector<double>& tab)
    {
        tab.assign(2 * SMR_MAX_LEVELS * 6, 0.0);
        for (int l = 0; l <= cfg.max_level && l < SMR_MAX_LEVELS; ++l)
        {
            const double h = cfg.cell_length(l), hf = cfg.cell_length(l + 1);
            tab[static_cast<size_t>(l) * 6]                    = h_factor(cfg.dim, h, h);
            tab[static_cast<size_t>(SMR_MAX_LEVELS + l) * 6] = h_factor(cfg.dim, hf, h);
        }
    }

    // calls fn(op) with the FluxWenoOp instantiation of (dim, kind, n_comp)
    template <class Fn>
    static void with_weno_op(int dim, int kind, int nc, const double* const* u, double* const* o, const int64_t* aux, const double* tab,
                             const double* params, double scale, Fn&& fn)
    {
        const double v[3] = {params ? params[0] : 0.0, params && dim > 1 ? params[1] : 0.0, params && dim > 2 ? params[2] : 0.0};
        const bool nonlin = kind == SMR_SCHEME_CONVECTION_WENO5_NONLINEAR;
#define SMR_WENO_CASE(D, N, C)                                                                                             \
    if (dim == D && nonlin == (N != 0) && nc == C)                                                                         \
    {                                                                                                                      \
        fn(FluxWenoOp<D, N, C>{{u[0], C > 1 ? u[1] : nullptr, C > 2 ? u[2] : nullptr}, {o[0], C > 1 ? o[1] : nullptr, C > 2 ? o[2] : nullptr}, \
                               aux, tab, {v[0], v[1], v[2]}, scale});                                                       \
        return;                                                                                                            \
    }
        SMR_WENO_CASE(1, 0, 1)
        SMR_WENO_CASE(2, 0, 1)
        SMR_WENO_CASE(3, 0, 1)
        SMR_WENO_CASE(1, 1, 1)
        SMR_WENO_CASE(2, 1, 1)
        SMR_WENO_CASE(3, 1, 1)
        SMR_WENO_CASE(2, 1, 2)
        SMR_WENO_CASE(3, 1, 3)
#undef SMR_WENO_CASE
        throw std::invalid_argument("make_convection_weno5: scalar fields, or (without a velocity) vector fields with n_comp == dim");
    }

    // WENO5 schemes (six-cell line stencil): ghost width 3; boundary conditions that fill three layers (Dirichlet<3>) are not built, so
    // they run on fully periodic meshes (demos/FiniteVolume/linear_convection.cpp)
    static void apply_wide_scheme(MeshObj& mo, FieldObj* const* in, FieldObj* const* out, int nc, int kind, const double* params, double scale)
    {
        const MeshConfig& cfg = mo.mesh.cfg;
        if (!cfg.all_periodic() || cfg.ghost_width() < 3)
        {
            throw std::invalid_argument("make_convection_weno5 needs a fully periodic mesh with max_stencil_size(6)");
        }
        ensure_plan(mo);
        for (int c = 0; c < nc; ++c)
        {
            if (!in[c]->ghosts_valid)
            {
                do_update_ghost(*in[c]);
            }
        }
        if (!mo.fluxw.ready)
        {
            const double t0 = now();
            build_fluxw_plan(mo.mesh, mo.fluxw, mo.filter);
            g.stats.host_batch_seconds += now() - t0;
            SMR_CUDA(cudaStreamSynchronize(g.stream));
            upload_arena(mo.fluxw.arena, mo.d_fluxw);
        }
        const size_t bytes = static_cast<size_t>(mo.mesh.nref) * sizeof(double);
        for (int c = 0; c < nc; ++c)
        {
            if (bytes > out[c]->data.cap)
            {
                SMR_CUDA(cudaStreamSynchronize(g.stream));
            }
            out[c]->data.ensure(bytes);
            out[c]->n            = mo.mesh.nref;
            out[c]->ghosts_valid = false;
        }
        Section sec;
        mg_barrier();
        for (int c = 0; c < nc; ++c)
        {
            SMR_CUDA(cudaMemsetAsync(out[c]->data.p, 0, bytes, g.stream)); // output.fill(0)
        }
        mg_barrier();
        static thread_local std::vector<double> wtab;
        wide_tab(cfg, wtab);
        mo.d_fluxtab.ensure(wtab.size() * sizeof(double));
        SMR_CUDA(cudaMemcpyAsync(mo.d_fluxtab.p, wtab.data(), wtab.size() * sizeof(double), cudaMemcpyHostToDevice, g.stream));
        const double* up[3] = {nullptr, nullptr, nullptr};
        double* op[3]       = {nullptr, nullptr, nullptr};
        for (int c = 0; c < nc; ++c)
        {
            up[c] = static_cast<const double*>(in[c]->data.p);
            op[c] = static_cast<double*>(out[c]->data.p);
        }
        const int64_t* aux = reinterpret_cast<const int64_t*>(static_cast<const char*>(mo.d_fluxw.p) + mo.fluxw.items.aux);
        with_weno_op(cfg.dim, kind, nc, up, op, aux, static_cast<const double*>(mo.d_fluxtab.p), params, scale,
                     [&](const auto& opr) { launch<smr_item_fluxw>(SMR_FAM_FV, mo.d_fluxw.p, mo.fluxw.items, opr); });
    }

    template <class F>
    static int guarded(F&& f)
    {
        try
        {
            f();
            return SMR_OK;
        }
        catch (const std::out_of_range& e)
        {
            g.err = e.what();
            return SMR_ERR_OUT_OF_RANGE;
        }
        catch (const std::invalid_argument& e)
        {
            g.err = e.what();
            return SMR_ERR_INVALID;
        }
        catch (const CudaError& e)
        {
            g.err = e.what();
            return SMR_ERR_CUDA;
        }
        catch (const std::exception& e)
        {
            g.err = e.what();
            return SMR_ERR_INTERNAL;
        }
    }
} // namespace smr

using namespace smr;

extern "C"
{
    int smr_init(int device)
    {
        return guarded(
            [&]
            {
                if (device < 0)
                {
                    g.device = false;
                    return;
                }
                int n = 0;
                if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
                {
                    g.device = false;
                    throw CudaError("no CUDA device visible");
                }
                if (device >= n)
                {
                    throw CudaError("device index out of range");
                }
                SMR_CUDA(cudaSetDevice(device));
                Arena::alloc_fn = pinned_alloc;
                Arena::free_fn  = pinned_free;
                g.dev    = device;
                g.stream = nullptr; // legacy default stream unless smr_set_stream() is called
                g.device = true;
            });
    }

    int smr_finalize(void)
    {
        return guarded(
            [&]
            {
                if (g.device)
                {
                    cudaStreamSynchronize(g.stream);
                }
                g.fields.clear();
                g.meshes.clear();
            });
    }

    const char* smr_last_error(void)
    {
        return g.err.c_str();
    }

    int smr_device_available(void)
    {
        return g.device ? 1 : 0;
    }

    int smr_set_stream(void* s)
    {
        return guarded(
            [&]
            {
                require_device();
                SMR_CUDA(cudaStreamSynchronize(g.stream));
                g.stream = static_cast<cudaStream_t>(s);
            });
    }

    int smr_synchronize(void)
    {
        return guarded(
            [&]
            {
                require_device();
                wf_fetch_error();
                SMR_CUDA(cudaStreamSynchronize(g.stream));
                wf_check_error();
            });
    }

    int smr_mesh_create_uniform(const smr_mesh_config* cfg, int level, smr_mesh_t* out)
    {
        return guarded(
            [&]
            {
                MeshConfig c;
                config_to_mesh_cfg(cfg, c);
                if (level < 0 || level > c.max_level)
                {
                    throw std::invalid_argument("start level out of range");
                }
                const double t0 = now();
                auto mo         = std::make_unique<MeshObj>();
                mo->mesh.init_uniform(c, level);
                mo->graduated = true; // a uniform mesh is trivially graduated
                g.stats.host_mesh_seconds += now() - t0;
                *out = g.next_id++;
                g.meshes[*out] = std::move(mo);
            });
    }

    int smr_mesh_create_from_intervals(const smr_mesh_config* cfg, const int32_t* levels, const smr_interval* ivl, int64_t n, smr_mesh_t* out)
    {
        return guarded(
            [&]
            {
                MeshConfig c;
                config_to_mesh_cfg(cfg, c);
                const int nlev = Mesh::levels_for(c);
                std::vector<SetBuilder> b(nlev);
                for (int64_t i = 0; i < n; ++i)
                {
                    if (levels[i] < 0 || levels[i] > c.max_level)
                    {
                        throw std::invalid_argument("interval level out of range");
                    }
                    b[levels[i]].add(mk_key(c.dim > 1 ? ivl[i].y : 0, c.dim > 2 ? ivl[i].z : 0), ivl[i].start, ivl[i].end);
                }
                CellArray ca(nlev);
                for (int l = 0; l < nlev; ++l)
                {
                    ca[l] = b[l].build();
                }
                const double t0 = now();
                auto mo         = std::make_unique<MeshObj>();
                mo->mesh.init_from_cells(c, std::move(ca));
                g.stats.host_mesh_seconds += now() - t0;
                *out = g.next_id++;
                g.meshes[*out] = std::move(mo);
            });
    }

    int smr_mesh_destroy(smr_mesh_t m)
    {
        return guarded(
            [&]
            {
                MeshObj& mo = get_mesh(m);
                if (!mo.fields.empty())
                {
                    throw std::invalid_argument("mesh still has fields attached");
                }
                if (g.device)
                {
                    cudaStreamSynchronize(g.stream);
                }
                g.meshes.erase(m);
            });
    }

    int smr_mesh_config_get(smr_mesh_t m, smr_mesh_config* out)
    {
        return guarded(
            [&]
            {
                const MeshConfig& c     = get_mesh(m).mesh.cfg;
                out->dim                = c.dim;
                out->min_level          = c.min_level;
                out->max_level          = c.max_level;
                out->pred_radius        = c.pred_radius;
                out->max_stencil_radius = c.max_stencil_radius;
                out->graduation_width   = c.graduation_width;
                for (int d = 0; d < 3; ++d)
                {
                    out->n_cells0[d] = c.n0[d];
                    out->origin[d]   = c.origin[d];
                }
                out->scaling_factor = c.scaling;
                for (int d = 0; d < 3; ++d)
                {
                    out->periodic[d] = c.periodic[d] ? 1 : 0;
                }
                out->refine_boundary = c.refine_boundary ? 1 : 0;
            });
    }

    static void check_mesh_id(int id)
    {
        if (id < 0 || id > 4)
        {
            throw std::invalid_argument("invalid mesh id");
        }
    }

    int smr_mesh_nb_cells(smr_mesh_t m, int mesh_id, int level, int64_t* out)
    {
        return guarded(
            [&]
            {
                const Mesh& mesh = get_mesh(m).mesh;
                check_mesh_id(mesh_id);
                const CellArray& ca = mesh.sub(mesh_id);
                int64_t n           = 0;
                for (int l = 0; l < mesh.nlev; ++l)
                {
                    if (level < 0 || level == l)
                    {
                        n += ca[l].n_cells();
                    }
                }
                *out = n;
            });
    }

    int smr_mesh_nb_intervals(smr_mesh_t m, int mesh_id, int level, int64_t* out)
    {
        return guarded(
            [&]
            {
                const Mesh& mesh = get_mesh(m).mesh;
                check_mesh_id(mesh_id);
                const CellArray& ca = mesh.sub(mesh_id);
                int64_t n           = 0;
                for (int l = 0; l < mesh.nlev; ++l)
                {
                    if (level < 0 || level == l)
                    {
                        n += static_cast<int64_t>(ca[l].n_intervals());
                    }
                }
                *out = n;
            });
    }

    int smr_mesh_get_intervals(smr_mesh_t m, int mesh_id, int level, smr_interval* out)
    {
        return guarded(
            [&]
            {
                const Mesh& mesh = get_mesh(m).mesh;
                check_mesh_id(mesh_id);
                if (level < 0 || level >= mesh.nlev)
                {
                    throw std::invalid_argument("level out of range");
                }
                LevelSet s = mesh.sub(mesh_id)[level];
                if (s.off.size() != s.xs.size())
                {
                    // union cells are not necessarily inside the reference mesh: report -1 offsets
                    s.off.assign(s.xs.size(), -1);
                }
                size_t k = 0;
                for (size_t r = 0; r < s.rows(); ++r)
                {
                    for (int q = s.ptr[r]; q < s.ptr[r + 1]; ++q)
                    {
                        out[k++] = {key_y(s.key[r]), key_z(s.key[r]), s.xs[q], s.xe[q], s.off[q]};
                    }
                }
            });
    }

    int smr_mesh_generation(smr_mesh_t m, uint64_t* out)
    {
        return guarded(
            [&]
            {
                *out = get_mesh(m).mesh.generation;
            });
    }

    int smr_mesh_get_index(smr_mesh_t m, int level, int i, int j, int k, int64_t* out)
    {
        return guarded(
            [&]
            {
                const Mesh& mesh = get_mesh(m).mesh;
                *out             = -1;
                if (level < 0 || level >= mesh.nlev)
                {
                    throw std::out_of_range("level out of range");
                }
                const int64_t o = mesh.ref[level].offset_of(mk_key(mesh.cfg.dim > 1 ? j : 0, mesh.cfg.dim > 2 ? k : 0), i, i);
                if (o < 0)
                {
                    missing("get_index", level, i, j, k);
                }
                *out = o;
            });
    }

    int smr_field_create(smr_mesh_t m, const char* name, smr_field_t* out)
    {
        return guarded(
            [&]
            {
                MeshObj& mo = get_mesh(m);
                auto f      = std::make_unique<FieldObj>();
                f->mesh     = &mo;
                f->name     = name ? name : "";
                mo.fields.push_back(f.get());
                *out           = g.next_id++;
                g.fields[*out] = std::move(f);
            });
    }

    int smr_field_destroy(smr_field_t fh)
    {
        return guarded(
            [&]
            {
                FieldObj& f = get_field(fh);
                if (g.device)
                {
                    cudaStreamSynchronize(g.stream);
                }
                auto& v = f.mesh->fields;
                v.erase(std::remove(v.begin(), v.end(), &f), v.end());
                g.fields.erase(fh);
            });
    }

    int smr_field_resize(smr_field_t fh)
    {
        return guarded(
            [&]
            {
                require_device();
                FieldObj& f     = get_field(fh);
                const int64_t n = f.mesh->mesh.nref;
                if (static_cast<size_t>(n) * sizeof(double) > f.data.cap)
                {
                    SMR_CUDA(cudaStreamSynchronize(g.stream));
                }
                f.data.ensure(static_cast<size_t>(n) * sizeof(double));
                if (f.n != n)
                {
                    f.ghosts_valid = false;
                }
                f.n = n;
            });
    }

    int smr_field_fill(smr_field_t fh, double v)
    {
        return guarded(
            [&]
            {
                require_device();
                FieldObj& f = get_field(fh);
                check_field_ready(f);
                f.ghosts_valid = false;
                mg_barrier();
                if (v == 0.0)
                {
                    SMR_CUDA(cudaMemsetAsync(f.data.p, 0, static_cast<size_t>(f.n) * sizeof(double), g.stream));
                    mg_barrier();
                }
                else
                {
                    std::vector<double> h(static_cast<size_t>(f.n), v);
                    SMR_CUDA(cudaMemcpyAsync(f.data.p, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, g.stream));
                    mg_barrier();
                    SMR_CUDA(cudaStreamSynchronize(g.stream));
                }
            });
    }

    int smr_field_size(smr_field_t fh, int64_t* out)
    {
        return guarded(
            [&]
            {
                *out = get_field(fh).mesh->mesh.nref;
            });
    }

    int smr_field_upload(smr_field_t fh, const double* host, int64_t n)
    {
        return guarded(
            [&]
            {
                require_device();
                FieldObj& f = get_field(fh);
                check_field_ready(f);
                if (n != f.n)
                {
                    throw std::invalid_argument("upload size does not match the field size");
                }
                f.ghosts_valid = false;
                mg_barrier(); // nobody is still storing into this buffer
                SMR_CUDA(cudaMemcpyAsync(f.data.p, host, static_cast<size_t>(n) * sizeof(double), cudaMemcpyHostToDevice, g.stream));
                mg_barrier();
                g.stats.h2d_bytes += static_cast<uint64_t>(n) * sizeof(double);
            });
    }

    int smr_field_download(smr_field_t fh, double* host, int64_t n)
    {
        return guarded(
            [&]
            {
                require_device();
                FieldObj& f = get_field(fh);
                check_field_ready(f);
                if (n != f.n)
                {
                    throw std::invalid_argument("download size does not match the field size");
                }
                SMR_CUDA(cudaMemcpyAsync(host, f.data.p, static_cast<size_t>(n) * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
                SMR_CUDA(cudaStreamSynchronize(g.stream));
                g.stats.d2h_bytes += static_cast<uint64_t>(n) * sizeof(double);
            });
    }

    int smr_field_swap(smr_field_t ah, smr_field_t bh)
    {
        return guarded(
            [&]
            {
                FieldObj& a = get_field(ah);
                FieldObj& b = get_field(bh);
                if (a.mesh != b.mesh)
                {
                    throw std::invalid_argument("fields live on different meshes");
                }
                a.data.swap(b.data);
                std::swap(a.n, b.n);
                std::swap(a.ghosts_valid, b.ghosts_valid);
            });
    }

    int smr_field_set_bc(smr_field_t fh, int bc_type, double value)
    {
        return guarded(
            [&]
            {
                FieldObj& f = get_field(fh);
                if (bc_type != SMR_BCTYPE_DIRICHLET && bc_type != SMR_BCTYPE_NEUMANN)
                {
                    throw std::invalid_argument("unknown boundary condition type");
                }
                if (f.bc_type != bc_type || f.bc_value != value)
                {
                    f.ghosts_valid = false; // the boundary ghosts depend on the condition: the next update must not be skipped
                }
                f.bc_type  = bc_type;
                f.bc_value = value;
            });
    }

    int smr_update_ghost_mr(smr_field_t fh)
    {
        return guarded(
            [&]
            {
                require_device();
                FieldObj& f = get_field(fh);
                if (f.ghosts_valid && f.n == f.mesh->mesh.nref)
                {
                    // the ghosts are a pure function of the leaves (and of cells no phase writes, which are unchanged):
                    // recomputing them would store the same bits (update_ghost_mr_if_needed, update_ghost_mr.hpp:242-249)
                    ++g.stats.ghost_updates_skipped;
                    return;
                }
                ensure_plan(*f.mesh);
                Section sec;
                do_update_ghost(f);
            });
    }

    int smr_fv_upwind(smr_field_t out, smr_field_t in, const double* a, double dt)
    {
        return guarded(
            [&]
            {
                do_fv(get_field(out), get_field(in), a, dt, false);
            });
    }

    int smr_fv_upwind_burgers(smr_field_t out, smr_field_t in, const double* k, double dt)
    {
        return guarded(
            [&]
            {
                do_fv(get_field(out), get_field(in), k, dt, true);
            });
    }

    int smr_scheme_apply(smr_field_t outh, smr_field_t inh, int kind, const double* params, double scale)
    {
        return guarded(
            [&]
            {
                require_device();
                FieldObj& out = get_field(outh);
                FieldObj& in  = get_field(inh);
                if (out.mesh != in.mesh || &out == &in)
                {
                    throw std::invalid_argument("scheme output must be a different field on the same mesh");
                }
                MeshObj& mo = *in.mesh;
                check_field_ready(in);
                const MeshConfig& cfg = mo.mesh.cfg;
                if (kind != SMR_SCHEME_CONVECTION_UPWIND && kind != SMR_SCHEME_DIFFUSION_ORDER2 && kind != SMR_SCHEME_CONVECTION_UPWIND_NONLINEAR
                    && kind != SMR_SCHEME_CONVECTION_WENO5 && kind != SMR_SCHEME_CONVECTION_WENO5_NONLINEAR)
                {
                    throw std::invalid_argument("unknown scheme kind");
                }
                if (kind == SMR_SCHEME_CONVECTION_WENO5 || kind == SMR_SCHEME_CONVECTION_WENO5_NONLINEAR)
                {
                    FieldObj* ip[1] = {&in};
                    FieldObj* op[1] = {&out};
                    apply_wide_scheme(mo, ip, op, 1, kind, params, scale);
                    return;
                }
                const bool nonlin = kind == SMR_SCHEME_CONVECTION_UPWIND_NONLINEAR;
                static const bool force_general = std::getenv("SMR_FLUX_GENERAL") != nullptr;
                const int level    = mo.mesh.min_leaf_level();
                // periodic meshes: the interfaces through the boundary (interface.hpp:83-92, 179-189, 280-290) are classified in the
                // general records; the uniform-level strip kernels know boundaries only
                const bool general = nonlin || force_general || level != mo.mesh.max_leaf_level() || cfg.any_periodic();
                // update_ghosts_if_needed (schemes/fv/FV_scheme.hpp:187-197)
                ensure_plan(mo);
                if (!in.ghosts_valid)
                {
                    do_update_ghost(in);
                }
                if (general && !mo.flux.ready)
                {
                    const double t0 = now();
                    build_flux_plan(mo.mesh, mo.flux, mo.filter);
                    g.stats.host_batch_seconds += now() - t0;
                    SMR_CUDA(cudaStreamSynchronize(g.stream));
                    upload_arena(mo.flux.arena, mo.d_flux);
                }
                if (static_cast<size_t>(mo.mesh.nref) * sizeof(double) > out.data.cap)
                {
                    SMR_CUDA(cudaStreamSynchronize(g.stream));
                }
                out.data.ensure(static_cast<size_t>(mo.mesh.nref) * sizeof(double));
                out.n = mo.mesh.nref;
                Section sec;
                out.ghosts_valid = false;
                mg_barrier();
                SMR_CUDA(cudaMemsetAsync(out.data.p, 0, static_cast<size_t>(out.n) * sizeof(double), g.stream)); // output.fill(0)
                mg_barrier();
                const double* u = static_cast<const double*>(in.data.p);
                double* o       = static_cast<double*>(out.data.p);
                if (general)
                {
                    // coefficient tables per level (flux_based_scheme__lin_hom.hpp:94-165, __nonlin.hpp:434-505)
                    static thread_local std::vector<double> tab;
                    tab.assign(2 * SMR_MAX_LEVELS * 6, 0.0);
                    for (int l = 0; l <= cfg.max_level && l < SMR_MAX_LEVELS; ++l)
                    {
                        const double h = cfg.cell_length(l), hf = cfg.cell_length(l + 1);
                        double* cs     = tab.data() + l * 6;
                        double* cj     = tab.data() + (SMR_MAX_LEVELS + l) * 6;
                        if (nonlin)
                        {
                            cs[0] = h_factor(cfg.dim, h, h);
                            cj[0] = h_factor(cfg.dim, hf, h);
                            continue;
                        }
                        for (int d = 0; d < cfg.dim; ++d)
                        {
                            double fc[2];
                            scheme_coeffs(kind, params, scale, d, h, fc);
                            cs[2 * d]     = h_factor(cfg.dim, h, h) * fc[0];
                            cs[2 * d + 1] = h_factor(cfg.dim, h, h) * fc[1];
                            scheme_coeffs(kind, params, scale, d, hf, fc); // flux computed at level + 1
                            cj[2 * d]     = h_factor(cfg.dim, hf, h) * fc[0];
                            cj[2 * d + 1] = h_factor(cfg.dim, hf, h) * fc[1];
                        }
                    }
                    mo.d_fluxtab.ensure(tab.size() * sizeof(double));
                    SMR_CUDA(cudaMemcpyAsync(mo.d_fluxtab.p, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, g.stream));
                    if (nonlin)
                    {
                        launch_flux_general<1>(cfg.dim, mo, u, o, scale);
                    }
                    else
                    {
                        launch_flux_general<0>(cfg.dim, mo, u, o, scale);
                    }
                    return;
                }
                FluxParams p;
                const double h      = cfg.cell_length(level);
                const double factor = h_factor(cfg.dim, h, h);
                for (int d = 0; d < 3; ++d)
                {
                    double fc[2] = {0, 0};
                    if (d < cfg.dim)
                    {
                        scheme_coeffs(kind, params, scale, d, h, fc);
                    }
                    p.lc[d][0] = factor * fc[0];
                    p.lc[d][1] = factor * fc[1];
                    p.n[d]     = cfg.n0[d] << level;
                }
                if (cfg.dim > 1)
                {
                    if (g.profile)
                    {
                        g.prof_cells[SMR_FAM_FV] += static_cast<uint64_t>(mo.plan.fv_strip.n_cells) * (SMR_STRIP_ROWS - 1);
                    }
                    launch_dim<FluxLinHomStripOp, smr_item_fvstrip>(SMR_FAM_FV, cfg.dim, mo.d_arena.p, mo.plan.fv_strip, -1, u, o, p);
                    launch_dim<FluxLinHomOp, smr_item_fv>(SMR_FAM_FV, cfg.dim, mo.d_arena.p, mo.plan.fv_single, -1, u, o, p);
                }
                else
                {
                    launch_dim<FluxLinHomOp, smr_item_fv>(SMR_FAM_FV, cfg.dim, mo.d_arena.p, mo.plan.fv, -1, u, o, p);
                }
            });
    }

    int smr_scheme_apply_vector(const smr_field_t* outh, const smr_field_t* inh, int n_comp, int kind, const double* params, double scale)
    {
        return guarded(
            [&]
            {
                require_device();
                if (kind != SMR_SCHEME_CONVECTION_UPWIND_NONLINEAR && kind != SMR_SCHEME_CONVECTION_WENO5_NONLINEAR)
                {
                    throw std::invalid_argument("vector fields: make_convection_upwind<VectorField>() and make_convection_weno5<VectorField>() are the vector schemes; apply the linear schemes per component");
                }
                if (n_comp < 2 || n_comp > 3)
                {
                    throw std::invalid_argument("make_convection_upwind() needs n_comp == dim (convection_nonlin.hpp:39-40)");
                }
                std::vector<FieldObj*> in, out;
                for (int c = 0; c < n_comp; ++c)
                {
                    in.push_back(&get_field(inh[c]));
                    out.push_back(&get_field(outh[c]));
                }
                MeshObj& mo           = *in[0]->mesh;
                const MeshConfig& cfg = mo.mesh.cfg;
                if (n_comp != cfg.dim)
                {
                    throw std::invalid_argument("make_convection_upwind() needs n_comp == dim (convection_nonlin.hpp:39-40)");
                }
                for (int c = 0; c < n_comp; ++c)
                {
                    if (in[c]->mesh != &mo || out[c]->mesh != &mo)
                    {
                        throw std::invalid_argument("scheme output must be a different field on the same mesh");
                    }
                    for (int k = 0; k < n_comp; ++k)
                    {
                        if (out[c] == in[k] || (k != c && out[c] == out[k]))
                        {
                            throw std::invalid_argument("scheme output must be a different field on the same mesh");
                        }
                    }
                    check_field_ready(*in[c]);
                }
                if (kind == SMR_SCHEME_CONVECTION_WENO5_NONLINEAR)
                {
                    apply_wide_scheme(mo, in.data(), out.data(), n_comp, kind, params, scale);
                    return;
                }
                ensure_plan(mo);
                for (int c = 0; c < n_comp; ++c)
                {
                    if (!in[c]->ghosts_valid)
                    {
                        do_update_ghost(*in[c]);
                    }
                }
                if (!mo.flux.ready)
                {
                    const double t0 = now();
                    build_flux_plan(mo.mesh, mo.flux, mo.filter);
                    g.stats.host_batch_seconds += now() - t0;
                    SMR_CUDA(cudaStreamSynchronize(g.stream));
                    upload_arena(mo.flux.arena, mo.d_flux);
                }
                const size_t bytes = static_cast<size_t>(mo.mesh.nref) * sizeof(double);
                for (int c = 0; c < n_comp; ++c)
                {
                    if (bytes > out[c]->data.cap)
                    {
                        SMR_CUDA(cudaStreamSynchronize(g.stream));
                    }
                    out[c]->data.ensure(bytes);
                    out[c]->n            = mo.mesh.nref;
                    out[c]->ghosts_valid = false;
                }
                Section sec;
                mg_barrier();
                for (int c = 0; c < n_comp; ++c)
                {
                    SMR_CUDA(cudaMemsetAsync(out[c]->data.p, 0, bytes, g.stream)); // output.fill(0)
                }
                mg_barrier();
                static thread_local std::vector<double> vtab;
                vtab.assign(2 * SMR_MAX_LEVELS * 6, 0.0);
                for (int l = 0; l <= cfg.max_level && l < SMR_MAX_LEVELS; ++l)
                {
                    const double h = cfg.cell_length(l), hf = cfg.cell_length(l + 1);
                    vtab[static_cast<size_t>(l) * 6]                    = h_factor(cfg.dim, h, h);
                    vtab[static_cast<size_t>(SMR_MAX_LEVELS + l) * 6] = h_factor(cfg.dim, hf, h);
                }
                mo.d_fluxtab.ensure(vtab.size() * sizeof(double));
                SMR_CUDA(cudaMemcpyAsync(mo.d_fluxtab.p, vtab.data(), vtab.size() * sizeof(double), cudaMemcpyHostToDevice, g.stream));
                const int64_t* aux = reinterpret_cast<const int64_t*>(static_cast<const char*>(mo.d_flux.p) + mo.flux.items.aux);
                const double* tab  = static_cast<const double*>(mo.d_fluxtab.p);
                const double* up[3] = {nullptr, nullptr, nullptr};
                double* op[3]       = {nullptr, nullptr, nullptr};
                for (int c = 0; c < n_comp; ++c)
                {
                    up[c] = static_cast<const double*>(in[c]->data.p);
                    op[c] = static_cast<double*>(out[c]->data.p);
                }
                if (cfg.dim == 2)
                {
                    launch<smr_item_flux>(SMR_FAM_FV, mo.d_flux.p, mo.flux.items, FluxVecOp<2>{{up[0], up[1], up[2]}, {op[0], op[1], op[2]}, aux, tab, scale});
                }
                else
                {
                    launch<smr_item_flux>(SMR_FAM_FV, mo.d_flux.p, mo.flux.items, FluxVecOp<3>{{up[0], up[1], up[2]}, {op[0], op[1], op[2]}, aux, tab, scale});
                }
            });
    }

    int smr_field_lincomb(smr_field_t outh, double a, smr_field_t xh, double b, smr_field_t yh)
    {
        return guarded(
            [&]
            {
                require_device();
                FieldObj& out = get_field(outh);
                FieldObj& x   = get_field(xh);
                FieldObj& y   = get_field(yh);
                if (out.mesh != x.mesh || out.mesh != y.mesh)
                {
                    throw std::invalid_argument("fields live on different meshes");
                }
                MeshObj& mo = *x.mesh;
                check_field_ready(x);
                check_field_ready(y);
                ensure_plan(mo);
                if (static_cast<size_t>(mo.mesh.nref) * sizeof(double) > out.data.cap)
                {
                    SMR_CUDA(cudaStreamSynchronize(g.stream));
                }
                out.data.ensure(static_cast<size_t>(mo.mesh.nref) * sizeof(double));
                out.n = mo.mesh.nref;
                Section sec;
                out.ghosts_valid = false;
                launch<smr_item_fv>(SMR_FAM_FV,
                                    mo.d_arena.p,
                                    mo.plan.fv,
                                    LinCombOp{static_cast<const double*>(x.data.p), static_cast<const double*>(y.data.p), static_cast<double*>(out.data.p), a, b, a == 1.0});
            });
    }

    static std::vector<FieldObj*> collect_fields(const smr_field_t* fields, int n)
    {
        if (n < 1 || n > 8)
        {
            throw std::invalid_argument("make_MRAdapt needs between 1 and 8 fields");
        }
        std::vector<FieldObj*> v;
        for (int i = 0; i < n; ++i)
        {
            v.push_back(&get_field(fields[i]));
        }
        return v;
    }

    int smr_adapt_iteration(const smr_field_t* fields, int n_fields, double epsilon, double regularity, int ite, int* unchanged)
    {
        return guarded(
            [&]
            {
                auto v     = collect_fields(fields, n_fields);
                *unchanged = do_harten(v, epsilon, regularity, ite) ? 1 : 0;
            });
    }

    int smr_adapt(const smr_field_t* fields, int n_fields, double epsilon, double regularity, int* n_iterations)
    {
        return smr_adapt_ex(fields, n_fields, epsilon, regularity, 0, n_iterations);
    }

    int smr_adapt_ex(const smr_field_t* fields, int n_fields, double epsilon, double regularity, int relative_detail, int* n_iterations)
    {
        return guarded(
            [&]
            {
                auto v                = collect_fields(fields, n_fields);
                const MeshConfig& cfg = v[0]->mesh->mesh.cfg;
                int done              = 0;
                if (cfg.min_level != cfg.max_level)
                {
                    for (int ite = 0; ite < cfg.max_level - cfg.min_level; ++ite)
                    {
                        ++done;
                        if (do_harten(v, epsilon, regularity, ite, relative_detail != 0))
                        {
                            break;
                        }
                    }
                }
                if (n_iterations)
                {
                    *n_iterations = done;
                }
            });
    }

    int smr_mesh_update_from_tags(smr_mesh_t m, const uint8_t* tags, int64_t n, int* unchanged)
    {
        return guarded(
            [&]
            {
                MeshObj& mo = get_mesh(m);
                if (n != mo.mesh.nref)
                {
                    throw std::invalid_argument("tag array size does not match nb_cells(reference)");
                }
                const double t0 = now();
                CellArray ca    = cells_from_tags(mo.mesh, tags);
                make_graduation(mo.mesh.cfg, ca);
                const bool same = same_cells(ca, mo.mesh.cells);
                *unchanged      = same ? 1 : 0;
                mo.graduated    = true;
                if (!same)
                {
                    if (g.device)
                    {
                        SMR_CUDA(cudaStreamSynchronize(g.stream));
                    }
                    Mesh nm;
                    nm.generation = mo.mesh.generation;
                    nm.init_from_cells(mo.mesh.cfg, std::move(ca));
                    mo.mesh       = std::move(nm);
                    mo.invalidate_plans();
                    mo.retire_csr();
                    ++g.stats.mesh_rebuilds;
                }
                g.stats.host_mesh_seconds += now() - t0;
            });
    }

    // ---------------------------------------------------------------------------------------------------------------
    // multi-GPU (one process per GPU; peers reached through CUDA IPC mappings of each rank's pool)
    // ---------------------------------------------------------------------------------------------------------------
    int smr_mg_init(int rank, int world, uint64_t pool_bytes)
    {
        return guarded(
            [&]
            {
                if (world < 1 || world > SMR_MAX_RANKS || rank < 0 || rank >= world)
                {
                    throw std::invalid_argument("smr_mg_init: need 0 <= rank < world <= 8");
                }
                if (!g.meshes.empty() || !g.fields.empty())
                {
                    throw std::invalid_argument("smr_mg_init must be called before any mesh or field exists");
                }
                g.mg_rank      = rank;
                g.mg_world     = world;
                g.mg_connected = false;
                if (world > 1 && g.device)
                {
                    if (pool_bytes < (1u << 20))
                    {
                        throw std::invalid_argument("smr_mg_init: pool_bytes too small");
                    }
                    SMR_CUDA(cudaMalloc(&g.mg_pool, pool_bytes));
                    SMR_CUDA(cudaMemset(g.mg_pool, 0, 4096));
                    g.mg_pool_bytes = pool_bytes;
                    g_pool.init(static_cast<char*>(g.mg_pool), pool_bytes, 4096);
                    g_pool_on = true;
                }
            });
    }

    int smr_mg_get_handle(void* out64)
    {
        return guarded(
            [&]
            {
                require_device();
                if (!g.mg_pool)
                {
                    throw std::invalid_argument("smr_mg_get_handle: no pool (world == 1?)");
                }
                static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
                cudaIpcMemHandle_t h;
                SMR_CUDA(cudaIpcGetMemHandle(&h, g.mg_pool));
                std::memcpy(out64, &h, 64);
            });
    }

    int smr_mg_connect(const void* handles)
    {
        return guarded(
            [&]
            {
                require_device();
                if (g.mg_world < 2)
                {
                    return;
                }
                PeerTable t{};
                t.rank  = g.mg_rank;
                t.world = g.mg_world;
                t.flags = static_cast<unsigned long long*>(g.mg_pool);
                t.error = reinterpret_cast<unsigned long long*>(static_cast<char*>(g.mg_pool) + 1024);
                for (int p = 0; p < g.mg_world; ++p)
                {
                    if (p == g.mg_rank)
                    {
                        g.mg_peer_base[p] = g.mg_pool;
                    }
                    else
                    {
                        cudaIpcMemHandle_t h;
                        std::memcpy(&h, static_cast<const char*>(handles) + 64 * p, 64);
                        SMR_CUDA(cudaIpcOpenMemHandle(&g.mg_peer_base[p], h, cudaIpcMemLazyEnablePeerAccess));
                    }
                    t.delta[p] = static_cast<long long>(static_cast<char*>(g.mg_peer_base[p]) - static_cast<char*>(g.mg_pool));
                }
                SMR_CUDA(set_peer_table(t));
                g.mg_connected = true;
            });
    }

    // every rank re-stores the reference cells it owns into all peers: afterwards each rank holds the complete field
    static void do_broadcast(FieldObj& f)
    {
        MeshObj& mo = *f.mesh;
        check_field_ready(f);
        ensure_plan(mo);
        if (g.mg_world < 2)
        {
            return;
        }
        SMR_CUDA(cudaStreamSynchronize(g.stream));
        build_broadcast(mo.mesh, mo.filter, g.transfer);
        upload_arena(g.transfer.arena, g.d_transfer);
        double* u = static_cast<double*>(f.data.p);
        launch<smr_item_copy>(SMR_FAM_COPY, g.d_transfer.p, g.transfer.copy, CopyOp{u, u});
    }

    int smr_mg_broadcast(smr_field_t fh)
    {
        return guarded(
            [&]
            {
                require_device();
                do_broadcast(get_field(fh));
                SMR_CUDA(cudaStreamSynchronize(g.stream));
                mg_check_error();
            });
    }

    int smr_mg_rebalance(const smr_field_t* fields, int n_fields)
    {
        return guarded(
            [&]
            {
                require_device();
                if (n_fields < 1)
                {
                    throw std::invalid_argument("smr_mg_rebalance needs at least one field");
                }
                MeshObj& mo = *get_field(fields[0]).mesh;
                for (int i = 0; i < n_fields; ++i)
                {
                    do_broadcast(get_field(fields[i]));
                }
                SMR_CUDA(cudaStreamSynchronize(g.stream));
                mg_check_error();
                mo.filter.rank  = g.mg_rank;
                mo.filter.world = g.mg_world;
                mo.filter.compute_cuts(mo.mesh);
                mo.invalidate_plans();
            });
    }

    int smr_mg_leaf_owners(smr_mesh_t m, int32_t* out, int64_t n)
    {
        return guarded(
            [&]
            {
                MeshObj& mo = get_mesh(m);
                if (n != mo.mesh.nleaves)
                {
                    throw std::invalid_argument("smr_mg_leaf_owners: n must equal nb_cells(cells)");
                }
                PlanFilter flt = mo.filter;
                if (flt.world != g.mg_world || flt.cut2.empty())
                {
                    flt.rank  = g.mg_rank;
                    flt.world = g.mg_world;
                    flt.compute_cuts(mo.mesh);
                }
                int64_t k = 0;
                for (int l = 0; l < mo.mesh.nlev; ++l)
                {
                    const LevelSet& c = mo.mesh.cells[l];
                    for (size_t r = 0; r < c.rows(); ++r)
                    {
                        const int o = flt.owner(l, flt.axis_coord(key_y(c.key[r]), key_z(c.key[r])));
                        for (int q = c.ptr[r]; q < c.ptr[r + 1]; ++q)
                        {
                            for (int x = c.xs[q]; x < c.xe[q]; ++x)
                            {
                                out[k++] = o;
                            }
                        }
                    }
                }
            });
    }

    int smr_debug_host_rebuild(smr_mesh_t m, int reps, double* mesh_seconds, double* plan_seconds, int64_t* arena_bytes)
    {
        return guarded(
            [&]
            {
                MeshObj& mo = get_mesh(m);
                double tm = 0, tp = 0;
                for (int r = 0; r < reps; ++r)
                {
                    CellArray ca = mo.mesh.cells;
                    double t0    = now();
                    Mesh nm;
                    nm.init_from_cells(mo.mesh.cfg, std::move(ca));
                    tm += now() - t0;
                    t0 = now();
                    MeshPlan plan;
                    build_plan(nm, plan);
                    tp += now() - t0;
                    *arena_bytes = static_cast<int64_t>(plan.arena.size);
                }
                *mesh_seconds = tm / reps;
                *plan_seconds = tp / reps;
            });
    }

    int smr_debug_flux_records(smr_mesh_t m, int32_t* out, int64_t capacity, int64_t* n_records)
    {
        return guarded(
            [&]
            {
                MeshObj& mo = get_mesh(m);
                FluxPlan fp;
                build_flux_plan(mo.mesh, fp);
                *n_records = fp.items.n_items;
                if (out == nullptr)
                {
                    return;
                }
                if (capacity < fp.items.n_items)
                {
                    throw std::invalid_argument("output array too small");
                }
                const smr_item_flux* it = reinterpret_cast<const smr_item_flux*>(fp.arena.p + fp.items.items);
                const Mesh& mesh        = mo.mesh;
                for (int i = 0; i < fp.items.n_items; ++i)
                {
                    // the record stores offsets, not coordinates: recover (x, y, z) of its first cell from the reference mesh
                    const LevelSet& ref = mesh.ref[it[i].level];
                    const auto pos      = std::upper_bound(ref.off.begin(), ref.off.end(), it[i].c) - ref.off.begin() - 1;
                    const auto row      = std::upper_bound(ref.ptr.begin(), ref.ptr.end(), static_cast<int32_t>(pos)) - ref.ptr.begin() - 1;
                    int32_t* o          = out + 6 * static_cast<int64_t>(i);
                    o[0]                = it[i].level;
                    o[1]                = ref.xs[static_cast<size_t>(pos)] + static_cast<int32_t>(it[i].c - ref.off[static_cast<size_t>(pos)]);
                    o[2]                = key_y(ref.key[static_cast<size_t>(row)]);
                    o[3]                = key_z(ref.key[static_cast<size_t>(row)]);
                    o[4]                = it[i].n;
                    o[5]                = it[i].kinds;
                }
            });
    }

    int smr_debug_fluxw_apply(smr_mesh_t m, const double* u, int n_comp, int kind, const double* velocity, double scale, double* out)
    {
        return guarded(
            [&]
            {
                MeshObj& mo           = get_mesh(m);
                const MeshConfig& cfg = mo.mesh.cfg;
                if (!cfg.all_periodic() || cfg.ghost_width() < 3)
                {
                    throw std::invalid_argument("make_convection_weno5 needs a fully periodic mesh with max_stencil_size(6)");
                }
                if (n_comp < 1 || n_comp > 3)
                {
                    throw std::invalid_argument("n_comp must be 1, 2 or 3");
                }
                FluxPlan fp;
                build_fluxw_plan(mo.mesh, fp);
                std::vector<double> tab;
                wide_tab(cfg, tab);
                const smr_item_fluxw* it = reinterpret_cast<const smr_item_fluxw*>(fp.arena.p + fp.items.items);
                const int64_t* aux       = reinterpret_cast<const int64_t*>(fp.arena.p + fp.items.aux);
                const double* up[3]      = {nullptr, nullptr, nullptr};
                double* op[3]            = {nullptr, nullptr, nullptr};
                for (int c = 0; c < n_comp; ++c)
                {
                    up[c] = u + static_cast<int64_t>(c) * mo.mesh.nref;
                    op[c] = out + static_cast<int64_t>(c) * mo.mesh.nref;
                }
                std::fill(out, out + static_cast<int64_t>(n_comp) * mo.mesh.nref, 0.0);
                with_weno_op(cfg.dim, kind, n_comp, up, op, aux, tab.data(), velocity, scale,
                             [&](const auto& opr)
                             {
                                 double acc[3];
                                 for (int i = 0; i < fp.items.n_items; ++i)
                                 {
                                     for (int k = 0; k < it[i].n; ++k)
                                     {
                                         opr.compute(it[i], k, acc);
                                         for (int c = 0; c < n_comp; ++c)
                                         {
                                             op[c][it[i].c + k] = acc[c];
                                         }
                                     }
                                 }
                             });
            });
    }

    int smr_adapt_last_size(smr_mesh_t m, int64_t* out)
    {
        return guarded(
            [&]
            {
                *out = get_mesh(m).last_size;
            });
    }

    int smr_adapt_last_tags(smr_mesh_t m, uint8_t* host, int64_t n)
    {
        return guarded(
            [&]
            {
                MeshObj& mo = get_mesh(m);
                if (n != mo.last_size || n == 0)
                {
                    throw std::invalid_argument("size does not match the last adaptation's reference size");
                }
                if (!mo.h_tag_valid)
                {
                    require_device();
                    SMR_CUDA(cudaMemcpyAsync(mo.h_tag.p, mo.d_tag.p, static_cast<size_t>(n), cudaMemcpyDeviceToHost, g.stream));
                    SMR_CUDA(cudaStreamSynchronize(g.stream));
                    mo.h_tag_valid = true;
                }
                std::memcpy(host, mo.h_tag.p, static_cast<size_t>(n));
            });
    }

    int smr_adapt_last_detail(smr_mesh_t m, double* host, int64_t n)
    {
        return guarded(
            [&]
            {
                require_device();
                MeshObj& mo = get_mesh(m);
                if (n != mo.last_size * mo.last_ncomp || n == 0)
                {
                    throw std::invalid_argument("size does not match the last adaptation's reference size x n_fields");
                }
                SMR_CUDA(cudaMemcpyAsync(host, mo.d_detail.p, static_cast<size_t>(n) * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
                SMR_CUDA(cudaStreamSynchronize(g.stream));
            });
    }

    int smr_stats_get(smr_stats* out)
    {
        return guarded(
            [&]
            {
                resolve_sections();
                *out = g.stats;
            });
    }

    int smr_stats_reset(void)
    {
        return guarded(
            [&]
            {
                resolve_sections();
                g.stats = smr_stats{};
            });
    }

    int smr_profile_enable(int on)
    {
        g.profile = on != 0;
        for (int f = 0; f < SMR_FAM_COUNT; ++f)
        {
            g.prof_seconds[f]  = 0;
            g.prof_launches[f] = 0;
            g.prof_cells[f]    = 0;
            g.prof_bytes[f]    = 0;
        }
        return SMR_OK;
    }

    int smr_profile_get_bytes(int family, uint64_t* bytes)
    {
        return guarded(
            [&]
            {
                if (family < 0 || family >= SMR_FAM_COUNT)
                {
                    throw std::invalid_argument("invalid kernel family");
                }
                *bytes = g.prof_bytes[family];
            });
    }

    int smr_set_fused(int on)
    {
        g.fuse = on != 0;
        return SMR_OK;
    }

    int smr_profile_get(int family, uint64_t* launches, double* seconds, uint64_t* cells)
    {
        return guarded(
            [&]
            {
                if (family < 0 || family >= SMR_FAM_COUNT)
                {
                    throw std::invalid_argument("invalid kernel family");
                }
                *launches = g.prof_launches[family];
                *seconds  = g.prof_seconds[family];
                *cells    = g.prof_cells[family];
            });
    }

    int smr_field_init_ball(smr_field_t fh, const double* center, double radius, double inside, double outside, int overwrite_outside)
    {
        return guarded(
            [&]
            {
                require_device();
                FieldObj& f = get_field(fh);
                check_field_ready(f);
                MeshObj& mo = *f.mesh;
                ensure_plan(mo);
                Section sec;
                f.ghosts_valid        = false;
                const MeshConfig& cfg = mo.mesh.cfg;
                double* u             = static_cast<double*>(f.data.p);
                auto run              = [&](auto op)
                {
                    op.u = u;
                    for (int d = 0; d < 3; ++d)
                    {
                        op.origin[d] = cfg.origin[d];
                        op.center[d] = d < cfg.dim ? center[d] : 0.0;
                    }
                    op.scaling           = cfg.scaling;
                    op.radius            = radius;
                    op.inside            = inside;
                    op.outside           = outside;
                    op.overwrite_outside = overwrite_outside != 0;
                    launch<smr_item_fv>(SMR_FAM_INIT, mo.d_arena.p, mo.plan.fv, op);
                };
                switch (cfg.dim)
                {
                    case 1:
                        run(InitBallOp<1>{});
                        break;
                    case 2:
                        run(InitBallOp<2>{});
                        break;
                    default:
                        run(InitBallOp<3>{});
                        break;
                }
            });
    }
}
