/* samurai_b200 -- C ABI of the B200-native implementation of samurai's per-time-step hot path.
 *
 * samurai (hpc-maths/samurai v0.33.0) is a header-only C++ library with no plugin/FFI layer; its "boundary" is its
 * public C++ API.  This header is what a `<samurai/...>` header set binds to: every entry point
 * names the reference interface it replaces (paths relative to the reference's include/samurai/).
 *
 * Conventions: plain C types, opaque handles, every call returns 0 on success and non-zero on failure with the
 * message available from smr_last_error().  Errors the reference raises as C++ exceptions (std::out_of_range from
 * LevelCellArray::get_interval, level_cell_array.hpp:671-724; std::invalid_argument from attach_bc,
 * field/field_base.hpp:306-315) are reported with the codes below so the C++ side can re-throw the same type.
 * One host thread per process drives one GPU; work is asynchronous on one CUDA stream and the calls that return
 * data to the host synchronise it.  There is NO CPU fallback: compute entry points fail with SMR_ERR_CUDA when no
 * device is available.  Host-only entry points (mesh construction and queries) work without a GPU.
 */
#ifndef SAMURAI_B200_H
#define SAMURAI_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

    typedef uint64_t smr_mesh_t;
    typedef uint64_t smr_field_t;

    enum
    {
        SMR_OK               = 0,
        SMR_ERR_INVALID      = 1, /* std::invalid_argument */
        SMR_ERR_OUT_OF_RANGE = 2, /* std::out_of_range     */
        SMR_ERR_CUDA         = 3, /* CUDA runtime failure / no device */
        SMR_ERR_INTERNAL     = 4
    };

    /* mesh_id_t of MRMesh (mr/mesh.hpp:25-34) */
    enum
    {
        SMR_MESH_CELLS            = 0,
        SMR_MESH_CELLS_AND_GHOSTS = 1,
        SMR_MESH_PROJ_CELLS       = 2,
        SMR_MESH_UNION_CELLS      = 3,
        SMR_MESH_REFERENCE        = 4
    };

    /* mesh_config<dim, prediction_stencil_radius> (mesh_config.hpp:20-432) after parse_args() */
    typedef struct
    {
        int32_t dim;                /* 1, 2 or 3                                                   */
        int32_t min_level;          /* mesh_config::min_level                                      */
        int32_t max_level;          /* mesh_config::max_level                                      */
        int32_t pred_radius;        /* template parameter prediction_stencil_radius: 0 or 1        */
        int32_t max_stencil_radius; /* mesh_config::max_stencil_radius: 1 or 2; up to 3 on fully periodic meshes */
        int32_t graduation_width;   /* mesh_config::graduation_width                               */
        int32_t n_cells0[3];        /* box size in level-0 cells: length / scaling_factor          */
        double origin[3];           /* Box::min_corner                                             */
        double scaling_factor;      /* LevelCellArray::scaling_factor (box.hpp:280-, approximate_box) */
        int32_t periodic[3];        /* mesh_config::periodic(d) (mesh_config.hpp:171-196)           */
        int32_t refine_boundary;    /* args::refine_boundary (`--refine-boundary`): keep_boundary_refined after the criteria,
                                       mr/adapt.hpp:245-274, 340-345 */
    } smr_mesh_config;

    /* one x-interval of a sub-mesh: LevelCellArray entry (interval.hpp:50-64) flattened with its (y, z) */
    typedef struct
    {
        int32_t y, z;
        int32_t start, end; /* [start, end) */
        int64_t offset;     /* storage offset of cell `start` (= Interval::index + start)          */
    } smr_interval;

    /* ---- runtime (samurai.hpp:22-114 initialize/finalize) ------------------------------------------------------ */
    int smr_init(int device);  /* device < 0: host-only mode (no CUDA context is created)            */
    int smr_finalize(void);
    const char* smr_last_error(void);
    int smr_device_available(void);
    int smr_set_stream(void* cuda_stream); /* run on an existing cudaStream_t (e.g. torch's current stream) */
    int smr_synchronize(void);

    /* ---- mesh: samurai::mra::make_mesh(box, cfg) / MRMesh(ca, cfg) (mr/mesh.hpp:481-529, mesh.hpp:326-412) ------ */
    int smr_mesh_create_uniform(const smr_mesh_config* cfg, int level, smr_mesh_t* out); /* start_level = level */
    int smr_mesh_create_from_intervals(const smr_mesh_config* cfg, const int32_t* levels, const smr_interval* ivl, int64_t n, smr_mesh_t* out);
    int smr_mesh_destroy(smr_mesh_t m);
    int smr_mesh_config_get(smr_mesh_t m, smr_mesh_config* out);
    /* mesh.nb_cells(mesh_id) / nb_cells(level, mesh_id) (mesh.hpp:541-558); level < 0: all levels */
    int smr_mesh_nb_cells(smr_mesh_t m, int mesh_id, int level, int64_t* out);
    int smr_mesh_nb_intervals(smr_mesh_t m, int mesh_id, int level, int64_t* out);
    /* for_each_interval(mesh[mesh_id][level]) order (algorithm.hpp:75-132); `out` has room for nb_intervals */
    int smr_mesh_get_intervals(smr_mesh_t m, int mesh_id, int level, smr_interval* out);
    int smr_mesh_generation(smr_mesh_t m, uint64_t* out); /* bumped by every adaptation that changes the mesh */
    /* mesh.get_index(level, i, j, k) (mesh.hpp:772-789): -1 in *out and SMR_ERR_OUT_OF_RANGE when absent */
    int smr_mesh_get_index(smr_mesh_t m, int level, int i, int j, int k, int64_t* out);

    /* host half of the adaptation on its own: update_cell_array_from_tag + make_graduation + MRMesh(new_ca, mesh)
     * (algorithm/graduation.hpp:743-842, 573-726; mr/adapt.hpp:363-380).  `tags` is reference-sized (CellFlag bits,
     * cell_flag.hpp:11-17).  Fields on the mesh are NOT transferred: resize them.  Works without a GPU. */
    int smr_mesh_update_from_tags(smr_mesh_t m, const uint8_t* tags, int64_t n, int* unchanged);

    /* ---- fields: make_scalar_field<double>(name, mesh) (field/scalar_field.hpp:171-215) ------------------------- */
    int smr_field_create(smr_mesh_t m, const char* name, smr_field_t* out);
    int smr_field_destroy(smr_field_t f);
    int smr_field_resize(smr_field_t f);                 /* Field::resize() (field/access_base.hpp:105-115) */
    int smr_field_fill(smr_field_t f, double v);         /* Field::fill */
    int smr_field_size(smr_field_t f, int64_t* out);     /* nb_cells(reference) */
    int smr_field_upload(smr_field_t f, const double* host, int64_t n);   /* host -> device, n == size */
    int smr_field_download(smr_field_t f, double* host, int64_t n);       /* device -> host, synchronises */
    int smr_field_swap(smr_field_t a, smr_field_t b);    /* std::swap(u.array(), v.array()) */
    /* make_bc<Dirichlet<1>>(u, v) / make_bc<Neumann<1>>(u, v), constant value, Everywhere (bc/bc.hpp:751-815) */
    enum
    {
        SMR_BCTYPE_DIRICHLET = 0, /* bc/dirichlet.hpp:19-30 */
        SMR_BCTYPE_NEUMANN   = 1  /* bc/neumann.hpp:19-35   */
    };

    int smr_field_set_bc(smr_field_t f, int bc_type, double value);

    /* ---- hot path ------------------------------------------------------------------------------------------------ */
    /* update_ghost_mr(field) (algorithm/update_ghost_mr.hpp:260-270 -> :194-237) */
    int smr_update_ghost_mr(smr_field_t f);
    /* unp1 = u - dt * upwind(a, u)  (stencil_field.hpp:83-179 through field_base.hpp:230-242) */
    int smr_fv_upwind(smr_field_t unp1, smr_field_t u, const double* a, double dt);
    /* unp1 = u - dt * upwind_scalar_burgers(k, u)  (stencil_field.hpp:184-249) */
    int smr_fv_upwind_burgers(smr_field_t unp1, smr_field_t u, const double* k, double dt);

    /* flux-based schemes, explicit application: out = S(u) with out.fill(0) first (schemes/fv/FV_scheme.hpp:202-238;
     * linear homogeneous: flux_based/explicit_flux_based_scheme__lin_hom.hpp:39-119,233-319 + flux_based_scheme__lin_hom.hpp:74-229;
     * non-linear: explicit_flux_based_scheme__nonlin.hpp:35-85 + flux_based_scheme__nonlin.hpp:334-520; interfaces incl.
     * level jumps and boundaries: interface.hpp:35-306,440-509).  Ghosts of `u` are updated first if needed
     * (update_ghost_mr_if_needed).  `scale` is the scalar of `scale * scheme` (flux_based/algebraic_operators.hpp:7-82), 1 for none.
     * Every cell accumulates its contributions in the order of the reference's sequential scatter loops. */
    enum
    {
        SMR_SCHEME_CONVECTION_UPWIND = 0, /* make_convection_upwind<Field>(velocity), operators/convection_lin.hpp:15-89; params = velocity[dim] */
        SMR_SCHEME_DIFFUSION_ORDER2  = 1, /* make_diffusion_order2<Field>(K),         operators/diffusion.hpp:123-175;     params = K[dim]        */
        SMR_SCHEME_CONVECTION_UPWIND_NONLINEAR = 2, /* make_convection_upwind<Field>() on a scalar field (Burgers, flux u*u upwinded by the
                                                       mean velocity), operators/convection_nonlin.hpp:24-76; params unused */
        SMR_SCHEME_CONVECTION_WENO5 = 3, /* make_convection_weno5<Field>(velocity) on a scalar field: non-linear flux scheme with the line
                                           stencil {-2 .. 3}, operators/convection_lin.hpp:95-178 + weno_impl.hpp:26-63; params = velocity[dim];
                                           fully periodic meshes with max_stencil_size(6) (interfaces through the periodic boundary:
                                           interface.hpp:83-92, 179-189, 280-290) */
        SMR_SCHEME_CONVECTION_WENO5_NONLINEAR = 4 /* make_convection_weno5<Field>(): WENO5 of f(u) = u * u (scalar field) or u(d) * u
                                                     (vector field with n_comp == dim, through smr_scheme_apply_vector), upwinded by the mean
                                                     of the two cells next to the interface, operators/convection_nonlin.hpp:162-233;
                                                     params unused; same mesh requirements as SMR_SCHEME_CONVECTION_WENO5 */
    };

    int smr_scheme_apply(smr_field_t out, smr_field_t u, int kind, const double* params, double scale);
    /* the same for a vector field stored as n_comp SoA component fields on one mesh: out[c] = S(u)[c].  Vector schemes:
     * SMR_SCHEME_CONVECTION_UPWIND_NONLINEAR = make_convection_upwind<VectorField>() with n_comp == dim (flux u(d) * u upwinded by the
     * mean of component d, operators/convection_nonlin.hpp:24-76) and SMR_SCHEME_CONVECTION_WENO5_NONLINEAR.  The linear schemes act
     * per component: call smr_scheme_apply. */
    int smr_scheme_apply_vector(const smr_field_t* out, const smr_field_t* u, int n_comp, int kind, const double* params, double scale);
    /* out = a * x + b * y over the leaves: the field-expression tail `unp1 = u - dt * scheme(u)` is (1, u, -dt, rhs) */
    int smr_field_lincomb(smr_field_t out, double a, smr_field_t x, double b, smr_field_t y);

    /* make_MRAdapt(fields...)(mra_config) (mr/adapt.hpp:148-195, 277-389): adapts the mesh IN PLACE and transfers
     * `fields`; any other field living on the mesh must be smr_field_resize()d by the caller, as in the reference.
     * *n_iterations receives the number of harten iterations executed. */
    int smr_adapt(const smr_field_t* fields, int n_fields, double epsilon, double regularity, int* n_iterations);
    /* same with mra_config::relative_detail (mr/config.hpp, mr/rel_detail.hpp:73-112): details divided by max_leaves |f| */
    int smr_adapt_ex(const smr_field_t* fields, int n_fields, double epsilon, double regularity, int relative_detail, int* n_iterations);
    /* one harten iteration (mr/adapt.hpp:277-389); *unchanged = 1 when the mesh was already at its fixed point */
    int smr_adapt_iteration(const smr_field_t* fields, int n_fields, double epsilon, double regularity, int ite, int* unchanged);
    /* detail and tag arrays of the last harten iteration, reference-sized, as they were BEFORE the mesh update;
     * n must equal the reference size of the mesh that iteration ran on (smr_adapt_last_size) */
    int smr_adapt_last_size(smr_mesh_t m, int64_t* out);
    int smr_adapt_last_tags(smr_mesh_t m, uint8_t* host, int64_t n);
    int smr_adapt_last_detail(smr_mesh_t m, double* host, int64_t n);

    /* ---- multi-GPU: one process per GPU (reference: Boost.MPI domain decomposition, mesh.hpp:1032-1408 and the
     * exchange_subdomains_merged rounds of algorithm/update_ghost_mr.hpp:125-187) ---------------------------------------
     * Every rank holds the same global mesh; the domain is cut into `world` leaf-balanced slabs along the last axis
     * and each rank computes the records whose output row lies in its slab.  Halo values are stored straight into the
     * peers' field copies by the producing kernels (CUDA IPC mappings over NVLink) and a flag barrier separates the
     * phases; tags are replicated so all ranks derive the same new mesh with no host communication.
     * Call order: smr_init -> smr_mg_init -> exchange the 64-byte handles out of band -> smr_mg_connect. */
    int smr_mg_init(int rank, int world, uint64_t pool_bytes); /* pool holds every field/detail/tag buffer */
    int smr_mg_get_handle(void* out64);                        /* cudaIpcMemHandle_t of this rank's pool */
    int smr_mg_connect(const void* handles);                   /* world x 64 bytes, indexed by rank */
    int smr_mg_broadcast(smr_field_t f);                       /* make every rank's copy of f complete (before a download) */
    /* re-cut the slabs from the current leaves (load_balancing/strategies/sfc.hpp role; uniform weights) */
    int smr_mg_rebalance(const smr_field_t* fields, int n_fields);
    int smr_mg_leaf_owners(smr_mesh_t m, int32_t* out, int64_t n); /* owner rank of every leaf, for_each_cell order */

    /* ---- instrumentation (timers.hpp: "mesh adaptation", "ghost update", ...) ------------------------------------ */
    typedef struct
    {
        uint64_t kernel_launches; /* kernels of this library launched since the last reset                          */
        uint64_t h2d_bytes;
        uint64_t d2h_bytes;
        double host_mesh_seconds;  /* update_cell_array_from_tag + make_graduation + sub-mesh construction          */
        double host_batch_seconds; /* set-algebra traversal into index batches                                      */
        uint64_t mesh_rebuilds;
        double device_seconds;     /* CUDA-event time of every stretch of device work (kernels + async copies)      */
        uint64_t ghost_updates_skipped; /* smr_update_ghost_mr calls answered by the ghosts_updated flag            */
        uint64_t harten_iterations;     /* iterations of the adaptation loop (mr/adapt.hpp:277-389)                  */
        /* host stages: 0 tags -> new leaves, 1 mesh equality tests, 2 make_graduation, 3 sub-mesh construction,
         * 4 field-transfer batches, 5 per-mesh batches, 6 waiting for the device before a host stage, 7 releasing the old mesh */
        double host_stage_seconds[8];
    } smr_stats;

    int smr_stats_get(smr_stats* out); /* synchronises the stream to resolve device_seconds */
    int smr_stats_reset(void);

    /* per-kernel-family profile: when enabled every launch is bracketed by CUDA events and synchronised (slow; never
     * enable it inside a timed region).  family = SMR_FAM_* of csrc/items.h: 0 fv, 1 projection, 2 prediction,
     * 3 detail, 4 criteria, 5 maximum, 6 bc, 7 copy, 8 keep, 9 init, 10 fused wavefront. */
    int smr_profile_enable(int on);
    int smr_profile_get(int family, uint64_t* launches, double* seconds, uint64_t* cells);
    /* algorithmic bytes of the fused launches of a family; 0 for the families whose bytes follow
     * from `cells` alone.  family 10 = fused level wavefront. */
    int smr_profile_get_bytes(int family, uint64_t* bytes);
    /* 1 (default): ghost update, harten iteration and field transfer each run as ONE cooperative launch that walks the
     * level wavefront with grid-wide barriers; 0: one launch per sweep (same results bit for bit; used to profile the
     * kernel families separately).  Multi-GPU runs use the fused launches too: the phase barrier then also exchanges flags with the peers. */
    int smr_set_fused(int on);

    /* host-side profiling aid: rebuild the sub-meshes and the index batches of the current leaves `reps` times
     * (no device work) and report the mean seconds of each stage */
    int smr_debug_host_rebuild(smr_mesh_t m, int reps, double* mesh_seconds, double* plan_seconds, int64_t* arena_bytes);
    /* host-side testing aid (no device needed): the records of the flux-scheme batch of the current mesh, 6 ints each:
     * level, x, y, z of the first cell, length, face kinds (2 bits per face, face = 2*d + plus side; 0 same-level leaf,
     * 1 coarser leaf, 2 finer leaves, 3 domain boundary; the x-face kinds apply to the first / last cell only).
     * out == NULL: only count. */
    int smr_debug_flux_records(smr_mesh_t m, int32_t* out, int64_t capacity, int64_t* n_records);
    /* host-side testing aid (no device needed, tests only): evaluates the records of the six-cell-stencil flux batch (kinds
     * SMR_SCHEME_CONVECTION_WENO5 / _WENO5_NONLINEAR) of the current mesh on the host with the kernel's own per-cell function; `u` and
     * `out` are host arrays of n_comp x nb_cells(reference) doubles (one component after the other), ghosts of `u` already updated.
     * Checks the host-built records and the operation order against the oracle where no GPU is present; it is not a compute path
     * (single-threaded, rebuilds the records at every call). */
    int smr_debug_fluxw_apply(smr_mesh_t m, const double* u, int n_comp, int kind, const double* velocity, double scale, double* out);

    /* the demos' initial condition as a device kernel over the leaves: u = inside where |center(cell) - c|^2 <= r^2,
     * else outside (only written when overwrite_outside != 0)
     * (demos/FiniteVolume/advection_2d.cpp:23-45, advection_3d.cpp:32-45, scalar_burgers_2d.cpp:20-50) */
    int smr_field_init_ball(smr_field_t f, const double* center, double radius, double inside, double outside, int overwrite_outside);

#ifdef __cplusplus
}
#endif
#endif
