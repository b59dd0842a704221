/* Work-item records shared by the host batch builders and the CUDA kernels.
 *
 * A batch is the host-side traversal of one set-algebra subset of the reference, flattened into
 * records that carry precomputed storage offsets, so the kernels never search the mesh.  One record = one x-interval
 * of the subset; `n` = number of output cells it produces.  Batches also carry an exclusive prefix sum of `n` and,
 * per CTA of SMR_CTA_CELLS output cells, the index of the first record that CTA touches.
 */
#ifndef SAMURAI_B200_ITEMS_H
#define SAMURAI_B200_ITEMS_H
#include <stdint.h>

#include "../../include/samurai_b200.h"

#define SMR_CTA_THREADS 256
#define SMR_CELLS_PER_THREAD 4
#define SMR_CTA_CELLS (SMR_CTA_THREADS * SMR_CELLS_PER_THREAD)
#define SMR_MAX_LEVELS 24
#define SMR_MAX_RANKS 8

/* kernel families, for the per-family profile (smr_profile_get) */
enum
{
    SMR_FAM_FV = 0,
    SMR_FAM_PROJ,
    SMR_FAM_PRED,
    SMR_FAM_DETAIL,
    SMR_FAM_CRITERIA,
    SMR_FAM_MAXIMUM,
    SMR_FAM_BC,
    SMR_FAM_COPY,
    SMR_FAM_KEEP,
    SMR_FAM_INIT,
    SMR_FAM_WAVEFRONT, /* fused level wavefront (ghost update / harten iteration / field transfer in one launch) */
    SMR_FAM_COUNT
};

/* leaf interval for the FV field-expression kernels (stencil_field.hpp) */
typedef struct
{
    int64_t c;      /* offset of first cell in its own row (i-1 / i+1 are c-1 / c+1) */
    int64_t ym, yp; /* same x, rows j-1 / j+1 */
    int64_t zm, zp; /* same x, rows k-1 / k+1 (3D) */
    int32_t n;
    int32_t level;
    int32_t x, y, z; /* integer coordinates of the first cell (cell.hpp:32-77) */
    int32_t mask;    /* multi-GPU: peers that also receive this record's outputs (bit p = rank p) */
} smr_item_fv;

/* SMR_STRIP_ROWS consecutive leaf rows (same z, y0 .. y0+R-1) sharing the x-range [start, start+n): one thread walks
 * one column of the strip, so every value of the strip is fetched from L2 once instead of three times */
#define SMR_STRIP_ROWS 4
#ifndef STRIP_UPT
#define STRIP_UPT 4 /* strip columns per thread (measured best on B200 with 6 CTAs/SM) */
#endif
typedef struct
{
    int64_t row[SMR_STRIP_ROWS + 2]; /* offset of x = start in rows y0-1, y0, ..., y0+R */
    int64_t zm[SMR_STRIP_ROWS];      /* 3D: same x, plane k-1, rows y0 .. y0+R-1 */
    int64_t zp[SMR_STRIP_ROWS];
    int32_t n; /* columns */
    int32_t level;
    int32_t mask;
    int32_t x, y, z; /* coordinates of the first cell of the first row */
} smr_item_fvstrip;

/* leaf sub-interval for the flux-based schemes on multi-level meshes (schemes/fv/flux_based/, interface.hpp): all cells
 * of the record see the same kind of neighbour across each transverse face; the x faces only matter for the first
 * (x-) and last (x+) cell, interior x faces are same-level by construction */
enum
{
    SMR_FACE_SAME   = 0, /* same-level leaf                      interface.hpp:35-110  */
    SMR_FACE_COARSE = 1, /* coarser leaf: this cell is the fine side of a jump, the stencil uses the ghost at its own level */
    SMR_FACE_FINE   = 2, /* finer leaves: this cell is the coarse side, contributions come from level+1 rows (aux offsets) */
    SMR_FACE_BDRY   = 3  /* domain boundary                      boundary.hpp:6-33     */
};
#define SMR_FLUX_AUX_SLOTS 24 /* 6 faces x 4 level+1 row offsets */
typedef struct
{
    int64_t c;     /* offset of the first cell in its own row */
    int64_t nb[4]; /* same x in rows y-1, y+1, z-1, z+1 of the cell's own level */
    int64_t fine;  /* first of this record's SMR_FLUX_AUX_SLOTS entries in the batch's aux array (only if a face is FINE):
                      x faces: [face*4 + cy + 2*cz] = offset of stencil cell 0 in child row (2y+cy, 2z+cz);
                      y/z faces: [face*4 + 2*b + st] = offset of x = 2*start in stencil row st of the b-th child row */
    int32_t n;
    int32_t level;
    int32_t kinds; /* 2 bits per face, face = 2*d + (plus side) */
    int32_t mask;
} smr_item_flux;

/* leaf sub-interval for flux schemes with a six-cell line stencil {-2 .. 3} (make_convection_weno5, operators/convection_lin.hpp:95-178)
 * on fully periodic meshes: like smr_item_flux, with the three own-level rows on each side of every transverse direction.  A face on a
 * periodic boundary is classified by the leaf found at the wrapped position (interface.hpp:83-92, 179-189, 280-290) and reads the
 * periodic ghosts at the unwrapped one. */
#define SMR_FLUXW_AUX_SLOTS 56 /* x faces: [face*4 + cy + 2*cz] = offset of the stencil origin (cell left of the interface) in child
                                  row (2y+cy, 2z+cz); transverse face f = 2..5: [8 + ((f-2)*2 + b)*6 + st] = offset of x = 2*start
                                  in stencil row st (origin row - 2 + st) of the b-th child row of the other transverse direction */
#define SMR_FLUX_PLUS_THROUGH_SHIFT 16 /* kinds bit (16 + d): the plus face of direction d is a same-level interface through the periodic
                                         boundary (linear schemes: its x-interface is not in the cell's own interface interval) */
#define SMR_FLUXW_SWAP_SHIFT 12 /* kinds bit (12 + d): the minus face of direction d is a same-level interface through the periodic
                                   boundary: the reference visits it after the regular same-level interfaces, i.e. after the plus face */
typedef struct
{
    int64_t c;      /* offset of the first cell in its own row */
    int64_t nb[12]; /* same x in rows y-3, y-2, y-1, y+1, y+2, y+3 [0..5] and z-3 .. z+3 [6..11] of the cell's own level */
    int64_t fine;   /* first of this record's SMR_FLUXW_AUX_SLOTS aux entries (only if a face is FINE) */
    int32_t n;
    int32_t level;
    int32_t kinds;  /* 2 bits per face as in smr_item_flux + the swap bits */
    int32_t mask;
} smr_item_fluxw;

/* coarse interval filled by projection (numeric/projection.hpp:22-64) */
typedef struct
{
    int64_t dst;    /* coarse offset */
    int64_t src[4]; /* fine offset of x = 2*start in rows (2y+cy, 2z+cz), index cy + 2*cz */
    int32_t n;
    int32_t mask;
} smr_item_proj;

/* fine interval filled by prediction (numeric/prediction.hpp:259-361, 149-257) */
typedef struct
{
    int64_t dst;    /* fine offset of x = start */
    int64_t src[9]; /* coarse offset of x = start>>1 in rows (y>>1 + ry - 1, z>>1 + rz - 1), index ry + 3*rz */
    int32_t n;
    int32_t par; /* bit0: start & 1, bit1: y & 1, bit2: z & 1; bits 8-15: multi-GPU peer mask */
} smr_item_pred;

/* coarse interval whose 2^dim children get a detail (mr/operators.hpp:139-533) */
typedef struct
{
    int64_t coarse[9]; /* offset of x = start in rows (y + ry - 1, z + rz - 1), index ry + 3*rz */
    int64_t fine[4];   /* offset of x = 2*start in rows (2y+cy, 2z+cz), index cy + 2*cz */
    int32_t n;
    int32_t mask;
} smr_item_detail;

/* coarse interval for the tagging criteria and the keep propagation (mr/criteria.hpp, mr/operators.hpp:29-89) */
typedef struct
{
    int64_t coarse;
    int64_t fine[4];
    int32_t n;
    int32_t level; /* bits 0-7: level of the children; bits 8-15: multi-GPU peer mask */
} smr_item_tag;

typedef struct
{
    int64_t dst;
    int64_t src;
    int32_t n;
    int32_t mask;
} smr_item_copy;

/* one boundary ghost cell (algorithm/update_outer_ghost.hpp, bc/apply_field_bc.hpp) */
enum
{
    SMR_BC_COPY  = 0, /* f[dst] = f[src[0]]                            corner extrapolation / projection, predict_bc */
    SMR_BC_VALUE = 1, /* Dirichlet: 2*v - f[src[0]] ; Neumann: coef*v + f[src[0]]   (coef = dx)                     */
    SMR_BC_AVG   = 2, /* f[dst] = (0 + f[src[0]] + ... ) / n_src ; 0 if n_src == 0   project_bc                     */
    SMR_BC_EXTRAP4 = 3 /* f[dst] = f[src[0]] - f[src[1]] * 3 + f[src[2]] * 3   PolynomialExtrapolation<4>, second ghost layer */
};

typedef struct
{
    int64_t dst;
    double coef;
    int32_t kind; /* bits 0-7: SMR_BC_*; bits 8-15: multi-GPU peer mask */
    int32_t n_src;
    int64_t src_first; /* index into the batch's int64 source-offset array */
} smr_item_bc;

/* ------------------------------------------------------------------------------------------------------------------
 * Device-side record derivation.  The host no longer looks up storage offsets: it uploads the reference sub-mesh as a
 * per-level CSR (row keys, row pointers, interval starts / ends / storage offsets) plus one 24-byte *seed* per record
 * (the x-interval of the subset and its row), and derive_kernel (kernels.cuh) turns every seed into the record above
 * by searching the CSR -- the row lookups of the reference's `mesh[mesh_id_t::reference][level][interval, index]`
 * accessors (cell_array.hpp, level_cell_array.hpp:find).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct
{
    int32_t xs; /* first x of the interval (output resolution of the record) */
    int32_t n;  /* output cells */
    int32_t y, z;
    int32_t level; /* bits 0-7: level the record is written for; bits 8-15: multi-GPU peer mask */
    int32_t pad;
} smr_seed;

/* one level of the reference sub-mesh on the device: byte offsets into the CSR buffer */
typedef struct
{
    int64_t key;  /* int64 [rows]    (y, z) packed like intervals.hpp: mk_key */
    int64_t ptr;  /* int32 [rows+1]  */
    int64_t xs;   /* int32 [nivl]    */
    int64_t xe;   /* int32 [nivl]    */
    int64_t off;  /* int64 [nivl]    storage offset of cell xs */
    int32_t rows; /* 0: level absent */
    int32_t pad;
} smr_csr_level;

typedef struct
{
    smr_csr_level lv[SMR_MAX_LEVELS];
} smr_csr_table;

enum
{
    SMR_DERIVE_FV = 0,
    SMR_DERIVE_FVSTRIP,
    SMR_DERIVE_PROJ,   /* seed at the coarse level `level`: dst in csr_dst[level], children in csr_src[level+1] */
    SMR_DERIVE_PRED,   /* seed at the fine level: dst in csr_dst[level], parents in csr_src[level-1] */
    SMR_DERIVE_DETAIL, /* seed at the coarse level */
    SMR_DERIVE_TAG,    /* seed at the coarse level `level`, record level = level + 1 */
    SMR_DERIVE_COPY    /* dst in csr_dst[level], src in csr_src[level] */
};

typedef struct
{
    int32_t kind;
    int32_t n;           /* seeds */
    int32_t first_block; /* first CTA of the derive launch that works on this job */
    int32_t pad;
    int64_t seeds; /* byte offset into the arena */
    int64_t items; /* byte offset into the arena (device-only region) */
} smr_derive_job;

#endif
