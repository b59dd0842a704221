"""CPU oracle: a numpy restatement of samurai's per-time-step hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product (`samurai_b200/`) may import
this module; only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
cpu_baseline / `--impl reference` legs do, and only as the checker / baseline.

It is deliberately a DIFFERENT implementation from the product's host code:
cell sets are sorted arrays of packed integer keys (one key per cell) instead
of interval lists, so a bug in the product's interval algebra cannot hide
behind a shared implementation.  Floating-point work is plain numpy fp64
element-wise arithmetic (no FMA, IEEE round-to-nearest), written in the
reference's operation order so results are bit-comparable.

Parity pin: `tests/test_oracle_golden.py` checks this oracle against four of the
reference's golden HDF5 files (converted by tests/golden/make_golden.py):
  * advection_2d, prediction radius 0 and 1, initial and final meshes + fields
    (field-expression path, MR adaptation, ghost update, graduation);
  * heat.cpp --explicit (test_finite_volume_demo_heat_explicit.h5): linear flux
    schemes across level jumps, Neumann;
  * burgers_mra.cpp (test_finite_volume_demo_mra_burgers_hat.h5): NON-LINEAR flux
    scheme, ghost width 2 (further-ghost extrapolation, contiguous-boundary
    graduation rule), graduation width 2;
  * linear_convection.cpp explicit (test_finite_volume_demo_linear_convection_explicit.h5):
    WENO5 (six-cell stencil, ghost width 3), fully periodic mesh with interfaces
    through the boundary, TVD-RK3;
and against the reference's own periodic test property (tests/test_periodic_oracle.py).
Unpinned (no golden in the snapshot that this path reaches): upwind_scalar_burgers,
the vector and WENO5 forms of the non-linear convection, `--refine-boundary`,
everything 3D: there this file is the restatement only.

Reference (hpc-maths/samurai v0.33.0, paths relative to include/samurai/):
  mesh construction      mesh.hpp:326-341,426-439,894-911,1160-1262  mr/mesh.hpp:222-455
  ghost update           algorithm/update_ghost_mr.hpp:194-237
  outer ghosts / BC      algorithm/update_outer_ghost.hpp:20-432  bc/apply_field_bc.hpp:14-466
  projection/prediction  numeric/projection.hpp:22-64  numeric/prediction.hpp:22-361
  detail / tags          mr/operators.hpp:29-89,139-533  mr/criteria.hpp:19-113  mr/adapt.hpp:148-389
  mesh from tags         algorithm/graduation.hpp:245-330,573-842
  field transfer         algorithm/update_fields.hpp:27-54
  FV operators           stencil_field.hpp:16-243
  flux-based schemes     schemes/fv/flux_based/{flux_based_scheme,explicit_flux_based_scheme}__{lin_hom,nonlin}.hpp
                         interface.hpp:35-306,440-509  schemes/fv/operators/{convection_lin,convection_nonlin,diffusion}.hpp
  relative detail        mr/rel_detail.hpp:73-112
Supported scope: box domains, periodic or not per direction; scalar fp64 fields
and lists of component arrays (vector fields, several fields adapted together);
ghost widths 1 and 2 with Dirichlet<1>/Neumann<1> constant BCs, any width on
fully periodic meshes; prediction radius 0 or 1.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field as dc_field

import math

import numpy as np

# ----------------------------------------------------------------------------
# packed cell keys
# ----------------------------------------------------------------------------
BITS = 21
BIAS = 1 << 20
MASK = (1 << BITS) - 1

KEEP, COARSEN, REFINE = 1, 2, 4  # cell_flag.hpp:11-17

EMPTY = np.zeros(0, dtype=np.int64)


def pack(coords: np.ndarray) -> np.ndarray:
    """coords [N, dim] (x, y[, z]) -> keys sorted by (z, y, x) when sorted."""
    coords = np.asarray(coords, dtype=np.int64)
    key = np.zeros(coords.shape[0], dtype=np.int64)
    for d in range(coords.shape[1]):
        key |= (coords[:, d] + BIAS) << (BITS * d)
    return key


def unpack(keys: np.ndarray, dim: int) -> np.ndarray:
    out = np.empty((keys.shape[0], dim), dtype=np.int64)
    for d in range(dim):
        out[:, d] = ((keys >> (BITS * d)) & MASK) - BIAS
    return out


def shift_key(direction) -> int:
    s = 0
    for d, v in enumerate(direction):
        s += int(v) << (BITS * d)
    return s


def translate(s: np.ndarray, direction) -> np.ndarray:
    return s + shift_key(direction)


def union(*sets):
    sets = [s for s in sets if s.size]
    if not sets:
        return EMPTY
    if len(sets) == 1:
        return sets[0]
    return np.unique(np.concatenate(sets))


def inter(a, b):
    if a.size == 0 or b.size == 0:
        return EMPTY
    return np.intersect1d(a, b, assume_unique=True)


def diff(a, b):
    if a.size == 0 or b.size == 0:
        return a
    return np.setdiff1d(a, b, assume_unique=True)


def coarsen(s: np.ndarray, shift: int, dim: int) -> np.ndarray:
    """`.on(level - shift)`: interval >> shift (interval.hpp:227-233)."""
    if shift == 0 or s.size == 0:
        return s
    return np.unique(pack(unpack(s, dim) >> shift))


def refine(s: np.ndarray, shift: int, dim: int) -> np.ndarray:
    """`.on(level + shift)`: every cell -> its 2^(shift*dim) descendants."""
    if shift == 0 or s.size == 0:
        return s
    c = unpack(s, dim) << shift
    n = 1 << shift
    parts = []
    for off in itertools.product(range(n), repeat=dim):
        parts.append(pack(c + np.array(off, dtype=np.int64)))
    return np.sort(np.concatenate(parts))


def expand(s: np.ndarray, w: int, dim: int) -> np.ndarray:
    """Box expansion by w cells in every dimension (subset/expansion.hpp)."""
    if w == 0 or s.size == 0:
        return s
    parts = [translate(s, off) for off in itertools.product(range(-w, w + 1), repeat=dim)]
    return np.unique(np.concatenate(parts))


def box_cells(lo, hi) -> np.ndarray:
    """All cells with lo[d] <= c[d] < hi[d]."""
    axes = [np.arange(lo[d], hi[d], dtype=np.int64) for d in range(len(lo))]
    grids = np.meshgrid(*axes, indexing="ij")
    coords = np.stack([g.ravel() for g in grids], axis=1)
    return np.sort(pack(coords))


def cartesian_directions(dim):
    """for_each_cartesian_direction order (stencil.hpp:299-335): +e_d, -e_d."""
    out = []
    for d in range(dim):
        for sgn in (1, -1):
            v = [0] * dim
            v[d] = sgn
            out.append(tuple(v))
    return out


def diagonal_directions(dim):
    """for_each_diagonal_direction (stencil.hpp:337-349): static_nested_loop<dim,-1,2>, x fastest."""
    out = []
    for rev in itertools.product((-1, 0, 1), repeat=dim):
        v = tuple(reversed(rev))  # last dim outermost, x innermost
        if sum(abs(c) for c in v) > 1:
            out.append(v)
    return out


# ----------------------------------------------------------------------------
# mesh
# ----------------------------------------------------------------------------
@dataclass
class MeshConfig:
    dim: int = 2
    min_level: int = 4
    max_level: int = 10
    pred_radius: int = 1  # mesh_config<dim, prediction_stencil_radius>
    max_stencil_radius: int = 1
    graduation_width: int = 1
    n_cells0: tuple = None  # domain size in level-0 cells per dim (default 1 each)
    origin: tuple = None
    scaling: float = 1.0
    periodic: tuple = None  # mesh_config::periodic(d) (mesh_config.hpp:171-196)

    def __post_init__(self):
        if self.n_cells0 is None:
            self.n_cells0 = (1,) * self.dim
        if self.origin is None:
            self.origin = (0.0,) * self.dim
        if self.periodic is None:
            self.periodic = (False,) * self.dim
        self.periodic = tuple(bool(p) for p in self.periodic)
        # ghost width 2: further-ghost extrapolation (bc/apply_field_bc.hpp:499-563), the two-layer corner block (:313-466) and the
        # contiguous-boundary graduation rule (graduation.hpp:372-500) are restated for boundary conditions that fill ONE layer
        # (Dirichlet<1> / Neumann<1>); larger widths are not
        # (fully periodic meshes have no boundary: any width, e.g. 3 for the WENO5 stencil)
        assert self.max_stencil_radius in (1, 2) or all(self.periodic), "oracle restates ghost widths 1 and 2 at non-periodic boundaries"
        assert self.pred_radius in (0, 1)

    def periodic_directions(self, level):
        """get_periodic_directions (mesh.hpp:45-81, graduation.hpp:37-76): every combination of -N_d, 0, +N_d over the periodic
        dimensions except the null vector, N_d = number of cells of the domain at `level`."""
        choices = [((-(n << level)), 0, (n << level)) if p else (0,) for n, p in zip(self.n_cells0, self.periodic)]
        out = []
        import itertools

        for v in itertools.product(*choices):
            if any(v):
                out.append(list(v))
        return out

    @property
    def ghost_width(self):
        return max(self.max_stencil_radius, self.pred_radius)  # mesh_config.hpp:395

    def cell_length(self, level):
        return self.scaling / (1 << level)  # cell.hpp:16-20


class Mesh:
    """MRMesh restatement: cells / cells_and_ghosts / proj_cells / union / reference."""

    def __init__(self, cfg: MeshConfig, cells: dict):
        self.cfg = cfg
        dim = cfg.dim
        L = cfg.max_level
        nlev = L + 3
        self.nlev = nlev
        self.cells = [EMPTY] * nlev
        for l, s in cells.items():
            self.cells[l] = np.asarray(s, dtype=np.int64)
        self._build()

    # domain pyramid (mesh.hpp:1207-1228): the box at every level 0..max_level, kept implicit
    def in_domain(self, level, keys):
        if level > self.cfg.max_level:
            return np.zeros(keys.size, dtype=bool)
        c = unpack(keys, self.cfg.dim)
        ok = np.ones(keys.size, dtype=bool)
        for d in range(self.cfg.dim):
            ok &= (c[:, d] >= 0) & (c[:, d] < (self.cfg.n_cells0[d] << level))
        return ok

    def domain_inter(self, level, s):
        return s[self.in_domain(level, s)] if s.size else s

    def domain_diff(self, level, s):
        return s[~self.in_domain(level, s)] if s.size else s

    def expanded_domain_inter(self, level, s, w):
        """s ∩ expand(domain[level], w)."""
        if s.size == 0:
            return s
        c = unpack(s, self.cfg.dim)
        ok = np.ones(s.size, dtype=bool)
        for d in range(self.cfg.dim):
            ok &= (c[:, d] >= -w) & (c[:, d] < (self.cfg.n_cells0[d] << level) + w)
        return s[ok]

    def corner_cell(self, level, direction):
        """Mesh_base::construct_corners (mesh.hpp:914-997) `.on(level)`: the domain's corner cell(s) for a
        diagonal direction = domain \\ U_{d: dir[d]!=0} translate(domain, -dir[d] e_d), here for a box."""
        dim = self.cfg.dim
        axes = []
        for d in range(dim):
            n = self.cfg.n_cells0[d] << level
            if direction[d] > 0:
                axes.append(np.array([n - 1], dtype=np.int64))
            elif direction[d] < 0:
                axes.append(np.array([0], dtype=np.int64))
            else:
                axes.append(np.arange(n, dtype=np.int64))
        grids = np.meshgrid(*axes, indexing="ij")
        return np.sort(pack(np.stack([g.ravel() for g in grids], axis=1)))

    @staticmethod
    def uniform(cfg: MeshConfig, level=None):
        level = cfg.max_level if level is None else level
        dim = cfg.dim
        return Mesh(cfg, {level: box_cells([0] * dim, [n << level for n in cfg.n_cells0])})

    def _build(self):
        cfg, dim, nlev = self.cfg, self.cfg.dim, self.nlev
        L = cfg.max_level
        msr, pr = cfg.max_stencil_radius, cfg.pred_radius
        # construct_union (mesh.hpp:1231-1262)
        self.union = [EMPTY] * nlev
        for l in range(L, 0, -1):
            self.union[l - 1] = coarsen(union(self.cells[l], self.union[l]), 1, dim)
        # cells_and_ghosts (mr/mesh.hpp:240-253)
        self.cag = [expand(self.cells[l], msr, dim) for l in range(nlev)]
        cl = [self.cag[l] for l in range(nlev)]
        periodic = any(cfg.periodic)
        if periodic:
            # ghosts of the periodic images (mr/mesh.hpp:276-331): translate the leaves, expand, keep what lies in the expanded domain
            for l in range(nlev):
                if self.cells[l].size and l <= L:
                    for d in cfg.periodic_directions(l):
                        s = expand(translate(self.cells[l], d), msr, dim)
                        cl[l] = union(cl[l], self.expanded_domain_inter(l, s, msr))
        self.proj = [EMPTY] * nlev
        if cfg.max_level != cfg.min_level:
            # prediction ghosts one and two levels below (mr/mesh.hpp:329-359)
            for l in range(nlev):
                if self.cells[l].size == 0 or l == 0:
                    continue
                s1 = self.expanded_domain_inter(l - 1, expand(coarsen(self.cag[l], 1, dim), pr, dim), pr)
                cl[l - 1] = union(cl[l - 1], s1)
                if l - 1 > 0:
                    s2 = expand(coarsen(self.cells[l], 2, dim), pr, dim)
                    cl[l - 2] = union(cl[l - 2], s2)
                if periodic:
                    # periodic part (mr/mesh.hpp:366-380): same sets for the translated leaves, inside the domain only
                    for d in cfg.periodic_directions(l):
                        p1 = expand(coarsen(translate(self.cag[l], d), 1, dim), pr, dim)
                        cl[l - 1] = union(cl[l - 1], self.domain_inter(l - 1, p1))
                        if l - 1 > 0:
                            p2 = expand(coarsen(translate(self.cells[l], d), 2, dim), pr, dim)
                            cl[l - 2] = union(cl[l - 2], self.domain_inter(l - 2, p2))
            ref = list(cl)
            # children of projected ghosts, coarse -> fine cascade (mr/mesh.hpp:415-452);
            # for_each_level re-evaluates max_level() every iteration and skips empty levels
            nonempty = [l for l in range(nlev) if ref[l].size]
            l = nonempty[0] if nonempty else nlev
            while l < nlev - 1 and l <= max(k for k in range(nlev) if ref[k].size):
                if ref[l].size:
                    e = inter(ref[l], self.union[l])
                    if e.size:
                        cl[l + 1] = union(cl[l + 1], refine(e, 1, dim))
                    ref[l + 1] = cl[l + 1]
                    self.proj[l] = e
                l += 1
        self.ref = cl
        # renumbering (mesh.hpp:894-911, cell_array.hpp:484-493): level asc, then (z, y, x)
        self.start = np.zeros(nlev + 1, dtype=np.int64)
        for l in range(nlev):
            self.start[l + 1] = self.start[l] + self.ref[l].size
        self.nref = int(self.start[-1])

    # storage offsets ---------------------------------------------------------
    def index(self, level, keys, strict=True):
        ref = self.ref[level]
        pos = np.searchsorted(ref, keys)
        ok = (pos < ref.size)
        ok[ok] &= ref[pos[ok]] == keys[ok]
        if strict:
            if not ok.all():
                bad = unpack(np.asarray(keys)[~ok][:5], self.cfg.dim)
                raise IndexError(f"cells not in reference mesh at level {level}: {bad.tolist()}")
            return self.start[level] + pos
        return np.where(ok, self.start[level] + pos, -1)

    def contains(self, level, keys):
        ref = self.ref[level]
        pos = np.searchsorted(ref, keys)
        ok = pos < ref.size
        ok[ok] &= ref[pos[ok]] == keys[ok]
        return ok

    def nb_cells(self):
        return int(sum(c.size for c in self.cells))

    def leaf_levels(self):
        return [l for l in range(self.nlev) if self.cells[l].size]

    def same_cells(self, other_cells):
        for l in range(self.nlev):
            o = other_cells[l] if l < len(other_cells) else EMPTY
            if self.cells[l].size != o.size or not np.array_equal(self.cells[l], o):
                return False
        return True

    def leaf_table(self):
        """(level, coords[N,dim], storage index) for every leaf, in for_each_cell order."""
        lv, co, ix = [], [], []
        for l in self.leaf_levels():
            k = self.cells[l]
            lv.append(np.full(k.size, l, dtype=np.int64))
            co.append(unpack(k, self.cfg.dim))
            ix.append(self.index(l, k))
        return np.concatenate(lv), np.concatenate(co), np.concatenate(ix)

    def cell_centers(self, level, keys):
        c = unpack(keys, self.cfg.dim)
        h = self.cfg.cell_length(level)
        return np.array(self.cfg.origin)[None, :] + h * (c + 0.5)  # cell.hpp:131-134


# ----------------------------------------------------------------------------
# boundary conditions
# ----------------------------------------------------------------------------
@dataclass
class Bc:
    kind: str = "dirichlet"  # or "neumann"
    value: float = 0.0


# ----------------------------------------------------------------------------
# numerics
# ----------------------------------------------------------------------------
def interp_coeffs(radius, sign):
    """numeric/prediction.hpp:22-40."""
    if radius == 0:
        return np.array([1.0])
    assert radius == 1
    return np.array([sign / 8.0, 1.0, -sign / 8.0])


def _children_offsets(dim):
    """static_nested_loop<dim-1,0,2> rows x (2i, 2i+1): x fastest, then y, then z."""
    out = []
    for rev in itertools.product((0, 1), repeat=dim):
        out.append(tuple(reversed(rev)))
    return out


def _stencil_offsets(dim, r):
    """static_nested_loop<dim,-r,r+1>: last dim outermost, x innermost."""
    out = []
    for rev in itertools.product(range(-r, r + 1), repeat=dim):
        out.append(tuple(reversed(rev)))
    return out


def projection(mesh: Mesh, f, coarse_level, coarse_keys):
    """projection_op_ (numeric/projection.hpp:22-64): sum rows of (f[2i] + f[2i+1]), * 1/2^dim."""
    dim = mesh.cfg.dim
    if coarse_keys.size == 0:
        return
    c = unpack(coarse_keys, dim)
    total = np.zeros(coarse_keys.size)
    rows = [tuple(reversed(r)) for r in itertools.product((0, 1), repeat=dim - 1)] if dim > 1 else [()]
    for row in rows:
        k0 = pack(2 * c + np.array((0,) + row, dtype=np.int64))
        i0 = mesh.index(coarse_level + 1, k0)
        i1 = mesh.index(coarse_level + 1, k0 + 1)
        total = total + (f[i0] + f[i1])
    inv = 1.0 / float(1 << dim)
    f[mesh.index(coarse_level, coarse_keys)] = total * inv


def predict_values(mesh: Mesh, fsrc, src_mesh: Mesh, fine_level, fine_keys):
    """prediction_op (numeric/prediction.hpp:286-361 and :149-257): value of fine cells from level-1."""
    dim = mesh.cfg.dim
    r = mesh.cfg.pred_radius
    cf = unpack(fine_keys, dim)
    cc = cf >> 1
    if r == 0:
        return fsrc[src_mesh.index(fine_level - 1, pack(cc))].copy()
    par = cf & 1  # 0 even -> interp_coeffs(+1), 1 odd -> interp_coeffs(-1)
    cpair = np.stack([interp_coeffs(r, 1.0), interp_coeffs(r, -1.0)])  # [2, 3]
    val = np.zeros(fine_keys.size)
    for st in _stencil_offsets(dim, r):
        src = fsrc[src_mesh.index(fine_level - 1, pack(cc + np.array(st, dtype=np.int64)))]
        coeff = np.ones(fine_keys.size)
        for d in range(dim):  # coeff *= c_x, then c_y, then c_z
            coeff = coeff * cpair[par[:, d], st[d] + r]
        val = val + src * coeff
    return val


def compute_detail(mesh: Mesh, f, detail, level, coarse_keys):
    """compute_detail_op (mr/operators.hpp:146-175, 226-358, 360-533)."""
    dim = mesh.cfg.dim
    r = mesh.cfg.pred_radius
    if coarse_keys.size == 0:
        return
    c = unpack(coarse_keys, dim)
    ce, co = interp_coeffs(r, 1.0), interp_coeffs(r, -1.0)
    cpair = [ce, co]
    children = _children_offsets(dim)
    idx_child = [mesh.index(level + 1, pack(2 * c + np.array(ch, dtype=np.int64))) for ch in children]
    d = [f[i].copy() for i in idx_child]
    if r == 0:
        src = f[mesh.index(level, coarse_keys)]
        for k in range(len(children)):
            d[k] = d[k] - src
    else:
        for st in _stencil_offsets(dim, r):
            src = f[mesh.index(level, pack(c + np.array(st, dtype=np.int64)))]
            for k, ch in enumerate(children):
                coeff = cpair[ch[0]][st[0] + r]
                for dd in range(1, dim):  # (c_x * c_y) * c_z
                    coeff = coeff * cpair[ch[dd]][st[dd] + r]
                d[k] = d[k] - coeff * src
    for k in range(len(children)):
        detail[idx_child[k]] = d[k]


# ----------------------------------------------------------------------------
# ghost update
# ----------------------------------------------------------------------------
def update_outer_ghosts(mesh: Mesh, f, bc: Bc, level):
    """algorithm/update_outer_ghost.hpp:338-432 for ghost width 1."""
    cfg, dim = mesh.cfg, mesh.cfg.dim
    L, lmin = cfg.max_level, cfg.min_level
    if dim > 1 and lmin <= level <= L:
        for direction in diagonal_directions(dim):
            if any(direction[d] != 0 and cfg.periodic[d] for d in range(dim)):
                continue  # a periodic direction has no corner ghost (update_outer_ghost.hpp:352-366)
            corner = mesh.corner_cell(level, direction)
            # update_outer_corners_by_polynomial_extrapolation, stencil size 2:
            # u[corner + direction] = u[corner]   (bc/polynomial_extrapolation.hpp:63-66)
            cc = inter(mesh.cells[level], corner)
            if cc.size:
                f[mesh.index(level, translate(cc, direction))] = f[mesh.index(level, cc)]
            if cfg.ghost_width == 2 and cc.size:
                _corner_block_width2(mesh, f, level, direction, cc)
            # project_corner_below (update_outer_ghost.hpp:267-336)
            if level > 0:
                for delta_l in (1, 2):
                    proj_level = level - delta_l
                    fine_outer = inter(translate(corner, direction), mesh.ref[level])
                    ghosts = inter(coarsen(fine_outer, delta_l, dim), mesh.ref[proj_level])
                    if ghosts.size:
                        g = unpack(ghosts, dim)
                        add = np.array([((1 << delta_l) - 1) if direction[d] == -1 else 0 for d in range(dim)], dtype=np.int64)
                        child = pack((g << delta_l) + add)
                        ok = mesh.contains(level, child)
                        if ok.any():
                            f[mesh.index(proj_level, ghosts[ok])] = f[mesh.index(level, child[ok])]
                    if proj_level == 0:
                        break
    for direction in cartesian_directions(dim):
        if any(direction[d] != 0 and cfg.periodic[d] for d in range(dim)):
            continue  # update_outer_ghost.hpp:372-375
        if level < L:
            _project_bc(mesh, f, level, direction)
        if level >= lmin:
            _apply_field_bc(mesh, f, bc, level, direction)
        if lmin <= level < L:
            _predict_bc(mesh, f, level + 1, direction)
    # the B.C. fills one ghost layer: the farther ones by polynomial extrapolation (update_outer_ghost.hpp:412-423)
    if level >= lmin and cfg.ghost_width > 1:
        for direction in cartesian_directions(dim):
            if any(direction[d] != 0 and cfg.periodic[d] for d in range(dim)):
                continue
            _further_ghosts(mesh, f, level, direction)


def _extrapolate4(mesh: Mesh, f, level, centres, direction, target_offset=None):
    """PolynomialExtrapolation<4> (bc/polynomial_extrapolation.hpp:67-70) on the line stencil -1, 0, 1, 2 along `direction`
    (stencil.hpp:211-231 rotated by convert_for_direction): u[c + 2 d] = u[c - d] - u[c] * 3 + u[c + d] * 3."""
    if centres.size == 0:
        return
    d = list(direction)
    u0 = f[mesh.index(level, translate(centres, [-x for x in d]))]
    u1 = f[mesh.index(level, centres)]
    u2 = f[mesh.index(level, translate(centres, d))]
    f[mesh.index(level, translate(centres, [2 * x for x in d]))] = u0 - u1 * 3.0 + u2 * 3.0


def _further_ghosts(mesh: Mesh, f, level, direction):
    """update_further_ghosts_by_polynomial_extrapolation (bc/apply_field_bc.hpp:499-563), ghost width 2, B.C. of stencil 2."""
    dim = mesh.cfg.dim
    ref = mesh.ref[level]
    if ref.size == 0:
        return
    d = list(direction)
    has2 = translate(ref, [-2 * x for x in d])  # cells c with c + 2 d in the reference mesh
    has1 = translate(ref, [-x for x in d])
    # 1. beyond the boundary leaves (apply_extrapolation_bc_cells<4>, :125-150)
    _extrapolate4(mesh, f, level, inter(_bdry_leaves(mesh, level, direction), has2), direction)
    # 2. beyond the inner ghosts of the boundary layer that lie under finer cells (apply_extrapolation_bc_ghosts<4>, :220-242)
    inside = ref[mesh.in_domain(level, ref)]
    layer = inside[~mesh.in_domain(level, translate(inside, d))]  # domain \ translate(domain, -direction)
    cand = inter(inter(inter(layer, has2), has1), mesh.union[level])
    _extrapolate4(mesh, f, level, diff(cand, mesh.cells[level]), direction)


def _corner_block_width2(mesh: Mesh, f, level, direction, cc):
    """update_outer_corners_by_polynomial_extrapolation for ghost width 2 (bc/apply_field_bc.hpp:313-466): the second diagonal ghost
    by the 4-point extrapolation along the diagonal, then the off-diagonal ghosts of the corner block copied from the diagonal ones."""
    dim = mesh.cfg.dim
    d = list(direction)
    ref = mesh.ref[level]
    # step 1, layer 2: needs the farthest ghost (apply_extrapolation_bc_cells<4>)
    c2 = inter(cc, translate(ref, [-2 * x for x in d]))
    _extrapolate4(mesh, f, level, c2, direction)
    # step 2: for layer k the diagonal ghost corner + k d is copied to the block cells whose offsets along the other non-zero
    # dimensions are g_p - (k - 1), g_p in {0, 1}, not all equal to k - 1
    nz = [k for k in range(dim) if d[k] != 0]
    if len(nz) < 2:
        return
    gw = 2
    for k in (1, 2):
        src = translate(cc, [k * x for x in d])
        for combo in itertools.product(range(gw), repeat=len(nz) - 1):
            # the reference enumerates g_1 fastest (combo index modulo ghost_width first)
            g = list(reversed(combo))
            delta = [0] * dim
            for p in range(1, len(nz)):
                delta[nz[p]] += (g[p - 1] - (k - 1)) * d[nz[p]]
            if not any(delta):
                continue
            dst = translate(src, delta)
            ok = mesh.contains(level, dst) & mesh.contains(level, src)
            if ok.any():
                f[mesh.index(level, dst[ok])] = f[mesh.index(level, src[ok])]


def _project_bc(mesh: Mesh, f, level, direction):
    """project_bc, layer 1 (update_outer_ghost.hpp:20-152)."""
    dim = mesh.cfg.dim
    outside = mesh.domain_diff(level, translate(mesh.union[level], direction))
    ghosts = inter(outside, mesh.ref[level])
    if ghosts.size == 0:
        return
    gi = mesh.index(level, ghosts)
    total = np.zeros(ghosts.size)
    count = np.zeros(ghosts.size, dtype=np.int64)
    g = unpack(ghosts, dim)
    for dl in (1, 2):
        todo = count == 0
        if not todo.any():
            break
        n = 1 << dl
        # children traversed row by row: (z, y) outer, x inner
        for rev in itertools.product(range(n), repeat=dim):
            off = np.array(tuple(reversed(rev)), dtype=np.int64)
            ck = pack((g[todo] << dl) + off)
            ok = mesh.contains(level + dl, ck)
            if ok.any():
                sel = np.flatnonzero(todo)[ok]
                total[sel] = total[sel] + f[mesh.index(level + dl, ck[ok])]
                count[sel] += 1
    has = count > 0
    out = np.zeros(ghosts.size)  # field = 0 first (update_outer_ghost.hpp:77)
    out[has] = total[has] / count[has]
    f[gi] = out


def _bdry_leaves(mesh: Mesh, level, direction):
    """leaves at `level` whose neighbour in `direction` is outside the domain (boundary.hpp:6-33)."""
    k = mesh.cells[level]
    return k[~mesh.in_domain(level, translate(k, direction))] if k.size else k


def _apply_field_bc(mesh: Mesh, f, bc: Bc, level, direction):
    """apply_field_bc (bc/apply_field_bc.hpp:53-101); DirichletImpl<1>, NeumannImpl<1>."""
    cells = _bdry_leaves(mesh, level, direction)
    if cells.size == 0:
        return
    ci = mesh.index(level, cells)
    gi = mesh.index(level, translate(cells, direction))
    if bc.kind == "dirichlet":
        f[gi] = 2 * bc.value - f[ci]  # bc/dirichlet.hpp:29
    elif bc.kind == "neumann":
        dx = mesh.cfg.cell_length(level)
        f[gi] = dx * bc.value + f[ci]  # bc/neumann.hpp:27-28
    else:
        raise ValueError(bc.kind)


def _predict_bc(mesh: Mesh, f, pred_level, direction):
    """predict_bc (update_outer_ghost.hpp:194-262): copy parent BC ghost into its children."""
    dim = mesh.cfg.dim
    bc_ghosts = mesh.domain_diff(pred_level - 1, translate(_bdry_leaves(mesh, pred_level - 1, direction), direction))
    fine = inter(refine(bc_ghosts, 1, dim), mesh.ref[pred_level])
    if fine.size == 0:
        return
    parent = pack(unpack(fine, dim) >> 1)
    f[mesh.index(pred_level, fine)] = f[mesh.index(pred_level - 1, parent)]


def projection_set(mesh: Mesh, level):
    """cells of proj_cells[level-1] with a child in reference[level] (update_ghost_mr.hpp:219)."""
    dim = mesh.cfg.dim
    return inter(coarsen(mesh.ref[level], 1, dim), mesh.proj[level - 1])


def prediction_set(mesh: Mesh, level):
    """update_ghost_mr.hpp:226-228."""
    dim = mesh.cfg.dim
    pred_ghosts = diff(mesh.ref[level], union(mesh.cells[level], mesh.proj[level]))
    e = mesh.domain_inter(level, pred_ghosts)
    if e.size == 0:
        return e
    parent = pack(unpack(e, dim) >> 1)
    return e[mesh.contains(level - 1, parent)]


def periodic_pairs(mesh: Mesh, level, d):
    """iterate_over_periodic_ghosts for dimension d (algorithm/update_periodic.hpp:34-125): returns (ghost keys, source keys).
    Ghosts within ghost_width beyond the upper boundary mirror the cells just inside the lower boundary (set1), ghosts before the
    lower boundary mirror the cells just inside the upper one (set2); in the other dimensions the slabs span the domain grown by
    ghost_width; a pair exists where both cells are in the reference sub-mesh."""
    cfg, dim = mesh.cfg, mesh.cfg.dim
    ref = mesh.ref[level]
    if ref.size == 0:
        return EMPTY, EMPTY
    gw = cfg.ghost_width
    c = unpack(ref, dim)
    n = [cfg.n_cells0[k] << level for k in range(dim)]
    inside_others = np.ones(ref.size, dtype=bool)
    for k in range(dim):
        if k != d:
            inside_others &= (c[:, k] >= -gw) & (c[:, k] < n[k] + gw)
    shift = [0] * dim
    shift[d] = n[d]
    ghosts, sources = [], []
    # set1: translate(ref & lca_min_p, +shift) & (ref & lca_max_p)
    src = ref[inside_others & (c[:, d] >= 0) & (c[:, d] < gw)]
    dst = ref[inside_others & (c[:, d] >= n[d]) & (c[:, d] < n[d] + gw)]
    g = inter(translate(src, shift), dst)
    ghosts.append(g)
    sources.append(translate(g, [-s for s in shift]))
    # set2: translate(ref & lca_max_m, -shift) & (ref & lca_min_m)
    src = ref[inside_others & (c[:, d] >= n[d] - gw) & (c[:, d] < n[d])]
    dst = ref[inside_others & (c[:, d] >= -gw) & (c[:, d] < 0)]
    g = inter(translate(src, [-s for s in shift]), dst)
    ghosts.append(g)
    sources.append(translate(g, shift))
    return np.concatenate(ghosts), np.concatenate(sources)


def update_ghost_periodic(mesh: Mesh, f, level):
    """update_ghost_periodic(level, field) (algorithm/update_periodic.hpp:23-32): dimension after dimension."""
    for d in range(mesh.cfg.dim):
        if mesh.cfg.periodic[d]:
            g, s = periodic_pairs(mesh, level, d)
            if g.size:
                f[mesh.index(level, g)] = f[mesh.index(level, s)]


def update_tag_periodic(mesh: Mesh, tag, level):
    """update_tag_periodic (algorithm/update_periodic.hpp:211-300): a ghost and its mirrored cell take the OR of their tags."""
    for d in range(mesh.cfg.dim):
        if mesh.cfg.periodic[d]:
            g, s = periodic_pairs(mesh, level, d)
            if g.size:
                gi, si = mesh.index(level, g), mesh.index(level, s)
                v = tag[gi] | tag[si]
                tag[gi] = v
                tag[si] = v


def update_ghost_mr(mesh: Mesh, f, bc: Bc):
    """update_ghost_mr_aggregated (algorithm/update_ghost_mr.hpp:194-237), serial."""
    L = mesh.cfg.max_level
    periodic = any(mesh.cfg.periodic)
    for level in range(L, -1, -1):
        if periodic:
            update_ghost_periodic(mesh, f, level)
        update_outer_ghosts(mesh, f, bc, level)
        if periodic:
            update_ghost_periodic(mesh, f, level)
        if level > 0:
            projection(mesh, f, level - 1, projection_set(mesh, level))
    for level in range(1, L + 1):
        e = prediction_set(mesh, level)
        if e.size:
            f[mesh.index(level, e)] = predict_values(mesh, f, mesh, level, e)
        if periodic:
            update_ghost_periodic(mesh, f, level)


# ----------------------------------------------------------------------------
# MR adaptation
# ----------------------------------------------------------------------------
def detail_set(mesh: Mesh, level):
    """mr/adapt.hpp:312-315."""
    dim = mesh.cfg.dim
    below = union(coarsen(mesh.cells[level + 1], 1, dim), coarsen(mesh.cells[level + 2], 2, dim) if level + 2 < mesh.nlev else EMPTY)
    return inter(mesh.ref[level], below)


def tag_set(mesh: Mesh, level):
    """coarse cells (level-1) having a leaf child at `level` (mr/adapt.hpp:333, 351)."""
    dim = mesh.cfg.dim
    return inter(mesh.ref[level - 1], coarsen(mesh.cells[level], 1, dim))


def mr_criteria(mesh: Mesh, detail, tag, level, eps, regularity):
    """mr_criteria_op (mr/criteria.hpp:19-113), fine level = `level`, coarse = level-1."""
    cfg, dim = mesh.cfg, mesh.cfg.dim
    coarse = tag_set(mesh, level)
    if coarse.size == 0:
        return
    exponent = dim * (cfg.max_level - level)
    eps_l = eps / (1 << exponent)  # mr/adapt.hpp:328-329
    reg = regularity + dim
    fine_eps = pow(2.0, reg) * eps_l
    coarse_eps = fine_eps / (1 << dim)
    c = unpack(coarse, dim)
    ci = mesh.index(level - 1, coarse)
    fi = [mesh.index(level, pack(2 * c + np.array(ch, dtype=np.int64))) for ch in _children_offsets(dim)]
    # several adapted fields / components: coarsen needs every component below its threshold, refine needs any above
    # (mr/criteria.hpp:33-37, 49-53, 81-85: loops over n_comp)
    comps = detail if isinstance(detail, (list, tuple)) else [detail]
    if level > cfg.min_level:
        cond = np.ones(coarse.size, dtype=bool)
        for dc in comps:
            cond &= ~(np.abs(dc[ci]) > coarse_eps)
            for i in fi:
                cond &= ~(np.abs(dc[i]) > eps_l)
        for i in fi:
            tag[i[cond]] = COARSEN
    if level < cfg.max_level:
        for i in fi:
            ref = np.zeros(coarse.size, dtype=bool)
            for dc in comps:
                ref |= np.abs(dc[i]) > fine_eps
            tag[i] |= ref.astype(np.uint8) * np.uint8(REFINE)


def maximum(mesh: Mesh, tag, level):
    """maximum_op (mr/operators.hpp:29-89) on parents (level-1) of leaves at `level`."""
    dim = mesh.cfg.dim
    coarse = tag_set(mesh, level)
    if coarse.size == 0:
        return
    c = unpack(coarse, dim)
    ci = mesh.index(level - 1, coarse)
    fi = [mesh.index(level, pack(2 * c + np.array(ch, dtype=np.int64))) for ch in _children_offsets(dim)]
    any_keep = np.zeros(coarse.size, dtype=bool)
    all_coarsen = np.ones(coarse.size, dtype=bool)
    for i in fi:
        any_keep |= (tag[i] & KEEP) != 0
        all_coarsen &= (tag[i] & COARSEN) != 0
    for i in fi:
        tag[i[any_keep]] |= KEEP
    tag[ci[any_keep]] |= KEEP
    m2 = ~any_keep & all_coarsen
    tag[ci[m2]] |= KEEP
    m3 = ~any_keep & ~all_coarsen
    for i in fi:
        tag[i[m3]] &= np.uint8(0xFF & ~COARSEN)


def update_cell_array_from_tag(mesh: Mesh, tag):
    """algorithm/graduation.hpp:743-842 (set semantics)."""
    cfg, dim, nlev = mesh.cfg, mesh.cfg.dim, mesh.nlev
    add = [[] for _ in range(nlev)]
    rem = [[] for _ in range(nlev)]
    for l in mesh.leaf_levels():
        k = mesh.cells[l]
        t = tag[mesh.index(l, k)]
        ref_ = ((t & REFINE) != 0) & (l < cfg.max_level)
        coa = ~ref_ & ((t & COARSEN) != 0) & ((t & KEEP) == 0) & (l > cfg.min_level)
        if ref_.any():
            rem[l].append(k[ref_])
            add[l + 1].append(refine(k[ref_], 1, dim))
        if coa.any():
            rem[l].append(k[coa])
            c = unpack(k[coa], dim)
            even = ((c & 1) == 0).all(axis=1)
            if even.any():
                add[l - 1].append(np.unique(pack(c[even] >> 1)))
    new = [EMPTY] * nlev
    for l in range(cfg.min_level, cfg.max_level + 1):
        s = union(mesh.cells[l], *add[l])
        r = union(*rem[l]) if rem[l] else EMPTY
        new[l] = diff(s, r)
    return new


def make_graduation(cfg: MeshConfig, ca):
    """make_graduation (algorithm/graduation.hpp:573-726) for max_stencil_radius == 1, serial, non-periodic."""
    dim = cfg.dim
    w = cfg.graduation_width
    nlev = len(ca)
    ca = list(ca)
    while True:
        levels = [l for l in range(nlev) if ca[l].size]
        if not levels:
            return ca
        lo, hi = levels[0], levels[-1]
        out = [[] for _ in range(nlev)]
        for fine in range(hi, lo + 1, -1):  # fine_level = max_level ... min_level+2
            if ca[fine].size == 0:
                continue
            # the fine cells and, on periodic meshes, their images (graduation.hpp:309-320)
            for fine_set in [ca[fine]] + [translate(ca[fine], d) for d in cfg.periodic_directions(fine)]:
                proj = coarsen(expand(fine_set, 2 * w, dim), 2, dim)
                coarse_level = fine - 2
                while True:
                    if proj.size:
                        r = inter(proj, ca[coarse_level])
                        if r.size:
                            out[coarse_level].append(r)
                    if coarse_level == lo or proj.size == 0:
                        break
                    proj = coarsen(proj, 1, dim)
                    coarse_level -= 1
        # cells two steps inward of a boundary leaf must not lie in a coarser leaf (graduation.hpp:372-455, max_stencil_radius 2:
        # n_contiguous_boundary_cells = max(2, 2 * (2 - 2)) = 2; the level -> level + 1 part only exists for radius > 2)
        if cfg.max_stencil_radius == 2:
            for direction in cartesian_directions(dim):
                if any(direction[k] != 0 and cfg.periodic[k] for k in range(dim)):
                    continue
                for level in range(hi, lo, -1):
                    if ca[level].size == 0 or ca[level - 1].size == 0:
                        continue
                    k = ca[level]
                    c = unpack(translate(k, direction), dim)
                    outside = np.zeros(k.size, dtype=bool)
                    for a in range(dim):
                        outside |= (c[:, a] < 0) | (c[:, a] >= (cfg.n_cells0[a] << level))
                    bdry = k[outside]
                    if bdry.size:
                        r = inter(coarsen(translate(bdry, [-2 * x for x in direction]), 1, dim), ca[level - 1])
                        if r.size:
                            out[level - 1].append(r)
        if not any(out):
            return ca
        rem = [union(*o) if o else EMPTY for o in out]
        # new_ca[level] = (ca[level] U children(remove[level-1])) \ remove[level]  (graduation.hpp:706-716)
        new = [diff(union(ca[l], refine(rem[l - 1], 1, dim) if l > 0 and rem[l - 1].size else EMPTY), rem[l]) for l in range(nlev)]
        changed = any(new[l].size != ca[l].size or not np.array_equal(new[l], ca[l]) for l in range(nlev))
        ca = new
        if not changed:
            return ca


def update_fields(old: Mesh, new: Mesh, f):
    """update_fields (algorithm/update_fields.hpp:27-54,101-127): copy, project, predict onto the new mesh."""
    cfg, dim = old.cfg, old.cfg.dim
    g = np.zeros(new.nref)
    for l in range(cfg.min_level, cfg.max_level + 1):
        s = inter(old.ref[l], new.cells[l])
        if s.size:
            g[new.index(l, s)] = f[old.index(l, s)]
    for l in range(cfg.min_level + 1, cfg.max_level + 1):
        sc = inter(coarsen(old.cells[l], 1, dim), new.cells[l - 1])
        if sc.size:
            c = unpack(sc, dim)
            total = np.zeros(sc.size)
            rows = [tuple(reversed(r)) for r in itertools.product((0, 1), repeat=dim - 1)] if dim > 1 else [()]
            for row in rows:
                k0 = pack(2 * c + np.array((0,) + row, dtype=np.int64))
                total = total + (f[old.index(l, k0)] + f[old.index(l, k0 + 1)])
            g[new.index(l - 1, sc)] = total * (1.0 / float(1 << dim))
        # set_refine = (new cells[l] ∩ old cells[l-1]).on(l-1): each coarse cell writes ALL 2^dim children
        sr = inter(coarsen(new.cells[l], 1, dim), old.cells[l - 1])
        if sr.size:
            fine = refine(sr, 1, dim)
            g[new.index(l, fine)] = predict_values(new, f, old, l, fine)
    return g


@dataclass
class AdaptStats:
    iterations: int = 0
    changed: bool = False


def compute_relative_detail(mesh: Mesh, f, detail):
    """compute_relative_detail (mr/rel_detail.hpp:73-112): detail *= 1 / max_leaves |f| over the whole array."""
    m = np.finfo(np.float64).tiny  # std::numeric_limits<double>::min()
    for l in mesh.leaf_levels():
        m = max(m, float(np.max(np.abs(f[mesh.index(l, mesh.cells[l])]))))
    if m < np.finfo(np.float64).eps:
        m = 1.0
    detail *= 1.0 / m


def keep_boundary_refined(mesh: Mesh, tag):
    """keep_boundary_refined (mr/adapt.hpp:245-274, `--refine-boundary`): after the criteria, the leaves of max_level within
    max_stencil_radius cells of the domain boundary, direction after direction (boundary.hpp:6-22: cells minus translate(domain, -w * dir)),
    get tag = keep (assignment)."""
    cfg, dim, L = mesh.cfg, mesh.cfg.dim, mesh.cfg.max_level
    cells = mesh.cells[L]
    if cells.size == 0:
        return
    w = cfg.max_stencil_radius
    c = unpack(cells, dim)
    for d in range(dim):
        n = cfg.n_cells0[d] << L
        for sel in (c[:, d] >= n - w, c[:, d] < w):  # direction +e_d, then -e_d
            if sel.any():
                tag[mesh.index(L, cells[sel])] = KEEP


def adapt(mesh: Mesh, f, bc: Bc, eps=1e-4, regularity=1.0, trace=None, relative_detail=False, refine_boundary=False):
    """Adapt::operator() + harten (mr/adapt.hpp:148-195, 277-389). Returns (mesh, field)."""
    cfg = mesh.cfg
    lmin, L = cfg.min_level, cfg.max_level
    if lmin == L:
        return mesh, f
    for ite in range(L - lmin):
        detail = np.zeros(mesh.nref)
        tag = np.zeros(mesh.nref, dtype=np.uint8)
        for l in mesh.leaf_levels():
            tag[mesh.index(l, mesh.cells[l])] = KEEP
        update_ghost_mr(mesh, f, bc)
        for level in range(max(lmin - 1, 0), L - ite):
            compute_detail(mesh, f, detail, level, detail_set(mesh, level))
        if relative_detail:
            compute_relative_detail(mesh, f, detail)
        for level in range(lmin, L - ite + 1):
            mr_criteria(mesh, detail, tag, level, eps, regularity)
        if refine_boundary:  # mr/adapt.hpp:340-345
            keep_boundary_refined(mesh, tag)
        for level in range(L, 0, -1):
            update_tag_periodic(mesh, tag, level)  # mr/adapt.hpp:353
            maximum(mesh, tag, level)
        if trace is not None:
            trace.append(dict(ite=ite, mesh=mesh, field=f.copy(), detail=detail, tag=tag.copy()))
        new_ca = update_cell_array_from_tag(mesh, tag)
        new_ca = make_graduation(cfg, new_ca)
        if mesh.same_cells(new_ca):
            break
        new_mesh = Mesh(cfg, {l: new_ca[l] for l in range(len(new_ca)) if new_ca[l].size})
        f = update_fields(mesh, new_mesh, f)
        mesh = new_mesh
    return mesh, f


def adapt_fields(mesh: Mesh, fields, bcs, eps=1e-4, regularity=1.0):
    """make_MRAdapt(u, v, ...)(mra_config): several fields adapted together (mr/adapt.hpp:103-107, 148-195, 277-389 with a
    Field_tuple: one detail array per field, one tag array).  Returns (mesh, [fields])."""
    cfg = mesh.cfg
    lmin, L = cfg.min_level, cfg.max_level
    fields = list(fields)
    if lmin == L:
        return mesh, fields
    for ite in range(L - lmin):
        details = [np.zeros(mesh.nref) for _ in fields]
        tag = np.zeros(mesh.nref, dtype=np.uint8)
        for l in mesh.leaf_levels():
            tag[mesh.index(l, mesh.cells[l])] = KEEP
        for f, bc in zip(fields, bcs):
            update_ghost_mr(mesh, f, bc)
        for level in range(max(lmin - 1, 0), L - ite):
            for f, d in zip(fields, details):
                compute_detail(mesh, f, d, level, detail_set(mesh, level))
        for level in range(lmin, L - ite + 1):
            mr_criteria(mesh, details, tag, level, eps, regularity)
        for level in range(L, 0, -1):
            update_tag_periodic(mesh, tag, level)
            maximum(mesh, tag, level)
        new_ca = update_cell_array_from_tag(mesh, tag)
        new_ca = make_graduation(cfg, new_ca)
        if mesh.same_cells(new_ca):
            break
        new_mesh = Mesh(cfg, {l: new_ca[l] for l in range(len(new_ca)) if new_ca[l].size})
        fields = [update_fields(mesh, new_mesh, f) for f in fields]
        mesh = new_mesh
    return mesh, fields


# ----------------------------------------------------------------------------
# FV operators (field-expression path)
# ----------------------------------------------------------------------------
def _upwind_flux(a, ul, ur):
    return (0.5 * a) * (ul + ur) + (0.5 * abs(a)) * (ul - ur)  # stencil_field.hpp:92-97


def _burgers_flux(a, ul, ur):
    """upwind_scalar_burgers_op::flux (stencil_field.hpp:193-217)."""
    out = np.zeros_like(ul)
    mask1 = (a * ul) < (a * ur)
    mask2 = (ul * ur) > 0.0
    mn = np.minimum(np.abs(ul), np.abs(ur))
    mx = np.maximum(np.abs(ul), np.abs(ur))
    m = mask1 & mask2
    out[m] = 0.5 * mn[m] * mn[m]
    m = ~mask1
    out[m] = 0.5 * mx[m] * mx[m]
    return out


def fv_step(mesh: Mesh, u, a, dt, scheme="upwind"):
    """unp1 = u - dt * upwind(a, u) over the leaves (field_base.hpp:230-242, stencil_field.hpp:30-54).

    Non-leaf entries of the result are NaN: the reference leaves them unspecified
    (`unp1.resize()` + swap), so nothing downstream may depend on them.
    """
    dim = mesh.cfg.dim
    flux = _upwind_flux if scheme == "upwind" else _burgers_flux
    unp1 = np.full(mesh.nref, np.nan)
    for l in mesh.leaf_levels():
        k = mesh.cells[l]
        ic = mesh.index(l, k)
        uc = u[ic]
        dx = mesh.cfg.cell_length(l)
        acc = None
        for d in range(dim):
            e = [0] * dim
            e[d] = 1
            um = u[mesh.index(l, translate(k, [-v for v in e]))]
            up = u[mesh.index(l, translate(k, e))]
            lo = flux(a[d], um, uc)
            hi = flux(a[d], uc, up)
            acc = (-lo + hi) if acc is None else ((acc + -lo) + hi)
        unp1[ic] = uc - dt * (acc / dx)
    return unp1


# ----------------------------------------------------------------------------
# flux-based linear homogeneous schemes (uniform-level meshes)
# ----------------------------------------------------------------------------
def convection_upwind_coeffs(velocity):
    """make_convection_upwind<Field>(velocity) (schemes/fv/operators/convection_lin.hpp:15-89): flux coeffs {left, right}."""
    def fn(d, h):
        v = float(velocity[d])
        return (v, 0.0) if v >= 0 else (0.0, v)
    return fn


def diffusion_order2_coeffs(K):
    """make_diffusion_order2<Field>(K) (schemes/fv/operators/diffusion.hpp:123-175)."""
    def fn(d, h):
        left, right = -1 / h, 1 / h
        left *= -float(K[d])
        right *= -float(K[d])
        return (left, right)
    return fn


def _runs(keys, dim):
    """Split sorted cell keys into x-intervals (runs of consecutive x in one row), in for_each_interval order."""
    if keys.size == 0:
        return []
    brk = np.flatnonzero(np.diff(keys) != 1) + 1
    return np.split(keys, brk)


def flux_linhom_apply(mesh: Mesh, u, coeff_fn):
    """Explicit<FluxBasedScheme<LinearHomogeneous>>::apply in its sequential context, restated literally
    (flux_based/explicit_flux_based_scheme__lin_hom.hpp:39-119, 233-319; coefficient / level loop
    flux_based_scheme__lin_hom.hpp:74-229; interface sets interface.hpp:35-110, 125-306, 440-509).
    For every direction: same-level interface intervals of every level, then per level the two level-jump orientations,
    then the boundary interfaces; every interval applies, for c = 0, 1:  out[left] += lc[c]*in[st_c]; out[right] += rc[c]*in[st_c]
    (coarse side of a jump: lc*in[2ii] + lc*in[2ii+1])."""
    cfg, dim = mesh.cfg, mesh.cfg.dim
    out = np.zeros(mesh.nref)
    leaf_lv = mesh.leaf_levels()
    lo = max(cfg.min_level, leaf_lv[0] - 1)
    hi = min(cfg.max_level, leaf_lv[-1] + 1)

    def hfac(h_face, h_cell):
        return pow(h_face, dim - 1) / pow(h_cell, dim)

    for d in range(dim):
        e = [0] * dim
        e[d] = 1
        me = [-v for v in e]

        def shift_of(level, sign):
            sh = [0] * dim
            sh[d] = sign * (cfg.n_cells0[d] << level)
            return sh

        # ---- same level
        for level in range(lo, hi + 1):
            cells = mesh.cells[level]
            if cells.size == 0:
                continue
            h = cfg.cell_length(level)
            fc = coeff_fn(d, h)
            f = hfac(h, h)
            lc = (f * fc[0], f * fc[1])
            rc = (-lc[0], -lc[1])
            pairs = [(cells, cells)]
            if cfg.periodic[d]:  # interfaces through the periodic boundary, seen once from each side (interface.hpp:83-92)
                pairs += [(cells, translate(cells, shift_of(level, 1))), (translate(cells, shift_of(level, -1)), cells)]
            for left_set, right_set in pairs:
                iface = inter(left_set, translate(right_set, me))
                for run in _runs(iface, dim):
                    li = mesh.index(level, run)
                    ri = mesh.index(level, translate(run, e))
                    st = (li, ri)
                    for cc in range(2):
                        out[li] = out[li] + lc[cc] * u[st[cc]]
                        out[ri] = out[ri] + rc[cc] * u[st[cc]]
        # ---- level jumps level -> level+1
        for level in range(lo, hi):
            coarse, fine = mesh.cells[level], mesh.cells[level + 1]
            if coarse.size == 0 or fine.size == 0:
                continue
            h_l, h_f = cfg.cell_length(level), cfg.cell_length(level + 1)
            fc = coeff_fn(d, h_f)  # flux computed at level+1
            rcoarse = refine(coarse, 1, dim)
            # orientation A: coarse on the left, fine on the right
            lcA = tuple(hfac(h_f, h_l) * v for v in fc)
            rcA = tuple(-hfac(h_f, h_f) * v for v in fc)
            pairsA = [(coarse, fine)]
            if cfg.periodic[d]:  # interface.hpp:179-189
                pairsA += [(coarse, translate(fine, shift_of(level + 1, 1))), (translate(coarse, shift_of(level, -1)), fine)]
            for cs_, fs_ in pairsA:
                ghosts = inter(refine(cs_, 1, dim), translate(fs_, me))
                for run in _runs(ghosts, dim):
                    st = (mesh.index(level + 1, run), mesh.index(level + 1, translate(run, e)))
                    right = st[1]
                    if run.size == 1 or d == 0:
                        left = mesh.index(level, pack(unpack(run, dim) >> 1))
                        for cc in range(2):
                            out[left] = out[left] + lcA[cc] * u[st[cc]]
                            out[right] = out[right] + rcA[cc] * u[st[cc]]
                    else:
                        assert run.size % 2 == 0
                        left = mesh.index(level, pack(unpack(run[0::2], dim) >> 1))
                        for cc in range(2):
                            out[left] = out[left] + (lcA[cc] * u[st[cc][0::2]] + lcA[cc] * u[st[cc][1::2]])
                            out[right] = out[right] + rcA[cc] * u[st[cc]]
            # orientation B: fine on the left, coarse on the right; stencil {-dir, 0} around the ghost child
            lcB = tuple(hfac(h_f, h_f) * v for v in fc)
            rcB = tuple(-hfac(h_f, h_l) * v for v in fc)
            pairsB = [(coarse, fine)]
            if cfg.periodic[d]:  # interface.hpp:280-290
                pairsB += [(coarse, translate(fine, shift_of(level + 1, -1))), (translate(coarse, shift_of(level, 1)), fine)]
            for cs_, fs_ in pairsB:
                ghosts = inter(refine(cs_, 1, dim), translate(fs_, e))
                for run in _runs(ghosts, dim):
                    st = (mesh.index(level + 1, translate(run, me)), mesh.index(level + 1, run))
                    left = st[0]
                    if run.size == 1 or d == 0:
                        right = mesh.index(level, pack(unpack(run, dim) >> 1))
                        for cc in range(2):
                            out[left] = out[left] + lcB[cc] * u[st[cc]]
                            out[right] = out[right] + rcB[cc] * u[st[cc]]
                    else:
                        assert run.size % 2 == 0
                        right = mesh.index(level, pack(unpack(run[0::2], dim) >> 1))
                        for cc in range(2):
                            out[left] = out[left] + lcB[cc] * u[st[cc]]
                            out[right] = out[right] + (rcB[cc] * u[st[cc][0::2]] + rcB[cc] * u[st[cc][1::2]])
        # ---- boundary interfaces: per level, direction then opposite direction (none in a periodic direction,
        # flux_based_scheme__lin_hom.hpp:189-192)
        if cfg.periodic[d]:
            continue
        for level in leaf_lv:
            cells = mesh.cells[level]
            h = cfg.cell_length(level)
            fc = coeff_fn(d, h)
            f = hfac(h, h)
            bc_ = (f * fc[0], f * fc[1])
            bd = cells[~mesh.in_domain(level, translate(cells, e))]
            for run in _runs(bd, dim):
                bi = mesh.index(level, run)
                st = (bi, mesh.index(level, translate(run, e)))
                for cc in range(2):
                    out[bi] = out[bi] + bc_[cc] * u[st[cc]]
            bd = cells[~mesh.in_domain(level, translate(cells, me))]
            for run in _runs(bd, dim):
                bi = mesh.index(level, run)
                st = (mesh.index(level, translate(run, me)), bi)
                for cc in range(2):
                    out[bi] = out[bi] + (-bc_[cc]) * u[st[cc]]
    return out


def burgers_upwind_flux(scale=1.0):
    """`scale * make_convection_upwind<Field>()` for a scalar field (schemes/fv/operators/convection_nonlin.hpp:24-76;
    scalar factor: flux_based/algebraic_operators.hpp:38-46): v = .5*(uL+uR); flux = (v >= 0 ? uL*uL : uR*uR) [* scale]."""
    def fn(ul, ur):
        v = 0.5 * (ul + ur)
        f = np.where(v >= 0, ul * ul, ur * ur)
        return f * scale if scale != 1 else f
    return fn


def burgers_upwind_flux_vector(dim, scale=1.0):
    """`scale * make_convection_upwind<VectorField>()`, n_comp == dim (schemes/fv/operators/convection_nonlin.hpp:24-76): in direction d,
    v = .5*(uL[d]+uR[d]); flux = v >= 0 ? uL[d]*uL : uR[d]*uR (every component)."""
    def make(d):
        def fn(ul, ur):
            v = 0.5 * (ul[d] + ur[d])
            out = []
            for c in range(len(ul)):
                f = np.where(v >= 0, ul[d] * ul[c], ur[d] * ur[c])
                out.append(f * scale if scale != 1 else f)
            return out
        return fn
    return [make(d) for d in range(dim)]


def _weno5(f):
    """compute_weno5_flux (schemes/fv/operators/weno_impl.hpp:26-63, Jiang & Shu 1996) of the five flux values f[0..4], same
    operation order; pow(x, 2) taken as x * x."""
    j = 2
    q0 = 1. / 3 * f[j - 2] - 7. / 6 * f[j - 1] + 11. / 6 * f[j]
    q1 = -1. / 6 * f[j - 1] + 5. / 6 * f[j] + 1. / 3 * f[j + 1]
    q2 = 1. / 3 * f[j] + 5. / 6 * f[j + 1] - 1. / 6 * f[j + 2]
    sq = lambda x: x * x
    IS0 = 13. / 12 * sq(f[j - 2] - 2 * f[j - 1] + f[j]) + 1. / 4 * sq(f[j - 2] - 4 * f[j - 1] + 3 * f[j])
    IS1 = 13. / 12 * sq(f[j - 1] - 2 * f[j] + f[j + 1]) + 1. / 4 * sq(f[j - 1] - f[j + 1])
    IS2 = 13. / 12 * sq(f[j] - 2 * f[j + 1] + f[j + 2]) + 1. / 4 * sq(3 * f[j] - 4 * f[j + 1] + f[j + 2])
    eps = 1e-6
    a0 = 0.1 / sq(eps + IS0)
    a1 = 0.6 / sq(eps + IS1)
    a2 = 0.3 / sq(eps + IS2)
    sa = a0 + a1 + a2
    return (a0 / sa) * q0 + (a1 / sa) * q1 + (a2 / sa) * q2


def weno5_flux(velocity):
    """make_convection_weno5 for a scalar field (schemes/fv/operators/convection_lin.hpp:95-178): per direction d a function of the six
    stencil values u[-2..3]; f = velocity[d] * (u0..u4) if velocity[d] >= 0 else velocity[d] * (u5..u1), then WENO5."""
    def make(d):
        v = float(velocity[d])

        def fn(u0, u1, u2, u3, u4, u5):
            return _weno5([u0 * v, u1 * v, u2 * v, u3 * v, u4 * v] if v >= 0 else [u5 * v, u4 * v, u3 * v, u2 * v, u1 * v])
        return fn
    return [make(d) for d in range(len(velocity))]


def weno5_flux_nonlinear(dim, n_comp=1, scale=1.0):
    """`scale * make_convection_weno5<Field>()` (schemes/fv/operators/convection_nonlin.hpp:162-233): f(u) = u * u (scalar) or
    u(d) * u (vector, n_comp == dim); upwinded by v = .5 * (u[2] + u[3]) (component d for vectors): WENO5 of f(u0..u4) if v >= 0
    else of f(u5..u1), component by component."""
    def make(d):
        def fn(*u):
            if n_comp == 1:
                v = 0.5 * (u[2] + u[3])
                f = [x * x for x in u]
                out = np.where(v >= 0, _weno5(f[0:5]), _weno5([f[5], f[4], f[3], f[2], f[1]]))
                return out * scale if scale != 1 else out
            v = 0.5 * (u[2][d] + u[3][d])
            outs = []
            for c in range(n_comp):
                f = [x[d] * x[c] for x in u]
                o = np.where(v >= 0, _weno5(f[0:5]), _weno5([f[5], f[4], f[3], f[2], f[1]]))
                outs.append(o * scale if scale != 1 else o)
            return outs
        return fn
    return [make(d) for d in range(dim)]


WENO5_OFFSETS = (-2, -1, 0, 1, 2, 3)  # line_stencil<dim, d>(-2, -1, 0, 1, 2, 3), convection_lin.hpp:115


def _scatter_pairs(out, left, a, right, b):
    """for k: out[left[k]] += a[k]; out[right[k]] += b[k] -- the reference's sequential scatter, vectorised where that keeps every
    cell's order of additions: disjoint sides (np.add.at is sequential per index), or the x-direction run right == left + 1 with
    unique cells, where a cell first receives b (as the right cell of interface k - 1) and then a (as the left cell of interface k)."""
    if left.size > 1 and np.array_equal(right[:-1], left[1:]) and np.all(np.diff(left) == 1):
        out[right] += b
        out[left] += a
    elif left.size == 1 or not np.intersect1d(left, right).size:
        np.add.at(out, left, a)
        np.add.at(out, right, b)
    else:
        for k in range(left.size):
            out[left[k]] = out[left[k]] + a[k]
            out[right[k]] = out[right[k]] + b[k]


def flux_nonlin_apply(mesh: Mesh, u, flux_fn, offsets=(0, 1)):
    """Explicit<FluxBasedScheme<NonLinear>>::apply with finer_level_flux disabled, sequential order
    (flux_based/explicit_flux_based_scheme__nonlin.hpp:35-85; flux_based_scheme__nonlin.hpp:334-364 interior,
    :366-402 boundary, :407-520 level loop and factors; fluxes[1] = -fluxes[0]: flux_definition.hpp:125-136).
    Per interface cell, in interval order: out[left] += flux*left_factor, then out[right] += (-flux)*right_factor.
    `offsets`: the line stencil of the flux (cells origin + o * e_d, origin = the cell left of the interface at the level the flux
    is computed on); `flux_fn`: one function of the stencil values, or one per direction.  Periodic directions add the interfaces
    through the boundary, seen once from each side (interface.hpp:83-92, 179-189, 280-290), and have no boundary interfaces
    (flux_based_scheme__nonlin.hpp:537-540)."""
    cfg, dim = mesh.cfg, mesh.cfg.dim
    vector = isinstance(u, (list, tuple))  # vector field: one array per component; the flux functions get/return per-component lists
    outs = [np.zeros(mesh.nref) for _ in (u if vector else [u])]
    out = outs[0]
    leaf_lv = mesh.leaf_levels()
    lo = max(cfg.min_level, leaf_lv[0] - 1)
    hi = min(cfg.max_level, leaf_lv[-1] + 1)

    def hfac(h_face, h_cell):
        return pow(h_face, dim - 1) / pow(h_cell, dim)

    for d in range(dim):
        fn = flux_fn[d] if isinstance(flux_fn, (list, tuple)) else flux_fn
        e = [0] * dim
        e[d] = 1
        me = [-v for v in e]

        def scatter(slevel, origin, run_left, run_right, lf, rf):
            # sequential per-cell accumulation (several fine cells may hit the same coarse cell)
            idx = [mesh.index(slevel, translate(origin, [o * x for x in e])) for o in offsets]
            if vector:
                fls = fn(*[[uc[i] for uc in u] for i in idx])  # one flux array per component
            else:
                fls = [fn(*[u[i] for i in idx])]
            for oc, fl in zip(outs, fls):
                a = fl * lf
                b = (-fl) * rf
                _scatter_pairs(oc, run_left, a, run_right, b)

        def shift_of(level, sign):
            sh = [0] * dim
            sh[d] = sign * (cfg.n_cells0[d] << level)
            return sh

        for level in range(lo, hi + 1):
            cells = mesh.cells[level]
            if cells.size == 0:
                continue
            h = cfg.cell_length(level)
            f = hfac(h, h)
            pairs = [(cells, cells)]
            if cfg.periodic[d]:
                pairs += [(cells, translate(cells, shift_of(level, 1))), (translate(cells, shift_of(level, -1)), cells)]
            for left_set, right_set in pairs:
                iface = inter(left_set, translate(right_set, me))
                for run in _runs(iface, dim):
                    li = mesh.index(level, run)
                    ri = mesh.index(level, translate(run, e))
                    scatter(level, run, li, ri, f, f)
        for level in range(lo, hi):
            coarse, fine = mesh.cells[level], mesh.cells[level + 1]
            if coarse.size == 0 or fine.size == 0:
                continue
            h_l, h_f = cfg.cell_length(level), cfg.cell_length(level + 1)
            pairs = [(coarse, fine)]
            if cfg.periodic[d]:
                pairs += [(coarse, translate(fine, shift_of(level + 1, 1))), (translate(coarse, shift_of(level, -1)), fine)]
            for cs, fs in pairs:
                ghosts = inter(refine(cs, 1, dim), translate(fs, me))
                for run in _runs(ghosts, dim):
                    st1 = mesh.index(level + 1, translate(run, e))
                    left = mesh.index(level, pack(unpack(run, dim) >> 1))
                    scatter(level + 1, run, left, st1, hfac(h_f, h_l), hfac(h_f, h_f))
            pairs = [(coarse, fine)]
            if cfg.periodic[d]:
                pairs += [(coarse, translate(fine, shift_of(level + 1, -1))), (translate(coarse, shift_of(level, 1)), fine)]
            for cs, fs in pairs:
                ghosts = inter(refine(cs, 1, dim), translate(fs, e))
                for run in _runs(ghosts, dim):
                    st0 = mesh.index(level + 1, translate(run, me))
                    right = mesh.index(level, pack(unpack(run, dim) >> 1))
                    scatter(level + 1, translate(run, me), st0, right, hfac(h_f, h_f), hfac(h_f, h_l))
        if cfg.periodic[d]:
            continue
        assert tuple(offsets) == (0, 1), "boundary interfaces are restated for two-cell stencils"
        for level in leaf_lv:
            cells = mesh.cells[level]
            h = cfg.cell_length(level)
            f = hfac(h, h)
            bd = cells[~mesh.in_domain(level, translate(cells, e))]
            for run in _runs(bd, dim):
                bi = mesh.index(level, run)
                ni = mesh.index(level, translate(run, e))
                fls = fn([uc[bi] for uc in u], [uc[ni] for uc in u]) if vector else [fn(u[bi], u[ni])]
                for oc, fl in zip(outs, fls):
                    oc[bi] = oc[bi] + fl * f
            bd = cells[~mesh.in_domain(level, translate(cells, me))]
            for run in _runs(bd, dim):
                bi = mesh.index(level, run)
                ni = mesh.index(level, translate(run, me))
                fls = fn([uc[ni] for uc in u], [uc[bi] for uc in u]) if vector else [fn(u[ni], u[bi])]
                for oc, fl in zip(outs, fls):
                    oc[bi] = oc[bi] + (-fl) * f  # flux_values[1] *= -(-factor)
    return outs if vector else out


def lincomb_leaves(mesh: Mesh, a, x, b, y):
    """`unp1 = a*x + b*y` over the leaves (field expression, field_base.hpp:230-242); other entries NaN."""
    out = np.full(mesh.nref, np.nan)
    for l in mesh.leaf_levels():
        i = mesh.index(l, mesh.cells[l])
        out[i] = a * x[i] + b * y[i] if a != 1.0 else x[i] + b * y[i]
    return out


# ----------------------------------------------------------------------------
# demo drivers
# ----------------------------------------------------------------------------
def init_disc(mesh: Mesh, center, radius):
    """advection_2d.cpp:23-45 / advection_3d.cpp:32-45: 1 inside the ball, else 0."""
    u = np.zeros(mesh.nref)
    for l in mesh.leaf_levels():
        k = mesh.cells[l]
        x = mesh.cell_centers(l, k)
        r2 = None
        for d in range(mesh.cfg.dim):
            t = (x[:, d] - center[d]) * (x[:, d] - center[d])
            r2 = t if r2 is None else r2 + t
        u[mesh.index(l, k)] = np.where(r2 <= radius * radius, 1.0, 0.0)
    return u


def run_advection(cfg: MeshConfig, Tf, eps=2e-4, regularity=1.0, a=None, cfl=0.5, center=None, radius=0.2,
                  max_steps=None, on_step=None):
    """demos/FiniteVolume/advection_2d.cpp:61-155 (and advection_3d.cpp). Returns dict with init and final state."""
    dim = cfg.dim
    a = [1.0] * dim if a is None else list(a)
    center = [0.3] * dim if center is None else center
    bc = Bc("dirichlet", 0.0)
    mesh = Mesh.uniform(cfg)
    u = init_disc(mesh, center, radius)
    dt = cfl * cfg.cell_length(cfg.max_level)
    mesh, u = adapt(mesh, u, bc, eps, regularity)
    init_state = (mesh, u.copy())
    t, nt = 0.0, 0
    while t != Tf:
        mesh, u = adapt(mesh, u, bc, eps, regularity)
        t += dt
        if t > Tf:
            dt += Tf - t
            t = Tf
        update_ghost_mr(mesh, u, bc)
        u = fv_step(mesh, u, a, dt)
        nt += 1
        if on_step is not None:
            on_step(nt, mesh, u)
        if max_steps is not None and nt >= max_steps:
            break
    return dict(init=init_state, final=(mesh, u), steps=nt)


def heat_exact(x, t, K=1.0):
    """demos/FiniteVolume/heat.cpp:14-24."""
    r = np.ones(x.shape[0])
    for d in range(x.shape[1]):
        r = r * (1 / (2 * math.sqrt(math.pi * K * t)) * np.exp(-x[:, d] * x[:, d] / (4 * K * t)))
    return r


def run_heat(cfg: MeshConfig, Tf=0.1, K=1.0, cfl=0.95, eps=1e-4, regularity=1.0, max_steps=None, on_step=None):
    """demos/FiniteVolume/heat.cpp:112-236 with --explicit --init-sol=dirac: u0 = exact(t0 = 1e-2), Neumann(0),
    dt = cfl dx^2 / (2^dim K), per step MRadaptation then unp1 = u - dt * diff(u) (diff = make_diffusion_order2)."""
    dim = cfg.dim
    bc = Bc("neumann", 0.0)
    mesh = Mesh.uniform(cfg)
    t = 1e-2
    u = np.zeros(mesh.nref)
    L = cfg.max_level
    u[mesh.index(L, mesh.cells[L])] = heat_exact(mesh.cell_centers(L, mesh.cells[L]), t, K)
    dx = cfg.cell_length(L)
    dt = cfl * (dx * dx) / (pow(2, dim) * K)
    coeffs = diffusion_order2_coeffs([K] * dim)
    mesh, u = adapt(mesh, u, bc, eps, regularity)
    init_state = (mesh, u.copy())
    nt = 0
    while t != Tf:
        t += dt
        if t > Tf:
            dt += Tf - t
            t = Tf
        mesh, u = adapt(mesh, u, bc, eps, regularity)
        update_ghost_mr(mesh, u, bc)
        rhs = flux_linhom_apply(mesh, u, coeffs)
        # unp1 = u - dt * diff(u)  (xtensor: dt * rhs evaluated per element, then subtracted)
        unp1 = np.full(mesh.nref, np.nan)
        for l in mesh.leaf_levels():
            i = mesh.index(l, mesh.cells[l])
            unp1[i] = u[i] - dt * rhs[i]
        u = unp1
        nt += 1
        if on_step is not None:
            on_step(nt, mesh, u)
        if max_steps is not None and nt >= max_steps:
            break
    return dict(init=init_state, final=(mesh, u), steps=nt)


def run_linear_convection(cfg: MeshConfig, Tf=0.1, cfl=0.95, eps=1e-4, regularity=1.0, velocity=None, max_steps=None, on_step=None):
    """demos/FiniteVolume/linear_convection.cpp, explicit branch (:84-222): fully periodic box, u0 = 1 in [-0.8, -0.3] x [0.3, 0.8]
    (1D: [-0.8, -0.3]) else 0, velocity (1, -1), conv = make_convection_weno5 (six-cell line stencil: max_stencil_size(6), ghost
    width 3), dt = cfl * dx / sum|v|; per step MRadaptation then TVD-RK3:
        u1 = u - dt*conv(u);  u2 = 3/4 u + 1/4 (u1 - dt*conv(u1));  unp1 = 1/3 u + 2/3 (u2 - dt*conv(u2))."""
    dim = cfg.dim
    if velocity is None:
        velocity = [1.0] * dim
        if dim == 2:
            velocity[1] = -1.0
    bc = Bc("neumann", 0.0)  # never applied: no boundary on a fully periodic mesh
    mesh = Mesh.uniform(cfg)
    L = cfg.max_level
    c = mesh.cell_centers(L, mesh.cells[L])
    inside = (c[:, 0] >= -0.8) & (c[:, 0] <= -0.3)
    if dim == 2:
        inside &= (c[:, 1] >= 0.3) & (c[:, 1] <= 0.8)
    u = np.zeros(mesh.nref)
    u[mesh.index(L, mesh.cells[L])] = np.where(inside, 1.0, 0.0)
    dt = cfl * cfg.cell_length(L) / sum(abs(v) for v in velocity)
    flux = weno5_flux(velocity)

    def conv(mesh, f):
        update_ghost_mr(mesh, f, bc)
        return flux_nonlin_apply(mesh, f, flux, WENO5_OFFSETS)

    def leaves_expr(mesh, fn):
        out = np.full(mesh.nref, np.nan)
        for l in mesh.leaf_levels():
            i = mesh.index(l, mesh.cells[l])
            out[i] = fn(i)
        return out

    mesh, u = adapt(mesh, u, bc, eps, regularity)
    init_state = (mesh, u.copy())
    t, nt = 0.0, 0
    while t != Tf:
        t += dt
        if t > Tf:
            dt += Tf - t
            t = Tf
        mesh, u = adapt(mesh, u, bc, eps, regularity)
        c0 = conv(mesh, u)
        u1 = leaves_expr(mesh, lambda i: u[i] - dt * c0[i])
        c1 = conv(mesh, u1)
        u2 = leaves_expr(mesh, lambda i: 3. / 4 * u[i] + 1. / 4 * (u1[i] - dt * c1[i]))
        c2 = conv(mesh, u2)
        u = leaves_expr(mesh, lambda i: 1. / 3 * u[i] + 2. / 3 * (u2[i] - dt * c2[i]))
        nt += 1
        if on_step is not None:
            on_step(nt, mesh, u)
        if max_steps is not None and nt >= max_steps:
            break
    return dict(init=init_state, final=(mesh, u), steps=nt)
