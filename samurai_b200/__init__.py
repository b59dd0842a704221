"""samurai_b200 -- Python binding (ctypes) of the C ABI in include/samurai_b200.h.

The names mirror samurai's C++ API for the hot path (MRMesh / make_scalar_field / make_bc / make_MRAdapt /
update_ghost_mr / upwind), so the parity tests read like the reference's demos (demos/FiniteVolume/advection_2d.cpp).
There is no CPU fallback: compute calls raise when the CUDA library or a device is missing.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsamurai_b200.so")

CELLS, CELLS_AND_GHOSTS, PROJ_CELLS, UNION_CELLS, REFERENCE = range(5)
DIRICHLET, NEUMANN = 0, 1
KEEP, COARSEN, REFINE = 1, 2, 4


class SamuraiError(RuntimeError):
    pass


class MeshConfigC(C.Structure):
    _fields_ = [
        ("dim", C.c_int32),
        ("min_level", C.c_int32),
        ("max_level", C.c_int32),
        ("pred_radius", C.c_int32),
        ("max_stencil_radius", C.c_int32),
        ("graduation_width", C.c_int32),
        ("n_cells0", C.c_int32 * 3),
        ("origin", C.c_double * 3),
        ("scaling_factor", C.c_double),
        ("periodic", C.c_int32 * 3),
        ("refine_boundary", C.c_int32),
    ]


class IntervalC(C.Structure):
    _fields_ = [("y", C.c_int32), ("z", C.c_int32), ("start", C.c_int32), ("end", C.c_int32), ("offset", C.c_int64)]


INTERVAL_DTYPE = np.dtype([("y", "<i4"), ("z", "<i4"), ("start", "<i4"), ("end", "<i4"), ("offset", "<i8")])


class StatsC(C.Structure):
    _fields_ = [
        ("kernel_launches", C.c_uint64),
        ("h2d_bytes", C.c_uint64),
        ("d2h_bytes", C.c_uint64),
        ("host_mesh_seconds", C.c_double),
        ("host_batch_seconds", C.c_double),
        ("mesh_rebuilds", C.c_uint64),
        ("device_seconds", C.c_double),
        ("ghost_updates_skipped", C.c_uint64),
        ("harten_iterations", C.c_uint64),
        ("host_stage_seconds", C.c_double * 8),
    ]


# every symbol include/samurai_b200.h declares: (name, argtypes)
_u64, _i64, _i32, _dbl, _vp = C.c_uint64, C.c_int64, C.c_int, C.c_double, C.c_void_p
_P = C.POINTER
SYMBOLS = {
    "smr_init": [_i32],
    "smr_finalize": [],
    "smr_last_error": [],
    "smr_device_available": [],
    "smr_set_stream": [_vp],
    "smr_synchronize": [],
    "smr_mesh_create_uniform": [_P(MeshConfigC), _i32, _P(_u64)],
    "smr_mesh_create_from_intervals": [_P(MeshConfigC), _vp, _vp, _i64, _P(_u64)],
    "smr_mesh_destroy": [_u64],
    "smr_mesh_config_get": [_u64, _P(MeshConfigC)],
    "smr_mesh_nb_cells": [_u64, _i32, _i32, _P(_i64)],
    "smr_mesh_nb_intervals": [_u64, _i32, _i32, _P(_i64)],
    "smr_mesh_get_intervals": [_u64, _i32, _i32, _vp],
    "smr_mesh_generation": [_u64, _P(_u64)],
    "smr_mesh_get_index": [_u64, _i32, _i32, _i32, _i32, _P(_i64)],
    "smr_mesh_update_from_tags": [_u64, _vp, _i64, _P(_i32)],
    "smr_field_create": [_u64, C.c_char_p, _P(_u64)],
    "smr_field_destroy": [_u64],
    "smr_field_resize": [_u64],
    "smr_field_fill": [_u64, _dbl],
    "smr_field_size": [_u64, _P(_i64)],
    "smr_field_upload": [_u64, _vp, _i64],
    "smr_field_download": [_u64, _vp, _i64],
    "smr_field_swap": [_u64, _u64],
    "smr_field_set_bc": [_u64, _i32, _dbl],
    "smr_update_ghost_mr": [_u64],
    "smr_fv_upwind": [_u64, _u64, _vp, _dbl],
    "smr_fv_upwind_burgers": [_u64, _u64, _vp, _dbl],
    "smr_scheme_apply": [_u64, _u64, _i32, _vp, _dbl],
    "smr_scheme_apply_vector": [_vp, _vp, _i32, _i32, _vp, _dbl],
    "smr_field_lincomb": [_u64, _dbl, _u64, _dbl, _u64],
    "smr_adapt": [_vp, _i32, _dbl, _dbl, _P(_i32)],
    "smr_adapt_ex": [_vp, _i32, _dbl, _dbl, _i32, _P(_i32)],
    "smr_adapt_iteration": [_vp, _i32, _dbl, _dbl, _i32, _P(_i32)],
    "smr_adapt_last_size": [_u64, _P(_i64)],
    "smr_adapt_last_tags": [_u64, _vp, _i64],
    "smr_adapt_last_detail": [_u64, _vp, _i64],
    "smr_stats_get": [_P(StatsC)],
    "smr_stats_reset": [],
    "smr_mg_init": [_i32, _i32, _u64],
    "smr_mg_get_handle": [_vp],
    "smr_mg_connect": [_vp],
    "smr_mg_broadcast": [_u64],
    "smr_mg_rebalance": [_vp, _i32],
    "smr_mg_leaf_owners": [_u64, _vp, _i64],
    "smr_debug_host_rebuild": [_u64, _i32, _P(_dbl), _P(_dbl), _P(_i64)],
    "smr_debug_flux_records": [_u64, _vp, _i64, _P(_i64)],
    "smr_debug_fluxw_apply": [_u64, _vp, _i32, _i32, _vp, _dbl, _vp],
    "smr_profile_enable": [_i32],
    "smr_profile_get": [_i32, _P(_u64), _P(_dbl), _P(_u64)],
    "smr_profile_get_bytes": [_i32, _P(_u64)],
    "smr_set_fused": [_i32],
    "smr_field_init_ball": [_u64, _vp, _dbl, _dbl, _dbl, _i32],
}

FAMILIES = ["fv", "projection", "prediction", "detail", "criteria", "maximum", "bc", "copy", "keep", "init", "wavefront"]

_lib = None
_initialized = None


def load_library():
    """Load libsamurai_b200.so (built in-tree by __graft_entry__.build()). Fails loudly when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SamuraiError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback exists)")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_char_p if name == "smr_last_error" else C.c_int
    _lib = lib
    return lib


def _check(rc):
    if rc != 0:
        msg = load_library().smr_last_error().decode()
        if rc == 2:
            raise IndexError(msg)  # std::out_of_range
        if rc == 1:
            raise ValueError(msg)  # std::invalid_argument
        raise SamuraiError(msg)


def initialize(device=None):
    """samurai::initialize (samurai.hpp:22-98). device=None: cuda:0 when visible, otherwise host-only mode."""
    global _initialized
    lib = load_library()
    if device is None:
        rc = lib.smr_init(0)
        if rc != 0:
            _check(lib.smr_init(-1))
            _initialized = -1
            return False
        _initialized = 0
        return True
    _check(lib.smr_init(device))
    _initialized = device
    return device >= 0


def initialize_multi(rank, world, device=None, pool_bytes=8 << 30, exchange=None):
    """One process per GPU. `exchange(bytes) -> list[bytes]` all-gathers the 64-byte IPC handles (rank order);
    by default torch.distributed.all_gather_object is used. device < 0: host-only (partition logic without a GPU)."""
    lib = load_library()
    dev = rank if device is None else device
    have = initialize(dev)
    _check(lib.smr_mg_init(rank, world, pool_bytes if have else 0))
    if have and world > 1:
        buf = C.create_string_buffer(64)
        _check(lib.smr_mg_get_handle(buf))
        if exchange is None:
            import torch.distributed as dist

            def exchange(b):
                out = [None] * world
                dist.all_gather_object(out, b)
                return out

        allh = b"".join(exchange(buf.raw))
        assert len(allh) == 64 * world
        _check(lib.smr_mg_connect(allh))
    return have


def mg_broadcast(field):
    _check(load_library().smr_mg_broadcast(field._h))


def mg_rebalance(*fields):
    arr = (C.c_uint64 * len(fields))(*[f._h for f in fields])
    _check(load_library().smr_mg_rebalance(arr, len(fields)))


def mg_leaf_owners(mesh):
    n = mesh.nb_cells(CELLS)
    out = np.empty(n, dtype=np.int32)
    _check(load_library().smr_mg_leaf_owners(mesh._h, out.ctypes.data, n))
    return out


def finalize():
    _check(load_library().smr_finalize())


def device_available():
    return bool(load_library().smr_device_available())


def synchronize():
    _check(load_library().smr_synchronize())


def set_stream(ptr):
    _check(load_library().smr_set_stream(C.c_void_p(ptr)))


def stats(reset=False):
    s = StatsC()
    load_library().smr_stats_get(C.byref(s))
    if reset:
        load_library().smr_stats_reset()
    return {k: (list(getattr(s, k)) if k == "host_stage_seconds" else getattr(s, k)) for k, _ in StatsC._fields_}


def profile_enable(on=True):
    _check(load_library().smr_profile_enable(1 if on else 0))


def profile_get():
    """{family: (launches, seconds, cells)} accumulated since profile_enable()."""
    out = {}
    for i, name in enumerate(FAMILIES):
        n, s, c = C.c_uint64(), C.c_double(), C.c_uint64()
        _check(load_library().smr_profile_get(i, C.byref(n), C.byref(s), C.byref(c)))
        out[name] = (n.value, s.value, c.value)
    return out


def profile_bytes(family="wavefront"):
    """algorithmic bytes of the fused launches of a family accumulated since profile_enable()."""
    b = C.c_uint64()
    _check(load_library().smr_profile_get_bytes(FAMILIES.index(family), C.byref(b)))
    return b.value


def set_fused(on=True):
    """one cooperative launch per level wavefront (default) or one launch per sweep (same results)."""
    _check(load_library().smr_set_fused(1 if on else 0))


class mesh_config:
    """samurai::mesh_config<dim, prediction_stencil_radius> fluent builder (mesh_config.hpp:20-432)."""

    def __init__(self, dim, prediction_stencil_radius=1):
        self.dim = dim
        self.pred_radius = prediction_stencil_radius
        self._min_level = 0
        self._max_level = 6
        self._max_stencil_radius = 1
        self._graduation_width = 1
        self._disable_minimal_ghost_width = False
        self._periodic = [False, False, False]

    def min_level(self, v):
        self._min_level = v
        return self

    def max_level(self, v):
        self._max_level = v
        return self

    def max_stencil_size(self, size):
        self._max_stencil_radius = size // 2 + (size % 2)
        return self

    def max_stencil_radius(self, r):
        self._max_stencil_radius = r
        return self

    def graduation_width(self, w):
        self._graduation_width = w
        return self

    def disable_minimal_ghost_width(self):
        self._disable_minimal_ghost_width = True
        return self

    def refine_boundary(self, on=True):
        """args::refine_boundary (`--refine-boundary`, arguments.hpp:67): keep_boundary_refined after the criteria (mr/adapt.hpp:245-274)"""
        self._refine_boundary = bool(on)
        return self

    def periodic(self, *flags):
        """mesh_config::periodic(bool) / periodic(array) (mesh_config.hpp:171-196)"""
        if len(flags) == 1 and isinstance(flags[0], (list, tuple)):
            flags = tuple(flags[0])
        if len(flags) == 1:
            flags = flags * self.dim
        for d in range(self.dim):
            self._periodic[d] = bool(flags[d])
        return self

    def to_c(self, box_min, box_max):
        msr = self._max_stencil_radius
        if not self._disable_minimal_ghost_width:
            msr = max(msr, 2)  # mesh_config.hpp:388-393
        c = MeshConfigC()
        c.dim, c.min_level, c.max_level = self.dim, self._min_level, self._max_level
        c.pred_radius, c.max_stencil_radius, c.graduation_width = self.pred_radius, msr, self._graduation_width
        for d in range(3):
            c.periodic[d] = 1 if (d < self.dim and self._periodic[d]) else 0
        c.refine_boundary = 1 if getattr(self, "_refine_boundary", False) else 0
        lengths = [float(box_max[d]) - float(box_min[d]) for d in range(self.dim)]
        # approximate_box (box.hpp:280-360) for boxes whose lengths are integer multiples of the smallest one
        scaling = min(lengths)
        for d in range(3):
            if d < self.dim:
                n = lengths[d] / scaling
                if abs(n - round(n)) > 1e-12:
                    raise ValueError("box lengths must be integer multiples of the smallest length")
                c.n_cells0[d] = int(round(n))
                c.origin[d] = float(box_min[d])
            else:
                c.n_cells0[d] = 1
                c.origin[d] = 0.0
        c.scaling_factor = scaling
        return c


class MRMesh:
    """samurai::MRMesh (mr/mesh.hpp:74-122); make_mesh(box, cfg) starts uniform at max_level (mr/mesh.hpp:510-518)."""

    def __init__(self, handle, cfg_c):
        self._h = handle
        self.cfg = cfg_c

    @staticmethod
    def make_mesh(box_min, box_max, config: mesh_config, start_level=None):
        lib = load_library()
        c = config.to_c(box_min, box_max)
        h = C.c_uint64()
        _check(lib.smr_mesh_create_uniform(C.byref(c), c.max_level if start_level is None else start_level, C.byref(h)))
        return MRMesh(h.value, c)

    @staticmethod
    def from_intervals(box_min, box_max, config: mesh_config, levels, intervals):
        lib = load_library()
        c = config.to_c(box_min, box_max)
        levels = np.ascontiguousarray(levels, dtype=np.int32)
        intervals = np.ascontiguousarray(intervals, dtype=INTERVAL_DTYPE)
        h = C.c_uint64()
        _check(lib.smr_mesh_create_from_intervals(C.byref(c), levels.ctypes.data, intervals.ctypes.data, len(levels), C.byref(h)))
        return MRMesh(h.value, c)

    @property
    def dim(self):
        return self.cfg.dim

    def min_level(self):
        return self.cfg.min_level

    def max_level(self):
        return self.cfg.max_level

    def cell_length(self, level):
        return self.cfg.scaling_factor / (1 << level)

    def min_cell_length(self):
        return self.cell_length(self.cfg.max_level)

    def nb_cells(self, mesh_id=CELLS, level=-1):
        out = C.c_int64()
        _check(load_library().smr_mesh_nb_cells(self._h, mesh_id, level, C.byref(out)))
        return out.value

    def generation(self):
        out = C.c_uint64()
        _check(load_library().smr_mesh_generation(self._h, C.byref(out)))
        return out.value

    def intervals(self, mesh_id, level):
        lib = load_library()
        n = C.c_int64()
        _check(lib.smr_mesh_nb_intervals(self._h, mesh_id, level, C.byref(n)))
        out = np.zeros(n.value, dtype=INTERVAL_DTYPE)
        if n.value:
            _check(lib.smr_mesh_get_intervals(self._h, mesh_id, level, out.ctypes.data))
        return out

    def get_index(self, level, i, j=0, k=0):
        out = C.c_int64()
        _check(load_library().smr_mesh_get_index(self._h, level, i, j, k, C.byref(out)))
        return out.value

    def debug_host_rebuild(self, reps=3):
        a, b, n = C.c_double(), C.c_double(), C.c_int64()
        _check(load_library().smr_debug_host_rebuild(self._h, reps, C.byref(a), C.byref(b), C.byref(n)))
        return a.value, b.value, n.value

    def debug_flux_records(self):
        """[N, 6] int32: level, x, y, z, n, face kinds of every record of the flux-scheme batch (host only)."""
        n = C.c_int64()
        _check(load_library().smr_debug_flux_records(self._h, None, 0, C.byref(n)))
        out = np.zeros((n.value, 6), dtype=np.int32)
        if n.value:
            _check(load_library().smr_debug_flux_records(self._h, out.ctypes.data, n.value, C.byref(n)))
        return out

    def debug_fluxw_apply(self, u, velocity=None, scale=1.0):
        """Host evaluation of the WENO5 flux records (tests only; ghosts of `u` must be up to date).  `u`: one array, or a list of
        component arrays; without a velocity the non-linear form make_convection_weno5<Field>()."""
        comps = list(u) if isinstance(u, (list, tuple)) else [u]
        soa = np.ascontiguousarray(np.stack([np.asarray(c, dtype=np.float64) for c in comps]))
        kind = CONVECTION_WENO5 if velocity is not None else CONVECTION_WENO5_NONLINEAR
        v = np.ascontiguousarray(list(velocity if velocity is not None else []) + [0.0] * (3 - len(velocity if velocity is not None else [])),
                                 dtype=np.float64)
        out = np.zeros_like(soa)
        _check(load_library().smr_debug_fluxw_apply(self._h, soa.ctypes.data, len(comps), kind, v.ctypes.data, float(scale), out.ctypes.data))
        return [out[c] for c in range(len(comps))] if isinstance(u, (list, tuple)) else out[0]

    def update_from_tags(self, tags):
        tags = np.ascontiguousarray(tags, dtype=np.uint8)
        unchanged = C.c_int()
        _check(load_library().smr_mesh_update_from_tags(self._h, tags.ctypes.data, tags.size, C.byref(unchanged)))
        return bool(unchanged.value)

    def cell_table(self, mesh_id=CELLS):
        """(level, coords[N,dim], storage offset) per cell in for_each_cell order."""
        lv, co, off = [], [], []
        for level in range(self.cfg.max_level + 3):
            iv = self.intervals(mesh_id, level)
            if iv.size == 0:
                continue
            n = (iv["end"] - iv["start"]).astype(np.int64)
            tot = int(n.sum())
            rep = np.repeat(np.arange(iv.size), n)
            k = np.arange(tot) - np.repeat(np.cumsum(n) - n, n)
            x = iv["start"][rep] + k
            cols = [x, iv["y"][rep], iv["z"][rep]][: self.cfg.dim]
            co.append(np.stack(cols, axis=1).astype(np.int64))
            off.append(iv["offset"][rep] + k)
            lv.append(np.full(tot, level, dtype=np.int64))
        if not lv:
            return np.zeros(0, np.int64), np.zeros((0, self.cfg.dim), np.int64), np.zeros(0, np.int64)
        return np.concatenate(lv), np.concatenate(co), np.concatenate(off)

    def destroy(self):
        if self._h:
            _check(load_library().smr_mesh_destroy(self._h))
            self._h = 0


class ScalarField:
    """samurai::ScalarField<mesh_t, double> (field/scalar_field.hpp), device resident."""

    def __init__(self, name, mesh: MRMesh):
        h = C.c_uint64()
        _check(load_library().smr_field_create(mesh._h, name.encode(), C.byref(h)))
        self._h = h.value
        self.name = name
        self.mesh = mesh

    def size(self):
        out = C.c_int64()
        _check(load_library().smr_field_size(self._h, C.byref(out)))
        return out.value

    def resize(self):
        _check(load_library().smr_field_resize(self._h))

    def fill(self, v):
        _check(load_library().smr_field_fill(self._h, float(v)))

    def upload(self, host):
        host = np.ascontiguousarray(host, dtype=np.float64)
        _check(load_library().smr_field_upload(self._h, host.ctypes.data, host.size))

    def download(self, out=None):
        n = self.size()
        if out is None:
            out = np.empty(n, dtype=np.float64)
        _check(load_library().smr_field_download(self._h, out.ctypes.data, n))
        return out

    def init_ball(self, center, radius, inside=1.0, outside=0.0, overwrite_outside=True):
        """The demos' init(): 1 inside the disc/ball, 0 outside, evaluated at the leaf centres on the device."""
        c = np.ascontiguousarray(list(center) + [0.0] * (3 - len(center)), dtype=np.float64)
        _check(load_library().smr_field_init_ball(self._h, c.ctypes.data, float(radius), float(inside), float(outside), int(overwrite_outside)))

    def destroy(self):
        if self._h:
            _check(load_library().smr_field_destroy(self._h))
            self._h = 0


def make_scalar_field(name, mesh):
    return ScalarField(name, mesh)


class VectorField:
    """samurai::VectorField<mesh_t, double, n_comp> (field/vector_field.hpp) on the device: one SoA array per component
    (BASELINE north_star), i.e. n_comp scalar device fields that are ghost-updated, adapted and advanced together.  The
    reference's host layout is AoS `[cell][comp]`; upload()/download() convert."""

    def __init__(self, name, mesh: MRMesh, n_comp):
        self.name, self.mesh, self.n_comp = name, mesh, n_comp
        self.components = [ScalarField(f"{name}_{c}", mesh) for c in range(n_comp)]

    def resize(self):
        for f in self.components:
            f.resize()

    def fill(self, v):
        for f in self.components:
            f.fill(v)

    def upload(self, aos):
        """aos: [nb_cells(reference), n_comp]"""
        aos = np.asarray(aos, dtype=np.float64)
        for c, f in enumerate(self.components):
            f.upload(np.ascontiguousarray(aos[:, c]))

    def download(self):
        return np.stack([f.download() for f in self.components], axis=1)

    def destroy(self):
        for f in self.components:
            f.destroy()


def make_vector_field(name, mesh, n_comp):
    return VectorField(name, mesh, n_comp)


def _scalars(field):
    return field.components if isinstance(field, VectorField) else [field]


def make_bc(field, kind, *values):
    """samurai::make_bc<Dirichlet<1>>(u, v...) / make_bc<Neumann<1>>(u, v...) (bc/bc.hpp:751-815): one constant per component."""
    comps = _scalars(field)
    if len(values) == 1 and len(comps) > 1:
        values = values * len(comps)
    if len(values) != len(comps):
        raise ValueError("one boundary value per component")
    for f, v in zip(comps, values):
        _check(load_library().smr_field_set_bc(f._h, kind, float(v)))


def swap(a, b):
    """std::swap(u.array(), unp1.array())."""
    for x, y in zip(_scalars(a), _scalars(b)):
        _check(load_library().smr_field_swap(x._h, y._h))


def update_ghost_mr(*fields):
    """update_ghost_mr(fields...) (algorithm/update_ghost_mr.hpp:260-270)"""
    for field in fields:
        for f in _scalars(field):
            _check(load_library().smr_update_ghost_mr(f._h))


def upwind_step(unp1, u, a, dt):
    """unp1 = u - dt * samurai::upwind(a, u) (every component of a vector field is transported by the same velocity)."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    for x, y in zip(_scalars(unp1), _scalars(u)):
        _check(load_library().smr_fv_upwind(x._h, y._h, a.ctypes.data, float(dt)))


def upwind_scalar_burgers_step(unp1: ScalarField, u: ScalarField, k, dt):
    """unp1 = u - dt * samurai::upwind_scalar_burgers(k, u)."""
    k = np.ascontiguousarray(k, dtype=np.float64)
    _check(load_library().smr_fv_upwind_burgers(unp1._h, u._h, k.ctypes.data, float(dt)))


CONVECTION_UPWIND, DIFFUSION_ORDER2, CONVECTION_UPWIND_NONLINEAR, CONVECTION_WENO5, CONVECTION_WENO5_NONLINEAR = 0, 1, 2, 3, 4


class FluxScheme:
    """A flux-based scheme object: `rhs = scheme(u)` / `scheme.apply(out, u)` (schemes/fv/FV_scheme.hpp:202-238);
    `scalar * scheme` scales the flux (flux_based/algebraic_operators.hpp:7-82)."""

    def __init__(self, kind, params, name, scale=1.0):
        self.kind, self.params, self.name, self.scale = kind, np.ascontiguousarray(params, dtype=np.float64), name, float(scale)

    def apply(self, out, u):
        if isinstance(u, VectorField):
            if self.kind in (CONVECTION_UPWIND_NONLINEAR, CONVECTION_WENO5_NONLINEAR):  # the schemes that couple the components (flux u(d) * u)
                n = u.n_comp
                oh = (C.c_uint64 * n)(*[f._h for f in out.components])
                uh = (C.c_uint64 * n)(*[f._h for f in u.components])
                _check(load_library().smr_scheme_apply_vector(oh, uh, n, self.kind, self.params.ctypes.data, self.scale))
            else:  # linear schemes act on every component separately
                for o, f in zip(out.components, u.components):
                    _check(load_library().smr_scheme_apply(o._h, f._h, self.kind, self.params.ctypes.data, self.scale))
            return
        _check(load_library().smr_scheme_apply(out._h, u._h, self.kind, self.params.ctypes.data, self.scale))

    def __rmul__(self, scalar):
        return FluxScheme(self.kind, self.params, f"{scalar} * {self.name}", self.scale * float(scalar))

    def __call__(self, u):
        if isinstance(u, VectorField):
            out = VectorField(f"{self.name}({u.name})", u.mesh, u.n_comp)
        else:
            out = ScalarField(f"{self.name}({u.name})", u.mesh)
        self.apply(out, u)
        return out


def make_convection_upwind(velocity=None):
    """samurai::make_convection_upwind<Field>(velocity) (schemes/fv/operators/convection_lin.hpp:15-89); without a velocity
    the non-linear Burgers form make_convection_upwind<Field>() (operators/convection_nonlin.hpp:24-76, scalar fields)."""
    if velocity is None:
        return FluxScheme(CONVECTION_UPWIND_NONLINEAR, [0.0, 0.0, 0.0], "convection(u)")
    return FluxScheme(CONVECTION_UPWIND, velocity, "convection")


def make_convection_weno5(velocity=None):
    """samurai::make_convection_weno5<Field>(velocity) (schemes/fv/operators/convection_lin.hpp:95-178): WENO5 (Jiang & Shu) linear
    convection, a non-linear flux scheme with a six-cell line stencil; fully periodic meshes with max_stencil_size(6).  Without a
    velocity the Burgers form make_convection_weno5<Field>() (operators/convection_nonlin.hpp:162-233; scalar or n_comp == dim)."""
    if velocity is None:
        return FluxScheme(CONVECTION_WENO5_NONLINEAR, [0.0, 0.0, 0.0], "convection(u)")
    v = list(velocity) + [0.0] * (3 - len(velocity))
    return FluxScheme(CONVECTION_WENO5, v, "convection")


def make_diffusion_order2(K):
    """samurai::make_diffusion_order2<Field>(K) (schemes/fv/operators/diffusion.hpp:123-175)."""
    return FluxScheme(DIFFUSION_ORDER2, K, "diffusion")


def lincomb(out: ScalarField, a, x: ScalarField, b, y: ScalarField):
    """out = a*x + b*y on the leaves; `unp1 = u - dt * scheme(u)` is lincomb(unp1, 1, u, -dt, rhs)."""
    _check(load_library().smr_field_lincomb(out._h, float(a), x._h, float(b), y._h))


class mra_config:
    """samurai::mra_config (mr/config.hpp:10-68)."""

    def __init__(self):
        self._eps, self._reg, self._rel = 1e-4, 1.0, False

    def relative_detail(self, v=True):
        self._rel = bool(v)
        return self

    def epsilon(self, v):
        self._eps = v
        return self

    def regularity(self, v):
        self._reg = v
        return self


class MRAdapt:
    """samurai::make_MRAdapt(fields...) (mr/adapt.hpp:391-397)."""

    def __init__(self, *fields):
        self.fields = [f for field in fields for f in _scalars(field)]  # a vector field contributes all its components
        fields = self.fields
        self._arr = (C.c_uint64 * len(fields))(*[f._h for f in fields])

    def __call__(self, cfg: mra_config):
        n = C.c_int()
        _check(load_library().smr_adapt_ex(self._arr, len(self.fields), cfg._eps, cfg._reg, int(cfg._rel), C.byref(n)))
        return n.value

    def iteration(self, cfg: mra_config, ite):
        unchanged = C.c_int()
        _check(load_library().smr_adapt_iteration(self._arr, len(self.fields), cfg._eps, cfg._reg, ite, C.byref(unchanged)))
        return bool(unchanged.value)

    def last_tags(self):
        mesh = self.fields[0].mesh
        n = C.c_int64()
        _check(load_library().smr_adapt_last_size(mesh._h, C.byref(n)))
        out = np.empty(n.value, dtype=np.uint8)
        _check(load_library().smr_adapt_last_tags(mesh._h, out.ctypes.data, out.size))
        return out

    def last_detail(self):
        mesh = self.fields[0].mesh
        n = C.c_int64()
        _check(load_library().smr_adapt_last_size(mesh._h, C.byref(n)))
        out = np.empty(n.value * len(self.fields), dtype=np.float64)
        _check(load_library().smr_adapt_last_detail(mesh._h, out.ctypes.data, out.size))
        return out


def make_MRAdapt(*fields):
    return MRAdapt(*fields)
