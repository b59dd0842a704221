"""ctypes wrapper of oracle/_build/libsamurai_cpu.so (oracle/cpu_path.cpp): the compiled, all-cores CPU execution of the
per-step path.  TEST / BENCH INFRASTRUCTURE: only tests/, bench.py's CPU arms and __graft_entry__.build() touch it."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libsamurai_cpu.so")
_lib = None

DIRICHLET, NEUMANN = 0, 1


def build(force=False):
    """make -C oracle (g++ -O3 -march=x86-64-v3 -ffp-contract=off -fopenmp)."""
    if force and os.path.exists(LIB):
        os.remove(LIB)
    subprocess.run(["make", "-C", HERE, "-s"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB):
        build()
    L = ctypes.CDLL(LIB)
    vp, i32, i64, f64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double
    L.cpu_sim_create.restype = vp
    L.cpu_sim_create.argtypes = [i32, i32, i32, i32, i32, f64, f64, i32, f64]
    L.cpu_sim_create_from_leaves.restype = vp
    L.cpu_sim_create_from_leaves.argtypes = [i32, i32, i32, i32, f64, f64, i32, f64, vp, i64, vp]
    L.cpu_sim_destroy.argtypes = [vp]
    L.cpu_sim_threads.restype = i32
    L.cpu_sim_set_threads.restype = i32
    L.cpu_sim_set_threads.argtypes = [i32]
    L.cpu_sim_init_ball.argtypes = [vp, vp, f64, f64, f64]
    L.cpu_sim_adapt.argtypes = [vp]
    L.cpu_sim_update_ghost.argtypes = [vp]
    L.cpu_sim_steps.argtypes = [vp, i32, vp, f64, vp]
    L.cpu_sim_nb_cells.restype = i64
    L.cpu_sim_nb_cells.argtypes = [vp, i32]
    L.cpu_sim_nb_leaf_intervals.restype = i64
    L.cpu_sim_nb_leaf_intervals.argtypes = [vp]
    L.cpu_sim_get_leaves.argtypes = [vp, vp, vp]
    L.cpu_sim_get_field.argtypes = [vp, vp]
    L.cpu_sim_times.argtypes = [vp, vp, vp, vp, i32]
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class CpuSim:
    """The demo loop of demos/FiniteVolume/advection_{2,3}d.cpp on host cores."""

    def __init__(self, dim, min_level, max_level, pred_radius=1, eps=2e-4, regularity=1.0, bc=DIRICHLET, bc_value=0.0, start_level=None,
                 leaves=None, leaf_values=None):
        L = load()
        self.L = L
        self.dim, self.min_level, self.max_level = dim, min_level, max_level
        if leaves is not None:
            iv = np.ascontiguousarray(leaves, dtype=np.int32)
            vals = np.ascontiguousarray(leaf_values, dtype=np.float64)
            self.h = L.cpu_sim_create_from_leaves(dim, min_level, max_level, pred_radius, eps, regularity, bc, bc_value, _ptr(iv), iv.shape[0], _ptr(vals))
        else:
            self.h = L.cpu_sim_create(dim, min_level, max_level, pred_radius, max_level if start_level is None else start_level, eps, regularity,
                                      bc, bc_value)
        if not self.h:
            raise RuntimeError("cpu_sim_create failed")

    def close(self):
        if self.h:
            self.L.cpu_sim_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    @staticmethod
    def threads():
        return load().cpu_sim_threads()

    @staticmethod
    def set_threads(n):
        """returns the previous setting"""
        return load().cpu_sim_set_threads(int(n))

    def _ok(self, rc):
        if rc != 0:
            raise RuntimeError("cpu path failed (see stderr)")

    def init_ball(self, center, radius, inside=1.0, outside=0.0):
        c = np.zeros(3)
        c[: len(center)] = center
        self._ok(self.L.cpu_sim_init_ball(self.h, _ptr(c), radius, inside, outside))

    def adapt(self):
        self._ok(self.L.cpu_sim_adapt(self.h))

    def update_ghost(self):
        self._ok(self.L.cpu_sim_update_ghost(self.h))

    def steps(self, n, a, dt):
        av = np.zeros(3)
        av[: len(a)] = a
        out = ctypes.c_int64(0)
        self._ok(self.L.cpu_sim_steps(self.h, n, _ptr(av), dt, ctypes.byref(out)))
        return out.value

    def nb_cells(self, reference=False):
        return self.L.cpu_sim_nb_cells(self.h, 1 if reference else 0)

    def leaves(self):
        """(intervals [n, 5] = level, y, z, xs, xe ; leaf values in for_each_cell order)"""
        n = self.L.cpu_sim_nb_leaf_intervals(self.h)
        iv = np.zeros((n, 5), np.int32)
        vals = np.zeros(self.nb_cells(), np.float64)
        self._ok(self.L.cpu_sim_get_leaves(self.h, _ptr(iv), _ptr(vals)))
        return iv, vals

    def field(self):
        out = np.zeros(self.nb_cells(True), np.float64)
        self.L.cpu_sim_get_field(self.h, _ptr(out))
        return out

    def times(self, reset=False):
        a, b, c = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        self.L.cpu_sim_times(self.h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), 1 if reset else 0)
        return {"host_mesh_s": a.value, "host_batches_s": b.value, "fp_s": c.value}

    def cell_length(self, level):
        return 1.0 / (1 << level)
