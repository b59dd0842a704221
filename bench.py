#!/usr/bin/env python
"""bench.py -- samurai hot path on B200: cell-updates/s per FV step incl. ghost update + MR detail/tagging.

Workload (BASELINE.json configs[1]): demos/FiniteVolume/advection_2d.cpp scaled to max_level=14 (min_level 4,
eps 2e-4, prediction radius 1, Dirichlet 0, disc r=0.2 at (0.3,0.3), a=(1,1), cfl 0.5). One "step" is one pass of
the demo's time loop:  MRadaptation(cfg) -> update_ghost_mr(u) -> unp1 = u - dt*upwind(a,u) -> swap.
One cell-update = one leaf cell advanced one step (nb_cells(mesh_id_t::cells), the reference benchmarks' counter).

  value   device-resident loop: fields live in HBM, K steps timed with CUDA events (includes the host-side mesh work
          the GPU waits for; the split device / host-mesh / host-batch is reported next to it)
  e2e     same K steps through the public API with HOST buffers: u uploaded from pinned memory and unp1 read back
          every step
  roofline  dominant kernel family of the timed loop (CUDA-event time per launch, measured in a separate profiled
          pass), algorithmic bytes from DESIGN.md; plus `uniform_sweep`: the FV kernel on a uniform level-13 mesh
          (BASELINE.json configs[4]) where the HBM roofline is the physically meaningful bound
  cpu_baseline  the numpy oracle ("port") on host cores, a bounded sample of the same adapted mesh

`--impl reference` times the oracle port alone (no GPU code on the path).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# torchrun pins OMP_NUM_THREADS=1; the host-side mesh/batch construction is OpenMP-parallel, so give every rank its share
_world = int(os.environ.get("WORLD_SIZE", 1))
if _world > 1:
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // _world))

METRIC = "cell-updates/s per FV step (incl. ghost+MR detail)"
UNIT = "cell-updates/s"

# algorithmic bytes per output cell of each kernel family (DESIGN.md "Kernels"), fp64 scalar field
def alg_bytes_per_cell(family, dim):
    nchild = 1 << dim
    return {
        "fv": 16.0,                          # read u, write unp1
        "projection": 8.0 * (nchild + 1),    # 2^dim children read, 1 coarse write
        "prediction": 8.0 * (1 + 1.0 / nchild),  # 1 fine write + its parent read once
        "detail": 8.0 * (1 + 2 * nchild),    # coarse read once, children read, details written
        "criteria": 8.0 * (nchild + 1) + 2.0 * nchild,  # detail reads + tag read/write
        "maximum": 2.0 * nchild + 2.0,       # tag bytes r/w
        "bc": 24.0,                          # 1 read + 1 write + item record amortised
        "copy": 16.0,
        "keep": 1.0,
        "init": 8.0,
    }[family]


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  NVML is read in-process
    (nvidia_ml_py, the library nvidia-smi itself uses): an external `nvidia-smi -lms 200` loop was measured to stall this
    process's CUDA calls by several ms per step.  Falls back to the nvidia-smi loop if NVML cannot be loaded."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, period=0.1):
        self.gpu = gpu_index
        self.period = period
        self.proc = None
        self.lines = []
        self.samples = []  # (sm_mhz, sm_max_mhz, reasons bitmask)
        self.nvml = None
        self._stop = threading.Event()
        self.t = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else self.gpu
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self._stop.is_set():
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
                rs = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((float(sm), float(mx), int(rs)))
            except Exception:
                pass
            self._stop.wait(self.period)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def wait_first_sample(self, timeout=8.0):
        """keep the sampler's start-up (NVML initialisation, ~1 s for nvidia-smi) out of the timed region"""
        t0 = time.perf_counter()
        while not self.lines and not self.samples and time.perf_counter() - t0 < timeout and (self.proc is not None or self.nvml is not None):
            time.sleep(0.02)

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.t.join(timeout=2)
            n = self.nvml
            try:  # one more reading right at the end of the timed region (short runs may fit between two periodic samples)
                rs = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)),
                                     float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)), int(rs)))
            except Exception:
                pass
            names = (("hw_slowdown", getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
                     ("hw_thermal_slowdown", getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
                     ("sw_thermal_slowdown", getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
                     ("sw_power_cap", getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)))
            reasons = sorted({name for _, _, rs in self.samples for name, bit in names if rs & bit})
            sm = [x[0] for x in self.samples]
            mx = [x[1] for x in self.samples]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                    "samples": len(sm), "source": "nvml in-process, %.0f ms period" % (1e3 * self.period)}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi -lms 200"}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ----------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the numpy oracle port
# ----------------------------------------------------------------------------------------------------------------------
def oracle_adapted_state(so, min_level, max_level, eps, start_level=8):
    """Adapted advection_2d state built by the oracle alone (bottom-up: uniform start_level, refine to max_level)."""
    cfg = so.MeshConfig(dim=2, min_level=min_level, max_level=max_level, pred_radius=1)
    bc = so.Bc("dirichlet", 0.0)
    mesh = so.Mesh.uniform(cfg, level=min(start_level, max_level))
    u = so.init_disc(mesh, [0.3, 0.3], 0.2)
    for _ in range(max_level - min_level + 2):
        before = mesh
        mesh, u = so.adapt(mesh, u, bc, eps, 1.0)
        # re-impose the exact initial condition on the refined leaves (sharp disc)
        u = so.init_disc(mesh, [0.3, 0.3], 0.2)
        if mesh is before:
            break
    return cfg, bc, mesh, u


def oracle_steps(so, cfg, bc, mesh, u, n_steps, dt):
    """n_steps of the demo loop on the oracle; returns (cell_updates, seconds, mesh, u)."""
    cells = 0
    t0 = time.perf_counter()
    for _ in range(n_steps):
        mesh, u = so.adapt(mesh, u, bc, ARGS.eps, 1.0)
        so.update_ghost_mr(mesh, u, bc)
        u = so.fv_step(mesh, u, [1.0] * cfg.dim, dt)
        cells += mesh.nb_cells()
    return cells, time.perf_counter() - t0, mesh, u


def workload_string(args):
    return (f"advection_{args.dim}d min_level={args.min_level} max_level={args.max_level} eps={args.eps} pred_radius=1 Dirichlet(0) "
            f"ball r=0.2@(0.3,..), a=(1,..), cfl={0.5 if args.dim == 2 else 0.25}; one step = MRadaptation + update_ghost_mr + upwind + swap")


def cpu_adapted_sim(args):
    """The workload's adapted start state built by the compiled CPU path alone (bottom-up: uniform coarse level, refine to
    max_level re-imposing the exact initial condition, like the demo's initial MRadaptation from the other side)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_path

    start = min(8 if args.dim == 2 else 5, args.max_level)
    sim = cpu_path.CpuSim(args.dim, args.min_level, args.max_level, 1, eps=args.eps, regularity=1.0, start_level=start)
    sim.init_ball([0.3] * args.dim, 0.2)
    for _ in range(args.max_level - args.min_level + 2):
        n0 = sim.nb_cells()
        sim.adapt()
        sim.init_ball([0.3] * args.dim, 0.2)
        if sim.nb_cells() == n0:
            break
    return cpu_path, sim


def run_reference(args):
    """The CPU arm: the reference library itself cannot be built in this image, so this times the
    compiled all-cores CPU path (oracle/cpu_path.cpp, `kind: port`, bit-identical to the numpy oracle that is pinned on the
    reference's golden files) on the SAME workload as the product arm, all host threads."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    os.environ.pop("OMP_NUM_THREADS", None)  # torchrun pins it to 1; this arm runs alone on rank 0 with every host core
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_path

    cpu_path.CpuSim.set_threads(os.cpu_count() or 1)
    cpu_path, sim = cpu_adapted_sim(args)
    a = [1.0] * args.dim
    dt = (0.5 if args.dim == 2 else 0.25) / (1 << args.max_level)
    sim.steps(args.warmup, a, dt)
    sim.times(reset=True)
    t0 = time.perf_counter()
    cells = sim.steps(args.steps, a, dt)
    secs = time.perf_counter() - t0
    value = cells / secs
    tm = sim.times()
    cores = cpu_path.CpuSim.threads()
    sample = (f"{args.steps} steps of the full workload ({sim.nb_cells()} leaves, {sim.nb_cells(True)} reference cells at the end); compiled C++/OpenMP port "
              f"(oracle/cpu_path.cpp, -O3 -march=x86-64-v3 -ffp-contract=off), {cores} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(args), "leaves": sim.nb_cells(), "reference_cells": sim.nb_cells(True)},
        "split_ms_per_step": {"fp_loops": 1e3 * tm["fp_s"] / args.steps, "host_mesh": 1e3 * tm["host_mesh_s"] / args.steps,
                              "host_batches": 1e3 * tm["host_batches_s"] / args.steps},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# product arm
# ----------------------------------------------------------------------------------------------------------------------
class Sim:
    """The advection_2d demo (demos/FiniteVolume/advection_2d.cpp:61-155) on the C ABI."""

    def __init__(self, sb, args, dim=None):
        dim = args.dim if dim is None else dim
        self.sb = sb
        self.dim = dim
        cfg = sb.mesh_config(dim, 1).min_level(args.min_level).max_level(args.max_level).max_stencil_size(2).disable_minimal_ghost_width()
        self.mesh = sb.MRMesh.make_mesh([0.0] * dim, [1.0] * dim, cfg)
        self.u = sb.make_scalar_field("u", self.mesh)
        self.u.resize()
        self.u.init_ball([0.3] * dim, 0.2)
        sb.make_bc(self.u, sb.DIRICHLET, 0.0)
        self.unp1 = sb.make_scalar_field("unp1", self.mesh)
        self.adapt = sb.make_MRAdapt(self.u)
        self.mra = sb.mra_config().epsilon(args.eps)
        self.a = [1.0] * dim
        self.dt = (0.5 if dim == 2 else 0.25) * self.mesh.min_cell_length()  # advection_2d.cpp:76 / advection_3d.cpp:76

    def step(self):
        sb = self.sb
        self.adapt(self.mra)
        sb.update_ghost_mr(self.u)
        self.unp1.resize()
        sb.upwind_step(self.unp1, self.u, self.a, self.dt)
        sb.swap(self.u, self.unp1)
        return self.mesh.nb_cells()


def uniform_sweep(sb, torch, level, iters=20):
    """BASELINE.json configs[4]: uniform-level upwind sweep, the shape where HBM bandwidth is the bound."""
    cfg = sb.mesh_config(2, 1).min_level(level).max_level(level).max_stencil_size(2).disable_minimal_ghost_width()
    mesh = sb.MRMesh.make_mesh([0.0, 0.0], [1.0, 1.0], cfg)
    u = sb.make_scalar_field("u", mesh)
    u.resize()
    u.fill(0.0)
    u.init_ball([0.3, 0.3], 0.2)
    sb.make_bc(u, sb.DIRICHLET, 0.0)
    v = sb.make_scalar_field("v", mesh)
    v.resize()
    sb.update_ghost_mr(u)
    n = mesh.nb_cells()
    dt = 0.5 * mesh.min_cell_length()
    for _ in range(3):
        sb.upwind_step(v, u, [1.0, 1.0], dt)
        sb.swap(u, v)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        sb.upwind_step(v, u, [1.0, 1.0], dt)
        sb.swap(u, v)
    e1.record()
    torch.cuda.synchronize()
    secs = e0.elapsed_time(e1) * 1e-3 / iters
    # heat/diffusion sweep (make_diffusion_order2, explicit): rhs = S(u) then unp1 = u - dt * rhs, the reference's unfused form
    rhs = sb.make_scalar_field("rhs", mesh)
    diff = sb.make_diffusion_order2([1.0, 1.0])
    for _ in range(3):
        diff.apply(rhs, u)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        diff.apply(rhs, u)
    e1.record()
    torch.cuda.synchronize()
    dsecs = e0.elapsed_time(e1) * 1e-3 / iters
    rhs.destroy()
    gsecs = None  # the ghost update of a uniform mesh is part of uniform_full_step (timing it here only hit the ghosts_updated skip path)
    u.destroy()
    v.destroy()
    mesh.destroy()
    return n, secs, gsecs, dsecs


def weno5_sweep(sb, torch, level, iters=10):
    """rhs = make_convection_weno5(velocity)(u) on a uniform, fully periodic 2D level-`level` mesh with max_stencil_size(6)
    (demos/FiniteVolume/linear_convection.cpp's operator, row f1): per cell four WENO5 fluxes of ~60 fp64 operations each without FMA
    contraction, so the fp64 pipe bounds it, not HBM (8 B zero-fill + 8 B read + 8 B write per cell)."""
    cfg = sb.mesh_config(2, 1).min_level(level).max_level(level).periodic([True, True]).max_stencil_size(6)
    mesh = sb.MRMesh.make_mesh([-1.0, -1.0], [1.0, 1.0], cfg)
    u = sb.make_scalar_field("u", mesh)
    u.resize()
    u.fill(0.0)
    u.init_ball([-0.5, 0.5], 0.3)
    rhs = sb.make_scalar_field("rhs", mesh)
    conv = sb.make_convection_weno5([1.0, -1.0])
    n = mesh.nb_cells()
    for _ in range(3):
        conv.apply(rhs, u)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        conv.apply(rhs, u)
    e1.record()
    torch.cuda.synchronize()
    secs = e0.elapsed_time(e1) * 1e-3 / iters
    for f in (rhs, u):
        f.destroy()
    mesh.destroy()
    return n, secs


def uniform_full_step(sb, torch, dim, level, iters=10):
    """The WHOLE step on a uniform level-`level` mesh with min_level = level - 1 (BASELINE.json configs[4] shapes): MRadaptation
    (zero fill, keep tags, ghost update = projection level -> level-1 + BC, detail, criteria, keep propagation, change flag),
    update_ghost_mr, unp1 = u - dt * upwind(a, u), swap.  epsilon < 0 makes every detail significant, so nothing coarsens and the
    mesh stays uniform while all the multiresolution work is done; the change flag stays clear, so no tag leaves the device.
    Returns the device time per step (CUDA events) and the algorithmic bytes the step moves."""
    cfg = sb.mesh_config(dim, 1).min_level(level - 1).max_level(level).max_stencil_size(2).disable_minimal_ghost_width()
    mesh = sb.MRMesh.make_mesh([0.0] * dim, [1.0] * dim, cfg)
    u = sb.make_scalar_field("u", mesh)
    u.resize()
    u.fill(0.0)
    u.init_ball([0.3] * dim, 0.2)
    sb.make_bc(u, sb.DIRICHLET, 0.0)
    v = sb.make_scalar_field("v", mesh)
    v.resize()
    adapt = sb.make_MRAdapt(u)
    mra = sb.mra_config().epsilon(-1.0)
    a = [1.0] * dim
    dt = (0.5 if dim == 2 else 0.25) * mesh.min_cell_length()

    def step():
        adapt(mra)
        sb.update_ghost_mr(u)
        sb.upwind_step(v, u, a, dt)
        sb.swap(u, v)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    n, nref = mesh.nb_cells(), mesh.nb_cells(sb.REFERENCE)
    assert n == (1 << (dim * level)), "the mesh must have stayed uniform"
    st0 = sb.stats(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    secs = e0.elapsed_time(e1) * 1e-3 / iters
    st = sb.stats()
    nc = 1 << dim
    # algorithmic bytes per step, N = leaves, nref = reference cells (levels L, L-1 and the L-2 prediction ghosts):
    parts = {
        "zero_fill_detail_tag": 9.0 * nref,              # mr/adapt.hpp:165-168
        "keep_tags": 1.0 * n,
        "projection": 8.0 * n + 8.0 * n / nc,            # ghost update: level -> level-1 (read children, write parents)
        "detail": 8.0 * n + 8.0 * n / nc + 8.0 * n,      # children + parents read once, details written
        "criteria": 8.0 * n + 8.0 * n / nc + 2.0 * n,    # details of children and parents, tags read + written
        "keep_propagation": 2.0 * n + 2.0 * n / nc,      # tags of children r/w, parents r/w
        "change_flag": 1.0 * n,
        "fv": 16.0 * n,
    }
    total = sum(parts.values())
    u.destroy()
    v.destroy()
    mesh.destroy()
    return {"dim": dim, "level": level, "cells": n, "reference_cells": nref, "ms_per_step": 1e3 * secs, "device_ms_per_step": 1e3 * st["device_seconds"] / iters,
            "algorithmic_bytes": total, "bytes_per_cell": total / n, "parts_bytes_per_cell": {k: x / n for k, x in parts.items()},
            "achieved_GBps": total / secs / 1e9, "cell_updates_per_s": n / secs, "launches_per_step": st["kernel_launches"] / iters,
            "d2h_bytes_per_step": st["d2h_bytes"] / iters, "harten_iterations_per_step": st["harten_iterations"] / iters}


def config4_3d(sb, torch, dist, args, world, barrier):
    """BASELINE.json configs[3]: advection_3d (levels 4-`--level-3d`, eps 2e-4, ball r=0.2 at 0.3, a=(1,1,1), cfl 0.25), the case
    the multi-GPU efficiency target is stated on.  Same loop and same timing rules as the headline; every rank calls this."""
    class A3:
        pass

    a3 = A3()
    a3.dim, a3.min_level, a3.max_level, a3.eps = 3, args.min_level, args.level_3d, args.eps
    t0 = time.perf_counter()
    sim = Sim(sb, a3)
    sim.adapt(sim.mra)
    sb.synchronize()
    init_secs = time.perf_counter() - t0
    if world > 1:
        sb.mg_rebalance(sim.u)
    for _ in range(3):
        sim.step()
    if world > 1:
        sb.mg_rebalance(sim.u)
    sb.stats(reset=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    cells = 0
    steps = args.steps_3d
    for _ in range(steps):
        cells += sim.step()
    e1.record()
    barrier()
    secs = e0.elapsed_time(e1) * 1e-3
    st = sb.stats()
    if world > 1:
        t = torch.tensor([secs], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        secs = t.item()
    out = {"workload": f"advection_3d min_level={a3.min_level} max_level={a3.max_level} eps={a3.eps} pred_radius=1 Dirichlet(0) ball r=0.2@(0.3,..), "
                       f"a=(1,1,1), cfl=0.25; one step = MRadaptation + update_ghost_mr + upwind + swap",
           "n_gpus": world, "steps": steps, "value": cells / secs, "unit": UNIT, "ms_per_step": 1e3 * secs / steps,
           "leaves": sim.mesh.nb_cells(), "reference_cells": sim.mesh.nb_cells(sb.REFERENCE), "initial_adaptation_s": init_secs,
           "split_ms_per_step": {"device": 1e3 * st["device_seconds"] / steps, "host_mesh": 1e3 * st["host_mesh_seconds"] / steps,
                                 "host_batches": 1e3 * st["host_batch_seconds"] / steps},
           "gpu_launches_per_step": st["kernel_launches"] / steps}
    sim.u.destroy()
    sim.unp1.destroy()
    sim.mesh.destroy()
    return out


def run_product(args):
    import torch
    import torch.distributed as dist

    rank, local_rank, world = dist_env()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    import __graft_entry__

    if not os.path.exists(__graft_entry__.LIB):
        __graft_entry__.build()
    import samurai_b200 as sb

    if world > 1:
        # every field / detail / tag buffer lives in a per-rank pool mapped into the peers: size it for the uniform
        # max_level start mesh (reference cells ~ 4/3 * 4^L; u, its transfer twin, unp1, detail, tags) with slack
        nref0 = int((2 ** args.dim) ** args.max_level * (1.35 if args.dim == 2 else 1.16))
        if args.level_3d > 0:
            nref0 = max(nref0, int(8 ** args.level_3d * 1.16))
        pool = int(nref0 * (8 * 4 + 1) * 1.6) + (1 << 28)
        ok = sb.initialize_multi(rank, world, device=local_rank, pool_bytes=pool)
    else:
        ok = sb.initialize(local_rank)
    if not ok:
        raise SystemExit("no CUDA device: bench.py has no CPU fallback for the product arm")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_file):
        peak_gbs, peak_src = json.load(open(peaks_file))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak_gbs, peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    t0 = time.perf_counter()
    sim = Sim(sb, args)
    sim.adapt(sim.mra)  # the demo's initial MRadaptation (from the uniform max_level mesh)
    sb.synchronize()
    init_secs = time.perf_counter() - t0
    if world > 1:
        sb.mg_rebalance(sim.u)  # the start mesh was cut into equal slabs; re-cut for the adapted leaves

    for _ in range(args.warmup):
        sim.step()
    if world > 1:
        sb.mg_rebalance(sim.u)

    # ---- timed, device-resident ------------------------------------------------------------------------------------
    sampler = ClockSampler(local_rank, period=args.clock_period)
    sampler.start()
    sampler.wait_first_sample()
    sb.stats(reset=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    cells = 0
    for _ in range(args.steps):
        cells += sim.step()
    e1.record()
    barrier()
    secs = e0.elapsed_time(e1) * 1e-3
    st = sb.stats()
    clocks = sampler.stop()
    leaves_now, ref_now = sim.mesh.nb_cells(), sim.mesh.nb_cells(sb.REFERENCE)

    # ---- e2e: host buffers in, host buffers out, every step --------------------------------------------------------
    if world > 1:
        sb.mg_broadcast(sim.u)
    host = sim.u.download()
    pinned = torch.empty(int(ref_now * 1.5) + 1024, dtype=torch.float64).pin_memory().numpy()
    pinned[: host.size] = host
    n_host = host.size
    sb.stats(reset=True)
    barrier()
    e0.record()
    e2e_cells = 0
    for _ in range(args.steps):
        sim.u.upload(pinned[:n_host])
        e2e_cells += sim.step()
        n_host = sim.u.size()
        if n_host > pinned.size:
            pinned = torch.empty(int(n_host * 1.5), dtype=torch.float64).pin_memory().numpy()
        if world > 1:
            sb.mg_broadcast(sim.u)  # the host copy must be the complete field on every rank
        sim.u.download(pinned[:n_host])
    e1.record()
    barrier()
    e2e_secs = e0.elapsed_time(e1) * 1e-3
    st_e2e = sb.stats()

    # ---- max over ranks --------------------------------------------------------------------------------------------
    if world > 1:
        t = torch.tensor([secs, e2e_secs], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        secs, e2e_secs = t.tolist()  # cells / e2e_cells already count the GLOBAL leaves (every rank holds the whole mesh)

    # ---- parity against the CPU path + cpu baseline (collective: every rank steps the product) -------------------------------
    cpu = parity = None
    if not args.no_cpu_baseline:
        cpu, parity = parity_and_cpu_baseline(sb, sim, args, rank, world, barrier)

    # ---- dominant kernel of the loop, measured live (collective at N > 1: every rank launches the same sequence) ----------------
    mg_prof = None
    if world > 1:
        sb.profile_enable(True)
        for _ in range(3):
            sim.step()
        prof = sb.profile_get()
        wf_bytes = sb.profile_bytes("wavefront")
        sb.profile_enable(False)
        fam_time = {k: v[1] for k, v in prof.items() if v[0]}
        if fam_time:
            total_prof = sum(fam_time.values())
            dom = max(fam_time, key=fam_time.get)
            n_l, s_l, c_l = prof[dom]
            dom_bytes = wf_bytes if dom == "wavefront" else alg_bytes_per_cell(dom, args.dim) * c_l
            mg_prof = {"bound": "hbm", "kernel": dom, "achieved": dom_bytes / s_l / 1e9, "peak": peak_gbs, "unit": "GB/s", "frac": dom_bytes / s_l / 1e9 / peak_gbs,
                       "traffic": None, "peak_source": peak_src, "share_of_device_time": fam_time[dom] / total_prof, "launches": n_l,
                       "us_per_launch": 1e6 * s_l / n_l, "algorithmic_bytes_per_launch": dom_bytes / n_l,
                       "note": "rank 0's share of the phases (its slab); the launch spans the cross-GPU phase barriers, so its duration includes "
                               "waiting for the slowest rank; latency bound like the single-GPU launch"}

    # ---- BASELINE configs[3]: the 3D case (every rank) -----------------------------------------------------------------------
    c4 = None
    if args.level_3d > 0:
        c4 = config4_3d(sb, torch, dist if world > 1 else None, args, world, barrier)

    line = None
    if rank == 0 and world > 1:
        line = {
            "metric": METRIC, "value": cells / secs, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_string(args),
                       "leaves": leaves_now, "reference_cells": ref_now,
                       "parallelism": f"{world} leaf-balanced slabs (one per GPU), halo values stored into the peers by the producing kernels over "
                                      f"NVLink (CUDA IPC); one fused cooperative launch per ghost update / harten iteration on every GPU, its phase "
                                      f"barrier exchanging flags with the peers; tags replicated; host mesh work replicated on every rank "
                                      f"({max(1, (os.cpu_count() or 1) // world)} host threads per rank); same global problem as N=1"},
            "split_ms_per_step": {"device": 1e3 * st["device_seconds"] / args.steps, "host_mesh": 1e3 * st["host_mesh_seconds"] / args.steps,
                                  "host_batches": 1e3 * st["host_batch_seconds"] / args.steps},
            "gpu_launches": int(st["kernel_launches"]),
            "initial_adaptation_s": init_secs,
            "e2e": {"value": e2e_cells / e2e_secs, "unit": UNIT, "h2d_bytes_per_step": int(st_e2e["h2d_bytes"] / args.steps),
                    "d2h_bytes_per_step": int(st_e2e["d2h_bytes"] / args.steps), "ms_per_step": 1e3 * e2e_secs / args.steps},
            "roofline": mg_prof, "cpu_baseline": cpu, "parity": parity, "config4_3d": c4, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    elif rank == 0:
        # ---- per-family profile pass (separate from the timed region: every launch is synchronised) ----------------
        # (a) as timed: the fused level wavefront (one cooperative launch per ghost update / harten iteration / transfer)
        sb.profile_enable(True)
        for _ in range(4):
            sim.step()
        prof = sb.profile_get()
        wf_bytes = sb.profile_bytes("wavefront")
        sb.profile_enable(False)
        fam_time = {k: v[1] for k, v in prof.items() if v[0]}
        total_prof = sum(fam_time.values()) or 1.0
        dom = max(fam_time, key=fam_time.get)
        n_l, s_l, c_l = prof[dom]
        dom_bytes = wf_bytes if dom == "wavefront" else alg_bytes_per_cell(dom, args.dim) * c_l
        achieved = dom_bytes / s_l / 1e9
        fused = {k: {"launches": v[0], "us_per_launch": 1e6 * v[1] / v[0], "units_per_launch": v[2] / v[0], "share": v[1] / total_prof,
                     "GBps": (wf_bytes if k == "wavefront" else alg_bytes_per_cell(k, args.dim) * v[2]) / v[1] / 1e9} for k, v in prof.items() if v[0]}
        # (b) the same steps with one launch per sweep, to see the kernel families separately
        sb.set_fused(False)
        sb.profile_enable(True)
        for _ in range(2):
            sim.step()
        prof_u = sb.profile_get()
        sb.profile_enable(False)
        sb.set_fused(True)
        tot_u = sum(v[1] for v in prof_u.values() if v[0]) or 1.0
        families = {k: {"launches": v[0], "us_per_launch": 1e6 * v[1] / v[0], "cells_per_launch": v[2] / v[0],
                        "share": v[1] / tot_u, "GBps": alg_bytes_per_cell(k, args.dim) * v[2] / v[1] / 1e9} for k, v in prof_u.items() if v[0]}

        # ---- uniform sweep (configs[4]) -----------------------------------------------------------------------------
        sweep = None
        if args.sweep_level > 0:
            n_u, s_u, g_u, d_u = uniform_sweep(sb, torch, args.sweep_level)
            sweep = {"workload": f"uniform 2D level {args.sweep_level} upwind sweep, {n_u} cells, working set {16 * n_u / 1e6:.0f} MB > L2",
                     "cell_updates_per_s": n_u / s_u, "ms_per_sweep": 1e3 * s_u, "bound": "hbm", "achieved": 16.0 * n_u / s_u / 1e9,
                     "peak": peak_gbs, "unit": "GB/s", "frac": 16.0 * n_u / s_u / 1e9 / peak_gbs,
                     "diffusion_order2": {"ms_per_apply": 1e3 * d_u, "note": "rhs = make_diffusion_order2(K)(u): fill(0) + gather, 8 B zero-fill + 8 B read + 8 B write per cell",
                                          "achieved": 24.0 * n_u / d_u / 1e9, "frac": 24.0 * n_u / d_u / 1e9 / peak_gbs}}

            try:
                n_w, s_w = weno5_sweep(sb, torch, max(args.sweep_level - 1, 2))
                sweep["weno5"] = {"workload": f"rhs = make_convection_weno5(a)(u), uniform periodic 2D level {max(args.sweep_level - 1, 2)}, {n_w} cells",
                                  "ms_per_apply": 1e3 * s_w, "cells_per_s": n_w / s_w, "achieved": 24.0 * n_w / s_w / 1e9, "unit": "GB/s",
                                  "frac": 24.0 * n_w / s_w / 1e9 / peak_gbs,
                                  "note": "fp64-pipe bound (four WENO5 fluxes per cell, no FMA contraction for bit-exactness), not HBM bound"}
            except sb.SamuraiError as e:
                sweep["weno5"] = {"error": str(e)}

        # ---- flux-based scheme on the same adapted mesh (a8-a10): rhs = diffusion(u); unp1 = u - dt * rhs ----------------
        flux = None
        try:
            diff = sb.make_diffusion_order2([1.0] * args.dim)
            rhs = diff(sim.u)  # builds the face-classified batch of this mesh (host) and applies once
            st0 = sb.stats(reset=True)
            sb.profile_enable(True)
            for _ in range(5):
                diff.apply(rhs, sim.u)
            n_f, s_f, c_f = sb.profile_get()["fv"]
            sb.profile_enable(False)
            flux = {"scheme": "make_diffusion_order2 across level jumps (FluxGenOp)", "leaves": leaves_now, "us_per_apply": 1e6 * s_f / max(n_f, 1),
                    "GBps": 16.0 * c_f / s_f / 1e9 if s_f > 0 else None,
                    "note": "gather kernel alone (8 B read + 8 B write per leaf); the adapted mesh is L2 resident, see uniform_sweep.diffusion_order2 for the HBM-bound shape"}
            rhs.destroy()
        except Exception as e:  # noqa: BLE001
            flux = {"error": str(e)}

        # ---- the whole step on uniform meshes (configs[4] shapes): the HBM-bound form of the metric ---------------------
        ustep = None
        if args.sweep_level > 0:
            ustep = {}
            for d, lvl in ((2, args.sweep_level), (3, args.sweep_level_3d)):
                if lvl <= 0:
                    continue
                r = uniform_full_step(sb, torch, d, lvl)
                r.update({"bound": "hbm", "achieved": r["achieved_GBps"], "peak": peak_gbs, "unit": "GB/s", "frac": r["achieved_GBps"] / peak_gbs})
                ustep[f"{d}d_level{lvl}"] = r

        # ---- the numpy oracle itself on the full-size mesh (the checker pinned on the reference's golden files) --------
        np_parity = None
        if not args.no_cpu_baseline and args.numpy_parity_steps > 0:
            np_parity = numpy_oracle_parity(sb, sim, args, args.numpy_parity_steps)

        line = {
            "metric": METRIC, "value": cells / secs, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_string(args),
                       "leaves": leaves_now, "reference_cells": ref_now, "parallelism": f"replicas x{world}" if world > 1 else "single GPU",
                       "l2_policy": "every kernel is launched once per mesh state; between timed steps the mesh and all index batches change; "
                                    "the uniform sweep uses a working set > L2"},
            "split_ms_per_step": {"device": 1e3 * st["device_seconds"] / args.steps, "host_mesh": 1e3 * st["host_mesh_seconds"] / args.steps,
                                  "host_batches": 1e3 * st["host_batch_seconds"] / args.steps},
            "host_stages_ms_per_step": dict(zip(("tags_to_leaves", "mesh_equality", "graduation", "sub_meshes", "transfer_batches", "mesh_batches",
                                                 "wait_device", "release_old_mesh"), [1e3 * v / args.steps for v in st["host_stage_seconds"]])),
            "harten_iterations_per_step": st["harten_iterations"] / args.steps, "mesh_rebuilds_per_step": st["mesh_rebuilds"] / args.steps,
            "device_only_value": cells / world / st["device_seconds"] if st["device_seconds"] > 0 else None,
            "gpu_launches": int(st["kernel_launches"]),
            "initial_adaptation_s": init_secs,
            "e2e": {"value": e2e_cells / e2e_secs, "unit": UNIT, "h2d_bytes_per_step": int(st_e2e["h2d_bytes"] / args.steps),
                    "d2h_bytes_per_step": int(st_e2e["d2h_bytes"] / args.steps), "ms_per_step": 1e3 * e2e_secs / args.steps},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                         # dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` of two steady-state harten
                         # launches of this workload (30.4 MB and 31.4 MB; profiles/r02_ncu_wavefront.md): below the algorithmic
                         # bytes because the adapted working set stays in L2 between the sweeps.  Not re-measured by this run.
                         "traffic": 30.9e6 if (dom == "wavefront" and args.dim == 2 and args.max_level == 14) else None,
                         "peak_source": peak_src, "share_of_device_time": fam_time[dom] / total_prof,
                         "launches": n_l, "us_per_launch": 1e6 * s_l / n_l, "algorithmic_bytes_per_launch": dom_bytes / n_l,
                         "note": "adapted-mesh step: ~1e6 cells over ~40 dependent level sweeps per launch (grid barrier between sweeps): latency bound, "
                                 "the working set lives in L2; see uniform_sweep for the HBM-bound shape"},
            "kernel_families_fused": fused,
            "kernel_families_per_sweep_launches": families,
            "uniform_sweep": sweep,
            "uniform_full_step": ustep,
            "config4_3d": c4,
            "flux_scheme_on_adapted_mesh": flux,
            "cpu_baseline": cpu,
            "parity": parity,
            "parity_numpy_oracle": np_parity,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def product_leaves(sb, sim, max_level):
    """Leaf intervals (level, y, z, xs, xe) and leaf values of the product, in for_each_cell order."""
    rows, offs = [], []
    for level in range(max_level + 1):
        iv = sim.mesh.intervals(sb.CELLS, level)
        if iv.size == 0:
            continue
        rows.append(np.stack([np.full(iv.size, level), iv["y"], iv["z"], iv["start"], iv["end"]], axis=1).astype(np.int32))
        n = (iv["end"] - iv["start"]).astype(np.int64)
        rep = np.repeat(np.arange(iv.size), n)
        offs.append(iv["offset"][rep] + (np.arange(int(n.sum())) - np.repeat(np.cumsum(n) - n, n)))
    u = sim.u.download()
    return np.concatenate(rows), u[np.concatenate(offs)]


def parity_and_cpu_baseline(sb, sim, args, rank, world, barrier):
    """Every rank calls this.  Rank 0 hands the product's current state (leaves + leaf values) to the compiled CPU path
    (oracle/cpu_path.cpp, bit-identical to the numpy oracle), times `--cpu-steps` steps of it on all host cores (the
    cpu_baseline: a bounded sample of the same workload), then all ranks advance the product by the same steps and rank 0
    compares: leaves bit-identical, leaf values within 1e-12 relative (north_star).  At N > 1 this is the check that the
    N-GPU run equals the single-process result (the reference CI's procedure, .github/workflows/ci.yml:298-330)."""
    n_steps = args.cpu_steps
    if world > 1:
        sb.mg_broadcast(sim.u)
    cpu = parity = None
    cs = None
    if rank == 0:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import cpu_path

        saved = cpu_path.CpuSim.set_threads(os.cpu_count() or 1)  # the other ranks wait at the barrier below
        iv, vals = product_leaves(sb, sim, args.max_level)
        cs = cpu_path.CpuSim(args.dim, args.min_level, args.max_level, 1, eps=args.eps, regularity=1.0, leaves=iv, leaf_values=vals)
        a = [1.0] * args.dim
        cs.steps(1, a, sim.dt)  # warm-up (first-touch of the arenas)
        t0 = time.perf_counter()
        done = cs.steps(n_steps, a, sim.dt)
        secs = time.perf_counter() - t0
        tm = cs.times()
        cores = cpu_path.CpuSim.threads()
        cpu_path.CpuSim.set_threads(saved)
        # at N > 1 the run is only the parity reference (torchrun pins the ranks' host threads, its timing means nothing)
        cpu = None if world > 1 else {"value": done / secs, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n_steps} steps from the product's own adapted state ({cs.nb_cells()} leaves); compiled C++/OpenMP port oracle/cpu_path.cpp "
                         f"(-O3 -march=x86-64-v3 -ffp-contract=off), all host threads" + (f"; {world} ranks were idle meanwhile" if world > 1 else ""),
               "ms_per_step": 1e3 * secs / n_steps}
    barrier()
    for _ in range(n_steps + 1):
        sim.step()
    if world > 1:
        sb.mg_broadcast(sim.u)
    if rank == 0:
        iv2, vals2 = product_leaves(sb, sim, args.max_level)
        civ, cvals = cs.leaves()
        same = bool(iv2.shape == civ.shape and np.array_equal(iv2, civ))
        err = float(np.max(np.abs(vals2 - cvals) / np.maximum(np.abs(cvals), 1.0))) if same else None
        parity = {"against": "compiled CPU path started from the same state (bit-identical to the numpy oracle, tests/test_cpu_path.py)",
                  "steps": n_steps + 1, "mesh_identical": same, "max_rel_err": err, "leaves": int(vals2.size),
                  "leaf_sum_product": float(np.sum(vals2)), "leaf_sum_cpu": float(np.sum(cvals))}
        cs.close()
    return cpu, parity


def numpy_oracle_parity(sb, sim, args, n_steps):
    """The numpy oracle (the checker pinned on the reference's golden files) on the product's current full-size mesh."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import samurai_oracle as so

    lv, co, off = sim.mesh.cell_table(sb.CELLS)
    pu = sim.u.download()
    cfg = so.MeshConfig(dim=args.dim, min_level=args.min_level, max_level=args.max_level, pred_radius=1)
    cells = {int(l): np.sort(so.pack(co[lv == l])) for l in np.unique(lv)}
    omesh = so.Mesh(cfg, cells)
    ou = np.zeros(omesh.nref)
    olv, oco, oix = omesh.leaf_table()
    # both tables are in for_each_cell order
    assert np.array_equal(olv, lv) and np.array_equal(oco, co)
    ou[oix] = pu[off]
    bc = so.Bc("dirichlet", 0.0)
    cells_done, secs, omesh, ou = oracle_steps(so, cfg, bc, omesh, ou, n_steps, sim.dt)
    for _ in range(n_steps):
        sim.step()
    lv2, co2, off2 = sim.mesh.cell_table(sb.CELLS)
    olv, oco, oix = omesh.leaf_table()
    same_mesh = bool(np.array_equal(olv, lv2) and np.array_equal(oco, co2))
    pu2 = sim.u.download()
    err = float(np.max(np.abs(pu2[off2] - ou[oix]) / np.maximum(np.abs(ou[oix]), 1.0))) if same_mesh else None
    return {"steps": n_steps, "leaves": int(omesh.nb_cells()), "mesh_identical": same_mesh, "max_rel_err": err,
            "numpy_oracle_cell_updates_per_s": cells_done / secs}


def main():
    global ARGS
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dim", type=int, default=2, help="2: advection_2d (configs[1], the headline); 3: advection_3d (configs[3])")
    ap.add_argument("--min-level", type=int, default=4)
    ap.add_argument("--max-level", type=int, default=14)
    ap.add_argument("--eps", type=float, default=2e-4)
    ap.add_argument("--sweep-level", type=int, default=13)
    ap.add_argument("--cpu-steps", type=int, default=10, help="steps of the compiled CPU path timed as cpu_baseline (and compared with the product)")
    ap.add_argument("--numpy-parity-steps", type=int, default=1, help="steps of the numpy oracle compared with the product at full size (N=1)")
    ap.add_argument("--level-3d", type=int, default=10, help="max_level of the 3D advection block (BASELINE configs[3]; 0: skip)")
    ap.add_argument("--steps-3d", type=int, default=20)
    ap.add_argument("--sweep-level-3d", type=int, default=9, help="level of the 3D uniform full-step measurement (0: skip)")
    ap.add_argument("--ref-max-level", type=int, default=14, help="largest max_level the CPU reference arm samples")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--clock-period", type=float, default=0.1, help="seconds between NVML clock samples during the timed region")
    ARGS = ap.parse_args()
    if ARGS.impl == "reference":
        run_reference(ARGS)
    else:
        run_product(ARGS)


if __name__ == "__main__":
    main()
