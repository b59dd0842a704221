"""Per-step wall times of the bench workload (sync after every step) to find outliers."""
import os, sys, time, argparse
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import samurai_b200 as sb
ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=120)
ap.add_argument("--dim", type=int, default=2)
ap.add_argument("--min-level", type=int, default=4)
ap.add_argument("--max-level", type=int, default=14)
ap.add_argument("--eps", type=float, default=2e-4)
a = ap.parse_args()
sb.initialize(0)
sim = bench.Sim(sb, a)
sim.adapt(sim.mra)
sb.synchronize()
ts = []
allparts = []
for i in range(a.steps):
    sb.stats(reset=True)
    t0 = time.perf_counter()
    parts = []
    ta = time.perf_counter(); sim.adapt(sim.mra); parts.append(time.perf_counter() - ta)
    ta = time.perf_counter(); sb.update_ghost_mr(sim.u); parts.append(time.perf_counter() - ta)
    ta = time.perf_counter(); sim.unp1.resize(); parts.append(time.perf_counter() - ta)
    ta = time.perf_counter(); sb.upwind_step(sim.unp1, sim.u, sim.a, sim.dt); parts.append(time.perf_counter() - ta)
    ta = time.perf_counter(); sb.swap(sim.u, sim.unp1); parts.append(time.perf_counter() - ta)
    ta = time.perf_counter(); sb.synchronize(); parts.append(time.perf_counter() - ta)
    allparts.append(parts)
    t1 = time.perf_counter()
    st = sb.stats()
    ts.append((1e3 * (t1 - t0), 1e3 * st["device_seconds"], 1e3 * (st["host_mesh_seconds"] + st["host_batch_seconds"]), sim.mesh.nb_cells()))
ts = np.array(ts)
print("median wall %.2f ms, mean %.2f, p90 %.2f, max %.2f; median device %.2f host %.2f" % (np.median(ts[:, 0]), ts[:, 0].mean(), np.percentile(ts[:, 0], 90), ts[:, 0].max(), np.median(ts[:, 1]), np.median(ts[:, 2])))
for i in np.argsort(-ts[:, 0])[:12]:
    print("step %3d wall %.2f device %.2f host %.2f leaves %d" % (i, *ts[i]), " adapt/ghost/resize/upwind/swap/sync ms:", " ".join("%.2f" % (1e3 * v) for v in allparts[i]))
print("unaccounted median %.2f" % np.median(ts[:, 0] - ts[:, 1] - ts[:, 2]))
