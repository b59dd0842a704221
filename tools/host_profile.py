"""Time the host-side mesh + batch construction on a dumped adapted mesh (no GPU needed)."""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import samurai_b200 as sb
sb.initialize(-1)
L = int(sys.argv[1]) if len(sys.argv) > 1 else 14
d = np.load(os.path.join(ROOT, "gpurun_out", f"mesh_L{L}.npz"))
cfg = sb.mesh_config(2, 1).min_level(4).max_level(L).max_stencil_size(2).disable_minimal_ghost_width()
m = sb.MRMesh.from_intervals([0, 0], [1, 1], cfg, d["levels"], d["intervals"])
print("leaves", m.nb_cells(), "ref", m.nb_cells(sb.REFERENCE))
tm, tp, nb = m.debug_host_rebuild(5)
print(f"mesh build {tm*1e3:.2f} ms   plan build {tp*1e3:.2f} ms   arena {nb/1e6:.2f} MB")
