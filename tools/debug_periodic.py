import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np
import parity_utils as pu
from parity_utils import sb, so
dim, lmin, L, per, msr = 2, 2, 6, tuple(bool(int(x)) for x in os.environ.get("PER", "1,1").split(",")), 1
sb.initialize(0)
if os.environ.get("UNFUSED"): sb.set_fused(False)
ocfg = pu.oracle_cfg(dim, lmin, L, 1, per, msr)
bc = so.Bc("dirichlet", 0.0)
omesh = so.Mesh.uniform(ocfg)
ou = so.init_disc(omesh, [0.3] * dim, 0.2)
pmesh = sb.MRMesh.make_mesh([0.0] * dim, [1.0] * dim, pu.product_cfg(dim, lmin, L, 1, per, msr))
u = sb.make_scalar_field("u", pmesh); u.resize(); u.upload(ou)
if not all(per): sb.make_bc(u, sb.DIRICHLET, 0.0)
unp1 = sb.make_scalar_field("unp1", pmesh)
adapt = sb.make_MRAdapt(u)
mcfg = sb.mra_config().epsilon(2e-4)
dt = 0.5 * pmesh.min_cell_length()
def where(omesh, idx):
    out = []
    for l in range(omesh.nlev):
        if omesh.ref[l].size:
            oi = omesh.index(l, omesh.ref[l])
            m = np.isin(oi, idx)
            for k in omesh.ref[l][m][:6]:
                out.append((l, so.unpack(np.array([k]), dim)[0].tolist()))
    return out
for it in range(-1, 8):
    trace = []
    omesh2, ou2 = so.adapt(omesh, ou, bc, 2e-4, 1.0, trace=trace)
    for ite in range(L - lmin):
        # compare ghosts before the iteration's detail: run product iteration
        unchanged = adapt.iteration(mcfg, ite)
        ref = trace[ite]
        det = adapt.last_detail(); tags = adapt.last_tags()
        bad = np.flatnonzero(np.abs(det - ref["detail"]) > 1e-12)
        badt = np.flatnonzero(tags != ref["tag"])
        print(f"step {it} ite {ite}: nref {det.size} detail bad {bad.size} tag bad {badt.size}")
        if badt.size:
            print("  tags differ at", where(ref["mesh"], badt)[:10], tags[badt][:8], ref["tag"][badt][:8])
        if bad.size or badt.size:
            print("  detail differs at", where(ref["mesh"], bad)[:10])
            pf = u.download(); of = ref["field"]; m = np.isfinite(of)
            bg = np.flatnonzero(m & (np.abs(pf - np.where(m, of, 0)) > 1e-12))
            print("  field (after the iteration's ghost update) differs at", where(ref["mesh"], bg)[:12])
            # compare the field the detail was computed from (after the ghost update inside harten)
            sys.exit(0)
        if unchanged: break
    omesh, ou = omesh2, ou2
    sb.update_ghost_mr(u); so.update_ghost_mr(omesh, ou, bc)
    pf = u.download(); m = np.isfinite(ou)
    badg = np.flatnonzero(m & (np.abs(pf - np.where(m, ou, 0)) > 1e-12))
    print(f"step {it}: ghosts bad {badg.size}", where(omesh, badg)[:8])
    unp1.resize(); sb.upwind_step(unp1, u, [1.0] * dim, dt); ou = so.fv_step(omesh, ou, [1.0] * dim, dt); sb.swap(u, unp1)
