"""dev: fused wavefront vs separate launches on a tall-z box (R = 1 path), same process, two contexts"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import oracle as O
from oracle import mesh as M
from hemocell_b200 import lib as H
import util as U

nx, ny, nz = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (24, 24, 256))]
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 6
par = M.Parameters(dx=0.5e-6, dt=-1)
ct = O.rbc_celltype(par)
cells = U.deformed_cells(ct, [(nx / 2, ny / 2, nz / 4), (nx / 2, ny / 2, 3 * nz / 4)], 3, amp=0.0)
ctx = H.Context(nx, ny, nz, (1, 1, 1), par.tau)
ctx.set_flags(np.zeros(nx * ny * nz, dtype=np.uint8))
ctx.set_body_force((1e-6, 2e-6, -1e-6))
ctx.set_force_limit(par.f_limit)
t = ctx.add_celltype(ct.model, ct.cc, ct.k)
ctx.add_cells(t, cells, np.arange(2))
ctx.set_timescales(1, 1, 1); ctx.set_material_timescale(t, 2)
ctx.iterate(steps)
pos = ctx.cells_download(H.P_POS); pop = ctx.lattice_download(H.LAT_POP)
np.save(f"/tmp/fused_{os.environ.get('HCG_FUSED', '1')}_pos.npy", pos)
np.save(f"/tmp/fused_{os.environ.get('HCG_FUSED', '1')}_pop.npy", pop)
print("done", os.environ.get("HCG_FUSED", "1"), float(np.abs(pos).sum()), float(np.abs(pop).sum()))
if os.environ.get("HCG_FUSED", "1") == "1" and os.path.exists("/tmp/fused_0_pos.npy"):
    U.assert_close(pos, np.load("/tmp/fused_0_pos.npy"), "positions fused vs separate", rtol=1e-13)
    U.assert_close(pop, np.load("/tmp/fused_0_pop.npy"), "populations fused vs separate", rtol=1e-12)
    print("fused == separate")
