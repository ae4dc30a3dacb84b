import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as O
from oracle import mesh as M
import util as U
from hemocell_b200 import lib as H
par = M.Parameters(dx=0.5e-6, dt=1e-7)
nx, ny, nz = 48, 38, 38
y, z = np.meshgrid(np.arange(ny), np.arange(nz), indexing="ij")
r2 = (y - (ny - 1) / 2.0) ** 2 + (z - (nz - 1) / 2.0) ** 2
fl3 = np.zeros((nx, ny, nz), dtype=np.uint8); fl3[:, r2 > 17.0 ** 2] = 1
fl = fl3.reshape(-1)
dom = O.make_domain(nx, ny, nz, (1, 0, 0), par.tau)
body = (8 * par.nu_lbm * 0.02 / 17.0 ** 2, 0.0, 0.0)
rbc, plt = O.rbc_celltype(par), O.plt_celltype(par)
rbc_cells = U.deformed_cells(rbc, [(10.0, 18.5, 13.0), (12.0, 18.0, 24.5), (30.0, 12.5, 18.5), (34.0, 25.0, 19.0)], 11, amp=0.0, stretch=(1.0, 1.0, 1.0))
plt_cells = U.deformed_cells(plt, [(22.0, 18.5, 6.0), (42.0, 30.0, 22.0)], 12, amp=0.0, stretch=(1.0, 1.0, 1.0))
O.set_parallel(1)
vts = int(sys.argv[1]) if len(sys.argv) > 1 else 5
mts = int(sys.argv[2]) if len(sys.argv) > 2 else 20
sim = O.OracleSim(dom, fl, par.f_limit, body); sim.vel_timescale = vts
ctx = U.gpu_context(dom, fl, None, body); ctx.set_force_limit(par.f_limit)
for k, (ct, cc, ids) in enumerate([(rbc, rbc_cells, [0, 1, 2, 3]), (plt, plt_cells, [4, 5])]):
    sim.add_celltype(ct, mts); sim.add_cells(k, cc, ids)
    t = U.gpu_add_type(ctx, ct); ctx.add_cells(t, cc, ids); ctx.set_material_timescale(t, mts)
ctx.set_timescales(vts, 1, 1)
t0 = time.time()
for blk in range(30):
    for _ in range(50): sim.iterate()
    ctx.iterate(50)
    p = ctx.cells_download(H.P_POS).reshape(-1, 3); d = np.abs(p - sim.pos)
    i = np.unravel_index(np.argmax(d), d.shape)
    pop = np.abs(ctx.lattice_download(H.LAT_POP) - sim.pop).max()
    print(sim.iter, "max pos diff %.3e at particle %d comp %d (pos %.3f)  pop diff %.3e  xmax %.2f alive %d/%d  t=%.1f" % (d.max(), i[0], i[1], sim.pos[i[0], i[1]], pop, sim.pos[:,0].max(), ctx.count()[0], len(sim.ctype), time.time()-t0), flush=True)
