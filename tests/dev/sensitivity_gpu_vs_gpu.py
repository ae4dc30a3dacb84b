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
fscale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.02
body = (8 * par.nu_lbm * fscale / 17.0 ** 2, 0.0, 0.0)
rbc, plt = O.rbc_celltype(par), O.plt_celltype(par)
rbc_cells = U.deformed_cells(rbc, [(10.0, 18.5, 13.0), (12.0, 18.0, 24.5), (30.0, 12.5, 18.5), (34.0, 25.0, 19.0)], 11, amp=0.0, stretch=(1.0, 1.0, 1.0))
plt_cells = U.deformed_cells(plt, [(22.0, 18.5, 6.0), (42.0, 30.0, 22.0)], 12, amp=0.0, stretch=(1.0, 1.0, 1.0))
def mk(mode):
    ctx = U.gpu_context(dom, fl, None, body); ctx.set_force_limit(par.f_limit); ctx.set_spread_mode(mode, 20)
    for k, (ct, cc, ids) in enumerate([(rbc, rbc_cells, [0, 1, 2, 3]), (plt, plt_cells, [4, 5])]):
        t = U.gpu_add_type(ctx, ct); ctx.add_cells(t, cc, ids); ctx.set_material_timescale(t, 20)
    ctx.set_timescales(5, 1, 1)
    return ctx
a, b = mk(1), mk(0)
for blk in range(30):
    a.iterate(50); b.iterate(50)
    pa = a.cells_download(H.P_POS); pb = b.cells_download(H.P_POS)
    print(a.iteration, "GPU(sorted spread) vs GPU(plain atomics): max pos diff %.3e" % np.abs(pa - pb).max(), flush=True)
