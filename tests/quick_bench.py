"""Scratch micro-benchmark used during development (not the driver's bench.py)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import oracle as O
from oracle import mesh as M
from hemocell_b200 import lib as H

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    ncell = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
    par = M.Parameters(dx=0.5e-6, dt=-1)
    ctx = H.Context(n, n, n, (1, 1, 1), par.tau)
    ctx.set_flags(np.zeros(n**3, dtype=np.uint8))
    ctx.set_body_force((1e-7, 1e-7, 1e-7))
    ctx.set_force_limit(par.f_limit)
    # fluid only
    ctx.fluid_warmup(5)
    ms = ctx.iterate_timed(20)
    print(f"fluid only {n}^3: {ms/20:.3f} ms/step  {n**3/(ms/20)/1e3:.0f} MLUPS  {n**3*304/(ms/20)/1e6:.0f} GB/s(304B)")
    ct = O.rbc_celltype(par)
    rng = np.random.default_rng(0)
    g = int(round(ncell ** (1/3))) + 1
    centers = []
    for i in range(g):
        for j in range(g):
            for k in range(g):
                if len(centers) < ncell:
                    centers.append(((i + 0.5) * n / g, (j + 0.5) * n / g, (k + 0.5) * n / g))
    v0 = ct.verts - 0.5 * (ct.verts.min(0) + ct.verts.max(0))
    cells = np.array([v0 + np.array(c) for c in centers])
    t = ctx.add_celltype(ct.model, ct.cc, ct.k)
    ctx.add_cells(t, cells, np.arange(len(centers)))
    for (vts, mts) in ((1, 20), (5, 20)):
        ctx.set_timescales(vts, 1, 1); ctx.set_material_timescale(t, mts)
        ctx.set_iteration(0)
        ctx.iterate(20)
        ms = ctx.iterate_timed(40)
        print(f"with {len(centers)} RBC vel/{vts} mat/{mts}: {ms/40:.3f} ms/step {n**3/(ms/40)/1e3:.0f} MLUPS")
        ctx.timers_enable(True); ctx.timers_reset(); ctx.iterate(40); ctx.timers_enable(False)
        for k, (tot, calls) in ctx.timers().items():
            print(f"   {k:32s} {tot/40:.3f} ms/step ({calls} calls)")
    ctx.close()

main()
