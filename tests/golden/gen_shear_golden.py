#!/usr/bin/env python
"""Generate tests/golden/shear_oracle.json: examples/oneCellShear (one RBC at (9.5, 9.5, 4.5) um, rotated
(90, 0, 0), in a 40x40x20 box sheared at 111 1/s through regularized velocity planes, dt = 0.5e-7) run
with the CPU ORACLE.  Recorded every `tmeas` steps, as the reference's stretch.log does
(examples/oneCellShear/oneCellShear.cpp:147-161): bounding-box diameters, volume %, area %, largest
diameter and the deformation index  DI = (D^2 - 1)/(D^2 + 1) * 100, D = D_max / (2 * 3.91 um).
usage: tests/golden/gen_shear_golden.py [iterations] [tmeas]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle as O
from oracle import mesh as M
import util as U


def observables(sim, ct, par, vol_eq, area_eq):
    um = par.dx / 1e-6
    p = sim.pos
    d = p[:, None, :] - p[None, :, :]
    dmax = float(np.sqrt((d * d).sum(-1).max())) * um
    ext = (p.max(0) - p.min(0)) * um
    tri = np.asarray(ct.cc["triangle_list"])
    a = 0.5 * np.linalg.norm(np.cross(p[tri[:, 1]] - p[tri[:, 0]], p[tri[:, 2]] - p[tri[:, 0]]), axis=1).sum()
    v = M.mesh_volume(p, ct.cc["triangle_list"])
    D = dmax / (2 * 3.91)
    return {"iter": int(sim.iter), "diam_um": ext.tolist(), "volume_pct": v / vol_eq * 100, "area_pct": a / area_eq * 100,
            "largest_diam_um": dmax, "deformation_index_pct": (D * D - 1) / (D * D + 1) * 100, "center_um": (p.mean(0) * um).tolist()}


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    tmeas = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    nx, ny, nz = 40, 40, 20
    par = M.Parameters(dx=0.5e-6, dt=0.5e-7)
    vh = (nz - 1) * 111.0 * par.dt * 0.5
    bc = np.zeros((6, 3)); bc[4] = (vh, 0, 0); bc[5] = (-vh, 0, 0)
    fl = U.couette_flags(nx, ny, nz).reshape(-1)
    dom = O.make_domain(nx, ny, nz, (1, 1, 0), par.tau, bc)
    ct = O.rbc_celltype(par)
    cells, ids = M.place_cells(ct.verts, np.array([[9.5, 9.5, 4.5, 90.0, 0.0, 0.0]]), par.dx, (nx, ny, nz), fl)
    sim = O.OracleSim(dom, fl, par.f_limit)
    sim.add_celltype(ct, 1); sim.add_cells(0, cells, ids)
    O.set_parallel(1)
    tri = np.asarray(ct.cc["triangle_list"]); p = sim.pos
    area_eq = 0.5 * np.linalg.norm(np.cross(p[tri[:, 1]] - p[tri[:, 0]], p[tri[:, 2]] - p[tri[:, 0]]), axis=1).sum()
    vol_eq = M.mesh_volume(p, ct.cc["triangle_list"])
    out = {"source": "examples/oneCellShear (config.xml, RBC.xml, RBC.pos) run with the CPU oracle", "tmeas": tmeas, "trace": []}
    t0 = time.time()
    for _ in range(iters // tmeas):
        for _ in range(tmeas):
            sim.iterate()
        out["trace"].append(observables(sim, ct, par, vol_eq, area_eq))
        print(out["trace"][-1], f"{time.time() - t0:.0f}s", flush=True)
        json.dump(out, open(os.path.join(ROOT, "tests", "golden", "shear_oracle.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
