#!/usr/bin/env python
"""Generate tests/golden/stretch_oracle.json: the reference's stretch-cell validation
(tests/validation/stretch_cell/test_stretch_cell.cpp) run with the CPU ORACLE.
One RBC in a closed 52x26x26 box with u = 0 regularized walls, 7 forced vertices per side,
10 000 iterations (about 3 minutes per force on one core).
usage: tests/golden/gen_stretch_golden.py [iterations]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import oracle as O
from oracle import mesh as M

BOUNDS = {25: ((7.3, 7.9), (9.2, 9.7)), 75: ((7.0, 7.5), (11.0, 12.0)), 125: ((6.5, 7.0), (12.25, 12.75))}
N_FORCED = 7


def setup(par):
    nx, ny, nz = 52, 26, 26
    fl = np.zeros((nx, ny, nz), dtype=np.uint8)
    fl[:, :, 0] = 6; fl[:, :, nz - 1] = 7
    fl[:, 0, :] = 4; fl[:, ny - 1, :] = 5
    fl[0, :, :] = 2; fl[nx - 1, :, :] = 3
    return (nx, ny, nz), fl.reshape(-1)


def run(force_pN, iters, checkpoints=(200,)):
    par = M.Parameters(dx=0.5e-6, dt=1e-7)
    dims, fl = setup(par)
    dom = O.make_domain(*dims, (0, 0, 0), par.tau, np.zeros((6, 3)))
    sim = O.OracleSim(dom, fl, par.f_limit)
    ct = O.rbc_celltype(par)
    sim.add_celltype(ct, 1)
    pos, ids = M.place_cells(ct.verts, np.array([[12.0, 6, 6, 90, 0, 0]]), par.dx, dims, fl)
    sim.add_cells(0, pos, ids)
    ef = force_pN * (1e-12 / par.df) / N_FORCED
    order = np.argsort(sim.pos[:, 0], kind="stable")
    lower, upper = order[:N_FORCED], order[::-1][:N_FORCED]
    out = {"force_pN": force_pN, "iterations": iters, "lower": lower.tolist(), "upper": upper.tolist(), "trace": {}}
    for _ in range(iters):
        sim.pforce[lower, 0] -= ef
        sim.pforce[upper, 0] += ef
        sim.iterate()
        if sim.iter in checkpoints or sim.iter == iters:
            ext = (sim.pos.max(0) - sim.pos.min(0)) * 0.5          # um
            out["trace"][str(sim.iter)] = {"axial_um": ext[0], "transverse_um": ext[1],
                                           "volume_ratio": M.mesh_volume(sim.pos, ct.cc["triangle_list"]) / ct.cc["volume_eq"]}
    return out


if __name__ == "__main__":
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    res = {"bounds_um": {str(k): {"transverse": v[0], "axial": v[1]} for k, v in BOUNDS.items()},
           "source": "tests/validation/stretch_cell/test_stretch_cell.cpp:158-162 (reference)",
           "runs": [run(f, iters) for f in (25, 75, 125)]}
    json.dump(res, open(os.path.join(ROOT, "tests", "golden", "stretch_oracle.json"), "w"), indent=1)
    print(json.dumps(res["runs"], indent=1))
