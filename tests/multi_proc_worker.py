"""dev/test worker: one PROCESS per GPU (the bench.py / torchrun topology, where the peer transport maps the
neighbour through CUDA IPC).  argv: rank n_ranks workdir transport.  Rank 0 writes the NCCL id to
workdir/uid.bin; every rank writes workdir/out_<rank>.npz."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import oracle as O
from oracle import mesh as M
from hemocell_b200 import lib as H
import util as U

rank, R, work, transport = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
cfg = np.load(os.path.join(work, "problem.npz"))
nx, ny, nz = [int(v) for v in cfg["dims"]]
periodic = tuple(int(v) for v in cfg["periodic"])
par = M.Parameters(dx=0.5e-6, dt=0.5e-7)
ct = O.rbc_celltype(par)
uid_path = os.path.join(work, "uid.bin")
if rank == 0:
    uid = H.Context.unique_id()
    with open(uid_path + ".tmp", "wb") as f:
        f.write(uid)
    os.rename(uid_path + ".tmp", uid_path)
else:
    t0 = time.time()
    while not os.path.exists(uid_path):
        if time.time() - t0 > 120:
            raise SystemExit("no NCCL id")
        time.sleep(0.05)
    uid = open(uid_path, "rb").read()
nxl = nx // R
ndev = H.load().hcg_device_count()
local = ndev < R or os.environ.get("HCG_TEST_LOCAL") == "1"      # fewer GPUs than ranks: share them (host-staged communicator)
ctx = H.Context(nx, ny, nz, periodic, par.tau, device=rank % max(ndev, 1), rank=rank, n_ranks=R)
ctx.set_transport(transport)
ctx.comm_init(uid, local=local)
fl3 = cfg["flags"].reshape(nx, ny, nz)
ctx.set_flags(np.ascontiguousarray(fl3[rank * nxl:(rank + 1) * nxl]))
for o in range(6):
    ctx.set_bc_velocity(o, cfg["bc"][o])
ctx.set_body_force(tuple(cfg["body"])); ctx.init_equilibrium(1.0, tuple(cfg["u0"])); ctx.set_force_limit(par.f_limit)
ctx.set_exchange(4.0, int(cfg["sync_every"]), 0.5)
t = ctx.add_celltype(ct.model, ct.cc, ct.k)
ctx.add_cells(t, cfg["cells"], cfg["ids"])
cad = int(cfg["cadence"])
ctx.set_timescales(cad, 1, 1); ctx.set_material_timescale(t, cad)
ctx.iterate(int(cfg["steps"]))
cid, _, alive = ctx.cells_info()
st = ctx.exchange_stats()
np.savez(os.path.join(work, f"out_{rank}.npz"), pop=ctx.lattice_download(H.LAT_POP), pos=ctx.cells_download(H.P_POS),
         ids=cid, alive=alive, count=np.array(ctx.count()), migrated_in=st["migrated_in"])
ctx.close()
print("worker", rank, "done", flush=True)
