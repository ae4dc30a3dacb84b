#!/usr/bin/env python
"""bench.py -- throughput of the HemoCell per-timestep IB-LBM hot path on B200.

Metric (BASELINE.json): MLUPS (D3Q19 fp64, with RBCs) and cell-steps/s.
Workload at N GPUs: the cases/performance_testing unit, 256^3 lattice nodes per GPU (domain 256*N x 256 x 256, slabs
along x), fully periodic, tau = 1, body force (f,f,f), the reference's own cases/performance_testing/hematocrit_33/RBC.pos
(10 935 rows packed into a 135 um cube; placed on the 128 um unit with the reference's reader rule - a cell survives iff
every vertex lies inside the domain, pinned by the 42-cell known answer of tests/validation/pipeflow - 7736 randomly
oriented RBC = 4.97 M LSP survive; fixtures/, sha256 in the line), tiled along x for N > 1 as
examples/cube/preprocess/analysis.py:cell_positions does; material update every 20 steps, velocity interpolation every
`--cadence` steps (1 = configs/, 5 = configs_timestep_5/).  `--workload synthetic` = round 1's crystal packing (8464 aligned discs).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # CPU oracle on the host cores
    python bench.py --workload cube                          # extra line: examples/cube (BASELINE configs[1]), one GPU

The JSON line carries `roofline` for the WHOLE lattice update (every lattice kernel of a step - collision + moments pass, or
the moment-only kernel - against SURVEY 8(d)'s 304 B/LU), `roofline_kernels` (each kernel against its own algorithmic bytes),
`roofline_step` (whole step against 304 + (264 + 72 + 216/c) N_LSP/N_nodes B/LU), `parity_check` (a small slab-decomposed
problem of the same code path checked against the CPU oracle before the timed region, on all N ranks), `e2e` (host buffers,
copies inside the timed region), `cpu_baseline` (rank 0, N = 1), `gpu_launches` and the SM clocks sampled during the timed region.

One JSON line on stdout (rank 0).  A "step" is one HemoCell::iterate().
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "MLUPS (D3Q19 fp64, with RBCs)"
UNIT_N = 256                       # lattice nodes per edge of one weak-scaling unit
DX = 0.5e-6
B_LU = 304.0                       # algorithmic bytes per lattice update (19 reads + 19 writes, fp64)
B_LU_TAU1 = 216.0                  # tau = 1 collision from the kept raw moments: 32 B moments + 32 B force read, 19 writes
B_MOM = 248.0                      # moments pass: 19 reads + 32 B force read + 32 B velocity write + 32 B force reset
B_MOM_TAU1 = 280.0                 # ... + 32 B raw moments kept for the next tau = 1 collision


# ----------------------------------------------------------------------------- workload
def synthetic_rows(unit_n=UNIT_N, seed=1234):
    """Seeded hematocrit packing of one unit: RBC discs on a 15 x 15 x 38 lattice (x, y, z spacing
    17.07 x 17.07 x 6.74 lu, disc axis along z), jittered by +-0.25 lu.  8550 cells * 81.1 um^3 /
    (128 um)^3 = 33 % hematocrit.  Rows are .pos records: centre in um, angles in degrees."""
    s = unit_n / 256.0
    nxy, nz = max(1, int(round(15 * s))), max(1, int(round(38 * s)))
    rng = np.random.default_rng(seed)
    ix, iy, iz = np.meshgrid(np.arange(nxy), np.arange(nxy), np.arange(nz), indexing="ij")
    ctr = np.stack([(ix + 0.5) * unit_n / nxy, (iy + 0.5) * unit_n / nxy, (iz + 0.5) * unit_n / nz], -1).reshape(-1, 3)
    ctr = ctr + rng.uniform(-0.25, 0.25, ctr.shape)
    rows = np.zeros((ctr.shape[0], 6))
    rows[:, 0:3] = ctr * (DX / 1e-6)          # lattice units -> um
    rows[:, 3] = 90.0                          # disc axis y -> z, as examples/oneCellShear/RBC.pos
    return rows


def body_force(nu_lbm, n):
    """performance_testing.cpp:74-78: 8 nu (u_max/2) / R^2 with u_max = Re nu / (2 R), Re 0.5, R = n/2"""
    u_max = 0.5 * nu_lbm / n           # lbm_pipe_parameters(cfg, nx): pipe_radius = nx
    r = n / 2.0
    f = 8 * nu_lbm * (u_max * 0.5) / r / r
    return (f, f, f)


def cube_setup(H, par):
    """examples/cube (BASELINE.json configs[1]): 100^3 nodes (50 um), x periodic, bounce-back planes at y = 0 / ny-1,
    regularized moving walls at z = 0 / nz-1 (shear rate 1 /s as in the config template), tau = 1, RBCs + PLTs at
    ~30 % hematocrit from a seeded lattice packing (the reference seeds with the irreproducible tools/packCells)."""
    n = 100
    fl = np.zeros((n, n, n), dtype=np.uint8)
    fl[:, :, 0] = H.VEL_ZN; fl[:, :, n - 1] = H.VEL_ZP
    fl[:, 0, :] = H.BOUNCEBACK; fl[:, n - 1, :] = H.BOUNCEBACK
    vhalf = (n - 1) * 1.0 * par["dt"] * 0.5
    bc = np.zeros((6, 3)); bc[4] = (vhalf, 0, 0); bc[5] = (-vhalf, 0, 0)
    rng = np.random.default_rng(4321)
    ix, iy, iz = np.meshgrid(np.arange(5), np.arange(5), np.arange(14), indexing="ij")
    ctr = np.stack([(ix + 0.5) * n / 5, 8.0 + (iy + 0.5) * (n - 16.0) / 5, 4.0 + (iz + 0.5) * (n - 8.0) / 14], -1).reshape(-1, 3)
    ctr = ctr + rng.uniform(-0.2, 0.2, ctr.shape)
    rbc_rows = np.zeros((ctr.shape[0], 6)); rbc_rows[:, 0:3] = ctr * (DX / 1e-6); rbc_rows[:, 3] = 90.0
    # platelets in the gaps between RBC columns (x, y offset by half a pitch)
    jx, jy, jz = np.meshgrid(np.arange(3), np.arange(3), np.arange(3), indexing="ij")
    pc = np.stack([(jx + 1.0) * n / 5, 8.0 + (jy + 1.0) * (n - 16.0) / 5, 15.0 + jz * 30.0], -1).reshape(-1, 3)
    plt_rows = np.zeros((pc.shape[0], 6)); plt_rows[:, 0:3] = pc * (DX / 1e-6)
    return n, fl, bc, rbc_rows, plt_rows


# ----------------------------------------------------------------------------- helpers
def pinned_empty(n_doubles):
    """page-locked host buffer through cudart (the C ABI takes plain host pointers)"""
    rt = C.CDLL("libcudart.so.12")
    ptr = C.c_void_p()
    rc = rt.cudaHostAlloc(C.byref(ptr), C.c_size_t(8 * n_doubles), C.c_uint(0))
    if rc != 0:
        raise RuntimeError(f"cudaHostAlloc failed: {rc}")
    buf = (C.c_double * n_doubles).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=np.float64)
    _PINNED.append((rt, ptr, buf))
    return arr


_PINNED = []


def pinned_empty_f32(n):
    rt = C.CDLL("libcudart.so.12")
    ptr = C.c_void_p()
    rc = rt.cudaHostAlloc(C.byref(ptr), C.c_size_t(4 * max(n, 1)), C.c_uint(0))
    if rc != 0:
        raise RuntimeError(f"cudaHostAlloc failed: {rc}")
    buf = (C.c_float * max(n, 1)).from_address(ptr.value)
    _PINNED.append((rt, ptr, buf))
    return np.frombuffer(buf, dtype=np.float32)


try:
    _AFFINITY0 = os.sched_getaffinity(0)
except Exception:
    _AFFINITY0 = set()


def bind_to_gpu_numa_node(index):
    """what `numactl` / the MPI launcher does for the reference: run this rank on the cores next to its GPU, so that its pinned
    host buffers are allocated on that socket's memory (8 ranks copying at once otherwise meet on one socket's memory controller)"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return len(allowed)
    except Exception:
        pass
    return 0


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.lines, self.p = [], None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for ln in self.p.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, t0, t1):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        sm, mx, reasons = [], None, set()
        for ts, ln in self.lines:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in ln.split(",")]
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except Exception:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel="k_collide_stream"):
    p = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(p):
        d = json.load(open(p))
        if kernel in d:
            return d[kernel].get("dram_bytes_per_launch_256cubed")
        if d.get("kernel") == kernel:
            return d.get("dram_bytes_per_launch_256cubed")
    return None


# ----------------------------------------------------------------------------- our arm
FIXTURE_POS = os.path.join(ROOT, "fixtures", "performance_testing_hematocrit_33_RBC.pos")

# algorithmic bytes per lattice update of each lattice kernel
KERNEL_B_LU = {"k_collide_stream": 304.0, "k_collide_tau1": B_LU_TAU1, "k_moments": B_MOM, "k_moment_step": 160.0, "k_moment_tile": 160.0}


def unit_rows(workload):
    """.pos rows of one 256^3 unit and the sha256 of their source"""
    import hashlib
    if workload == "synthetic":
        rows = synthetic_rows()
        return rows, "sha256:" + hashlib.sha256(np.ascontiguousarray(rows).tobytes()).hexdigest(), "synthetic crystal packing (seed 1234)"
    from hemocell_b200 import lib as H
    rows = H.read_pos(FIXTURE_POS)
    return rows, "sha256:" + hashlib.sha256(open(FIXTURE_POS, "rb").read()).hexdigest(), \
        "cases/performance_testing/hematocrit_33/RBC.pos (reference data file, fixtures/)"


def parity_check(args, dist, H, par):
    """Before the timed region: a small slab-decomposed problem on the SAME code path as the benchmark (fully periodic,
    tau = 1, body force, RBCs sitting on every slab face incl. the periodic one, velocity cadence as benchmarked), 20
    iterate() steps on all ranks against the CPU oracle on rank 0 (checker leg; outside the timed region)."""
    rank, world = args.rank, args.world
    nxl, ny, nz, steps = 32, 28, 28, 20
    nx = nxl * world
    ct = H.HostCellType(H.MODEL_RBC, H.RBC_FROM_SPHERE, par, H.RBC_MATERIAL)
    um = DX / 1e-6
    # one RBC centred on every slab face (the first one on the periodic face x = 0), one in the bulk of slab 0
    rows = [(((k * nxl) % nx + (0.4 if k else 0.0)) * um, (13.0 + 0.7 * (k % 3)) * um, (14.0 - 0.5 * (k % 2)) * um, 90.0, 20.0 * k, 0.0) for k in range(world)]
    rows.append((16.3 * um, 14.2 * um, 13.6 * um, 70.0, 20.0, 10.0))
    rows = np.array(rows)
    # the reader's placement rule prunes cells that stick out of the domain: place on a domain shifted by half a slab
    # and shift back, so that the face cells (incl. the one on the periodic face) survive with unwrapped coordinates
    shifted = rows.copy(); shifted[:, 0] += 0.5 * nxl * um
    cells, ids = ct.place(shifted, DX, (nx + nxl, ny, nz))
    cells = cells - np.array([0.5 * nxl, 0.0, 0.0])
    assert len(ids) == len(rows)
    body = body_force(par["nu_lbm"], UNIT_N)
    body = tuple(50.0 * b for b in body)
    ctx = H.Context(nx, ny, nz, (1, 1, 1), par["tau"], device=args.local_rank, rank=rank, n_ranks=world)
    if world > 1:
        import torch
        idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idbuf.copy_(torch.frombuffer(bytearray(H.Context.unique_id()), dtype=torch.uint8))
        dist.broadcast(idbuf, 0)
        ctx.comm_init(bytes(idbuf.cpu().numpy().tobytes()))
    ctx.set_flags(np.zeros(ctx.Nl, dtype=np.uint8))
    ctx.set_body_force(body); ctx.set_force_limit(par["f_limit"])
    t = ct.add_to(ctx)
    if world > 1:
        ctx.set_exchange(4.0, 5, 1.0)
    ctx.add_cells(t, cells, ids)
    ctx.set_timescales(args.cadence, 1, 1); ctx.set_material_timescale(t, 5 * args.cadence)
    ctx.iterate(steps)
    pop = ctx.lattice_download(H.LAT_POP).reshape(19, ctx.nxl, ny, nz)
    pos = ctx.cells_download(H.P_POS).reshape(-1, ct.V, 3)
    cid, _, alive = ctx.cells_info()
    mine = {int(c): pos[k] for k, (c, a) in enumerate(zip(cid, alive)) if c >= 0 and a}
    mode = ctx.lattice_mode() if hasattr(ctx, "lattice_mode") else None
    ctx.close()
    if world > 1:
        gathered = [None] * world if rank == 0 else None
        dist.gather_object((ctx.x0, pop, mine), gathered, dst=0)
    else:
        gathered = [(0, pop, mine)]
    if rank != 0:
        return None
    import oracle as O
    from oracle import mesh as M
    opar = M.Parameters(DX, -1.0)
    oct_ = O.rbc_celltype(opar)
    dom = O.make_domain(nx, ny, nz, (1, 1, 1), opar.tau)
    sim = O.OracleSim(dom, np.zeros(nx * ny * nz, dtype=np.uint8), opar.f_limit, body)
    sim.vel_timescale = args.cadence
    sim.add_celltype(oct_, 5 * args.cadence)
    sim.add_cells(0, cells, ids)
    for _ in range(steps):
        sim.iterate()
    ref_pop = sim.pop.reshape(19, nx, ny, nz)
    ref_pos = sim.pos.reshape(-1, ct.V, 3)
    scale = float(np.abs(ref_pop).max())
    e_pop, e_pos, seen = 0.0, 0.0, set()
    for x0, gp, gm in gathered:
        e_pop = max(e_pop, float(np.abs(gp - ref_pop[:, x0:x0 + gp.shape[1]]).max()) / scale)
        for c, pp in gm.items():
            k = int(np.where(sim.cell_id == c)[0][0])
            e_pos = max(e_pos, float(np.abs(pp - ref_pos[k]).max()) / float(np.abs(ref_pos[k]).max()))
            seen.add(c)
    ok = e_pop <= 1e-10 and e_pos <= 1e-10 and seen == set(int(i) for i in ids)
    return {"max_rel": max(e_pop, e_pos), "max_rel_populations": e_pop, "max_rel_positions": e_pos, "tolerance": 1e-10, "ok": bool(ok),
            "against": "CPU oracle (oracle/hemo_oracle.c), rank 0", "steps": steps, "lattice": [nx, ny, nz], "slabs": world,
            "cells": int(len(ids)), "cells_on_slab_faces": world, "every_cell_found": seen == set(int(i) for i in ids),
            "lattice_update": mode}


def run_cuda(args):
    rank, world = args.rank, args.world
    from hemocell_b200 import lib as H
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(args.local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", args.local_rank))
    numa_cpus = bind_to_gpu_numa_node(args.local_rank)
    par = H.parameters(DX, -1.0)
    pcheck = None if args.no_parity_check else parity_check(args, dist, H, par)
    ct = H.HostCellType(H.MODEL_RBC, H.RBC_FROM_SPHERE, par, H.RBC_MATERIAL)
    rows, pos_hash, pos_src = unit_rows(args.workload)
    nx = UNIT_N * world
    # every unit gets the same cells, shifted along x (examples/cube/preprocess/analysis.py:cell_positions)
    unit_cells, unit_ids = ct.place(rows, DX, (UNIT_N, UNIT_N, UNIT_N))
    n_unit = len(unit_ids)
    ctx = H.Context(nx, UNIT_N, UNIT_N, (1, 1, 1), par["tau"], device=args.local_rank, rank=rank, n_ranks=world)
    if world > 1:
        import torch
        idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idbuf.copy_(torch.frombuffer(bytearray(H.Context.unique_id()), dtype=torch.uint8))
        dist.broadcast(idbuf, 0)
        ctx.comm_init(bytes(idbuf.cpu().numpy().tobytes()))
    ctx.set_flags(np.zeros(ctx.Nl, dtype=np.uint8))
    ctx.set_body_force(body_force(par["nu_lbm"], UNIT_N))
    ctx.set_force_limit(par["f_limit"])
    t = ct.add_to(ctx)
    if world > 1:
        # own unit plus the cells of the neighbouring units that reach into this slab's hold region (the library keeps
        # what it holds); spare slots for later arrivals: 15 % of the held cells
        ctx.set_exchange(4.0, 20, 0.15)
        parts_c, parts_i = [unit_cells + np.array([rank * UNIT_N, 0.0, 0.0])], [unit_ids + rank * len(rows)]
        lo, hi = unit_cells[:, :, 0].min(1), unit_cells[:, :, 0].max(1)
        for u, sel in (((rank - 1) % world, hi > UNIT_N - 12.0), ((rank + 1) % world, lo < 12.0)):
            if u == rank:
                continue
            parts_c.append(unit_cells[sel] + np.array([u * UNIT_N, 0.0, 0.0])); parts_i.append(unit_ids[sel] + u * len(rows))
        cells, ids = np.concatenate(parts_c), np.concatenate(parts_i)
    else:
        cells, ids = unit_cells, unit_ids
    pos_host = pinned_empty(cells.size)
    pos_host[:] = cells.reshape(-1)
    ctx.add_cells(t, pos_host.reshape(cells.shape), ids)
    ctx.set_timescales(args.cadence, 1, 1)
    ctx.set_material_timescale(t, 20)
    n_cells_global = n_unit * world
    nodes = nx * UNIT_N * UNIT_N

    def barrier():
        ctx.synchronize()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        import torch
        tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    # ---- device-resident leg: W warm-up steps, then exactly K timed steps
    ctx.iterate(args.warmup)
    launches0 = ctx.launch_count()
    ctx.timers_enable(True); ctx.timers_reset()
    sampler = ClockSampler(args.local_rank) if rank == 0 else None
    barrier()
    t0 = time.time()
    ms = ctx.iterate_timed(args.steps)            # CUDA events on the launching stream
    barrier()
    t1 = time.time()
    ms = max_over_ranks(ms)
    clocks = sampler.stop(t0, t1) if sampler else None
    launches = ctx.launch_count() - launches0
    timers = ctx.timers()
    ctx.timers_enable(False)
    alive_cells = ctx.count()[0]

    # ---- end-to-end leg through the C ABI with HOST buffers (pinned), the loop of the case file
    # (performance_testing.cpp:126-133): cell state H2D, then K x { iterate(); setExternalVector(force) (24 B host argument);
    # cell count of the step D2H into pinned memory, not waited for }, one wait, positions and forces D2H (writeOutput)
    npart = ctx.capacity()[1]
    out_pos = pinned_empty_f32(3 * npart); out_frc = pinned_empty_f32(3 * npart)   # what writeOutput stores: float32, SI units
    state_host = pinned_empty(3 * npart)          # the particle state as the host holds it (pinned)
    counts = pinned_empty(2 * args.steps)         # 2 int64 per step
    c_fp = C.POINTER(C.c_float)
    ctx.L.hcg_cells_download(ctx.h, C.c_int32(H.P_POS), state_host.ctypes.data_as(H.c_dp))
    ctx.set_iteration(0)
    bf = body_force(par["nu_lbm"], UNIT_N)
    barrier()
    te0 = time.time()
    ctx.cells_upload(H.P_POS, state_host)
    for k in range(args.steps):
        ctx.iterate_async(1)
        ctx.set_body_force(bf)                    # setExternalVector after every iterate
        ctx.count_async(counts.ctypes.data + 16 * k)
    ctx.synchronize()
    ctx._ck(ctx.L.hcg_cells_download_f32(ctx.h, C.c_int32(H.P_POS), C.c_double(DX), out_pos.ctypes.data_as(c_fp)))
    ctx._ck(ctx.L.hcg_cells_download_f32(ctx.h, C.c_int32(H.P_FORCE), C.c_double(par["df"]), out_frc.ctypes.data_as(c_fp)))
    barrier()
    e2e_ms = max_over_ranks((time.time() - te0) * 1e3)
    h2d = (8 * 3 * npart) / args.steps + 24
    d2h = (2 * 4 * 3 * npart) / args.steps + 16
    last_count = int(counts.view(np.int64)[2 * (args.steps - 1)])
    if dist is not None:
        import torch
        tt = torch.tensor([last_count], dtype=torch.int64, device="cuda")
        dist.all_reduce(tt)
        last_count = int(tt.item())

    if rank != 0:
        ctx.close()
        return None
    peak, peak_src = measured_peak()
    nodes_local = UNIT_N ** 3
    # ---- rooflines.  (1) the whole lattice update = every lattice kernel of a step, against SURVEY 8(d)'s 304 B/LU
    lat = {k[len("kernel:"):]: v for k, v in timers.items() if k.startswith("kernel:") and k[len("kernel:"):] in KERNEL_B_LU}
    lat_ms_per_step = sum(v[0] for v in lat.values()) / args.steps
    lat_ach = B_LU * nodes_local / (lat_ms_per_step * 1e-3) / 1e9 if lat_ms_per_step > 0 else None
    kernels = {}
    for name, (kms, calls) in lat.items():
        if not calls:
            continue
        b = KERNEL_B_LU[name]
        if name == "k_moments" and "k_collide_tau1" in lat:
            b = B_MOM_TAU1
        ach = b * nodes_local / (kms / calls * 1e-3) / 1e9
        kernels[name] = {"bytes_per_lu": b, "launch_ms": kms / calls, "launches_timed": calls, "achieved": ach, "frac": ach / peak,
                         "traffic": ncu_traffic(name)}
    traffic = [kernels[k]["traffic"] for k in kernels]
    traffic_update = None
    if kernels and all(v is not None for v in traffic):
        traffic_update = sum(kernels[k]["traffic"] * kernels[k]["launches_timed"] for k in kernels) / args.steps
    lsp = n_unit * ct.V
    b_step = B_LU + (264.0 + 72.0 + 216.0 / args.cadence) * lsp / nodes_local
    step_ach = b_step * nodes_local / (ms / args.steps * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": nodes * args.steps / (ms * 1e-3) / 1e6, "unit": "MLUPS",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "cell_steps_per_s": (alive_cells if world == 1 else n_cells_global) * args.steps / (ms * 1e-3),
        "config": {"workload": f"cases/performance_testing unit: {nx}x{UNIT_N}x{UNIT_N} D3Q19 fp64, fully periodic, tau=1, "
                               f"body force, {n_cells_global} RBC (642 LSP each; {pos_src}), material every 20, "
                               f"velocity every {args.cadence}",
                   "lattice": [nx, UNIT_N, UNIT_N], "cells": n_cells_global, "lsp": n_cells_global * ct.V,
                   "cells_alive_after_run": int(last_count), "pos_rows": int(len(rows)), "pos_source": pos_src, "pos_sha256": pos_hash,
                   "hematocrit": n_unit * 81.116 / (UNIT_N * DX * 1e6) ** 3,
                   "velocity_cadence": args.cadence, "material_cadence": 20, "decomposition": f"{world} x-slabs",
                   "l2": "inputs (0.5 - 5 GB of lattice state per GPU) are far larger than the 126 MB L2; no flush needed"},
        "roofline": {"bound": "hbm", "kernel": " + ".join(sorted(kernels)) + " (whole lattice update)", "achieved": lat_ach, "peak": peak,
                     "unit": "GB/s", "frac": lat_ach / peak if lat_ach else None, "traffic": traffic_update,
                     "peak_source": peak_src, "bytes_per_lu": B_LU, "launch_ms": lat_ms_per_step,
                     "note": "achieved = 304 B/LU (SURVEY 8d contract: 19 populations read + written) x 256^3 / time of ALL lattice kernels of a step; "
                             "a fraction above 1 means the update moved fewer bytes than the contract (tau = 1: the populations are never stored)"},
        "roofline_kernels": kernels,
        "roofline_step": {"bound": "hbm", "bytes_per_lu": b_step, "achieved": step_ach, "peak": peak, "unit": "GB/s", "frac": step_ach / peak,
                          "formula": "304 + (264 + 72 + 216/c) N_LSP/N_nodes B/LU (SURVEY 8d)"},
        "kernel_ms_per_step": {k: v[0] / args.steps for k, v in timers.items()},
        "e2e": {"value": nodes * args.steps / (e2e_ms * 1e-3) / 1e6, "unit": "MLUPS",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps,
                "loop": "cell positions H2D (fp64); K x {iterate, body force, async cell-count D2H}; wait; positions + forces D2H as "
                        "writeOutput stores them (float32, SI)", "cpus_bound_to_gpu_numa_node": numa_cpus},
        "parity_check": pcheck,
        "gpu_launches": launches, "clocks": clocks,
    }
    ctx.close()
    return line


# ----------------------------------------------------------------------------- BASELINE configs 2 - 4 (strong scaling over x-slabs)
def read_pos_multi(path, counts):
    """a .pos file that holds several cell types one after the other (examples/pipeflow/initial_states/*/cells.pos: the
    counts on the first lines, then the rows of each type)"""
    tok = open(path).read().split()
    n = [int(tok[k]) for k in range(counts)]
    vals = np.array(tok[counts:counts + 6 * sum(n)], dtype=np.float64).reshape(-1, 6)
    out, at = [], 0
    for k in n:
        out.append(vals[at:at + k]); at += k
    return out


def case_spec(name, H):
    """domain, boundary nodes, driving force, cells and cadences of a BASELINE.json configuration, from the reference's case file"""
    fx = os.path.join(ROOT, "fixtures")
    if name == "pipeflow":
        # examples/pipeflow/pipeflow.cpp:51-146 on the D = 64 um, L = 128 um vessel of initial_states/D64_Ht21:
        # bounce-back outside the radius, x periodic, body force 8 nu (u_max / 2) / R^2 at Re 0.5, dt = 1e-7 (tau 1.82)
        par = H.parameters(DX, 1e-7)
        nx, ny, nz, R = 256, 130, 130, 64.0
        yy, zz = np.meshgrid(np.arange(ny), np.arange(nz), indexing="ij")
        wall = ((yy - 64.5) ** 2 + (zz - 64.5) ** 2) >= R * R
        fl = np.zeros((nx, ny, nz), dtype=np.uint8); fl[:, wall] = H.BOUNCEBACK
        u_max = 0.5 * par["nu_lbm"] / (2 * R)
        body = (8 * par["nu_lbm"] * (u_max * 0.5) / R / R, 0.0, 0.0)
        rbc_rows, plt_rows = read_pos_multi(os.path.join(fx, "pipeflow_D64_Ht21_cells.pos"), 2)
        return dict(label="examples/pipeflow, D64_Ht21 vessel", par=par, dims=(nx, ny, nz), periodic=(1, 0, 0), flags=fl, bc=None, body=body,
                    cells=[("RBC", rbc_rows, 0.0), ("PLT", plt_rows, 0.0)], material=20, velocity=5, rep=None,
                    pos=["pipeflow_D64_Ht21_cells.pos"])
    if name == "stenosis":
        # cases/stenosis/stenosis.cpp:37-74, 112-230: 600 x 348 x 160, bounce-back y / z faces + the analytic stenosis shape,
        # x periodic, body force dpdz_lbm, nu = 3e-6, dt = 1e-8 (tau 0.86), Ht20 cells, minimum wall distance 1 um
        par = H.parameters(DX, 1e-8, 3.0e-6)
        nx, ny, nz = 600, 348, 160
        rc, ytop, xbl = 15, 316, 100
        xtr, xcirc, ycirc = xbl + 2 * rc, xbl + rc, ytop - rc
        ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
        shape = ((ix - xcirc) ** 2 + (iy - ycirc) ** 2 <= rc * rc) | ((ix <= xtr) & (ix >= xbl) & (iy <= ycirc)) | \
                ((ix <= (iy - 514.16683048) / -1.60677134525) & (ix >= 127.73502714) & (iy <= 308.92584909))
        fl = np.zeros((nx, ny, nz), dtype=np.uint8)
        fl[shape] = H.BOUNCEBACK
        fl[:, :, 0] = H.BOUNCEBACK; fl[:, :, nz - 1] = H.BOUNCEBACK; fl[:, 0, :] = H.BOUNCEBACK; fl[:, ny - 1, :] = H.BOUNCEBACK
        flow_q = 1800.0 * 130e-6 * 80e-6 * 80e-6 / 6
        dpdz = flow_q * 12 * 3.0e-3 / (80e-6 ** 3 * 130e-6)
        body = (dpdz * DX * DX * par["dt"] ** 2 / par["dm"], 0.0, 0.0)
        return dict(label="cases/stenosis, Ht20", par=par, dims=(nx, ny, nz), periodic=(1, 0, 0), flags=fl, bc=None, body=body,
                    cells=[("RBC", H.read_pos(os.path.join(fx, "stenosis_Ht20_RBC.pos")), 1.0), ("PLT", H.read_pos(os.path.join(fx, "stenosis_Ht20_PLT.pos")), 0.0)],
                    material=10, velocity=10, rep=None, pos=["stenosis_Ht20_RBC.pos", "stenosis_Ht20_PLT.pos"])
    if name == "cube":
        # examples/cube/cube.cpp:48-133: 100^3, x periodic, bounce-back y planes, moving z walls, tau = 1; the reference seeds it
        # with the irreproducible tools/packCells, here a seeded lattice packing at ~30 % hematocrit (RBC + PLT)
        par = H.parameters(DX, -1.0)
        n, fl, bc, rbc_rows, plt_rows = cube_setup(H, par)
        return dict(label="examples/cube, seeded packing", par=par, dims=(n, n, n), periodic=(1, 0, 0), flags=fl, bc=bc, body=(0.0, 0.0, 0.0),
                    cells=[("RBC", rbc_rows, 0.0), ("PLT", plt_rows, 0.0)], material=20, velocity=5, rep=None, pos=[])
    raise ValueError(name)


def case_parity_check(args, dist, H, spec):
    """a small problem with the boundary kinds, relaxation time, cell types and cadences of the case (and repulsion when the case
    has it), cut into the same number of slabs, 20 iterate() steps against the CPU oracle on rank 0"""
    rank, world = args.rank, args.world
    par = spec["par"]
    nxl, ny, nz, steps = 32, 30, 30, 20
    nx = nxl * world
    fl = np.zeros((nx, ny, nz), dtype=np.uint8)
    bc = np.zeros((6, 3))
    if spec["bc"] is not None:                            # cube-like: bounce-back y planes, moving z walls
        fl[:, :, 0] = H.VEL_ZN; fl[:, :, nz - 1] = H.VEL_ZP; fl[:, 0, :] = H.BOUNCEBACK; fl[:, ny - 1, :] = H.BOUNCEBACK
        bc[4] = (0.02, 0, 0); bc[5] = (-0.02, 0, 0)
    else:                                                 # vessel-like: bounce-back outside a cylinder along x
        yy, zz = np.meshgrid(np.arange(ny), np.arange(nz), indexing="ij")
        fl[:, ((yy - 14.5) ** 2 + (zz - 14.5) ** 2) >= 13.5 ** 2] = H.BOUNCEBACK
    um = DX / 1e-6
    rbc = H.HostCellType(H.MODEL_RBC, H.RBC_FROM_SPHERE, par, H.RBC_MATERIAL)
    plt = H.HostCellType(H.MODEL_PLT, H.ELLIPSOID_FROM_SPHERE, par, H.PLT_MATERIAL, H.PLT_INNER_EDGES)
    rrows = np.array([(((k * nxl) % nx + (0.3 if k else 0.0)) * um, 14.5 * um, 14.5 * um, 0.0, 90.0, 15.0 * k) for k in range(world)])
    prows = np.array([(16.0 * um, 14.5 * um, 14.5 * um, 30.0, 0.0, 0.0)])
    fl_big = np.concatenate([fl, fl[:nxl]], axis=0)       # place on a domain shifted by half a slab: the face cells stay whole
    def place(ct, rows, id0):
        sh = rows.copy(); sh[:, 0] += 0.5 * nxl * um
        cells, ids = ct.place(sh, DX, (nx + nxl, ny, nz), fl_big.reshape(-1), 0.0, id0)
        return cells - np.array([0.5 * nxl, 0.0, 0.0]), ids
    rc, rid = place(rbc, rrows, 0); pc, pid = place(plt, prows, len(rrows))
    body = tuple(20.0 * b for b in spec["body"]) if any(spec["body"]) else (0.0, 0.0, 0.0)
    ctx = H.Context(nx, ny, nz, spec["periodic"], par["tau"], device=args.local_rank, rank=rank, n_ranks=world)
    if world > 1:
        import torch
        idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idbuf.copy_(torch.frombuffer(bytearray(H.Context.unique_id()), dtype=torch.uint8))
        dist.broadcast(idbuf, 0)
        ctx.comm_init(bytes(idbuf.cpu().numpy().tobytes()))
    ctx.set_flags(np.ascontiguousarray(fl[ctx.x0:ctx.x0 + ctx.nxl]))
    for o in range(6):
        ctx.set_bc_velocity(o, bc[o])
    ctx.set_body_force(body); ctx.set_force_limit(par["f_limit"])
    t0, t1 = rbc.add_to(ctx), plt.add_to(ctx)
    if world > 1:
        ctx.set_exchange(4.0, 5, 1.0)
    ctx.add_cells(t0, rc, rid); ctx.add_cells(t1, pc, pid)
    vel = min(spec["velocity"], 5)
    ctx.set_timescales(vel, vel, vel); ctx.set_material_timescale(t0, vel); ctx.set_material_timescale(t1, vel)
    if spec["rep"]:
        ctx.set_repulsion(True, spec["rep"]["k"], spec["rep"]["cut"]); ctx.set_wall_repulsion(True, spec["rep"]["k"], spec["rep"]["cut"])
    ctx.iterate(steps)
    pop = ctx.lattice_download(H.LAT_POP).reshape(19, ctx.nxl, ny, nz)
    cid, ctp, alive = ctx.cells_info()
    pos = ctx.cells_download(H.P_POS)
    mine, at = {}, 0
    for c, tp, al in zip(cid, ctp, alive):
        V = rbc.V if tp == t0 else plt.V
        if c >= 0 and al:
            mine[int(c)] = pos.reshape(-1, 3)[at:at + V].copy()
        at += V
    x0 = ctx.x0
    ctx.close()
    if world > 1:
        gathered = [None] * world if rank == 0 else None
        dist.gather_object((x0, pop, mine), gathered, dst=0)
    else:
        gathered = [(x0, pop, mine)]
    if rank != 0:
        return None
    import oracle as O
    from oracle import mesh as M
    opar = M.Parameters(DX, par["dt"] if spec["par"]["tau"] != 1.0 else -1.0, nu_p=par.get("nu_p", 1.1e-6))
    ort, opt = O.rbc_celltype(opar), O.plt_celltype(opar)
    dom = O.make_domain(nx, ny, nz, spec["periodic"], opar.tau, bc)
    sim = O.OracleSim(dom, fl.reshape(-1), opar.f_limit, body)
    sim.vel_timescale = vel
    sim.add_celltype(ort, vel); sim.add_celltype(opt, vel)
    sim.add_cells(0, rc, rid); sim.add_cells(1, pc, pid)
    if spec["rep"]:
        sim.rep_enabled = sim.wall_enabled = True; sim.rep_timescale = sim.wall_timescale = vel
        sim.rep_k = sim.wall_k = spec["rep"]["k"]; sim.rep_cutoff = sim.wall_cutoff = spec["rep"]["cut"]
    for _ in range(steps):
        sim.iterate()
    ref_pop = sim.pop.reshape(19, nx, ny, nz)
    off = sim._offsets()
    scale = float(np.abs(ref_pop).max())
    e_pop, e_pos, seen = 0.0, 0.0, set()
    for gx0, gp, gm in gathered:
        e_pop = max(e_pop, float(np.abs(gp - ref_pop[:, gx0:gx0 + gp.shape[1]]).max()) / scale)
        for c, pp in gm.items():
            k = int(np.where(sim.cell_id == c)[0][0])
            rp = sim.pos[off[k]:off[k + 1]]
            e_pos = max(e_pos, float(np.abs(pp - rp).max()) / float(np.abs(rp).max())); seen.add(c)
    want = set(int(i) for i in list(rid) + list(pid))
    # 20 coupled steps: the order of the spreading atomics differs from the oracle's loop and the stiff membranes amplify that
    # round-off in the populations (tests/test_gpu_multi.py uses the same 1e-9 for its 120-step runs); positions stay at 1e-10
    ok = e_pop <= 1e-9 and e_pos <= 1e-10 and seen == want
    return {"max_rel": max(e_pop, e_pos), "max_rel_populations": e_pop, "max_rel_positions": e_pos,
            "tolerance": {"populations": 1e-9, "positions": 1e-10}, "ok": bool(ok),
            "against": "CPU oracle (oracle/hemo_oracle.c), rank 0", "steps": steps, "lattice": [nx, ny, nz], "slabs": world,
            "cells": len(want), "every_cell_found": seen == want,
            "setup": "reduced problem with the case's boundary kinds, tau, cell types, cadences" + (" and both repulsions" if spec["rep"] else "")}


def run_case(args):
    """BASELINE.json configs[1..3] (examples/cube, examples/pipeflow, cases/stenosis): the whole domain of the case cut into
    x-slabs over the N ranks (strong scaling), same measurements as the main line"""
    import hashlib
    rank, world = args.rank, args.world
    from hemocell_b200 import lib as H
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(args.local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", args.local_rank))
    bind_to_gpu_numa_node(args.local_rank)
    spec = case_spec(args.workload, H)
    par = spec["par"]
    if args.repulsion:
        # HemoCell::setRepulsion / enableBoundaryParticles with the constants of examples/pipeflow/config.xml (kRep 2e-22, cut-off 0.7 um)
        spec["rep"] = dict(k=2e-22 / par["df"], cut=0.7e-6 / DX)
    pcheck = None if args.no_parity_check else case_parity_check(args, dist, H, spec)
    nx, ny, nz = spec["dims"]
    fl = spec["flags"]
    ctx = H.Context(nx, ny, nz, spec["periodic"], par["tau"], device=args.local_rank, rank=rank, n_ranks=world)
    if world > 1:
        import torch
        idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idbuf.copy_(torch.frombuffer(bytearray(H.Context.unique_id()), dtype=torch.uint8))
        dist.broadcast(idbuf, 0)
        ctx.comm_init(bytes(idbuf.cpu().numpy().tobytes()))
    ctx.set_flags(np.ascontiguousarray(fl[ctx.x0:ctx.x0 + ctx.nxl]))
    if spec["bc"] is not None:
        for o in range(6):
            ctx.set_bc_velocity(o, spec["bc"][o])
    ctx.set_body_force(spec["body"]); ctx.set_force_limit(par["f_limit"])
    types, n_cells, n_lsp, id0 = [], {}, 0, 0
    if world > 1:
        ctx.set_exchange(4.0, 20, 0.3)
    for cname, rows, min_dist in spec["cells"]:
        ct = (H.HostCellType(H.MODEL_RBC, H.RBC_FROM_SPHERE, par, H.RBC_MATERIAL) if cname == "RBC" else
              H.HostCellType(H.MODEL_PLT, H.ELLIPSOID_FROM_SPHERE, par, H.PLT_MATERIAL, H.PLT_INNER_EDGES))
        t = ct.add_to(ctx)
        cells, ids = ct.place(rows, DX, (nx, ny, nz), fl.reshape(-1), min_dist, id0)
        id0 += len(rows)
        ctx.add_cells(t, cells, ids)
        ctx.set_material_timescale(t, spec["material"])
        types.append(ct); n_cells[cname] = len(ids); n_lsp += len(ids) * ct.V
    ctx.set_timescales(spec["velocity"], spec["material"], spec["material"])
    if spec["rep"]:
        ctx.set_repulsion(True, spec["rep"]["k"], spec["rep"]["cut"]); ctx.set_wall_repulsion(True, spec["rep"]["k"], spec["rep"]["cut"])
    nodes = nx * ny * nz
    cells_total = sum(n_cells.values())

    def barrier():
        ctx.synchronize()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(v):
        if dist is None:
            return v
        import torch
        tt = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    ctx.iterate(args.warmup)
    launches0 = ctx.launch_count()
    ctx.timers_enable(True); ctx.timers_reset()
    sampler = ClockSampler(args.local_rank) if rank == 0 else None
    barrier(); t0 = time.time()
    ms = ctx.iterate_timed(args.steps)
    barrier(); t1 = time.time()
    ms = max_over_ranks(ms)
    clocks = sampler.stop(t0, t1) if sampler else None
    launches = ctx.launch_count() - launches0
    timers = ctx.timers(); ctx.timers_enable(False)
    # end-to-end: the loop of the case file through the C ABI with host buffers (see run_cuda)
    npart = ctx.capacity()[1]
    state_host = pinned_empty(max(3 * npart, 1)); out_pos = pinned_empty_f32(3 * npart); out_frc = pinned_empty_f32(3 * npart)
    c_fp = C.POINTER(C.c_float)
    counts = pinned_empty(2 * args.steps)
    ctx.L.hcg_cells_download(ctx.h, C.c_int32(H.P_POS), state_host.ctypes.data_as(H.c_dp))
    barrier(); te0 = time.time()
    ctx.cells_upload(H.P_POS, state_host[:3 * npart])
    for k in range(args.steps):
        ctx.iterate_async(1)
        ctx.set_body_force(spec["body"])
        ctx.count_async(counts.ctypes.data + 16 * k)
    ctx.synchronize()
    ctx._ck(ctx.L.hcg_cells_download_f32(ctx.h, C.c_int32(H.P_POS), C.c_double(DX), out_pos.ctypes.data_as(c_fp)))
    ctx._ck(ctx.L.hcg_cells_download_f32(ctx.h, C.c_int32(H.P_FORCE), C.c_double(par["df"]), out_frc.ctypes.data_as(c_fp)))
    barrier()
    e2e_ms = max_over_ranks((time.time() - te0) * 1e3)
    alive = int(counts.view(np.int64)[2 * (args.steps - 1)])
    if dist is not None:
        import torch
        tt = torch.tensor([alive], dtype=torch.int64, device="cuda"); dist.all_reduce(tt); alive = int(tt.item())
    nodes_local = ctx.nxl * ny * nz
    if rank != 0:
        ctx.close()
        return None
    peak, peak_src = measured_peak()
    lat = {k[len("kernel:"):]: v for k, v in timers.items() if k.startswith("kernel:") and k[len("kernel:"):] in KERNEL_B_LU}
    lat_ms = sum(v[0] for v in lat.values()) / args.steps
    lat_ach = B_LU * nodes_local / (lat_ms * 1e-3) / 1e9 if lat_ms > 0 else None
    kernels = {}
    for kname, (kms, calls) in lat.items():
        if calls:
            b = B_MOM_TAU1 if (kname == "k_moments" and "k_collide_tau1" in lat) else KERNEL_B_LU[kname]
            ach = b * nodes_local / (kms / calls * 1e-3) / 1e9
            kernels[kname] = {"bytes_per_lu": b, "launch_ms": kms / calls, "launches_timed": calls, "achieved": ach, "frac": ach / peak}
    b_step = B_LU + (264.0 + 72.0 + 216.0 / spec["velocity"]) * n_lsp / nodes
    step_ach = b_step * nodes_local / (ms / args.steps * 1e-3) / 1e9
    sha = {f: "sha256:" + hashlib.sha256(open(os.path.join(ROOT, "fixtures", f), "rb").read()).hexdigest() for f in spec["pos"]}
    line = {
        "metric": METRIC, "value": nodes * args.steps / (ms * 1e-3) / 1e6, "unit": "MLUPS", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "cell_steps_per_s": cells_total * args.steps / (ms * 1e-3),
        "config": {"workload": f"{spec['label']}: {nx}x{ny}x{nz} D3Q19 fp64, tau={par['tau']:.3f}, " +
                               " + ".join(f"{v} {k}" for k, v in n_cells.items()) + f", material every {spec['material']}, velocity every {spec['velocity']}" +
                               (", cell-cell + wall repulsion" if spec["rep"] else ""),
                   "lattice": [nx, ny, nz], "cells": cells_total, "cells_by_type": n_cells, "lsp": n_lsp, "cells_alive_after_run": alive,
                   "fluid_fraction": float((fl == 0).mean()), "pos_sha256": sha, "velocity_cadence": spec["velocity"],
                   "material_cadence": spec["material"], "repulsion": bool(spec["rep"]), "decomposition": f"{world} x-slabs of the one domain",
                   "l2": "lattice state per GPU far larger than the 126 MB L2 (cube: 0.3 GB); no flush"},
        "roofline": {"bound": "hbm", "kernel": " + ".join(sorted(kernels)) + " (whole lattice update, averaged over the cadence)", "achieved": lat_ach,
                     "peak": peak, "unit": "GB/s", "frac": lat_ach / peak if lat_ach else None, "traffic": None, "peak_source": peak_src,
                     "bytes_per_lu": B_LU, "launch_ms": lat_ms},
        "roofline_kernels": kernels,
        "roofline_step": {"bound": "hbm", "bytes_per_lu": b_step, "achieved": step_ach, "peak": peak, "unit": "GB/s", "frac": step_ach / peak,
                          "formula": "304 + (264 + 72 + 216/c) N_LSP/N_nodes B/LU (SURVEY 8d)"},
        "kernel_ms_per_step": {k: v[0] / args.steps for k, v in timers.items()},
        "e2e": {"value": nodes * args.steps / (e2e_ms * 1e-3) / 1e6, "unit": "MLUPS", "h2d_bytes_per_step": 8 * 3 * npart / args.steps + 24,
                "d2h_bytes_per_step": 2 * 4 * 3 * npart / args.steps + 16, "ms_per_step": e2e_ms / args.steps},
        "parity_check": pcheck, "gpu_launches": launches, "clocks": clocks}
    ctx.close()
    return line


# ----------------------------------------------------------------------------- CPU arm
def run_cpu(steps, warmup, cadence, budget_s=150.0, workload="performance_testing"):
    """The reference cannot be built (Palabos/MPI/HDF5 absent): time the CPU oracle (a port) with
    OpenMP on the host cores, on a bounded sub-box of the same workload (same packing density)."""
    import oracle as O
    from oracle import mesh as M
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    O.set_parallel(max(cores, 1) if cores > 1 else 1)      # explicit thread count: torchrun exports OMP_NUM_THREADS=1
    par = M.Parameters(DX, -1.0)
    ct = O.rbc_celltype(par)

    def make(n):
        # the same .pos rows on an n^3 sub-box of the unit: the reader's rule keeps the cells that lie wholly inside it
        rows = synthetic_rows(n) if workload == "synthetic" else M.read_pos(FIXTURE_POS)
        if workload != "synthetic":
            rows = rows[np.all(rows[:, :3] < n * DX / 1e-6 + 4.0, axis=1)]
        cells, ids = M.place_cells(ct.verts, rows, DX, (n, n, n))
        dom = O.make_domain(n, n, n, (1, 1, 1), par.tau)
        sim = O.OracleSim(dom, np.zeros(n ** 3, dtype=np.uint8), par.f_limit, body_force(par.nu_lbm, UNIT_N))
        sim.vel_timescale = cadence
        sim.add_celltype(ct, 20)
        sim.add_cells(0, cells, ids)
        return sim, len(ids)

    # pick the sample so that (steps + warmup) iterations fit the budget
    sim, ncell = make(64)
    t = time.time(); sim.iterate(); sim.iterate(); per64 = (time.time() - t) / 2
    n = 64
    for cand in (128, 96):
        if per64 * (cand / 64.0) ** 3 * (steps + warmup) <= budget_s:
            n = cand
            break
    if n != 64:
        sim, ncell = make(n)
    for _ in range(warmup):
        sim.iterate()
    t0 = time.time()
    for _ in range(steps):
        sim.iterate()
    dt = time.time() - t0
    return {"value": n ** 3 * steps / dt / 1e6, "unit": "MLUPS", "cores": cores, "kind": "port",
            "sample": f"{n}^3 periodic sub-box of the same workload, {ncell} RBC, {steps} iterate() steps, "
                      f"OpenMP on {cores} threads, CPU oracle (restatement, not the reference build)",
            "ms_per_step": dt / steps * 1e3, "cell_steps_per_s": ncell * steps / dt}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--cadence", type=int, default=1, help="velocity interpolation every n steps (stepParticleEvery)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="performance_testing", choices=["performance_testing", "synthetic", "cube", "pipeflow", "stenosis"],
                    help="performance_testing = the reference's unit with its own RBC.pos (default; weak scaling); synthetic = round 1's crystal "
                         "packing; cube / pipeflow / stenosis = BASELINE configs 1 - 3 (one domain cut into N slabs: strong scaling)")
    ap.add_argument("--repulsion", action="store_true", help="cube / pipeflow / stenosis: cell-cell and wall repulsion on (kRep 2e-22, 0.7 um)")
    ap.add_argument("--no-parity-check", action="store_true")
    args = ap.parse_args()
    args.rank = int(os.environ.get("RANK", "0"))
    args.world = int(os.environ.get("WORLD_SIZE", "1"))
    args.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        if args.rank != 0:
            return
        cb = run_cpu(args.steps, args.warmup, args.cadence, workload=args.workload)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "MLUPS", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "cases/performance_testing unit (bounded CPU sample, see cpu_baseline.sample)",
                           "velocity_cadence": args.cadence, "material_cadence": 20},
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return
    if args.workload in ("cube", "pipeflow", "stenosis"):
        line = run_case(args)
        if args.rank == 0:
            print(json.dumps(line), flush=True)
        return
    line = run_cuda(args)
    if args.rank != 0:
        return
    if args.world == 1 and not args.no_cpu_baseline:
        try:
            os.sched_setaffinity(0, _AFFINITY0)           # the CPU baseline gets every host core again (the GPU leg ran NUMA-bound)
        except Exception:
            pass
        line["cpu_baseline"] = run_cpu(6, 1, args.cadence, budget_s=30.0, workload=args.workload)
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
