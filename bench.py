#!/usr/bin/env python
"""bench.py -- throughput of the HemoCell per-timestep IB-LBM hot path on B200.

Metric (BASELINE.json): MLUPS (D3Q19 fp64, with RBCs) and cell-steps/s.
Workload at N GPUs: the cases/performance_testing weak-scaling unit, 256^3 lattice nodes per
GPU (domain 256*N x 256 x 256, slabs along x), fully periodic, tau = 1, body force (f,f,f),
RBCs seeded to ~33 % hematocrit per unit (synthetic, seeded lattice packing), material update
every 20 steps, velocity interpolation every `--cadence` steps (1 = configs/, 5 = configs_timestep_5/).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # CPU oracle on the host cores
    python bench.py --workload cube                          # extra line: examples/cube (BASELINE configs[1]), one GPU

The JSON line carries `roofline` for the collision kernel that ran (generic pull kernel: 304 B/LU; tau = 1 kernel after a
moments pass: 216 B/LU; the opt-in moment-only update, HCG_MOMENT_ONLY=1: 160 B/LU), `roofline_moments` for the moments pass, `e2e` (host buffers, copies inside the timed region),
`cpu_baseline` (rank 0, N = 1), `gpu_launches` and the SM clocks sampled during the timed region.

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
def run_cuda(args):
    rank, world = args.rank, args.world
    from hemocell_b200 import lib as H
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(args.local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", args.local_rank))
    par = H.parameters(DX, -1.0)
    ct = H.HostCellType(H.MODEL_RBC, H.RBC_FROM_SPHERE, par, H.RBC_MATERIAL)
    rows = synthetic_rows()
    nx = UNIT_N * world
    # every unit gets the same packing, shifted along x (examples/cube/preprocess/analysis.py:cell_positions)
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
        # own unit plus the neighbouring units; the library keeps the cells within its hold region
        ctx.set_exchange(4.0, 20, 0.3)
        units = sorted({(rank - 1) % world, rank, (rank + 1) % world})
        cells = np.concatenate([unit_cells + np.array([u * UNIT_N, 0.0, 0.0]) for u in units])
        ids = np.concatenate([unit_ids + u * len(rows) for u in units])
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

    # ---- end-to-end leg through the C ABI with HOST buffers (pinned): state upload, K x
    # (body force H2D + iterate(1) + cell-count D2H), final read-back of positions and forces
    npart = ctx.capacity()[1]
    out_pos = pinned_empty(3 * npart); out_frc = pinned_empty(3 * npart)
    state_host = pinned_empty(3 * npart)          # the particle state as the host holds it (pinned)
    ctx.L.hcg_cells_download(ctx.h, C.c_int32(H.P_POS), state_host.ctypes.data_as(H.c_dp))
    ctx.set_iteration(0)
    barrier()
    te0 = time.time()
    ctx.cells_upload(H.P_POS, state_host)
    bf = body_force(par["nu_lbm"], UNIT_N)
    for _ in range(args.steps):
        ctx.set_body_force(bf)                    # setExternalVector after every iterate (performance_testing.cpp:132-135)
        ctx.iterate(1)
        ctx.count()
    ctx.L.hcg_cells_download(ctx.h, C.c_int32(H.P_POS), out_pos.ctypes.data_as(H.c_dp))
    ctx.L.hcg_cells_download(ctx.h, C.c_int32(H.P_FORCE), out_frc.ctypes.data_as(H.c_dp))
    barrier()
    e2e_ms = max_over_ranks((time.time() - te0) * 1e3)
    h2d = (8 * 3 * npart) / args.steps + 24
    d2h = (2 * 8 * 3 * npart) / args.steps + 16

    if rank != 0:
        ctx.close()
        return None
    # dominant kernel = the collide-and-stream kernel that ran most: the generic pull kernel (19 r + 19 w = 304 B/LU,
    # the contract figure) or, at tau = 1 after a moments pass, k_collide_tau1 (32 B raw moments + 32 B force read,
    # 19 populations written = 216 B/LU)
    gen = timers.get("kernel:k_collide_stream", (0.0, 0)); t1 = timers.get("kernel:k_collide_tau1", (0.0, 0))
    mo = timers.get("kernel:k_moment_step", (0.0, 0))
    if mo[0] > max(gen[0], t1[0]):
        # opt-in moment-only update (HCG_MOMENT_ONLY=1): 64 B read + 96 B written per lattice update
        k1_name, (k1_ms, k1_calls), b_lu = "k_moment_step", mo, 160.0
    elif t1[0] > gen[0]:
        k1_name, (k1_ms, k1_calls), b_lu = "k_collide_tau1", t1, B_LU_TAU1
    else:
        k1_name, (k1_ms, k1_calls), b_lu = "k_collide_stream", gen, B_LU
    peak, peak_src = measured_peak()
    nodes_local = UNIT_N ** 3
    k1_avg_ms = k1_ms / max(k1_calls, 1)
    achieved = b_lu * nodes_local / (k1_avg_ms * 1e-3) / 1e9 if k1_calls else None
    mom_ms, mom_calls = timers.get("kernel:k_moments", (0.0, 0))
    b_mom = B_MOM_TAU1 if k1_name == "k_collide_tau1" else B_MOM
    mom_ach = b_mom * nodes_local / (mom_ms / max(mom_calls, 1) * 1e-3) / 1e9 if mom_calls else None
    line = {
        "metric": METRIC, "value": nodes * args.steps / (ms * 1e-3) / 1e6, "unit": "MLUPS",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "cell_steps_per_s": (alive_cells if world == 1 else n_cells_global) * args.steps / (ms * 1e-3),
        "config": {"workload": f"cases/performance_testing unit: {nx}x{UNIT_N}x{UNIT_N} D3Q19 fp64, fully periodic, tau=1, "
                               f"body force, {n_cells_global} RBC (642 LSP each, ~33% hematocrit), material every 20, "
                               f"velocity every {args.cadence}",
                   "lattice": [nx, UNIT_N, UNIT_N], "cells": n_cells_global, "lsp": n_cells_global * ct.V,
                   "velocity_cadence": args.cadence, "material_cadence": 20, "decomposition": f"{world} x-slabs",
                   "l2": "inputs (5.1 GB of populations per GPU) are far larger than the 126 MB L2; no flush needed"},
        "roofline": {"bound": "hbm", "kernel": k1_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if achieved else None, "traffic": ncu_traffic(k1_name),
                     "peak_source": peak_src, "bytes_per_lu": b_lu, "launch_ms": k1_avg_ms, "launches_timed": k1_calls},
        "roofline_moments": {"bound": "hbm", "kernel": "k_moments", "achieved": mom_ach, "peak": peak, "unit": "GB/s",
                             "frac": mom_ach / peak if mom_ach else None, "bytes_per_lu": b_mom,
                             "launch_ms": mom_ms / max(mom_calls, 1), "launches_timed": mom_calls},
        "kernel_ms_per_step": {k: v[0] / args.steps for k, v in timers.items()},
        "e2e": {"value": nodes * args.steps / (e2e_ms * 1e-3) / 1e6, "unit": "MLUPS",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": launches, "clocks": clocks,
    }
    ctx.close()
    return line


def run_cube(args):
    """extra line (not the default): examples/cube on one GPU.  A 10^6-node lattice is launch- and latency-bound on a
    B200, so this line says little about the kernels; it is here because BASELINE.json lists the configuration."""
    from hemocell_b200 import lib as H
    par = H.parameters(DX, -1.0)
    n, fl, bc, rbc_rows, plt_rows = cube_setup(H, par)
    rbc = H.HostCellType(H.MODEL_RBC, H.RBC_FROM_SPHERE, par, H.RBC_MATERIAL)
    plt = H.HostCellType(H.MODEL_PLT, H.ELLIPSOID_FROM_SPHERE, par, H.PLT_MATERIAL, H.PLT_INNER_EDGES)
    ctx = H.Context(n, n, n, (1, 0, 0), par["tau"], device=args.local_rank)
    ctx.set_flags(fl.reshape(-1))
    for o in range(6):
        ctx.set_bc_velocity(o, bc[o])
    ctx.set_force_limit(par["f_limit"])
    t0_, t1_ = rbc.add_to(ctx), plt.add_to(ctx)
    rc, rid = rbc.place(rbc_rows, DX, (n, n, n), fl.reshape(-1))
    pc, pid = plt.place(plt_rows, DX, (n, n, n), fl.reshape(-1), cell_id0=len(rbc_rows))
    ctx.add_cells(t0_, rc, rid); ctx.add_cells(t1_, pc, pid)
    ctx.set_timescales(5, 1, 1); ctx.set_material_timescale(t0_, 20); ctx.set_material_timescale(t1_, 20)
    ctx.iterate(args.warmup)
    launches0 = ctx.launch_count()
    ctx.timers_enable(True); ctx.timers_reset()
    sampler = ClockSampler(args.local_rank)
    ctx.synchronize(); t0 = time.time()
    ms = ctx.iterate_timed(args.steps)
    ctx.synchronize(); t1 = time.time()
    clocks = sampler.stop(t0, t1)
    launches = ctx.launch_count() - launches0
    timers = ctx.timers(); ctx.timers_enable(False)
    ncell = ctx.count()[0]
    npart = ctx.capacity()[1]
    host = pinned_empty(3 * npart); out_pos = pinned_empty(3 * npart)
    ctx.L.hcg_cells_download(ctx.h, C.c_int32(H.P_POS), host.ctypes.data_as(H.c_dp))
    ctx.synchronize(); te0 = time.time()
    ctx.cells_upload(H.P_POS, host)
    for _ in range(args.steps):
        ctx.iterate(1); ctx.count()
    ctx.L.hcg_cells_download(ctx.h, C.c_int32(H.P_POS), out_pos.ctypes.data_as(H.c_dp))
    ctx.synchronize(); e2e_ms = (time.time() - te0) * 1e3
    nodes = n ** 3
    peak, peak_src = measured_peak()
    gen = timers.get("kernel:k_collide_stream", (0.0, 0)); t1k = timers.get("kernel:k_collide_tau1", (0.0, 0))
    k_ms = (gen[0] + t1k[0]) / max(gen[1] + t1k[1], 1)
    bytes_lu = (B_LU * gen[1] + B_LU_TAU1 * t1k[1]) / max(gen[1] + t1k[1], 1)
    ach = bytes_lu * nodes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else None
    line = {"metric": METRIC, "value": nodes * args.steps / (ms * 1e-3) / 1e6, "unit": "MLUPS", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "cell_steps_per_s": ncell * args.steps / (ms * 1e-3),
            "config": {"workload": f"examples/cube: {n}^3 D3Q19 fp64, x periodic, bounce-back y planes, moving z walls, tau=1, "
                                   f"{len(rid)} RBC + {len(pid)} PLT (seeded packing), material every 20, velocity every 5",
                       "lattice": [n, n, n], "cells": int(ncell), "lsp": int(npart), "velocity_cadence": 5, "material_cadence": 20,
                       "l2": "the 0.3 GB of populations exceed the 126 MB L2; no flush"},
            "roofline": {"bound": "hbm", "kernel": "k_collide_stream / k_collide_tau1 (mix of the launches timed)", "achieved": ach, "peak": peak,
                         "unit": "GB/s", "frac": ach / peak if ach else None, "traffic": None, "peak_source": peak_src,
                         "bytes_per_lu": bytes_lu, "launch_ms": k_ms, "launches_timed": gen[1] + t1k[1]},
            "kernel_ms_per_step": {k: v[0] / args.steps for k, v in timers.items()},
            "e2e": {"value": nodes * args.steps / (e2e_ms * 1e-3) / 1e6, "unit": "MLUPS", "h2d_bytes_per_step": 8 * 3 * npart / args.steps,
                    "d2h_bytes_per_step": 8 * 3 * npart / args.steps + 16, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches, "clocks": clocks}
    ctx.close()
    return line


# ----------------------------------------------------------------------------- CPU arm
def run_cpu(steps, warmup, cadence, budget_s=150.0):
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
        rows = synthetic_rows(n)
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
    ap.add_argument("--workload", default="performance_testing", choices=["performance_testing", "cube"],
                    help="performance_testing = the weak-scaling unit the metric is quoted on (default); cube = examples/cube, 1 GPU")
    args = ap.parse_args()
    args.rank = int(os.environ.get("RANK", "0"))
    args.world = int(os.environ.get("WORLD_SIZE", "1"))
    args.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        if args.rank != 0:
            return
        cb = run_cpu(args.steps, args.warmup, args.cadence)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "MLUPS", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "cases/performance_testing unit (bounded CPU sample, see cpu_baseline.sample)",
                           "velocity_cadence": args.cadence, "material_cadence": 20},
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return
    if args.workload == "cube":
        if args.rank == 0:
            print(json.dumps(run_cube(args)), flush=True)
        return
    line = run_cuda(args)
    if args.rank != 0:
        return
    if args.world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = run_cpu(6, 1, args.cadence, budget_s=30.0)
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
