"""GPU parity of SURVEY.md section 8 row f4: Zou-He velocity / pressure nodes with per-node values and the pre-inlet
coupling (helper/preInlet.cpp, examples/pipeflow_with_preinlet), against the CPU oracle, through the C ABI."""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

import oracle as O
import util as U
import preinlet_case as PC

pytestmark = pytest.mark.gpu


def _lib():
    from hemocell_b200 import lib as H
    return H


def _io_flags(nx, ny, nz, axis):
    """bounce-back duct around `axis`, Zou-He velocity nodes on its low face, pressure nodes on its high face,
    plus one velocity node and one pressure node of every other orientation inside the fluid"""
    fl = np.zeros((nx, ny, nz), dtype=np.uint8)
    n = (nx, ny, nz)
    for a in range(3):
        if a == axis:
            continue
        sl = [slice(None)] * 3
        sl[a] = 0; fl[tuple(sl)] = 1
        sl[a] = n[a] - 1; fl[tuple(sl)] = 1
    lo = [slice(None)] * 3; lo[axis] = 0
    hi = [slice(None)] * 3; hi[axis] = n[axis] - 1
    inner = fl[tuple(lo)] == 0
    fl[tuple(lo)][inner] = 8 + 2 * axis            # outward normal -axis
    fl[tuple(hi)][inner] = 14 + 2 * axis + 1       # outward normal +axis
    for o in range(6):
        fl[3 + o, 4, 5] = 8 + o
        fl[4 + o, 6, 3] = 14 + o
    return fl.reshape(-1)


@pytest.mark.parametrize("axis,tau,shape", [(0, 0.8, (20, 12, 10)), (1, 1.0, (14, 18, 12)), (2, 1.3, (12, 10, 22))])
def test_zouhe_nodes_collide_stream_parity(axis, tau, shape):
    H = _lib()
    nx, ny, nz = shape
    N = nx * ny * nz
    fl = _io_flags(nx, ny, nz, axis)
    dom = O.make_domain(nx, ny, nz, (0, 0, 0), tau)
    rng = np.random.default_rng(21 + axis)
    pop = U.mask_inflow(dom, U.smooth_state(dom, 12))
    force = np.ascontiguousarray(1e-5 * rng.standard_normal(3 * N))
    nodes = np.nonzero(fl >= 8)[0]
    val = np.column_stack([0.02 * rng.standard_normal((nodes.size, 3)), 1.0 + 2e-3 * rng.standard_normal(nodes.size)])
    bc = np.zeros((4, N)); bc[3] = 1.0
    bc[:, nodes] = val.T
    bc = np.ascontiguousarray(bc.reshape(-1))
    ctx = U.gpu_context(dom, fl)
    ctx.set_bc_nodes(nodes, val)
    ctx.lattice_upload(H.LAT_POP, pop)
    ctx.lattice_upload(H.LAT_FORCE, force)
    ref = pop.copy()
    for step in range(1, 6):
        if tau == 1.0 and step > 1:
            ctx.lattice_download(H.LAT_DENSITY)          # a moments pass: the next collision takes the tau = 1 fast path
        O.collide_and_stream(dom, fl, ref, force, bc_node=bc)
        ctx.op("collide_stream")
        if step in (1, 5):
            U.assert_close(ctx.lattice_download(H.LAT_POP), ref, f"populations after {step} steps (axis {axis})")
    rho, vel = O.moments(dom, fl, ref, force, bc_node=bc)
    U.assert_close(ctx.lattice_download(H.LAT_VELOCITY), vel, "velocity field")
    U.assert_close(ctx.lattice_download(H.LAT_DENSITY), rho, "density field")
    pick = np.concatenate([nodes[:40], rng.integers(0, N, 60)])
    U.assert_close(ctx.node_velocity(pick), vel.reshape(3, N)[:, pick].T, "hcg_lattice_node_velocity")
    ctx.close()


def _gpu_pair(c, pre_device=0):
    H = _lib()
    pre = U.gpu_context(c['domp'], c['flp'], body=PC.BODY, device=pre_device)
    main = U.gpu_context(c['domm'], c['flm'])
    for ctx in (pre, main):
        ctx.init_equilibrium(1.0, PC.U0)
        ctx.set_force_limit(c['par'].f_limit)
    t = U.gpu_add_type(pre, c['rbc'])
    pre.add_cells(t, c['cells'], [0, 1])
    t = U.gpu_add_type(main, c['rbc'])
    main.reserve_cells(t, 4)
    main.add_cells(t, np.zeros((0, c['rbc'].V, 3)), np.zeros(0, dtype=np.int64))
    main.preinlet_map(pre, c['pre_idx'], c['main_idx'])
    return H, pre, main


def _by_id(ctx, H, field):
    ids, _, alive = ctx.cells_info()
    a = ctx.cells_download(field).reshape(len(ids), -1, 3)
    keep = np.nonzero((alive != 0) & (ids >= 0))[0]
    order = keep[np.argsort(ids[keep])]
    return ids[order], a[order]


def _oracle_by_id(sim, arr):
    a = arr.reshape(len(sim.cell_id), -1, 3)
    order = np.argsort(sim.cell_id)
    return sim.cell_id[order], a[order]


def _n_gpus():
    import ctypes as C
    try:
        rt = C.CDLL("libcudart.so.12"); n = C.c_int(0)
        return n.value if rt.cudaGetDeviceCount(C.byref(n)) == 0 else 0
    except OSError:
        return 0


@pytest.mark.parametrize("pre_device", [0, 1])
def test_preinlet_two_domains_match_the_oracle(pre_device):
    """periodic force-driven pre-inlet duct -> Zou-He inlet of the main duct with a pressure outlet, tau = 1, two RBCs in
    the pre-inlet of which one is handed over at the first step: 30 passes of the main loop of
    pipeflow_with_preinlet.cpp (iterate both, applyPreInlet) on the GPU vs the oracle.  pre_device = 1: the pre-inlet
    context lives on a second GPU of the box (velocity buffer and cells cross by peer copies)"""
    if pre_device >= max(_n_gpus(), 1):
        pytest.skip("needs two GPUs")
    c = PC.build()
    opre, omain, cpl = PC.oracle_pair(c)
    H, pre, main = _gpu_pair(c, pre_device)
    steps, handed_o, handed_g = 30, 0, 0
    for _ in range(steps):
        handed_o += PC.oracle_step(opre, omain, cpl)
        pre.iterate(1); main.iterate(1)
        main.preinlet_apply_velocity()
        handed_g += main.preinlet_apply_cells(0, float(PC.NXP), c['shift'], PC.SLAB[0], PC.SLAB[1], PC.ID_STRIDE)
    assert handed_o == handed_g == 1
    assert main.count()[0] == 1 and pre.count()[0] == 2
    U.assert_close(pre.lattice_download(H.LAT_POP), opre.pop, "pre-inlet populations", rtol=1e-9, floor=1e-11)
    U.assert_close(main.lattice_download(H.LAT_POP), omain.pop, "main populations", rtol=1e-9, floor=1e-11)
    for ctx, sim, name in ((pre, opre, "pre-inlet"), (main, omain, "main")):
        gi, gp = _by_id(ctx, H, H.P_POS)
        oi, op = _oracle_by_id(sim, sim.pos)
        np.testing.assert_array_equal(gi, oi)
        U.assert_close(gp, op, name + " positions", rtol=1e-11)
        _, gv = _by_id(ctx, H, H.P_VEL)
        _, ov = _oracle_by_id(sim, sim.vel)
        U.assert_close(gv, ov, name + " velocities", rtol=1e-8, floor=1e-10)
        _, gf = _by_id(ctx, H, H.P_FORCE)
        _, of = _oracle_by_id(sim, sim.pforce)
        U.assert_close(gf, of, name + " membrane forces", rtol=1e-6, floor=1e-8)
    # the inlet nodes carry the pre-inlet's coupling-plane velocity
    U.assert_close(main.node_velocity(c['main_idx']), opre.node_velocity(c['pre_idx']), "inlet node velocity", rtol=1e-9, floor=1e-11)
    main.close(); pre.close()


def test_preinlet_error_paths():
    H = _lib()
    c = PC.build()
    _, pre, main = _gpu_pair(c)
    with pytest.raises(H.HcgError):
        pre.preinlet_apply_velocity()                       # no map on this context
    with pytest.raises(H.HcgError):
        main.preinlet_map(pre, [pre.Nl], [0])               # node outside the lattice
    with pytest.raises(H.HcgError):
        main.set_bc_nodes([main.Nl], [[0, 0, 0, 1.0]])
    with pytest.raises(H.HcgError):
        main.set_flags(np.full(main.Nl, 20, dtype=np.uint8))   # unknown flag
    main.close(); pre.close()


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_pipeflow_with_preinlet_unmodified_binary(tmp_path):
    """the REFERENCE's examples/pipeflow_with_preinlet/pipeflow_with_preinlet.cpp compiled unmodified against include/hemocell.h
    (hemo::PreInlet, Zou-He inlet nodes fed by the periodic pre-inlet, 3-plane Zou-He pressure outlet), with its own config,
    STL and .pos files, shortened to 600 iterations: both domains on one GPU; the run finishes, the main domain holds cells at
    every measurement, the pre-inlet drives a flow through the inlet (mean velocity grows from rest, stays finite and
    sub-sonic), forces stay finite, and the fluid / particle HDF5 files are written"""
    name = "pipeflow_with_preinlet"
    src = os.path.join(ROOT, "build", "refcases", name)
    if not os.path.exists(os.path.join(src, name)):
        pytest.skip("build/refcases not present (built from /root/reference in the authoring container)")
    for f in [name, "config.xml", "RBC.xml", "PLT.xml", "RBC.pos", "PLT.pos", "normal.stl"]:
        shutil.copy(os.path.join(src, f), tmp_path / f)
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "hemocell_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    cfg = (tmp_path / "config.xml").read_text()
    for key, val in (("tmax", 600), ("tmeas", 100), ("tcheckpoint", 100000), ("tbalance", 100000)):
        cfg = re.sub(rf"<{key}>.*?</{key}>", f"<{key}> {val} </{key}>", cfg)
    (tmp_path / "config.xml").write_text(cfg)
    env["HEMOCELL_H5_DEFLATE"] = "1"
    r = subprocess.run([str(tmp_path / name), "config.xml"], cwd=tmp_path, capture_output=True, text=True, timeout=600, env=env)
    out = r.stdout + r.stderr
    for f in list((tmp_path / "tmp").glob("**/*")):
        if f.is_file() and "log" in f.name and f.suffix not in (".h5", ".csv", ".bin"):
            out += f.read_text(errors="ignore")
    dump = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(dump):
        open(os.path.join(dump, "preinlet_case.log"), "w").write(out)
    assert r.returncode == 0, out[-4000:]
    assert "(main) Simulation finished" in out
    assert "inlet nodes coupled to the pre-inlet" in out
    cells = [int(x) for x in re.findall(r"# of cells: (\d+)", out)]
    vmax = [float(x) for x in re.findall(r"Velocity  -  max\.: (\S+) m/s", out)]
    vmean = [float(x) for x in re.findall(r"m/s, mean: (\S+) m/s", out)]
    fmax = [float(x) for x in re.findall(r"pN, max\.: (\S+) pN", out)]
    pre_cells = [int(x) for x in re.findall(r"(\d+) RBC cells placed inside the pre-inlet", out)]
    assert len(cells) >= 6 and len(vmean) >= 6, out[-3000:]
    # the shipped RBC.pos was packed for a wider box: 8 of its 236 cells fit the main pipe, 6 the pre-inlet; two of those sit in
    # the pre-inlet's hand-over slab and are mirrored into the main domain at the first applyPreInlet()
    assert pre_cells and pre_cells[0] >= 1                      # the .pos file seeds the pre-inlet as well
    handed = [int(x) for x in re.findall(r"\((\d+) so far\)", out)]
    assert handed and handed[-1] >= 1
    # (the case's "# of cells" gathers over both domains, as the reference's MPI gather does)
    assert all(c >= 8 + pre_cells[0] for c in cells) and cells[-1] >= 8 + pre_cells[0] + 1, cells
    assert all(np.isfinite(v) and v > 0 for v in vmean) and vmean[-1] > vmean[0], vmean
    dx, dt = 5e-7, 1e-7
    assert all(v * dt / dx < 0.2 for v in vmax), vmax            # lattice velocity well below the speed of sound
    assert all(np.isfinite(f) for f in fmax), fmax
    import glob
    h5 = glob.glob(str(tmp_path / "**" / "hdf5" / "*" / "*.h5"), recursive=True)
    assert any(os.path.basename(str(f)).startswith("Fluid") for f in h5) and any(os.path.basename(str(f)).startswith("RBC") for f in h5)
    print("pipeflow_with_preinlet: cells", cells, "mean velocity m/s", vmean, "handed over:", re.findall(r"\((\d+) so far\)", out)[-1:])


def _preinlet_run(d, tmax, cfg_path=None):
    name = "pipeflow_with_preinlet"
    src = os.path.join(ROOT, "build", "refcases", name)
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "hemocell_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""),
               HEMOCELL_H5_DEFLATE="1")
    if cfg_path is None:
        d.mkdir()
        for f in [name, "config.xml", "RBC.xml", "PLT.xml", "RBC.pos", "PLT.pos", "normal.stl"]:
            shutil.copy(os.path.join(src, f), d / f)
        cfg = (d / "config.xml").read_text()
        for key, val in (("tmax", tmax), ("tmeas", 100), ("tcheckpoint", 200), ("tbalance", 100000)):
            cfg = re.sub(rf"<{key}>.*?</{key}>", f"<{key}> {val} </{key}>", cfg)
        (d / "config.xml").write_text(cfg)
        cfg_path = "config.xml"
    r = subprocess.run([str(d / name), str(cfg_path)], cwd=d, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    out = r.stdout
    stats = re.findall(r"Stats\. @ (\d+) .*?# of cells: (\d+).*?Velocity  -  max\.: (\S+) m/s, mean: (\S+) m/s.*?Force  -  min\.: (\S+) pN, max\.: (\S+) pN \(\S+ lf\), mean: (\S+) pN",
                       out, re.S)
    return out, {int(s[0]): [float(v) for v in s[1:]] for s in stats}


def test_preinlet_checkpoint_restart_reproduces_uninterrupted_run(tmp_path):
    """saveCheckPoint / loadCheckPoint with a pre-inlet (PRE_lattice / PRE_particleField of the reference,
    core/hemoCellFields.cpp:254-265, 297-314): the unmodified pipeflow_with_preinlet binary interrupted at iteration 200 and
    restarted from checkpoint.xml reaches the iteration-300 statistics of the uninterrupted run, and hands over no cell twice"""
    if not os.path.exists(os.path.join(ROOT, "build", "refcases", "pipeflow_with_preinlet", "pipeflow_with_preinlet")):
        pytest.skip("build/refcases not present (built from /root/reference in the authoring container)")
    _, straight = _preinlet_run(tmp_path / "straight", 300)
    _, first = _preinlet_run(tmp_path / "first", 200)
    assert sorted(straight) == [100, 200, 300] and sorted(first) == [100, 200]
    U.assert_close(first[200], straight[200], "two runs of the same case", rtol=1e-4, floor=1e-6)
    d = tmp_path / "first"
    cp = d / "tmp" / "checkpoint" / "checkpoint.xml"
    assert cp.exists() and (d / "tmp" / "checkpoint" / "pre.bin").exists()
    cp.write_text(re.sub(r"<tmax>.*?</tmax>", "<tmax> 300 </tmax>", cp.read_text()))
    out, resumed = _preinlet_run(d, 300, cfg_path=cp)
    assert "CHECKPOINT found" in out or "Loading Checkpoint" in out
    assert sorted(resumed) == [300]
    assert resumed[300][0] == straight[300][0]                              # number of cells
    U.assert_close(resumed[300][1:], straight[300][1:], "iteration-300 statistics of the restarted run", rtol=1e-4, floor=1e-6)
