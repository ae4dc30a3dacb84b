"""GPU parity tests: every operator of the hot path, called through the C ABI
(libhemocell_gpu.so), against the CPU oracle on the same seeded inputs.
Tolerance: 1e-12 relative (north_star), with an absolute floor of 1e-14 x field scale."""
import numpy as np
import pytest

import oracle as O
from oracle import mesh as M
import util as U

pytestmark = pytest.mark.gpu


def _lib():
    from hemocell_b200 import lib as H
    return H


# --------------------------------------------------------------------------- lattice
# several shapes (the opt-in row-pipelined kernel, HCG_K1_ROWS=1, is eligible for nz = 32 / 64)
@pytest.mark.parametrize("shape", [(24, 20, 16), (12, 14, 32), (7, 9, 64)])
@pytest.mark.parametrize("periodic,flagkind,tau", [
    ((1, 1, 1), "bb_slab", 1.16), ((1, 1, 1), "none", 1.0), ((1, 1, 0), "couette", 1.16),
    ((0, 0, 0), "box", 0.86), ((1, 0, 0), "channel_bb", 1.82)])
def test_collide_stream_parity(periodic, flagkind, tau, shape):
    H = _lib()
    nx, ny, nz = shape
    bc = np.zeros((6, 3))
    if flagkind == "bb_slab":
        fl = np.zeros((nx, ny, nz), dtype=np.uint8); fl[nx//4:nx//4+3, 3:ny//2+1, 2:7] = 1; fl[:, 0, :] = 1
    elif flagkind == "none":
        fl = np.zeros((nx, ny, nz), dtype=np.uint8)
    elif flagkind == "couette":
        fl = U.couette_flags(nx, ny, nz); bc[4] = (0.02, 0.0, 0.0); bc[5] = (-0.02, 0.001, 0.0)
    elif flagkind == "box":
        fl = U.box_flags(nx, ny, nz)
    else:
        fl = np.zeros((nx, ny, nz), dtype=np.uint8)
        fl[:, 0, :] = 1; fl[:, ny - 1, :] = 1; fl[:, :, 0] = 1; fl[:, :, nz - 1] = 1
    fl = fl.reshape(-1)
    dom = O.make_domain(nx, ny, nz, periodic, tau, bc)
    rng = np.random.default_rng(7)
    pop = U.mask_inflow(dom, U.smooth_state(dom, 11))
    force = np.ascontiguousarray(1e-5 * rng.standard_normal(3 * nx * ny * nz))
    ctx = U.gpu_context(dom, fl, bc)
    ctx.lattice_upload(H.LAT_POP, pop)
    ctx.lattice_upload(H.LAT_FORCE, force)
    U.assert_close(ctx.lattice_download(H.LAT_POP), pop, "upload/download round trip", rtol=0, floor=0)
    ref = pop.copy()
    for step in range(1, 7):
        O.collide_and_stream(dom, fl, ref, force)
        ctx.op("collide_stream")
        if step in (1, 6):
            U.assert_close(ctx.lattice_download(H.LAT_POP), ref, f"populations after {step} steps ({flagkind})")
    rho, vel = O.moments(dom, fl, ref, force)
    U.assert_close(ctx.lattice_download(H.LAT_VELOCITY), vel, "velocity field")
    U.assert_close(ctx.lattice_download(H.LAT_DENSITY), rho, "density field")
    ctx.close()


def test_equilibrium_init_and_body_force():
    H = _lib()
    nx, ny, nz = 16, 12, 10
    dom = O.make_domain(nx, ny, nz, (1, 1, 1), 1.0)
    fl = np.zeros(nx * ny * nz, dtype=np.uint8)
    body = (1e-6, -2e-6, 3e-6)
    ctx = U.gpu_context(dom, fl, body=body)
    ctx.init_equilibrium(1.0, (0.01, 0.0, -0.02))
    ref = O.init_equilibrium(dom, 1.0, (0.01, 0.0, -0.02))
    U.assert_close(ctx.lattice_download(H.LAT_POP), ref, "initializeAtEquilibrium")
    force = np.empty(3 * nx * ny * nz)
    for k in range(3):
        force[k * nx * ny * nz:(k + 1) * nx * ny * nz] = body[k]
    U.assert_close(ctx.lattice_download(H.LAT_FORCE), force, "body force", rtol=0, floor=0)
    for _ in range(5):
        O.collide_and_stream(dom, fl, ref, force)
    ctx.fluid_warmup(5)
    U.assert_close(ctx.lattice_download(H.LAT_POP), ref, "5 warm-up steps with body force")
    ctx.close()


# --------------------------------------------------------------------------- IBM
def _one_rbc_setup(periodic=(1, 1, 0), n_cells=2, seed=3):
    nx, ny, nz = 40, 32, 24
    par = M.Parameters(dx=0.5e-6, dt=0.5e-7)
    bc = np.zeros((6, 3)); bc[4] = (0.01, 0, 0); bc[5] = (-0.01, 0, 0)
    fl = U.couette_flags(nx, ny, nz) if not periodic[2] else np.zeros((nx, ny, nz), dtype=np.uint8)
    fl = fl.reshape(-1)
    dom = O.make_domain(nx, ny, nz, periodic, par.tau, bc)
    ct = O.rbc_celltype(par)
    centers = [(12.3, 14.1, 11.7), (38.6, 15.2, 12.4)][:n_cells]     # the second one straddles the periodic x face
    cells = U.deformed_cells(ct, centers, seed)
    return par, dom, fl, bc, ct, cells


@pytest.mark.parametrize("spread_mode", [1, 0])     # 1 = node-sorted pairs + warp reduction, 0 = plain atomics
def test_spread_interpolate_advance_parity(spread_mode):
    H = _lib()
    par, dom, fl, bc, ct, cells = _one_rbc_setup()
    N = dom.nx * dom.ny * dom.nz
    rng = np.random.default_rng(5)
    pos = np.ascontiguousarray(cells.reshape(-1, 3))
    pforce = np.ascontiguousarray(rng.standard_normal(pos.shape) * par.f_limit * 0.5)   # some exceed the cap
    frep = np.ascontiguousarray(rng.standard_normal(pos.shape) * par.f_limit * 0.01)
    pop = U.mask_inflow(dom, U.smooth_state(dom, 21))
    ctx = U.gpu_context(dom, fl, bc)
    ctx.set_force_limit(par.f_limit)
    ctx.set_spread_mode(spread_mode, 20)
    t = U.gpu_add_type(ctx, ct)
    ctx.add_cells(t, cells, np.arange(cells.shape[0]))
    ctx.cells_upload(H.P_FORCE, pforce); ctx.cells_upload(H.P_FREP, frep)
    ctx.lattice_upload(H.LAT_POP, pop)
    U.assert_close(ctx.cells_download(H.P_POS), pos, "position round trip", rtol=0, floor=0)
    # spread
    node_force = np.zeros(3 * N)
    pf_ref = pforce.copy()
    O.spread(dom, fl, pos, pf_ref, frep, par.f_limit, node_force)
    ctx.op("spread")
    U.assert_close(ctx.cells_download(H.P_FORCE), pf_ref, "capped particle force")
    U.assert_close(ctx.lattice_download(H.LAT_FORCE), node_force, "spread node force")
    assert abs(node_force.sum()) > 0
    # collide-stream then interpolate (kernel at the pre-advance position, force still on the nodes)
    ref = pop.copy()
    O.collide_and_stream(dom, fl, ref, node_force)
    ctx.op("collide_stream")
    vel = O.interpolate(dom, fl, pos, ref, node_force)
    ctx.op("interpolate")
    U.assert_close(ctx.cells_download(H.P_VEL), vel, "interpolated velocity")
    # advance
    pos2 = pos.copy()
    O.advance(dom, fl, pos2, vel)
    ctx.op("advance")
    U.assert_close(ctx.cells_download(H.P_POS), pos2, "advanced position", rtol=1e-15, floor=0)
    ctx.close()


def test_advance_deletes_cell_on_boundary_node():
    H = _lib()
    par, dom, fl, bc, ct, cells = _one_rbc_setup()
    ctx = U.gpu_context(dom, fl, bc)
    t = U.gpu_add_type(ctx, ct)
    ctx.add_cells(t, cells, np.arange(cells.shape[0]))
    vel = np.zeros((cells.shape[0] * ct.V, 3))
    vel[5, 2] = -(cells.reshape(-1, 3)[5, 2])        # vertex 5 of cell 0 lands on the z = 0 wall plane
    ctx.cells_upload(H.P_VEL, vel)
    ctx.op("advance")
    assert ctx.count() == (1, ct.V)
    _, _, alive = ctx.cells_info()
    assert list(alive) == [0, 1]
    ctx.close()


# --------------------------------------------------------------------------- mechanics
@pytest.mark.parametrize("kind", ["rbc", "plt", "rbc_visc"])
def test_mechanics_parity(kind):
    H = _lib()
    par = M.Parameters(dx=0.5e-6, dt=1e-7)
    if kind == "plt":
        ct = O.plt_celltype(par, dict(M.PLT_MATERIAL, eta_m=2e-9))
    elif kind == "rbc_visc":
        ct = O.rbc_celltype(par, dict(M.RBC_MATERIAL, eta_m=5e-10))
    else:
        ct = O.rbc_celltype(par)
    centers = [(30.2, 40.7, 50.1), (201.5, 120.25, 77.0), (12.0, 12.0, 12.0)]
    cells = U.deformed_cells(ct, centers, 9, amp=0.02 if kind != "plt" else 0.01)
    pos = np.ascontiguousarray(cells.reshape(-1, 3))
    rng = np.random.default_rng(1)
    vel = np.ascontiguousarray(1e-3 * rng.standard_normal(pos.shape))
    f_ref = np.zeros_like(pos)
    comp = O.mechanics(ct, pos, vel, f_ref, components=True)
    dom = O.make_domain(256, 128, 96, (1, 1, 1), par.tau)
    ctx = U.gpu_context(dom, np.zeros(256 * 128 * 96, dtype=np.uint8))
    t = U.gpu_add_type(ctx, ct)
    ctx.add_cells(t, cells, np.arange(len(centers)))
    ctx.cells_upload(H.P_VEL, vel)
    ctx.cells_upload(H.P_FORCE, rng.standard_normal(pos.shape))      # must be overwritten, not accumulated
    ctx.op("mechanics", 1, 1)
    scale = np.abs(f_ref).max()
    got = ctx.cells_download(H.P_FORCE)
    U.assert_close(got, f_ref, f"total vertex force ({kind})")
    names = ["area", "volume", "bending", "link", "visc", "inner"]
    for k, fld in enumerate([H.P_F_AREA, H.P_F_VOLUME, H.P_F_BEND, H.P_F_LINK, H.P_F_VISC, H.P_F_INNER]):
        g = ctx.cells_download(fld)
        # components are summed in a different association than the total; floor on the total's scale
        err = np.abs(g - comp[k])
        tol = U.RTOL * np.maximum(np.abs(g), np.abs(comp[k])) + U.FLOOR * scale
        assert np.all(err <= tol), f"{names[k]} force ({kind}): max err {err.max():.3e}"
    assert np.abs(comp[0]).max() > 0 and np.abs(comp[1]).max() > 0 and np.abs(comp[3]).max() > 0
    # non-forced call respects the material cadence (iter % timescale)
    ctx.set_material_timescale(t, 20)
    ctx.set_iteration(7)
    ctx.cells_upload(H.P_FORCE, np.ones_like(pos))
    ctx.op("mechanics", 0, 0)
    U.assert_close(ctx.cells_download(H.P_FORCE), np.ones_like(pos), "mechanics skipped off-cadence", rtol=0, floor=0)
    ctx.close()


# --------------------------------------------------------------------------- repulsion
def test_repulsion_parity():
    H = _lib()
    par = M.Parameters(dx=0.5e-6, dt=1e-7)
    nx, ny, nz = 48, 32, 32
    fl = np.zeros((nx, ny, nz), dtype=np.uint8)
    fl[:, 0, :] = 1; fl[:, ny - 1, :] = 1
    fl = fl.reshape(-1)
    dom = O.make_domain(nx, ny, nz, (1, 0, 1), par.tau)
    ct = O.rbc_celltype(par)
    # three cells in near contact; one wraps across periodic x, one hugs the y = 0 wall
    centers = [(10.0, 14.0, 16.0), (10.5, 16.3, 16.2), (46.0, 2.0, 15.0)]
    cells = U.deformed_cells(ct, centers, 4, amp=0.0, stretch=(1, 1, 1))
    pos = np.ascontiguousarray(cells.reshape(-1, 3))
    cell_of = np.repeat(np.arange(3), ct.V)
    k_rep = 2e-22 / par.df; cutoff = 0.7e-6 / par.dx
    ref = O.repulsion(dom, pos, cell_of, k_rep, cutoff)
    assert np.abs(ref).max() > 0
    ctx = U.gpu_context(dom, fl)
    t = U.gpu_add_type(ctx, ct)
    ctx.add_cells(t, cells, np.arange(3))
    ctx.set_repulsion(True, k_rep, cutoff)
    ctx.cells_upload(H.P_FREP, np.ones_like(pos))          # zeroed by the operator
    ctx.op("repulsion")
    U.assert_close(ctx.cells_download(H.P_FREP), ref, "cell-cell repulsion")
    # wall repulsion accumulates on top
    ref2 = ref.copy()
    O.wall_repulsion(dom, fl, pos, 3 * k_rep, 1.3 * cutoff, ref2)
    assert np.abs(ref2 - ref).max() > 0
    ctx.set_wall_repulsion(True, 3 * k_rep, 1.3 * cutoff)
    ctx.op("wall_repulsion")
    U.assert_close(ctx.cells_download(H.P_FREP), ref2, "wall repulsion")
    ctx.close()


# --------------------------------------------------------------------------- iterate
# nz = 32 is also eligible for the opt-in overlapped collide+moments path (HCG_K1_ROWS=1 HCG_OVERLAP=1)
@pytest.mark.parametrize("nz", [24, 32])
@pytest.mark.parametrize("cadence", [(1, 1), (5, 10)])
def test_iterate_parity(cadence, nz):
    """full HemoCell::iterate() for 20 steps: shear flow, one RBC + one PLT, body force"""
    H = _lib()
    vel_ts, mat_ts = cadence
    nx, ny = 40, 32
    par = M.Parameters(dx=0.5e-6, dt=0.5e-7)
    bc = np.zeros((6, 3)); bc[4] = (0.02, 0, 0); bc[5] = (-0.02, 0, 0)
    fl = U.couette_flags(nx, ny, nz).reshape(-1)
    dom = O.make_domain(nx, ny, nz, (1, 1, 0), par.tau, bc)
    body = (2e-7, 0.0, 0.0)
    rbc, plt = O.rbc_celltype(par), O.plt_celltype(par)
    rbc_cells = U.deformed_cells(rbc, [(14.0, 16.0, 12.0)], 2, amp=0.0, stretch=(1.05, 0.98, 0.97))
    plt_cells = U.deformed_cells(plt, [(30.0, 10.0, 9.0), (38.9, 20.0, 14.0)], 3, amp=0.0, stretch=(1.03, 1, 0.98))
    sim = O.OracleSim(dom, fl, par.f_limit, body)
    sim.vel_timescale = vel_ts
    sim.add_celltype(rbc, mat_ts); sim.add_celltype(plt, mat_ts)
    sim.add_cells(0, rbc_cells, [0]); sim.add_cells(1, plt_cells, [1, 2])
    ctx = U.gpu_context(dom, fl, bc, body)
    ctx.set_force_limit(par.f_limit)
    t0, t1 = U.gpu_add_type(ctx, rbc), U.gpu_add_type(ctx, plt)
    ctx.add_cells(t0, rbc_cells, [0]); ctx.add_cells(t1, plt_cells, [1, 2])
    ctx.set_timescales(vel_ts, 1, 1)
    ctx.set_material_timescale(t0, mat_ts); ctx.set_material_timescale(t1, mat_ts)
    for chunk in (1, 19):
        for _ in range(chunk):
            sim.iterate()
        ctx.iterate(chunk)
        assert ctx.iteration == sim.iter
        U.assert_close(ctx.cells_download(H.P_POS), sim.pos, f"positions @ {sim.iter}", rtol=1e-13)
        U.assert_close(ctx.cells_download(H.P_VEL), sim.vel, f"velocities @ {sim.iter}", rtol=1e-9, floor=1e-12)
        U.assert_close(ctx.cells_download(H.P_FORCE), sim.pforce, f"forces @ {sim.iter}", rtol=1e-8, floor=1e-10)
        U.assert_close(ctx.lattice_download(H.LAT_POP), sim.pop, f"populations @ {sim.iter}", rtol=1e-10, floor=1e-12)
        U.assert_close(ctx.lattice_download(H.LAT_FORCE), sim.force, f"node force reset @ {sim.iter}", rtol=0, floor=0)
    ctx.close()


# --------------------------------------------------------------------------- experimental lattice paths
def test_experimental_row_pipeline_matches_default():
    """the opt-in bulk-async row kernel + overlapped moments pass give the default path's results"""
    import subprocess, sys, os
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fused_repro.py")
    base = dict(os.environ)
    r0 = subprocess.run([sys.executable, script, "24", "24", "64", "6"], env=dict(base, HCG_FUSED="0", HCG_K1_ROWS="0", HCG_OVERLAP="0"),
                        capture_output=True, text=True, timeout=300)
    assert r0.returncode == 0, r0.stdout + r0.stderr
    r1 = subprocess.run([sys.executable, script, "24", "24", "64", "6"], env=dict(base, HCG_FUSED="1", HCG_K1_ROWS="1", HCG_OVERLAP="1"),
                        capture_output=True, text=True, timeout=300)
    assert r1.returncode == 0 and "fused == separate" in r1.stdout, r1.stdout + r1.stderr


def test_output_observables_stretch_and_pineq():
    """output-path kernels: per-cell largest vertex distance (helper/cellInfo.cpp:103-121) and the
    off-equilibrium momentum flux behind the ShearStress / StrainRate fluid fields, vs numpy"""
    H = _lib()
    par, dom, fl, bc, ct, cells = _one_rbc_setup()
    N = dom.nx * dom.ny * dom.nz
    two = np.concatenate([cells, U.deformed_cells(ct, [cells[0].mean(0) + np.array([0.0, 0.0, 0.0])], 9, stretch=(1.2, 0.9, 0.95))])
    pop = U.mask_inflow(dom, U.smooth_state(dom, 33))
    ctx = U.gpu_context(dom, fl, bc)
    t = U.gpu_add_type(ctx, ct)
    ctx.add_cells(t, two, np.arange(two.shape[0]))
    ctx.lattice_upload(H.LAT_POP, pop)
    ref = []
    for c in two:
        d = c[:, None, :] - c[None, :, :]
        ref.append(np.sqrt((d * d).sum(-1).max()))
    U.assert_close(ctx.stretch(), np.array(ref), "cell stretch", rtol=1e-14)
    C = np.array([[0,0,0],[-1,0,0],[0,-1,0],[0,0,-1],[-1,-1,0],[-1,1,0],[-1,0,-1],[-1,0,1],[0,-1,-1],[0,-1,1],
                  [1,0,0],[0,1,0],[0,0,1],[1,1,0],[1,-1,0],[1,0,1],[1,0,-1],[0,1,1],[0,1,-1]], dtype=np.float64)
    f = pop.reshape(19, N)
    rhoBar = f.sum(0); j = C.T @ f; inv = 1.0 / (1.0 + rhoBar)
    pairs = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]
    pi = np.stack([(C[:, a] * C[:, b]) @ f - (rhoBar / 3.0 if a == b else 0.0) - inv * j[a] * j[b] for a, b in pairs])
    pi[:, fl.reshape(-1) == 1] = 0.0
    U.assert_close(ctx.lattice_download(H.LAT_PINEQ), pi, "PiNeq", rtol=1e-10, floor=1e-12)
    ctx.close()


@pytest.mark.parametrize("flagkind", ["none", "bb_slab", "couette", "box"])
def test_tau1_fast_path_matches_oracle(flagkind):
    """omega = 1: after a moments pass the collision reads the node's raw moments instead of its 19 populations
    (k_collide_tau1); walls take the generic path inside the same kernel"""
    H = _lib()
    nx, ny, nz = 20, 18, 32
    bc = np.zeros((6, 3))
    periodic = (1, 1, 1)
    fl = np.zeros((nx, ny, nz), dtype=np.uint8)
    if flagkind == "bb_slab":
        fl[nx//4:nx//4+3, 3:ny//2+1, 2:7] = 1; fl[:, 0, :] = 1
    elif flagkind == "couette":
        fl = U.couette_flags(nx, ny, nz); bc[4] = (0.02, 0.0, 0.0); bc[5] = (-0.02, 0.001, 0.0); periodic = (1, 1, 0)
    elif flagkind == "box":
        fl = U.box_flags(nx, ny, nz); periodic = (0, 0, 0)
    fl = fl.reshape(-1)
    dom = O.make_domain(nx, ny, nz, periodic, 1.0, bc)
    rng = np.random.default_rng(17)
    pop = U.mask_inflow(dom, U.smooth_state(dom, 19))
    force = np.ascontiguousarray(1e-5 * rng.standard_normal(3 * nx * ny * nz))
    ctx = U.gpu_context(dom, fl, bc)
    ctx.lattice_upload(H.LAT_POP, pop)
    ctx.lattice_upload(H.LAT_FORCE, force)
    ref = pop.copy()
    launches = []
    for step in range(1, 6):
        rho, vel = O.moments(dom, fl, ref, force)
        U.assert_close(ctx.lattice_download(H.LAT_DENSITY), rho, f"density before step {step}")   # moments pass: W becomes valid
        O.collide_and_stream(dom, fl, ref, force)
        ctx.op("collide_stream")
        U.assert_close(ctx.lattice_download(H.LAT_POP), ref, f"populations after {step} steps ({flagkind}, tau = 1)")
    ctx.close()


def test_body_force_field_parity():
    """per-node driving force (cases/kolmogorovFlow: the two half-domains pushed in opposite directions): the node
    force is reset to the field after every step, on interpolation steps and on the others"""
    H = _lib()
    nx, ny, nz = 24, 20, 32
    par = M.Parameters(dx=0.5e-6, dt=-1.0)
    dom = O.make_domain(nx, ny, nz, (1, 1, 1), par.tau)
    fl = np.zeros(nx * ny * nz, dtype=np.uint8)
    N = nx * ny * nz
    field = np.zeros((3, nx, ny, nz))
    field[0, :, :, nz // 2:] = 3e-6; field[0, :, :, :nz // 2] = -3e-6; field[1, :, :5, :] = 1e-6
    field = np.ascontiguousarray(field.reshape(-1))

    class Sim(O.OracleSim):
        def _reset_force(self):
            self.force[:] = field

    rbc = O.rbc_celltype(par)
    cells = U.deformed_cells(rbc, [(11.0, 10.0, 15.0)], 2, amp=0.0, stretch=(1.04, 0.98, 0.98))
    sim = Sim(dom, fl, par.f_limit)
    sim.vel_timescale = 2
    sim.add_celltype(rbc, 4); sim.add_cells(0, cells, [0])
    ctx = U.gpu_context(dom, fl)
    ctx.set_force_limit(par.f_limit)
    ctx.set_body_force_field(field)
    t = U.gpu_add_type(ctx, rbc)
    ctx.add_cells(t, cells, [0])
    ctx.set_timescales(2, 1, 1); ctx.set_material_timescale(t, 4)
    U.assert_close(ctx.lattice_download(H.LAT_FORCE), field, "force field after upload", rtol=0, floor=0)
    for _ in range(9):
        sim.iterate()
    ctx.iterate(9)
    U.assert_close(ctx.lattice_download(H.LAT_FORCE), field, "node force reset to the field", rtol=0, floor=0)
    U.assert_close(ctx.lattice_download(H.LAT_POP), sim.pop, "populations", rtol=1e-10, floor=1e-12)
    U.assert_close(ctx.cells_download(H.P_POS), sim.pos, "positions", rtol=1e-13)
    ctx.set_body_force((1e-6, 0.0, 0.0))                                   # back to a uniform force
    assert np.all(ctx.lattice_download(H.LAT_FORCE).reshape(3, N)[0] == 1e-6)
    ctx.close()
