"""CPU checks of the oracle's Zou-He nodes and pre-inlet coupling (the checker of tests/test_gpu_zz_preinlet.py)."""
import numpy as np

import oracle as O
import preinlet_case as PC


def test_zouhe_channel_conserves_mass_and_meets_its_boundary_values():
    """velocity inlet (parabolic profile) + pressure outlet (rho = 1) on a bounce-back channel: at steady state the
    mass flux is the same through every interior plane, the outlet density is the imposed one, the inlet node
    velocity is the imposed one, and the pressure falls monotonically along the channel"""
    nx, ny, nz = 24, 11, 4
    dom = O.make_domain(nx, ny, nz, (0, 0, 1), 0.8)
    N = nx * ny * nz
    fl = np.zeros((nx, ny, nz), np.uint8)
    fl[:, 0, :] = 1; fl[:, -1, :] = 1
    fl[0, 1:-1, :] = 8; fl[-1, 1:-1, :] = 15
    bc = np.zeros((4, nx, ny, nz)); bc[3] = 1.0
    y = np.arange(ny)
    prof = 0.02 * 4 * (y - 0.5) * (ny - 1.5 - y) / (ny - 2) ** 2
    bc[0, 0, :, :] = prof[:, None]
    bc = np.ascontiguousarray(bc.reshape(-1)); fl = np.ascontiguousarray(fl.reshape(-1))
    pop = O.init_equilibrium(dom); force = np.zeros(3 * N)
    for _ in range(2500):
        O.collide_and_stream(dom, fl, pop, force, bc_node=bc)
    rho, vel = O.moments(dom, fl, pop, force, bc_node=bc)
    u = vel.reshape(3, nx, ny, nz); rho = rho.reshape(nx, ny, nz)
    flux = np.array([(rho[x, 1:-1, 0] * u[0, x, 1:-1, 0]).sum() for x in range(1, nx)])
    assert np.all(np.abs(flux - flux[0]) < 1e-7 * abs(flux[0]))      # steady state: the same flux through every plane
    assert abs(flux[0] - prof[1:-1].sum()) < 0.02 * prof[1:-1].sum()      # inlet density is within 2 % of 1
    np.testing.assert_array_equal(u[0, 0, 1:-1, 0], prof[1:-1])
    np.testing.assert_array_equal(rho[-1, 1:-1, :], 1.0)
    assert np.all(np.abs(u[1:, -1]) == 0.0)                              # no tangential velocity on the outlet
    mid = rho[1:, ny // 2, 0]
    assert np.all(np.diff(mid) < 0) and mid[0] > 1.005


def test_zouhe_completion_gives_exact_moments():
    """after the completion + a collision without force, density and momentum of a Zou-He node are the imposed
    ones (BGK conserves them), for every orientation"""
    rng = np.random.default_rng(5)
    n = 5
    dom = O.make_domain(n, n, n, (0, 0, 0), 0.9)
    N = n ** 3
    C = np.array([[0,0,0],[-1,0,0],[0,-1,0],[0,0,-1],[-1,-1,0],[-1,1,0],[-1,0,-1],[-1,0,1],[0,-1,-1],[0,-1,1],
                  [1,0,0],[0,1,0],[0,0,1],[1,1,0],[1,-1,0],[1,0,1],[1,0,-1],[0,1,1],[0,1,-1]])
    centre = 2 + n * (2 + n * 2)
    for o in range(6):
        for pressure in (0, 1):
            fl = np.zeros(N, np.uint8); fl[centre] = (14 if pressure else 8) + o
            pop = O.init_equilibrium(dom, 1.0, (0.01, -0.02, 0.015)) + 1e-4 * rng.standard_normal(19 * N)
            bc = np.zeros((4, N)); bc[3] = 1.0
            bc[:, centre] = (0.03, -0.01, 0.02, 1.004)
            before = pop.reshape(19, N)[:, centre].copy()
            O.collide_and_stream(dom, fl, pop, np.zeros(3 * N), bc_node=np.ascontiguousarray(bc.reshape(-1)))
            # gather the post-collision populations of the centre node from where they streamed to
            p = pop.reshape(19, n, n, n)
            f = np.array([p[q, 2 + C[q, 0], 2 + C[q, 1], 2 + C[q, 2]] for q in range(19)])
            rho = 1.0 + f.sum(); j = f @ C
            d, sgn = o // 2, (1 if o & 1 else -1)
            if pressure:
                assert abs(rho - 1.004) < 1e-14
                for k in range(3):
                    if k != d:
                        assert abs(j[k]) < 1e-15
                known = [q for q in range(19) if C[q, d] * sgn >= 0]
                rho_on = sum(before[q] + (1/3 if q == 0 else 1/18 if (C[q] ** 2).sum() == 1 else 1/36) for q in known if C[q, d] == 0)
                rho_out = sum(before[q] + (1/18 if (C[q] ** 2).sum() == 1 else 1/36) for q in known if C[q, d] * sgn > 0)
                assert abs(j[d] / rho - sgn * ((rho_on + 2 * rho_out) / 1.004 - 1.0)) < 1e-14
            else:
                assert np.all(np.abs(j / rho - np.array([0.03, -0.01, 0.02])) < 1e-14)


def test_preinlet_handover_rule_and_coupling():
    c = PC.build()
    pre, main, cpl = PC.oracle_pair(c)
    handed = []
    for it in range(12):
        n = PC.oracle_step(pre, main, cpl)
        if n:
            handed.append((it, n))
    # cell 0's periodic image lies inside the slab from the start: handed over once, at the first step, as image k = +1 under id 0 + 2*stride
    assert handed == [(0, 1)]
    assert list(main.cell_id) == [0 + 2 * PC.ID_STRIDE] and list(pre.cell_id) == [0, 1]
    # the copy starts at the pre-inlet cell's position in main coordinates (one lap ahead)
    assert main.pos.shape == (c['rbc'].V, 3)
    d = main.pos.mean(0) - (pre.pos[:c['rbc'].V].mean(0) + np.array([PC.NXP - PC.XC, 0, 0]))
    assert np.all(np.abs(d) < 0.05)          # they drift apart slowly (different flow fields), not by a lattice shift
    # inlet nodes carry the pre-inlet's velocity of the coupling plane
    b = main.bc_node.reshape(4, main.N)
    np.testing.assert_array_equal(b[0:3, c['main_idx']].T, pre.node_velocity(c['pre_idx']))
    assert b[0, c['main_idx']].max() > 0.015
    assert np.isfinite(main.pop).all() and np.isfinite(pre.pop).all()
    # a second lap hands the same cell over again under a new id
    V = c['rbc'].V
    pre.pos[:V, 0] -= PC.NXP                 # as if the cell had gone round once more against the flow direction ...
    assert cpl.apply_cells() == 1 and sorted(main.cell_id) == [2 * PC.ID_STRIDE, 4 * PC.ID_STRIDE]


def test_zouhe_channel_reproduces_plane_poiseuille_flow():
    """known answer that pins the Zou-He nodes to physics (Palabos itself is not in the tree): plane Poiseuille flow between
    bounce-back walls, parabolic Zou-He velocity inlet, Zou-He pressure outlet.  Half-way bounce-back puts the walls half a
    node outside the last fluid node (channel width H = ny - 2); at steady state the pressure gradient must be
    dp/dx = -12 nu rho U_mean / H^2 with p = rho / 3, and the profile stays the imposed parabola along the channel."""
    nx, ny, nz = 40, 19, 3
    tau = 0.9
    nu = (tau - 0.5) / 3.0
    dom = O.make_domain(nx, ny, nz, (0, 0, 1), tau)
    N = nx * ny * nz
    fl = np.zeros((nx, ny, nz), np.uint8)
    fl[:, 0, :] = 1; fl[:, -1, :] = 1
    fl[0, 1:-1, :] = 8; fl[-1, 1:-1, :] = 15
    H = ny - 2
    y = np.arange(ny) - 0.5                       # distance from the lower wall
    umax = 0.03
    prof = 4 * umax * y * (H - y) / H ** 2
    bc = np.zeros((4, nx, ny, nz)); bc[3] = 1.0
    bc[0, 0, 1:-1, :] = prof[1:-1, None]
    bc = np.ascontiguousarray(bc.reshape(-1)); fl = np.ascontiguousarray(fl.reshape(-1))
    pop = O.init_equilibrium(dom); force = np.zeros(3 * N)
    for _ in range(6000):
        O.collide_and_stream(dom, fl, pop, force, bc_node=bc)
    rho, vel = O.moments(dom, fl, pop, force, bc_node=bc)
    u = vel.reshape(3, nx, ny, nz); rho = rho.reshape(nx, ny, nz)
    xs = np.arange(8, nx - 8)
    dpdx = np.polyfit(xs, rho[xs, ny // 2, 0] / 3.0, 1)[0]
    umean = prof[1:-1].sum() / H
    expect = -12 * nu * umean / H ** 2
    assert abs(dpdx / expect - 1.0) < 0.02, (dpdx, expect)
    mid = u[0, nx // 2, 1:-1, 0]
    assert np.max(np.abs(mid - prof[1:-1])) < 0.01 * umax
    assert np.max(np.abs(u[1, nx // 2])) < 1e-6 * umax


def test_product_handover_rule_matches_the_oracle_rule():
    """the host logic of hcg_preinlet_apply_cells (hch_preinlet_select, no device needed) takes the same cells under the same
    periodic image as the oracle's PreInletCoupling.apply_cells, on random cell extents, both flow directions, repeated calls"""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from hemocell_b200 import lib as H
    rng = np.random.default_rng(11)
    for trial in range(40):
        n = int(rng.integers(1, 60))
        period = float(rng.integers(20, 90))
        shift = float(rng.integers(-50, 120))
        slab_lo = float(rng.integers(0, 100)); slab_hi = slab_lo + float(rng.integers(5, 40))
        lo = rng.uniform(-300, 300, n); hi = lo + rng.uniform(0.5, 18, n)
        if trial % 5 == 0:                                   # exact hits on the slab edges
            lo[0] = slab_lo - shift - 2 * period; hi[0] = slab_hi - shift - 2 * period
        alive = rng.random(n) > 0.15
        last = np.full(n, np.iinfo(np.int64).min, dtype=np.int64)
        seen = {}
        for step in range(4):
            lap, take = H.preinlet_select(lo, hi, alive, last, shift, period, slab_lo, slab_hi)
            for i in range(n):                               # the oracle's rule, cell by cell (oracle/__init__.py: PreInletCoupling.apply_cells)
                a, b = lo[i] + shift, hi[i] + shift
                k = np.ceil((slab_lo - a) / period)
                inside = alive[i] and not (b + k * period > slab_hi)
                want = inside and seen.get(i) != int(k)
                assert bool(take[i]) == bool(want), (trial, step, i)
                if inside:
                    assert lap[i] == int(k)
                    assert slab_lo <= a + k * period and b + k * period <= slab_hi
                if want:
                    seen[i] = int(k); last[i] = int(k)
            drift = rng.uniform(-6, 6)
            lo = lo + drift; hi = hi + drift
    lap, take = H.preinlet_select([1.0], [3.0], [1], None, 0.0, 10.0, 0.0, 5.0)
    assert take[0] and lap[0] == 0


def test_preinlet_coupling_carries_the_flow_into_the_main_domain():
    """physics check of the coupling (fluid only): the force-driven periodic pre-inlet develops a duct flow; through the Zou-He inlet
    the main domain must carry the same volume flux at every cross-section (up to the small compressibility of the pressure drop),
    with the pre-inlet's profile at the inlet and density 1 at the outlet"""
    c = PC.build()
    pre = O.OracleSim(c['domp'], c['flp'], c['par'].f_limit, (4e-5, 0.0, 0.0))
    main = O.OracleSim(c['domm'], c['flm'], c['par'].f_limit)
    cpl = O.PreInletCoupling(pre, main, c['pre_idx'], c['main_idx'], 0, float(PC.NXP), c['shift'], PC.SLAB[0], PC.SLAB[1], 1)
    for _ in range(1500):
        pre.iterate(); main.iterate()
        cpl.apply_velocity()
    nxm, ny, nz = PC.NXM, PC.NY, PC.NZ
    _, vp = O.moments(pre.dom, pre.flags, pre.pop, pre.force)
    rho, vm = O.moments(main.dom, main.flags, main.pop, main.force, main.bc_node)
    up = vp.reshape(3, PC.NXP, ny, nz)[0]; um = vm.reshape(3, nxm, ny, nz)[0]; rho = rho.reshape(nxm, ny, nz)
    q_pre = up[PC.XC].sum()
    assert q_pre > 0.2                                             # a developed flow: ~200 fluid nodes x ~1.5e-3
    q_main = np.array([um[x, 1:-1, 1:-1].sum() for x in range(1, nxm - 1)])
    assert np.all(np.abs(q_main / q_pre - 1.0) < 0.03), q_main / q_pre
    np.testing.assert_array_equal(um[0][c['flm'].reshape(nxm, ny, nz)[0] == 8], up[PC.XC][c['flm'].reshape(nxm, ny, nz)[0] == 8])
    assert np.all(rho[-1][c['flm'].reshape(nxm, ny, nz)[-1] == 15] == 1.0)
    assert rho[1, ny // 2, nz // 2] > rho[nxm - 2, ny // 2, nz // 2] > 0.999      # pressure falls towards the outlet
