"""Shared two-domain pre-inlet case of the CPU (oracle-only) and GPU tests: a periodic, force-driven square duct
(the pre-inlet) feeds the Zou-He velocity inlet of a non-periodic duct with a Zou-He pressure outlet, as
examples/pipeflow_with_preinlet does with helper/preInlet.cpp (direction Xneg: flow in +x)."""
import numpy as np
import oracle as O
from oracle import mesh as M
import util as U

NY = NZ = 16
NXP, NXM = 24, 32
XC = NXP - 2                 # coupling plane of the pre-inlet <-> plane 0 of the main domain
SLAB = (0.0, 12.0)           # hand-over slab behind the inlet (particleEnvelope planes), main coordinates
ID_STRIDE = 2                # number_of_cells of the pre-inlet
BODY = (2e-5, 0.0, 0.0)
U0 = (0.02, 0.0, 0.0)


def duct(nx):
    fl = np.zeros((nx, NY, NZ), dtype=np.uint8)
    fl[:, 0, :] = 1; fl[:, NY - 1, :] = 1; fl[:, :, 0] = 1; fl[:, :, NZ - 1] = 1
    return fl


def build(tau_dt=-1.0):
    par = M.Parameters(dx=1.0e-6, dt=tau_dt)
    flp, flm = duct(NXP), duct(NXM)
    inner = flm[0] == 0
    flm[0][inner] = 8                 # Zou-He velocity nodes, outward normal -x
    flm[NXM - 1][inner] = 15          # Zou-He pressure nodes, outward normal +x
    domp = O.make_domain(NXP, NY, NZ, (1, 0, 0), par.tau)
    domm = O.make_domain(NXM, NY, NZ, (0, 0, 0), par.tau)
    rbc = O.rbc_celltype(par)
    # cell 0 sits where its periodic image is already inside the hand-over slab, cell 1 is half a lap behind
    cells = U.deformed_cells(rbc, [(4.8, 7.5, 7.5), (14.0, 7.4, 7.6)], 3, amp=0.0, stretch=(1, 1, 1))
    yy, zz = np.nonzero(flm[0] == 8)
    pre_idx = zz + NZ * (yy + NY * XC)
    main_idx = zz + NZ * (yy + NY * 0)
    shift = (-float(XC), 0.0, 0.0)
    return dict(par=par, flp=flp.reshape(-1), flm=flm.reshape(-1), domp=domp, domm=domm, rbc=rbc, cells=cells,
                pre_idx=pre_idx, main_idx=main_idx, shift=shift)


def oracle_pair(c):
    pre = O.OracleSim(c['domp'], c['flp'], c['par'].f_limit, BODY)
    main = O.OracleSim(c['domm'], c['flm'], c['par'].f_limit)
    pre.pop = O.init_equilibrium(c['domp'], 1.0, U0); main.pop = O.init_equilibrium(c['domm'], 1.0, U0)
    for s in (pre, main):
        s.add_celltype(c['rbc'], 1)
    pre.add_cells(0, c['cells'], [0, 1])
    cpl = O.PreInletCoupling(pre, main, c['pre_idx'], c['main_idx'], 0, float(NXP), c['shift'], SLAB[0], SLAB[1], ID_STRIDE)
    return pre, main, cpl


def oracle_step(pre, main, cpl):
    """one pass of the main loop of pipeflow_with_preinlet.cpp:162-171: iterate() on both sides, then applyPreInlet()"""
    pre.iterate(); main.iterate()
    cpl.apply_velocity()
    return cpl.apply_cells()
