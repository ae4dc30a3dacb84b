"""Shared builders for the parity tests: the same seeded inputs go to the CPU oracle and,
through the C ABI, to the CUDA path."""
import numpy as np
import oracle as O
from oracle import mesh as M

RTOL = 1e-12        # north_star: per-operator fields match to 1e-12 relative in fp64
FLOOR = 1e-14       # absolute floor as a fraction of the field scale (near-cancelling sums)


def assert_close(got, ref, what, rtol=RTOL, floor=FLOOR):
    got = np.asarray(got, dtype=np.float64).reshape(-1)
    ref = np.asarray(ref, dtype=np.float64).reshape(-1)
    assert got.shape == ref.shape, what
    assert np.all(np.isfinite(got)), what + ": non-finite values"
    scale = float(np.max(np.abs(ref))) if ref.size else 0.0
    tol = rtol * np.maximum(np.abs(got), np.abs(ref)) + floor * scale
    err = np.abs(got - ref)
    bad = err > tol
    if bad.any():
        i = int(np.argmax(err - tol))
        raise AssertionError(f"{what}: {int(bad.sum())}/{ref.size} entries differ; worst idx {i}: got {got[i]!r} "
                             f"ref {ref[i]!r} err {err[i]:.3e} tol {tol[i]:.3e} (scale {scale:.3e})")


def smooth_state(dom, seed, amp_u=0.03, amp_rho=1e-3):
    """equilibrium populations of a smooth density/velocity field (post-stream layout)"""
    rng = np.random.default_rng(seed)
    nx, ny, nz = dom.nx, dom.ny, dom.nz
    x, y, z = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing='ij')
    ph = rng.uniform(0, 2 * np.pi, 6)
    rho = 1.0 + amp_rho * np.sin(2 * np.pi * x / nx + ph[0]) * np.cos(2 * np.pi * y / ny + ph[1])
    u = np.stack([amp_u * np.sin(2 * np.pi * y / ny + ph[2]) * np.cos(2 * np.pi * z / nz + ph[3]),
                  amp_u * np.sin(2 * np.pi * z / nz + ph[4]),
                  amp_u * np.cos(2 * np.pi * x / nx + ph[5])])
    C = np.array([[0,0,0],[-1,0,0],[0,-1,0],[0,0,-1],[-1,-1,0],[-1,1,0],[-1,0,-1],[-1,0,1],[0,-1,-1],[0,-1,1],
                  [1,0,0],[0,1,0],[0,0,1],[1,1,0],[1,-1,0],[1,0,1],[1,0,-1],[0,1,1],[0,1,-1]], dtype=np.float64)
    T = np.array([1/3] + [1/18]*3 + [1/36]*6 + [1/18]*3 + [1/36]*6)
    N = nx * ny * nz
    pop = np.empty((19, N))
    rhoBar = (rho - 1.0).reshape(-1)
    j = (rho[None] * u).reshape(3, -1)
    jsq = (j * j).sum(0)
    inv = 1.0 / rho.reshape(-1)
    for q in range(19):
        cj = C[q, 0] * j[0] + C[q, 1] * j[1] + C[q, 2] * j[2]
        pop[q] = T[q] * (rhoBar + 3 * cj + inv * (4.5 * cj * cj - 1.5 * jsq))
    # add a non-equilibrium perturbation so that relaxation is exercised
    pop += 1e-5 * rng.standard_normal(pop.shape)
    return np.ascontiguousarray(pop.reshape(-1))


def mask_inflow(dom, pop):
    """zero the populations that would have streamed in through a non-periodic face: by
    definition (include/hemocell_gpu.h) what enters there is the rest equilibrium, stored 0"""
    C = [(0,0,0),(-1,0,0),(0,-1,0),(0,0,-1),(-1,-1,0),(-1,1,0),(-1,0,-1),(-1,0,1),(0,-1,-1),(0,-1,1),
         (1,0,0),(0,1,0),(0,0,1),(1,1,0),(1,-1,0),(1,0,1),(1,0,-1),(0,1,1),(0,1,-1)]
    n = (dom.nx, dom.ny, dom.nz)
    p = pop.reshape(19, dom.nx, dom.ny, dom.nz)
    for q, c in enumerate(C):
        for ax in range(3):
            if dom.periodic[ax] or c[ax] == 0:
                continue
            sl = [slice(None)] * 3
            sl[ax] = 0 if c[ax] > 0 else n[ax] - 1
            p[(q,) + tuple(sl)] = 0.0
    return pop


def couette_flags(nx, ny, nz):
    fl = np.zeros((nx, ny, nz), dtype=np.uint8)
    fl[:, :, 0] = 6
    fl[:, :, nz - 1] = 7
    return fl


def box_flags(nx, ny, nz):
    fl = np.zeros((nx, ny, nz), dtype=np.uint8)
    fl[:, :, 0] = 6; fl[:, :, nz - 1] = 7
    fl[:, 0, :] = 4; fl[:, ny - 1, :] = 5
    fl[0, :, :] = 2; fl[nx - 1, :, :] = 3
    return fl


def gpu_context(dom, flags, bc_vel=None, body=None, device=0):
    from hemocell_b200 import lib as H
    ctx = H.Context(dom.nx, dom.ny, dom.nz, [dom.periodic[k] for k in range(3)], 1.0 / dom.omega, device=device)
    ctx.set_flags(flags)
    if bc_vel is not None:
        for o in range(6):
            ctx.set_bc_velocity(o, bc_vel[o])
    if body is not None:
        ctx.set_body_force(body)
    return ctx


def gpu_add_type(ctx, ct):
    return ctx.add_celltype(ct.model, ct.cc, ct.k)


def deformed_cells(ct, centers, seed, amp=0.03, stretch=(1.08, 0.97, 0.96)):
    """n cells: reference mesh, anisotropically stretched, rotated a bit, noise added"""
    rng = np.random.default_rng(seed)
    out = []
    v0 = ct.verts - 0.5 * (ct.verts.min(0) + ct.verts.max(0))
    for c in centers:
        a = M.rotation_xyz(*rng.uniform(-1, 1, 3))
        v = (v0 * np.array(stretch)) @ a.T
        v = v + amp * rng.standard_normal(v.shape)
        out.append(v + np.asarray(c, dtype=np.float64))
    return np.ascontiguousarray(np.array(out))
