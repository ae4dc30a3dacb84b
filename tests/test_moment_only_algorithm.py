"""CPU check of the algorithm behind the opt-in moment-only update (hemocell_b200/csrc/lattice.cu: k_moment_step): at tau = 1 on a
periodic lattice without walls, the next post-stream moments of a node are sums over the populations its 19 upstream
neighbours emit, and those are functions of the neighbours' four raw moments and force alone.  The numpy restatement below is
the kernel's arithmetic; the checker is the oracle's population path (collide_and_stream + moments)."""
import numpy as np

import oracle as O

C = np.array([[0,0,0],[-1,0,0],[0,-1,0],[0,0,-1],[-1,-1,0],[-1,1,0],[-1,0,-1],[-1,0,1],[0,-1,-1],[0,-1,1],
              [1,0,0],[0,1,0],[0,0,1],[1,1,0],[1,-1,0],[1,0,1],[1,0,-1],[0,1,1],[0,1,-1]])
T = np.array([1/3] + [1/18]*3 + [1/36]*6 + [1/18]*3 + [1/36]*6)


def moment_step(W, F):
    """W [4, nx, ny, nz] = (rhoBar, j), F [3, nx, ny, nz] -> W' (k_moment_step)"""
    rho = 1.0 + W[0]; inv = 1.0 / rho
    u = W[1:4] * inv + 0.5 * F
    j = rho * u
    jsq = (j * j).sum(0); uF = (u * F).sum(0)
    out = np.zeros_like(W)
    for q in range(19):
        cj = sum(C[q, k] * j[k] for k in range(3)); cu = sum(C[q, k] * u[k] for k in range(3)); cF = sum(C[q, k] * F[k] for k in range(3))
        fq = T[q] * (W[0] + 3.0 * cj + inv * (4.5 * cj * cj - 1.5 * jsq)) + T[q] * 0.5 * (3.0 * (cF - uF) + 9.0 * cu * cF)
        arriving = np.roll(fq, shift=tuple(C[q]), axis=(0, 1, 2))          # S_q(m) = f*_q(m - c_q)
        out[0] += arriving
        for k in range(3):
            out[1 + k] += C[q, k] * arriving
    return out


def test_moment_only_update_equals_the_population_path():
    nx, ny, nz = 12, 10, 8
    N = nx * ny * nz
    dom = O.make_domain(nx, ny, nz, (1, 1, 1), 1.0)
    fl = np.zeros(N, np.uint8)
    rng = np.random.default_rng(4)
    pop = O.init_equilibrium(dom, 1.0, (0.01, -0.02, 0.015)) + 1e-4 * rng.standard_normal(19 * N)
    p = pop.reshape(19, nx, ny, nz)
    W = np.stack([p.sum(0)] + [sum(C[q, k] * p[q] for q in range(19)) for k in range(3)])
    for step in range(4):
        force = np.ascontiguousarray(1e-4 * rng.standard_normal(3 * N))
        O.collide_and_stream(dom, fl, pop, force)
        W = moment_step(W, force.reshape(3, nx, ny, nz))
        p = pop.reshape(19, nx, ny, nz)
        ref = np.stack([p.sum(0)] + [sum(C[q, k] * p[q] for q in range(19)) for k in range(3)])
        assert np.max(np.abs(W - ref)) < 1e-15 + 1e-13 * np.max(np.abs(ref)), step
        # the node velocity the IBM reads: u = j / rho + F / 2 of the post-stream state with this step's force
        rho, vel = O.moments(dom, fl, pop, force)
        u = W[1:4] / (1.0 + W[0]) + 0.5 * force.reshape(3, nx, ny, nz)
        assert np.max(np.abs(u.reshape(-1) - vel)) < 1e-15
