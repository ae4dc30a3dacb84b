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


def test_kernel_body_compiled_for_the_host_matches_the_numpy_restatement(tmp_path):
    """the per-node body of k_moment_step (hemocell_b200/csrc/moment_step.cuh, host + device code) compiled for the CPU by nvcc
    and run over a small lattice: node indexing with ghost planes, the y / z wrap, the arithmetic, the node velocity and the
    force reset agree with the numpy restatement above (which the previous test ties to the oracle's population path)"""
    import ctypes, os, shutil, subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        import pytest
        pytest.skip("nvcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = tmp_path / "libmoment_host.so"
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
                           os.path.join(root, "tests", "cpp", "moment_host.cu"), "-o", str(so)])
    lib = ctypes.CDLL(str(so))
    nx, ny, nz = 7, 6, 5
    rng = np.random.default_rng(9)
    W = np.zeros((4, nx, ny, nz)); W[0] = 1e-3 * rng.standard_normal((nx, ny, nz)); W[1:] = 0.02 * rng.standard_normal((3, nx, ny, nz))
    F = 1e-4 * rng.standard_normal((3, nx, ny, nz))
    body = np.array([3e-6, -1e-6, 2e-6])

    def padded(a4):      # [4, nx, ny, nz] -> AoS [(nx + 2), ny, nz, 4] with periodic ghost planes
        a = np.moveaxis(a4, 0, -1)
        return np.ascontiguousarray(np.concatenate([a[-1:], a, a[:1]], axis=0))

    Win = padded(W); Fin = padded(np.concatenate([F, np.zeros((1, nx, ny, nz))]))
    Wout = np.full_like(Win, np.nan); Fout = np.full_like(Win, np.nan); Uo = np.full_like(Win, np.nan)
    dp = ctypes.POINTER(ctypes.c_double)
    lib.moment_host(nx, ny, nz, Win.ctypes.data_as(dp), Fin.ctypes.data_as(dp), Wout.ctypes.data_as(dp), Fout.ctypes.data_as(dp),
                    Uo.ctypes.data_as(dp), body.ctypes.data_as(dp), 1)
    ref = moment_step(W, F)
    got = np.moveaxis(Wout[1:-1], -1, 0)
    assert np.max(np.abs(got - ref)) < 1e-15 + 1e-13 * np.max(np.abs(ref))
    u = ref[1:4] / (1.0 + ref[0]) + 0.5 * F
    assert np.max(np.abs(np.moveaxis(Uo[1:-1], -1, 0)[:3] - u)) < 1e-15
    assert np.max(np.abs(Uo[1:-1, ..., 3] - (1.0 + ref[0]))) < 1e-15
    assert np.all(Fout[1:-1, ..., :3] == body) and np.all(Fout[1:-1, ..., 3] == 0.0)
    assert np.isnan(Wout[0]).all() and np.isnan(Wout[-1]).all()            # ghost planes are the caller's business
    # the re-associated evaluation of the 19 post-collision populations (k_moment_tile) against guo_collide_tau1
    n = 5000
    w = np.zeros((n, 4)); w[:, 0] = 5e-3 * rng.standard_normal(n); w[:, 1:] = 0.05 * rng.standard_normal((n, 3))
    f = 1e-3 * rng.standard_normal((n, 3))
    fast = np.empty((n, 19)); ref = np.empty((n, 19))
    lib.tau1_pops_both(n, w.ctypes.data_as(dp), f.ctypes.data_as(dp), fast.ctypes.data_as(dp), ref.ctypes.data_as(dp))
    assert np.max(np.abs(fast - ref)) < 2e-17 + 1e-14 * np.max(np.abs(ref))
