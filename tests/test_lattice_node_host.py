"""CPU check of the CUDA kernels' per-node arithmetic: hemocell_b200/csrc/lattice_node.cuh is host + device code; here nvcc
compiles it for the CPU (tests/cpp/lattice_node_host.cu) and every node kind - fluid (generic and tau = 1 collision),
bounce-back, regularized velocity planes, Zou-He velocity and pressure nodes of all orientations - is compared with the
oracle's collision on the same populations.  The GPU parity tests remain the proof of the kernels themselves (memory
layout, streaming, exchanges); this keeps their arithmetic under test where no GPU exists."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle as O
import util as U

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
C = np.array([[0,0,0],[-1,0,0],[0,-1,0],[0,0,-1],[-1,-1,0],[-1,1,0],[-1,0,-1],[-1,0,1],[0,-1,-1],[0,-1,1],
              [1,0,0],[0,1,0],[0,0,1],[1,1,0],[1,-1,0],[1,0,1],[1,0,-1],[0,1,1],[0,1,-1]])


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    so = tmp_path_factory.mktemp("node_host") / "liblattice_node_host.so"
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
                           os.path.join(ROOT, "tests", "cpp", "lattice_node_host.cu"), "-o", str(so)])
    return ctypes.CDLL(str(so))


@pytest.mark.parametrize("tau", [0.8, 1.0, 1.6])
def test_node_arithmetic_matches_the_oracle(lib, tau):
    nx, ny, nz = 10, 9, 8
    N = nx * ny * nz
    rng = np.random.default_rng(17)
    fl = np.zeros((nx, ny, nz), dtype=np.uint8)
    fl[2:4, 2:5, 1:3] = 1                                       # a bounce-back block
    k = 0
    for o in range(6):                                         # isolated nodes of every boundary kind and orientation
        for base in (2, 8, 14):
            fl[1 + k % 8, 6 + (k // 8) % 2, 3 + k % 4] = base + o
            k += 1
    fl = fl.reshape(-1)
    bc_vel = 0.02 * rng.standard_normal((6, 3))
    dom = O.make_domain(nx, ny, nz, (1, 1, 1), tau, bc_vel)
    pop = U.smooth_state(dom, 5)
    force = np.ascontiguousarray(1e-4 * rng.standard_normal(3 * N))
    bc = np.zeros((4, N)); bc[3] = 1.0
    nodes = np.nonzero(fl >= 8)[0]
    bc[:, nodes] = np.column_stack([0.02 * rng.standard_normal((nodes.size, 3)), 1.0 + 2e-3 * rng.standard_normal(nodes.size)]).T
    bc = np.ascontiguousarray(bc.reshape(-1))
    # oracle: collide and stream, then undo the (periodic) streaming to get the post-collision populations
    ref = pop.copy()
    O.collide_and_stream(dom, fl, ref, force, bc_node=bc)
    r = ref.reshape(19, nx, ny, nz)
    post = np.stack([np.roll(r[q], shift=tuple(-C[q]), axis=(0, 1, 2)) for q in range(19)]).reshape(19, N)
    dp = ctypes.POINTER(ctypes.c_double)
    for tau1 in ([0, 1] if tau == 1.0 else [0]):
        out = np.empty((19, N))
        bv = np.ascontiguousarray(bc_vel)
        lib.node_collide_host(ctypes.c_int64(N), fl.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), pop.ctypes.data_as(dp),
                              force.ctypes.data_as(dp), ctypes.c_double(1.0 / tau), bv.ctypes.data_as(dp), bc.ctypes.data_as(dp),
                              ctypes.c_int(tau1), out.ctypes.data_as(dp))
        for name, sel in (("fluid", fl == 0), ("bounce-back", fl == 1), ("regularized planes", (fl >= 2) & (fl < 8)),
                          ("Zou-He velocity", (fl >= 8) & (fl < 14)), ("Zou-He pressure", fl >= 14)):
            assert sel.any()
            U.assert_close(out[:, sel], post[:, sel], f"{name} nodes (tau {tau}, tau1 path {tau1})", rtol=1e-12)
