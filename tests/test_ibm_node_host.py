"""CPU check of the immersed-boundary kernels' per-particle arithmetic (hemocell_b200/csrc/ibm_node.cuh, host + device code compiled for
the CPU by nvcc): the phi2 kernel of one particle and the unrolled velocity interpolation against the oracle, on a single slab and on
the two slabs of a decomposed lattice, with periodic wrap, non-periodic faces and non-fluid nodes."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle as O
import util as U

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    so = tmp_path_factory.mktemp("ibm_host") / "libibm_node_host.so"
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
                           os.path.join(ROOT, "tests", "cpp", "ibm_node_host.cu"), "-o", str(so)])
    L = ctypes.CDLL(str(so))
    L.ibm_kernel_host.restype = ctypes.c_int
    L.spread_corners_host.restype = ctypes.c_uint
    return L


def _slab(a, x0, nxl, nx, periodic_x, fill):
    """[nx, ...] -> padded slab [nxl + 2, ...]: ghost planes = neighbouring planes (periodic images) or `fill` outside the domain"""
    idx = np.arange(x0 - 1, x0 + nxl + 1)
    out = np.empty((nxl + 2,) + a.shape[1:], dtype=a.dtype)
    for k, x in enumerate(idx):
        if 0 <= x < nx or periodic_x:
            out[k] = a[x % nx]
        else:
            out[k] = fill
    return np.ascontiguousarray(out)


@pytest.mark.parametrize("periodic", [(1, 1, 1), (1, 1, 0), (0, 0, 0)])
def test_interpolation_and_kernel_match_the_oracle(lib, periodic):
    nx, ny, nz = 12, 9, 8
    N = nx * ny * nz
    rng = np.random.default_rng(23)
    fl = np.zeros((nx, ny, nz), dtype=np.uint8)
    fl[3:5, 2:4, 2:5] = 1
    if not periodic[2]:
        fl[:, :, 0] = 6; fl[:, :, -1] = 7
    flf = np.ascontiguousarray(fl.reshape(-1))
    dom = O.make_domain(nx, ny, nz, periodic, 0.9)
    pop = U.mask_inflow(dom, U.smooth_state(dom, 8))
    force = np.ascontiguousarray(1e-4 * rng.standard_normal(3 * N))
    rho, vel = O.moments(dom, flf, pop, force)
    u4 = np.concatenate([vel.reshape(3, nx, ny, nz), rho.reshape(1, nx, ny, nz)]).transpose(1, 2, 3, 0)     # [nx, ny, nz, 4]
    npart = 400
    pos = np.column_stack([rng.uniform(-1.5, nx + 0.5, npart), rng.uniform(-1.5, ny + 0.5, npart), rng.uniform(-1.5, nz + 0.5, npart)])
    pos[:20] = np.round(pos[:20])                              # vertices exactly on nodes (zero weights)
    pos = np.ascontiguousarray(pos)
    want = O.interpolate(dom, flf, pos, pop, force)
    per = (ctypes.c_int * 3)(*[int(p) for p in periodic])
    dp = ctypes.POINTER(ctypes.c_double); u8 = ctypes.POINTER(ctypes.c_uint8)
    # (a) one slab = the whole lattice; (b) two slabs: a vertex is interpolated by every rank that can address all its corners
    for nranks, slabs in ((1, [(0, nx)]), (2, [(0, 6), (6, 6)])):
        done = np.zeros(npart, dtype=bool)
        for x0, nxl in slabs:
            fs = _slab(fl, x0, nxl, nx, periodic[0], 1)
            us = _slab(u4, x0, nxl, nx, periodic[0], 0.0)
            got = np.zeros((npart, 3)); ok = np.zeros(npart, dtype=np.uint8)
            lib.ibm_interp_host(nx, ny, nz, per, x0, nxl, nranks, 1, fs.ctypes.data_as(u8), us.ctypes.data_as(dp), ctypes.c_int64(npart),
                                pos.ctypes.data_as(dp), got.ctypes.data_as(dp), ok.ctypes.data_as(u8))
            sel = ok.astype(bool)
            if nranks > 1:
                # a rank only ever evaluates the vertices of cells it holds, i.e. near its slab: both corners on its real or ghost planes
                sel &= (pos[:, 0] >= x0 - 1) & (pos[:, 0] < x0 + nxl)
            # the oracle leaves the velocity of a vertex without any fluid node at 0 / unchanged: compare where it has support
            has = np.isfinite(got[sel]).all(axis=1)
            U.assert_close(got[sel][has], want[sel][has], f"interpolated velocity (slab at {x0}, {nranks} rank(s))", rtol=1e-12)
            done |= sel
            # the (node, weight) pairs of the fallback spreading kernel against ora_ibm_kernel
            for p in np.nonzero(sel)[0][:60]:
                node = np.zeros(8, dtype=np.int64); w = np.zeros(8)
                n = lib.ibm_kernel_host(nx, ny, nz, per, x0, nxl, nranks, fs.ctypes.data_as(u8), pos[p].ctypes.data_as(dp),
                                        node.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), w.ctypes.data_as(dp))
                onode = np.zeros(8, dtype=np.int64); ow = np.zeros(8)
                m = O.lib().ora_ibm_kernel(ctypes.byref(dom), flf.ctypes.data_as(u8), pos[p].ctypes.data_as(dp),
                                           onode.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), ow.ctypes.data_as(dp))
                assert n == m, (p, pos[p], n, m)
                # local padded index -> global index
                lx = node[:n] // (ny * nz); rem = node[:n] % (ny * nz)
                gx = (lx - 1 + x0) % nx
                np.testing.assert_array_equal(gx * ny * nz + rem, onode[:m])
                U.assert_close(w[:n], ow[:m], "kernel weights", rtol=1e-14)
                # the same pairs as k_spread_sorted stages and reads them back (node key + wrap flags, factorised normalised weights):
                # the corners on this slab's REAL planes, in the fallback kernel's order
                valid = np.zeros(8, dtype=np.uint8); cn = np.zeros(8, dtype=np.int64); cw = np.zeros(8)
                lib.spread_corners_host(nx, ny, nz, per, x0, nxl, nranks, fs.ctypes.data_as(u8), pos[p].ctypes.data_as(dp),
                                        valid.ctypes.data_as(u8), cn.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), cw.ctypes.data_as(dp))
                v = valid.astype(bool)
                real = (lx >= 1) & (lx <= nxl)
                assert v.sum() == real.sum(), (p, pos[p])
                np.testing.assert_array_equal(cn[v], node[:n][real])
                if real.any():
                    U.assert_close(cw[v], ow[:m][real], "staged corner weights", rtol=1e-14)
        if nranks == 1:
            assert done.all() or not all(periodic)
