"""CPU check of the RBC mechanics kernel's algorithm (hemocell_b200/csrc/mechanics.cu):

* the packed ring table hcg_celltype_add hands the kernel (hemocell_b200/csrc/mech_tables.h, compiled here with g++) names, for every
  (vertex, ring slot), the right neighbour, edge, triangle, ring size, third-vertex code and orientation, and covers every edge twice
  and every triangle three times;
* the kernel's four steps - per-triangle area-force magnitude and signed-volume term, per-edge link scalar, per-patch bending force,
  one ring walk per vertex driven by that table - restated in numpy with the kernel's operand order reproduce the oracle's
  RbcHighOrderModel forces (reference mechanics/rbcHighOrderModel.cpp:38-207) within the 1e-12 of the GPU parity test, on the total
  and on every force family, with and without membrane viscosity."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle as O
from oracle import mesh as M
import util as U

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    so = tmp_path_factory.mktemp("mech_tables") / "libmech_tables_host.so"
    subprocess.check_call([gxx, "-O2", "-std=c++17", "-shared", "-fPIC", os.path.join(ROOT, "tests", "cpp", "mech_tables_host.cpp"), "-o", str(so)])
    L = ctypes.CDLL(str(so))
    L.mech_ring_table.restype = ctypes.c_int
    return L


def ring_table(L, ct):
    a = ct._arrs
    tri, edges, vv, nvv = a['triangles'], a['edges'], a['vv'], a['nvv']
    V, T, E = ct.V, len(tri), len(edges)
    out = np.zeros(6 * V, dtype=np.uint64)
    err = ctypes.create_string_buffer(200)
    ip = ctypes.POINTER(ctypes.c_int)
    rc = L.mech_ring_table(V, T, E, tri.ctypes.data_as(ip), edges.ctypes.data_as(ip), vv.ctypes.data_as(ip), nvv.ctypes.data_as(ip),
                           out.ctypes.data_as(ctypes.POINTER(ctypes.c_ulonglong)), err, 200)
    assert rc == 0, err.value
    rg = out.reshape(6, V)
    f = lambda sh, m: ((rg >> np.uint64(sh)) & np.uint64(m)).astype(np.int64)
    return dict(ring=f(0, 0xffff), edge=f(16, 0xffff), tri=f(32, 0xffff), nn=f(48, 7), last=f(51, 3), neg=f(53, 1))


@pytest.mark.parametrize("min_triangles", [600, 200])          # 642 / 1280 / 1920 and the coarser 162 / 320 / 480 mesh
def test_ring_table_names_the_right_elements(lib, min_triangles):
    ct = O.rbc_celltype(M.Parameters(dx=0.5e-6, dt=1e-7), dict(M.RBC_MATERIAL, minNumTriangles=min_triangles))
    a = ct._arrs
    tri, edges, vv, nvv = a['triangles'], a['edges'], a['vv'].reshape(-1, 6), a['nvv']
    V, T, E = ct.V, len(tri), len(edges)
    tb = ring_table(lib, ct)
    edge_hits = np.zeros(E + 1, dtype=int); tri_hits = np.zeros(T + 1, dtype=int)
    for v in range(V):
        n = nvv[v]
        for j in range(6):
            r, e, t, rn, last, neg = (tb[k][j, v] for k in ("ring", "edge", "tri", "nn", "last", "neg"))
            edge_hits[e] += 1; tri_hits[t] += 1
            if j >= n:                                      # unused slot: null edge, null triangle, weight 0, a valid position
                assert (r, e, t, rn) == (vv[v, 0], E, T, 0)
                continue
            ia, ib = vv[v, j], vv[v, (j + 1) % n]
            assert r == ia and rn == nvv[ia]
            assert set(edges[e]) == {v, ia}
            assert set(tri[t]) == {v, ia, ib}
            assert tri[t][2] == (v, ia, ib)[last]
            cyc = [(v, ia, ib), (ia, ib, v), (ib, v, ia)]
            assert (tuple(tri[t]) in cyc) == (neg == 0)
    assert np.all(edge_hits[:E] == 2) and np.all(tri_hits[:T] == 3)
    assert edge_hits[E] == tri_hits[T] == 6 * V - int(nvv.sum())


def test_ring_table_rejects_broken_topology(lib):
    ct = O.rbc_celltype(M.Parameters(dx=0.5e-6, dt=1e-7))
    a = ct._arrs
    vv = a['vv'].copy().reshape(-1, 6)
    vv[5, 0], vv[5, 2] = vv[5, 2], vv[5, 0]                 # ring no longer cyclic: two consecutive neighbours share no triangle
    out = np.zeros(6 * ct.V, dtype=np.uint64)
    err = ctypes.create_string_buffer(200)
    ip = ctypes.POINTER(ctypes.c_int)
    vvc = np.ascontiguousarray(vv.reshape(-1))
    rc = lib.mech_ring_table(ct.V, len(a['triangles']), len(a['edges']), a['triangles'].ctypes.data_as(ip), a['edges'].ctypes.data_as(ip),
                             vvc.ctypes.data_as(ip), a['nvv'].ctypes.data_as(ip), out.ctypes.data_as(ctypes.POINTER(ctypes.c_ulonglong)), err, 200)
    assert rc == 1 and b"triangle" in err.value


def kernel_arithmetic(ct, tb, X, VEL):
    """the four steps of k_mechanics<RBC> for one cell; returns (total, [area, volume, bending, link, viscosity])"""
    a, k, cc = ct._arrs, ct.k, ct.cc
    tri, edges, nvv = a['triangles'], a['edges'], a['nvv']
    V = ct.V
    v0, v1, v2 = X[tri[:, 0]], X[tri[:, 1]], X[tri[:, 2]]
    # 1. per triangle (every product and sum rounded separately, in this order)
    VT = (((((-(v2[:, 0] * v1[:, 1]) * v0[:, 2]) + (v1[:, 0] * v2[:, 1]) * v0[:, 2]) + (v2[:, 0] * v0[:, 1]) * v1[:, 2])
           - (v0[:, 0] * v2[:, 1]) * v1[:, 2]) - (v1[:, 0] * v0[:, 1]) * v2[:, 2]) + (v0[:, 0] * v1[:, 1]) * v2[:, 2]
    cr = np.cross(v1 - v0, v2 - v0)
    area = 0.5 * np.sqrt((cr * cr).sum(1))
    ar = (area - a['ta']) / a['ta']
    AFM = np.append(k['k_area'] * (ar + ar / np.abs(0.09 - ar * ar)), 0.0)
    # 2. sequential volume sum | per edge | per patch
    vol = 0.0
    for x in VT:
        vol += x
    vol *= 1.0 / 6.0
    ev = X[edges[:, 1]] - X[edges[:, 0]]
    ln = np.sqrt((ev * ev).sum(1)); il = 1.0 / ln
    fr = (ln - a['el']) / a['el']
    EF = np.append((k['k_link'] * (fr + fr / np.abs(9.0 - fr * fr))) * il, 0.0)
    eta = k['eta_m']
    uv = ev * il[:, None]
    pr = ((VEL[edges[:, 1]] - VEL[edges[:, 0]]) * uv).sum(1)
    g = eta * (pr[:, None] * uv); m2 = (g * g).sum(1)
    EG = (eta * pr) * il
    EG = np.append(np.where(m2 > 156.25, EG * (12.5 / np.sqrt(np.maximum(m2, 1e-300))), EG), 0.0)
    BF = np.zeros((V, 3))
    inv_edge_mean = 1.0 / cc['edge_mean_eq']
    for i in range(V):
        nn = nvv[i]; ring = tb['ring'][:nn, i]; xi = X[i]
        s = X[ring[0]].copy(); prev = X[ring[0]] - xi; first = prev; pn = np.zeros(3)
        for j in range(nn):
            if j + 1 < nn:
                r = X[ring[j + 1]]; s = s + r; nxt = r - xi
            else:
                nxt = first
            tn = np.cross(prev, nxt); pn = pn + tn * (1.0 / np.sqrt(tn @ tn)); prev = nxt
        dev = s / nn - xi
        pn = pn * (1.0 / np.sqrt(pn @ pn))
        dD = (pn @ dev - a['pd'][i]) * inv_edge_mean
        BF[i] = (k['k_bend'] * (dD + dD / abs(0.0555 - dD * dD))) * pn
    # 3. volume force
    vf = (vol - cc['volume_eq']) / cc['volume_eq']
    vf_half = ((-k['k_volume'] * vf / abs(0.01 - vf * vf)) * 0.5) * (1.0 / cc['area_mean_eq'])
    # 4. one ring walk per vertex
    ninv = np.array([0.0] + [-1.0 / n for n in range(1, 8)])
    third = 1.0 / 3.0
    tot = np.zeros((V, 3)); fam = [np.zeros((V, 3)) for _ in range(5)]
    for v in range(V):
        xv = X[v]
        fa = np.zeros(3); fw = np.zeros(3); fb = BF[v].copy(); fl = np.zeros(3); fs = np.zeros(3)
        for j in range(6):
            ra = X[tb['ring'][j, v]]; rb = X[tb['ring'][(j + 1) % 6, v]]
            e, t, last = tb['edge'][j, v], tb['tri'][j, v], tb['last'][j, v]
            da, db = ra - xv, rb - xv
            fl = fl + da * EF[e]; fs = fs + da * EG[e]
            fb = fb + BF[tb['ring'][j, v]] * ninv[tb['nn'][j, v]]
            p, q, l = (ra, rb, xv) if last == 0 else ((xv, rb, ra) if last == 1 else (xv, ra, rb))
            fa = fa + AFM[t] * (((p + q) + l) * third - xv)
            fw = fw + (-vf_half if tb['neg'][j, v] else vf_half) * np.cross(da, db)
        F = ((fa + fw) + fb) + fl
        if eta != 0.0:
            F = F + fs
        tot[v] = F
        for q_, x in enumerate((fa, fw, fb, fl, fs)):
            fam[q_][v] = x
    return tot, fam


@pytest.mark.parametrize("kind", ["rbc", "rbc_visc"])
def test_ring_walk_reproduces_the_oracle_forces(lib, kind):
    par = M.Parameters(dx=0.5e-6, dt=1e-7)
    ct = O.rbc_celltype(par, dict(M.RBC_MATERIAL, eta_m=5e-10)) if kind == "rbc_visc" else O.rbc_celltype(par)
    tb = ring_table(lib, ct)
    centers = [(30.2, 40.7, 50.1), (201.5, 120.25, 77.0)]                # the second: absolute coordinates that cost the centroid five digits
    cells = U.deformed_cells(ct, centers, 9, amp=0.02)
    pos = np.ascontiguousarray(cells.reshape(-1, 3))
    vel = np.ascontiguousarray(1e-3 * np.random.default_rng(1).standard_normal(pos.shape))
    f_ref = np.zeros_like(pos)
    comp = O.mechanics(ct, pos, vel, f_ref, components=True)
    V = ct.V
    got = np.zeros_like(pos); fam = [np.zeros_like(pos) for _ in range(5)]
    for ci in range(len(centers)):
        sl = slice(ci * V, (ci + 1) * V)
        got[sl], f5 = kernel_arithmetic(ct, tb, pos[sl], vel[sl])
        for q in range(5):
            fam[q][sl] = f5[q]
    scale = np.abs(f_ref).max()
    U.assert_close(got, f_ref, f"total vertex force ({kind})")
    for q, name in enumerate(["area", "volume", "bending", "link", "visc"]):
        err = np.abs(fam[q] - comp[q])
        tol = U.RTOL * np.maximum(np.abs(fam[q]), np.abs(comp[q])) + U.FLOOR * scale
        assert np.all(err <= tol), f"{name} force ({kind}): max err {err.max():.3e}"
