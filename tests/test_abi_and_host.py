"""CPU tests of the boundary: the C-ABI library loads and exports every symbol the headers
declare; the product's C++ host set-up code agrees with the numpy oracle; creating a context
without a GPU fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import numpy as np
import pytest

import oracle as O
from oracle import mesh as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header, prefix):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from hemocell_b200 import lib as H
    L = H.load()
    dev = _declared("hemocell_gpu.h", "hcg")
    host = _declared("hemocell_host.h", "hch")
    assert len(dev) >= 45 and len(host) >= 9
    for name in dev + host:
        assert hasattr(L, name), f"{name} declared in include/ but not exported"
    assert sorted(H.SYMBOLS) == dev
    assert sorted(H.HOST_SYMBOLS) == host
    assert b"sm_100a" in L.hcg_version()


def test_no_cpu_fallback():
    from hemocell_b200 import lib as H
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(H.HcgError, match="no CUDA device"):
        H.Context(8, 8, 8, (1, 1, 1), 1.0)


def test_argument_validation():
    from hemocell_b200 import lib as H
    L = H.load()
    d = H.HcgDomain(); d.nx, d.ny, d.nz = 10, 10, 10; d.tau = 0.4; d.n_ranks = 1
    h = C.c_void_p()
    assert L.hcg_create(C.byref(d), C.byref(h)) == -1            # tau <= 0.5
    d.tau = 1.0; d.n_ranks = 3; d.rank = 3
    assert L.hcg_create(C.byref(d), C.byref(h)) == -1            # rank outside [0, n_ranks)
    assert b"rank" in L.hcg_last_error(None)
    # uneven slabs: the first nx % n_ranks ranks own one plane more, the slabs tile [0, nx)
    for nx, R in ((103, 2), (10, 3), (256, 8), (7, 7)):
        x_next = 0
        for r in range(R):
            x0, nxl = C.c_int32(), C.c_int32()
            L.hcg_slab(C.c_int32(nx), C.c_int32(r), C.c_int32(R), C.byref(x0), C.byref(nxl))
            assert x0.value == x_next and nxl.value in (nx // R, nx // R + 1)
            x_next += nxl.value
        assert x_next == nx


@pytest.mark.parametrize("kind", ["rbc", "plt"])
def test_host_cpp_matches_numpy_oracle(kind):
    from hemocell_b200 import lib as H
    par_o = M.Parameters(0.5e-6, 1e-7)
    par = H.parameters(0.5e-6, 1e-7)
    for k in ("tau", "nu_lbm", "df", "f_limit", "kBT_lbm"):
        assert par[k] == getattr(par_o, k)
    if kind == "rbc":
        h = H.HostCellType(H.MODEL_RBC, H.RBC_FROM_SPHERE, par, H.RBC_MATERIAL)
        ct = O.rbc_celltype(par_o)
    else:
        h = H.HostCellType(H.MODEL_PLT, H.ELLIPSOID_FROM_SPHERE, par, H.PLT_MATERIAL, H.PLT_INNER_EDGES)
        ct = O.plt_celltype(par_o)
    v = h.view.contents
    V, T, E, I = v.n_vertices, v.n_triangles, v.n_edges, v.n_inner_edges
    assert (V, T, E, I) == (ct.V, ct.cc["triangle_list"].shape[0], ct.cc["edge_list"].shape[0], ct.cc["inner_edge_list"].shape[0])
    np.testing.assert_allclose(h.verts, ct.verts, rtol=0, atol=1e-13)
    for name, shape, key in [("triangles", (T, 3), "triangle_list"), ("edges", (E, 2), "edge_list"),
                             ("vertex_vertexes", (V, 6), "vertex_vertexes"), ("vertex_n_vertexes", (V,), "vertex_n_vertexes"),
                             ("edge_bending_triangles", (E, 2), "edge_bending_triangles_list"),
                             ("edge_bending_outer_points", (E, 2), "edge_bending_triangles_outer_points")]:
        assert np.array_equal(h.table(name, shape, np.int32), ct.cc[key]), name
    for name, shape, key in [("edge_length_eq", (E,), "edge_length_eq_list"), ("edge_angle_eq", (E,), "edge_angle_eq_list"),
                             ("triangle_area_eq", (T,), "triangle_area_eq_list"),
                             ("patch_dist_eq", (V,), "surface_patch_center_dist_eq_list"),
                             ("inner_edge_length_eq", (I,), "inner_edge_length_eq_list")]:
        np.testing.assert_allclose(h.table(name, shape, np.float64), ct.cc[key], rtol=1e-12, atol=1e-14, err_msg=name)
    np.testing.assert_allclose([v.volume_eq, v.area_mean_eq, v.edge_mean_eq], [ct.cc[k] for k in ("volume_eq", "area_mean_eq", "edge_mean_eq")], rtol=1e-13)
    np.testing.assert_allclose([v.k_volume, v.k_area, v.k_link, v.k_bend, v.eta_m],
                               [ct.k[k] for k in ("k_volume", "k_area", "k_link", "k_bend", "eta_m")], rtol=1e-15)


def test_host_placement_matches_oracle(tmp_path):
    from hemocell_b200 import lib as H
    par = H.parameters(0.5e-6, 1e-7)
    h = H.HostCellType(H.MODEL_RBC, H.RBC_FROM_SPHERE, par, H.RBC_MATERIAL)
    ct = O.rbc_celltype(M.Parameters(0.5e-6, 1e-7))
    rng = np.random.default_rng(2)
    rows = np.concatenate([rng.uniform(0, 32, (40, 3)), rng.uniform(-180, 180, (40, 3))], axis=1)
    dims = (64, 48, 40)
    fl = np.zeros(dims, dtype=np.uint8); fl[:, 0, :] = 1; fl[:, -1, :] = 1; fl[:, :, 0] = 6; fl[:, :, -1] = 7
    p = tmp_path / "RBC.pos"
    p.write_text("40\n" + "\n".join(" ".join(repr(float(x)) for x in r) for r in rows) + "\n")
    rows_rd = H.read_pos(str(p))
    assert np.array_equal(rows_rd, rows)
    for md in (0.0, 1.0):
        pos_h, ids_h = h.place(rows_rd, 0.5e-6, dims, fl, md, cell_id0=7)
        pos_o, ids_o = M.place_cells(ct.verts, rows, 0.5e-6, dims, fl.reshape(-1), md, cell_id0=7)
        assert ids_h.tolist() == ids_o.tolist() and len(ids_h) > 0
        np.testing.assert_allclose(pos_h, pos_o, rtol=0, atol=1e-12)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree not present (authoring container only)")
def test_voxeliser_and_placement_reproduce_the_references_42_cells():
    """known answer from the reference's own validation test (tests/validation/pipeflow/test_pipeflow.cpp:88-92):
    after voxelising examples/pipeflow/tube.stl at refDirN 50 and placing the shipped RBC / PLT position files,
    exactly 42 cells survive.  Pins the voxeliser conventions (dx = extent / refDirN, margin 1, N + 1 nodes) and the
    placement filter, including the reference's quirk that the 0.5 um minimum wall distance is stored in an
    unsigned int and therefore is 0 (core/hemoCellField.h:64)."""
    from hemocell_b200 import lib as H
    fl, dx = H.voxelize_stl("/root/reference/examples/pipeflow/tube.stl", 50, 1)
    assert fl.shape == (103, 53, 53) and abs(dx - 0.2) < 1e-12
    area = int((fl[0] == 0).sum())
    assert area == int((fl[50] == 0).sum()) and abs(np.sqrt(area / np.pi) - 24.9) < 0.1     # open ends, radius ~ 25 lu
    par = H.parameters(0.5e-6, 1e-7)
    rbc = H.HostCellType(H.MODEL_RBC, H.RBC_FROM_SPHERE, par, H.RBC_MATERIAL)
    plt = H.HostCellType(H.MODEL_PLT, H.ELLIPSOID_FROM_SPHERE, par, H.PLT_MATERIAL, H.PLT_INNER_EDGES)
    d = "/root/reference/tests/validation/pipeflow"
    rr, pr = H.read_pos(d + "/RBC.pos"), H.read_pos(d + "/PLT.pos")
    _, rid = rbc.place(rr, 0.5e-6, fl.shape, fl.reshape(-1), min_dist_um=float(int(0.5)))
    _, pid = plt.place(pr, 0.5e-6, fl.shape, fl.reshape(-1), min_dist_um=0.0, cell_id0=len(rr))
    assert len(rid) + len(pid) == 42, (len(rid), len(pid))


def test_voxeliser_binary_stl_box(tmp_path):
    """a 10 x 6 x 4 box written as a binary STL: dx = extent / refDirN, margin of one node, nodes on the (inflated)
    surface count as inside, and the two x ends are opened as helper/voxelizeDomain.cpp:142-156 does"""
    import struct
    from hemocell_b200 import lib as H
    lo, hi = np.array([2.0, -1.0, 5.0]), np.array([12.0, 5.0, 9.0])
    c = [np.array([x, y, z]) for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])]
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]     # outward orientation
    tris = []
    for a, b, cc, d in quads:
        tris += [(c[a], c[b], c[cc]), (c[a], c[cc], c[d])]
    p = tmp_path / "box.stl"
    with open(p, "wb") as f:
        f.write(b"binary box".ljust(80, b" ")); f.write(struct.pack("<I", len(tris)))
        for t in tris:
            n = np.cross(t[1] - t[0], t[2] - t[0]); n = n / np.linalg.norm(n)
            f.write(struct.pack("<12fH", *n, *t[0], *t[1], *t[2], 0))
    fl, dx = H.voxelize_stl(p, 20, 0)                       # 20 cells along x: dx = 0.5
    assert abs(dx - 0.5) < 1e-12 and fl.shape == (23, 15, 11)
    fluid = fl == 0
    assert fluid[:, 1:14, 1:10].all() and not fluid[:, 0, :].any() and not fluid[:, :, 10].any()   # open x ends, closed elsewhere
    assert int(fluid.sum()) == 23 * 13 * 9
