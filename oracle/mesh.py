"""CPU ORACLE (numpy) for the set-up side of the HemoCell hot path: reference-shape meshes,
CommonCellConstants topology tables, Parameters unit conversion and the .pos reader.

TEST INFRASTRUCTURE ONLY -- the product's own implementation is C++
(hemocell_b200/csrc/host_*.cpp); this file restates the reference independently so the two
can be compared.  Citations are relative to /root/reference (UvaCsl/HemoCell).

Palabos pieces that are NOT in the reference tree and are restated from memory (SURVEY.md
Appendix C/E, "parity unpinned"): TriangleSet::rotate Euler convention, vertex numbering of
DEFscaledMesh/TriangleBoundary3D (first appearance), constructSphere (octahedron),
TriangularSurfaceMesh::inflate (treated as a no-op, see INFLATE).
"""
import math
import numpy as np

PI = 3.14159265358979323846
# TriangularSurfaceMesh::inflate() default displacement in lattice units.  Unknown here (Appendix
# E.8); 1e-3 along the normalised sum of the adjacent unit triangle normals puts the RBC volume
# at 81.117 um^3, the only value consistent with scripts/ci/stretchCell_sanity.sh:15-26
# (81.12 <= V <= 81.19 while 100 % <= V/V_eq <= 100.1 %; without inflation V_eq = 81.052).
INFLATE = 1.0e-3


# ----------------------------------------------------------------------------- shapes
def _refine(tris, min_tri):
    """refinement loop of helper/meshGeneratingFunctions.hh:108-141 (same in Palabos constructSphere)"""
    tris = [list(t) for t in tris]
    while len(tris) < min_tri:
        size = len(tris)
        for i in range(size):
            va, vb, vc = tris[i]
            vd = 0.5 * (va + vb)
            ve = 0.5 * (vb + vc)
            vf = 0.5 * (vc + va)
            vd = vd / np.sqrt(np.dot(vd, vd))
            ve = ve / np.sqrt(np.dot(ve, ve))
            vf = vf / np.sqrt(np.dot(vf, vf))
            tris[i] = [vd, ve, vf]
            tris.append([va, vd, vf])
            tris.append([vd, vb, ve])
            tris.append([vf, ve, vc])
    return np.array(tris, dtype=np.float64)  # [T,3,3]


def sphere_icosahedron(min_tri):
    """constructSphereIcosahedron, helper/meshGeneratingFunctions.hh:31-151 (unit radius, origin)"""
    tau = -0.8506508084
    one = -0.5257311121
    v = {1: (tau, one, 0), 2: (-tau, one, 0), 3: (-tau, -one, 0), 4: (tau, -one, 0),
         5: (one, 0, tau), 6: (one, 0, -tau), 7: (-one, 0, -tau), 8: (-one, 0, tau),
         9: (0, tau, one), 10: (0, -tau, one), 11: (0, -tau, -one), 12: (0, tau, -one)}
    v = {k: np.array(x, dtype=np.float64) for k, x in v.items()}
    faces = [(5, 8, 9), (5, 10, 8), (6, 12, 7), (6, 7, 11), (1, 4, 5), (1, 6, 4), (3, 2, 8), (3, 7, 2),
             (9, 12, 1), (9, 2, 12), (10, 4, 11), (10, 11, 3), (9, 1, 5), (12, 6, 1), (5, 4, 10),
             (6, 11, 4), (8, 2, 9), (7, 12, 2), (8, 10, 3), (7, 3, 11)]
    return _refine([[v[a], v[b], v[c]] for a, b, c in faces], min_tri)


def sphere_octahedron(min_tri):
    """Palabos constructSphere (not in tree; restated from memory)"""
    va, vb, vc = np.array([1., 0, 0]), np.array([0., 1, 0]), np.array([-1., 0, 0])
    vd, ve, vf = np.array([0., -1, 0]), np.array([0., 0, 1]), np.array([0., 0, -1])
    faces = [(ve, va, vb), (ve, vb, vc), (ve, vc, vd), (ve, vd, va),
             (vf, vb, va), (vf, vc, vb), (vf, vd, vc), (vf, va, vd)]
    return _refine([list(f) for f in faces], min_tri)


def euler_zxz(phi, theta, psi):
    """Palabos TriangleSet::rotate(phi, theta, psi): classical z-x-z Euler matrix (from memory)"""
    a = np.empty((3, 3))
    a[0, 0] = math.cos(psi) * math.cos(phi) - math.cos(theta) * math.sin(phi) * math.sin(psi)
    a[0, 1] = math.cos(psi) * math.sin(phi) + math.cos(theta) * math.cos(phi) * math.sin(psi)
    a[0, 2] = math.sin(psi) * math.sin(theta)
    a[1, 0] = -math.sin(psi) * math.cos(phi) - math.cos(theta) * math.sin(phi) * math.cos(psi)
    a[1, 1] = -math.sin(psi) * math.sin(phi) + math.cos(theta) * math.cos(phi) * math.cos(psi)
    a[1, 2] = math.cos(psi) * math.sin(theta)
    a[2, 0] = math.sin(theta) * math.sin(phi)
    a[2, 1] = -math.sin(theta) * math.cos(phi)
    a[2, 2] = math.cos(theta)
    return a


def _rotate(tris, phi, theta, psi):
    a = euler_zxz(phi, theta, psi)
    out = np.empty_like(tris)
    for i in range(3):
        out[..., i] = a[i, 0] * tris[..., 0] + a[i, 1] * tris[..., 1] + a[i, 2] * tris[..., 2]
    return out


def _to_rbc(p, R=1.0):
    """spherePointToRBCPoint, helper/meshGeneratingFunctions.hh:153-168"""
    x, y, z = p[..., 0], p[..., 1], p[..., 2]
    r2 = x * x + y * y
    sign = (0 < z).astype(np.float64) - (z < 0).astype(np.float64)
    r2 = np.where(1 - r2 < 0, 1.0, r2)
    C0, C2, C4 = 0.054322, 1.001279, -0.561381
    out = np.empty_like(p)
    out[..., 0] = x * R
    out[..., 1] = y * R
    out[..., 2] = sign * R * np.sqrt(1 - r2) * (C0 + C2 * r2 + C4 * r2 * r2)
    return out


def _to_ellipsoid(p, R, aspect):
    """spherePointToEllipsoidPoint, helper/meshGeneratingFunctions.hh:170-183"""
    x, y, z = p[..., 0], p[..., 1], p[..., 2]
    r2 = x * x + y * y
    sign = (0 < z).astype(np.float64) - (z < 0).astype(np.float64)
    r2 = np.where(1 - r2 < 0, 1.0, r2)
    out = np.empty_like(p)
    out[..., 0] = x * R
    out[..., 1] = y * R
    out[..., 2] = sign * aspect * R * np.sqrt(1 - r2)
    return out


def _index(tris):
    """TriangleSet -> (vertices, triangles): vertices numbered by first appearance while
    scanning the triangle list (DEFscaledMesh / TriangleBoundary3D, from memory)."""
    key = {}
    verts = []
    tri_idx = np.empty((tris.shape[0], 3), dtype=np.int32)
    for t in range(tris.shape[0]):
        for k in range(3):
            p = tris[t, k]
            kk = (round(p[0] * 1e9), round(p[1] * 1e9), round(p[2] * 1e9))
            if kk not in key:
                key[kk] = len(verts)
                verts.append(p.copy())
            tri_idx[t, k] = key[kk]
    return np.array(verts), tri_idx


def inflate(verts, tris, amount=None):
    """TriangularSurfaceMesh::inflate (helper/meshGeneratingFunctions.h:92), from memory: every
    vertex moves by `amount` along its vertex normal (normalised sum of adjacent unit normals)"""
    amount = INFLATE if amount is None else amount
    if amount == 0.0:
        return verts
    n, _ = tri_normals_areas(verts, tris)
    vn = np.zeros_like(verts)
    for t in range(tris.shape[0]):
        for k in range(3):
            vn[tris[t, k]] += n[t]
    vn = vn / np.sqrt((vn * vn).sum(1))[:, None]
    return verts + amount * vn


def rbc_from_sphere(radius_lu, min_tri):
    """constructMeshElement(shape = RBC_FROM_SPHERE = 1), helper/meshGeneratingFunctions.h:68-94 and
    constructRBCFromSphere, .hh:213-241, with eulerAngles = 0."""
    s = sphere_icosahedron(min_tri)
    s = _rotate(s, PI / 2.0, PI / 2.0, 0.0)
    s = _to_rbc(s)
    s = s * radius_lu
    s = _rotate(s, PI / 2.0, PI / 2.0, 0.0)
    v, t = _index(s)
    return inflate(v, t), t


def ellipsoid_from_sphere(radius_lu, aspect, min_tri):
    """constructMeshElement(shape = ELLIPSOID_FROM_SPHERE = 6), constructEllipsoidFromSphere .hh:244-271"""
    s = sphere_octahedron(min_tri)
    s = _rotate(s, PI / 2.0, PI / 2.0, 0.0)
    s = _to_ellipsoid(s, radius_lu, aspect)
    s = _rotate(s, PI / 2.0, PI / 2.0, 0.0)
    v, t = _index(s)
    return inflate(v, t), t


# ----------------------------------------------------------------------------- metrics
def tri_normals_areas(verts, tris):
    v0, v1, v2 = verts[tris[:, 0]], verts[tris[:, 1]], verts[tris[:, 2]]
    n = np.cross(v1 - v0, v2 - v0)
    nn = np.sqrt((n * n).sum(1))
    return n / nn[:, None], 0.5 * nn


def mesh_volume(verts, tris):
    v0, v1, v2 = verts[tris[:, 0]], verts[tris[:, 1]], verts[tris[:, 2]]
    return float(np.einsum('ij,ij->i', v0, np.cross(v1, v2)).sum() / 6.0)


# ----------------------------------------------------------------------------- constants
def _adjacent_triangles(tris):
    """directed edge (a,b) -> triangle that contains it in that cyclic order"""
    d = {}
    for t, (a, b, c) in enumerate(tris):
        d[(int(a), int(b))] = t
        d[(int(b), int(c))] = t
        d[(int(c), int(a))] = t
    return d


def common_cell_constants(verts, tris, inner_edges=()):
    """CommonCellConstants::CommonCellConstantsConstructor, mechanics/commonCellConstants.cpp:70-409.
    Returns a dict of numpy arrays named after the reference members."""
    V, T = verts.shape[0], tris.shape[0]
    edges = []
    for a, b, c in tris:                                   # :81-93
        if a < b: edges.append((a, b))
        if b < c: edges.append((b, c))
        if c < a: edges.append((c, a))
    edges = np.array(edges, dtype=np.int32)
    E = edges.shape[0]
    de = _adjacent_triangles(tris)
    elen = np.sqrt(((verts[edges[:, 1]] - verts[edges[:, 0]]) ** 2).sum(1))   # :96-99
    normals, areas = tri_normals_areas(verts, tris)
    # getAdjacentTriangleIds(e0, e1): Palabos order unknown here (Appendix E.9).  Chosen so that the
    # PLT dihedral force of pltSimpleModel.cpp:156-182 is restoring: first = triangle holding the
    # directed edge e1->e0, second = the one holding e0->e1.
    bend_tris = np.empty((E, 2), dtype=np.int32)
    outer = np.empty((E, 2), dtype=np.int32)
    angle_eq = np.empty(E)
    for e, (a, b) in enumerate(edges):
        a, b = int(a), int(b)
        t0, t1 = de[(b, a)], de[(a, b)]
        bend_tris[e] = (t0, t1)
        ev = verts[b] - verts[a]
        uv = ev / np.sqrt(np.dot(ev, ev))
        cr = np.cross(normals[t0], normals[t1])
        angle_eq[e] = math.atan2(np.dot(cr, uv), np.dot(normals[t0], normals[t1]))   # :105-140
        for k, t in enumerate((t0, t1)):                   # :173-189
            for i in range(3):
                if tris[t][i] != a and tris[t][i] != b:
                    outer[e, k] = tris[t][i]
    inner = np.array(inner_edges, dtype=np.int32).reshape(-1, 2)
    inner_len = (np.sqrt(((verts[inner[:, 1]] - verts[inner[:, 0]]) ** 2).sum(1))
                 if inner.shape[0] else np.zeros(0))
    # vertex neighbours in order of first appearance in edge_list (:213-229) ...
    vv = -np.ones((V, 6), dtype=np.int32)
    nvv = np.zeros(V, dtype=np.int32)
    for a, b in edges:
        vv[a, nvv[a]] = b; nvv[a] += 1
        vv[b, nvv[b]] = a; nvv[b] += 1
    # ... re-ordered into a ring (:241-280): next = third vertex of the triangle in which
    # (vertex, current) appear consecutively in that cyclic order
    for v in range(V):
        cur = int(vv[v, 0])
        for n in range(1, nvv[v]):
            t = de[(v, cur)]
            nxt = [int(w) for w in tris[t] if w != v and w != cur][0]
            cur = nxt
            vv[v, n] = cur
    # surface patch centre deviation (:283-314)
    patch = np.empty(V)
    for i in range(V):
        n = nvv[i]
        ring = verts[vv[i, :n]]
        s = np.zeros(3)
        for j in range(n):
            s = s + ring[j]
        mid = s / n
        dev = mid - verts[i]
        pn = np.zeros(3)
        for j in range(n):
            tn = np.cross(ring[j] - verts[i], ring[(j + 1) % n] - verts[i])
            tn = tn / np.sqrt(np.dot(tn, tn))
            pn = pn + tn
        pn = pn / np.sqrt(np.dot(pn, pn))
        patch[i] = np.dot(pn, dev)
    return dict(
        triangle_list=tris.astype(np.int32), edge_list=edges, edge_length_eq_list=elen,
        edge_angle_eq_list=angle_eq, surface_patch_center_dist_eq_list=patch,
        edge_bending_triangles_list=bend_tris, edge_bending_triangles_outer_points=outer,
        triangle_area_eq_list=areas, vertex_vertexes=vv, vertex_n_vertexes=nvv,
        volume_eq=mesh_volume(verts, tris), area_mean_eq=float(areas.sum() / T),
        edge_mean_eq=float(elen.sum() / E), angle_mean_eq=float(angle_eq.sum() / E),
        inner_edge_list=inner, inner_edge_length_eq_list=inner_len)


# ----------------------------------------------------------------------------- parameters
class Parameters:
    """hemo::Parameters::lbm_base_parameters, mechanics/constantConversion.cpp:36-59"""

    def __init__(self, dx, dt, nu_p=1.1e-6, rho_p=1025.0, kBT_p=4.100531391e-21):
        self.dx, self.nu_p, self.rho_p, self.kBT_p = dx, nu_p, rho_p, kBT_p
        if dt < 0.0:
            self.tau = 1.0
            self.nu_lbm = 1.0 / 3.0 * (self.tau - 0.5)
            self.dt = self.nu_lbm / nu_p * (dx * dx)
        else:
            self.dt = dt
            self.nu_lbm = nu_p * dt / (dx * dx)
            self.tau = 3.0 * self.nu_lbm + 0.5
        self.dm = rho_p * (dx * dx * dx)
        self.df = self.dm * dx / (self.dt * self.dt)
        self.f_limit = 50.0 / 1.0e12 / self.df
        self.kBT_lbm = kBT_p / (self.df * dx)

    def stiffness(self, kLink, kBend, kVolume, kArea, eta_m, n_tri):
        """CellMechanics::calculate_k*, mechanics/cellMechanics.h:50-78"""
        plc = 7.5e-9 / self.dx
        eq = 5e-7 / self.dx
        scale = 1280.0 / n_tri
        return dict(k_link=kLink * self.kBT_lbm / plc, k_bend=kBend * self.kBT_lbm / eq,
                    k_volume=kVolume * scale * self.kBT_lbm / eq,
                    k_area=kArea * scale * self.kBT_lbm / eq,
                    eta_m=eta_m * self.dx / self.dt / self.df)


RBC_MATERIAL = dict(kBend=80.0, kVolume=20.0, kArea=5.0, kLink=15.0, eta_m=0.0,
                    minNumTriangles=600, radius=3.91e-6)       # examples/*/RBC.xml (all identical)
PLT_MATERIAL = dict(kBend=250.0, kVolume=100.0, kArea=8.0, kLink=25.0, eta_m=0.0,
                    minNumTriangles=66, radius=1.25e-6, aspectRatio=0.434782608696)  # examples/pipeflow/PLT.xml
PLT_INNER_EDGES = [(60, 65), (62, 64), (37, 42), (54, 56), (34, 40), (25, 46), (50, 59), (29, 47),
                   (61, 63), (26, 45), (33, 43), (27, 35), (32, 39), (49, 51), (0, 4), (48, 52),
                   (6, 10), (53, 55), (19, 21), (57, 58), (15, 13)]


# ----------------------------------------------------------------------------- .pos reader
def rotation_xyz(alpha, beta, gamma):
    """rotateTriangularMeshXYZ, io/readPositionsBloodCells.cpp:39-98: a = Rz * (Ry * Rx)"""
    a = np.array([[1.0, 0.0, 0.0], [0.0, math.cos(alpha), math.sin(alpha)], [0.0, -math.sin(alpha), math.cos(alpha)]])
    b = np.array([[math.cos(beta), 0.0, -math.sin(beta)], [0.0, 1.0, 0.0], [math.sin(beta), 0.0, math.cos(beta)]])
    c = np.zeros((3, 3))
    for i in range(3):
        for j in range(3):
            for k in range(3):
                c[i, j] += a[k, j] * b[i, k]
    b = np.array([[math.cos(gamma), math.sin(gamma), 0.0], [-math.sin(gamma), math.cos(gamma), 0.0], [0.0, 0.0, 1.0]])
    a = np.zeros((3, 3))
    for i in range(3):
        for j in range(3):
            for k in range(3):
                a[i, j] += c[k, j] * b[i, k]
    return a


def read_pos(path):
    with open(path) as fh:
        tok = fh.read().split()
    n = int(tok[0])
    return np.array(tok[1:1 + 6 * n], dtype=np.float64).reshape(n, 6)


def place_cells(verts, pos_rows, dx, dims, flags=None, min_dist_um=0.0, cell_id0=0):
    """ReadPositionsBloodCellField3D::processGenericBlocks + positionCellInParticleField +
    syncEnvelopes/deleteIncompleteCells of loadParticles (io/readPositionsBloodCells.cpp:120-170,
    186-361; core/hemoCell.cpp:191-197) collapsed onto the single global lattice: a cell survives
    iff every vertex lies in (-0.5, n-0.5] on every axis, its nearest node is not a boundary node
    and no boundary node lies within the deny cube.  Returns (positions [n,V,3], cell ids)."""
    nx, ny, nz = dims
    lo = 0.5 * (verts.min(0) + verts.max(0))
    mesh = verts - lo                                        # centred on its bbox (:317-318)
    pos_ratio = 1e-6 / dx
    deny = int((min_dist_um * 1e-6) / dx)
    out, ids = [], []
    for c, row in enumerate(pos_rows):
        ang = row[3:6] * (PI / 180.0) * -1.0                 # :228-229
        m = mesh.copy()
        ctr = 0.5 * (m.min(0) + m.max(0))                    # meshRotation :100-107
        m = m - ctr
        a = rotation_xyz(ang[0], ang[1], ang[2])
        r = np.empty_like(m)
        for i in range(3):
            r[:, i] = a[i, 0] * m[:, 0] + a[i, 1] * m[:, 1] + a[i, 2] * m[:, 2]
        m = r + ctr
        p = row[0:3] * pos_ratio + m                         # :129, :349
        ok = np.all((p > -0.5) & (p <= np.array([nx, ny, nz]) - 0.5))
        if ok and flags is not None:
            q = (p + 0.5).astype(np.int64)
            for dxx in range(-deny, deny + 1):
                for dyy in range(-deny, deny + 1):
                    for dzz in range(-deny, deny + 1):
                        qq = q + np.array([dxx, dyy, dzz])
                        inside = np.all((qq >= 0) & (qq < np.array([nx, ny, nz])), axis=1)
                        qq = qq[inside]
                        if np.any(flags[qq[:, 2] + nz * (qq[:, 1] + ny * qq[:, 0])] != 0):
                            ok = False
        if ok:
            out.append(p)
            ids.append(cell_id0 + c)
    if not out:
        return np.zeros((0, verts.shape[0], 3)), np.zeros(0, dtype=np.int64)
    return np.array(out), np.array(ids, dtype=np.int64)
