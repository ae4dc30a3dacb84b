"""CPU ORACLE package -- test infrastructure, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this.  `hemo_oracle.c` restates the per-operator arithmetic of the reference in plain C;
`mesh.py` restates the set-up side in numpy; `OracleSim` below restates the operator order and
cadences of HemoCell::iterate() (core/hemoCell.cpp:299-376).  Parity status: "parity unpinned"
per operator (the reference cannot be built here, see hemo_oracle.h); pinned only by the
known-answer checks in tests/test_oracle_known_answers.py.
"""
import ctypes as C
import os
import subprocess
import numpy as np

from . import mesh  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = C.POINTER(C.c_double)
c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_u8p = C.POINTER(C.c_uint8)


class OraDomain(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
                ("periodic", C.c_int32 * 3), ("omega", C.c_double),
                ("bc_vel", (C.c_double * 3) * 6)]


class OraCellType(C.Structure):
    _fields_ = [("model", C.c_int32), ("n_vertices", C.c_int32), ("n_triangles", C.c_int32),
                ("n_edges", C.c_int32), ("n_inner_edges", C.c_int32),
                ("triangles", c_i32p), ("edges", c_i32p), ("inner_edges", c_i32p),
                ("vertex_vertexes", c_i32p), ("vertex_n_vertexes", c_i32p),
                ("edge_bending_triangles", c_i32p), ("edge_bending_outer_points", c_i32p),
                ("edge_length_eq", c_dp), ("edge_angle_eq", c_dp), ("triangle_area_eq", c_dp),
                ("patch_dist_eq", c_dp), ("inner_edge_length_eq", c_dp),
                ("volume_eq", C.c_double), ("area_mean_eq", C.c_double), ("edge_mean_eq", C.c_double),
                ("k_volume", C.c_double), ("k_area", C.c_double), ("k_link", C.c_double),
                ("k_bend", C.c_double), ("eta_m", C.c_double)]


def build(force=False):
    so = os.path.join(_HERE, "libhemo_oracle.so")
    src = os.path.join(_HERE, "hemo_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.ora_ibm_kernel.restype = C.c_int
        _LIB.ora_advance.restype = C.c_int64
    return _LIB


def set_parallel(on):
    """OpenMP on the oracle loops (bench.py cpu_baseline only; tests keep the serial order)"""
    lib().ora_set_parallel(C.c_int(int(on)))


def _p(a, t=c_dp):
    return a.ctypes.data_as(t)


def make_domain(nx, ny, nz, periodic, tau, bc_vel=None):
    d = OraDomain()
    d.nx, d.ny, d.nz = nx, ny, nz
    for k in range(3):
        d.periodic[k] = int(bool(periodic[k]))
    d.omega = 1.0 / tau
    if bc_vel is not None:
        for o in range(6):
            for k in range(3):
                d.bc_vel[o][k] = float(bc_vel[o][k])
    return d


class CellType:
    """Topology + stiffness of one cell type, backed by numpy arrays kept alive here."""

    def __init__(self, model, verts, cc, k):
        self.model = model            # 0 RbcHighOrderModel, 1 PltSimpleModel
        self.verts = np.ascontiguousarray(verts, dtype=np.float64)
        self.cc = cc
        self.k = dict(k)
        a = self._arrs = dict(
            triangles=np.ascontiguousarray(cc['triangle_list'], dtype=np.int32),
            edges=np.ascontiguousarray(cc['edge_list'], dtype=np.int32),
            inner=np.ascontiguousarray(cc['inner_edge_list'], dtype=np.int32).reshape(-1, 2),
            vv=np.ascontiguousarray(cc['vertex_vertexes'], dtype=np.int32),
            nvv=np.ascontiguousarray(cc['vertex_n_vertexes'], dtype=np.int32),
            bt=np.ascontiguousarray(cc['edge_bending_triangles_list'], dtype=np.int32),
            bo=np.ascontiguousarray(cc['edge_bending_triangles_outer_points'], dtype=np.int32),
            el=np.ascontiguousarray(cc['edge_length_eq_list'], dtype=np.float64),
            ea=np.ascontiguousarray(cc['edge_angle_eq_list'], dtype=np.float64),
            ta=np.ascontiguousarray(cc['triangle_area_eq_list'], dtype=np.float64),
            pd=np.ascontiguousarray(cc['surface_patch_center_dist_eq_list'], dtype=np.float64),
            il=np.ascontiguousarray(cc['inner_edge_length_eq_list'], dtype=np.float64))
        t = self.c = OraCellType()
        t.model = model
        t.n_vertices = self.verts.shape[0]
        t.n_triangles = a['triangles'].shape[0]
        t.n_edges = a['edges'].shape[0]
        t.n_inner_edges = a['inner'].shape[0]
        t.triangles, t.edges, t.inner_edges = _p(a['triangles'], c_i32p), _p(a['edges'], c_i32p), _p(a['inner'], c_i32p)
        t.vertex_vertexes, t.vertex_n_vertexes = _p(a['vv'], c_i32p), _p(a['nvv'], c_i32p)
        t.edge_bending_triangles, t.edge_bending_outer_points = _p(a['bt'], c_i32p), _p(a['bo'], c_i32p)
        t.edge_length_eq, t.edge_angle_eq, t.triangle_area_eq = _p(a['el']), _p(a['ea']), _p(a['ta'])
        t.patch_dist_eq, t.inner_edge_length_eq = _p(a['pd']), _p(a['il'])
        t.volume_eq, t.area_mean_eq, t.edge_mean_eq = cc['volume_eq'], cc['area_mean_eq'], cc['edge_mean_eq']
        t.k_volume, t.k_area, t.k_link, t.k_bend, t.eta_m = (k['k_volume'], k['k_area'], k['k_link'],
                                                              k['k_bend'], k['eta_m'])

    @property
    def V(self):
        return self.verts.shape[0]


def rbc_celltype(par, material=None):
    m = dict(mesh.RBC_MATERIAL if material is None else material)
    v, t = mesh.rbc_from_sphere(m['radius'] / par.dx, m['minNumTriangles'])
    cc = mesh.common_cell_constants(v, t)
    k = par.stiffness(m['kLink'], m['kBend'], m['kVolume'], m['kArea'], m['eta_m'], t.shape[0])
    return CellType(0, v, cc, k)


def plt_celltype(par, material=None):
    m = dict(mesh.PLT_MATERIAL if material is None else material)
    v, t = mesh.ellipsoid_from_sphere(m['radius'] / par.dx, m['aspectRatio'], m['minNumTriangles'])
    cc = mesh.common_cell_constants(v, t, mesh.PLT_INNER_EDGES)
    k = par.stiffness(m['kLink'], m['kBend'], m['kVolume'], m['kArea'], m['eta_m'], t.shape[0])
    return CellType(1, v, cc, k)


# ------------------------------------------------------------------ thin operator wrappers
def init_equilibrium(dom, rho=1.0, u=(0.0, 0.0, 0.0)):
    N = dom.nx * dom.ny * dom.nz
    pop = np.empty(19 * N)
    lib().ora_init_equilibrium(C.byref(dom), C.c_double(rho), (C.c_double * 3)(*u), _p(pop))
    return pop


def collide_and_stream(dom, flags, pop, force, scratch=None, bc_node=None):
    """bc_node: [4*N] (u_x, u_y, u_z, rho) planes for Zou-He velocity / pressure nodes (flags 8..19)"""
    if scratch is None:
        scratch = np.empty_like(pop)
    lib().ora_collide_and_stream_io(C.byref(dom), _p(flags, c_u8p), _p(pop), _p(force), _p(scratch),
                                    None if bc_node is None else _p(bc_node))


def moments(dom, flags, pop, force, bc_node=None):
    N = dom.nx * dom.ny * dom.nz
    rho, vel = np.empty(N), np.empty(3 * N)
    lib().ora_moments_io(C.byref(dom), _p(flags, c_u8p), _p(pop), _p(force), _p(rho), _p(vel),
                         None if bc_node is None else _p(bc_node))
    return rho, vel


def spread(dom, flags, pos, pforce, frep, f_limit, node_force):
    lib().ora_spread(C.byref(dom), _p(flags, c_u8p), C.c_int64(pos.shape[0]), _p(pos), _p(pforce),
                     _p(frep), C.c_double(f_limit), _p(node_force))


def interpolate(dom, flags, pos, pop, node_force):
    vel = np.empty_like(pos)
    lib().ora_interpolate(C.byref(dom), _p(flags, c_u8p), C.c_int64(pos.shape[0]), _p(pos), _p(pop),
                          _p(node_force), _p(vel))
    return vel


def advance(dom, flags, pos, vel):
    hit = np.zeros(pos.shape[0], dtype=np.uint8)
    n = lib().ora_advance(C.byref(dom), _p(flags, c_u8p), C.c_int64(pos.shape[0]), _p(pos), _p(vel),
                          _p(hit, c_u8p))
    return n, hit


def mechanics(ct, pos, vel, force, components=False):
    """pos/vel/force: [ncells*V, 3]; force is accumulated into.  Returns the 6 component arrays
    (area, volume, bending, link, visc, inner) when components=True."""
    ncells = pos.shape[0] // ct.V
    comp = None
    arr = None
    if components:
        comp = [np.zeros_like(pos) for _ in range(6)]
        arr = (c_dp * 6)(*[_p(c) for c in comp])
    lib().ora_mechanics(C.byref(ct.c), C.c_int64(ncells), _p(pos), _p(vel), _p(force), arr)
    return comp


def repulsion(dom, pos, cell_of, k, cutoff):
    frep = np.empty_like(pos)
    cell_of = np.ascontiguousarray(cell_of, dtype=np.int64)
    lib().ora_repulsion(C.byref(dom), C.c_int64(pos.shape[0]), _p(pos), _p(cell_of, c_i64p),
                        C.c_double(k), C.c_double(cutoff), _p(frep))
    return frep


def wall_repulsion(dom, flags, pos, k, cutoff, frep):
    lib().ora_wall_repulsion(C.byref(dom), _p(flags, c_u8p), C.c_int64(pos.shape[0]), _p(pos),
                             C.c_double(k), C.c_double(cutoff), _p(frep))


class OracleSim:
    """Restatement of HemoCell::iterate() (core/hemoCell.cpp:299-376) on one global lattice.

    Particles of all cell types live in one [np,3] array, cells contiguous and in vertexId order,
    types in the order given.  Deleted cells (a vertex advanced onto a boundary node,
    hemoCellParticleField.cpp:579-584 + deleteIncompleteCells) are removed from the arrays.
    """

    def __init__(self, dom, flags, f_limit, body_force=(0.0, 0.0, 0.0)):
        self.dom, self.flags = dom, np.ascontiguousarray(flags, dtype=np.uint8)
        self.N = dom.nx * dom.ny * dom.nz
        self.f_limit = f_limit
        self.body_force = np.array(body_force, dtype=np.float64)
        self.pop = init_equilibrium(dom)
        self.force = np.empty(3 * self.N)
        self._reset_force()
        self.scratch = np.empty(19 * self.N)
        self.types, self.timescale = [], []
        self.pos = np.zeros((0, 3)); self.vel = np.zeros((0, 3))
        self.pforce = np.zeros((0, 3)); self.frep = np.zeros((0, 3))
        self.ctype = np.zeros(0, dtype=np.int32)      # per cell
        self.cell_id = np.zeros(0, dtype=np.int64)    # per cell
        self.iter = 0
        self.vel_timescale = 1
        self.rep_enabled = False; self.rep_timescale = 1; self.rep_k = 0.0; self.rep_cutoff = 0.0
        self.wall_enabled = False; self.wall_timescale = 1; self.wall_k = 0.0; self.wall_cutoff = 0.0
        self.bc_node = None     # [4*N] (u_x, u_y, u_z, rho) of the Zou-He velocity / pressure nodes (flags 8..19)

    def set_bc_nodes(self, node_idx, val):
        """setBoundaryVelocity / setBoundaryDensity on single Zou-He nodes (helper/preInlet.cpp:380,
        pipeflow_with_preinlet.cpp:132); val [n, 4]"""
        if self.bc_node is None:
            self.bc_node = np.zeros(4 * self.N); self.bc_node[3 * self.N:] = 1.0
        b = self.bc_node.reshape(4, self.N)
        b[:, np.asarray(node_idx, dtype=np.int64)] = np.asarray(val, dtype=np.float64).reshape(-1, 4).T

    def node_velocity(self, node_idx):
        """Cell::computeVelocity of the listed nodes from the current populations and node force"""
        _, vel = moments(self.dom, self.flags, self.pop, self.force, self.bc_node)
        return np.ascontiguousarray(vel.reshape(3, self.N)[:, np.asarray(node_idx, dtype=np.int64)].T)

    def insert_cells(self, t, pos, vel, pforce, frep, cell_ids):
        """cells that arrive later (pre-inlet hand-over): appended to the block of their type"""
        off = self._offsets()
        c_at = int(np.searchsorted(self.ctype, t, side='right'))
        p_at = int(off[c_at])
        ins = lambda a, b: np.ascontiguousarray(np.concatenate([a[:p_at], np.asarray(b).reshape(-1, 3), a[p_at:]]))
        self.pos, self.vel = ins(self.pos, pos), ins(self.vel, vel)
        self.pforce, self.frep = ins(self.pforce, pforce), ins(self.frep, frep)
        n = len(cell_ids)
        self.ctype = np.concatenate([self.ctype[:c_at], np.full(n, t, dtype=np.int32), self.ctype[c_at:]])
        self.cell_id = np.concatenate([self.cell_id[:c_at], np.asarray(cell_ids, dtype=np.int64), self.cell_id[c_at:]])

    def _reset_force(self):
        for k in range(3):
            self.force[k * self.N:(k + 1) * self.N] = self.body_force[k]

    def add_celltype(self, ct, timescale=1):
        self.types.append(ct); self.timescale.append(timescale)
        return len(self.types) - 1

    def add_cells(self, t, positions, cell_ids):
        """positions [n, V, 3]; all cells of type t must be added in one call, types in order"""
        n = positions.shape[0]
        self.pos = np.ascontiguousarray(np.concatenate([self.pos, positions.reshape(-1, 3)]))
        z = np.zeros((n * self.types[t].V, 3))
        self.vel = np.concatenate([self.vel, z]); self.pforce = np.concatenate([self.pforce, z])
        self.frep = np.concatenate([self.frep, z])
        self.ctype = np.concatenate([self.ctype, np.full(n, t, dtype=np.int32)])
        self.cell_id = np.concatenate([self.cell_id, np.asarray(cell_ids, dtype=np.int64)])

    def _offsets(self):
        sizes = np.array([self.types[t].V for t in self.ctype], dtype=np.int64)
        return np.concatenate([[0], np.cumsum(sizes)])

    def cell_of_particle(self):
        off = self._offsets()
        return np.repeat(np.arange(len(self.ctype), dtype=np.int64), np.diff(off))

    def apply_mechanics(self, forced=False):
        off = self._offsets()
        for t, ct in enumerate(self.types):          # hemoCellParticleField.cpp:654-670
            if not (forced or self.iter % self.timescale[t] == 0):
                continue
            sel = np.where(self.ctype == t)[0]
            if sel.size == 0:
                continue
            a, b = off[sel[0]], off[sel[-1] + 1]      # cells of one type are contiguous
            self.pforce[a:b] = 0.0
            pos = np.ascontiguousarray(self.pos[a:b]); vel = np.ascontiguousarray(self.vel[a:b])
            f = np.zeros_like(pos)
            mechanics(ct, pos, vel, f)
            self.pforce[a:b] = f

    def iterate(self):
        d, fl = self.dom, self.flags
        if self.rep_enabled and self.iter % self.rep_timescale == 0:
            self.frep = repulsion(d, self.pos, self.cell_of_particle(), self.rep_k, self.rep_cutoff)
        if self.wall_enabled and self.iter % self.wall_timescale == 0:
            wall_repulsion(d, fl, self.pos, self.wall_k, self.wall_cutoff, self.frep)
        spread(d, fl, self.pos, self.pforce, self.frep, self.f_limit, self.force)
        collide_and_stream(d, fl, self.pop, self.force, self.scratch, self.bc_node)
        if self.iter % self.vel_timescale == 0:
            self.vel = interpolate(d, fl, self.pos, self.pop, self.force)
        _, hit = advance(d, fl, self.pos, self.vel)
        if hit.any():
            dead = np.unique(self.cell_of_particle()[hit.astype(bool)])
            keep_c = np.ones(len(self.ctype), dtype=bool); keep_c[dead] = False
            keep_p = keep_c[self.cell_of_particle()]
            self.pos = np.ascontiguousarray(self.pos[keep_p]); self.vel = np.ascontiguousarray(self.vel[keep_p])
            self.pforce = np.ascontiguousarray(self.pforce[keep_p]); self.frep = np.ascontiguousarray(self.frep[keep_p])
            self.ctype = self.ctype[keep_c]; self.cell_id = self.cell_id[keep_c]
        self.apply_mechanics()
        self._reset_force()                            # setExternalVector(..., 0) + case-file body force
        self.iter += 1


class PreInletCoupling:
    """Restatement of PreInlet::applyPreInlet (helper/preInlet.cpp:255-397) between two OracleSims: `pre`, the
    periodic force-driven pre-inlet, and `main`, whose Zou-He velocity nodes main_idx take the velocity of the
    pre-inlet nodes pre_idx after every step.  Cells are handed over whole (see include/hemocell_gpu.h,
    hcg_preinlet_apply_cells): the periodic image k of a pre-inlet cell is copied the first time it lies wholly
    inside [slab_lo, slab_hi] along `axis` in main coordinates (= position + shift + k*period), id + z(k)*id_stride with z = 0, 1, 2, 3, 4 for k = 0, -1, 1, -2, 2 (unique, non-negative)."""

    def __init__(self, pre, main, pre_idx, main_idx, axis, period, shift, slab_lo, slab_hi, id_stride):
        self.pre, self.main = pre, main
        self.pre_idx = np.asarray(pre_idx, dtype=np.int64); self.main_idx = np.asarray(main_idx, dtype=np.int64)
        self.axis, self.period, self.shift = axis, float(period), np.asarray(shift, dtype=np.float64)
        self.slab_lo, self.slab_hi, self.id_stride = float(slab_lo), float(slab_hi), int(id_stride)
        self.last_lap = {}
        if main.bc_node is None:
            main.set_bc_nodes(np.zeros(0, dtype=np.int64), np.zeros((0, 4)))

    def apply_velocity(self):
        u = self.pre.node_velocity(self.pre_idx)
        b = self.main.bc_node.reshape(4, self.main.N)
        b[0:3, self.main_idx] = u.T

    def apply_cells(self):
        pre, main, ax = self.pre, self.main, self.axis
        off = pre._offsets()
        added = 0
        for c in range(len(pre.ctype)):
            x = pre.pos[off[c]:off[c + 1]]
            lo, hi = x[:, ax].min() + self.shift[ax], x[:, ax].max() + self.shift[ax]
            k = np.ceil((self.slab_lo - lo) / self.period)
            if hi + k * self.period > self.slab_hi:
                continue
            k = int(k)
            cid = int(pre.cell_id[c])
            if self.last_lap.get(cid) == k:
                continue
            self.last_lap[cid] = k
            sh = self.shift.copy(); sh[ax] += k * self.period
            sl = slice(off[c], off[c + 1])
            main.insert_cells(int(pre.ctype[c]), pre.pos[sl] + sh, pre.vel[sl], pre.pforce[sl], pre.frep[sl],
                              [cid + (2 * abs(k) - (1 if k < 0 else 0)) * self.id_stride])
            added += 1
        return added
