"""ctypes binding of libhemocell_gpu.so (include/hemocell_gpu.h) -- the harness side of the C ABI.

This is what the pytest / bench harness uses to drive the product; it contains no compute and
no fallback: if the CUDA library is missing, import fails loudly.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhemocell_gpu.so")

c_dp = C.POINTER(C.c_double)
c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_u8p = C.POINTER(C.c_uint8)

FLUID, BOUNCEBACK, VEL_XN, VEL_XP, VEL_YN, VEL_YP, VEL_ZN, VEL_ZP = range(8)
ZH_VEL_XN, ZH_VEL_XP, ZH_VEL_YN, ZH_VEL_YP, ZH_VEL_ZN, ZH_VEL_ZP = range(8, 14)        # Zou-He velocity nodes
ZH_PRES_XN, ZH_PRES_XP, ZH_PRES_YN, ZH_PRES_YP, ZH_PRES_ZN, ZH_PRES_ZP = range(14, 20)  # Zou-He pressure nodes
MODEL_RBC, MODEL_PLT = 0, 1
LAT_POP, LAT_FORCE, LAT_VELOCITY, LAT_DENSITY, LAT_PINEQ = range(5)
P_POS, P_VEL, P_FORCE, P_FREP, P_F_AREA, P_F_VOLUME, P_F_BEND, P_F_LINK, P_F_VISC, P_F_INNER = range(10)


class HcgDomain(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("periodic", C.c_int32 * 3),
                ("tau", C.c_double), ("device", C.c_int32), ("rank", C.c_int32), ("n_ranks", C.c_int32)]


class HcgCellType(C.Structure):
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


class HcgTimer(C.Structure):
    _fields_ = [("name", C.c_char * 40), ("ms_total", C.c_double), ("calls", C.c_int64)]


# every symbol include/hemocell_gpu.h declares (tests check that the .so exports all of them)
SYMBOLS = """hcg_last_error hcg_version hcg_create hcg_device_count hcg_slab hcg_destroy hcg_comm_unique_id hcg_comm_init hcg_comm_init_local
hcg_lattice_set_flags hcg_lattice_set_bc_velocity hcg_lattice_init_equilibrium hcg_lattice_set_body_force hcg_lattice_set_body_force_field
hcg_lattice_upload hcg_lattice_download hcg_celltype_add hcg_cells_add hcg_cells_count hcg_cells_count_async hcg_cells_capacity
hcg_cells_upload hcg_cells_download hcg_cells_download_f32 hcg_cells_info hcg_cells_owned hcg_allreduce hcg_cells_add_force hcg_celltype_set_stiffness
hcg_set_force_limit hcg_set_timescales hcg_set_material_timescale hcg_set_repulsion hcg_set_wall_repulsion
hcg_set_spread_mode hcg_set_exchange hcg_set_transport hcg_exchange_stats hcg_set_iteration hcg_get_iteration hcg_iterate hcg_iterate_async hcg_fluid_warmup hcg_op_repulsion hcg_op_wall_repulsion
hcg_op_spread hcg_op_collide_stream hcg_op_interpolate hcg_op_sync hcg_op_advance hcg_op_mechanics
hcg_op_zero_force hcg_cells_bbox hcg_cells_volume_area hcg_cells_stretch hcg_fluid_velocity_stats hcg_timers_enable
hcg_timers hcg_timers_reset hcg_launch_count hcg_synchronize hcg_iterate_timed
hcg_lattice_set_bc_nodes hcg_lattice_node_velocity hcg_cells_reserve hcg_preinlet_map hcg_preinlet_apply_velocity
hcg_preinlet_apply_cells hcg_preinlet_laps hcg_set_moment_only""".split()

_lib = None


def load():
    """dlopen the CUDA library; raises if it was not built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        _lib.hcg_last_error.restype = C.c_char_p
        _lib.hcg_last_error.argtypes = [C.c_void_p]
        _lib.hcg_version.restype = C.c_char_p
    return _lib


class HcgError(RuntimeError):
    pass


def _p(a, t=c_dp):
    return a.ctypes.data_as(t)


class Context:
    """One hcg_ctx (one GPU).  Thin: every method is one C-ABI call."""

    def __init__(self, nx, ny, nz, periodic, tau, device=0, rank=0, n_ranks=1):
        self.L = load()
        d = HcgDomain()
        d.nx, d.ny, d.nz = nx, ny, nz
        for k in range(3):
            d.periodic[k] = int(bool(periodic[k]))
        d.tau, d.device, d.rank, d.n_ranks = tau, device, rank, n_ranks
        self.dom = d
        self.h = C.c_void_p()
        st = self.L.hcg_create(C.byref(d), C.byref(self.h))
        if st != 0:
            msg = self.L.hcg_last_error(self.h if self.h else None)
            raise HcgError(f"hcg_create failed ({st}): {msg.decode() if msg else ''}")
        x0, nxl = C.c_int32(), C.c_int32()
        self.L.hcg_slab(C.c_int32(nx), C.c_int32(rank), C.c_int32(n_ranks), C.byref(x0), C.byref(nxl))
        self.x0, self.nxl = x0.value, nxl.value
        self.Nl = self.nxl * ny * nz
        self._keep = []

    def _ck(self, st):
        if st != 0:
            raise HcgError(f"status {st}: {self.L.hcg_last_error(self.h).decode()}")

    def close(self):
        if self.h:
            self.L.hcg_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- multi-GPU
    @staticmethod
    def unique_id():
        buf = (C.c_uint8 * 128)()
        st = load().hcg_comm_unique_id(buf)
        if st != 0:
            raise HcgError("hcg_comm_unique_id failed")
        return bytes(buf)

    def comm_init(self, id128, local=False):
        """local=True: in-process communicator (all ranks are contexts of this process, possibly on one GPU)"""
        buf = (C.c_uint8 * 128).from_buffer_copy(id128)
        self._ck((self.L.hcg_comm_init_local if local else self.L.hcg_comm_init)(self.h, buf))

    # ---- lattice
    def set_flags(self, flags):
        f = np.ascontiguousarray(flags, dtype=np.uint8).reshape(-1)
        assert f.size == self.Nl
        self._ck(self.L.hcg_lattice_set_flags(self.h, _p(f, c_u8p)))

    def set_bc_velocity(self, orientation, u):
        self._ck(self.L.hcg_lattice_set_bc_velocity(self.h, C.c_int32(orientation), (C.c_double * 3)(*u)))

    def set_bc_nodes(self, node_idx, val):
        """per-node (u_x, u_y, u_z, rho) of Zou-He velocity / pressure nodes; node_idx = local node indices"""
        idx = np.ascontiguousarray(node_idx, dtype=np.int64).reshape(-1)
        v = np.ascontiguousarray(val, dtype=np.float64).reshape(-1)
        assert v.size == 4 * idx.size
        self._ck(self.L.hcg_lattice_set_bc_nodes(self.h, C.c_int64(idx.size), _p(idx, c_i64p), _p(v)))

    def node_velocity(self, node_idx):
        idx = np.ascontiguousarray(node_idx, dtype=np.int64).reshape(-1)
        out = np.empty((idx.size, 3))
        self._ck(self.L.hcg_lattice_node_velocity(self.h, C.c_int64(idx.size), _p(idx, c_i64p), _p(out)))
        return out

    # ---- pre-inlet coupling (this context = main domain)
    def preinlet_map(self, pre, pre_idx, main_idx):
        a = np.ascontiguousarray(pre_idx, dtype=np.int64).reshape(-1)
        b = np.ascontiguousarray(main_idx, dtype=np.int64).reshape(-1)
        assert a.size == b.size
        self._ck(self.L.hcg_preinlet_map(self.h, pre.h, C.c_int64(a.size), _p(a, c_i64p), _p(b, c_i64p)))

    def preinlet_apply_velocity(self):
        self._ck(self.L.hcg_preinlet_apply_velocity(self.h))

    def preinlet_apply_cells(self, axis, period, shift, slab_lo, slab_hi, id_stride):
        n = C.c_int64()
        self._ck(self.L.hcg_preinlet_apply_cells(self.h, C.c_int32(axis), C.c_double(period), (C.c_double * 3)(*shift),
                                                 C.c_double(slab_lo), C.c_double(slab_hi), C.c_int64(id_stride), C.byref(n)))
        return n.value

    def reserve_cells(self, ctype, spare):
        self._ck(self.L.hcg_cells_reserve(self.h, C.c_int32(ctype), C.c_int64(spare)))

    def init_equilibrium(self, rho=1.0, u=(0.0, 0.0, 0.0)):
        self._ck(self.L.hcg_lattice_init_equilibrium(self.h, C.c_double(rho), (C.c_double * 3)(*u)))

    def set_body_force(self, f):
        self._ck(self.L.hcg_lattice_set_body_force(self.h, (C.c_double * 3)(*f)))

    def set_body_force_field(self, f):
        a = np.ascontiguousarray(f, dtype=np.float64).reshape(-1)
        assert a.size == 3 * self.Nl
        self._ck(self.L.hcg_lattice_set_body_force_field(self.h, _p(a)))

    def lattice_upload(self, field, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64).reshape(-1)
        assert a.size == {LAT_POP: 19, LAT_FORCE: 3}[field] * self.Nl
        self._ck(self.L.hcg_lattice_upload(self.h, C.c_int32(field), _p(a)))

    def lattice_download(self, field):
        n = {LAT_POP: 19, LAT_FORCE: 3, LAT_VELOCITY: 3, LAT_DENSITY: 1, LAT_PINEQ: 6}[field] * self.Nl
        out = np.empty(n)
        self._ck(self.L.hcg_lattice_download(self.h, C.c_int32(field), _p(out)))
        return out

    # ---- cells
    def add_celltype(self, model, tables, k):
        """tables: dict with the CommonCellConstants member names; k: dict of k_* and eta_m"""
        a = dict(
            triangles=np.ascontiguousarray(tables['triangle_list'], dtype=np.int32),
            edges=np.ascontiguousarray(tables['edge_list'], dtype=np.int32),
            inner=np.ascontiguousarray(tables['inner_edge_list'], dtype=np.int32).reshape(-1, 2),
            vv=np.ascontiguousarray(tables['vertex_vertexes'], dtype=np.int32),
            nvv=np.ascontiguousarray(tables['vertex_n_vertexes'], dtype=np.int32),
            bt=np.ascontiguousarray(tables['edge_bending_triangles_list'], dtype=np.int32),
            bo=np.ascontiguousarray(tables['edge_bending_triangles_outer_points'], dtype=np.int32),
            el=np.ascontiguousarray(tables['edge_length_eq_list'], dtype=np.float64),
            ea=np.ascontiguousarray(tables['edge_angle_eq_list'], dtype=np.float64),
            ta=np.ascontiguousarray(tables['triangle_area_eq_list'], dtype=np.float64),
            pd=np.ascontiguousarray(tables['surface_patch_center_dist_eq_list'], dtype=np.float64),
            il=np.ascontiguousarray(tables['inner_edge_length_eq_list'], dtype=np.float64))
        t = HcgCellType()
        t.model = model
        t.n_vertices = a['nvv'].shape[0]
        t.n_triangles = a['triangles'].shape[0]
        t.n_edges = a['edges'].shape[0]
        t.n_inner_edges = a['inner'].shape[0]
        t.triangles, t.edges, t.inner_edges = _p(a['triangles'], c_i32p), _p(a['edges'], c_i32p), _p(a['inner'], c_i32p)
        t.vertex_vertexes, t.vertex_n_vertexes = _p(a['vv'], c_i32p), _p(a['nvv'], c_i32p)
        t.edge_bending_triangles, t.edge_bending_outer_points = _p(a['bt'], c_i32p), _p(a['bo'], c_i32p)
        t.edge_length_eq, t.edge_angle_eq, t.triangle_area_eq = _p(a['el']), _p(a['ea']), _p(a['ta'])
        t.patch_dist_eq, t.inner_edge_length_eq = _p(a['pd']), _p(a['il'])
        t.volume_eq, t.area_mean_eq, t.edge_mean_eq = tables['volume_eq'], tables['area_mean_eq'], tables['edge_mean_eq']
        t.k_volume, t.k_area, t.k_link, t.k_bend, t.eta_m = k['k_volume'], k['k_area'], k['k_link'], k['k_bend'], k['eta_m']
        out = C.c_int32(-1)
        self._ck(self.L.hcg_celltype_add(self.h, C.byref(t), C.byref(out)))
        return out.value

    def add_cells(self, ctype, positions, cell_ids):
        pos = np.ascontiguousarray(positions, dtype=np.float64)
        ids = np.ascontiguousarray(cell_ids, dtype=np.int64)
        self._ck(self.L.hcg_cells_add(self.h, C.c_int32(ctype), C.c_int64(ids.shape[0]), _p(ids, c_i64p), _p(pos)))

    def capacity(self):
        nc, npt = C.c_int64(), C.c_int64()
        self._ck(self.L.hcg_cells_capacity(self.h, C.byref(nc), C.byref(npt)))
        return nc.value, npt.value

    def count(self):
        nc, npt = C.c_int64(), C.c_int64()
        self._ck(self.L.hcg_cells_count(self.h, C.byref(nc), C.byref(npt)))
        return nc.value, npt.value

    def cells_upload(self, field, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64).reshape(-1)
        assert a.size == 3 * self.capacity()[1]
        self._ck(self.L.hcg_cells_upload(self.h, C.c_int32(field), _p(a)))

    def cells_download(self, field):
        out = np.empty((self.capacity()[1], 3))
        self._ck(self.L.hcg_cells_download(self.h, C.c_int32(field), _p(out)))
        return out

    def cells_info(self):
        nc = self.capacity()[0]
        ids, ct, alive = np.empty(nc, dtype=np.int64), np.empty(nc, dtype=np.int32), np.empty(nc, dtype=np.uint8)
        self._ck(self.L.hcg_cells_info(self.h, _p(ids, c_i64p), _p(ct, c_i32p), _p(alive, c_u8p)))
        return ids, ct, alive

    def add_force(self, index, f):
        idx = np.ascontiguousarray(index, dtype=np.int64)
        ff = np.ascontiguousarray(f, dtype=np.float64).reshape(-1)
        self._ck(self.L.hcg_cells_add_force(self.h, C.c_int64(idx.shape[0]), _p(idx, c_i64p), _p(ff)))

    # ---- knobs
    def set_force_limit(self, f):
        self._ck(self.L.hcg_set_force_limit(self.h, C.c_double(f)))

    def set_timescales(self, velocity=1, repulsion=1, wall=1):
        self._ck(self.L.hcg_set_timescales(self.h, C.c_int32(velocity), C.c_int32(repulsion), C.c_int32(wall)))

    def set_material_timescale(self, ctype, every):
        self._ck(self.L.hcg_set_material_timescale(self.h, C.c_int32(ctype), C.c_int32(every)))

    def set_repulsion(self, on, k, cutoff):
        self._ck(self.L.hcg_set_repulsion(self.h, C.c_int32(int(on)), C.c_double(k), C.c_double(cutoff)))

    def set_wall_repulsion(self, on, k, cutoff):
        self._ck(self.L.hcg_set_wall_repulsion(self.h, C.c_int32(int(on)), C.c_double(k), C.c_double(cutoff)))

    def set_moment_only(self, on):
        self._ck(self.L.hcg_set_moment_only(self.h, C.c_int32(int(on))))

    def set_spread_mode(self, mode=1, resort_every=20):
        self._ck(self.L.hcg_set_spread_mode(self.h, C.c_int32(mode), C.c_int32(resort_every)))

    def set_exchange(self, margin=4.0, sync_every=20, slack=0.3):
        self._ck(self.L.hcg_set_exchange(self.h, C.c_double(margin), C.c_int32(sync_every), C.c_double(slack)))

    def set_transport(self, transport):
        """1 = NVLink peer memory (default), 0 = NCCL send/recv; before comm_init"""
        self._ck(self.L.hcg_set_transport(self.h, C.c_int32(transport)))

    def exchange_stats(self):
        v = [C.c_int64() for _ in range(4)]
        self._ck(self.L.hcg_exchange_stats(self.h, *[C.byref(x) for x in v]))
        return dict(zip(["shared_left", "shared_right", "migrated_in", "migrated_out"], [x.value for x in v]))

    def set_iteration(self, it):
        self._ck(self.L.hcg_set_iteration(self.h, C.c_int64(it)))

    @property
    def iteration(self):
        it = C.c_int64()
        self._ck(self.L.hcg_get_iteration(self.h, C.byref(it)))
        return it.value

    # ---- run
    def iterate(self, n=1):
        self._ck(self.L.hcg_iterate(self.h, C.c_int64(n)))

    def iterate_async(self, n=1):
        self._ck(self.L.hcg_iterate_async(self.h, C.c_int64(n)))

    def count_async(self, out_addr):
        """out_addr: address of 2 int64 in page-locked host memory"""
        self._ck(self.L.hcg_cells_count_async(self.h, C.c_void_p(out_addr)))

    def iterate_timed(self, n):
        ms = C.c_double()
        self._ck(self.L.hcg_iterate_timed(self.h, C.c_int64(n), C.byref(ms)))
        return ms.value

    def fluid_warmup(self, n):
        self._ck(self.L.hcg_fluid_warmup(self.h, C.c_int64(n)))

    def op(self, name, *args):
        fn = getattr(self.L, "hcg_op_" + name)
        self._ck(fn(self.h, *[C.c_int32(int(a)) for a in args]))

    # ---- observables / timing
    def bbox(self):
        out = np.empty((self.capacity()[0], 6))
        self._ck(self.L.hcg_cells_bbox(self.h, _p(out)))
        return out

    def volume_area(self):
        nc = self.capacity()[0]
        v, a = np.empty(nc), np.empty(nc)
        self._ck(self.L.hcg_cells_volume_area(self.h, _p(v), _p(a)))
        return v, a

    def stretch(self):
        out = np.zeros(self.capacity()[0])
        self._ck(self.L.hcg_cells_stretch(self.h, _p(out)))
        return out

    def velocity_stats(self):
        a, b, m = C.c_double(), C.c_double(), C.c_double()
        self._ck(self.L.hcg_fluid_velocity_stats(self.h, C.byref(a), C.byref(b), C.byref(m)))
        return a.value, b.value, m.value

    def timers_enable(self, on=True):
        self._ck(self.L.hcg_timers_enable(self.h, C.c_int32(int(on))))

    def timers(self):
        n = C.c_int32(64)
        arr = (HcgTimer * 64)()
        self._ck(self.L.hcg_timers(self.h, arr, C.byref(n)))
        return {arr[k].name.decode(): (arr[k].ms_total, arr[k].calls) for k in range(min(n.value, 64))}

    def timers_reset(self):
        self._ck(self.L.hcg_timers_reset(self.h))

    def launch_count(self):
        n = C.c_int64()
        self._ck(self.L.hcg_launch_count(self.h, C.byref(n)))
        return n.value

    def synchronize(self):
        self._ck(self.L.hcg_synchronize(self.h))


# ------------------------------------------------------------------ host-side set-up (C++ in the same .so)
HOST_SYMBOLS = """hch_parameters hch_celltype_build hch_celltype_view hch_celltype_vertices hch_celltype_scalar
hch_celltype_free hch_read_pos hch_place_cells hch_slab_membership hch_slab_membership_at hch_last_error
hch_h5_create hch_h5_attribute hch_h5_dataset hch_h5_close hch_voxelize_stl hch_preinlet_select""".split()

RBC_MATERIAL = dict(kBend=80.0, kVolume=20.0, kArea=5.0, kLink=15.0, eta_m=0.0, minNumTriangles=600,
                    radius=3.91e-6, aspectRatio=0.3)
PLT_MATERIAL = dict(kBend=250.0, kVolume=100.0, kArea=8.0, kLink=25.0, eta_m=0.0, minNumTriangles=66,
                    radius=1.25e-6, aspectRatio=0.434782608696)
PLT_INNER_EDGES = [(60, 65), (62, 64), (37, 42), (54, 56), (34, 40), (25, 46), (50, 59), (29, 47), (61, 63),
                   (26, 45), (33, 43), (27, 35), (32, 39), (49, 51), (0, 4), (48, 52), (6, 10), (53, 55),
                   (19, 21), (57, 58), (15, 13)]
RBC_FROM_SPHERE, ELLIPSOID_FROM_SPHERE = 1, 6


def parameters(dx, dt, nu_p=1.1e-6, rho_p=1025.0, kBT_p=4.100531391e-21):
    """hemo::Parameters::lbm_base_parameters -> dict(tau, nu_lbm, dt, dm, df, f_limit, kBT_lbm)"""
    out = (C.c_double * 7)()
    fn = load().hch_parameters
    fn.restype = None
    fn(C.c_double(dx), C.c_double(dt), C.c_double(nu_p), C.c_double(rho_p), C.c_double(kBT_p), out)
    d = dict(zip(["tau", "nu_lbm", "dt", "dm", "df", "f_limit", "kBT_lbm"], list(out)))
    d.update(dx=dx, nu_p=nu_p, rho_p=rho_p, kBT_p=kBT_p)
    return d


class HostCellType:
    """hemo::CellTypeTables built by the product's C++ host code"""

    def __init__(self, model, construct_type, par, material, inner_edges=()):
        L = load()
        L.hch_celltype_build.restype = C.c_void_p
        L.hch_celltype_view.restype = C.POINTER(HcgCellType)
        L.hch_celltype_view.argtypes = [C.c_void_p]
        L.hch_celltype_vertices.argtypes = [C.c_void_p, c_dp]
        L.hch_celltype_scalar.restype = C.c_double
        L.hch_celltype_scalar.argtypes = [C.c_void_p, C.c_int32]
        L.hch_celltype_free.argtypes = [C.c_void_p]
        L.hch_last_error.restype = C.c_char_p
        ie = np.ascontiguousarray(np.array(inner_edges, dtype=np.int32).reshape(-1, 2))
        m = material
        self.h = L.hch_celltype_build(
            C.c_int32(model), C.c_int32(construct_type), C.c_double(par["dx"]), C.c_double(par["dt"]),
            C.c_double(par["nu_p"]), C.c_double(par["rho_p"]), C.c_double(par["kBT_p"]),
            C.c_double(m["kBend"]), C.c_double(m["kVolume"]), C.c_double(m["kArea"]), C.c_double(m["kLink"]),
            C.c_double(m["eta_m"]), C.c_double(m["radius"]), C.c_double(m["aspectRatio"]),
            C.c_int32(m["minNumTriangles"]), _p(ie, c_i32p), C.c_int32(ie.shape[0]))
        if not self.h:
            raise HcgError("hch_celltype_build: " + L.hch_last_error().decode())
        self.L = L
        self.model = model
        self.view = L.hch_celltype_view(self.h)
        self.V = self.view.contents.n_vertices
        self.verts = np.empty((self.V, 3))
        L.hch_celltype_vertices(self.h, _p(self.verts))

    def table(self, name, shape, dtype):
        ptr = getattr(self.view.contents, name)
        n = int(np.prod(shape))
        if n == 0:
            return np.zeros(shape, dtype=dtype)
        return np.ctypeslib.as_array(ptr, shape=(n,)).reshape(shape).astype(dtype, copy=True)

    def scalar(self, which):
        return self.L.hch_celltype_scalar(self.h, which)

    def add_to(self, ctx):
        out = C.c_int32(-1)
        ctx._ck(ctx.L.hcg_celltype_add(ctx.h, self.view, C.byref(out)))
        return out.value

    def place(self, rows, dx, dims, flags=None, min_dist_um=0.0, cell_id0=0):
        rows = np.ascontiguousarray(rows, dtype=np.float64).reshape(-1, 6)
        n = rows.shape[0]
        out = np.empty((n, self.V, 3))
        ids = np.empty(n, dtype=np.int64)
        fl = None
        if flags is not None:
            fl = np.ascontiguousarray(flags, dtype=np.uint8).reshape(-1)
        fn = self.L.hch_place_cells
        fn.restype = C.c_int64
        k = fn(C.c_void_p(self.h), _p(rows), C.c_int64(n), C.c_double(dx), C.c_int32(dims[0]), C.c_int32(dims[1]),
               C.c_int32(dims[2]), _p(fl, c_u8p) if fl is not None else None, C.c_double(min_dist_um),
               C.c_int64(cell_id0), _p(out), _p(ids, c_i64p))
        if k < 0:
            raise HcgError("hch_place_cells: " + self.L.hch_last_error().decode())
        return np.ascontiguousarray(out[:k]), ids[:k].copy()

    def __del__(self):
        try:
            if self.h:
                self.L.hch_celltype_free(self.h)
                self.h = None
        except Exception:
            pass


def slab_membership(xlo, xhi, nx, periodic_x, nxl, rank, n_ranks, margin):
    """hch_slab_membership -> (held, share_left, share_right) boolean arrays"""
    xlo = np.ascontiguousarray(xlo, dtype=np.float64); xhi = np.ascontiguousarray(xhi, dtype=np.float64)
    n = xlo.shape[0]
    out = [np.zeros(n, dtype=np.uint8) for _ in range(3)]
    fn = load().hch_slab_membership
    fn.restype = None
    fn(C.c_int64(n), _p(xlo), _p(xhi), C.c_int32(nx), C.c_int32(int(periodic_x)), C.c_int32(nxl), C.c_int32(rank),
       C.c_int32(n_ranks), C.c_double(margin), *[_p(o, c_u8p) for o in out])
    return [o.astype(bool) for o in out]


def preinlet_select(lo, hi, alive, last_lap, shift, period, slab_lo, slab_hi):
    """hch_preinlet_select -> (lap, take)"""
    lo = np.ascontiguousarray(lo, dtype=np.float64); hi = np.ascontiguousarray(hi, dtype=np.float64)
    alive = np.ascontiguousarray(alive, dtype=np.uint8)
    n = lo.shape[0]
    lap = np.zeros(n, dtype=np.int64); take = np.zeros(n, dtype=np.uint8)
    ll = None if last_lap is None else np.ascontiguousarray(last_lap, dtype=np.int64)
    fn = load().hch_preinlet_select
    fn.restype = None
    fn(C.c_int64(n), _p(lo), _p(hi), _p(alive, c_u8p), None if ll is None else _p(ll, c_i64p), C.c_double(shift), C.c_double(period),
       C.c_double(slab_lo), C.c_double(slab_hi), _p(lap, c_i64p), _p(take, c_u8p))
    return lap, take.astype(bool)


def read_pos(path):
    fn = load().hch_read_pos
    fn.restype = C.c_int64
    n = fn(path.encode(), None, C.c_int64(0))
    if n < 0:
        raise HcgError("hch_read_pos: cannot read " + path)
    rows = np.empty((n, 6))
    fn(path.encode(), _p(rows), C.c_int64(n))
    return rows


class H5Writer:
    """hch_h5_*: the product's HDF5 container writer (hemocell_b200/host/hemo_h5.cpp)"""
    TYPES = {np.dtype("float32"): 0, np.dtype("float64"): 1, np.dtype("int32"): 2, np.dtype("int64"): 3}

    def __init__(self, path, deflate_level=7):
        self.L = load()
        self.L.hch_h5_create.restype = C.c_void_p
        self.L.hch_last_error.restype = C.c_char_p
        self.h = self.L.hch_h5_create(str(path).encode(), C.c_int32(deflate_level))
        if not self.h:
            raise HcgError("hch_h5_create: " + self.L.hch_last_error().decode())

    def attribute(self, name, values):
        a = np.ascontiguousarray(values)
        rc = self.L.hch_h5_attribute(C.c_void_p(self.h), name.encode(), C.c_int32(self.TYPES[a.dtype]),
                                     a.ctypes.data_as(C.c_void_p), C.c_int64(a.size))
        if rc:
            raise HcgError("hch_h5_attribute: " + self.L.hch_last_error().decode())

    def dataset(self, name, array, chunk=None):
        a = np.ascontiguousarray(array)
        dims = (C.c_uint64 * a.ndim)(*a.shape)
        ch = (C.c_uint64 * a.ndim)(*chunk) if chunk is not None else None
        rc = self.L.hch_h5_dataset(C.c_void_p(self.h), name.encode(), C.c_int32(self.TYPES[a.dtype]), C.c_int32(a.ndim),
                                   dims, a.ctypes.data_as(C.c_void_p), ch)
        if rc:
            raise HcgError("hch_h5_dataset: " + self.L.hch_last_error().decode())

    def close(self):
        rc = self.L.hch_h5_close(C.c_void_p(self.h))
        self.h = None
        if rc:
            raise HcgError("hch_h5_close: " + self.L.hch_last_error().decode())


def voxelize_stl(path, ref_dir_n, ref_dir):
    """hch_voxelize_stl -> (flags [nx, ny, nz] uint8, dx in STL units per lattice unit)"""
    L = load()
    L.hch_last_error.restype = C.c_char_p
    dims = (C.c_int32 * 3)()
    dx = C.c_double()
    if L.hch_voxelize_stl(str(path).encode(), C.c_int32(ref_dir_n), C.c_int32(ref_dir), dims, None, C.c_int64(0), C.byref(dx)):
        raise HcgError("hch_voxelize_stl: " + L.hch_last_error().decode())
    n = dims[0] * dims[1] * dims[2]
    fl = np.empty(n, dtype=np.uint8)
    if L.hch_voxelize_stl(str(path).encode(), C.c_int32(ref_dir_n), C.c_int32(ref_dir), dims, fl.ctypes.data_as(c_u8p), C.c_int64(n), C.byref(dx)):
        raise HcgError("hch_voxelize_stl: " + L.hch_last_error().decode())
    return fl.reshape(dims[0], dims[1], dims[2]), dx.value
