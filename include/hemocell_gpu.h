/*
 * hemocell_gpu.h -- C ABI of the B200-native HemoCell hot path (libhemocell_gpu.so).
 *
 * The reference (UvaCsl/HemoCell) has no FFI boundary: its per-timestep path is reached
 * through C++ classes.  Every entry point below names the reference interface it replaces
 * (paths relative to the reference tree).  Host arrays are caller-owned, plain pointers and
 * sizes; no pointer handed out by the library outlives hcg_destroy(); uploads/downloads are
 * synchronous with respect to the context's streams.  Nothing here throws or exit()s: every
 * call returns an hcg_status and hcg_last_error() holds the message.
 *
 * Conventions
 *   node index    idx = z + nz*(y + ny*x)  (patch/palabos.patch:245), x,y,z GLOBAL coordinates
 *   populations   pop[q*N + idx], q = 0..18 in Palabos D3Q19 order, stored as f_q - t_q,
 *                 POST-STREAM (what Palabos holds between collideAndStream() calls)
 *   node vectors  a[d*N + idx], d = 0..2
 *   flags         uint8: HCG_FLUID, HCG_BOUNCEBACK, HCG_VEL_* (velocity plane, OUTWARD normal),
 *                 HCG_ZH_VEL_* / HCG_ZH_PRES_* (Zou-He velocity / pressure node with per-node values)
 *   particles     AoS xyz, cells contiguous, vertices in vertexId order, cell types in the
 *                 order they were added:  a[3*(base(cell) + vertexId) + d]
 *   multi-GPU     one context per rank/GPU; the lattice is cut into n_ranks slabs along x.
 *                 Lattice up/downloads take/return the rank's own slab (nx_local*ny*nz nodes,
 *                 same index formula with x local); particle calls are per rank.
 */
#ifndef HEMOCELL_GPU_H
#define HEMOCELL_GPU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hcg_ctx hcg_ctx;
typedef int32_t hcg_status;               /* 0 = ok, < 0 = error */
#define HCG_OK 0
#define HCG_ERR_ARG (-1)
#define HCG_ERR_CUDA (-2)
#define HCG_ERR_STATE (-3)
#define HCG_ERR_NCCL (-4)
#define HCG_ERR_CAPACITY (-5)

enum { HCG_FLUID = 0, HCG_BOUNCEBACK = 1, HCG_VEL_XN = 2, HCG_VEL_XP = 3, HCG_VEL_YN = 4,
       HCG_VEL_YP = 5, HCG_VEL_ZN = 6, HCG_VEL_ZP = 7,
       /* Zou-He nodes with per-node values (hcg_lattice_set_bc_nodes), OUTWARD normal -x,+x,-y,+y,-z,+z:
        * velocity nodes = createZouHeBoundaryCondition3D()->addVelocityBoundary{0,1,2}{N,P} (helper/preInlet.cpp:399-436),
        * pressure nodes = WrappedZouHeBoundaryManager3D addPressureBoundary{0,1,2}{N,P} + setBoundaryDensity
        * (examples/pipeflow_with_preinlet/pipeflow_with_preinlet.cpp:125-133) */
       HCG_ZH_VEL_XN = 8, HCG_ZH_VEL_XP = 9, HCG_ZH_VEL_YN = 10, HCG_ZH_VEL_YP = 11, HCG_ZH_VEL_ZN = 12, HCG_ZH_VEL_ZP = 13,
       HCG_ZH_PRES_XN = 14, HCG_ZH_PRES_XP = 15, HCG_ZH_PRES_YN = 16, HCG_ZH_PRES_YP = 17, HCG_ZH_PRES_ZN = 18, HCG_ZH_PRES_ZP = 19 };
enum { HCG_MODEL_RBC_HIGHORDER = 0, HCG_MODEL_PLT_SIMPLE = 1,
       HCG_MODEL_HOST = 2 /* no device kernel: the caller evaluates the constitutive model itself (a user subclass of
                             CellMechanics, mechanics/cellMechanics.h:45) and uploads HCG_P_FORCE at the material cadence */ };
/* lattice fields */
enum { HCG_LAT_POP = 0, HCG_LAT_FORCE = 1, HCG_LAT_VELOCITY = 2, HCG_LAT_DENSITY = 3,
       HCG_LAT_PINEQ = 4 /* off-equilibrium momentum flux, 6 components xx xy xz yy yz zz (output path:
                            Cell::computeShearStress / computeStrainRateFromStress, io/FluidHdf5IO.hh:406-541) */ };
/* particle fields */
enum { HCG_P_POS = 0, HCG_P_VEL = 1, HCG_P_FORCE = 2, HCG_P_FREP = 3, HCG_P_F_AREA = 4,
       HCG_P_F_VOLUME = 5, HCG_P_F_BEND = 6, HCG_P_F_LINK = 7, HCG_P_F_VISC = 8, HCG_P_F_INNER = 9 };

/* MultiBlockLattice3D ctor + periodicity().toggle + GuoExternalForceBGKdynamics(1/tau)
 * (cases/performance_testing/performance_testing.cpp:55-67) */
typedef struct {
  int32_t nx, ny, nz;          /* global lattice */
  int32_t periodic[3];
  double tau;
  int32_t device;              /* CUDA device ordinal of this context */
  int32_t rank, n_ranks;       /* slab decomposition along x: rank r owns nx/n_ranks planes, the first nx % n_ranks
                                  ranks one more (hcg_slab) */
} hcg_domain;

/* CommonCellConstants (mechanics/commonCellConstants.h:39-86) + the k_* of the model
 * (mechanics/rbcHighOrderModel.h:34-51, mechanics/cellMechanics.h:50-78) */
typedef struct {
  int32_t model;               /* HCG_MODEL_* */
  int32_t n_vertices, n_triangles, n_edges, n_inner_edges;
  const int32_t* triangles;                   /* [T][3] triangle_list */
  const int32_t* edges;                       /* [E][2] edge_list */
  const int32_t* inner_edges;                 /* [I][2] inner_edge_list */
  const int32_t* vertex_vertexes;             /* [V][6] ring ordered, -1 padded */
  const int32_t* vertex_n_vertexes;           /* [V] */
  const int32_t* edge_bending_triangles;      /* [E][2] */
  const int32_t* edge_bending_outer_points;   /* [E][2] */
  const double* edge_length_eq;               /* [E] */
  const double* edge_angle_eq;                /* [E] */
  const double* triangle_area_eq;             /* [T] */
  const double* patch_dist_eq;                /* [V] surface_patch_center_dist_eq_list */
  const double* inner_edge_length_eq;         /* [I] */
  double volume_eq, area_mean_eq, edge_mean_eq;
  double k_volume, k_area, k_link, k_bend, eta_m;
} hcg_celltype;

typedef struct {               /* helper/profiler.h:47-76 key names, CUDA-event times */
  char name[40];
  double ms_total;
  int64_t calls;
} hcg_timer;

const char* hcg_last_error(const hcg_ctx*);   /* ctx may be NULL: error of a failed hcg_create */
const char* hcg_version(void);

/* ---- lifetime: HemoCell ctor/dtor + initializeLattice (core/hemoCell.cpp:69-127, 438-583) */
hcg_status hcg_create(const hcg_domain* d, hcg_ctx** out);
/* number of CUDA devices visible to the process (0 when there is none) */
int32_t    hcg_device_count(void);
/* the x-slab of a rank: first plane and number of planes (pure function of nx, rank, n_ranks) */
void       hcg_slab(int32_t nx, int32_t rank, int32_t n_ranks, int32_t* x0_out, int32_t* nxl_out);
void       hcg_destroy(hcg_ctx*);

/* ---- multi-GPU plumbing (replaces plb::plbInit/MPI_Init + ParallelBlockCommunicator3D).
 * The caller distributes the 128-byte NCCL unique id of rank 0 (e.g. torch.distributed
 * broadcast); one context per process/GPU. */
hcg_status hcg_comm_unique_id(void* out128);
hcg_status hcg_comm_init(hcg_ctx*, const void* id128);
/* Same role for a run whose n_ranks contexts all live in ONE process (one host thread per context; collective: returns
 * when every rank has joined the group named by id128, any 128 bytes unique to the run).  The contexts may share a GPU -
 * NCCL refuses two ranks on one device - so the slab decomposition (the reference's `mpirun -n 2` vs `-n 4` identity
 * check, scripts/ci/pipeflow_sanity.sh:25-32) can be verified on a single-GPU box.  Messages are matched like NCCL's
 * grouped send/recv and moved with cudaMemcpyPeerAsync (host-synchronous); the peer-store transport works unchanged. */
hcg_status hcg_comm_init_local(hcg_ctx*, const void* id128);

/* ---- lattice set-up */
/* defineDynamics(lattice, domain, new BounceBack) / setVelocityConditionOnBlockBoundaries
 * (examples/cube/cube.cpp:74-88, helper/hemocellInit.hh:64-73): flags of this rank's slab */
hcg_status hcg_lattice_set_flags(hcg_ctx*, const uint8_t* flags);
/* setBoundaryVelocity(lattice, plane, u) (helper/hemocellInit.hh:75-77): one wall velocity per
 * velocity-plane orientation, index = flag - HCG_VEL_XN */
hcg_status hcg_lattice_set_bc_velocity(hcg_ctx*, int32_t orientation, const double u[3]);
/* setBoundaryVelocity(lattice, Box3D(point), u) / setBoundaryDensity(lattice, box, rho) on Zou-He nodes
 * (helper/preInlet.cpp:380, examples/pipeflow_with_preinlet/pipeflow_with_preinlet.cpp:132): values of n nodes of this
 * rank's slab, node_idx = local node index, val = [n][4] (u_x, u_y, u_z, rho).  Velocity nodes use u, pressure nodes
 * rho; nodes never set hold (0, 0, 0, 1). */
hcg_status hcg_lattice_set_bc_nodes(hcg_ctx*, int64_t n, const int64_t* node_idx, const double* val);
/* Cell::computeVelocity of n nodes of this rank's slab from the CURRENT populations and node force
 * (helper/preInlet.cpp:372: the velocity the pre-inlet sends to the main domain's inlet nodes): u_out [n][3] */
hcg_status hcg_lattice_node_velocity(hcg_ctx*, int64_t n, const int64_t* node_idx, double* u_out);
/* HemoCell::latticeEquilibrium -> initializeAtEquilibrium (core/hemoCell.cpp:129-133) */
hcg_status hcg_lattice_init_equilibrium(hcg_ctx*, double rho, const double u[3]);
/* the setExternalVector(lattice, bbox, forceBeginsAt, f) every case file re-applies after
 * iterate() (examples/pipeflow/pipeflow.cpp:144-146): the value the node force is reset to */
hcg_status hcg_lattice_set_body_force(hcg_ctx*, const double f[3]);
/* the same with a different force per region, e.g. the two half-domains driven in opposite directions of
 * cases/kolmogorovFlow/kolmogorovFlow.cpp:138-142: f[d*N + idx] over this rank's slab.  A later
 * hcg_lattice_set_body_force() returns to the uniform value. */
hcg_status hcg_lattice_set_body_force_field(hcg_ctx*, const double* f);
hcg_status hcg_lattice_upload(hcg_ctx*, int32_t field /*POP|FORCE*/, const double* in);
hcg_status hcg_lattice_download(hcg_ctx*, int32_t field /*POP|FORCE|VELOCITY|DENSITY|PINEQ*/, double* out);

/* ---- cell types and cells */
/* HemoCell::addCellType<Model>(name, constructType) (hemocell.h:122-128) */
hcg_status hcg_celltype_add(hcg_ctx*, const hcg_celltype* t, int32_t* ctype_out);
/* HemoCell::loadParticles (core/hemoCell.cpp:191-197) after placement: all cells of one type,
 * call once per type in type order.  pos: [n_cells][V][3]. */
hcg_status hcg_cells_add(hcg_ctx*, int32_t ctype, int64_t n_cells, const int64_t* cell_id, const double* pos);
/* spare cell slots of a type for cells that arrive later (pre-inlet hand-over); must precede hcg_cells_add of the type,
 * which may then be called with n_cells = 0 */
hcg_status hcg_cells_reserve(hcg_ctx*, int32_t ctype, int64_t spare_cells);
hcg_status hcg_cells_count(hcg_ctx*, int64_t* n_cells_alive, int64_t* n_particles_alive);
/* the same two numbers without waiting: out2 = {cells, particles} must be page-locked host memory and is valid after the
 * next synchronising call (hcg_synchronize, hcg_iterate, any download); lets a host loop read a per-step result without
 * draining the launch queue every step */
hcg_status hcg_cells_count_async(hcg_ctx*, int64_t* out2);
/* whole-array particle access in storage order incl. deleted cells (alive_out marks them);
 * needed by per-operator parity tests and checkpoint restore.  n = hcg_cells_capacity */
hcg_status hcg_cells_capacity(hcg_ctx*, int64_t* n_cells, int64_t* n_particles);
hcg_status hcg_cells_upload(hcg_ctx*, int32_t field /*POS|VEL|FORCE|FREP*/, const double* in);
hcg_status hcg_cells_download(hcg_ctx*, int32_t field, double* out);
/* the same field as the output writers store it (io/ParticleHdf5IO.cpp: float datasets, SI units when outputInSiUnits):
 * converted and scaled on the device, 12 bytes per particle over PCIe instead of 24 */
hcg_status hcg_cells_download_f32(hcg_ctx*, int32_t field, double scale, float* out);
hcg_status hcg_cells_info(hcg_ctx*, int64_t* cell_id_out, int32_t* ctype_out, uint8_t* alive_out);
/* multi-GPU: 1 for the cell slots this rank OWNS (it holds the nearest node of the cell's vertex 0; replicas held for
 * the neighbour's sake are 0), so that per-cell output and statistics count every cell once.  All 1 on a single rank. */
hcg_status hcg_cells_owned(hcg_ctx*, uint8_t* owned_out);
/* element-wise reduction of a few host doubles over all ranks (the MPI reductions behind the reference's
 * HemoCellGatheringFunctional, helper/cellInfo.cpp, fluidInfo.cpp, particleInfo.cpp); op: 0 sum, 1 min, 2 max.
 * Collective; a no-op on a single rank. */
hcg_status hcg_allreduce(hcg_ctx*, double* inout, int64_t n, int32_t op);
/* HemoCellStretch::applyForce (helper/hemoCellStretch.cpp:63-78, 102-110): force[lsp] += f */
hcg_status hcg_cells_add_force(hcg_ctx*, int64_t n, const int64_t* particle_index, const double* f /*[n][3]*/);
/* stiffness update without re-uploading topology (CellMechanics k_* members) */
hcg_status hcg_celltype_set_stiffness(hcg_ctx*, int32_t ctype, double k_volume, double k_area,
                                      double k_link, double k_bend, double eta_m);

/* ---- knobs */
/* Parameters::f_limit (mechanics/constantConversion.cpp:55-57) */
hcg_status hcg_set_force_limit(hcg_ctx*, double f_limit_lbm);
/* setParticleVelocityUpdateTimeScaleSeparation / setRepulsionTimeScaleSeperation /
 * enableBoundaryParticles step / setMaterialTimeScaleSeparation (hemocell.h:158-176) */
hcg_status hcg_set_timescales(hcg_ctx*, int32_t velocity, int32_t repulsion, int32_t wall_repulsion);
hcg_status hcg_set_material_timescale(hcg_ctx*, int32_t ctype, int32_t every);
/* HemoCell::setRepulsion(k, cutoff) (core/hemoCell.cpp:420-426); cutoff in lattice units;
 * enabled != 0 switches the operator on inside hcg_iterate */
hcg_status hcg_set_repulsion(hcg_ctx*, int32_t enabled, double k, double cutoff_lu);
/* HemoCell::enableBoundaryParticles (core/hemoCell.cpp:428-436) */
hcg_status hcg_set_wall_repulsion(hcg_ctx*, int32_t enabled, double k, double cutoff_lu);
/* spreading strategy: 1 (default) = per-cell node-sorted (vertex, corner) pairs + warp-level reduction in
 * front of the fp64 atomics, permutation rebuilt every `resort_every` steps; 0 = one atomic per pair */
hcg_status hcg_set_spread_mode(hcg_ctx*, int32_t mode, int32_t resort_every);
/* EXPERIMENTAL (not yet measured): at tau = 1 on a fully periodic lattice without walls (cases/performance_testing) the lattice
 * state can be the four raw moments per node instead of the 19 populations; 1 = one kernel per step reads the neighbours'
 * moments and forces and writes the new moments (populations are materialised on demand), 2 = the same on slab-decomposed runs
 * (face planes read from the neighbours over NVLink, or exchanged with send/recv; every rank must use the same setting),
 * 0 = stored populations.  Default: on (environment HCG_MOMENT_ONLY=0 switches it off); it applies only where every rank's
 * lattice is plain periodic fluid at tau = 1, which the ranks agree on when the flags are set. */
hcg_status hcg_set_moment_only(hcg_ctx*, int32_t on);
/* multi-GPU particle exchange (replaces particleEnvelope of config.xml and the comm. structure of
 * HemoCellFields::calculateCommunicationStructure, core/hemoCellFields.cpp:363-372): a rank holds
 * every cell within `margin_lu` of its slab (2 lu of kernel support + the drift allowed between two looks at the
 * membership), membership is re-evaluated when the fastest vertex of any rank could have drifted 1 lu at its present speed,
 * at least `sync_every` and at most 8 x `sync_every` steps after the previous time; `slack` = fraction of spare cell
 * slots for arrivals.  Must precede hcg_cells_add. */
hcg_status hcg_set_exchange(hcg_ctx*, double margin_lu, int32_t sync_every, double slack);
/* multi-GPU transport of the per-step exchanges (lattice ghost planes, node velocity ghosts, shared-cell
 * velocity sync; replaces the MPI messages of Palabos' block communicator and of
 * HemoCellParticleDataTransfer::send/receive, core/hemoCellParticleDataTransfer.cpp:67-180):
 * 1 (default) = NVLink peer memory: the producing kernels store into the neighbour's buffers, a flag
 * kernel synchronises (needs P2P access between neighbouring GPUs of one box); 0 = NCCL send/recv.
 * Must precede hcg_comm_init.  Environment HCG_TRANSPORT=nccl|peer overrides. */
hcg_status hcg_set_transport(hcg_ctx*, int32_t transport);
hcg_status hcg_exchange_stats(hcg_ctx*, int64_t* shared_left, int64_t* shared_right,
                              int64_t* migrated_in, int64_t* migrated_out);
hcg_status hcg_set_iteration(hcg_ctx*, int64_t iter);
hcg_status hcg_get_iteration(hcg_ctx*, int64_t* iter);

/* ---- run: HemoCell::iterate() x n (core/hemoCell.cpp:299-376), incl. the case file's
 * lattice->collideAndStream() warm-up loop when fluid_only != 0 */
hcg_status hcg_iterate(hcg_ctx*, int64_t n_steps);
/* the same steps enqueued without the final wait (the case-file loop `iterate(); setExternalVector(...)` then keeps the
 * launch queue full); errors of the device work surface at the next synchronising call */
hcg_status hcg_iterate_async(hcg_ctx*, int64_t n_steps);
hcg_status hcg_fluid_warmup(hcg_ctx*, int64_t n_steps);

/* ---- pre-inlet (helper/preInlet.cpp): a second, periodic, force-driven context `pre` feeds the inlet of `main`.
 * Both are single-rank contexts, on one GPU or on two GPUs of the box; they iterate independently (own streams). */
/* PreInlet::initializePreInletVelocityBoundary (helper/preInlet.cpp:399-436): n node pairs, pre-inlet node pre_idx[k]
 * (a fluid node of its coupling plane) drives main's Zou-He velocity node main_idx[k] */
hcg_status hcg_preinlet_map(hcg_ctx* main, hcg_ctx* pre, int64_t n, const int64_t* pre_idx, const int64_t* main_idx);
/* PreInlet::applyPreInletVelocityBoundary (helper/preInlet.cpp:344-397), device to device: Cell::computeVelocity of the
 * pre-inlet nodes (current populations + node force) -> boundary velocity of the mapped main nodes; stream-ordered
 * between the two contexts by events, no host synchronisation */
hcg_status hcg_preinlet_apply_velocity(hcg_ctx* main);
/* PreInlet::applyPreInletParticleBoundary (helper/preInlet.cpp:255-342) in whole cells: a cell of the pre-inlet is
 * copied (positions + shift, velocity, force, repulsion force) into a free slot of main the first time its periodic
 * image k (position + k*period along `axis`) lies wholly inside [slab_lo, slab_hi] (main coordinates, = the
 * reference's inflow slab of particleEnvelope planes behind the inlet); its id becomes id + z(k)*id_stride, z = 0, 1, 2, 3, 4, ... for k = 0, -1, 1, -2, 2, ... (unique and non-negative; the
 * reference's cellId += offset*number_of_cells per wrap, core/hemoCellParticleDataTransfer.cpp:33-66).  The pre-inlet
 * keeps its cell.  Deviation from the reference: partially entered cells are not mirrored vertex by vertex. */
hcg_status hcg_preinlet_apply_cells(hcg_ctx* main, int32_t axis, double period, const double shift[3],
                                    double slab_lo, double slab_hi, int64_t id_stride, int64_t* n_added);

/* checkpoint support: read (set = 0) or write (set != 0) the periodic image handed over last per pre-inlet cell slot
 * (INT64_MIN = never), n = the pre-inlet context's cell capacity */
hcg_status hcg_preinlet_laps(hcg_ctx* main, int64_t n, int64_t* laps, int32_t set);

/* ---- per-operator entry points (single-step parity; same order as iterate()) */
hcg_status hcg_op_repulsion(hcg_ctx*);       /* HemoCellFields::applyRepulsionForce          */
hcg_status hcg_op_wall_repulsion(hcg_ctx*);  /* HemoCellFields::applyBoundaryRepulsionForce  */
hcg_status hcg_op_spread(hcg_ctx*);          /* HemoCellFields::spreadParticleForce          */
hcg_status hcg_op_collide_stream(hcg_ctx*);  /* lattice->collideAndStream()                  */
hcg_status hcg_op_interpolate(hcg_ctx*);     /* HemoCellFields::interpolateFluidVelocity     */
hcg_status hcg_op_sync(hcg_ctx*);            /* HemoCellFields::syncEnvelopes: velocity swap of the shared cells + whole-cell
                                                migration (collective over the ranks; nothing to do on one rank) */
hcg_status hcg_op_advance(hcg_ctx*);         /* HemoCellFields::advanceParticles             */
hcg_status hcg_op_mechanics(hcg_ctx*, int32_t forced, int32_t components); /* applyConstitutiveModel */
hcg_status hcg_op_zero_force(hcg_ctx*);      /* setExternalVector(lattice, bbox, 0 | body)   */

/* ---- observables (helper/cellInfo.cpp, helper/fluidInfo.cpp): per cell in storage order */
hcg_status hcg_cells_bbox(hcg_ctx*, double* bbox /*[n_cells][6] xmin xmax ymin ymax zmin zmax*/);
hcg_status hcg_cells_volume_area(hcg_ctx*, double* volume, double* area);
/* CellInformationFunctionals::calculateCellStretch (helper/cellInfo.cpp:103-121): max pairwise vertex distance */
hcg_status hcg_cells_stretch(hcg_ctx*, double* stretch);
hcg_status hcg_fluid_velocity_stats(hcg_ctx*, double* vmin, double* vmax, double* vmean);

/* ---- timing (helper/profiler.cpp): CUDA-event totals under the reference's key names */
hcg_status hcg_timers_enable(hcg_ctx*, int32_t on);
hcg_status hcg_timers(hcg_ctx*, hcg_timer* out, int32_t* n_inout);
hcg_status hcg_timers_reset(hcg_ctx*);
/* number of kernels launched by this context so far */
hcg_status hcg_launch_count(hcg_ctx*, int64_t* n);
hcg_status hcg_synchronize(hcg_ctx*);

/* ---- device-resident benchmark support: step with CUDA events bracketing n_steps,
 * returns elapsed milliseconds on the context's stream */
hcg_status hcg_iterate_timed(hcg_ctx*, int64_t n_steps, double* ms_out);

#ifdef __cplusplus
}
#endif
#endif
