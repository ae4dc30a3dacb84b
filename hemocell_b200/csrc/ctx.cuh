// Internal definitions shared by the CUDA translation units of libhemocell_gpu.so.
// Device memory layout:
//   lattice: slab of nxl planes + 1 ghost plane on each x side; plane = ny*nz nodes;
//            local node index  n = z + nz*(y + ny*lx), lx = x - x0 + 1 in [0, nxl+1]
//   populations g[q*S + n], S = (nxl+2)*ny*nz: the PRE-STREAMED (post-collision) value that
//            will arrive at node n + c_q; the reference's post-stream state is S_q(n) = g_q(n - c_q)
//   node force F[4*n + d], velocity U[4*n + d] (AoS, one 32-byte sector per node; U slot 3 = density), flags[n]
//   particles: SoA by component, p = cell_base + vertexId
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <map>
#include "../../include/hemocell_gpu.h"
#include "../../include/hemocell_host.h"

#define HCG_MAX_TYPES 8
#define HCG_MAX_RING 6

struct CellTypeDev {
  int model, V, T, E, I;
  int *tri, *edge, *inner, *ring, *nring, *bend_tri, *bend_outer;
  double *edge_len_eq, *edge_ang_eq, *tri_area_eq, *patch_eq, *inner_len_eq;
  // per-vertex gather tables (built on the host in hcg_celltype_add)
  int *vt;    // [V][6]  incident triangles, ascending, -1 padded (PLT)
  unsigned long long *rg;   // [6][V]  RBC, per ring slot: ring vertex | edge << 16 | triangle << 32 | ring size << 48 | codes (hcg_celltype_add)
  int *vpe;   // [V][12] PLT: edges touching v as end or outer point: 4*e + role, ascending
  int *vin;   // [V][4]  PLT inner edges: 2*e + side
  double volume_eq, area_mean_eq, edge_mean_eq;
  double k_volume, k_area, k_link, k_bend, eta_m;
};

struct CellTypeHost {
  CellTypeDev d;
  int64_t n_cells = 0, first_cell = 0, first_particle = 0;   // n_cells = slots in use (high-water mark)
  int64_t cap_cells = 0;                                      // slots reserved (multi-GPU arrivals)
  int64_t reserve = 0;                                        // extra spare slots asked for by hcg_cells_reserve (pre-inlet arrivals)
  uint16_t* perm = nullptr;                                   // node-sorted (vertex, corner) pairs per cell (spread_sorted.cu)
  int timescale = 1;
  std::vector<void*> allocs;
};

// multi-GPU particle exchange state (csrc/multi.cu)
struct MultiFace { int32_t* d_cells = nullptr; int64_t* d_off = nullptr; int n = 0, cap = 0; int64_t total = 0; };
struct MultiState {
  double margin = 4.0;            // hold region = slab +- margin lattice units
  int sync_every = 20;            // membership re-evaluation: at least this many steps apart ...
  int64_t next_sync_iter = 0;     // ... and due at this iteration (multi_rebalance sets it from the fastest vertex: see there)
  double* d_vmax = nullptr;       // largest velocity component of the held particles (device scalar)
  double slack = 0.3;             // spare cell slots per type for arrivals
  MultiFace face[2];              // cells shared through the left / right face, sorted by global id
  MultiFace all;                  // union of the two lists (each shared cell once)
  uint8_t* d_cell_shared = nullptr; int64_t cell_shared_cap = 0;   // per cell slot: 1 = on a shared list
  std::vector<uint8_t> h_held, h_shared[2];
  std::vector<std::vector<int32_t>> free_slots;
  double* sync_buf = nullptr; size_t sync_cap = 0;
  double* mig_buf = nullptr; size_t mig_cap = 0;
  int64_t* d_cnt = nullptr; double** d_arr = nullptr;
  int64_t migrated_in = 0, migrated_out = 0;
  int64_t* d_meta = nullptr; size_t meta_cap = 0;  // migration meta messages (ids, types)
  MultiFace tmp_send[2], tmp_recv[2];              // device lists of the cells leaving / arriving in a rebalance
  double* d_bbox = nullptr; size_t bbox_cap = 0;   // bounding boxes of the cell slots (multi_rebalance)
  size_t sync_half = 0;           // receive-buffer half of the velocity sync in flight (multi_sync_pack_post .. wait_unpack_advance)
};

// NVLink peer-memory transport (csrc/peer.cu): the slab neighbours' lattice buffers, flag words and
// velocity-sync receive buffers mapped into this context (CUDA IPC across processes, direct peer
// access inside one process), so that halo planes and shared-cell velocities are STORED into the
// neighbour by the producing kernel and a one-CTA flag kernel replaces the NCCL send/recv pair.
#define HCG_PEER_NPTR 10           // g[0], g[1], U, flag words, sync_recv[left face], sync_recv[right face], W / F buffers of the moment-only update (2 each)
struct PeerLink {
  int rank = -1;                   // neighbour rank, -1 = none (non-periodic end)
  void* ptr[HCG_PEER_NPTR] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};
struct PeerMap { unsigned char handle[64]; void* base; };
struct PeerState {
  int transport = 1;               // 1 = peer memory (default), 0 = NCCL send/recv
  bool ready = false;
  bool usable = true;              // false: this rank could not export / map peer memory (see peer_setup)
  PeerLink link[2];                // left, right
  unsigned long long* flags = nullptr;   // my flag words: [0] written by the left neighbour, [1] by the right
  unsigned long long epoch = 0, sync_count = 0;
  double* sync_recv[2] = {nullptr, nullptr}; size_t sync_recv_cap[2] = {0, 0};
  std::vector<PeerMap> maps;       // IPC handles opened by this context
  std::vector<void*> retired;      // outgrown receive buffers: a neighbour may still map them, freed at destroy
  void* d_blob = nullptr;
};

struct PreInletState;   // csrc/preinlet.cu

struct TimerSlot { std::string name; double ms = 0; int64_t calls = 0; };
struct TimerPending { int slot; cudaEvent_t a, b; };

struct hcg_ctx {
  hcg_domain dom;
  int nxl, x0;                 // slab
  int64_t P, S, Nl;            // plane size, padded slab size, real nodes
  double omega;
  double bc_vel[6][3]; double* d_bc;
  double body[3];
  double f_limit;
  // lattice
  double *g[2]; int cur;       // double-buffered pre-streamed populations
  double *F, *U, *rho;         // node force, interpolation velocity (+ density scratch)
  bool f_clean = false;         // the node force holds exactly the reset value (body / F0) everywhere: re-applying it is free
  double* F0 = nullptr;         // optional per-node driving force the node force is reset to (AoS [n][4]); null = uniform body[3]
  double* W = nullptr; bool w_valid = false;   // tau = 1 fast path: raw moments (rhoBar, j) of the current populations, AoS [n][4]
  // moment-only update at tau = 1 (lattice.cu: k_moment_step; opt-in): second W / F buffers = the inputs of the last such step
  double* W2 = nullptr; double* F2 = nullptr;
  bool pops_stale = false;     // the populations lag behind W (materialised on demand by lat_ensure_pops)
  double* Wphys[2] = {nullptr, nullptr}; double* Fphys[2] = {nullptr, nullptr};   // the two W / F buffers by identity (what the slab neighbours map)
  int mo_flip = 0;             // W == Wphys[mo_flip], F == Fphys[mo_flip]: flips with every moment-only step, in lockstep on all ranks
  int mo_mode = -1;            // -1 = follow HCG_MOMENT_ONLY (default on), 0 = off, >= 1 = on (hcg_set_moment_only)
  bool mo_ok = false;          // tau = 1 and the whole lattice - on every rank - is plain periodic fluid (agreed in refresh_nonfluid)
  uint8_t* flags;
  bool u_valid, has_velbc, has_nonfluid;
  bool has_iobc = false;       // Zou-He velocity / pressure nodes present (flags >= HCG_ZH_VEL_XN)
  double* bcn = nullptr;       // their per-node values, AoS [n][4] = (u_x, u_y, u_z, rho) over the padded slab; allocated on first use
  bool real_nonfluid = false;  // any non-fluid flag on this rank's real nodes (has_nonfluid also covers the ghost planes)
  // lattices with walls: cells with no non-fluid node within reach skip the flag look-ups of the IBM kernels (ibm.cu: far_classify)
  uint8_t* wall_coarse = nullptr; int wc_dim[3] = {0, 0, 0}; bool wall_coarse_valid = false;   // 8^3 blocks of the padded slab holding a non-fluid node
  int* far_typeV = nullptr; int far_ntypes = -1;
  int* bbox_typeV = nullptr; int bbox_ntypes = -1;   // mech_bbox scratch
  bool in_iterate = false;     // inside hcg_iterate*: the step cadence bounds how far a cell can drift between classifications
  uint8_t* cell_far = nullptr; int64_t cell_far_cap = 0; int far_steps_left = 0;               // per cell slot; valid for far_steps_left more advances
  // particles
  int64_t np, ncells, cap_p, cap_c;
  double *pos[3], *vel[3], *frc[3], *frep[3];
  double *comp[6][3];          // optional per-component force arrays
  bool comp_alloc;
  int32_t* p_cell;             // particle -> cell slot
  uint8_t* cell_alive;         // per cell
  int32_t* cell_type;          // per cell (device)
  int64_t* cell_base;          // per cell (device) first particle
  int64_t* cell_gid = nullptr; bool cell_gid_dirty = true; int64_t cell_gid_cap = 0;   // per cell (device) global id, on demand
  std::vector<int64_t> h_cell_id; std::vector<int32_t> h_cell_type; std::vector<int64_t> h_cell_base;
  std::vector<CellTypeHost> types;
  // repulsion
  bool rep_on, wall_on; double rep_k, rep_cut, wall_k, wall_cut;
  int ts_vel, ts_rep, ts_wall;
  int *bin_count, *bin_start, *bin_items; int64_t* wall_nodes; int64_t n_wall; bool wall_built;
  void* scan_tmp; size_t scan_tmp_bytes;
  int64_t iter;
  int spread_mode; bool perm_valid; int perm_every;   // 1 = node-sorted pairs + warp reduction, 0 = plain atomics
  // execution
  cudaStream_t stream, stream_halo;
  cudaStream_t stream_lo = nullptr;   // low-priority stream: bulk work that overlaps the exchange chain of the main stream
  cudaEvent_t ev_a, ev_b;
  void* nccl;                  // ncclComm_t
  void* local = nullptr;       // host-staged communicator endpoint (comm.cu; hcg_comm_init_local)
  int* comm_scratch = nullptr; // device scratch of the small host-value reductions
  MultiState multi;
  PeerState peer;
  double* halo_send[2]; double* halo_recv[2];
  int* d_qsets = nullptr;      // population index sets of the halo exchange (lattice.cu), on this context's device
  bool timers_on; std::vector<TimerSlot> timers; std::map<std::string,int> timer_idx;
  std::vector<cudaEvent_t> ev_pool; std::vector<TimerPending> ev_pending;
  int64_t launches;
  unsigned long long* count_dev = nullptr; int* count_typeV = nullptr; int count_ntypes = -1;   // hcg_cells_count scratch
  int sm_count = 0, smem_optin = 0;
  int* fused_done = nullptr;   // per-plane completion counters (collision kernel -> overlapped moments kernel)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::string err;
  double* staging; size_t staging_bytes;   // device scratch for AoS<->SoA transposes
  PreInletState* preinlet = nullptr;       // coupling to a pre-inlet context (this context is the main domain)
};

hcg_status hcg_fail(hcg_ctx* c, hcg_status code, const std::string& msg);
#define CUDA_TRY(c, expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) \
  return hcg_fail((c), HCG_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); } while (0)
#define KERNEL_CHECK(c) do { (c)->launches++; cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) \
  return hcg_fail((c), HCG_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e_)); } while (0)

// host -> device copy that has LANDED when the call returns: a plain cudaMemcpy from pageable memory may return once the data is
// staged, and the context's non-blocking stream is not ordered behind the legacy default stream
inline cudaError_t hcg_h2d(hcg_ctx* c, void* dst, const void* src, size_t bytes) {
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream);
  return e != cudaSuccess ? e : cudaStreamSynchronize(c->stream);
}

// CUDA-event timers under the reference's Profiler key names (helper/profiler.cpp).  Events are
// recorded on the launching stream without any host synchronisation; they are resolved in
// resolve_timers() (hcg_timers), so enabling them does not serialise the step.
struct OpTimer {
  hcg_ctx* c; int slot; cudaEvent_t a, b; bool on;
  OpTimer(hcg_ctx* c_, const char* name) : c(c_), slot(-1), on(c_->timers_on) {
    if (!on) return;
    auto it = c->timer_idx.find(name);
    if (it == c->timer_idx.end()) { slot = (int)c->timers.size(); c->timers.push_back({name, 0.0, 0}); c->timer_idx[name] = slot; }
    else slot = it->second;
    if (c->ev_pool.size() < 2) { cudaEvent_t e; cudaEventCreate(&e); c->ev_pool.push_back(e); cudaEventCreate(&e); c->ev_pool.push_back(e); }
    a = c->ev_pool.back(); c->ev_pool.pop_back(); b = c->ev_pool.back(); c->ev_pool.pop_back();
    cudaEventRecord(a, c->stream);
  }
  ~OpTimer() { if (!on) return; cudaEventRecord(b, c->stream); c->ev_pending.push_back({slot, a, b}); }
};
inline void resolve_timers(hcg_ctx* c) {
  cudaStreamSynchronize(c->stream);
  for (auto& p : c->ev_pending) {
    float ms = 0; cudaEventElapsedTime(&ms, p.a, p.b);
    c->timers[p.slot].ms += ms; c->timers[p.slot].calls++;
    c->ev_pool.push_back(p.a); c->ev_pool.push_back(p.b);
  }
  c->ev_pending.clear();
}

// lattice.cu
hcg_status lat_collide_stream(hcg_ctx* c, bool reset_force);
hcg_status lat_collide_rows(hcg_ctx* c, bool reset_force, int row0, int row1, cudaStream_t st, int* done);
hcg_status lat_moments(hcg_ctx* c, bool reset_force, bool want_rho);
hcg_status lat_collide_moments_overlapped(hcg_ctx* c, bool* done_out);
hcg_status lat_reset_force(hcg_ctx* c);
hcg_status lat_init_equilibrium(hcg_ctx* c, double rho, const double u[3]);
hcg_status lat_halo_exchange_pop(hcg_ctx* c);
hcg_status lat_halo_exchange_u(hcg_ctx* c);
hcg_status lat_exchange_byte_planes(hcg_ctx* c, uint8_t* buf, int ghost_default);   // per-node byte field: face planes -> neighbours' ghosts
hcg_status lat_pop_to_reference(hcg_ctx* c, double* dst_dev);      // S_q(n) = g_q(n - c_q), compact slab
hcg_status lat_pop_from_reference(hcg_ctx* c, const double* src_dev);
hcg_status lat_pineq(hcg_ctx* c, double* dst_dev);                 // off-equilibrium momentum flux, compact SoA [6][Nl]
hcg_status lat_velocity_stats(hcg_ctx* c, double* vmin, double* vmax, double* vmean);
bool lat_moment_eligible(hcg_ctx* c);
hcg_status lat_moment_step(hcg_ctx* c, bool write_u);
hcg_status lat_ensure_pops(hcg_ctx* c);
hcg_status lat_moment_buffers(hcg_ctx* c);   // W, W2, F2 of the moment-only update (idempotent)
hcg_status lat_bcn_ensure(hcg_ctx* c);
hcg_status lat_bcn_scatter(hcg_ctx* c, int64_t n, const int64_t* idx_dev, const double* val_dev, bool keep_rho, cudaStream_t st);
hcg_status lat_node_velocity(hcg_ctx* c, int64_t n, const int64_t* idx_dev, double* out_dev, cudaStream_t st);   // out [n][4] = (u, rho)
// ibm.cu
hcg_status ibm_far_classify(hcg_ctx* c, int valid_steps);  // which cells cannot meet a non-fluid node during the next valid_steps advances
inline const uint8_t* ibm_far(const hcg_ctx* c) { return (c->far_steps_left > 0) ? c->cell_far : nullptr; }
hcg_status ibm_spread(hcg_ctx* c);
hcg_status ibm_interpolate(hcg_ctx* c);
hcg_status ibm_advance(hcg_ctx* c);
hcg_status ibm_interpolate_advance(hcg_ctx* c);
hcg_status ibm_interpolate_advance_unshared(hcg_ctx* c);   // multi-GPU: interpolate all, advance the cells no neighbour holds
hcg_status ibm_advance_shared(hcg_ctx* c);                 // ... and the shared ones after the velocity sync
hcg_status ibm_interpolate_shared(hcg_ctx* c);             // overlap variant: shared cells only (main stream)
hcg_status ibm_interpolate_advance_unshared_on(hcg_ctx* c, cudaStream_t st);   // ... unshared cells only, on a second stream
// mechanics.cu
hcg_status mech_apply(hcg_ctx* c, int ctype, bool components);
hcg_status mech_bbox(hcg_ctx* c, double* out_dev);
hcg_status mech_volume_area(hcg_ctx* c, double* vol_dev, double* area_dev);
hcg_status mech_stretch(hcg_ctx* c, double* out_dev);   // max pairwise vertex distance per cell slot
// spread_sorted.cu
bool spread_sorted_supported(const CellTypeHost& th);
hcg_status spread_sorted_rebuild(hcg_ctx* c);
hcg_status spread_sorted(hcg_ctx* c);
// multi.cu
hcg_status multi_velocity_sync(hcg_ctx* c);                // = multi_field_sync(c, 0)
hcg_status multi_velocity_sync_advance(hcg_ctx* c);        // ... fused with the advance of the shared cells
hcg_status multi_sync_pack_post(hcg_ctx* c);               // the same in two halves: pack + publish ...
hcg_status multi_sync_wait_unpack_advance(hcg_ctx* c);     // ... wait + unpack + advance
hcg_status multi_field_sync(hcg_ctx* c, int field);        // 0 = velocity + alive flags, 1 = repulsion force
hcg_status multi_upload_cell_gid(hcg_ctx* c);
hcg_status multi_rebalance(hcg_ctx* c, bool initial);
hcg_status multi_neighbour_exchange(hcg_ctx* c, const void* sendL, size_t nsL, const void* sendR, size_t nsR,
                                    void* recvR, size_t nrR, void* recvL, size_t nrL);
// comm.cu: send/recv + reductions over NCCL or over the in-process communicator
bool comm_up(const hcg_ctx* c);
hcg_status comm_group_begin(hcg_ctx* c);
void comm_send(hcg_ctx* c, const void* p, size_t bytes, int peer);
void comm_recv(hcg_ctx* c, void* p, size_t bytes, int peer);
hcg_status comm_group_end(hcg_ctx* c, const char* what);
hcg_status comm_allreduce_f64(hcg_ctx* c, double* dev, size_t n, int op);
hcg_status comm_allreduce_min_host(hcg_ctx* c, int* value, int n = 1);
hcg_status comm_nccl_init(hcg_ctx* c, const void* id128);
hcg_status comm_local_init(hcg_ctx* c, const void* id128);
void comm_destroy(hcg_ctx* c);
// peer.cu
inline bool peer_on(const hcg_ctx* c) { return c->dom.n_ranks > 1 && c->peer.transport == 1 && c->peer.ready; }
hcg_status peer_setup(hcg_ctx* c);                        // collective over slab neighbours (NCCL must be up)
hcg_status peer_barrier(hcg_ctx* c);                      // publish "everything before this is stored", wait for both neighbours
hcg_status peer_post(hcg_ctx* c);                         // ... the two halves as separate launches
hcg_status peer_wait(hcg_ctx* c);
hcg_status peer_reserve_sync(hcg_ctx* c, size_t doubles_left, size_t doubles_right, bool* changed);
void peer_destroy(hcg_ctx* c);
// preinlet.cu
void preinlet_destroy(hcg_ctx* c);
// repulsion.cu
hcg_status rep_apply(hcg_ctx* c);
hcg_status rep_wall_apply(hcg_ctx* c);

// D3Q19, Palabos order; opposite(i) = i + 9
__device__ __constant__ static const int d_cx[19] = {0,-1,0,0,-1,-1,-1,-1,0,0, 1,0,0,1,1,1,1,0,0};
__device__ __constant__ static const int d_cy[19] = {0,0,-1,0,-1,1,0,0,-1,-1, 0,1,0,1,-1,0,0,1,1};
__device__ __constant__ static const int d_cz[19] = {0,0,0,-1,0,0,-1,1,-1,1, 0,0,1,0,0,1,-1,1,-1};
