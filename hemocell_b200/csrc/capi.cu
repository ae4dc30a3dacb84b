// C ABI of libhemocell_gpu.so (include/hemocell_gpu.h): context lifetime, up/downloads,
// the iterate() scheduler (reference core/hemoCell.cpp:299-376) and per-operator entry points.
#include "ctx.cuh"
#include <nccl.h>
#include <algorithm>
#include "mech_tables.h"
#include <cstring>
#include <cstdio>
#include <cstdlib>

hcg_status lat_pad3(hcg_ctx* c, const double* src_dev, double* dst);
hcg_status lat_unpad(hcg_ctx* c, const double* src, double* dst_dev, int first, int ncomp);

static thread_local std::string g_create_error;

hcg_status hcg_fail(hcg_ctx* c, hcg_status code, const std::string& msg) {
  if (c) c->err = msg; else g_create_error = msg;
  return code;
}

// one byte per node (flags, wall mask): my face planes -> the neighbours' ghost planes; ghosts beyond a
// non-periodic end of the domain keep `ghost_default`
hcg_status lat_exchange_byte_planes(hcg_ctx* c, uint8_t* buf, int ghost_default) {
  const int R = c->dom.n_ranks, r = c->dom.rank; const bool px = c->dom.periodic[0];
  const int64_t P = c->P;
  CUDA_TRY(c, cudaMemsetAsync(buf, ghost_default, P, c->stream));
  CUDA_TRY(c, cudaMemsetAsync(buf + (int64_t)(c->nxl+1)*P, ghost_default, P, c->stream));
  if (R == 1) {
    if (px) {
      CUDA_TRY(c, cudaMemcpyAsync(buf, buf + (int64_t)c->nxl*P, P, cudaMemcpyDeviceToDevice, c->stream));
      CUDA_TRY(c, cudaMemcpyAsync(buf + (int64_t)(c->nxl+1)*P, buf + P, P, cudaMemcpyDeviceToDevice, c->stream));
    }
    return HCG_OK;
  }
  if (!comm_up(c)) return HCG_OK;     // done again from hcg_comm_init
  const int left = (r == 0) ? (px ? R - 1 : -1) : r - 1;
  const int right = (r == R - 1) ? (px ? 0 : -1) : r + 1;
  hcg_status s = comm_group_begin(c); if (s) return s;
  if (left >= 0) comm_send(c, buf + P, P, left);
  if (right >= 0) comm_send(c, buf + (int64_t)c->nxl*P, P, right);
  if (right >= 0) comm_recv(c, buf + (int64_t)(c->nxl+1)*P, P, right);
  if (left >= 0) comm_recv(c, buf, P, left);
  return comm_group_end(c, "byte-plane exchange");
}

namespace {

inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1)/t); }

__global__ void k_aos_to_soa(const double* __restrict__ in, double* x, double* y, double* z, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= n) return;
  x[i] = in[3*i]; y[i] = in[3*i+1]; z[i] = in[3*i+2];
}
__global__ void k_soa_to_aos(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                             double* out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[3*i] = x[i]; out[3*i+1] = y[i]; out[3*i+2] = z[i];
}
// particle field as the output writers want it: float32 triples scaled to SI (io/ParticleHdf5IO.cpp writes float datasets)
__global__ void k_soa_to_aos_f32(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                                 float* out, int64_t n, double scale) {
  const int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[3*i] = (float)(x[i]*scale); out[3*i+1] = (float)(y[i]*scale); out[3*i+2] = (float)(z[i]*scale);
}
__global__ void k_add_force(int64_t n, const int64_t* __restrict__ idx, const double* __restrict__ f,
                            double* fx, double* fy, double* fz, int64_t np) {
  const int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t p = idx[i];
  if (p < 0 || p >= np) return;
  atomicAdd(fx + p, f[3*i]); atomicAdd(fy + p, f[3*i+1]); atomicAdd(fz + p, f[3*i+2]);
}
__global__ void k_pad_flags(const uint8_t* __restrict__ src, uint8_t* dst, int64_t Nl, int64_t P) {
  const int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (i < Nl) dst[i + P] = src[i];
}
// multi-GPU: a cell is counted by the rank that owns its vertex 0 (replicas are not counted twice)
__global__ void k_count_alive(const uint8_t* __restrict__ alive, const int32_t* __restrict__ ctype,
                              const int* __restrict__ typeV, int64_t n, unsigned long long* out,
                              const int64_t* __restrict__ cell_base, const double* __restrict__ x,
                              int nranks, int nx, int px, int x0, int nxl) {
  const int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= n || !alive[i]) return;
  if (nranks > 1) {
    int gx = (int)floor(x[cell_base[i]] + 0.5);
    if (px) { gx %= nx; if (gx < 0) gx += nx; }
    int rel = gx - x0; if (px && rel < 0) rel += nx;
    if (rel < 0 || rel >= nxl) return;
  }
  atomicAdd(out, 1ULL); atomicAdd(out + 1, (unsigned long long)typeV[ctype[i]]);
}

__global__ void k_cells_owned(const uint8_t* __restrict__ alive, int64_t n, const int64_t* __restrict__ cell_base,
                              const double* __restrict__ x, int nranks, int nx, int px, int x0, int nxl, uint8_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t o = alive[i] ? 1 : 0;
  if (o && nranks > 1) {
    int gx = (int)floor(x[cell_base[i]] + 0.5);
    if (px) { gx %= nx; if (gx < 0) gx += nx; }
    int rel = gx - x0; if (px && rel < 0) rel += nx;
    if (rel < 0 || rel >= nxl) o = 0;
  }
  out[i] = o;
}

hcg_status ensure_staging(hcg_ctx* c, size_t bytes) {
  if (c->staging_bytes >= bytes) return HCG_OK;
  if (c->staging) cudaFree(c->staging);
  c->staging = nullptr; c->staging_bytes = 0;
  CUDA_TRY(c, cudaMalloc(&c->staging, bytes));
  c->staging_bytes = bytes;
  return HCG_OK;
}

template <class T> hcg_status upload_table(hcg_ctx* c, CellTypeHost& th, T** dst, const T* src, size_t n) {
  const size_t bytes = sizeof(T)*(n > 0 ? n : 1);
  CUDA_TRY(c, cudaMalloc(dst, bytes));
  th.allocs.push_back(*dst);
  if (n > 0) CUDA_TRY(c, hcg_h2d(c, *dst, src, sizeof(T)*n));
  return HCG_OK;
}

hcg_status exchange_flags(hcg_ctx* c) { return lat_exchange_byte_planes(c, c->flags, HCG_BOUNCEBACK); }

// has_nonfluid = the IBM kernels must look at node flags: real nodes or ghost planes (neighbour's face /
// outside of a non-periodic domain) hold something that is not plain fluid
hcg_status refresh_nonfluid(hcg_ctx* c) {
  std::vector<uint8_t> gh(2*(size_t)c->P);
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  CUDA_TRY(c, cudaMemcpy(gh.data(), c->flags, c->P, cudaMemcpyDeviceToHost));
  CUDA_TRY(c, cudaMemcpy(gh.data() + c->P, c->flags + (int64_t)(c->nxl+1)*c->P, c->P, cudaMemcpyDeviceToHost));
  bool nf = c->real_nonfluid;
  // ghost planes beyond a non-periodic end of the domain are never addressed by the IBM kernels
  const bool useL = c->dom.periodic[0] || c->dom.rank > 0, useR = c->dom.periodic[0] || c->dom.rank < c->dom.n_ranks - 1;
  for (int64_t i = 0; i < c->P && !nf && useL; i++) nf = gh[i] != HCG_FLUID;
  for (int64_t i = 0; i < c->P && !nf && useR; i++) nf = gh[c->P + i] != HCG_FLUID;
  c->has_nonfluid = nf;
  // moment-only lattice update (lattice.cu): every rank must take the same path, so the ranks agree here on whether the
  // whole lattice is plain periodic fluid at tau = 1 (collective, like the flag exchange above)
  int ok = (c->omega == 1.0 && c->dom.periodic[0] && c->dom.periodic[1] && c->dom.periodic[2] && !c->real_nonfluid && !nf &&
            !c->has_velbc && !c->has_iobc) ? 1 : 0;
  if (c->dom.n_ranks > 1) {
    if (comm_up(c)) { hcg_status s = comm_allreduce_min_host(c, &ok); if (s) return s; }
    else ok = 0;                                          // decided in hcg_comm_init
  }
  c->mo_ok = ok != 0;
  return HCG_OK;
}

hcg_status do_mechanics(hcg_ctx* c, bool forced, bool components) {
  OpTimer t(c, "applyConstitutiveModel");
  for (size_t k = 0; k < c->types.size(); k++) {
    if (forced || c->iter % c->types[k].timescale == 0) {      // hemoCellParticleField.cpp:655
      hcg_status s = mech_apply(c, (int)k, components); if (s) return s;
    }
  }
  return HCG_OK;
}

// spreadParticleForce: node-sorted pair path when every cell type supports it, else plain atomics
hcg_status do_spread(hcg_ctx* c) {
  bool sorted = c->spread_mode == 1;
  for (auto& t : c->types) if (t.n_cells > 0 && !spread_sorted_supported(t)) sorted = false;
  if (!sorted) return ibm_spread(c);
  if (!c->perm_valid || c->iter % c->perm_every == 0) {
    hcg_status s = spread_sorted_rebuild(c); if (s) return s;
    c->perm_valid = true;
    // lattices with walls: which cells cannot meet a non-fluid node until the next rebuild (iterate() only: the cadence is
    // what bounds the drift; the per-operator entry points leave every cell checked)
    if (c->in_iterate && (s = ibm_far_classify(c, c->perm_every))) return s;
  }
  return spread_sorted(c);
}

// one HemoCell::iterate() (core/hemoCell.cpp:299-376)
hcg_status step(hcg_ctx* c) {
  hcg_status s;
  // multi-GPU: every rank runs the same operator sequence (the exchanges are collective), cells or not
  const bool have_p = c->np > 0 || (c->dom.n_ranks > 1 && !c->types.empty());
  if (have_p && c->rep_on && c->iter % c->ts_rep == 0) { OpTimer t(c, "applyRepulsionForce"); if ((s = rep_apply(c))) return s; }
  if (have_p && c->wall_on && c->iter % c->ts_wall == 0) { OpTimer t(c, "applyBoundaryRepulsionForce"); if ((s = rep_wall_apply(c))) return s; }
  if (have_p) { OpTimer t(c, "spreadParticleForce"); if ((s = do_spread(c))) return s; }
  const bool interp = have_p && (c->iter % c->ts_vel == 0);
  bool fused = false;
  if (lat_moment_eligible(c)) {
    // tau = 1, fully periodic, no walls, opt-in: the moments are the state, one kernel replaces collision + moments pass
    OpTimer t(c, "collideAndStream"); if ((s = lat_moment_step(c, interp))) return s;
    fused = true;
  }
  if (interp && !fused) { OpTimer t(c, "collideAndStream+moments"); if ((s = lat_collide_moments_overlapped(c, &fused))) return s; }
  if (!fused) { OpTimer t(c, "collideAndStream"); if ((s = lat_collide_stream(c, !interp))) return s; }
  if (interp) {
    // interpolation and advance share one pass over the particles (advance uses the velocity just interpolated)
    if (c->dom.n_ranks == 1) {
      OpTimer t(c, "interpolateFluidVelocity"); if (!fused && (s = lat_moments(c, true, false))) return s; if ((s = ibm_interpolate_advance(c))) return s;
    } else {
      // cells no neighbour holds move in the interpolation pass; the shared ones after the velocity sync
      static int overlap = -1;
      if (overlap < 0) { const char* e = getenv("HCG_SYNC_OVERLAP"); overlap = e ? atoi(e) : 0; }   // opt-in: measured no faster (2 B200: 2.305 vs 2.291 ms), the exchange chain is work, not idle time
      if (!overlap) {
        { OpTimer t(c, "interpolateFluidVelocity"); if (!fused && (s = lat_moments(c, true, false))) return s; if ((s = ibm_interpolate_advance_unshared(c))) return s; }
        { OpTimer t(c, "syncEnvelopes"); if ((s = multi_velocity_sync_advance(c))) return s; }
      } else {
        // the velocity exchange of the shared cells (pack -> neighbour -> unpack -> advance) runs on the main stream
        // while the cells no neighbour holds are interpolated and advanced on a low-priority stream
        if (!c->stream_lo) {
          int lo = 0, hi = 0;
          CUDA_TRY(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));
          CUDA_TRY(c, cudaStreamCreateWithPriority(&c->stream_lo, cudaStreamNonBlocking, lo));
        }
        if (!c->ev_fork) { CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming)); CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming)); }
        { OpTimer t(c, "interpolateFluidVelocity"); if (!fused && (s = lat_moments(c, true, false))) return s; }
        CUDA_TRY(c, cudaEventRecord(c->ev_fork, c->stream));
        CUDA_TRY(c, cudaStreamWaitEvent(c->stream_lo, c->ev_fork, 0));
        { OpTimer t(c, "syncEnvelopes");
          if ((s = ibm_interpolate_shared(c))) return s;
          if ((s = ibm_interpolate_advance_unshared_on(c, c->stream_lo))) return s;
          CUDA_TRY(c, cudaEventRecord(c->ev_join, c->stream_lo));
          if ((s = multi_velocity_sync(c))) return s; }
        { OpTimer t(c, "advanceParticles"); if ((s = ibm_advance_shared(c))) return s;
          CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_join, 0)); }
      }
    }
  } else if (have_p) {
    OpTimer t(c, "advanceParticles"); if ((s = ibm_advance(c))) return s;
  }
  if (have_p && (s = do_mechanics(c, false, false))) return s;
  c->f_clean = true;            // the collision (or moments) kernel of this step wrote the reset value on every node
  c->iter++;
  if (c->dom.n_ranks > 1 && c->iter >= c->multi.next_sync_iter) {
    OpTimer t(c, "syncEnvelopes (membership + migration)"); if ((s = multi_rebalance(c, false))) return s;
  }
  return HCG_OK;
}

}  // namespace

extern "C" {

const char* hcg_last_error(const hcg_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }
const char* hcg_version(void) { return "hemocell_b200 0.1 (sm_100a)"; }

int32_t hcg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

void hcg_slab(int32_t nx, int32_t rank, int32_t n_ranks, int32_t* x0_out, int32_t* nxl_out) {
  const int32_t base = nx / n_ranks, rem = nx % n_ranks;
  if (x0_out) *x0_out = rank*base + (rank < rem ? rank : rem);
  if (nxl_out) *nxl_out = base + (rank < rem ? 1 : 0);
}

hcg_status hcg_create(const hcg_domain* d, hcg_ctx** out) {
  if (!d || !out) return hcg_fail(nullptr, HCG_ERR_ARG, "null argument");
  if (d->nx < 1 || d->ny < 3 || d->nz < 3) return hcg_fail(nullptr, HCG_ERR_ARG, "lattice too small");
  if (d->n_ranks < 1 || d->rank < 0 || d->rank >= d->n_ranks || d->nx < d->n_ranks)
    return hcg_fail(nullptr, HCG_ERR_ARG, "bad rank/n_ranks");
  if (!(d->tau > 0.5)) return hcg_fail(nullptr, HCG_ERR_ARG, "tau must be > 0.5");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return hcg_fail(nullptr, HCG_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
  if (d->device < 0 || d->device >= ndev) return hcg_fail(nullptr, HCG_ERR_ARG, "bad device ordinal");
  hcg_ctx* c = new hcg_ctx();
  c->dom = *d;
  { int32_t x0, nxl; hcg_slab(d->nx, d->rank, d->n_ranks, &x0, &nxl); c->x0 = x0; c->nxl = nxl; }
  if (d->n_ranks > 1 && c->nxl < 3) { delete c; return hcg_fail(nullptr, HCG_ERR_ARG, "slab thinner than 3 planes"); }
  c->P = (int64_t)d->ny * d->nz; c->S = (int64_t)(c->nxl + 2) * c->P; c->Nl = (int64_t)c->nxl * c->P;
  c->omega = 1.0 / d->tau;
  memset(c->bc_vel, 0, sizeof(c->bc_vel)); memset(c->body, 0, sizeof(c->body));
  c->f_limit = 1e300;
  c->cur = 0; c->u_valid = false; c->has_velbc = false; c->has_nonfluid = d->n_ranks > 1; c->rho = nullptr;
  c->mo_ok = d->n_ranks == 1 && c->omega == 1.0 && d->periodic[0] && d->periodic[1] && d->periodic[2];
  c->np = c->ncells = c->cap_p = c->cap_c = 0;
  for (int k = 0; k < 3; k++) c->pos[k] = c->vel[k] = c->frc[k] = c->frep[k] = nullptr;
  memset(c->comp, 0, sizeof(c->comp)); c->comp_alloc = false;
  c->p_cell = nullptr; c->cell_alive = nullptr; c->cell_type = nullptr; c->cell_base = nullptr;
  c->rep_on = c->wall_on = false; c->rep_k = c->rep_cut = c->wall_k = c->wall_cut = 0.0;
  c->ts_vel = c->ts_rep = c->ts_wall = 1;
  c->bin_count = c->bin_start = c->bin_items = nullptr; c->wall_nodes = nullptr; c->n_wall = 0; c->wall_built = false;
  c->scan_tmp = nullptr; c->scan_tmp_bytes = 0;
  c->iter = 0; c->spread_mode = 1; c->perm_valid = false; c->perm_every = 20; c->nccl = nullptr; c->timers_on = false; c->launches = 0;
  c->staging = nullptr; c->staging_bytes = 0;
  c->halo_send[0] = c->halo_send[1] = c->halo_recv[0] = c->halo_recv[1] = nullptr;
  *out = c;
  CUDA_TRY(c, cudaSetDevice(d->device));
  CUDA_TRY(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CUDA_TRY(c, cudaStreamCreateWithFlags(&c->stream_halo, cudaStreamNonBlocking));
  CUDA_TRY(c, cudaEventCreate(&c->ev_a)); CUDA_TRY(c, cudaEventCreate(&c->ev_b));
  CUDA_TRY(c, cudaMalloc(&c->g[0], sizeof(double)*19*c->S));
  CUDA_TRY(c, cudaMalloc(&c->g[1], sizeof(double)*19*c->S));
  CUDA_TRY(c, cudaMalloc(&c->F, sizeof(double)*4*c->S));
  CUDA_TRY(c, cudaMalloc(&c->U, sizeof(double)*4*c->S));
  CUDA_TRY(c, cudaMalloc(&c->flags, c->S));
  CUDA_TRY(c, cudaMalloc(&c->d_bc, sizeof(double)*18));
  CUDA_TRY(c, cudaMemsetAsync(c->d_bc, 0, sizeof(double)*18, c->stream));
  CUDA_TRY(c, cudaMemsetAsync(c->g[0], 0, sizeof(double)*19*c->S, c->stream));
  CUDA_TRY(c, cudaMemsetAsync(c->g[1], 0, sizeof(double)*19*c->S, c->stream));
  CUDA_TRY(c, cudaMemsetAsync(c->F, 0, sizeof(double)*4*c->S, c->stream));
  CUDA_TRY(c, cudaMemsetAsync(c->U, 0, sizeof(double)*4*c->S, c->stream));
  CUDA_TRY(c, cudaMemsetAsync(c->flags, HCG_FLUID, c->S, c->stream));
  hcg_status s = exchange_flags(c); if (s) return s;
  // tau = 1 on a fully periodic box: the buffers of the moment-only update exist from the start (the slab neighbours map them in hcg_comm_init)
  if (c->omega == 1.0 && d->periodic[0] && d->periodic[1] && d->periodic[2] && (s = lat_moment_buffers(c))) return s;
  const double u0[3] = {0, 0, 0};
  if (d->n_ranks == 1) { s = lat_init_equilibrium(c, 1.0, u0); if (s) return s; }
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return HCG_OK;
}

void hcg_destroy(hcg_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->dom.device);
  cudaDeviceSynchronize();
  preinlet_destroy(c);
  peer_destroy(c);
  comm_destroy(c);
  cudaFree(c->g[0]); cudaFree(c->g[1]); cudaFree(c->F); cudaFree(c->U); if (c->W) cudaFree(c->W); if (c->F0) cudaFree(c->F0); if (c->bcn) cudaFree(c->bcn); if (c->W2) cudaFree(c->W2); if (c->F2) cudaFree(c->F2); if (c->d_qsets) cudaFree(c->d_qsets); cudaFree(c->flags); cudaFree(c->d_bc);
  if (c->rho) cudaFree(c->rho);
  if (c->count_dev) cudaFree(c->count_dev); if (c->count_typeV) cudaFree(c->count_typeV);
  if (c->wall_coarse) cudaFree(c->wall_coarse); if (c->cell_far) cudaFree(c->cell_far); if (c->far_typeV) cudaFree(c->far_typeV); if (c->bbox_typeV) cudaFree(c->bbox_typeV);
  if (c->fused_done) cudaFree(c->fused_done);
  for (int k = 0; k < 3; k++) { cudaFree(c->pos[k]); cudaFree(c->vel[k]); cudaFree(c->frc[k]); cudaFree(c->frep[k]); }
  for (int k = 0; k < 6; k++) for (int d = 0; d < 3; d++) if (c->comp[k][d]) cudaFree(c->comp[k][d]);
  if (c->multi.d_cell_shared) cudaFree(c->multi.d_cell_shared);
  if (c->multi.d_meta) cudaFree(c->multi.d_meta); if (c->multi.d_bbox) cudaFree(c->multi.d_bbox); if (c->multi.d_vmax) cudaFree(c->multi.d_vmax);
  for (int f = 0; f < 2; f++) { cudaFree(c->multi.tmp_send[f].d_cells); cudaFree(c->multi.tmp_send[f].d_off); cudaFree(c->multi.tmp_recv[f].d_cells); cudaFree(c->multi.tmp_recv[f].d_off); }
  if (c->cell_gid) cudaFree(c->cell_gid);
  cudaFree(c->p_cell); cudaFree(c->cell_alive); cudaFree(c->cell_type); cudaFree(c->cell_base);
  cudaFree(c->bin_count); cudaFree(c->bin_start); cudaFree(c->bin_items); cudaFree(c->wall_nodes); cudaFree(c->scan_tmp);
  cudaFree(c->staging);
  for (auto& t : c->types) { for (void* p : t.allocs) cudaFree(p); if (t.perm) cudaFree(t.perm); }
  for (auto& p : c->ev_pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
  for (auto e : c->ev_pool) cudaEventDestroy(e);
  cudaEventDestroy(c->ev_a); cudaEventDestroy(c->ev_b);
  cudaStreamDestroy(c->stream); cudaStreamDestroy(c->stream_halo); if (c->stream_lo) cudaStreamDestroy(c->stream_lo);
  delete c;
}

hcg_status hcg_comm_unique_id(void* out128) {
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess) return hcg_fail(nullptr, HCG_ERR_NCCL, "ncclGetUniqueId failed");
  static_assert(sizeof(ncclUniqueId) == 128, "unique id size");
  memcpy(out128, &id, 128);
  return HCG_OK;
}

static hcg_status comm_finish_init(hcg_ctx* c) {
  if (const char* e = getenv("HCG_TRANSPORT")) c->peer.transport = (strcmp(e, "nccl") == 0) ? 0 : 1;
  hcg_status s = peer_setup(c); if (s) return s;
  if (c->peer.transport == 1) {
    // every rank must use the same transport: fall back to send/recv everywhere if any rank cannot map its neighbours
    int all = c->peer.ready ? 1 : 0;
    if ((s = comm_allreduce_min_host(c, &all))) return s;
    if (!all) {
      if (c->dom.rank == 0) fprintf(stderr, "(hemocell_gpu) peer-memory transport unavailable on this box: using NCCL send/recv\n");
      c->peer.ready = false; c->peer.transport = 0;
    }
  }
  s = exchange_flags(c); if (s) return s;
  if ((s = refresh_nonfluid(c))) return s;
  const double u0[3] = {0, 0, 0};
  s = lat_init_equilibrium(c, 1.0, u0); if (s) return s;
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return HCG_OK;
}

hcg_status hcg_comm_init(hcg_ctx* c, const void* id128) {
  if (!c || !id128) return HCG_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  if (c->dom.n_ranks == 1) return HCG_OK;
  if (comm_up(c)) return hcg_fail(c, HCG_ERR_STATE, "communicator already initialised");
  hcg_status s = comm_nccl_init(c, id128); if (s) return s;
  return comm_finish_init(c);
}

hcg_status hcg_comm_init_local(hcg_ctx* c, const void* id128) {
  if (!c || !id128) return HCG_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  if (c->dom.n_ranks == 1) return HCG_OK;
  if (comm_up(c)) return hcg_fail(c, HCG_ERR_STATE, "communicator already initialised");
  hcg_status s = comm_local_init(c, id128); if (s) return s;
  return comm_finish_init(c);
}

hcg_status hcg_lattice_set_flags(hcg_ctx* c, const uint8_t* flags) {
  if (!c || !flags) return HCG_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  { hcg_status sp = lat_ensure_pops(c); if (sp) return sp; }    // (moment-only mode: materialise the populations under the old flags)
  bool vel = false, io = false;
  for (int64_t i = 0; i < c->Nl; i++) {
    if (flags[i] > HCG_ZH_PRES_ZP) return hcg_fail(c, HCG_ERR_ARG, "unknown node flag");
    vel = vel || flags[i] >= HCG_VEL_XN;
    io = io || flags[i] >= HCG_ZH_VEL_XN;
  }
  c->has_velbc = vel; c->has_iobc = io;
  if (io) { hcg_status sb = lat_bcn_ensure(c); if (sb) return sb; }
  bool nonfluid = false;
  for (int64_t i = 0; i < c->Nl && !nonfluid; i++) nonfluid = flags[i] != HCG_FLUID;
  c->real_nonfluid = nonfluid; c->has_nonfluid = true;
  hcg_status s = ensure_staging(c, c->Nl); if (s) return s;
  CUDA_TRY(c, cudaMemcpyAsync(c->staging, flags, c->Nl, cudaMemcpyHostToDevice, c->stream));
  k_pad_flags<<<nblk(c->Nl, 256), 256, 0, c->stream>>>((const uint8_t*)c->staging, c->flags, c->Nl, c->P);
  KERNEL_CHECK(c);
  s = exchange_flags(c); if (s) return s;
  c->wall_built = false; c->wall_coarse_valid = false; c->far_steps_left = 0;
  return refresh_nonfluid(c);                   // (collective for n_ranks > 1: every rank sets its flags)
}

hcg_status hcg_lattice_set_bc_velocity(hcg_ctx* c, int32_t o, const double u[3]) {
  if (!c || !u || o < 0 || o > 5) return HCG_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  for (int k = 0; k < 3; k++) c->bc_vel[o][k] = u[k];
  CUDA_TRY(c, hcg_h2d(c, c->d_bc, c->bc_vel, sizeof(c->bc_vel)));
  return HCG_OK;
}

hcg_status hcg_lattice_set_bc_nodes(hcg_ctx* c, int64_t n, const int64_t* node_idx, const double* val) {
  if (!c || n < 0 || (n > 0 && (!node_idx || !val))) return HCG_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  for (int64_t k = 0; k < n; k++) if (node_idx[k] < 0 || node_idx[k] >= c->Nl) return hcg_fail(c, HCG_ERR_ARG, "node index outside this rank's slab");
  hcg_status s = lat_bcn_ensure(c); if (s) return s;
  if (n == 0) return HCG_OK;
  if ((s = ensure_staging(c, (size_t)n*(sizeof(int64_t) + 4*sizeof(double))))) return s;
  int64_t* d_idx = (int64_t*)c->staging; double* d_val = (double*)((char*)c->staging + (size_t)n*sizeof(int64_t));
  CUDA_TRY(c, cudaMemcpyAsync(d_idx, node_idx, (size_t)n*sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(c, cudaMemcpyAsync(d_val, val, (size_t)n*4*sizeof(double), cudaMemcpyHostToDevice, c->stream));
  if ((s = lat_bcn_scatter(c, n, d_idx, d_val, false, c->stream))) return s;
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return HCG_OK;
}

hcg_status hcg_lattice_node_velocity(hcg_ctx* c, int64_t n, const int64_t* node_idx, double* u_out) {
  if (!c || n < 0 || (n > 0 && (!node_idx || !u_out))) return HCG_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  for (int64_t k = 0; k < n; k++) if (node_idx[k] < 0 || node_idx[k] >= c->Nl) return hcg_fail(c, HCG_ERR_ARG, "node index outside this rank's slab");
  if (n == 0) return HCG_OK;
  hcg_status s = ensure_staging(c, (size_t)n*(sizeof(int64_t) + 4*sizeof(double))); if (s) return s;
  int64_t* d_idx = (int64_t*)c->staging; double* d_val = (double*)((char*)c->staging + (size_t)n*sizeof(int64_t));
  CUDA_TRY(c, cudaMemcpyAsync(d_idx, node_idx, (size_t)n*sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
  if ((s = lat_node_velocity(c, n, d_idx, d_val, c->stream))) return s;
  std::vector<double> h((size_t)4*n);
  CUDA_TRY(c, cudaMemcpyAsync(h.data(), d_val, (size_t)n*4*sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  for (int64_t k = 0; k < n; k++) for (int d = 0; d < 3; d++) u_out[3*k + d] = h[4*k + d];
  return HCG_OK;
}

hcg_status hcg_lattice_init_equilibrium(hcg_ctx* c, double rho, const double u[3]) {
  if (!c || !u) return HCG_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  hcg_status s = lat_init_equilibrium(c, rho, u); if (s) return s;
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return HCG_OK;
}

hcg_status hcg_lattice_set_body_force(hcg_ctx* c, const double f[3]) {
  if (!c || !f) return HCG_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  // the case files re-apply the same driving force after every iterate() (examples/pipeflow/pipeflow.cpp:144-146):
  // iterate() has already reset the node force to it
  if (c->f_clean && !c->F0 && f[0] == c->body[0] && f[1] == c->body[1] && f[2] == c->body[2]) return HCG_OK;
  for (int k = 0; k < 3; k++) c->body[k] = f[k];
  if (c->F0) { CUDA_TRY(c, cudaStreamSynchronize(c->stream)); cudaFree(c->F0); c->F0 = nullptr; }   // back to a uniform force
  hcg_status s = lat_reset_force(c); if (s) return s;
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return HCG_OK;
}

hcg_status hcg_lattice_set_body_force_field(hcg_ctx* c, const double* f) {
  if (!c || !f) return HCG_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  hcg_status s;
  if (!c->F0) {
    CUDA_TRY(c, cudaMalloc(&c->F0, sizeof(double)*4*c->S));
    CUDA_TRY(c, cudaMemsetAsync(c->F0, 0, sizeof(double)*4*c->S, c->stream));
  }
  if ((s = ensure_staging(c, sizeof(double)*3*c->Nl))) return s;
  CUDA_TRY(c, cudaMemcpyAsync(c->staging, f, sizeof(double)*3*c->Nl, cudaMemcpyHostToDevice, c->stream));
  if ((s = lat_pad3(c, c->staging, c->F0))) return s;
  if ((s = lat_reset_force(c))) return s;
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return HCG_OK;
}

hcg_status hcg_lattice_upload(hcg_ctx* c, int32_t field, const double* in) {
  if (!c || !in) return HCG_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  hcg_status s;
  if (field == HCG_LAT_POP) {
    if ((s = ensure_staging(c, sizeof(double)*19*c->Nl))) return s;
    CUDA_TRY(c, cudaMemcpyAsync(c->staging, in, sizeof(double)*19*c->Nl, cudaMemcpyHostToDevice, c->stream));
    if ((s = lat_pop_from_reference(c, c->staging))) return s;
  } else if (field == HCG_LAT_FORCE) {
    if ((s = ensure_staging(c, sizeof(double)*3*c->Nl))) return s;
    CUDA_TRY(c, cudaMemcpyAsync(c->staging, in, sizeof(double)*3*c->Nl, cudaMemcpyHostToDevice, c->stream));
    if ((s = lat_pad3(c, c->staging, c->F))) return s;
    c->f_clean = false;
  } else return hcg_fail(c, HCG_ERR_ARG, "lattice_upload: field must be POP or FORCE");
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return HCG_OK;
}

hcg_status hcg_lattice_download(hcg_ctx* c, int32_t field, double* out) {
  if (!c || !out) return HCG_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  hcg_status s;
  if (field == HCG_LAT_POP) {
    if ((s = ensure_staging(c, sizeof(double)*19*c->Nl))) return s;
    if ((s = lat_pop_to_reference(c, c->staging))) return s;
    CUDA_TRY(c, cudaMemcpyAsync(out, c->staging, sizeof(double)*19*c->Nl, cudaMemcpyDeviceToHost, c->stream));
  } else if (field == HCG_LAT_FORCE) {
    if ((s = ensure_staging(c, sizeof(double)*3*c->Nl))) return s;
    if ((s = lat_unpad(c, c->F, c->staging, 0, 3))) return s;
    CUDA_TRY(c, cudaMemcpyAsync(out, c->staging, sizeof(double)*3*c->Nl, cudaMemcpyDeviceToHost, c->stream));
  } else if (field == HCG_LAT_VELOCITY || field == HCG_LAT_DENSITY) {
    // Cell::computeVelocity / computeDensity of the current populations and node force
    if ((s = lat_moments(c, false, true))) return s;
    if ((s = ensure_staging(c, sizeof(double)*3*c->Nl))) return s;
    if (field == HCG_LAT_VELOCITY) {
      if ((s = lat_unpad(c, c->U, c->staging, 0, 3))) return s;
      CUDA_TRY(c, cudaMemcpyAsync(out, c->staging, sizeof(double)*3*c->Nl, cudaMemcpyDeviceToHost, c->stream));
    } else {
      if ((s = lat_unpad(c, c->U, c->staging, 3, 1))) return s;
      CUDA_TRY(c, cudaMemcpyAsync(out, c->staging, sizeof(double)*c->Nl, cudaMemcpyDeviceToHost, c->stream));
    }
  } else if (field == HCG_LAT_PINEQ) {
    if ((s = ensure_staging(c, sizeof(double)*6*c->Nl))) return s;
    if ((s = lat_pineq(c, c->staging))) return s;
    CUDA_TRY(c, cudaMemcpyAsync(out, c->staging, sizeof(double)*6*c->Nl, cudaMemcpyDeviceToHost, c->stream));
  } else return hcg_fail(c, HCG_ERR_ARG, "lattice_download: unknown field");
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return HCG_OK;
}

hcg_status hcg_celltype_add(hcg_ctx* c, const hcg_celltype* t, int32_t* ctype_out) {
  if (!c || !t) return HCG_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  if (c->types.size() >= HCG_MAX_TYPES) return hcg_fail(c, HCG_ERR_CAPACITY, "too many cell types");
  if (t->model != HCG_MODEL_RBC_HIGHORDER && t->model != HCG_MODEL_PLT_SIMPLE && t->model != HCG_MODEL_HOST) return hcg_fail(c, HCG_ERR_ARG, "unknown model");
  const int V = t->n_vertices, T = t->n_triangles, E = t->n_edges, I = t->n_inner_edges;
  if (V < 4 || T < 4 || E < 6 || I < 0) return hcg_fail(c, HCG_ERR_ARG, "degenerate mesh");
  // ---- per-vertex gather tables
  std::vector<int> vt(6*(size_t)V, -1), vpe(12*(size_t)V, -1), vin(4*(size_t)V, -1);
  std::vector<int> nvt(V, 0), nvpe(V, 0), nvin(V, 0);
  for (int k = 0; k < T; k++) for (int m = 0; m < 3; m++) {
    const int v = t->triangles[3*k+m];
    if (v < 0 || v >= V) return hcg_fail(c, HCG_ERR_ARG, "triangle index out of range");
    if (nvt[v] >= 6) return hcg_fail(c, HCG_ERR_ARG, "vertex with more than 6 triangles");
    vt[6*v + nvt[v]++] = k;
  }
  for (int e = 0; e < E; e++) for (int m = 0; m < 2; m++) {
    const int v = t->edges[2*e+m];
    if (v < 0 || v >= V) return hcg_fail(c, HCG_ERR_ARG, "edge index out of range");
  }
  for (int v = 0; v < V; v++) {
    const int nn = t->vertex_n_vertexes[v];
    if (nn < 3 || nn > 6) return hcg_fail(c, HCG_ERR_ARG, "vertex ring size must be 3..6");
    for (int j = 0; j < nn; j++) {
      const int r = t->vertex_vertexes[6*v + j];
      if (r < 0 || r >= V || r == v) return hcg_fail(c, HCG_ERR_ARG, "ring vertex out of range");
    }
  }
  // RBC: the packed ring table of the mechanics kernel (mech_tables.h)
  std::vector<unsigned long long> rg(6*(size_t)V, 0ull);
  if (t->model == HCG_MODEL_RBC_HIGHORDER) {
    if (const char* why = mech_tables::build_ring_table(V, T, E, t->triangles, t->edges, t->vertex_vertexes, t->vertex_n_vertexes, rg))
      return hcg_fail(c, HCG_ERR_ARG, why);
  }
  if (t->model == HCG_MODEL_PLT_SIMPLE) {
    for (int e = 0; e < E; e++) {
      const int role_v[4] = {t->edges[2*e], t->edges[2*e+1], t->edge_bending_outer_points[2*e], t->edge_bending_outer_points[2*e+1]};
      for (int r = 0; r < 4; r++) {
        const int v = role_v[r];
        if (v < 0 || v >= V) return hcg_fail(c, HCG_ERR_ARG, "bending point out of range");
        if (nvpe[v] >= 12) return hcg_fail(c, HCG_ERR_ARG, "vertex touches more than 12 bending edges");
        vpe[12*v + nvpe[v]++] = 4*e + r;
      }
    }
    for (int e = 0; e < I; e++) for (int m = 0; m < 2; m++) {
      const int v = t->inner_edges[2*e+m];
      if (v < 0 || v >= V) return hcg_fail(c, HCG_ERR_ARG, "inner edge index out of range");
      if (nvin[v] >= 4) return hcg_fail(c, HCG_ERR_ARG, "vertex with more than 4 inner edges");
      vin[4*v + nvin[v]++] = 2*e + m;
    }
  }
  CellTypeHost th;
  CellTypeDev& d = th.d;
  memset(&d, 0, sizeof(d));
  d.model = t->model; d.V = V; d.T = T; d.E = E; d.I = I;
  hcg_status s;
#define UP(dst, src, n) if ((s = upload_table(c, th, &d.dst, src, (size_t)(n)))) return s
  UP(tri, t->triangles, 3*T); UP(edge, t->edges, 2*E); UP(inner, t->inner_edges, 2*I);
  UP(ring, t->vertex_vertexes, 6*V); UP(nring, t->vertex_n_vertexes, V);
  UP(bend_tri, t->edge_bending_triangles, 2*E); UP(bend_outer, t->edge_bending_outer_points, 2*E);
  UP(edge_len_eq, t->edge_length_eq, E); UP(edge_ang_eq, t->edge_angle_eq, E);
  UP(tri_area_eq, t->triangle_area_eq, T); UP(patch_eq, t->patch_dist_eq, V); UP(inner_len_eq, t->inner_edge_length_eq, I);
  UP(vt, vt.data(), vt.size()); UP(rg, rg.data(), rg.size());
  UP(vpe, vpe.data(), vpe.size()); UP(vin, vin.data(), vin.size());
#undef UP
  d.volume_eq = t->volume_eq; d.area_mean_eq = t->area_mean_eq; d.edge_mean_eq = t->edge_mean_eq;
  d.k_volume = t->k_volume; d.k_area = t->k_area; d.k_link = t->k_link; d.k_bend = t->k_bend; d.eta_m = t->eta_m;
  th.first_cell = c->ncells; th.first_particle = c->np;
  c->types.push_back(th);
  if (ctype_out) *ctype_out = (int32_t)c->types.size() - 1;
  return HCG_OK;
}

hcg_status hcg_celltype_set_stiffness(hcg_ctx* c, int32_t ctype, double kv, double ka, double kl, double kb, double eta) {
  if (!c || ctype < 0 || ctype >= (int)c->types.size()) return HCG_ERR_ARG;
  CellTypeDev& d = c->types[ctype].d;
  d.k_volume = kv; d.k_area = ka; d.k_link = kl; d.k_bend = kb; d.eta_m = eta;
  return HCG_OK;
}

hcg_status hcg_cells_add(hcg_ctx* c, int32_t ctype, int64_t n_cells, const int64_t* cell_id, const double* pos) {
  if (!c || ctype < 0 || ctype >= (int)c->types.size() || n_cells < 0) return HCG_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  for (size_t k = ctype + 1; k < c->types.size(); k++)
    if (c->types[k].n_cells > 0) return hcg_fail(c, HCG_ERR_STATE, "cells must be added in cell-type order");
  CellTypeHost& th = c->types[ctype];
  if (th.n_cells > 0) return hcg_fail(c, HCG_ERR_STATE, "cells of a type must be added in one call");
  if (n_cells > 0 && (!cell_id || !pos)) return HCG_ERR_ARG;
  const int V = th.d.V;
  // multi-GPU: keep the cells whose x-extent intersects this rank's hold region (the caller may
  // pass any superset, e.g. the global list) and reserve spare slots for later arrivals
  std::vector<int64_t> keep_id; std::vector<double> keep_pos;
  int64_t n_slots = n_cells;
  if (c->dom.n_ranks > 1) {
    std::vector<double> lo(n_cells), hi(n_cells);
    for (int64_t i = 0; i < n_cells; i++) {
      double a = pos[3*i*V], b = a;
      for (int v = 1; v < V; v++) { const double x = pos[3*(i*V + v)]; a = x < a ? x : a; b = x > b ? x : b; }
      lo[i] = a; hi[i] = b;
    }
    std::vector<uint8_t> held(n_cells), sl(n_cells), sr(n_cells);
    hch_slab_membership_at(n_cells, lo.data(), hi.data(), c->dom.nx, c->dom.periodic[0], c->x0, c->nxl, c->dom.rank,
                        c->dom.n_ranks, c->multi.margin, held.data(), sl.data(), sr.data());
    for (int64_t i = 0; i < n_cells; i++) if (held[i]) {
      keep_id.push_back(cell_id[i]);
      keep_pos.insert(keep_pos.end(), pos + 3*i*V, pos + 3*(i + 1)*V);
    }
    n_cells = (int64_t)keep_id.size();
    cell_id = keep_id.data(); pos = keep_pos.data();
    n_slots = n_cells + (int64_t)(c->multi.slack*n_cells) + 64;
  }
  n_slots += th.reserve;
  if (n_slots == 0) return HCG_OK;
  const int64_t add_p = n_slots*V, new_p = c->np + add_p, new_c = c->ncells + n_slots;
  if (new_p > 2000000000LL) return hcg_fail(c, HCG_ERR_CAPACITY, "more than 2e9 particles on one GPU");
  // grow SoA arrays
  auto grow = [&](double** arr) -> cudaError_t {
    double* n; cudaError_t e = cudaMalloc(&n, sizeof(double)*new_p); if (e) return e;
    e = cudaMemsetAsync(n, 0, sizeof(double)*new_p, c->stream); if (e) return e;
    if (*arr && c->np) { e = cudaMemcpyAsync(n, *arr, sizeof(double)*c->np, cudaMemcpyDeviceToDevice, c->stream); if (e) return e; }
    e = cudaStreamSynchronize(c->stream); if (e) return e;
    if (*arr) cudaFree(*arr);
    *arr = n; return cudaSuccess;
  };
  for (int k = 0; k < 3; k++) { CUDA_TRY(c, grow(&c->pos[k])); CUDA_TRY(c, grow(&c->vel[k])); CUDA_TRY(c, grow(&c->frc[k])); CUDA_TRY(c, grow(&c->frep[k])); }
  if (c->comp_alloc) { for (int k = 0; k < 6; k++) for (int d = 0; d < 3; d++) { cudaFree(c->comp[k][d]); c->comp[k][d] = nullptr; } c->comp_alloc = false; }
  th.first_cell = c->ncells; th.first_particle = c->np; th.n_cells = n_cells; th.cap_cells = n_slots;
  for (size_t k = ctype + 1; k < c->types.size(); k++) { c->types[k].first_cell = new_c; c->types[k].first_particle = new_p; }
  // host-side cell tables (slots beyond n_cells are spare: id -1, not alive)
  std::vector<int32_t> pc(add_p);
  std::vector<uint8_t> alive_new(n_slots, 0);
  c->multi.free_slots.resize(c->types.size());
  for (int64_t i = 0; i < n_slots; i++) {
    c->h_cell_id.push_back(i < n_cells ? cell_id[i] : -1); c->h_cell_type.push_back(ctype); c->h_cell_base.push_back(c->np + i*V);
    c->multi.h_held.push_back(i < n_cells); c->multi.h_shared[0].push_back(0); c->multi.h_shared[1].push_back(0);
    alive_new[i] = i < n_cells;
    for (int v = 0; v < V; v++) pc[i*V + v] = (int32_t)(c->ncells + i);
  }
  // device cell tables (rebuilt)
  std::vector<uint8_t> alive_all(new_c, 0);
  if (c->ncells) CUDA_TRY(c, cudaMemcpy(alive_all.data(), c->cell_alive, c->ncells, cudaMemcpyDeviceToHost));
  std::copy(alive_new.begin(), alive_new.end(), alive_all.begin() + c->ncells);
  cudaFree(c->cell_alive); cudaFree(c->cell_type); cudaFree(c->cell_base);
  int32_t* npc;
  CUDA_TRY(c, cudaMalloc(&npc, sizeof(int32_t)*new_p));
  if (c->p_cell && c->np) CUDA_TRY(c, cudaMemcpy(npc, c->p_cell, sizeof(int32_t)*c->np, cudaMemcpyDeviceToDevice));
  CUDA_TRY(c, hcg_h2d(c, npc + c->np, pc.data(), sizeof(int32_t)*add_p));
  cudaFree(c->p_cell); c->p_cell = npc;
  CUDA_TRY(c, cudaMalloc(&c->cell_alive, new_c));
  CUDA_TRY(c, hcg_h2d(c, c->cell_alive, alive_all.data(), new_c));
  CUDA_TRY(c, cudaMalloc(&c->cell_type, sizeof(int32_t)*new_c));
  CUDA_TRY(c, cudaMalloc(&c->cell_base, sizeof(int64_t)*new_c));
  CUDA_TRY(c, hcg_h2d(c, c->cell_type, c->h_cell_type.data(), sizeof(int32_t)*new_c));
  CUDA_TRY(c, hcg_h2d(c, c->cell_base, c->h_cell_base.data(), sizeof(int64_t)*new_c));
  // positions
  const int64_t live_p = n_cells*V;
  hcg_status s = ensure_staging(c, sizeof(double)*3*(live_p + 1)); if (s) return s;
  if (live_p) {
    CUDA_TRY(c, cudaMemcpyAsync(c->staging, pos, sizeof(double)*3*live_p, cudaMemcpyHostToDevice, c->stream));
    k_aos_to_soa<<<nblk(live_p, 256), 256, 0, c->stream>>>(c->staging, c->pos[0] + c->np, c->pos[1] + c->np, c->pos[2] + c->np, live_p);
    KERNEL_CHECK(c);
  }
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  c->np = new_p; c->ncells = new_c; c->cap_p = new_p; c->cap_c = new_c;
  if (c->multi.d_arr) { cudaFree(c->multi.d_arr); c->multi.d_arr = nullptr; }     // array pointers changed
  c->perm_valid = false; c->cell_gid_dirty = true; c->far_steps_left = 0;
  if (c->dom.n_ranks > 1 && (s = multi_rebalance(c, true))) return s;            // builds the shared lists
  if (c->bin_items) { cudaFree(c->bin_items); cudaFree(c->bin_count); cudaFree(c->bin_start); cudaFree(c->scan_tmp);
                      c->bin_items = c->bin_count = c->bin_start = nullptr; c->scan_tmp = nullptr; }
  return HCG_OK;
}

hcg_status hcg_cells_reserve(hcg_ctx* c, int32_t ctype, int64_t spare_cells) {
  if (!c || ctype < 0 || ctype >= (int)c->types.size() || spare_cells < 0) return HCG_ERR_ARG;
  if (c->types[ctype].cap_cells > 0) return hcg_fail(c, HCG_ERR_STATE, "hcg_cells_reserve must precede hcg_cells_add of the type");
  c->types[ctype].reserve = spare_cells;
  return HCG_OK;
}

hcg_status hcg_cells_capacity(hcg_ctx* c, int64_t* n_cells, int64_t* n_particles) {
  if (!c) return HCG_ERR_ARG;
  if (n_cells) *n_cells = c->ncells;
  if (n_particles) *n_particles = c->np;
  return HCG_OK;
}

// launches the count of alive cells (multi-GPU: cells this rank owns) and alive particles into c->count_dev (2 words)
static hcg_status count_launch(hcg_ctx* c) {
  if (!c->count_dev) CUDA_TRY(c, cudaMalloc(&c->count_dev, sizeof(unsigned long long)*2));
  if ((int)c->types.size() != c->count_ntypes) {
    std::vector<int> hv; for (auto& t : c->types) hv.push_back(t.d.V);
    if (c->count_typeV) { CUDA_TRY(c, cudaStreamSynchronize(c->stream)); cudaFree(c->count_typeV); c->count_typeV = nullptr; }
    CUDA_TRY(c, cudaMalloc(&c->count_typeV, sizeof(int)*std::max<size_t>(hv.size(), 1)));
    if (!hv.empty()) CUDA_TRY(c, hcg_h2d(c, c->count_typeV, hv.data(), sizeof(int)*hv.size()));
    c->count_ntypes = (int)c->types.size();
  }
  CUDA_TRY(c, cudaMemsetAsync(c->count_dev, 0, sizeof(unsigned long long)*2, c->stream));
  if (c->ncells > 0) {
    k_count_alive<<<nblk(c->ncells, 256), 256, 0, c->stream>>>(c->cell_alive, c->cell_type, c->count_typeV, c->ncells, c->count_dev,
        c->cell_base, c->pos[0], c->dom.n_ranks, c->dom.nx, c->dom.periodic[0], c->x0, c->nxl);
    KERNEL_CHECK(c);
  }
  return HCG_OK;
}

hcg_status hcg_cells_count(hcg_ctx* c, int64_t* n_cells_alive, int64_t* n_particles_alive) {
  if (!c) return HCG_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  unsigned long long h[2] = {0, 0};
  hcg_status s = count_launch(c); if (s) return s;
  CUDA_TRY(c, cudaMemcpyAsync(h, c->count_dev, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  if (n_cells_alive) *n_cells_alive = (int64_t)h[0];
  if (n_particles_alive) *n_particles_alive = (int64_t)h[1];
  return HCG_OK;
}

hcg_status hcg_cells_count_async(hcg_ctx* c, int64_t* out2) {
  if (!c || !out2) return HCG_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  hcg_status s = count_launch(c); if (s) return s;
  CUDA_TRY(c, cudaMemcpyAsync(out2, c->count_dev, sizeof(int64_t)*2, cudaMemcpyDeviceToHost, c->stream));
  return HCG_OK;
}

static hcg_status field_arrays(hcg_ctx* c, int32_t field, double** a) {
  switch (field) {
    case HCG_P_POS: a[0] = c->pos[0]; a[1] = c->pos[1]; a[2] = c->pos[2]; return HCG_OK;
    case HCG_P_VEL: a[0] = c->vel[0]; a[1] = c->vel[1]; a[2] = c->vel[2]; return HCG_OK;
    case HCG_P_FORCE: a[0] = c->frc[0]; a[1] = c->frc[1]; a[2] = c->frc[2]; return HCG_OK;
    case HCG_P_FREP: a[0] = c->frep[0]; a[1] = c->frep[1]; a[2] = c->frep[2]; return HCG_OK;
    default:
      if (field >= HCG_P_F_AREA && field <= HCG_P_F_INNER) {
        if (!c->comp_alloc) return hcg_fail(c, HCG_ERR_STATE, "component forces not computed: call hcg_op_mechanics(ctx, forced, 1) first");
        for (int d = 0; d < 3; d++) a[d] = c->comp[field - HCG_P_F_AREA][d];
        return HCG_OK;
      }
  }
  return hcg_fail(c, HCG_ERR_ARG, "unknown particle field");
}

hcg_status hcg_cells_upload(hcg_ctx* c, int32_t field, const double* in) {
  if (!c || !in) return HCG_ERR_ARG;
  if (field > HCG_P_FREP) return hcg_fail(c, HCG_ERR_ARG, "cells_upload: field must be POS, VEL, FORCE or FREP");
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  if (c->np == 0) return HCG_OK;
  double* a[3]; hcg_status s = field_arrays(c, field, a); if (s) return s;
  if ((s = ensure_staging(c, sizeof(double)*3*c->np))) return s;
  CUDA_TRY(c, cudaMemcpyAsync(c->staging, in, sizeof(double)*3*c->np, cudaMemcpyHostToDevice, c->stream));
  k_aos_to_soa<<<nblk(c->np, 256), 256, 0, c->stream>>>(c->staging, a[0], a[1], a[2], c->np);
  KERNEL_CHECK(c);
  if (field == HCG_P_POS) c->far_steps_left = 0;        // positions set from outside: every cell is checked against the walls again
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return HCG_OK;
}

hcg_status hcg_cells_download(hcg_ctx* c, int32_t field, double* out) {
  if (!c || !out) return HCG_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  if (c->np == 0) return HCG_OK;
  double* a[3]; hcg_status s = field_arrays(c, field, a); if (s) return s;
  if ((s = ensure_staging(c, sizeof(double)*3*c->np))) return s;
  k_soa_to_aos<<<nblk(c->np, 256), 256, 0, c->stream>>>(a[0], a[1], a[2], c->staging, c->np);
  KERNEL_CHECK(c);
  CUDA_TRY(c, cudaMemcpyAsync(out, c->staging, sizeof(double)*3*c->np, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return HCG_OK;
}

hcg_status hcg_cells_download_f32(hcg_ctx* c, int32_t field, double scale, float* out) {
  if (!c || !out) return HCG_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  if (c->np == 0) return HCG_OK;
  double* a[3]; hcg_status s = field_arrays(c, field, a); if (s) return s;
  if ((s = ensure_staging(c, sizeof(float)*3*c->np))) return s;
  k_soa_to_aos_f32<<<nblk(c->np, 256), 256, 0, c->stream>>>(a[0], a[1], a[2], (float*)c->staging, c->np, scale);
  KERNEL_CHECK(c);
  CUDA_TRY(c, cudaMemcpyAsync(out, c->staging, sizeof(float)*3*c->np, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return HCG_OK;
}

hcg_status hcg_cells_info(hcg_ctx* c, int64_t* cell_id_out, int32_t* ctype_out, uint8_t* alive_out) {
  if (!c) return HCG_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  if (cell_id_out) std::copy(c->h_cell_id.begin(), c->h_cell_id.end(), cell_id_out);
  if (ctype_out) std::copy(c->h_cell_type.begin(), c->h_cell_type.end(), ctype_out);
  if (alive_out && c->ncells) {
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    CUDA_TRY(c, cudaMemcpy(alive_out, c->cell_alive, c->ncells, cudaMemcpyDeviceToHost));
  }
  return HCG_OK;
}

hcg_status hcg_cells_add_force(hcg_ctx* c, int64_t n, const int64_t* idx, const double* f) {
  if (!c || n < 0 || (n > 0 && (!idx || !f))) return HCG_ERR_ARG;
  if (n == 0) return HCG_OK;
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  hcg_status s = ensure_staging(c, (sizeof(int64_t) + 3*sizeof(double))*n); if (s) return s;
  int64_t* di = (int64_t*)c->staging; double* df = (double*)(di + n);
  CUDA_TRY(c, cudaMemcpyAsync(di, idx, sizeof(int64_t)*n, cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(c, cudaMemcpyAsync(df, f, sizeof(double)*3*n, cudaMemcpyHostToDevice, c->stream));
  k_add_force<<<nblk(n, 128), 128, 0, c->stream>>>(n, di, df, c->frc[0], c->frc[1], c->frc[2], c->np);
  KERNEL_CHECK(c);
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));   // host buffers are caller-owned
  return HCG_OK;
}

hcg_status hcg_set_force_limit(hcg_ctx* c, double f) { if (!c || !(f > 0)) return HCG_ERR_ARG; c->f_limit = f; return HCG_OK; }
hcg_status hcg_set_timescales(hcg_ctx* c, int32_t v, int32_t r, int32_t w) {
  if (!c || v < 1 || r < 1 || w < 1) return HCG_ERR_ARG;
  c->ts_vel = v; c->ts_rep = r; c->ts_wall = w; return HCG_OK;
}
hcg_status hcg_set_material_timescale(hcg_ctx* c, int32_t ctype, int32_t every) {
  if (!c || ctype < 0 || ctype >= (int)c->types.size() || every < 1) return HCG_ERR_ARG;
  c->types[ctype].timescale = every; return HCG_OK;
}
hcg_status hcg_set_repulsion(hcg_ctx* c, int32_t on, double k, double cut) {
  if (!c || (on && !(cut > 0))) return HCG_ERR_ARG;
  c->rep_on = on != 0; c->rep_k = k; c->rep_cut = cut; return HCG_OK;
}
hcg_status hcg_set_wall_repulsion(hcg_ctx* c, int32_t on, double k, double cut) {
  if (!c || (on && !(cut > 0))) return HCG_ERR_ARG;
  c->wall_on = on != 0; c->wall_k = k; c->wall_cut = cut; return HCG_OK;
}
hcg_status hcg_set_moment_only(hcg_ctx* c, int32_t on) {
  if (!c) return HCG_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  if (!on) { hcg_status s = lat_ensure_pops(c); if (s) return s; }      // back to stored populations
  c->mo_mode = on < 0 ? 0 : on;                                          // 1 = single rank, 2 = also slab-decomposed runs
  return HCG_OK;
}
hcg_status hcg_set_spread_mode(hcg_ctx* c, int32_t mode, int32_t resort_every) {
  if (!c || mode < 0 || mode > 1 || resort_every < 1) return HCG_ERR_ARG;
  c->spread_mode = mode; c->perm_every = resort_every; c->perm_valid = false; c->far_steps_left = 0;
  return HCG_OK;
}
hcg_status hcg_set_exchange(hcg_ctx* c, double margin_lu, int32_t sync_every, double slack) {
  if (!c || !(margin_lu >= 2.0) || sync_every < 1 || slack < 0) return hcg_fail(c, HCG_ERR_ARG, "exchange: margin >= 2 lu, sync_every >= 1, slack >= 0");
  if (c->ncells > 0) return hcg_fail(c, HCG_ERR_STATE, "hcg_set_exchange must precede hcg_cells_add");
  if (c->dom.n_ranks > 1 && 2*margin_lu + 20 > c->nxl) return hcg_fail(c, HCG_ERR_ARG, "slab too thin for this margin");
  c->multi.margin = margin_lu; c->multi.sync_every = sync_every; c->multi.slack = slack;
  return HCG_OK;
}
hcg_status hcg_set_transport(hcg_ctx* c, int32_t transport) {
  if (!c || transport < 0 || transport > 1) return HCG_ERR_ARG;
  if (comm_up(c)) return hcg_fail(c, HCG_ERR_STATE, "hcg_set_transport must precede hcg_comm_init");
  c->peer.transport = transport;
  return HCG_OK;
}
hcg_status hcg_exchange_stats(hcg_ctx* c, int64_t* shared_left, int64_t* shared_right, int64_t* migrated_in, int64_t* migrated_out) {
  if (!c) return HCG_ERR_ARG;
  if (shared_left) *shared_left = c->multi.face[0].n;
  if (shared_right) *shared_right = c->multi.face[1].n;
  if (migrated_in) *migrated_in = c->multi.migrated_in;
  if (migrated_out) *migrated_out = c->multi.migrated_out;
  return HCG_OK;
}
hcg_status hcg_set_iteration(hcg_ctx* c, int64_t it) { if (!c || it < 0) return HCG_ERR_ARG; c->iter = it; c->multi.next_sync_iter = it + c->multi.sync_every; return HCG_OK; }
hcg_status hcg_get_iteration(hcg_ctx* c, int64_t* it) { if (!c || !it) return HCG_ERR_ARG; *it = c->iter; return HCG_OK; }

hcg_status hcg_iterate(hcg_ctx* c, int64_t n) {
  if (!c || n < 0) return HCG_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  // sanityCheck (core/hemoCell.cpp:600-627): material / repulsion cadences are multiples of the velocity cadence
  for (auto& t : c->types) if (t.timescale % c->ts_vel) return hcg_fail(c, HCG_ERR_STATE, "material timescale must be a multiple of the velocity timescale");
  if (c->rep_on && c->ts_rep % c->ts_vel) return hcg_fail(c, HCG_ERR_STATE, "repulsion timescale must be a multiple of the velocity timescale");
  c->in_iterate = true;
  for (int64_t i = 0; i < n; i++) { hcg_status s = step(c); if (s) { c->in_iterate = false; return s; } }
  c->in_iterate = false;
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return HCG_OK;
}

hcg_status hcg_iterate_async(hcg_ctx* c, int64_t n) {
  if (!c || n < 0) return HCG_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  for (auto& t : c->types) if (t.timescale % c->ts_vel) return hcg_fail(c, HCG_ERR_STATE, "material timescale must be a multiple of the velocity timescale");
  if (c->rep_on && c->ts_rep % c->ts_vel) return hcg_fail(c, HCG_ERR_STATE, "repulsion timescale must be a multiple of the velocity timescale");
  c->in_iterate = true;
  for (int64_t i = 0; i < n; i++) { hcg_status s = step(c); if (s) { c->in_iterate = false; return s; } }
  c->in_iterate = false;
  return HCG_OK;
}

hcg_status hcg_iterate_timed(hcg_ctx* c, int64_t n, double* ms_out) {
  if (!c || n < 0 || !ms_out) return HCG_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  cudaEvent_t a, b;
  CUDA_TRY(c, cudaEventCreate(&a)); CUDA_TRY(c, cudaEventCreate(&b));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  CUDA_TRY(c, cudaEventRecord(a, c->stream));
  c->in_iterate = true;
  for (int64_t i = 0; i < n; i++) { hcg_status s = step(c); if (s) { c->in_iterate = false; return s; } }
  c->in_iterate = false;
  CUDA_TRY(c, cudaEventRecord(b, c->stream));
  CUDA_TRY(c, cudaEventSynchronize(b));
  float ms = 0; CUDA_TRY(c, cudaEventElapsedTime(&ms, a, b));
  cudaEventDestroy(a); cudaEventDestroy(b);
  *ms_out = ms;
  return HCG_OK;
}

hcg_status hcg_fluid_warmup(hcg_ctx* c, int64_t n) {
  if (!c || n < 0) return HCG_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  // case files call lattice->collideAndStream() directly: the node force is NOT reset
  for (int64_t i = 0; i < n; i++) { hcg_status s = lat_collide_stream(c, false); if (s) return s; }
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return HCG_OK;
}

#define OP_PROLOGUE if (!c) return HCG_ERR_ARG; CUDA_TRY(c, cudaSetDevice(c->dom.device)); hcg_status s
#define OP_EPILOGUE CUDA_TRY(c, cudaStreamSynchronize(c->stream)); return HCG_OK
hcg_status hcg_op_repulsion(hcg_ctx* c) { OP_PROLOGUE; if ((s = rep_apply(c))) return s; OP_EPILOGUE; }
hcg_status hcg_op_wall_repulsion(hcg_ctx* c) { OP_PROLOGUE; if ((s = rep_wall_apply(c))) return s; OP_EPILOGUE; }
hcg_status hcg_op_spread(hcg_ctx* c) { OP_PROLOGUE; c->f_clean = false; if ((s = do_spread(c))) return s; OP_EPILOGUE; }
hcg_status hcg_op_collide_stream(hcg_ctx* c) { OP_PROLOGUE; if ((s = lat_collide_stream(c, false))) return s; OP_EPILOGUE; }
hcg_status hcg_op_interpolate(hcg_ctx* c) {
  OP_PROLOGUE; if ((s = lat_moments(c, false, false))) return s; if ((s = ibm_interpolate(c))) return s; OP_EPILOGUE;
}
// syncEnvelopes as iterate() performs it after the interpolation (core/hemoCell.cpp:333-341): the holders of every shared cell
// exchange the velocities (and alive flags) of the vertices they own; then, like the reference's envelope update, membership is
// re-evaluated and whole cells migrate.  Collective over the ranks; a single-rank context has nothing to exchange.
hcg_status hcg_op_sync(hcg_ctx* c) {
  OP_PROLOGUE;
  if (c->dom.n_ranks > 1) {
    if ((s = multi_velocity_sync(c))) return s;
    if ((s = multi_rebalance(c, false))) return s;
  }
  OP_EPILOGUE;
}
hcg_status hcg_op_advance(hcg_ctx* c) { OP_PROLOGUE; if ((s = ibm_advance(c))) return s; OP_EPILOGUE; }
hcg_status hcg_op_mechanics(hcg_ctx* c, int32_t forced, int32_t components) {
  OP_PROLOGUE; if ((s = do_mechanics(c, forced != 0, components != 0))) return s; OP_EPILOGUE;
}
hcg_status hcg_op_zero_force(hcg_ctx* c) { OP_PROLOGUE; if ((s = lat_reset_force(c))) return s; OP_EPILOGUE; }

hcg_status hcg_cells_owned(hcg_ctx* c, uint8_t* owned) {
  if (!c || !owned) return HCG_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  if (c->ncells == 0) return HCG_OK;
  hcg_status s = ensure_staging(c, (size_t)c->ncells); if (s) return s;
  k_cells_owned<<<nblk(c->ncells, 256), 256, 0, c->stream>>>(c->cell_alive, c->ncells, c->cell_base, c->pos[0], c->dom.n_ranks,
      c->dom.nx, c->dom.periodic[0], c->x0, c->nxl, (uint8_t*)c->staging);
  KERNEL_CHECK(c);
  CUDA_TRY(c, cudaMemcpyAsync(owned, c->staging, (size_t)c->ncells, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return HCG_OK;
}

hcg_status hcg_allreduce(hcg_ctx* c, double* inout, int64_t n, int32_t op) {
  if (!c || !inout || n < 0 || op < 0 || op > 2) return HCG_ERR_ARG;
  if (c->dom.n_ranks == 1 || n == 0) return HCG_OK;
  if (!comm_up(c)) return hcg_fail(c, HCG_ERR_STATE, "hcg_allreduce: hcg_comm_init first");
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  hcg_status s = ensure_staging(c, sizeof(double)*(size_t)n); if (s) return s;
  CUDA_TRY(c, cudaMemcpyAsync(c->staging, inout, sizeof(double)*n, cudaMemcpyHostToDevice, c->stream));
  if ((s = comm_allreduce_f64(c, c->staging, (size_t)n, op))) return s;
  CUDA_TRY(c, cudaMemcpyAsync(inout, c->staging, sizeof(double)*n, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return HCG_OK;
}

hcg_status hcg_cells_bbox(hcg_ctx* c, double* bbox) {
  if (!c || !bbox) return HCG_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  if (c->ncells == 0) return HCG_OK;
  hcg_status s = ensure_staging(c, sizeof(double)*6*c->ncells); if (s) return s;
  if ((s = mech_bbox(c, c->staging))) return s;
  CUDA_TRY(c, cudaMemcpyAsync(bbox, c->staging, sizeof(double)*6*c->ncells, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return HCG_OK;
}

hcg_status hcg_cells_volume_area(hcg_ctx* c, double* volume, double* area) {
  if (!c || !volume || !area) return HCG_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  if (c->ncells == 0) return HCG_OK;
  hcg_status s = ensure_staging(c, sizeof(double)*2*c->ncells); if (s) return s;
  if ((s = mech_volume_area(c, c->staging, c->staging + c->ncells))) return s;
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  CUDA_TRY(c, cudaMemcpy(volume, c->staging, sizeof(double)*c->ncells, cudaMemcpyDeviceToHost));
  CUDA_TRY(c, cudaMemcpy(area, c->staging + c->ncells, sizeof(double)*c->ncells, cudaMemcpyDeviceToHost));
  return HCG_OK;
}

hcg_status hcg_cells_stretch(hcg_ctx* c, double* stretch) {
  if (!c || !stretch) return HCG_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  if (c->ncells == 0) return HCG_OK;
  hcg_status s = ensure_staging(c, sizeof(double)*c->ncells); if (s) return s;
  CUDA_TRY(c, cudaMemsetAsync(c->staging, 0, sizeof(double)*c->ncells, c->stream));
  if ((s = mech_stretch(c, c->staging))) return s;
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  CUDA_TRY(c, cudaMemcpy(stretch, c->staging, sizeof(double)*c->ncells, cudaMemcpyDeviceToHost));
  return HCG_OK;
}

hcg_status hcg_fluid_velocity_stats(hcg_ctx* c, double* vmin, double* vmax, double* vmean) {
  if (!c || !vmin || !vmax || !vmean) return HCG_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  hcg_status s = lat_moments(c, false, false); if (s) return s;
  return lat_velocity_stats(c, vmin, vmax, vmean);
}

hcg_status hcg_timers_enable(hcg_ctx* c, int32_t on) { if (!c) return HCG_ERR_ARG; c->timers_on = on != 0; return HCG_OK; }
hcg_status hcg_timers_reset(hcg_ctx* c) { if (!c) return HCG_ERR_ARG; resolve_timers(c); c->timers.clear(); c->timer_idx.clear(); return HCG_OK; }
hcg_status hcg_timers(hcg_ctx* c, hcg_timer* out, int32_t* n) {
  if (!c || !n) return HCG_ERR_ARG;
  resolve_timers(c);
  const int have = (int)c->timers.size();
  if (out) for (int k = 0; k < have && k < *n; k++) {
    memset(out[k].name, 0, sizeof(out[k].name));
    strncpy(out[k].name, c->timers[k].name.c_str(), sizeof(out[k].name) - 1);
    out[k].ms_total = c->timers[k].ms; out[k].calls = c->timers[k].calls;
  }
  *n = have;
  return HCG_OK;
}
hcg_status hcg_launch_count(hcg_ctx* c, int64_t* n) { if (!c || !n) return HCG_ERR_ARG; *n = c->launches; return HCG_OK; }
hcg_status hcg_synchronize(hcg_ctx* c) { if (!c) return HCG_ERR_ARG; CUDA_TRY(c, cudaSetDevice(c->dom.device)); CUDA_TRY(c, cudaStreamSynchronize(c->stream)); return HCG_OK; }

}  // extern "C"
