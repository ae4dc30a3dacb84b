// Multi-GPU particle handling for the x-slab decomposition (one context per rank/GPU).
// Replaces HemoCellFields::syncEnvelopes + HemoCellParticleDataTransfer (reference
// core/hemoCellFields.cpp:377-499, core/hemoCellParticleDataTransfer.cpp:33-466) and the
// deleteNonLocalParticles / deleteIncompleteCells bookkeeping (hemoCellFields.cpp:676-688).
//
// Scheme ("replicated whole cells", the reference's own strategy):
//  * a rank HOLDS every cell whose x-extent intersects [x0 - M, x0 + nxl + M); cells near a slab
//    face are therefore held by both neighbours, as complete bit-identical copies;
//  * every holder spreads onto its own real nodes and computes the membrane forces of the whole
//    cell redundantly from identical inputs -> no force halo is ever communicated;
//  * on velocity-interpolation steps each vertex velocity is authoritative on the rank that OWNS
//    the vertex (x in [x0 - 0.5, x0 + nxl - 0.5)); the two holders of a shared cell swap their
//    velocity arrays (NCCL send/recv) and keep the neighbour's value for vertices they do not own;
//    the cells' alive flags are AND-ed in the same message;
//  * every `sync_every` steps membership is re-evaluated from the cells' bounding boxes: cells that
//    entered a neighbour's hold region are shipped whole (pos, vel, force, frep), cells that left
//    the own region are dropped.  M = 2 (kernel support) + drift allowance.
#include "ctx.cuh"
#include "ibm_node.cuh"
#include <nccl.h>
#include <algorithm>
#include <cmath>
#include <cstring>

namespace {

inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1)/t); }

// does [lo, hi] (unwrapped) intersect [a, b) on a circle of length nx (or on the line)?
inline bool band_hit(double lo, double hi, double a, double b, int nx, bool periodic) {
  for (int k = periodic ? -1 : 0; k <= (periodic ? 1 : 0); k++) {
    const double l = lo + (double)k*nx, h = hi + (double)k*nx;
    if (l < b && h >= a) return true;
  }
  return false;
}

// owner test shared by both holders of a cell (pre-advance position, identical bits on both sides):
// the vertex' nearest node lies inside the slab [x0, x0 + nxl)
__device__ __forceinline__ bool owns_vertex(double x, int nx, int px, int x0, int nxl) {
  int gx = (int)floor(x + 0.5);
  if (px) { gx %= nx; if (gx < 0) gx += nx; }
  int rel = gx - x0; if (px && rel < 0) rel += nx;
  return rel >= 0 && rel < nxl;
}

// message = [n alive flags][vx of all shared vertices][vy ...][vz ...]; only the entries of vertices this
// rank OWNS are written (the receiver reads exactly those: the vertices it does not own), which halves the
// NVLink traffic; component blocks keep the stores coalesced.
__global__ void __launch_bounds__(256)
k_pack_sync(const int32_t* __restrict__ cells, const int64_t* __restrict__ off, int n, int64_t total,
            const int64_t* __restrict__ cell_base, const uint8_t* __restrict__ alive,
            const double* __restrict__ x,
            const double* __restrict__ vx, const double* __restrict__ vy, const double* __restrict__ vz,
            double* buf, int nx, int px, int x0, int nxl) {
  const int i = blockIdx.x;
  if (i >= n) return;
  const int c = cells[i];
  const int64_t b = cell_base[c], o = off[i];
  const int V = (int)(off[i+1] - o);
  if (threadIdx.x == 0) buf[i] = alive[c] ? 1.0 : 0.0;
  double* v = buf + n + o;
  for (int k = threadIdx.x; k < V; k += blockDim.x) {
    if (!owns_vertex(x[b+k], nx, px, x0, nxl)) continue;
    v[k] = vx[b+k]; v[total + k] = vy[b+k]; v[2*total + k] = vz[b+k];
  }
}

__global__ void __launch_bounds__(256)
k_unpack_sync(const int32_t* __restrict__ cells, const int64_t* __restrict__ off, int n, int64_t total,
              const int64_t* __restrict__ cell_base, uint8_t* alive,
              const double* __restrict__ x, double* vx, double* vy, double* vz,
              const double* __restrict__ buf, int nx, int px, int x0, int nxl) {
  const int i = blockIdx.x;
  if (i >= n) return;
  const int c = cells[i];
  const int64_t b = cell_base[c], o = off[i];
  const int V = (int)(off[i+1] - o);
  if (threadIdx.x == 0 && buf[i] == 0.0) alive[c] = 0;
  const double* v = buf + n + o;
  for (int k = threadIdx.x; k < V; k += blockDim.x) {
    if (owns_vertex(x[b+k], nx, px, x0, nxl)) continue;      // authoritative here
    vx[b+k] = v[k]; vy[b+k] = v[total + k]; vz[b+k] = v[2*total + k];
  }
}

// The two faces in one launch, and the unpack fused with the advance of the shared cells (a shared cell sits on exactly one
// face list - multi_rebalance refuses slabs thin enough for a cell to reach both faces - so one CTA owns it): per velocity
// sync 3 launches (pack, flag barrier, unpack + advance) instead of 6.
struct SyncFace { const int32_t* cells; const int64_t* off; int n; int64_t total; double* send; const double* recv; };
__global__ void __launch_bounds__(256)
k_pack_sync2(SyncFace f0, SyncFace f1, const int64_t* __restrict__ cell_base, const uint8_t* __restrict__ alive,
             const double* __restrict__ x, const double* __restrict__ vx, const double* __restrict__ vy, const double* __restrict__ vz,
             int nx, int px, int x0, int nxl) {
  const bool second = (int)blockIdx.x >= f0.n;
  const SyncFace& f = second ? f1 : f0;
  const int i = second ? (int)blockIdx.x - f0.n : (int)blockIdx.x;
  if (i >= f.n || !f.send) return;
  const int c = f.cells[i];
  const int64_t b = cell_base[c], o = f.off[i];
  const int V = (int)(f.off[i+1] - o);
  if (threadIdx.x == 0) f.send[i] = alive[c] ? 1.0 : 0.0;
  double* v = f.send + f.n + o;
  for (int k = threadIdx.x; k < V; k += blockDim.x) {
    if (!owns_vertex(x[b+k], nx, px, x0, nxl)) continue;
    v[k] = vx[b+k]; v[f.total + k] = vy[b+k]; v[2*f.total + k] = vz[b+k];
  }
}
__global__ void __launch_bounds__(256)
k_unpack_advance2(SyncFace f0, SyncFace f1, IbmArgs a, const uint8_t* __restrict__ flags, const int64_t* __restrict__ cell_base,
                  uint8_t* alive, double* x, double* y, double* z, double* vx, double* vy, double* vz) {
  const bool second = (int)blockIdx.x >= f0.n;
  const SyncFace& f = second ? f1 : f0;
  const int i = second ? (int)blockIdx.x - f0.n : (int)blockIdx.x;
  if (i >= f.n) return;
  const int c = f.cells[i];
  const int64_t b = cell_base[c], o = f.off[i];
  const int V = (int)(f.off[i+1] - o);
  const bool live = alive[c] && (!f.recv || f.recv[i] != 0.0);   // deleted on either holder = deleted
  __syncthreads();                                        // every thread has read the flag before thread 0 may clear it
  if (!live) { if (threadIdx.x == 0) alive[c] = 0; return; }
  const double* v = f.recv ? f.recv + f.n + o : nullptr;
  for (int k = threadIdx.x; k < V; k += blockDim.x) {
    const int64_t p = b + k;
    double v0 = vx[p], v1 = vy[p], v2 = vz[p];
    if (v && !owns_vertex(x[p], a.nx, a.px, a.x0, a.nxl)) { v0 = v[k]; v1 = v[f.total + k]; v2 = v[2*f.total + k]; vx[p] = v0; vy[p] = v1; vz[p] = v2; }
    const double qx = x[p] + v0, qy = y[p] + v1, qz = z[p] + v2;
    x[p] = qx; y[p] = qy; z[p] = qz;
    // particle on a boundary node => its cell is deleted (hemoCellParticleField.cpp:572-584)
    int lx; bool out;
    int yy = (int)floor(qy + 0.5), zz = (int)floor(qz + 0.5);
    if (local_x((int)floor(qx + 0.5), a, lx, out) && wrap_yz(yy, a.ny, a.py) && wrap_yz(zz, a.nz, a.pz)) {
      if (flags[(int64_t)zz + (int64_t)a.nz*((int64_t)yy + (int64_t)a.ny*lx)] != HCG_FLUID) alive[c] = 0;
    }
  }
}

// largest velocity component over the particles of live cells (non-negative doubles order like their bit patterns)
__global__ void k_vmax(const double* __restrict__ vx, const double* __restrict__ vy, const double* __restrict__ vz,
                       const int32_t* __restrict__ p_cell, const uint8_t* __restrict__ alive, int64_t np, unsigned long long* out) {
  double m = 0.0;
  for (int64_t p = (int64_t)blockIdx.x*blockDim.x + threadIdx.x; p < np; p += (int64_t)gridDim.x*blockDim.x)
    if (alive[p_cell[p]]) m = fmax(m, fmax(fabs(vx[p]), fmax(fabs(vy[p]), fabs(vz[p]))));
  for (int s = 16; s > 0; s >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, s));
  if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}

// whole-cell migration payload: per cell 12*V doubles (pos, vel, force, frep as xyz triples)
__global__ void k_pack_cells(const int32_t* __restrict__ cells, const int64_t* __restrict__ off, int n,
                             const int64_t* __restrict__ cell_base,
                             const double* const* __restrict__ arr /* 12 SoA arrays */, double* buf) {
  const int i = blockIdx.x;
  if (i >= n) return;
  const int64_t b = cell_base[cells[i]], o = off[i];
  const int V = (int)(off[i+1] - o);
  double* out = buf + 12*o;
  for (int k = threadIdx.x; k < 12*V; k += blockDim.x) { const int a = k / V, v = k - a*V; out[k] = arr[a][b+v]; }
}
__global__ void k_unpack_cells(const int32_t* __restrict__ cells, const int64_t* __restrict__ off, int n,
                               const int64_t* __restrict__ cell_base, double* const* __restrict__ arr,
                               const double* __restrict__ buf, uint8_t* alive) {
  const int i = blockIdx.x;
  if (i >= n) return;
  const int c = cells[i];
  const int64_t b = cell_base[c], o = off[i];
  const int V = (int)(off[i+1] - o);
  const double* in = buf + 12*o;
  for (int k = threadIdx.x; k < 12*V; k += blockDim.x) { const int a = k / V, v = k - a*V; arr[a][b+v] = in[k]; }
  if (threadIdx.x == 0) alive[c] = 1;
}

// grouped neighbour exchange of raw bytes; order keeps the 2-rank periodic case matched
hcg_status neighbour_exchange(hcg_ctx* c, const void* sendL, size_t nsL, const void* sendR, size_t nsR,
                              void* recvR, size_t nrR, void* recvL, size_t nrL) {
  const int R = c->dom.n_ranks, r = c->dom.rank; const bool px = c->dom.periodic[0];
  const int left = (r == 0) ? (px ? R - 1 : -1) : r - 1;
  const int right = (r == R - 1) ? (px ? 0 : -1) : r + 1;
  hcg_status s = comm_group_begin(c); if (s) return s;
  if (left >= 0 && nsL) comm_send(c, sendL, nsL, left);
  if (right >= 0 && nsR) comm_send(c, sendR, nsR, right);
  if (right >= 0 && nrR) comm_recv(c, recvR, nrR, right);
  if (left >= 0 && nrL) comm_recv(c, recvL, nrL, left);
  return comm_group_end(c, "neighbour exchange");
}

hcg_status ensure_buf(hcg_ctx* c, double** p, size_t* cap, size_t doubles) {
  if (*cap >= doubles) return HCG_OK;
  if (*p) cudaFree(*p);
  *p = nullptr; *cap = 0;
  const size_t want = doubles + doubles/4 + 1024;
  CUDA_TRY(c, cudaMalloc(p, sizeof(double)*want));
  *cap = want;
  return HCG_OK;
}

// upload a face's shared list (cells sorted by global id) and its particle offsets
hcg_status upload_list(hcg_ctx* c, MultiFace& f, const std::vector<int32_t>& cells) {
  f.n = (int)cells.size();
  std::vector<int64_t> off(f.n + 1, 0);
  for (int i = 0; i < f.n; i++) off[i+1] = off[i] + c->types[c->h_cell_type[cells[i]]].d.V;
  f.total = off[f.n];
  if (f.cap < f.n + 1) {
    if (f.d_cells) { cudaFree(f.d_cells); cudaFree(f.d_off); }
    f.cap = f.n + 1 + f.n/2 + 64;
    CUDA_TRY(c, cudaMalloc(&f.d_cells, sizeof(int32_t)*f.cap));
    CUDA_TRY(c, cudaMalloc(&f.d_off, sizeof(int64_t)*f.cap));
  }
  if (f.n) CUDA_TRY(c, cudaMemcpyAsync(f.d_cells, cells.data(), sizeof(int32_t)*f.n, cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(c, cudaMemcpyAsync(f.d_off, off.data(), sizeof(int64_t)*(f.n + 1), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));     // host vectors go out of scope
  return HCG_OK;
}

}  // namespace

hcg_status multi_neighbour_exchange(hcg_ctx* c, const void* sendL, size_t nsL, const void* sendR, size_t nsR,
                                    void* recvR, size_t nrR, void* recvL, size_t nrL) {
  return neighbour_exchange(c, sendL, nsL, sendR, nsR, recvR, nrR, recvL, nrL);
}

// pure host logic, exported for CPU tests (include/hemocell_host.h)
extern "C" void hch_slab_membership_at(int64_t n, const double* xlo, const double* xhi, int32_t nx, int32_t periodic_x,
                                       int32_t x0_, int32_t nxl, int32_t rank, int32_t n_ranks, double margin,
                                       uint8_t* held, uint8_t* share_left, uint8_t* share_right) {
  const double x0 = (double)x0_, x1 = x0 + nxl;
  const bool px = periodic_x != 0;
  const bool has_left = n_ranks > 1 && (rank > 0 || px), has_right = n_ranks > 1 && (rank < n_ranks - 1 || px);
  for (int64_t i = 0; i < n; i++) {
    const bool h = n_ranks == 1 ? true : band_hit(xlo[i], xhi[i], x0 - margin, x1 + margin, nx, px);
    held[i] = h;
    share_left[i] = h && has_left && band_hit(xlo[i], xhi[i], x0 - margin, x0 + margin, nx, px);
    share_right[i] = h && has_right && band_hit(xlo[i], xhi[i], x1 - margin, x1 + margin, nx, px);
  }
}
extern "C" void hch_slab_membership(int64_t n, const double* xlo, const double* xhi, int32_t nx, int32_t periodic_x,
                                    int32_t nxl, int32_t rank, int32_t n_ranks, double margin,
                                    uint8_t* held, uint8_t* share_left, uint8_t* share_right) {
  hch_slab_membership_at(n, xlo, xhi, nx, periodic_x, nxl*rank, nxl, rank, n_ranks, margin, held, share_left, share_right);
}

hcg_status multi_upload_cell_gid(hcg_ctx* c) {
  if (!c->cell_gid_dirty && c->cell_gid) return HCG_OK;
  const int64_t nc = c->ncells;
  if (c->cell_gid_cap < nc || !c->cell_gid) {
    if (c->cell_gid) cudaFree(c->cell_gid);
    CUDA_TRY(c, cudaMalloc(&c->cell_gid, sizeof(int64_t)*(size_t)std::max<int64_t>(nc, 1)));
    c->cell_gid_cap = nc;
  }
  if (nc) CUDA_TRY(c, hcg_h2d(c, c->cell_gid, c->h_cell_id.data(), sizeof(int64_t)*nc));
  c->cell_gid_dirty = false;
  return HCG_OK;
}

hcg_status multi_velocity_sync(hcg_ctx* c) { return multi_field_sync(c, 0); }

// velocity sync of the shared cells + their advance (the step's syncEnvelopes + advanceParticles for the cells a neighbour
// also holds).  Peer transport: pack (both faces) -> flag barrier -> unpack + advance; otherwise the separate kernels.
static hcg_status sync_faces(hcg_ctx* c, SyncFace f[2], size_t half) {
  MultiState& m = c->multi;
  for (int k = 0; k < 2; k++) {
    const bool on = m.face[k].n > 0 && c->peer.link[k].rank >= 0;
    f[k].cells = m.face[k].d_cells; f[k].off = m.face[k].d_off; f[k].n = on ? m.face[k].n : 0; f[k].total = m.face[k].total;
    f[k].send = nullptr; f[k].recv = nullptr;
    if (!on) continue;
    double* dst = (double*)c->peer.link[k].ptr[4 + (1 - k)];
    if (!dst) return hcg_fail(c, HCG_ERR_STATE, "peer transport: neighbour receive buffer not mapped");
    const size_t span = (size_t)m.face[k].n + 3*(size_t)m.face[k].total;
    f[k].send = dst + half*span; f[k].recv = c->peer.sync_recv[k] + half*span;
  }
  return HCG_OK;
}
// pack the shared cells' velocities (both faces) into the neighbours' receive buffers and publish them; the caller may put
// independent work (the interpolation of the unshared cells) before multi_sync_wait_unpack_advance
hcg_status multi_sync_pack_post(hcg_ctx* c) {
  c->multi.sync_half = (size_t)(c->peer.sync_count++ & 1ULL);
  SyncFace f[2];
  hcg_status s = sync_faces(c, f, c->multi.sync_half); if (s) return s;
  const int nblocks = f[0].n + f[1].n;
  if (nblocks > 0) {
    k_pack_sync2<<<nblocks, 256, 0, c->stream>>>(f[0], f[1], c->cell_base, c->cell_alive, c->pos[0], c->vel[0], c->vel[1], c->vel[2],
                                                 c->dom.nx, c->dom.periodic[0], c->x0, c->nxl);
    KERNEL_CHECK(c);
  }
  return peer_post(c);
}
hcg_status multi_sync_wait_unpack_advance(hcg_ctx* c) {
  hcg_status s = peer_wait(c); if (s) return s;
  SyncFace f[2];
  if ((s = sync_faces(c, f, c->multi.sync_half))) return s;
  const int nblocks = f[0].n + f[1].n;
  if (nblocks > 0) {
    IbmArgs a;
    a.nx = c->dom.nx; a.ny = c->dom.ny; a.nz = c->dom.nz; a.px = c->dom.periodic[0]; a.py = c->dom.periodic[1]; a.pz = c->dom.periodic[2];
    a.nxl = c->nxl; a.x0 = c->x0; a.nranks = c->dom.n_ranks; a.P = c->P; a.S = c->S; a.np = c->np; a.f_limit = c->f_limit;
    k_unpack_advance2<<<nblocks, 256, 0, c->stream>>>(f[0], f[1], a, c->flags, c->cell_base, c->cell_alive, c->pos[0], c->pos[1], c->pos[2],
                                                      c->vel[0], c->vel[1], c->vel[2]);
    KERNEL_CHECK(c);
  }
  return HCG_OK;
}
hcg_status multi_velocity_sync_advance(hcg_ctx* c) {
  if (c->dom.n_ranks == 1) return HCG_OK;
  hcg_status s;
  if (!peer_on(c)) {
    if ((s = multi_field_sync(c, 0))) return s;
    return ibm_advance_shared(c);
  }
  if ((s = multi_sync_pack_post(c))) return s;
  return multi_sync_wait_unpack_advance(c);
}

// per-vertex swap of a particle field of the shared cells: the rank that OWNS a vertex is authoritative.
// field 0: velocity (every interpolation step; the alive flags are AND-ed in the same message);
// field 1: repulsion force (after applyRepulsionForce / applyBoundaryRepulsionForce)
hcg_status multi_field_sync(hcg_ctx* c, int field) {
  if (c->dom.n_ranks == 1) return HCG_OK;
  MultiState& m = c->multi;
  double* const* arr = field == 0 ? c->vel : c->frep;
  hcg_status s;
  size_t ns[2], off_send[2], off_recv[2];
  size_t tot = 0;
  if (peer_on(c)) {
    // peer-memory path: the pack kernel stores straight into the neighbour's receive buffer of the facing
    // face (left neighbour: its right-face buffer, ptr[5]; right neighbour: its left-face buffer, ptr[4])
    // Receive buffers are double-buffered by sync parity: the neighbour's unpack of sync k may still be
    // running when I pack sync k+1 (no barrier in between), but never when I pack sync k+2.
    const size_t half = (size_t)(c->peer.sync_count++ & 1ULL);
    for (int f = 0; f < 2; f++) {
      if (!m.face[f].n || c->peer.link[f].rank < 0) continue;
      double* dst = (double*)c->peer.link[f].ptr[4 + (1 - f)];
      if (!dst) return hcg_fail(c, HCG_ERR_STATE, "peer transport: neighbour receive buffer not mapped");
      dst += half*((size_t)m.face[f].n + 3*(size_t)m.face[f].total);
      k_pack_sync<<<m.face[f].n, 256, 0, c->stream>>>(m.face[f].d_cells, m.face[f].d_off, m.face[f].n, m.face[f].total,
          c->cell_base, c->cell_alive, c->pos[0], arr[0], arr[1], arr[2], dst,
          c->dom.nx, c->dom.periodic[0], c->x0, c->nxl);
      KERNEL_CHECK(c);
    }
    if ((s = peer_barrier(c))) return s;
    for (int f = 0; f < 2; f++) {
      if (!m.face[f].n || c->peer.link[f].rank < 0) continue;
      k_unpack_sync<<<m.face[f].n, 256, 0, c->stream>>>(m.face[f].d_cells, m.face[f].d_off, m.face[f].n, m.face[f].total,
          c->cell_base, c->cell_alive, c->pos[0], arr[0], arr[1], arr[2],
          c->peer.sync_recv[f] + half*((size_t)m.face[f].n + 3*(size_t)m.face[f].total),
          c->dom.nx, c->dom.periodic[0], c->x0, c->nxl);
      KERNEL_CHECK(c);
    }
    // the neighbour may overwrite my receive buffer only after it has seen my NEXT barrier, which follows these kernels
    return HCG_OK;
  }
  for (int f = 0; f < 2; f++) { ns[f] = (size_t)m.face[f].n + 3*(size_t)m.face[f].total; off_send[f] = tot; tot += ns[f]; }
  for (int f = 0; f < 2; f++) { off_recv[f] = tot; tot += ns[f]; }      // both sides hold the same lists
  if ((s = ensure_buf(c, &m.sync_buf, &m.sync_cap, tot))) return s;
  for (int f = 0; f < 2; f++) {
    if (!m.face[f].n) continue;
    k_pack_sync<<<m.face[f].n, 256, 0, c->stream>>>(m.face[f].d_cells, m.face[f].d_off, m.face[f].n, m.face[f].total,
        c->cell_base, c->cell_alive, c->pos[0], arr[0], arr[1], arr[2], m.sync_buf + off_send[f],
        c->dom.nx, c->dom.periodic[0], c->x0, c->nxl);
    KERNEL_CHECK(c);
  }
  if ((s = neighbour_exchange(c, m.sync_buf + off_send[0], 8*ns[0], m.sync_buf + off_send[1], 8*ns[1],
                              m.sync_buf + off_recv[1], 8*ns[1], m.sync_buf + off_recv[0], 8*ns[0]))) return s;
  for (int f = 0; f < 2; f++) {
    if (!m.face[f].n) continue;
    k_unpack_sync<<<m.face[f].n, 256, 0, c->stream>>>(m.face[f].d_cells, m.face[f].d_off, m.face[f].n, m.face[f].total,
        c->cell_base, c->cell_alive, c->pos[0], arr[0], arr[1], arr[2], m.sync_buf + off_recv[f],
        c->dom.nx, c->dom.periodic[0], c->x0, c->nxl);
    KERNEL_CHECK(c);
  }
  return HCG_OK;
}

// membership re-evaluation + whole-cell migration (host coordinated; every sync_every steps)
hcg_status multi_rebalance(hcg_ctx* c, bool initial) {
  if (c->dom.n_ranks == 1) return HCG_OK;
  MultiState& m = c->multi;
  hcg_status s;
  const int64_t nc = c->ncells;
  const int nx = c->dom.nx; const bool px = c->dom.periodic[0];
  // 0. agree on the alive flags first (a boundary hit is seen only by the rank that holds the node)
  if (!initial && (s = multi_velocity_sync(c))) return s;
  // 1. bounding boxes and alive flags of every slot
  double vmax = 0.0;
  std::vector<double> bbox(6*(size_t)std::max<int64_t>(nc, 1));
  std::vector<uint8_t> alive(std::max<int64_t>(nc, 1));
  if (nc) {
    if (m.bbox_cap < (size_t)nc) {
      if (m.d_bbox) cudaFree(m.d_bbox);
      m.d_bbox = nullptr; m.bbox_cap = 0;
      CUDA_TRY(c, cudaMalloc(&m.d_bbox, sizeof(double)*6*(size_t)nc));
      m.bbox_cap = (size_t)nc;
    }
    if ((s = mech_bbox(c, m.d_bbox))) return s;
    if (!m.d_vmax) CUDA_TRY(c, cudaMalloc(&m.d_vmax, sizeof(double)));
    CUDA_TRY(c, cudaMemsetAsync(m.d_vmax, 0, sizeof(double), c->stream));
    if (c->np > 0) {
      k_vmax<<<296, 256, 0, c->stream>>>(c->vel[0], c->vel[1], c->vel[2], c->p_cell, c->cell_alive, c->np, (unsigned long long*)m.d_vmax);
      KERNEL_CHECK(c);
    }
    CUDA_TRY(c, cudaMemcpyAsync(&vmax, m.d_vmax, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(bbox.data(), m.d_bbox, sizeof(double)*6*nc, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(alive.data(), c->cell_alive, nc, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  }
  std::vector<double> lo(nc), hi(nc);
  for (int64_t i = 0; i < nc; i++) { lo[i] = bbox[6*i]; hi[i] = bbox[6*i+1]; }
  std::vector<uint8_t> held(nc), shl(nc), shr(nc);
  hch_slab_membership_at(nc, lo.data(), hi.data(), nx, px, c->x0, c->nxl, c->dom.rank, c->dom.n_ranks, m.margin,
                      held.data(), shl.data(), shr.data());
  // a cell held through BOTH faces (slab thinner than cell extent + 2 margin) would need a three-way velocity exchange: the
  // per-face swap below assumes that every vertex this rank does not own belongs to the one neighbour of that face
  for (int64_t i = 0; i < nc; i++)
    if (m.h_held[i] && held[i] && alive[i] && shl[i] && shr[i])
      return hcg_fail(c, HCG_ERR_ARG, "slab decomposition too fine: a cell reaches both faces of a slab (need nx / n_ranks >= cell extent + 2 * margin; use fewer ranks)");
  // 2. drops and departures
  std::vector<int32_t> send[2];
  bool alive_dirty = false;
  for (int64_t i = 0; i < nc; i++) {
    if (!m.h_held[i]) continue;
    if (!held[i] || !alive[i]) {          // left my hold region, or deleted on every holder: forget it
      m.h_held[i] = 0; m.h_shared[0][i] = m.h_shared[1][i] = 0;
      if (alive[i]) { alive[i] = 0; alive_dirty = true; }
      c->h_cell_id[i] = -1;
      m.free_slots[c->h_cell_type[i]].push_back((int32_t)i);
      continue;
    }
    const uint8_t sh[2] = {shl[i], shr[i]};
    for (int f = 0; f < 2; f++) {
      if (sh[f] && !m.h_shared[f][i] && alive[i] && !initial) send[f].push_back((int32_t)i);
      m.h_shared[f][i] = sh[f];
    }
  }
  // 3. exchange counts, then meta (id, type) and payload
  int64_t h_cnt[4] = {(int64_t)send[0].size(), (int64_t)send[1].size(), 0, 0};
  if (!m.d_cnt) CUDA_TRY(c, cudaMalloc(&m.d_cnt, sizeof(int64_t)*4));
  CUDA_TRY(c, cudaMemcpyAsync(m.d_cnt, h_cnt, sizeof(h_cnt), cudaMemcpyHostToDevice, c->stream));
  if ((s = neighbour_exchange(c, m.d_cnt, 8, m.d_cnt + 1, 8, m.d_cnt + 3, 8, m.d_cnt + 2, 8))) return s;
  CUDA_TRY(c, cudaMemcpyAsync(h_cnt, m.d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  const int64_t n_recv[2] = {h_cnt[2], h_cnt[3]};      // from left, from right
  const int R = c->dom.n_ranks, r = c->dom.rank;
  const bool has[2] = {r > 0 || px, r < R - 1 || px};
  // meta
  std::vector<int64_t> meta_send[2], meta_recv[2];
  int64_t *d_ms[2] = {nullptr, nullptr}, *d_mr[2] = {nullptr, nullptr};
  {
    // scratch of the migration messages lives across calls (no cudaMalloc / cudaFree - a device-wide sync - per rebalance)
    size_t want = 0;
    for (int f = 0; f < 2; f++) want += 2*send[f].size() + 2*(size_t)(has[f] ? n_recv[f] : 0);
    if (m.meta_cap < want) {
      if (m.d_meta) cudaFree(m.d_meta);
      m.d_meta = nullptr; m.meta_cap = 0;
      CUDA_TRY(c, cudaMalloc(&m.d_meta, sizeof(int64_t)*(want + want/2 + 256)));
      m.meta_cap = want + want/2 + 256;
    }
  }
  size_t meta_at = 0;
  for (int f = 0; f < 2; f++) {
    for (int32_t sl : send[f]) { meta_send[f].push_back(c->h_cell_id[sl]); meta_send[f].push_back(c->h_cell_type[sl]); }
    if (!meta_send[f].empty()) {
      d_ms[f] = m.d_meta + meta_at; meta_at += meta_send[f].size();
      CUDA_TRY(c, cudaMemcpyAsync(d_ms[f], meta_send[f].data(), sizeof(int64_t)*meta_send[f].size(), cudaMemcpyHostToDevice, c->stream));
    }
    meta_recv[f].resize(2*(size_t)(has[f] ? n_recv[f] : 0));
    if (!meta_recv[f].empty()) { d_mr[f] = m.d_meta + meta_at; meta_at += meta_recv[f].size(); }
  }
  if ((s = neighbour_exchange(c, d_ms[0], 8*meta_send[0].size(), d_ms[1], 8*meta_send[1].size(),
                              d_mr[1], 8*meta_recv[1].size(), d_mr[0], 8*meta_recv[0].size()))) return s;
  for (int f = 0; f < 2; f++)
    if (!meta_recv[f].empty()) CUDA_TRY(c, cudaMemcpyAsync(meta_recv[f].data(), d_mr[f], 8*meta_recv[f].size(), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  // payload
  MultiFace* tmp_send = m.tmp_send; MultiFace* tmp_recv = m.tmp_recv;      // device lists reused across calls
  std::vector<int32_t> arrive[2];
  for (int f = 0; f < 2; f++) {
    for (size_t k = 0; k < meta_recv[f].size()/2; k++) {
      const int t = (int)meta_recv[f][2*k+1];
      if (t < 0 || t >= (int)c->types.size()) return hcg_fail(c, HCG_ERR_STATE, "migration: unknown cell type received");
      int32_t slot;
      if (!m.free_slots[t].empty()) { slot = m.free_slots[t].back(); m.free_slots[t].pop_back(); }
      else {
        CellTypeHost& th = c->types[t];
        if (th.n_cells >= th.cap_cells) return hcg_fail(c, HCG_ERR_CAPACITY, "migration: no free cell slot (raise the slack)");
        slot = (int32_t)(th.first_cell + th.n_cells++);
      }
      c->h_cell_id[slot] = meta_recv[f][2*k];
      m.h_held[slot] = 1; m.h_shared[0][slot] = m.h_shared[1][slot] = 0;
      m.h_shared[f][slot] = 1;            // shared with the rank it came from, through that face
      arrive[f].push_back(slot);
    }
    if ((s = upload_list(c, tmp_send[f], send[f]))) return s;
    if ((s = upload_list(c, tmp_recv[f], arrive[f]))) return s;
  }
  size_t need = 0, o_s[2], o_r[2];
  for (int f = 0; f < 2; f++) { o_s[f] = need; need += 12*(size_t)tmp_send[f].total; }
  for (int f = 0; f < 2; f++) { o_r[f] = need; need += 12*(size_t)tmp_recv[f].total; }
  if ((s = ensure_buf(c, &m.mig_buf, &m.mig_cap, need))) return s;
  if (!m.d_arr) {
    double* h_arr[12];
    for (int k = 0; k < 3; k++) { h_arr[k] = c->pos[k]; h_arr[3+k] = c->vel[k]; h_arr[6+k] = c->frc[k]; h_arr[9+k] = c->frep[k]; }
    CUDA_TRY(c, cudaMalloc(&m.d_arr, sizeof(h_arr)));
    CUDA_TRY(c, hcg_h2d(c, m.d_arr, h_arr, sizeof(h_arr)));
  }
  for (int f = 0; f < 2; f++) if (tmp_send[f].n) {
    k_pack_cells<<<tmp_send[f].n, 256, 0, c->stream>>>(tmp_send[f].d_cells, tmp_send[f].d_off, tmp_send[f].n, c->cell_base,
                                                        (const double* const*)m.d_arr, m.mig_buf + o_s[f]);
    KERNEL_CHECK(c);
  }
  if ((s = neighbour_exchange(c, m.mig_buf + o_s[0], 8*12*(size_t)tmp_send[0].total, m.mig_buf + o_s[1], 8*12*(size_t)tmp_send[1].total,
                              m.mig_buf + o_r[1], 8*12*(size_t)tmp_recv[1].total, m.mig_buf + o_r[0], 8*12*(size_t)tmp_recv[0].total))) return s;
  if (alive_dirty) CUDA_TRY(c, cudaMemcpyAsync(c->cell_alive, alive.data(), nc, cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  for (int f = 0; f < 2; f++) if (tmp_recv[f].n) {
    k_unpack_cells<<<tmp_recv[f].n, 256, 0, c->stream>>>(tmp_recv[f].d_cells, tmp_recv[f].d_off, tmp_recv[f].n, c->cell_base,
                                                          m.d_arr, m.mig_buf + o_r[f], c->cell_alive);
    KERNEL_CHECK(c);
  }
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  m.migrated_in += (int64_t)arrive[0].size() + (int64_t)arrive[1].size();
  m.migrated_out += (int64_t)send[0].size() + (int64_t)send[1].size();
  // 4. shared lists, sorted by global cell id so that both holders pack in the same order
  for (int f = 0; f < 2; f++) {
    std::vector<int32_t> list;
    for (int64_t i = 0; i < nc; i++)
      if (m.h_held[i] && m.h_shared[f][i] && c->h_cell_id[i] >= 0) list.push_back((int32_t)i);
    std::sort(list.begin(), list.end(), [&](int32_t a, int32_t b) { return c->h_cell_id[a] < c->h_cell_id[b]; });
    if ((s = upload_list(c, m.face[f], list))) return s;
  }
  c->cell_gid_dirty = true; c->far_steps_left = 0;       // slots changed hands: every cell is checked against the walls until the next classification
  // union list + per-slot flag: the step advances unshared cells in the interpolation pass and the
  // shared ones after the velocity sync
  {
    std::vector<int32_t> list;
    std::vector<uint8_t> flag((size_t)std::max<int64_t>(nc, 1), 0);
    for (int64_t i = 0; i < nc; i++)
      if (m.h_held[i] && (m.h_shared[0][i] || m.h_shared[1][i]) && c->h_cell_id[i] >= 0) { list.push_back((int32_t)i); flag[i] = 1; }
    if ((s = upload_list(c, m.all, list))) return s;
    if (m.cell_shared_cap < nc || !m.d_cell_shared) {
      if (m.d_cell_shared) cudaFree(m.d_cell_shared);
      CUDA_TRY(c, cudaMalloc(&m.d_cell_shared, (size_t)std::max<int64_t>(nc, 1)));
      m.cell_shared_cap = nc;
    }
    CUDA_TRY(c, hcg_h2d(c, m.d_cell_shared, flag.data(), (size_t)std::max<int64_t>(nc, 1)));
  }
  // one small reduction settles (a) whether the peer mappings must be re-published - only when some rank's receive buffer was
  // outgrown (they carry 50 % headroom) - and (b) when membership is looked at next: the hold margin leaves 2 lu of drift
  // beyond the kernel support; the next rebalance comes when the fastest vertex of any rank could have used 1 lu of it at
  // its present speed (the other half is left for acceleration in between), at least sync_every and at most 8 sync_every
  // steps from now.  Slow suspensions (the benchmark: 1e-3 lu per step) thus pay the host round trip 8 times less often.
  bool changed = false;
  if (c->peer.transport == 1 &&
      (s = peer_reserve_sync(c, 2*((size_t)m.face[0].n + 3*(size_t)m.face[0].total), 2*((size_t)m.face[1].n + 3*(size_t)m.face[1].total), &changed))) return s;
  int agree[2] = {(c->peer.transport != 1 || (!changed && c->peer.ready)) ? 1 : 0, -(int)std::min(1.0e9, std::ceil(vmax*1.0e6))};
  if ((s = comm_allreduce_min_host(c, agree, 2))) return s;
  const double vmax_all = -agree[1]*1.0e-6;
  const int64_t interval = std::min<int64_t>(8*(int64_t)m.sync_every, std::max<int64_t>(m.sync_every, (int64_t)std::floor(1.0/(vmax_all + 1.0e-3))));
  m.next_sync_iter = c->iter + (initial ? m.sync_every : interval);
  if (c->peer.transport == 1 && !agree[0]) {
    if ((s = peer_setup(c))) return s;
    if (!c->peer.ready) return hcg_fail(c, HCG_ERR_STATE, "peer transport: re-mapping the neighbours' buffers failed after a rebalance");
  }
  return HCG_OK;
}
