// Pre-inlet coupling (replaces helper/preInlet.cpp:255-397 of the reference: applyPreInletVelocityBoundary and
// applyPreInletParticleBoundary, MPI messages between the pre-inlet ranks and the main-domain ranks).
//
// Here the periodic pre-inlet is a second context (same GPU or another GPU of the box) that iterates on its own
// stream next to the main domain.  After each step
//   * the velocity of the pre-inlet's coupling nodes becomes the boundary velocity of the main domain's Zou-He
//     inlet nodes: gather kernel on the pre-inlet's stream -> (peer copy) -> scatter kernel on the main stream,
//     ordered by events, no host synchronisation;
//   * whole cells that have entered the hand-over slab are copied into free cell slots of the main domain
//     (host coordinated like the multi-GPU migration of csrc/multi.cu: bounding boxes -> host -> pack -> copy -> unpack).
#include "ctx.cuh"
#include <cmath>
#include <cstdlib>
#include <algorithm>

struct PreInletState {
  hcg_ctx* pre = nullptr;
  int pre_device = 0;
  int64_t n = 0;
  int64_t* d_src_idx = nullptr;     // pre-inlet device
  int64_t* d_dst_idx = nullptr;     // main device
  double* buf_src = nullptr;        // pre-inlet device, [n][4]
  double* buf_dst = nullptr;        // main device (== buf_src on one device)
  cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
  bool first = true;
  std::vector<int64_t> last_lap;    // per pre-inlet cell slot: periodic image handed over last (INT64_MIN = none)
  int64_t handed = 0;
};

namespace {

struct Arr12 { double* a[12]; };

inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1)/t); }

// one CTA per listed cell: buf[12*off[i] + a*V + v] = arr[a][base(cell) + v]
__global__ void k_pre_pack(const int32_t* __restrict__ slots, const int64_t* __restrict__ off, int n,
                           const int64_t* __restrict__ cell_base, Arr12 arr, double* __restrict__ buf) {
  const int i = blockIdx.x;
  if (i >= n) return;
  const int64_t b = cell_base[slots[i]];
  const int V = (int)(off[i+1] - off[i]);
  double* o = buf + 12*off[i];
  for (int k = threadIdx.x; k < 12*V; k += blockDim.x) { const int a = k / V, v = k - a*V; o[k] = arr.a[a][b + v]; }
}
// ... and back into the slots of the receiving context, positions shifted into its coordinates
__global__ void k_pre_unpack(const int32_t* __restrict__ slots, const int64_t* __restrict__ off, int n,
                             const int64_t* __restrict__ cell_base, Arr12 arr, const double* __restrict__ buf,
                             const double* __restrict__ shift /* [n][3] */, uint8_t* alive) {
  const int i = blockIdx.x;
  if (i >= n) return;
  const int64_t b = cell_base[slots[i]];
  const int V = (int)(off[i+1] - off[i]);
  const double* in = buf + 12*off[i];
  for (int k = threadIdx.x; k < 12*V; k += blockDim.x) {
    const int a = k / V, v = k - a*V;
    arr.a[a][b + v] = a < 3 ? in[k] + shift[3*i + a] : in[k];
  }
  if (threadIdx.x == 0) alive[slots[i]] = 1;
}

Arr12 arrays_of(hcg_ctx* c) {
  Arr12 r;
  for (int k = 0; k < 3; k++) { r.a[k] = c->pos[k]; r.a[3+k] = c->vel[k]; r.a[6+k] = c->frc[k]; r.a[9+k] = c->frep[k]; }
  return r;
}

}  // namespace

// pure host logic of the hand-over (CPU-tested against the oracle's rule, tests/test_preinlet_oracle.py): cell i is taken when it is
// alive, the periodic image k = ceil((slab_lo - (lo + shift)) / period) of its extent [lo, hi] + shift lies wholly inside
// [slab_lo, slab_hi], and that image was not the one handed over last (last_lap)
extern "C" void hch_preinlet_select(int64_t n, const double* lo, const double* hi, const uint8_t* alive, const int64_t* last_lap,
                                    double shift, double period, double slab_lo, double slab_hi, int64_t* lap_out, uint8_t* take_out) {
  for (int64_t i = 0; i < n; i++) {
    take_out[i] = 0; lap_out[i] = 0;
    if (!alive[i]) continue;
    const double a = lo[i] + shift, b = hi[i] + shift;
    const double k = std::ceil((slab_lo - a)/period);
    if (b + k*period > slab_hi) continue;
    lap_out[i] = (int64_t)k;
    if (last_lap && last_lap[i] == (int64_t)k) continue;
    take_out[i] = 1;
  }
}

void preinlet_destroy(hcg_ctx* c) {
  PreInletState* p = c->preinlet;
  if (!p) return;
  cudaSetDevice(p->pre_device); cudaFree(p->d_src_idx); if (p->buf_src != p->buf_dst) cudaFree(p->buf_src);   // (the pre-inlet context itself may be gone already)
  if (p->ev_ready) cudaEventDestroy(p->ev_ready);
  cudaSetDevice(c->dom.device);
  cudaFree(p->d_dst_idx); cudaFree(p->buf_dst);
  if (p->ev_done) cudaEventDestroy(p->ev_done);
  delete p;
  c->preinlet = nullptr;
}

extern "C" {

hcg_status hcg_preinlet_map(hcg_ctx* c, hcg_ctx* pre, int64_t n, const int64_t* pre_idx, const int64_t* main_idx) {
  if (!c || !pre || c == pre || n < 0 || (n > 0 && (!pre_idx || !main_idx))) return HCG_ERR_ARG;
  if (c->dom.n_ranks != 1 || pre->dom.n_ranks != 1)
    return hcg_fail(c, HCG_ERR_STATE, "pre-inlet coupling: both domains must be single-rank contexts");
  for (int64_t k = 0; k < n; k++) {
    if (pre_idx[k] < 0 || pre_idx[k] >= pre->Nl || main_idx[k] < 0 || main_idx[k] >= c->Nl)
      return hcg_fail(c, HCG_ERR_ARG, "pre-inlet coupling: node index outside the lattice");
  }
  preinlet_destroy(c);
  PreInletState* p = new PreInletState();
  c->preinlet = p;
  p->pre = pre; p->pre_device = pre->dom.device; p->n = n;
  const bool same = pre->dom.device == c->dom.device;
  CUDA_TRY(c, cudaSetDevice(pre->dom.device));
  if (n) {
    CUDA_TRY(c, cudaMalloc(&p->d_src_idx, sizeof(int64_t)*n));
    CUDA_TRY(c, hcg_h2d(c, p->d_src_idx, pre_idx, sizeof(int64_t)*n));
    CUDA_TRY(c, cudaMalloc(&p->buf_src, sizeof(double)*4*n));
  }
  CUDA_TRY(c, cudaEventCreateWithFlags(&p->ev_ready, cudaEventDisableTiming));   // recorded on the pre-inlet's stream: lives on its device
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  if (n) {
    CUDA_TRY(c, cudaMalloc(&p->d_dst_idx, sizeof(int64_t)*n));
    CUDA_TRY(c, hcg_h2d(c, p->d_dst_idx, main_idx, sizeof(int64_t)*n));
    if (same) p->buf_dst = p->buf_src; else CUDA_TRY(c, cudaMalloc(&p->buf_dst, sizeof(double)*4*n));
  }
  CUDA_TRY(c, cudaEventCreateWithFlags(&p->ev_done, cudaEventDisableTiming));    // recorded on the main stream
  return lat_bcn_ensure(c);
}

hcg_status hcg_preinlet_apply_velocity(hcg_ctx* c) {
  if (!c) return HCG_ERR_ARG;
  PreInletState* p = c->preinlet;
  if (!p || !p->pre) return hcg_fail(c, HCG_ERR_STATE, "hcg_preinlet_map has not been called");
  if (p->n == 0) return HCG_OK;
  hcg_ctx* pre = p->pre;
  hcg_status s;
  CUDA_TRY(c, cudaSetDevice(pre->dom.device));
  if (!p->first) CUDA_TRY(c, cudaStreamWaitEvent(pre->stream, p->ev_done, 0));   // the previous scatter has consumed the buffer
  if ((s = lat_node_velocity(pre, p->n, p->d_src_idx, p->buf_src, pre->stream))) return hcg_fail(c, s, pre->err);
  CUDA_TRY(c, cudaEventRecord(p->ev_ready, pre->stream));
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  CUDA_TRY(c, cudaStreamWaitEvent(c->stream, p->ev_ready, 0));
  if (p->buf_dst != p->buf_src)
    CUDA_TRY(c, cudaMemcpyAsync(p->buf_dst, p->buf_src, sizeof(double)*4*p->n, cudaMemcpyDefault, c->stream));
  if ((s = lat_bcn_scatter(c, p->n, p->d_dst_idx, p->buf_dst, true, c->stream))) return s;
  CUDA_TRY(c, cudaEventRecord(p->ev_done, c->stream));
  p->first = false;
  return HCG_OK;
}

hcg_status hcg_preinlet_apply_cells(hcg_ctx* c, int32_t axis, double period, const double shift[3],
                                    double slab_lo, double slab_hi, int64_t id_stride, int64_t* n_added) {
  if (n_added) *n_added = 0;
  if (!c || axis < 0 || axis > 2 || !(period > 0) || !shift || !(slab_hi > slab_lo)) return HCG_ERR_ARG;
  PreInletState* p = c->preinlet;
  if (!p || !p->pre) return hcg_fail(c, HCG_ERR_STATE, "hcg_preinlet_map has not been called");
  hcg_ctx* pre = p->pre;
  if (pre->types.size() != c->types.size()) return hcg_fail(c, HCG_ERR_STATE, "pre-inlet hand-over: the two domains must register the same cell types");
  for (size_t t = 0; t < c->types.size(); t++)
    if (pre->types[t].d.V != c->types[t].d.V) return hcg_fail(c, HCG_ERR_STATE, "pre-inlet hand-over: cell types differ");
  const int64_t npc = pre->ncells;
  if (npc == 0) return HCG_OK;
  hcg_status s;
  // 1. bounding boxes + alive flags of the pre-inlet's cells
  CUDA_TRY(c, cudaSetDevice(pre->dom.device));
  std::vector<double> bbox(6*(size_t)npc);
  std::vector<uint8_t> alive((size_t)npc);
  {
    double* d_bbox;
    CUDA_TRY(c, cudaMalloc(&d_bbox, sizeof(double)*6*npc));
    if ((s = mech_bbox(pre, d_bbox))) { cudaFree(d_bbox); return hcg_fail(c, s, pre->err); }
    CUDA_TRY(c, cudaStreamSynchronize(pre->stream));
    CUDA_TRY(c, cudaMemcpy(bbox.data(), d_bbox, sizeof(double)*6*npc, cudaMemcpyDeviceToHost));
    CUDA_TRY(c, cudaMemcpy(alive.data(), pre->cell_alive, npc, cudaMemcpyDeviceToHost));
    cudaFree(d_bbox);
  }
  if ((int64_t)p->last_lap.size() < npc) p->last_lap.resize(npc, INT64_MIN);
  // 2. candidates: the periodic image k of the cell lies wholly inside [slab_lo, slab_hi] (main coordinates, along `axis`)
  std::vector<int32_t> src; std::vector<int64_t> lap;
  {
    std::vector<double> lo(npc), hi(npc); std::vector<uint8_t> live(npc), take(npc); std::vector<int64_t> k(npc);
    for (int64_t i = 0; i < npc; i++) { lo[i] = bbox[6*i + 2*axis]; hi[i] = bbox[6*i + 2*axis + 1]; live[i] = alive[i] && pre->h_cell_id[i] >= 0; }
    hch_preinlet_select(npc, lo.data(), hi.data(), live.data(), p->last_lap.data(), shift[axis], period, slab_lo, slab_hi, k.data(), take.data());
    for (int64_t i = 0; i < npc; i++) if (take[i]) { src.push_back((int32_t)i); lap.push_back(k[i]); }
  }
  if (src.empty()) return HCG_OK;
  // 3. free slots in the main domain: spare slots first, then slots of cells that have been deleted
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  std::vector<uint8_t> main_alive;
  c->multi.free_slots.resize(c->types.size());
  std::vector<int32_t> dst; std::vector<int64_t> off(1, 0); std::vector<double> sh;
  std::vector<int32_t> src_ok; std::vector<int64_t> lap_ok;
  for (size_t q = 0; q < src.size(); q++) {
    const int t = pre->h_cell_type[src[q]];
    CellTypeHost& th = c->types[t];
    int32_t slot = -1;
    if (!c->multi.free_slots[t].empty()) { slot = c->multi.free_slots[t].back(); c->multi.free_slots[t].pop_back(); }
    else if (th.n_cells < th.cap_cells) slot = (int32_t)(th.first_cell + th.n_cells++);
    else {
      if (main_alive.empty() && c->ncells) {
        main_alive.resize(c->ncells);
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        CUDA_TRY(c, cudaMemcpy(main_alive.data(), c->cell_alive, c->ncells, cudaMemcpyDeviceToHost));
      }
      for (int64_t i = th.first_cell; i < th.first_cell + th.n_cells; i++)
        if (!main_alive[i]) { slot = (int32_t)i; main_alive[i] = 1; break; }
    }
    if (slot < 0) return hcg_fail(c, HCG_ERR_CAPACITY, "pre-inlet hand-over: no free cell slot in the main domain (hcg_cells_reserve)");
    dst.push_back(slot); src_ok.push_back(src[q]); lap_ok.push_back(lap[q]);
    off.push_back(off.back() + th.d.V);
    for (int d = 0; d < 3; d++) sh.push_back(shift[d] + (d == axis ? lap[q]*period : 0.0));
    c->h_cell_id[slot] = pre->h_cell_id[src[q]] + (2*std::llabs(lap[q]) - (lap[q] < 0 ? 1 : 0))*id_stride;   // images 0, -1, 1, -2, 2, ... -> 0, 1, 2, 3, 4, ...: unique and non-negative (-1 marks a free slot)
    p->last_lap[src[q]] = lap[q];
  }
  const int n = (int)dst.size();
  const size_t nd = 12*(size_t)off.back();
  // 4. pack on the pre-inlet, copy, unpack (positions shifted) on the main domain
  CUDA_TRY(c, cudaSetDevice(pre->dom.device));
  int32_t* d_src; int64_t* d_off_s; double* d_buf_s;
  CUDA_TRY(c, cudaMalloc(&d_src, sizeof(int32_t)*n)); CUDA_TRY(c, cudaMalloc(&d_off_s, sizeof(int64_t)*(n + 1)));
  CUDA_TRY(c, cudaMalloc(&d_buf_s, sizeof(double)*nd));
  CUDA_TRY(c, cudaMemcpyAsync(d_src, src_ok.data(), sizeof(int32_t)*n, cudaMemcpyHostToDevice, pre->stream));
  CUDA_TRY(c, cudaMemcpyAsync(d_off_s, off.data(), sizeof(int64_t)*(n + 1), cudaMemcpyHostToDevice, pre->stream));
  k_pre_pack<<<n, 256, 0, pre->stream>>>(d_src, d_off_s, n, pre->cell_base, arrays_of(pre), d_buf_s);
  pre->launches++;
  CUDA_TRY(c, cudaGetLastError());
  CUDA_TRY(c, cudaStreamSynchronize(pre->stream));
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  int32_t* d_dst; int64_t* d_off_d; double* d_buf_d; double* d_sh;
  CUDA_TRY(c, cudaMalloc(&d_dst, sizeof(int32_t)*n)); CUDA_TRY(c, cudaMalloc(&d_off_d, sizeof(int64_t)*(n + 1)));
  CUDA_TRY(c, cudaMalloc(&d_buf_d, sizeof(double)*nd)); CUDA_TRY(c, cudaMalloc(&d_sh, sizeof(double)*3*n));
  CUDA_TRY(c, cudaMemcpyAsync(d_dst, dst.data(), sizeof(int32_t)*n, cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(c, cudaMemcpyAsync(d_off_d, off.data(), sizeof(int64_t)*(n + 1), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(c, cudaMemcpyAsync(d_sh, sh.data(), sizeof(double)*3*n, cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(c, cudaMemcpyAsync(d_buf_d, d_buf_s, sizeof(double)*nd, cudaMemcpyDefault, c->stream));
  k_pre_unpack<<<n, 256, 0, c->stream>>>(d_dst, d_off_d, n, c->cell_base, arrays_of(c), d_buf_d, d_sh, c->cell_alive);
  KERNEL_CHECK(c);
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  cudaFree(d_dst); cudaFree(d_off_d); cudaFree(d_buf_d); cudaFree(d_sh);
  CUDA_TRY(c, cudaSetDevice(pre->dom.device));
  cudaFree(d_src); cudaFree(d_off_s); cudaFree(d_buf_s);
  CUDA_TRY(c, cudaSetDevice(c->dom.device));
  c->perm_valid = false; c->cell_gid_dirty = true; c->far_steps_left = 0;
  p->handed += n;
  if (n_added) *n_added = n;
  return HCG_OK;
}

/* hand-over bookkeeping for checkpoints: the periodic image handed over last per pre-inlet cell slot (INT64_MIN = none);
 * set != 0 writes, else reads; n = the pre-inlet's cell capacity */
hcg_status hcg_preinlet_laps(hcg_ctx* c, int64_t n, int64_t* laps, int32_t set) {
  if (!c || n < 0 || (n > 0 && !laps)) return HCG_ERR_ARG;
  PreInletState* p = c->preinlet;
  if (!p || !p->pre) return hcg_fail(c, HCG_ERR_STATE, "hcg_preinlet_map has not been called");
  if ((int64_t)p->last_lap.size() < n) p->last_lap.resize(n, INT64_MIN);
  for (int64_t i = 0; i < n; i++) { if (set) p->last_lap[i] = laps[i]; else laps[i] = p->last_lap[i]; }
  return HCG_OK;
}

}  // extern "C"
