// NVLink / NVSwitch peer-memory transport for the x-slab decomposition.
//
// The reference moves ghost layers and particle envelopes with MPI messages
// (Palabos block communicator; HemoCellFields::syncEnvelopes, core/hemoCellFields.cpp:377-499).
// Here every context maps its two slab neighbours' buffers once (CUDA IPC between processes,
// cudaDeviceEnablePeerAccess inside one process) and from then on
//   * k_collide_stream stores the 5 outgoing populations of its face planes straight into the
//     neighbour's ghost plane, k_moments does the same with the node velocity,
//   * k_pack_sync stores the shared cells' velocities into the neighbour's receive buffer,
//   * k_peer_barrier (one CTA) publishes an epoch word in the neighbour's memory with
//     st.release.sys and spins with ld.acquire.sys on its own words.
// No send/recv kernels, no staging copies; NCCL is only used to exchange the mapping blobs (and for
// the host-coordinated migration every `sync_every` steps).
#include "ctx.cuh"
#include <cstring>
#include <cstdlib>
#include <unistd.h>

namespace {

struct PeerBlob {
  int32_t pid, device, rank, pad;
  uint64_t host;                                   // boot-unique host tag (same box check)
  uint64_t raw[HCG_PEER_NPTR];
  unsigned char handle[HCG_PEER_NPTR][64];
  uint8_t valid[HCG_PEER_NPTR]; uint8_t pad2[8 - HCG_PEER_NPTR % 8];
};
static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// thread 0 / 1 handle the left / right neighbour: everything this GPU stored before the launch is
// made visible system-wide, the epoch goes into the neighbour's word, then wait for the neighbour's.
__global__ void k_peer_barrier(unsigned long long* left_word, unsigned long long* right_word,
                               const unsigned long long* mine, unsigned long long epoch) {
  const int t = threadIdx.x;
  if (t > 1) return;
  unsigned long long* out = t == 0 ? left_word : right_word;
  if (!out) return;
  __threadfence_system();
  st_release_sys(out, epoch);
  long long t_start = 0;
  for (unsigned spin = 0; ld_acquire_sys(mine + t) < epoch; spin++) {
    if ((spin & 1023u) == 1023u) {                 // a lost neighbour must not hang the device for ever
      const long long now = clock64();
      if (t_start == 0) t_start = now; else if (now - t_start > 120000000000LL) __trap();   // ~60 s
    }
  }
}

// the two halves of k_peer_barrier as separate launches, so that independent work can sit between "my stores are published"
// and "the neighbours' stores have arrived"
__global__ void k_peer_post(unsigned long long* left_word, unsigned long long* right_word, unsigned long long epoch) {
  const int t = threadIdx.x;
  if (t > 1) return;
  unsigned long long* out = t == 0 ? left_word : right_word;
  if (!out) return;
  __threadfence_system();
  st_release_sys(out, epoch);
}
__global__ void k_peer_wait(bool has_left, bool has_right, const unsigned long long* mine, unsigned long long epoch) {
  const int t = threadIdx.x;
  if (t > 1 || !(t == 0 ? has_left : has_right)) return;
  long long t_start = 0;
  for (unsigned spin = 0; ld_acquire_sys(mine + t) < epoch; spin++) {
    if ((spin & 1023u) == 1023u) {
      const long long now = clock64();
      if (t_start == 0) t_start = now; else if (now - t_start > 120000000000LL) __trap();   // ~60 s: a lost neighbour must not hang the device
    }
  }
}

hcg_status map_pointer(hcg_ctx* c, const PeerBlob& b, int k, void** out) {
  *out = nullptr;
  if (!b.valid[k]) return HCG_OK;
  if (b.pid == (int32_t)getpid()) {                // same process (one thread per GPU): plain peer access
    if (b.device != c->dom.device) {
      int can = 0;
      CUDA_TRY(c, cudaDeviceCanAccessPeer(&can, c->dom.device, b.device));
      if (!can) return hcg_fail(c, HCG_ERR_CUDA, "peer transport: no P2P access between the slab neighbours' GPUs (select the NCCL transport)");
      cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CUDA_TRY(c, e);
      cudaGetLastError();
    }
    *out = (void*)b.raw[k];
    return HCG_OK;
  }
  for (auto& m : c->peer.maps) if (!memcmp(m.handle, b.handle[k], 64)) { *out = m.base; return HCG_OK; }
  cudaIpcMemHandle_t h; memcpy(&h, b.handle[k], 64);
  void* base = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return hcg_fail(c, HCG_ERR_CUDA, std::string("peer transport: cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
  PeerMap m; memcpy(m.handle, b.handle[k], 64); m.base = base;
  c->peer.maps.push_back(m);
  *out = base;
  return HCG_OK;
}

uint64_t host_tag() {
  char name[256] = {0};
  gethostname(name, sizeof(name) - 1);
  uint64_t h = 1469598103934665603ULL;
  for (const char* p = name; *p; p++) { h ^= (unsigned char)*p; h *= 1099511628211ULL; }
  return h;
}

}  // namespace

hcg_status peer_reserve_sync(hcg_ctx* c, size_t left, size_t right, bool* changed) {
  const size_t want[2] = {left, right};
  if (changed) *changed = false;
  for (int f = 0; f < 2; f++) {
    if (c->peer.sync_recv_cap[f] >= want[f] && c->peer.sync_recv[f]) continue;
    // the neighbour may still hold a mapping of the old buffer: keep it allocated until destroy;
    // the new one is published by the caller (peer_setup)
    if (c->peer.sync_recv[f]) c->peer.retired.push_back(c->peer.sync_recv[f]);
    c->peer.sync_recv[f] = nullptr; c->peer.sync_recv_cap[f] = 0;
    const size_t cap = want[f] + want[f]/2 + 4096;
    CUDA_TRY(c, cudaMalloc(&c->peer.sync_recv[f], sizeof(double)*cap));
    c->peer.sync_recv_cap[f] = cap;
    if (changed) *changed = true;
  }
  return HCG_OK;
}

// Export my buffers, swap blobs with both slab neighbours, map theirs.  Collective over neighbours.
hcg_status peer_setup(hcg_ctx* c) {
  PeerState& p = c->peer;
  p.ready = false;
  if (c->dom.n_ranks == 1 || p.transport != 1) return HCG_OK;
  if (!comm_up(c)) return hcg_fail(c, HCG_ERR_STATE, "peer transport: hcg_comm_init first");
  const int R = c->dom.n_ranks, r = c->dom.rank; const bool px = c->dom.periodic[0];
  p.link[0].rank = (r == 0) ? (px ? R - 1 : -1) : r - 1;
  p.link[1].rank = (r == R - 1) ? (px ? 0 : -1) : r + 1;
  if (!p.flags) {
    CUDA_TRY(c, cudaMalloc(&p.flags, sizeof(unsigned long long)*8));
    CUDA_TRY(c, cudaMemset(p.flags, 0, sizeof(unsigned long long)*8));
  }
  hcg_status s;
  if (!p.sync_recv[0] || !p.sync_recv[1]) { if ((s = peer_reserve_sync(c, 0, 0, nullptr))) return s; }
  PeerBlob mine; memset(&mine, 0, sizeof(mine));
  mine.pid = (int32_t)getpid(); mine.device = c->dom.device; mine.rank = r; mine.host = host_tag();
  void* ptrs[HCG_PEER_NPTR] = {c->g[0], c->g[1], c->U, p.flags, p.sync_recv[0], p.sync_recv[1], c->Wphys[0], c->Wphys[1], c->Fphys[0], c->Fphys[1]};
  // Failures that only mean "no peer memory on this box" (IPC disabled in the container, no P2P path, neighbour on
  // another host) must not leave the neighbours hanging in the collective below: they are recorded in p.usable and
  // hcg_comm_init lets all ranks agree on a transport afterwards.
  p.usable = true;
  if (const char* e = getenv("HCG_PEER_SIMULATE_FAILURE")) if (atoi(e) == r + 1) p.usable = false;   // test hook: rank e-1 cannot use peer memory
  for (int k = 0; k < HCG_PEER_NPTR; k++) {
    if (!ptrs[k]) continue;
    mine.raw[k] = (uint64_t)ptrs[k]; mine.valid[k] = 1;
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, ptrs[k]) != cudaSuccess) { cudaGetLastError(); p.usable = false; mine.valid[k] = 0; continue; }
    memcpy(mine.handle[k], &h, 64);
  }
  // blobs travel through device memory (NCCL): [mine][from right][from left]
  if (!p.d_blob) CUDA_TRY(c, cudaMalloc(&p.d_blob, 3*sizeof(PeerBlob)));
  char* d = (char*)p.d_blob;
  CUDA_TRY(c, cudaMemcpyAsync(d, &mine, sizeof(mine), cudaMemcpyHostToDevice, c->stream));
  if ((s = multi_neighbour_exchange(c, d, sizeof(PeerBlob), d, sizeof(PeerBlob),
                                    d + sizeof(PeerBlob), sizeof(PeerBlob), d + 2*sizeof(PeerBlob), sizeof(PeerBlob)))) return s;
  PeerBlob got[2];                                 // [0] = from the left neighbour, [1] = from the right
  CUDA_TRY(c, cudaMemcpyAsync(&got[1], d + sizeof(PeerBlob), sizeof(PeerBlob), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaMemcpyAsync(&got[0], d + 2*sizeof(PeerBlob), sizeof(PeerBlob), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  for (int f = 0; f < 2; f++) {
    if (p.link[f].rank < 0) { for (auto& q : p.link[f].ptr) q = nullptr; continue; }
    if (got[f].rank != p.link[f].rank) return hcg_fail(c, HCG_ERR_STATE, "peer transport: neighbour blob from an unexpected rank");
    if (got[f].host != mine.host) { p.usable = false; continue; }
    for (int k = 0; k < HCG_PEER_NPTR; k++) {
      if (ptrs[k] && !got[f].valid[k]) { p.usable = false; continue; }
      if (map_pointer(c, got[f], k, &p.link[f].ptr[k]) != HCG_OK) { cudaGetLastError(); p.usable = false; }
    }
  }
  p.ready = p.usable;
  return HCG_OK;
}

hcg_status peer_barrier(hcg_ctx* c) {
  PeerState& p = c->peer;
  if (c->local) {
    // host-staged communicator (ranks may share a GPU): a spinning flag kernel would deadlock against any device-wide
    // synchronising call (cudaFree, cudaMalloc) of the rank it waits for, so the neighbours meet on the host instead:
    // my stream is drained (my peer stores have landed), then one token goes each way
    if (!p.d_blob) CUDA_TRY(c, cudaMalloc(&p.d_blob, 3*sizeof(PeerBlob)));
    char* d = (char*)p.d_blob;
    return multi_neighbour_exchange(c, d, 8, d + 8, 8, d + 16, 8, d + 24, 8);
  }
  p.epoch++;
  // my left neighbour watches its word [1] (written by ITS right neighbour = me), and vice versa
  unsigned long long* lw = p.link[0].rank >= 0 ? (unsigned long long*)p.link[0].ptr[3] + 1 : nullptr;
  unsigned long long* rw = p.link[1].rank >= 0 ? (unsigned long long*)p.link[1].ptr[3] + 0 : nullptr;
  k_peer_barrier<<<1, 32, 0, c->stream>>>(lw, rw, p.flags, p.epoch);
  KERNEL_CHECK(c);
  return HCG_OK;
}

// split barrier: peer_post publishes "everything before this is stored", peer_wait waits for both neighbours' posts
hcg_status peer_post(hcg_ctx* c) {
  PeerState& p = c->peer;
  if (c->local) return HCG_OK;                           // host-staged communicator: the whole barrier happens in peer_wait
  p.epoch++;
  unsigned long long* lw = p.link[0].rank >= 0 ? (unsigned long long*)p.link[0].ptr[3] + 1 : nullptr;
  unsigned long long* rw = p.link[1].rank >= 0 ? (unsigned long long*)p.link[1].ptr[3] + 0 : nullptr;
  k_peer_post<<<1, 32, 0, c->stream>>>(lw, rw, p.epoch);
  KERNEL_CHECK(c);
  return HCG_OK;
}
hcg_status peer_wait(hcg_ctx* c) {
  PeerState& p = c->peer;
  if (c->local) return peer_barrier(c);
  k_peer_wait<<<1, 32, 0, c->stream>>>(p.link[0].rank >= 0, p.link[1].rank >= 0, p.flags, p.epoch);
  KERNEL_CHECK(c);
  return HCG_OK;
}

void peer_destroy(hcg_ctx* c) {
  PeerState& p = c->peer;
  for (auto& m : p.maps) cudaIpcCloseMemHandle(m.base);
  p.maps.clear();
  if (p.flags) cudaFree(p.flags);
  for (int f = 0; f < 2; f++) if (p.sync_recv[f]) cudaFree(p.sync_recv[f]);
  if (p.d_blob) cudaFree(p.d_blob);
  for (void* q : p.retired) cudaFree(q);
  p.retired.clear();
  p.flags = nullptr; p.sync_recv[0] = p.sync_recv[1] = nullptr; p.d_blob = nullptr; p.ready = false;
}
