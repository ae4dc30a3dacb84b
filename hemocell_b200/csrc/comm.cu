// Point-to-point / reduction plumbing between the slab contexts of one run.
//
// Two back ends behind one set of calls (comm_group_begin / comm_send / comm_recv / comm_group_end,
// comm_allreduce_*):
//   * NCCL (hcg_comm_init): one process or thread per GPU, grouped ncclSend / ncclRecv on the context's stream;
//   * host-staged (hcg_comm_init_local): messages travel as files in a memory-backed directory (/dev/shm), named
//     by (run key, source, destination, sequence number), so they are matched FIFO per (source, destination) pair
//     exactly as NCCL matches grouped send/recv.  It works between threads of one process and between processes,
//     and - unlike NCCL - with several ranks on ONE GPU, so a slab-decomposed run (the reference's `mpirun -n 2`
//     vs `-n 4` identity check, scripts/ci/pipeflow_sanity.sh:25-32) can be verified on a single-GPU box.  It is
//     host-synchronous and meant for verification; it carries the same call sites as NCCL (set-up blobs,
//     migration, reductions and - with transport 0 - the halo planes).  The peer-store transport (kernels storing
//     into the neighbour, flag barrier) works unchanged on top of it: neighbours on the same GPU are mapped as
//     plain pointers (same process) or through CUDA IPC (other process).
// The reference moves the same data with MPI messages (Palabos block communicator; HemoCellFields::syncEnvelopes,
// core/hemoCellFields.cpp:377-499).
#include "ctx.cuh"
#include <nccl.h>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <sys/stat.h>
#include <unistd.h>

namespace {

struct LocalOp { void* ptr; size_t bytes; int peer; };
struct LocalEndpoint {
  std::string prefix;                       // <dir>/hcg_<key>_
  std::vector<uint64_t> seq_out, seq_in;    // per peer
  std::vector<LocalOp> sends, recvs;
  std::vector<unsigned char> host;
};

constexpr double kLocalTimeoutS = 120.0;

std::string msg_name(const LocalEndpoint& ep, int src, int dst, uint64_t seq) {
  return ep.prefix + std::to_string(src) + "_" + std::to_string(dst) + "_" + std::to_string(seq);
}

hcg_status local_put(hcg_ctx* c, LocalEndpoint& ep, int peer, const void* data, size_t bytes, const char* what) {
  const std::string name = msg_name(ep, c->dom.rank, peer, ep.seq_out[peer]++), tmp = name + ".part";
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) return hcg_fail(c, HCG_ERR_NCCL, std::string(what) + ": cannot create " + tmp);
  const bool ok = bytes == 0 || fwrite(data, 1, bytes, f) == bytes;
  fclose(f);
  if (getenv("HCG_COMM_DEBUG")) fprintf(stderr, "[comm %d pid %d] put %s bytes %zu first %d %d %d %d\n", c->dom.rank, (int)getpid(), name.c_str(), bytes, bytes >= 16 ? ((const int*)data)[0] : -1, bytes >= 16 ? ((const int*)data)[1] : -1, bytes >= 16 ? ((const int*)data)[2] : -1, bytes >= 16 ? ((const int*)data)[3] : -1);
  if (!ok || rename(tmp.c_str(), name.c_str()) != 0) { unlink(tmp.c_str()); return hcg_fail(c, HCG_ERR_NCCL, std::string(what) + ": cannot write " + name); }
  return HCG_OK;
}

hcg_status local_get(hcg_ctx* c, LocalEndpoint& ep, int peer, void* data, size_t bytes, const char* what) {
  const std::string name = msg_name(ep, peer, c->dom.rank, ep.seq_in[peer]++);
  const auto t0 = std::chrono::steady_clock::now();
  struct stat st;
  for (unsigned spin = 0; stat(name.c_str(), &st) != 0; spin++) {
    if (spin < 2000) std::this_thread::yield(); else std::this_thread::sleep_for(std::chrono::microseconds(200));
    if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > kLocalTimeoutS)
      return hcg_fail(c, HCG_ERR_NCCL, std::string(what) + ": host-staged exchange timed out waiting for rank " + std::to_string(peer));
  }
  if ((size_t)st.st_size != bytes) { unlink(name.c_str()); return hcg_fail(c, HCG_ERR_STATE, std::string(what) + ": host-staged exchange: message size mismatch"); }
  FILE* f = fopen(name.c_str(), "rb");
  if (!f) return hcg_fail(c, HCG_ERR_NCCL, std::string(what) + ": cannot open " + name);
  const bool ok = bytes == 0 || fread(data, 1, bytes, f) == bytes;
  fclose(f); unlink(name.c_str());
  if (getenv("HCG_COMM_DEBUG")) fprintf(stderr, "[comm %d pid %d] got %s bytes %zu first %d %d %d %d\n", c->dom.rank, (int)getpid(), name.c_str(), bytes, bytes >= 16 ? ((const int*)data)[0] : -1, bytes >= 16 ? ((const int*)data)[1] : -1, bytes >= 16 ? ((const int*)data)[2] : -1, bytes >= 16 ? ((const int*)data)[3] : -1);
  if (!ok) return hcg_fail(c, HCG_ERR_NCCL, std::string(what) + ": short read of " + name);
  return HCG_OK;
}

hcg_status local_group_end(hcg_ctx* c, const char* what) {
  LocalEndpoint& ep = *(LocalEndpoint*)c->local;
  hcg_status rc = HCG_OK;
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));               // my send buffers are final
  for (auto& s : ep.sends) {
    if (ep.host.size() < s.bytes) ep.host.resize(s.bytes);
    CUDA_TRY(c, cudaMemcpyAsync(ep.host.data(), s.ptr, s.bytes, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if ((rc = local_put(c, ep, s.peer, ep.host.data(), s.bytes, what))) break;
  }
  if (!rc) for (auto& q : ep.recvs) {
    if (ep.host.size() < q.bytes) ep.host.resize(q.bytes);
    if ((rc = local_get(c, ep, q.peer, ep.host.data(), q.bytes, what))) break;
    // on the context's stream and waited for: a plain cudaMemcpy from pageable memory may return before the DMA has landed,
    // and the (non-blocking) stream that reads the buffer next is not ordered behind the legacy default stream
    CUDA_TRY(c, cudaMemcpyAsync(q.ptr, ep.host.data(), q.bytes, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  }
  ep.sends.clear(); ep.recvs.clear();
  return rc;
}

// every rank contributes n doubles; all ranks combine the contributions in rank order (bit-identical results)
hcg_status local_allreduce(hcg_ctx* c, std::vector<double>& v, int op) {
  LocalEndpoint& ep = *(LocalEndpoint*)c->local;
  const int R = c->dom.n_ranks, r = c->dom.rank;
  hcg_status s;
  for (int p = 0; p < R; p++) if (p != r && (s = local_put(c, ep, p, v.data(), sizeof(double)*v.size(), "allreduce"))) return s;
  std::vector<double> out, in(v.size());
  for (int p = 0; p < R; p++) {
    const std::vector<double>* src = &v;
    if (p != r) { if ((s = local_get(c, ep, p, in.data(), sizeof(double)*in.size(), "allreduce"))) return s; src = &in; }
    if (p == 0) { out = *src; continue; }
    for (size_t i = 0; i < out.size(); i++) {
      const double a = out[i], b = (*src)[i];
      out[i] = op == 0 ? a + b : (op == 1 ? (b < a ? b : a) : (b > a ? b : a));
    }
  }
  v = out;
  return HCG_OK;
}

}  // namespace

bool comm_up(const hcg_ctx* c) { return c->nccl != nullptr || c->local != nullptr; }

hcg_status comm_group_begin(hcg_ctx* c) {
  if (c->local) { LocalEndpoint* ep = (LocalEndpoint*)c->local; ep->sends.clear(); ep->recvs.clear(); return HCG_OK; }
  if (!c->nccl) return hcg_fail(c, HCG_ERR_STATE, "n_ranks > 1 but hcg_comm_init was not called");
  ncclGroupStart();
  return HCG_OK;
}
void comm_send(hcg_ctx* c, const void* p, size_t bytes, int peer) {
  if (c->local) { ((LocalEndpoint*)c->local)->sends.push_back({const_cast<void*>(p), bytes, peer}); return; }
  ncclSend(p, bytes, ncclUint8, peer, (ncclComm_t)c->nccl, c->stream);
}
void comm_recv(hcg_ctx* c, void* p, size_t bytes, int peer) {
  if (c->local) { ((LocalEndpoint*)c->local)->recvs.push_back({p, bytes, peer}); return; }
  ncclRecv(p, bytes, ncclUint8, peer, (ncclComm_t)c->nccl, c->stream);
}
hcg_status comm_group_end(hcg_ctx* c, const char* what) {
  if (c->local) return local_group_end(c, what);
  ncclResult_t rc = ncclGroupEnd();
  if (rc != ncclSuccess) return hcg_fail(c, HCG_ERR_NCCL, std::string(what) + ": " + ncclGetErrorString(rc));
  return HCG_OK;
}

// in place on a device buffer, ordered on the context's stream; op 0 sum, 1 min, 2 max
hcg_status comm_allreduce_f64(hcg_ctx* c, double* dev, size_t n, int op) {
  if (c->dom.n_ranks == 1 || n == 0) return HCG_OK;
  if (c->local) {
    std::vector<double> v(n);
    CUDA_TRY(c, cudaMemcpyAsync(v.data(), dev, sizeof(double)*n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    hcg_status s = local_allreduce(c, v, op); if (s) return s;
    CUDA_TRY(c, cudaMemcpyAsync(dev, v.data(), sizeof(double)*n, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return HCG_OK;
  }
  if (!c->nccl) return hcg_fail(c, HCG_ERR_STATE, "allreduce: hcg_comm_init first");
  const ncclRedOp_t ops[3] = {ncclSum, ncclMin, ncclMax};
  ncclResult_t rc = ncclAllReduce(dev, dev, n, ncclDouble, ops[op], (ncclComm_t)c->nccl, c->stream);
  if (rc != ncclSuccess) return hcg_fail(c, HCG_ERR_NCCL, std::string("ncclAllReduce: ") + ncclGetErrorString(rc));
  return HCG_OK;
}

// host values, element-wise minimum over the ranks (set-up agreement, rebalance decisions)
hcg_status comm_allreduce_min_host(hcg_ctx* c, int* value, int n) {
  if (c->dom.n_ranks == 1 || n <= 0) return HCG_OK;
  if (c->local) {
    std::vector<double> v(value, value + n);
    hcg_status s = local_allreduce(c, v, 1); if (s) return s;
    for (int k = 0; k < n; k++) value[k] = (int)v[k];
    return HCG_OK;
  }
  if (!c->nccl) return hcg_fail(c, HCG_ERR_STATE, "allreduce: hcg_comm_init first");
  if (!c->comm_scratch) CUDA_TRY(c, cudaMalloc(&c->comm_scratch, sizeof(int)*16));
  if (n > 16) return hcg_fail(c, HCG_ERR_ARG, "allreduce: at most 16 values");
  int* d = c->comm_scratch;
  CUDA_TRY(c, cudaMemcpyAsync(d, value, sizeof(int)*n, cudaMemcpyHostToDevice, c->stream));
  ncclResult_t rc = ncclAllReduce(d, d, n, ncclInt, ncclMin, (ncclComm_t)c->nccl, c->stream);
  if (rc != ncclSuccess) return hcg_fail(c, HCG_ERR_NCCL, std::string("ncclAllReduce: ") + ncclGetErrorString(rc));
  CUDA_TRY(c, cudaMemcpyAsync(value, d, sizeof(int)*n, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return HCG_OK;
}

hcg_status comm_nccl_init(hcg_ctx* c, const void* id128) {
  ncclUniqueId id; memcpy(&id, id128, 128);
  ncclComm_t comm;
  ncclResult_t rc = ncclCommInitRank(&comm, c->dom.n_ranks, id, c->dom.rank);
  if (rc != ncclSuccess) return hcg_fail(c, HCG_ERR_NCCL, std::string("ncclCommInitRank: ") + ncclGetErrorString(rc));
  c->nccl = comm;
  return HCG_OK;
}

// join the run named by id128 (any 128 bytes unique to it); returns when every rank has answered
hcg_status comm_local_init(hcg_ctx* c, const void* id128) {
  auto* ep = new LocalEndpoint();
  uint64_t h = 1469598103934665603ULL;
  for (int i = 0; i < 128; i++) { h ^= ((const unsigned char*)id128)[i]; h *= 1099511628211ULL; }
  char key[32]; snprintf(key, sizeof(key), "%016llx", (unsigned long long)h);
  const char* dir = getenv("HCG_COMM_DIR");
  struct stat st;
  std::string d = dir ? dir : (stat("/dev/shm", &st) == 0 ? "/dev/shm" : "/tmp");
  ep->prefix = d + "/hcg_" + key + "_";
  ep->seq_out.assign(c->dom.n_ranks, 0); ep->seq_in.assign(c->dom.n_ranks, 0);
  c->local = ep;
  std::vector<double> hello(1, 1.0);
  hcg_status s = local_allreduce(c, hello, 0); if (s) return s;
  if ((int)hello[0] != c->dom.n_ranks) return hcg_fail(c, HCG_ERR_STATE, "hcg_comm_init_local: handshake failed");
  return HCG_OK;
}

void comm_destroy(hcg_ctx* c) {
  if (c->comm_scratch) { cudaFree(c->comm_scratch); c->comm_scratch = nullptr; }
  if (c->nccl) { ncclCommDestroy((ncclComm_t)c->nccl); c->nccl = nullptr; }
  if (c->local) { delete (LocalEndpoint*)c->local; c->local = nullptr; }
}
