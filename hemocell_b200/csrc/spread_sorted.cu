// K2 (fast path): IBM force spreading with per-cell node-sorted (vertex, corner) pairs and a
// warp-level segmented reduction in front of the fp64 global RED.
//
// The plain kernel (k_spread, ibm.cu) issues 8 nodes x 3 components = 24 fp64 atomics per LSP and is
// bound by the L2 atomic rate (~165 G RED/s measured on B200).  Membrane vertices sit ~1 lu apart, so
// within one cell every lattice node is hit by ~3-4 (vertex, corner) pairs.  Here one CTA handles one
// cell: the 8*V pairs are visited in an order sorted by target node (the permutation is rebuilt every
// few steps by k_pair_sort and only needs to be *a* permutation for correctness - a stale order just
// merges less), runs of equal node are summed with shuffles, and only the run tails touch global
// memory.  Same arithmetic per pair as HemoCellParticleField::spreadParticleForce
// (reference core/hemoCellParticleField.cpp:841-863): force cap, phi2 weights in the reference's
// accumulation order for the normalisation, (frep + f) * w.
#include "ctx.cuh"
#include <cub/block/block_radix_sort.cuh>
#include <climits>
#include <cstdlib>

#include "spread_node.cuh"   // SpArgs, sp_local_x, sp_wrap, sp_phi2, sp_stage_vertex, sp_pair_node (host + device)

namespace {

// ---------------------------------------------------------------- permutation rebuild
template <int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS)
k_pair_sort(SpArgs a, const uint8_t* __restrict__ alive, const double* __restrict__ x, const double* __restrict__ y,
            const double* __restrict__ z, uint16_t* __restrict__ perm) {
  typedef cub::BlockRadixSort<uint32_t, THREADS, ITEMS, uint16_t> Sort;
  __shared__ typename Sort::TempStorage tmp;
  __shared__ int s_min[3];
  const int64_t cell = a.first_cell + blockIdx.x;
  if (!alive[cell]) return;
  const int64_t base = a.first_particle + (int64_t)blockIdx.x*a.V;
  const int npair = 8*a.V;
  if (threadIdx.x < 3) s_min[threadIdx.x] = INT_MAX;
  __syncthreads();
  int mn[3] = {INT_MAX, INT_MAX, INT_MAX};
  for (int v = threadIdx.x; v < a.V; v += THREADS) {
    mn[0] = min(mn[0], (int)floor(x[base+v])); mn[1] = min(mn[1], (int)floor(y[base+v])); mn[2] = min(mn[2], (int)floor(z[base+v]));
  }
  for (int d = 0; d < 3; d++) atomicMin(&s_min[d], mn[d]);
  __syncthreads();
  uint32_t keys[ITEMS]; uint16_t vals[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    const int e = threadIdx.x*ITEMS + i;
    if (e < npair) {
      const int v = e >> 3, c = e & 7;
      int rx = (int)floor(x[base+v]) + (c >> 2) - s_min[0];
      int ry = (int)floor(y[base+v]) + ((c >> 1) & 1) - s_min[1];
      int rz = (int)floor(z[base+v]) + (c & 1) - s_min[2];
      rx = min(rx, 63); ry = min(ry, 63); rz = min(rz, 63);
      keys[i] = ((uint32_t)rx << 12) | ((uint32_t)ry << 6) | (uint32_t)rz;
      vals[i] = (uint16_t)e;
    } else { keys[i] = 0xFFFFFFFFu; vals[i] = 0xFFFF; }
  }
  Sort(tmp).Sort(keys, vals, 0, 19);
  uint16_t* out = perm + (int64_t)blockIdx.x*npair;
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    const int e = threadIdx.x*ITEMS + i;
    if (e < npair) out[e] = vals[i];          // padding keys sort to the end
  }
}

__global__ void k_perm_identity(uint16_t* perm, int npair, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (i < total) perm[i] = (uint16_t)(i % npair);
}

// ---------------------------------------------------------------- spread
// One CTA per cell.  Phase 1 (thread = vertex): force cap, the kernel's per-axis weights and node offsets, the
// normalisation, a mask of the corners that add to a real node.  What phase 2 needs per (vertex, corner) pair is left in
// shared memory FACTORISED, 80 B per vertex (four cells per SM): the node key K0 of the lower corner with a word of flags
// (corner mask; per axis, whether the upper corner wraps around a periodic axis instead of lying one stride on), the four
// products wx[dx]*wy[dy], the two wz[dz] / (sum of the admitted weights) and the three (frep + f)_k.  Phase 2 (thread =
// chunk of 8 consecutive node-sorted pairs) then takes SIX 64-bit shared-memory loads per pair where rebuilding node and
// weight from the staged position took eleven - the kernel is bound by shared-memory wavefronts and their latency, vertex
// indices being random across a warp - merges runs of equal node in registers and issues one fp64 RED triple per run.
// bulk-async staging helpers (1-D TMA path: cp.async.bulk global -> shared, completion on an mbarrier)
__device__ __forceinline__ uint32_t sp_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sp_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(sp_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void sp_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(sp_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sp_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = sp_smem_u32(bar);
  uint32_t ok = 0;
  long long t_start = 0;
  for (unsigned spin = 0; !ok; spin++) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (!ok && (spin & 255u) == 255u) {       // a byte-count mismatch must not hang the device: trap after ~2 s
      const long long now = clock64();
      if (t_start == 0) t_start = now; else if (now - t_start > 4000000000LL) __trap();
    }
  }
}
__device__ __forceinline__ void sp_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(sp_smem_u32(dst)), "l"(src), "r"(bytes), "r"(sp_smem_u32(bar)) : "memory");
}

// BULK: the cell's nine particle arrays (V contiguous doubles each) arrive in shared memory as nine bulk-async
// copies issued by one thread, and the first node-sorted pairs are already in registers when they land:
// phase 1 no longer waits on global loads (they were ~35 % of the kernel's stall samples).  Needs V even
// (16-byte aligned segments); otherwise the direct-load variant runs.
template <int THREADS, bool CHECK_FLAGS, bool BULK>
// three cells per SM: with the registers of two (90) the kernel measured 0.45 ms instead of 0.39
__global__ void __launch_bounds__(THREADS, 4)
k_spread_sorted(SpArgs a, const uint8_t* __restrict__ flags, const uint8_t* __restrict__ alive,
                const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                double* fx, double* fy, double* fz,
                const double* __restrict__ rx, const double* __restrict__ ry, const double* __restrict__ rz,
                const uint16_t* __restrict__ perm, double* __restrict__ F, const uint8_t* __restrict__ far) {
  extern __shared__ __align__(16) double sm[];
  const int V = a.V;
  const int64_t cell = a.first_cell + blockIdx.x;
  if (!alive[cell]) return;
  const bool chk = !(CHECK_FLAGS && far && far[cell]);   // no non-fluid node within this cell's reach: skip the flag look-ups
  const int64_t base = a.first_particle + (int64_t)blockIdx.x*V;
  // twelve arrays of V doubles.  Slots 0-8 receive the inputs (position, repulsion force, membrane force: BULK only);
  // every thread reads its vertex's nine values into registers before it writes that vertex's outputs over them.
  double* IN = sm;                                                 // [9][V]  x y z  rx ry rz  fx fy fz
  double* AB = sm;                                                 // [4][V]  wx[dx]*wy[dy] at (2*dx + dy)
  double* G = sm + 4*V;                                            // [3][V]  (frep + f)_k
  double* CZ = sm + 7*V;                                           // [2][V]  wz[dz]/total
  int2* KV = reinterpret_cast<int2*>(sm + 9*V);                    // [V]  K0 ; corner mask | wrap flags of the x, y, z offsets << 8
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 10*V);
  const SpStrides st = sp_strides(a);

  const uint4* pp = reinterpret_cast<const uint4*>(perm + (int64_t)blockIdx.x*8*V);
  constexpr int NPRE = 3;                                          // chunks of 8 pairs prefetched per thread
  uint4 qpre[NPRE];
  if (BULK) {
    if (threadIdx.x == 0) {
      sp_mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      const uint32_t nb = (uint32_t)(V*sizeof(double));
      sp_mbar_expect_tx(bar, 9*nb);
      sp_bulk_g2s(IN, x + base, nb, bar); sp_bulk_g2s(IN + V, y + base, nb, bar); sp_bulk_g2s(IN + 2*V, z + base, nb, bar);
      sp_bulk_g2s(IN + 3*V, rx + base, nb, bar); sp_bulk_g2s(IN + 4*V, ry + base, nb, bar); sp_bulk_g2s(IN + 5*V, rz + base, nb, bar);
      sp_bulk_g2s(IN + 6*V, fx + base, nb, bar); sp_bulk_g2s(IN + 7*V, fy + base, nb, bar); sp_bulk_g2s(IN + 8*V, fz + base, nb, bar);
    }
#pragma unroll
    for (int r = 0; r < NPRE; r++) { const int j = threadIdx.x + r*THREADS; if (j < V) qpre[r] = __ldg(pp + j); }
    __syncthreads();                                               // the barrier is initialised for everyone
    sp_mbar_wait(bar, 0);
  }

  for (int v = threadIdx.x; v < V; v += THREADS) {
    const int64_t p = base + v;
    double px, py, pz, f0, f1, f2, r0, r1, r2;
    if (BULK) { px = IN[v]; py = IN[V + v]; pz = IN[2*V + v]; r0 = IN[3*V + v]; r1 = IN[4*V + v]; r2 = IN[5*V + v];
                f0 = IN[6*V + v]; f1 = IN[7*V + v]; f2 = IN[8*V + v]; }
    else { px = x[p]; py = y[p]; pz = z[p]; f0 = fx[p]; f1 = fy[p]; f2 = fz[p]; r0 = rx[p]; r1 = ry[p]; r2 = rz[p]; }
    const double mag = sqrt(f0*f0 + f1*f1 + f2*f2);
    if (mag > a.f_limit) {                      // permanent cap (hemoCellParticleField.cpp:848-852)
      const double s = a.f_limit/mag;
      f0 *= s; f1 *= s; f2 *= s;
      fx[p] = f0; fy[p] = f1; fz[p] = f2;
    }
    G[v] = r0 + f0; G[V + v] = r1 + f1; G[2*V + v] = r2 + f2;     // (stored first: six values less to keep across the staging)
    SpVertex sv;
    sp_stage_vertex<CHECK_FLAGS>(a, st, flags, chk, px, py, pz, sv);
#pragma unroll
    for (int d = 0; d < 2; d++) { AB[(2*d)*V + v] = sv.ab[2*d]; AB[(2*d + 1)*V + v] = sv.ab[2*d + 1]; CZ[d*V + v] = sv.cz[d]; }
    KV[v] = make_int2(sv.k0, (int)sv.fl);
  }
  __syncthreads();

  auto process = [&](const uint4 q) {            // 8 consecutive sorted pairs
    const unsigned e8[4] = {q.x, q.y, q.z, q.w};
    int cur = -1; double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const unsigned e = (e8[k >> 1] >> ((k & 1)*16)) & 0xFFFFu;
      const int v = e >> 3, c = e & 7;
      const int2 kv = KV[v];
      int key;
      if (!sp_pair_node(st, kv.x, (unsigned)kv.y, c, key)) continue;
      const double w = AB[(c >> 1)*V + v]*CZ[(c & 1)*V + v];
      const double v0 = G[v]*w, v1 = G[V + v]*w, v2 = G[2*V + v]*w;
      if (key != cur) {
        if (cur >= 0) { double* Fn = F + 4*(int64_t)cur; atomicAdd(Fn, a0); atomicAdd(Fn + 1, a1); atomicAdd(Fn + 2, a2); }
        cur = key; a0 = v0; a1 = v1; a2 = v2;
      } else { a0 += v0; a1 += v1; a2 += v2; }
    }
    if (cur >= 0) { double* Fn = F + 4*(int64_t)cur; atomicAdd(Fn, a0); atomicAdd(Fn + 1, a1); atomicAdd(Fn + 2, a2); }
  };
  if (BULK) {
#pragma unroll
    for (int r = 0; r < NPRE; r++) { const int j = threadIdx.x + r*THREADS; if (j < V) process(qpre[r]); }
    for (int j = threadIdx.x + NPRE*THREADS; j < V; j += THREADS) process(__ldg(pp + j));
  } else {
    for (int j = threadIdx.x; j < V; j += THREADS) process(__ldg(pp + j));
  }
}

SpArgs make_args(const hcg_ctx* c, const CellTypeHost& th) {
  SpArgs a;
  a.nx = c->dom.nx; a.ny = c->dom.ny; a.nz = c->dom.nz;
  a.px = c->dom.periodic[0]; a.py = c->dom.periodic[1]; a.pz = c->dom.periodic[2];
  a.nxl = c->nxl; a.x0 = c->x0; a.nranks = c->dom.n_ranks;
  a.P = c->P; a.f_limit = c->f_limit; a.V = th.d.V;
  a.first_cell = th.first_cell; a.first_particle = th.first_particle;
  return a;
}
inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1)/t); }

}  // namespace

// supported when 8*V pairs fit the compiled sort shapes (RBC 642 -> 256 x 21; PLT 66 -> 64 x 9)
bool spread_sorted_supported(const CellTypeHost& th) { return 8*th.d.V <= 256*21; }

hcg_status spread_sorted_rebuild(hcg_ctx* c) {
  for (auto& th : c->types) {
    if (th.n_cells == 0 || !spread_sorted_supported(th)) continue;
    const int npair = 8*th.d.V;
    if (!th.perm) {
      const int64_t total = (int64_t)th.cap_cells*npair;
      CUDA_TRY(c, cudaMalloc(&th.perm, sizeof(uint16_t)*total));
      k_perm_identity<<<nblk(total, 256), 256, 0, c->stream>>>(th.perm, npair, total);
      KERNEL_CHECK(c);
    }
    SpArgs a = make_args(c, th);
    if (npair <= 64*9) k_pair_sort<64, 9><<<(unsigned)th.n_cells, 64, 0, c->stream>>>(a, c->cell_alive, c->pos[0], c->pos[1], c->pos[2], th.perm);
    else k_pair_sort<256, 21><<<(unsigned)th.n_cells, 256, 0, c->stream>>>(a, c->cell_alive, c->pos[0], c->pos[1], c->pos[2], th.perm);
    KERNEL_CHECK(c);
  }
  return HCG_OK;
}

hcg_status spread_sorted(hcg_ctx* c) {
  static int bulk_env = -1;
  if (bulk_env < 0) { const char* e = getenv("HCG_SPREAD_BULK"); bulk_env = e ? atoi(e) : 1; }
  for (auto& th : c->types) {
    if (th.n_cells == 0) continue;
    SpArgs a = make_args(c, th);
    const int V = th.d.V;
    const bool bulk = bulk_env && (V % 2 == 0) && (th.first_particle % 2 == 0) && V >= 256;
    const size_t smem = sizeof(double)*10*(size_t)V + 16;
    const bool chk = c->has_nonfluid;
#define SP_LAUNCH(T, C, B) do { \
      CUDA_TRY(c, cudaFuncSetAttribute(k_spread_sorted<T, C, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      k_spread_sorted<T, C, B><<<(unsigned)th.n_cells, T, smem, c->stream>>>(a, c->flags, c->cell_alive, c->pos[0], c->pos[1], c->pos[2], \
          c->frc[0], c->frc[1], c->frc[2], c->frep[0], c->frep[1], c->frep[2], th.perm, c->F, ibm_far(c)); } while (0)
    if (V >= 256 && bulk) { if (chk) SP_LAUNCH(256, true, true); else SP_LAUNCH(256, false, true); }
    else if (V >= 256) { if (chk) SP_LAUNCH(256, true, false); else SP_LAUNCH(256, false, false); }
    else { if (chk) SP_LAUNCH(64, true, false); else SP_LAUNCH(64, false, false); }
#undef SP_LAUNCH
    KERNEL_CHECK(c);
  }
  return HCG_OK;
}
