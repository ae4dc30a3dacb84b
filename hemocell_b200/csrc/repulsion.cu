// K4a: cell-cell and wall repulsion on a node-binned neighbour list (counting sort by nearest
// node; CSR bin_start/bin_items; no 10-per-node cap).  Gather form: every particle sums the
// pair forces acting on itself, which reproduces the reference's symmetric half-stencil update
// (R on i, -R on j) including the doubled force for pairs that share a node.
// Replaces HemoCellParticleField::update_pg / applyRepulsionForce / populateBoundaryParticles /
// applyBoundaryRepulsionForce (reference core/hemoCellParticleField.cpp:137-168, 677-743, 865-918).
#include "ctx.cuh"
#include <cub/device/device_scan.cuh>

namespace {

struct RepArgs {
  int nx, ny, nz, px, py, pz;
  int nxl, x0, nranks;
  int64_t P, np;
  double k, cutoff;
};

__device__ __forceinline__ bool wrapc(int& v, int n, int periodic) {
  if (v >= 0 && v < n) return true;
  if (!periodic) return false;
  v %= n; if (v < 0) v += n;
  return true;
}
// global node x -> plane of the padded slab (0 and nxl + 1 are the ghost planes, used with n_ranks > 1 only);
// false if the node is not addressable from this rank or lies outside a non-periodic domain
__device__ __forceinline__ bool slab_x(const RepArgs& a, int gx, int& lx) {
  if (!wrapc(gx, a.nx, a.px)) return false;
  int rel = gx - a.x0; if (rel < 0) rel += a.nx;
  if (rel < a.nxl) { lx = rel + 1; return true; }
  if (a.nranks > 1) {
    if (rel == a.nx - 1) { lx = 0; return true; }
    if (rel == a.nxl) { lx = a.nxl + 1; return true; }
  }
  return false;
}
// bin = nearest node, as an index into the padded slab; -1 = not binned here
__device__ __forceinline__ int64_t bin_of(const RepArgs& a, double x, double y, double z, bool* owned) {
  int by = (int)floor(y + 0.5), bz = (int)floor(z + 0.5), lx;
  if (!slab_x(a, (int)floor(x + 0.5), lx) || !wrapc(by, a.ny, a.py) || !wrapc(bz, a.nz, a.pz)) return -1;
  if (owned) *owned = lx >= 1 && lx <= a.nxl;
  return bz + (int64_t)a.nz*(by + (int64_t)a.ny*lx);
}

__global__ void k_bin_count(RepArgs a, const double* __restrict__ x, const double* __restrict__ y,
                            const double* __restrict__ z, const int32_t* __restrict__ p_cell,
                            const uint8_t* __restrict__ alive, int* count) {
  const int64_t p = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (p >= a.np || !alive[p_cell[p]]) return;
  const int64_t b = bin_of(a, x[p], y[p], z[p], nullptr);
  if (b >= 0) atomicAdd(count + b, 1);
}
__global__ void k_bin_fill(RepArgs a, const double* __restrict__ x, const double* __restrict__ y,
                           const double* __restrict__ z, const int32_t* __restrict__ p_cell,
                           const uint8_t* __restrict__ alive, const int* __restrict__ start, int* cursor, int* items) {
  const int64_t p = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (p >= a.np || !alive[p_cell[p]]) return;
  const int64_t b = bin_of(a, x[p], y[p], z[p], nullptr);
  if (b >= 0) items[start[b] + atomicAdd(cursor + b, 1)] = (int)p;
}
// Deterministic order inside every bin: ascending (global cell id, vertex).  The same key on every rank
// and on a single GPU, so the pair forces are summed in the same order whatever the decomposition.
__device__ __forceinline__ long long order_key(int p, const int32_t* __restrict__ p_cell, const int64_t* __restrict__ cell_gid,
                                               const int64_t* __restrict__ cell_base) {
  const int c = p_cell[p];
  return (long long)cell_gid[c]*65536LL + (long long)(p - cell_base[c]);
}
__global__ void k_bin_sort(int64_t nbins, const int* __restrict__ start, int* items, const int32_t* __restrict__ p_cell,
                           const int64_t* __restrict__ cell_gid, const int64_t* __restrict__ cell_base) {
  const int64_t b = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (b >= nbins) return;
  const int s = start[b], e = start[b+1];
  for (int i = s + 1; i < e; i++) {
    const int v = items[i]; const long long kv = order_key(v, p_cell, cell_gid, cell_base);
    int j = i - 1;
    while (j >= s && order_key(items[j], p_cell, cell_gid, cell_base) > kv) { items[j+1] = items[j]; j--; }
    items[j+1] = v;
  }
}

// every OWNED particle (nearest node on a real plane of this rank) sums the pair forces acting on itself;
// the copies of a shared cell's other vertices get theirs from the neighbour (multi_field_sync)
__global__ void __launch_bounds__(128)
k_repulse(RepArgs a, const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
          const int32_t* __restrict__ p_cell, const uint8_t* __restrict__ alive,
          const int* __restrict__ start, const int* __restrict__ items,
          double* rx, double* ry, double* rz) {
  const int64_t p = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (p >= a.np) return;
  const int cell = p_cell[p];
  if (!alive[cell]) return;
  const double xi = x[p], yi = y[p], zi = z[p];
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  const int gx = (int)floor(xi + 0.5);
  int by = (int)floor(yi + 0.5), bz = (int)floor(zi + 0.5), lx;
  const bool binned = slab_x(a, gx, lx) && wrapc(by, a.ny, a.py) && wrapc(bz, a.nz, a.pz);
  if (a.nranks > 1 && !(binned && lx >= 1 && lx <= a.nxl)) return;       // not mine: value arrives with the sync
  if (binned) {
    for (int dx = -1; dx <= 1; dx++) for (int dy = -1; dy <= 1; dy++) for (int dz = -1; dz <= 1; dz++) {
      int yy = by + dy, zz = bz + dz, lxx;
      if (!slab_x(a, gx + dx, lxx) || !wrapc(yy, a.ny, a.py) || !wrapc(zz, a.nz, a.pz)) continue;
      const int64_t nb = zz + (int64_t)a.nz*(yy + (int64_t)a.ny*lxx);
      const double mult = (dx == 0 && dy == 0 && dz == 0) ? 2.0 : 1.0;   // same-node pairs are visited twice
      for (int s = start[nb]; s < start[nb+1]; s++) {
        const int j = items[s];
        if (j == p || p_cell[j] == cell) continue;
        double d0 = xi - x[j], d1 = yi - y[j], d2 = zi - z[j];
        if (a.px) d0 -= a.nx*rint(d0/a.nx);
        if (a.py) d1 -= a.ny*rint(d1/a.ny);
        if (a.pz) d2 -= a.nz*rint(d2/a.nz);
        const double dist = sqrt(d0*d0 + d1*d1 + d2*d2);
        if (dist < a.cutoff) {
          const double sc = a.k*(1/(dist/a.cutoff));
          a0 += mult*(sc*(d0/dist)); a1 += mult*(sc*(d1/dist)); a2 += mult*(sc*(d2/dist));
        }
      }
    }
  }
  rx[p] = a0; ry[p] = a1; rz[p] = a2;
}

// wall "boundary particles": boundary nodes that touch a fluid node (populateBoundaryParticles); real planes
// of the padded slab (the ghost planes of the mask come from the neighbours)
__global__ void k_wall_mask(RepArgs a, const uint8_t* __restrict__ flags /* padded slab */, uint8_t* mask) {
  const int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= (int64_t)a.nxl*a.P) return;
  const int64_t n = i + a.P;
  uint8_t m = 0;
  if (flags[n] != HCG_FLUID) {
    const int z = (int)(n % a.nz), y = (int)((n / a.nz) % a.ny), lx = (int)(n / a.P);
    for (int dx = -1; dx <= 1 && !m; dx++) for (int dy = -1; dy <= 1 && !m; dy++) for (int dz = -1; dz <= 1 && !m; dz++) {
      int yy = y + dy, zz = z + dz;
      if (!wrapc(yy, a.ny, a.py) || !wrapc(zz, a.nz, a.pz)) continue;
      // x: the ghost planes hold the periodic image / the neighbour's face, or non-fluid beyond a non-periodic end
      if (flags[zz + (int64_t)a.nz*(yy + (int64_t)a.ny*(lx + dx))] == HCG_FLUID) m = 1;
    }
  }
  mask[n] = m;
}

__global__ void __launch_bounds__(128)
k_wall_repulse(RepArgs a, const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
               const int32_t* __restrict__ p_cell, const uint8_t* __restrict__ alive,
               const uint8_t* __restrict__ mask, double* rx, double* ry, double* rz) {
  const int64_t p = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (p >= a.np || !alive[p_cell[p]]) return;
  const double xi = x[p], yi = y[p], zi = z[p];
  const int ux = (int)floor(xi + 0.5), uy = (int)floor(yi + 0.5), uz = (int)floor(zi + 0.5);
  int by = uy, bz = uz, lx;
  if (!slab_x(a, ux, lx) || !wrapc(by, a.ny, a.py) || !wrapc(bz, a.nz, a.pz)) return;
  if (lx < 1 || lx > a.nxl) return;                                    // owned particles only
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  for (int dx = -1; dx <= 1; dx++) for (int dy = -1; dy <= 1; dy++) for (int dz = -1; dz <= 1; dz++) {
    int yy = by + dy, zz = bz + dz, lxx;
    if (!slab_x(a, ux + dx, lxx) || !wrapc(yy, a.ny, a.py) || !wrapc(zz, a.nz, a.pz)) continue;
    if (!mask[zz + (int64_t)a.nz*(yy + (int64_t)a.ny*lxx)]) continue;
    const double d0 = xi - (double)(ux + dx), d1 = yi - (double)(uy + dy), d2 = zi - (double)(uz + dz);
    const double dist = sqrt(d0*d0 + d1*d1 + d2*d2);
    if (dist < a.cutoff) {
      const double sc = a.k*(1/(dist/a.cutoff));
      a0 += sc*(d0/dist); a1 += sc*(d1/dist); a2 += sc*(d2/dist);
    }
  }
  rx[p] += a0; ry[p] += a1; rz[p] += a2;     // accumulates, never zeroes (Appendix D.5)
}

RepArgs make_args(const hcg_ctx* c, double k, double cut) {
  RepArgs a;
  a.nx = c->dom.nx; a.ny = c->dom.ny; a.nz = c->dom.nz;
  a.px = c->dom.periodic[0]; a.py = c->dom.periodic[1]; a.pz = c->dom.periodic[2];
  a.nxl = c->nxl; a.x0 = c->x0; a.nranks = c->dom.n_ranks; a.P = c->P;
  a.np = c->np; a.k = k; a.cutoff = cut;
  return a;
}
inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1)/t); }

hcg_status build_bins(hcg_ctx* c, const RepArgs& a) {
  const int64_t N = c->S;                    // one bin per node of the padded slab
  if (!c->bin_count) {
    CUDA_TRY(c, cudaMalloc(&c->bin_count, sizeof(int)*(N + 1)));
    CUDA_TRY(c, cudaMalloc(&c->bin_start, sizeof(int)*(N + 1)));
    CUDA_TRY(c, cudaMalloc(&c->bin_items, sizeof(int)*(c->cap_p > 0 ? c->cap_p : 1)));
    c->scan_tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, c->scan_tmp_bytes, c->bin_count, c->bin_start, (int)(N + 1), c->stream);
    CUDA_TRY(c, cudaMalloc(&c->scan_tmp, c->scan_tmp_bytes));
  }
  hcg_status s = multi_upload_cell_gid(c); if (s) return s;
  CUDA_TRY(c, cudaMemsetAsync(c->bin_count, 0, sizeof(int)*(N + 1), c->stream));
  k_bin_count<<<nblk(c->np, 256), 256, 0, c->stream>>>(a, c->pos[0], c->pos[1], c->pos[2], c->p_cell, c->cell_alive, c->bin_count);
  KERNEL_CHECK(c);
  CUDA_TRY(c, cub::DeviceScan::ExclusiveSum(c->scan_tmp, c->scan_tmp_bytes, c->bin_count, c->bin_start, (int)(N + 1), c->stream));
  c->launches++;
  CUDA_TRY(c, cudaMemsetAsync(c->bin_count, 0, sizeof(int)*(N + 1), c->stream));
  k_bin_fill<<<nblk(c->np, 256), 256, 0, c->stream>>>(a, c->pos[0], c->pos[1], c->pos[2], c->p_cell, c->cell_alive,
                                                     c->bin_start, c->bin_count, c->bin_items);
  KERNEL_CHECK(c);
  k_bin_sort<<<nblk(N, 256), 256, 0, c->stream>>>(N, c->bin_start, c->bin_items, c->p_cell, c->cell_gid, c->cell_base);
  KERNEL_CHECK(c);
  return HCG_OK;
}

}  // namespace

hcg_status rep_apply(hcg_ctx* c) {
  if (c->np == 0 && c->dom.n_ranks == 1) return HCG_OK;
  RepArgs a = make_args(c, c->rep_k, c->rep_cut);
  if (c->np > 0) {
    hcg_status s = build_bins(c, a); if (s) return s;
    k_repulse<<<nblk(c->np, 128), 128, 0, c->stream>>>(a, c->pos[0], c->pos[1], c->pos[2], c->p_cell, c->cell_alive,
        c->bin_start, c->bin_items, c->frep[0], c->frep[1], c->frep[2]);
    KERNEL_CHECK(c);
  }
  if (c->dom.n_ranks > 1) return multi_field_sync(c, 1);     // copies of shared cells: the owner's value per vertex
  return HCG_OK;
}

hcg_status rep_wall_apply(hcg_ctx* c) {
  if (c->np == 0 && c->dom.n_ranks == 1) return HCG_OK;
  RepArgs a = make_args(c, c->wall_k, c->wall_cut);
  if (!c->wall_built) {
    if (!c->wall_nodes) CUDA_TRY(c, cudaMalloc((void**)&c->wall_nodes, c->S));   // used as the uint8 mask over the padded slab
    CUDA_TRY(c, cudaMemsetAsync(c->wall_nodes, 0, c->S, c->stream));
    k_wall_mask<<<nblk(c->Nl, 256), 256, 0, c->stream>>>(a, c->flags, (uint8_t*)c->wall_nodes);
    KERNEL_CHECK(c);
    hcg_status s = lat_exchange_byte_planes(c, (uint8_t*)c->wall_nodes, 0); if (s) return s;   // ghost planes of the mask
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->wall_built = true;
  }
  if (c->np > 0) {
    k_wall_repulse<<<nblk(c->np, 128), 128, 0, c->stream>>>(a, c->pos[0], c->pos[1], c->pos[2], c->p_cell, c->cell_alive,
        (const uint8_t*)c->wall_nodes, c->frep[0], c->frep[1], c->frep[2]);
    KERNEL_CHECK(c);
  }
  if (c->dom.n_ranks > 1) return multi_field_sync(c, 1);
  return HCG_OK;
}
