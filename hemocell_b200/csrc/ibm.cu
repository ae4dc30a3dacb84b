// K2 / K4b: immersed-boundary spread + interpolate (phi2 kernel) and particle advance.
// Replaces HemoCellParticleField::spreadParticleForce / interpolateFluidVelocity /
// advanceParticles (reference core/hemoCellParticleField.cpp:841-863, 819-839, 566-588) and
// interpolationCoefficientsPhi2 (core/immersedBoundaryMethod.h:62-138).
#include "ctx.cuh"
#include <vector>
#include "ibm_node.cuh"   // IbmArgs, local_x, wrap_yz, phi2, ibm_kernel, interp_vertex (host + device)

namespace {

__global__ void __launch_bounds__(256)
k_spread(IbmArgs a, const uint8_t* __restrict__ flags, const int32_t* __restrict__ p_cell,
         const uint8_t* __restrict__ alive,
         const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
         double* fx, double* fy, double* fz,
         const double* __restrict__ rx, const double* __restrict__ ry, const double* __restrict__ rz,
         double* __restrict__ F) {
  const int64_t p = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (p >= a.np) return;
  if (!alive[p_cell[p]]) return;
  double f0 = fx[p], f1 = fy[p], f2 = fz[p];
  // force cap, permanent (hemoCellParticleField.cpp:848-852)
  const double mag = sqrt(f0*f0 + f1*f1 + f2*f2);
  if (mag > a.f_limit) {
    const double s = a.f_limit/mag;
    f0 *= s; f1 *= s; f2 *= s;
    fx[p] = f0; fy[p] = f1; fz[p] = f2;
  }
  int64_t node[8]; double w[8];
  const int n = ibm_kernel(a, flags, x[p], y[p], z[p], node, w);
  const double t0 = rx[p] + f0, t1 = ry[p] + f1, t2 = rz[p] + f2;
  const int64_t lo = a.P, hi = (int64_t)(a.nxl + 1)*a.P;      // real planes only
  for (int k = 0; k < n; k++) {
    if (node[k] < lo || node[k] >= hi) continue;
    double* Fn = F + 4*node[k];            // node vectors are AoS [n][4]: one 32-byte sector per node
    atomicAdd(Fn, t0*w[k]);
    atomicAdd(Fn + 1, t1*w[k]);
    atomicAdd(Fn + 2, t2*w[k]);
  }
}

// hold_back (multi-GPU): cells flagged there wait for the neighbour's velocities before they move
// advance of the cells on a list (one CTA per cell): the shared cells after the velocity sync
__global__ void __launch_bounds__(256)
k_advance_list(IbmArgs a, const uint8_t* __restrict__ flags, const int32_t* __restrict__ cells, const int64_t* __restrict__ off,
               int n, const int64_t* __restrict__ cell_base, uint8_t* alive, double* x, double* y, double* z,
               const double* __restrict__ vx, const double* __restrict__ vy, const double* __restrict__ vz);


// One launch per cell type, one thread per vertex.  The kernel is bound by the latency of its dependent loads, so the
// chain is kept two deep: the vertex's cell follows from its index (the cells of a type are V consecutive vertices each -
// no particle -> cell table), and position, liveness and the cell's flags are requested together before the eight node
// loads they decide.  (Tried and dropped, profiles/README.md r2y: two lanes per vertex, one per z offset of the corners, so
// that a lane pair reads two neighbouring node records - 0.34 ms instead of 0.23; five or six CTAs per SM at 48 / 40
// registers - 0.237 / 0.245 ms.)
struct TypeSpan { int64_t first_particle, first_cell, n; int V; double inv_V; };

template <bool ADVANCE, bool INTERP, bool CHECK_FLAGS>
__global__ void __launch_bounds__(256, 4)
k_interp_advance(IbmArgs a, TypeSpan ts, const uint8_t* __restrict__ flags,
                 uint8_t* alive, double* x, double* y, double* z,
                 double* vx, double* vy, double* vz, const double* __restrict__ U,
                 const uint8_t* __restrict__ hold_back, int skip_held, const uint8_t* __restrict__ far) {
  const int64_t q = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (q >= ts.n) return;
  const int64_t p = ts.first_particle + q;
  int64_t lc = (int64_t)((double)q*ts.inv_V);                       // q / V, corrected for the rounding of the reciprocal
  if (lc*ts.V > q) lc--; else if ((lc + 1)*ts.V <= q) lc++;
  const int64_t cell = ts.first_cell + lc;
  double px = x[p], py = y[p], pz = z[p];
  const uint8_t live = alive[cell];
  const uint8_t is_far = (CHECK_FLAGS && far) ? far[cell] : 0;
  const uint8_t held = hold_back ? hold_back[cell] : 0;
  if (!live) return;
  const bool chk = !(CHECK_FLAGS && is_far);      // cells with no non-fluid node within reach skip the flag look-ups
  if (skip_held && held) return;                  // shared cells are interpolated by k_interp_list on the main stream
  double v0, v1, v2;
  if (INTERP) {
    if (interp_vertex<CHECK_FLAGS>(a, flags, U, px, py, pz, v0, v1, v2, chk)) { vx[p] = v0; vy[p] = v1; vz[p] = v2; }
    else { v0 = vx[p]; v1 = vy[p]; v2 = vz[p]; }
  } else { v0 = vx[p]; v1 = vy[p]; v2 = vz[p]; }
  if (ADVANCE && !held) {
    px += v0; py += v1; pz += v2;
    x[p] = px; y[p] = py; z[p] = pz;
    // particle on a boundary node => its cell is deleted (hemoCellParticleField.cpp:572-584, 512-553)
    if (CHECK_FLAGS && chk) {                      // without non-fluid nodes within reach no particle can land on one
      int lx; bool out;
      int yy = (int)floor(py + 0.5), zz = (int)floor(pz + 0.5);
      if (local_x((int)floor(px + 0.5), a, lx, out) && wrap_yz(yy, a.ny, a.py) && wrap_yz(zz, a.nz, a.pz)) {
        if (flags[(int64_t)zz + (int64_t)a.nz*((int64_t)yy + (int64_t)a.ny*lx)] != HCG_FLUID) alive[cell] = 0;
      }
    }
  }
}

__global__ void __launch_bounds__(256)
k_advance_list(IbmArgs a, const uint8_t* __restrict__ flags, const int32_t* __restrict__ cells, const int64_t* __restrict__ off,
               int n, const int64_t* __restrict__ cell_base, uint8_t* alive, double* x, double* y, double* z,
               const double* __restrict__ vx, const double* __restrict__ vy, const double* __restrict__ vz) {
  const int i = blockIdx.x;
  if (i >= n) return;
  const int cell = cells[i];
  if (!alive[cell]) return;
  const int64_t b = cell_base[cell];
  const int V = (int)(off[i+1] - off[i]);
  for (int k = threadIdx.x; k < V; k += blockDim.x) {
    const int64_t p = b + k;
    const double px = x[p] + vx[p], py = y[p] + vy[p], pz = z[p] + vz[p];
    x[p] = px; y[p] = py; z[p] = pz;
    int lx; bool out;
    int yy = (int)floor(py + 0.5), zz = (int)floor(pz + 0.5);
    if (local_x((int)floor(px + 0.5), a, lx, out) && wrap_yz(yy, a.ny, a.py) && wrap_yz(zz, a.nz, a.pz)) {
      if (flags[(int64_t)zz + (int64_t)a.nz*((int64_t)yy + (int64_t)a.ny*lx)] != HCG_FLUID) alive[cell] = 0;
    }
  }
}


// interpolation of the cells on a list (one CTA per cell): the shared cells of the multi-GPU run, ahead of the
// velocity exchange, while the unshared majority is interpolated and advanced on a second stream
template <bool CHECK_FLAGS>
__global__ void __launch_bounds__(256)
k_interp_list(IbmArgs a, const uint8_t* __restrict__ flags, const int32_t* __restrict__ cells, const int64_t* __restrict__ off,
              int n, const int64_t* __restrict__ cell_base, const uint8_t* __restrict__ alive,
              const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
              double* vx, double* vy, double* vz, const double* __restrict__ U) {
  const int i = blockIdx.x;
  if (i >= n) return;
  const int cell = cells[i];
  if (!alive[cell]) return;
  const int64_t b = cell_base[cell];
  const int V = (int)(off[i+1] - off[i]);
  for (int k = threadIdx.x; k < V; k += blockDim.x) {
    const int64_t p = b + k;
    double v0, v1, v2;
    if (interp_vertex<CHECK_FLAGS>(a, flags, U, x[p], y[p], z[p], v0, v1, v2)) { vx[p] = v0; vy[p] = v1; vz[p] = v2; }
  }
}

// 8^3 blocks of the padded slab (ghost planes included) that hold a non-fluid node
__global__ void k_wall_coarse(const uint8_t* __restrict__ flags, int nxp, int ny, int nz, int bx, int by, int bz, uint8_t* __restrict__ out) {
  const int b = blockIdx.x*blockDim.x + threadIdx.x;
  if (b >= bx*by*bz) return;
  const int k = b % bz, j = (b / bz) % by, i = b / (bz*by);
  uint8_t any = 0;
  for (int x = 8*i; x < min(8*i + 8, nxp) && !any; x++)
    for (int y = 8*j; y < min(8*j + 8, ny) && !any; y++)
      for (int z = 8*k; z < min(8*k + 8, nz); z++)
        if (flags[(int64_t)z + (int64_t)nz*((int64_t)y + (int64_t)ny*x)] != HCG_FLUID) { any = 1; break; }
  out[b] = any;
}
// one warp per cell: far = the bounding box grown by `margin` nodes lies inside the slab (ghost planes included) without
// wrapping around a periodic axis and touches no block that holds a non-fluid node.  Conservative: anything else is "near".
__global__ void k_far_classify(IbmArgs a, const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                               const int64_t* __restrict__ cell_base, const int32_t* __restrict__ cell_type, const int* __restrict__ typeV,
                               const uint8_t* __restrict__ alive, int64_t ncells, const uint8_t* __restrict__ coarse, int by, int bz,
                               double margin, uint8_t* __restrict__ far) {
  const int64_t cell = (int64_t)blockIdx.x*(blockDim.x/32) + threadIdx.x/32;
  if (cell >= ncells) return;
  const int lane = threadIdx.x & 31;
  if (!alive[cell]) { if (lane == 0) far[cell] = 0; return; }
  const int V = typeV[cell_type[cell]]; const int64_t b = cell_base[cell];
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
  for (int i = lane; i < V; i += 32) {
    const double p[3] = {x[b+i], y[b+i], z[b+i]};
    for (int d = 0; d < 3; d++) { mn[d] = fmin(mn[d], p[d]); mx[d] = fmax(mx[d], p[d]); }
  }
  for (int s = 16; s > 0; s >>= 1)
    for (int d = 0; d < 3; d++) { mn[d] = fmin(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], s)); mx[d] = fmax(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], s)); }
  if (lane != 0) return;
  // x: whole periods off (unwrapped coordinates), then local plane index incl. the ghost planes
  const double kx = a.px ? floor(mn[0]/a.nx)*a.nx : 0.0;
  const int x0l = (int)floor(mn[0] - kx - margin) - a.x0 + 1, x1l = (int)ceil(mx[0] - kx + margin) - a.x0 + 1;
  const int y0 = (int)floor(mn[1] - margin), y1 = (int)ceil(mx[1] + margin), z0 = (int)floor(mn[2] - margin), z1 = (int)ceil(mx[2] + margin);
  uint8_t ok = x0l >= 0 && x1l <= a.nxl + 1 && y0 >= 0 && y1 <= a.ny - 1 && z0 >= 0 && z1 <= a.nz - 1;
  for (int i = x0l >> 3; ok && i <= (x1l >> 3); i++)
    for (int j = y0 >> 3; ok && j <= (y1 >> 3); j++)
      for (int k = z0 >> 3; k <= (z1 >> 3); k++)
        if (coarse[(int64_t)k + (int64_t)bz*((int64_t)j + (int64_t)by*i)]) { ok = 0; break; }
  far[cell] = ok;
}

IbmArgs make_args(const hcg_ctx* c) {
  IbmArgs a;
  a.nx = c->dom.nx; a.ny = c->dom.ny; a.nz = c->dom.nz;
  a.px = c->dom.periodic[0]; a.py = c->dom.periodic[1]; a.pz = c->dom.periodic[2];
  a.nxl = c->nxl; a.x0 = c->x0; a.nranks = c->dom.n_ranks;
  a.P = c->P; a.S = c->S; a.np = c->np; a.f_limit = c->f_limit;
  return a;
}
inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1)/t); }

}  // namespace

// Lattices with walls: classify the cells that cannot meet a non-fluid node during the next `valid_steps` advances (kernel
// support 1 node + drift of at most 0.1 lu per step + 1 node of safety); the IBM kernels skip their per-corner flag look-ups
// for those.  Called at the cadence of the spreading permutation; positions set from outside invalidate it.
hcg_status ibm_far_classify(hcg_ctx* c, int valid_steps) {
  c->far_steps_left = 0;
  if (!c->has_nonfluid || c->ncells == 0 || c->np == 0) return HCG_OK;
  const int nxp = c->nxl + 2, ny = c->dom.ny, nz = c->dom.nz;
  const int bx = (nxp + 7)/8, by = (ny + 7)/8, bz = (nz + 7)/8;
  if (!c->wall_coarse || c->wc_dim[0] != bx || c->wc_dim[1] != by || c->wc_dim[2] != bz) {
    if (c->wall_coarse) cudaFree(c->wall_coarse);
    c->wall_coarse = nullptr;
    CUDA_TRY(c, cudaMalloc(&c->wall_coarse, (size_t)bx*by*bz));
    c->wc_dim[0] = bx; c->wc_dim[1] = by; c->wc_dim[2] = bz; c->wall_coarse_valid = false;
  }
  if (!c->wall_coarse_valid) {
    k_wall_coarse<<<nblk((int64_t)bx*by*bz, 128), 128, 0, c->stream>>>(c->flags, nxp, ny, nz, bx, by, bz, c->wall_coarse);
    KERNEL_CHECK(c);
    c->wall_coarse_valid = true;
  }
  if (c->cell_far_cap < c->ncells) {
    if (c->cell_far) cudaFree(c->cell_far);
    c->cell_far = nullptr; c->cell_far_cap = 0;
    CUDA_TRY(c, cudaMalloc(&c->cell_far, (size_t)c->ncells));
    c->cell_far_cap = c->ncells;
  }
  if ((int)c->types.size() != c->far_ntypes) {
    std::vector<int> hv; for (auto& t : c->types) hv.push_back(t.d.V);
    if (c->far_typeV) { CUDA_TRY(c, cudaStreamSynchronize(c->stream)); cudaFree(c->far_typeV); c->far_typeV = nullptr; }
    CUDA_TRY(c, cudaMalloc(&c->far_typeV, sizeof(int)*hv.size()));
    CUDA_TRY(c, hcg_h2d(c, c->far_typeV, hv.data(), sizeof(int)*hv.size()));
    c->far_ntypes = (int)c->types.size();
  }
  IbmArgs a = make_args(c);
  const double margin = 2.0 + 0.1*valid_steps;
  k_far_classify<<<(unsigned)((c->ncells + 7)/8), 256, 0, c->stream>>>(a, c->pos[0], c->pos[1], c->pos[2], c->cell_base, c->cell_type, c->far_typeV,
      c->cell_alive, c->ncells, c->wall_coarse, by, bz, margin, c->cell_far);
  KERNEL_CHECK(c);
  c->far_steps_left = valid_steps;
  return HCG_OK;
}

hcg_status ibm_spread(hcg_ctx* c) {
  if (c->np == 0) return HCG_OK;
  IbmArgs a = make_args(c);
  k_spread<<<nblk(c->np, 256), 256, 0, c->stream>>>(a, c->flags, c->p_cell, c->cell_alive,
      c->pos[0], c->pos[1], c->pos[2], c->frc[0], c->frc[1], c->frc[2],
      c->frep[0], c->frep[1], c->frep[2], c->F);
  KERNEL_CHECK(c);
  return HCG_OK;
}

// interpolation and / or advance of every vertex: one launch per cell type
template <bool ADVANCE, bool INTERP>
static hcg_status launch_interp_advance(hcg_ctx* c, cudaStream_t st, const uint8_t* hold_back, int skip_held) {
  IbmArgs a = make_args(c);
  for (auto& th : c->types) {
    if (th.n_cells == 0) continue;
    TypeSpan ts{th.first_particle, th.first_cell, th.n_cells*(int64_t)th.d.V, th.d.V, 1.0/(double)th.d.V};
    if (c->has_nonfluid) k_interp_advance<ADVANCE, INTERP, true><<<nblk(ts.n, 256), 256, 0, st>>>(a, ts, c->flags, c->cell_alive,
        c->pos[0], c->pos[1], c->pos[2], c->vel[0], c->vel[1], c->vel[2], c->U, hold_back, skip_held, ibm_far(c));
    else k_interp_advance<ADVANCE, INTERP, false><<<nblk(ts.n, 256), 256, 0, st>>>(a, ts, c->flags, c->cell_alive,
        c->pos[0], c->pos[1], c->pos[2], c->vel[0], c->vel[1], c->vel[2], c->U, hold_back, skip_held, ibm_far(c));
    KERNEL_CHECK(c);
  }
  return HCG_OK;
}

hcg_status ibm_interpolate(hcg_ctx* c) {
  if (c->np == 0) return HCG_OK;
  return launch_interp_advance<false, true>(c, c->stream, nullptr, 0);
}

hcg_status ibm_advance(hcg_ctx* c) {
  if (c->np == 0) return HCG_OK;
  struct FarTick { hcg_ctx* c; ~FarTick() { if (c->far_steps_left > 0) c->far_steps_left--; } } far_tick{c};
  return launch_interp_advance<true, false>(c, c->stream, nullptr, 0);
}

hcg_status ibm_interpolate_advance(hcg_ctx* c) {
  if (c->np == 0) return HCG_OK;
  struct FarTick { hcg_ctx* c; ~FarTick() { if (c->far_steps_left > 0) c->far_steps_left--; } } far_tick{c};
  return launch_interp_advance<true, true>(c, c->stream, nullptr, 0);
}

hcg_status ibm_interpolate_advance_unshared(hcg_ctx* c) {
  if (c->np == 0) return HCG_OK;
  struct FarTick { hcg_ctx* c; ~FarTick() { if (c->far_steps_left > 0) c->far_steps_left--; } } far_tick{c};
  if (!c->multi.d_cell_shared) return hcg_fail(c, HCG_ERR_STATE, "shared-cell flags missing (multi_rebalance has not run)");
  return launch_interp_advance<true, true>(c, c->stream, c->multi.d_cell_shared, 0);
}

// multi-GPU overlap: (1) the shared cells on the main stream ...
hcg_status ibm_interpolate_shared(hcg_ctx* c) {
  const MultiFace& f = c->multi.all;
  if (c->np == 0 || f.n == 0) return HCG_OK;
  IbmArgs a = make_args(c);
  if (c->has_nonfluid) k_interp_list<true><<<f.n, 256, 0, c->stream>>>(a, c->flags, f.d_cells, f.d_off, f.n, c->cell_base, c->cell_alive,
      c->pos[0], c->pos[1], c->pos[2], c->vel[0], c->vel[1], c->vel[2], c->U);
  else k_interp_list<false><<<f.n, 256, 0, c->stream>>>(a, c->flags, f.d_cells, f.d_off, f.n, c->cell_base, c->cell_alive,
      c->pos[0], c->pos[1], c->pos[2], c->vel[0], c->vel[1], c->vel[2], c->U);
  KERNEL_CHECK(c);
  return HCG_OK;
}
// ... (2) interpolation + advance of the cells no neighbour holds, on `st`
hcg_status ibm_interpolate_advance_unshared_on(hcg_ctx* c, cudaStream_t st) {
  if (c->np == 0) return HCG_OK;
  struct FarTick { hcg_ctx* c; ~FarTick() { if (c->far_steps_left > 0) c->far_steps_left--; } } far_tick{c};
  if (!c->multi.d_cell_shared) return hcg_fail(c, HCG_ERR_STATE, "shared-cell flags missing (multi_rebalance has not run)");
  return launch_interp_advance<true, true>(c, st, c->multi.d_cell_shared, 1);
}

hcg_status ibm_advance_shared(hcg_ctx* c) {
  const MultiFace& f = c->multi.all;
  if (c->np == 0 || f.n == 0) return HCG_OK;
  IbmArgs a = make_args(c);
  k_advance_list<<<f.n, 256, 0, c->stream>>>(a, c->flags, f.d_cells, f.d_off, f.n, c->cell_base, c->cell_alive,
      c->pos[0], c->pos[1], c->pos[2], c->vel[0], c->vel[1], c->vel[2]);
  KERNEL_CHECK(c);
  return HCG_OK;
}
