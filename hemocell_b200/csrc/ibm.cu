// K2 / K4b: immersed-boundary spread + interpolate (phi2 kernel) and particle advance.
// Replaces HemoCellParticleField::spreadParticleForce / interpolateFluidVelocity /
// advanceParticles (reference core/hemoCellParticleField.cpp:841-863, 819-839, 566-588) and
// interpolationCoefficientsPhi2 (core/immersedBoundaryMethod.h:62-138).
#include "ctx.cuh"

namespace {

struct IbmArgs {
  int nx, ny, nz, px, py, pz;
  int nxl, x0, nranks;
  int64_t P, S, np;
  double f_limit;
};

// global (unwrapped) node x -> local plane index incl. ghosts; false if not held by this rank
// or outside a non-periodic domain
__device__ __forceinline__ bool local_x(int gx, const IbmArgs& a, int& lx, bool& outside) {
  outside = false;
  if (gx < 0 || gx >= a.nx) {
    if (!a.px) { outside = true; return false; }
    gx %= a.nx; if (gx < 0) gx += a.nx;
  }
  int rel = gx - a.x0; if (rel < 0) rel += a.nx;
  if (rel < a.nxl) { lx = rel + 1; return true; }
  if (a.nranks > 1) {
    if (rel == a.nx - 1) { lx = 0; return true; }
    if (rel == a.nxl) { lx = a.nxl + 1; return true; }
  }
  return false;
}
__device__ __forceinline__ bool wrap_yz(int& v, int n, int periodic) {
  if (v >= 0 && v < n) return true;
  if (!periodic) return false;
  v %= n; if (v < 0) v += n;
  return true;
}
__device__ __forceinline__ double phi2(double x) { x = 1.0 - fabs(x); return x > 0.0 ? x : 0.0; }

// Kernel of one particle: up to 8 (node, weight) pairs in the reference's x-outer/z-inner
// order, zero weights and boundary nodes skipped, normalised.  Returns the count, or -1 when a
// candidate node is not addressable from this rank (particle irrelevant here).
template <bool CHECK_FLAGS = true>
__device__ __forceinline__ int ibm_kernel(const IbmArgs& a, const uint8_t* __restrict__ flags,
                                          double px, double py, double pz, int64_t node[8], double w[8]) {
  const int bx = (int)floor(px), by = (int)floor(py), bz = (int)floor(pz);
  int n = 0; double total = 0.0;
#pragma unroll
  for (int dx = 0; dx < 2; dx++) {
    const double wx = phi2(px - (double)(bx + dx));
    if (wx == 0.0) continue;
    int lx; bool out;
    if (!local_x(bx + dx, a, lx, out)) { if (out) continue; return -1; }
#pragma unroll
    for (int dy = 0; dy < 2; dy++) {
      const double wy = phi2(py - (double)(by + dy));
      int y = by + dy;
      if (wy == 0.0 || !wrap_yz(y, a.ny, a.py)) continue;
#pragma unroll
      for (int dz = 0; dz < 2; dz++) {
        const double wz = phi2(pz - (double)(bz + dz));
        int z = bz + dz;
        if (wz == 0.0 || !wrap_yz(z, a.nz, a.pz)) continue;
        const double weight = wx*wy*wz;
        if (weight == 0.0) continue;
        const int64_t id = (int64_t)z + (int64_t)a.nz*((int64_t)y + (int64_t)a.ny*lx);
        if (CHECK_FLAGS && flags[id] != HCG_FLUID) continue;     // !CHECK_FLAGS: every addressable node is plain fluid
        total += weight;
        node[n] = id; w[n] = weight; n++;
      }
    }
  }
  const double coeff = 1.0/total;
  for (int k = 0; k < n; k++) w[k] *= coeff;
  return n;
}

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

// one node of the AoS velocity field (u0, u1, u2, rho) as ONE 256-bit load (LDG.E.256 on sm_100a): the
// interpolation is bound by the number of L1 requests of its scattered gathers, not by bytes
__device__ __forceinline__ void ld_node4(const double* p, double& a, double& b, double& c) {
  double d;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

// interpolation of one vertex: the 8 corners fully unrolled (registers only, no local-memory arrays), raw weights in
// the reference's x-outer / z-inner order, zero for corners outside the kernel, the domain or (CHECK_FLAGS) on
// non-fluid nodes.  Returns false when a candidate node is not addressable from this rank (velocity left alone).
template <bool CHECK_FLAGS>
__device__ __forceinline__ bool interp_vertex(const IbmArgs& a, const uint8_t* __restrict__ flags, const double* __restrict__ U,
                                              double px, double py, double pz, double& v0, double& v1, double& v2) {
  const int bx = (int)floor(px), by = (int)floor(py), bz = (int)floor(pz);
  double ax[2], ay[2], az[2]; int64_t jx[2]; int jy[2], jz[2];
  bool addressable = true;
#pragma unroll
  for (int d = 0; d < 2; d++) {
    ax[d] = phi2(px - (double)(bx + d)); jx[d] = 0;
    if (ax[d] != 0.0) {
      int lx; bool out;
      if (local_x(bx + d, a, lx, out)) jx[d] = (int64_t)lx*a.P;
      else { ax[d] = 0.0; if (!out) addressable = false; }
    }
    ay[d] = phi2(py - (double)(by + d)); int yy = by + d;
    if (ay[d] != 0.0 && !wrap_yz(yy, a.ny, a.py)) ay[d] = 0.0;
    jy[d] = yy*a.nz;
    az[d] = phi2(pz - (double)(bz + d)); int zz = bz + d;
    if (az[d] != 0.0 && !wrap_yz(zz, a.nz, a.pz)) az[d] = 0.0;
    jz[d] = zz;
  }
  if (!addressable) return false;
  double w[8]; double total = 0.0;
#pragma unroll
  for (int c = 0; c < 8; c++) {
    const int dx = c >> 2, dy = (c >> 1) & 1, dz = c & 1;
    w[c] = ax[dx]*ay[dy]*az[dz];
    if (w[c] == 0.0) continue;
    if (CHECK_FLAGS && flags[jx[dx] + jy[dy] + jz[dz]] != HCG_FLUID) { w[c] = 0.0; continue; }
    total += w[c];
  }
  const double coeff = 1.0/total;
  v0 = v1 = v2 = 0.0;
#pragma unroll
  for (int c = 0; c < 8; c++) {
    if (w[c] == 0.0) continue;
    const int dx = c >> 2, dy = (c >> 1) & 1, dz = c & 1;
    const double wn = w[c]*coeff;
    double u0, u1, u2;
    ld_node4(U + 4*(jx[dx] + jy[dy] + jz[dz]), u0, u1, u2);
    v0 += u0*wn; v1 += u1*wn; v2 += u2*wn;
  }
  return true;
}

template <bool ADVANCE, bool INTERP, bool CHECK_FLAGS>
__global__ void __launch_bounds__(256)
k_interp_advance(IbmArgs a, const uint8_t* __restrict__ flags, const int32_t* __restrict__ p_cell,
                 uint8_t* alive, double* x, double* y, double* z,
                 double* vx, double* vy, double* vz, const double* __restrict__ U,
                 const uint8_t* __restrict__ hold_back, int skip_held) {
  const int64_t p = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (p >= a.np) return;
  const int cell = p_cell[p];
  if (!alive[cell]) return;
  if (skip_held && hold_back[cell]) return;       // shared cells are interpolated by k_interp_list on the main stream
  double px = x[p], py = y[p], pz = z[p];
  double v0, v1, v2;
  if (INTERP) {
    if (interp_vertex<CHECK_FLAGS>(a, flags, U, px, py, pz, v0, v1, v2)) { vx[p] = v0; vy[p] = v1; vz[p] = v2; }
    else { v0 = vx[p]; v1 = vy[p]; v2 = vz[p]; }
  } else { v0 = vx[p]; v1 = vy[p]; v2 = vz[p]; }
  if (ADVANCE && !(hold_back && hold_back[cell])) {
    px += v0; py += v1; pz += v2;
    x[p] = px; y[p] = py; z[p] = pz;
    // particle on a boundary node => its cell is deleted (hemoCellParticleField.cpp:572-584, 512-553)
    if (CHECK_FLAGS) {                             // without non-fluid nodes no particle can land on one
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

hcg_status ibm_spread(hcg_ctx* c) {
  if (c->np == 0) return HCG_OK;
  IbmArgs a = make_args(c);
  k_spread<<<nblk(c->np, 256), 256, 0, c->stream>>>(a, c->flags, c->p_cell, c->cell_alive,
      c->pos[0], c->pos[1], c->pos[2], c->frc[0], c->frc[1], c->frc[2],
      c->frep[0], c->frep[1], c->frep[2], c->F);
  KERNEL_CHECK(c);
  return HCG_OK;
}

hcg_status ibm_interpolate(hcg_ctx* c) {
  if (c->np == 0) return HCG_OK;
  IbmArgs a = make_args(c);
  if (c->has_nonfluid) k_interp_advance<false, true, true><<<nblk(c->np, 256), 256, 0, c->stream>>>(a, c->flags, c->p_cell, c->cell_alive,
      c->pos[0], c->pos[1], c->pos[2], c->vel[0], c->vel[1], c->vel[2], c->U, nullptr, 0);
  else k_interp_advance<false, true, false><<<nblk(c->np, 256), 256, 0, c->stream>>>(a, c->flags, c->p_cell, c->cell_alive,
      c->pos[0], c->pos[1], c->pos[2], c->vel[0], c->vel[1], c->vel[2], c->U, nullptr, 0);
  KERNEL_CHECK(c);
  return HCG_OK;
}

hcg_status ibm_advance(hcg_ctx* c) {
  if (c->np == 0) return HCG_OK;
  IbmArgs a = make_args(c);
  if (c->has_nonfluid) k_interp_advance<true, false, true><<<nblk(c->np, 256), 256, 0, c->stream>>>(a, c->flags, c->p_cell, c->cell_alive,
      c->pos[0], c->pos[1], c->pos[2], c->vel[0], c->vel[1], c->vel[2], c->U, nullptr, 0);
  else k_interp_advance<true, false, false><<<nblk(c->np, 256), 256, 0, c->stream>>>(a, c->flags, c->p_cell, c->cell_alive,
      c->pos[0], c->pos[1], c->pos[2], c->vel[0], c->vel[1], c->vel[2], c->U, nullptr, 0);
  KERNEL_CHECK(c);
  return HCG_OK;
}

hcg_status ibm_interpolate_advance(hcg_ctx* c) {
  if (c->np == 0) return HCG_OK;
  IbmArgs a = make_args(c);
  if (c->has_nonfluid) k_interp_advance<true, true, true><<<nblk(c->np, 256), 256, 0, c->stream>>>(a, c->flags, c->p_cell, c->cell_alive,
      c->pos[0], c->pos[1], c->pos[2], c->vel[0], c->vel[1], c->vel[2], c->U, nullptr, 0);
  else k_interp_advance<true, true, false><<<nblk(c->np, 256), 256, 0, c->stream>>>(a, c->flags, c->p_cell, c->cell_alive,
      c->pos[0], c->pos[1], c->pos[2], c->vel[0], c->vel[1], c->vel[2], c->U, nullptr, 0);
  KERNEL_CHECK(c);
  return HCG_OK;
}

hcg_status ibm_interpolate_advance_unshared(hcg_ctx* c) {
  if (c->np == 0) return HCG_OK;
  if (!c->multi.d_cell_shared) return hcg_fail(c, HCG_ERR_STATE, "shared-cell flags missing (multi_rebalance has not run)");
  IbmArgs a = make_args(c);
  if (c->has_nonfluid) k_interp_advance<true, true, true><<<nblk(c->np, 256), 256, 0, c->stream>>>(a, c->flags, c->p_cell, c->cell_alive,
      c->pos[0], c->pos[1], c->pos[2], c->vel[0], c->vel[1], c->vel[2], c->U, c->multi.d_cell_shared, 0);
  else k_interp_advance<true, true, false><<<nblk(c->np, 256), 256, 0, c->stream>>>(a, c->flags, c->p_cell, c->cell_alive,
      c->pos[0], c->pos[1], c->pos[2], c->vel[0], c->vel[1], c->vel[2], c->U, c->multi.d_cell_shared, 0);
  KERNEL_CHECK(c);
  return HCG_OK;
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
  if (!c->multi.d_cell_shared) return hcg_fail(c, HCG_ERR_STATE, "shared-cell flags missing (multi_rebalance has not run)");
  IbmArgs a = make_args(c);
  if (c->has_nonfluid) k_interp_advance<true, true, true><<<nblk(c->np, 256), 256, 0, st>>>(a, c->flags, c->p_cell, c->cell_alive,
      c->pos[0], c->pos[1], c->pos[2], c->vel[0], c->vel[1], c->vel[2], c->U, c->multi.d_cell_shared, 1);
  else k_interp_advance<true, true, false><<<nblk(c->np, 256), 256, 0, st>>>(a, c->flags, c->p_cell, c->cell_alive,
      c->pos[0], c->pos[1], c->pos[2], c->vel[0], c->vel[1], c->vel[2], c->U, c->multi.d_cell_shared, 1);
  KERNEL_CHECK(c);
  return HCG_OK;
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
