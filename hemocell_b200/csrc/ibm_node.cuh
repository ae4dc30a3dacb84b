// Per-particle arithmetic of the immersed-boundary kernels (ibm.cu): node addressing on an x-slab with ghost planes, the phi2 kernel
// of one particle, the unrolled velocity interpolation of one vertex.  Host + device code: inlined by the kernels on the device,
// compiled for the CPU by tests/cpp/ibm_node_host.cu so that the CPU suite checks it against the oracle (tests/test_ibm_node_host.py).
#pragma once
#include <stdint.h>
#include <math.h>
#include "../../include/hemocell_gpu.h"

struct IbmArgs {
  int nx, ny, nz, px, py, pz;
  int nxl, x0, nranks;
  int64_t P, S, np;
  double f_limit;
};

// global (unwrapped) node x -> local plane index incl. ghosts; false if not held by this rank
// or outside a non-periodic domain
__host__ __device__ __forceinline__ bool local_x(int gx, const IbmArgs& a, int& lx, bool& outside) {
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
__host__ __device__ __forceinline__ bool wrap_yz(int& v, int n, int periodic) {
  if (v >= 0 && v < n) return true;
  if (!periodic) return false;
  v %= n; if (v < 0) v += n;
  return true;
}
__host__ __device__ __forceinline__ double phi2(double x) { x = 1.0 - fabs(x); return x > 0.0 ? x : 0.0; }

// Kernel of one particle: up to 8 (node, weight) pairs in the reference's x-outer/z-inner
// order, zero weights and boundary nodes skipped, normalised.  Returns the count, or -1 when a
// candidate node is not addressable from this rank (particle irrelevant here).
template <bool CHECK_FLAGS = true>
__host__ __device__ __forceinline__ int ibm_kernel(const IbmArgs& a, const uint8_t* __restrict__ flags,
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

// one node of the AoS velocity field (u0, u1, u2, rho) as ONE 256-bit load (LDG.E.256 on sm_100a): the
// interpolation is bound by the number of L1 requests of its scattered gathers, not by bytes
__host__ __device__ __forceinline__ void ld_node4(const double* p, double& a, double& b, double& c) {
#ifdef __CUDA_ARCH__
  double d; (void)d;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
#else
  a = p[0]; b = p[1]; c = p[2];       // host build (tests/cpp/ibm_node_host.cu)
#endif
}

// interpolation of one vertex: the 8 corners fully unrolled (registers only, no local-memory arrays), raw weights in
// the reference's x-outer / z-inner order, zero for corners outside the kernel, the domain or (CHECK_FLAGS) on
// non-fluid nodes.  Returns false when a candidate node is not addressable from this rank (velocity left alone).
template <bool CHECK_FLAGS>
__host__ __device__ __forceinline__ bool interp_vertex(const IbmArgs& a, const uint8_t* __restrict__ flags, const double* __restrict__ U,
                                              double px, double py, double pz, double& v0, double& v1, double& v2, bool check = true) {
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
    if (CHECK_FLAGS && check && flags[jx[dx] + jy[dy] + jz[dz]] != HCG_FLUID) { w[c] = 0.0; continue; }
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
