// Per-corner arithmetic of the node-sorted spreading kernels (spread_sorted.cu): slab addressing and the node / raw phi2 weight of one
// (vertex, corner) pair.  Host + device code: inlined by the kernels on the device, compiled for the CPU by
// tests/cpp/ibm_node_host.cu (tests/test_ibm_node_host.py).
#pragma once
#include <stdint.h>
#include <math.h>
#include "../../include/hemocell_gpu.h"

struct SpArgs {
  int nx, ny, nz, px, py, pz;
  int nxl, x0, nranks;
  int64_t P;
  double f_limit;
  int V;
  int64_t first_cell, first_particle;
};

__host__ __device__ __forceinline__ bool sp_local_x(int gx, const SpArgs& a, int& lx, bool& outside) {
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
__host__ __device__ __forceinline__ bool sp_wrap(int& v, int n, int periodic) {
  if (v >= 0 && v < n) return true;
  if (!periodic) return false;
  v %= n; if (v < 0) v += n;
  return true;
}
__host__ __device__ __forceinline__ double sp_phi2(double x) { x = 1.0 - fabs(x); return x > 0.0 ? x : 0.0; }

// node (local index incl. ghosts) and raw weight of one corner; false if the corner carries nothing
__host__ __device__ __forceinline__ bool corner_node(const SpArgs& a, const uint8_t* __restrict__ flags, double px, double py,
                                            double pz, int corner, int& node, double& weight, bool& unaddressable) {
  unaddressable = false;
  const int dx = corner >> 2, dy = (corner >> 1) & 1, dz = corner & 1;
  const int bx = (int)floor(px) + dx, by = (int)floor(py) + dy, bz = (int)floor(pz) + dz;
  const double wx = sp_phi2(px - (double)bx), wy = sp_phi2(py - (double)by), wz = sp_phi2(pz - (double)bz);
  weight = wx*wy*wz;
  if (wx == 0.0) return false;
  int lx; bool out;
  if (!sp_local_x(bx, a, lx, out)) { if (!out) unaddressable = true; return false; }
  int y = by, z = bz;
  if (wy == 0.0 || !sp_wrap(y, a.ny, a.py)) return false;
  if (wz == 0.0 || !sp_wrap(z, a.nz, a.pz)) return false;
  if (weight == 0.0) return false;
  node = z + a.nz*(y + a.ny*lx);
  return flags[node] == HCG_FLUID;
}

