// Per-vertex arithmetic of the node-sorted spreading kernel (spread_sorted.cu): slab addressing, the record phase 1 stages per vertex
// and the node / weight phase 2 reads back from it per (vertex, corner) pair.  Host + device code: inlined by the kernel on the device, compiled for the CPU by
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

// ---- what k_spread_sorted stages per vertex (phase 1) and how a (vertex, corner) pair is read back from it (phase 2)
// node of corner (dx, dy, dz) = k0 + dx*(DX or WX) + dy*(DY or WY) + dz*(1 or WZ): the offset of the upper corner along an axis is the
// plane / row / node stride, or the way back to the start of a periodic axis
struct SpStrides { int DX, WX, DY, WY, WZ; };
__host__ __device__ __forceinline__ SpStrides sp_strides(const SpArgs& a) {
  SpStrides s; s.DX = a.ny*a.nz; s.WX = (1 - a.nxl)*s.DX; s.DY = a.nz; s.WY = -(a.ny - 1)*a.nz; s.WZ = -(a.nz - 1); return s;
}
struct SpVertex {
  int k0;            // node (local index incl. ghost planes) of the lower corner
  unsigned fl;       // bits 0-7: corners that add to a real node of this rank; bits 8 / 9 / 10: the x / y / z upper corner wraps
  double ab[4];      // wx[dx]*wy[dy] at 2*dx + dy
  double cz[2];      // wz[dz] / (sum of the admitted raw weights, reference accumulation order)
};
// check: look the corners' node flags up (false: no non-fluid node within the cell's reach)
template <bool CHECK_FLAGS>
__host__ __device__ __forceinline__ void sp_stage_vertex(const SpArgs& a, const SpStrides& st, const uint8_t* __restrict__ flags, bool check,
                                                double px, double py, double pz, SpVertex& o) {
  bool skip = false;
  const int bx = (int)floor(px), by = (int)floor(py), bz = (int)floor(pz);
  double ax[2], ay[2], az[2]; int jx[2], jy[2], jz[2]; bool realx[2];
#pragma unroll
  for (int d = 0; d < 2; d++) {
    ax[d] = sp_phi2(px - (double)(bx + d)); jx[d] = 0; realx[d] = false;
    if (ax[d] != 0.0) {
      int lx; bool out;
      if (sp_local_x(bx + d, a, lx, out)) { jx[d] = lx*a.ny*a.nz; realx[d] = lx >= 1 && lx <= a.nxl; }
      else { ax[d] = 0.0; if (!out) skip = true; }
    }
    ay[d] = sp_phi2(py - (double)(by + d)); int yy = by + d;
    if (ay[d] != 0.0 && !sp_wrap(yy, a.ny, a.py)) ay[d] = 0.0;
    jy[d] = yy*a.nz;
    az[d] = sp_phi2(pz - (double)(bz + d)); int zz = bz + d;
    if (az[d] != 0.0 && !sp_wrap(zz, a.nz, a.pz)) az[d] = 0.0;
    jz[d] = zz;
  }
  double total = 0.0; unsigned mask = 0;
#pragma unroll
  for (int c = 0; c < 8; c++) {               // corner order == the reference's x-outer / z-inner order
    const int dx = c >> 2, dy = (c >> 1) & 1, dz = c & 1;
    const double w = ax[dx]*ay[dy]*az[dz];
    if (w == 0.0) continue;
    if (CHECK_FLAGS && check && flags[jx[dx] + jy[dy] + jz[dz]] != HCG_FLUID) continue;
    total += w;
    if (realx[dx]) mask |= 1u << c;            // ghost planes count in the normalisation only
  }
  const double co = 1.0/total;
#pragma unroll
  for (int d = 0; d < 2; d++) { o.ab[2*d] = ax[d]*ay[0]; o.ab[2*d + 1] = ax[d]*ay[1]; o.cz[d] = az[d]*co; }
  // lower x corner not addressable (left of a non-periodic domain): its plane is the virtual one below the upper corner's
  const int kx0 = ax[0] != 0.0 ? jx[0] : jx[1] - st.DX;
  unsigned fl = skip ? 0u : mask;               // multi-GPU: a candidate node is not addressable here
  if (ax[1] != 0.0 && jx[1] - kx0 != st.DX) {
#ifdef __CUDA_ARCH__
    if (jx[1] - kx0 != st.WX) __trap();          // (no third plane offset exists)
#endif
    fl |= 1u << 8;
  }
  if (jy[1] - jy[0] != st.DY) fl |= 1u << 9;
  if (jz[1] - jz[0] != 1) fl |= 1u << 10;
  o.k0 = kx0 + jy[0] + jz[0]; o.fl = fl;
}
// corner c of a staged vertex: false if it adds nothing here, else its node
__host__ __device__ __forceinline__ bool sp_pair_node(const SpStrides& st, int k0, unsigned fl, int c, int& node) {
  if (!((fl >> c) & 1u)) return false;
  node = k0 + ((c & 4) ? ((fl & 0x100u) ? st.WX : st.DX) : 0) + ((c & 2) ? ((fl & 0x200u) ? st.WY : st.DY) : 0)
            + ((c & 1) ? ((fl & 0x400u) ? st.WZ : 1) : 0);
  return true;
}
