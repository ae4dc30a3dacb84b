// Host-side harness of the immersed-boundary kernels' per-particle arithmetic: compiles hemocell_b200/csrc/ibm_node.cuh (host + device
// code, inlined by k_spread / k_interp_advance / k_interp_list on the device) for the CPU (tests/test_ibm_node_host.py).
#include "../../hemocell_b200/csrc/ibm_node.cuh"
#include "../../hemocell_b200/csrc/spread_node.cuh"

static IbmArgs make(int nx, int ny, int nz, const int* periodic, int x0, int nxl, int nranks) {
  IbmArgs a;
  a.nx = nx; a.ny = ny; a.nz = nz; a.px = periodic[0]; a.py = periodic[1]; a.pz = periodic[2];
  a.nxl = nxl; a.x0 = x0; a.nranks = nranks;
  a.P = (int64_t)ny*nz; a.S = (int64_t)(nxl + 2)*a.P; a.np = 0; a.f_limit = 1e300;
  return a;
}

// flags / U: padded slab (ghost plane on each x side), U is AoS [n][4].  vel [np][3]; ok[p] = 0 when a corner is not addressable.
extern "C" void ibm_interp_host(int nx, int ny, int nz, const int* periodic, int x0, int nxl, int nranks, int check_flags,
                                const uint8_t* flags, const double* U, int64_t np, const double* pos, double* vel, uint8_t* ok) {
  const IbmArgs a = make(nx, ny, nz, periodic, x0, nxl, nranks);
  for (int64_t p = 0; p < np; p++) {
    double v0 = 0, v1 = 0, v2 = 0;
    const bool r = check_flags ? interp_vertex<true>(a, flags, U, pos[3*p], pos[3*p+1], pos[3*p+2], v0, v1, v2)
                               : interp_vertex<false>(a, flags, U, pos[3*p], pos[3*p+1], pos[3*p+2], v0, v1, v2);
    ok[p] = r; vel[3*p] = v0; vel[3*p+1] = v1; vel[3*p+2] = v2;
  }
}
// the (node, weight) pairs of one particle as the spreading fallback kernel uses them; returns the count (-1: not addressable)
extern "C" int ibm_kernel_host(int nx, int ny, int nz, const int* periodic, int x0, int nxl, int nranks,
                               const uint8_t* flags, const double* p3, int64_t* node, double* w) {
  const IbmArgs a = make(nx, ny, nz, periodic, x0, nxl, nranks);
  return ibm_kernel<true>(a, flags, p3[0], p3[1], p3[2], node, w);
}

// the 8 (vertex, corner) pairs of one particle as k_spread_sorted sees them: phase 1 stages the vertex (sp_stage_vertex), phase 2 reads
// node and normalised weight of a corner back (sp_pair_node, ab * cz).  valid[c]: the corner adds to a REAL node of this slab
// (ghost-plane corners count in the normalisation only); returns the staged flag word.
extern "C" unsigned spread_corners_host(int nx, int ny, int nz, const int* periodic, int x0, int nxl, int nranks,
                                        const uint8_t* flags, const double* p3, uint8_t* valid, int64_t* node, double* weight) {
  SpArgs a;
  a.nx = nx; a.ny = ny; a.nz = nz; a.px = periodic[0]; a.py = periodic[1]; a.pz = periodic[2];
  a.nxl = nxl; a.x0 = x0; a.nranks = nranks; a.P = (int64_t)ny*nz; a.f_limit = 1e300; a.V = 1; a.first_cell = 0; a.first_particle = 0;
  const SpStrides st = sp_strides(a);
  SpVertex sv;
  sp_stage_vertex<true>(a, st, flags, true, p3[0], p3[1], p3[2], sv);
  for (int c = 0; c < 8; c++) {
    int nd = -1;
    valid[c] = sp_pair_node(st, sv.k0, sv.fl, c, nd);
    node[c] = nd; weight[c] = valid[c] ? sv.ab[c >> 1]*sv.cz[c & 1] : 0.0;
  }
  return sv.fl;
}
