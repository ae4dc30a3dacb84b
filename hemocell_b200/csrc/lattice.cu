// K1: fused D3Q19 Guo-BGK stream(pull)+collide, moments pass, halo exchange, layout conversion.
// Replaces Palabos MultiBlockLattice3D<double,ForcedD3Q19Descriptor>::collideAndStream() with
// GuoExternalForceBGKdynamics / BounceBack / regularized velocity planes as HemoCell drives it
// (reference core/hemoCell.cpp:317; arithmetic restated in SURVEY.md Appendix C).
#include "ctx.cuh"
#include "moment_step.cuh"
#include "lattice_node.cuh"
#include <nccl.h>
#include <cfloat>
#include <algorithm>
#include <type_traits>
#include <cstdlib>
#include <cstring>
#include <cstdio>

namespace {

// (per-node arithmetic - feq, moments19, guo_collide, guo_collide_tau1, regularized_complete, zouhe_complete - lives in lattice_node.cuh)
struct LatArgs {
  int nxl, ny, nz, py, pz;
  int64_t P, S;
  int nxlL;             // planes of the LEFT slab neighbour (uneven decompositions: may differ from nxl by one)
  int64_t SL, SR;       // padded slab sizes (population stride) of the left / right neighbour
  double omega;
  const double* bc;     // [6][3] wall velocity per orientation, device memory
  const double* bcn;    // per-node boundary values of the Zou-He nodes, AoS [n][4] = (u_x, u_y, u_z, rho); null without such nodes
  double body[3];
};


// pull the 19 post-stream populations of node (lx,y,z): S_q(n) = g_q(n - c_q).
// x always has a ghost plane; y/z wrap when periodic, else the value entering through the face
// is the rest equilibrium (stored 0).
__device__ __forceinline__ void pull19(const double* __restrict__ g, const LatArgs& a, int64_t n,
                                       int y, int z, double f[19]) {
  constexpr int CX[19] = {0,-1,0,0,-1,-1,-1,-1,0,0, 1,0,0,1,1,1,1,0,0};
  constexpr int CY[19] = {0,0,-1,0,-1,1,0,0,-1,-1, 0,1,0,1,-1,0,0,1,1};
  constexpr int CZ[19] = {0,0,0,-1,0,0,-1,1,-1,1, 0,0,1,0,0,1,-1,1,-1};
  // offsets for source = this - c: dy index 0 -> source y+1 (cy=-1), 2 -> source y-1 (cy=+1)
  const int nz = a.nz, ny = a.ny;
  int64_t oyp, oym, ozp, ozm; bool vyp = true, vym = true, vzp = true, vzm = true;
  if (y + 1 < ny) oyp = nz; else { oyp = -(int64_t)(ny - 1)*nz; vyp = a.py; }
  if (y > 0) oym = -nz; else { oym = (int64_t)(ny - 1)*nz; vym = a.py; }
  if (z + 1 < nz) ozp = 1; else { ozp = -(nz - 1); vzp = a.pz; }
  if (z > 0) ozm = -1; else { ozm = nz - 1; vzm = a.pz; }
#pragma unroll
  for (int q = 0; q < 19; q++) {
    int64_t off = n - (int64_t)CX[q]*a.P;
    bool ok = true;
    if (CY[q] == 1) { off += oym; ok = ok && vym; } else if (CY[q] == -1) { off += oyp; ok = ok && vyp; }
    if (CZ[q] == 1) { off += ozm; ok = ok && vzm; } else if (CZ[q] == -1) { off += ozp; ok = ok && vzp; }
    f[q] = ok ? __ldg(g + (int64_t)q*a.S + off) : 0.0;
  }
}





// boundary completion of a non-fluid, non-bounce-back node.  BC: 1 = regularized velocity planes only,
// 2 = also Zou-He velocity / pressure nodes with per-node values (a.bcn).
template <int BC>
__device__ __forceinline__ void complete_boundary(double f[19], const LatArgs& a, int64_t n, uint8_t fl) {
  if (BC == 2 && fl >= HCG_ZH_VEL_XN) {
    const double2 b01 = *reinterpret_cast<const double2*>(a.bcn + 4*n), b23 = *reinterpret_cast<const double2*>(a.bcn + 4*n + 2);
    if (fl >= HCG_ZH_PRES_XN) zouhe_complete(f, fl - HCG_ZH_PRES_XN, true, b01.x, b01.y, b23.x, b23.y);
    else zouhe_complete(f, fl - HCG_ZH_VEL_XN, false, b01.x, b01.y, b23.x, b23.y);
  } else {
    const double uw[3] = {a.bc[3*(fl-2)], a.bc[3*(fl-2)+1], a.bc[3*(fl-2)+2]};
    regularized_complete(f, fl - 2, uw);
  }
}

// node velocity / density of a boundary node as Cell::computeVelocity reports it (moments pass)
template <bool IOBC>
__device__ __forceinline__ void boundary_velocity(const double f[19], const LatArgs& a, int64_t n, uint8_t fl,
                                                  double& u0, double& u1, double& u2, double& rho) {
  if (IOBC && fl >= HCG_ZH_VEL_XN) {
    const double2 b01 = *reinterpret_cast<const double2*>(a.bcn + 4*n), b23 = *reinterpret_cast<const double2*>(a.bcn + 4*n + 2);
    if (fl >= HCG_ZH_PRES_XN) {
      constexpr int CC[19][3] = {{0,0,0},{-1,0,0},{0,-1,0},{0,0,-1},{-1,-1,0},{-1,1,0},{-1,0,-1},{-1,0,1},{0,-1,-1},{0,-1,1},
                                 {1,0,0},{0,1,0},{0,0,1},{1,1,0},{1,-1,0},{1,0,1},{1,0,-1},{0,1,1},{0,1,-1}};
      constexpr double TW[19] = {1.0/3.0, 1.0/18.0,1.0/18.0,1.0/18.0,1.0/36.0,1.0/36.0,1.0/36.0,1.0/36.0,1.0/36.0,1.0/36.0,
                                 1.0/18.0,1.0/18.0,1.0/18.0,1.0/36.0,1.0/36.0,1.0/36.0,1.0/36.0,1.0/36.0,1.0/36.0};
      const int o = fl - HCG_ZH_PRES_XN, dir = o >> 1, sgn = (o & 1) ? 1 : -1;
      double rho_on = 0.0, rho_out = 0.0;
#pragma unroll
      for (int q = 0; q < 19; q++) {
        const int cd = dir == 0 ? CC[q][0] : (dir == 1 ? CC[q][1] : CC[q][2]);
        const int cn = cd*sgn;
        if (cn == 0) rho_on += f[q] + TW[q]; else if (cn > 0) rho_out += f[q] + TW[q];
      }
      rho = b23.y;
      const double un = sgn*((rho_on + 2.0*rho_out)/rho - 1.0);
      u0 = dir == 0 ? un : 0.0; u1 = dir == 1 ? un : 0.0; u2 = dir == 2 ? un : 0.0;
    } else { u0 = b01.x; u1 = b01.y; u2 = b23.x; }
  } else { u0 = a.bc[3*(fl-2)]; u1 = a.bc[3*(fl-2)+1]; u2 = a.bc[3*(fl-2)+2]; }
}

// One thread per real node.  RESET: write the body force back after reading F (steps without
// velocity interpolation).
// PEER (multi-GPU, csrc/peer.cu): the face planes additionally store their 5 outgoing populations
// into the slab neighbour's ghost plane over NVLink (peerL / peerR = the neighbours' output buffers),
// which replaces the separate halo exchange.
template <bool RESET, int BC /*0 none, 1 regularized planes, 2 + Zou-He nodes*/, int MINB, bool PEER>
__global__ void __launch_bounds__(256, MINB)
k_collide_stream(const double* __restrict__ gin, double* __restrict__ gout, double* __restrict__ F,
                 const uint8_t* __restrict__ flags, LatArgs a, int64_t first, int64_t count,
                 double* __restrict__ peerL, double* __restrict__ peerR) {
  const int64_t k = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (k >= count) return;
  const int64_t i = first + k;
  const int64_t n = i + a.P;
  const int rem = (int)(i % a.P);
  const int y = rem / a.nz, z = rem - y*a.nz;
  double f[19];
  pull19(gin, a, n, y, z, f);
  const uint8_t fl = flags[n];
  if (fl == HCG_BOUNCEBACK) {
#pragma unroll
    for (int q = 1; q <= 9; q++) { const double t = f[q]; f[q] = f[q+9]; f[q+9] = t; }
  } else {
    const double2 fa = *reinterpret_cast<const double2*>(F + 4*n);
    double Fn[3] = {fa.x, fa.y, F[4*n + 2]};
    if (BC >= 1 && fl >= HCG_VEL_XN) complete_boundary<BC>(f, a, n, fl);
    guo_collide(f, Fn, a.omega);
  }
  if (RESET) { double2* Fw = reinterpret_cast<double2*>(F + 4*n); Fw[0] = make_double2(a.body[0], a.body[1]); Fw[1] = make_double2(a.body[2], 0.0); }
#pragma unroll
  for (int q = 0; q < 19; q++) gout[(int64_t)q*a.S + n] = f[q];
  if (PEER) {
    // my first real plane -> left neighbour's RIGHT ghost (c_x = -1 set); my last -> right neighbour's LEFT ghost (c_x = +1 set)
    if (i < a.P && peerL) {
      const int64_t off = (int64_t)(a.nxlL + 1)*a.P + rem;
      peerL[1*a.SL + off] = f[1]; peerL[4*a.SL + off] = f[4]; peerL[5*a.SL + off] = f[5]; peerL[6*a.SL + off] = f[6]; peerL[7*a.SL + off] = f[7];
    }
    if (i >= (int64_t)(a.nxl - 1)*a.P && peerR) {
      const int64_t off = rem;
      peerR[10*a.SR + off] = f[10]; peerR[13*a.SR + off] = f[13]; peerR[14*a.SR + off] = f[14]; peerR[15*a.SR + off] = f[15]; peerR[16*a.SR + off] = f[16];
    }
  }
}

#ifndef TAU1_MINB
#define TAU1_MINB 4
#endif
// tau = 1 fast path (guo_collide_tau1 in lattice_node.cuh): fluid nodes read the raw moments W and the force instead of pulling 19 populations

template <bool RESET, int BC, bool PEER>
__global__ void __launch_bounds__(256, TAU1_MINB)
k_collide_tau1(const double* __restrict__ gin, double* __restrict__ gout, double* __restrict__ F,
               const double* __restrict__ W, const uint8_t* __restrict__ flags, LatArgs a, int64_t first, int64_t count,
               double* __restrict__ peerL, double* __restrict__ peerR) {
  const int64_t k = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (k >= count) return;
  const int64_t i = first + k;
  const int64_t n = i + a.P;
  const int rem = (int)(i % a.P);
  double f[19];
  const uint8_t fl = flags[n];
  if (fl == HCG_FLUID) {
    double w0, w1, w2, w3, f0, f1, f2, f3;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(w0), "=d"(w1), "=d"(w2), "=d"(w3) : "l"(W + 4*n));
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(f0), "=d"(f1), "=d"(f2), "=d"(f3) : "l"(F + 4*n) : "memory");
    const double j[3] = {w1, w2, w3}, Fn[3] = {f0, f1, f2};
    guo_collide_tau1(f, w0, j, Fn);
  } else {
    const int y = rem / a.nz, z = rem - y*a.nz;
    pull19(gin, a, n, y, z, f);
    if (fl == HCG_BOUNCEBACK) {
#pragma unroll
      for (int q = 1; q <= 9; q++) { const double t = f[q]; f[q] = f[q+9]; f[q+9] = t; }
    } else {
      const double2 fa = *reinterpret_cast<const double2*>(F + 4*n);
      double Fn[3] = {fa.x, fa.y, F[4*n + 2]};
      if (BC >= 1 && fl >= HCG_VEL_XN) complete_boundary<BC>(f, a, n, fl);
      guo_collide(f, Fn, a.omega);
    }
  }
  if (RESET) { double2* Fw = reinterpret_cast<double2*>(F + 4*n); Fw[0] = make_double2(a.body[0], a.body[1]); Fw[1] = make_double2(a.body[2], 0.0); }
#pragma unroll
  for (int q = 0; q < 19; q++) gout[(int64_t)q*a.S + n] = f[q];
  if (PEER) {
    if (i < a.P && peerL) {
      const int64_t off = (int64_t)(a.nxlL + 1)*a.P + rem;
      peerL[1*a.SL + off] = f[1]; peerL[4*a.SL + off] = f[4]; peerL[5*a.SL + off] = f[5]; peerL[6*a.SL + off] = f[6]; peerL[7*a.SL + off] = f[7];
    }
    if (i >= (int64_t)(a.nxl - 1)*a.P && peerR) {
      const int64_t off = rem;
      peerR[10*a.SR + off] = f[10]; peerR[13*a.SR + off] = f[13]; peerR[14*a.SR + off] = f[14]; peerR[15*a.SR + off] = f[15]; peerR[16*a.SR + off] = f[16];
    }
  }
}

// Moments pass: velocity the IBM interpolation sees, u = j/rho + F/2 of the POST-stream
// populations with the spread force still on the node (Cell::computeVelocity through
// core/hemoCellParticleField.cpp:833).  BounceBack: 0; velocity plane: wall velocity.
// WOUT (tau = 1 fast path, see k_collide_tau1): also keep the raw moments (rhoBar, j) of the node in W.
template <bool RESET, bool PEER, bool WOUT, bool IOBC>
__global__ void __launch_bounds__(256)
k_moments(const double* __restrict__ g, double* __restrict__ F, double* __restrict__ U,
          const uint8_t* __restrict__ flags, LatArgs a, int64_t first, int64_t count,
          double* __restrict__ peerL, double* __restrict__ peerR, double* __restrict__ W) {
  const int64_t k = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (k >= count) return;
  const int64_t i = first + k;
  const int64_t n = i + a.P;
  const int rem = (int)(i % a.P);
  const int y = rem / a.nz, z = rem - y*a.nz;
  double f[19];
  pull19(g, a, n, y, z, f);
  double rhoBar, j[3];
  moments19(f, rhoBar, j);
  if (WOUT) { double2* Ww = reinterpret_cast<double2*>(W + 4*n); Ww[0] = make_double2(rhoBar, j[0]); Ww[1] = make_double2(j[1], j[2]); }
  const uint8_t fl = flags[n];
  double u0, u1, u2, rho = 1.0 + rhoBar;
  if (fl == HCG_FLUID) {
    const double invRho = 1.0/rho;
    const double2 fa = *reinterpret_cast<const double2*>(F + 4*n);
    u0 = j[0]*invRho + 0.5*fa.x; u1 = j[1]*invRho + 0.5*fa.y; u2 = j[2]*invRho + 0.5*F[4*n + 2];
  } else if (fl == HCG_BOUNCEBACK) { u0 = u1 = u2 = 0.0; rho = 1.0; }
  else boundary_velocity<IOBC>(f, a, n, fl, u0, u1, u2, rho);
  double2* Uw = reinterpret_cast<double2*>(U + 4*n);
  Uw[0] = make_double2(u0, u1); Uw[1] = make_double2(u2, rho);      // slot 3 carries the density
  if (PEER) {                                                        // node velocity of the face planes -> neighbours' ghost planes
    if (i < a.P && peerL) { double2* Pw = reinterpret_cast<double2*>(peerL + 4*((int64_t)(a.nxlL + 1)*a.P + rem)); Pw[0] = make_double2(u0, u1); Pw[1] = make_double2(u2, rho); }
    if (i >= (int64_t)(a.nxl - 1)*a.P && peerR) { double2* Pw = reinterpret_cast<double2*>(peerR + 4*(int64_t)rem); Pw[0] = make_double2(u0, u1); Pw[1] = make_double2(u2, rho); }
  }
  if (RESET) { double2* Fw = reinterpret_cast<double2*>(F + 4*n); Fw[0] = make_double2(a.body[0], a.body[1]); Fw[1] = make_double2(a.body[2], 0.0); }
}

// ---------------------------------------------------------------------------------------------
// Moment-only update at tau = 1 (EXPERIMENTAL, opt-in: HCG_MOMENT_ONLY=1 or hcg_set_moment_only; written at the end of round 1,
// NOT yet run or measured on a GPU).  With omega = 1 the post-collision population f*_q(n) is a function of the four raw
// moments and the force of node n alone (guo_collide_tau1), and the next post-stream moments of node m are sums over
// S_q(m) = f*_q(m - c_q).  On a fully periodic lattice without walls the populations therefore need not be stored at all: one
// kernel reads (W, F) of the 19 upstream neighbours (each node's 64 B come from HBM once and 18 more times from L1 / L2),
// evaluates the 19 arriving populations in the summation order of moments19 and writes the new moments, the node velocity
// the IBM interpolation reads (on interpolation steps) and the reset force: 64 B read + 96 B written per lattice update
// instead of 216 + 280.  W and F are double-buffered (a neighbour still reads the old force while this node resets it); the
// previous pair stays intact until the next step, so the populations can be materialised on demand (lat_ensure_pops) for
// downloads, output, checkpoints and any operator that needs them.
template <bool WRITE_U>
__global__ void __launch_bounds__(256)
k_moment_step(const double* __restrict__ Win, const double* Fin, double* __restrict__ Wout, double* __restrict__ Fout,
              double* __restrict__ U, LatArgs a, int64_t count) {
  const int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= count) return;
  MomentArgs m; m.ny = a.ny; m.nz = a.nz; m.P = a.P; m.body[0] = a.body[0]; m.body[1] = a.body[1]; m.body[2] = a.body[2];
  moment_node<WRITE_U>(Win, Fin, Wout, Fout, U, m, i);      // csrc/moment_step.cuh (also compiled for the host by tests/cpp/moment_host.cu)
}

// Tile-marching form of the same update (default).  k_moment_step re-evaluates each neighbour's prologue 19 times and pulls
// 38 x 32 B per node through L1 (measured: 1.20 ms at 256^3, L1 data pipe 63 %, fp64 pipe 39 %).  Here a CTA owns a (TY x TZ)
// column of the (y, z) plane plus a one-node halo and marches along x: per plane every thread evaluates the 19 post-collision
// populations of ITS node once (tau1_pops_fast, ~100 fp64 operations) and hands them to the neighbours through shared memory
// (19 stores + 19 loads of 8 B per node instead of 1216 B of L1 traffic); the populations with c_x = +1 / 0 / -1 are added
// to the moment accumulators of the planes x + 1 / x / x - 1, which live in registers, so plane x - 1 is complete - and written
// (new moments, node velocity, reset force: 96 B) - as soon as plane x has been evaluated.  Shared memory is double-buffered:
// one __syncthreads per plane.  HBM traffic = 64 B read (+ halo re-reads served by L2) + 96 B written per lattice update.
// mbarrier / bulk-async (1-D TMA) helpers, shared by k_moment_tile and the row-pipelined kernels below
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
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
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// one 32-byte node record as ONE 256-bit store (STG.E.256 on sm_100a): half the store requests of two 128-bit stores, whole sectors
__device__ __forceinline__ void st_node4(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
struct MomentPeer {      // slab neighbours' buffers (null: read the own ghost planes, filled by the send/recv exchange)
  const double *WL, *FL, *WR, *FR;   // the left neighbour's last real plane / the right neighbour's first real plane of (W, F)
  double *UL, *UR;                   // the left neighbour's right ghost plane / the right neighbour's left ghost plane of U
};
template <bool WRITE_U, int TY, int TZ, int MINB>
__global__ void __launch_bounds__(((TY + 2)*(TZ + 2) + 31)/32*32, MINB)
k_moment_tile(const double* __restrict__ Win, const double* __restrict__ Fin, double* __restrict__ Wout, double* __restrict__ Fout,
              double* __restrict__ U, LatArgs a, int xc, MomentPeer pr) {
  constexpr int HY = TY + 2, HZ = TZ + 2, HN = HY*HZ;
  // shared memory: the 19 populations of the current plane [19][HN], then a two-deep ring of the (W, F) inputs of the planes
  // ahead as node records [2 stages][W | F][HY][HZ][4], filled row by row with bulk-async copies (1-D TMA: one request per
  // 1 KB row instead of four 16-byte cp.async per node, whole sectors, no registers held while the loads fly)
  extern __shared__ __align__(128) double sm_pop[];
  constexpr int NT = (HN + 31)/32*32;
  double* ring = sm_pop + 19*HN;                         // [2][2][HN][4]
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + 2*2*HN*4);
  const int t = threadIdx.x;
  const int hy = t / HZ, hz = t - hy*HZ;
  const int y = (int)blockIdx.y*TY + hy - 1, z = (int)blockIdx.x*TZ + hz - 1;
  const int ny = a.ny, nz = a.nz;
  const bool halo_ok = t < HN && y <= ny && z <= nz;
  const bool interior = halo_ok && hy >= 1 && hy <= TY && hz >= 1 && hz <= TZ && y < ny && z < nz;
  const int yw = y < 0 ? ny - 1 : (y >= ny ? 0 : y), zw = z < 0 ? nz - 1 : (z >= nz ? 0 : z);
  const int64_t col = (int64_t)yw*nz + zw;
  const int x_lo = 1 + (int)blockIdx.z*xc, x_hi = min(x_lo + xc - 1, a.nxl);
  if (x_lo > a.nxl) return;
  if (t == 0) {
    mbar_init(full, 1); mbar_init(full + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // the row of halo'd nodes a lane of warp 0 fetches: y row `t` of the tile (z0 - 1 .. z0 + TZ, wrapped at the ends of the z axis)
  const int z0 = (int)blockIdx.x*TZ;
  const int zmain0 = max(z0 - 1, 0), zmain1 = min(z0 + TZ, nz - 1);            // contiguous part [zmain0, zmain1]
  const bool wrapL = z0 == 0, wrapR = z0 + TZ >= nz;                            // hz = 0 <- node nz - 1; hz = nz - z0 + 1 <- node 0
  const int rows_ok = min(HY, ny - (int)blockIdx.y*TY + 2);                     // rows with y <= ny
  const uint32_t row_bytes = 32u*(uint32_t)((zmain1 - zmain0 + 1) + (wrapL ? 1 : 0) + (wrapR ? 1 : 0));
  auto fetch = [&](int lx) {                             // (W, F) of the tile's nodes on plane lx -> ring stage lx & 1 (warp 0 issues)
    if (t >= 32 || lx > x_hi + 1) return;
    uint64_t* bar = full + (lx & 1);
    if (t == 0) mbar_expect_tx(bar, 2u*row_bytes*(uint32_t)rows_ok);
    __syncwarp();
    if (t < rows_ok) {
      const int ry = (int)blockIdx.y*TY + t - 1;
      const int ryw = ry < 0 ? ny - 1 : (ry >= ny ? 0 : ry);
      const double* w = Win + 4*(int64_t)lx*a.P; const double* f = Fin + 4*(int64_t)lx*a.P;
      // slab faces (multi-GPU, peer transport): the ghost planes are the neighbours' own face planes, read over NVLink
      if (lx == 0 && pr.WL) { w = pr.WL; f = pr.FL; }
      else if (lx == a.nxl + 1 && pr.WR) { w = pr.WR; f = pr.FR; }
      const int64_t rowoff = 4*(int64_t)ryw*nz;
      double* dw = ring + (size_t)(lx & 1)*2*HN*4 + (size_t)t*HZ*4; double* df = dw + HN*4;
      const int hz0 = zmain0 - (z0 - 1);                 // slot of the first contiguous node
      const uint32_t nb = 32u*(uint32_t)(zmain1 - zmain0 + 1);
      bulk_g2s(dw + 4*hz0, w + rowoff + 4*zmain0, nb, bar); bulk_g2s(df + 4*hz0, f + rowoff + 4*zmain0, nb, bar);
      if (wrapL) { bulk_g2s(dw, w + rowoff + 4*(int64_t)(nz - 1), 32u, bar); bulk_g2s(df, f + rowoff + 4*(int64_t)(nz - 1), 32u, bar); }
      if (wrapR) { const int hzr = nz - z0 + 1; bulk_g2s(dw + 4*hzr, w + rowoff, 32u, bar); bulk_g2s(df + 4*hzr, f + rowoff, 32u, bar); }
    }
  };
  fetch(x_lo - 1);
  fetch(x_lo);
  // moment accumulators (rhoBar, j) of the planes lx - 1 and lx; this column's own force on plane lx - 1
  double a0r = 0, a0x = 0, a0y = 0, a0z = 0, a1r = 0, a1x = 0, a1y = 0, a1z = 0;
  double p0 = 0, p1 = 0, p2 = 0;
  const int i = hy*HZ + hz;
  double* sb = sm_pop;
  for (int lx = x_lo - 1; lx <= x_hi + 1; lx++) {
    mbar_wait(full + (lx & 1), (uint32_t)(((lx - x_lo + 1) >> 1) & 1));   // plane lx has landed (the copies of plane lx + 1 may still fly)
    double c0 = 0, c1 = 0, c2 = 0;
    if (halo_ok) {
      // the records lie 32 bytes apart: lanes 4..7 of every group of eight read their two 16-byte halves in the opposite
      // order, so that each 128-bit load of a quarter-warp touches all 32 banks once (no 2-way conflict)
      const double2* sl = reinterpret_cast<const double2*>(ring + (size_t)(lx & 1)*2*HN*4 + (size_t)t*4);
      const int h = (t >> 2) & 1;
      const double2 w_first = sl[h], w_second = sl[1 - h], f_first = sl[2*HN + h], f_second = sl[2*HN + 1 - h];
      const double2 wa = h ? w_second : w_first, wb = h ? w_first : w_second, fa = h ? f_second : f_first, fb = h ? f_first : f_second;
      c0 = fa.x; c1 = fa.y; c2 = fb.x;
      double p[19];
      tau1_pops_fast(wa.x, wa.y, wb.x, wb.y, fa.x, fa.y, fb.x, p);
#pragma unroll
      for (int q = 0; q < 19; q++) sb[q*HN + t] = p[q];
    }
    __syncthreads();
    fetch(lx + 2);                                       // into the stage just consumed (every thread's reads of it are behind the barrier)
    if (interior) {
      // arriving population q comes from the node at -c_q: index i - c_y HZ - c_z
      const double q1 = sb[1*HN + i], q4 = sb[4*HN + i + HZ], q5 = sb[5*HN + i - HZ], q6 = sb[6*HN + i + 1], q7 = sb[7*HN + i - 1];
      // c_x = -1 -> plane lx - 1 (complete now)
      const double sm = ((q1 + q4) + (q5 + q6)) + q7;
      a0r += sm; a0x -= sm; a0y += q5 - q4; a0z += q7 - q6;
      if (lx - 1 >= x_lo) {
        const int64_t n = (int64_t)(lx - 1)*a.P + col;
        st_node4(Wout + 4*n, a0r, a0x, a0y, a0z);
        if (WRITE_U) {
          const double rho = 1.0 + a0r, inv = 1.0/rho;
          const double u0 = a0x*inv + 0.5*p0, u1 = a0y*inv + 0.5*p1, u2 = a0z*inv + 0.5*p2;
          st_node4(U + 4*n, u0, u1, u2, rho);
          // node velocity of my face planes -> the neighbours' ghost planes (the interpolation of their shared cells reads it)
          if (lx - 1 == 1 && pr.UL) st_node4(pr.UL + 4*col, u0, u1, u2, rho);
          if (lx - 1 == a.nxl && pr.UR) st_node4(pr.UR + 4*col, u0, u1, u2, rho);
        }
        st_node4(Fout + 4*n, a.body[0], a.body[1], a.body[2], 0.0);
      }
      // c_x = 0 -> plane lx; c_x = +1 -> plane lx + 1; then the accumulators move one plane on
      const double q0 = sb[0*HN + i];
      const double q2 = sb[2*HN + i + HZ], q11 = sb[11*HN + i - HZ], q3 = sb[3*HN + i + 1], q12 = sb[12*HN + i - 1];
      const double q8 = sb[8*HN + i + HZ + 1], q9 = sb[9*HN + i + HZ - 1], q17 = sb[17*HN + i - HZ - 1], q18 = sb[18*HN + i - HZ + 1];
      const double s0 = (((q0 + q2) + (q3 + q8)) + ((q9 + q11) + (q12 + q17))) + q18;
      a0r = a1r + s0; a0x = a1x; a0y = a1y + (((q11 - q2) + (q17 - q8)) + (q18 - q9)); a0z = a1z + (((q12 - q3) + (q17 - q8)) + (q9 - q18));
      const double q10 = sb[10*HN + i], q13 = sb[13*HN + i - HZ], q14 = sb[14*HN + i + HZ], q15 = sb[15*HN + i - 1], q16 = sb[16*HN + i + 1];
      const double sp = ((q10 + q13) + (q14 + q15)) + q16;
      a1r = sp; a1x = sp; a1y = q13 - q14; a1z = q15 - q16;
      p0 = c0; p1 = c1; p2 = c2;
    }
    __syncthreads();                                     // the populations of this plane are consumed: the buffer is free
  }
}

// Off-equilibrium momentum flux of the post-stream populations (output path only: the "ShearStress"
// and "StrainRate" fluid fields of io/FluidHdf5IO.hh:406-433, 503-541 are multiples of it):
//   PiNeq_ab = sum_q c_qa c_qb fbar_q - cs2 rhoBar delta_ab - j_a j_b / rho,  order xx xy xz yy yz zz
// (Palabos momentTemplates::compute_rhoBar_j_PiNeq, restated from memory).  BounceBack nodes: 0.
// dst is compact SoA [6][Nl]; slot 6 (optional, dst7 != 0) receives rho for the strain-rate prefactor.
__global__ void __launch_bounds__(256)
k_pineq(const double* __restrict__ g, const uint8_t* __restrict__ flags, LatArgs a, int64_t count,
        double* __restrict__ dst) {
  constexpr int CX[19] = {0,-1,0,0,-1,-1,-1,-1,0,0, 1,0,0,1,1,1,1,0,0};
  constexpr int CY[19] = {0,0,-1,0,-1,1,0,0,-1,-1, 0,1,0,1,-1,0,0,1,1};
  constexpr int CZ[19] = {0,0,0,-1,0,0,-1,1,-1,1, 0,0,1,0,0,1,-1,1,-1};
  const int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int64_t n = i + a.P;
  const int rem = (int)(i % a.P);
  const int y = rem / a.nz, z = rem - y*a.nz;
  double f[19];
  pull19(g, a, n, y, z, f);
  double rhoBar, j[3];
  moments19(f, rhoBar, j);
  double pi[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int q = 0; q < 19; q++) {
    pi[0] += CX[q]*CX[q]*f[q]; pi[1] += CX[q]*CY[q]*f[q]; pi[2] += CX[q]*CZ[q]*f[q];
    pi[3] += CY[q]*CY[q]*f[q]; pi[4] += CY[q]*CZ[q]*f[q]; pi[5] += CZ[q]*CZ[q]*f[q];
  }
  const double invRho = 1.0/(1.0 + rhoBar), cs2 = 1.0/3.0;
  pi[0] -= cs2*rhoBar + invRho*j[0]*j[0]; pi[1] -= invRho*j[0]*j[1]; pi[2] -= invRho*j[0]*j[2];
  pi[3] -= cs2*rhoBar + invRho*j[1]*j[1]; pi[4] -= invRho*j[1]*j[2]; pi[5] -= cs2*rhoBar + invRho*j[2]*j[2];
  const bool bb = flags[n] == HCG_BOUNCEBACK;
#pragma unroll
  for (int k = 0; k < 6; k++) dst[(int64_t)k*count + i] = bb ? 0.0 : pi[k];
}

// ---------------------------------------------------------------------------------------------
// Row-pipelined lattice kernels (EXPERIMENTAL, opt-in with HCG_K1_ROWS=1; nz even, a few rows must fit shared memory).
// Measured on B200 (256^3): 1.05 ms vs 0.945 ms for the plain one-thread-per-node kernel, so the plain kernel is the default.
//
// One task = one z-row (fixed lx, y; nz nodes).  Every source row g_q(lx - c_x, y - c_y, :) is nz
// contiguous doubles, so a task is 19 bulk-async copies (cp.async.bulk, the 1-D TMA path) plus the
// node-force row, landing in a ring of shared-memory stages and signalled through one mbarrier per
// stage.  A persistent CTA walks a contiguous range of rows: thread 0 keeps `stages - 1` rows in
// flight, all threads read the populations of "their" node out of shared memory (the z shift of the
// pull becomes a shared-memory index, wrap included), collide in registers and store coalesced.
// The loads of a row are in flight while the previous rows are being collided, independent of the
// register-limited occupancy that bounds the plain one-thread-per-node kernel.
// Guo-forced BGK collision, opposite pairs sharing their symmetric part (same algebra as guo_collide,
// re-associated: differences are a few ulp of the intermediate terms).
__device__ __forceinline__ void guo_collide_pairs(double f[19], const double F[3], double omega) {
  constexpr int CX[19] = {0,-1,0,0,-1,-1,-1,-1,0,0, 1,0,0,1,1,1,1,0,0};
  constexpr int CY[19] = {0,0,-1,0,-1,1,0,0,-1,-1, 0,1,0,1,-1,0,0,1,1};
  constexpr int CZ[19] = {0,0,0,-1,0,0,-1,1,-1,1, 0,0,1,0,0,1,-1,1,-1};
  constexpr double T0 = 1.0/3.0, T1 = 1.0/18.0, T2 = 1.0/36.0;
  double rhoBar, j[3];
  moments19(f, rhoBar, j);
  const double rho = 1.0 + rhoBar, invRho = 1.0/rho;
  const double ux = j[0]*invRho + 0.5*F[0], uy = j[1]*invRho + 0.5*F[1], uz = j[2]*invRho + 0.5*F[2];
  const double om1 = 1.0 - omega, fpre = 1.0 - 0.5*omega;
  const double uSqr = ux*ux + uy*uy + uz*uz;
  const double uF = ux*F[0] + uy*F[1] + uz*F[2];
  // feq = t (rhoBar + 3 rho cu + 4.5 rho cu^2 - 1.5 rho u^2);  Guo = t fpre (3 cF - 3 uF + 9 cu cF)
  const double base = omega*(rhoBar - 1.5*rho*uSqr) - 3.0*fpre*uF;     // symmetric, c-independent
  const double a2 = 4.5*omega*rho, a1 = 3.0*omega*rho, g2 = 9.0*fpre, g1 = 3.0*fpre;
  f[0] = om1*f[0] + T0*base;
#pragma unroll
  for (int q = 1; q <= 9; q++) {
    const double t = (q <= 3) ? T1 : T2;
    const double cu = CX[q]*ux + CY[q]*uy + CZ[q]*uz;
    const double cF = CX[q]*F[0] + CY[q]*F[1] + CZ[q]*F[2];
    const double sym = t*(base + cu*(a2*cu + g2*cF));
    const double asym = t*(a1*cu + g1*cF);
    f[q] = om1*f[q] + (sym + asym);
    f[q+9] = om1*f[q+9] + (sym - asym);
  }
}

// Warp-specialised: warp 0 is the producer (its lanes issue the 19 population rows + the force row of
// every row of a stage), the other warps are consumers.  full[s]: bytes landed; empty[s]: every consumer
// warp has pulled its nodes of stage s into registers (arrives BEFORE colliding, so the refill overlaps
// the arithmetic).  No CTA-wide barrier inside the loop.
template <int NT /*32 producer + consumer threads*/, int MODE /*0 = collide + stream, 1 = moments*/, bool RESET, bool VELBC>
__global__ void __launch_bounds__(NT)
k_rows(const double* __restrict__ gin, double* __restrict__ gout, double* __restrict__ F, double* __restrict__ U,
       const uint8_t* __restrict__ flags, LatArgs a, int row0, int row1, int stages, int R, int interleave, int* done) {
  constexpr int CY[19] = {0,0,-1,0,-1,1,0,0,-1,-1, 0,1,0,1,-1,0,0,1,1};
  constexpr int CZ[19] = {0,0,0,-1,0,0,-1,1,-1,1, 0,0,1,0,0,1,-1,1,-1};
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int nz = a.nz, ny = a.ny;
  const bool fsm = false;                                     // (flag row in the stage measured slower than a direct load: off)
  const int row_doubles = 23*nz + (fsm ? nz/8 : 0);           // 19 population rows + the [nz][4] force row (+ nz flag bytes)
  const size_t stage_doubles = (size_t)R*row_doubles;
  double* sbase = reinterpret_cast<double*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(sbase + (size_t)stages*stage_doubles);
  uint64_t* empty = full + stages;
  const int ncons = (int)blockDim.x - 32, nwarp_cons = ncons >> 5;

  // contiguous range of row groups (R rows each) for this CTA
  const int ngroups = (row1 - row0 + R - 1)/R;
  // interleave: CTA b takes groups b, b + G, b + 2G, ... (all SMs stream through one window of the
  // lattice: DRAM-page / TLB locality); else one contiguous range per CTA
  const int per = (ngroups + (int)gridDim.x - 1)/(int)gridDim.x;
  const int g0 = (interleave & 1) ? (int)blockIdx.x : (int)blockIdx.x*per;
  const int gstep = (interleave & 1) ? (int)gridDim.x : 1;
  const int ng = (interleave & 1) ? (ngroups - g0 + gstep - 1)/gstep : min(g0 + per, ngroups) - g0;
  if (ng <= 0) return;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; s++) { mbar_init(full + s, 1); mbar_init(empty + s, (uint32_t)nwarp_cons); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (threadIdx.x < 32) {
    // ------------------------------------------------------------------ producer warp
    const int lane = threadIdx.x;                             // lane q < 19: population q; lane 19: node force
    int cx = 0, cy = 0;
    if (lane < 19) { cx = d_cx[lane]; cy = d_cy[lane]; }
    for (int it = 0; it < ng; it++) {
      const int s = it % stages;
      if (it >= stages) mbar_wait(empty + s, (uint32_t)(((it / stages) - 1) & 1));
      const int r_first = row0 + (g0 + it*gstep)*R;
      const int nr = min(R, row1 - r_first);
      double* st = sbase + (size_t)s*stage_doubles;
      if (lane == 0) {
        int bytes = 0;
        for (int r = 0; r < nr; r++) {
          const int task = r_first + r, y = task % ny;
          int nvalid = 19;
          if (!a.py) { if (y == 0) nvalid -= 5; if (y == ny - 1) nvalid -= 5; }   // 5 populations per c_y sign
          bytes += nvalid*nz*8 + nz*32 + (fsm ? nz : 0);
        }
        mbar_expect_tx(full + s, (uint32_t)bytes);
      }
      __syncwarp();
      if (lane < 20 || (lane == 20 && fsm)) {
        for (int r = 0; r < nr; r++) {
          const int task = r_first + r;
          const int lx = task / ny + 1, y = task - (lx - 1)*ny;
          double* dst = st + (size_t)r*row_doubles + (size_t)lane*nz;
          if (lane < 19) {
            int sy = y - cy; bool ok = true;
            if (sy < 0) { ok = a.py; sy += ny; } else if (sy >= ny) { ok = a.py; sy -= ny; }
            if (ok) bulk_g2s(dst, gin + (int64_t)lane*a.S + ((int64_t)(lx - cx)*ny + sy)*nz, (uint32_t)(nz*8), full + s);
          } else if (lane == 19) {
            bulk_g2s(dst, F + 4*((int64_t)lx*ny + y)*nz, (uint32_t)(nz*32), full + s);
          } else {
            bulk_g2s(st + (size_t)r*row_doubles + (size_t)23*nz, flags + ((int64_t)lx*ny + y)*nz, (uint32_t)nz, full + s);
          }
        }
      }
    }
    return;
  }

  // -------------------------------------------------------------------- consumer warps
  const int lane = (int)threadIdx.x & 31;
  const int wbase = (int)threadIdx.x - 32 - lane;             // warp-uniform index of the warp's first node
  // `done` (optional): per-plane count of finished (row group, consumer warp) pairs, read by the moments
  // kernel running concurrently two planes behind (lat_collide_moments_overlapped).  A group is published
  // one task late, when its stores have long drained, so the fence costs next to nothing.
  int pending = -1;
  for (int it = 0; it < ng; it++) {
    const int s = it % stages;
    const int r_first = row0 + (g0 + it*gstep)*R;
    const int nr = min(R, row1 - r_first);
    const int nnode = nr*nz;
    const double* st = sbase + (size_t)s*stage_doubles;
    // flags of this thread's first node: issued before the wait so that the latency overlaps it
    uint8_t fl0 = 0;
    if (!fsm && wbase + lane < nnode) {
      const int k = wbase + lane, r = k / nz, z = k - r*nz, task = r_first + r;
      const int lx = task / ny + 1, y = task - (lx - 1)*ny;
      fl0 = flags[((int64_t)lx*ny + y)*nz + z];
    }
    mbar_wait(full + s, (uint32_t)((it / stages) & 1));
    if (wbase >= nnode) {                                     // nothing to do in this stage: still owes its arrival
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(empty + s)) : "memory");
      continue;
    }
    for (int kb = wbase; kb < nnode; kb += ncons) {           // kb is warp-uniform
      const int k = kb + lane;
      const bool act = k < nnode;
      double f[19]; double2 fa = make_double2(0.0, 0.0); double fb = 0.0; int64_t n = 0; uint8_t fls = 0;
      if (act) {
        const int r = k / nz, z = k - r*nz, task = r_first + r;
        const int lx = task / ny + 1, y = task - (lx - 1)*ny;
        n = ((int64_t)lx*ny + y)*nz + z;
        const bool vyp = a.py || (y + 1 < ny), vym = a.py || (y > 0);
        const double* sr = st + (size_t)r*row_doubles;
        int zm = z - 1, zp = z + 1; bool vzm = true, vzp = true;   // source z for c_z = +1 / -1
        if (zm < 0) { zm = nz - 1; vzm = a.pz; }
        if (zp >= nz) { zp = 0; vzp = a.pz; }
#pragma unroll
        for (int q = 0; q < 19; q++) {
          bool ok = true; int zz = z;
          if (CY[q] == 1) ok = vym; else if (CY[q] == -1) ok = vyp;
          if (CZ[q] == 1) { zz = zm; ok = ok && vzm; } else if (CZ[q] == -1) { zz = zp; ok = ok && vzp; }
          f[q] = ok ? sr[q*nz + zz] : 0.0;
        }
        fa = *reinterpret_cast<const double2*>(sr + 19*nz + 4*z);
        fb = sr[19*nz + 4*z + 2];
        if (fsm) fls = reinterpret_cast<const uint8_t*>(sr + 23*nz)[z];
      }
      if (kb + ncons >= nnode) {                              // the warp has pulled its last nodes of stage s
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(empty + s)) : "memory");
      }
      if (MODE == 0 && kb == wbase && pending >= 0) {
        __threadfence();
        __syncwarp();
        if (lane == 0) atomicAdd(done + pending, 1);
        pending = -1;
      }
      if (!act) continue;
      const uint8_t fl = fsm ? fls : ((kb == wbase) ? fl0 : flags[n]);
      if (MODE == 0) {
        if (fl == HCG_BOUNCEBACK) {
#pragma unroll
          for (int q = 1; q <= 9; q++) { const double t = f[q]; f[q] = f[q+9]; f[q+9] = t; }
        } else {
          const double Fn[3] = {fa.x, fa.y, fb};
          if (VELBC && fl >= HCG_VEL_XN) { const double uw[3] = {a.bc[3*(fl-2)], a.bc[3*(fl-2)+1], a.bc[3*(fl-2)+2]}; regularized_complete(f, fl - 2, uw); }
          guo_collide_pairs(f, Fn, a.omega);
        }
#pragma unroll
        if (interleave & 2) { for (int q = 0; q < 19; q++) gout[(int64_t)q*a.S + n] = f[q]; }
        else { for (int q = 0; q < 19; q++) __stcs(gout + (int64_t)q*a.S + n, f[q]); }
      } else {
        double rhoBar, j[3];
        moments19(f, rhoBar, j);
        double u0, u1, u2, rho = 1.0 + rhoBar;
        if (fl == HCG_FLUID) {
          const double invRho = 1.0/rho;
          u0 = j[0]*invRho + 0.5*fa.x; u1 = j[1]*invRho + 0.5*fa.y; u2 = j[2]*invRho + 0.5*fb;
        } else if (fl == HCG_BOUNCEBACK) { u0 = u1 = u2 = 0.0; rho = 1.0; }
        else { u0 = a.bc[3*(fl-2)]; u1 = a.bc[3*(fl-2)+1]; u2 = a.bc[3*(fl-2)+2]; }
        double2* Uw = reinterpret_cast<double2*>(U + 4*n);
        Uw[0] = make_double2(u0, u1); Uw[1] = make_double2(u2, rho);
      }
      if (RESET) { double2* Fw = reinterpret_cast<double2*>(F + 4*n); Fw[0] = make_double2(a.body[0], a.body[1]); Fw[1] = make_double2(a.body[2], 0.0); }
    }
    if (MODE == 0 && done) pending = r_first / ny + 1;        // plane of this group (groups never straddle planes here)
  }
  if (MODE == 0 && pending >= 0) { __threadfence(); __syncwarp(); if (lane == 0) atomicAdd(done + pending, 1); }
}

// Moments pass that runs CONCURRENTLY with the collision kernel of the same step (second stream), a few
// planes behind it: the populations it pulls were written moments ago and are still in the 126 MB L2,
// so the pass costs its U/F traffic instead of a second 152 B/node sweep over HBM.  CTA j handles 256
// consecutive nodes starting at plane `first_plane` (the ring-closing planes come last); thread 0 spins
// on the plane counters published by k_rows until the three source planes are complete.
__global__ void __launch_bounds__(256)
k_moments_wait(const double* __restrict__ g, double* __restrict__ F, double* __restrict__ U,
               const uint8_t* __restrict__ flags, LatArgs a, int64_t first, int64_t count, int64_t rot,
               const int* done, int expected, int wrapx) {
  int64_t k = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  const int64_t kfirst = (int64_t)blockIdx.x*blockDim.x;
  // rotate so that the planes whose sources close the periodic ring are handled last
  auto rotate = [&](int64_t v) { v += rot; if (v >= count) v -= count; return v; };
  if (threadIdx.x == 0) {
    const int64_t klast = min(kfirst + (int64_t)blockDim.x, count) - 1;
    const int p0 = (int)((first + rotate(kfirst)) / a.P) + 1, p1 = (int)((first + rotate(klast)) / a.P) + 1;
    long long t_start = 0;
    for (int pp = 0; pp < 2; pp++) {
      const int p = pp ? p1 : p0;
      if (pp && p1 == p0) break;
      int need[3] = {p - 1, p, p + 1};
      if (need[0] == 0) need[0] = wrapx ? a.nxl : -1;
      if (need[2] == a.nxl + 1) need[2] = wrapx ? 1 : -1;
      for (int m = 0; m < 3; m++) {
        if (need[m] < 0) continue;
        unsigned spin = 0;
        while (ld_acquire_gpu(done + need[m]) < expected) {
          if ((++spin & 1023u) == 0) {
            const long long now = clock64();
            if (t_start == 0) t_start = now; else if (now - t_start > 4000000000LL) __trap();
          }
        }
      }
    }
  }
  __syncthreads();
  if (k >= count) return;
  const int64_t i = first + rotate(k);
  const int64_t n = i + a.P;
  const int rem = (int)(i % a.P);
  const int y = rem / a.nz, z = rem - y*a.nz;
  double f[19];
  // the ring-closing planes read the wrapped plane directly: the ghost planes are filled only after this step
  if (wrapx && (n < 2*a.P || n >= (int64_t)a.nxl*a.P)) {
    constexpr int CX[19] = {0,-1,0,0,-1,-1,-1,-1,0,0, 1,0,0,1,1,1,1,0,0};
    constexpr int CY[19] = {0,0,-1,0,-1,1,0,0,-1,-1, 0,1,0,1,-1,0,0,1,1};
    constexpr int CZ[19] = {0,0,0,-1,0,0,-1,1,-1,1, 0,0,1,0,0,1,-1,1,-1};
    const int lx = (int)(n / a.P);
#pragma unroll
    for (int q = 0; q < 19; q++) {
      int sx = lx - CX[q], sy = y - CY[q], sz = z - CZ[q]; bool ok = true;
      if (sx == 0) sx = a.nxl; else if (sx == a.nxl + 1) sx = 1;
      if (sy < 0) { sy += a.ny; ok = ok && a.py; } else if (sy >= a.ny) { sy -= a.ny; ok = ok && a.py; }
      if (sz < 0) { sz += a.nz; ok = ok && a.pz; } else if (sz >= a.nz) { sz -= a.nz; ok = ok && a.pz; }
      f[q] = ok ? __ldcg(g + (int64_t)q*a.S + ((int64_t)sx*a.ny + sy)*a.nz + sz) : 0.0;
    }
  } else pull19(g, a, n, y, z, f);
  double rhoBar, j[3];
  moments19(f, rhoBar, j);
  const uint8_t fl = flags[n];
  double u0, u1, u2, rho = 1.0 + rhoBar;
  if (fl == HCG_FLUID) {
    const double invRho = 1.0/rho;
    const double2 fa = *reinterpret_cast<const double2*>(F + 4*n);
    u0 = j[0]*invRho + 0.5*fa.x; u1 = j[1]*invRho + 0.5*fa.y; u2 = j[2]*invRho + 0.5*F[4*n + 2];
  } else if (fl == HCG_BOUNCEBACK) { u0 = u1 = u2 = 0.0; rho = 1.0; }
  else { u0 = a.bc[3*(fl-2)]; u1 = a.bc[3*(fl-2)+1]; u2 = a.bc[3*(fl-2)+2]; }
  double2* Uw = reinterpret_cast<double2*>(U + 4*n);
  __stcs(Uw, make_double2(u0, u1)); __stcs(Uw + 1, make_double2(u2, rho));
  double2* Fw = reinterpret_cast<double2*>(F + 4*n);              // force reset rides along
  __stcs(Fw, make_double2(a.body[0], a.body[1])); __stcs(Fw + 1, make_double2(a.body[2], 0.0));
}

// Cell::computeVelocity of a list of nodes (local slab indices) from the current populations and node force:
// out[4*k] = (u_x, u_y, u_z, rho).  Fluid: u = j/rho + F/2; bounce-back: 0; boundary nodes as in the moments pass.
__global__ void __launch_bounds__(256)
k_node_velocity(const double* __restrict__ g, const double* __restrict__ F, const uint8_t* __restrict__ flags, LatArgs a,
                int64_t count, const int64_t* __restrict__ idx, double* __restrict__ out) {
  const int64_t k = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (k >= count) return;
  const int64_t i = idx[k];
  const int64_t n = i + a.P;
  const int rem = (int)(i % a.P);
  const int y = rem / a.nz, z = rem - y*a.nz;
  double f[19];
  pull19(g, a, n, y, z, f);
  double rhoBar, j[3];
  moments19(f, rhoBar, j);
  const uint8_t fl = flags[n];
  double u0, u1, u2, rho = 1.0 + rhoBar;
  if (fl == HCG_FLUID) {
    const double invRho = 1.0/rho;
    u0 = j[0]*invRho + 0.5*F[4*n]; u1 = j[1]*invRho + 0.5*F[4*n + 1]; u2 = j[2]*invRho + 0.5*F[4*n + 2];
  } else if (fl == HCG_BOUNCEBACK) { u0 = u1 = u2 = 0.0; rho = 1.0; }
  else if (a.bcn) boundary_velocity<true>(f, a, n, fl, u0, u1, u2, rho);
  else boundary_velocity<false>(f, a, n, fl, u0, u1, u2, rho);
  out[4*k] = u0; out[4*k + 1] = u1; out[4*k + 2] = u2; out[4*k + 3] = rho;
}
// per-node boundary values: scatter val[k] (4 doubles; slot 3 kept when keep_rho) to node idx[k]; fill with (0, 0, 0, 1)
__global__ void k_bcn_scatter(double* __restrict__ bcn, int64_t P, int64_t count, const int64_t* __restrict__ idx,
                              const double* __restrict__ val, int keep_rho) {
  const int64_t k = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (k >= count) return;
  double* b = bcn + 4*(idx[k] + P);
  b[0] = val[4*k]; b[1] = val[4*k + 1]; b[2] = val[4*k + 2];
  if (!keep_rho) b[3] = val[4*k + 3];
}
__global__ void k_bcn_fill(double* bcn, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (i < total) { double2* b = reinterpret_cast<double2*>(bcn + 4*i); b[0] = make_double2(0.0, 0.0); b[1] = make_double2(0.0, 1.0); }
}
__global__ void k_fill4(double* F, int64_t total, double b0, double b1, double b2) {
  const int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= total) return;
  double2* Fw = reinterpret_cast<double2*>(F + 4*i);
  Fw[0] = make_double2(b0, b1); Fw[1] = make_double2(b2, 0.0);
}

__global__ void k_fill_pop(double* g, int64_t S, int64_t lo, int64_t hi, const double* vals19) {
  const int64_t i = lo + (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= hi) return;
#pragma unroll
  for (int q = 0; q < 19; q++) g[(int64_t)q*S + i] = vals19[q];
}

// periodic self-exchange (n_ranks == 1): left ghost <- last real plane, right ghost <- first
__global__ void k_halo_self(double* buf, int64_t S, int64_t P /* elements per plane */, int nxl, const int* qL, int nL, const int* qR, int nR) {
  const int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= P) return;
  const int k = blockIdx.y;
  if (k < nL) { const int q = qL[k]; buf[(int64_t)q*S + i] = buf[(int64_t)q*S + (int64_t)nxl*P + i]; }
  else { const int q = qR[k - nL]; buf[(int64_t)q*S + (int64_t)(nxl+1)*P + i] = buf[(int64_t)q*S + P + i]; }
}

// reference layout (post-stream, compact slab) <-> device layout (pre-streamed, padded slab)
__global__ void k_to_reference(const double* __restrict__ g, double* __restrict__ dst, LatArgs a) {
  const int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= (int64_t)a.nxl*a.P) return;
  const int rem = (int)(i % a.P);
  const int y = rem / a.nz, z = rem - y*a.nz;
  double f[19];
  pull19(g, a, i + a.P, y, z, f);
  const int64_t Nl = (int64_t)a.nxl*a.P;
#pragma unroll
  for (int q = 0; q < 19; q++) dst[(int64_t)q*Nl + i] = f[q];
}
// g_q(n) = S_q(n + c_q); `s` is a padded buffer holding S with valid ghosts
__global__ void k_from_reference(const double* __restrict__ s, double* __restrict__ g, LatArgs a) {
  constexpr int CX[19] = {0,-1,0,0,-1,-1,-1,-1,0,0, 1,0,0,1,1,1,1,0,0};
  constexpr int CY[19] = {0,0,-1,0,-1,1,0,0,-1,-1, 0,1,0,1,-1,0,0,1,1};
  constexpr int CZ[19] = {0,0,0,-1,0,0,-1,1,-1,1, 0,0,1,0,0,1,-1,1,-1};
  const int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= (int64_t)a.nxl*a.P) return;
  const int64_t n = i + a.P;
  const int rem = (int)(i % a.P);
  const int y = rem / a.nz, z = rem - y*a.nz;
#pragma unroll
  for (int q = 0; q < 19; q++) {
    int yy = y + CY[q], zz = z + CZ[q]; bool ok = true;
    if (yy < 0) { yy = a.ny - 1; ok = ok && a.py; } else if (yy >= a.ny) { yy = 0; ok = ok && a.py; }
    if (zz < 0) { zz = a.nz - 1; ok = ok && a.pz; } else if (zz >= a.nz) { zz = 0; ok = ok && a.pz; }
    const int64_t src = n + (int64_t)CX[q]*a.P + (int64_t)(yy - y)*a.nz + (zz - z);
    g[(int64_t)q*a.S + n] = ok ? s[(int64_t)q*a.S + src] : 0.0;
  }
}
// compact SoA [3][Nl] (C ABI) <-> padded AoS [n][4] (device node vectors)
__global__ void k_pad4(const double* __restrict__ src, double* __restrict__ dst, int64_t Nl, int64_t P) {
  const int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= Nl) return;
  double* d = dst + 4*(P + i);
  d[0] = src[i]; d[1] = src[Nl + i]; d[2] = src[2*Nl + i]; d[3] = 0.0;
}
__global__ void k_unpad4(const double* __restrict__ src, double* __restrict__ dst, int64_t Nl, int64_t P, int first, int ncomp) {
  const int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= Nl) return;
  for (int q = 0; q < ncomp; q++) dst[(int64_t)q*Nl + i] = src[4*(P + i) + first + q];
}
__global__ void k_pad(const double* __restrict__ src, double* __restrict__ dst, int64_t Nl, int64_t S, int64_t P, int ncomp) {
  const int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= Nl) return;
  for (int q = 0; q < ncomp; q++) dst[(int64_t)q*S + P + i] = src[(int64_t)q*Nl + i];
}
__global__ void k_unpad(const double* __restrict__ src, double* __restrict__ dst, int64_t Nl, int64_t S, int64_t P, int ncomp) {
  const int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= Nl) return;
  for (int q = 0; q < ncomp; q++) dst[(int64_t)q*Nl + i] = src[(int64_t)q*S + P + i];
}

// |u| statistics over non-boundary nodes (helper/fluidInfo.cpp:33-65)
__global__ void k_vel_stats(const double* __restrict__ U, const uint8_t* __restrict__ flags, int64_t S, int64_t P,
                            int64_t Nl, double* out /* min,max,sum,count */) {
  __shared__ double smin[256], smax[256], ssum[256], scnt[256];
  double mn = DBL_MAX, mx = 0.0, sm = 0.0, ct = 0.0;
  for (int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x; i < Nl; i += (int64_t)gridDim.x*blockDim.x) {
    const int64_t n = i + P;
    if (flags[n] != HCG_FLUID) continue;
    const double u = sqrt(U[4*n]*U[4*n] + U[4*n+1]*U[4*n+1] + U[4*n+2]*U[4*n+2]);
    mn = fmin(mn, u); mx = fmax(mx, u); sm += u; ct += 1.0;
  }
  const int t = threadIdx.x;
  smin[t] = mn; smax[t] = mx; ssum[t] = sm; scnt[t] = ct;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (t < s) { smin[t] = fmin(smin[t], smin[t+s]); smax[t] = fmax(smax[t], smax[t+s]); ssum[t] += ssum[t+s]; scnt[t] += scnt[t+s]; }
    __syncthreads();
  }
  if (t == 0) {
    // doubles are non-negative: integer atomics on the bit pattern order correctly
    atomicMin((unsigned long long*)&out[0], (unsigned long long)__double_as_longlong(smin[0]));
    atomicMax((unsigned long long*)&out[1], (unsigned long long)__double_as_longlong(smax[0]));
    atomicAdd(&out[2], ssum[0]); atomicAdd(&out[3], scnt[0]);
  }
}

LatArgs make_args(const hcg_ctx* c) {
  LatArgs a;
  a.nxl = c->nxl; a.ny = c->dom.ny; a.nz = c->dom.nz; a.py = c->dom.periodic[1]; a.pz = c->dom.periodic[2];
  a.P = c->P; a.S = c->S; a.omega = c->omega;
  {
    const int R = c->dom.n_ranks, r = c->dom.rank;
    int32_t x0, nl, nr;
    hcg_slab(c->dom.nx, (r + R - 1) % R, R, &x0, &nl); hcg_slab(c->dom.nx, (r + 1) % R, R, &x0, &nr);
    a.nxlL = nl; a.SL = (int64_t)(nl + 2)*c->P; a.SR = (int64_t)(nr + 2)*c->P;
  }
  a.bc = c->d_bc; a.bcn = c->bcn;
  for (int k = 0; k < 3; k++) a.body[k] = c->body[k];
  return a;
}
inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1)/t); }

// population index sets of the halo exchange {10,13,14,15,16, 1,4,5,6,7, 0,1,2}: one device copy per context (contexts of one
// process may live on different GPUs without peer access, e.g. a pre-inlet on a second GPU)
const int h_qsets[13] = {10,13,14,15,16, 1,4,5,6,7, 0,1,2};

hcg_status exchange(hcg_ctx* c, double* buf, int64_t P /* elements per plane */, const int* hL, const int* dL, int nL, const int* hR, const int* dR, int nR) {
  const bool px = c->dom.periodic[0];
  const int R = c->dom.n_ranks, r = c->dom.rank;
  if (R == 1) {
    if (!px) return HCG_OK;
    dim3 grid(nblk(P, 256), nL + nR);
    k_halo_self<<<grid, 256, 0, c->stream>>>(buf, c->S, P, c->nxl, dL, nL, dR, nR);
    KERNEL_CHECK(c);
    return HCG_OK;
  }
  const int left = (r == 0) ? (px ? R - 1 : -1) : r - 1;
  const int right = (r == R - 1) ? (px ? 0 : -1) : r + 1;
  const int64_t S = c->S;
  const size_t B = sizeof(double)*(size_t)P;
  hcg_status s = comm_group_begin(c); if (s) return s;
  // my first real plane -> left neighbour's RIGHT ghost (sets R); my last -> right neighbour's LEFT ghost (sets L)
  if (left >= 0) for (int k = 0; k < nR; k++) comm_send(c, buf + (int64_t)hR[k]*S + P, B, left);
  if (right >= 0) for (int k = 0; k < nL; k++) comm_send(c, buf + (int64_t)hL[k]*S + (int64_t)c->nxl*P, B, right);
  if (right >= 0) for (int k = 0; k < nR; k++) comm_recv(c, buf + (int64_t)hR[k]*S + (int64_t)(c->nxl+1)*P, B, right);
  if (left >= 0) for (int k = 0; k < nL; k++) comm_recv(c, buf + (int64_t)hL[k]*S, B, left);
  return comm_group_end(c, "halo exchange");
}

hcg_status ensure_qsets(hcg_ctx* c) {
  if (!c->d_qsets) {
    CUDA_TRY(c, cudaMalloc(&c->d_qsets, sizeof(h_qsets)));
    CUDA_TRY(c, hcg_h2d(c, c->d_qsets, h_qsets, sizeof(h_qsets)));
  }
  return HCG_OK;
}

}  // namespace

hcg_status lat_halo_exchange_pop(hcg_ctx* c) {
  hcg_status s = ensure_qsets(c); if (s) return s;
  // pull kernel: left ghost read by c_x = +1 populations, right ghost by c_x = -1
  return exchange(c, c->g[c->cur], c->P, h_qsets, c->d_qsets, 5, h_qsets + 5, c->d_qsets + 5, 5);
}
hcg_status lat_halo_exchange_u(hcg_ctx* c) {
  hcg_status s = ensure_qsets(c); if (s) return s;
  // node vectors are AoS [n][4]: one contiguous block of 4*P doubles per plane ("population" 0)
  return exchange(c, c->U, 4*c->P, h_qsets + 10, c->d_qsets + 10, 1, h_qsets + 10, c->d_qsets + 10, 1);
}

// row-pipelined path: eligibility and launch shape
namespace {
struct RowCfg { bool ok; int stages, grid, R, nt, interleave; size_t smem; };
RowCfg row_config(hcg_ctx* c, int nrows) {
  static int env_mode = -2, env_stages = 0, env_ctas = 0, env_nt = 0, env_il = 1;
  if (env_mode == -2) {
    const char* e = getenv("HCG_K1_ROWS"); env_mode = e ? atoi(e) : 0;      // opt-in: measured slower than k_collide_stream
    e = getenv("HCG_K1_STAGES"); env_stages = e ? atoi(e) : 0;
    e = getenv("HCG_K1_CTAS"); env_ctas = e ? atoi(e) : 0;
    e = getenv("HCG_K1_NT"); env_nt = e ? atoi(e) : 0;
    e = getenv("HCG_K1_INTERLEAVE"); env_il = e ? atoi(e) : 1;
  }
  RowCfg r; r.ok = false; r.stages = 0; r.grid = 0; r.smem = 0; r.R = 1; r.nt = 288;
  const int nz = c->dom.nz;
  if (env_mode <= 0 || (nz & 1) || nz < 16 || peer_on(c) || c->has_iobc) return r;   // (the row kernel has no peer-store variant)
  if (c->sm_count <= 0) {
    int dev = c->dom.device, v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev); c->sm_count = v > 0 ? v : 148;
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev); c->smem_optin = v > 0 ? v : 0;
  }
  r.nt = (env_nt == 544) ? 544 : 288;
  const int ncons = r.nt - 32;
  r.R = (ncons + nz - 1)/nz;                                  // rows per stage: one node per consumer thread
  if (r.R < 1) r.R = 1;
  const size_t stage = (size_t)r.R*23*nz*sizeof(double);
  int ctas = env_ctas > 0 ? env_ctas : 2;
  int stages = env_stages > 0 ? env_stages : 2;
  // shrink to what one SM holds (228 KB per SM, 1 KB reserved per CTA)
  while (ctas > 1 && ctas*(stages*stage + 128 + 1024) > (size_t)228*1024) ctas--;
  while (stages > 2 && stages*stage + 128 > (size_t)c->smem_optin) stages--;
  if (stages*stage + 128 > (size_t)c->smem_optin) return r;
  r.ok = true; r.stages = stages; r.smem = stages*stage + 128; r.interleave = env_il;
  const int ngroups = (nrows + r.R - 1)/r.R;
  r.grid = c->sm_count*ctas; if (r.grid > ngroups) r.grid = ngroups;
  return r;
}
template <int MODE, bool RESET, bool VELBC>
hcg_status launch_rows(hcg_ctx* c, const RowCfg& rc, const double* gin, double* gout, const LatArgs& a, int row0, int row1, cudaStream_t st, int* done = nullptr) {
  if (rc.nt == 544) {
    CUDA_TRY(c, cudaFuncSetAttribute(k_rows<544, MODE, RESET, VELBC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rc.smem));
    k_rows<544, MODE, RESET, VELBC><<<rc.grid, 544, rc.smem, st>>>(gin, gout, c->F, c->U, c->flags, a, row0, row1, rc.stages, rc.R, rc.interleave, done);
  } else {
    CUDA_TRY(c, cudaFuncSetAttribute(k_rows<288, MODE, RESET, VELBC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rc.smem));
    k_rows<288, MODE, RESET, VELBC><<<rc.grid, 288, rc.smem, st>>>(gin, gout, c->F, c->U, c->flags, a, row0, row1, rc.stages, rc.R, rc.interleave, done);
  }
  KERNEL_CHECK(c);
  return HCG_OK;
}
}  // namespace

// collide + stream of the rows [row0, row1) of the slab (row = (lx - 1)*ny + y) on stream `st`
hcg_status lat_collide_rows(hcg_ctx* c, bool reset_force, int row0, int row1, cudaStream_t st, int* done) {
  if (row1 <= row0) return HCG_OK;
  LatArgs a = make_args(c);
  double* gin = c->g[c->cur]; double* gout = c->g[1 - c->cur];
  const RowCfg rc = row_config(c, row1 - row0);
  if (rc.ok) {
    if (c->has_velbc) return reset_force ? launch_rows<0, true, true>(c, rc, gin, gout, a, row0, row1, st, done)
                                         : launch_rows<0, false, true>(c, rc, gin, gout, a, row0, row1, st, done);
    return reset_force ? launch_rows<0, true, false>(c, rc, gin, gout, a, row0, row1, st, done)
                       : launch_rows<0, false, false>(c, rc, gin, gout, a, row0, row1, st, done);
  }
  if (done) return hcg_fail(c, HCG_ERR_STATE, "plane counters need the row-pipelined kernel");
  const int64_t first = (int64_t)row0*a.nz, n = (int64_t)(row1 - row0)*a.nz;
  const unsigned nb = nblk(n, 256);
  if (c->W && c->w_valid && c->omega == 1.0) {
    // tau = 1 and the raw moments of the current populations are at hand (k_moments<.., WOUT>)
    double* pL = nullptr; double* pR = nullptr;
    const bool peer = peer_on(c);
    if (peer) { pL = c->peer.link[0].rank >= 0 ? (double*)c->peer.link[0].ptr[1 - c->cur] : nullptr;
                pR = c->peer.link[1].rank >= 0 ? (double*)c->peer.link[1].ptr[1 - c->cur] : nullptr; }
#define K1T_LAUNCH(R, V, P) k_collide_tau1<R, V, P><<<nb, 256, 0, st>>>(gin, gout, c->F, c->W, c->flags, a, first, n, pL, pR)
    if (peer) {
      if (c->has_iobc) { if (reset_force) K1T_LAUNCH(true, 2, true); else K1T_LAUNCH(false, 2, true); }
      else if (c->has_velbc) { if (reset_force) K1T_LAUNCH(true, 1, true); else K1T_LAUNCH(false, 1, true); }
      else { if (reset_force) K1T_LAUNCH(true, 0, true); else K1T_LAUNCH(false, 0, true); }
    } else {
      if (c->has_iobc) { if (reset_force) K1T_LAUNCH(true, 2, false); else K1T_LAUNCH(false, 2, false); }
      else if (c->has_velbc) { if (reset_force) K1T_LAUNCH(true, 1, false); else K1T_LAUNCH(false, 1, false); }
      else { if (reset_force) K1T_LAUNCH(true, 0, false); else K1T_LAUNCH(false, 0, false); }
    }
#undef K1T_LAUNCH
    KERNEL_CHECK(c);
    return HCG_OK;
  }
  if (peer_on(c)) {
    double* pL = c->peer.link[0].rank >= 0 ? (double*)c->peer.link[0].ptr[1 - c->cur] : nullptr;
    double* pR = c->peer.link[1].rank >= 0 ? (double*)c->peer.link[1].ptr[1 - c->cur] : nullptr;
#define K1_LAUNCH(R, V) k_collide_stream<R, V, 2, true><<<nb, 256, 0, st>>>(gin, gout, c->F, c->flags, a, first, n, pL, pR)
    if (c->has_iobc) { if (reset_force) K1_LAUNCH(true, 2); else K1_LAUNCH(false, 2); }
    else if (c->has_velbc) { if (reset_force) K1_LAUNCH(true, 1); else K1_LAUNCH(false, 1); }
    else { if (reset_force) K1_LAUNCH(true, 0); else K1_LAUNCH(false, 0); }
#undef K1_LAUNCH
  } else {
    static int minb = -1;
    if (minb < 0) { const char* e = getenv("HCG_K1_MINB"); minb = e ? atoi(e) : 3; }   // 3 CTAs/SM (80 registers, ~90 B spilled): 0.935 vs 0.950 ms at 2
#define K1_LAUNCH(R, V) do { if (minb == 3) k_collide_stream<R, V, 3, false><<<nb, 256, 0, st>>>(gin, gout, c->F, c->flags, a, first, n, nullptr, nullptr); \
      else k_collide_stream<R, V, 2, false><<<nb, 256, 0, st>>>(gin, gout, c->F, c->flags, a, first, n, nullptr, nullptr); } while (0)
    if (c->has_iobc) { if (reset_force) K1_LAUNCH(true, 2); else K1_LAUNCH(false, 2); }
    else if (c->has_velbc) { if (reset_force) K1_LAUNCH(true, 1); else K1_LAUNCH(false, 1); }
    else { if (reset_force) K1_LAUNCH(true, 0); else K1_LAUNCH(false, 0); }
#undef K1_LAUNCH
  }
  KERNEL_CHECK(c);
  return HCG_OK;
}

hcg_status lat_collide_stream(hcg_ctx* c, bool reset_force) {
  { hcg_status sp = lat_ensure_pops(c); if (sp) return sp; }
  {
    OpTimer tk(c, (c->W && c->w_valid && c->omega == 1.0) ? "kernel:k_collide_tau1" : "kernel:k_collide_stream");
    // a per-node driving force (F0) is restored by a copy after the kernel instead of the in-kernel reset
    hcg_status s = lat_collide_rows(c, reset_force && !c->F0, 0, c->nxl*c->dom.ny, c->stream, nullptr); if (s) return s;
    if (reset_force && c->F0 && (s = lat_reset_force(c))) return s;
  }
  c->cur = 1 - c->cur;
  c->u_valid = false; c->w_valid = false;
  if (peer_on(c)) return peer_barrier(c);          // the face planes were stored into the neighbours by the kernel itself
  return lat_halo_exchange_pop(c);
}

// moments of the rows [row0, row1): plain one-thread-per-node kernel (read-dominated, no register
// pressure: it out-runs the row-pipelined variant)
static bool tau1_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("HCG_TAU1"); on = e ? atoi(e) : 1; }
  return on != 0;
}
static hcg_status moments_rows(hcg_ctx* c, bool reset_force, int row0, int row1) {
  if (row1 <= row0) return HCG_OK;
  LatArgs a = make_args(c);
  const int64_t first = (int64_t)row0*a.nz, n = (int64_t)(row1 - row0)*a.nz;
  const unsigned nb = nblk(n, 256);
  double* pL = nullptr; double* pR = nullptr;
  const bool peer = peer_on(c);
  if (peer) { pL = c->peer.link[0].rank >= 0 ? (double*)c->peer.link[0].ptr[2] : nullptr;
              pR = c->peer.link[1].rank >= 0 ? (double*)c->peer.link[1].ptr[2] : nullptr; }
  const bool wout = c->W != nullptr;
#define MOM_LAUNCH(R, P, WO) do { if (c->has_iobc) k_moments<R, P, WO, true><<<nb, 256, 0, c->stream>>>(c->g[c->cur], c->F, c->U, c->flags, a, first, n, pL, pR, c->W); \
    else k_moments<R, P, WO, false><<<nb, 256, 0, c->stream>>>(c->g[c->cur], c->F, c->U, c->flags, a, first, n, pL, pR, c->W); } while (0)
  if (peer) {
    if (wout) { if (reset_force) MOM_LAUNCH(true, true, true); else MOM_LAUNCH(false, true, true); }
    else { if (reset_force) MOM_LAUNCH(true, true, false); else MOM_LAUNCH(false, true, false); }
  } else {
    if (wout) { if (reset_force) MOM_LAUNCH(true, false, true); else MOM_LAUNCH(false, false, true); }
    else { if (reset_force) MOM_LAUNCH(true, false, false); else MOM_LAUNCH(false, false, false); }
  }
#undef MOM_LAUNCH
  KERNEL_CHECK(c);
  return HCG_OK;
}

// ---- moment-only update (k_moment_step): eligibility, the step, and materialising the populations on demand
static int moment_only_env() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("HCG_MOMENT_ONLY"); on = e ? atoi(e) : 1; }    // default on; HCG_MOMENT_ONLY=0: stored populations
  return on;
}
// eligible: tau = 1, every rank's lattice is plain periodic fluid (mo_ok), uniform driving force, and the raw moments of the
// current state are at hand (the first step after any change of the populations runs collision + moments pass and leaves them)
bool lat_moment_eligible(hcg_ctx* c) {
  const int level = c->mo_mode < 0 ? moment_only_env() : c->mo_mode;
  return level > 0 && tau1_enabled() && c->mo_ok && !c->F0 && c->W && c->w_valid;
}
static int moment_kernel_env() {                          // HCG_MOMENT_KERNEL=simple: one thread per node, 19 neighbour loads (k_moment_step)
  const char* e = getenv("HCG_MOMENT_KERNEL");            // (read per step: lets a test compare the two kernels in one process)
  return (e && strcmp(e, "simple") == 0) ? 0 : 1;
}
static int moment_chunk_env() {                           // planes per CTA of k_moment_tile
  const char* e = getenv("HCG_MOMENT_XC"); int xc = e ? atoi(e) : 32;
  return xc < 1 ? 32 : xc;
}
// W (moments of the current state), W2 / F2 (second buffers of the moment-only update); by identity in Wphys / Fphys
hcg_status lat_moment_buffers(hcg_ctx* c) {
  auto ensure = [&](double** p) -> cudaError_t {
    if (*p) return cudaSuccess;
    cudaError_t e = cudaMalloc(p, sizeof(double)*4*c->S); if (e != cudaSuccess) return e;
    return cudaMemsetAsync(*p, 0, sizeof(double)*4*c->S, c->stream);
  };
  CUDA_TRY(c, ensure(&c->W)); CUDA_TRY(c, ensure(&c->W2)); CUDA_TRY(c, ensure(&c->F2));
  if (!c->Wphys[0]) { c->Wphys[0] = c->W; c->Wphys[1] = c->W2; c->Fphys[0] = c->F; c->Fphys[1] = c->F2; c->mo_flip = 0; }
  return HCG_OK;
}
hcg_status lat_moment_step(hcg_ctx* c, bool write_u) {
  hcg_status s = ensure_qsets(c); if (s) return s;
  if ((s = lat_moment_buffers(c))) return s;
  // ghost planes of the state and of the (spread) force.  Peer transport: the kernel reads the neighbours' face planes where
  // they lie (after a flag barrier: their spreading and their previous update are complete) and stores the node velocity of
  // its own face planes into their ghost planes; otherwise: periodic images / the send-recv exchange into the own ghost planes
  MomentPeer pr = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  const PeerLink* L = c->peer.link;
  const bool peer = peer_on(c) && moment_kernel_env() && (L[0].rank < 0 || (L[0].ptr[6] && L[0].ptr[7] && L[0].ptr[8] && L[0].ptr[9]))
                    && (L[1].rank < 0 || (L[1].ptr[6] && L[1].ptr[7] && L[1].ptr[8] && L[1].ptr[9]));
  if (peer) {
    LatArgs b = make_args(c);
    const int fl = c->mo_flip;                              // the neighbours flip in lockstep: their current buffers carry the same index
    if (L[0].rank >= 0) { pr.WL = (const double*)L[0].ptr[6 + fl] + 4*(int64_t)b.nxlL*c->P; pr.FL = (const double*)L[0].ptr[8 + fl] + 4*(int64_t)b.nxlL*c->P;
                          pr.UL = (double*)L[0].ptr[2] + 4*(int64_t)(b.nxlL + 1)*c->P; }
    if (L[1].rank >= 0) { pr.WR = (const double*)L[1].ptr[6 + fl] + 4*c->P; pr.FR = (const double*)L[1].ptr[8 + fl] + 4*c->P; pr.UR = (double*)L[1].ptr[2]; }
    if ((s = peer_barrier(c))) return s;
  } else if (c->dom.n_ranks == 1 && moment_kernel_env()) {
    // one rank, periodic in x: the "neighbours" are this slab's own end planes - the kernel reads them where they lie and writes
    // the node velocity of its end planes into its own ghost planes (no ghost-fill launches at all)
    pr.WL = c->W + 4*(int64_t)c->nxl*c->P; pr.FL = c->F + 4*(int64_t)c->nxl*c->P; pr.UL = c->U + 4*(int64_t)(c->nxl + 1)*c->P;
    pr.WR = c->W + 4*c->P; pr.FR = c->F + 4*c->P; pr.UR = c->U;
  } else {
    if ((s = exchange(c, c->W, 4*c->P, h_qsets + 10, c->d_qsets + 10, 1, h_qsets + 10, c->d_qsets + 10, 1))) return s;
    if ((s = exchange(c, c->F, 4*c->P, h_qsets + 10, c->d_qsets + 10, 1, h_qsets + 10, c->d_qsets + 10, 1))) return s;
  }
  const bool self_wrap = !peer && pr.WL != nullptr;
  LatArgs a = make_args(c);
  if (moment_kernel_env()) {
    OpTimer tk(c, "kernel:k_moment_tile");
    const int xc = std::min(moment_chunk_env(), c->nxl);
    hcg_status st = HCG_OK;
    auto launch = [&](auto ty, auto tz, auto minb) {
      constexpr int TY = decltype(ty)::value, TZ = decltype(tz)::value, MINB = decltype(minb)::value;
      constexpr int NT = ((TY + 2)*(TZ + 2) + 31)/32*32;
      const size_t smem = sizeof(double)*(19*(TY + 2)*(TZ + 2) + 2*2*(TY + 2)*(TZ + 2)*4) + 16;   // populations + input ring + 2 mbarriers
      dim3 grid((unsigned)((a.nz + TZ - 1)/TZ), (unsigned)((a.ny + TY - 1)/TY), (unsigned)((c->nxl + xc - 1)/xc));
      cudaError_t e;
      if (write_u) {
        e = cudaFuncSetAttribute(k_moment_tile<true, TY, TZ, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) k_moment_tile<true, TY, TZ, MINB><<<grid, NT, smem, c->stream>>>(c->W, c->F, c->W2, c->F2, c->U, a, xc, pr);
      } else {
        e = cudaFuncSetAttribute(k_moment_tile<false, TY, TZ, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) k_moment_tile<false, TY, TZ, MINB><<<grid, NT, smem, c->stream>>>(c->W, c->F, c->W2, c->F2, c->U, a, xc, pr);
      }
      if (e != cudaSuccess) st = hcg_fail(c, HCG_ERR_CUDA, std::string("k_moment_tile: ") + cudaGetErrorString(e));
    };
    const char* shape = getenv("HCG_MOMENT_TILE");        // experiment knob: tile shape / CTAs per SM
    if (shape && !strcmp(shape, "4x32")) launch(std::integral_constant<int, 4>(), std::integral_constant<int, 32>(), std::integral_constant<int, 3>());
    else if (shape && !strcmp(shape, "6x32")) launch(std::integral_constant<int, 6>(), std::integral_constant<int, 32>(), std::integral_constant<int, 2>());
    else if (shape && !strcmp(shape, "4x64")) launch(std::integral_constant<int, 4>(), std::integral_constant<int, 64>(), std::integral_constant<int, 1>());
    else if (shape && !strcmp(shape, "16x32")) launch(std::integral_constant<int, 16>(), std::integral_constant<int, 32>(), std::integral_constant<int, 1>());
    else launch(std::integral_constant<int, 8>(), std::integral_constant<int, 32>(), std::integral_constant<int, 2>());
    if (st) return st;
    KERNEL_CHECK(c);
  } else {
    OpTimer tk(c, "kernel:k_moment_step");
    const unsigned nb = nblk(c->Nl, 256);
    if (write_u) k_moment_step<true><<<nb, 256, 0, c->stream>>>(c->W, c->F, c->W2, c->F2, c->U, a, c->Nl);
    else k_moment_step<false><<<nb, 256, 0, c->stream>>>(c->W, c->F, c->W2, c->F2, c->U, a, c->Nl);
    KERNEL_CHECK(c);
  }
  std::swap(c->W, c->W2); std::swap(c->F, c->F2);           // the second buffers now hold the inputs of this step (kept for lat_ensure_pops)
  c->mo_flip ^= 1;
  c->pops_stale = true; c->u_valid = write_u; c->f_clean = true; c->w_valid = true;
  if (write_u && !self_wrap) return peer ? peer_barrier(c) : lat_halo_exchange_u(c);
  return HCG_OK;
}
// the populations of the current state from the inputs of the last moment-only step: g_q(n) = f*_q(n) (k_collide_tau1)
hcg_status lat_ensure_pops(hcg_ctx* c) {
  if (!c->pops_stale) return HCG_OK;
  LatArgs a = make_args(c);
  k_collide_tau1<false, 0, false><<<nblk(c->Nl, 256), 256, 0, c->stream>>>(c->g[1 - c->cur], c->g[c->cur], c->F2, c->W2, c->flags, a, 0, c->Nl, nullptr, nullptr);
  KERNEL_CHECK(c);
  c->pops_stale = false;
  return lat_halo_exchange_pop(c);
}

hcg_status lat_pineq(hcg_ctx* c, double* dst_dev) {
  { hcg_status sp = lat_ensure_pops(c); if (sp) return sp; }
  LatArgs a = make_args(c);
  k_pineq<<<nblk(c->Nl, 256), 256, 0, c->stream>>>(c->g[c->cur], c->flags, a, c->Nl, dst_dev);
  KERNEL_CHECK(c);
  return HCG_OK;
}

hcg_status lat_moments(hcg_ctx* c, bool reset_force, bool want_rho) {
  (void)want_rho;   // the density always rides in slot 3 of the node velocity
  { hcg_status sp = lat_ensure_pops(c); if (sp) return sp; }
  {
    OpTimer tk(c, "kernel:k_moments");
    if (!c->W && c->omega == 1.0 && tau1_enabled()) {      // tau = 1: keep the raw moments for the next collision
      CUDA_TRY(c, cudaMalloc(&c->W, sizeof(double)*4*c->S));
      CUDA_TRY(c, cudaMemsetAsync(c->W, 0, sizeof(double)*4*c->S, c->stream));
    }
    hcg_status s = moments_rows(c, reset_force && !c->F0, 0, c->nxl*c->dom.ny); if (s) return s;
    if (reset_force && c->F0 && (s = lat_reset_force(c))) return s;
  }
  c->u_valid = true; c->w_valid = c->W != nullptr;
  if (peer_on(c)) return peer_barrier(c);
  return lat_halo_exchange_u(c);
}

// collideAndStream + the interpolation-step moments pass (with the force reset), overlapped: the collision
// kernel (row-pipelined, publishes per-plane completion counters) on the main stream, the moments kernel on
// the second stream trailing it through L2.  Returns *done_out = false when the shape is not eligible
// (the caller then runs the two passes back to back).
hcg_status lat_collide_moments_overlapped(hcg_ctx* c, bool* done_out) {
  *done_out = false;
  { hcg_status sp = lat_ensure_pops(c); if (sp) return sp; }
  static int env_on = -1;
  if (env_on < 0) { const char* e = getenv("HCG_OVERLAP"); env_on = e ? atoi(e) : 0; }   // opt-in (needs HCG_K1_ROWS=1): measured slower, DESIGN.md §4
  const int ny = c->dom.ny, nxl = c->nxl;
  RowCfg rc = row_config(c, nxl*ny);
  if (!env_on || !rc.ok || !rc.interleave || ny % rc.R || nxl < 4 || c->F0) return HCG_OK;
  const int R = c->dom.n_ranks;
  const int wrapx = (R == 1 && c->dom.periodic[0]) ? 1 : 0;
  if (!c->fused_done) CUDA_TRY(c, cudaMalloc(&c->fused_done, sizeof(int)*(nxl + 2)));
  if (!c->ev_fork) { CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming)); CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming)); }
  CUDA_TRY(c, cudaMemsetAsync(c->fused_done, 0, sizeof(int)*(nxl + 2), c->stream));
  CUDA_TRY(c, cudaEventRecord(c->ev_fork, c->stream));
  CUDA_TRY(c, cudaStreamWaitEvent(c->stream_halo, c->ev_fork, 0));
  hcg_status s;
  {
    OpTimer tk(c, "kernel:k_collide_stream");
    if ((s = lat_collide_rows(c, false, 0, nxl*ny, c->stream, c->fused_done))) return s;
  }
  c->cur = 1 - c->cur; c->w_valid = false;
  {
    // planes m_lo..m_hi; multi-GPU: the face planes need the neighbour's halo and follow after the exchange
    const int m_lo = R > 1 ? 2 : 1, m_hi = R > 1 ? nxl - 1 : nxl;
    LatArgs a = make_args(c);
    const int64_t first = (int64_t)(m_lo - 1)*c->P, count = (int64_t)(m_hi - m_lo + 1)*c->P;
    const int64_t rot = (wrapx && count > c->P) ? c->P : 0;   // start at plane 2: plane 1 (needs plane nxl) comes last
    const int expected = (ny/rc.R)*((rc.nt - 32)/32);
    k_moments_wait<<<nblk(count, 256), 256, 0, c->stream_halo>>>(c->g[c->cur], c->F, c->U, c->flags, a, first, count, rot,
                                                                 c->fused_done, expected, wrapx);
    KERNEL_CHECK(c);
  }
  CUDA_TRY(c, cudaEventRecord(c->ev_join, c->stream_halo));
  CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_join, 0));
  if ((s = lat_halo_exchange_pop(c))) return s;
  if (R > 1) {                                                // face planes: their ghost neighbours have just arrived
    if ((s = moments_rows(c, true, 0, ny))) return s;
    if ((s = moments_rows(c, true, (nxl - 1)*ny, nxl*ny))) return s;
  }
  c->u_valid = true;
  *done_out = true;
  return lat_halo_exchange_u(c);
}

hcg_status lat_reset_force(hcg_ctx* c) {
  c->f_clean = true;
  if (c->F0) { CUDA_TRY(c, cudaMemcpyAsync(c->F, c->F0, sizeof(double)*4*c->S, cudaMemcpyDeviceToDevice, c->stream)); return HCG_OK; }
  k_fill4<<<nblk(c->S, 256), 256, 0, c->stream>>>(c->F, c->S, c->body[0], c->body[1], c->body[2]);
  KERNEL_CHECK(c);
  return HCG_OK;
}

hcg_status lat_init_equilibrium(hcg_ctx* c, double rho, const double u[3]) {
  static const int C[19][3] = {{0,0,0},{-1,0,0},{0,-1,0},{0,0,-1},{-1,-1,0},{-1,1,0},{-1,0,-1},{-1,0,1},{0,-1,-1},{0,-1,1},
                               {1,0,0},{0,1,0},{0,0,1},{1,1,0},{1,-1,0},{1,0,1},{1,0,-1},{0,1,1},{0,1,-1}};
  double vals[19];
  const double rhoBar = rho - 1.0, invRho = 1.0/rho;
  const double j[3] = {rho*u[0], rho*u[1], rho*u[2]};
  const double jSqr = j[0]*j[0] + j[1]*j[1] + j[2]*j[2];
  for (int q = 0; q < 19; q++) {
    const double t = (q == 0) ? 1.0/3.0 : ((q <= 3 || (q >= 10 && q <= 12)) ? 1.0/18.0 : 1.0/36.0);
    const double cj = C[q][0]*j[0] + C[q][1]*j[1] + C[q][2]*j[2];
    vals[q] = t*(rhoBar + 3.0*cj + invRho*(4.5*cj*cj - 1.5*jSqr));
  }
  double* dv;
  CUDA_TRY(c, cudaMalloc(&dv, sizeof(vals)));
  CUDA_TRY(c, cudaMemcpyAsync(dv, vals, sizeof(vals), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(c, cudaMemsetAsync(c->g[c->cur], 0, sizeof(double)*19*c->S, c->stream));
  // a uniform field is its own pre-streamed image; real planes only, ghosts by exchange
  k_fill_pop<<<nblk((int64_t)c->nxl*c->P, 256), 256, 0, c->stream>>>(c->g[c->cur], c->S, c->P, (int64_t)(c->nxl+1)*c->P, dv);
  KERNEL_CHECK(c);
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  cudaFree(dv);
  c->u_valid = false; c->w_valid = false; c->pops_stale = false;
  return lat_halo_exchange_pop(c);
}

hcg_status lat_pop_to_reference(hcg_ctx* c, double* dst_dev) {
  { hcg_status sp = lat_ensure_pops(c); if (sp) return sp; }
  LatArgs a = make_args(c);
  k_to_reference<<<nblk((int64_t)c->nxl*c->P, 256), 256, 0, c->stream>>>(c->g[c->cur], dst_dev, a);
  KERNEL_CHECK(c);
  return HCG_OK;
}

hcg_status lat_pop_from_reference(hcg_ctx* c, const double* src_dev) {
  c->pops_stale = false;                                  // the uploaded populations replace whatever state there was
  LatArgs a = make_args(c);
  double* s = c->g[1 - c->cur];
  CUDA_TRY(c, cudaMemsetAsync(s, 0, sizeof(double)*19*c->S, c->stream));
  k_pad<<<nblk(c->Nl, 256), 256, 0, c->stream>>>(src_dev, s, c->Nl, c->S, c->P, 19);
  KERNEL_CHECK(c);
  hcg_status st = ensure_qsets(c); if (st) return st;
  // g_q(n) = S_q(n + c_q): c_x = -1 populations read the LEFT ghost, c_x = +1 the RIGHT ghost
  st = exchange(c, s, c->P, h_qsets + 5, c->d_qsets + 5, 5, h_qsets, c->d_qsets, 5); if (st) return st;
  CUDA_TRY(c, cudaMemsetAsync(c->g[c->cur], 0, sizeof(double)*19*c->S, c->stream));
  k_from_reference<<<nblk((int64_t)c->nxl*c->P, 256), 256, 0, c->stream>>>(s, c->g[c->cur], a);
  KERNEL_CHECK(c);
  c->u_valid = false; c->w_valid = false;
  return lat_halo_exchange_pop(c);
}

hcg_status lat_velocity_stats(hcg_ctx* c, double* vmin, double* vmax, double* vmean) {
  double* d; double h[4] = {DBL_MAX, 0.0, 0.0, 0.0};
  CUDA_TRY(c, cudaMalloc(&d, sizeof(h)));
  CUDA_TRY(c, cudaMemcpyAsync(d, h, sizeof(h), cudaMemcpyHostToDevice, c->stream));
  k_vel_stats<<<296, 256, 0, c->stream>>>(c->U, c->flags, c->S, c->P, c->Nl, d);
  KERNEL_CHECK(c);
  CUDA_TRY(c, cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  cudaFree(d);
  *vmin = h[0]; *vmax = h[1]; *vmean = h[3] > 0 ? h[2]/h[3] : 0.0;   // local slab; caller reduces over ranks
  return HCG_OK;
}

// utilities used by capi.cu
hcg_status lat_pad3(hcg_ctx* c, const double* src_dev, double* dst) {
  k_pad4<<<nblk(c->Nl, 256), 256, 0, c->stream>>>(src_dev, dst, c->Nl, c->P);
  KERNEL_CHECK(c); return HCG_OK;
}
hcg_status lat_unpad(hcg_ctx* c, const double* src, double* dst_dev, int first, int ncomp) {
  k_unpad4<<<nblk(c->Nl, 256), 256, 0, c->stream>>>(src, dst_dev, c->Nl, c->P, first, ncomp);
  KERNEL_CHECK(c); return HCG_OK;
}

// per-node boundary values of the Zou-He nodes (allocated on first use, (0, 0, 0, 1) everywhere)
hcg_status lat_bcn_ensure(hcg_ctx* c) {
  if (c->bcn) return HCG_OK;
  CUDA_TRY(c, cudaMalloc(&c->bcn, sizeof(double)*4*c->S));
  k_bcn_fill<<<nblk(c->S, 256), 256, 0, c->stream>>>(c->bcn, c->S);
  KERNEL_CHECK(c);
  return HCG_OK;
}
hcg_status lat_bcn_scatter(hcg_ctx* c, int64_t n, const int64_t* idx_dev, const double* val_dev, bool keep_rho, cudaStream_t st) {
  if (n <= 0) return HCG_OK;
  k_bcn_scatter<<<nblk(n, 256), 256, 0, st>>>(c->bcn, c->P, n, idx_dev, val_dev, keep_rho ? 1 : 0);
  KERNEL_CHECK(c);
  return HCG_OK;
}
hcg_status lat_node_velocity(hcg_ctx* c, int64_t n, const int64_t* idx_dev, double* out_dev, cudaStream_t st) {
  if (n <= 0) return HCG_OK;
  { hcg_status sp = lat_ensure_pops(c); if (sp) return sp; }
  LatArgs a = make_args(c);
  k_node_velocity<<<nblk(n, 256), 256, 0, st>>>(c->g[c->cur], c->F, c->flags, a, n, idx_dev, out_dev);
  KERNEL_CHECK(c);
  return HCG_OK;
}
