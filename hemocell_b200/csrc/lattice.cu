// K1: fused D3Q19 Guo-BGK stream(pull)+collide, moments pass, halo exchange, layout conversion.
// Replaces Palabos MultiBlockLattice3D<double,ForcedD3Q19Descriptor>::collideAndStream() with
// GuoExternalForceBGKdynamics / BounceBack / regularized velocity planes as HemoCell drives it
// (reference core/hemoCell.cpp:317; arithmetic restated in SURVEY.md Appendix C).
#include "ctx.cuh"
#include <nccl.h>
#include <cfloat>
#include <cstdlib>

namespace {

struct LatArgs {
  int nxl, ny, nz, py, pz;
  int64_t P, S;
  double omega;
  const double* bc;     // [6][3] wall velocity per orientation, device memory
  double body[3];
};

__device__ __forceinline__ double feq(double t, double cj, double rhoBar, double invRho, double jSqr) {
  return t * (rhoBar + 3.0*cj + invRho*(4.5*cj*cj - 1.5*jSqr));
}

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

__device__ __forceinline__ void moments19(const double f[19], double& rhoBar, double j[3]) {
  constexpr int CX[19] = {0,-1,0,0,-1,-1,-1,-1,0,0, 1,0,0,1,1,1,1,0,0};
  constexpr int CY[19] = {0,0,-1,0,-1,1,0,0,-1,-1, 0,1,0,1,-1,0,0,1,1};
  constexpr int CZ[19] = {0,0,0,-1,0,0,-1,1,-1,1, 0,0,1,0,0,1,-1,1,-1};
  rhoBar = 0.0; j[0] = j[1] = j[2] = 0.0;
#pragma unroll
  for (int q = 0; q < 19; q++) {
    rhoBar += f[q];
    if (CX[q] == 1) j[0] += f[q]; else if (CX[q] == -1) j[0] -= f[q];
    if (CY[q] == 1) j[1] += f[q]; else if (CY[q] == -1) j[1] -= f[q];
    if (CZ[q] == 1) j[2] += f[q]; else if (CZ[q] == -1) j[2] -= f[q];
  }
}

__device__ __forceinline__ void guo_collide(double f[19], const double F[3], double omega) {
  constexpr int CX[19] = {0,-1,0,0,-1,-1,-1,-1,0,0, 1,0,0,1,1,1,1,0,0};
  constexpr int CY[19] = {0,0,-1,0,-1,1,0,0,-1,-1, 0,1,0,1,-1,0,0,1,1};
  constexpr int CZ[19] = {0,0,0,-1,0,0,-1,1,-1,1, 0,0,1,0,0,1,-1,1,-1};
  constexpr double T0 = 1.0/3.0, T1 = 1.0/18.0, T2 = 1.0/36.0;
  double rhoBar, j[3];
  moments19(f, rhoBar, j);
  const double rho = 1.0 + rhoBar, invRho = 1.0/rho;
  const double ux = j[0]*invRho + 0.5*F[0], uy = j[1]*invRho + 0.5*F[1], uz = j[2]*invRho + 0.5*F[2];
  const double jx = rho*ux, jy = rho*uy, jz = rho*uz;
  const double jSqr = jx*jx + jy*jy + jz*jz;
  const double om1 = 1.0 - omega, fpre = 1.0 - omega/2.0;
  const double uF = ux*F[0] + uy*F[1] + uz*F[2];
#pragma unroll
  for (int q = 0; q < 19; q++) {
    const double t = (q == 0) ? T0 : ((q <= 3 || (q >= 10 && q <= 12)) ? T1 : T2);
    const double cj = CX[q]*jx + CY[q]*jy + CZ[q]*jz;
    const double cu = CX[q]*ux + CY[q]*uy + CZ[q]*uz;
    const double cF = CX[q]*F[0] + CY[q]*F[1] + CZ[q]*F[2];
    // sum_d ((c_d - u_d)*3 + cu*c_d*9) F_d
    const double ft = 3.0*(cF - uF) + 9.0*cu*cF;
    f[q] = om1*f[q] + omega*feq(t, cj, rhoBar, invRho, jSqr) + t*fpre*ft;
  }
}

// regularized velocity plane (see oracle/hemo_oracle.c:regularized_velocity_complete)
__device__ __forceinline__ void regularized_complete(double f[19], int o, const double uw[3]) {
  constexpr int CC[19][3] = {{0,0,0},{-1,0,0},{0,-1,0},{0,0,-1},{-1,-1,0},{-1,1,0},{-1,0,-1},{-1,0,1},{0,-1,-1},{0,-1,1},
                             {1,0,0},{0,1,0},{0,0,1},{1,1,0},{1,-1,0},{1,0,1},{1,0,-1},{0,1,1},{0,1,-1}};
  constexpr double TW[19] = {1.0/3.0, 1.0/18.0,1.0/18.0,1.0/18.0,1.0/36.0,1.0/36.0,1.0/36.0,1.0/36.0,1.0/36.0,1.0/36.0,
                             1.0/18.0,1.0/18.0,1.0/18.0,1.0/36.0,1.0/36.0,1.0/36.0,1.0/36.0,1.0/36.0,1.0/36.0};
  const int dir = o >> 1, sgn = (o & 1) ? 1 : -1;
  double rho_on = 0.0, rho_out = 0.0;
#pragma unroll
  for (int q = 0; q < 19; q++) {
    const int cd = dir == 0 ? CC[q][0] : (dir == 1 ? CC[q][1] : CC[q][2]);
    const int cn = cd*sgn;
    if (cn == 0) rho_on += f[q] + TW[q]; else if (cn > 0) rho_out += f[q] + TW[q];
  }
  const double rho = (rho_on + 2.0*rho_out) / (1.0 + sgn*uw[dir]);
  const double rhoBar = rho - 1.0, invRho = 1.0/rho;
  const double jx = rho*uw[0], jy = rho*uw[1], jz = rho*uw[2];
  const double jSqr = jx*jx + jy*jy + jz*jz;
  double eq[19], fneq[19];
#pragma unroll
  for (int q = 0; q < 19; q++) {
    eq[q] = feq(TW[q], CC[q][0]*jx + CC[q][1]*jy + CC[q][2]*jz, rhoBar, invRho, jSqr);
    fneq[q] = f[q] - eq[q];
  }
#pragma unroll
  for (int q = 1; q < 19; q++) {
    const int cd = dir == 0 ? CC[q][0] : (dir == 1 ? CC[q][1] : CC[q][2]);
    if (cd*sgn < 0) fneq[q] = f[q <= 9 ? q + 9 : q - 9] - eq[q <= 9 ? q + 9 : q - 9];
  }
  double Pxx = 0, Pxy = 0, Pxz = 0, Pyy = 0, Pyz = 0, Pzz = 0;
#pragma unroll
  for (int q = 0; q < 19; q++) {
    Pxx += CC[q][0]*CC[q][0]*fneq[q]; Pxy += CC[q][0]*CC[q][1]*fneq[q]; Pxz += CC[q][0]*CC[q][2]*fneq[q];
    Pyy += CC[q][1]*CC[q][1]*fneq[q]; Pyz += CC[q][1]*CC[q][2]*fneq[q]; Pzz += CC[q][2]*CC[q][2]*fneq[q];
  }
  constexpr double cs2 = 1.0/3.0;
#pragma unroll
  for (int q = 0; q < 19; q++) {
    const double Q = (CC[q][0]*CC[q][0] - cs2)*Pxx + 2.0*CC[q][0]*CC[q][1]*Pxy + 2.0*CC[q][0]*CC[q][2]*Pxz
                   + (CC[q][1]*CC[q][1] - cs2)*Pyy + 2.0*CC[q][1]*CC[q][2]*Pyz + (CC[q][2]*CC[q][2] - cs2)*Pzz;
    f[q] = eq[q] + TW[q]*4.5*Q;
  }
}

// One thread per real node.  RESET: write the body force back after reading F (steps without
// velocity interpolation).
template <bool RESET, bool VELBC, int MINB>
__global__ void __launch_bounds__(256, MINB)
k_collide_stream(const double* __restrict__ gin, double* __restrict__ gout, double* __restrict__ F,
                 const uint8_t* __restrict__ flags, LatArgs a) {
  const int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= (int64_t)a.nxl*a.P) return;
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
    if (VELBC && fl >= HCG_VEL_XN) { const double uw[3] = {a.bc[3*(fl-2)], a.bc[3*(fl-2)+1], a.bc[3*(fl-2)+2]}; regularized_complete(f, fl - 2, uw); }
    guo_collide(f, Fn, a.omega);
  }
  if (RESET) { double2* Fw = reinterpret_cast<double2*>(F + 4*n); Fw[0] = make_double2(a.body[0], a.body[1]); Fw[1] = make_double2(a.body[2], 0.0); }
#pragma unroll
  for (int q = 0; q < 19; q++) gout[(int64_t)q*a.S + n] = f[q];
}

// Moments pass: velocity the IBM interpolation sees, u = j/rho + F/2 of the POST-stream
// populations with the spread force still on the node (Cell::computeVelocity through
// core/hemoCellParticleField.cpp:833).  BounceBack: 0; velocity plane: wall velocity.
template <bool RESET>
__global__ void __launch_bounds__(256)
k_moments(const double* __restrict__ g, double* __restrict__ F, double* __restrict__ U,
          const uint8_t* __restrict__ flags, LatArgs a) {
  const int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= (int64_t)a.nxl*a.P) return;
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
    const double2 fa = *reinterpret_cast<const double2*>(F + 4*n);
    u0 = j[0]*invRho + 0.5*fa.x; u1 = j[1]*invRho + 0.5*fa.y; u2 = j[2]*invRho + 0.5*F[4*n + 2];
  } else if (fl == HCG_BOUNCEBACK) { u0 = u1 = u2 = 0.0; rho = 1.0; }
  else { u0 = a.bc[3*(fl-2)]; u1 = a.bc[3*(fl-2)+1]; u2 = a.bc[3*(fl-2)+2]; }
  double2* Uw = reinterpret_cast<double2*>(U + 4*n);
  Uw[0] = make_double2(u0, u1); Uw[1] = make_double2(u2, rho);      // slot 3 carries the density
  if (RESET) { double2* Fw = reinterpret_cast<double2*>(F + 4*n); Fw[0] = make_double2(a.body[0], a.body[1]); Fw[1] = make_double2(a.body[2], 0.0); }
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
  a.bc = c->d_bc;
  for (int k = 0; k < 3; k++) a.body[k] = c->body[k];
  return a;
}
inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1)/t); }

int* d_qsets = nullptr;   // {10,13,14,15,16, 1,4,5,6,7, 0,1,2}
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
  if (!c->nccl) return hcg_fail(c, HCG_ERR_STATE, "n_ranks > 1 but hcg_comm_init was not called");
  ncclComm_t comm = (ncclComm_t)c->nccl;
  const int left = (r == 0) ? (px ? R - 1 : -1) : r - 1;
  const int right = (r == R - 1) ? (px ? 0 : -1) : r + 1;
  const int64_t S = c->S;
  ncclGroupStart();
  // my first real plane -> left neighbour's RIGHT ghost (sets R); my last -> right neighbour's LEFT ghost (sets L)
  if (left >= 0) for (int k = 0; k < nR; k++) ncclSend(buf + (int64_t)hR[k]*S + P, P, ncclDouble, left, comm, c->stream);
  if (right >= 0) for (int k = 0; k < nL; k++) ncclSend(buf + (int64_t)hL[k]*S + (int64_t)c->nxl*P, P, ncclDouble, right, comm, c->stream);
  if (right >= 0) for (int k = 0; k < nR; k++) ncclRecv(buf + (int64_t)hR[k]*S + (int64_t)(c->nxl+1)*P, P, ncclDouble, right, comm, c->stream);
  if (left >= 0) for (int k = 0; k < nL; k++) ncclRecv(buf + (int64_t)hL[k]*S, P, ncclDouble, left, comm, c->stream);
  ncclResult_t rc = ncclGroupEnd();
  if (rc != ncclSuccess) return hcg_fail(c, HCG_ERR_NCCL, std::string("halo exchange: ") + ncclGetErrorString(rc));
  return HCG_OK;
}

hcg_status ensure_qsets(hcg_ctx* c) {
  if (!d_qsets) {
    CUDA_TRY(c, cudaMalloc(&d_qsets, sizeof(h_qsets)));
    CUDA_TRY(c, cudaMemcpy(d_qsets, h_qsets, sizeof(h_qsets), cudaMemcpyHostToDevice));
  }
  return HCG_OK;
}

}  // namespace

hcg_status lat_halo_exchange_pop(hcg_ctx* c) {
  hcg_status s = ensure_qsets(c); if (s) return s;
  // pull kernel: left ghost read by c_x = +1 populations, right ghost by c_x = -1
  return exchange(c, c->g[c->cur], c->P, h_qsets, d_qsets, 5, h_qsets + 5, d_qsets + 5, 5);
}
hcg_status lat_halo_exchange_u(hcg_ctx* c) {
  hcg_status s = ensure_qsets(c); if (s) return s;
  // node vectors are AoS [n][4]: one contiguous block of 4*P doubles per plane ("population" 0)
  return exchange(c, c->U, 4*c->P, h_qsets + 10, d_qsets + 10, 1, h_qsets + 10, d_qsets + 10, 1);
}

hcg_status lat_collide_stream(hcg_ctx* c, bool reset_force) {
  LatArgs a = make_args(c);
  const int64_t n = (int64_t)c->nxl*c->P;
  double* gin = c->g[c->cur]; double* gout = c->g[1 - c->cur];
  const unsigned nb = nblk(n, 256);
  {
  OpTimer tk(c, "kernel:k_collide_stream");
  static int variant = -1;
  if (variant < 0) { const char* e = getenv("HCG_K1_MINB"); variant = e ? atoi(e) : 2; }
#define K1_LAUNCH(R, V, M) k_collide_stream<R, V, M><<<nb, 256, 0, c->stream>>>(gin, gout, c->F, c->flags, a)
#define K1_PICK(M) do { if (c->has_velbc) { if (reset_force) K1_LAUNCH(true, true, M); else K1_LAUNCH(false, true, M); } \
                        else { if (reset_force) K1_LAUNCH(true, false, M); else K1_LAUNCH(false, false, M); } } while (0)
  if (variant == 4) K1_PICK(4); else if (variant == 3) K1_PICK(3); else if (variant == 2) K1_PICK(2); else K1_PICK(1);
  }
  KERNEL_CHECK(c);
  c->cur = 1 - c->cur;
  c->u_valid = false;
  return lat_halo_exchange_pop(c);
}

hcg_status lat_moments(hcg_ctx* c, bool reset_force, bool want_rho) {
  (void)want_rho;   // the density always rides in slot 3 of the node velocity
  LatArgs a = make_args(c);
  const int64_t n = (int64_t)c->nxl*c->P;
  const unsigned nb = nblk(n, 256);
  {
  OpTimer tk(c, "kernel:k_moments");
  if (reset_force) k_moments<true><<<nb, 256, 0, c->stream>>>(c->g[c->cur], c->F, c->U, c->flags, a);
  else k_moments<false><<<nb, 256, 0, c->stream>>>(c->g[c->cur], c->F, c->U, c->flags, a);
  }
  KERNEL_CHECK(c);
  c->u_valid = true;
  return lat_halo_exchange_u(c);
}

hcg_status lat_reset_force(hcg_ctx* c) {
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
  c->u_valid = false;
  return lat_halo_exchange_pop(c);
}

hcg_status lat_pop_to_reference(hcg_ctx* c, double* dst_dev) {
  LatArgs a = make_args(c);
  k_to_reference<<<nblk((int64_t)c->nxl*c->P, 256), 256, 0, c->stream>>>(c->g[c->cur], dst_dev, a);
  KERNEL_CHECK(c);
  return HCG_OK;
}

hcg_status lat_pop_from_reference(hcg_ctx* c, const double* src_dev) {
  LatArgs a = make_args(c);
  double* s = c->g[1 - c->cur];
  CUDA_TRY(c, cudaMemsetAsync(s, 0, sizeof(double)*19*c->S, c->stream));
  k_pad<<<nblk(c->Nl, 256), 256, 0, c->stream>>>(src_dev, s, c->Nl, c->S, c->P, 19);
  KERNEL_CHECK(c);
  hcg_status st = ensure_qsets(c); if (st) return st;
  // g_q(n) = S_q(n + c_q): c_x = -1 populations read the LEFT ghost, c_x = +1 the RIGHT ghost
  st = exchange(c, s, c->P, h_qsets + 5, d_qsets + 5, 5, h_qsets, d_qsets, 5); if (st) return st;
  CUDA_TRY(c, cudaMemsetAsync(c->g[c->cur], 0, sizeof(double)*19*c->S, c->stream));
  k_from_reference<<<nblk((int64_t)c->nxl*c->P, 256), 256, 0, c->stream>>>(s, c->g[c->cur], a);
  KERNEL_CHECK(c);
  c->u_valid = false;
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
