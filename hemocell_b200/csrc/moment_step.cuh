// Per-node body of the moment-only update at tau = 1 (k_moment_step in lattice.cu; see the comment there).  Host + device: the
// same code is compiled for the CPU by tests/cpp/moment_host.cu, so that indexing, wrap and arithmetic are checked without a GPU
// (tests/test_moment_only_algorithm.py).
// Layout: W, F, U are AoS [n][4] over the padded slab (one ghost plane on each x side, plane = ny*nz nodes); node i is the i-th
// REAL node, n = i + P.  The lattice is periodic in y and z (wrap here) and in x (through the ghost planes, filled by the caller).
#pragma once
#include <stdint.h>

struct MomentArgs { int ny, nz; int64_t P; double body[3]; };

__host__ __device__ inline double mo_feq(double t, double cj, double rhoBar, double invRho, double jSqr) {
  return t * (rhoBar + 3.0*cj + invRho*(4.5*cj*cj - 1.5*jSqr));
}

template <bool WRITE_U>
__host__ __device__ inline void moment_node(const double* Win, const double* Fin, double* Wout, double* Fout, double* U,
                                            const MomentArgs& a, int64_t i) {
  constexpr int CX[19] = {0,-1,0,0,-1,-1,-1,-1,0,0, 1,0,0,1,1,1,1,0,0};
  constexpr int CY[19] = {0,0,-1,0,-1,1,0,0,-1,-1, 0,1,0,1,-1,0,0,1,1};
  constexpr int CZ[19] = {0,0,0,-1,0,0,-1,1,-1,1, 0,0,1,0,0,1,-1,1,-1};
  constexpr double T0 = 1.0/3.0, T1 = 1.0/18.0, T2 = 1.0/36.0;
  const int64_t n = i + a.P;
  const int rem = (int)(i % a.P);
  const int y = rem / a.nz, z = rem - y*a.nz;
  const int nz = a.nz, ny = a.ny;
  // source = this - c (periodic in y and z; x through the ghost planes)
  const int64_t oyp = (y + 1 < ny) ? nz : -(int64_t)(ny - 1)*nz, oym = (y > 0) ? -nz : (int64_t)(ny - 1)*nz;
  const int64_t ozp = (z + 1 < nz) ? 1 : -(nz - 1), ozm = (z > 0) ? -1 : nz - 1;
  double rb = 0.0, j0 = 0.0, j1 = 0.0, j2 = 0.0;
  double own0 = 0.0, own1 = 0.0, own2 = 0.0;
#pragma unroll
  for (int q = 0; q < 19; q++) {
    int64_t off = n - (int64_t)CX[q]*a.P;
    if (CY[q] == 1) off += oym; else if (CY[q] == -1) off += oyp;
    if (CZ[q] == 1) off += ozm; else if (CZ[q] == -1) off += ozp;
    double w0, w1, w2, w3, f0, f1, f2, f3;
#ifdef __CUDA_ARCH__
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(w0), "=d"(w1), "=d"(w2), "=d"(w3) : "l"(Win + 4*off));
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(f0), "=d"(f1), "=d"(f2), "=d"(f3) : "l"(Fin + 4*off) : "memory");
#else
    w0 = Win[4*off]; w1 = Win[4*off + 1]; w2 = Win[4*off + 2]; w3 = Win[4*off + 3];
    f0 = Fin[4*off]; f1 = Fin[4*off + 1]; f2 = Fin[4*off + 2]; f3 = Fin[4*off + 3];
#endif
    (void)f3;
    if (q == 0) { own0 = f0; own1 = f1; own2 = f2; }
    const double rho = 1.0 + w0, invRho = 1.0/rho;
    const double ux = w1*invRho + 0.5*f0, uy = w2*invRho + 0.5*f1, uz = w3*invRho + 0.5*f2;
    const double jx = rho*ux, jy = rho*uy, jz = rho*uz;
    const double jSqr = jx*jx + jy*jy + jz*jz;
    const double uF = ux*f0 + uy*f1 + uz*f2;
    const double t = (q == 0) ? T0 : ((q <= 3 || (q >= 10 && q <= 12)) ? T1 : T2);
    const double cj = CX[q]*jx + CY[q]*jy + CZ[q]*jz;
    const double cu = CX[q]*ux + CY[q]*uy + CZ[q]*uz;
    const double cF = CX[q]*f0 + CY[q]*f1 + CZ[q]*f2;
    const double ft = 3.0*(cF - uF) + 9.0*cu*cF;
    const double fq = mo_feq(t, cj, w0, invRho, jSqr) + t*0.5*ft;      // guo_collide_tau1, population q of the upstream node
    rb += fq;                                                          // moments19's order
    if (CX[q] == 1) j0 += fq; else if (CX[q] == -1) j0 -= fq;
    if (CY[q] == 1) j1 += fq; else if (CY[q] == -1) j1 -= fq;
    if (CZ[q] == 1) j2 += fq; else if (CZ[q] == -1) j2 -= fq;
  }
#ifdef __CUDA_ARCH__
  double2* Ww = reinterpret_cast<double2*>(Wout + 4*n);
  Ww[0] = make_double2(rb, j0); Ww[1] = make_double2(j1, j2);
  if (WRITE_U) {
    const double rho = 1.0 + rb, invRho = 1.0/rho;
    double2* Uw = reinterpret_cast<double2*>(U + 4*n);
    Uw[0] = make_double2(j0*invRho + 0.5*own0, j1*invRho + 0.5*own1); Uw[1] = make_double2(j2*invRho + 0.5*own2, rho);
  }
  double2* Fw = reinterpret_cast<double2*>(Fout + 4*n);
  Fw[0] = make_double2(a.body[0], a.body[1]); Fw[1] = make_double2(a.body[2], 0.0);
#else
  Wout[4*n] = rb; Wout[4*n + 1] = j0; Wout[4*n + 2] = j1; Wout[4*n + 3] = j2;
  if (WRITE_U) {
    const double rho = 1.0 + rb, invRho = 1.0/rho;
    U[4*n] = j0*invRho + 0.5*own0; U[4*n + 1] = j1*invRho + 0.5*own1; U[4*n + 2] = j2*invRho + 0.5*own2; U[4*n + 3] = rho;
  }
  Fout[4*n] = a.body[0]; Fout[4*n + 1] = a.body[1]; Fout[4*n + 2] = a.body[2]; Fout[4*n + 3] = 0.0;
#endif
}

// The 19 post-collision populations of a tau = 1 fluid node from its raw moments and force, re-associated so that opposite
// directions share their symmetric part (k_moment_tile in lattice.cu evaluates every node once and hands the populations to the
// neighbours through shared memory):
//   f*_q = t_q [ A + c.B + (c.u)(c.G) ],  A = rhoBar - 1.5 (j.j / rho + u.F),  B = 3 j + 1.5 F,  G = 4.5 (j + F),
// with u = j_in / rho + F / 2 and j = rho u, which is guo_collide_tau1 (lattice_node.cuh) term by term: (c.j)^2 / rho = (c.j)(c.u).
// ~100 fp64 operations for all 19 populations instead of ~30 per population; results agree with guo_collide_tau1 to rounding.
__host__ __device__ __forceinline__ void tau1_pops_fast(double w0, double w1, double w2, double w3, double f0, double f1, double f2,
                                                        double p[19]) {
  constexpr double T0 = 1.0/3.0, T1 = 1.0/18.0, T2 = 1.0/36.0;
  const double rho = 1.0 + w0, inv = 1.0/rho;
  const double ux = w1*inv + 0.5*f0, uy = w2*inv + 0.5*f1, uz = w3*inv + 0.5*f2;
  const double Jx = rho*ux, Jy = rho*uy, Jz = rho*uz;
  const double jSqr = Jx*Jx + Jy*Jy + Jz*Jz, uF = ux*f0 + uy*f1 + uz*f2;
  const double A = w0 - 1.5*(inv*jSqr + uF);
  const double Bx = 3.0*Jx + 1.5*f0, By = 3.0*Jy + 1.5*f1, Bz = 3.0*Jz + 1.5*f2;
  const double Gx = 4.5*(Jx + f0), Gy = 4.5*(Jy + f1), Gz = 4.5*(Jz + f2);
  p[0] = T0*A;
  { const double S = A + ux*Gx; p[10] = T1*(S + Bx); p[1] = T1*(S - Bx); }
  { const double S = A + uy*Gy; p[11] = T1*(S + By); p[2] = T1*(S - By); }
  { const double S = A + uz*Gz; p[12] = T1*(S + Bz); p[3] = T1*(S - Bz); }
  { const double S = A + (ux + uy)*(Gx + Gy), as = Bx + By; p[13] = T2*(S + as); p[4] = T2*(S - as); }     // ( 1, 1, 0) / (-1,-1, 0)
  { const double S = A + (ux - uy)*(Gx - Gy), as = Bx - By; p[14] = T2*(S + as); p[5] = T2*(S - as); }     // ( 1,-1, 0) / (-1, 1, 0)
  { const double S = A + (ux + uz)*(Gx + Gz), as = Bx + Bz; p[15] = T2*(S + as); p[6] = T2*(S - as); }     // ( 1, 0, 1) / (-1, 0,-1)
  { const double S = A + (ux - uz)*(Gx - Gz), as = Bx - Bz; p[16] = T2*(S + as); p[7] = T2*(S - as); }     // ( 1, 0,-1) / (-1, 0, 1)
  { const double S = A + (uy + uz)*(Gy + Gz), as = By + Bz; p[17] = T2*(S + as); p[8] = T2*(S - as); }     // ( 0, 1, 1) / ( 0,-1,-1)
  { const double S = A + (uy - uz)*(Gy - Gz), as = By - Bz; p[18] = T2*(S + as); p[9] = T2*(S - as); }     // ( 0, 1,-1) / ( 0,-1, 1)
}
