// Per-node arithmetic of the lattice kernels (lattice.cu): equilibrium, moments, Guo-BGK collision (generic and tau = 1),
// regularized velocity planes, Zou-He velocity / pressure nodes.  Host + device code: the kernels inline it on the device, and
// tests/cpp/lattice_node_host.cu compiles the same functions for the CPU so that the CPU test suite checks them against the oracle
// without a GPU (tests/test_lattice_node_host.py).
#pragma once

__host__ __device__ __forceinline__ double feq(double t, double cj, double rhoBar, double invRho, double jSqr) {
  return t * (rhoBar + 3.0*cj + invRho*(4.5*cj*cj - 1.5*jSqr));
}

__host__ __device__ __forceinline__ void moments19(const double f[19], double& rhoBar, double j[3]) {
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

__host__ __device__ __forceinline__ void guo_collide(double f[19], const double F[3], double omega) {
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
__host__ __device__ __forceinline__ void regularized_complete(double f[19], int o, const double uw[3]) {
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

// Zou-He velocity (pressure = false) / pressure (true) node with OUTWARD normal o (see oracle/hemo_oracle.c:zouhe_complete;
// helper/preInlet.cpp:399-436, examples/pipeflow_with_preinlet/pipeflow_with_preinlet.cpp:125-133): density (or normal
// velocity) from the known populations, bounce-back of the non-equilibrium part for the populations entering the
// domain, tangential momentum excess removed through the unknown diagonals.  bc = (u_x, u_y, u_z, rho) of the node.
__host__ __device__ __forceinline__ void zouhe_complete(double f[19], int o, bool pressure, double b0, double b1, double b2, double b3) {
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
  double rho, u[3];
  if (pressure) {
    rho = b3; u[0] = u[1] = u[2] = 0.0;
    const double un = sgn*((rho_on + 2.0*rho_out)/rho - 1.0);
    if (dir == 0) u[0] = un; else if (dir == 1) u[1] = un; else u[2] = un;
  } else {
    u[0] = b0; u[1] = b1; u[2] = b2;
    const double un = dir == 0 ? b0 : (dir == 1 ? b1 : b2);
    rho = (rho_on + 2.0*rho_out)/(1.0 + sgn*un);
  }
  const double rhoBar = rho - 1.0, invRho = 1.0/rho;
  const double jx = rho*u[0], jy = rho*u[1], jz = rho*u[2];
  const double jSqr = jx*jx + jy*jy + jz*jz;
#pragma unroll
  for (int q = 1; q < 19; q++) {
    const int cd = dir == 0 ? CC[q][0] : (dir == 1 ? CC[q][1] : CC[q][2]);
    if (cd*sgn < 0) {
      const int p = q <= 9 ? q + 9 : q - 9;
      const double cjq = CC[q][0]*jx + CC[q][1]*jy + CC[q][2]*jz;
      f[q] = f[p] - feq(TW[p], -cjq, rhoBar, invRho, jSqr) + feq(TW[q], cjq, rhoBar, invRho, jSqr);
    }
  }
  double jf[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int q = 0; q < 19; q++) { jf[0] += CC[q][0]*f[q]; jf[1] += CC[q][1]*f[q]; jf[2] += CC[q][2]*f[q]; }
  const double jt[3] = {jx, jy, jz};
#pragma unroll
  for (int k = 0; k < 3; k++) {
    if (k == dir) continue;
    const double diff = 0.5*(jf[k] - jt[k]);
#pragma unroll
    for (int q = 1; q < 19; q++) {
      const int cd = dir == 0 ? CC[q][0] : (dir == 1 ? CC[q][1] : CC[q][2]);
      if (cd*sgn < 0 && CC[q][k] != 0) f[q] -= CC[q][k]*diff;
    }
  }
}

// tau = 1 fast path.  With omega = 1 the BGK collision forgets the incoming populations: the post-collision
// state of a fluid node is f*_q = feq_q(rhoBar, j + rho F/2) + Guo_q(u, F), a function of the node's four raw
// moments and its force alone.  When the moments pass of the previous step kept (rhoBar, j) in W (it pulls the
// same 19 populations this kernel would pull), a fluid node reads 32 B of W + 32 B of F instead of 152 B of
// populations; wall nodes (bounce-back, velocity planes) take the generic pull path.  Same arithmetic as
// guo_collide with om1 = 0, omega = 1 (0*f + 1*feq is exact), so the results are those of k_collide_stream.
__host__ __device__ __forceinline__ void guo_collide_tau1(double f[19], double rhoBar, const double j[3], const double F[3]) {
  constexpr int CX[19] = {0,-1,0,0,-1,-1,-1,-1,0,0, 1,0,0,1,1,1,1,0,0};
  constexpr int CY[19] = {0,0,-1,0,-1,1,0,0,-1,-1, 0,1,0,1,-1,0,0,1,1};
  constexpr int CZ[19] = {0,0,0,-1,0,0,-1,1,-1,1, 0,0,1,0,0,1,-1,1,-1};
  constexpr double T0 = 1.0/3.0, T1 = 1.0/18.0, T2 = 1.0/36.0;
  const double rho = 1.0 + rhoBar, invRho = 1.0/rho;
  const double ux = j[0]*invRho + 0.5*F[0], uy = j[1]*invRho + 0.5*F[1], uz = j[2]*invRho + 0.5*F[2];
  const double jx = rho*ux, jy = rho*uy, jz = rho*uz;
  const double jSqr = jx*jx + jy*jy + jz*jz;
  const double fpre = 1.0 - 1.0/2.0;
  const double uF = ux*F[0] + uy*F[1] + uz*F[2];
#pragma unroll
  for (int q = 0; q < 19; q++) {
    const double t = (q == 0) ? T0 : ((q <= 3 || (q >= 10 && q <= 12)) ? T1 : T2);
    const double cj = CX[q]*jx + CY[q]*jy + CZ[q]*jz;
    const double cu = CX[q]*ux + CY[q]*uy + CZ[q]*uz;
    const double cF = CX[q]*F[0] + CY[q]*F[1] + CZ[q]*F[2];
    const double ft = 3.0*(cF - uF) + 9.0*cu*cF;
    f[q] = feq(t, cj, rhoBar, invRho, jSqr) + t*fpre*ft;
  }
}
