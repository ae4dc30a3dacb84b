/*
 * hemo_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See hemo_oracle.h for the status statement ("parity unpinned" per operator)
 * and the array conventions.  Compile with -ffp-contract=off so that every
 * expression below is evaluated exactly as written.
 *
 * Each function restates one piece of the reference; citations are relative
 * to /root/reference (UvaCsl/HemoCell) or, for the lattice, to SURVEY.md
 * Appendix C (Palabos v2.3.0 is not vendored in the reference tree).
 */
#include "hemo_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* OpenMP is used only by bench.py's cpu_baseline / --impl reference legs (ora_set_parallel(1));
 * the parity tests run the loops serially, in the reference's order. */
#ifdef _OPENMP
#include <omp.h>
#endif
static int ora_parallel = 0;
/* on > 1 also fixes the thread count (a launcher such as torchrun exports OMP_NUM_THREADS=1 to its children) */
void ora_set_parallel(int on) {
  ora_parallel = on != 0;
#ifdef _OPENMP
  if (on > 1) omp_set_num_threads(on);
#endif
}

/* D3Q19 of Palabos descriptors::D3Q19Descriptor (opposite = i+9,
 * patch/palabos.patch:492-497) */
static const int C[19][3] = {
  {0,0,0},
  {-1,0,0},{0,-1,0},{0,0,-1},{-1,-1,0},{-1,1,0},{-1,0,-1},{-1,0,1},{0,-1,-1},{0,-1,1},
  {1,0,0},{0,1,0},{0,0,1},{1,1,0},{1,-1,0},{1,0,1},{1,0,-1},{0,1,1},{0,1,-1}};
static const double TW[19] = {
  1.0/3.0,
  1.0/18.0,1.0/18.0,1.0/18.0,1.0/36.0,1.0/36.0,1.0/36.0,1.0/36.0,1.0/36.0,1.0/36.0,
  1.0/18.0,1.0/18.0,1.0/18.0,1.0/36.0,1.0/36.0,1.0/36.0,1.0/36.0,1.0/36.0,1.0/36.0};
static inline int opp(int i) { return i == 0 ? 0 : (i <= 9 ? i + 9 : i - 9); }

static inline int64_t nidx(const ora_domain* d, int x, int y, int z) {
  return (int64_t)z + (int64_t)d->nz * ((int64_t)y + (int64_t)d->ny * (int64_t)x);
}
static inline int64_t nnodes(const ora_domain* d) { return (int64_t)d->nx * d->ny * d->nz; }

/* wrap a node coordinate; returns 0 if it falls outside a non-periodic axis */
static inline int wrap1(int* v, int n, int periodic) {
  if (*v >= 0 && *v < n) return 1;
  if (!periodic) return 0;
  *v %= n; if (*v < 0) *v += n;
  return 1;
}

/* second-order equilibrium in the f - t formulation (dynamicsTemplates::bgk_ma2_equilibrium) */
static inline double feq(int i, double rhoBar, double invRho, const double j[3], double jSqr) {
  double cj = C[i][0]*j[0] + C[i][1]*j[1] + C[i][2]*j[2];
  return TW[i] * (rhoBar + 3.0*cj + invRho*(4.5*cj*cj - 1.5*jSqr));
}

/* initializeAtEquilibrium (core/hemoCell.cpp:129-133 -> Palabos): f_i = feq(rhoBar = rho-1, j = rho u) */
void ora_init_equilibrium(const ora_domain* d, double rho, const double u[3], double* pop) {
  int64_t N = nnodes(d);
  double rhoBar = rho - 1.0, invRho = 1.0 / rho;
  double j[3] = {rho*u[0], rho*u[1], rho*u[2]};
  double jSqr = j[0]*j[0] + j[1]*j[1] + j[2]*j[2];
  for (int i = 0; i < 19; i++) {
    double v = feq(i, rhoBar, invRho, j, jSqr);
    for (int64_t n = 0; n < N; n++) pop[i*N + n] = v;
  }
}

/* GuoExternalForceBGKdynamics::collide (called through core/hemoCell.cpp:317; dynamics
 * installed at core/hemoCell.cpp:459 and in every case file).  SURVEY.md Appendix C. */
static void guo_bgk_collide(double f[19], const double F[3], double omega) {
  double rhoBar = 0.0, j[3] = {0.0, 0.0, 0.0};
  for (int i = 0; i < 19; i++) {
    rhoBar += f[i];
    j[0] += C[i][0]*f[i]; j[1] += C[i][1]*f[i]; j[2] += C[i][2]*f[i];
  }
  double rho = 1.0 + rhoBar, invRho = 1.0 / rho;
  double u[3];
  for (int k = 0; k < 3; k++) u[k] = j[k]*invRho + 0.5*F[k];     /* computeVelocity */
  for (int k = 0; k < 3; k++) j[k] = rho*u[k];
  double jSqr = j[0]*j[0] + j[1]*j[1] + j[2]*j[2];
  for (int i = 0; i < 19; i++) {                                   /* bgk_ma2_collision */
    f[i] *= (1.0 - omega);
    f[i] += omega * feq(i, rhoBar, invRho, j, jSqr);
  }
  for (int i = 0; i < 19; i++) {                                   /* addGuoForce, amplitude 1 */
    double cu = C[i][0]*u[0] + C[i][1]*u[1] + C[i][2]*u[2];
    double ft = 0.0;
    for (int k = 0; k < 3; k++)
      ft += ((C[i][k] - u[k])*3.0 + cu*C[i][k]*9.0) * F[k];
    ft *= TW[i] * (1.0 - omega/2.0);
    f[i] += ft;
  }
}

/* Regularized ("local") velocity boundary, createLocalBoundaryCondition3D
 * (examples/oneCellShear/oneCellShear.cpp:72-73, helper/hemocellInit.hh:64-77): density from
 * the known populations and the imposed wall velocity, bounce-back of the non-equilibrium
 * parts for the unknowns, Pi_neq, then every population regularised; the base (Guo BGK)
 * dynamics collides afterwards.  `o` = flag-2 encodes the OUTWARD normal (-x,+x,-y,+y,-z,+z). */
static void regularized_velocity_complete(double f[19], int o, const double uw[3]) {
  int dir = o / 2, sgn = (o & 1) ? +1 : -1;          /* outward normal = sgn * e_dir */
  double rho_on = 0.0, rho_out = 0.0;                /* c.n_out == 0, c.n_out > 0 (known, leaving) */
  for (int i = 0; i < 19; i++) {
    int cn = C[i][dir]*sgn;
    if (cn == 0) rho_on += f[i] + TW[i];
    else if (cn > 0) rho_out += f[i] + TW[i];
  }
  double rho = (rho_on + 2.0*rho_out) / (1.0 + sgn*uw[dir]);
  double rhoBar = rho - 1.0, invRho = 1.0 / rho;
  double j[3] = {rho*uw[0], rho*uw[1], rho*uw[2]};
  double jSqr = j[0]*j[0] + j[1]*j[1] + j[2]*j[2];
  double fneq[19], eq[19];
  for (int i = 0; i < 19; i++) eq[i] = feq(i, rhoBar, invRho, j, jSqr);
  for (int i = 0; i < 19; i++) {
    int cn = C[i][dir]*sgn;
    if (cn >= 0) fneq[i] = f[i] - eq[i];
  }
  for (int i = 0; i < 19; i++) {
    int cn = C[i][dir]*sgn;
    if (cn < 0) fneq[i] = fneq[opp(i)];              /* unknown: entering the domain */
  }
  double Pi[6] = {0,0,0,0,0,0};                      /* xx xy xz yy yz zz */
  for (int i = 0; i < 19; i++) {
    Pi[0] += C[i][0]*C[i][0]*fneq[i]; Pi[1] += C[i][0]*C[i][1]*fneq[i];
    Pi[2] += C[i][0]*C[i][2]*fneq[i]; Pi[3] += C[i][1]*C[i][1]*fneq[i];
    Pi[4] += C[i][1]*C[i][2]*fneq[i]; Pi[5] += C[i][2]*C[i][2]*fneq[i];
  }
  const double cs2 = 1.0/3.0;
  for (int i = 0; i < 19; i++) {
    double q = (C[i][0]*C[i][0] - cs2)*Pi[0] + 2.0*C[i][0]*C[i][1]*Pi[1]
             + 2.0*C[i][0]*C[i][2]*Pi[2] + (C[i][1]*C[i][1] - cs2)*Pi[3]
             + 2.0*C[i][1]*C[i][2]*Pi[4] + (C[i][2]*C[i][2] - cs2)*Pi[5];
    f[i] = eq[i] + TW[i]*4.5*q;                       /* t_i/(2 cs^4) Q:Pi */
  }
}

/* Zou-He velocity / pressure boundary nodes (the pre-inlet coupling and the outlet of
 * examples/pipeflow_with_preinlet/pipeflow_with_preinlet.cpp:125-133, helper/preInlet.cpp:399-436:
 * createZouHeBoundaryCondition3D -> addVelocityBoundary{0,1,2}{N,P} on single nodes,
 * WrappedZouHeBoundaryManager3D -> addPressureBoundary0N + setBoundaryDensity).  Palabos is not in the
 * tree; restated from the published scheme (Zou & He 1997; 3-D form of Hecht & Harting 2010) as Palabos'
 * ZouHeDynamics::completePopulations implements it:
 *   velocity node: rho from the known populations and the imposed u (same closure as the regularized plane);
 *   pressure node: rho imposed, normal velocity from the same closure, tangential velocity 0;
 *   unknown populations (c.n_out < 0): bounce-back of the non-equilibrium part,
 *   then the tangential momentum excess is removed through the unknown diagonal populations
 *   (half of it on each of the two diagonals that carry the component), which makes rho and j exact.
 * The base dynamics (Guo BGK) collides afterwards.  bc = (u_x, u_y, u_z, rho) of the node. */
static void zouhe_complete(double f[19], int o, int pressure, const double bc[4]) {
  int dir = o / 2, sgn = (o & 1) ? +1 : -1;
  double rho_on = 0.0, rho_out = 0.0;
  for (int i = 0; i < 19; i++) {
    int cn = C[i][dir]*sgn;
    if (cn == 0) rho_on += f[i] + TW[i];
    else if (cn > 0) rho_out += f[i] + TW[i];
  }
  double rho, u[3];
  if (pressure) {
    rho = bc[3];
    u[0] = u[1] = u[2] = 0.0;
    u[dir] = sgn * ((rho_on + 2.0*rho_out) / rho - 1.0);
  } else {
    u[0] = bc[0]; u[1] = bc[1]; u[2] = bc[2];
    rho = (rho_on + 2.0*rho_out) / (1.0 + sgn*u[dir]);
  }
  double rhoBar = rho - 1.0, invRho = 1.0 / rho;
  double j[3] = {rho*u[0], rho*u[1], rho*u[2]};
  double jSqr = j[0]*j[0] + j[1]*j[1] + j[2]*j[2];
  for (int i = 1; i < 19; i++) {
    if (C[i][dir]*sgn < 0) f[i] = f[opp(i)] - feq(opp(i), rhoBar, invRho, j, jSqr) + feq(i, rhoBar, invRho, j, jSqr);
  }
  double jf[3] = {0.0, 0.0, 0.0};
  for (int i = 0; i < 19; i++) { jf[0] += C[i][0]*f[i]; jf[1] += C[i][1]*f[i]; jf[2] += C[i][2]*f[i]; }
  for (int k = 0; k < 3; k++) {
    if (k == dir) continue;
    double diff = 0.5*(jf[k] - j[k]);
    for (int i = 1; i < 19; i++)
      if (C[i][dir]*sgn < 0 && C[i][k] != 0) f[i] -= C[i][k]*diff;
  }
}

/* MultiBlockLattice3D::collideAndStream (core/hemoCell.cpp:317): collide every node with its
 * dynamics, then stream f_i(x + c_i) <- f_i(x) (Palabos' swap scheme is equivalent to this
 * push).  BounceBack::collide = swap with the opposite population (full-way bounce back).
 * Periodic axes wrap; what leaves a non-periodic face is dropped and what would enter
 * through it is the rest equilibrium (stored value 0).
 * bc_node (may be NULL): per-node boundary values [4][N] = (u_x, u_y, u_z, rho) of the Zou-He nodes. */
void ora_collide_and_stream_io(const ora_domain* d, const uint8_t* flags, double* pop,
                               const double* force, double* scratch, const double* bc_node) {
  int64_t N = nnodes(d);
  #pragma omp parallel for schedule(static) if(ora_parallel)
  for (int64_t n = 0; n < N; n++) {
    double f[19], F[3];
    for (int i = 0; i < 19; i++) f[i] = pop[i*N + n];
    uint8_t fl = flags[n];
    if (fl == ORA_BB) {
      for (int i = 1; i <= 9; i++) { double t = f[i]; f[i] = f[i+9]; f[i+9] = t; }
    } else {
      for (int k = 0; k < 3; k++) F[k] = force[k*N + n];
      if (fl >= ORA_ZH_VEL_XN) {
        double bc[4] = {0.0, 0.0, 0.0, 1.0};
        if (bc_node) for (int k = 0; k < 4; k++) bc[k] = bc_node[k*N + n];
        if (fl >= ORA_ZH_PRES_XN) zouhe_complete(f, fl - ORA_ZH_PRES_XN, 1, bc);
        else zouhe_complete(f, fl - ORA_ZH_VEL_XN, 0, bc);
      } else if (fl >= ORA_VEL_XN) regularized_velocity_complete(f, fl - 2, d->bc_vel[fl - 2]);
      guo_bgk_collide(f, F, d->omega);
    }
    for (int i = 0; i < 19; i++) pop[i*N + n] = f[i];
  }
  memset(scratch, 0, sizeof(double)*19*N);
  #pragma omp parallel for schedule(static) if(ora_parallel)
  for (int x = 0; x < d->nx; x++) for (int y = 0; y < d->ny; y++) for (int z = 0; z < d->nz; z++) {
    int64_t n = nidx(d, x, y, z);
    for (int i = 0; i < 19; i++) {
      int xx = x + C[i][0], yy = y + C[i][1], zz = z + C[i][2];
      if (!wrap1(&xx, d->nx, d->periodic[0])) continue;
      if (!wrap1(&yy, d->ny, d->periodic[1])) continue;
      if (!wrap1(&zz, d->nz, d->periodic[2])) continue;
      scratch[i*N + nidx(d, xx, yy, zz)] = pop[i*N + n];
    }
  }
  memcpy(pop, scratch, sizeof(double)*19*N);
}

void ora_collide_and_stream(const ora_domain* d, const uint8_t* flags, double* pop,
                            const double* force, double* scratch) {
  ora_collide_and_stream_io(d, flags, pop, force, scratch, NULL);
}

/* Cell::computeVelocity / computeDensity as used by IBM interpolation
 * (core/hemoCellParticleField.cpp:833), FluidInfo (helper/fluidInfo.cpp:46) and the pre-inlet velocity
 * coupling (helper/preInlet.cpp:372):
 * Guo dynamics: u = j/rho + F/2; BounceBack: u = 0, rho = 1; velocity BC: u = wall velocity;
 * Zou-He velocity node: the imposed u; Zou-He pressure node: the imposed rho, normal velocity from the
 * known populations, tangential velocity 0. */
void ora_moments_io(const ora_domain* d, const uint8_t* flags, const double* pop,
                    const double* force, double* rho_out, double* vel, const double* bc_node) {
  int64_t N = nnodes(d);
  #pragma omp parallel for schedule(static) if(ora_parallel)
  for (int64_t n = 0; n < N; n++) {
    double rhoBar = 0.0, j[3] = {0,0,0};
    for (int i = 0; i < 19; i++) {
      double f = pop[i*N + n];
      rhoBar += f; j[0] += C[i][0]*f; j[1] += C[i][1]*f; j[2] += C[i][2]*f;
    }
    double rho = 1.0 + rhoBar, invRho = 1.0/rho;
    uint8_t fl = flags[n];
    if (fl >= ORA_ZH_PRES_XN) {
      int o = fl - ORA_ZH_PRES_XN, dir = o / 2, sgn = (o & 1) ? +1 : -1;
      double rho_on = 0.0, rho_o = 0.0;
      for (int i = 0; i < 19; i++) {
        int cn = C[i][dir]*sgn;
        if (cn == 0) rho_on += pop[i*N + n] + TW[i]; else if (cn > 0) rho_o += pop[i*N + n] + TW[i];
      }
      rho = bc_node ? bc_node[3*N + n] : 1.0;
      for (int k = 0; k < 3; k++) vel[k*N + n] = 0.0;
      vel[dir*N + n] = sgn * ((rho_on + 2.0*rho_o) / rho - 1.0);
    } else {
      for (int k = 0; k < 3; k++) {
        double u;
        if (fl == ORA_BB) u = 0.0;
        else if (fl >= ORA_ZH_VEL_XN) u = bc_node ? bc_node[k*N + n] : 0.0;
        else if (fl >= ORA_VEL_XN) u = d->bc_vel[fl-2][k];
        else u = j[k]*invRho + 0.5*force[k*N + n];
        vel[k*N + n] = u;
      }
    }
    if (rho_out) rho_out[n] = (fl == ORA_BB) ? 1.0 : rho;
  }
}

void ora_moments(const ora_domain* d, const uint8_t* flags, const double* pop,
                 const double* force, double* rho_out, double* vel) {
  ora_moments_io(d, flags, pop, force, rho_out, vel, NULL);
}

/* interpolationCoefficientsPhi2 (core/immersedBoundaryMethod.h:62-138), phi2 (:37-41).
 * Candidate nodes centre + {-1,0,1}^3 in x-outer / z-inner order, weight = product of
 * max(0, 1-|delta|), zero weights and boundary nodes skipped, then normalised.
 * The single global lattice replaces the reference's block + envelope: nodes are wrapped
 * on periodic axes and skipped outside non-periodic ones.  Returns the entry count. */
static inline double phi2(double x) { x = fabs(x); x = 1.0 - x; return x > 0.0 ? x : 0.0; }

int ora_ibm_kernel(const ora_domain* d, const uint8_t* flags, const double p[3],
                   int64_t node[8], double w[8]) {
  /* the reference truncates (plint)(p + 0.5); floor() gives the same candidate set for the
   * two nodes per axis that can carry weight and is also right for unwrapped p < 0 */
  int c[3] = {(int)floor(p[0] + 0.5), (int)floor(p[1] + 0.5), (int)floor(p[2] + 0.5)};
  int n = 0; double total = 0.0;
  for (int dx = -1; dx < 2; dx++) for (int dy = -1; dy < 2; dy++) for (int dz = -1; dz < 2; dz++) {
    int q[3] = {c[0] + dx, c[1] + dy, c[2] + dz};
    double weight = phi2(p[0] - q[0]) * phi2(p[1] - q[1]) * phi2(p[2] - q[2]);
    if (weight == 0.0) continue;
    if (!wrap1(&q[0], d->nx, d->periodic[0])) continue;
    if (!wrap1(&q[1], d->ny, d->periodic[1])) continue;
    if (!wrap1(&q[2], d->nz, d->periodic[2])) continue;
    int64_t id = nidx(d, q[0], q[1], q[2]);
    if (flags[id] != ORA_FLUID) continue;            /* getDynamics().isBoundary() */
    total += weight;
    node[n] = id; w[n] = weight; n++;
  }
  double coeff = 1.0 / total;
  for (int k = 0; k < n; k++) w[k] *= coeff;
  return n;
}

/* spreadParticleForce (core/hemoCellParticleField.cpp:841-863) incl. the in-place force cap */
void ora_spread(const ora_domain* d, const uint8_t* flags, int64_t np, const double* pos,
                double* pforce, const double* frep, double f_limit, double* node_force) {
  int64_t N = nnodes(d);
  #pragma omp parallel for schedule(static) if(ora_parallel)
  for (int64_t p = 0; p < np; p++) {
    int64_t node[8]; double w[8];
    int n = ora_ibm_kernel(d, flags, pos + 3*p, node, w);
    double* f = pforce + 3*p;
    double mag = sqrt(f[0]*f[0] + f[1]*f[1] + f[2]*f[2]);
    if (mag > f_limit) { double s = f_limit/mag; f[0] *= s; f[1] *= s; f[2] *= s; }
    for (int k = 0; k < n; k++)
      for (int c = 0; c < 3; c++) {
        const double add = (frep[3*p + c] + f[c]) * w[k];
        #pragma omp atomic
        node_force[c*N + node[k]] += add;
      }
  }
}

/* interpolateFluidVelocity (core/hemoCellParticleField.cpp:819-839): v = sum_j u_j w_j with the
 * kernel built at the same (pre-advance) position that spread used */
void ora_interpolate(const ora_domain* d, const uint8_t* flags, int64_t np, const double* pos,
                     const double* pop, const double* node_force, double* vel) {
  int64_t N = nnodes(d);
  #pragma omp parallel for schedule(static) if(ora_parallel)
  for (int64_t p = 0; p < np; p++) {
    int64_t node[8]; double w[8];
    int n = ora_ibm_kernel(d, flags, pos + 3*p, node, w);
    double v[3] = {0,0,0};
    for (int k = 0; k < n; k++) {
      double rhoBar = 0.0, j[3] = {0,0,0};
      for (int i = 0; i < 19; i++) {
        double f = pop[i*N + node[k]];
        rhoBar += f; j[0] += C[i][0]*f; j[1] += C[i][1]*f; j[2] += C[i][2]*f;
      }
      double invRho = 1.0/(1.0 + rhoBar);
      for (int c = 0; c < 3; c++) {
        double u = j[c]*invRho + 0.5*node_force[c*N + node[k]];
        v[c] += u * w[k];
      }
    }
    vel[3*p] = v[0]; vel[3*p+1] = v[1]; vel[3*p+2] = v[2];
  }
}

/* HemoCellParticle::advance (core/hemoCellParticle.h:188-203, Euler) and the boundary test of
 * advanceParticles (core/hemoCellParticleField.cpp:566-588) */
int64_t ora_advance(const ora_domain* d, const uint8_t* flags, int64_t np, double* pos,
                    const double* vel, uint8_t* hit) {
  int64_t nhit = 0;
  #pragma omp parallel for schedule(static) reduction(+:nhit) if(ora_parallel)
  for (int64_t p = 0; p < np; p++) {
    for (int c = 0; c < 3; c++) pos[3*p + c] += vel[3*p + c];
    int q[3] = {(int)floor(pos[3*p] + 0.5), (int)floor(pos[3*p+1] + 0.5), (int)floor(pos[3*p+2] + 0.5)};
    uint8_t h = 0;
    if (wrap1(&q[0], d->nx, d->periodic[0]) && wrap1(&q[1], d->ny, d->periodic[1]) &&
        wrap1(&q[2], d->nz, d->periodic[2])) {
      if (flags[nidx(d, q[0], q[1], q[2])] != ORA_FLUID) h = 1;
    }
    if (hit) hit[p] = h;
    nhit += h;
  }
  return nhit;
}

/* ------------------------------------------------------------------ mechanics */
static inline void v3sub(const double* a, const double* b, double* r) { r[0]=a[0]-b[0]; r[1]=a[1]-b[1]; r[2]=a[2]-b[2]; }
static inline void v3cross(const double* a, const double* b, double* r) {   /* helper/array.h:199-203 */
  r[0] = a[1]*b[2] - a[2]*b[1]; r[1] = a[2]*b[0] - a[0]*b[2]; r[2] = a[0]*b[1] - a[1]*b[0];
}
static inline double v3dot(const double* a, const double* b) {              /* helper/array.h:220-226 */
  double r = 0; r += a[0]*b[0]; r += a[1]*b[1]; r += a[2]*b[2]; return r;
}
static inline double v3norm(const double* a) {                              /* helper/array.h:238-244 */
  double r = 0; r += a[0]*a[0]; r += a[1]*a[1]; r += a[2]*a[2]; return sqrt(r);
}
/* helper/array.h:271-285 */
static void tri_area_normal(const double* v0, const double* v1, const double* v2, double* area, double* n) {
  double e01[3], e02[3]; v3sub(v1, v0, e01); v3sub(v2, v0, e02);
  v3cross(e01, e02, n);
  double nn = v3norm(n);
  if (nn != 0.0) { *area = 0.5*nn; n[0] /= nn; n[1] /= nn; n[2] /= nn; }
  else { *area = 0.0; n[0] = n[1] = n[2] = 0.0; }
}
#define ADD3(dst, s, v) do { (dst)[0] += (s)*(v)[0]; (dst)[1] += (s)*(v)[1]; (dst)[2] += (s)*(v)[2]; } while (0)

static void mech_one_cell(const ora_celltype* t, const double* x, const double* vel, double* F,
                          double* Fa, double* Fv, double* Fb, double* Fl, double* Fvi, double* Fin) {
  const int T = t->n_triangles, V = t->n_vertices, E = t->n_edges;
  double* areas = (double*)malloc(sizeof(double)*T);
  double* normals = (double*)malloc(sizeof(double)*3*T);
  double volume = 0.0;
  /* per-triangle: signed volume, area force (rbcHighOrderModel.cpp:56-98 == pltSimpleModel.cpp:60-101) */
  for (int k = 0; k < T; k++) {
    const int* tr = t->triangles + 3*k;
    const double *v0 = x + 3*tr[0], *v1 = x + 3*tr[1], *v2 = x + 3*tr[2];
    const double v210 = v2[0]*v1[1]*v0[2];
    const double v120 = v1[0]*v2[1]*v0[2];
    const double v201 = v2[0]*v0[1]*v1[2];
    const double v021 = v0[0]*v2[1]*v1[2];
    const double v102 = v1[0]*v0[1]*v2[2];
    const double v012 = v0[0]*v1[1]*v2[2];
    volume += (-v210+v120+v201-v021-v102+v012);
    double area, n[3];
    tri_area_normal(v0, v1, v2, &area, n);
    const double areaRatio = (area - t->triangle_area_eq[k]) / t->triangle_area_eq[k];
    const double afm = t->k_area * (areaRatio + areaRatio/fabs(0.09 - areaRatio*areaRatio));
    double c[3];
    c[0] = (v0[0]+v1[0]+v2[0])/3.0; c[1] = (v0[1]+v1[1]+v2[1])/3.0; c[2] = (v0[2]+v1[2]+v2[2])/3.0;
    for (int m = 0; m < 3; m++) {
      const double* vm = x + 3*tr[m];
      double av[3]; v3sub(c, vm, av);
      double f[3] = {afm*av[0], afm*av[1], afm*av[2]};
      ADD3(F + 3*tr[m], 1.0, f);
      if (Fa) ADD3(Fa + 3*tr[m], 1.0, f);
    }
    areas[k] = area; normals[3*k] = n[0]; normals[3*k+1] = n[1]; normals[3*k+2] = n[2];
  }
  volume *= (1.0/6.0);
  /* global volume force (rbcHighOrderModel.cpp:100-124) */
  const double volume_frac = (volume - t->volume_eq)/t->volume_eq;
  const double volume_force = -t->k_volume * volume_frac/fabs(0.01 - volume_frac*volume_frac);
  for (int k = 0; k < T; k++) {
    const int* tr = t->triangles + 3*k;
    const double s = areas[k]/t->area_mean_eq;
    double f[3] = {(volume_force*normals[3*k])*s, (volume_force*normals[3*k+1])*s, (volume_force*normals[3*k+2])*s};
    for (int m = 0; m < 3; m++) { ADD3(F + 3*tr[m], 1.0, f); if (Fv) ADD3(Fv + 3*tr[m], 1.0, f); }
  }
  if (t->model == 0) {
    /* per-vertex bending (rbcHighOrderModel.cpp:127-166) */
    for (int i = 0; i < V; i++) {
      const int nn = t->vertex_n_vertexes[i];
      const int* ring = t->vertex_vertexes + 6*i;
      double sum[3] = {0,0,0};
      for (int j = 0; j < nn; j++) { sum[0] += x[3*ring[j]]; sum[1] += x[3*ring[j]+1]; sum[2] += x[3*ring[j]+2]; }
      double mid[3] = {sum[0]/nn, sum[1]/nn, sum[2]/nn};
      double dev[3]; v3sub(mid, x + 3*i, dev);
      double pn[3] = {0,0,0};
      for (int j = 0; j < nn; j++) {              /* j = nn-1 wraps to ring[0], same order as the reference */
        double a[3], b[3], tn[3];
        v3sub(x + 3*ring[j], x + 3*i, a);
        v3sub(x + 3*ring[(j+1) % nn], x + 3*i, b);
        v3cross(a, b, tn);
        double l = v3norm(tn);
        tn[0] /= l; tn[1] /= l; tn[2] /= l;
        pn[0] += tn[0]; pn[1] += tn[1]; pn[2] += tn[2];
      }
      double l = v3norm(pn); pn[0] /= l; pn[1] /= l; pn[2] /= l;
      const double ndev = v3dot(pn, dev);
      const double dDev = (ndev - t->patch_dist_eq[i]) / t->edge_mean_eq;
      const double s = t->k_bend * (dDev + dDev/fabs(0.0555 - dDev*dDev));
      double bf[3] = {s*pn[0], s*pn[1], s*pn[2]};
      ADD3(F + 3*i, 1.0, bf); if (Fb) ADD3(Fb + 3*i, 1.0, bf);
      double nb[3] = {-bf[0]/nn, -bf[1]/nn, -bf[2]/nn};
      for (int j = 0; j < nn; j++) { ADD3(F + 3*ring[j], 1.0, nb); if (Fb) ADD3(Fb + 3*ring[j], 1.0, nb); }
    }
  }
  /* per-edge link (+ membrane viscosity; + PLT dihedral bending) */
  for (int e = 0; e < E; e++) {
    const int a = t->edges[2*e], b = t->edges[2*e+1];
    const double *p0 = x + 3*a, *p1 = x + 3*b;
    double ev[3]; v3sub(p1, p0, ev);
    const double len = v3norm(ev);
    double uv[3] = {ev[0]/len, ev[1]/len, ev[2]/len};
    const double frac = (len - t->edge_length_eq[e]) / t->edge_length_eq[e];
    const double fs = t->k_link * (frac + frac/fabs(9.0 - frac*frac));
    double f[3] = {uv[0]*fs, uv[1]*fs, uv[2]*fs};
    ADD3(F + 3*a, 1.0, f); ADD3(F + 3*b, -1.0, f);
    if (Fl) { ADD3(Fl + 3*a, 1.0, f); ADD3(Fl + 3*b, -1.0, f); }
    if (t->model == 1 || t->eta_m != 0.0) {
      /* rbcHighOrderModel.cpp:186-201 (only if eta_m != 0), pltSimpleModel.cpp:139-153 (always) */
      double rv[3]; v3sub(vel + 3*b, vel + 3*a, rv);
      const double pr = v3dot(rv, uv);
      double fv[3] = {t->eta_m*(pr*uv[0]), t->eta_m*(pr*uv[1]), t->eta_m*(pr*uv[2])};
      const double mag = v3norm(fv);
      if (mag > 50.0/4.0) { double s = (50.0/4.0)/mag; fv[0] *= s; fv[1] *= s; fv[2] *= s; }
      ADD3(F + 3*a, 1.0, fv); ADD3(F + 3*b, -1.0, fv);
      if (Fvi) { ADD3(Fvi + 3*a, 1.0, fv); ADD3(Fvi + 3*b, -1.0, fv); }
    }
    if (t->model == 1) {
      /* pltSimpleModel.cpp:156-182 */
      const int b0 = t->edge_bending_triangles[2*e], b1 = t->edge_bending_triangles[2*e+1];
      const int* t0 = t->triangles + 3*b0; const int* t1 = t->triangles + 3*b1;
      double V1[3], V2[3], ar;
      tri_area_normal(x + 3*t0[0], x + 3*t0[1], x + 3*t0[2], &ar, V1);
      tri_area_normal(x + 3*t1[0], x + 3*t1[1], x + 3*t1[2], &ar, V2);
      double cr[3]; v3cross(V1, V2, cr);
      const double angle = atan2(v3dot(cr, uv), v3dot(V1, V2));    /* helper/geometryUtils.h:49-52 */
      const double af = angle - t->edge_angle_eq[e];
      const double fm = t->k_bend * (af + af/fabs(2.467 - af*af));
      double bf[3] = {fm*(V1[0]+V2[0])*0.5, fm*(V1[1]+V2[1])*0.5, fm*(V1[2]+V2[2])*0.5};
      const int o0 = t->edge_bending_outer_points[2*e], o1 = t->edge_bending_outer_points[2*e+1];
      ADD3(F + 3*a, 1.0, bf); ADD3(F + 3*b, 1.0, bf); ADD3(F + 3*o0, -1.0, bf); ADD3(F + 3*o1, -1.0, bf);
      if (Fb) { ADD3(Fb + 3*a, 1.0, bf); ADD3(Fb + 3*b, 1.0, bf); ADD3(Fb + 3*o0, -1.0, bf); ADD3(Fb + 3*o1, -1.0, bf); }
    }
  }
  if (t->model == 1) {
    /* inner links, linear part only (pltSimpleModel.cpp:188-205) */
    for (int e = 0; e < t->n_inner_edges; e++) {
      const int a = t->inner_edges[2*e], b = t->inner_edges[2*e+1];
      double ev[3]; v3sub(x + 3*b, x + 3*a, ev);
      const double len = sqrt(ev[0]*ev[0] + ev[1]*ev[1] + ev[2]*ev[2]);
      double uv[3] = {ev[0]/len, ev[1]/len, ev[2]/len};
      const double frac = (len - t->inner_edge_length_eq[e]) / t->inner_edge_length_eq[e];
      const double fs = t->k_link * 5.0 * frac;
      double f[3] = {uv[0]*fs, uv[1]*fs, uv[2]*fs};
      ADD3(F + 3*a, 1.0, f); ADD3(F + 3*b, -1.0, f);
      if (Fin) { ADD3(Fin + 3*a, 1.0, f); ADD3(Fin + 3*b, -1.0, f); }
    }
  }
  free(areas); free(normals);
}

void ora_mechanics(const ora_celltype* t, int64_t n_cells, const double* pos, const double* vel,
                   double* force, double* const* comp) {
  const int64_t V = t->n_vertices;
  #pragma omp parallel for schedule(dynamic, 8) if(ora_parallel)
  for (int64_t c = 0; c < n_cells; c++) {
    int64_t o = 3*V*c;
    mech_one_cell(t, pos + o, vel + o, force + o,
                  comp ? comp[0] + o : 0, comp ? comp[1] + o : 0, comp ? comp[2] + o : 0,
                  comp ? comp[3] + o : 0, comp ? comp[4] + o : 0, comp ? comp[5] + o : 0);
  }
}

/* ------------------------------------------------------------------ repulsion */
/* bins by nearest node (update_pg, hemoCellParticleField.cpp:137-168); no 10-per-node cap */
typedef struct { int64_t* head; int64_t* next; } bins_t;
static int bin_of(const ora_domain* d, const double* p, int q[3]) {
  q[0] = (int)floor(p[0] + 0.5); q[1] = (int)floor(p[1] + 0.5); q[2] = (int)floor(p[2] + 0.5);
  return wrap1(&q[0], d->nx, d->periodic[0]) && wrap1(&q[1], d->ny, d->periodic[1]) &&
         wrap1(&q[2], d->nz, d->periodic[2]);
}
static bins_t build_bins(const ora_domain* d, int64_t np, const double* pos) {
  bins_t b; int64_t N = nnodes(d);
  b.head = (int64_t*)malloc(sizeof(int64_t)*N); b.next = (int64_t*)malloc(sizeof(int64_t)*(np > 0 ? np : 1));
  for (int64_t n = 0; n < N; n++) b.head[n] = -1;
  for (int64_t p = np - 1; p >= 0; p--) {            /* reverse insert => lists in ascending order */
    int q[3];
    if (!bin_of(d, pos + 3*p, q)) { b.next[p] = -2; continue; }
    int64_t n = nidx(d, q[0], q[1], q[2]);
    b.next[p] = b.head[n]; b.head[n] = p;
  }
  return b;
}
static inline void min_image(const ora_domain* d, double dv[3]) {
  const int n[3] = {d->nx, d->ny, d->nz};
  for (int k = 0; k < 3; k++)
    if (d->periodic[k]) dv[k] -= n[k]*rint(dv[k]/n[k]);
}
static void rep_pairs(const ora_domain* d, const bins_t* b, int64_t l, int64_t nb, const double* pos,
                      const int64_t* cell_of, double k, double cutoff, double* frep) {
  for (int64_t i = b->head[l]; i >= 0; i = b->next[i])
    for (int64_t j = b->head[nb]; j >= 0; j = b->next[j]) {
      if (i == j) continue;
      if (cell_of[i] == cell_of[j]) continue;
      double dv[3]; v3sub(pos + 3*i, pos + 3*j, dv); min_image(d, dv);
      const double dist = sqrt(dv[0]*dv[0] + dv[1]*dv[1] + dv[2]*dv[2]);
      if (dist < cutoff) {
        const double s = k * (1/(dist/cutoff));
        for (int c = 0; c < 3; c++) {
          const double r = s * (dv[c]/dist);
          frep[3*i + c] = frep[3*i + c] + r;
          frep[3*j + c] = frep[3*j + c] - r;
        }
      }
    }
}
/* applyRepulsionForce (hemoCellParticleField.cpp:677-743): origin bin + 13 forward
 * neighbours; the origin bin paired with itself visits (i,j) and (j,i) => doubled force for
 * same-node pairs (preserved). */
void ora_repulsion(const ora_domain* d, int64_t np, const double* pos, const int64_t* cell_of,
                   double k, double cutoff, double* frep) {
  memset(frep, 0, sizeof(double)*3*np);
  bins_t b = build_bins(d, np, pos);
  static const int ST[14][3] = {{0,0,0},{0,0,1},{0,1,0},{0,1,1},
    {1,-1,-1},{1,-1,0},{1,-1,1},{1,0,-1},{1,0,0},{1,0,1},{1,1,-1},{1,1,0},{1,1,1},{0,1,-1}};
  for (int x = 0; x < d->nx; x++) for (int y = 0; y < d->ny; y++) for (int z = 0; z < d->nz; z++) {
    int64_t l = nidx(d, x, y, z);
    if (b.head[l] < 0) continue;
    for (int s = 0; s < 14; s++) {
      int xx = x + ST[s][0], yy = y + ST[s][1], zz = z + ST[s][2];
      if (!wrap1(&xx, d->nx, d->periodic[0])) continue;
      if (!wrap1(&yy, d->ny, d->periodic[1])) continue;
      if (!wrap1(&zz, d->nz, d->periodic[2])) continue;
      rep_pairs(d, &b, l, nidx(d, xx, yy, zz), pos, cell_of, k, cutoff, frep);
    }
  }
  free(b.head); free(b.next);
}

/* populateBoundaryParticles + applyBoundaryRepulsionForce (hemoCellParticleField.cpp:865-918):
 * every boundary node with a non-boundary node in its 3^3 neighbourhood pushes the LSPs binned
 * within +-1 node; one-sided; accumulates onto force_repulsion WITHOUT zeroing (Appendix D.5) */
void ora_wall_repulsion(const ora_domain* d, const uint8_t* flags, int64_t np, const double* pos,
                        double k, double cutoff, double* frep) {
  bins_t b = build_bins(d, np, pos);
  for (int x = 0; x < d->nx; x++) for (int y = 0; y < d->ny; y++) for (int z = 0; z < d->nz; z++) {
    if (flags[nidx(d, x, y, z)] == ORA_FLUID) continue;
    int touches = 0;
    for (int dx = -1; dx <= 1 && !touches; dx++) for (int dy = -1; dy <= 1 && !touches; dy++)
      for (int dz = -1; dz <= 1 && !touches; dz++) {
        int xx = x + dx, yy = y + dy, zz = z + dz;
        if (!wrap1(&xx, d->nx, d->periodic[0]) || !wrap1(&yy, d->ny, d->periodic[1]) ||
            !wrap1(&zz, d->nz, d->periodic[2])) continue;
        if (flags[nidx(d, xx, yy, zz)] == ORA_FLUID) touches = 1;
      }
    if (!touches) continue;
    const double wp[3] = {(double)x, (double)y, (double)z};
    for (int dx = -1; dx <= 1; dx++) for (int dy = -1; dy <= 1; dy++) for (int dz = -1; dz <= 1; dz++) {
      int xx = x + dx, yy = y + dy, zz = z + dz;
      if (!wrap1(&xx, d->nx, d->periodic[0]) || !wrap1(&yy, d->ny, d->periodic[1]) ||
          !wrap1(&zz, d->nz, d->periodic[2])) continue;
      for (int64_t i = b.head[nidx(d, xx, yy, zz)]; i >= 0; i = b.next[i]) {
        double dv[3]; v3sub(pos + 3*i, wp, dv); min_image(d, dv);
        const double dist = sqrt(dv[0]*dv[0] + dv[1]*dv[1] + dv[2]*dv[2]);
        if (dist < cutoff) {
          const double s = k * (1/(dist/cutoff));
          for (int c = 0; c < 3; c++) frep[3*i + c] = frep[3*i + c] + s*(dv[c]/dist);
        }
      }
    }
  }
  free(b.head); free(b.next);
}
