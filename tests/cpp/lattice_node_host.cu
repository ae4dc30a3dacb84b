// Host-side harness of the lattice kernels' per-node arithmetic: compiles hemocell_b200/csrc/lattice_node.cuh (host + device code,
// inlined by k_collide_stream / k_collide_tau1 / k_moments on the device) for the CPU and applies it node by node in the order
// the kernels do, so that the CPU test suite checks it against the oracle without a GPU (tests/test_lattice_node_host.py).
#include <stdint.h>
#include "../../include/hemocell_gpu.h"
#include "../../hemocell_b200/csrc/lattice_node.cuh"

// pop: post-stream populations [19][N]; out: post-collision populations [19][N] (before streaming)
extern "C" void node_collide_host(int64_t N, const uint8_t* flags, const double* pop, const double* force, double omega,
                                  const double* bc_vel /*[6][3]*/, const double* bc_node /*[4][N] or NULL*/, int tau1, double* out) {
  for (int64_t n = 0; n < N; n++) {
    double f[19];
    for (int q = 0; q < 19; q++) f[q] = pop[(int64_t)q*N + n];
    const uint8_t fl = flags[n];
    if (fl == HCG_BOUNCEBACK) {
      for (int q = 1; q <= 9; q++) { const double t = f[q]; f[q] = f[q+9]; f[q+9] = t; }
    } else {
      const double Fn[3] = {force[n], force[N + n], force[2*N + n]};
      if (fl >= HCG_ZH_VEL_XN) {
        const double b[4] = {bc_node ? bc_node[n] : 0.0, bc_node ? bc_node[N + n] : 0.0, bc_node ? bc_node[2*N + n] : 0.0, bc_node ? bc_node[3*N + n] : 1.0};
        if (fl >= HCG_ZH_PRES_XN) zouhe_complete(f, fl - HCG_ZH_PRES_XN, true, b[0], b[1], b[2], b[3]);
        else zouhe_complete(f, fl - HCG_ZH_VEL_XN, false, b[0], b[1], b[2], b[3]);
      } else if (fl >= HCG_VEL_XN) {
        const double uw[3] = {bc_vel[3*(fl-2)], bc_vel[3*(fl-2)+1], bc_vel[3*(fl-2)+2]};
        regularized_complete(f, fl - 2, uw);
      }
      if (tau1 && fl == HCG_FLUID) {                     // k_collide_tau1: the moments come from the W field = moments19 of the same populations
        double rb, j[3];
        moments19(f, rb, j);
        guo_collide_tau1(f, rb, j, Fn);
      } else guo_collide(f, Fn, omega);
    }
    for (int q = 0; q < 19; q++) out[(int64_t)q*N + n] = f[q];
  }
}
