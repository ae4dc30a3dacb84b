// Host-side harness of the moment-only update: compiles csrc/moment_step.cuh for the CPU so that its indexing and arithmetic can be
// checked without a GPU (tests/test_moment_only_algorithm.py builds this with nvcc and calls it through ctypes).
#include "../../hemocell_b200/csrc/moment_step.cuh"

extern "C" void moment_host(int nx, int ny, int nz, const double* Win, const double* Fin, double* Wout, double* Fout, double* U,
                            const double* body, int write_u) {
  MomentArgs a; a.ny = ny; a.nz = nz; a.P = (int64_t)ny*nz; a.body[0] = body[0]; a.body[1] = body[1]; a.body[2] = body[2];
  const int64_t Nl = (int64_t)nx*ny*nz;
  for (int64_t i = 0; i < Nl; i++) {
    if (write_u) moment_node<true>(Win, Fin, Wout, Fout, U, a, i);
    else moment_node<false>(Win, Fin, Wout, Fout, U, a, i);
  }
}

// the re-associated population evaluation of k_moment_tile next to guo_collide_tau1 (the arithmetic the oracle is checked against)
#include "../../hemocell_b200/csrc/lattice_node.cuh"
extern "C" void tau1_pops_both(int n, const double* w4, const double* f3, double* fast19, double* ref19) {
  for (int i = 0; i < n; i++) {
    tau1_pops_fast(w4[4*i], w4[4*i+1], w4[4*i+2], w4[4*i+3], f3[3*i], f3[3*i+1], f3[3*i+2], fast19 + 19*i);
    const double j[3] = {w4[4*i+1], w4[4*i+2], w4[4*i+3]};
    guo_collide_tau1(ref19 + 19*i, w4[4*i], j, f3 + 3*i);
  }
}
