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

// the velocity-state variant: Vin / Vout hold (rhoBar, j / rho)
extern "C" void moment_host_vel(int nx, int ny, int nz, const double* Vin, const double* Fin, double* Vout, double* Fout, double* U,
                                const double* body, int write_u) {
  MomentArgs a; a.ny = ny; a.nz = nz; a.P = (int64_t)ny*nz; a.body[0] = body[0]; a.body[1] = body[1]; a.body[2] = body[2];
  const int64_t Nl = (int64_t)nx*ny*nz;
  for (int64_t i = 0; i < Nl; i++) {
    if (write_u) moment_node_vel<true>(Vin, Fin, Vout, Fout, U, a, i);
    else moment_node_vel<false>(Vin, Fin, Vout, Fout, U, a, i);
  }
}
