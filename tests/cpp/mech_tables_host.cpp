// Host build of the RBC mechanics kernel's packed ring table (hemocell_b200/csrc/mech_tables.h) for
// tests/test_mechanics_ring_algorithm.py.
#include <cstring>
#include "../../hemocell_b200/csrc/mech_tables.h"

extern "C" int mech_ring_table(int V, int T, int E, const int* triangles, const int* edges, const int* vertex_vertexes,
                               const int* vertex_n_vertexes, unsigned long long* out, char* err, int errlen) {
  std::vector<unsigned long long> rg;
  const char* why = mech_tables::build_ring_table(V, T, E, triangles, edges, vertex_vertexes, vertex_n_vertexes, rg);
  if (why) { if (err && errlen > 0) { std::strncpy(err, why, errlen - 1); err[errlen - 1] = 0; } return 1; }
  std::memcpy(out, rg.data(), sizeof(unsigned long long)*rg.size());
  return 0;
}
