// Packed ring table of the RBC mechanics kernel (mechanics.cu, step 4): host code, used by hcg_celltype_add and compiled on
// its own by tests/cpp/mech_tables_host.cpp (tests/test_mechanics_ring_algorithm.py).
//
// One 64-bit word per (ring slot j, vertex v), laid out [6][V]:
//   bits  0-15  ring vertex r_j                      (reference mesh: vertex_vertexes, cyclic order)
//   bits 16-31  edge (v, r_j)
//   bits 32-47  triangle (v, r_j, r_j+1)
//   bits 48-50  ring size of r_j                      (the share -1/n of r_j's bending force that v takes)
//   bits 51-52  which of (v, r_j, r_j+1) is the triangle's THIRD vertex: the centroid is summed (v0 + v1) + v2 as the
//               reference does (mechanics/rbcHighOrderModel.cpp:72-92), and that sum cancels five digits
//   bit  53     (v, r_j, r_j+1) runs against the triangle's orientation (sign of the cross product)
// Slots past the ring size repeat r_0 with the null edge E, the null triangle T and ring size 0: they add +0.0.
#pragma once
#include <algorithm>
#include <unordered_map>
#include <utility>
#include <vector>

namespace mech_tables {

constexpr int kMaxIndex = 65534;    // V, E, T must fit 16 bits with the null elements E and T

inline const char* build_ring_table(int V, int T, int E, const int* triangles, const int* edges, const int* vertex_vertexes,
                                    const int* vertex_n_vertexes, std::vector<unsigned long long>& rg) {
  typedef unsigned long long u64;
  if (V > kMaxIndex || E > kMaxIndex || T > kMaxIndex) return "mesh too large for the packed gather tables";
  rg.assign(6*(size_t)V, 0ull);
  std::unordered_map<u64, int> emap, tmap;
  auto ekey = [](int x, int y) { if (x > y) std::swap(x, y); return ((u64)x << 20) | (u64)y; };
  auto tkey = [](int x, int y, int z) { int q[3] = {x, y, z}; std::sort(q, q + 3); return ((u64)q[0] << 40) | ((u64)q[1] << 20) | (u64)q[2]; };
  for (int e = 0; e < E; e++) emap[ekey(edges[2*e], edges[2*e+1])] = e;
  for (int k = 0; k < T; k++) tmap[tkey(triangles[3*k], triangles[3*k+1], triangles[3*k+2])] = k;
  for (int v = 0; v < V; v++) {
    const int nn = vertex_n_vertexes[v];
    if (nn < 3 || nn > 6) return "vertex ring size must be 3..6";
    const int* ring = vertex_vertexes + 6*v;
    for (int j = 0; j < nn; j++) if (ring[j] < 0 || ring[j] >= V || ring[j] == v) return "ring vertex out of range";
    for (int j = 0; j < 6; j++) {
      u64 w;
      if (j < nn) {
        const int ia = ring[j], ib = ring[(j + 1) % nn];
        auto ei = emap.find(ekey(v, ia));
        auto ti = tmap.find(tkey(v, ia, ib));
        if (ei == emap.end()) return "ring neighbour without an edge";
        if (ti == tmap.end()) return "consecutive ring neighbours without a triangle";
        const int* q = triangles + 3*ti->second;
        const int last = q[2] == v ? 0 : (q[2] == ia ? 1 : 2);
        const bool same = (q[0] == v && q[1] == ia) || (q[0] == ia && q[1] == ib) || (q[0] == ib && q[1] == v);
        const int rn = vertex_n_vertexes[ia];
        if (rn < 3 || rn > 6) return "vertex ring size must be 3..6";
        w = (u64)ia | ((u64)ei->second << 16) | ((u64)ti->second << 32) | ((u64)rn << 48) | ((u64)last << 51) | ((u64)(same ? 0 : 1) << 53);
      } else {
        w = (u64)ring[0] | ((u64)E << 16) | ((u64)T << 32);
      }
      rg[(size_t)j*V + v] = w;
    }
  }
  return nullptr;
}

}  // namespace mech_tables
