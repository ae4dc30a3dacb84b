#include "hemo_voxel.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <tuple>

namespace hemo { namespace host {
namespace {
typedef std::array<double, 3> P3;
typedef std::array<P3, 3> Tri;

std::vector<Tri> readSTL(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("(Voxelizer) Error: " + path + " is not an existing stl file.");
  std::string all((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  std::vector<Tri> tris;
  // binary STL: 80-byte header, uint32 count, 50 bytes per facet
  if (all.size() >= 84) {
    uint32_t n; memcpy(&n, all.data() + 80, 4);
    if (all.size() == 84 + (size_t)n*50 && all.compare(0, 5, "solid") != 0) {
      for (uint32_t i = 0; i < n; i++) {
        float v[12]; memcpy(v, all.data() + 84 + (size_t)i*50, 48);
        Tri t; for (int k = 0; k < 3; k++) for (int d = 0; d < 3; d++) t[k][d] = v[3 + 3*k + d];
        tris.push_back(t);
      }
      return tris;
    }
  }
  std::istringstream in(all);
  std::string w; Tri t; int k = 0;
  while (in >> w) {
    if (w == "vertex") {
      if (!(in >> t[k][0] >> t[k][1] >> t[k][2])) throw std::runtime_error("(Voxelizer) malformed STL " + path);
      if (++k == 3) { tris.push_back(t); k = 0; }
    }
  }
  if (tris.empty()) throw std::runtime_error("(Voxelizer) no triangles in " + path);
  return tris;
}
}  // namespace

VoxelizedSTL voxelizeSTL(const std::string& path, int refDirN, int refDir, int margin, bool openXEnds) {
  if (refDirN < 1 || refDir < 0 || refDir > 2) throw std::runtime_error("(Voxelizer) bad refDirN / refDir");
  std::vector<Tri> tris = readSTL(path);
  // ---- indexed mesh (vertices merged), to lattice units
  std::vector<P3> vert; std::vector<std::array<int, 3>> tri;
  {
    std::map<std::tuple<long long, long long, long long>, int> key;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (auto& t : tris) for (auto& p : t) for (int d = 0; d < 3; d++) { lo[d] = std::min(lo[d], p[d]); hi[d] = std::max(hi[d], p[d]); }
    const double scale = std::max({hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]});
    for (auto& t : tris) {
      std::array<int, 3> id;
      for (int k = 0; k < 3; k++) {
        auto kk = std::make_tuple(std::llround((t[k][0] - lo[0])/scale*1e9), std::llround((t[k][1] - lo[1])/scale*1e9), std::llround((t[k][2] - lo[2])/scale*1e9));
        auto it = key.find(kk);
        if (it == key.end()) { it = key.emplace(kk, (int)vert.size()).first; vert.push_back(t[k]); }
        id[k] = it->second;
      }
      if (id[0] != id[1] && id[1] != id[2] && id[0] != id[2]) tri.push_back(id);
    }
    VoxelizedSTL dummy; (void)dummy;
  }
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (auto& p : vert) for (int d = 0; d < 3; d++) { lo[d] = std::min(lo[d], p[d]); hi[d] = std::max(hi[d], p[d]); }
  VoxelizedSTL out;
  out.dx = (hi[refDir] - lo[refDir])/(double)refDirN;
  for (int d = 0; d < 3; d++) out.location[d] = lo[d] - margin*out.dx;
  for (auto& p : vert) for (int d = 0; d < 3; d++) p[d] = (p[d] - lo[d])/out.dx + margin;
  int n[3];
  for (int d = 0; d < 3; d++) n[d] = (int)((hi[d] - lo[d])/out.dx) + 1 + 2*margin;
  out.nx = n[0]; out.ny = n[1]; out.nz = n[2];
  // ---- inflate: 1e-3 lu along the normalised sum of the adjacent unit normals (outward for a consistently oriented STL)
  {
    std::vector<P3> vn(vert.size(), P3{0, 0, 0});
    double signedVol = 0;
    for (auto& t : tri) {
      const P3 &a = vert[t[0]], &b = vert[t[1]], &c = vert[t[2]];
      const P3 u = {b[0]-a[0], b[1]-a[1], b[2]-a[2]}, v = {c[0]-a[0], c[1]-a[1], c[2]-a[2]};
      P3 nn = {u[1]*v[2]-u[2]*v[1], u[2]*v[0]-u[0]*v[2], u[0]*v[1]-u[1]*v[0]};
      signedVol += a[0]*(b[1]*c[2]-b[2]*c[1]) + a[1]*(b[2]*c[0]-b[0]*c[2]) + a[2]*(b[0]*c[1]-b[1]*c[0]);
      const double l = std::sqrt(nn[0]*nn[0] + nn[1]*nn[1] + nn[2]*nn[2]);
      if (l == 0) continue;
      for (int k = 0; k < 3; k++) for (int d = 0; d < 3; d++) vn[t[k]][d] += nn[d]/l;
    }
    const double sgn = signedVol >= 0 ? 1.0 : -1.0;       // normals of an inward-oriented file point the other way
    for (size_t i = 0; i < vert.size(); i++) {
      const double l = std::sqrt(vn[i][0]*vn[i][0] + vn[i][1]*vn[i][1] + vn[i][2]*vn[i][2]);
      if (l > 0) for (int d = 0; d < 3; d++) vert[i][d] += sgn*1.0e-3*vn[i][d]/l;
    }
  }
  // ---- inside test: for every (y, z) column the crossings of the +x ray with the surface, parity fill
  out.flag.assign((size_t)out.nx*out.ny*out.nz, 0);
  std::vector<std::vector<double>> hits((size_t)out.ny*out.nz);
  for (auto& t : tri) {
    const P3 &a = vert[t[0]], &b = vert[t[1]], &c = vert[t[2]];
    const int y0 = std::max(0, (int)std::ceil(std::min({a[1], b[1], c[1]}))), y1 = std::min(out.ny - 1, (int)std::floor(std::max({a[1], b[1], c[1]})));
    const int z0 = std::max(0, (int)std::ceil(std::min({a[2], b[2], c[2]}))), z1 = std::min(out.nz - 1, (int)std::floor(std::max({a[2], b[2], c[2]})));
    // projected (y, z) triangle; orientation-normalised edge functions with a half-open rule so that a ray through a
    // shared edge counts exactly one of the two triangles
    const double area = (b[1]-a[1])*(c[2]-a[2]) - (b[2]-a[2])*(c[1]-a[1]);
    if (area == 0) continue;
    for (int y = y0; y <= y1; y++) for (int z = z0; z <= z1; z++) {
      auto edge = [&](const P3& p, const P3& q) {
        double e = (q[1]-p[1])*((double)z-p[2]) - (q[2]-p[2])*((double)y-p[1]);
        if (area < 0) e = -e;
        if (e != 0) return e > 0;
        const double dy = (area < 0 ? -1 : 1)*(q[1]-p[1]), dz = (area < 0 ? -1 : 1)*(q[2]-p[2]);
        return dz < 0 || (dz == 0 && dy > 0);                  // top-left rule
      };
      if (!(edge(a, b) && edge(b, c) && edge(c, a))) continue;
      const double w0 = ((b[1]-y)*(c[2]-z) - (b[2]-z)*(c[1]-y))/area, w1 = ((c[1]-y)*(a[2]-z) - (c[2]-z)*(a[1]-y))/area;
      const double w2 = 1.0 - w0 - w1;
      hits[(size_t)y*out.nz + z].push_back(w0*a[0] + w1*b[0] + w2*c[0]);
    }
  }
  for (int y = 0; y < out.ny; y++) for (int z = 0; z < out.nz; z++) {
    auto& h = hits[(size_t)y*out.nz + z];
    if (h.empty()) continue;
    if (h.size() % 2) throw std::runtime_error("(Voxelizer) the STL surface is not closed (odd number of crossings)");
    std::sort(h.begin(), h.end());
    for (size_t k = 0; k + 1 < h.size(); k += 2)
      for (int x = std::max(0, (int)std::ceil(h[k])); x <= std::min(out.nx - 1, (int)std::floor(h[k+1])); x++)
        out.flag[(size_t)z + (size_t)out.nz*((size_t)y + (size_t)out.ny*x)] = 1;
  }
  // ---- helper/voxelizeDomain.cpp:142-156: open the two x ends (CopyFromNeighbor, applied in ascending x)
  if (openXEnds && out.nx >= 4) {
    auto at = [&](int x, int y, int z) -> int32_t& { return out.flag[(size_t)z + (size_t)out.nz*((size_t)y + (size_t)out.ny*x)]; };
    for (int x = 0; x <= 1; x++) for (int y = 0; y < out.ny; y++) for (int z = 0; z < out.nz; z++) at(x, y, z) = at(x + 1, y, z);
    for (int x = out.nx - 2; x <= out.nx - 1; x++) for (int y = 0; y < out.ny; y++) for (int z = 0; z < out.nz; z++) at(x, y, z) = at(x - 1, y, z);
  }
  return out;
}

}}  // namespace hemo::host
