// See hemo_mesh.h.  Palabos pieces that are not part of the reference tree (TriangleSet::rotate,
// DEFscaledMesh vertex numbering, constructSphere, TriangularSurfaceMesh::inflate) are restated
// from their published behaviour; the PLT <InnerEdges> vertex pairs of examples/pipeflow/PLT.xml
// come out as exact mirror pairs under this numbering, which pins it.
#include "hemo_mesh.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <map>
#include <stdexcept>
#include <tuple>

namespace hemo {
namespace host {

namespace {
const T PI = 3.14159265358979323846;
typedef std::array<Vec3, 3> Tri;

inline Vec3 add(const Vec3& a, const Vec3& b) { return {a[0]+b[0], a[1]+b[1], a[2]+b[2]}; }
inline Vec3 sub(const Vec3& a, const Vec3& b) { return {a[0]-b[0], a[1]-b[1], a[2]-b[2]}; }
inline Vec3 mul(T s, const Vec3& a) { return {s*a[0], s*a[1], s*a[2]}; }
inline T dot(const Vec3& a, const Vec3& b) { return a[0]*b[0] + a[1]*b[1] + a[2]*b[2]; }
inline T norm(const Vec3& a) { return std::sqrt(dot(a, a)); }
inline Vec3 cross(const Vec3& a, const Vec3& b) {
  return {a[1]*b[2] - a[2]*b[1], a[2]*b[0] - a[0]*b[2], a[0]*b[1] - a[1]*b[0]};
}
inline Vec3 unit(const Vec3& a) { T n = norm(a); return {a[0]/n, a[1]/n, a[2]/n}; }

// refinement loop shared by both sphere generators (helper/meshGeneratingFunctions.hh:108-141)
void refine(std::vector<Tri>& tris, int minNumOfTriangles) {
  int size;
  while ((size = (int)tris.size()) < minNumOfTriangles) {
    for (int i = 0; i < size; i++) {
      Vec3 va = tris[i][0], vb = tris[i][1], vc = tris[i][2];
      Vec3 vd = unit(mul(0.5, add(va, vb)));
      Vec3 ve = unit(mul(0.5, add(vb, vc)));
      Vec3 vf = unit(mul(0.5, add(vc, va)));
      tris[i] = {vd, ve, vf};
      tris.push_back({va, vd, vf});
      tris.push_back({vd, vb, ve});
      tris.push_back({vf, ve, vc});
    }
  }
}

// constructSphereIcosahedron (helper/meshGeneratingFunctions.hh:31-151), unit radius at origin
std::vector<Tri> constructSphereIcosahedron(int minNumOfTriangles) {
  const T tau = -0.8506508084, one = -0.5257311121;
  const Vec3 v[13] = {{0,0,0}, {tau, one, 0}, {-tau, one, 0}, {-tau, -one, 0}, {tau, -one, 0},
                      {one, 0, tau}, {one, 0, -tau}, {-one, 0, -tau}, {-one, 0, tau},
                      {0, tau, one}, {0, -tau, one}, {0, -tau, -one}, {0, tau, -one}};
  const int f[20][3] = {{5,8,9},{5,10,8},{6,12,7},{6,7,11},{1,4,5},{1,6,4},{3,2,8},{3,7,2},{9,12,1},{9,2,12},
                        {10,4,11},{10,11,3},{9,1,5},{12,6,1},{5,4,10},{6,11,4},{8,2,9},{7,12,2},{8,10,3},{7,3,11}};
  std::vector<Tri> tris;
  for (auto& t : f) tris.push_back({v[t[0]], v[t[1]], v[t[2]]});
  refine(tris, minNumOfTriangles);
  return tris;
}

// Palabos constructSphere: octahedron refined onto the unit sphere
std::vector<Tri> constructSphere(int minNumOfTriangles) {
  const Vec3 va = {1,0,0}, vb = {0,1,0}, vc = {-1,0,0}, vd = {0,-1,0}, ve = {0,0,1}, vf = {0,0,-1};
  std::vector<Tri> tris = {{ve,va,vb},{ve,vb,vc},{ve,vc,vd},{ve,vd,va},{vf,vb,va},{vf,vc,vb},{vf,vd,vc},{vf,va,vd}};
  refine(tris, minNumOfTriangles);
  return tris;
}

// Palabos TriangleSet::rotate(phi, theta, psi): z-x-z Euler angles
void rotate(std::vector<Tri>& tris, T phi, T theta, T psi) {
  T a[3][3];
  a[0][0] =  std::cos(psi)*std::cos(phi) - std::cos(theta)*std::sin(phi)*std::sin(psi);
  a[0][1] =  std::cos(psi)*std::sin(phi) + std::cos(theta)*std::cos(phi)*std::sin(psi);
  a[0][2] =  std::sin(psi)*std::sin(theta);
  a[1][0] = -std::sin(psi)*std::cos(phi) - std::cos(theta)*std::sin(phi)*std::cos(psi);
  a[1][1] = -std::sin(psi)*std::sin(phi) + std::cos(theta)*std::cos(phi)*std::cos(psi);
  a[1][2] =  std::cos(psi)*std::sin(theta);
  a[2][0] =  std::sin(theta)*std::sin(phi);
  a[2][1] = -std::sin(theta)*std::cos(phi);
  a[2][2] =  std::cos(theta);
  for (auto& t : tris) for (auto& p : t) {
    Vec3 x = p;
    for (int i = 0; i < 3; i++) p[i] = a[i][0]*x[0] + a[i][1]*x[1] + a[i][2]*x[2];
  }
}

// spherePointToRBCPoint (helper/meshGeneratingFunctions.hh:153-168)
Vec3 spherePointToRBCPoint(const Vec3& point, T R = 1.0) {
  Vec3 p = point;
  T r2 = p[0]*p[0] + p[1]*p[1];
  const T val = p[2];
  const int sign = (T(0) < val) - (val < T(0));
  p[0] *= R; p[1] *= R;
  if (1 - r2 < 0) r2 = 1;
  const T C0 = 0.054322, C2 = 1.001279, C4 = -0.561381;
  p[2] = sign * R * std::sqrt(1 - r2) * (C0 + C2*r2 + C4*r2*r2);
  return p;
}
// spherePointToEllipsoidPoint (:170-183)
Vec3 spherePointToEllipsoidPoint(const Vec3& point, T R, T aspectRatio) {
  Vec3 p = point;
  T r2 = p[0]*p[0] + p[1]*p[1];
  const T val = p[2];
  const int sign = (T(0) < val) - (val < T(0));
  if (1 - r2 < 0) r2 = 1;
  p[0] *= R; p[1] *= R;
  p[2] = sign * aspectRatio * R * std::sqrt(1 - r2);
  return p;
}

// DEFscaledMesh + TriangleBoundary3D: vertices numbered by first appearance in the triangle scan
TriangularSurfaceMesh indexMesh(const std::vector<Tri>& tris) {
  TriangularSurfaceMesh m;
  std::map<std::tuple<long long, long long, long long>, int> key;
  for (auto& t : tris) {
    std::array<int, 3> id;
    for (int k = 0; k < 3; k++) {
      auto kk = std::make_tuple(std::llround(t[k][0]*1e9), std::llround(t[k][1]*1e9), std::llround(t[k][2]*1e9));
      auto it = key.find(kk);
      if (it == key.end()) { it = key.emplace(kk, (int)m.vertices.size()).first; m.vertices.push_back(t[k]); }
      id[k] = it->second;
    }
    m.triangles.push_back(id);
  }
  return m;
}

// TriangularSurfaceMesh::inflate(): every vertex moves 1e-3 lu along the normalised sum of the
// adjacent unit triangle normals (amount chosen to meet scripts/ci/stretchCell_sanity.sh)
void inflate(TriangularSurfaceMesh& m, T amount = 1.0e-3) {
  std::vector<Vec3> vn(m.vertices.size(), Vec3{0, 0, 0});
  for (auto& t : m.triangles) {
    Vec3 n = unit(cross(sub(m.vertices[t[1]], m.vertices[t[0]]), sub(m.vertices[t[2]], m.vertices[t[0]])));
    for (int k = 0; k < 3; k++) vn[t[k]] = add(vn[t[k]], n);
  }
  for (size_t i = 0; i < m.vertices.size(); i++) m.vertices[i] = add(m.vertices[i], mul(amount, unit(vn[i])));
}

// rotateTriangularMeshXYZ (io/readPositionsBloodCells.cpp:39-98): a = Rz * (Ry * Rx)
void rotationXYZ(T alpha, T beta, T gamma, T a[3][3]) {
  T b[3][3], c[3][3];
  a[0][0] = 1; a[0][1] = 0; a[0][2] = 0;
  a[1][0] = 0; a[1][1] = std::cos(alpha); a[1][2] = std::sin(alpha);
  a[2][0] = 0; a[2][1] = -std::sin(alpha); a[2][2] = std::cos(alpha);
  b[0][0] = std::cos(beta); b[0][1] = 0; b[0][2] = -std::sin(beta);
  b[1][0] = 0; b[1][1] = 1; b[1][2] = 0;
  b[2][0] = std::sin(beta); b[2][1] = 0; b[2][2] = std::cos(beta);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { c[i][j] = 0; for (int k = 0; k < 3; k++) c[i][j] += a[k][j]*b[i][k]; }
  b[0][0] = std::cos(gamma); b[0][1] = std::sin(gamma); b[0][2] = 0;
  b[1][0] = -std::sin(gamma); b[1][1] = std::cos(gamma); b[1][2] = 0;
  b[2][0] = 0; b[2][1] = 0; b[2][2] = 1;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { a[i][j] = 0; for (int k = 0; k < 3; k++) a[i][j] += c[k][j]*b[i][k]; }
}
}  // namespace

T TriangularSurfaceMesh::getVolume() const {
  T v = 0;
  for (auto& t : triangles) v += dot(vertices[t[0]], cross(vertices[t[1]], vertices[t[2]]));
  return v/6.0;
}
T TriangularSurfaceMesh::getSurface() const {
  T s = 0;
  for (auto& t : triangles) s += 0.5*norm(cross(sub(vertices[t[1]], vertices[t[0]]), sub(vertices[t[2]], vertices[t[0]])));
  return s;
}
void TriangularSurfaceMesh::boundingBox(Vec3& lo, Vec3& hi) const {
  lo = hi = vertices[0];
  for (auto& v : vertices) for (int d = 0; d < 3; d++) { lo[d] = std::min(lo[d], v[d]); hi[d] = std::max(hi[d], v[d]); }
}

TriangularSurfaceMesh constructMeshElement(int shape, T radius, int cellNumTriangles, T aspectRatio) {
  std::vector<Tri> tris;
  if (shape == RBC_FROM_SPHERE) {                      // constructRBCFromSphere(.., initialSphereShape = 1), .hh:213-241
    tris = constructSphereIcosahedron(cellNumTriangles);
    rotate(tris, PI/2.0, PI/2.0, 0.0);
    for (auto& t : tris) for (auto& p : t) p = spherePointToRBCPoint(p);
    for (auto& t : tris) for (auto& p : t) p = mul(radius, p);
    rotate(tris, PI/2.0, PI/2.0, 0.0);
  } else if (shape == ELLIPSOID_FROM_SPHERE) {         // constructEllipsoidFromSphere(.., 0), .hh:244-271
    tris = constructSphere(cellNumTriangles);
    rotate(tris, PI/2.0, PI/2.0, 0.0);
    for (auto& t : tris) for (auto& p : t) p = spherePointToEllipsoidPoint(p, radius, aspectRatio);
    rotate(tris, PI/2.0, PI/2.0, 0.0);
  } else {
    throw std::invalid_argument("constructMeshElement: only RBC_FROM_SPHERE and ELLIPSOID_FROM_SPHERE are supported");
  }
  TriangularSurfaceMesh m = indexMesh(tris);
  inflate(m);
  return m;
}

void Parameters::lbm_base_parameters(T dx_, T dt_, T nu_p_, T rho_p_, T kBT_p_) {
  dt = dt_; dx = dx_; nu_p = nu_p_; rho_p = rho_p_; kBT_p = kBT_p_;
  if (dt < 0.0) {
    tau = 1.0;
    nu_lbm = 1.0/3.0 * (tau - 0.5);
    dt = nu_lbm / nu_p * (dx*dx);
  } else {
    nu_lbm = nu_p * dt / (dx*dx);
    tau = 3.0 * nu_lbm + 0.5;
  }
  dm = rho_p * (dx*dx*dx);
  df = dm * dx / (dt*dt);
  f_limit = 50.0 / 1.0e12 / df;          // FORCE_LIMIT = 50 pN (config/constant_defaults.h:73-75)
  kBT_lbm = kBT_p/(df*dx);
}
void Parameters::lbm_pipe_parameters(T Re, int nY) { re = Re; pipe_radius = nY; u_lbm_max = re * nu_lbm / (pipe_radius*2); }
void Parameters::lbm_shear_parameters(T shearrate_p, T nx) {
  re = (nx * (shearrate_p * (nx*0.5))) / nu_p;
  shearrate_lbm = shearrate_p*dt;
  u_lbm_max = shearrate_lbm;
}

Stiffness calculate_stiffness(const MaterialModel& m, const Parameters& p, int nTriangles) {
  Stiffness k;
  const T plc = 7.5e-9/p.dx;
  const T eqLength = 5e-7/p.dx;
  const T NfacesScaling = 1280.0/nTriangles;
  k.k_link = m.kLink * p.kBT_lbm/plc;
  k.k_bend = m.kBend * p.kBT_lbm / eqLength;
  k.k_volume = m.kVolume * NfacesScaling * p.kBT_lbm / eqLength;
  k.k_area = m.kArea * NfacesScaling * p.kBT_lbm/(eqLength);
  k.eta_m = m.eta_m * p.dx / p.dt / p.df;
  return k;
}

CommonCellConstants CommonCellConstants::CommonCellConstantsConstructor(
    const TriangularSurfaceMesh& mesh, const std::vector<std::array<int, 2>>& innerEdges) {
  CommonCellConstants cc;
  const int V = mesh.getNumVertices(), Tn = mesh.getNumTriangles();
  cc.triangle_list = mesh.triangles;
  for (auto& t : cc.triangle_list) {                       // commonCellConstants.cpp:81-93
    if (t[0] < t[1]) cc.edge_list.push_back({t[0], t[1]});
    if (t[1] < t[2]) cc.edge_list.push_back({t[1], t[2]});
    if (t[2] < t[0]) cc.edge_list.push_back({t[2], t[0]});
  }
  const int E = (int)cc.edge_list.size();
  std::map<std::pair<int, int>, int> directed;             // (a,b) -> triangle holding a->b
  for (int k = 0; k < Tn; k++) {
    auto& t = cc.triangle_list[k];
    directed[{t[0], t[1]}] = k; directed[{t[1], t[2]}] = k; directed[{t[2], t[0]}] = k;
  }
  std::vector<Vec3> normal(Tn);
  for (int k = 0; k < Tn; k++) {
    auto& t = cc.triangle_list[k];
    Vec3 n = cross(sub(mesh.vertices[t[1]], mesh.vertices[t[0]]), sub(mesh.vertices[t[2]], mesh.vertices[t[0]]));
    const T nn = norm(n);
    cc.triangle_area_eq_list.push_back(0.5*nn);
    normal[k] = {n[0]/nn, n[1]/nn, n[2]/nn};
  }
  for (auto& e : cc.edge_list) {
    const Vec3 ev = sub(mesh.vertices[e[1]], mesh.vertices[e[0]]);
    const T len = norm(ev);
    cc.edge_length_eq_list.push_back(len);
    // getAdjacentTriangleIds(e0, e1): first the triangle holding e1->e0, then the one holding
    // e0->e1 -- the order under which the dihedral force of pltSimpleModel.cpp:156-182 restores
    const int t0 = directed.at({e[1], e[0]}), t1 = directed.at({e[0], e[1]});
    cc.edge_bending_triangles_list.push_back({t0, t1});
    const Vec3 uv = {ev[0]/len, ev[1]/len, ev[2]/len};
    cc.edge_angle_eq_list.push_back(std::atan2(dot(cross(normal[t0], normal[t1]), uv), dot(normal[t0], normal[t1])));
    std::array<int, 2> op = {-1, -1};
    for (int i = 0; i < 3; i++) {
      if (cc.triangle_list[t0][i] != e[0] && cc.triangle_list[t0][i] != e[1]) op[0] = cc.triangle_list[t0][i];
      if (cc.triangle_list[t1][i] != e[0] && cc.triangle_list[t1][i] != e[1]) op[1] = cc.triangle_list[t1][i];
    }
    cc.edge_bending_triangles_outer_points.push_back(op);
  }
  cc.inner_edge_list = innerEdges;
  for (auto& e : innerEdges) cc.inner_edge_length_eq_list.push_back(norm(sub(mesh.vertices[e[1]], mesh.vertices[e[0]])));
  cc.volume_eq = mesh.getVolume();
  T s = 0; for (T a : cc.triangle_area_eq_list) s += a; cc.area_mean_eq = s / Tn;
  s = 0; for (T l : cc.edge_length_eq_list) s += l; cc.edge_mean_eq = s / E;
  s = 0; for (T a : cc.edge_angle_eq_list) s += a; cc.angle_mean_eq = s / E;
  // neighbours by first appearance in edge_list (:213-229), then ring order (:241-280)
  cc.vertex_vertexes.assign(V, {-1, -1, -1, -1, -1, -1});
  cc.vertex_n_vertexes.assign(V, 0);
  for (auto& e : cc.edge_list) for (int side = 0; side < 2; side++) {
    const int v = e[side], w = e[1 - side];
    if (cc.vertex_n_vertexes[v] >= 6) throw std::invalid_argument("vertex with more than 6 neighbours");
    cc.vertex_vertexes[v][cc.vertex_n_vertexes[v]++] = w;
  }
  for (int v = 0; v < V; v++) {
    int cur = cc.vertex_vertexes[v][0];
    for (int n = 1; n < cc.vertex_n_vertexes[v]; n++) {
      auto& t = cc.triangle_list[directed.at({v, cur})];
      int nxt = -1;
      for (int i = 0; i < 3; i++) if (t[i] != v && t[i] != cur) nxt = t[i];
      cur = nxt;
      cc.vertex_vertexes[v][n] = cur;
    }
  }
  for (int i = 0; i < V; i++) {                            // :283-314
    const int n = cc.vertex_n_vertexes[i];
    Vec3 sum = {0, 0, 0};
    for (int j = 0; j < n; j++) sum = add(sum, mesh.vertices[cc.vertex_vertexes[i][j]]);
    const Vec3 mid = {sum[0]/n, sum[1]/n, sum[2]/n};
    const Vec3 dev = sub(mid, mesh.vertices[i]);
    Vec3 pn = {0, 0, 0};
    for (int j = 0; j < n; j++) {
      Vec3 tn = cross(sub(mesh.vertices[cc.vertex_vertexes[i][j]], mesh.vertices[i]),
                      sub(mesh.vertices[cc.vertex_vertexes[i][(j + 1) % n]], mesh.vertices[i]));
      pn = add(pn, unit(tn));
    }
    pn = unit(pn);
    cc.surface_patch_center_dist_eq_list.push_back(dot(pn, dev));
  }
  return cc;
}

void CellTypeTables::build(int model_, int constructType, const MaterialModel& m, const Parameters& p) {
  model = model_;
  mesh = constructMeshElement(constructType, m.radius/p.dx, m.minNumTriangles, m.aspectRatio);
  cc = CommonCellConstants::CommonCellConstantsConstructor(mesh, m.innerEdges);
  k = calculate_stiffness(m, p, mesh.getNumTriangles());
  f_tri.clear(); f_edge.clear(); f_inner.clear(); f_vv.clear(); f_nvv.clear(); f_bt.clear(); f_bo.clear();
  for (auto& t : cc.triangle_list) for (int v : t) f_tri.push_back(v);
  for (auto& e : cc.edge_list) for (int v : e) f_edge.push_back(v);
  for (auto& e : cc.inner_edge_list) for (int v : e) f_inner.push_back(v);
  for (auto& r : cc.vertex_vertexes) for (int v : r) f_vv.push_back(v);
  for (int n : cc.vertex_n_vertexes) f_nvv.push_back(n);
  for (auto& e : cc.edge_bending_triangles_list) for (int v : e) f_bt.push_back(v);
  for (auto& e : cc.edge_bending_triangles_outer_points) for (int v : e) f_bo.push_back(v);
  c.model = model;
  c.n_vertices = mesh.getNumVertices(); c.n_triangles = mesh.getNumTriangles();
  c.n_edges = (int)cc.edge_list.size(); c.n_inner_edges = (int)cc.inner_edge_list.size();
  c.triangles = f_tri.data(); c.edges = f_edge.data(); c.inner_edges = f_inner.data();
  c.vertex_vertexes = f_vv.data(); c.vertex_n_vertexes = f_nvv.data();
  c.edge_bending_triangles = f_bt.data(); c.edge_bending_outer_points = f_bo.data();
  c.edge_length_eq = cc.edge_length_eq_list.data(); c.edge_angle_eq = cc.edge_angle_eq_list.data();
  c.triangle_area_eq = cc.triangle_area_eq_list.data(); c.patch_dist_eq = cc.surface_patch_center_dist_eq_list.data();
  c.inner_edge_length_eq = cc.inner_edge_length_eq_list.data();
  c.volume_eq = cc.volume_eq; c.area_mean_eq = cc.area_mean_eq; c.edge_mean_eq = cc.edge_mean_eq;
  c.k_volume = k.k_volume; c.k_area = k.k_area; c.k_link = k.k_link; c.k_bend = k.k_bend; c.eta_m = k.eta_m;
}

std::vector<std::array<T, 6>> readPositionsFile(const std::string& path) {
  std::ifstream in(path);
  if (!in.is_open()) throw std::invalid_argument("particle positions input file " + path + " does not exist");
  long n = 0; in >> n;
  std::vector<std::array<T, 6>> rows(n);
  for (long i = 0; i < n; i++) for (int k = 0; k < 6; k++) if (!(in >> rows[i][k])) throw std::invalid_argument("truncated .pos file " + path);
  return rows;
}

std::vector<int64_t> placeCells(const TriangularSurfaceMesh& mesh0, const std::vector<std::array<T, 6>>& rows,
                                T dx, int nx, int ny, int nz, const uint8_t* flags, T minDist_um,
                                int64_t cell_id0, std::vector<T>& out) {
  const int V = mesh0.getNumVertices();
  Vec3 lo, hi; mesh0.boundingBox(lo, hi);
  std::vector<Vec3> mesh(V);
  for (int i = 0; i < V; i++) for (int d = 0; d < 3; d++) mesh[i][d] = mesh0.vertices[i][d] - (lo[d] + hi[d])/2.0;   // :317-318
  const T posRatio = 1e-6/dx;
  const int deny = (int)((minDist_um*1e-6)/dx);
  const int n[3] = {nx, ny, nz};
  std::vector<int64_t> ids;
  std::vector<Vec3> p(V);
  for (size_t c = 0; c < rows.size(); c++) {
    T ang[3];
    for (int k = 0; k < 3; k++) { ang[k] = rows[c][3+k] * (PI/180.0); ang[k] *= -1.0; }     // :228-229
    Vec3 mlo = mesh[0], mhi = mesh[0];
    for (auto& v : mesh) for (int d = 0; d < 3; d++) { mlo[d] = std::min(mlo[d], v[d]); mhi[d] = std::max(mhi[d], v[d]); }
    Vec3 ctr; for (int d = 0; d < 3; d++) ctr[d] = (mhi[d] + mlo[d]) * 0.5;                   // meshRotation :100-107
    T a[3][3]; rotationXYZ(ang[0], ang[1], ang[2], a);
    bool ok = true;
    for (int i = 0; i < V && ok; i++) {
      Vec3 x = sub(mesh[i], ctr), r;
      for (int k = 0; k < 3; k++) r[k] = a[k][0]*x[0] + a[k][1]*x[1] + a[k][2]*x[2];
      for (int d = 0; d < 3; d++) {
        p[i][d] = rows[c][d]*posRatio + (r[d] + ctr[d]);                                     // :129, :349
        if (!(p[i][d] > -0.5 && p[i][d] <= n[d] - 0.5)) ok = false;
      }
      if (ok && flags) {
        const int q[3] = {(int)(p[i][0] + 0.5), (int)(p[i][1] + 0.5), (int)(p[i][2] + 0.5)};
        for (int px = -deny; px <= deny && ok; px++) for (int py = -deny; py <= deny && ok; py++)
          for (int pz = -deny; pz <= deny && ok; pz++) {
            const int qq[3] = {q[0]+px, q[1]+py, q[2]+pz};
            if (qq[0] < 0 || qq[0] >= nx || qq[1] < 0 || qq[1] >= ny || qq[2] < 0 || qq[2] >= nz) continue;
            if (flags[(int64_t)qq[2] + (int64_t)nz*((int64_t)qq[1] + (int64_t)ny*qq[0])] != HCG_FLUID) ok = false;
          }
      }
    }
    if (!ok) continue;
    ids.push_back(cell_id0 + (int64_t)c);
    for (int i = 0; i < V; i++) for (int d = 0; d < 3; d++) out.push_back(p[i][d]);
  }
  return ids;
}

}  // namespace host
}  // namespace hemo
