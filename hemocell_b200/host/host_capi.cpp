// extern "C" view of the host-side set-up code (include/hemocell_host.h)
#include "hemocell_host.h"
#include "hemo_mesh.h"
#include "hemo_h5.h"
#include "hemo_voxel.h"
#include <cstring>
#include <exception>
#include <string>

struct hch_celltype { hemo::host::CellTypeTables t; };
static thread_local std::string g_err;

extern "C" {

const char* hch_last_error(void) { return g_err.c_str(); }

void hch_parameters(double dx, double dt, double nu_p, double rho_p, double kBT_p, double* out7) {
  hemo::host::Parameters p; p.lbm_base_parameters(dx, dt, nu_p, rho_p, kBT_p);
  out7[0] = p.tau; out7[1] = p.nu_lbm; out7[2] = p.dt; out7[3] = p.dm; out7[4] = p.df; out7[5] = p.f_limit; out7[6] = p.kBT_lbm;
}

hch_celltype* hch_celltype_build(int32_t model, int32_t construct_type, double dx, double dt, double nu_p,
                                 double rho_p, double kBT_p, double kBend, double kVolume, double kArea,
                                 double kLink, double eta_m, double radius_m, double aspect_ratio,
                                 int32_t min_num_triangles, const int32_t* inner_edges, int32_t n_inner_edges) {
  try {
    hemo::host::Parameters p; p.lbm_base_parameters(dx, dt, nu_p, rho_p, kBT_p);
    hemo::host::MaterialModel m;
    m.kBend = kBend; m.kVolume = kVolume; m.kArea = kArea; m.kLink = kLink; m.eta_m = eta_m;
    m.radius = radius_m; m.aspectRatio = aspect_ratio; m.minNumTriangles = min_num_triangles;
    for (int i = 0; i < n_inner_edges; i++) m.innerEdges.push_back({inner_edges[2*i], inner_edges[2*i+1]});
    hch_celltype* h = new hch_celltype();
    h->t.build(model, construct_type, m, p);
    return h;
  } catch (std::exception& e) { g_err = e.what(); return nullptr; }
}
const hcg_celltype* hch_celltype_view(const hch_celltype* h) { return h ? &h->t.c : nullptr; }
int32_t hch_celltype_vertices(const hch_celltype* h, double* out) {
  if (!h) return -1;
  const int V = h->t.mesh.getNumVertices();
  if (out) for (int i = 0; i < V; i++) for (int d = 0; d < 3; d++) out[3*i+d] = h->t.mesh.vertices[i][d];
  return V;
}
double hch_celltype_scalar(const hch_celltype* h, int32_t which) {
  if (!h) return 0.0;
  if (which == 0) return h->t.cc.volume_eq;
  if (which == 1) return h->t.mesh.getSurface();
  return h->t.cc.angle_mean_eq;
}
void hch_celltype_free(hch_celltype* h) { delete h; }

int64_t hch_read_pos(const char* path, double* rows6, int64_t cap) {
  try {
    auto rows = hemo::host::readPositionsFile(path);
    if (rows6) for (size_t i = 0; i < rows.size() && (int64_t)i < cap; i++) memcpy(rows6 + 6*i, rows[i].data(), 6*sizeof(double));
    return (int64_t)rows.size();
  } catch (std::exception& e) { g_err = e.what(); return -1; }
}

int64_t hch_place_cells(const hch_celltype* h, const double* rows6, int64_t n_rows, double dx, int32_t nx,
                        int32_t ny, int32_t nz, const uint8_t* flags, double min_dist_um, int64_t cell_id0,
                        double* out_pos, int64_t* out_ids) {
  if (!h || !rows6 || !out_pos || !out_ids) { g_err = "null argument"; return -1; }
  try {
    std::vector<std::array<double, 6>> rows(n_rows);
    for (int64_t i = 0; i < n_rows; i++) memcpy(rows[i].data(), rows6 + 6*i, 6*sizeof(double));
    std::vector<double> out;
    auto ids = hemo::host::placeCells(h->t.mesh, rows, dx, nx, ny, nz, flags, min_dist_um, cell_id0, out);
    memcpy(out_pos, out.data(), out.size()*sizeof(double));
    memcpy(out_ids, ids.data(), ids.size()*sizeof(int64_t));
    return (int64_t)ids.size();
  } catch (std::exception& e) { g_err = e.what(); return -1; }
}

int32_t hch_voxelize_stl(const char* path, int32_t ref_dir_n, int32_t ref_dir, int32_t* dims_out, uint8_t* flags,
                         int64_t cap, double* dx_out) {
  if (!path || !dims_out) { g_err = "null argument"; return -1; }
  try {
    hemo::host::VoxelizedSTL v = hemo::host::voxelizeSTL(path, ref_dir_n, ref_dir);
    dims_out[0] = v.nx; dims_out[1] = v.ny; dims_out[2] = v.nz;
    if (dx_out) *dx_out = v.dx;
    if (flags) {
      if (cap < (int64_t)v.flag.size()) { g_err = "flags buffer too small"; return -1; }
      for (size_t i = 0; i < v.flag.size(); i++) flags[i] = v.flag[i] ? HCG_FLUID : HCG_BOUNCEBACK;
    }
    return 0;
  } catch (std::exception& e) { g_err = e.what(); return -1; }
}

struct hch_h5 { hemo::h5::Writer w; hch_h5(const char* p, int l) : w(p, l) {} };
hch_h5* hch_h5_create(const char* path, int32_t deflate_level) {
  if (!path) { g_err = "null argument"; return nullptr; }
  hch_h5* h = new hch_h5(path, deflate_level);
  if (!h->w.ok()) { g_err = h->w.error(); delete h; return nullptr; }
  return h;
}
int32_t hch_h5_attribute(hch_h5* h, const char* name, int32_t type, const void* data, int64_t n) {
  if (!h || !name || !data || type < 0 || type > 3 || n < 1) { g_err = "bad argument"; return -1; }
  h->w.attribute(name, (hemo::h5::Type)type, data, (size_t)n);
  return 0;
}
int32_t hch_h5_dataset(hch_h5* h, const char* name, int32_t type, int32_t rank, const uint64_t* dims,
                       const void* data, const uint64_t* chunk) {
  if (!h || !name || !dims || type < 0 || type > 3 || rank < 1 || rank > 8) { g_err = "bad argument"; return -1; }
  std::vector<uint64_t> d(dims, dims + rank), c;
  if (chunk) c.assign(chunk, chunk + rank);
  h->w.dataset(name, (hemo::h5::Type)type, d, data, c);
  if (!h->w.ok()) { g_err = h->w.error(); return -1; }
  return 0;
}
int32_t hch_h5_close(hch_h5* h) {
  if (!h) return -1;
  const bool ok = h->w.close();
  if (!ok) g_err = h->w.error();
  delete h;
  return ok ? 0 : -1;
}

}  // extern "C"
