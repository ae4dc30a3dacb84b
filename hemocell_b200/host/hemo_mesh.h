// Host-side set-up of the hot path (C++, mirrors the reference's names): reference-shape
// meshes, CommonCellConstants, Parameters, .pos placement.  Runs once per cell type; the
// tables it produces are what hcg_celltype_add() uploads.
//   reference: helper/meshGeneratingFunctions.{h,hh}, mechanics/commonCellConstants.{h,cpp},
//              mechanics/constantConversion.{h,cpp}, mechanics/cellMechanics.h,
//              io/readPositionsBloodCells.cpp, core/hemoCellField.cpp
#pragma once
#include <array>
#include <cstdint>
#include <string>
#include <vector>
#include "hemocell_gpu.h"

namespace hemo {
namespace host {

typedef double T;
using Vec3 = std::array<T, 3>;

// constructType of HemoCell::addCellType (config/constant_defaults.h)
enum { RBC_FROM_SPHERE = 1, ELLIPSOID_FROM_SPHERE = 6 };

// stands in for plb::TriangularSurfaceMesh<T> (vertices numbered by first appearance)
struct TriangularSurfaceMesh {
  std::vector<Vec3> vertices;
  std::vector<std::array<int, 3>> triangles;
  int getNumVertices() const { return (int)vertices.size(); }
  int getNumTriangles() const { return (int)triangles.size(); }
  T getVolume() const;
  T getSurface() const;
  void boundingBox(Vec3& lo, Vec3& hi) const;
};

// helper/meshGeneratingFunctions.h:68-94
TriangularSurfaceMesh constructMeshElement(int shape, T radius, int cellNumTriangles, T aspectRatio = 0.3);

// mechanics/constantConversion.{h,cpp}: hemo::Parameters (param::)
struct Parameters {
  T dx = 0, dt = 0, nu_p = 0, rho_p = 0, kBT_p = 0;
  T tau = 0, nu_lbm = 0, dm = 0, df = 0, f_limit = 0, kBT_lbm = 0;
  T re = 0, pipe_radius = 0, u_lbm_max = 0, shearrate_lbm = 0;
  void lbm_base_parameters(T dx_, T dt_, T nu_p_, T rho_p_, T kBT_p_);   // constantConversion.cpp:36-59
  void lbm_pipe_parameters(T Re, int nY);                                 // :76-82
  void lbm_shear_parameters(T shearrate_p, T nx);                         // :84-90
};

struct MaterialModel {           // <MaterialModel> of <CELL>.xml
  T kBend = 0, kVolume = 0, kArea = 0, kLink = 0, eta_m = 0;
  T radius = 0, aspectRatio = 0.3, volume = 0;
  int minNumTriangles = 0;
  std::vector<std::array<int, 2>> innerEdges;
};

// mechanics/commonCellConstants.h:39-86
struct CommonCellConstants {
  std::vector<std::array<int, 3>> triangle_list;
  std::vector<std::array<int, 2>> edge_list;
  std::vector<T> edge_length_eq_list, edge_angle_eq_list, surface_patch_center_dist_eq_list;
  std::vector<std::array<int, 2>> edge_bending_triangles_list, edge_bending_triangles_outer_points;
  std::vector<T> triangle_area_eq_list;
  std::vector<std::array<int, 6>> vertex_vertexes;
  std::vector<int> vertex_n_vertexes;
  T volume_eq = 0, area_mean_eq = 0, edge_mean_eq = 0, angle_mean_eq = 0;
  std::vector<std::array<int, 2>> inner_edge_list;
  std::vector<T> inner_edge_length_eq_list;
  static CommonCellConstants CommonCellConstantsConstructor(const TriangularSurfaceMesh& mesh,
                                                            const std::vector<std::array<int, 2>>& innerEdges);
};

// CellMechanics::calculate_k* (mechanics/cellMechanics.h:50-78)
struct Stiffness { T k_volume, k_area, k_link, k_bend, eta_m; };
Stiffness calculate_stiffness(const MaterialModel& m, const Parameters& p, int nTriangles);

// one cell type ready for the GPU: mesh + constants + stiffness, and the flat hcg_celltype view
struct CellTypeTables {
  int model = HCG_MODEL_RBC_HIGHORDER;
  TriangularSurfaceMesh mesh;
  CommonCellConstants cc;
  Stiffness k{};
  // flat storage backing the C struct
  std::vector<int32_t> f_tri, f_edge, f_inner, f_vv, f_nvv, f_bt, f_bo;
  hcg_celltype c{};
  void build(int model_, int constructType, const MaterialModel& m, const Parameters& p);
};

// io/readPositionsBloodCells.cpp: parse "<n>\n x y z rx ry rz ..." (um, degrees)
std::vector<std::array<T, 6>> readPositionsFile(const std::string& path);
// placement + the incomplete-cell purge of loadParticles, on the single global lattice
// (positionCellInParticleField :120-170, processGenericBlocks :290-361, core/hemoCell.cpp:191-197).
// flags: global lattice flags (may be null = all fluid).  Returns surviving cell ids (index in
// rows + cell_id0); positions appended to `out` as [cell][vertex][3].
std::vector<int64_t> placeCells(const TriangularSurfaceMesh& mesh, const std::vector<std::array<T, 6>>& rows,
                                T dx, int nx, int ny, int nz, const uint8_t* flags, T minDistFromSolid_um,
                                int64_t cell_id0, std::vector<T>& out);

}  // namespace host
}  // namespace hemo
