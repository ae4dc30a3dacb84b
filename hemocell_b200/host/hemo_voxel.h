// STL voxeliser behind hemo::getFlagMatrixFromSTL (reference helper/voxelizeDomain.cpp:63-158), which in the
// reference is Palabos' TriangleSet -> DEFscaledMesh -> TriangleBoundary3D -> inflate() -> VoxelizedDomain3D chain.
// Restated from the Palabos conventions (from memory; pinned by the reference's own pipeflow validation test,
// tests/validation/pipeflow/test_pipeflow.cpp:90-92: 42 cells survive placement in the voxelised tube.stl):
//   * dx = (extent of the mesh along refDir) / refDirN; the mesh is moved so that its lower corner sits at
//     (margin, margin, margin) lattice units, margin = 1;
//   * the lattice has (int)extent + 1 + 2*margin nodes per direction (N cells -> N + 1 nodes);
//   * every vertex moves 1e-3 lu outwards (inflate), a node is fluid when it lies inside the closed surface;
//   * HemoCell then opens the two x ends by copying slice 2 into slices 0..1 and slice nx-3 into nx-2..nx-1.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace hemo { namespace host {

struct VoxelizedSTL {
  int nx = 0, ny = 0, nz = 0;
  double dx = 0;                       // physical length of one lattice unit in the STL's units
  double location[3] = {0, 0, 0};      // physical position of node (0,0,0)
  std::vector<int32_t> flag;           // 1 = fluid (inside), 0 = solid; index z + nz*(y + ny*x)
};

// throws std::runtime_error (unreadable file, open surface)
VoxelizedSTL voxelizeSTL(const std::string& path, int refDirN, int refDir, int margin = 1, bool openXEnds = true);

}}  // namespace hemo::host
