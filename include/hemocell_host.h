/*
 * hemocell_host.h -- C view of the C++ host-side set-up code (hemocell_b200/host/), exported
 * by the same libhemocell_gpu.so.  Lets a non-C++ harness (ctypes) run the product's own mesh
 * generation, CommonCellConstants construction, unit conversion and .pos placement instead of
 * re-implementing them.  Reference interfaces replaced (paths in the reference tree):
 *   hch_parameters       hemo::Parameters::lbm_base_parameters   mechanics/constantConversion.cpp:36-59
 *   hch_celltype_build   HemoCellField ctor + CommonCellConstants + calculate_k*
 *                        core/hemoCellField.cpp:38-118, mechanics/commonCellConstants.cpp:70-409,
 *                        mechanics/cellMechanics.h:50-78
 *   hch_read_pos / hch_place_cells   readPositionsBloodCellField3D + loadParticles
 *                        io/readPositionsBloodCells.cpp:120-361, core/hemoCell.cpp:191-197
 */
#ifndef HEMOCELL_HOST_H
#define HEMOCELL_HOST_H
#include <stdint.h>
#include "hemocell_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hch_celltype hch_celltype;

/* out[7] = tau, nu_lbm, dt, dm, df, f_limit, kBT_lbm */
void hch_parameters(double dx, double dt, double nu_p, double rho_p, double kBT_p, double* out7);

/* construct_type: 1 = RBC_FROM_SPHERE, 6 = ELLIPSOID_FROM_SPHERE; returns NULL on error */
hch_celltype* hch_celltype_build(int32_t model, int32_t construct_type,
                                 double dx, double dt, double nu_p, double rho_p, double kBT_p,
                                 double kBend, double kVolume, double kArea, double kLink, double eta_m,
                                 double radius_m, double aspect_ratio, int32_t min_num_triangles,
                                 const int32_t* inner_edges /*[n][2]*/, int32_t n_inner_edges);
const hcg_celltype* hch_celltype_view(const hch_celltype*);
int32_t hch_celltype_vertices(const hch_celltype*, double* out /*[V][3] lattice units*/);
double hch_celltype_scalar(const hch_celltype*, int32_t which /*0 volume_eq,1 surface,2 angle_mean_eq*/);
void hch_celltype_free(hch_celltype*);

/* returns the number of rows in the file (rows6 may be NULL to query), -1 on error */
int64_t hch_read_pos(const char* path, double* rows6, int64_t capacity_rows);
/* returns the number of surviving cells; out_pos capacity n_rows*V*3, out_ids capacity n_rows */
int64_t hch_place_cells(const hch_celltype*, const double* rows6, int64_t n_rows, double dx,
                        int32_t nx, int32_t ny, int32_t nz, const uint8_t* flags /*may be NULL*/,
                        double min_dist_from_solid_um, int64_t cell_id0, double* out_pos, int64_t* out_ids);
/* slab membership of cells for the multi-GPU exchange (pure host logic of csrc/multi.cu; replaces the
 * envelope tests of HemoCellParticleField::isContainedABS, core/hemoCellParticleField.h:93-103):
 * from the cells' x-extents decide which this rank holds and which it shares through each face */
void hch_slab_membership(int64_t n, const double* xlo, const double* xhi, int32_t nx, int32_t periodic_x,
                         int32_t nxl, int32_t rank, int32_t n_ranks, double margin,
                         uint8_t* held, uint8_t* share_left, uint8_t* share_right);
/* the same for a slab given by its first plane x0 and thickness nxl (uneven decompositions, hcg_slab) */
void hch_slab_membership_at(int64_t n, const double* xlo, const double* xhi, int32_t nx, int32_t periodic_x,
                            int32_t x0, int32_t nxl, int32_t rank, int32_t n_ranks, double margin,
                            uint8_t* held, uint8_t* share_left, uint8_t* share_right);
/* hand-over rule of the pre-inlet (pure host logic of csrc/preinlet.cu; replaces the box tests of
 * HemoCellParticleDataTransfer::send_preinlet / HemoCellParticleField::addParticlePreinlet, core/hemoCellParticleDataTransfer.cpp:99-121,
 * core/hemoCellParticleField.cpp:237-282, in whole cells): from the extents [lo, hi] of the pre-inlet's cells along the flow axis decide
 * which periodic image k (lap_out) lies wholly inside the inflow slab [slab_lo, slab_hi] of the main domain (extent + shift +
 * k*period) and whether the cell is handed over now (take_out: alive, inside, and image k not handed over before: last_lap, may be NULL) */
void hch_preinlet_select(int64_t n, const double* lo, const double* hi, const uint8_t* alive, const int64_t* last_lap,
                         double shift, double period, double slab_lo, double slab_hi, int64_t* lap_out, uint8_t* take_out);
/* STL voxeliser behind hemo::getFlagMatrixFromSTL (helper/voxelizeDomain.cpp:63-158; Palabos TriangleSet ->
 * DEFscaledMesh -> VoxelizedDomain3D in the reference).  dims_out[3] = lattice size; flags (may be NULL to query the
 * size) receives HCG_FLUID / HCG_BOUNCEBACK per node, index z + nz*(y + ny*x); dx_out = STL units per lattice unit.
 * Returns 0, or -1 on error. */
int32_t hch_voxelize_stl(const char* path, int32_t ref_dir_n, int32_t ref_dir, int32_t* dims_out, uint8_t* flags,
                         int64_t flags_capacity, double* dx_out);
/* HDF5 container writer behind HemoCell::writeOutput (replaces the H5Fcreate / H5LTset_attribute_* /
 * H5Dcreate2 + H5Dwrite calls of io/ParticleHdf5IO.cpp:60-194 and io/FluidHdf5IO.hh:36-49).
 * type: 0 = float32, 1 = float64, 2 = int32, 3 = int64.  deflate_level < 0 writes contiguous datasets;
 * chunk may be NULL (contiguous).  hch_h5_close returns 0 on success and frees the handle. */
typedef struct hch_h5 hch_h5;
hch_h5* hch_h5_create(const char* path, int32_t deflate_level);
int32_t hch_h5_attribute(hch_h5*, const char* name, int32_t type, const void* data, int64_t n);
int32_t hch_h5_dataset(hch_h5*, const char* name, int32_t type, int32_t rank, const uint64_t* dims,
                       const void* data, const uint64_t* chunk);
int32_t hch_h5_close(hch_h5*);
const char* hch_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
