/*
 * hemo_oracle.h -- CPU ORACLE for the HemoCell IB-LBM hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under hemocell_b200/ (the product) may
 * include, link or call this.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, as the checker.
 *
 * PARITY STATUS: "parity unpinned" at the per-operator level.  The reference
 * (UvaCsl/HemoCell) cannot be built in this container (Palabos v2.3.0, MPI and
 * HDF5 are absent, setup.sh:9 downloads Palabos), and its own tests hold no
 * per-operator golden vectors.  This file is a plain-C
 * restatement of the reference algorithm, each function citing the reference
 * file:line it follows.  What *is* pinned (tests/test_oracle_known_answers.py):
 * the reference's two CI gates and two validation tests run with this oracle, their windows
 * unwidened - stretchCell diameter / volume / surface (scripts/ci/stretchCell_sanity.sh:6-33),
 * pipeflow cell count / apparent viscosity / particle force (scripts/ci/pipeflow_sanity.sh:6-22,
 * tests/validation/pipeflow/test_pipeflow.cpp:88-106), the stretch-cell force-displacement bounds
 * (tests/validation/stretch_cell/test_stretch_cell.cpp:158-162) - and the V/T/E counts.  The Zou-He velocity / pressure
 * nodes (Palabos code, not in the tree) are pinned by physics instead: imposed density and
 * momentum reproduced to 1e-14, plane Poiseuille flow within 2 %, flux continuity through the
 * pre-inlet coupling (tests/test_preinlet_oracle.py).
 *
 * Array conventions (identical to include/hemocell_gpu.h):
 *   node index      idx = z + nz*(y + ny*x)            (patch/palabos.patch:245)
 *   populations     pop[q*N + idx], q = 0..18, stored as f_q - t_q, POST-STREAM
 *                   (i.e. exactly what Palabos holds between collideAndStream calls)
 *   node force      force[d*N + idx], d = 0..2
 *   flags           uint8 per node: 0 fluid, 1 bounce-back, 2..7 velocity plane
 *                   whose OUTWARD normal is -x,+x,-y,+y,-z,+z; 8..13 Zou-He velocity node,
 *                   14..19 Zou-He pressure node (same normal order)
 *   particles       AoS xyz: pos[3*p + d]
 */
#ifndef HEMO_ORACLE_H
#define HEMO_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORA_FLUID = 0, ORA_BB = 1, ORA_VEL_XN = 2, ORA_VEL_XP = 3, ORA_VEL_YN = 4,
       ORA_VEL_YP = 5, ORA_VEL_ZN = 6, ORA_VEL_ZP = 7,
       /* Zou-He velocity nodes (per-node velocity) and pressure nodes (per-node density), OUTWARD normal
        * -x,+x,-y,+y,-z,+z (helper/preInlet.cpp:399-436, pipeflow_with_preinlet.cpp:125-133) */
       ORA_ZH_VEL_XN = 8, ORA_ZH_VEL_ZP = 13, ORA_ZH_PRES_XN = 14, ORA_ZH_PRES_ZP = 19 };

typedef struct {
  int32_t nx, ny, nz;
  int32_t periodic[3];
  double omega;            /* 1/tau */
  double bc_vel[6][3];     /* wall velocity per velocity-plane orientation (flag-2) */
} ora_domain;

/* per cell-type topology and constants == CommonCellConstants + k_* of the model
 * (mechanics/commonCellConstants.cpp:70-409, mechanics/cellMechanics.h:50-78) */
typedef struct {
  int32_t model;           /* 0 = RbcHighOrderModel, 1 = PltSimpleModel */
  int32_t n_vertices, n_triangles, n_edges, n_inner_edges;
  const int32_t* triangles;        /* [T][3] */
  const int32_t* edges;            /* [E][2] */
  const int32_t* inner_edges;      /* [I][2] */
  const int32_t* vertex_vertexes;  /* [V][6], -1 padded, ring ordered */
  const int32_t* vertex_n_vertexes;/* [V] */
  const int32_t* edge_bending_triangles;       /* [E][2] */
  const int32_t* edge_bending_outer_points;    /* [E][2] */
  const double* edge_length_eq;    /* [E] */
  const double* edge_angle_eq;     /* [E] */
  const double* triangle_area_eq;  /* [T] */
  const double* patch_dist_eq;     /* [V] surface_patch_center_dist_eq_list */
  const double* inner_edge_length_eq; /* [I] */
  double volume_eq, area_mean_eq, edge_mean_eq;
  double k_volume, k_area, k_link, k_bend, eta_m;
} ora_celltype;

/* OpenMP on/off (off by default; on only for the timed cpu_baseline legs of bench.py) */
void ora_set_parallel(int on);

/* ---- lattice (Palabos v2.3.0 behaviour restated) ---- */
void ora_init_equilibrium(const ora_domain* d, double rho, const double u[3], double* pop);
void ora_collide_and_stream(const ora_domain* d, const uint8_t* flags, double* pop,
                            const double* force, double* scratch /* 19*N */);
void ora_moments(const ora_domain* d, const uint8_t* flags, const double* pop,
                 const double* force, double* rho /* N or NULL */, double* vel /* 3*N */);
/* the same with Zou-He velocity / pressure nodes: bc_node [4][N] = (u_x, u_y, u_z, rho) per node (NULL: u = 0, rho = 1) */
void ora_collide_and_stream_io(const ora_domain* d, const uint8_t* flags, double* pop,
                               const double* force, double* scratch, const double* bc_node);
void ora_moments_io(const ora_domain* d, const uint8_t* flags, const double* pop,
                    const double* force, double* rho, double* vel, const double* bc_node);

/* ---- IBM (core/immersedBoundaryMethod.h:62-138, hemoCellParticleField.cpp:819-863) ---- */
int ora_ibm_kernel(const ora_domain* d, const uint8_t* flags, const double p[3],
                   int64_t node[8], double w[8]);
void ora_spread(const ora_domain* d, const uint8_t* flags, int64_t np, const double* pos,
                double* pforce /* capped in place */, const double* frep, double f_limit,
                double* node_force);
void ora_interpolate(const ora_domain* d, const uint8_t* flags, int64_t np, const double* pos,
                     const double* pop, const double* node_force, double* vel);
/* advance (hemoCellParticle.h:188-203, hemoCellParticleField.cpp:566-588): returns number
 * of particles that landed on a boundary node; hit[p] set to 1 for those. */
int64_t ora_advance(const ora_domain* d, const uint8_t* flags, int64_t np, double* pos,
                    const double* vel, uint8_t* hit);

/* ---- mechanics (rbcHighOrderModel.cpp:38-207, pltSimpleModel.cpp:44-208) ---- */
/* comp: NULL, or 6 arrays [area, volume, bending, link, visc, inner] each 3*V*ncells */
void ora_mechanics(const ora_celltype* t, int64_t n_cells, const double* pos,
                   const double* vel, double* force /* accumulated, caller zeroes */,
                   double* const* comp);

/* ---- repulsion (hemoCellParticleField.cpp:137-168, 677-743, 865-918) ---- */
void ora_repulsion(const ora_domain* d, int64_t np, const double* pos, const int64_t* cell_of,
                   double k, double cutoff, double* frep /* zeroed inside */);
void ora_wall_repulsion(const ora_domain* d, const uint8_t* flags, int64_t np,
                        const double* pos, double k, double cutoff, double* frep /* accumulated */);

#ifdef __cplusplus
}
#endif
#endif
