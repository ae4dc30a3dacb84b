/*
 * hemocell.h -- the HemoCell C++ API surface, on top of the B200 C ABI (hemocell_gpu.h).
 *
 * A case file written against the reference (examples/<case>/<case>.cpp) includes "hemocell.h" and the two
 * Palabos headers, builds a lattice, registers cell types and calls HemoCell::iterate().  This
 * header keeps that surface source-compatible for the per-timestep hot path:
 *
 *   hemo::HemoCell            reference hemocell.h:68-253, core/hemoCell.cpp
 *   hemo::HemoCellFields      core/hemoCellFields.h:52-331
 *   hemo::HemoCellField       core/hemoCellField.h:42-84
 *   hemo::CellMechanics, RbcHighOrderModel, PltSimpleModel
 *                             mechanics/cellMechanics.h:37-79, rbcHighOrderModel.h:34-51, pltSimpleModel.h
 *   hemo::Config / XMLElement config/config.h:37-78
 *   hemo::Parameters (param::) mechanics/constantConversion.h
 *   hemo::hlog, hemo::global, Profiler      config/logfile.h, config/config.h:82-98, helper/profiler.h
 *   CellInformationFunctionals, FluidInfo   helper/cellInfo.h, helper/fluidInfo.h
 *   plb:: shim                exactly the Palabos calls the reference's case files make (lattice
 *                             construction, periodicity, defineDynamics, velocity planes, setExternalVector,
 *                             initializeAtEquilibrium, collideAndStream); each call records into a host-side
 *                             domain description that is uploaded through the C ABI.
 *
 * Everything below runs on the host and only orchestrates; the operators themselves are the CUDA
 * kernels behind hcg_iterate().  There is no CPU implementation of the hot path here: without a
 * CUDA device the HemoCell constructor logs the error and exits, as the reference does on fatal errors.
 */
#ifndef HEMOCELL_FACADE_H
#define HEMOCELL_FACADE_H

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "hemocell_gpu.h"

/* ---- config/constant_defaults.h ------------------------------------------------------------------ */
#ifndef FORCE_LIMIT
#define FORCE_LIMIT 50.0                       /* pN */
#endif
#define RBC_FROM_SPHERE 1
#define ELLIPSOID_FROM_SPHERE 6
#define OUTPUT_POSITION 1
#define OUTPUT_VELOCITY 2
#define OUTPUT_FORCE 3
#define OUTPUT_TRIANGLES 4
#define OUTPUT_DENSITY 5
#define OUTPUT_FORCE_VOLUME 19
#define OUTPUT_FORCE_AREA 20
#define OUTPUT_FORCE_LINK 21
#define OUTPUT_FORCE_BENDING 22
#define OUTPUT_FORCE_VISC 23
#define OUTPUT_FORCE_INNER_LINK 24
#define OUTPUT_FORCE_REPULSION 25
#define OUTPUT_VERTEX_ID 7
#define OUTPUT_CELL_ID 8
#define OUTPUT_CELL_DENSITY 9
#define OUTPUT_SHEAR_STRESS 10
#define OUTPUT_INNER_LINKS 11
#define OUTPUT_OMEGA 12
#define OUTPUT_BOUNDARY 13
#define OUTPUT_SHEAR_RATE 16
#define OUTPUT_STRAIN_RATE 17
#define OUTPUT_RES_TIME 18
#ifndef PI
#define PI 3.14159265358979323846
#endif
#define param Parameters
#define DESCRIPTOR plb::descriptors::ForcedD3Q19Descriptor

typedef double T;

namespace plb { struct Box3D; struct DomainFunctional3D; }
namespace hemo {
class GpuLattice; class HemoCell; class HemoCellFields; class HemoCellField; class Config; class PreInlet;
/* non-template back end of the plb:: shim (hemocell_b200/host/facade.cpp) */
GpuLattice* gpu_lattice_create(long nx, long ny, long nz, double omega);
void gpu_lattice_destroy(GpuLattice*);
void gpu_lattice_size(const GpuLattice*, long out[3]);
void gpu_lattice_set_periodic(GpuLattice*, int axis, bool on);
bool gpu_lattice_get_periodic(const GpuLattice*, int axis);
void gpu_lattice_collide_and_stream(GpuLattice*);
void gpu_lattice_velocity_plane(GpuLattice*, const plb::Box3D& plane);
void gpu_lattice_velocity_all_faces(GpuLattice*);
void gpu_lattice_boundary_velocity(GpuLattice*, const plb::Box3D& domain, const double u[3]);
void gpu_lattice_external_vector(GpuLattice*, const plb::Box3D& domain, const double v[3]);
void gpu_lattice_define_flag(GpuLattice*, const plb::Box3D& domain, const plb::DomainFunctional3D* fun, int flag);
void gpu_lattice_equilibrium(GpuLattice*, double rho, const double u[3]);
void gpu_lattice_zouhe(GpuLattice*, const plb::Box3D& domain, int pressure, int orientation);   /* Zou-He velocity / pressure nodes */
void gpu_lattice_boundary_density(GpuLattice*, const plb::Box3D& domain, double rho);
int gpu_lattice_flag(const GpuLattice*, long x, long y, long z);   /* HCG_* node flag as recorded on the host */
std::string gpu_lattice_info(const GpuLattice*);
}

/* ================================================================================================== */
/* plb:: shim                                                                                         */
/* ================================================================================================== */
namespace plb {

typedef long int plint;
typedef long unsigned int pluint;

template <typename U, pluint n>
class Array {
 public:
  U data[n];
  Array() { for (pluint i = 0; i < n; i++) data[i] = U(); }
  Array(U a, U b) { static_assert(n == 2, "size"); data[0] = a; data[1] = b; }
  Array(U a, U b, U c) { static_assert(n == 3, "size"); data[0] = a; data[1] = b; data[2] = c; }
  U& operator[](pluint i) { return data[i]; }
  const U& operator[](pluint i) const { return data[i]; }
};

struct Box3D {
  plint x0, x1, y0, y1, z0, z1;
  Box3D() : x0(0), x1(0), y0(0), y1(0), z0(0), z1(0) {}
  Box3D(plint x0_, plint x1_, plint y0_, plint y1_, plint z0_, plint z1_) : x0(x0_), x1(x1_), y0(y0_), y1(y1_), z0(z0_), z1(z1_) {}
  plint getNx() const { return x1 - x0 + 1; }
  plint getNy() const { return y1 - y0 + 1; }
  plint getNz() const { return z1 - z0 + 1; }
  plint nCells() const { return getNx()*getNy()*getNz(); }
  Box3D enlarge(plint w) const { return Box3D(x0 - w, x1 + w, y0 - w, y1 + w, z0 - w, z1 + w); }
};
struct Dot3D { plint x, y, z; Dot3D() : x(0), y(0), z(0) {} Dot3D(plint a, plint b, plint c) : x(a), y(b), z(c) {} };

namespace descriptors {
template <typename U> struct ForcedD3Q19Descriptor {
  enum { d = 3, q = 19 };
  struct ExternalField { enum { numScalars = 3, numSpecies = 1, forceBeginsAt = 0, sizeOfForce = 3 }; };
};
}  // namespace descriptors

/* dynamics objects are tags here: the device kernel switches on a node flag */
template <typename U, template <typename V> class Descriptor>
struct Dynamics {
  virtual ~Dynamics() {}
  virtual bool isBoundary() const { return false; }
  virtual U getOmega() const { return U(); }
  virtual int nodeFlag() const { return HCG_FLUID; }
};
template <typename U, template <typename V> class Descriptor>
struct GuoExternalForceBGKdynamics : public Dynamics<U, Descriptor> {
  U omega;
  explicit GuoExternalForceBGKdynamics(U omega_) : omega(omega_) {}
  U getOmega() const override { return omega; }
};
template <typename U, template <typename V> class Descriptor>
struct BounceBack : public Dynamics<U, Descriptor> {
  explicit BounceBack(U rho_ = U(1)) { (void)rho_; }
  bool isBoundary() const override { return true; }
  int nodeFlag() const override { return HCG_BOUNCEBACK; }
};

struct DomainFunctional3D {
  virtual ~DomainFunctional3D() {}
  virtual bool operator()(plint iX, plint iY, plint iZ) const = 0;
  virtual DomainFunctional3D* clone() const = 0;
};

class MultiBlockManagement3D {
 public:
  MultiBlockManagement3D(plint nx, plint ny, plint nz, plint envelope) : box(0, nx - 1, 0, ny - 1, 0, nz - 1), env(envelope) {}
  Box3D getBoundingBox() const { return box; }
  plint getEnvelopeWidth() const { return env; }
  void changeEnvelopeWidth(plint w) { env = w; }
 private:
  Box3D box; plint env;
};
struct BlockCommunicator3D {};
struct CombinedStatistics {};
template <typename U, template <typename V> class Descriptor> struct MultiCellAccess3D {};
struct defaultMultiBlockPolicy3D {
  MultiBlockManagement3D getMultiBlockManagement(plint nx, plint ny, plint nz, plint envelope = 1) const { return MultiBlockManagement3D(nx, ny, nz, envelope); }
  BlockCommunicator3D* getBlockCommunicator() const { return nullptr; }
  CombinedStatistics* getCombinedStatistics() const { return nullptr; }
  template <typename U, template <typename V> class Descriptor> MultiCellAccess3D<U, Descriptor>* getMultiCellAccess() const { return nullptr; }
};

class PeriodicitySwitch3D {
 public:
  explicit PeriodicitySwitch3D(hemo::GpuLattice* l) : lat(l) {}
  void toggle(plint axis, bool on) { hemo::gpu_lattice_set_periodic(lat, (int)axis, on); }
  void toggleAll(bool on) { toggle(0, on); toggle(1, on); toggle(2, on); }
  bool get(plint axis) const { return hemo::gpu_lattice_get_periodic(lat, (int)axis); }
 private:
  hemo::GpuLattice* lat;
};

/* MultiBlockLattice3D<double, ForcedD3Q19Descriptor>: one global lattice, x-slab per GPU rank */
template <typename U, template <typename V> class Descriptor>
class MultiBlockLattice3D {
 public:
  MultiBlockLattice3D(const MultiBlockManagement3D& m, BlockCommunicator3D*, CombinedStatistics*, MultiCellAccess3D<U, Descriptor>*,
                      Dynamics<U, Descriptor>* background_)
      : impl(hemo::gpu_lattice_create(m.getBoundingBox().getNx(), m.getBoundingBox().getNy(), m.getBoundingBox().getNz(), background_->getOmega())),
        mgmt(m), background(background_), per(impl) {}
  MultiBlockLattice3D(plint nx, plint ny, plint nz, Dynamics<U, Descriptor>* background_)
      : impl(hemo::gpu_lattice_create(nx, ny, nz, background_->getOmega())), mgmt(nx, ny, nz, 1), background(background_), per(impl) {}
  ~MultiBlockLattice3D() { hemo::gpu_lattice_destroy(impl); delete background; }
  MultiBlockLattice3D(const MultiBlockLattice3D&) = delete;
  MultiBlockLattice3D& operator=(const MultiBlockLattice3D&) = delete;
  Box3D getBoundingBox() const { long n[3]; hemo::gpu_lattice_size(impl, n); return Box3D(0, n[0] - 1, 0, n[1] - 1, 0, n[2] - 1); }
  plint getNx() const { return getBoundingBox().getNx(); }
  plint getNy() const { return getBoundingBox().getNy(); }
  plint getNz() const { return getBoundingBox().getNz(); }
  PeriodicitySwitch3D& periodicity() { return per; }
  void toggleInternalStatistics(bool) {}
  void initialize() {}                       /* the device context is created on first device use (periodicity may still be toggled) */
  void collideAndStream() { hemo::gpu_lattice_collide_and_stream(impl); }   /* warm-up loops of the case files (force is NOT reset) */
  MultiBlockManagement3D& getMultiBlockManagement() { return mgmt; }
  void signalPeriodicity() {}
  Dynamics<U, Descriptor>& getBackgroundDynamics() { return *background; }
  hemo::GpuLattice* gpu() { return impl; }
 private:
  hemo::GpuLattice* impl;
  MultiBlockManagement3D mgmt;
  Dynamics<U, Descriptor>* background;
  PeriodicitySwitch3D per;
};

/* MultiScalarField3D<int>: the flag matrix of helper/voxelizeDomain.cpp (1 = fluid, 0 = solid), one global block */
template <typename U>
class MultiScalarField3D {
 public:
  MultiScalarField3D(plint nx_, plint ny_, plint nz_, U v = U()) : nx(nx_), ny(ny_), nz(nz_), data((size_t)nx_*ny_*nz_, v) {}
  Box3D getBoundingBox() const { return Box3D(0, nx - 1, 0, ny - 1, 0, nz - 1); }
  plint getNx() const { return nx; } plint getNy() const { return ny; } plint getNz() const { return nz; }
  U& get(plint x, plint y, plint z) { return data[(size_t)z + (size_t)nz*((size_t)y + (size_t)ny*x)]; }
  const U& get(plint x, plint y, plint z) const { return data[(size_t)z + (size_t)nz*((size_t)y + (size_t)ny*x)]; }
 private:
  plint nx, ny, nz; std::vector<U> data;
};
template <typename U>
U computeSum(MultiScalarField3D<U>& f, Box3D d) {
  U s = U();
  for (plint x = d.x0; x <= d.x1; x++) for (plint y = d.y0; y <= d.y1; y++) for (plint z = d.z0; z <= d.z1; z++) s += f.get(x, y, z);
  return s;
}
/* VoxelizedDomain3D<T>: only what the case files touch (the lattice size and the voxel matrix) */
template <typename U>
class VoxelizedDomain3D {
 public:
  VoxelizedDomain3D(plint nx, plint ny, plint nz, plint envelope) : mgmt(nx, ny, nz, envelope), voxels(nx, ny, nz, 0) {}
  MultiBlockManagement3D const& getMultiBlockManagement() const { return mgmt; }
  MultiScalarField3D<int>& getVoxelMatrix() { return voxels; }
 private:
  MultiBlockManagement3D mgmt; MultiScalarField3D<int> voxels;
};

namespace boundary { enum BcType { dirichlet, neumann, freeslip, density, outflow, normalOutflow }; }
template <typename U, template <typename V> class Descriptor>
class OnLatticeBoundaryCondition3D {
 public:
  /* regularized ("local") velocity condition on one face plane of the bounding box (helper/hemocellInit.hh:72-73) */
  void setVelocityConditionOnBlockBoundaries(MultiBlockLattice3D<U, Descriptor>& lattice, Box3D plane, boundary::BcType = boundary::dirichlet) {
    hemo::gpu_lattice_velocity_plane(lattice.gpu(), plane);
  }
  /* ... on all six faces (tests/validation/stretch_cell/test_stretch_cell.cpp:90) */
  void setVelocityConditionOnBlockBoundaries(MultiBlockLattice3D<U, Descriptor>& lattice, boundary::BcType = boundary::dirichlet) {
    hemo::gpu_lattice_velocity_all_faces(lattice.gpu());
  }
  virtual ~OnLatticeBoundaryCondition3D() {}
  /* Zou-He velocity / pressure nodes on any box of nodes, 0/1/2 = normal axis, N/P = OUTWARD normal -/+ (helper/preInlet.cpp:415-432,
   * examples/pipeflow_with_preinlet/pipeflow_with_preinlet.cpp:131); bounce-back nodes inside the box stay walls */
#define HEMO_ZH_BC(NAME, PRESSURE, ORIENT) \
  void NAME(Box3D domain, MultiBlockLattice3D<U, Descriptor>& lattice, boundary::BcType = boundary::dirichlet) { hemo::gpu_lattice_zouhe(lattice.gpu(), domain, PRESSURE, ORIENT); }
  HEMO_ZH_BC(addVelocityBoundary0N, 0, 0) HEMO_ZH_BC(addVelocityBoundary0P, 0, 1) HEMO_ZH_BC(addVelocityBoundary1N, 0, 2)
  HEMO_ZH_BC(addVelocityBoundary1P, 0, 3) HEMO_ZH_BC(addVelocityBoundary2N, 0, 4) HEMO_ZH_BC(addVelocityBoundary2P, 0, 5)
  HEMO_ZH_BC(addPressureBoundary0N, 1, 0) HEMO_ZH_BC(addPressureBoundary0P, 1, 1) HEMO_ZH_BC(addPressureBoundary1N, 1, 2)
  HEMO_ZH_BC(addPressureBoundary1P, 1, 3) HEMO_ZH_BC(addPressureBoundary2N, 1, 4) HEMO_ZH_BC(addPressureBoundary2P, 1, 5)
#undef HEMO_ZH_BC
};
template <typename U, template <typename V> class Descriptor> struct WrappedZouHeBoundaryManager3D {};
template <typename U, template <typename V> class Descriptor, class Manager>
class BoundaryConditionInstantiator3D : public OnLatticeBoundaryCondition3D<U, Descriptor> {};
template <typename U, template <typename V> class Descriptor>
OnLatticeBoundaryCondition3D<U, Descriptor>* createZouHeBoundaryCondition3D() { return new OnLatticeBoundaryCondition3D<U, Descriptor>(); }
template <typename U, template <typename V> class Descriptor>
void setBoundaryDensity(MultiBlockLattice3D<U, Descriptor>& lattice, Box3D domain, U rho) { hemo::gpu_lattice_boundary_density(lattice.gpu(), domain, rho); }
template <typename U, template <typename V> class Descriptor>
OnLatticeBoundaryCondition3D<U, Descriptor>* createLocalBoundaryCondition3D() { return new OnLatticeBoundaryCondition3D<U, Descriptor>(); }

template <typename U, template <typename V> class Descriptor>
void setBoundaryVelocity(MultiBlockLattice3D<U, Descriptor>& lattice, Box3D domain, Array<U, 3> velocity) {
  const double u[3] = {velocity[0], velocity[1], velocity[2]};
  hemo::gpu_lattice_boundary_velocity(lattice.gpu(), domain, u);
}
template <typename U, template <typename V> class Descriptor>
void setExternalVector(MultiBlockLattice3D<U, Descriptor>& lattice, Box3D domain, int vectorStartsAt, Array<U, 3> vec) {
  (void)vectorStartsAt;                       /* the only external vector of ForcedD3Q19Descriptor is the force */
  const double v[3] = {vec[0], vec[1], vec[2]};
  hemo::gpu_lattice_external_vector(lattice.gpu(), domain, v);
}
template <typename U, template <typename V> class Descriptor>
void defineDynamics(MultiBlockLattice3D<U, Descriptor>& lattice, Box3D domain, Dynamics<U, Descriptor>* dynamics) {
  hemo::gpu_lattice_define_flag(lattice.gpu(), domain, nullptr, dynamics->nodeFlag());
  delete dynamics;
}
template <typename U, template <typename V> class Descriptor>
void defineDynamics(MultiBlockLattice3D<U, Descriptor>& lattice, Box3D domain, DomainFunctional3D* functional, Dynamics<U, Descriptor>* dynamics) {
  hemo::gpu_lattice_define_flag(lattice.gpu(), domain, functional, dynamics->nodeFlag());
  delete dynamics; delete functional;
}
/* defineDynamics(lattice, flagMatrix, domain, dynamics, whichFlag): nodes whose flag equals whichFlag (examples/pipeflow/pipeflow.cpp:73) */
struct FlagMatrixDomain : public DomainFunctional3D {
  const MultiScalarField3D<int>& f; int which;
  FlagMatrixDomain(const MultiScalarField3D<int>& f_, int w) : f(f_), which(w) {}
  bool operator()(plint x, plint y, plint z) const override { return f.get(x, y, z) == which; }
  DomainFunctional3D* clone() const override { return new FlagMatrixDomain(f, which); }
};
template <typename U, template <typename V> class Descriptor>
void defineDynamics(MultiBlockLattice3D<U, Descriptor>& lattice, MultiScalarField3D<int>& flagMatrix, Box3D domain, Dynamics<U, Descriptor>* dynamics, int whichFlag) {
  FlagMatrixDomain fun(flagMatrix, whichFlag);
  hemo::gpu_lattice_define_flag(lattice.gpu(), domain, &fun, dynamics->nodeFlag());
  delete dynamics;
}
template <typename U, template <typename V> class Descriptor>
void initializeAtEquilibrium(MultiBlockLattice3D<U, Descriptor>& lattice, Box3D domain, U rho, Array<U, 3> velocity) {
  (void)domain;
  const double u[3] = {velocity[0], velocity[1], velocity[2]};
  hemo::gpu_lattice_equilibrium(lattice.gpu(), rho, u);
}
template <typename U, template <typename V> class Descriptor>
std::string getMultiBlockInfo(MultiBlockLattice3D<U, Descriptor>& lattice) { return hemo::gpu_lattice_info(lattice.gpu()); }

/* rank-0 streams and the process "MPI" view: one process per GPU, rank / size from the launcher's
 * environment (RANK / WORLD_SIZE / LOCAL_RANK of torchrun, or OMPI_COMM_WORLD_*) */
class Parallel_ostream {
 public:
  explicit Parallel_ostream(std::ostream& o) : os(o) {}
  template <typename V> Parallel_ostream& operator<<(const V& v);
  Parallel_ostream& operator<<(std::ostream& (*f)(std::ostream&));
 private:
  std::ostream& os;
};
extern Parallel_ostream pcout, pcerr;
/* file stream that only the main processor writes (Palabos io/parallelIO.h) */
class plb_ofstream {
 public:
  plb_ofstream() {}
  explicit plb_ofstream(const char* filename, std::ios_base::openmode mode = std::ios_base::out) { open(filename, mode); }
  void open(const char* filename, std::ios_base::openmode mode = std::ios_base::out);
  void close() { if (f.is_open()) f.close(); }
  bool is_open() { return f.is_open(); }
  template <typename V> plb_ofstream& operator<<(const V& v) { if (f.is_open()) f << v; return *this; }
  plb_ofstream& operator<<(std::ostream& (*fn)(std::ostream&)) { if (f.is_open()) f << fn; return *this; }
 private:
  std::ofstream f;
};
void plbInit(int* argc, char*** argv);
namespace global {
class MpiManager {
 public:
  int getRank() const; int getSize() const; int getLocalRank() const;
  bool isMainProcessor() const { return getRank() == 0; }
  void barrier();
};
MpiManager& mpi();
class Directories {
 public:
  void setOutputDir(const std::string& d) { out = d; }
  void setLogOutDir(const std::string& d) { log = d; }
  void setInputDir(const std::string& d) { in = d; }
  std::string getOutputDir() const { return out; }
  std::string getLogOutDir() const { return log; }
  std::string getInputDir() const { return in; }
 private:
  std::string out = "./tmp/", log = "./tmp/log/", in = "./";
};
Directories& directories();
}  // namespace global

template <typename V> Parallel_ostream& Parallel_ostream::operator<<(const V& v) { if (global::mpi().isMainProcessor()) os << v; return *this; }
inline Parallel_ostream& Parallel_ostream::operator<<(std::ostream& (*f)(std::ostream&)) { if (global::mpi().isMainProcessor()) os << f; return *this; }

}  // namespace plb

/* ================================================================================================== */
/* hemo::                                                                                             */
/* ================================================================================================== */
namespace hemo {

using plb::plint;
using plb::pluint;
using std::cout; using std::endl; using std::map; using std::string; using std::vector;
using plb::pcout; using plb::pcerr;

/* helper/array.h */
template <typename U, size_t n>
struct Array {
  U data[n];
  U& operator[](size_t i) { return data[i]; }
  const U& operator[](size_t i) const { return data[i]; }
  Array& operator+=(const Array& o) { for (size_t i = 0; i < n; i++) data[i] += o.data[i]; return *this; }
  Array& operator-=(const Array& o) { for (size_t i = 0; i < n; i++) data[i] -= o.data[i]; return *this; }
  Array& operator*=(U s) { for (size_t i = 0; i < n; i++) data[i] *= s; return *this; }
  Array& operator/=(U s) { for (size_t i = 0; i < n; i++) data[i] /= s; return *this; }
  Array operator/(U s) const { Array r = *this; r /= s; return r; }
  void resetToZero() { for (size_t i = 0; i < n; i++) data[i] = U(); }
  size_t size() const { return n; }
  Array operator+(const Array& o) const { Array r = *this; r += o; return r; }
  Array operator-(const Array& o) const { Array r = *this; r -= o; return r; }
  Array operator*(U s) const { Array r = *this; r *= s; return r; }
};
template <typename U> U dot(const Array<U, 3>& a, const Array<U, 3>& b) { return a[0]*b[0] + a[1]*b[1] + a[2]*b[2]; }
template <typename U> U norm(const Array<U, 3>& a) { return std::sqrt(dot(a, a)); }
template <typename U> Array<U, 3> crossProduct(const Array<U, 3>& a, const Array<U, 3>& b) {
  return Array<U, 3>{{a[1]*b[2] - a[2]*b[1], a[2]*b[0] - a[0]*b[2], a[0]*b[1] - a[1]*b[0]}};
}

/* ---- config/config.h ---------------------------------------------------------------------------- */
namespace xml { struct Node; }
class XMLElement {
 public:
  explicit XMLElement(const xml::Node* n) : orig(n) {}
  XMLElement operator[](const std::string& name) const;     /* throws std::invalid_argument when absent */
  template <typename V> V read() const {
    std::stringstream value(text());
    V ret = V();
    if (!(value >> ret)) std::cout << "Cannot convert value from XML element" << std::endl;
    return ret;
  }
  const xml::Node* getOrig() const { return orig; }
  std::string text() const;
 private:
  const xml::Node* orig;
};
class Config {
 public:
  bool checkpointed = false;
  explicit Config(const std::string& paramXmlFilename);      /* exits(1) when the file does not exist */
  ~Config();
  void reload(const std::string& paramXmlFilename);
  XMLElement operator[](const std::string& name) const;
  const xml::Node* root() const { return doc.get(); }
 private:
  void load(const std::string& f);
  std::unique_ptr<xml::Node> doc;
};
void loadDirectories(Config* cfg, bool edit_out_dir = true);

/* ---- helper/profiler.h: same key names; device operators report CUDA-event times ---------------------- */
class Profiler {
 public:
  explicit Profiler(const std::string& name_) : name(name_) {}
  void start(); void stop();
  double elapsed() const;                                    /* seconds since start() */
  void setDeviceTimers(hcg_ctx* ctx_) { ctx = ctx_; }
  void printStatistics();                                    /* to hlog */
  void outputStatistics();                                   /* to <logfile>.statistics */
  static std::string toString(double seconds);
 private:
  void render(std::ostream& o);
  std::string name; double t0 = 0, t_total = 0; bool running = false; hcg_ctx* ctx = nullptr;
};
struct ConfigValues {
  bool hemoCellInitialized = false;
  bool cellsDeletedInfo = false;
  std::string checkpointDirectory = "./checkpoint/";
  Profiler statistics = Profiler("HemoCell");
};
extern ConfigValues global;

/* ---- config/logfile.h ---------------------------------------------------------------------------- */
class Logfile {
 public:
  explicit Logfile(bool to_stdout_) : to_stdout(to_stdout_) {}
  template <typename V> Logfile& operator<<(const V& v) {
    if (plb::global::mpi().isMainProcessor()) { if (to_stdout) std::cout << v; if (file.is_open()) file << v; }
    return *this;
  }
  Logfile& operator<<(std::ostream& (*f)(std::ostream&)) {
    if (plb::global::mpi().isMainProcessor()) { if (to_stdout) std::cout << f; if (file.is_open()) file << f; }
    return *this;
  }
  void open(const std::string& path) { if (plb::global::mpi().isMainProcessor()) { file.close(); file.open(path, std::ios::app); } filename = path; }
  std::string filename;
 private:
  bool to_stdout; std::ofstream file;
};
extern Logfile hlog, hlogfile;

/* ---- mechanics/constantConversion.h -------------------------------------------------------------- */
class Parameters {
 public:
  static void lbm_base_parameters(Config& cfg);
  static void lbm_pipe_parameters(Config& cfg, int nY);
  static void lbm_pipe_parameters(Config& cfg, plb::MultiScalarField3D<int>* flagMatrix);   /* radius from the fluid area of the x0 slice */
  static void lbm_shear_parameters(Config& cfg, T nx);
  static void printParameters();
  static T dt, dx, dm, df, nu_p, rho_p, tau, re, nu_lbm, u_lbm_max, pipe_radius, kBT_p, kBT_lbm, shearrate_lbm, f_limit;
};

/* ---- helper/voxelizeDomain.h: STL geometry -> flag matrix (1 = fluid, 0 = solid) + the lattice size ------------------ */
void getFlagMatrixFromSTL(std::string meshFileName, plb::plint extendedEnvelopeWidth, plb::plint refDirLength, plb::plint refDir,
                          plb::VoxelizedDomain3D<T>*& voxelizedDomain, plb::MultiScalarField3D<int>*& flagMatrix, plb::plint blockSize, int particleEnvelope = 0);
template <class PtrVD, class PtrFM>   /* std::auto_ptr / std::unique_ptr overload of the case files */
void getFlagMatrixFromSTL(std::string meshFileName, plb::plint extendedEnvelopeWidth, plb::plint refDirLength, plb::plint refDir,
                          PtrVD& voxelizedDomain, PtrFM& flagMatrix, plb::plint blockSize, int particleEnvelope = 0) {
  plb::VoxelizedDomain3D<T>* vd = nullptr; plb::MultiScalarField3D<int>* fm = nullptr;
  getFlagMatrixFromSTL(meshFileName, extendedEnvelopeWidth, refDirLength, refDir, vd, fm, blockSize, particleEnvelope);
  voxelizedDomain = PtrVD(vd); flagMatrix = PtrFM(fm);
}

/* ---- core/hemoCellParticle.h (host view of one Lagrangian surface point; see HemoCellFields::getParticles) */
class HemoCellParticle {
 public:
  struct serializeValues_t {
    hemo::Array<T, 3> v, position, force, force_repulsion;
    plint cellId; uint16_t vertexId; unsigned int restime; unsigned char celltype;
  } sv;
  hemo::Array<T, 3>* force_volume = &sv.force; hemo::Array<T, 3>* force_bending = &sv.force;
  hemo::Array<T, 3>* force_link = &sv.force; hemo::Array<T, 3>* force_area = &sv.force;
  hemo::Array<T, 3>* force_visc = &sv.force; hemo::Array<T, 3>* force_inner_link = &sv.force;
};

/* ---- mechanics ------------------------------------------------------------------------------------ */
struct CommonCellConstantsView {                   /* mechanics/commonCellConstants.h:39-86 (read-only copy) */
  std::vector<hemo::Array<plint, 3>> triangle_list;
  std::vector<hemo::Array<plint, 2>> edge_list;
  std::vector<T> edge_length_eq_list, edge_angle_eq_list, triangle_area_eq_list, surface_patch_center_dist_eq_list;
  T volume_eq = 0, area_mean_eq = 0, edge_mean_eq = 0, angle_mean_eq = 0;
};
class CellMechanics {
 public:
  CellMechanics(HemoCellField& cellfield, Config& modelCfg_);
  virtual ~CellMechanics() {}
  /* Host-side force interface of the reference.  The built-in models below run as device kernels
   * (deviceModel() >= 0) and never get this call on the hot path. */
  virtual void ParticleMechanics(std::map<int, std::vector<HemoCellParticle*>>& particles_per_cell,
                                 const std::map<int, bool>& lpc, pluint ctype) = 0;
  virtual void statistics() = 0;
  virtual int deviceModel() const { return -1; }   /* HCG_MODEL_* of the kernel that replaces ParticleMechanics */
  T calculate_kLink(Config& cfg);
  T calculate_kBend(Config& cfg);
  T calculate_kVolume(Config& cfg);
  T calculate_kArea(Config& cfg);
  T calculate_etaM(Config& cfg);
  const CommonCellConstantsView cellConstants;
  Config& cfg;
 protected:
  HemoCellField& field_;
};
class RbcHighOrderModel : public CellMechanics {
 public:
  HemoCellField& cellField;
  const T k_volume, k_area, k_link, k_bend, eta_m;
  RbcHighOrderModel(Config& modelCfg_, HemoCellField& cellField_);
  void ParticleMechanics(std::map<int, std::vector<HemoCellParticle*>>&, const std::map<int, bool>&, pluint ctype) override;
  void statistics() override;
  int deviceModel() const override { return HCG_MODEL_RBC_HIGHORDER; }
};
class PltSimpleModel : public CellMechanics {
 public:
  HemoCellField& cellField;
  const T k_volume, k_area, k_link, k_bend, eta_m;
  PltSimpleModel(Config& modelCfg_, HemoCellField& cellField_);
  void ParticleMechanics(std::map<int, std::vector<HemoCellParticle*>>&, const std::map<int, bool>&, pluint ctype) override;
  void statistics() override;
  int deviceModel() const override { return HCG_MODEL_PLT_SIMPLE; }
};

/* ---- core/hemoCellField.h ------------------------------------------------------------------------ */
struct CellTypeImpl;
class HemoCellField {
 public:
  HemoCellField(HemoCellFields& cellFields_, const std::string& name_, unsigned int ctype_, int constructType);
  ~HemoCellField();
  std::string name;
  HemoCellFields& cellFields;
  vector<int> desiredOutputVariables;
  Config* materialCfg = nullptr;
  unsigned char ctype;
  int numVertex = 0;
  T volume = 0, volumeFractionOfLspPerNode = 0;
  unsigned int timescale = 1;
  unsigned int minimumDistanceFromSolid = 0;        /* [um]; an unsigned int in the reference too (core/hemoCellField.h:64): 0.5 um truncates to 0 */
  bool outputTriangles = false;
  vector<hemo::Array<plint, 3>> triangle_list;
  CellMechanics* mechanics = nullptr;
  int constructType;
  void setOutputVariables(const vector<int>& outputs);
  void statistics();
  int getNumberOfCells_Global();
  T getVolumeFraction();
  std::string getIdentifier() { return name; }
  CellTypeImpl* impl = nullptr;                     /* mesh + CommonCellConstants tables (host), device ctype id */
  /* MeshMetrics of the undeformed mesh in lattice units (core/hemoCellField.h meshmetric; the case files use getVolume / getSurface) */
  struct MeshMetrics {
    T volume = 0, surface = 0, meanLength = 0, maxLength = 0, minLength = 0;
    T getVolume() const { return volume; } T getSurface() const { return surface; }
    T getMeanLength() const { return meanLength; } T getMaxLength() const { return maxLength; } T getMinLength() const { return minLength; }
  };
  MeshMetrics* meshmetric = nullptr;
  hemo::Array<T, 6> getOriginalBoundingBox();       /* core/hemoCellField.cpp:149-165 */
};

/* ---- core/hemoCellFields.h ----------------------------------------------------------------------- */
class HemoCellFields {
 public:
  HemoCellFields(plb::MultiBlockLattice3D<T, DESCRIPTOR>& lattice_, unsigned int particleEnvelopeWidth, HemoCell& hemocell_);
  ~HemoCellFields();
  HemoCellField* addCellType(const std::string& name_, int constructType);
  HemoCellField* operator[](unsigned int index) { return cellFields[index]; }
  HemoCellField* operator[](const std::string& name);
  unsigned int size() { return (unsigned int)cellFields.size(); }
  /* the per-operator methods of core/hemoCellFields.h:101-158 (each maps onto one hcg_op_*) */
  void advanceParticles();
  void interpolateFluidVelocity();
  void spreadParticleForce();
  void applyRepulsionForce();
  void applyBoundaryRepulsionForce();
  void applyConstitutiveModel(bool forced = false);
  void syncEnvelopes();
  void deleteIncompleteCells(bool verbose = true) { (void)verbose; }   /* cells are whole by construction */
  void deleteNonLocalParticles(int envelope) { (void)envelope; }
  void populateBoundaryParticles() {}
  void separate_force_vectors() { separateForces = true; }
  void unify_force_vectors() { separateForces = false; }
  void calculateCommunicationStructure() {}
  void getParticles(vector<HemoCellParticle>& particles);    /* device -> host copy of every live LSP */
  /* knobs (core/hemoCellFields.h:182-208) */
  plb::MultiBlockLattice3D<T, DESCRIPTOR>* lattice;
  vector<int> desiredFluidOutputVariables;
  unsigned int particleVelocityUpdateTimescale = 1, repulsionTimescale = 1, boundaryRepulsionTimescale = 1;
  T repulsionConstant = 0, repulsionCutoff = 0, boundaryRepulsionConstant = 0, boundaryRepulsionCutoff = 0;
  pluint envelopeSize;
  int periodicity_limit[3] = {100, 100, 100};
  plint number_of_cells = 0;
  HemoCell& hemocell;
  bool separateForces = false;
  vector<HemoCellField*> cellFields;
  hcg_ctx* ctx();
};

/* ---- helper/preInlet.h --------------------------------------------------------------------------- */
}  // namespace hemo
enum Direction : int { Xpos, Xneg, Ypos, Yneg, Zpos, Zneg };
namespace hemo {
/* The reference splits the MPI ranks between the periodic pre-inlet and the main domain and couples them with messages.
 * Here ONE process holds both: the pre-inlet is a second device context (same GPU, or GPU $HEMOCELL_PREINLET_DEVICE)
 * that HemoCell::iterate() steps next to the main lattice, so `partOfpreInlet` is always false for the case file and the
 * calls the reference makes on its pre-inlet ranks (bounce-back fill, periodicity, driving force, loading its cells) happen
 * inside this class.  applyPreInlet() = hcg_preinlet_apply_velocity + hcg_preinlet_apply_cells (include/hemocell_gpu.h). */
inline plint cellsInBoundingBox(plb::Box3D const& box) { return std::abs((box.x1 - box.x0)*(box.y1 - box.y0)*(box.z1 - box.z0)); }
class PreInlet {
 public:
  PreInlet(HemoCell* hemocell_, plb::MultiScalarField3D<int>* flagMatrix_);
  ~PreInlet();
  plint getNumberOfNodes() { return cellsInBoundingBox(location); }
  void createBoundary();
  bool readNormalizedVelocities();
  void setDrivingForce();
  void setDrivingForceTimeDependent(double t);
  void calculateDrivingForce();
  double interpolate(vector<double>& xData, vector<double>& yData, double x, bool extrapolate);
  double average(vector<double> values);
  void applyPreInletVelocityBoundary();
  void applyPreInletParticleBoundary();
  void applyPreInlet() { applyPreInletVelocityBoundary(); applyPreInletParticleBoundary(); }
  void initializePreInletParticleBoundary();
  void initializePreInletVelocityBoundary();
  void initializePreInlet() { initializePreInletVelocityBoundary(); initializePreInletParticleBoundary(); }
  void autoPreinletFromBoundary(Direction);
  void preInletFromSlice(Direction direction_, plb::Box3D boundary);
  Direction direction = Direction::Zneg;
  plb::Box3D location, fluidInlet;
  int nProcs = 0;
  bool initialized = false;
  double drivingForce = 0.0, average_vel = 0.0, pulseEndTime = 1.0, pFrequency = 1.0;
  std::vector<double> normalizedVelocityTimes, normalizedVelocityValues;
  bool partOfpreInlet = false;
  int inflow_length = 0, preinlet_length = 0;
  HemoCell* hemocell;
  plb::MultiScalarField3D<int>* flagMatrix = nullptr;
  /* ---- B200 side ---- */
  GpuLattice* pre = nullptr;                      /* the pre-inlet's own lattice (local coordinates: global - location.{x0,y0,z0}) */
  hcg_ctx* preCtx();
  void createLattice(double omega);               /* HemoCell::initializeLattice */
  void iterate();                                 /* HemoCell::iterate: one step of the pre-inlet domain */
  int64_t cellsHandedOver = 0;
 private:
  void locate(plb::Box3D slice);
  void coupleNodes();
  void applyForce(double f);
  void saveCheckPoint(const std::string& dir);    /* the reference's PRE_lattice / PRE_particleField (core/hemoCellFields.cpp:297-314) */
  void loadCheckPoint(const std::string& dir);
  int axis() const { return (int)direction/2; }
  bool coupled = false, force_applied = false;
  friend class HemoCell;
};
void boundaryFromFlagMatrix(plb::MultiBlockLattice3D<T, DESCRIPTOR>* fluid, plb::MultiScalarField3D<int>* flagMatrix, bool partOfpreInlet);   /* helper/genericFunctions.cpp:138-163 */

/* ---- hemocell.h ---------------------------------------------------------------------------------- */
class HemoCell {
 public:
  HemoCell(char* configFileName, int argc, char* argv[]);
  ~HemoCell();
  void latticeEquilibrium(T rho, hemo::Array<T, 3> vel);
  void latticeEquilibrium(T rho, plb::Array<T, 3> vel) { latticeEquilibrium(rho, hemo::Array<T, 3>{{vel[0], vel[1], vel[2]}}); }
  void initializeCellfield();
  template <class Mechanics>
  void addCellType(std::string name, int constructType) {
    HemoCellField* cellfield = cellfields->addCellType(name, constructType);
    Mechanics* mechanics = new Mechanics(*cellfield->materialCfg, *cellfield);
    cellfield->mechanics = mechanics;
    registerCellType(cellfield);
    cellfield->statistics();
  }
  void setOutputs(std::string name, vector<int> outputs);
  void setFluidOutputs(vector<int> outputs);
  bool repulsionEnabled = false, boundaryRepulsionEnabled = false;
  void setRepulsion(T repulsionConstant, T repulsionCutoff);
  void setMaterialTimeScaleSeparation(std::string name, unsigned int separation);
  void setParticleVelocityUpdateTimeScaleSeparation(unsigned int separation);
  void setRepulsionTimeScaleSeperation(unsigned int separation);
  void enableBoundaryParticles(T boundaryRepulsionConstant, T boundaryRepulsionCutoff, unsigned int timestep = 1);
  void setInitialMinimumDistanceFromSolid(std::string name, T distance);
  void setSystemPeriodicity(unsigned int axis, bool bePeriodic);
  void setSystemPeriodicityLimit(unsigned int axis, int limit);
  void loadParticles();
  void loadCheckPoint();
  void saveCheckPoint();
  bool outputInSiUnits = true;
  void writeOutput();
  void iterate();
  void initializeLattice(const plb::MultiBlockManagement3D& management);
  plb::MultiBlockLattice3D<T, DESCRIPTOR>* lattice = nullptr;
  Config* cfg = nullptr;
  HemoCellFields* cellfields = nullptr;
  unsigned int iter = 0;
  PreInlet* preInlet = nullptr;                      /* owned (deleted with the HemoCell object, core/hemoCell.cpp:117-119) */
  bool partOfpreInlet = false;                       /* always false here: one process holds both domains (see PreInlet) */
  hcg_ctx* ctx();                                    /* the device context behind `lattice` (created by lattice->initialize()) */
 private:
  void registerCellType(HemoCellField* f);
  void sanityCheck();
  void pushSettings();
  bool sanityCheckDone = false, loadParticlesIsCalled = false;
  unsigned int lastOutputAt = 0; double lastOutput = 0;
  friend class PreInlet;
};

/* ---- helper/cellInfo.h, helper/fluidInfo.h ------------------------------------------------------------ */
struct CellInformation {
  hemo::Array<T, 3> position = {}; hemo::Array<T, 3> velocity = {};
  T volume = 0, area = 0, stretch = 0;
  hemo::Array<T, 6> bbox = {};
  pluint blockId = 0, cellType = 0; bool centerLocal = true; int base_cell_id = 0;
};
class CellInformationFunctionals {
 public:
  static map<int, CellInformation> info_per_cell;
  static void clear_list() { info_per_cell.clear(); }
  static void calculateCellVolume(HemoCell*);
  static void calculateCellArea(HemoCell*);
  static void calculateCellPosition(HemoCell*);
  static void calculateCellStretch(HemoCell*);            /* max pairwise vertex distance, helper/cellInfo.cpp:103-121 */
  static void calculateCellBoundingBox(HemoCell*);
  static void calculateCellInformation(HemoCell*);
  static pluint getTotalNumberOfCells(HemoCell*);
  static pluint getNumberOfCellsFromType(HemoCell*, std::string type);
};
/* helper/particleInfo.h: |v| and |force + force_repulsion| over the live LSPs */
struct ParticleStatistics { T min = 0, max = 0, avg = 0; pluint ncells = 0; };
class ParticleInfo {
 public:
  static ParticleStatistics calculateVelocityStatistics(HemoCell* hemocell);
  static ParticleStatistics calculateForceStatistics(HemoCell* hemocell);
};
/* helper/hemoCellStretch.h: opposite forces on the n outermost LSPs (along x) of the single cell in the domain */
class HemoCellStretch {
 public:
  HemoCellStretch(HemoCellField& cellfield_, unsigned int n_forced_lsps_, T external_force_);
  void applyForce();
  static vector<plint> lower_lsps, upper_lsps;      /* vertex ids */
  HemoCellField& cellfield;
  static unsigned int n_forced_lsps;
  static T external_force, scale;
};
void writeCellInfo_CSV(HemoCell& hemocell);         /* io/writeCellInfoCSV.h */
struct FluidStatistics { T min = 0, max = 0, avg = 0; pluint ncells = 0; };
class FluidInfo {
 public:
  static FluidStatistics calculateVelocityStatistics(HemoCell* hemocell);
};

/* helper/hemocellInit.hh:59-92 */
template <typename U, template <class V> class Descriptor>
void iniLatticeSquareCouette(plb::MultiBlockLattice3D<U, Descriptor>& lattice, plint nx, plint ny, plint nz,
                             plb::OnLatticeBoundaryCondition3D<U, Descriptor>& boundaryCondition, U shearRate) {
  plb::Box3D top(0, nx - 1, 0, ny - 1, nz - 1, nz - 1), bottom(0, nx - 1, 0, ny - 1, 0, 0);
  lattice.periodicity().toggle(0, true); lattice.periodicity().toggle(1, true); lattice.periodicity().toggle(2, false);
  boundaryCondition.setVelocityConditionOnBlockBoundaries(lattice, top);
  boundaryCondition.setVelocityConditionOnBlockBoundaries(lattice, bottom);
  const U vHalf = (nz - 1)*shearRate*0.5;
  plb::setBoundaryVelocity(lattice, top, plb::Array<U, 3>(-vHalf, 0.0, 0.0));
  plb::setBoundaryVelocity(lattice, bottom, plb::Array<U, 3>(vHalf, 0.0, 0.0));
  plb::setExternalVector(lattice, lattice.getBoundingBox(), Descriptor<U>::ExternalField::forceBeginsAt, plb::Array<U, 3>(0.0, 0.0, 0.0));
  lattice.initialize();
}

}  // namespace hemo

using namespace plb;   /* the reference's hemocell.h pulls plb:: into the case files the same way (palabos3D.h + using) */

#endif
