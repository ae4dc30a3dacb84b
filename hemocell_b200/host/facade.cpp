// The HemoCell C++ API surface (include/hemocell.h) implemented over the C ABI of
// libhemocell_gpu.so.  Host orchestration only: every operator call ends in an hcg_* entry point.
//   reference: core/hemoCell.cpp, core/hemoCellFields.cpp, core/hemoCellField.cpp, config/config.cpp,
//              config/logfile.cpp, mechanics/constantConversion.cpp, helper/cellInfo.cpp,
//              helper/fluidInfo.cpp, helper/profiler.cpp, io/writeCellInfoCSV.cpp
#include "hemo_mesh.h"          // hemo::host:: set-up code (before hemocell.h: its enum names are macros there)
#include "hemo_xml.h"
#include "hemo_h5.h"
#include "hemo_voxel.h"
#include "hemocell.h"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <functional>
#include <iomanip>
#include <thread>
#include <sys/stat.h>
#include <unistd.h>

namespace hemo {

// ------------------------------------------------------------------------------------------------
// process view: one process per GPU, rank / size from the launcher
// ------------------------------------------------------------------------------------------------
namespace {
int env_int(const char* a, const char* b, int dflt) {
  const char* v = getenv(a);
  if (!v && b) v = getenv(b);
  return v ? atoi(v) : dflt;
}
[[noreturn]] void fatal(const std::string& msg) {
  // reference convention: log, then exit(1) (e.g. core/hemoCell.cpp:75-79)
  std::cerr << msg << std::endl;
  hlog << msg << std::endl;
  exit(1);
}
void ck(hcg_ctx* c, hcg_status s, const char* what) {
  if (s == HCG_OK) return;
  fatal(std::string("(HemoCell) (GPU) ") + what + " failed: " + hcg_last_error(c));
}
bool file_exists(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0; }
void mkpath(const std::string& p) {
  std::string cur;
  for (size_t i = 0; i < p.size(); i++) {
    cur += p[i];
    if (p[i] == '/' || i + 1 == p.size()) mkdir(cur.c_str(), 0777);
  }
}
std::string zeroPadNumber(unsigned long n) { std::ostringstream o; o << std::setw(12) << std::setfill('0') << n; return o.str(); }   // helper/genericFunctions.h:63
double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// rank 0 -> all ranks of this launch through a file (all ranks share one box and one parent launcher)
void share_bytes(const char* tag, void* data, size_t n) {
  static int generation = 0;
  const int rank = env_int("RANK", "OMPI_COMM_WORLD_RANK", 0);
  const char* dir = getenv("HEMOCELL_RENDEZVOUS_DIR");
  std::ostringstream p;
  p << (dir ? dir : "/tmp") << "/hemocell_" << tag << "_" << (getenv("MASTER_PORT") ? getenv("MASTER_PORT") : "0") << "_" << getppid() << "_" << generation++;
  const std::string path = p.str();
  if (rank == 0) {
    std::ofstream f(path + ".tmp", std::ios::binary);
    f.write((const char*)data, (std::streamsize)n); f.close();
    rename((path + ".tmp").c_str(), path.c_str());
  } else {
    const double t0 = now_s();
    while (!file_exists(path)) {
      if (now_s() - t0 > 300) fatal("(HemoCell) timed out waiting for rank 0 at " + path);
      std::this_thread::sleep_for(std::chrono::milliseconds(20));
    }
    std::ifstream f(path, std::ios::binary);
    f.read((char*)data, (std::streamsize)n);
  }
}
}  // namespace

// ------------------------------------------------------------------------------------------------
// GpuLattice: host-side domain description recorded by the plb:: calls + the device context
// ------------------------------------------------------------------------------------------------
class GpuLattice {
 public:
  int nx, ny, nz;
  double omega;
  bool periodic[3] = {false, false, false};
  std::vector<uint8_t> flags;                 // global, z + nz*(y + ny*x)
  double bc[6][3];
  double body[3] = {0, 0, 0};
  std::vector<double> bodyfield;              // optional per-node driving force, global, [3][N]; empty = uniform `body`
  bool bodyfield_dirty = false; int bodyfield_generation = -1;
  double eq_rho = 1.0, eq_u[3] = {0, 0, 0};
  bool eq_pending = true, body_pending = true;
  hcg_ctx* ctx = nullptr;
  bool created_periodic[3] = {false, false, false};
  bool has_cells = false;
  int generation = 0;                         // bumped whenever a new device context is created
  std::map<int64_t, std::array<double, 4>> bcn;   // per-node (u, rho) of the Zou-He velocity / pressure nodes, keyed by global node index
  bool bcn_dirty = false;
  bool solo = false;                          // a single-rank context whatever the launcher says (the pre-inlet domain)
  int device_override = -1;
  GpuLattice* companion = nullptr;            // the pre-inlet lattice that warms up and iterates along with this one

  GpuLattice(int nx_, int ny_, int nz_, double omega_) : nx(nx_), ny(ny_), nz(nz_), omega(omega_), flags((size_t)nx_*ny_*nz_, HCG_FLUID) {
    memset(bc, 0, sizeof(bc));
  }
  ~GpuLattice() { if (ctx) hcg_destroy(ctx); }

  int64_t idx(int x, int y, int z) const { return (int64_t)z + (int64_t)nz*((int64_t)y + (int64_t)ny*x); }
  int rank() const { return solo ? 0 : plb::global::mpi().getRank(); }
  int size() const { return solo ? 1 : plb::global::mpi().getSize(); }
  // x-slab of this rank (hcg_slab: the first nx % size ranks own one plane more)
  int nxl() const { int32_t x0, n; hcg_slab(nx, rank(), size(), &x0, &n); return n; }
  int x0() const { int32_t x0, n; hcg_slab(nx, rank(), size(), &x0, &n); return x0; }

  // create the context on first device use; re-create if the periodicity was toggled afterwards
  void materialize() {
    if (ctx && !memcmp(periodic, created_periodic, sizeof(periodic))) { flush(); return; }
    if (ctx) {
      if (has_cells) fatal("(HemoCell) (Periodicity) the periodicity cannot change once particles are loaded");
      hcg_destroy(ctx); ctx = nullptr;
    }
    hcg_domain d;
    d.nx = nx; d.ny = ny; d.nz = nz;
    for (int k = 0; k < 3; k++) d.periodic[k] = periodic[k];
    d.tau = 1.0/omega;
    // more ranks than GPUs on the box (or HEMOCELL_COMM=host): the ranks share the GPUs and talk through the host-staged
    // communicator (hcg_comm_init_local) - NCCL refuses two ranks on one device
    const int ndev = hcg_device_count();
    const char* comm_env = getenv("HEMOCELL_COMM");
    const bool host_comm = (comm_env && !strcmp(comm_env, "host")) || (ndev > 0 && size() > ndev);
    d.device = device_override >= 0 ? device_override : (ndev > 0 ? plb::global::mpi().getLocalRank() % ndev : 0);
    d.rank = rank(); d.n_ranks = size();
    hcg_status s = hcg_create(&d, &ctx);
    if (s != HCG_OK) fatal(std::string("(HemoCell) (GPU) cannot create the device context: ") + hcg_last_error(ctx));
    if (d.n_ranks > 1) {
      unsigned char id[128];
      rendezvous(id);
      if (host_comm) ck(ctx, hcg_comm_init_local(ctx, id), "hcg_comm_init_local");
      else ck(ctx, hcg_comm_init(ctx, id), "hcg_comm_init");
    }
    memcpy(created_periodic, periodic, sizeof(periodic));
    generation++;
    const int64_t P = (int64_t)ny*nz;
    ck(ctx, hcg_lattice_set_flags(ctx, flags.data() + (int64_t)x0()*P), "hcg_lattice_set_flags");
    for (int o = 0; o < 6; o++) ck(ctx, hcg_lattice_set_bc_velocity(ctx, o, bc[o]), "hcg_lattice_set_bc_velocity");
    flags_dirty = false;
    eq_pending = true; body_pending = true; bcn_dirty = !bcn.empty();
    flush();
    if (!solo) { global.statistics.setDeviceTimers(ctx); hcg_timers_enable(ctx, 1); }
  }
  void flush() {
    if (!ctx) return;
    if (flags_dirty) {
      const int64_t P = (int64_t)ny*nz;
      ck(ctx, hcg_lattice_set_flags(ctx, flags.data() + (int64_t)x0()*P), "hcg_lattice_set_flags");
      for (int o = 0; o < 6; o++) ck(ctx, hcg_lattice_set_bc_velocity(ctx, o, bc[o]), "hcg_lattice_set_bc_velocity");
      flags_dirty = false;
    }
    if (bcn_dirty) {
      const int64_t P = (int64_t)ny*nz, lo = (int64_t)x0()*P, hi = lo + (int64_t)nxl()*P;
      std::vector<int64_t> idx; std::vector<double> val;
      for (auto& kv : bcn) if (kv.first >= lo && kv.first < hi) { idx.push_back(kv.first - lo); val.insert(val.end(), kv.second.begin(), kv.second.end()); }
      ck(ctx, hcg_lattice_set_bc_nodes(ctx, (int64_t)idx.size(), idx.data(), val.data()), "hcg_lattice_set_bc_nodes");
      bcn_dirty = false;
    }
    if (eq_pending) { ck(ctx, hcg_lattice_init_equilibrium(ctx, eq_rho, eq_u), "hcg_lattice_init_equilibrium"); eq_pending = false; }
    if (body_pending) {
      if (bodyfield.empty()) ck(ctx, hcg_lattice_set_body_force(ctx, body), "hcg_lattice_set_body_force");
      else if (bodyfield_dirty || bodyfield_generation != generation) {
        // this rank's slab of the global field, component-major
        const size_t N = (size_t)nx*ny*nz, Nl = (size_t)nxl()*ny*nz, off = (size_t)x0()*ny*nz;
        std::vector<double> slab(3*Nl);
        for (int k = 0; k < 3; k++) std::copy(bodyfield.begin() + k*N + off, bodyfield.begin() + k*N + off + Nl, slab.begin() + k*Nl);
        ck(ctx, hcg_lattice_set_body_force_field(ctx, slab.data()), "hcg_lattice_set_body_force_field");
        bodyfield_dirty = false; bodyfield_generation = generation;
      }
      body_pending = false;
    }
  }
  void touchFlags() { flags_dirty = true; }

 private:
  bool flags_dirty = false;
  void rendezvous(unsigned char id[128]) {
    if (rank() == 0) ck(nullptr, hcg_comm_unique_id(id), "hcg_comm_unique_id");
    share_bytes("nccl", id, 128);
  }
};

// ------------------------------------------------------------------------------------------------
// globals
// ------------------------------------------------------------------------------------------------
ConfigValues global;
Logfile hlog(true), hlogfile(false);
T Parameters::dt = 0, Parameters::dx = 0, Parameters::dm = 0, Parameters::df = 0, Parameters::nu_p = 0, Parameters::rho_p = 0,
  Parameters::tau = 0, Parameters::re = 0, Parameters::nu_lbm = 0, Parameters::u_lbm_max = 0, Parameters::pipe_radius = 0,
  Parameters::kBT_p = 0, Parameters::kBT_lbm = 0, Parameters::shearrate_lbm = 0, Parameters::f_limit = 0;
map<int, CellInformation> CellInformationFunctionals::info_per_cell;

// ------------------------------------------------------------------------------------------------
// Config (config/config.cpp:28-84)
// ------------------------------------------------------------------------------------------------
std::string XMLElement::text() const { return orig ? orig->text : std::string(); }
XMLElement XMLElement::operator[](const std::string& name) const {
  const xml::Node* child = orig ? orig->firstChild(name) : nullptr;
  if (!child) throw std::invalid_argument("XML child " + name + " does not exist.");
  return XMLElement(child);
}
Config::Config(const std::string& f) { load(f); }
Config::~Config() {}
void Config::reload(const std::string& f) { load(f); }
void Config::load(const std::string& f) {
  if (!file_exists(f)) { pcout << f + " is not an existing config file, exiting ..." << std::endl; exit(1); }
  try { doc = xml::parseFile(f); }
  catch (std::exception& e) { pcout << f << ": " << e.what() << ", exiting ..." << std::endl; exit(1); }
  // fresh start or checkpointed run: the first element is <Checkpoint> or <hemocell>
  checkpointed = !doc->children.empty() && doc->children[0]->name == "Checkpoint";
}
XMLElement Config::operator[](const std::string& name) const {
  if (checkpointed) return XMLElement(doc.get())["Checkpoint"]["hemocell"][name];
  return XMLElement(doc.get())["hemocell"][name];
}

void loadDirectories(Config* cfg, bool edit_out_dir) {
  auto& dirs = plb::global::directories();
  dirs.setInputDir("./");
  if (edit_out_dir) {
    std::string outDir;
    try {
      outDir = (*cfg)["parameters"]["outputDirectory"].read<std::string>();
      while (!outDir.empty() && outDir.back() == '/') outDir.pop_back();
      if (outDir.empty() || outDir[0] != '/') outDir = "./" + outDir;
    } catch (std::invalid_argument&) { outDir = "./tmp"; }
    // never overwrite an earlier run: tmp, tmp_0, tmp_1, ... (rank 0 decides, the others follow)
    std::string chosen = outDir;
    if (plb::global::mpi().isMainProcessor()) {
      for (int i = 0; file_exists(chosen); i++) chosen = outDir + "_" + std::to_string(i);
      mkpath(chosen + "/hdf5/");
    }
    if (plb::global::mpi().getSize() > 1) {
      char buf[1024] = {0};
      strncpy(buf, chosen.c_str(), sizeof(buf) - 1);
      share_bytes("outdir", buf, sizeof(buf));
      chosen = buf;
    }
    dirs.setOutputDir(chosen + "/");
  }
  try { dirs.setLogOutDir(dirs.getOutputDir() + "/" + (*cfg)["parameters"]["logDirectory"].read<std::string>() + "/"); }
  catch (std::invalid_argument&) { dirs.setLogOutDir(dirs.getOutputDir() + "/log/"); }
  std::string logfilename = "logfile";
  try { logfilename = (*cfg)["parameters"]["logFile"].read<std::string>(); } catch (std::invalid_argument&) {}
  if (plb::global::mpi().isMainProcessor()) mkpath(dirs.getLogOutDir());
  // logfile, logfile.0, logfile.1, ... (config/config.cpp:120-135)
  std::string base = dirs.getLogOutDir() + logfilename, name = base;
  for (int i = 0; file_exists(name); i++) name = base + "." + std::to_string(i);
  hlog.open(name); hlogfile.open(name);
  // config/config.cpp:168-173
  try { global.checkpointDirectory = dirs.getOutputDir() + "/" + (*cfg)["parameters"]["checkpointDirectory"].read<std::string>() + "/"; }
  catch (std::invalid_argument&) { global.checkpointDirectory = dirs.getOutputDir() + "/checkpoint/"; }
}

// ------------------------------------------------------------------------------------------------
// Parameters (mechanics/constantConversion.cpp:36-110)
// ------------------------------------------------------------------------------------------------
void Parameters::lbm_base_parameters(Config& cfg) {
  host::Parameters p;
  p.lbm_base_parameters(cfg["domain"]["dx"].read<T>(), cfg["domain"]["dt"].read<T>(), cfg["domain"]["nuP"].read<T>(),
                        cfg["domain"]["rhoP"].read<T>(), cfg["domain"]["kBT"].read<T>());
  if (cfg["domain"]["dt"].read<T>() < 0.0) hlog << "(HemoCell) dt is set to *auto*. Tau will be set to 1!" << std::endl;
  dt = p.dt; dx = p.dx; nu_p = p.nu_p; rho_p = p.rho_p; kBT_p = p.kBT_p; tau = p.tau; nu_lbm = p.nu_lbm;
  dm = p.dm; df = p.df; f_limit = p.f_limit; kBT_lbm = p.kBT_lbm;
}
void Parameters::lbm_pipe_parameters(Config& cfg, plb::MultiScalarField3D<int>* sf) {
  // mechanics/constantConversion.cpp:61-73: radius of the circle with the fluid area of the first x slice
  lbm_base_parameters(cfg);
  re = cfg["domain"]["Re"].read<T>();
  plb::Box3D d = sf->getBoundingBox(); d.x1 = d.x0;
  const T fluidArea = plb::computeSum<int>(*sf, d);
  pcout << fluidArea << endl;
  pipe_radius = std::sqrt(fluidArea/PI);
  hlog << "(Parameters) The channel has a calculated radius of " << pipe_radius << " LU, assuming a perfect circle cross-section." << endl;
  u_lbm_max = re*nu_lbm/(pipe_radius*2);
}
void getFlagMatrixFromSTL(std::string meshFileName, plb::plint extendedEnvelopeWidth, plb::plint refDirLength, plb::plint refDir,
                          plb::VoxelizedDomain3D<T>*& voxelizedDomain, plb::MultiScalarField3D<int>*& flagMatrix, plb::plint blockSize, int particleEnvelope) {
  (void)blockSize; (void)particleEnvelope;            // one x-slab per GPU rank: no sparse block structure to tune
  host::VoxelizedSTL v;
  try { v = host::voxelizeSTL(meshFileName, (int)refDirLength, (int)refDir); }
  catch (std::exception& e) { hlog << e.what() << endl; exit(1); }
  voxelizedDomain = new plb::VoxelizedDomain3D<T>(v.nx, v.ny, v.nz, extendedEnvelopeWidth);
  flagMatrix = new plb::MultiScalarField3D<int>(v.nx, v.ny, v.nz, 0);
  for (int x = 0; x < v.nx; x++) for (int y = 0; y < v.ny; y++) for (int z = 0; z < v.nz; z++) {
    const int f = v.flag[(size_t)z + (size_t)v.nz*((size_t)y + (size_t)v.ny*x)];
    flagMatrix->get(x, y, z) = f; voxelizedDomain->getVoxelMatrix().get(x, y, z) = f ? 3 : 1;   // voxelFlag::inside / outside
  }
  hlog << "(main) Voxelisation is done. Resulting domain parameters are: " << endl;
  hlog << "Size of the multi-block:     " << v.nx << "-by-" << v.ny << "-by-" << v.nz << endl;
}
void Parameters::lbm_pipe_parameters(Config& cfg, int nY) {
  lbm_base_parameters(cfg);
  re = cfg["domain"]["Re"].read<T>();
  pipe_radius = nY;
  hlog << "(Parameters) The channel has a predefined radius of " << pipe_radius << " LU." << std::endl;
  u_lbm_max = re * nu_lbm / (pipe_radius*2);
}
void Parameters::lbm_shear_parameters(Config& cfg, T nx) {
  lbm_base_parameters(cfg);
  const T shearrate_p = cfg["domain"]["shearrate"].read<T>();
  re = (nx * (shearrate_p * (nx*0.5))) / nu_p;
  shearrate_lbm = shearrate_p*dt;
  u_lbm_max = shearrate_lbm;
}
void Parameters::printParameters() {
  hlog << "(HemoCell) System parameters:" << std::endl;
  hlog << "\t dx: \t" << dx << std::endl;
  hlog << "\t dt: \t" << dt << std::endl;
  hlog << "\t dm: \t" << dm << std::endl;
  hlog << "\t dN: \t" << df << std::endl;
  hlog << "\t tau: \t" << tau << std::endl;
  hlog << "\t nu_lbm: \t" << nu_lbm << std::endl;
  hlog << "\t u_lb_max: \t" << u_lbm_max << std::endl;
  hlog << "\t f_limit: \t" << f_limit << std::endl;
}

// ------------------------------------------------------------------------------------------------
// Profiler (helper/profiler.cpp:156-180): wall clock for the run, CUDA-event times per operator
// ------------------------------------------------------------------------------------------------
void Profiler::start() { t0 = now_s(); running = true; }
void Profiler::stop() { if (running) { t_total += now_s() - t0; running = false; } }
double Profiler::elapsed() const { return t_total + (running ? now_s() - t0 : 0.0); }
std::string Profiler::toString(double s) { std::ostringstream o; o << std::fixed << std::setprecision(6) << s; return o.str(); }
void Profiler::render(std::ostream& o) {
  o << name << ": " << toString(elapsed()) << " s" << std::endl;
  if (!ctx) return;
  int32_t n = 0;
  hcg_timers(ctx, nullptr, &n);
  std::vector<hcg_timer> t((size_t)std::max(n, 1));
  hcg_timers(ctx, t.data(), &n);
  o << "  iterate (device operators, CUDA events)" << std::endl;
  for (int i = 0; i < n; i++)
    o << "    " << std::left << std::setw(34) << t[i].name << " " << toString(t[i].ms_total*1e-3) << " s  (" << t[i].calls << " calls)" << std::endl;
}
void Profiler::printStatistics() { std::ostringstream o; render(o); hlog << o.str(); }
void Profiler::outputStatistics() {
  if (!plb::global::mpi().isMainProcessor() || hlog.filename.empty()) return;
  std::ofstream f(hlog.filename + ".statistics");
  render(f);
}

// ------------------------------------------------------------------------------------------------
// cell types
// ------------------------------------------------------------------------------------------------
struct CellTypeImpl {
  host::CellTypeTables tables;
  host::MaterialModel material;
  int device_ctype = -1;
  int device_generation = -1;                 // GpuLattice::generation the type was uploaded to
  int64_t n_cells_loaded = 0;
  bool host_model = false;                    // user subclass of CellMechanics without a device kernel: ParticleMechanics runs on the host
};

static host::MaterialModel read_material(Config& m, int constructType) {
  host::MaterialModel mm;
  XMLElement mat = m["MaterialModel"];
  mm.kBend = mat["kBend"].read<T>(); mm.kVolume = mat["kVolume"].read<T>(); mm.kArea = mat["kArea"].read<T>();
  mm.kLink = mat["kLink"].read<T>();
  try { mm.eta_m = mat["eta_m"].read<T>(); } catch (std::invalid_argument&) { mm.eta_m = 0; }
  mm.radius = mat["radius"].read<T>();
  mm.minNumTriangles = (int)mat["minNumTriangles"].read<T>();
  if (constructType == ELLIPSOID_FROM_SPHERE) mm.aspectRatio = mat["aspectRatio"].read<T>();
  try { mm.volume = mat["Volume"].read<T>(); } catch (std::invalid_argument&) {}
  // <InnerEdges><Edge> a b </Edge>...</InnerEdges> (mechanics/commonCellConstants.cpp:359-375)
  try {
    XMLElement ie = mat["InnerEdges"];
    for (auto& ch : ie.getOrig()->children) {
      if (ch->name != "Edge") continue;
      std::stringstream ss(ch->text); int a, b;
      if (ss >> a >> b) mm.innerEdges.push_back({a, b});
    }
  } catch (std::invalid_argument&) {}
  return mm;
}

HemoCellField::HemoCellField(HemoCellFields& cellFields_, const std::string& name_, unsigned int ctype_, int constructType_)
    : name(name_), cellFields(cellFields_), desiredOutputVariables({OUTPUT_POSITION}), ctype((unsigned char)ctype_), constructType(constructType_) {
  if (ctype_ > 255) fatal("(HemoCell) (AddCellType) more celltypes than UCHAR_MAX (255) added, please convert celltype to int or add less celltypes");
  if (constructType != RBC_FROM_SPHERE && constructType != ELLIPSOID_FROM_SPHERE)
    fatal("(HemoCell) (AddCellType) only RBC_FROM_SPHERE and ELLIPSOID_FROM_SPHERE meshes are built in");
  materialCfg = new Config(name + ".xml");
  impl = new CellTypeImpl();
  impl->material = read_material(*materialCfg, constructType);
  host::Parameters p;
  p.lbm_base_parameters(param::dx, param::dt < 0 ? -1.0 : param::dt, param::nu_p, param::rho_p, param::kBT_p);
  // model id is fixed once the mechanics object exists (HemoCell::addCellType); tables do not depend on it
  impl->tables.build(HCG_MODEL_RBC_HIGHORDER, constructType, impl->material, p);
  numVertex = impl->tables.mesh.getNumVertices();
  meshmetric = new MeshMetrics();
  meshmetric->volume = impl->tables.mesh.getVolume(); meshmetric->surface = impl->tables.mesh.getSurface();
  {
    const auto& el = impl->tables.cc.edge_length_eq_list;
    if (!el.empty()) { meshmetric->meanLength = impl->tables.cc.edge_mean_eq; meshmetric->maxLength = *std::max_element(el.begin(), el.end()); meshmetric->minLength = *std::min_element(el.begin(), el.end()); }
  }
  for (auto& t : impl->tables.cc.triangle_list) triangle_list.push_back({{t[0], t[1], t[2]}});
  volume = impl->material.volume;
  if (volume > 0) volumeFractionOfLspPerNode = (volume/numVertex)/std::pow(param::dx*1e6, 3);
  else hlog << "(HemoCell) (WARNING) (AddCellType) Volume of celltype " << name << " not present, volume set to zero" << endl;
}
HemoCellField::~HemoCellField() { delete mechanics; delete materialCfg; delete impl; delete meshmetric; }
hemo::Array<T, 6> HemoCellField::getOriginalBoundingBox() {
  host::Vec3 lo, hi; impl->tables.mesh.boundingBox(lo, hi);
  return hemo::Array<T, 6>{{lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]}};
}
void HemoCellField::setOutputVariables(const vector<int>& outputs) {
  desiredOutputVariables = outputs;
  auto it = std::find(desiredOutputVariables.begin(), desiredOutputVariables.end(), OUTPUT_TRIANGLES);
  if (it != desiredOutputVariables.end()) { desiredOutputVariables.erase(it); outputTriangles = true; } else outputTriangles = false;
}
void HemoCellField::statistics() {
  hlog << "Cellfield  (+ material model) of " << name << endl;
  hlog << "  timescale separation: " << timescale << endl;
  if (mechanics) mechanics->statistics();
}
int HemoCellField::getNumberOfCells_Global() { return (int)CellInformationFunctionals::getNumberOfCellsFromType(&cellFields.hemocell, name); }
T HemoCellField::getVolumeFraction() {
  auto box = cellFields.lattice->getBoundingBox();
  return getNumberOfCells_Global()*volume/(box.nCells()*std::pow(param::dx*1e6, 3));
}

// CellMechanics --------------------------------------------------------------------------------------
static CommonCellConstantsView make_view(HemoCellField& f) {
  CommonCellConstantsView v;
  const host::CommonCellConstants& cc = f.impl->tables.cc;
  for (auto& t : cc.triangle_list) v.triangle_list.push_back({{t[0], t[1], t[2]}});
  for (auto& e : cc.edge_list) v.edge_list.push_back({{e[0], e[1]}});
  v.edge_length_eq_list = cc.edge_length_eq_list; v.edge_angle_eq_list = cc.edge_angle_eq_list;
  v.triangle_area_eq_list = cc.triangle_area_eq_list; v.surface_patch_center_dist_eq_list = cc.surface_patch_center_dist_eq_list;
  v.volume_eq = cc.volume_eq; v.area_mean_eq = cc.area_mean_eq; v.edge_mean_eq = cc.edge_mean_eq; v.angle_mean_eq = cc.angle_mean_eq;
  return v;
}
CellMechanics::CellMechanics(HemoCellField& cellfield, Config& modelCfg_) : cellConstants(make_view(cellfield)), cfg(modelCfg_), field_(cellfield) {}
T CellMechanics::calculate_kLink(Config& c) { return c["MaterialModel"]["kLink"].read<T>() * param::kBT_lbm/(7.5e-9/param::dx); }
T CellMechanics::calculate_kBend(Config& c) { return c["MaterialModel"]["kBend"].read<T>() * param::kBT_lbm/(5e-7/param::dx); }
T CellMechanics::calculate_kVolume(Config& c) {
  return c["MaterialModel"]["kVolume"].read<T>() * (1280.0/cellConstants.triangle_list.size()) * param::kBT_lbm/(5e-7/param::dx);
}
T CellMechanics::calculate_kArea(Config& c) {
  return c["MaterialModel"]["kArea"].read<T>() * (1280.0/cellConstants.triangle_list.size()) * param::kBT_lbm/(5e-7/param::dx);
}
T CellMechanics::calculate_etaM(Config& c) {
  T eta = 0; try { eta = c["MaterialModel"]["eta_m"].read<T>(); } catch (std::invalid_argument&) {}
  return eta * param::dx/param::dt/param::df;
}

static void device_only(const char* model) {
  fatal(std::string("(HemoCell) (") + model + ") ParticleMechanics(map<...>) is the reference's host interface; this model runs as the device kernel k_mechanics "
        "inside HemoCellFields::applyConstitutiveModel()");
}
RbcHighOrderModel::RbcHighOrderModel(Config& modelCfg_, HemoCellField& cellField_)
    : CellMechanics(cellField_, modelCfg_), cellField(cellField_), k_volume(calculate_kVolume(modelCfg_)), k_area(calculate_kArea(modelCfg_)),
      k_link(calculate_kLink(modelCfg_)), k_bend(calculate_kBend(modelCfg_)), eta_m(calculate_etaM(modelCfg_)) {}
void RbcHighOrderModel::ParticleMechanics(std::map<int, std::vector<HemoCellParticle*>>&, const std::map<int, bool>&, pluint) { device_only("RbcHighOrderModel"); }
void RbcHighOrderModel::statistics() {
  hlog << "(Cell-mechanics model) High Order model parameters for " << cellField.name << " cellfield" << std::endl;
  hlog << "\t k_link:   " << k_link << std::endl; hlog << "\t k_area:   " << k_area << std::endl;
  hlog << "\t k_bend: : " << k_bend << std::endl; hlog << "\t k_volume: " << k_volume << std::endl;
  hlog << "\t eta_m:    " << eta_m << std::endl;
}
PltSimpleModel::PltSimpleModel(Config& modelCfg_, HemoCellField& cellField_)
    : CellMechanics(cellField_, modelCfg_), cellField(cellField_), k_volume(calculate_kVolume(modelCfg_)), k_area(calculate_kArea(modelCfg_)),
      k_link(calculate_kLink(modelCfg_)), k_bend(calculate_kBend(modelCfg_)), eta_m(calculate_etaM(modelCfg_)) {}
void PltSimpleModel::ParticleMechanics(std::map<int, std::vector<HemoCellParticle*>>&, const std::map<int, bool>&, pluint) { device_only("PltSimpleModel"); }
void PltSimpleModel::statistics() {
  hlog << "(Cell-mechanics model) Reduced-model parameters for " << cellField.name << " cellfield" << std::endl;
  hlog << "\t k_link:   " << k_link << std::endl; hlog << "\t k_area:   " << k_area << std::endl;
  hlog << "\t k_bend: : " << k_bend << std::endl; hlog << "\t k_volume: " << k_volume << std::endl;
  hlog << "\t eta_m:    " << eta_m << std::endl;
}

// HemoCellFields ---------------------------------------------------------------------------------------
HemoCellFields::HemoCellFields(plb::MultiBlockLattice3D<T, DESCRIPTOR>& lattice_, unsigned int particleEnvelopeWidth, HemoCell& hemocell_)
    : lattice(&lattice_), envelopeSize(particleEnvelopeWidth), hemocell(hemocell_) {}
HemoCellFields::~HemoCellFields() { for (auto* f : cellFields) delete f; }
hcg_ctx* HemoCellFields::ctx() { return hemocell.ctx(); }
HemoCellField* HemoCellFields::addCellType(const std::string& name_, int constructType) {
  HemoCellField* f = new HemoCellField(*this, name_, (unsigned int)cellFields.size(), constructType);
  cellFields.push_back(f);
  return f;
}
HemoCellField* HemoCellFields::operator[](const std::string& name) {
  for (auto* f : cellFields) if (f->name == name) return f;
  fatal("(HemoCell) (CellField) cell type " + name + " does not exist");
}
void HemoCellFields::advanceParticles() { ck(ctx(), hcg_op_advance(ctx()), "advanceParticles"); }
void HemoCellFields::interpolateFluidVelocity() { ck(ctx(), hcg_op_interpolate(ctx()), "interpolateFluidVelocity"); }
void HemoCellFields::spreadParticleForce() { ck(ctx(), hcg_op_spread(ctx()), "spreadParticleForce"); }
void HemoCellFields::applyRepulsionForce() { ck(ctx(), hcg_op_repulsion(ctx()), "applyRepulsionForce"); }
void HemoCellFields::applyBoundaryRepulsionForce() { ck(ctx(), hcg_op_wall_repulsion(ctx()), "applyBoundaryRepulsionForce"); }
// CellMechanics::ParticleMechanics of the cell types whose model has no device kernel (a user subclass, mechanics/cellMechanics.h:45):
// the reference's call sequence of HemoCellParticleField::applyConstitutiveModel (core/hemoCellParticleField.cpp:633-675) on host
// copies of the particles - download positions / velocities, zero the type's forces, call the model, upload the forces.  Slow path
// by construction (two PCIe round trips per material step); the built-in models never take it.
static void host_constitutive_model(HemoCellFields& cf, bool forced, plint iter) {
  bool any = false;
  for (auto* f : cf.cellFields) any = any || (f->impl->host_model && (forced || iter % f->timescale == 0));
  if (!any) return;
  hcg_ctx* c = cf.ctx();
  int64_t nc = 0, np = 0;
  ck(c, hcg_cells_capacity(c, &nc, &np), "hcg_cells_capacity");
  if (np == 0) return;
  std::vector<double> pos(3*np), vel(3*np), frc(3*np), frep(3*np);
  ck(c, hcg_cells_download(c, HCG_P_POS, pos.data()), "download"); ck(c, hcg_cells_download(c, HCG_P_VEL, vel.data()), "download");
  ck(c, hcg_cells_download(c, HCG_P_FORCE, frc.data()), "download"); ck(c, hcg_cells_download(c, HCG_P_FREP, frep.data()), "download");
  std::vector<int64_t> ids(nc); std::vector<int32_t> types(nc); std::vector<uint8_t> alive(nc);
  ck(c, hcg_cells_info(c, ids.data(), types.data(), alive.data()), "hcg_cells_info");
  for (auto* f : cf.cellFields) {
    if (!f->impl->host_model || !(forced || iter % f->timescale == 0)) continue;
    std::vector<HemoCellParticle> store;
    std::vector<int64_t> first;                           // first particle (device order) of each cell in `store`
    int64_t p = 0;
    for (int64_t k = 0; k < nc; k++) {
      const int V = cf.cellFields[types[k]]->numVertex;
      if (types[k] == f->impl->device_ctype && alive[k] && ids[k] >= 0) {
        first.push_back(p);
        for (int v = 0; v < V; v++) {
          HemoCellParticle q;
          for (int d = 0; d < 3; d++) { q.sv.position[d] = pos[3*(p+v)+d]; q.sv.v[d] = vel[3*(p+v)+d]; q.sv.force[d] = 0.0; q.sv.force_repulsion[d] = frep[3*(p+v)+d]; }
          q.sv.cellId = ids[k]; q.sv.vertexId = (uint16_t)v; q.sv.restime = 0; q.sv.celltype = (unsigned char)types[k];
          store.push_back(q);
        }
      }
      p += V;
    }
    const int V = f->numVertex;
    std::map<int, std::vector<HemoCellParticle*>> ppc; std::map<int, bool> lpc;
    for (size_t k = 0; k < first.size(); k++) {
      auto& vec = ppc[(int)store[k*V].sv.cellId];
      vec.resize(V);
      for (int v = 0; v < V; v++) {
        HemoCellParticle* q = &store[k*V + v];
        q->force_volume = q->force_bending = q->force_link = q->force_area = q->force_visc = q->force_inner_link = &q->sv.force;
        vec[v] = q;
      }
      lpc[(int)store[k*V].sv.cellId] = true;
    }
    f->mechanics->ParticleMechanics(ppc, lpc, f->ctype);
    for (size_t k = 0; k < first.size(); k++) for (int v = 0; v < V; v++) for (int d = 0; d < 3; d++) frc[3*(first[k]+v)+d] = store[k*V + v].sv.force[d];
  }
  ck(c, hcg_cells_upload(c, HCG_P_FORCE, frc.data()), "upload");
}
void HemoCellFields::applyConstitutiveModel(bool forced) {
  ck(ctx(), hcg_op_mechanics(ctx(), forced ? 1 : 0, separateForces ? 1 : 0), "applyConstitutiveModel");
  host_constitutive_model(*this, forced, hemocell.iter);
}
void HemoCellFields::syncEnvelopes() { ck(ctx(), hcg_op_sync(ctx()), "syncEnvelopes"); }
void HemoCellFields::getParticles(vector<HemoCellParticle>& particles) {
  particles.clear();
  hcg_ctx* c = ctx();
  int64_t nc = 0, np = 0;
  ck(c, hcg_cells_capacity(c, &nc, &np), "hcg_cells_capacity");
  if (np == 0) return;
  std::vector<double> pos(3*np), vel(3*np), frc(3*np), frep(3*np);
  ck(c, hcg_cells_download(c, HCG_P_POS, pos.data()), "download"); ck(c, hcg_cells_download(c, HCG_P_VEL, vel.data()), "download");
  ck(c, hcg_cells_download(c, HCG_P_FORCE, frc.data()), "download"); ck(c, hcg_cells_download(c, HCG_P_FREP, frep.data()), "download");
  std::vector<int64_t> ids(nc); std::vector<int32_t> types(nc); std::vector<uint8_t> alive(nc);
  ck(c, hcg_cells_info(c, ids.data(), types.data(), alive.data()), "hcg_cells_info");
  int64_t p = 0;
  for (int64_t k = 0; k < nc; k++) {
    const int V = cellFields[types[k]]->numVertex;
    if (alive[k] && ids[k] >= 0) for (int v = 0; v < V; v++) {
      HemoCellParticle q;
      for (int d = 0; d < 3; d++) { q.sv.position[d] = pos[3*(p+v)+d]; q.sv.v[d] = vel[3*(p+v)+d]; q.sv.force[d] = frc[3*(p+v)+d]; q.sv.force_repulsion[d] = frep[3*(p+v)+d]; }
      q.sv.cellId = ids[k]; q.sv.vertexId = (uint16_t)v; q.sv.restime = 0; q.sv.celltype = (unsigned char)types[k];
      particles.push_back(q);
    }
    p += V;
  }
}

// HemoCell ---------------------------------------------------------------------------------------------
HemoCell::HemoCell(char* configFileName, int argc, char* argv[]) {
  plb::plbInit(&argc, &argv);
  if (global.hemoCellInitialized) { pcout << "(HemoCell) (Error) Hemocell object already created, refusing to construct another one" << endl; exit(1); }
  global.hemoCellInitialized = true;
  pcout << "(HemoCell) (Config) reading " << configFileName << endl;
  cfg = new Config(configFileName);
  if (cfg->checkpointed) pcout << "(HemoCell) (Config) Checkpointed config, deferring the loading of the directories (out,log,checkpoint) until loadCheckpoint is called" << endl;
  else loadDirectories(cfg);
  try { global.cellsDeletedInfo = (*cfg)["verbose"]["cellsDeletedInfo"].read<int>() != 0; } catch (std::invalid_argument&) {}
  hlog << "(HemoCell) B200-native build: " << hcg_version() << endl;
  global.statistics.start();
}
HemoCell::~HemoCell() {
  delete preInlet; preInlet = nullptr;
  delete cellfields; delete cfg; delete lattice;
  global.statistics.setDeviceTimers(nullptr);
  global.hemoCellInitialized = false;
}
hcg_ctx* HemoCell::ctx() {
  if (!lattice) fatal("(HemoCell) please create a lattice first");
  lattice->gpu()->materialize();
  return lattice->gpu()->ctx;
}
void HemoCell::latticeEquilibrium(T rho, hemo::Array<T, 3> vel) {
  hlog << "(HemoCell) (Fluid) Setting Fluid Equilibrium" << endl;
  plb::initializeAtEquilibrium(*lattice, lattice->getBoundingBox(), rho, plb::Array<T, 3>(vel[0], vel[1], vel[2]));
}
void HemoCell::initializeCellfield() {
  if (!lattice) fatal("(HemoCell) (CellField) please create a lattice before initializing the cellfield");
  cellfields = new HemoCellFields(*lattice, (*cfg)["domain"]["particleEnvelope"].read<int>(), *this);
}
void HemoCell::registerCellType(HemoCellField* f) {
  if (f->mechanics->deviceModel() < 0) {
    // a user subclass of CellMechanics: its ParticleMechanics(map<...>) is honoured on host copies of the particles
    hlog << "(HemoCell) (AddCellType) " << f->name << ": mechanics class without a device kernel, ParticleMechanics runs on the host (slow path)" << endl;
    f->impl->host_model = true;
    f->impl->tables.c.model = HCG_MODEL_HOST;
  } else f->impl->tables.c.model = f->mechanics->deviceModel();
}
// cell types go to the device when it is first needed (the periodicity may be toggled until then,
// which re-creates the context: hemocell.setSystemPeriodicity comes after addCellType in the case files)
static hcg_celltype device_celltype(HemoCellField* f) {
  hcg_celltype t = f->impl->tables.c;
  if (auto* m = dynamic_cast<RbcHighOrderModel*>(f->mechanics)) { t.k_volume = m->k_volume; t.k_area = m->k_area; t.k_link = m->k_link; t.k_bend = m->k_bend; t.eta_m = m->eta_m; }
  else if (auto* m2 = dynamic_cast<PltSimpleModel*>(f->mechanics)) { t.k_volume = m2->k_volume; t.k_area = m2->k_area; t.k_link = m2->k_link; t.k_bend = m2->k_bend; t.eta_m = m2->eta_m; }
  return t;
}
static void upload_celltypes(HemoCell& h) {
  hcg_ctx* c = h.ctx();
  const int gen = h.lattice->gpu()->generation;
  for (auto* f : h.cellfields->cellFields) {
    if (f->impl->device_generation == gen) continue;
    hcg_celltype t = f->impl->tables.c;
    // the k_* members of the model object are authoritative (a case may derive its own values)
    if (auto* m = dynamic_cast<RbcHighOrderModel*>(f->mechanics)) { t.k_volume = m->k_volume; t.k_area = m->k_area; t.k_link = m->k_link; t.k_bend = m->k_bend; t.eta_m = m->eta_m; }
    else if (auto* m2 = dynamic_cast<PltSimpleModel*>(f->mechanics)) { t.k_volume = m2->k_volume; t.k_area = m2->k_area; t.k_link = m2->k_link; t.k_bend = m2->k_bend; t.eta_m = m2->eta_m; }
    int32_t id = -1;
    ck(c, hcg_celltype_add(c, &t, &id), "hcg_celltype_add");
    f->impl->device_ctype = id; f->impl->device_generation = gen;
  }
  ck(c, hcg_set_force_limit(c, param::f_limit), "hcg_set_force_limit");
}
void HemoCell::setOutputs(std::string name, vector<int> outputs) {
  hlog << "(HemoCell) (CellField) Setting output variables for " << name << " cells" << endl;
  (*cellfields)[name]->setOutputVariables(outputs);
}
void HemoCell::setFluidOutputs(vector<int> outputs) {
  hlog << "(HemoCell) (Fluid) Setting output variables for fluid field" << endl;
  cellfields->desiredFluidOutputVariables = outputs;
}
void HemoCell::setRepulsion(T repulsionConstant, T repulsionCutoff) {
  hlog << "(HemoCell) (Repulsion) Setting repulsion constant to " << repulsionConstant << ". repulsionCutoff to" << repulsionCutoff << " µm" << endl;
  hlogfile << "(HemoCell) (Repulsion) Enabling repulsion." << endl;
  cellfields->repulsionConstant = repulsionConstant;
  cellfields->repulsionCutoff = repulsionCutoff*(1e-6/param::dx);
  repulsionEnabled = true;
}
void HemoCell::enableBoundaryParticles(T k, T cutoff, unsigned int timestep) {
  hlog << "(HemoCell) (Repulsion) Setting boundary repulsion constant to " << k << ". boundary repulsionCutoff to" << cutoff << " µm" << endl;
  hlogfile << "(HemoCell) (Repulsion) Enabling boundary repulsion" << endl;
  cellfields->boundaryRepulsionConstant = k;
  cellfields->boundaryRepulsionCutoff = cutoff*(1e-6/param::dx);
  cellfields->boundaryRepulsionTimescale = timestep;
  boundaryRepulsionEnabled = true;
}
void HemoCell::setMaterialTimeScaleSeparation(std::string name, unsigned int separation) {
  hlog << "(HemoCell) (Timescale Seperation) Setting seperation of " << name << " to " << separation << " timesteps" << endl;
  (*cellfields)[name]->timescale = separation;
}
void HemoCell::setParticleVelocityUpdateTimeScaleSeparation(unsigned int separation) {
  hlog << "(HemoCell) (Timescale separation) Setting update separation of all particles to " << separation << " timesteps" << endl;
  hlogfile << "(HemoCell) WARNING this introduces curvature artifacts. Make sure, it is smaller than material timescale separation!" << endl;
  cellfields->particleVelocityUpdateTimescale = separation;
}
void HemoCell::setRepulsionTimeScaleSeperation(unsigned int separation) {
  hlog << "(HemoCell) (Repulsion Timescale Seperation) Setting seperation to " << separation << " timesteps" << endl;
  cellfields->repulsionTimescale = separation;
}
void HemoCell::setInitialMinimumDistanceFromSolid(std::string name, T distance) {
  hlog << "(HemoCell) (Set Distance) Setting minimum distance from solid to " << distance << " micrometer for " << name << endl;
  if (loadParticlesIsCalled) pcout << "(HemoCell) (Set Distance) WARNING: this function is called after the particles are loaded, so it has no effect!" << endl;
  (*cellfields)[name]->minimumDistanceFromSolid = distance;
}
void HemoCell::setSystemPeriodicity(unsigned int axis, bool bePeriodic) {
  if (lattice == nullptr) { pcerr << "(HemoCell) (Periodicity) please create a lattice before trying to set the periodicity" << endl; exit(1); }
  if (cellfields == nullptr) { pcerr << "(HemoCell) (Periodicity) please create a particlefield (hemocell.initializeCellfields()) before trying to set the periodicity" << endl; exit(1); }
  lattice->periodicity().toggle(axis, bePeriodic);
}
void HemoCell::setSystemPeriodicityLimit(unsigned int axis, int limit) {
  hlog << "(HemoCell) (Periodicity) Setting periodicity limit of axis " << axis << " to " << limit << endl;
  cellfields->periodicity_limit[axis] = limit;
}
void HemoCell::initializeLattice(const plb::MultiBlockManagement3D& management) {
  delete lattice;
  hlog << "(HemoCell) Using default domain management." << endl;
  lattice = new plb::MultiBlockLattice3D<T, DESCRIPTOR>(management, nullptr, nullptr, nullptr,
                                                        new plb::GuoExternalForceBGKdynamics<T, DESCRIPTOR>(1.0/param::tau));
  if (preInlet) preInlet->createLattice(1.0/param::tau);      // core/hemoCell.cpp:476-570: the pre-inlet's own lattice
}

// settings that live in host-side knobs until the device needs them
void HemoCell::pushSettings() {
  hcg_ctx* c = ctx();
  upload_celltypes(*this);
  ck(c, hcg_set_timescales(c, (int)cellfields->particleVelocityUpdateTimescale, (int)cellfields->repulsionTimescale,
                           (int)cellfields->boundaryRepulsionTimescale), "hcg_set_timescales");
  for (auto* f : cellfields->cellFields) ck(c, hcg_set_material_timescale(c, f->impl->device_ctype, (int)f->timescale), "hcg_set_material_timescale");
  ck(c, hcg_set_repulsion(c, repulsionEnabled, cellfields->repulsionConstant, repulsionEnabled ? cellfields->repulsionCutoff : 1.0), "hcg_set_repulsion");
  ck(c, hcg_set_wall_repulsion(c, boundaryRepulsionEnabled, cellfields->boundaryRepulsionConstant,
                               boundaryRepulsionEnabled ? cellfields->boundaryRepulsionCutoff : 1.0), "hcg_set_wall_repulsion");
  ck(c, hcg_set_iteration(c, iter), "hcg_set_iteration");
  if (preInlet && preInlet->pre) {                            // the pre-inlet ranks of the reference run the same case file: same knobs
    hcg_ctx* q = preInlet->preCtx();
    ck(q, hcg_set_timescales(q, (int)cellfields->particleVelocityUpdateTimescale, (int)cellfields->repulsionTimescale,
                             (int)cellfields->boundaryRepulsionTimescale), "hcg_set_timescales");
    for (auto* f : cellfields->cellFields) ck(q, hcg_set_material_timescale(q, f->impl->device_ctype, (int)f->timescale), "hcg_set_material_timescale");
    ck(q, hcg_set_repulsion(q, repulsionEnabled, cellfields->repulsionConstant, repulsionEnabled ? cellfields->repulsionCutoff : 1.0), "hcg_set_repulsion");
    ck(q, hcg_set_wall_repulsion(q, boundaryRepulsionEnabled, cellfields->boundaryRepulsionConstant,
                                 boundaryRepulsionEnabled ? cellfields->boundaryRepulsionCutoff : 1.0), "hcg_set_wall_repulsion");
    ck(q, hcg_set_iteration(q, iter), "hcg_set_iteration");
  }
}

void HemoCell::loadParticles() {
  hlog << "(HemoCell) (CellField) Loading particle positions " << endl;
  loadParticlesIsCalled = true;
  hcg_ctx* c = ctx();
  upload_celltypes(*this);
  GpuLattice* g = lattice->gpu();
  int64_t cellid = 0;                                  // one counter over the .pos files of all types (readPositionsBloodCells.cpp:204-228)
  for (auto* f : cellfields->cellFields) {
    std::vector<std::array<T, 6>> rows;
    try { rows = host::readPositionsFile(f->name + ".pos"); }
    catch (std::invalid_argument&) { cout << "*** WARNING! particle positions input file " << f->name << ".pos does not exist!" << endl; }
    hlog << "(readPositionsBloodCells) Particle count in file (" << f->name << "): " << rows.size() << "." << endl;
    std::vector<T> pos;
    std::vector<int64_t> ids = host::placeCells(f->impl->tables.mesh, rows, param::dx, g->nx, g->ny, g->nz, g->flags.data(),
                                                f->minimumDistanceFromSolid, cellid, pos);
    if (preInlet && preInlet->pre) {
      // the same .pos file seeds the pre-inlet (rows in its local coordinates); the main domain keeps spare slots for the cells it will feed in
      ck(c, hcg_cells_reserve(c, f->impl->device_ctype, (int64_t)rows.size() + 64), "hcg_cells_reserve");
      GpuLattice* q = preInlet->pre;
      hcg_ctx* qc = preInlet->preCtx();
      hcg_celltype t = device_celltype(f);
      int32_t qid = -1;
      ck(qc, hcg_celltype_add(qc, &t, &qid), "hcg_celltype_add");
      if (qid != f->impl->device_ctype) fatal("(PreInlet) cell types of the pre-inlet and the main domain are out of step");
      ck(qc, hcg_set_force_limit(qc, param::f_limit), "hcg_set_force_limit");
      std::vector<std::array<T, 6>> prows = rows;
      const T um = param::dx*1e6;
      for (auto& r : prows) { r[0] -= preInlet->location.x0*um; r[1] -= preInlet->location.y0*um; r[2] -= preInlet->location.z0*um; }
      std::vector<T> ppos;
      std::vector<int64_t> pids = host::placeCells(f->impl->tables.mesh, prows, param::dx, q->nx, q->ny, q->nz, q->flags.data(),
                                                   f->minimumDistanceFromSolid, cellid, ppos);
      q->has_cells = q->has_cells || !pids.empty();
      ck(qc, hcg_cells_add(qc, qid, (int64_t)pids.size(), pids.data(), ppos.data()), "hcg_cells_add");
      hlog << "(readPositionsBloodCells) " << pids.size() << " " << f->name << " cells placed inside the pre-inlet." << endl;
    }
    cellid += (int64_t)rows.size();
    g->has_cells = g->has_cells || !ids.empty();
    ck(c, hcg_cells_add(c, f->impl->device_ctype, (int64_t)ids.size(), ids.data(), pos.data()), "hcg_cells_add");
    f->impl->n_cells_loaded = (int64_t)ids.size();
    hlog << "(readPositionsBloodCells) " << ids.size() << " " << f->name << " cells placed inside the domain." << endl;
  }
  cellfields->number_of_cells = cellid;
}

void HemoCell::sanityCheck() {
  // core/hemoCell.cpp:600-627: material / repulsion cadences must be multiples of the velocity cadence
  hlog << "(HemoCell) (SanityCheck) Performing Sanity check on simulation parameters and setup" << endl;
  for (auto* f : cellfields->cellFields)
    if (f->timescale % cellfields->particleVelocityUpdateTimescale != 0)
      fatal("(HemoCell) (SanityCheck) Error, Velocity timescale separation cannot divide this material timescale separation, exiting ...");
  if (repulsionEnabled && cellfields->repulsionTimescale % cellfields->particleVelocityUpdateTimescale != 0)
    fatal("(HemoCell) (SanityCheck) Error, Velocity timescale separation cannot divide this repulsion timescale separation, exiting ...");
  if (param::tau < 0.5 || param::tau > 2.0) hlog << "(HemoCell) (SanityCheck) WARNING: tau = " << param::tau << " is outside of the recommended range (0.5, 2.0]" << endl;
  pushSettings();
  sanityCheckDone = true;
}

void HemoCell::iterate() {
  if (!sanityCheckDone) sanityCheck();
  hcg_ctx* c = ctx();
  ck(c, hcg_iterate(c, 1), "hcg_iterate");
  host_constitutive_model(*cellfields, false, iter);          // user models without a device kernel (last operator of iterate(), core/hemoCell.cpp:362)
  if (preInlet) preInlet->iterate();                          // asynchronous, on the pre-inlet context's own stream (or GPU)
  iter++;
}

// ranks meet on the device communicator (the facade has no host-side MPI)
static void rank_barrier(hcg_ctx* c) {
  if (plb::global::mpi().getSize() <= 1) return;
  double one = 1.0;
  ck(c, hcg_allreduce(c, &one, 1, 0), "hcg_allreduce");
}

void HemoCell::saveCheckPoint() {
  hlog << "(HemoCell) (Saving Functions) Saving Checkpoint at timestep " << iter << endl;

  hcg_ctx* c = ctx();
  GpuLattice* g = lattice->gpu();
  const std::string dir = global.checkpointDirectory;
  mkpath(dir);
  const std::string base = dir + "rank" + std::to_string(plb::global::mpi().getRank());
  // keep the previous checkpoint as .old (core/hemoCellFields.cpp:240-262)
  if (file_exists(base + ".bin")) rename((base + ".bin").c_str(), (base + ".bin.old").c_str());
  std::ofstream f(base + ".bin", std::ios::binary);
  const int64_t Nl = (int64_t)g->nxl()*g->ny*g->nz;
  int64_t nc = 0, np = 0;
  ck(c, hcg_cells_capacity(c, &nc, &np), "hcg_cells_capacity");
  const int64_t hdr[8] = {0x48434732, (int64_t)iter, Nl, nc, np, (int64_t)cellfields->size(), (int64_t)plb::global::mpi().getSize(), (int64_t)g->x0()};
  f.write((const char*)hdr, sizeof(hdr));
  std::vector<double> buf((size_t)std::max<int64_t>(19*Nl, 3*np));
  ck(c, hcg_lattice_download(c, HCG_LAT_POP, buf.data()), "download"); f.write((const char*)buf.data(), 8*19*Nl);
  ck(c, hcg_lattice_download(c, HCG_LAT_FORCE, buf.data()), "download"); f.write((const char*)buf.data(), 8*3*Nl);
  for (int fld : {HCG_P_POS, HCG_P_VEL, HCG_P_FORCE, HCG_P_FREP}) { ck(c, hcg_cells_download(c, fld, buf.data()), "download"); f.write((const char*)buf.data(), 8*3*np); }
  std::vector<int64_t> ids(nc); std::vector<int32_t> types(nc); std::vector<uint8_t> alive(nc);
  ck(c, hcg_cells_info(c, ids.data(), types.data(), alive.data()), "hcg_cells_info");
  f.write((const char*)ids.data(), 8*nc); f.write((const char*)types.data(), 4*nc); f.write((const char*)alive.data(), nc);
  f.close();
  if (!f) fatal("(HemoCell) (Saving Functions) writing " + base + ".bin failed");
  if (preInlet && preInlet->pre) preInlet->saveCheckPoint(dir);
  rank_barrier(c);                                            // checkpoint.xml appears only once every rank's block is complete
  if (plb::global::mpi().isMainProcessor()) {
    // checkpoint.xml: <Checkpoint><General><Iteration>..</Iteration></General> + the original <hemocell> tree (config/config.cpp:53-78)
    xml::Node root; xml::Node* cp = root.addChild("Checkpoint");
    xml::Node* gen = cp->addChild("General");
    gen->addChild("Iteration", std::to_string(iter));
    gen->addChild("OutDirectory", plb::global::directories().getOutputDir());      // core/hemoCellFields.cpp:248: where a restart continues
    std::function<void(const xml::Node*, xml::Node*)> copy = [&](const xml::Node* s, xml::Node* d) {
      for (auto& ch : s->children) { xml::Node* n = d->addChild(ch->name, ch->text); n->attributes = ch->attributes; copy(ch.get(), n); }
    };
    const xml::Node* src = cfg->checkpointed ? cfg->root()->firstChild("Checkpoint") : cfg->root();
    if (const xml::Node* h = src->firstChild("hemocell")) copy(h, cp->addChild("hemocell"));
    if (file_exists(dir + "checkpoint.xml")) rename((dir + "checkpoint.xml").c_str(), (dir + "checkpoint.xml.old").c_str());
    std::ofstream x(dir + "checkpoint.xml"); x << xml::serialize(root);
  }
}

void HemoCell::loadCheckPoint() {
  hlog << "(HemoCell) (Saving Functions) Loading Checkpoint" << endl;
  if (!cfg->checkpointed) fatal("(HemoCell) loadCheckPoint() needs a checkpoint.xml as configuration file");
  const xml::Node* cp = cfg->root()->firstChild("Checkpoint");
  iter = (unsigned int)XMLElement(cp)["General"]["Iteration"].read<unsigned long>();
  // core/hemoCellFields.cpp:246-252: the run continues in the output directory recorded in the checkpoint
  try { plb::global::directories().setOutputDir(XMLElement(cp)["General"]["OutDirectory"].read<std::string>()); } catch (std::invalid_argument&) {}
  loadDirectories(cfg, false);
  hcg_ctx* c = ctx();
  upload_celltypes(*this);
  GpuLattice* g = lattice->gpu();
  const std::string base = global.checkpointDirectory + "/rank" + std::to_string(plb::global::mpi().getRank()) + ".bin";
  std::ifstream f(base, std::ios::binary);
  if (!f) fatal("(HemoCell) cannot open checkpoint data " + base);
  int64_t hdr[8]; f.read((char*)hdr, sizeof(hdr));
  const int64_t Nl = (int64_t)g->nxl()*g->ny*g->nz;
  if (!f || hdr[0] != 0x48434732 || hdr[2] != Nl || hdr[5] != (int64_t)cellfields->size() || hdr[6] != (int64_t)plb::global::mpi().getSize() || hdr[7] != (int64_t)g->x0())
    fatal("(HemoCell) checkpoint does not match this domain / cell types / number of ranks (a restart needs the decomposition it was written with)");
  const int64_t nc = hdr[3], np = hdr[4];
  std::vector<double> pop(19*Nl), frc(3*Nl);
  f.read((char*)pop.data(), 8*19*Nl); f.read((char*)frc.data(), 8*3*Nl);
  std::vector<double> P[4];
  for (auto& v : P) { v.resize(3*np); f.read((char*)v.data(), 8*3*np); }
  std::vector<int64_t> ids(nc); std::vector<int32_t> types(nc); std::vector<uint8_t> alive(nc);
  f.read((char*)ids.data(), 8*nc); f.read((char*)types.data(), 4*nc); f.read((char*)alive.data(), nc);
  // re-create the cells slot by slot (live ones only; slot order = type order), then overwrite the state
  std::vector<int64_t> base_of(nc); { int64_t p = 0; for (int64_t k = 0; k < nc; k++) { base_of[k] = p; p += (*cellfields)[(unsigned)types[k]]->numVertex; } }
  if (!f) fatal("(HemoCell) checkpoint data " + base + " is truncated");
  for (auto* fl : cellfields->cellFields) {
    std::vector<int64_t> tid; std::vector<double> tpos;
    for (int64_t k = 0; k < nc; k++) if (types[k] == fl->impl->device_ctype && alive[k] && ids[k] >= 0) {
      tid.push_back(ids[k]);
      const int V = fl->numVertex;
      tpos.insert(tpos.end(), P[0].begin() + 3*base_of[k], P[0].begin() + 3*(base_of[k] + V));
    }
    if (preInlet && preInlet->pre) ck(c, hcg_cells_reserve(c, fl->impl->device_ctype, 2*(int64_t)tid.size() + 256), "hcg_cells_reserve");
    ck(c, hcg_cells_add(c, fl->impl->device_ctype, (int64_t)tid.size(), tid.data(), tpos.data()), "hcg_cells_add");
    g->has_cells = g->has_cells || !tid.empty();
  }
  // velocity, membrane force and repulsion force (the reference serialises sv.v, sv.force, force_repulsion): scattered by cell
  // id into the slot layout of the new context (spare slots of the multi-GPU slack / pre-inlet reserve, any number of ranks)
  {
    int64_t cap_c = 0, cap_p = 0;
    ck(c, hcg_cells_capacity(c, &cap_c, &cap_p), "hcg_cells_capacity");
    if (cap_c > 0) {
      std::vector<int64_t> sid(cap_c); std::vector<int32_t> stype(cap_c); std::vector<uint8_t> salive(cap_c);
      ck(c, hcg_cells_info(c, sid.data(), stype.data(), salive.data()), "hcg_cells_info");
      std::map<int64_t, int64_t> saved;
      for (int64_t k = 0; k < nc; k++) if (alive[k] && ids[k] >= 0) saved[ids[k]] = k;
      std::vector<double> st[3];
      for (auto& v : st) v.assign((size_t)3*cap_p, 0.0);
      int64_t p = 0;
      for (int64_t slot = 0; slot < cap_c; slot++) {
        const int V = (*cellfields)[(unsigned)stype[slot]]->numVertex;
        auto it = (sid[slot] >= 0 && salive[slot]) ? saved.find(sid[slot]) : saved.end();
        if (it != saved.end())
          for (int a = 0; a < 3; a++) std::copy(P[1 + a].begin() + 3*base_of[it->second], P[1 + a].begin() + 3*(base_of[it->second] + V), st[a].begin() + 3*p);
        p += V;
      }
      ck(c, hcg_cells_upload(c, HCG_P_VEL, st[0].data()), "upload"); ck(c, hcg_cells_upload(c, HCG_P_FORCE, st[1].data()), "upload");
      ck(c, hcg_cells_upload(c, HCG_P_FREP, st[2].data()), "upload");
    }
  }
  ck(c, hcg_lattice_upload(c, HCG_LAT_POP, pop.data()), "upload"); ck(c, hcg_lattice_upload(c, HCG_LAT_FORCE, frc.data()), "upload");
  g->eq_pending = false; g->body_pending = false;
  loadParticlesIsCalled = true;
  if (preInlet && preInlet->pre) preInlet->loadCheckPoint(global.checkpointDirectory + "/");
}

// snapshot of the cell bookkeeping (ids, types, alive / owned flags), shared by output and observables
namespace {
struct CellSnapshot {
  int64_t nc = 0, np = 0;
  std::vector<int64_t> ids, base; std::vector<int32_t> types; std::vector<uint8_t> alive;
  std::vector<uint8_t> owned;       // multi-rank: cells this rank owns (each cell is owned by exactly one rank)
};
CellSnapshot snapshot(HemoCell* h) {
  CellSnapshot s; hcg_ctx* c = h->ctx();
  ck(c, hcg_cells_capacity(c, &s.nc, &s.np), "hcg_cells_capacity");
  s.ids.resize(s.nc); s.types.resize(s.nc); s.alive.resize(s.nc); s.base.resize(s.nc);
  if (s.nc) ck(c, hcg_cells_info(c, s.ids.data(), s.types.data(), s.alive.data()), "hcg_cells_info");
  s.owned.assign(s.nc, 1);
  if (s.nc) ck(c, hcg_cells_owned(c, s.owned.data()), "hcg_cells_owned");
  int64_t p = 0;
  for (int64_t k = 0; k < s.nc; k++) { s.base[k] = p; p += (*h->cellfields)[(unsigned)s.types[k]]->numVertex; }
  return s;
}
CellInformation& entry(const CellSnapshot& s, int64_t k) {
  CellInformation& ci = CellInformationFunctionals::info_per_cell[(int)s.ids[k]];
  ci.cellType = (pluint)s.types[k]; ci.base_cell_id = (int)s.ids[k]; ci.blockId = (pluint)plb::global::mpi().getRank();
  return ci;
}
}  // namespace

// ---- HDF5 output ---------------------------------------------------------------------------------------
// File names, dataset names, shapes, element types, root attributes, SI scaling and chunking follow
// io/ParticleHdf5IO.cpp:60-194 and io/FluidHdf5IO.hh:74-211; the container is written by hemo_h5 (no libhdf5 here).
namespace {
int h5_deflate_level() { const char* e = getenv("HEMOCELL_H5_DEFLATE"); return e ? atoi(e) : 7; }   // < 0: contiguous, uncompressed
std::string h5_name(const HemoCell& h, const std::string& identifier) {
  return plb::global::directories().getOutputDir() + "/hdf5/" + zeroPadNumber(h.iter) + '/' + identifier + "." + zeroPadNumber(h.iter) +
         ".p." + std::to_string(plb::global::mpi().getRank()) + ".h5";
}
void h5_common_attrs(h5::Writer& w, const HemoCell& h) {
  const double dx = param::dx, dt = param::dt; const int64_t it = h.iter; const int32_t id = plb::global::mpi().getRank();
  w.attribute("dx", h5::F64, &dx, 1); w.attribute("dt", h5::F64, &dt, 1);
  w.attribute("iteration", h5::I64, &it, 1); w.attribute("processorId", h5::I32, &id, 1);
}
void write_particle_h5(HemoCell& h, HemoCellField& field) {
  hcg_ctx* c = h.ctx();
  h5::Writer w(h5_name(h, field.name), h5_deflate_level());
  if (!w.ok()) fatal("(HemoCell) (Output) cannot create " + h5_name(h, field.name));
  h5_common_attrs(w, h);
  const int64_t nproc = plb::global::mpi().getSize();
  w.attribute("numberOfProcessors", h5::I64, &nproc, 1);
  int64_t nc = 0, np = 0;
  ck(c, hcg_cells_capacity(c, &nc, &np), "hcg_cells_capacity");
  std::vector<int64_t> ids(nc), base(nc); std::vector<int32_t> types(nc); std::vector<uint8_t> alive(nc);
  if (nc) ck(c, hcg_cells_info(c, ids.data(), types.data(), alive.data()), "hcg_cells_info");
  { int64_t p = 0; for (int64_t k = 0; k < nc; k++) { base[k] = p; p += (*h.cellfields)[(unsigned)types[k]]->numVertex; } }
  // cells of this type in ascending cell id (the reference walks a std::map keyed by cell id); multi-GPU: a cell
  // shared by two ranks is written by the one that owns it (hcg_cells_owned), so every cell appears in exactly one file
  const int t = field.impl->device_ctype, V = field.numVertex;
  std::vector<double> pos((size_t)3*np);
  if (np) ck(c, hcg_cells_download(c, HCG_P_POS, pos.data()), "download");
  std::vector<uint8_t> owned(nc, 1);
  if (nc) ck(c, hcg_cells_owned(c, owned.data()), "hcg_cells_owned");
  std::vector<std::pair<int64_t, int64_t>> order;       // (cell id, slot)
  for (int64_t k = 0; k < nc; k++) {
    if (!alive[k] || !owned[k] || ids[k] < 0 || types[k] != t) continue;
    order.push_back({ids[k], k});
  }
  std::sort(order.begin(), order.end());
  const uint64_t N = (uint64_t)order.size()*V;
  const std::vector<uint64_t> chunk3 = {std::max<uint64_t>(1, std::min<uint64_t>(1000, N)), 3}, chunk1 = {chunk3[0], 1};
  std::vector<double> buf((size_t)3*np);
  std::vector<float> out((size_t)3*N);
  auto vector_field = [&](int fld, const char* name, double scale) {
    if (np) ck(c, hcg_cells_download(c, fld, buf.data()), "download");
    size_t n = 0;
    for (auto& o : order) for (int v = 0; v < V; v++) for (int d = 0; d < 3; d++) out[n++] = (float)(buf[3*(base[o.second] + v) + d]*scale);
    w.dataset(name, h5::F32, {N, 3}, out.data(), chunk3);
  };
  const bool si = h.outputInSiUnits;
  for (int var : field.desiredOutputVariables) {
    switch (var) {
      case OUTPUT_POSITION: {
        vector_field(HCG_P_POS, "Position", si ? param::dx : 1.0);
        const int64_t nP = (int64_t)N; w.attribute("numberOfParticles", h5::I64, &nP, 1);
        break; }
      case OUTPUT_VELOCITY: vector_field(HCG_P_VEL, "Velocity", si ? param::dx/param::dt : 1.0); break;
      case OUTPUT_FORCE: {
        // force_total = sum of the constitutive parts + repulsion (core/hemoCellParticle.h), valid after the forced model update above
        if (np) ck(c, hcg_cells_download(c, HCG_P_FORCE, buf.data()), "download");
        std::vector<double> rep((size_t)3*np); if (np) ck(c, hcg_cells_download(c, HCG_P_FREP, rep.data()), "download");
        size_t n = 0; const double sc = si ? param::df : 1.0;
        for (auto& o : order) for (int v = 0; v < V; v++) for (int d = 0; d < 3; d++) { const size_t q = 3*(base[o.second] + v) + d; out[n++] = (float)((buf[q] + rep[q])*sc); }
        w.dataset("Total force", h5::F32, {N, 3}, out.data(), chunk3);
        break; }
      case OUTPUT_FORCE_VOLUME: vector_field(HCG_P_F_VOLUME, "Volume force", si ? param::df : 1.0); break;
      case OUTPUT_FORCE_AREA: vector_field(HCG_P_F_AREA, "Area force", si ? param::df : 1.0); break;
      case OUTPUT_FORCE_LINK: vector_field(HCG_P_F_LINK, "Link force", si ? param::df : 1.0); break;
      case OUTPUT_FORCE_BENDING: vector_field(HCG_P_F_BEND, "Bending force", si ? param::df : 1.0); break;
      case OUTPUT_FORCE_VISC: vector_field(HCG_P_F_VISC, "Viscous force", si ? param::df : 1.0); break;
      case OUTPUT_FORCE_INNER_LINK: vector_field(HCG_P_F_INNER, "Inner link force", si ? param::df : 1.0); break;
      case OUTPUT_FORCE_REPULSION: vector_field(HCG_P_FREP, "Repulsion force", si ? param::df : 1.0); break;
      case OUTPUT_VERTEX_ID: case OUTPUT_CELL_ID: case OUTPUT_RES_TIME: {
        std::vector<float> one((size_t)N); size_t n = 0;
        // residence time: writeOutput adds the iterations since the last output to every particle (core/hemoCell.cpp:226,
        // HemoCellFields::updateResidenceTime); without a pre-inlet every cell exists from iteration 0, so it is the iteration count
        const float res = (float)((double)h.iter*(si ? param::dt : 1.0));
        for (auto& o : order) for (int v = 0; v < V; v++) one[n++] = var == OUTPUT_VERTEX_ID ? (float)v : (var == OUTPUT_CELL_ID ? (float)o.first : res);
        w.dataset(var == OUTPUT_VERTEX_ID ? "Vertex Id" : (var == OUTPUT_CELL_ID ? "Cell Id" : "Res Time"), h5::F32, {N, 1}, one.data(), chunk1);
        break; }
      default: break;      // the reference skips variables without a particle output function (io/hemoCellParticleFieldOutputFunctions.cpp:46-50)
    }
  }
  if (field.outputTriangles) {
    const auto& tri = field.impl->tables.cc.triangle_list;
    const uint64_t nT = (uint64_t)order.size()*tri.size();
    std::vector<int32_t> tt((size_t)3*nT); size_t n = 0; int32_t counter = 0;
    for (size_t k = 0; k < order.size(); k++) { for (auto& q : tri) for (int d = 0; d < 3; d++) tt[n++] = q[d] + counter; counter += V; }
    w.dataset("Triangles", h5::I32, {nT, 3}, tt.data(), {std::max<uint64_t>(1, std::min<uint64_t>(1000, nT)), 3});
    const int64_t v = (int64_t)nT; w.attribute("numberOfTriangles", h5::I64, &v, 1);
  }
  if (std::find(field.desiredOutputVariables.begin(), field.desiredOutputVariables.end(), OUTPUT_INNER_LINKS) != field.desiredOutputVariables.end()) {
    const auto& il = field.impl->tables.cc.inner_edge_list;
    const uint64_t nL = (uint64_t)order.size()*il.size();
    if (nL) {
      std::vector<int32_t> ll((size_t)2*nL); size_t n = 0; int32_t counter = 0;
      for (size_t k = 0; k < order.size(); k++) { for (auto& q : il) for (int d = 0; d < 2; d++) ll[n++] = q[d] + counter; counter += V; }
      w.dataset("InnerLinks", h5::I32, {nL, 2}, ll.data(), {std::max<uint64_t>(1, std::min<uint64_t>(1000, nL)), 2});
      const int64_t v = (int64_t)nL; w.attribute("numberOfInnerLinks", h5::I64, &v, 1);
    }
  }
  if (!w.close()) fatal("(HemoCell) (Output) writing " + h5_name(h, field.name) + " failed: " + w.error());
}

void write_fluid_h5(HemoCell& h) {
  const std::vector<int>& vars = h.cellfields->desiredFluidOutputVariables;
  if (vars.empty()) return;
  hcg_ctx* c = h.ctx();
  GpuLattice* g = h.lattice->gpu();
  h5::Writer w(h5_name(h, "Fluid"), h5_deflate_level());
  if (!w.ok()) fatal("(HemoCell) (Output) cannot create " + h5_name(h, "Fluid"));
  h5_common_attrs(w, h);
  // the block of this rank plus an envelope of one node on every side "for paraview" (io/FluidHdf5IO.hh:103-120);
  // datasets are [Nz][Ny][Nx][C]
  const int nxl = g->nxl(), ny = g->ny, nz = g->nz, x0 = g->x0();
  const uint64_t Nx = nxl + 2, Ny = ny + 2, Nz = nz + 2, nCells = Nx*Ny*Nz;
  const int32_t ncells = (int32_t)nCells, sub[3] = {(int32_t)Nz, (int32_t)Ny, (int32_t)Nx};
  const bool si = h.outputInSiUnits;
  float dxdydz[3] = {1.f, 1.f, 1.f}, rel[3] = {-1.5f, -1.5f, (float)(x0 - 1.5)};
  if (si) for (int k = 0; k < 3; k++) { rel[k] *= (float)param::dx; dxdydz[k] = (float)param::dx; }
  w.attribute("numberOfCells", h5::I32, &ncells, 1); w.attribute("subdomainSize", h5::I32, sub, 3);
  w.attribute("relativePosition", h5::F32, rel, 3); w.attribute("dxdydz", h5::F32, dxdydz, 3);
  // chunk = min(1000, N) per axis x all components, as io/FluidHdf5IO.hh:133-137; HDF5 caps a chunk at 4 GiB, so very large
  // blocks are split further along z
  auto fluid_chunk = [&](int C) {
    std::vector<uint64_t> ch = {std::min<uint64_t>(1000, Nz), std::min<uint64_t>(1000, Ny), std::min<uint64_t>(1000, Nx), (uint64_t)C};
    while (ch[0] > 1 && ch[0]*ch[1]*ch[2]*ch[3]*4 > (1ull << 31)) ch[0] = (ch[0] + 1)/2;
    return ch;
  };
  const int64_t Nl = (int64_t)nxl*ny*nz;
  // envelope node -> source node of this rank's slab, or -1 (outside a non-periodic face, or in a neighbouring
  // rank's slab: those envelope values are left at the Palabos background default, rho = 1, u = 0)
  auto src = [&](int64_t ix, int64_t iy, int64_t iz) -> int64_t {
    int64_t x = ix - 1, y = iy - 1, z = iz - 1;
    if (x < 0 || x >= nxl) { if (g->size() > 1 || !g->periodic[0]) return -1; x = (x + nxl) % nxl; }
    if (y < 0 || y >= ny) { if (!g->periodic[1]) return -1; y = (y + ny) % ny; }
    if (z < 0 || z >= nz) { if (!g->periodic[2]) return -1; z = (z + nz) % nz; }
    return z + (int64_t)nz*(y + (int64_t)ny*x);
  };
  std::vector<int64_t> map((size_t)nCells);
  { size_t n = 0; for (uint64_t iz = 0; iz < Nz; iz++) for (uint64_t iy = 0; iy < Ny; iy++) for (uint64_t ix = 0; ix < Nx; ix++) map[n++] = src(ix, iy, iz); }
  std::vector<double> buf;
  std::vector<float> out;
  // gather `C` components of a downloaded SoA field [C][Nl] into [Nz][Ny][Nx][C] floats
  auto emit = [&](const std::string& name, int C, double scale, double outside) {
    out.assign((size_t)nCells*C, 0.f);
    for (size_t n = 0; n < (size_t)nCells; n++) for (int k = 0; k < C; k++) out[n*C + k] = (float)((map[n] >= 0 ? buf[(size_t)k*Nl + map[n]] : outside)*scale);
    w.dataset(name, h5::F32, {Nz, Ny, Nx, (uint64_t)C}, out.data(), fluid_chunk(C));
  };
  const uint8_t* flags = g->flags.data() + (int64_t)x0*ny*nz;
  const double omega = g->omega;
  for (int var : vars) {
    switch (var) {
      case OUTPUT_VELOCITY:
        buf.resize((size_t)3*Nl); ck(c, hcg_lattice_download(c, HCG_LAT_VELOCITY, buf.data()), "download");
        emit("Velocity", 3, si ? param::dx/param::dt : 1.0, 0.0); break;
      case OUTPUT_FORCE:
        buf.resize((size_t)3*Nl); ck(c, hcg_lattice_download(c, HCG_LAT_FORCE, buf.data()), "download");
        emit("Force", 3, si ? param::df : 1.0, 0.0); break;
      case OUTPUT_DENSITY:
        buf.resize((size_t)Nl); ck(c, hcg_lattice_download(c, HCG_LAT_DENSITY, buf.data()), "download");
        emit("Density", 1, si ? param::df/(param::dx*param::dx) : 1.0, 1.0); break;
      case OUTPUT_BOUNDARY:
        buf.resize((size_t)Nl); for (int64_t n = 0; n < Nl; n++) buf[n] = flags[n] != HCG_FLUID ? 1.0 : 0.0;   // isBoundary(): bounce-back, velocity-plane and Zou-He nodes alike (io/FluidHdf5IO.hh:288-305)
        emit("Boundary", 1, 1.0, 0.0); break;
      case OUTPUT_OMEGA:
        buf.resize((size_t)Nl); for (int64_t n = 0; n < Nl; n++) buf[n] = flags[n] == HCG_BOUNCEBACK ? 0.0 : omega;
        emit("Omega", 1, si ? param::df/(param::dx*param::dx) : 1.0, omega); break;
      case OUTPUT_CELL_DENSITY: {
        // LSP count per nearest node x volumeFractionOfLspPerNode (io/FluidHdf5IO.hh:376-404)
        int64_t nc = 0, np = 0; ck(c, hcg_cells_capacity(c, &nc, &np), "hcg_cells_capacity");
        std::vector<int64_t> ids(nc); std::vector<int32_t> types(nc); std::vector<uint8_t> alive(nc);
        if (nc) ck(c, hcg_cells_info(c, ids.data(), types.data(), alive.data()), "hcg_cells_info");
        std::vector<double> pos((size_t)3*np); if (np) ck(c, hcg_cells_download(c, HCG_P_POS, pos.data()), "download");
        for (unsigned i = 0; i < h.cellfields->size(); i++) {
          HemoCellField& f = *(*h.cellfields)[i];
          out.assign((size_t)nCells, 0.f);
          int64_t p = 0;
          for (int64_t k = 0; k < nc; k++) {
            const int V = (*h.cellfields)[(unsigned)types[k]]->numVertex;
            if (alive[k] && ids[k] >= 0 && types[k] == f.impl->device_ctype) for (int v = 0; v < V; v++) {
              int64_t q[3]; const int64_t dims[3] = {g->nx, ny, nz};
              for (int d = 0; d < 3; d++) { q[d] = (int64_t)std::floor(pos[3*(p + v) + d] + 0.5); if (g->periodic[d]) q[d] = ((q[d] % dims[d]) + dims[d]) % dims[d]; }
              const int64_t lx = q[0] - x0;
              if (lx < 0 || lx >= nxl || q[1] < 0 || q[1] >= ny || q[2] < 0 || q[2] >= nz) continue;
              out[(size_t)((lx + 1) + (q[1] + 1)*(int64_t)Nx + (q[2] + 1)*(int64_t)(Nx*Ny))] += 1.f;
            }
            p += V;
          }
          if (si) for (auto& v : out) v *= (float)f.volumeFractionOfLspPerNode;
          w.dataset("CellDensity_" + f.name, h5::F32, {Nz, Ny, Nx, 1}, out.data(), fluid_chunk(1));
        }
        break; }
      case OUTPUT_SHEAR_STRESS:
        // Cell::computeShearStress of the BGK dynamics: (omega/2 - 1) PiNeq (Palabos, restated from memory)
        buf.resize((size_t)6*Nl); ck(c, hcg_lattice_download(c, HCG_LAT_PINEQ, buf.data()), "download");
        emit("ShearStress", 6, (0.5*omega - 1.0)*(si ? param::df/(param::dx*param::dx) : 1.0), 0.0); break;
      case OUTPUT_STRAIN_RATE: {
        // computeStrainRateFromStress: S = -omega invCs2 / (2 rho) PiNeq (Palabos, restated from memory)
        buf.resize((size_t)6*Nl); ck(c, hcg_lattice_download(c, HCG_LAT_PINEQ, buf.data()), "download");
        std::vector<double> rho((size_t)Nl); ck(c, hcg_lattice_download(c, HCG_LAT_DENSITY, rho.data()), "download");
        for (int k = 0; k < 6; k++) for (int64_t n = 0; n < Nl; n++) buf[(size_t)k*Nl + n] *= -omega*3.0/(2.0*rho[n]);
        emit("StrainRate", 6, si ? 1.0/param::dt : 1.0, 0.0); break; }
      case OUTPUT_SHEAR_RATE: {
        // central differences of the node velocity, [d u_a / d x_b] at index 3a + b (io/FluidHdf5IO.hh:437-501)
        std::vector<double> u((size_t)3*Nl); ck(c, hcg_lattice_download(c, HCG_LAT_VELOCITY, u.data()), "download");
        out.assign((size_t)nCells*9, 0.f);
        const double sc = si ? 1.0/param::dt : 1.0;
        size_t n = 0;
        for (uint64_t iz = 0; iz < Nz; iz++) for (uint64_t iy = 0; iy < Ny; iy++) for (uint64_t ix = 0; ix < Nx; ix++, n++) {
          const int64_t nb[3][2] = {{src(ix + 1, iy, iz), src((int64_t)ix - 1, iy, iz)}, {src(ix, iy + 1, iz), src(ix, (int64_t)iy - 1, iz)}, {src(ix, iy, iz + 1), src(ix, iy, (int64_t)iz - 1)}};
          for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) {
            const double up = nb[b][0] >= 0 ? u[(size_t)a*Nl + nb[b][0]] : 0.0, um = nb[b][1] >= 0 ? u[(size_t)a*Nl + nb[b][1]] : 0.0;
            out[n*9 + 3*a + b] = (float)((up - um)/2*sc);
          }
        }
        w.dataset("ShearRate", h5::F32, {Nz, Ny, Nx, 9}, out.data(), fluid_chunk(9));
        break; }
      default: break;
    }
  }
  if (!w.close()) fatal("(HemoCell) (Output) writing " + h5_name(h, "Fluid") + " failed: " + w.error());
}
}  // namespace

// io/writeCellInfoCSV.cpp:47-70.  With several ranks every rank contributes the cells it owns; the rows travel through
// hcg_allreduce (each rank fills its own slice of a zero-initialised table) and rank 0 writes the files.
void writeCellInfo_CSV(HemoCell& hemocell) {
  HemoCell* self = &hemocell;
  const std::string out = plb::global::directories().getOutputDir();
  mkpath(out + "/csv");
  CellInformationFunctionals::calculateCellInformation(self);
  const int R = plb::global::mpi().getSize(), r = plb::global::mpi().getRank();
  const int W = 12;      // cellType, X, Y, Z, area, volume, block, cellId, baseCellId, vx, vy, vz
  std::vector<double> rows;
  {
    CellSnapshot s = snapshot(self);
    std::map<int, bool> owned;
    for (int64_t k = 0; k < s.nc; k++) if (s.alive[k] && s.ids[k] >= 0) owned[(int)s.ids[k]] = s.owned[k] != 0;
    for (auto& pr : CellInformationFunctionals::info_per_cell) {
      if (R > 1 && !owned[pr.first]) continue;
      CellInformation ci = pr.second;
      if (self->outputInSiUnits) { ci.position *= param::dx; ci.area *= param::dx*param::dx; ci.velocity *= param::dx/param::dt; ci.volume *= param::dx*param::dx*param::dx; }
      const double row[W] = {(double)ci.cellType, ci.position[0], ci.position[1], ci.position[2], ci.area, ci.volume, (double)ci.blockId,
                             (double)pr.first, (double)ci.base_cell_id, ci.velocity[0], ci.velocity[1], ci.velocity[2]};
      rows.insert(rows.end(), row, row + W);
    }
  }
  if (R > 1) {
    std::vector<double> counts(R, 0.0); counts[r] = (double)(rows.size()/W);
    ck(self->ctx(), hcg_allreduce(self->ctx(), counts.data(), R, 0), "hcg_allreduce");
    size_t total = 0, off = 0;
    for (int k = 0; k < R; k++) { if (k < r) off += (size_t)(counts[k] + 0.5); total += (size_t)(counts[k] + 0.5); }
    std::vector<double> all(total*W, 0.0);
    std::copy(rows.begin(), rows.end(), all.begin() + off*W);
    if (total) ck(self->ctx(), hcg_allreduce(self->ctx(), all.data(), (int64_t)all.size(), 0), "hcg_allreduce");
    rows.swap(all);
  }
  if (r != 0) return;
  std::vector<std::ofstream> csv(self->cellfields->size());
  for (unsigned i = 0; i < self->cellfields->size(); i++) {
    csv[i].open(out + "/csv/" + (*self->cellfields)[i]->name + "." + zeroPadNumber(self->iter) + ".csv", std::ofstream::trunc);
    csv[i] << "X,Y,Z,area,volume,atomic_block,cellId,baseCellId,velocity_x,velocity_y,velocity_z" << endl;
  }
  for (size_t k = 0; k < rows.size()/W; k++) {
    const double* q = rows.data() + k*W;
    auto& o = csv[(size_t)(q[0] + 0.5)];
    o << q[1] << "," << q[2] << "," << q[3] << "," << q[4] << "," << q[5] << "," << (long)q[6] << "," << (long)q[7] << "," << (long)q[8] << ","
      << q[9] << "," << q[10] << "," << q[11] << endl;
  }
}

void HemoCell::writeOutput() {
  const double el = global.statistics.elapsed();
  const std::string tpi = (iter != lastOutputAt) ? Profiler::toString((el - lastOutput)/(iter - lastOutputAt)) : "0.00";
  lastOutput = el; lastOutputAt = iter;
  pcout << "(HemoCell) (Output) writing output at timestep " << iter << " (" << param::dt * iter << " s). Approx. performance: " << tpi << " s / iteration." << endl;
  if (!sanityCheckDone) pushSettings();
  hcg_ctx* c = ctx();
  // the reference recomputes repulsion and (forced) membrane forces at every output: part of the trajectory (core/hemoCell.cpp:236-262)
  if (repulsionEnabled) cellfields->applyRepulsionForce();
  if (boundaryRepulsionEnabled) cellfields->applyBoundaryRepulsionForce();
  cellfields->separate_force_vectors();
  cellfields->applyConstitutiveModel(true);
  const std::string out = plb::global::directories().getOutputDir();
  // every rank creates the (shared) directories itself: ranks only meet inside the device exchanges, not on the host
  mkpath(out + "/hdf5/" + zeroPadNumber(iter)); mkpath(out + "/csv");
  // HDF5 particle files per cell type and the fluid file of this rank's block (io/ParticleHdf5IO.cpp, io/FluidHdf5IO.hh)
  for (unsigned i = 0; i < cellfields->size(); i++) write_particle_h5(*this, *(*cellfields)[i]);
  write_fluid_h5(*this);
  writeCellInfo_CSV(*this);
  cellfields->unify_force_vectors();
  (void)c;
}

// CellInformationFunctionals / FluidInfo ------------------------------------------------------------------
void CellInformationFunctionals::calculateCellVolume(HemoCell* h) {
  CellSnapshot s = snapshot(h); if (!s.nc) return;
  std::vector<double> vol(s.nc), area(s.nc);
  ck(h->ctx(), hcg_cells_volume_area(h->ctx(), vol.data(), area.data()), "hcg_cells_volume_area");
  for (int64_t k = 0; k < s.nc; k++) if (s.alive[k] && s.ids[k] >= 0) entry(s, k).volume = vol[k];
}
void CellInformationFunctionals::calculateCellArea(HemoCell* h) {
  CellSnapshot s = snapshot(h); if (!s.nc) return;
  std::vector<double> vol(s.nc), area(s.nc);
  ck(h->ctx(), hcg_cells_volume_area(h->ctx(), vol.data(), area.data()), "hcg_cells_volume_area");
  for (int64_t k = 0; k < s.nc; k++) if (s.alive[k] && s.ids[k] >= 0) entry(s, k).area = area[k];
}
void CellInformationFunctionals::calculateCellBoundingBox(HemoCell* h) {
  CellSnapshot s = snapshot(h); if (!s.nc) return;
  std::vector<double> bb(6*s.nc);
  ck(h->ctx(), hcg_cells_bbox(h->ctx(), bb.data()), "hcg_cells_bbox");
  for (int64_t k = 0; k < s.nc; k++) if (s.alive[k] && s.ids[k] >= 0) for (int d = 0; d < 6; d++) entry(s, k).bbox[d] = bb[6*k + d];
}
void CellInformationFunctionals::calculateCellPosition(HemoCell* h) {
  CellSnapshot s = snapshot(h); if (!s.np) return;
  std::vector<double> pos(3*s.np), vel(3*s.np);
  ck(h->ctx(), hcg_cells_download(h->ctx(), HCG_P_POS, pos.data()), "download");
  ck(h->ctx(), hcg_cells_download(h->ctx(), HCG_P_VEL, vel.data()), "download");
  for (int64_t k = 0; k < s.nc; k++) if (s.alive[k] && s.ids[k] >= 0) {
    const int V = (*h->cellfields)[(unsigned)s.types[k]]->numVertex;
    CellInformation& ci = entry(s, k);
    for (int d = 0; d < 3; d++) {
      double a = 0, b = 0;
      for (int v = 0; v < V; v++) { a += pos[3*(s.base[k]+v)+d]; b += vel[3*(s.base[k]+v)+d]; }
      ci.position[d] = a/V; ci.velocity[d] = b/V;
    }
  }
}
void CellInformationFunctionals::calculateCellStretch(HemoCell* h) {
  // max pairwise vertex distance per cell (helper/cellInfo.cpp:103-121), one CTA per cell on the device
  CellSnapshot s = snapshot(h); if (!s.nc) return;
  std::vector<double> st(s.nc);
  ck(h->ctx(), hcg_cells_stretch(h->ctx(), st.data()), "hcg_cells_stretch");
  for (int64_t k = 0; k < s.nc; k++) if (s.alive[k] && s.ids[k] >= 0) entry(s, k).stretch = st[k];
}
void CellInformationFunctionals::calculateCellInformation(HemoCell* h) {
  clear_list();
  calculateCellVolume(h); calculateCellArea(h); calculateCellPosition(h); calculateCellBoundingBox(h);
}
pluint CellInformationFunctionals::getTotalNumberOfCells(HemoCell* h) {
  int64_t n = 0, p = 0;
  ck(h->ctx(), hcg_cells_count(h->ctx(), &n, &p), "hcg_cells_count");
  double g = (double)n;                                    // every cell is counted by the one rank that owns it
  ck(h->ctx(), hcg_allreduce(h->ctx(), &g, 1, 0), "hcg_allreduce");
  if (h->preInlet && h->preInlet->pre) {                   // the reference gathers over all ranks, the pre-inlet's included
    hcg_ctx* q = h->preInlet->preCtx();
    int64_t qn = 0, qp = 0;
    ck(q, hcg_cells_count(q, &qn, &qp), "hcg_cells_count");
    g += (double)qn;
  }
  return (pluint)(g + 0.5);
}
pluint CellInformationFunctionals::getNumberOfCellsFromType(HemoCell* h, std::string type) {
  CellSnapshot s = snapshot(h);
  const int t = (*h->cellfields)[type]->impl->device_ctype;
  pluint n = 0;
  for (int64_t k = 0; k < s.nc; k++) if (s.alive[k] && s.owned[k] && s.ids[k] >= 0 && s.types[k] == t) n++;
  double g = (double)n;
  ck(h->ctx(), hcg_allreduce(h->ctx(), &g, 1, 0), "hcg_allreduce");
  if (h->preInlet && h->preInlet->pre) {
    hcg_ctx* q = h->preInlet->preCtx();
    int64_t qc = 0, qp = 0;
    ck(q, hcg_cells_capacity(q, &qc, &qp), "hcg_cells_capacity");
    std::vector<int64_t> ids(qc); std::vector<int32_t> types(qc); std::vector<uint8_t> alive(qc);
    if (qc) ck(q, hcg_cells_info(q, ids.data(), types.data(), alive.data()), "hcg_cells_info");
    for (int64_t k = 0; k < qc; k++) if (alive[k] && ids[k] >= 0 && types[k] == t) g += 1.0;
  }
  return (pluint)(g + 0.5);
}
// helper/particleInfo.cpp:28-119
static ParticleStatistics particle_stats(HemoCell* h, bool force) {
  CellSnapshot s = snapshot(h);
  ParticleStatistics r;
  if (!s.np && plb::global::mpi().getSize() == 1) return r;
  std::vector<double> a(3*s.np), b;
  if (s.np) ck(h->ctx(), hcg_cells_download(h->ctx(), force ? HCG_P_FORCE : HCG_P_VEL, a.data()), "download");
  if (force && s.np) { b.resize(3*s.np); ck(h->ctx(), hcg_cells_download(h->ctx(), HCG_P_FREP, b.data()), "download"); }
  bool first = true; double sum = 0;
  for (int64_t k = 0; k < s.nc; k++) if (s.alive[k] && s.owned[k] && s.ids[k] >= 0) {
    const int V = (*h->cellfields)[(unsigned)s.types[k]]->numVertex;
    for (int v = 0; v < V; v++) {
      const int64_t q = 3*(s.base[k] + v);
      double x = a[q], y = a[q+1], z = a[q+2];
      if (force) { x += b[q]; y += b[q+1]; z += b[q+2]; }
      const double m = std::sqrt(x*x + y*y + z*z);
      if (first) { r.min = r.max = m; first = false; }
      r.min = std::min(r.min, m); r.max = std::max(r.max, m); sum += m; r.ncells++;
    }
  }
  if (plb::global::mpi().getSize() > 1) {
    double tot[2] = {sum, (double)r.ncells}, mn = first ? 1e300 : r.min, mx = first ? -1e300 : r.max;
    ck(h->ctx(), hcg_allreduce(h->ctx(), tot, 2, 0), "hcg_allreduce");
    ck(h->ctx(), hcg_allreduce(h->ctx(), &mn, 1, 1), "hcg_allreduce");
    ck(h->ctx(), hcg_allreduce(h->ctx(), &mx, 1, 2), "hcg_allreduce");
    sum = tot[0]; r.ncells = (pluint)(tot[1] + 0.5); r.min = mn; r.max = mx;
  }
  if (r.ncells) r.avg = sum/r.ncells;
  return r;
}
ParticleStatistics ParticleInfo::calculateVelocityStatistics(HemoCell* h) { return particle_stats(h, false); }
ParticleStatistics ParticleInfo::calculateForceStatistics(HemoCell* h) { return particle_stats(h, true); }

// helper/hemoCellStretch.cpp ------------------------------------------------------------------------------
vector<plint> HemoCellStretch::lower_lsps = vector<plint>();
vector<plint> HemoCellStretch::upper_lsps = vector<plint>();
unsigned int HemoCellStretch::n_forced_lsps = 0;
T HemoCellStretch::external_force = 0.0;
T HemoCellStretch::scale = 1.0;
HemoCellStretch::HemoCellStretch(HemoCellField& cellfield_, unsigned int n_forced_lsps_, T external_force_) : cellfield(cellfield_) {
  HemoCell* h = &cellfield.cellFields.hemocell;
  if (CellInformationFunctionals::getTotalNumberOfCells(h) != 1) { pcout << "(HemoCellStretch) Refusing to run with more or less than 1 cell" << endl; exit(1); }
  n_forced_lsps = n_forced_lsps_;
  external_force = external_force_/n_forced_lsps;
  // the n LSPs with the smallest / largest x of cell 0 (stable order of the reference's bubble sort: ties keep vertex order)
  CellSnapshot s = snapshot(h);
  std::vector<double> pos(3*s.np);
  ck(h->ctx(), hcg_cells_download(h->ctx(), HCG_P_POS, pos.data()), "download");
  int64_t k0 = -1; for (int64_t k = 0; k < s.nc; k++) if (s.alive[k] && s.ids[k] == 0) k0 = k;
  if (k0 < 0) { cout << "Error -1 found in cell, exiting" << endl; exit(1); }
  const int V = cellfield.numVertex;
  std::vector<int> idx(V); for (int v = 0; v < V; v++) idx[v] = v;
  std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return pos[3*(s.base[k0] + a)] < pos[3*(s.base[k0] + b)]; });
  for (unsigned i = 0; i < n_forced_lsps; i++) { lower_lsps.push_back(idx[i]); upper_lsps.push_back(idx[V - 1 - i]); }
}
void HemoCellStretch::applyForce() {
  if (cellfield.timescale != 1) { pcout << "Refusing to stretch with particle update timestep larger than 1" << endl; exit(1); }
  HemoCell* h = &cellfield.cellFields.hemocell;
  // the single cell never changes its storage slot: resolve the particle indices once
  static std::vector<int64_t> which; static hcg_ctx* which_ctx = nullptr;
  if (which_ctx != h->ctx() || which.size() != lower_lsps.size() + upper_lsps.size()) {
    CellSnapshot s = snapshot(h);
    int64_t k0 = -1; for (int64_t k = 0; k < s.nc; k++) if (s.alive[k] && s.ids[k] == 0) k0 = k;
    if (k0 < 0) return;
    which.clear();
    for (plint v : lower_lsps) which.push_back(s.base[k0] + v);
    for (plint v : upper_lsps) which.push_back(s.base[k0] + v);
    which_ctx = h->ctx();
  }
  std::vector<double> f;
  for (size_t i = 0; i < lower_lsps.size(); i++) { f.push_back(-external_force*scale); f.push_back(0); f.push_back(0); }
  for (size_t i = 0; i < upper_lsps.size(); i++) { f.push_back(external_force*scale); f.push_back(0); f.push_back(0); }
  ck(h->ctx(), hcg_cells_add_force(h->ctx(), (int64_t)which.size(), which.data(), f.data()), "hcg_cells_add_force");
}

FluidStatistics FluidInfo::calculateVelocityStatistics(HemoCell* h) {
  FluidStatistics f;
  ck(h->ctx(), hcg_fluid_velocity_stats(h->ctx(), &f.min, &f.max, &f.avg), "hcg_fluid_velocity_stats");
  if (plb::global::mpi().getSize() > 1) {
    // the device statistics cover this rank's slab: weigh the means by the slabs' non-boundary node counts
    GpuLattice* g = h->lattice->gpu();
    const int64_t P = (int64_t)g->ny*g->nz; const uint8_t* fl = g->flags.data() + (int64_t)g->x0()*P;
    double nfluid = 0; for (int64_t i = 0; i < (int64_t)g->nxl()*P; i++) nfluid += fl[i] != HCG_BOUNCEBACK;
    double sum[2] = {f.avg*nfluid, nfluid}, mn = f.min, mx = f.max;
    ck(h->ctx(), hcg_allreduce(h->ctx(), sum, 2, 0), "hcg_allreduce");
    ck(h->ctx(), hcg_allreduce(h->ctx(), &mn, 1, 1), "hcg_allreduce");
    ck(h->ctx(), hcg_allreduce(h->ctx(), &mx, 1, 2), "hcg_allreduce");
    f.avg = sum[1] > 0 ? sum[0]/sum[1] : 0; f.min = mn; f.max = mx;
  }
  return f;
}

}  // namespace hemo

// ====================================================================================================
// PreInlet (helper/preInlet.cpp): the periodic pre-inlet as a second device context of the same process
// ====================================================================================================
namespace hemo {
using plb::Box3D;

PreInlet::PreInlet(HemoCell* hemocell_, plb::MultiScalarField3D<int>* flagMatrix_) : hemocell(hemocell_), flagMatrix(flagMatrix_) {
  preinlet_length = (*hemocell->cfg)["preInlet"]["parameters"]["lengthN"].read<int>();
}
PreInlet::~PreInlet() {
  if (hemocell && hemocell->lattice && hemocell->lattice->gpu()->companion == pre) hemocell->lattice->gpu()->companion = nullptr;
  delete pre;
}
hcg_ctx* PreInlet::preCtx() {
  if (!pre) fatal("(PreInlet) the pre-inlet lattice does not exist yet: call preInletFromSlice / autoPreinletFromBoundary before initializeLattice");
  pre->materialize();
  return pre->ctx;
}

// helper/preInlet.cpp:463-560 / 592-700: bounding box of the fluid nodes of the slice, one node of solid around it, preinlet_length
// planes outwards.  (The transverse extent is always clipped to the flag matrix here; the reference clips it to the lattice when one exists.)
void PreInlet::locate(Box3D slice) {
  inflow_length = (*hemocell->cfg)["domain"]["particleEnvelope"].read<int>();
  const Box3D bb = flagMatrix->getBoundingBox();
  bool found = false; Box3D d;
  for (plint x = std::max(slice.x0, bb.x0); x <= std::min(slice.x1, bb.x1); x++)
    for (plint y = std::max(slice.y0, bb.y0); y <= std::min(slice.y1, bb.y1); y++)
      for (plint z = std::max(slice.z0, bb.z0); z <= std::min(slice.z1, bb.z1); z++) {
        if (!flagMatrix->get(x, y, z)) continue;
        if (!found) { d = Box3D(x, x, y, y, z, z); found = true; }
        else { d.x0 = std::min(d.x0, x); d.x1 = std::max(d.x1, x); d.y0 = std::min(d.y0, y); d.y1 = std::max(d.y1, y); d.z0 = std::min(d.z0, z); d.z1 = std::max(d.z1, z); }
      }
  if (!found) { hlog << "(PreInlet) no preinlet found, is it in the correct location?" << endl; exit(1); }
  location = d.enlarge(1);
  switch (direction) {
    case Direction::Xneg: location.x0 -= preinlet_length; break;
    case Direction::Yneg: location.y0 -= preinlet_length; break;
    case Direction::Zneg: location.z0 -= preinlet_length; break;
    case Direction::Xpos: location.x1 += preinlet_length; break;
    case Direction::Ypos: location.y1 += preinlet_length; break;
    case Direction::Zpos: location.z1 += preinlet_length; break;
  }
  if (!hemocell->lattice) hlog << "(PreInlet) preInlet located before the lattice exists: transverse extent clipped to the flag matrix" << endl;
  if (axis() != 0) { location.x0 = std::max(location.x0, bb.x0); location.x1 = std::min(location.x1, bb.x1); }
  if (axis() != 1) { location.y0 = std::max(location.y0, bb.y0); location.y1 = std::min(location.y1, bb.y1); }
  if (axis() != 2) { location.z0 = std::max(location.z0, bb.z0); location.z1 = std::min(location.z1, bb.z1); }
}

void PreInlet::preInletFromSlice(Direction direction_, Box3D boundary) {
  direction = direction_; initialized = true;
  const bool flat = axis() == 0 ? boundary.x0 == boundary.x1 : (axis() == 1 ? boundary.y0 == boundary.y1 : boundary.z0 == boundary.z1);
  if (!flat) { hlog << "Not a flat slice, refusing to create preInlet" << endl; exit(1); }
  locate(boundary);
}

void PreInlet::autoPreinletFromBoundary(Direction dir_) {
  direction = dir_; initialized = true;
  Box3D f = flagMatrix->getBoundingBox();                       // the second plane from the face the pre-inlet attaches to (helper/preInlet.cpp:594-618)
  switch (direction) {
    case Direction::Xneg: f.x1 = f.x0 + 1; f.x0 = f.x1; break;
    case Direction::Yneg: f.y1 = f.y0 + 1; f.y0 = f.y1; break;
    case Direction::Zneg: f.z1 = f.z0 + 1; f.z0 = f.z1; break;
    case Direction::Xpos: f.x0 = f.x1 - 1; f.x1 = f.x0; break;
    case Direction::Ypos: f.y0 = f.y1 - 1; f.y1 = f.y0; break;
    case Direction::Zpos: f.z0 = f.z1 - 1; f.z1 = f.z0; break;
  }
  locate(f);
}

void PreInlet::createLattice(double omega) {
  if (!initialized) fatal("(PreInlet) preInletFromSlice / autoPreinletFromBoundary must be called before initializeLattice");
  if (plb::global::mpi().getSize() > 1) fatal("(PreInlet) the pre-inlet runs inside a single process next to the main domain; launch one rank");
  delete pre;
  pre = new GpuLattice((int)location.getNx(), (int)location.getNy(), (int)location.getNz(), omega);
  pre->solo = true;
  if (const char* e = getenv("HEMOCELL_PREINLET_DEVICE")) pre->device_override = atoi(e);
  hemocell->lattice->gpu()->companion = pre;
  hlog << "(PreInlet) pre-inlet lattice " << pre->nx << " x " << pre->ny << " x " << pre->nz << " at (" << location.x0 << ", " << location.y0
       << ", " << location.z0 << "), as a second device context of this process" << endl;
}

// helper/preInlet.cpp:399-436: the plane of the main domain the pre-inlet feeds becomes Zou-He velocity nodes (fluid nodes only)
void PreInlet::initializePreInletVelocityBoundary() {
  const Box3D bb = flagMatrix->getBoundingBox();
  Box3D d(std::max(bb.x0, location.x0), std::min(bb.x1, location.x1), std::max(bb.y0, location.y0), std::min(bb.y1, location.y1),
          std::max(bb.z0, location.z0), std::min(bb.z1, location.z1));
  switch (direction) {
    case Direction::Xneg: d.x0 = d.x1; break;  case Direction::Yneg: d.y0 = d.y1; break;  case Direction::Zneg: d.z0 = d.z1; break;
    case Direction::Xpos: d.x1 = d.x0; break;  case Direction::Ypos: d.y1 = d.y0; break;  case Direction::Zpos: d.z1 = d.z0; break;
  }
  fluidInlet = d;
  // outward normal of the inlet plane: towards the pre-inlet.  orientation index 0..5 = -x +x -y +y -z +z
  const int orientation = 2*axis() + ((int)direction % 2 == 0 ? 1 : 0);      // Xpos -> +x, Xneg -> -x, ...
  GpuLattice* g = hemocell->lattice->gpu();
  const double zero[3] = {0, 0, 0};
  for (plint x = d.x0; x <= d.x1; x++) for (plint y = d.y0; y <= d.y1; y++) for (plint z = d.z0; z <= d.z1; z++) {
    if (flagMatrix->get(x, y, z) != 1) continue;
    const Box3D point(x, x, y, y, z, z);
    gpu_lattice_zouhe(g, point, 0, orientation);
    gpu_lattice_boundary_velocity(g, point, zero);
  }
}
void PreInlet::initializePreInletParticleBoundary() {}      // (the reference maps MPI senders to receivers here)

// helper/preInlet.cpp:950-993 on the pre-inlet side: solid wherever the inlet cross-section is solid, periodic along the flow
void PreInlet::createBoundary() {
  if (!pre) fatal("(PreInlet) createBoundary before initializeLattice");
  const Box3D bb = flagMatrix->getBoundingBox();
  for (int lx = 0; lx < pre->nx; lx++) for (int ly = 0; ly < pre->ny; ly++) for (int lz = 0; lz < pre->nz; lz++) {
    plint gx = lx + location.x0, gy = ly + location.y0, gz = lz + location.z0;
    if (axis() == 0) gx = fluidInlet.x0; else if (axis() == 1) gy = fluidInlet.y0; else gz = fluidInlet.z0;
    const bool inside = gx >= bb.x0 && gx <= bb.x1 && gy >= bb.y0 && gy <= bb.y1 && gz >= bb.z0 && gz <= bb.z1;
    const bool fluid = inside && flagMatrix->get(gx, gy, gz) != 0;
    if (!fluid) pre->flags[pre->idx(lx, ly, lz)] = HCG_BOUNCEBACK;
  }
  pre->touchFlags();
  pre->periodic[0] = pre->periodic[1] = pre->periodic[2] = false;
  pre->periodic[axis()] = true;
}

// helper/preInlet.cpp:756-804: Poiseuille force for the Reynolds number of config.xml on the pre-inlet's cross-section
void PreInlet::calculateDrivingForce() {
  if (!pre) fatal("(PreInlet) calculateDrivingForce before initializeLattice");
  const double re = (*hemocell->cfg)["preInlet"]["parameters"]["Re"].read<T>();
  const int n[3] = {pre->nx, pre->ny, pre->nz};
  const int plane = ((int)direction % 2 == 1) ? 2 : n[axis()] - 1 - 2;        // N: third plane from the low end, P: from the high end
  plint fluidArea = 0;
  for (int lx = 0; lx < pre->nx; lx++) for (int ly = 0; ly < pre->ny; ly++) for (int lz = 0; lz < pre->nz; lz++) {
    const int l[3] = {lx, ly, lz};
    if (l[axis()] != plane) continue;
    if (pre->flags[pre->idx(lx, ly, lz)] == HCG_FLUID) fluidArea++;
  }
  const T pipe_radius = std::sqrt(fluidArea/PI);
  hlog << "(Parameters) Your preInlet pipe has a calculated radius of " << pipe_radius << " LU, assuming a perfect circle" << std::endl;
  const T u_lbm_max = re*param::nu_lbm/(pipe_radius*2);
  drivingForce = 8*param::nu_lbm*(u_lbm_max*0.5)/pipe_radius/pipe_radius;
}
void PreInlet::applyForce(double f) {
  if (!pre) return;
  double v[3] = {0, 0, 0};
  v[axis()] = ((int)direction % 2 == 1) ? f : -f;             // N: the pre-inlet sits on the negative side and pushes in +, P: the other way
  gpu_lattice_external_vector(pre, Box3D(0, pre->nx - 1, 0, pre->ny - 1, 0, pre->nz - 1), v);
}
void PreInlet::setDrivingForce() { applyForce(drivingForce); }

double PreInlet::average(vector<double> values) {
  double s = 0.0;
  for (double v : values) s += v;
  return s/(double)values.size();
}
// helper/preInlet.cpp:820-857: "<time> <normalised velocity>" lines of the pulse file
bool PreInlet::readNormalizedVelocities() {
  const std::string name = (*hemocell->cfg)["preInlet"]["parameters"]["pulseFileName"].read<std::string>();
  std::ifstream in(name);
  if (!in.is_open()) { cout << "*** WARNING! pulsatility data file " << name << " does not exist!" << endl; return false; }
  double t, v;
  while (in >> t >> v) { normalizedVelocityTimes.push_back(t); normalizedVelocityValues.push_back(v); }
  if (normalizedVelocityTimes.size() < 2) return false;
  average_vel = average(normalizedVelocityValues);
  pulseEndTime = normalizedVelocityTimes.back();
  try { pFrequency = (*hemocell->cfg)["preInlet"]["parameters"]["pFrequency"].read<double>(); }
  catch (const std::invalid_argument&) { pFrequency = 1.0/pulseEndTime; }
  return true;
}
// piecewise-linear look-up, clamped at both ends unless `extrapolate` (helper/preInlet.cpp:860-890)
double PreInlet::interpolate(vector<double>& xData, vector<double>& yData, double x, bool extrapolate) {
  const int n = (int)xData.size();
  int i = 0;
  if (x >= xData[n - 2]) i = n - 2; else while (x > xData[i + 1]) i++;
  double yl = yData[i], yr = yData[i + 1];
  if (!extrapolate) { if (x < xData[i]) yr = yl; if (x > xData[i + 1]) yl = yr; }
  return yl + (yr - yl)/(xData[i + 1] - xData[i])*(x - xData[i]);
}
void PreInlet::setDrivingForceTimeDependent(double t) {
  t = std::fmod(t*pFrequency*pulseEndTime, pulseEndTime);     // position inside the (periodic) pulse
  const double v = interpolate(normalizedVelocityTimes, normalizedVelocityValues, t, false);
  applyForce(v/average_vel*drivingForce);
}

void PreInlet::iterate() {
  if (!pre) return;
  hcg_ctx* q = preCtx();
  ck(q, hcg_iterate(q, 1), "hcg_iterate (pre-inlet)");
}

// node pairs of the coupling plane: fluid nodes of the inlet plane <-> the same nodes in the pre-inlet's local coordinates
void PreInlet::coupleNodes() {
  GpuLattice* g = hemocell->lattice->gpu();
  std::vector<int64_t> pi, mi;
  const Box3D& d = fluidInlet;
  for (plint x = d.x0; x <= d.x1; x++) for (plint y = d.y0; y <= d.y1; y++) for (plint z = d.z0; z <= d.z1; z++) {
    const uint8_t f = g->flags[g->idx((int)x, (int)y, (int)z)];
    if (f < HCG_ZH_VEL_XN || f > HCG_ZH_VEL_ZP) continue;
    const int lx = (int)(x - location.x0), ly = (int)(y - location.y0), lz = (int)(z - location.z0);
    if (pre->flags[pre->idx(lx, ly, lz)] != HCG_FLUID) continue;
    mi.push_back(g->idx((int)x, (int)y, (int)z)); pi.push_back(pre->idx(lx, ly, lz));
  }
  hcg_ctx* c = hemocell->ctx();
  ck(c, hcg_preinlet_map(c, preCtx(), (int64_t)mi.size(), pi.data(), mi.data()), "hcg_preinlet_map");
  hlog << "(PreInlet) " << mi.size() << " inlet nodes coupled to the pre-inlet" << endl;
  coupled = true;
}

// helper/preInlet.cpp:344-397
void PreInlet::applyPreInletVelocityBoundary() {
  if (!pre) return;
  // what the pre-inlet ranks of the reference do between iterate() and applyPreInlet() (`if (hemocell.partOfpreInlet) preInlet->setDrivingForce()`)
  if (!normalizedVelocityTimes.empty()) setDrivingForceTimeDependent(hemocell->iter*param::dt); else setDrivingForce();
  if (!coupled) coupleNodes();
  hcg_ctx* c = hemocell->ctx();
  ck(c, hcg_preinlet_apply_velocity(c), "hcg_preinlet_apply_velocity");
}

// helper/preInlet.cpp:255-342, in whole cells (include/hemocell_gpu.h: hcg_preinlet_apply_cells)
void PreInlet::applyPreInletParticleBoundary() {
  if (!pre || !hemocell->cellfields || hemocell->cellfields->size() == 0) return;
  if (!coupled) coupleNodes();
  static int every = -1;
  if (every < 0) { const char* e = getenv("HEMOCELL_PREINLET_EVERY"); every = e ? std::max(1, atoi(e)) : 1; }
  if (hemocell->iter % (unsigned)every != 0) return;
  const int a = axis();
  const double inlet = a == 0 ? fluidInlet.x0 : (a == 1 ? fluidInlet.y0 : fluidInlet.z0);
  const bool neg = (int)direction % 2 == 1;                     // pre-inlet on the negative side: the slab lies on the positive side of the inlet plane
  // the slab stops one node short of the inlet plane: the Zou-He inlet nodes count as boundary nodes for the IBM here, and a
  // vertex whose nearest node is one of them would delete its cell at the next advance (hemoCellParticleField.cpp:579-584)
  const double lo = neg ? inlet + 1 : inlet - inflow_length, hi = neg ? inlet + inflow_length : inlet - 1;
  const double shift[3] = {(double)location.x0, (double)location.y0, (double)location.z0};
  const double period = a == 0 ? pre->nx : (a == 1 ? pre->ny : pre->nz);
  int64_t n = 0;
  hcg_ctx* c = hemocell->ctx();
  ck(c, hcg_preinlet_apply_cells(c, a, period, shift, lo, hi, std::max<plint>(1, hemocell->cellfields->number_of_cells), &n), "hcg_preinlet_apply_cells");
  cellsHandedOver += n;
  if (n) { hemocell->lattice->gpu()->has_cells = true; hlog << "(PreInlet) iteration " << hemocell->iter << ": " << n << " cell(s) handed over to the main domain (" << cellsHandedOver << " so far)" << endl; }
}

// The pre-inlet's share of a checkpoint (PRE_lattice / PRE_particleField in the reference, core/hemoCellFields.cpp:297-314):
// populations, node force, live cells with their state, the hand-over bookkeeping and the id stride, in pre.bin beside the
// main domain's rank file (rotated to .old like it).
void PreInlet::saveCheckPoint(const std::string& dir) {
  hcg_ctx* q = preCtx();
  hcg_ctx* c = hemocell->ctx();
  if (!coupled) coupleNodes();
  const std::string base = dir + "pre";
  if (file_exists(base + ".bin")) rename((base + ".bin").c_str(), (base + ".bin.old").c_str());
  std::ofstream f(base + ".bin", std::ios::binary);
  const int64_t Nl = (int64_t)pre->nx*pre->ny*pre->nz;
  int64_t nc = 0, np = 0;
  ck(q, hcg_cells_capacity(q, &nc, &np), "hcg_cells_capacity");
  const int64_t hdr[8] = {0x48434750, (int64_t)hemocell->iter, Nl, nc, np, (int64_t)hemocell->cellfields->size(),
                          (int64_t)hemocell->cellfields->number_of_cells, cellsHandedOver};
  f.write((const char*)hdr, sizeof(hdr));
  std::vector<double> buf((size_t)std::max<int64_t>(19*Nl, 3*np));
  ck(q, hcg_lattice_download(q, HCG_LAT_POP, buf.data()), "download"); f.write((const char*)buf.data(), 8*19*Nl);
  ck(q, hcg_lattice_download(q, HCG_LAT_FORCE, buf.data()), "download"); f.write((const char*)buf.data(), 8*3*Nl);
  for (int fld : {HCG_P_POS, HCG_P_VEL, HCG_P_FORCE, HCG_P_FREP}) { ck(q, hcg_cells_download(q, fld, buf.data()), "download"); f.write((const char*)buf.data(), 8*3*np); }
  std::vector<int64_t> ids(nc), laps(nc); std::vector<int32_t> types(nc); std::vector<uint8_t> alive(nc);
  ck(q, hcg_cells_info(q, ids.data(), types.data(), alive.data()), "hcg_cells_info");
  ck(c, hcg_preinlet_laps(c, nc, laps.data(), 0), "hcg_preinlet_laps");
  f.write((const char*)ids.data(), 8*nc); f.write((const char*)types.data(), 4*nc); f.write((const char*)alive.data(), nc);
  f.write((const char*)laps.data(), 8*nc);
}
void PreInlet::loadCheckPoint(const std::string& dir) {
  hcg_ctx* q = preCtx();
  hcg_ctx* c = hemocell->ctx();
  std::ifstream f(dir + "pre.bin", std::ios::binary);
  if (!f) fatal("(PreInlet) cannot open the pre-inlet's checkpoint data " + dir + "pre.bin");
  int64_t hdr[8]; f.read((char*)hdr, sizeof(hdr));
  const int64_t Nl = (int64_t)pre->nx*pre->ny*pre->nz;
  HemoCellFields& cf = *hemocell->cellfields;
  if (hdr[0] != 0x48434750 || hdr[2] != Nl || hdr[5] != (int64_t)cf.size()) fatal("(PreInlet) checkpoint does not match this pre-inlet / cell types");
  const int64_t nc = hdr[3], np = hdr[4];
  cf.number_of_cells = (plint)hdr[6]; cellsHandedOver = hdr[7];
  std::vector<double> pop(19*Nl), frc(3*Nl);
  f.read((char*)pop.data(), 8*19*Nl); f.read((char*)frc.data(), 8*3*Nl);
  std::vector<double> P[4];
  for (auto& v : P) { v.resize(3*np); f.read((char*)v.data(), 8*3*np); }
  std::vector<int64_t> ids(nc), laps(nc); std::vector<int32_t> types(nc); std::vector<uint8_t> alive(nc);
  f.read((char*)ids.data(), 8*nc); f.read((char*)types.data(), 4*nc); f.read((char*)alive.data(), nc); f.read((char*)laps.data(), 8*nc);
  std::vector<int64_t> base_of(nc); { int64_t p = 0; for (int64_t k = 0; k < nc; k++) { base_of[k] = p; p += cf[(unsigned)types[k]]->numVertex; } }
  std::vector<double> vel, force, frep; std::vector<int64_t> new_laps;
  for (auto* fl : cf.cellFields) {
    hcg_celltype t = device_celltype(fl);
    int32_t qid = -1;
    ck(q, hcg_celltype_add(q, &t, &qid), "hcg_celltype_add");
    if (qid != fl->impl->device_ctype) fatal("(PreInlet) cell types of the pre-inlet and the main domain are out of step");
    std::vector<int64_t> tid; std::vector<double> tpos;
    for (int64_t k = 0; k < nc; k++) if (types[k] == qid && alive[k] && ids[k] >= 0) {
      tid.push_back(ids[k]); new_laps.push_back(laps[k]);
      const int V = fl->numVertex;
      tpos.insert(tpos.end(), P[0].begin() + 3*base_of[k], P[0].begin() + 3*(base_of[k] + V));
      vel.insert(vel.end(), P[1].begin() + 3*base_of[k], P[1].begin() + 3*(base_of[k] + V));
      force.insert(force.end(), P[2].begin() + 3*base_of[k], P[2].begin() + 3*(base_of[k] + V));
      frep.insert(frep.end(), P[3].begin() + 3*base_of[k], P[3].begin() + 3*(base_of[k] + V));
    }
    ck(q, hcg_cells_add(q, qid, (int64_t)tid.size(), tid.data(), tpos.data()), "hcg_cells_add");
    pre->has_cells = pre->has_cells || !tid.empty();
    int64_t cap_c = 0, cap_p = 0;
    ck(q, hcg_cells_capacity(q, &cap_c, &cap_p), "hcg_cells_capacity");
    vel.resize((size_t)3*cap_p, 0.0); force.resize((size_t)3*cap_p, 0.0); frep.resize((size_t)3*cap_p, 0.0);
    new_laps.resize((size_t)cap_c, INT64_MIN);
  }
  ck(q, hcg_set_force_limit(q, param::f_limit), "hcg_set_force_limit");
  if (!vel.empty()) {
    ck(q, hcg_cells_upload(q, HCG_P_VEL, vel.data()), "upload"); ck(q, hcg_cells_upload(q, HCG_P_FORCE, force.data()), "upload");
    ck(q, hcg_cells_upload(q, HCG_P_FREP, frep.data()), "upload");
  }
  setDrivingForce();                                  // the value the node force is reset to; the saved node force follows
  ck(q, hcg_lattice_upload(q, HCG_LAT_POP, pop.data()), "upload"); ck(q, hcg_lattice_upload(q, HCG_LAT_FORCE, frc.data()), "upload");
  pre->eq_pending = false; pre->body_pending = false;
  if (!coupled) coupleNodes();
  // live cells were re-created in slot order: their hand-over marks follow them
  ck(c, hcg_preinlet_laps(c, (int64_t)new_laps.size(), new_laps.data(), 1), "hcg_preinlet_laps");
  ck(c, hcg_preinlet_apply_velocity(c), "hcg_preinlet_apply_velocity");      // the inlet nodes' velocities are not in the main rank file
}

// helper/genericFunctions.cpp:138-163: bounce-back wherever the flag matrix is solid (main domain only)
void boundaryFromFlagMatrix(plb::MultiBlockLattice3D<T, DESCRIPTOR>* fluid, plb::MultiScalarField3D<int>* flagMatrix, bool partOfpreInlet) {
  if (partOfpreInlet) return;
  plb::defineDynamics(*fluid, *flagMatrix, flagMatrix->getBoundingBox(), new plb::BounceBack<T, DESCRIPTOR>(1.), 0);
}

}  // namespace hemo

// ====================================================================================================
// plb:: shim
// ====================================================================================================
namespace plb {

Parallel_ostream pcout(std::cout), pcerr(std::cerr);
void plbInit(int*, char***) {}
void plb_ofstream::open(const char* filename, std::ios_base::openmode mode) { if (global::mpi().isMainProcessor()) f.open(filename, mode); }
namespace global {
int MpiManager::getRank() const { return hemo::env_int("RANK", "OMPI_COMM_WORLD_RANK", 0); }
int MpiManager::getSize() const { return hemo::env_int("WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", 1); }
int MpiManager::getLocalRank() const { return hemo::env_int("LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", getRank()); }
void MpiManager::barrier() {}      // ranks meet in the NCCL / peer exchanges of the device path; host-side files are per rank
MpiManager& mpi() { static MpiManager m; return m; }
Directories& directories() { static Directories d; return d; }
}  // namespace global

}  // namespace plb

// non-template back end of the plb:: shim ---------------------------------------------------------------
namespace hemo {
using plb::Box3D;
namespace {
Box3D clip(const GpuLattice& g, Box3D b) {
  b.x0 = std::max<plint>(b.x0, 0); b.y0 = std::max<plint>(b.y0, 0); b.z0 = std::max<plint>(b.z0, 0);
  b.x1 = std::min<plint>(b.x1, g.nx - 1); b.y1 = std::min<plint>(b.y1, g.ny - 1); b.z1 = std::min<plint>(b.z1, g.nz - 1);
  return b;
}
// orientation of a face plane of the bounding box: 0..5 = -x +x -y +y -z +z, or -1
int plane_orientation(const GpuLattice& g, const Box3D& b) {
  if (b.x0 == b.x1) { if (b.x0 == 0) return 0; if (b.x0 == g.nx - 1) return 1; }
  if (b.y0 == b.y1) { if (b.y0 == 0) return 2; if (b.y0 == g.ny - 1) return 3; }
  if (b.z0 == b.z1) { if (b.z0 == 0) return 4; if (b.z0 == g.nz - 1) return 5; }
  return -1;
}
}  // namespace

GpuLattice* gpu_lattice_create(long nx, long ny, long nz, double omega) {
  if (nx < 1 || ny < 3 || nz < 3 || !(omega > 0 && omega < 2)) fatal("(HemoCell) (Fluid) invalid lattice size or relaxation frequency");
  return new GpuLattice((int)nx, (int)ny, (int)nz, omega);
}
void gpu_lattice_destroy(GpuLattice* g) { delete g; }
void gpu_lattice_size(const GpuLattice* g, long out[3]) { out[0] = g->nx; out[1] = g->ny; out[2] = g->nz; }
void gpu_lattice_set_periodic(GpuLattice* g, int axis, bool on) { if (axis >= 0 && axis < 3) g->periodic[axis] = on; }
bool gpu_lattice_get_periodic(const GpuLattice* g, int axis) { return axis >= 0 && axis < 3 && g->periodic[axis]; }
void gpu_lattice_collide_and_stream(GpuLattice* g) {
  g->materialize();
  ck(g->ctx, hcg_fluid_warmup(g->ctx, 1), "collideAndStream");
  if (g->companion) gpu_lattice_collide_and_stream(g->companion);     // the pre-inlet ranks of the reference run the same warm-up loop
}
void gpu_lattice_velocity_plane(GpuLattice* gp, const Box3D& plane_) {
  GpuLattice& g = *gp;
  const Box3D plane = clip(g, plane_);
  const int o = plane_orientation(g, plane);
  if (o < 0) fatal("(HemoCell) (BoundaryCondition) velocity conditions are supported on the face planes of the bounding box");
  for (plint x = plane.x0; x <= plane.x1; x++) for (plint y = plane.y0; y <= plane.y1; y++) for (plint z = plane.z0; z <= plane.z1; z++)
    g.flags[g.idx((int)x, (int)y, (int)z)] = (uint8_t)(HCG_VEL_XN + o);
  g.touchFlags();
}
void gpu_lattice_velocity_all_faces(GpuLattice* g) {
  // z faces last: rim nodes belong to the z planes, then y, then x (every rim node gets exactly one plane closure)
  gpu_lattice_velocity_plane(g, Box3D(0, 0, 0, g->ny - 1, 0, g->nz - 1));
  gpu_lattice_velocity_plane(g, Box3D(g->nx - 1, g->nx - 1, 0, g->ny - 1, 0, g->nz - 1));
  gpu_lattice_velocity_plane(g, Box3D(0, g->nx - 1, 0, 0, 0, g->nz - 1));
  gpu_lattice_velocity_plane(g, Box3D(0, g->nx - 1, g->ny - 1, g->ny - 1, 0, g->nz - 1));
  gpu_lattice_velocity_plane(g, Box3D(0, g->nx - 1, 0, g->ny - 1, 0, 0));
  gpu_lattice_velocity_plane(g, Box3D(0, g->nx - 1, 0, g->ny - 1, g->nz - 1, g->nz - 1));
}
void gpu_lattice_boundary_velocity(GpuLattice* gp, const Box3D& domain_, const double u[3]) {
  GpuLattice& g = *gp;
  const Box3D domain = clip(g, domain_);
  // the device keeps one wall velocity per face orientation: set it for every orientation present in the box
  bool seen[6] = {false, false, false, false, false, false};
  for (plint x = domain.x0; x <= domain.x1; x++) for (plint y = domain.y0; y <= domain.y1; y++) for (plint z = domain.z0; z <= domain.z1; z++) {
    const uint8_t f = g.flags[g.idx((int)x, (int)y, (int)z)];
    if (f >= HCG_VEL_XN && f <= HCG_VEL_ZP) seen[f - HCG_VEL_XN] = true;
    else if (f >= HCG_ZH_VEL_XN && f <= HCG_ZH_VEL_ZP) {                 // Zou-He velocity node: its own value
      auto it = g.bcn.find(g.idx((int)x, (int)y, (int)z));
      const double rho = it == g.bcn.end() ? 1.0 : it->second[3];
      g.bcn[g.idx((int)x, (int)y, (int)z)] = {u[0], u[1], u[2], rho};
      g.bcn_dirty = true;
    }
  }
  for (int o = 0; o < 6; o++) if (seen[o]) for (int k = 0; k < 3; k++) g.bc[o][k] = u[k];
  g.touchFlags();
}
void gpu_lattice_external_vector(GpuLattice* gp, const Box3D& domain_, const double vec[3]) {
  GpuLattice& g = *gp;
  const Box3D domain = clip(g, domain_);
  if (domain.nCells() != (plint)g.nx*g.ny*g.nz) {
    // a force on part of the lattice (cases/kolmogorovFlow/kolmogorovFlow.cpp:138-142): keep a per-node driving-force
    // field on the host, upload it when it changes; iterate() resets the node force to it.  Regions never set keep the
    // uniform value given before.
    const size_t N = (size_t)g.nx*g.ny*g.nz;
    if (g.bodyfield.empty()) { g.bodyfield.resize(3*N); for (int k = 0; k < 3; k++) std::fill(g.bodyfield.begin() + k*N, g.bodyfield.begin() + (k + 1)*N, g.body[k]); g.bodyfield_dirty = true; }
    for (plint x = domain.x0; x <= domain.x1; x++) for (plint y = domain.y0; y <= domain.y1; y++) for (plint z = domain.z0; z <= domain.z1; z++) {
      const size_t i = (size_t)g.idx((int)x, (int)y, (int)z);
      for (int k = 0; k < 3; k++) if (g.bodyfield[k*N + i] != vec[k]) { g.bodyfield[k*N + i] = vec[k]; g.bodyfield_dirty = true; }
    }
    g.body_pending = true;
    g.flush();
    return;
  }
  // after iterate() every node already carries the driving force again: re-applying the same value is free
  if (g.ctx && !g.body_pending && g.bodyfield.empty() && vec[0] == g.body[0] && vec[1] == g.body[1] && vec[2] == g.body[2]) return;
  for (int k = 0; k < 3; k++) g.body[k] = vec[k];
  g.bodyfield.clear(); g.bodyfield.shrink_to_fit(); g.bodyfield_dirty = false;
  g.body_pending = true;
  g.flush();
}
void gpu_lattice_define_flag(GpuLattice* gp, const Box3D& domain_, const plb::DomainFunctional3D* fun, int flag) {
  GpuLattice& g = *gp;
  const Box3D domain = clip(g, domain_);
  for (plint x = domain.x0; x <= domain.x1; x++) for (plint y = domain.y0; y <= domain.y1; y++) for (plint z = domain.z0; z <= domain.z1; z++)
    if (!fun || (*fun)(x, y, z)) g.flags[g.idx((int)x, (int)y, (int)z)] = (uint8_t)flag;
  g.touchFlags();
}
void gpu_lattice_equilibrium(GpuLattice* g, double rho, const double u[3]) {
  g->eq_rho = rho; for (int k = 0; k < 3; k++) g->eq_u[k] = u[k];
  g->eq_pending = true;
  g->flush();
  if (g->companion) gpu_lattice_equilibrium(g->companion, rho, u);
}
// Zou-He velocity (pressure = 0) / pressure (1) nodes, orientation 0..5 = outward normal -x +x -y +y -z +z.  Bounce-back nodes
// inside the box stay walls (Palabos would wrap the bounce-back dynamics; a wall node has nothing to impose).
void gpu_lattice_zouhe(GpuLattice* gp, const Box3D& domain_, int pressure, int orientation) {
  GpuLattice& g = *gp;
  const Box3D domain = clip(g, domain_);
  const uint8_t flag = (uint8_t)((pressure ? HCG_ZH_PRES_XN : HCG_ZH_VEL_XN) + orientation);
  for (plint x = domain.x0; x <= domain.x1; x++) for (plint y = domain.y0; y <= domain.y1; y++) for (plint z = domain.z0; z <= domain.z1; z++) {
    const int64_t i = g.idx((int)x, (int)y, (int)z);
    if (g.flags[i] == HCG_BOUNCEBACK) continue;
    g.flags[i] = flag;
    if (!g.bcn.count(i)) g.bcn[i] = {0.0, 0.0, 0.0, 1.0};
  }
  g.bcn_dirty = true;
  g.touchFlags();
}
void gpu_lattice_boundary_density(GpuLattice* gp, const Box3D& domain_, double rho) {
  GpuLattice& g = *gp;
  const Box3D domain = clip(g, domain_);
  for (plint x = domain.x0; x <= domain.x1; x++) for (plint y = domain.y0; y <= domain.y1; y++) for (plint z = domain.z0; z <= domain.z1; z++) {
    const int64_t i = g.idx((int)x, (int)y, (int)z);
    if (g.flags[i] < HCG_ZH_PRES_XN) continue;
    auto it = g.bcn.find(i);
    if (it == g.bcn.end()) g.bcn[i] = {0.0, 0.0, 0.0, rho}; else it->second[3] = rho;
    g.bcn_dirty = true;
  }
  if (g.ctx) g.flush();
}
int gpu_lattice_flag(const GpuLattice* g, long x, long y, long z) {
  if (x < 0 || y < 0 || z < 0 || x >= g->nx || y >= g->ny || z >= g->nz) return -1;
  return g->flags[g->idx((int)x, (int)y, (int)z)];
}
std::string gpu_lattice_info(const GpuLattice* g) {
  std::ostringstream o;
  const int R = plb::global::mpi().getSize();
  o << "Size of the multi-block:     " << g->nx << "-by-" << g->ny << "-by-" << g->nz << "\n"
    << "Number of atomic-blocks:     " << R << " (one x-slab per GPU)\n"
    << "Smallest atomic-block:       " << g->nx/R << "-by-" << g->ny << "-by-" << g->nz << "\n"
    << "Number of allocated cells:   " << (double)g->nx*g->ny*g->nz/1e6 << " million\n";
  return o.str();
}
}  // namespace hemo
