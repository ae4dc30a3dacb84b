// CPU-side checks of the C++ API surface (include/hemocell.h): config parsing, unit conversion, cell-type
// set-up.  No device call is made (the context is created lazily).  Prints "ok <name>" per check; exit code
// = number of failures.  usage: test_host_facade <scratch dir>
#include <cmath>
#include <cstdio>
#include <unistd.h>
#include "hemocell.h"
#include "rbcHighOrderModel.h"
#include "pltSimpleModel.h"

using namespace hemo;
static int failures = 0;
#define CHECK(name, cond) do { if (cond) std::printf("ok %s\n", name); else { std::printf("FAIL %s\n", name); failures++; } } while (0)
static bool close_rel(double a, double b, double tol) { return std::fabs(a - b) <= tol*std::max(std::fabs(a), std::fabs(b)); }

static void write(const std::string& path, const std::string& s) { std::ofstream f(path); f << s; }

int main(int argc, char** argv) {
  if (argc < 2) return 100;
  if (chdir(argv[1]) != 0) return 101;
  write("config.xml",
        "<?xml version=\"1.0\" ?>\n<!-- a comment -->\n<hemocell>\n<parameters><warmup> 3 </warmup><outputDirectory>out</outputDirectory></parameters>\n"
        "<ibm><radius>3.91e-6</radius><stepMaterialEvery> 20 </stepMaterialEvery></ibm>\n"
        "<domain attr=\"x &amp; y\"><shearrate> 111.0 </shearrate><rhoP>1025</rhoP><nuP>1.1e-6</nuP><dx>0.5e-6</dx><dt>0.5e-7</dt>\n"
        "<particleEnvelope>20</particleEnvelope><kBT>4.100531391e-21</kBT><Re>0.5</Re><empty/></domain>\n"
        "<preInlet><parameters><lengthN> 30 </lengthN><Re> 0.5 </Re></parameters></preInlet>\n"
        "<sim><tmax>10</tmax></sim></hemocell>\n");
  write("RBC.xml",
        "<?xml version=\"1.0\" ?><hemocell><MaterialModel><name>RBC</name><eta_m>0.0</eta_m><kBend>80.0</kBend><kVolume>20.0</kVolume>"
        "<kArea>5.0</kArea><kLink>15.0</kLink><minNumTriangles>600</minNumTriangles><radius>3.91e-6</radius><Volume>90</Volume></MaterialModel></hemocell>");
  write("PLT.xml",
        "<?xml version=\"1.0\" ?><hemocell><MaterialModel><name>PLT</name><eta_m>0.0</eta_m><kBend>250.0</kBend><kVolume>100.0</kVolume>"
        "<kArea>8.0</kArea><kLink>25.0</kLink><minNumTriangles>66</minNumTriangles><radius>1.25e-6</radius><aspectRatio>0.434782608696</aspectRatio>"
        "<Volume>11</Volume><InnerEdges><Edge> 0 1 </Edge><Edge>2 3</Edge></InnerEdges></MaterialModel></hemocell>");
  char prog[] = "test"; char cfgname[] = "config.xml"; char* av[] = {prog, cfgname};
  HemoCell hemocell(cfgname, 2, av);
  Config* cfg = hemocell.cfg;
  CHECK("config.read<int>", (*cfg)["parameters"]["warmup"].read<int>() == 3);
  CHECK("config.read<double>", (*cfg)["domain"]["shearrate"].read<T>() == 111.0);
  CHECK("config.read<string>", (*cfg)["parameters"]["outputDirectory"].read<std::string>() == "out");
  CHECK("config.not_checkpointed", !cfg->checkpointed);
  bool threw = false;
  try { (*cfg)["domain"]["nope"]; } catch (std::invalid_argument&) { threw = true; }
  CHECK("config.missing_key_throws_invalid_argument", threw);
  CHECK("output_directory_created", access(global::directories().getOutputDir().c_str(), F_OK) == 0);

  param::lbm_shear_parameters(*cfg, 40);
  CHECK("param.tau", close_rel(param::tau, 1.16, 1e-14));
  CHECK("param.nu_lbm", close_rel(param::nu_lbm, 0.22, 1e-14));
  CHECK("param.df", close_rel(param::df, 1025*std::pow(0.5e-6, 4)/std::pow(0.5e-7, 2), 1e-14));
  CHECK("param.f_limit", close_rel(param::f_limit, 50.0/1e12/param::df, 1e-14));
  CHECK("param.shearrate_lbm", close_rel(param::shearrate_lbm, 111.0*0.5e-7, 1e-14));
  param::lbm_pipe_parameters(*cfg, 256);
  CHECK("param.u_lbm_max", close_rel(param::u_lbm_max, 0.5*param::nu_lbm/512.0, 1e-14));

  hemocell.lattice = new MultiBlockLattice3D<T, DESCRIPTOR>(defaultMultiBlockPolicy3D().getMultiBlockManagement(40, 40, 20, 2),
      defaultMultiBlockPolicy3D().getBlockCommunicator(), defaultMultiBlockPolicy3D().getCombinedStatistics(),
      defaultMultiBlockPolicy3D().getMultiCellAccess<T, DESCRIPTOR>(), new GuoExternalForceBGKdynamics<T, DESCRIPTOR>(1.0/param::tau));
  CHECK("lattice.bbox", hemocell.lattice->getBoundingBox().getNx() == 40 && hemocell.lattice->getNz() == 20);
  hemocell.lattice->periodicity().toggle(0, true);
  CHECK("lattice.periodicity", hemocell.lattice->periodicity().get(0) && !hemocell.lattice->periodicity().get(2));
  defineDynamics(*hemocell.lattice, Box3D(0, 39, 0, 0, 0, 19), new BounceBack<T, DESCRIPTOR>(1.));
  hemocell.latticeEquilibrium(1., plb::Array<T, 3>(0., 0., 0.));
  hemocell.lattice->initialize();                      // still no device needed

  hemocell.initializeCellfield();
  hemocell.addCellType<RbcHighOrderModel>("RBC", RBC_FROM_SPHERE);
  hemocell.addCellType<PltSimpleModel>("PLT", ELLIPSOID_FROM_SPHERE);
  hemocell.setMaterialTimeScaleSeparation("RBC", 20);
  HemoCellField* rbc = (*hemocell.cellfields)["RBC"];
  HemoCellField* plt = (*hemocell.cellfields)[1];
  CHECK("rbc.mesh", rbc->numVertex == 642 && rbc->triangle_list.size() == 1280);
  CHECK("plt.mesh", plt->numVertex == 66 && plt->triangle_list.size() == 128);
  CHECK("rbc.timescale", rbc->timescale == 20);
  CHECK("rbc.edges", rbc->mechanics->cellConstants.edge_list.size() == 1920);
  const RbcHighOrderModel* m = dynamic_cast<RbcHighOrderModel*>(rbc->mechanics);
  const double kBT_lbm = 4.100531391e-21/(param::df*param::dx);
  CHECK("rbc.k_link", m && close_rel(m->k_link, 15.0*kBT_lbm/(7.5e-9/param::dx), 1e-13));
  CHECK("rbc.k_bend", m && close_rel(m->k_bend, 80.0*kBT_lbm/(5e-7/param::dx), 1e-13));
  CHECK("rbc.k_volume", m && close_rel(m->k_volume, 20.0*kBT_lbm/(5e-7/param::dx), 1e-13));
  const PltSimpleModel* pm = dynamic_cast<PltSimpleModel*>(plt->mechanics);
  CHECK("plt.k_area_face_scaling", pm && close_rel(pm->k_area, 8.0*(1280.0/128.0)*kBT_lbm/(5e-7/param::dx), 1e-13));
  CHECK("rbc.volume_eq_um3", std::fabs(rbc->mechanics->cellConstants.volume_eq*std::pow(0.5, 3) - 81.1) < 0.2);
  CHECK("cellfields.size", hemocell.cellfields->size() == 2);
  // meshmetric / original bounding box (core/hemoCellField.h), used by the stretchCell case files
  CHECK("rbc.meshmetric.volume", close_rel(rbc->meshmetric->getVolume(), rbc->mechanics->cellConstants.volume_eq, 1e-12));
  CHECK("rbc.meshmetric.surface_um2", std::fabs(rbc->meshmetric->getSurface()*0.25 - 129.2) < 0.3);
  { auto bb = rbc->getOriginalBoundingBox(); CHECK("rbc.original_bbox", std::fabs((bb[1] - bb[0])*0.5 - 7.82) < 0.01 && std::fabs((bb[3] - bb[2])*0.5 - 2.294) < 0.01 && close_rel(bb[5] - bb[4], bb[1] - bb[0], 1e-12)); }   // disc axis along y
  // the reference stores the minimum wall distance in an unsigned int (core/hemoCellField.h:64): 0.5 um -> 0
  hemocell.setInitialMinimumDistanceFromSolid("RBC", 0.5);
  CHECK("quirk.min_distance_truncates", rbc->minimumDistanceFromSolid == 0);
  hemocell.setInitialMinimumDistanceFromSolid("RBC", 1);
  CHECK("min_distance_1um", rbc->minimumDistanceFromSolid == 1);

  // helper/voxelizeDomain.h on an ASCII STL written here: a 20 x 10 x 10 box, refDirN 50 along y -> dx = 0.2
  {
    std::ofstream f("box.stl");
    f << "solid box\n";
    const double lo[3] = {-10, -10, -5}, hi[3] = {10, 0, 5};
    auto P = [&](int ix, int iy, int iz) { std::ostringstream o; o << "vertex " << (ix ? hi[0] : lo[0]) << " " << (iy ? hi[1] : lo[1]) << " " << (iz ? hi[2] : lo[2]) << "\n"; return o.str(); };
    const int q[6][4][3] = {{{0,0,0},{0,0,1},{0,1,1},{0,1,0}}, {{1,0,0},{1,1,0},{1,1,1},{1,0,1}}, {{0,0,0},{1,0,0},{1,0,1},{0,0,1}},
                            {{0,1,0},{0,1,1},{1,1,1},{1,1,0}}, {{0,0,0},{0,1,0},{1,1,0},{1,0,0}}, {{0,0,1},{1,0,1},{1,1,1},{0,1,1}}};
    for (auto& fc : q) for (int t = 0; t < 2; t++) {
      f << "facet normal 0 0 0\nouter loop\n" << P(fc[0][0], fc[0][1], fc[0][2]) << P(fc[t+1][0], fc[t+1][1], fc[t+1][2]) << P(fc[t+2][0], fc[t+2][1], fc[t+2][2]) << "endloop\nendfacet\n";
    }
    f << "endsolid box\n";
  }
  plb::VoxelizedDomain3D<T>* vd = nullptr; plb::MultiScalarField3D<int>* fm = nullptr;
  getFlagMatrixFromSTL("box.stl", 2, 50, 1, vd, fm, -1, 25);
  CHECK("voxelizer.size", fm && fm->getNx() == 103 && fm->getNy() == 53 && fm->getNz() == 53);
  CHECK("voxelizer.fluid_inside", fm && fm->get(50, 26, 26) == 1 && fm->get(50, 1, 1) == 1 && fm->get(50, 0, 26) == 0 && fm->get(50, 26, 52) == 0);
  CHECK("voxelizer.open_x_ends", fm && fm->get(0, 26, 26) == 1 && fm->get(102, 26, 26) == 1);
  CHECK("voxelizer.management", vd && vd->getMultiBlockManagement().getBoundingBox().getNx() == 103);
  param::lbm_pipe_parameters(*cfg, fm);
  CHECK("param.pipe_radius_from_fluid_area", close_rel(param::pipe_radius, std::sqrt(51.0*51.0/PI), 1e-12));
  // helper/preInlet.h on that flag matrix (host side only: the pre-inlet is a second lattice of this process, no device yet)
  {
    hemocell.preInlet = new hemo::PreInlet(&hemocell, fm);
    CHECK("preinlet.length", hemocell.preInlet->preinlet_length == 30 && !hemocell.partOfpreInlet);
    Box3D slice = fm->getBoundingBox(); slice.x0 = slice.x1 = 2;
    hemocell.preInlet->preInletFromSlice(Direction::Xneg, slice);
    const Box3D loc = hemocell.preInlet->location;
    // fluid cross-section 1..51 in y and z, one solid node around it, lengthN planes towards -x (helper/preInlet.cpp:511-519)
    CHECK("preinlet.location", loc.x0 == 1 - 30 && loc.x1 == 3 && loc.y0 == 0 && loc.y1 == 52 && loc.z0 == 0 && loc.z1 == 52);
    CHECK("preinlet.inflow_length", hemocell.preInlet->inflow_length == 20);
    hemocell.cellfields->lattice = nullptr;
    hemocell.initializeLattice(vd->getMultiBlockManagement());
    long n[3]; hemo::gpu_lattice_size(hemocell.preInlet->pre, n);
    CHECK("preinlet.lattice_size", n[0] == 33 && n[1] == 53 && n[2] == 53);
    hemocell.lattice->periodicity().toggleAll(false);
    hemocell.preInlet->initializePreInlet();
    const Box3D in = hemocell.preInlet->fluidInlet;
    CHECK("preinlet.fluidInlet_plane", in.x0 == 3 && in.x1 == 3 && in.y0 == 0 && in.y1 == 52);
    boundaryFromFlagMatrix(hemocell.lattice, fm, hemocell.partOfpreInlet);
    hemocell.preInlet->createBoundary();
    int zh = 0, bb = 0;
    for (int y = 0; y < 53; y++) for (int z = 0; z < 53; z++) {
      const int f = hemo::gpu_lattice_flag(hemocell.lattice->gpu(), 3, y, z);
      zh += f == HCG_ZH_VEL_XN; bb += f == HCG_BOUNCEBACK;
    }
    CHECK("preinlet.inlet_nodes", zh == 51*51 && bb == 53*53 - 51*51);
    CHECK("preinlet.main_stays_fluid_elsewhere", hemo::gpu_lattice_flag(hemocell.lattice->gpu(), 4, 26, 26) == HCG_FLUID);
    int pre_bb = 0;
    for (int x = 0; x < 33; x++) for (int y = 0; y < 53; y++) for (int z = 0; z < 53; z++) pre_bb += hemo::gpu_lattice_flag(hemocell.preInlet->pre, x, y, z) == HCG_BOUNCEBACK;
    CHECK("preinlet.extruded_cross_section", pre_bb == 33*(53*53 - 51*51));
    CHECK("preinlet.periodic_along_the_flow", hemo::gpu_lattice_get_periodic(hemocell.preInlet->pre, 0) && !hemo::gpu_lattice_get_periodic(hemocell.preInlet->pre, 1)
                                              && !hemo::gpu_lattice_get_periodic(hemocell.lattice->gpu(), 0));
    hemocell.preInlet->calculateDrivingForce();
    const double r = std::sqrt(51.0*51.0/PI), u = 0.5*param::nu_lbm/(2*r);
    CHECK("preinlet.driving_force", close_rel(hemocell.preInlet->drivingForce, 8*param::nu_lbm*(u*0.5)/r/r, 1e-13));
    // Zou-He outlet as pipeflow_with_preinlet.cpp:125-133 writes it (three planes, density 1)
    Box3D lb = hemocell.lattice->getBoundingBox();
    OnLatticeBoundaryCondition3D<T, DESCRIPTOR>* boundary = new BoundaryConditionInstantiator3D<T, DESCRIPTOR, WrappedZouHeBoundaryManager3D<T, DESCRIPTOR>>();
    boundary->addPressureBoundary0P(Box3D(lb.x1 - 2, lb.x1, lb.y0, lb.y1, lb.z0, lb.z1), *hemocell.lattice, boundary::density);
    setBoundaryDensity(*hemocell.lattice, Box3D(lb.x1 - 2, lb.x1, lb.y0, lb.y1, lb.z0, lb.z1), 1.0);
    CHECK("zouhe.pressure_outlet", hemo::gpu_lattice_flag(hemocell.lattice->gpu(), lb.x1 - 1, 26, 26) == HCG_ZH_PRES_XP
                                   && hemo::gpu_lattice_flag(hemocell.lattice->gpu(), lb.x1, 0, 0) == HCG_BOUNCEBACK);
    delete boundary;
    // pulsatile driving force helpers (helper/preInlet.cpp:860-890)
    std::vector<double> xs = {0.0, 1.0, 2.0}, ys = {1.0, 3.0, 2.0};
    CHECK("preinlet.interpolate", close_rel(hemocell.preInlet->interpolate(xs, ys, 0.5, false), 2.0, 1e-15) && close_rel(hemocell.preInlet->interpolate(xs, ys, 1.5, false), 2.5, 1e-15)
                                  && close_rel(hemocell.preInlet->interpolate(xs, ys, 5.0, false), 2.0, 1e-15) && close_rel(hemocell.preInlet->average(ys), 2.0, 1e-15));
  }
  // the other directions on a hand-made flag matrix: a 20 x 30 x 24 block, fluid inside a one-node solid shell that is open at both z ends
  {
    delete hemocell.preInlet; hemocell.preInlet = nullptr;
    plb::MultiScalarField3D<int> box(20, 30, 24, 0);
    for (int x = 1; x < 19; x++) for (int y = 1; y < 29; y++) for (int z = 0; z < 24; z++) box.get(x, y, z) = 1;
    plb::VoxelizedDomain3D<T> vbox(20, 30, 24, 2);
    hemocell.preInlet = new hemo::PreInlet(&hemocell, &box);
    hemocell.preInlet->autoPreinletFromBoundary(Direction::Zpos);            // slice = second plane from the +z face
    const Box3D loc = hemocell.preInlet->location;
    CHECK("preinlet.zpos.location", loc.x0 == 0 && loc.x1 == 19 && loc.y0 == 0 && loc.y1 == 29 && loc.z0 == 21 && loc.z1 == 23 + 30);
    hemocell.initializeLattice(vbox.getMultiBlockManagement());
    long n[3]; hemo::gpu_lattice_size(hemocell.preInlet->pre, n);
    CHECK("preinlet.zpos.lattice_size", n[0] == 20 && n[1] == 30 && n[2] == 33);
    hemocell.preInlet->initializePreInlet();
    const Box3D in = hemocell.preInlet->fluidInlet;
    CHECK("preinlet.zpos.inlet_plane", in.z0 == 21 && in.z1 == 21 && in.x0 == 0 && in.x1 == 19);
    boundaryFromFlagMatrix(hemocell.lattice, &box, false);
    hemocell.preInlet->createBoundary();
    int zh = 0;
    for (int x = 0; x < 20; x++) for (int y = 0; y < 30; y++) zh += hemo::gpu_lattice_flag(hemocell.lattice->gpu(), x, y, 21) == HCG_ZH_VEL_ZP;
    CHECK("preinlet.zpos.inlet_nodes_outward_plus_z", zh == 18*28);
    CHECK("preinlet.zpos.periodic_z_only", hemo::gpu_lattice_get_periodic(hemocell.preInlet->pre, 2) && !hemo::gpu_lattice_get_periodic(hemocell.preInlet->pre, 0));
    CHECK("preinlet.zpos.walls_extruded", hemo::gpu_lattice_flag(hemocell.preInlet->pre, 0, 5, 30) == HCG_BOUNCEBACK && hemo::gpu_lattice_flag(hemocell.preInlet->pre, 10, 15, 30) == HCG_FLUID);
    hemocell.preInlet->calculateDrivingForce();
    const double r = std::sqrt(18.0*28.0/PI), u = 0.5*param::nu_lbm/(2*r);
    CHECK("preinlet.zpos.driving_force", close_rel(hemocell.preInlet->drivingForce, 8*param::nu_lbm*(u*0.5)/r/r, 1e-13));
    // and a pre-inlet on the negative y side of the same block (opened at the y ends instead)
    delete hemocell.preInlet; hemocell.preInlet = nullptr;
    for (int x = 1; x < 19; x++) for (int z = 0; z < 24; z++) { box.get(x, 0, z) = z > 0 && z < 23; box.get(x, 29, z) = z > 0 && z < 23; }
    for (int x = 0; x < 20; x++) for (int y = 0; y < 30; y++) { box.get(x, y, 0) = 0; box.get(x, y, 23) = 0; }
    hemocell.preInlet = new hemo::PreInlet(&hemocell, &box);
    Box3D sl = box.getBoundingBox(); sl.y0 = sl.y1 = 1;
    hemocell.preInlet->preInletFromSlice(Direction::Yneg, sl);
    const Box3D l2 = hemocell.preInlet->location;
    CHECK("preinlet.yneg.location", l2.y0 == 0 - 30 && l2.y1 == 2 && l2.x0 == 0 && l2.x1 == 19 && l2.z0 == 0 && l2.z1 == 23);
    hemocell.initializeLattice(vbox.getMultiBlockManagement());
    hemocell.preInlet->initializePreInlet();
    CHECK("preinlet.yneg.inlet_plane", hemocell.preInlet->fluidInlet.y0 == 2 && hemocell.preInlet->fluidInlet.y1 == 2);
    int zy = 0;
    for (int x = 0; x < 20; x++) for (int z = 0; z < 24; z++) zy += hemo::gpu_lattice_flag(hemocell.lattice->gpu(), x, 2, z) == HCG_ZH_VEL_YN;
    CHECK("preinlet.yneg.inlet_nodes_outward_minus_y", zy == 18*22);
    delete hemocell.preInlet; hemocell.preInlet = nullptr;
  }
  delete vd; delete fm;
  std::printf("%d failures\n", failures);
  return failures;
}
