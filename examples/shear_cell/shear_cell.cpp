// One cell type in plane Couette flow, written against the HemoCell API exactly as a case file of the
// reference would be (include/hemocell.h keeps that surface; compare examples/oneCellShear in the
// reference, which itself compiles unmodified against this header: see examples/Makefile `refcases`).
// Every `tmeas` steps the cell's observables go to shear.log:
//   iter  dx dy dz [um]  volume%  area%  D_max [um]  deformation index [%]
#include <iomanip>
#include "hemocell.h"
#include "rbcHighOrderModel.h"
#include "pltSimpleModel.h"
#include "cellInfo.h"
#include "fluidInfo.h"
#include "helper/hemocellInit.hh"
#include "palabos3D.h"
#include "palabos3D.hh"

using namespace hemo;

int main(int argc, char* argv[]) {
  if (argc < 2) { cout << "Usage: " << argv[0] << " <configuration.xml>" << endl; return -1; }
  HemoCell hemocell(argv[1], argc, argv);
  Config* cfg = hemocell.cfg;

  // box: height H um between the moving plates, 2H x 2H in the periodic directions
  const T height_um = (*cfg)["domain"]["height"].read<T>();
  const plint nz = (plint)(height_um*(1e-6/(*cfg)["domain"]["dx"].read<T>()));
  const plint nx = 2*nz, ny = 2*nz;
  param::lbm_shear_parameters(*cfg, ny);
  param::printParameters();

  hemocell.lattice = new MultiBlockLattice3D<T, DESCRIPTOR>(
      defaultMultiBlockPolicy3D().getMultiBlockManagement(nx, ny, nz, 2),
      defaultMultiBlockPolicy3D().getBlockCommunicator(), defaultMultiBlockPolicy3D().getCombinedStatistics(),
      defaultMultiBlockPolicy3D().getMultiCellAccess<T, DESCRIPTOR>(),
      new GuoExternalForceBGKdynamics<T, DESCRIPTOR>(1.0/param::tau));
  OnLatticeBoundaryCondition3D<T, DESCRIPTOR>* bc = createLocalBoundaryCondition3D<T, DESCRIPTOR>();
  hemocell.lattice->toggleInternalStatistics(false);
  iniLatticeSquareCouette(*hemocell.lattice, nx, ny, nz, *bc, param::shearrate_lbm);
  hlog << getMultiBlockInfo(*hemocell.lattice) << endl;

  hemocell.initializeCellfield();
  const std::string cellName = (*cfg)["ibm"]["cellType"].read<std::string>();
  if (cellName == "PLT") hemocell.addCellType<PltSimpleModel>("PLT", ELLIPSOID_FROM_SPHERE);
  else hemocell.addCellType<RbcHighOrderModel>(cellName, RBC_FROM_SPHERE);
  hemocell.setMaterialTimeScaleSeparation(cellName, (*cfg)["ibm"]["stepMaterialEvery"].read<int>());
  hemocell.setParticleVelocityUpdateTimeScaleSeparation((*cfg)["ibm"]["stepParticleEvery"].read<int>());
  hemocell.setOutputs(cellName, {OUTPUT_POSITION, OUTPUT_TRIANGLES, OUTPUT_FORCE});
  hemocell.setFluidOutputs({OUTPUT_VELOCITY});
  hemocell.loadParticles();

  for (plint i = 0; i < (*cfg)["parameters"]["warmup"].read<plint>(); i++) hemocell.lattice->collideAndStream();

  const unsigned int tmax = (*cfg)["sim"]["tmax"].read<unsigned int>();
  const unsigned int tmeas = (*cfg)["sim"]["tmeas"].read<unsigned int>();
  const T lu2um = param::dx/1e-6;
  CellInformationFunctionals::calculateCellVolume(&hemocell);
  CellInformationFunctionals::calculateCellArea(&hemocell);
  const int cid = CellInformationFunctionals::info_per_cell.begin()->first;
  const T volume_eq = CellInformationFunctionals::info_per_cell[cid].volume, area_eq = CellInformationFunctionals::info_per_cell[cid].area;
  const T D0 = 2.0*(*cfg)["ibm"]["radius"].read<T>()*1e6;

  plb_ofstream log("shear.log");
  log << std::setprecision(15);
  while (hemocell.iter < tmax) {
    hemocell.iterate();
    if (hemocell.iter % tmeas) continue;
    hemocell.writeOutput();
    CellInformationFunctionals::calculateCellStretch(&hemocell);
    const CellInformation& ci = CellInformationFunctionals::info_per_cell[cid];
    const T dmax = ci.stretch*lu2um, r2 = (dmax/D0)*(dmax/D0);
    log << hemocell.iter << " " << (ci.bbox[1] - ci.bbox[0])*lu2um << " " << (ci.bbox[3] - ci.bbox[2])*lu2um << " " << (ci.bbox[5] - ci.bbox[4])*lu2um
        << " " << ci.volume/volume_eq*100.0 << " " << ci.area/area_eq*100.0 << " " << dmax << " " << (r2 - 1.0)/(r2 + 1.0)*100.0 << endl;
    const FluidStatistics fs = FluidInfo::calculateVelocityStatistics(&hemocell);
    hlog << "(shear_cell) iter " << hemocell.iter << ": cells " << CellInformationFunctionals::getTotalNumberOfCells(&hemocell)
         << ", D_max " << dmax << " um, mean |u| " << fs.avg*param::dx/param::dt << " m/s" << endl;
  }
  log.close();
  hemo::global.statistics.printStatistics();
  hemo::global.statistics.outputStatistics();
  delete bc;
  return 0;
}
