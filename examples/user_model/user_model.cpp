// A user-defined constitutive model plugged into HemoCell through the reference's CellMechanics interface
// (mechanics/cellMechanics.h:37-79): the class below overrides ParticleMechanics(map<cellId, vector<HemoCellParticle*>>, ...) and
// has no device kernel (deviceModel() is not overridden), so the facade evaluates it on host copies of the particles at the
// material cadence - the plugin path of the north star, as opposed to the built-in RBC / PLT models, which run as k_mechanics.
// The model is a worm-like-chain link force on every mesh edge (the link term a model like RbcHighOrderModel also contains),
// written purely against the public API: cellConstants, the k_* helpers, sv.position and the force_* pointers.
// Output: every `tmeas` steps one line per vertex "iter vertex x y z fx fy fz" (lattice units) in user_model.log.
#include <iomanip>
#include "hemocell.h"
#include "cellMechanics.h"
#include "helper/hemocellInit.hh"
#include "palabos3D.h"
#include "palabos3D.hh"

using namespace hemo;

class LinkOnlyModel : public CellMechanics {
 public:
  HemoCellField& cellField;
  const T k_link;
  LinkOnlyModel(Config& modelCfg_, HemoCellField& cellField_) : CellMechanics(cellField_, modelCfg_), cellField(cellField_), k_link(calculate_kLink(modelCfg_)) {}
  void ParticleMechanics(std::map<int, std::vector<HemoCellParticle*>>& particles_per_cell, const std::map<int, bool>& lpc, pluint ctype) override {
    for (const auto& pair : particles_per_cell) {
      const int& cid = pair.first;
      const std::vector<HemoCellParticle*>& cell = pair.second;
      if (cell[0]->sv.celltype != ctype) continue;
      if (lpc.find(cid) == lpc.end()) continue;
      int edge_n = 0;
      for (const hemo::Array<plint, 2>& edge : cellConstants.edge_list) {
        const hemo::Array<T, 3>& p0 = cell[edge[0]]->sv.position;
        const hemo::Array<T, 3>& p1 = cell[edge[1]]->sv.position;
        const hemo::Array<T, 3> d = p1 - p0;
        const T len = norm(d);
        const T frac = (len - cellConstants.edge_length_eq_list[edge_n])/cellConstants.edge_length_eq_list[edge_n];
        const T scalar = k_link*(frac + frac/std::fabs(9.0 - frac*frac));
        const hemo::Array<T, 3> force = (d/len)*scalar;
        *cell[edge[0]]->force_link += force;
        *cell[edge[1]]->force_link -= force;
        edge_n++;
      }
    }
  }
  void statistics() override { hlog << "(Cell-mechanics model) user link-only model for " << cellField.name << ", k_link " << k_link << std::endl; }
};

int main(int argc, char* argv[]) {
  if (argc < 2) { cout << "Usage: " << argv[0] << " <configuration.xml>" << endl; return -1; }
  HemoCell hemocell(argv[1], argc, argv);
  Config* cfg = hemocell.cfg;
  const T height_um = (*cfg)["domain"]["height"].read<T>();
  const plint nz = (plint)(height_um*(1e-6/(*cfg)["domain"]["dx"].read<T>()));
  const plint nx = 2*nz, ny = 2*nz;
  param::lbm_shear_parameters(*cfg, ny);
  hemocell.lattice = new MultiBlockLattice3D<T, DESCRIPTOR>(
      defaultMultiBlockPolicy3D().getMultiBlockManagement(nx, ny, nz, 2),
      defaultMultiBlockPolicy3D().getBlockCommunicator(), defaultMultiBlockPolicy3D().getCombinedStatistics(),
      defaultMultiBlockPolicy3D().getMultiCellAccess<T, DESCRIPTOR>(),
      new GuoExternalForceBGKdynamics<T, DESCRIPTOR>(1.0/param::tau));
  OnLatticeBoundaryCondition3D<T, DESCRIPTOR>* bc = createLocalBoundaryCondition3D<T, DESCRIPTOR>();
  hemocell.lattice->toggleInternalStatistics(false);
  iniLatticeSquareCouette(*hemocell.lattice, nx, ny, nz, *bc, param::shearrate_lbm);
  hemocell.initializeCellfield();
  hemocell.addCellType<LinkOnlyModel>("RBC", RBC_FROM_SPHERE);
  hemocell.setMaterialTimeScaleSeparation("RBC", (*cfg)["ibm"]["stepMaterialEvery"].read<int>());
  hemocell.setParticleVelocityUpdateTimeScaleSeparation((*cfg)["ibm"]["stepParticleEvery"].read<int>());
  hemocell.setOutputs("RBC", {OUTPUT_POSITION, OUTPUT_FORCE});
  hemocell.setFluidOutputs({OUTPUT_VELOCITY});
  hemocell.loadParticles();

  const unsigned int tmax = (*cfg)["sim"]["tmax"].read<unsigned int>();
  const unsigned int tmeas = (*cfg)["sim"]["tmeas"].read<unsigned int>();
  plb_ofstream log("user_model.log");
  log << std::setprecision(17);
  while (hemocell.iter < tmax) {
    hemocell.iterate();
    if (hemocell.iter % tmeas) continue;
    std::vector<HemoCellParticle> particles;
    hemocell.cellfields->getParticles(particles);
    for (const HemoCellParticle& p : particles)
      log << hemocell.iter << " " << p.sv.vertexId << " " << p.sv.position[0] << " " << p.sv.position[1] << " " << p.sv.position[2]
          << " " << p.sv.force[0] << " " << p.sv.force[1] << " " << p.sv.force[2] << endl;
  }
  log.close();
  delete bc;
  return 0;
}
