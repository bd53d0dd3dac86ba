#pragma once
#include "pointers.h"
namespace LAMMPS_NS {
class KSpace : protected Pointers {
 public:
  double energy, virial[6];
  double g_ewald, accuracy, accuracy_relative, slab_volfactor, scale, qqrd2e;
  int slabflag, order, compute_flag, differentiation_flag;
  int nx_pppm, ny_pppm, nz_pppm;
  KSpace(LAMMPS *l) : Pointers(l) {}
  virtual void init() {}
  virtual void setup() {}
  virtual void compute(int, int) {}
};
}  // namespace LAMMPS_NS
