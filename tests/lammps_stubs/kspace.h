#pragma once
#include "pointers.h"
#define KSpaceStyle(key, Class)
namespace LAMMPS_NS {
class KSpace : protected Pointers {
 public:
  double energy, virial[6];
  double g_ewald, accuracy, accuracy_relative, slab_volfactor, scale, qqrd2e;
  int slabflag, order, compute_flag, differentiation_flag;
  int nx_pppm, ny_pppm, nz_pppm;
  KSpace(LAMMPS *l) : Pointers(l), energy(0), virial{0, 0, 0, 0, 0, 0}, g_ewald(0), accuracy(0), accuracy_relative(0),
      slab_volfactor(1), scale(1), qqrd2e(0), slabflag(0), order(5), compute_flag(1), differentiation_flag(0),
      nx_pppm(0), ny_pppm(0), nz_pppm(0) {}
  virtual void init() {}
  virtual void setup() {}
  virtual void compute(int, int) {}
};
}  // namespace LAMMPS_NS
