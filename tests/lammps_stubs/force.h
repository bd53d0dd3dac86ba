#pragma once
#include "pointers.h"
namespace LAMMPS_NS {
class Pair; class KSpace;
class Force : protected Pointers {
 public:
  double qqrd2e, qqr2e, qe2f, dielectric, boltz;
  int newton, newton_pair, newton_bond;
  Pair *pair;
  KSpace *kspace;
  Force(LAMMPS *l) : Pointers(l) {}
  Pair *pair_match(const std::string &, int, int = 0);
};
}  // namespace LAMMPS_NS
