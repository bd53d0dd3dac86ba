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
  std::string pair_style;
  Force(LAMMPS *l) : Pointers(l), qqrd2e(332.06371), qqr2e(332.06371), qe2f(23.060549), dielectric(1.0), boltz(0.0019872067),
      newton(1), newton_pair(1), newton_bond(1), pair(nullptr), kspace(nullptr) {}  // `units real`
  Pair *pair_match(const std::string &word, int exact, int = 0) {
    if (!pair) return nullptr;
    if (exact ? pair_style == word : pair_style.find(word) != std::string::npos) return pair;
    return nullptr;
  }
};
}  // namespace LAMMPS_NS
