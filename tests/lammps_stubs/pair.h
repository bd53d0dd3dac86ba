#pragma once
#include "pointers.h"
namespace LAMMPS_NS {
class Pair : protected Pointers {
 public:
  double eng_vdwl, eng_coul, virial[6];
  int eflag_global, vflag_global, eflag_either, vflag_either, evflag;
  double **cutsq;
  double cut_coul;
  Pair(LAMMPS *l) : Pointers(l), eng_vdwl(0), eng_coul(0), virial{0, 0, 0, 0, 0, 0}, eflag_global(1), vflag_global(1),
      eflag_either(1), vflag_either(1), evflag(1), cutsq(nullptr), cut_coul(0) {}
  virtual void *extract(const char *what, int &dim) {
    dim = 0;
    return strcmp(what, "cut_coul") == 0 ? (void *)&cut_coul : nullptr;
  }
};
}  // namespace LAMMPS_NS
