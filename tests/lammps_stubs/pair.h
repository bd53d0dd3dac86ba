#pragma once
#include "pointers.h"
namespace LAMMPS_NS {
class Pair : protected Pointers {
 public:
  double eng_vdwl, eng_coul, virial[6];
  int eflag_global, vflag_global, eflag_either, vflag_either, evflag;
  double **cutsq;
  Pair(LAMMPS *l) : Pointers(l) {}
  virtual void *extract(const char *, int &) { return nullptr; }
  void ev_tally(int, int, int, int, double, double, double, double, double, double);
};
}  // namespace LAMMPS_NS
