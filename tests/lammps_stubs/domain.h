#pragma once
#include "pointers.h"
namespace LAMMPS_NS {
class Domain : protected Pointers {
 public:
  int xperiodic, yperiodic, zperiodic;
  int periodicity[3];
  double xprd, yprd, zprd;
  double boxlo[3], boxhi[3], prd[3];
  Domain(LAMMPS *l) : Pointers(l) {}
};
}  // namespace LAMMPS_NS
