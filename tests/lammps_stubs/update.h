#pragma once
#include "pointers.h"
namespace LAMMPS_NS {
class Update : protected Pointers {
 public:
  bigint ntimestep;
  int eflag_global, vflag_global;
  Update(LAMMPS *l) : Pointers(l) {}
};
}  // namespace LAMMPS_NS
