#pragma once
#include "pointers.h"
namespace LAMMPS_NS {
class Update : protected Pointers {
 public:
  bigint ntimestep, laststep;
  int eflag_global, vflag_global;
  Update(LAMMPS *l) : Pointers(l), ntimestep(0), laststep(0), eflag_global(0), vflag_global(0) {}
};
}  // namespace LAMMPS_NS
