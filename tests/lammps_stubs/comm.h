#pragma once
#include "pointers.h"
namespace LAMMPS_NS {
class Comm : protected Pointers {
 public:
  int me, nprocs;
  Comm(LAMMPS *l) : Pointers(l), me(0), nprocs(1) {}
};
}  // namespace LAMMPS_NS
