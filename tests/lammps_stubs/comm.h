#pragma once
#include "pointers.h"
namespace LAMMPS_NS {
class Comm : protected Pointers {
 public:
  int me, nprocs;
  Comm(LAMMPS *l) : Pointers(l) {}
};
}  // namespace LAMMPS_NS
