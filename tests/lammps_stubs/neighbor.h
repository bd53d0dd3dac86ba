#pragma once
#include "pointers.h"
namespace LAMMPS_NS {
class Neighbor : protected Pointers {
 public:
  Neighbor(LAMMPS *l) : Pointers(l) {}
  int request(void *, int = 0);
};
}  // namespace LAMMPS_NS
