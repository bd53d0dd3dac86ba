#pragma once
#include "pointers.h"
namespace LAMMPS_NS {
class Group : protected Pointers {
 public:
  int ngroup;
  int *bitmask;
  Group(LAMMPS *l) : Pointers(l) {}
  int find(const std::string &);
};
}  // namespace LAMMPS_NS
