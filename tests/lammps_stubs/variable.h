#pragma once
#include "pointers.h"
namespace LAMMPS_NS {
class Variable : protected Pointers {
 public:
  Variable(LAMMPS *l) : Pointers(l) {}
  int find(const char *);
  int equalstyle(int);
  double compute_equal(int);
};
}  // namespace LAMMPS_NS
