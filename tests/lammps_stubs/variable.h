#pragma once
#include "pointers.h"
namespace LAMMPS_NS {
class Variable : protected Pointers {
 public:
  std::string name;  // one equal-style variable with a constant value is enough for the v_ token
  double value;
  Variable(LAMMPS *l) : Pointers(l), value(0.0) {}
  int find(const char *n) { return name == n ? 0 : -1; }
  int equalstyle(int) { return 1; }
  double compute_equal(int) { return value; }
};
}  // namespace LAMMPS_NS
