#pragma once
#include "pointers.h"
namespace LAMMPS_NS {
class Variable;
class Input : protected Pointers {
 public:
  Variable *variable;
  Input(LAMMPS *l) : Pointers(l), variable(nullptr) {}
};
}  // namespace LAMMPS_NS
