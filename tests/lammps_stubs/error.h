#pragma once
#include "pointers.h"
namespace LAMMPS_NS {
class Error : protected Pointers {
 public:
  Error(LAMMPS *l) : Pointers(l) {}
  [[noreturn]] void all(const std::string &, int, const std::string &);
  [[noreturn]] void one(const std::string &, int, const std::string &);
  void warning(const std::string &, int, const std::string &, int = 1);
};
}  // namespace LAMMPS_NS
