#pragma once
#include "pointers.h"
#include <stdexcept>
namespace LAMMPS_NS {
class Error : protected Pointers {
 public:
  Error(LAMMPS *l) : Pointers(l) {}
  [[noreturn]] void all(const std::string &, int, const std::string &msg) { throw std::runtime_error(msg); }
  [[noreturn]] void one(const std::string &, int, const std::string &msg) { throw std::runtime_error(msg); }
  void warning(const std::string &, int, const std::string &msg, int = 1) { fprintf(stderr, "WARNING: %s\n", msg.c_str()); }
};
}  // namespace LAMMPS_NS
