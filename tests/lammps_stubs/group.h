#pragma once
#include "pointers.h"
#include <vector>
namespace LAMMPS_NS {
class Group : protected Pointers {
 public:
  int ngroup;
  int *bitmask;
  std::vector<std::string> names;
  std::vector<int> bits;
  Group(LAMMPS *l) : Pointers(l), ngroup(0), bitmask(nullptr) {}
  int add(const std::string &n) { names.push_back(n); bits.push_back(1 << ngroup); bitmask = bits.data(); return ngroup++; }
  int find(const std::string &n) {
    for (int i = 0; i < ngroup; ++i)
      if (names[i] == n) return i;
    return -1;
  }
};
}  // namespace LAMMPS_NS
