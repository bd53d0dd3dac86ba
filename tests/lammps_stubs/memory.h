#pragma once
#include "pointers.h"
namespace LAMMPS_NS {
class Memory : protected Pointers {
 public:
  Memory(LAMMPS *l) : Pointers(l) {}
  template <class T> T *create(T *&a, int n, const char *) { a = new T[n]; return a; }
  template <class T> T **create(T **&a, int n1, int n2, const char *) { (void)n1; (void)n2; a = nullptr; return a; }
  template <class T> void destroy(T *&a) { delete[] a; a = nullptr; }
  template <class T> void destroy(T **&a) { a = nullptr; }
  template <class T> T *grow(T *&a, int n, const char *) { (void)n; return a; }
};
}  // namespace LAMMPS_NS
