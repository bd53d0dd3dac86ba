#pragma once
#include "pointers.h"
namespace LAMMPS_NS {
class Atom : protected Pointers {
 public:
  bigint natoms;
  int nlocal, nghost, nmax, ntypes;
  tagint *tag;
  int *type, *mask;
  double **x, **f, *q;
  tagint *molecule;
  int map(tagint);
  Atom(LAMMPS *l) : Pointers(l) {}
};
}  // namespace LAMMPS_NS
