#pragma once
#include "lmptype.h"
#include <mpi.h>
namespace LAMMPS_NS {
class LAMMPS;
class Memory; class Error; class Universe; class Input; class Atom; class Update; class Neighbor; class Comm;
class Domain; class Force; class Modify; class Group; class Output; class Timer;
class Pointers {
 public:
  Pointers(LAMMPS *) {}
  virtual ~Pointers() {}
 protected:
  LAMMPS *lmp;
  Memory *memory;
  Error *error;
  Universe *universe;
  Input *input;
  Atom *atom;
  Update *update;
  Neighbor *neighbor;
  Comm *comm;
  Domain *domain;
  Force *force;
  Modify *modify;
  Group *group;
  Output *output;
  Timer *timer;
  MPI_Comm world;
  FILE *screen, *logfile;
};
}  // namespace LAMMPS_NS
