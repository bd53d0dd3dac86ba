#pragma once
#include "lmptype.h"
#include <mpi.h>
namespace LAMMPS_NS {
class Memory; class Error; class Universe; class Input; class Atom; class Update; class Neighbor; class Comm;
class Domain; class Force; class Modify; class Group; class Output; class Timer;
class LAMMPS {  // the mock host fills these before any Pointers-derived object is created
 public:
  Memory *memory = nullptr; Error *error = nullptr; Universe *universe = nullptr; Input *input = nullptr;
  Atom *atom = nullptr; Update *update = nullptr; Neighbor *neighbor = nullptr; Comm *comm = nullptr;
  Domain *domain = nullptr; Force *force = nullptr; Modify *modify = nullptr; Group *group = nullptr;
  Output *output = nullptr; Timer *timer = nullptr;
  MPI_Comm world = MPI_COMM_WORLD;
  FILE *screen = nullptr, *logfile = nullptr;
  std::string log;  // what utils::logmesg received
};
class Pointers {
 public:
  Pointers(LAMMPS *l) : lmp(l), memory(l->memory), error(l->error), universe(l->universe), input(l->input),
      atom(l->atom), update(l->update), neighbor(l->neighbor), comm(l->comm), domain(l->domain), force(l->force),
      modify(l->modify), group(l->group), output(l->output), timer(l->timer), world(l->world), screen(l->screen),
      logfile(l->logfile) {}
  virtual ~Pointers() {}
 protected:
  LAMMPS *lmp;
  Memory *&memory;
  Error *&error;
  Universe *&universe;
  Input *&input;
  Atom *&atom;
  Update *&update;
  Neighbor *&neighbor;
  Comm *&comm;
  Domain *&domain;
  Force *&force;
  Modify *&modify;
  Group *&group;
  Output *&output;
  Timer *&timer;
  MPI_Comm &world;
  FILE *&screen, *&logfile;
};
}  // namespace LAMMPS_NS
