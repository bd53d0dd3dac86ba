#pragma once
#include "pointers.h"
#include "error.h"
#include "group.h"
#define FixStyle(key, Class)
namespace LAMMPS_NS {
class NeighList;
namespace FixConst {
enum { INITIAL_INTEGRATE = 1, POST_INTEGRATE = 2, PRE_EXCHANGE = 4, PRE_NEIGHBOR = 8, POST_NEIGHBOR = 16,
       PRE_FORCE = 32, PRE_REVERSE = 64, POST_FORCE = 128, FINAL_INTEGRATE = 256, END_OF_STEP = 512 };
}
class Fix : protected Pointers {
 public:
  char *id, *style;
  int igroup, groupbit;
  int scalar_flag, extscalar, global_freq, nevery, vector_flag, size_vector;
  Fix(LAMMPS *l, int, char **arg) : Pointers(l), id(arg[0]), style(arg[2]) {  // fix ID group-ID style ...
    igroup = group->find(arg[1]);
    if (igroup == -1) error->all(FLERR, "Could not find fix group ID");
    groupbit = group->bitmask[igroup];
  }
  virtual int setmask() = 0;
  virtual void init() {}
  virtual void setup(int) {}
  virtual void post_integrate() {}
  virtual void init_list(int, NeighList *) {}
  virtual void setup_post_neighbor() {}
  virtual void setup_pre_force(int) {}
  virtual void post_neighbor() {}
  virtual void pre_force(int) {}
  virtual void post_force(int) {}
  virtual void end_of_step() {}
  virtual double compute_scalar() { return 0.0; }
  virtual int modify_param(int, char **) { return 0; }
};
}  // namespace LAMMPS_NS
