/* LAMMPS-side shim for the B200 charge solve: keeps the deck syntax of
   USER-CONP2 (`fix ID group conp|conq|cond Nevery group2 eta value logfile
   [keywords]`, reference fix_conp.cpp:79-176) and the fix hook order
   (setmask: POST_NEIGHBOR | PRE_FORCE | POST_FORCE | END_OF_STEP, reference
   fix_conp.cpp:233-241) and forwards every numerical step to libconp_b200.so
   through include/conp_b200.h.  Compiles only inside a LAMMPS (27May2021)
   source tree; see INTEGRATION.md.  Nothing here is copied from the reference:
   the reference's members (matrices, cross-lists, timers) have no counterpart
   because that state lives on the GPU. */
#ifdef FIX_CLASS
// clang-format off
FixStyle(conp,FixConpB200)
FixStyle(conq,FixConqB200)
FixStyle(cond,FixCondB200)
// clang-format on
#else
#ifndef LMP_FIX_CONP_B200_H
#define LMP_FIX_CONP_B200_H

#include "fix.h"

#include <cstdio>
#include <string>
#include <vector>

struct conp_ctx;

namespace LAMMPS_NS {

class FixConpB200 : public Fix {
 public:
  FixConpB200(class LAMMPS *, int, char **);
  ~FixConpB200() override;
  int setmask() override;
  void init() override;
  void setup_post_neighbor() override;
  void setup_pre_force(int) override;
  void post_neighbor() override;
  void pre_force(int) override;
  void post_force(int) override;
  void end_of_step() override;
  double compute_scalar() override;
  int modify_param(int, char **) override;

  conp_ctx *context() { return ctx; }  // used by PPPMCONPB200 (the KSpaceModule seam)

 protected:
  virtual int variant() const { return 0; }  // CONP_VARIANT_CONP
  void check(int status);                    // non-zero status -> error->all(FLERR, conp_last_error())
  void linalg_init();                        // reference linalg_init  (fix_conp.cpp:393-424)
  void linalg_setup();                       // reference linalg_setup (fix_conp.cpp:426-464)
  void scatter_charges();
  void read_matrix_file(std::vector<double> &full);             // `org` / `inv` (reference a_read :721-773)
  void write_matrix(const char *name, const char *fmt);          // `matout` (reference :833-849, 960-977)
  void gather_electrode_charges(std::vector<double> &qall);      // `qinit` snapshot (reference :1107-1114)

  conp_ctx *ctx;
  int everynum, jgroup, jgroupbit;
  double eta, potdiff, evscale, scalar_output;
  char *potdiffstr;
  int potdiffvar;
  int ff_flag, a_matrix_f, pairmode;
  bool smartlist, zneutrflag, matoutflag, pppmflag, qinitflag, lowmemflag, nullneutralflag;
  bool one_electrode_flag, postforceflag, init_done, setup_done;
  std::string a_matrix_file, logfile;
  std::vector<int> is_eletype;            // etypes keyword
  std::vector<double> eta_i, u0_i;        // fix_modify ... ehgo coeff
  double kappa;
  // global electrode numbering (what the reference calls eleall)
  int elenum_all;
  std::vector<int> eleall2tag, eleall_side, eleall_type;
  std::vector<double> eleall_x, eleallq;
  int *tag2eleall;
  int maxtag_all;
  class Pair *coulpair;
  class PPPMCONPB200 *pppm;               // force->kspace when the `pppm` keyword is given
  FILE *outf;                             // the fix's log file (arg 7), rank 0 only
  double Btime, Ctime, Ktime, pair_share, kspace_share;   // reference timers (fix_conp.cpp:553-568)
  int stage_samples;
};

class FixConqB200 : public FixConpB200 {
 public:
  FixConqB200(class LAMMPS *l, int n, char **a) : FixConpB200(l, n, a) {}
 protected:
  int variant() const override { return 1; }  // CONP_VARIANT_CONQ
};

class FixCondB200 : public FixConpB200 {
 public:
  FixCondB200(class LAMMPS *l, int n, char **a) : FixConpB200(l, n, a) {}
 protected:
  int variant() const override { return 2; }  // CONP_VARIANT_COND
};

}    // namespace LAMMPS_NS
#endif
#endif
