/* `fix ID group zmirror Nevery group2` (reference fix_zmirror.h/.cpp): every Nevery steps, after the
   integrator has moved the atoms, the atoms of group2 become the mirror image of the atoms of `group`
   in the plane z = (zlo + zhi)/2, matched by tag offset (k-th tag of group -> k-th tag of group2).  Used
   by the doubled-cell (noslab zneutr) decks.  Host-side atom plumbing only: nothing here touches the GPU
   library; it is part of the shim so that those decks run unchanged. */
#ifdef FIX_CLASS
// clang-format off
FixStyle(zmirror,FixZmirrorB200)
// clang-format on
#else
#ifndef LMP_FIX_ZMIRROR_B200_H
#define LMP_FIX_ZMIRROR_B200_H

#include "fix.h"

#include <vector>

namespace LAMMPS_NS {

class FixZmirrorB200 : public Fix {
 public:
  FixZmirrorB200(class LAMMPS *, int, char **);
  int setmask() override;
  void setup(int) override;
  void post_integrate() override;
  void end_of_step() override;

 protected:
  void mirror();
  int everynum, jgroupbit, ngroup;
  tagint send_mintag, recv_mintag;
  bool ran_postint;
  std::vector<double> table, mine;   // [ngroup][4]: x, y, z, "seen" of the source atoms, by tag offset
};

}    // namespace LAMMPS_NS
#endif
#endif
