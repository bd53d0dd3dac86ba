/* `kspace_style pppm/conp` shim (reference pppm_conp.h:19-21, 26-78): a LAMMPS
   PPPM that (a) hands its mesh tables to the library once (the KSpaceModule
   seam: conp_setup / a_cal / b_cal are now library calls made by the fix) and
   (b) overrides particle_map()/make_rho() so the force pass reuses the
   electrolyte + electrode density the charge solve already spread
   (reference pppm_conp.cpp:428-450). */
#ifdef KSPACE_CLASS
// clang-format off
KSpaceStyle(pppm/conp,PPPMCONPB200)
// clang-format on
#else
#ifndef LMP_PPPM_CONP_B200_H
#define LMP_PPPM_CONP_B200_H

#include "pppm.h"

struct conp_ctx;

namespace LAMMPS_NS {

class PPPMCONPB200 : public PPPM {
 public:
  PPPMCONPB200(class LAMMPS *l) : PPPM(l), ctx(nullptr), tables_sent(false) {}
  void setup() override;        // PPPM::setup, then conp_pppm_setup(mesh, order, rho_coeff, greensfn, shift, shiftone)
  void attach(conp_ctx *c) { ctx = c; tables_sent = false; }   // called by FixConpB200 when the `pppm` keyword is given

 protected:
  void particle_map() override; // no-op after the first solve: the library mapped the atoms
  void make_rho() override;     // density_brick <- conp_get_density(ctx, 2, ...) restricted to this rank's brick
  conp_ctx *ctx;
  bool tables_sent;
};

}    // namespace LAMMPS_NS
#endif
#endif
