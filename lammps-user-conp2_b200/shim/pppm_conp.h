/* `kspace_style pppm/conp` shim (reference pppm_conp.h:19-21, 26-78): a LAMMPS
   PPPM that (a) hands its mesh tables to the library once (the KSpaceModule
   seam: conp_setup / a_cal / b_cal are now library calls made by the fix) and
   (b) overrides make_rho() so the force pass reuses the electrolyte + electrode
   density the charge solve already spread (reference pppm_conp.cpp:428-450). */
#ifdef KSPACE_CLASS
// clang-format off
KSpaceStyle(pppm/conp,PPPMCONPB200)
// clang-format on
#else
#ifndef LMP_PPPM_CONP_B200_H
#define LMP_PPPM_CONP_B200_H

#include "pppm.h"

#include <vector>

struct conp_ctx;

namespace LAMMPS_NS {

class PPPMCONPB200 : public PPPM {
 public:
  PPPMCONPB200(class LAMMPS *l) : PPPM(l), ctx(nullptr), tables_sent(false), density_ready(false) {}
  // PPPM::setup, then (once, after attach) conp_pppm_setup(mesh, order, rho_coeff, greensfn, shift, shiftone)
  void setup() override;
  // called by FixConpB200::linalg_init when the `pppm` keyword is given (reference: dynamic_cast of
  // force->kspace to KSpaceModule + register_fix, fix_conp.cpp:401-409)
  void attach(conp_ctx *c) { ctx = c; tables_sent = false; density_ready = false; }
  // called by the fix after every charge solve: the library holds the densities of this step
  void charges_updated() { density_ready = true; }

 protected:
  void make_rho() override;   // density_brick <- conp_get_density_region(ctx, 2, owned sub-brick)
  conp_ctx *ctx;
  bool tables_sent, density_ready;
  std::vector<double> region;
};

}    // namespace LAMMPS_NS
#endif
#endif
