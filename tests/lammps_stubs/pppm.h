#ifdef KSPACE_CLASS
#else
#pragma once
#include "kspace.h"
#define KSpaceStyle(key, Class)
namespace LAMMPS_NS {
class PPPM : public KSpace {
 public:
  PPPM(LAMMPS *l) : KSpace(l) {}
  void setup() override {}
 protected:
  int nlower, nupper;
  int nxlo_in, nylo_in, nzlo_in, nxhi_in, nyhi_in, nzhi_in;
  int nxlo_out, nylo_out, nzlo_out, nxhi_out, nyhi_out, nzhi_out;
  int nfft, ngrid;
  double shift, shiftone;
  double **rho_coeff;
  double *greensfn;
  double ***density_brick;
  virtual void particle_map() {}
  virtual void make_rho() {}
};
}  // namespace LAMMPS_NS
#endif
