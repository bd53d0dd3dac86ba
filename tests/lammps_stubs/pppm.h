#ifdef KSPACE_CLASS
#else
#pragma once
#include "kspace.h"
#include <algorithm>
#include <vector>
namespace LAMMPS_NS {
// single-rank PPPM: the brick is the whole mesh plus LAMMPS' ghost layers; the tables are loaded by the mock
// host (mock_lammps_main.cpp) from what the Python tests generated, setup() has nothing left to compute
class PPPM : public KSpace {
 public:
  PPPM(LAMMPS *l) : KSpace(l), setups(0), nlower(0), nupper(0), nfft(0), ngrid(0), shift(0), shiftone(0),
      rho_coeff(nullptr), greensfn(nullptr), density_brick(nullptr) {}
  void setup() override { ++setups; }
  // mock host API ------------------------------------------------------------------
  void mock_tables(const int mesh[3], int order_, const double *rc, const double *gf, double shift_, double shiftone_) {
    nx_pppm = mesh[0]; ny_pppm = mesh[1]; nz_pppm = mesh[2]; order = order_;
    nlower = -(order - 1) / 2; nupper = order / 2;
    shift = shift_; shiftone = shiftone_;
    rc_store.assign(rc, rc + (size_t)order * order);
    rc_rows.resize(order);
    for (int l = 0; l < order; ++l) rc_rows[l] = rc_store.data() + (size_t)l * order - nlower;  // k = nlower..nupper
    rho_coeff = rc_rows.data();
    nxlo_in = nylo_in = nzlo_in = nxlo_fft = nylo_fft = nzlo_fft = 0;
    nxhi_in = nxhi_fft = nx_pppm - 1; nyhi_in = nyhi_fft = ny_pppm - 1; nzhi_in = nzhi_fft = nz_pppm - 1;
    nxlo_out = nxlo_in + nlower - 1; nxhi_out = nxhi_in + nupper + 1;   // stencil reach + skin cell, as LAMMPS sizes them
    nylo_out = nylo_in + nlower - 1; nyhi_out = nyhi_in + nupper + 1;
    nzlo_out = nzlo_in + nlower - 1; nzhi_out = nzhi_in + nupper + 1;
    nfft = nx_pppm * ny_pppm * nz_pppm;
    gf_store.assign(gf, gf + nfft);
    greensfn = gf_store.data();
    const int ex = nxhi_out - nxlo_out + 1, ey = nyhi_out - nylo_out + 1, ez = nzhi_out - nzlo_out + 1;
    ngrid = ex * ey * ez;
    brick.assign(ngrid, -1.0);
    rows.resize((size_t)ez * ey);
    planes.resize(ez);
    for (int z = 0; z < ez; ++z) {
      for (int y = 0; y < ey; ++y) rows[(size_t)z * ey + y] = brick.data() + ((size_t)z * ey + y) * ex - nxlo_out;
      planes[z] = rows.data() + (size_t)z * ey - nylo_out;
    }
    density_brick = planes.data() - nzlo_out;
  }
  void mock_make_rho() { make_rho(); }   // what PPPM::compute does first
  double mock_density(int ix, int iy, int iz) const { return density_brick[iz][iy][ix]; }
  int setups;

 protected:
  int nlower, nupper;
  int nxlo_in, nylo_in, nzlo_in, nxhi_in, nyhi_in, nzhi_in;
  int nxlo_out, nylo_out, nzlo_out, nxhi_out, nyhi_out, nzhi_out;
  int nxlo_fft, nylo_fft, nzlo_fft, nxhi_fft, nyhi_fft, nzhi_fft;
  int nfft, ngrid;
  double shift, shiftone;
  double **rho_coeff;
  double *greensfn;
  double ***density_brick;
  virtual void particle_map() {}
  virtual void make_rho() { std::fill(brick.begin(), brick.end(), 0.0); }
  std::vector<double> rc_store, gf_store, brick;
  std::vector<double *> rc_rows, rows;
  std::vector<double **> planes;
};
}  // namespace LAMMPS_NS
#endif
