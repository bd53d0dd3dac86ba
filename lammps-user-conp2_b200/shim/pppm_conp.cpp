/* See pppm_conp.h.  Requires the host PPPM to run with one full-mesh brick per
   rank being addressable (nprocs == 1) or a gather of the local bricks; the
   per-step charge solve itself never touches LAMMPS' FFT3d/GridComm. */
#include "pppm_conp.h"

#include "atom.h"
#include "error.h"
#include "memory.h"

#include "conp_b200.h"

#include <vector>

using namespace LAMMPS_NS;

void PPPMCONPB200::setup()
{
  PPPM::setup();   // computes greensfn for the current box (reference inherits the same call)
  if (!ctx || tables_sent) return;
  /* rho_coeff[l][k], k = nlower..nupper -> dense [order][order] */
  std::vector<double> rc((size_t) order * order);
  for (int l = 0; l < order; ++l)
    for (int k = nlower; k <= nupper; ++k) rc[(size_t) l * order + (k - nlower)] = rho_coeff[l][k];
  /* greensfn on the full mesh in FFT-grid order (z slowest, x fastest); single-rank FFT decomposition
     assumed here, otherwise gather nfft values per rank with MPI_Gatherv in the same order */
  std::vector<double> gf((size_t) nx_pppm * ny_pppm * nz_pppm);
  for (size_t i = 0; i < gf.size(); ++i) gf[i] = greensfn[i];
  const int mesh[3] = {nx_pppm, ny_pppm, nz_pppm};
  if (conp_pppm_setup(ctx, mesh, order, rc.data(), gf.data(), shift, shiftone))
    error->all(FLERR, conp_last_error(ctx));
  tables_sent = true;
}

void PPPMCONPB200::particle_map()
{
  if (!ctx) PPPM::particle_map();   // reference :428-432 (first_bcal / fixconp == nullptr)
}

void PPPMCONPB200::make_rho()
{
  if (!ctx) { PPPM::make_rho(); return; }
  /* reference :434-450: density_brick = elyte_density_brick + ele_density_brick */
  std::vector<double> full((size_t) nx_pppm * ny_pppm * nz_pppm);
  if (conp_get_density(ctx, 2, full.data())) error->all(FLERR, conp_last_error(ctx));
  for (int iz = nzlo_out; iz <= nzhi_out; ++iz)
    for (int iy = nylo_out; iy <= nyhi_out; ++iy)
      for (int ix = nxlo_out; ix <= nxhi_out; ++ix) {
        const int gx = (ix % nx_pppm + nx_pppm) % nx_pppm, gy = (iy % ny_pppm + ny_pppm) % ny_pppm,
                  gz = (iz % nz_pppm + nz_pppm) % nz_pppm;
        const bool owned = ix >= nxlo_in && ix <= nxhi_in && iy >= nylo_in && iy <= nyhi_in && iz >= nzlo_in && iz <= nzhi_in;
        density_brick[iz][iy][ix] = owned ? full[((size_t) gz * ny_pppm + gy) * nx_pppm + gx] : 0.0;
      }
}
