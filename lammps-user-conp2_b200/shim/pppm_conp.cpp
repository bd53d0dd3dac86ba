/* See pppm_conp.h.  The per-step charge solve never touches LAMMPS' FFT3d/GridComm; this class only moves
   tables in (once) and the density out (every step, this rank's sub-brick only). */
#include "pppm_conp.h"

#include "atom.h"
#include "comm.h"
#include "error.h"
#include "memory.h"

#include "conp_b200.h"

#include <cstring>
#include <mpi.h>

using namespace LAMMPS_NS;

void PPPMCONPB200::setup()
{
  PPPM::setup();   // computes greensfn for the current box (the reference inherits the same call)
  if (!ctx || tables_sent) return;
  /* rho_coeff[l][k], k = nlower..nupper -> dense [order][order] */
  std::vector<double> rc((size_t) order * order);
  for (int l = 0; l < order; ++l)
    for (int k = nlower; k <= nupper; ++k) rc[(size_t) l * order + (k - nlower)] = rho_coeff[l][k];
  /* greensfn: every rank holds the nfft entries of its FFT-decomposition sub-box, ordered z slowest /
     x fastest over [nzlo_fft, nzhi_fft] x [nylo_fft, nyhi_fft] x [nxlo_fft, nxhi_fft] (PPPM::compute_gf_ik).
     The library wants the whole mesh: every rank drops its entries into a zeroed full-mesh array and the
     arrays are summed (setup only; on one rank the sum is a copy). */
  const size_t nmesh = (size_t) nx_pppm * ny_pppm * nz_pppm;
  std::vector<double> mine(nmesh, 0.0), gf(nmesh, 0.0);
  size_t n = 0;
  for (int m = nzlo_fft; m <= nzhi_fft; ++m)
    for (int l = nylo_fft; l <= nyhi_fft; ++l)
      for (int k = nxlo_fft; k <= nxhi_fft; ++k) mine[((size_t) m * ny_pppm + l) * nx_pppm + k] = greensfn[n++];
  if (comm->nprocs > 1) {
    /* MPI counts are int: sum in slices */
    const size_t slice = (size_t) 1 << 26;
    for (size_t o = 0; o < nmesh; o += slice) {
      const int cnt = (int) ((nmesh - o < slice) ? nmesh - o : slice);
      MPI_Allreduce(mine.data() + o, gf.data() + o, cnt, MPI_DOUBLE, MPI_SUM, world);
    }
  } else gf.swap(mine);
  const int mesh[3] = {nx_pppm, ny_pppm, nz_pppm};
  int status = conp_pppm_setup(ctx, mesh, order, rc.data(), gf.data(), shift, shiftone), any = 0;
  MPI_Allreduce(&status, &any, 1, MPI_INT, MPI_MAX, world);
  if (any) error->all(FLERR, status ? conp_last_error(ctx) : "pppm/conp setup failed on another rank");
  tables_sent = true;
  density_ready = false;
}

/* particle_map() is NOT overridden: PPPM::fieldforce indexes part2grid for every local atom, and the library
   keeps its own cell map on the GPU (the reference can skip the call only because elyte_particle_map /
   aaa_map_rho filled part2grid in the same step, pppm_conp.cpp:428-432). */

void PPPMCONPB200::make_rho()
{
  if (!ctx || !density_ready) { PPPM::make_rho(); return; }   // reference :435 (first_bcal / no fix)
  /* reference :434-450: density_brick = elyte_density_brick + ele_density_brick.  The library returns the
     finished (ghost contributions already folded in) density on this rank's owned sub-brick; the ghost
     layers are zero, so the reverse_comm that follows in PPPM::compute adds nothing. */
  const int lo[3] = {nxlo_in, nylo_in, nzlo_in}, hi[3] = {nxhi_in, nyhi_in, nzhi_in};
  const int ex = hi[0] - lo[0] + 1, ey = hi[1] - lo[1] + 1, ez = hi[2] - lo[2] + 1;
  region.resize((size_t) ex * ey * ez);
  int status = conp_get_density_region(ctx, 2, lo, hi, region.data()), any = 0;
  MPI_Allreduce(&status, &any, 1, MPI_INT, MPI_MAX, world);
  if (any) error->all(FLERR, status ? conp_last_error(ctx) : "pppm/conp density hand-off failed on another rank");
  memset(&(density_brick[nzlo_out][nylo_out][nxlo_out]), 0, ngrid * sizeof(FFT_SCALAR));
  size_t n = 0;
  for (int iz = lo[2]; iz <= hi[2]; ++iz)
    for (int iy = lo[1]; iy <= hi[1]; ++iy)
      for (int ix = lo[0]; ix <= hi[0]; ++ix) density_brick[iz][iy][ix] = region[n++];
}
