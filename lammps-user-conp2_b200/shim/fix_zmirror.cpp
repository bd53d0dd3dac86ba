/* See fix_zmirror.h.  Same command syntax, hooks and error messages as the reference (fix_zmirror.cpp:30-64,
   66-99); the exchange is one summed table of source positions indexed by tag offset instead of the
   reference's map()-or-Allgatherv two-path logic (:124-220) -- every rank ends with the same coordinates. */
#include "fix_zmirror.h"

#include "atom.h"
#include "domain.h"
#include "error.h"
#include "group.h"
#include "update.h"
#include "utils.h"

#include <mpi.h>

using namespace LAMMPS_NS;
using namespace FixConst;

FixZmirrorB200::FixZmirrorB200(LAMMPS *lmp, int narg, char **arg) :
    Fix(lmp, narg, arg), ngroup(0), send_mintag(0), recv_mintag(0), ran_postint(false)
{
  if (narg != 5) error->all(FLERR, "Illegal fix zmirror command (incorrect no. of parameters)");
  everynum = utils::inumeric(FLERR, arg[3], false, lmp);
  const int jgroup = group->find(arg[4]);
  if (jgroup == -1) error->all(FLERR, "Fix zmirror group ID does not exist");
  jgroupbit = group->bitmask[jgroup];
}

int FixZmirrorB200::setmask() { return POST_INTEGRATE | END_OF_STEP; }

/* tag ranges of the two groups (reference :66-99) */
void FixZmirrorB200::setup(int)
{
  const int nlocal = atom->nlocal;
  tagint lo[2] = {2147483647, 2147483647}, hi[2] = {0, 0}, glo[2], ghi[2];
  for (int i = 0; i < nlocal; ++i) {
    if (atom->mask[i] & groupbit) { lo[0] = MIN(lo[0], atom->tag[i]); hi[0] = MAX(hi[0], atom->tag[i]); }
    if (atom->mask[i] & jgroupbit) { lo[1] = MIN(lo[1], atom->tag[i]); hi[1] = MAX(hi[1], atom->tag[i]); }
  }
  MPI_Allreduce(lo, glo, 2, MPI_INT, MPI_MIN, world);
  MPI_Allreduce(hi, ghi, 2, MPI_INT, MPI_MAX, world);
  send_mintag = glo[0]; recv_mintag = glo[1];
  ngroup = ghi[0] - glo[0] + 1;
  if (ghi[1] - glo[1] + 1 != ngroup) error->all(FLERR, "Groups do not have same number of tags");
  table.assign(4 * (size_t) ngroup, 0.0);
  mine.assign(4 * (size_t) ngroup, 0.0);
}

void FixZmirrorB200::mirror()
{
  const int nlocal = atom->nlocal;
  double **x = atom->x;
  const double zoffset = 2 * domain->boxlo[2] + domain->zprd;   // reference :133
  std::fill(mine.begin(), mine.end(), 0.0);
  for (int i = 0; i < nlocal; ++i)
    if (atom->mask[i] & groupbit) {
      double *t = &mine[4 * (size_t) (atom->tag[i] - send_mintag)];
      t[0] = x[i][0]; t[1] = x[i][1]; t[2] = x[i][2]; t[3] = 1.0;
    }
  MPI_Allreduce(mine.data(), table.data(), 4 * ngroup, MPI_DOUBLE, MPI_SUM, world);
  double seen = 0.0;
  for (int k = 0; k < ngroup; ++k) seen += table[4 * (size_t) k + 3];
  if ((int) (seen + 0.5) != ngroup) error->all(FLERR, "Incorrect number of atoms communicated");   // reference :182
  for (int i = 0; i < nlocal; ++i)
    if (atom->mask[i] & jgroupbit) {
      const double *t = &table[4 * (size_t) (atom->tag[i] - recv_mintag)];
      x[i][0] = t[0]; x[i][1] = t[1]; x[i][2] = zoffset - t[2];   // reference :214-216
    }
}

void FixZmirrorB200::post_integrate()
{
  if (update->ntimestep % everynum == 0) { mirror(); ran_postint = true; }
}

void FixZmirrorB200::end_of_step()
{
  if (!ran_postint) post_integrate();   // reference :222-226
  else ran_postint = false;
}
