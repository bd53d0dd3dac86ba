/* See fix_conp.h.  Every conp_* call names the reference code it stands in
   for (file:line in the USER-CONP2 tree). */
#include "fix_conp.h"

#include "atom.h"
#include "comm.h"
#include "domain.h"
#include "error.h"
#include "force.h"
#include "group.h"
#include "input.h"
#include "kspace.h"
#include "memory.h"
#include "pair.h"
#include "update.h"
#include "utils.h"
#include "variable.h"

#include "conp_b200.h"
#include "pppm_conp.h"

#include <cmath>
#include <cstring>
#include <mpi.h>

using namespace LAMMPS_NS;
using namespace FixConst;

/* keyword parsing: same tokens and messages as reference fix_conp.cpp:86-176 */
FixConpB200::FixConpB200(LAMMPS *lmp, int narg, char **arg) :
    Fix(lmp, narg, arg), ctx(nullptr), potdiffstr(nullptr), potdiffvar(-1), tag2eleall(nullptr), coulpair(nullptr),
    pppm(nullptr), outf(nullptr)
{
  if (narg < 8) error->all(FLERR, "Illegal fix conp command (too few input parameters)");
  everynum = utils::inumeric(FLERR, arg[3], false, lmp);
  jgroup = group->find(arg[4]);
  if (jgroup == -1) error->all(FLERR, "Fix conp group ID does not exist");
  jgroupbit = group->bitmask[jgroup];
  eta = utils::numeric(FLERR, arg[5], false, lmp);
  potdiff = 0.0;
  if (utils::strmatch(arg[6], "^v_")) potdiffstr = utils::strdup(arg[6] + 2);
  else potdiff = utils::numeric(FLERR, arg[6], false, lmp);
  logfile = arg[7];
  ff_flag = CONP_FF_NORMAL; a_matrix_f = 0; pairmode = CONP_PAIR_ETA;
  smartlist = zneutrflag = matoutflag = pppmflag = qinitflag = false;
  lowmemflag = nullneutralflag = true;
  kappa = 1.0;
  is_eletype.assign(atom->ntypes + 1, 0);
  eta_i.assign(atom->ntypes + 1, 0.0);
  u0_i.assign(atom->ntypes + 1, 0.0);
  for (int iarg = 8; iarg < narg; ++iarg) {
    if (strcmp(arg[iarg], "ffield") == 0) {
      if (ff_flag == CONP_FF_NOSLAB)
        error->all(FLERR, "Invalid fix conp command (ffield and noslab cannot both be chosen)");
      ff_flag = CONP_FF_FFIELD;
    } else if (strcmp(arg[iarg], "noslab") == 0) {
      if (ff_flag == CONP_FF_FFIELD)
        error->all(FLERR, "Invalid fix conp command (ffield and noslab cannot both be chosen)");
      ff_flag = CONP_FF_NOSLAB;
    } else if (strcmp(arg[iarg], "org") == 0 || strcmp(arg[iarg], "inv") == 0) {
      if (a_matrix_f != 0) error->all(FLERR, "Invalid fix conp command (A matrix file specified more than once)");
      a_matrix_f = (strcmp(arg[iarg], "org") == 0) ? 1 : 2;
      if (++iarg >= narg) error->all(FLERR, "Invalid fix conp command (No A matrix filename given)");
      a_matrix_file = arg[iarg];
    } else if (strcmp(arg[iarg], "etypes") == 0) {
      if (++iarg >= narg - 1) error->all(FLERR, "Invalid fix conp command (Insufficient input entries for etypes)");
      const int n = utils::inumeric(FLERR, arg[iarg], false, lmp);
      for (int i = 0; i < n; ++i) {
        if (++iarg >= narg) error->all(FLERR, "Invalid fix conp command (Insufficient input entries for etypes)");
        const int t = utils::inumeric(FLERR, arg[iarg], false, lmp);
        if (t < 1 || t > atom->ntypes) error->all(FLERR, "Invalid fix conp command (Invalid atom type in etypes)");
        is_eletype[t] = 1;
      }
      smartlist = true;
    } else if (strcmp(arg[iarg], "zneutr") == 0) zneutrflag = true;
    else if (strcmp(arg[iarg], "matout") == 0) matoutflag = true;
    else if (strcmp(arg[iarg], "pppm") == 0) pppmflag = true;
    else if (strcmp(arg[iarg], "qinit") == 0) qinitflag = true;
    else if (strcmp(arg[iarg], "himem") == 0) lowmemflag = false;
    else if (strcmp(arg[iarg], "nonneutral") == 0) nullneutralflag = false;
    else if (strcmp(arg[iarg], "ehgo") == 0) pairmode = CONP_PAIR_EHGO;
    else error->all(FLERR, std::string("Invalid fix conp commmand (unknown option: ") + arg[iarg] + ")");
  }
  scalar_flag = 1; extscalar = 0; global_freq = 1;
  scalar_output = 0.0;
  postforceflag = init_done = setup_done = false;
  one_electrode_flag = false;
  elenum_all = 0; maxtag_all = 0;
  Btime = Ctime = Ktime = pair_share = kspace_share = 0.0;
  stage_samples = 0;
  if (comm->me == 0) outf = fopen(logfile.c_str(), "w");   // reference fix_conp.cpp:119

  /* one context per MPI rank / GPU; rank 0 makes the NCCL id, MPI_Bcast on `world` carries it
     (replaces nothing in the reference: its collectives run on `world` directly) */
  char uid[CONP_UNIQUE_ID_BYTES];
  if (comm->me == 0 && comm->nprocs > 1) check(conp_get_unique_id(uid));
  MPI_Bcast(uid, CONP_UNIQUE_ID_BYTES, MPI_BYTE, 0, world);
  /* one GPU per rank: node-local rank -> device index (CONP_GPUS_PER_NODE overrides the device count) */
  MPI_Comm node;
  MPI_Comm_split_type(world, MPI_COMM_TYPE_SHARED, comm->me, MPI_INFO_NULL, &node);
  int local_rank = 0, ndev = 0;
  MPI_Comm_rank(node, &local_rank);
  MPI_Comm_free(&node);
  const char *v = getenv("CONP_GPUS_PER_NODE");
  if (v) ndev = atoi(v);
  else conp_device_count(&ndev);
  if (ndev < 1) error->all(FLERR, "fix conp (B200): no sm_100 CUDA device on this node; there is no CPU fallback");
  int status = conp_create(&ctx, local_rank % ndev, comm->me, comm->nprocs, comm->nprocs > 1 ? uid : nullptr), any = 0;
  MPI_Allreduce(&status, &any, 1, MPI_INT, MPI_MAX, world);
  if (any) error->all(FLERR, status ? conp_last_error(nullptr) : "conp_create failed on another rank");
}

FixConpB200::~FixConpB200()
{
  conp_destroy(ctx);
  if (outf) fclose(outf);   // reference :208
  delete[] potdiffstr;
  delete[] tag2eleall;
}

void FixConpB200::check(int status)
{
  int any = 0;
  MPI_Allreduce(&status, &any, 1, MPI_INT, MPI_MAX, world);  // collective abort, like error->all in the reference
  if (any) error->all(FLERR, status ? conp_last_error(ctx) : "fix conp failed on another rank");
}

int FixConpB200::setmask() { return POST_NEIGHBOR | PRE_FORCE | POST_FORCE | END_OF_STEP; }

void FixConpB200::init()
{
  coulpair = (Pair *) force->pair_match("coul", 0);
  if (coulpair == nullptr) coulpair = (Pair *) force->pair_match("coul", 0, 1);
  if (coulpair == nullptr) error->all(FLERR, "Fix conp couldn't detect a Coulombic pair style");
  if (potdiffstr) {
    potdiffvar = input->variable->find(potdiffstr);
    if (potdiffvar < 0) error->all(FLERR, "Fix conp potential difference variable does not exist");
    if (!input->variable->equalstyle(potdiffvar))
      error->all(FLERR, "Fix conp potential difference variable is invalid style");
  }
  one_electrode_flag = (groupbit == jgroupbit);
  /* no neighbor->request(): the library builds its own cell lists on the GPU */
}

int FixConpB200::modify_param(int narg, char **arg)
{
  /* fix_modify ID ehgo kappa K | ehgo coeff <types> eta u0|auto  (reference fix_conp.cpp:1482-1515) */
  if (pairmode == CONP_PAIR_ETA) error->all(FLERR, "Can't fix_modify conp parameters in basic pair mode");
  const double s2overpis = sqrt(2.0) / 1.77245385090551602729;
  const double evs = force->qe2f / force->qqr2e;
  if (strcmp(arg[0], "ehgo") == 0) {
    if (strcmp(arg[1], "kappa") == 0) {
      if (narg != 3) error->all(FLERR, "Invalid number of inputs for EHGO coeff setting");
      kappa = utils::numeric(FLERR, arg[2], false, lmp);
      return 3;
    } else if (strcmp(arg[1], "coeff") == 0) {
      if (narg != 5) error->all(FLERR, "Invalid number of inputs for EHGO coeff setting");
      int ilo, ihi;
      utils::bounds(FLERR, arg[2], 1, atom->ntypes, ilo, ihi, error);
      const double eta_one = utils::numeric(FLERR, arg[3], false, lmp);
      const double u0_one = (strcmp(arg[4], "auto") == 0) ? s2overpis * eta_one / evs
                                                           : utils::numeric(FLERR, arg[4], false, lmp);
      if (ilo > ihi) error->all(FLERR, "Couldn't set EHGO coeffs with mintype more than maxtype");
      for (int i = ilo; i <= ihi; ++i) { eta_i[i] = eta_one; u0_i[i] = u0_one * evs; }
      return 5;
    } else error->all(FLERR, "Invalid entry for EHGO coeff setting");
  }
  return 0;
}

/* FixConp::linalg_init (reference fix_conp.cpp:393-424): everything that does not need the k-space style
   to be set up -- geometry, pair data, the global electrode list */
void FixConpB200::linalg_init()
{
  if (pppmflag) {   // reference :400-404
    pppm = dynamic_cast<PPPMCONPB200 *>(force->kspace);
    if (pppm == nullptr)
      error->all(FLERR, "Fix conp couldn't detect a pppm/conp kspace style (which is required with the pppm flag)");
  }
  const int nlocal = atom->nlocal;
  int *mask = atom->mask, *type = atom->type;
  tagint *tag = atom->tag;
  double **x = atom->x;

  /* global electrode list in rank-concatenated order (reference post_neighbor :478-525) */
  std::vector<int> ltag, lside, ltype;
  std::vector<double> lx;
  for (int i = 0; i < nlocal; ++i) {
    const int side = (mask[i] & groupbit) ? 1 : ((mask[i] & jgroupbit) ? -1 : 0);  // electrode_check :599-605
    if (!side) continue;
    ltag.push_back(tag[i]); lside.push_back(side); ltype.push_back(type[i]);
    lx.insert(lx.end(), {x[i][0], x[i][1], x[i][2]});
  }
  const int nprocs = comm->nprocs;
  int n = (int) ltag.size();
  std::vector<int> counts(nprocs), displs(nprocs), counts3(nprocs), displs3(nprocs);
  MPI_Allgather(&n, 1, MPI_INT, counts.data(), 1, MPI_INT, world);
  elenum_all = 0;
  for (int p = 0; p < nprocs; ++p) { displs[p] = elenum_all; elenum_all += counts[p]; counts3[p] = 3 * counts[p]; displs3[p] = 3 * displs[p]; }
  eleall2tag.resize(elenum_all); eleall_side.resize(elenum_all); eleall_type.resize(elenum_all);
  eleall_x.resize(3 * (size_t) elenum_all); eleallq.resize(elenum_all);
  MPI_Allgatherv(ltag.data(), n, MPI_INT, eleall2tag.data(), counts.data(), displs.data(), MPI_INT, world);
  MPI_Allgatherv(lside.data(), n, MPI_INT, eleall_side.data(), counts.data(), displs.data(), MPI_INT, world);
  MPI_Allgatherv(ltype.data(), n, MPI_INT, eleall_type.data(), counts.data(), displs.data(), MPI_INT, world);
  MPI_Allgatherv(lx.data(), 3 * n, MPI_DOUBLE, eleall_x.data(), counts3.data(), displs3.data(), MPI_DOUBLE, world);
  tagint maxtag = 0;
  for (int i = 0; i < nlocal; ++i) maxtag = MAX(maxtag, tag[i]);
  MPI_Allreduce(&maxtag, &maxtag_all, 1, MPI_INT, MPI_MAX, world);
  delete[] tag2eleall;
  tag2eleall = new int[maxtag_all + 1];
  for (int t = 0; t <= maxtag_all; ++t) tag2eleall[t] = -1;
  for (int e = 0; e < elenum_all; ++e) tag2eleall[eleall2tag[e]] = e;

  /* KSpaceModuleEwald::conp_setup inputs (reference km_ewald.cpp:63-89) */
  double qsqsum = 0.0, qsq_all;
  for (int i = 0; i < nlocal; ++i) qsqsum += atom->q[i] * atom->q[i];
  MPI_Allreduce(&qsqsum, &qsq_all, 1, MPI_DOUBLE, MPI_SUM, world);
  const double q2 = qsq_all * force->qqrd2e / force->dielectric;
  double prd[3] = {domain->xprd, domain->yprd, domain->zprd};
  int periodic[3] = {domain->xperiodic, domain->yperiodic, domain->zperiodic};
  check(conp_set_cell(ctx, domain->boxlo, prd, periodic, force->kspace->slabflag, force->kspace->slab_volfactor, ff_flag));
  check(conp_set_ewald(ctx, force->kspace->g_ewald, force->kspace->accuracy, q2, (long long) atom->natoms, lowmemflag));

  /* pair data (reference fix_conp.cpp:1232-1238) + EHGO tables (:1517-1559) */
  const int ntypes = atom->ntypes, n1 = ntypes + 1;
  int itmp;
  double *p_cut_coul = (double *) coulpair->extract("cut_coul", itmp);
  std::vector<double> cutsq((size_t) n1 * n1, 0.0), eta_ij, fo_ij;
  for (int i = 1; i <= ntypes; ++i)
    for (int j = 1; j <= ntypes; ++j) cutsq[(size_t) i * n1 + j] = coulpair->cutsq[i][j];
  if (pairmode == CONP_PAIR_EHGO) {
    bool any = false;
    for (int i = 1; i <= ntypes; ++i) any = any || eta_i[i] != 0.0 || u0_i[i] != 0.0;
    if (!any) {
      pairmode = CONP_PAIR_ETA;
      error->warning(FLERR, "No EHGO settings found, switching back to ETA mode");
    } else {
      eta_ij.assign((size_t) n1 * n1, 0.0); fo_ij.assign((size_t) n1 * n1, 0.0);
      const double s2 = sqrt(2.0) / 1.77245385090551602729;
      for (int i = 1; i <= ntypes; ++i)
        for (int j = 1; j <= i; ++j) {
          double e, f = 0.0;
          if (eta_i[i] != 0.0 && eta_i[j] != 0.0) {
            const double prod = eta_i[i] * eta_i[j];
            e = prod / sqrt(eta_i[i] * eta_i[i] + eta_i[j] * eta_i[j]);
            f = 0.5 * kappa * ((u0_i[i] - s2 * eta_i[i]) + (u0_i[j] - s2 * eta_i[j])) * sqrt(8.0) * e * e * e / (prod * sqrt(prod));
          } else e = eta_i[i] + eta_i[j];
          eta_ij[(size_t) i * n1 + j] = eta_ij[(size_t) j * n1 + i] = e;
          fo_ij[(size_t) i * n1 + j] = fo_ij[(size_t) j * n1 + i] = f;
        }
    }
  }
  check(conp_set_pair(ctx, pairmode, eta, *p_cut_coul, ntypes, cutsq.data(),
                      pairmode == CONP_PAIR_EHGO ? eta_ij.data() : nullptr, pairmode == CONP_PAIR_EHGO ? fo_ij.data() : nullptr,
                      pairmode == CONP_PAIR_EHGO ? u0_i.data() : nullptr, smartlist, smartlist ? is_eletype.data() : nullptr));
  check(conp_set_electrodes(ctx, elenum_all, eleall2tag.data(), eleall_type.data(), eleall_side.data(), eleall_x.data()));
  if (pppm) pppm->attach(ctx);   // its next setup() hands the mesh tables over (KSpaceModule::register_fix, :409)
  init_done = true;
}

/* FixConp::linalg_setup (reference fix_conp.cpp:426-464): A matrix, inverse, unit-voltage charges */
void FixConpB200::linalg_setup()
{
  /* A matrix: a_cal / a_read (reference fix_conp.cpp:438-445, 721-861) */
  if (a_matrix_f == 0) {
    if (outf) fprintf(outf, "A matrix calculating ...\n");   // reference :787
    check(conp_build_A(ctx));
    conp_info info;
    conp_get_info(ctx, &info);
    if (outf) fprintf(outf, "A matrix calculation time  = %g\n", info.setup_build_ms * 1e-3);   // reference :857
  } else {
    /* rank 0 reads the `%20d` tag row + N rows (reference :725-748), permutes to eleall order, broadcasts */
    std::vector<double> full((size_t) elenum_all * elenum_all);
    read_matrix_file(full);  // helper: tokenise, "Too many/Too few entries in A matrix file" errors
    check(conp_load_matrix(ctx, full.data(), a_matrix_f == 2));
  }
  if (matoutflag && a_matrix_f == 0) write_matrix("amatrix", "%20.12f");       // reference :833-849
  double ee = 0.0;
  check(conp_invert_project(ctx, nullneutralflag, zneutrflag, one_electrode_flag, &ee));  // inv + inv_project :932-1067
  evscale = force->qe2f / force->qqr2e;
  if (comm->me == 0 && !one_electrode_flag && a_matrix_f < 2)
    utils::logmesg(lmp, fmt::format("conp output: <e,e> = {:.8g}\n", ee * evscale));
  if (matoutflag && a_matrix_f < 2) write_matrix("inv_a_matrix", "%20.10f");   // reference :960-977
  std::vector<double> qinit;
  if (qinitflag) { qinit.resize(elenum_all); gather_electrode_charges(qinit); }  // reference :1107-1114
  double totsetq = 0.0;
  check(conp_set_unit_voltage(ctx, evscale, qinitflag ? qinit.data() : nullptr, one_electrode_flag, nullneutralflag,
                              zneutrflag, &totsetq));                          // b_setq_cal + get_setq :609-637, 1071-1116
  if (comm->me == 0) utils::logmesg(lmp, fmt::format("conp output: <d,d> = {:.8g}\n", -totsetq));
  setup_done = true;
}

/* reference fix_conp.cpp:382-385 */
void FixConpB200::setup_post_neighbor()
{
  if (!init_done) linalg_init();
  post_neighbor();
}

/* reference fix_conp.cpp:387-391.  kspace->setup() is where `pppm/conp` hands its mesh tables to the library;
   conp_pppm_setup invalidates the per-rank atom data (the exchange buffers depend on the mesh), so the
   post_neighbor hand-over is repeated before the first solve. */
void FixConpB200::setup_pre_force(int vflag)
{
  force->kspace->setup();
  if (!setup_done) linalg_setup();
  post_neighbor();
  /* event-timed stage breakdown of the first solves: feeds the reference's Log-file timer lines */
  conp_stage_times(ctx, 1, nullptr);
  stage_samples = 8;
  pre_force(vflag);
}

/* FixConp::post_neighbor (reference :468-539): static per-atom data of the locally owned atoms */
void FixConpB200::post_neighbor()
{
  check(conp_post_neighbor(ctx, atom->nlocal, atom->q, atom->type, atom->mask, groupbit | jgroupbit));
}

/* FixConp::pre_force (reference :543-573) */
void FixConpB200::pre_force(int)
{
  if (update->ntimestep % everynum) return;
  if (potdiffstr) potdiff = input->variable->compute_equal(potdiffvar);          // reference :1143
  const int mode = pppmflag ? CONP_KSPACE_PPPM : CONP_KSPACE_EWALD;
  /* On several GPUs a rank's step writes straight into its peers' device buffers (positions, b, S.b).
     No rank may start the next solve while another is still in conp_post_force / conp_get_density of
     the previous one; LAMMPS' own per-step collectives normally guarantee that, the barrier makes it
     unconditional (a few microseconds against a step of hundreds). */
  if (comm->nprocs > 1) MPI_Barrier(world);
  const double t1 = MPI_Wtime();
  check(conp_pre_force(ctx, &atom->x[0][0], mode, variant(), potdiff, eleallq.data(), &scalar_output));
  Btime += MPI_Wtime() - t1;
  if (stage_samples > 0 && --stage_samples == 0) {   // back to the CUDA-graph replay
    double st[8];
    conp_stage_times(ctx, 0, st);
    /* st: pack, bin, pair, kspace, gather, exchange, gemv, epilogue [ms] */
    double tot = 0.0;
    for (double v : st) tot += v;
    pair_share = tot > 0.0 ? st[2] / tot : 0.0;
    kspace_share = tot > 0.0 ? (st[3] + st[4]) / tot : 0.0;
  }
  if (pppm) pppm->charges_updated();
  scatter_charges();
  /* reference :553-568: timers written at the last step of the run.  Btime is this rank's wall time inside
     conp_pre_force (b_cal, matvec and epilogue are one call here); the Coulomb / k-space lines are its
     shares measured with CUDA events on the first solves of the run. */
  if (update->laststep == update->ntimestep) {
    double Btime_all = 0.0;
    MPI_Reduce(&Btime, &Btime_all, 1, MPI_DOUBLE, MPI_SUM, 0, world);
    if (outf) {
      Btime = Btime_all / comm->nprocs;
      Ctime = Btime * pair_share;
      Ktime = Btime * kspace_share;
      fprintf(outf, "B vector calculation time = %g\n", Btime);
      fprintf(outf, "Coulomb calculation time = %g\n", Ctime);
      fprintf(outf, "Kspace calculation time = %g\n", Ktime);
      fflush(outf);
    }
  }
}

/* charges of local AND ghost electrode atoms (reference :1153-1158) */
void FixConpB200::scatter_charges()
{
  const int nall = atom->nlocal + atom->nghost;
  int *mask = atom->mask;
  tagint *tag = atom->tag;
  double *q = atom->q;
  for (int i = 0; i < nall; ++i)
    if (mask[i] & (groupbit | jgroupbit)) q[i] = eleallq[tag2eleall[tag[i]]];
}

/* force_cal (reference :1163-1201): self energy into kspace->energy, pair correction via ev_tally */
void FixConpB200::post_force(int)
{
  postforceflag = true;
  std::vector<double> f(3 * (size_t) atom->nlocal);
  double en[8];
  check(conp_post_force(ctx, force->qqrd2e, f.data(), en));
  double **fa = atom->f;
  for (int i = 0; i < atom->nlocal; ++i) { fa[i][0] += f[3 * i]; fa[i][1] += f[3 * i + 1]; fa[i][2] += f[3 * i + 2]; }
  if (force->kspace->energy) force->kspace->energy += en[1];
  if (force->pair->eflag_global) force->pair->eng_coul += en[0] / comm->nprocs;   // ev_tally ecoul, summed by LAMMPS
  if (force->pair->vflag_global) for (int k = 0; k < 6; ++k) force->pair->virial[k] += en[2 + k] / comm->nprocs;
}

void FixConpB200::end_of_step()
{
  if (!postforceflag) post_force(0);
  postforceflag = false;
}

double FixConpB200::compute_scalar() { return scalar_output; }

/* ---- matrix text I/O and qinit helpers ------------------------------------------- */

void FixConpB200::read_matrix_file(std::vector<double> &full)
{
  const size_t n = elenum_all;
  std::vector<int> tags(n);
  if (comm->me == 0) {
    FILE *fp = fopen(a_matrix_file.c_str(), "r");
    if (fp == nullptr) error->one(FLERR, "Invalid fix conp command (Cannot open A matrix file)");
    size_t i = 0;
    double v;
    std::vector<double> file(n * n);
    while (fscanf(fp, "%lf", &v) == 1) {
      if (i < n) tags[i] = (int) v;
      else if (i - n < n * n) file[i - n] = v;
      else error->one(FLERR, "Too many entries in A matrix file");
      ++i;
    }
    fclose(fp);
    if (i != n + n * n) error->one(FLERR, "Too few entries in A matrix file");
    /* the reference adopts the file's tag order as eleall order (:750-759); permuting the matrix into the
       current eleall order gives the same charge per tag */
    std::vector<int> pos(n);
    for (size_t e = 0; e < n; ++e) pos[e] = -1;
    for (size_t k = 0; k < n; ++k) {
      const int e = (tags[k] >= 0 && tags[k] <= maxtag_all) ? tag2eleall[tags[k]] : -1;
      if (e < 0) error->one(FLERR, "A matrix file does not list the electrode atoms of this run");
      pos[e] = (int) k;
    }
    for (size_t a = 0; a < n; ++a)
      for (size_t b = 0; b < n; ++b) full[a * n + b] = file[(size_t) pos[a] * n + pos[b]];
  }
  MPI_Bcast(full.data(), (int) (n * n), MPI_DOUBLE, 0, world);
}

void FixConpB200::write_matrix(const char *name, const char *fmt)
{
  if (comm->nprocs != 1) error->all(FLERR, "fix conp matout: run on one rank (each rank holds only its row block)");
  const size_t n = elenum_all;
  std::vector<double> rows(n * n);
  check(conp_get_matrix(ctx, rows.data()));
  FILE *fp = fopen(name, "w");
  fprintf(fp, " ");
  for (size_t i = 0; i < n; ++i) fprintf(fp, "%20d", eleall2tag[i]);
  fprintf(fp, "\n");
  for (size_t i = 0; i < n; ++i) {
    fprintf(fp, " ");
    for (size_t j = 0; j < n; ++j) fprintf(fp, fmt, rows[i * n + j]);
    fprintf(fp, "\n");
  }
  fclose(fp);
}

void FixConpB200::gather_electrode_charges(std::vector<double> &qall)
{
  std::vector<double> mine(elenum_all, 0.0);
  for (int i = 0; i < atom->nlocal; ++i)
    if (atom->mask[i] & (groupbit | jgroupbit)) mine[tag2eleall[atom->tag[i]]] = atom->q[i];
  MPI_Allreduce(mine.data(), qall.data(), elenum_all, MPI_DOUBLE, MPI_SUM, world);
}
