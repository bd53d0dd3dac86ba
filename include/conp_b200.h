/* conp_b200.h -- C ABI of the B200 (sm_100a) electrode charge solve.
 *
 * This is the drop-in boundary for USER-CONP2's per-step path.  The LAMMPS
 * shim classes (lammps-user-conp2_b200/shim/: FixConp, FixConq, FixCond,
 * PPPMCONP) keep the reference's deck syntax and hook order and forward to
 * these entry points; the Python host mirror (conp_b200/fix_conp.py) and the
 * tests bind the same symbols through ctypes.
 *
 * Each entry point names the reference interface it replaces (file:line
 * relative to the upstream USER-CONP2 tree).
 *
 * Conventions
 *  - every function returns CONP_OK (0) or a CONP_ERR_* code; nothing aborts
 *    or throws across the boundary.  conp_last_error() gives the message; the
 *    shim turns a non-zero code into a collective error->all(FLERR, msg)
 *    (the reference's only error mechanism, e.g. fix_conp.cpp:956).
 *  - pointer arguments are HOST memory borrowed for the duration of the call
 *    (pinned memory recommended) unless the name says `_device`.
 *  - one context per GPU / MPI rank, not thread-safe; all ranks call the
 *    collective entry points (marked [collective]) in the same order.
 *  - all floating point is FP64, all indices are 32-bit int (LAMMPS tagint in
 *    default builds, fix_conp.cpp:474).
 *  - there is NO CPU fallback: without a CUDA device conp_create() fails.
 */
#ifndef CONP_B200_H
#define CONP_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define CONP_ABI_VERSION 2
#define CONP_UNIQUE_ID_BYTES 128

typedef struct conp_ctx conp_ctx;

enum conp_status {
  CONP_OK = 0,
  CONP_ERR_ARG = 1,      /* "Illegal fix conp command ..." class of errors            */
  CONP_ERR_STATE = 2,    /* entry points called out of order                          */
  CONP_ERR_CUDA = 3,     /* CUDA / cuFFT / cuSOLVER failure                           */
  CONP_ERR_NUMERIC = 4,  /* "Inversion failed!"                   fix_conp.cpp:956    */
  CONP_ERR_RANGE = 5,    /* "Out of range atoms - cannot compute PPPM" pppm_conp.cpp:167 */
  CONP_ERR_COMM = 6,     /* NCCL failure                                              */
  CONP_ERR_NOMEM = 7
};

enum { CONP_FF_NORMAL = 0, CONP_FF_FFIELD = 1, CONP_FF_NOSLAB = 2 }; /* fix_conp.cpp:68 */
enum { CONP_PAIR_ETA = 0, CONP_PAIR_EHGO = 1 };                      /* fix_conp.cpp:69 */
enum { CONP_KSPACE_EWALD = 0, CONP_KSPACE_PPPM = 1 };                /* `pppm` keyword, fix_conp.cpp:164, 401-408 */
enum { CONP_VARIANT_CONP = 0, CONP_VARIANT_CONQ = 1, CONP_VARIANT_COND = 2 };

/* sizes and counters, filled by conp_get_info() */
typedef struct conp_info {
  int abi_version;
  int device, rank, nranks;
  int n_ele;              /* elenum_all                                   */
  int row_begin, row_end; /* this GPU's row block of A / S                */
  int n_elyte;            /* charged+uncharged non-electrode atoms, all ranks */
  int kxmax, kymax, kzmax;
  int kcount, kcount_flat, kcount_expand; /* km_ewald.cpp:360-361           */
  int mesh[3], order;
  long long matrix_pitch; /* doubles per stored row of S                  */
  long long launches;     /* kernels launched by this library so far      */
  double setup_build_ms, setup_invert_ms; /* last conp_build_A / conp_invert_project */
  double ee, dd;          /* "<e,e>" and "<d,d>" log values, fix_conp.cpp:1006-1009, 458-461 */
  double totsetq;
  int symmetric_matvec;   /* 1: S is symmetric and the per-step product reads half of it (symv) */
  int reserved0;
  double asymmetry;       /* max|S - S^T| / max|S| measured before use (-1: not measured)      */
} conp_info;

/* ---- lifetime ---------------------------------------------------------- */

int conp_abi_version(void);
/* Number of usable (sm_100) CUDA devices of this node, so that a host without the CUDA runtime
 * headers can map its node-local MPI rank to a device (shim/fix_conp.cpp). */
int conp_device_count(int *count_out);

/* Rank 0 creates the NCCL unique id; the host broadcasts the 128 bytes to
 * the other ranks (MPI_Bcast on LAMMPS' `world`) before conp_create(). */
int conp_get_unique_id(void *id_out /* CONP_UNIQUE_ID_BYTES */);

/* One context per GPU.  unique_id may be NULL when nranks == 1.
 * Replaces the constructors FixConp::FixConp (fix_conp.cpp:79-201) and
 * KSpaceModuleEwald::KSpaceModuleEwald (km_ewald.cpp:41-54). [collective] */
int conp_create(conp_ctx **out, int device, int rank, int nranks, const void *unique_id);
void conp_destroy(conp_ctx *ctx);
const char *conp_last_error(const conp_ctx *ctx); /* ctx may be NULL: last conp_create failure */
int conp_get_info(const conp_ctx *ctx, conp_info *out);

/* ---- one-time setup (FixConp::linalg_init / linalg_setup) --------------- */

/* domain->boxlo/prd/periodicity, force->kspace->slabflag/slab_volfactor and
 * the fix's ffield/noslab keyword (km_ewald.cpp:66-89, fix_conp.cpp:120-133). */
int conp_set_cell(conp_ctx *ctx, const double boxlo[3], const double prd[3], const int periodic[3],
                  int slabflag, double slab_volfactor, int ff_flag);

/* KSpaceModuleEwald::conp_setup (km_ewald.cpp:63-132): k-vector enumeration
 * from LAMMPS' absolute accuracy, q2 = qqrd2e*sum(q^2)/dielectric over all
 * atoms at setup, natoms = atom->natoms.  Needed in both KSpace modes: the A
 * matrix always comes from the Ewald module (pppm_conp.cpp:91-101).
 * lowmem is accepted for deck compatibility (`himem` keyword,
 * fix_conp.cpp:167); it selects a CPU table layout in the reference and does
 * not change results. */
int conp_set_ewald(conp_ctx *ctx, double g_ewald, double accuracy_abs, double q2, long long natoms,
                   int lowmem);

/* Pair-potential data of the Coulomb pair style + the fix (fix_conp.cpp:252-258,
 * 1232-1238, 1300-1305): cutsq is (ntypes+1)^2 row-major; eta_ij/fo_ij/u0_i
 * are the EHGO tables (fix_conp.cpp:1517-1559; may be NULL in ETA mode);
 * is_eletype[ntypes+1] holds the `etypes` keyword (NULL / smartlist 0 = all
 * pairs listed, fix_conp.cpp:304-361). */
int conp_set_pair(conp_ctx *ctx, int pairmode, double eta, double cut_coul, int ntypes,
                  const double *cutsq, const double *eta_ij, const double *fo_ij, const double *u0_i,
                  int smartlist, const int *is_eletype);

/* Global electrode description in eleall order (FixConp::post_neighbor,
 * fix_conp.cpp:468-539): side = electrode_check() (+1 fix group, -1 group2;
 * fix_conp.cpp:599-605); xyz is N x 3.  The library splits the rows into
 * nranks contiguous equal blocks (conp_info.row_begin/row_end). [collective] */
int conp_set_electrodes(conp_ctx *ctx, int n_ele, const int *tag, const int *type, const int *side,
                        const double *xyz);

/* Host PPPM tables for `kspace_style pppm/conp` (PPPMCONP inherits them from
 * LAMMPS PPPM: rho_coeff [order][order] with the k index shifted by -nlower,
 * greensfn on the full mesh [nz][ny][nx] x-fastest, shift/shiftone;
 * pppm_conp.cpp:146-148, 199-203, 245-249).  Also caches the electrode
 * stencils (aaa_map_rho, pppm_conp.cpp:318-344) and allocates the bricks
 * (setup_allocate :346-356). [collective]; conp_post_neighbor must follow. */
int conp_pppm_setup(conp_ctx *ctx, const int mesh[3], int order, const double *rho_coeff,
                    const double *greensfn, double shift, double shiftone);

/* FixConp::a_cal (fix_conp.cpp:777-861) = KSpaceModuleEwald::a_cal
 * (km_ewald.cpp:147-151, 426-666) + self term + alist_coul_cal (:1209-1279)
 * + symmetrisation, for this GPU's row block. [collective] */
int conp_build_A(conp_ctx *ctx);

/* `org <file>` / `inv <file>` keywords (FixConp::a_read, fix_conp.cpp:721-773):
 * full n_ele x n_ele row-major matrix; is_inverse != 0 skips the inversion. */
int conp_load_matrix(conp_ctx *ctx, const double *full_matrix, int is_inverse);
/* `matout` keyword (fix_conp.cpp:833-849, 960-977): this rank's row block of
 * the current matrix (A before conp_invert_project, S after), n_rows x n_ele. */
int conp_get_matrix(conp_ctx *ctx, double *rows_out);

/* FixConp::inv + inv_project (fix_conp.cpp:932-1067): in-place inverse, then
 * (unless one_electrode) the electroneutrality projection(s).
 * ee_out receives "<e,e>" (may be NULL). [collective] */
int conp_invert_project(conp_ctx *ctx, int nullneutral, int zneutr, int one_electrode, double *ee_out);

/* b_setq_cal + cond_setup + get_setq (fix_conp.cpp:609-637, 1071-1116,
 * fix_cond.cpp:46-55): d vector, elesetq = S.d, totsetq; q_init (may be NULL)
 * is the `qinit` snapshot in eleall order; with one_electrode the projection
 * is applied afterwards (:1115). [collective] */
int conp_set_unit_voltage(conp_ctx *ctx, double evscale, const double *q_init, int one_electrode,
                          int nullneutral, int zneutr, double *totsetq_out);

/* ---- per-step path ------------------------------------------------------ */

/* FixConp::post_neighbor (fix_conp.cpp:468-539) for the atoms this rank
 * owns: static per-atom data until the next reneighbouring.  mask/groupbits
 * as in LAMMPS (atom->mask, groupbit | jgroupbit): atoms with
 * (mask & ele_bits) != 0 are electrode atoms and are skipped; mask may be
 * NULL when only non-electrode atoms are passed.  [collective] (the ranks
 * exchange their counts of charged atoms and size the exchange buffers).
 * Call order: after conp_set_electrodes / conp_pppm_setup -- both invalidate
 * the per-rank atom data (the exchange buffers depend on N and on the mesh),
 * so a host that runs its kspace->setup() after setup_post_neighbor, as the
 * reference's hook order does (fix_conp.cpp:382-391), calls this again before
 * the first solve. */
int conp_post_neighbor(conp_ctx *ctx, int nlocal, const double *q, const int *type, const int *mask,
                       int ele_bits);

/* FixConp::pre_force (fix_conp.cpp:543-573) = b_cal (:677-695: k-space part
 * km_ewald.cpp:153-167 or pppm_conp.cpp:269-316, real-space part
 * fix_conp.cpp:1281-1365) + update_charge of the chosen variant
 * (fix_conp.cpp:1120-1161, fix_conq.cpp:41-90, fix_cond.cpp:70-126).
 * x = atom->x[0] (nlocal x 3).  value = dV [V] for conp, right-electrode
 * charge for conq, D for cond.  q_ele_out[n_ele] = new electrode charges in
 * eleall order; scalar_out = the fix's compute_scalar(). [collective]
 * On several GPUs the ranks' steps write into each other's device buffers: a
 * rank must not enter conp_pre_force while another is still inside
 * conp_post_force / conp_get_* of the previous solve (the shim puts an
 * MPI_Barrier in front; inside one solve the library orders itself). */
int conp_pre_force(conp_ctx *ctx, const double *x, int kspace_mode, int variant, double value,
                   double *q_ele_out, double *scalar_out);

/* Same solve with the positions already in device memory (x_device: nlocal x 3
 * on this context's GPU) and no device->host copy: the device-resident timing
 * leg of bench.py.  Results are fetched with conp_get_charges(). [collective] */
int conp_solve_device(conp_ctx *ctx, const double *x_device, int kspace_mode, int variant,
                      double value);
int conp_get_charges(conp_ctx *ctx, double *q_ele_out, double *scalar_out);

/* b vector of the last solve (bbb_all, fix_conp.cpp:694) and its k-space-only
 * part (kspmod->b_cal output incl. slab term). */
int conp_get_b(conp_ctx *ctx, double *b_out, double *b_kspace_out);

/* PPPMCONP::ele_make_rho / make_rho overrides (pppm_conp.cpp:385-450): bricks
 * on the global periodic mesh [nz][ny][nx]; which = 0 electrolyte, 1 electrode,
 * 2 sum (what the host PPPM force pass consumes). */
int conp_get_density(conp_ctx *ctx, int which, double *brick_out);
/* The same hand-off for the part of the mesh ONE host rank owns -- what PPPMCONP::make_rho copies into its
 * density_brick (pppm_conp.cpp:434-450): mesh indices lo[a] .. hi[a] inclusive per axis (x, y, z; inside
 * [0, n)), out[(hi[2]-lo[2]+1)][(hi[1]-lo[1]+1)][(hi[0]-lo[0]+1)], x fastest.  No allocation per call and
 * no full-mesh transfer: planes that hold no charge are written as zeros by the gather kernel, the
 * device->host copy is the region only.  On several GPUs the ranks first exchange their z-slabs of the
 * electrolyte density and add up the electrode density (once per solve). [collective] */
int conp_get_density_region(conp_ctx *ctx, int which, const int lo[3], const int hi[3], double *out);
/* u_brick (pppm_conp.cpp:260-266) for potential probes. */
int conp_get_potential_brick(conp_ctx *ctx, double *brick_out);

/* `compute potential/atom` (compute_potential_atom.cpp:120-182), the diagnostic that reuses the charge solve's
 * mesh and erfc pair kernels.
 * conp_mesh_potential: the mesh sum of PPPMCONP::compute_particle_potential (pppm_conp.cpp:452-484) at n
 * arbitrary positions, u_out[i] = sum_stencil w u_brick, with u_brick the potential of ALL charges of the
 * last solve (electrolyte + updated electrode charges) -- what LAMMPS' PPPM holds after a force pass with
 * per-atom energy; internal units e/Angstrom, no self or slab term.
 * conp_electrode_potential: the complete compute for the electrode atoms (eleall order), in the
 * reference's units (qqr2e/qe2f applied by the caller: the value returned here is in e/Angstrom):
 * pairflag -> compute_pair_potential (:223-318) with the `eta` keyword's Gaussian terms, kspaceflag ->
 * mesh sum, -2 g q/sqrt(pi) self term, +eta q sqrt(2)/sqrt(pi), and slabcorr (:333-358; qsumflag = the
 * `noqsum` keyword off).  In conp runs the result is +-dV/2 (plus a common shift) on the two electrodes:
 * an on-device residual check of the solve.  Both need `kspace_style pppm/conp` (same error message as the
 * reference without it). [collective] */
int conp_mesh_potential(conp_ctx *ctx, int n, const double *xyz, double *u_out);
int conp_electrode_potential(conp_ctx *ctx, int pairflag, int kspaceflag, double eta, int qsumflag,
                             double *phi_out);

/* FixConp::post_force -> force_cal (fix_conp.cpp:1163-1201, 1368-1444).
 * f_out (nlocal x 3, may be NULL) receives the Gaussian-correction force on
 * this rank's non-electrode atoms (zero rows for electrode atoms);
 * energies_out[8] = { pair ecoul tally, self energy added to kspace->energy,
 * virial xx yy zz xy xz yz }, summed over ranks. [collective] */
int conp_post_force(conp_ctx *ctx, double qqrd2e, double *f_out, double *energies_out);

/* ---- instrumentation ------------------------------------------------------ */

void *conp_stream(conp_ctx *ctx); /* cudaStream_t all kernels are launched on */
int conp_sync(conp_ctx *ctx);
/* CUDA events on the context's stream (slots 0..15) */
int conp_timer_record(conp_ctx *ctx, int slot);
int conp_timer_elapsed_ms(conp_ctx *ctx, int slot_begin, int slot_end, float *ms_out);
/* per-stage event timing of the solves since the last reset:
 * out[0..7] = mean ms of {upload/pack, bin, pair, kspace, gather/extract,
 * exchange, gemv, epilogue+electrode spread}; returns number of solves. */
int conp_stage_times(conp_ctx *ctx, int enable, double *out8);
/* stand-alone kernels for roofline measurement (bench.py):
 *   conp_bench_gemv: one q = S.b pass on the resident row block.
 *   conp_bench_dgemm_tflops: cuBLAS DGEMM n^3 ceiling for the Gram. */
int conp_bench_gemv(conp_ctx *ctx, int reps, float *ms_per_rep_out);
/* out[N] = S.v for a host vector v[N] through the same kernel the step uses (the
 * ddot_ loop of get_setq, fix_conp.cpp:1090-1096, applied to any vector): lets a
 * host check the resident matrix, e.g. S.e = 0 after the projection. [collective] */
int conp_matvec(conp_ctx *ctx, const double *v, double *out);
/* Host-only (no GPU needed): the strip decomposition the symmetric matvec uses for the row block
 * [row0, row0+nrows) of an n x n matrix on a device with num_sms SMs.  strips_out receives
 * min(*nstrips_out, max_strips) pairs [begin, end); *slice_len_out is the length of a strip's column
 * slice.  Returns 0 and *nstrips_out = -1 when the symmetric kernel does not apply (n < 64). */
int conp_plan_symv(int n, int row0, int nrows, int num_sms, int max_strips, int *strips_out, int *nstrips_out,
                   int *slice_len_out);
int conp_bench_dgemm_tflops(conp_ctx *ctx, int n, double *tflops_out);
/* Host-only: the row partition of A and S (what fix_conp.cpp:816-823 does with elenum_list/displs): rank's
 * rows [*row_begin, *row_end) of n_ele; every rank but the last ones owns *rows_per_rank rows (a multiple
 * of 16).  conp_set_electrodes uses exactly this rule. */
int conp_row_block(int n_ele, int nranks, int rank, int *row_begin, int *row_end, int *rows_per_rank);
/* Host-only (no GPU needed): the tile decomposition of the owner-computes PPPM spread (replaces the scatter
 * loop of elyte_make_rho, pppm_conp.cpp:172-228) for a mesh / cell grid: geom_out[12] = {tz, ty, tx, ntz, nty,
 * ntx, halo_z, halo_y, halo_x, ncx, ncy, ncz}; run_start_out[ntiles + 1] and runs_out[2 * nruns] list, per
 * tile (z-major, x fastest), the [c0, c1) ranges of sort cells whose charges can reach the tile.  rc = the
 * cell-grid search radius (cells are rc/2 wide), zin_lo/nzi/zs_lo/zs_n = plane pruning and slab of the rank. */
int conp_plan_spread(const int mesh[3], int order, double shift, const double boxlo[3], const double prd[3],
                     const int periodic[3], double slab_volfactor, double rc, int zin_lo, int nzi, int zs_lo,
                     int zs_n, int num_sms, int *geom_out, int *run_start_out, int max_tiles, int *runs_out,
                     int max_runs, int *ntiles_out, int *nruns_out);
/* Host-only: the static candidate list of the real-space pair kernels (replaces the half neighbour list walked by
 * blist_coul_cal / alist_coul_cal, fix_conp.cpp:1225-1365) for n fixed points (the electrode atoms) and search
 * radius rc: sort-cell grid nc_out[3] (cells rc/2 wide, x fastest), and per point the x-contiguous runs of cells,
 * seen through a periodic image shift, that the cut-off sphere can reach: run_start_out[n + 1] and
 * runs_out[5 * nruns] = {c0, c1, sx, sy, sz} (cells [c0, c1), image shift in box lengths).  Points of a
 * non-periodic axis outside the box are binned into the edge cells. */
/* Host-only: the work plan of the z-sweep PPPM spread (the second, mesh-aligned sort + sliding window of planes
 * that replaces the scatter loop of elyte_make_rho, pppm_conp.cpp:172-228, at large charge counts) for a rank's
 * slab [zs_lo, zs_lo + zs_n) of the nzi compact planes: geom_out[8] = {usable, columns along x, columns along y
 * (a column = 8 rows x 32 mesh columns), first binned origin plane, number of binned origin planes, origin planes
 * wrap (periodic z), bins, grid size}; items_out[3 * nitems] = {column, first, end slab plane of the segment}. */
int conp_plan_sweep(const int mesh[3], int order, int nzi, int zs_lo, int zs_n, int num_sms, int *geom_out,
                    int *items_out, int max_items, int *nitems_out);
/* Host-only: the work plan of the windowed z-convolution that replaces the z-part of the FFT -> greensfn -> FFT
 * chain of elyte_poisson (pppm_conp.cpp:230-267).  krad[ncol] = per (kx,ky) column the circular distance beyond
 * which the tabulated kernel is dropped; zout[nzo] = mesh planes the electrode stencils read; the rank's slab of
 * input planes is compact planes [zs_lo, zs_lo + nzl) of the nzi planes that can hold charge (mesh plane =
 * zin_lo + compact plane, mod nz).  Columns are handled in groups of 8: a "narrow" group stages the slab planes
 * within its radius of an output plane in shared memory, the other columns ("wide") read all planes.
 * groups_out[32 * ngroups] = {c0, rblock, nint, np, lo[8], hi[8], base[8], pad[4]} per narrow group,
 * wide_out[nwide] the wide columns, aout_out[nzo] the output planes in compact coordinates,
 * caps_out[2] = {rcap, npcap} (largest radius / staged planes the narrow path is sized for). */
int conp_plan_zconv(int ncol, int nz, int nzi, int zs_lo, int nzl, int zin_lo, const int *krad, int nzo,
                    const int *zout, int real_kernel, int *groups_out, int max_groups, int *wide_out, int max_wide,
                    int *aout_out, int *caps_out, int *ngroups_out, int *nwide_out);
int conp_plan_pair_runs(const double boxlo[3], const double prd[3], const int periodic[3], double rc, int n,
                        const double *xyz, int *nc_out, int *run_start_out, int *runs_out, int max_runs,
                        int *nruns_out);

#ifdef __cplusplus
}
#endif
#endif /* CONP_B200_H */
