// C-ABI entry points (include/conp_b200.h) and the host-side orchestration of
// the setup and per-step pipelines.  No numerical work happens on the host
// besides O(N) bookkeeping; there is no CPU fallback.
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

using namespace conp;

namespace {
constexpr double MY_PI = 3.14159265358979323846;
constexpr double MY_PIS = 1.77245385090551602729;
constexpr double ERFC_MAX = 5.8;
constexpr int NSTAGE = 8;
std::string g_create_error;

size_t round_up(size_t a, size_t b) { return (a + b - 1) / b * b; }
}  // namespace

struct conp_ctx {
  int device = 0, rank = 0, nranks = 1, num_sms = 148;
  cudaStream_t stream = nullptr;
  // side stream of the step's fork/join: buffer clears and the real-space pair kernel run beside the
  // sort -> spread -> FFT chain (they meet again at the b gather); captured into the same CUDA graph
  cudaStream_t side = nullptr;
  cudaEvent_t ev_begin = nullptr, ev_cleared = nullptr, ev_sorted = nullptr, ev_pair = nullptr;
  bool overlap = true;
  Comm *comm = nullptr;
  // direct NVLink exchanges (multi-GPU): b, S.b, the output-plane spectra and the packed charges live
  // in an IPC-mapped arena that every peer writes into; nullptr => NCCL collectives
  PeerArena *p2p = nullptr;
  size_t p2p_bytes = 0, off_b = 0, off_sb = 0, off_uhat = 0, off_stage = 0, off_packed = 0, off_parts = 0;
  // routed position exchange (pack_route_kernel): inbox types / sender-local indices / per-sender counts
  size_t off_ptype = 0, off_psrc = 0, off_rcnt = 0;
  bool route_allowed = true, routed = false;
  DevBuf<unsigned char> d_rel_all;  // [nranks][ncells] relevance masks of every rank
  DevBuf<int> d_sendcnt, d_psrc;
  DevBuf<PosQ> d_own;               // this rank's wrapped charges (all of them)
  std::string err;
  long long launches = 0;

  bool have_cell = false, have_ewald = false, have_pair = false, have_ele = false, have_pppm = false;
  bool have_A = false, inverted = false, have_setq = false, have_atoms = false, solved = false;

  // cell ---------------------------------------------------------------
  double boxlo[3] = {0, 0, 0}, prd[3] = {1, 1, 1};
  int periodic[3] = {1, 1, 1};
  int slabflag = 0, ff_flag = 0;
  double slab_volfactor = 1.0;

  // ewald ----------------------------------------------------------------
  double g_ewald = 0;
  EwaldHost ew;
  DevBuf<short> d_kx, d_ky, d_kz;
  DevBuf<double> d_ug, d_sfac;
  DevBuf<double2> d_etab, d_jtab;
  // tensor-core form of the Ewald sums: -1 automatic (large M*K), 0 never, 1 always (CONP_EWALD_GEMM)
  EwaldGemm eg;
  int eg_mode = -1;
  bool eg_valid = false;

  // pair -----------------------------------------------------------------
  int pairmode = 0, ntypes = 0, smartlist = 0;
  double eta = 0, cut_coul = 0;
  std::vector<double> h_cutsq, h_u0;
  DevBuf<double> d_cuteff_b, d_cuteff_a, d_cutsq_listed, d_eta_ij, d_fo_ij, d_u0;
  double rc_b = 0, rc_a = 0, rc_f = 0;

  // electrodes -------------------------------------------------------------
  int N = 0, rpr = 0, r0 = 0, r1 = 0, ncols_pad = 0;
  size_t pitch = 0, vlen = 0;
  std::vector<int> h_tag, h_type, h_side;
  std::vector<double> h_xyz;
  DevBuf<double> d_ex, d_ey, d_ez, d_exyz;
  DevBuf<int> d_etype, d_eside;

  // matrix / vectors ---------------------------------------------------------
  DevBuf<double> d_mat, d_fullS;
  // symmetric S: the matvec reads half of the matrix (symv_tma_kernel in gemv.cu)
  bool sym = false, sym_allowed = true, sym_plan_ok = false;
  double asym_rel = -1.0;  // measured max |S - S^T| / max |S| before symmetrising (-1: not measured)
  SymvPlan sy;
  DevBuf<double> d_rowpart, d_colpart;
  DevBuf<int2> d_strips;
  DevBuf<double> d_b, d_bk, d_breal, d_sb, d_q, d_setq, d_dvec, d_setz, d_qinit, d_scal;
  bool have_qinit = false;
  double totsetq = 0, vmult = 0, evscale = 0, ee = 0, dd = 0;
  int one_electrode = 0;
  // electroneutrality polish of the epilogue (see charge_epilogue): on when the projection is
  bool neutral_polish = false, projected_here = false;
  bool A_half_band = false;  // several GPUs: d_mat holds A's cyclic half band only until the blocks are assembled
  int n_left = 0;
  double sum_setz = 0;
  double build_ms = 0, invert_ms = 0;

  // atoms ------------------------------------------------------------------------
  int nlocal = 0, m_local = 0, m_total = 0;
  double qsum_elyte = 0;  // sum of the non-electrode charges, all ranks (compute potential/atom, slab term)
  int mpad = 0, m_slots = 0;  // multi-GPU: every rank's packed block has mpad slots (last one carries sum q z)
  std::vector<int> h_idx, m_counts, m_offsets;
  DevBuf<int> d_mcounts;
  DevBuf<double> d_xraw, d_qraw;
  DevBuf<int> d_typeraw, d_idx;
  DevBuf<PosQ> d_packed, d_sorted;
  DevBuf<float4> d_sortedf;
  DevBuf<int> d_ptype, d_stype, d_ssrc, d_cellof, d_slot, d_cellcount, d_cellstart, d_nearlist, d_nearcount;
  DevBuf<double> d_fpacked;
  // static electrode cell structures for the real-space kernels (own rows)
  CellGrid grid_b;
  bool static_cells = false;
  DevBuf<EPos> d_esorted;
  DevBuf<int> d_ecellstart, d_runstart;
  DevBuf<PairRun> d_runs;
  DevBuf<unsigned char> d_nearmask;
  // several GPUs, PPPM, non-periodic z: cells whose charges this rank needs at all (reachable from its
  // rows, or in the z-range of its slab of input planes); the others are not sorted
  DevBuf<unsigned char> d_relevant;
  bool have_relevant = false;

  // pppm ---------------------------------------------------------------------------
  PPPMGeom pg;
  size_t ngrid = 0, nhalf = 0, plane = 0, ncol = 0;
  std::vector<double> h_ghalf;          // symmetrised greensfn/(nx ny nz), half spectrum (full-mesh path on demand)
  std::vector<int> h_zout;              // output planes (sorted)
  std::vector<double> h_rho;            // rho_coeff [order][order] (kernel argument of the tile spread)
  DevBuf<double> d_rho, d_brick, d_ubrick, d_ebrick, d_weights, d_Kr;
  // force-pass hand-off (conp_get_density_region): gathered bricks on several GPUs, staging of the region
  DevBuf<double> d_brick_all, d_ebrick_all, d_region;
  bool density_gathered = false;
  DevBuf<int> d_part2grid, d_widx, d_poff, d_flag, d_zmap, d_zout, d_krad, d_zc_wide, d_zc_aout;
  DevBuf<ZconvGroup> d_zc_narrow;
  ZconvPlan zplan;
  // owner-computes spread (pppm.cu): tiles of the slab + candidate cell runs (CONP_SPREAD selects the kernel)
  SpreadPlan splan;
  DevBuf<int> d_sp_runstart, d_sp_counter;
  DevBuf<int2> d_sp_runs;
  DevBuf<int4> d_sp_origin;
  DevBuf<double> d_sp_weights;
  // z-sweep spread
  SweepPlan swplan;
  DevBuf<int> d_sw_binof, d_sw_slot, d_sw_count, d_sw_start;
  DevBuf<double> d_sw_records;
  DevBuf<int4> d_sw_items;
  bool spread_atomic = false;
  int spread_mode = -1;  // CONP_SPREAD = atomic | smem | mma | sweep (default: by size, see conp_post_neighbor)
  DevBuf<double> d_pw;
  DevBuf<cufftDoubleComplex> d_rhat, d_uhat, d_Kc;
  cufftHandle plan_f = 0, plan_b = 0;
  bool plans = false, k_real = true;

  cusolverDnHandle_t solver = nullptr;
  cublasHandle_t blas = nullptr;

  // epilogue scratch + CUDA-graph replay of the step ----------------------------------
  DevBuf<double> d_partials;
  DevBuf<unsigned int> d_counter;
  PinnedBuf<double> h_value;  // ring of staged `value` arguments
  int value_slot = 0;
  bool use_graph = true;
  cudaGraphExec_t graph_exec[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
  long long graph_launches[2][3] = {{0, 0, 0}, {0, 0, 0}};
  int eager_runs[2][3] = {{0, 0, 0}, {0, 0, 0}};

  // timing ----------------------------------------------------------------------------
  cudaEvent_t ev[16];
  cudaEvent_t sev[NSTAGE + 1];
  cudaEvent_t kev[6];           // CONP_DEBUG: spread | fft | zconv | all-reduce | ifft inside the k-space stage
  double kev_ms[5] = {0, 0, 0, 0, 0};
  bool debug = false;             // CONP_DEBUG: extra timers and plan printouts on stderr
  bool signal_in_kernel = false;  // CONP_SIGNAL_IN_KERNEL=1: producers raise the flags themselves (per-block fences)
  bool fused_signal = true;     // CONP_FUSED_SIGNAL=0: stand-alone one-block signal kernels behind the producers
  bool uhat_nccl = false;       // CONP_UHAT_NCCL=1: NCCL all-reduce for the spectra even on the peer-to-peer path
  bool spread_unsorted = true;  // CONP_SPREAD_UNSORTED=0: the red.global spread reads the cell-sorted charges
  bool stage_timing = false;
  // CONP_TRACE=1 (debugging): one-thread kernels between the stages of the captured step add up %globaltimer
  // offsets from the step's start -- the only way to see the critical path inside a graph replay, where the
  // eager per-stage events (launch-rate bound for kernels of a few us) say little.  Each stamp costs ~2 us.
  bool trace = false;
  DevBuf<unsigned long long> d_trace;  // [0] start of the current step, [1 + i] sum of offsets of mark i, [40] steps
  double stage_ms[NSTAGE] = {0, 0, 0, 0, 0, 0, 0, 0};
  int stage_n = 0;

  // scalars in d_scal: [0] scalar_output [1] potdiff [2] qz_sum [3] projection total [4..11] energies
  // [12] value (dV | QR | D) of the current solve [13] mean residual of S.b taken off every charge
  double *scal(int i) { return d_scal.p + i; }
};

namespace {

template <class F>
int guard(conp_ctx *c, F &&f) {
  if (!c) return CONP_ERR_ARG;
  try {
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) CONP_THROW(CONP_ERR_CUDA, "cudaSetDevice(%d): %s", c->device, cudaGetErrorString(e));
    f();
    return CONP_OK;
  } catch (const Error &e) {
    c->err = e.msg;
    return e.code;
  } catch (const std::exception &e) {
    c->err = e.what();
    return CONP_ERR_CUDA;
  }
}

void need(bool cond, const char *what) {
  if (!cond) CONP_THROW(CONP_ERR_STATE, "%s", what);
}

PairTables pair_tables(conp_ctx *c, const double *cuteff) {
  PairTables pt;
  pt.ntypes = c->ntypes;
  pt.pairmode = c->pairmode;
  pt.g_ewald = c->g_ewald;
  pt.eta = c->eta;
  pt.cuteff = cuteff;
  pt.eta_ij = c->d_eta_ij.p;
  pt.fo_ij = c->d_fo_ij.p;
  return pt;
}

double slab_pref(const conp_ctx *c) {
  if (!c->slabflag) return 0.0;
  const double volume = c->prd[0] * c->prd[1] * c->prd[2] * c->slab_volfactor;
  return 4.0 * MY_PI / volume;  // km_ewald.cpp:839, pppm_conp.cpp:307
}

// electrodes never move: sort this rank's rows into the cell grid once and mark
// the cells from which a point charge can reach them
void ensure_static_cells(conp_ctx *c) {
  if (c->static_cells) return;
  c->grid_b = make_cell_grid(c->boxlo, c->prd, c->periodic, c->rc_b > 0 ? c->rc_b : 1.0);
  std::vector<EPos> sorted;
  std::vector<int> cs;
  std::vector<unsigned char> mask;
  build_electrode_cells(c->grid_b, c->r0, c->r1, c->h_xyz.data(), c->h_type.data(), sorted, cs);
  build_near_mask(c->grid_b, c->r0, c->r1, c->h_xyz.data(), mask);
  if (sorted.empty()) sorted.resize(1);
  std::vector<int> run_start;
  std::vector<PairRun> runs;
  build_pair_runs(c->grid_b, c->r0, c->r1, c->h_xyz.data(), run_start, runs);
  c->d_runstart.upload(run_start, c->stream);
  c->d_runs.upload(runs, c->stream);
  c->d_esorted.upload(sorted, c->stream);
  c->d_ecellstart.upload(cs, c->stream);
  c->d_nearmask.upload(mask, c->stream);
  c->d_nearcount.zero(1, c->stream);
  c->have_relevant = false;
  std::vector<int> rs;
  std::vector<int2> rr;
  if (c->have_pppm) {
    c->splan.use_mma = c->spread_mode != 1;
    plan_pppm_spread_tiles(c->pg, c->grid_b, c->num_sms, rs, rr, c->splan);
    if (rr.empty()) rr.push_back(make_int2(0, 0));
    c->d_sp_runstart.upload(rs, c->stream);
    c->d_sp_runs.upload(rr, c->stream);
    c->d_sp_counter.zero(8, c->stream);
    c->splan.run_start = c->d_sp_runstart.p;
    c->splan.runs = c->d_sp_runs.p;
    c->splan.counter = c->d_sp_counter.p;
    {
      std::vector<int4> items;
      plan_pppm_sweep(c->pg, c->num_sms, items, c->swplan);
      if (c->swplan.usable) {
        c->d_sw_items.upload(items, c->stream);
        c->d_sw_count.zero((size_t)c->swplan.nbins + 8, c->stream);
        c->d_sw_start.zero((size_t)c->swplan.nbins + 8, c->stream);
        c->swplan.items = c->d_sw_items.p;
        c->swplan.bin_count = c->d_sw_count.p;
        c->swplan.bin_start = c->d_sw_start.p;
        c->swplan.counter = c->d_sp_counter.p + 1;
        if (c->debug)
          fprintf(stderr, "[conp] rank %d sweep spread: %d x %d columns, origin planes %d..+%d, %d bins, %d items, grid %d\n",
                  c->rank, c->swplan.ncoly, c->swplan.ncolx, c->swplan.pz_lo, c->swplan.npz, c->swplan.nbins,
                  c->swplan.nitems, c->swplan.grid);
      }
    }
    if (c->debug)
      fprintf(stderr, "[conp] rank %d spread tiles: %d x %d x %d tiles of %d x %d x %d (z,y,x), %zu cell runs, "
              "%zu B smem per warp, grid %d\n", c->rank, c->splan.ntz, c->splan.nty, c->splan.ntx, c->splan.tz,
              c->splan.ty, c->splan.tx, rr.size(), c->splan.smem, c->splan.grid);
  }
  // Several GPUs: the cells whose charges this rank reads at all -- within reach of its electrode rows (pair
  // kernels), or in a candidate cell run of one of its spread tiles (its slab of PPPM planes).  Static.  The
  // receiver-side filter of the all-gather path and the sender-side routing both use it.
  if (c->nranks > 1 && getenv("CONP_SORT_ALL") == nullptr) {
    std::vector<unsigned char> rel((size_t)c->grid_b.ncells, 0);
    for (const PairRun &r : runs)
      for (int cell = r.c0; cell < r.c1; ++cell) rel[cell] = 1;
    if (c->have_pppm)
      for (int k = 0; k < c->splan.ntiles && k + 1 < (int)rs.size(); ++k)
        for (int r = rs[k]; r < rs[k + 1]; ++r)
          for (int cell = rr[r].x; cell < rr[r].y; ++cell) rel[cell] = 1;
    c->d_relevant.upload(rel, c->stream);
    c->have_relevant = true;
    c->d_rel_all.reserve((size_t)c->nranks * c->grid_b.ncells);
    comm_allgather(c->comm, c->d_relevant.p, c->d_rel_all.p, (size_t)c->grid_b.ncells, c->stream);  // collective
  }
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  c->static_cells = true;
}

void drop_graphs(conp_ctx *c) {
  for (auto &row : c->graph_exec)
    for (auto &g : row)
      if (g) { cudaGraphExecDestroy(g); g = nullptr; }
  for (auto &row : c->eager_runs)
    for (auto &n : row) n = 0;
}

ChargeEpilogue make_epilogue(conp_ctx *c, int variant, bool fused) {
  ChargeEpilogue ep;
  std::memset(&ep, 0, sizeof(ep));
  ep.enabled = 1;
  ep.variant = variant;
  ep.n = c->N;
  ep.row_offset = (fused && !c->sym) ? c->r0 : 0;
  ep.one_electrode = c->one_electrode;
  ep.neutral = c->neutral_polish ? 1 : 0;
  ep.n_left = c->n_left;
  ep.sum_setz = c->sum_setz;
  ep.totsetq = c->totsetq;
  ep.lz = c->prd[2];
  ep.vmult = c->vmult;
  ep.value = c->scal(12);
  ep.dipole = c->scal(2);
  ep.side = c->d_eside.p;
  ep.setz = c->d_setz.p;
  ep.setq = c->d_setq.p;
  ep.qinit = c->have_qinit ? c->d_qinit.p : nullptr;
  ep.sb = c->d_sb.p;
  ep.q_out = c->d_q.p;
  ep.scalar_out = c->scal(0);
  ep.partials = c->d_partials.p;
  ep.counter = c->d_counter.p;
  return ep;
}

// Detach every view and free the arena.  Collective (p2p_destroy closes the peers' mappings): used by the
// setup entry points that change the size of an exchanged buffer; conp_post_neighbor builds the next one.
void drop_p2p(conp_ctx *c) {
  if (!c->p2p) return;
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  c->d_b.release(); c->d_sb.release(); c->d_uhat.release(); c->d_packed.release();
  if (c->routed) { c->d_ptype.release(); c->d_psrc.release(); }
  c->routed = false;
  p2p_destroy(c->p2p);
  c->p2p = nullptr;
  c->p2p_bytes = 0;
}

// (Re)build the peer-to-peer arena when the exchanged buffers changed size.  Collective: every rank
// sees the same sizes, so every rank takes the same branch.
void ensure_p2p(conp_ctx *c) {
  if (c->nranks == 1) return;
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t n_u = c->have_pppm ? 2 * (size_t)c->pg.nzo * c->ncol : 0;  // doubles in the spectra
  const size_t slice = (std::max(n_u, c->vlen) + c->nranks - 1) / c->nranks;
  const size_t b_bytes = up(sizeof(double) * c->vlen);
  const size_t u_bytes = up(sizeof(double) * std::max<size_t>(n_u, 2));
  const size_t st_bytes = up(sizeof(double) * std::max<size_t>(slice * c->nranks, 2));
  const size_t pk_bytes = up(sizeof(PosQ) * (size_t)std::max(c->m_slots, 1));
  const size_t pt_bytes = up(sizeof(double) * c->vlen * c->nranks);  // one partial S.b per rank
  const size_t ti_bytes = up(sizeof(int) * (size_t)std::max(c->m_slots, 1));  // inbox types / source indices
  const size_t need = 2 * b_bytes + u_bytes + st_bytes + pk_bytes + pt_bytes + 2 * ti_bytes + 256;
  if (c->p2p && need == c->p2p_bytes) return;
  drop_p2p(c);
  // the views about to be attached may still own private memory from a single-context phase
  c->p2p = p2p_create(c->comm, need, c->stream);
  c->p2p_bytes = need;
  if (!c->p2p) return;  // IPC not available: stay on NCCL
  c->off_b = 0;
  c->off_sb = b_bytes;
  c->off_uhat = 2 * b_bytes;
  c->off_stage = c->off_uhat + u_bytes;
  c->off_packed = c->off_stage + st_bytes;
  c->off_parts = c->off_packed + pk_bytes;
  c->off_ptype = c->off_parts + pt_bytes;
  c->off_psrc = c->off_ptype + ti_bytes;
  c->off_rcnt = c->off_psrc + ti_bytes;
  char *base = p2p_local(c->p2p);
  c->d_b.attach((double *)(base + c->off_b), c->vlen);
  c->d_sb.attach((double *)(base + c->off_sb), c->vlen);
  if (c->have_pppm) c->d_uhat.attach((cufftDoubleComplex *)(base + c->off_uhat), n_u / 2);
  c->d_packed.attach((PosQ *)(base + c->off_packed), (size_t)std::max(c->m_slots, 1));
  c->routed = c->route_allowed;
  if (c->routed) {
    c->d_ptype.attach((int *)(base + c->off_ptype), (size_t)std::max(c->m_slots, 1));
    c->d_psrc.attach((int *)(base + c->off_psrc), (size_t)std::max(c->m_slots, 1));
  }
}

__global__ void stamp_kernel(unsigned long long *tr, int i) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  if (i == 0) { tr[0] = t; tr[40]++; }
  tr[1 + i] += t - tr[0];
}

void trace_mark(conp_ctx *c, int i) {
  if (c->trace && !c->stage_timing) stamp_kernel<<<1, 1, 0, c->stream>>>(c->d_trace.p, i);
}

void stage_mark(conp_ctx *c, int i) {
  if (c->stage_timing) CUDA_CHECK(cudaEventRecord(c->sev[i], c->stream));
  trace_mark(c, i);
}

// q-side matvec out = S.b.  `out` is a full-length (vlen) vector: this rank's rows on return of the
// GEMV branch, the complete product (one GPU) or this rank's partial sum (to be all-reduced) on return
// of the symmetric branch.  The epilogue can only be fused on one GPU.
int enqueue_matvec(conp_ctx *c, cudaStream_t s, const double *b, double *out, const ChargeEpilogue *ep) {
  const int nr = c->r1 - c->r0;
  if (c->sym)
    return launch_symv(s, c->d_mat.p, c->pitch, c->N, c->r0, nr, b, c->sy, c->d_rowpart.p, c->d_colpart.p, out,
                       (int)c->vlen, ep, PeerSync(), PeerSync(), 0);
  return launch_gemv(s, c->d_mat.p, c->pitch, nr, c->ncols_pad, b, out + c->r0, c->num_sms, ep);
}

// Decide whether the symmetric matvec may be used for the matrix `full` (N x N, ld = N, the complete
// inverse on this GPU).  own_inverse: the matrix was inverted here from a symmetric A, so an asymmetry
// at the rounding level (x condition number) is noise and is averaged away; a matrix read from a file
// is used exactly as given, and takes the symmetric path only if it is symmetric to the last bit.
void decide_symmetry(conp_ctx *c, double *full, bool own_inverse) {
  c->sym = false;
  if (!c->sym_allowed || !c->sym_plan_ok) return;
  cudaStream_t s = c->stream;
  DevBuf<double> m;
  m.zero(2, s);
  c->launches += launch_asymmetry(s, c->N, full, c->N, m.p);
  double h[2] = {0, 0};
  CUDA_CHECK(cudaMemcpyAsync(h, m.p, sizeof(h), cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  c->asym_rel = h[1] > 0 ? h[0] / h[1] : 0.0;
  double bad = own_inverse ? (c->asym_rel <= 1e-9 ? 0.0 : 1.0) : (h[0] == 0.0 ? 0.0 : 1.0);
  if (c->nranks > 1) {  // agree across ranks
    CUDA_CHECK(cudaMemcpyAsync(m.p, &bad, sizeof(double), cudaMemcpyHostToDevice, s));
    comm_allreduce_sum_f64(c->comm, m.p, 1, s);
    CUDA_CHECK(cudaMemcpyAsync(&bad, m.p, sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
  }
  if (bad != 0.0) return;
  if (own_inverse && h[0] != 0.0) c->launches += launch_symmetrise(s, c->N, full, c->N);
  c->sym = true;
}

// Ewald mode: the GEMM form pays off once the direct O(M K) sum is more than a few launches' worth
bool use_ewald_gemm(const conp_ctx *c) {
  if (c->eg_mode >= 0) return c->eg_mode == 1;
  return (long long)c->m_total * c->ew.kcount >= 20000000LL;
}

// operands and workspaces of the GEMM form (outside graph capture: allocates)
void ensure_ewald_gemm(conp_ctx *c) {
  if (c->eg_valid) return;
  // every rank sums its own charges (sfac_reduce, km_ewald.cpp:782-786)
  ewald_gemm_plan(c->eg, c->ew, std::max(c->m_local, 1), c->r1 - c->r0, c->num_sms, c->stream);
  c->launches += ewald_gemm_electrodes(c->stream, c->eg, c->ew, c->r0, c->r1, c->d_etab.p);
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  c->eg_valid = true;
}

// --------------------------------------------------------------------------
// the per-step pipeline (device side, asynchronous on c->stream).  Inputs:
// positions in c->d_xraw, the variant's value in scal(12).  Everything here is
// stream-ordered and capturable into a CUDA graph.
// --------------------------------------------------------------------------
void enqueue_step(conp_ctx *c, int kspace_mode, int variant) {
  cudaStream_t s = c->stream;
  const int nr = c->r1 - c->r0;
  const bool multi = c->nranks > 1;
  const double *x_dev = c->d_xraw.p;

  // fork: `t` runs beside the main stream (t == s when the overlap is off or stages are being timed)
  const bool fork = c->overlap && !c->stage_timing;
  cudaStream_t t = fork ? c->side : s;
  stage_mark(c, 0);
  if (fork) {
    CUDA_CHECK(cudaEventRecord(c->ev_begin, s));
    CUDA_CHECK(cudaStreamWaitEvent(t, c->ev_begin, 0));
  }
  if (kspace_mode == CONP_KSPACE_PPPM) {  // clears of the step's bricks, off the critical path
    if (c->spread_atomic)  // the owner-computes kernels store every point of the brick themselves
      c->launches += launch_fill_zero(t, c->d_brick.p, (size_t)std::max(c->pg.zs_n, 1) * c->plane);
    CUDA_CHECK(cudaMemsetAsync(c->d_flag.p, 0, sizeof(int), t));
    c->launches += launch_fill_zero(t, c->d_ebrick.p, (size_t)c->pg.nzo * c->plane);
    if (fork) CUDA_CHECK(cudaEventRecord(c->ev_cleared, t));
  }
  // ---- counting sort of the point charges: pack (+histogram, sum q z), scan, scatter ----
  CUDA_CHECK(cudaMemsetAsync(c->scal(2), 0, sizeof(double), s));
  const CellGrid &g = c->grid_b;
  CUDA_CHECK(cudaMemsetAsync(c->d_cellcount.p, 0, sizeof(int) * ((size_t)g.ncells + 8), s));
  PosQ *packed_local = c->d_packed.p + c->m_offsets[c->rank];
  int *ptype_local = c->d_ptype.p + c->m_offsets[c->rank];
  // several GPUs, peer-to-peer path: the exchanges are done by the producing / consuming kernels
  const bool fused = multi && c->p2p != nullptr;
  // producer side of a fused exchange: store into the peers; the flags go up in the kernel's last block
  // or, by default, from a one-block p2p_signal kernel behind it (see PeerSync::signal_in_kernel)
  auto producer = [&](int chan) {
    PeerSync ps = p2p_sync(c->p2p, chan);
    ps.signal_in_kernel = c->signal_in_kernel ? 1 : 0;
    return ps;
  };
  const bool late_signal = fused && !c->signal_in_kernel;
  // ... or, by default, from block 0 of the consuming kernel (PeerSync::raise_first): one launch less per exchange
  const bool raise_in_consumer = late_signal && c->fused_signal;
  PeerSync ps_pos = fused ? p2p_sync(c->p2p, 0) : PeerSync();
  const size_t off_qz = c->off_packed + sizeof(PosQ) * (size_t)(c->m_offsets[c->rank] + c->mpad - 1);
  // routed position exchange: a charge goes only to the ranks whose relevance mask covers its cell
  const bool routed = fused && c->routed && c->have_relevant;
  if (!multi) {
    c->launches += launch_pack_count(s, g, c->m_local, x_dev, c->d_idx.p, c->d_qraw.p, c->d_typeraw.p, packed_local,
                                     ptype_local, c->d_cellof.p, c->d_slot.p, c->d_cellcount.p, c->scal(2),
                                     PeerSync(), 0, 0);
  } else if (routed) {
    CUDA_CHECK(cudaMemsetAsync(c->d_sendcnt.p, 0, sizeof(int) * c->nranks, s));
    c->launches += launch_pack_route(s, g, c->m_local, x_dev, c->d_idx.p, c->d_qraw.p, c->d_typeraw.p, c->d_own.p,
                                     c->rank, c->nranks, c->mpad, c->d_rel_all.p, p2p_sync(c->p2p, 0),
                                     c->off_packed, c->off_ptype, c->off_psrc, c->d_sendcnt.p, c->scal(2));
    // per-receiver counts, this rank's sum(q z) (padding slot of its inbox block on every rank), flags
    if (raise_in_consumer) {
      ps_pos.raise_first = 1;
      ps_pos.value_off = p2p_arena_offset(off_qz); ps_pos.value = c->scal(2);
      ps_pos.count_off = p2p_arena_offset(c->off_rcnt); ps_pos.counts = c->d_sendcnt.p;
    } else {
      c->launches += p2p_signal(c->p2p, 0, s, off_qz, c->scal(2), c->off_rcnt, c->d_sendcnt.p);
    }
  } else {
    // every rank's positions go to every rank; the sum(q z) partial rides in the block's last (padding)
    // slot.  Types and charges are static between reneighbourings and were gathered in conp_post_neighbor.
    c->launches += launch_pack_count(s, g, c->m_local, x_dev, c->d_idx.p, c->d_qraw.p, c->d_typeraw.p, packed_local,
                                     ptype_local, nullptr, nullptr, nullptr, c->scal(2),
                                     fused ? producer(0) : PeerSync(),
                                     c->off_packed + sizeof(PosQ) * (size_t)c->m_offsets[c->rank], c->mpad);
    if (raise_in_consumer) {  // + this rank's sum(q z) into the block's padding slot on every rank
      ps_pos.raise_first = 1;
      ps_pos.value_off = p2p_arena_offset(off_qz); ps_pos.value = c->scal(2);
    } else if (late_signal) {
      c->launches += p2p_signal(c->p2p, 0, s, off_qz, c->scal(2));
    }
  }
  stage_mark(c, 1);
  if (multi) {
    if (!fused) {
      CUDA_CHECK(cudaMemcpyAsync(&packed_local[c->mpad - 1].x, c->scal(2), sizeof(double),
                                 cudaMemcpyDeviceToDevice, s));
      comm_allgather(c->comm, packed_local, c->d_packed.p, sizeof(PosQ) * (size_t)c->mpad, s);
    }
    // PPPM mode: only the charges this rank can use (near its rows, or in its slab) are sorted
    // all-gather path: only the charges this rank can use are sorted (Ewald mode on that path sums the
    // structure factors of the rank's own block, which is not sorted either)
    const unsigned char *relevant = (!routed && c->have_relevant) ? c->d_relevant.p : nullptr;
    const int *counts = routed ? (const int *)(p2p_local(c->p2p) + c->off_rcnt) : c->d_mcounts.p;
    c->launches += launch_bin_positions(s, g, c->m_slots, c->mpad, counts, c->d_packed.p, c->d_cellof.p,
                                        c->d_slot.p, c->d_cellcount.p, relevant, ps_pos);
  }
  // The z-sweep spread reads the packed charges itself (it has its own, mesh-aligned sort), so the physical
  // cell sort then feeds only the real-space pair kernel and moves to the side stream with it.
  const bool use_sweep = kspace_mode == CONP_KSPACE_PPPM && !c->spread_atomic && c->swplan.usable &&
                         c->spread_mode != 1 && c->spread_mode != 2;
  // red.global spread: it reads the packed charges / the inbox as they arrived (unsorted; red.global does not
  // care: cfg4 +1.7 %), so here too the sort is only the pair kernel's business.  CONP_SPREAD_UNSORTED=0: sorted
  // (one GPU only).
  const bool spread_inbox = kspace_mode == CONP_KSPACE_PPPM && c->spread_atomic && (multi || c->spread_unsorted);
  const bool sort_aside = fork && (use_sweep || spread_inbox);
  cudaStream_t u = sort_aside ? t : s;
  if (sort_aside) {
    CUDA_CHECK(cudaEventRecord(c->ev_sorted, s));
    CUDA_CHECK(cudaStreamWaitEvent(t, c->ev_sorted, 0));
  }
  c->launches += launch_cell_scan(u, g.ncells, c->d_cellcount.p, c->d_cellstart.p, c->d_packed.p, c->mpad,
                                  c->nranks, multi ? c->scal(2) : nullptr);
  c->launches += launch_cell_scatter(u, g, c->m_slots, c->d_packed.p, c->d_ptype.p, c->d_cellof.p, c->d_slot.p,
                                     c->d_cellstart.p, c->d_sorted.p, c->d_stype.p, c->d_ssrc.p, c->d_sortedf.p);
  stage_mark(c, 2);

  // ---- real-space part of b (blist_coul_cal), beside the k-space chain -----------
  if (fork && !sort_aside) {
    CUDA_CHECK(cudaEventRecord(c->ev_sorted, s));
    CUDA_CHECK(cudaStreamWaitEvent(t, c->ev_sorted, 0));
  }
  if (c->rc_b > 0.0 && c->m_total > 0 && nr > 0) {
    c->launches += launch_pair_b(t, g, pair_tables(c, c->d_cuteff_b.p), c->r0, c->r1, c->d_ex.p, c->d_ey.p,
                                 c->d_ez.p, c->d_etype.p, c->d_runstart.p, c->d_runs.p, c->d_sorted.p,
                                 c->d_stype.p, c->d_sortedf.p, c->d_cellstart.p, c->d_breal.p);
  } else {
    CUDA_CHECK(cudaMemsetAsync(c->d_breal.p + c->r0, 0, sizeof(double) * std::max(nr, 1), t));
  }
  if (fork) CUDA_CHECK(cudaEventRecord(c->ev_pair, t));
  stage_mark(c, 3);

  // ---- k-space part of b --------------------------------------------------------
  const double spref = slab_pref(c);
  // PPPM mode on the peer-to-peer path: the gather kernel stores b into every peer itself
  const bool fused_b = fused && kspace_mode == CONP_KSPACE_PPPM;
  if (kspace_mode == CONP_KSPACE_PPPM) {
    const PPPMGeom &pg = c->pg;
    if (fork) CUDA_CHECK(cudaStreamWaitEvent(s, c->ev_cleared, 0));
    auto kmark = [&](int i) {
      if (c->stage_timing && c->debug) CUDA_CHECK(cudaEventRecord(c->kev[i], s));
      trace_mark(c, 10 + i);
    };
    kmark(0);
    if (use_sweep) {
      // all packed charges on one GPU; on several, the inbox slots bin_positions accepted (cell_of >= 0)
      c->launches += launch_pppm_spread_sweep(s, pg, c->swplan, c->h_rho.data(), c->d_packed.p,
                                              multi ? c->m_slots : c->m_total, multi ? c->d_cellof.p : nullptr,
                                              c->d_brick.p, c->d_flag.p);
    } else if (!c->spread_atomic) {
      // the sorted charges: all of them on one GPU, the relevant subset (count in cell_start[ncells]) on several
      c->launches += launch_pppm_spread_tiles(s, pg, c->splan, c->h_rho.data(), c->d_sorted.p, c->d_cellstart.p,
                                              multi ? c->m_slots : c->m_total,
                                              multi ? c->d_cellstart.p + g.ncells : nullptr, c->d_brick.p,
                                              c->d_flag.p);
    } else if (!multi) {
      c->launches += launch_pppm_spread(s, pg, c->d_rho.p, c->m_total, spread_inbox ? c->d_packed.p : c->d_sorted.p,
                                        c->d_brick.p, c->d_flag.p);
    } else {
      // what arrived: this rank's slab charges plus the pair kernel's halo (the kernel skips planes outside the
      // slab).  Grid for the uniform share + 50 %; the kernel grid-strides beyond it.
      const int *counts = routed ? (const int *)(p2p_local(c->p2p) + c->off_rcnt) : c->d_mcounts.p;
      const long long bound = (long long)c->m_total / c->nranks * 3 / 2 + 1024;
      c->launches += launch_pppm_spread(s, pg, c->d_rho.p, (int)std::min<long long>(bound, c->m_slots), c->d_packed.p,
                                        c->d_brick.p, c->d_flag.p, counts, c->nranks, c->mpad, c->d_cellof.p);
    }
    kmark(1);
    if (pg.zs_n > 0) CUFFT_CHECK(cufftExecD2Z(c->plan_f, c->d_brick.p, c->d_rhat.p));
    kmark(2);
    const bool uhat_p2p = fused && !c->uhat_nccl;
    c->launches += launch_pppm_zconv(s, (int)c->ncol, pg.nz, pg.zs_n, pg.zs_lo, pg.nzo, c->d_krad.p, c->zplan,
                                     c->d_rhat.p, c->k_real ? c->d_Kr.p : nullptr, c->d_Kc.p, c->d_uhat.p,
                                     uhat_p2p ? producer(1) : PeerSync());
    if (uhat_p2p && late_signal && !raise_in_consumer) c->launches += p2p_signal(c->p2p, 1, s);
    kmark(3);
    // every rank holds the partial sum over its slab: one small all-reduce completes the spectra (zconv has
    // announced its partial; the owner of a slice pulls it from every rank, sums, and stores it everywhere)
    if (multi) {
      if (uhat_p2p)
        c->launches += p2p_allreduce_pull_f64(c->p2p, c->off_uhat, 2 * (size_t)pg.nzo * c->ncol, 1, 2, s,
                                              raise_in_consumer ? 1 : 0);
      else
        comm_allreduce_sum_f64(c->comm, (double *)c->d_uhat.p, 2 * (size_t)pg.nzo * c->ncol, s);
    }
    kmark(4);
    CUFFT_CHECK(cufftExecZ2D(c->plan_b, c->d_uhat.p, c->d_ubrick.p));
    kmark(5);
    c->launches += 2;  // at least one kernel per cuFFT exec (library)
    stage_mark(c, 4);
    if (fork) CUDA_CHECK(cudaStreamWaitEvent(s, c->ev_pair, 0));  // join
    c->launches += launch_pppm_gather_b(s, c->pg, c->r0, c->r1, c->d_poff.p, c->d_pw.p, c->d_ubrick.p,
                                        c->d_ez.p, c->scal(2), spref, c->d_breal.p, c->d_bk.p, c->d_b.p,
                                        fused_b ? producer(3) : PeerSync(), c->off_b);
    // (the symmetric matvec's consumers wait for b themselves and can raise the flags first; the stand-alone
    // wait kernel of the general GEMV path cannot)
    if (fused_b && late_signal && !(raise_in_consumer && fused && c->sym)) c->launches += p2p_signal(c->p2p, 3, s);
  } else {
    const EwaldHost &e = c->ew;
    const bool gemm = use_ewald_gemm(c);
    // sincos_b + sfac_reduce (km_ewald.cpp:668-786): every rank sums the structure factors of its share of the
    // charges and the partial S(k) are added across ranks
    // = its own charges (the block it packed itself; the sort order does not matter for a sum)
    const PosQ *own = routed ? c->d_own.p : packed_local;
    const int mo = c->m_local;
    c->launches += launch_axis_tables(s, mo, nullptr, nullptr, nullptr, own, e.unitk, e.kxmax, e.kymax, e.kzmax,
                                      c->d_jtab.p);
    if (gemm)
      c->launches += ewald_gemm_sfac(s, c->eg, e, 0, mo, own, c->d_jtab.p, c->d_kz.p, c->d_sfac.p);
    else
      c->launches += launch_ewald_sfac(s, mo, own, c->d_jtab.p, e.kxmax, e.kymax, e.kzmax, e.kcount, c->d_kx.p,
                                       c->d_ky.p, c->d_kz.p, c->d_sfac.p);
    if (multi) comm_allreduce_sum_f64(c->comm, c->d_sfac.p, 2 * (size_t)std::max(e.kcount, 1), s);
    stage_mark(c, 4);
    if (fork) CUDA_CHECK(cudaStreamWaitEvent(s, c->ev_pair, 0));  // join
    if (gemm)
      c->launches += ewald_gemm_bextract(s, c->eg, e, c->r0, c->r1, c->d_kz.p, c->d_ug.p, c->d_sfac.p,
                                         c->d_ez.p, c->scal(2), spref, c->d_breal.p, c->d_bk.p, c->d_b.p);
    else
      c->launches += launch_ewald_bextract(s, c->r0, c->r1, c->d_etab.p, e.kxmax, e.kymax, e.kzmax, e.kcount,
                                           c->d_kx.p, c->d_ky.p, c->d_kz.p, c->d_ug.p, c->d_sfac.p, c->d_ez.p,
                                           c->scal(2), spref, c->d_breal.p, c->d_bk.p, c->d_b.p);
  }
  stage_mark(c, 5);

  // ---- exchange b, matvec (+ fused epilogue on one GPU), exchange S.b ------------------
  // Peer-to-peer path with the symmetric matvec: no stand-alone exchange kernels at all -- b arrives
  // while symv_tma_kernel is already streaming S (its consumers poll the flags), symv_reduce_kernel
  // stores this rank's partial product into every rank's staging slot, and update_charge_sum_kernel
  // waits, adds the slots in rank order and runs the epilogue.
  const bool fused_mv = fused && c->sym;
  if (multi && !fused_b) {
    if (c->p2p) {
      c->launches += p2p_push(c->p2p, c->off_b + sizeof(double) * (size_t)c->rank * c->rpr,
                              sizeof(double) * c->rpr, 3, s);
      if (!fused_mv) c->launches += p2p_wait(c->p2p, 3, s);
    } else {
      comm_allgather(c->comm, c->d_b.p + c->r0, c->d_b.p, sizeof(double) * c->rpr, s);
    }
  } else if (fused_b && !fused_mv) {
    c->launches += p2p_wait(c->p2p, 3, s);
  }
  stage_mark(c, 6);
  if (!multi) {
    const ChargeEpilogue ep = make_epilogue(c, variant, true);
    c->launches += enqueue_matvec(c, s, c->d_b.p, c->d_sb.p, &ep);
    stage_mark(c, 7);
  } else if (fused_mv) {
    PeerSync wait_b = p2p_sync(c->p2p, 3), wait_parts = p2p_sync(c->p2p, 4);
    wait_b.raise_first = (raise_in_consumer && fused_b) ? 1 : 0;  // Ewald mode: p2p_push has signalled already
    wait_parts.raise_first = raise_in_consumer ? 1 : 0;
    c->launches += launch_symv(s, c->d_mat.p, c->pitch, c->N, c->r0, nr, c->d_b.p, c->sy, c->d_rowpart.p,
                               c->d_colpart.p, c->d_sb.p, (int)c->vlen, nullptr, wait_b, producer(4), c->off_parts);
    if (late_signal && !raise_in_consumer) c->launches += p2p_signal(c->p2p, 4, s);
    stage_mark(c, 7);
    c->launches += launch_update_charge_sum(s, make_epilogue(c, variant, false), wait_parts,
                                            (const double *)(p2p_local(c->p2p) + c->off_parts), (int)c->vlen,
                                            c->d_sb.p);
  } else {
    c->launches += enqueue_matvec(c, s, c->d_b.p, c->d_sb.p, nullptr);
    if (c->sym) {  // every rank holds a partial sum over its half band: all-reduce instead of all-gather
      comm_allreduce_sum_f64(c->comm, c->d_sb.p, c->vlen, s);
    } else if (c->p2p) {
      c->launches += p2p_allgather(c->p2p, c->off_sb, sizeof(double) * c->rpr, sizeof(double) * c->rpr, 4, s);
    } else {
      comm_allgather(c->comm, c->d_sb.p + c->r0, c->d_sb.p, sizeof(double) * c->rpr, s);
    }
    stage_mark(c, 7);
    c->launches += launch_update_charge(s, make_epilogue(c, variant, false));
  }
  const double *qinit = c->have_qinit ? c->d_qinit.p : nullptr;
  if (kspace_mode == CONP_KSPACE_PPPM) {  // charges + kspmod->update_charge() -> ele_make_rho
    c->launches += launch_pppm_ele_spread(s, c->pg, c->N, c->r0, c->r1, c->d_widx.p, c->d_weights.p, c->d_sb.p,
                                          c->d_setq.p, qinit, c->scal(0), c->d_q.p, c->d_ebrick.p);
  } else {
    c->launches += launch_finalize_q(s, c->N, c->d_sb.p, c->d_setq.p, qinit, c->scal(0), c->d_q.p);
  }
  stage_mark(c, 8);
}

void solve_device(conp_ctx *c, const double *x_dev, int kspace_mode, int variant, double value) {
  need(c->have_setq, "conp_pre_force: setup incomplete (conp_set_unit_voltage not called)");
  need(c->have_atoms, "conp_pre_force: conp_post_neighbor not called");
  if (kspace_mode != CONP_KSPACE_PPPM && kspace_mode != CONP_KSPACE_EWALD)
    CONP_THROW(CONP_ERR_ARG, "conp_pre_force: unknown kspace mode %d", kspace_mode);
  if (kspace_mode == CONP_KSPACE_PPPM) need(c->have_pppm, "conp_pre_force: PPPM mode needs conp_pppm_setup");
  if (variant < 0 || variant > 2) CONP_THROW(CONP_ERR_ARG, "conp_pre_force: unknown variant %d", variant);
  cudaStream_t s = c->stream;
  if (x_dev != c->d_xraw.p && c->nlocal > 0)
    CUDA_CHECK(cudaMemcpyAsync(c->d_xraw.p, x_dev, sizeof(double) * 3 * (size_t)c->nlocal, cudaMemcpyDeviceToDevice,
                               s));
  double *slot = c->h_value.p + (c->value_slot++ & 63);
  *slot = value;
  CUDA_CHECK(cudaMemcpyAsync(c->scal(12), slot, sizeof(double), cudaMemcpyHostToDevice, s));
  if (kspace_mode == CONP_KSPACE_EWALD) {
    const EwaldHost &e = c->ew;
    c->d_jtab.reserve(((size_t)e.kxmax + e.kymax + e.kzmax + 3) * (size_t)std::max(c->m_total, 1));
    if (use_ewald_gemm(c)) ensure_ewald_gemm(c);
  }

  cudaGraphExec_t &exec = c->graph_exec[kspace_mode][variant];
  if (!c->use_graph || c->stage_timing) {
    enqueue_step(c, kspace_mode, variant);
  } else if (exec) {
    CUDA_CHECK(cudaGraphLaunch(exec, s));
    c->launches += c->graph_launches[kspace_mode][variant];
  } else if (c->eager_runs[kspace_mode][variant]++ < 1) {
    enqueue_step(c, kspace_mode, variant);  // first run eager: sets kernel attributes, sizes buffers
  } else {
    const long long before = c->launches;
    cudaGraph_t graph = nullptr;
    CUDA_CHECK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    try {
      enqueue_step(c, kspace_mode, variant);
    } catch (...) {
      cudaStreamEndCapture(s, &graph);
      if (graph) cudaGraphDestroy(graph);
      throw;
    }
    CUDA_CHECK(cudaStreamEndCapture(s, &graph));
    c->graph_launches[kspace_mode][variant] = c->launches - before;
    cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
      exec = nullptr;
      CONP_THROW(CONP_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
    }
    CUDA_CHECK(cudaGraphLaunch(exec, s));
  }
  c->solved = true;
  c->density_gathered = false;
  if (c->stage_timing) {
    CUDA_CHECK(cudaStreamSynchronize(s));
    for (int i = 0; i < NSTAGE; ++i) {
      float ms = 0;
      CUDA_CHECK(cudaEventElapsedTime(&ms, c->sev[i], c->sev[i + 1]));
      c->stage_ms[i] += ms;
    }
    if (kspace_mode == CONP_KSPACE_PPPM && c->debug)
      for (int i = 0; i < 5; ++i) {
        float ms = 0;
        CUDA_CHECK(cudaEventElapsedTime(&ms, c->kev[i], c->kev[i + 1]));
        c->kev_ms[i] += ms;
      }
    c->stage_n++;
  }
}

void check_range_flag(conp_ctx *c) {
  if (!c->have_pppm) return;
  int flag = 0;
  CUDA_CHECK(cudaMemcpyAsync(&flag, c->d_flag.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  if (flag) CONP_THROW(CONP_ERR_RANGE, "Out of range atoms - cannot compute PPPM");
}

// full-matrix projection(s) of inv_project on S (ld = N)
void project_full(conp_ctx *c, double *S, int nullneutral, int zneutr) {
  const int N = c->N;
  DevBuf<double> w;
  w.reserve(N);
  // first pass always evaluates e^T S e (reported even without neutralisation, fix_conp.cpp:986-1009)
  c->launches += launch_project(c->stream, N, S, N, nullptr, w.p, c->scal(3), nullneutral);
  if (nullneutral) c->projected_here = true;
  double tot = 0;
  CUDA_CHECK(cudaMemcpyAsync(&tot, c->scal(3), sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  c->ee = tot;
  if (nullneutral && zneutr) {
    const double zhalf = 0.5 * c->prd[2] + c->boxlo[2];
    std::vector<int> pos(N);
    for (int i = 0; i < N; ++i) pos[i] = c->h_xyz[3 * i + 2] > zhalf;  // :1039
    DevBuf<int> dpos;
    dpos.upload(pos, c->stream);
    c->launches += launch_project(c->stream, N, S, N, dpos.p, w.p, c->scal(3), 1);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
  }
}

void store_rows_from_full(conp_ctx *c, double *full) {
  decide_symmetry(c, full, true);
  const int nr = c->r1 - c->r0;
  c->d_mat.zero((size_t)std::max(nr, 1) * c->pitch, c->stream);
  if (nr > 0)
    CUDA_CHECK(cudaMemcpy2DAsync(c->d_mat.p, c->pitch * sizeof(double), full + (size_t)c->r0 * c->N,
                                 (size_t)c->N * sizeof(double), (size_t)c->N * sizeof(double), nr,
                                 cudaMemcpyDeviceToDevice, c->stream));
}

// Full periodic mesh potential on demand (the per-step path only evaluates the electrode planes): the
// reference's 3-D transform, pppm_conp.cpp:230-267, of the electrolyte density or (include_ele) of the
// electrolyte + electrode density -- what PPPM's u_brick holds after a force pass with per-atom energy.
void full_mesh_potential(conp_ctx *c, bool include_ele, DevBuf<double> &full) {
  cudaStream_t s = c->stream;
  const PPPMGeom &g = c->pg;
  DevBuf<double> gd;
  DevBuf<cufftDoubleComplex> work;
  full.zero(c->ngrid, s);
  work.reserve(c->nhalf);
  gd.upload(c->h_ghalf, s);
  c->launches += launch_expand_planes(s, c->plane, g.zs_n, g.nz, g.zin_lo + g.zs_lo, nullptr, c->d_brick.p, full.p);
  if (include_ele) {
    DevBuf<double> fe;
    fe.zero(c->ngrid, s);
    c->launches += launch_expand_planes(s, c->plane, g.nzo, g.nz, 0, c->d_zout.p, c->d_ebrick.p, fe.p);
    c->launches += launch_add_bricks(s, c->ngrid, full.p, fe.p, full.p);
    CUDA_CHECK(cudaStreamSynchronize(s));
  }
  if (c->nranks > 1) comm_allreduce_sum_f64(c->comm, full.p, c->ngrid, s);  // z-slabs / own rows of every rank
  cufftHandle pf, pb;
  CUFFT_CHECK(cufftPlan3d(&pf, g.nz, g.ny, g.nx, CUFFT_D2Z));
  CUFFT_CHECK(cufftPlan3d(&pb, g.nz, g.ny, g.nx, CUFFT_Z2D));
  CUFFT_CHECK(cufftSetStream(pf, s));
  CUFFT_CHECK(cufftSetStream(pb, s));
  CUFFT_CHECK(cufftExecD2Z(pf, full.p, work.p));
  c->launches += launch_pppm_green_mul(s, c->nhalf, work.p, gd.p);
  CUFFT_CHECK(cufftExecZ2D(pb, work.p, full.p));
  CUDA_CHECK(cudaStreamSynchronize(s));
  cufftDestroy(pf);
  cufftDestroy(pb);
}

}  // namespace

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

int conp_abi_version(void) { return CONP_ABI_VERSION; }

int conp_device_count(int *count_out) {
  if (!count_out) return CONP_ERR_ARG;
  int ndev = 0, ok = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess) ndev = 0;
  for (int d = 0; d < ndev; ++d) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, d) == cudaSuccess && prop.major >= 10) ++ok;
  }
  *count_out = ok;
  return ok > 0 ? CONP_OK : CONP_ERR_CUDA;
}

int conp_get_unique_id(void *id_out) {
  try {
    return comm_get_unique_id(id_out);
  } catch (const Error &e) {
    g_create_error = e.msg;
    return e.code;
  }
}

int conp_create(conp_ctx **out, int device, int rank, int nranks, const void *unique_id) {
  if (!out || nranks < 1 || rank < 0 || rank >= nranks) {
    g_create_error = "conp_create: bad arguments";
    return CONP_ERR_ARG;
  }
  *out = nullptr;
  conp_ctx *c = nullptr;
  try {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
      CONP_THROW(CONP_ERR_CUDA, "conp_create: no CUDA device (%s); this library has no CPU fallback",
                 e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= ndev) CONP_THROW(CONP_ERR_ARG, "conp_create: device %d of %d", device, ndev);
    CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
      CONP_THROW(CONP_ERR_CUDA, "conp_create: device %s is sm_%d%d; kernels are built for sm_100a only", prop.name,
                 prop.major, prop.minor);
    c = new conp_ctx;
    c->device = device;
    c->rank = rank;
    c->nranks = nranks;
    c->num_sms = prop.multiProcessorCount;
    {  // the k-space chain is the critical path of the step: main stream at the highest priority
      int prio_lo = 0, prio_hi = 0;
      CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
      CUDA_CHECK(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_hi));
      CUDA_CHECK(cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, prio_lo));
    }
    for (cudaEvent_t *e : {&c->ev_begin, &c->ev_cleared, &c->ev_sorted, &c->ev_pair})
      CUDA_CHECK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    c->overlap = getenv("CONP_NO_OVERLAP") == nullptr;
    for (auto &ev : c->ev) CUDA_CHECK(cudaEventCreate(&ev));
    for (auto &ev : c->sev) CUDA_CHECK(cudaEventCreate(&ev));
    for (auto &ev : c->kev) CUDA_CHECK(cudaEventCreate(&ev));
    c->debug = getenv("CONP_DEBUG") != nullptr;
    c->spread_unsorted = getenv("CONP_SPREAD_UNSORTED") == nullptr || atoi(getenv("CONP_SPREAD_UNSORTED")) != 0;
    c->trace = getenv("CONP_TRACE") != nullptr && atoi(getenv("CONP_TRACE")) != 0;
    if (c->trace) {
      c->d_trace.reserve(48);
      CUDA_CHECK(cudaMemset(c->d_trace.p, 0, sizeof(unsigned long long) * 48));
    }
    if (getenv("CONP_EWALD_GEMM")) c->eg_mode = atoi(getenv("CONP_EWALD_GEMM")) != 0 ? 1 : 0;
    c->signal_in_kernel = getenv("CONP_SIGNAL_IN_KERNEL") != nullptr && atoi(getenv("CONP_SIGNAL_IN_KERNEL")) != 0;
    c->fused_signal = getenv("CONP_FUSED_SIGNAL") == nullptr || atoi(getenv("CONP_FUSED_SIGNAL")) != 0;
    c->uhat_nccl = getenv("CONP_UHAT_NCCL") != nullptr && atoi(getenv("CONP_UHAT_NCCL")) != 0;
    c->route_allowed = getenv("CONP_ROUTE") == nullptr || atoi(getenv("CONP_ROUTE")) != 0;
    if (const char *e = getenv("CONP_SPREAD"))
      c->spread_mode = !strcmp(e, "atomic") ? 0 : !strcmp(e, "smem") ? 1 : !strcmp(e, "mma") ? 2 :
                       !strcmp(e, "sweep") ? 3 : -1;
    c->d_scal.zero(16, c->stream);
    c->d_partials.zero(3 * 1024, c->stream);
    c->d_counter.zero(1, c->stream);
    c->h_value.reserve(64);
    c->use_graph = getenv("CONP_NO_GRAPH") == nullptr;
    c->sym_allowed = getenv("CONP_NO_SYMV") == nullptr;
    c->comm = comm_create(rank, nranks, unique_id);
    CUSOLVER_CHECK(cusolverDnCreate(&c->solver));
    CUSOLVER_CHECK(cusolverDnSetStream(c->solver, c->stream));
    // cuBLAS only measures the DGEMM ceiling quoted next to the Gram (conp_bench_dgemm_tflops); no product
    // path calls it
    CUBLAS_CHECK(cublasCreate(&c->blas));
    CUBLAS_CHECK(cublasSetStream(c->blas, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    *out = c;
    return CONP_OK;
  } catch (const Error &e) {
    g_create_error = e.msg;
    if (c) conp_destroy(c);
    return e.code;
  }
}

void conp_destroy(conp_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  drop_graphs(c);
  if (c->plans) {
    cufftDestroy(c->plan_f);
    cufftDestroy(c->plan_b);
  }
  if (c->solver) cusolverDnDestroy(c->solver);
  if (c->blas) cublasDestroy(c->blas);
  try { drop_p2p(c); } catch (...) {}
  comm_destroy(c->comm);
  for (auto &ev : c->ev) cudaEventDestroy(ev);
  for (auto &ev : c->sev) cudaEventDestroy(ev);
  for (auto &ev : c->kev) cudaEventDestroy(ev);
  for (cudaEvent_t e : {c->ev_begin, c->ev_cleared, c->ev_sorted, c->ev_pair})
    if (e) cudaEventDestroy(e);
  if (c->side) cudaStreamDestroy(c->side);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

const char *conp_last_error(const conp_ctx *c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int conp_get_info(const conp_ctx *c, conp_info *o) {
  if (!c || !o) return CONP_ERR_ARG;
  std::memset(o, 0, sizeof(*o));
  o->abi_version = CONP_ABI_VERSION;
  o->device = c->device; o->rank = c->rank; o->nranks = c->nranks;
  o->n_ele = c->N; o->row_begin = c->r0; o->row_end = c->r1; o->n_elyte = c->m_total;
  o->kxmax = c->ew.kxmax; o->kymax = c->ew.kymax; o->kzmax = c->ew.kzmax;
  o->kcount = c->ew.kcount; o->kcount_flat = c->ew.kcount_flat; o->kcount_expand = c->ew.kcount_expand;
  if (c->have_pppm) { o->mesh[0] = c->pg.nx; o->mesh[1] = c->pg.ny; o->mesh[2] = c->pg.nz; o->order = c->pg.order; }
  o->matrix_pitch = (long long)c->pitch;
  o->launches = c->launches;
  o->setup_build_ms = c->build_ms; o->setup_invert_ms = c->invert_ms;
  o->ee = c->ee; o->dd = c->dd; o->totsetq = c->totsetq;
  o->symmetric_matvec = c->sym ? 1 : 0;
  o->asymmetry = c->asym_rel;
  return CONP_OK;
}

// ---- setup -----------------------------------------------------------------

int conp_set_cell(conp_ctx *c, const double boxlo[3], const double prd[3], const int periodic[3], int slabflag,
                  double slab_volfactor, int ff_flag) {
  return guard(c, [&] {
    for (int a = 0; a < 3; ++a) {
      if (!(prd[a] > 0.0) || !std::isfinite(boxlo[a]))
        CONP_THROW(CONP_ERR_ARG, "Non-numeric box dimensions - simulation unstable");  // pppm_conp.cpp:136-137
      c->boxlo[a] = boxlo[a]; c->prd[a] = prd[a]; c->periodic[a] = periodic[a] ? 1 : 0;
    }
    if (ff_flag < 0 || ff_flag > 2) CONP_THROW(CONP_ERR_ARG, "conp_set_cell: bad ff_flag");
    c->slabflag = slabflag ? 1 : 0;
    c->slab_volfactor = c->slabflag ? slab_volfactor : 1.0;
    c->ff_flag = ff_flag;
    c->have_cell = true;
  });
}

int conp_set_ewald(conp_ctx *c, double g_ewald, double accuracy_abs, double q2, long long natoms, int lowmem) {
  return guard(c, [&] {
    need(c->have_cell, "conp_set_ewald: call conp_set_cell first");
    (void)lowmem;
    c->g_ewald = g_ewald;
    ewald_setup_host(c->ew, g_ewald, accuracy_abs, q2, natoms, c->prd, c->slab_volfactor);
    c->d_kx.upload(c->ew.kx, c->stream);
    c->d_ky.upload(c->ew.ky, c->stream);
    c->d_kz.upload(c->ew.kz, c->stream);
    c->d_ug.upload(c->ew.ug, c->stream);
    c->d_sfac.zero(2 * (size_t)std::max(c->ew.kcount, 1), c->stream);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->have_ewald = true;
    c->eg_valid = false;
  });
}

int conp_set_pair(conp_ctx *c, int pairmode, double eta, double cut_coul, int ntypes, const double *cutsq,
                  const double *eta_ij, const double *fo_ij, const double *u0_i, int smartlist,
                  const int *is_eletype) {
  return guard(c, [&] {
    need(c->have_ewald, "conp_set_pair: call conp_set_ewald first (needs g_ewald)");
    if (ntypes < 1 || !cutsq) CONP_THROW(CONP_ERR_ARG, "Fix conp couldn't detect a Coulombic pair style");
    if (pairmode != CONP_PAIR_ETA && pairmode != CONP_PAIR_EHGO) CONP_THROW(CONP_ERR_ARG, "bad pairmode");
    if (pairmode == CONP_PAIR_EHGO && (!eta_ij || !fo_ij || !u0_i))
      CONP_THROW(CONP_ERR_ARG, "EHGO pair mode needs eta_ij, fo_ij and u0_i tables");
    if (smartlist && !is_eletype) CONP_THROW(CONP_ERR_ARG, "Invalid fix conp command (Insufficient input entries for etypes)");
    const int n1 = ntypes + 1;
    c->pairmode = pairmode; c->eta = eta; c->cut_coul = cut_coul; c->ntypes = ntypes; c->smartlist = smartlist;
    c->h_cutsq.assign(cutsq, cutsq + (size_t)n1 * n1);
    std::vector<double> zero((size_t)n1 * n1, 0.0);
    c->d_eta_ij.upload(eta_ij ? eta_ij : zero.data(), (size_t)n1 * n1, c->stream);
    c->d_fo_ij.upload(fo_ij ? fo_ij : zero.data(), (size_t)n1 * n1, c->stream);
    c->h_u0.assign(n1, 0.0);
    if (u0_i) c->h_u0.assign(u0_i, u0_i + n1);
    c->d_u0.upload(c->h_u0, c->stream);
    // cut_coulsq = min(cut_coul^2, ERFC_MAX^2/g^2)  fix_conp.cpp:1236-1238, 1303-1305
    double cut_coulsq = cut_coul * cut_coul;
    const double cut_erfc = ERFC_MAX * ERFC_MAX / (c->g_ewald * c->g_ewald);
    if (cut_coulsq > cut_erfc) cut_coulsq = cut_erfc;
    std::vector<double> cb((size_t)n1 * n1, 0.0), ca((size_t)n1 * n1, 0.0), cf((size_t)n1 * n1, 0.0);
    double mb = 0, ma = 0, mf = 0;
    for (int it = 1; it <= ntypes; ++it)
      for (int jt = 1; jt <= ntypes; ++jt) {
        const size_t ij = (size_t)it * n1 + jt;
        const bool ei = smartlist && is_eletype[it], ej = smartlist && is_eletype[jt];
        const bool listed_b = !smartlist || (ei != ej);             // request_smartlist :327-332
        const bool listed_a = !smartlist || (it == jt && ei);       // :322-325
        const double cs = cutsq[ij];
        if (listed_b) { cb[ij] = std::min(cs, cut_coulsq); cf[ij] = cs; }
        if (listed_a) ca[ij] = std::min(cs, cut_coulsq);
        mb = std::max(mb, cb[ij]); ma = std::max(ma, ca[ij]); mf = std::max(mf, cf[ij]);
      }
    c->d_cuteff_b.upload(cb, c->stream);
    c->d_cuteff_a.upload(ca, c->stream);
    c->d_cutsq_listed.upload(cf, c->stream);
    c->rc_b = std::sqrt(mb);
    c->rc_a = std::sqrt(ma);
    // post_force guard eta^2 r^2 < ERFC_MAX (sic), fix_conp.cpp:1418-1419
    c->rc_f = std::sqrt(std::min(mf, ERFC_MAX / (eta * eta) * (1.0 + 1e-9)));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->have_pair = true;
    c->static_cells = false;
  });
}

int conp_set_electrodes(conp_ctx *c, int n_ele, const int *tag, const int *type, const int *side,
                        const double *xyz) {
  return guard(c, [&] {
    need(c->have_cell && c->have_ewald, "conp_set_electrodes: call conp_set_cell/conp_set_ewald first");
    if (n_ele < 1 || !type || !side || !xyz) CONP_THROW(CONP_ERR_ARG, "conp_set_electrodes: empty electrode");
    if (n_ele > 65535) CONP_THROW(CONP_ERR_ARG, "conp_set_electrodes: n_ele > 65535 not supported");
    const int N = n_ele;
    // b and S.b may be views into the exchange arena sized for the previous electrode: rebuild it
    // (conp_post_neighbor must follow, it sizes the arena for the new vectors)
    drop_p2p(c);
    c->have_atoms = false;
    c->N = N;
    c->h_tag.assign(N, 0);
    if (tag) c->h_tag.assign(tag, tag + N);
    c->h_type.assign(type, type + N);
    c->h_side.assign(side, side + N);
    c->h_xyz.assign(xyz, xyz + 3 * (size_t)N);
    for (int i = 0; i < N; ++i) {
      if (side[i] != 1 && side[i] != -1) CONP_THROW(CONP_ERR_ARG, "conp_set_electrodes: side must be +1/-1");
      if (type[i] < 1 || (c->have_pair && type[i] > c->ntypes))
        CONP_THROW(CONP_ERR_ARG, "conp_set_electrodes: atom type out of range");
    }
    // equal contiguous row blocks, multiple of 16 rows so every block start is 128-byte aligned
    int rpr_all = 0;
    conp_row_block(N, c->nranks, c->rank, &c->r0, &c->r1, &rpr_all);
    c->rpr = rpr_all;
    c->ncols_pad = (int)round_up(N, 16);
    c->pitch = c->ncols_pad;
    c->vlen = std::max((size_t)c->nranks * c->rpr, (size_t)c->ncols_pad) + 16;
    // symmetric matvec: every rank must take the same branch, so the decision uses the largest block
    c->sym = false;
    c->asym_rel = -1.0;
    {
      std::vector<int2> strips;
      c->sym_plan_ok = plan_symv(N, 0, std::min(c->rpr, N), c->num_sms, strips).usable;
      c->sy = plan_symv(N, c->r0, c->r1 - c->r0, c->num_sms, strips);
      c->sym_plan_ok = c->sym_plan_ok && c->sy.usable;
      if (c->sym_plan_ok) {
        c->d_rowpart.zero(c->vlen, c->stream);
        c->d_colpart.zero(std::max<size_t>((size_t)c->sy.nstrips * c->sy.L, 2), c->stream);
        if (strips.empty()) strips.push_back(make_int2(0, 0));
        c->d_strips.upload(strips, c->stream);
        c->sy.strips = c->d_strips.p;
      }
    }
    std::vector<double> hx(N), hy(N), hz(N);
    for (int i = 0; i < N; ++i) { hx[i] = xyz[3 * i]; hy[i] = xyz[3 * i + 1]; hz[i] = xyz[3 * i + 2]; }
    cudaStream_t s = c->stream;
    c->d_ex.upload(hx, s); c->d_ey.upload(hy, s); c->d_ez.upload(hz, s);
    c->d_exyz.upload(c->h_xyz, s);
    c->d_etype.upload(c->h_type, s);
    c->d_eside.upload(c->h_side, s);
    for (DevBuf<double> *v : {&c->d_b, &c->d_bk, &c->d_breal, &c->d_sb, &c->d_q, &c->d_setq, &c->d_dvec, &c->d_setz,
                              &c->d_qinit})
      v->zero(c->vlen, s);
    // electrode axis tables (used by the A build and by Ewald-mode b extraction)
    const EwaldHost &e = c->ew;
    const size_t T = (size_t)e.kxmax + e.kymax + e.kzmax + 3;
    c->d_etab.reserve(T * N);
    c->launches += launch_axis_tables(s, N, c->d_ex.p, c->d_ey.p, c->d_ez.p, nullptr, e.unitk, e.kxmax, e.kymax,
                                      e.kzmax, c->d_etab.p);
    CUDA_CHECK(cudaStreamSynchronize(s));
    c->have_ele = true;
    c->eg_valid = false;
    c->static_cells = false;
    c->have_A = c->inverted = c->have_setq = false;
  });
}

int conp_pppm_setup(conp_ctx *c, const int mesh[3], int order, const double *rho_coeff, const double *greensfn,
                    double shift, double shiftone) {
  return guard(c, [&] {
    need(c->have_ele, "conp_pppm_setup: call conp_set_electrodes first");
    c->static_cells = false;  // the per-rank cell relevance mask depends on the slab of planes
    // the output-plane spectra live in the exchange arena, whose size depends on the mesh: an arena built by
    // an earlier conp_post_neighbor (the shim's hook order: setup_post_neighbor, then kspace->setup()) is
    // dropped here and conp_post_neighbor must be called again before the next solve
    drop_p2p(c);
    c->have_atoms = false;
    if (order < 1 || order > 7) CONP_THROW(CONP_ERR_ARG, "conp_pppm_setup: PPPM order must be 1..7");
    for (int a = 0; a < 3; ++a)
      if (mesh[a] < order) CONP_THROW(CONP_ERR_ARG, "conp_pppm_setup: mesh smaller than the stencil");
    PPPMGeom &g = c->pg;
    g.nx = mesh[0]; g.ny = mesh[1]; g.nz = mesh[2]; g.order = order;
    g.nlower = -((order - 1) / 2);
    const double prd_slab[3] = {c->prd[0], c->prd[1], c->prd[2] * c->slab_volfactor};
    for (int a = 0; a < 3; ++a) { g.boxlo[a] = c->boxlo[a]; g.delinv[a] = mesh[a] / prd_slab[a]; }
    g.delvolinv = g.delinv[0] * g.delinv[1] * g.delinv[2];
    g.shift = shift; g.shiftone = shiftone;
    const size_t nx = g.nx, ny = g.ny, nz = g.nz, nxh = nx / 2 + 1;
    c->ngrid = nx * ny * nz;
    c->nhalf = nxh * ny * nz;
    c->plane = nx * ny;
    c->ncol = nxh * ny;
    const size_t ncol = c->ncol;
    cudaStream_t s = c->stream;
    std::vector<int> h_krad;
    c->d_rho.upload(rho_coeff, (size_t)order * order, s);
    c->h_rho.assign(rho_coeff, rho_coeff + (size_t)order * order);
    // half-spectrum Green's function, symmetrised and pre-scaled by 1/(nx ny nz):
    // Re IFFT(G rho^) of the reference's complex transform (pppm_conp.cpp:235-266)
    // equals the real transform with G_sym(k) = (G(k) + G(-k))/2.
    std::vector<double> &gh = c->h_ghalf;
    gh.resize(c->nhalf);
    const double scaleinv = 1.0 / ((double)nx * ny * nz);
    for (size_t kz = 0; kz < nz; ++kz)
      for (size_t ky = 0; ky < ny; ++ky)
        for (size_t kx = 0; kx < nxh; ++kx) {
          const size_t a = (kz * ny + ky) * nx + kx;
          const size_t b = (((nz - kz) % nz) * ny + ((ny - ky) % ny)) * nx + ((nx - kx) % nx);
          gh[(kz * ny + ky) * nxh + kx] = 0.5 * (greensfn[a] + greensfn[b]) * scaleinv;
        }
    // ---- cached electrode stencils and the output planes they touch ----------
    c->d_part2grid.reserve(3 * (size_t)c->N);
    c->d_weights.reserve(3 * (size_t)c->N * order);
    c->launches += launch_pppm_ele_stencil(s, g, c->d_rho.p, c->N, c->d_ex.p, c->d_ey.p, c->d_ez.p,
                                           c->d_part2grid.p, c->d_weights.p);
    std::vector<int> p2g(3 * (size_t)c->N);
    CUDA_CHECK(cudaMemcpyAsync(p2g.data(), c->d_part2grid.p, sizeof(int) * p2g.size(), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    std::vector<int> zmap(nz, -1);
    for (int i = 0; i < c->N; ++i)
      for (int n = 0; n < order; ++n) {
        int mz = (p2g[3 * (size_t)i + 2] + g.nlower + n) % (int)nz;
        if (mz < 0) mz += (int)nz;
        zmap[mz] = 0;
      }
    c->h_zout.clear();
    for (int mz = 0; mz < (int)nz; ++mz)
      if (zmap[mz] == 0) { zmap[mz] = (int)c->h_zout.size(); c->h_zout.push_back(mz); }
    g.nzo = (int)c->h_zout.size();
    c->d_zmap.upload(zmap, s);
    c->d_zout.upload(c->h_zout, s);
    g.zmap = c->d_zmap.p;
    c->d_widx.reserve(3 * (size_t)c->N * order);
    c->launches += launch_pppm_ele_index(s, g, c->N, c->d_part2grid.p, c->d_widx.p);
    c->d_poff.reserve((size_t)c->N * order * order * order);
    c->d_pw.reserve((size_t)c->N * order * order * order);
    c->launches += launch_pppm_point_table(s, g, c->N, c->d_widx.p, c->d_weights.p, c->d_poff.p, c->d_pw.p);
    // ---- input planes: what the box can reach (all of them if z is periodic) ----
    if (c->periodic[2]) {
      g.zin_lo = 0;
      g.nzi = (int)nz;
    } else {
      const int base_hi = (int)(c->prd[2] * g.delinv[2] + shift) - 16384;
      const int lo = g.nlower - 2, hi = base_hi + order / 2 + 2;
      g.zin_lo = lo;
      g.nzi = hi - lo + 1;
      if (g.nzi >= (int)nz) { g.zin_lo = 0; g.nzi = (int)nz; }
    }
    // ---- z-convolution kernel table K[col][d] = sum_kz gh[kz][col] exp(+2 pi i kz d / nz) ----
    {
      std::vector<cufftDoubleComplex> gc(c->nhalf);
      for (size_t i = 0; i < c->nhalf; ++i) { gc[i].x = gh[i]; gc[i].y = 0.0; }
      DevBuf<cufftDoubleComplex> dgc;
      dgc.upload(gc, s);
      c->d_Kc.reserve(c->nhalf);
      cufftHandle p1;
      int n1[1] = {(int)nz};
      CUFFT_CHECK(cufftPlanMany(&p1, 1, n1, n1, (int)ncol, 1, n1, 1, (int)nz, CUFFT_Z2Z, (int)ncol));
      CUFFT_CHECK(cufftSetStream(p1, s));
      CUFFT_CHECK(cufftExecZ2Z(p1, dgc.p, c->d_Kc.p, CUFFT_INVERSE));
      CUDA_CHECK(cudaMemcpyAsync(gc.data(), c->d_Kc.p, sizeof(cufftDoubleComplex) * c->nhalf, cudaMemcpyDeviceToHost, s));
      CUDA_CHECK(cudaStreamSynchronize(s));
      cufftDestroy(p1);
      double mre = 0, mim = 0;
      for (size_t i = 0; i < c->nhalf; ++i) { mre = std::max(mre, std::fabs(gc[i].x)); mim = std::max(mim, std::fabs(gc[i].y)); }
      c->k_real = mim <= 1e-14 * mre;
      // per-column window radius.  |K| falls off like the Ewald Gaussian (or exp(-k|d|) for the
      // smallest k_xy) until it hits the ~1e-16 noise floor of this FFT-built table, so the radius
      // is read off at the 1e-13 level, where the table is still clean, and extended by 20 % + 2
      // planes (a Gaussian that is 1e-13 at R is < 1e-18 at 1.2 R).
      {
        std::vector<int> &krad = h_krad;
        krad.assign(ncol, 0);
        for (size_t col = 0; col < ncol; ++col) {
          const cufftDoubleComplex *row = gc.data() + col * nz;
          double kmax = 0;
          for (size_t d = 0; d < nz; ++d) kmax = std::max(kmax, std::hypot(row[d].x, row[d].y));
          const double thr = 1e-13 * kmax;
          int R = 0;
          for (size_t d = 0; d < nz; ++d)
            if (kmax > 0 && std::hypot(row[d].x, row[d].y) >= thr) R = std::max(R, (int)std::min(d, nz - d));
          krad[col] = (int)std::ceil(1.2 * R) + 2;
        }
        if (getenv("CONP_DEBUG")) {
          long long tot = 0;
          int full = 0;
          for (int r : krad) { tot += r; full += (2 * r + 1 >= (int)nz); }
          fprintf(stderr, "[conp] zconv window radius: mean %.1f planes, %d of %zu columns full (nzi %d, nzo %d)\n",
                  (double)tot / ncol, full, ncol, g.nzi, g.nzo);
        }
        c->d_krad.upload(krad, s);
        CUDA_CHECK(cudaStreamSynchronize(s));
      }
      if (c->k_real) {
        std::vector<double> kr(c->nhalf);
        for (size_t i = 0; i < c->nhalf; ++i) kr[i] = gc[i].x;
        c->d_Kr.upload(kr, s);
        CUDA_CHECK(cudaStreamSynchronize(s));
        c->d_Kc.release();
      }
    }
    // ---- this rank's slab of input planes (all of them on one GPU) ------------------
    g.zs_lo = (int)(((long long)g.nzi * c->rank) / c->nranks);
    g.zs_n = (int)(((long long)g.nzi * (c->rank + 1)) / c->nranks) - g.zs_lo;
    {  // split the (kx,ky) column groups between the narrow and the wide z-convolution paths
      std::vector<ZconvGroup> narrow;
      std::vector<int> wide, aout;
      plan_pppm_zconv(h_krad, (int)ncol, (int)nz, g.nzi, g.zs_lo, g.zs_n, g.zin_lo, c->h_zout, c->k_real, narrow,
                      wide, aout, c->zplan);
      if (getenv("CONP_DEBUG"))
        fprintf(stderr, "[conp] zconv plan: %d narrow groups (R <= %d, <= %d planes staged), %d wide columns\n",
                c->zplan.n_narrow, c->zplan.rcap, c->zplan.npcap, c->zplan.n_wide);
      if (narrow.empty()) narrow.resize(1);
      if (wide.empty()) wide.push_back(0);
      if (aout.empty()) aout.push_back(0);
      c->d_zc_narrow.upload(narrow, s);
      c->d_zc_wide.upload(wide, s);
      c->d_zc_aout.upload(aout, s);
      c->zplan.narrow = c->d_zc_narrow.p;
      c->zplan.wide = c->d_zc_wide.p;
      c->zplan.aout = c->d_zc_aout.p;
      CUDA_CHECK(cudaStreamSynchronize(s));
    }
    // ---- compact bricks and batched 2-D plans -------------------------------------
    c->d_brick.zero((size_t)std::max(g.zs_n, 1) * c->plane, s);
    c->d_ubrick.zero((size_t)g.nzo * c->plane, s);
    c->d_ebrick.zero((size_t)g.nzo * c->plane, s);
    c->d_rhat.reserve((size_t)std::max(g.zs_n, 1) * ncol);
    c->d_uhat.reserve((size_t)g.nzo * ncol);
    c->d_flag.zero(1, s);
    if (c->plans) { cufftDestroy(c->plan_f); cufftDestroy(c->plan_b); c->plans = false; }
    int n2[2] = {g.ny, g.nx}, er[2] = {g.ny, g.nx}, ec[2] = {g.ny, (int)nxh};
    CUFFT_CHECK(cufftPlanMany(&c->plan_f, 2, n2, er, 1, (int)c->plane, ec, 1, (int)ncol, CUFFT_D2Z,
                              std::max(g.zs_n, 1)));
    CUFFT_CHECK(cufftPlanMany(&c->plan_b, 2, n2, ec, 1, (int)ncol, er, 1, (int)c->plane, CUFFT_Z2D, g.nzo));
    CUFFT_CHECK(cufftSetStream(c->plan_f, s));
    CUFFT_CHECK(cufftSetStream(c->plan_b, s));
    c->plans = true;
    CUDA_CHECK(cudaStreamSynchronize(s));
    drop_graphs(c);
    c->have_pppm = true;
  });
}

int conp_build_A(conp_ctx *c) {
  return guard(c, [&] {
    need(c->have_ele && c->have_pair, "conp_build_A: electrodes / pair data missing");
    cudaStream_t s = c->stream;
    const int N = c->N, nr = c->r1 - c->r0;
    const EwaldHost &e = c->ew;
    CUDA_CHECK(cudaEventRecord(c->ev[14], s));
    c->d_mat.zero((size_t)std::max(nr, 1) * c->pitch, s);

    // ---- k-space Gram on FP64 tensor cores, chunked over k --------------------
    const int KC = 4096;  // k-vectors per chunk -> panel of 2*KC k-rows
    const size_t ld = round_up(N, 128) + 128;
    DevBuf<double> panel;
    panel.zero((size_t)2 * KC * ld, s);
    DevBuf<int2> dsegs;
    for (int k0 = 0; k0 < e.kcount; k0 += KC) {
      const int kc_real = std::min(KC, e.kcount - k0);
      const int kc = (int)round_up(kc_real, 8);  // 2*kc multiple of 16
      if (kc_real != KC) CUDA_CHECK(cudaMemsetAsync(panel.p, 0, sizeof(double) * 2 * (size_t)kc * ld, s));
      std::vector<int2> segs;
      int a = k0;
      while (a < k0 + kc_real) {
        int b = a + 1;
        while (b < k0 + kc_real && e.kx[b] == e.kx[a] && e.ky[b] == e.ky[a] && e.kz[b] == e.kz[b - 1] + 1) ++b;
        segs.push_back(make_int2(a, b));
        a = b;
      }
      dsegs.upload(segs, s);
      c->launches += launch_ewald_panel(s, N, c->d_etab.p, e.kxmax, e.kymax, e.kzmax, k0, kc, (int)segs.size(),
                                        dsegs.p, c->d_kx.p, c->d_ky.p, c->d_kz.p, c->d_ug.p, panel.p, ld);
      c->launches += launch_gram_accumulate(s, nr, N, 2 * kc, panel.p + c->r0, panel.p, ld, c->d_mat.p, c->pitch,
                                            c->r0);
      CUDA_CHECK(cudaStreamSynchronize(s));  // segs / dsegs reuse
    }
    // the Gram is symmetric and only its cyclic half band was computed: one GPU completes it now, several
    // GPUs when the row blocks are assembled for the inversion (conp_invert_project)
    if (c->nranks == 1) c->launches += launch_gram_mirror(s, N, c->d_mat.p, c->pitch);
    c->A_half_band = c->nranks > 1;
    // ---- diagonal, self and slab terms -----------------------------------------
    const double diag_k = e.ug_tot - 2.0 / MY_PIS * c->g_ewald;  // km_ewald.cpp:632
    const double self_eta = std::sqrt(2.0) / MY_PIS * c->eta;    // fix_conp.cpp:796-800
    c->launches += launch_a_finish(s, c->r0, c->r1, N, c->d_mat.p, c->pitch, diag_k, c->pairmode, self_eta,
                                   c->d_u0.p, c->d_etype.p, slab_pref(c), c->d_ez.p);
    // ---- real-space electrode-electrode pairs -------------------------------------
    if (c->rc_a > 0.0) {
      CellGrid g = make_cell_grid(c->boxlo, c->prd, c->periodic, c->rc_a);
      std::vector<EPos> sorted;
      std::vector<int> cs;
      build_electrode_cells(g, 0, N, c->h_xyz.data(), c->h_type.data(), sorted, cs);
      std::vector<int> run_start;
      std::vector<PairRun> runs;
      build_pair_runs(g, c->r0, c->r1, c->h_xyz.data(), run_start, runs);
      DevBuf<EPos> dsorted;
      DevBuf<int> dcs, drs;
      DevBuf<PairRun> druns;
      dsorted.upload(sorted, s);
      dcs.upload(cs, s);
      drs.upload(run_start, s);
      druns.upload(runs, s);
      c->launches += launch_pair_A(s, g, pair_tables(c, c->d_cuteff_a.p), dsorted.p, dcs.p, c->r0, c->r1, c->d_ex.p,
                                   c->d_ey.p, c->d_ez.p, c->d_etype.p, drs.p, druns.p, c->d_mat.p, c->pitch);
      CUDA_CHECK(cudaStreamSynchronize(s));
    }
    CUDA_CHECK(cudaEventRecord(c->ev[15], s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, c->ev[14], c->ev[15]));
    c->build_ms = ms;
    c->have_A = true;
    c->inverted = false;
    c->projected_here = false;
  });
}

int conp_load_matrix(conp_ctx *c, const double *full, int is_inverse) {
  return guard(c, [&] {
    need(c->have_ele, "conp_load_matrix: call conp_set_electrodes first");
    if (!full) CONP_THROW(CONP_ERR_ARG, "Invalid fix conp command (Cannot open A matrix file)");
    const int N = c->N, nr = c->r1 - c->r0;
    c->d_mat.zero((size_t)std::max(nr, 1) * c->pitch, c->stream);
    c->sym = false;
    c->asym_rel = -1.0;
    c->projected_here = false;
    c->A_half_band = false;
    c->d_fullS.release();
    if (is_inverse) {
      // a ready-made inverse: look at the whole matrix once to see whether the symmetric product applies.
      // The full copy is kept until conp_invert_project / conp_set_unit_voltage know whether this is a
      // one-electrode run, whose inv file holds the UNprojected inverse (fix_conp.cpp:958-977) and is
      // projected after get_setq (:1115).
      DevBuf<double> F;
      F.upload(full, (size_t)N * N, c->stream);
      decide_symmetry(c, F.p, false);
      if (nr > 0)
        CUDA_CHECK(cudaMemcpy2DAsync(c->d_mat.p, c->pitch * sizeof(double), F.p + (size_t)c->r0 * N,
                                     (size_t)N * sizeof(double), (size_t)N * sizeof(double), nr,
                                     cudaMemcpyDeviceToDevice, c->stream));
      CUDA_CHECK(cudaStreamSynchronize(c->stream));
      c->d_fullS.adopt(F);
    } else if (nr > 0) {
      CUDA_CHECK(cudaMemcpy2DAsync(c->d_mat.p, c->pitch * sizeof(double), full + (size_t)c->r0 * N,
                                   (size_t)N * sizeof(double), (size_t)N * sizeof(double), nr,
                                   cudaMemcpyHostToDevice, c->stream));
    }
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    drop_graphs(c);
    c->have_A = true;
    c->inverted = is_inverse != 0;
  });
}

int conp_get_matrix(conp_ctx *c, double *rows_out) {
  return guard(c, [&] {
    need(c->have_A, "conp_get_matrix: no matrix yet");
    need(!c->A_half_band, "conp_get_matrix: on several GPUs the rows of A are complete only after "
                          "conp_invert_project (matout: run on one rank)");
    const int N = c->N, nr = c->r1 - c->r0;
    if (nr > 0)
      CUDA_CHECK(cudaMemcpy2DAsync(rows_out, (size_t)N * sizeof(double), c->d_mat.p, c->pitch * sizeof(double),
                                   (size_t)N * sizeof(double), nr, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
  });
}

int conp_invert_project(conp_ctx *c, int nullneutral, int zneutr, int one_electrode, double *ee_out) {
  return guard(c, [&] {
    need(c->have_A, "conp_invert_project: no A matrix (conp_build_A / conp_load_matrix)");
    cudaStream_t s = c->stream;
    const int N = c->N, nr = c->r1 - c->r0;
    c->one_electrode = one_electrode ? 1 : 0;
    CUDA_CHECK(cudaEventRecord(c->ev[14], s));
    if (!c->inverted) {
      // assemble the full matrix on every GPU (the reference replicates it per rank, fix_conp.cpp:816-823)
      DevBuf<double> F;
      F.reserve((size_t)N * N);
      if (nr > 0)
        CUDA_CHECK(cudaMemcpy2DAsync(F.p + (size_t)c->r0 * N, (size_t)N * sizeof(double), c->d_mat.p,
                                     c->pitch * sizeof(double), (size_t)N * sizeof(double), nr,
                                     cudaMemcpyDeviceToDevice, s));
      if (c->nranks > 1) {
        std::vector<size_t> bytes(c->nranks), offs(c->nranks);
        for (int r = 0; r < c->nranks; ++r) {
          const int a = std::min(N, r * c->rpr), b = std::min(N, a + c->rpr);
          bytes[r] = (size_t)(b - a) * N * sizeof(double);
          offs[r] = (size_t)a * N * sizeof(double);
        }
        comm_allgatherv(c->comm, F.p + (size_t)c->r0 * N, F.p, bytes.data(), offs.data(), c->rank, c->nranks, s);
        if (c->A_half_band) {  // rows built here hold the half band only (conp_build_A): mirror the rest
          c->launches += launch_gram_mirror(s, N, F.p, N);
          c->A_half_band = false;
        }
      }
      // LU + solve against the identity (LAPACK dgetrf_/dgetri_ in the reference, fix_conp.cpp:947-949)
      cusolverDnParams_t params;
      CUSOLVER_CHECK(cusolverDnCreateParams(&params));
      size_t ws_dev = 0, ws_host = 0;
      CUSOLVER_CHECK(cusolverDnXgetrf_bufferSize(c->solver, params, N, N, CUDA_R_64F, F.p, N, CUDA_R_64F, &ws_dev,
                                                 &ws_host));
      DevBuf<unsigned char> wdev;
      wdev.reserve(std::max<size_t>(ws_dev, 16));
      std::vector<unsigned char> whost(std::max<size_t>(ws_host, 16));
      DevBuf<int64_t> ipiv;
      ipiv.reserve(N);
      DevBuf<int> dinfo;
      dinfo.zero(2, s);
      CUSOLVER_CHECK(cusolverDnXgetrf(c->solver, params, N, N, CUDA_R_64F, F.p, N, ipiv.p, CUDA_R_64F, wdev.p,
                                      ws_dev, whost.data(), ws_host, dinfo.p));
      int info[2] = {0, 0};
      CUDA_CHECK(cudaMemcpyAsync(info, dinfo.p, sizeof(int), cudaMemcpyDeviceToHost, s));
      CUDA_CHECK(cudaStreamSynchronize(s));
      if (info[0] != 0) {
        cusolverDnDestroyParams(params);
        CONP_THROW(CONP_ERR_NUMERIC, "Inversion failed!");
      }
      c->d_fullS.reserve((size_t)N * N);
      c->launches += launch_pad_identity(s, N, c->d_fullS.p, N);
      CUSOLVER_CHECK(cusolverDnXgetrs(c->solver, params, CUBLAS_OP_N, N, N, CUDA_R_64F, F.p, N, ipiv.p, CUDA_R_64F,
                                      c->d_fullS.p, N, dinfo.p + 1));
      CUDA_CHECK(cudaMemcpyAsync(info + 1, dinfo.p + 1, sizeof(int), cudaMemcpyDeviceToHost, s));
      CUDA_CHECK(cudaStreamSynchronize(s));
      cusolverDnDestroyParams(params);
      if (info[1] != 0) CONP_THROW(CONP_ERR_NUMERIC, "Inversion failed!");
      c->launches += 2;
      F.release();
      if (!c->one_electrode) project_full(c, c->d_fullS.p, nullneutral, zneutr);  // fix_conp.cpp:958
      store_rows_from_full(c, c->d_fullS.p);
      CUDA_CHECK(cudaStreamSynchronize(s));
      c->inverted = true;
    }
    // the full copy is only needed for the deferred projection of a one-electrode run
    if (!c->one_electrode) c->d_fullS.release();
    CUDA_CHECK(cudaEventRecord(c->ev[15], s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, c->ev[14], c->ev[15]));
    c->invert_ms = ms;
    if (ee_out) *ee_out = c->ee;
  });
}

int conp_set_unit_voltage(conp_ctx *c, double evscale, const double *q_init, int one_electrode, int nullneutral,
                          int zneutr, double *totsetq_out) {
  return guard(c, [&] {
    need(c->have_A && c->inverted, "conp_set_unit_voltage: matrix not inverted (conp_invert_project)");
    cudaStream_t s = c->stream;
    const int N = c->N;
    c->evscale = evscale;
    c->launches += launch_d_vector(s, N, c->d_ez.p, c->d_eside.p, c->ff_flag, evscale, c->boxlo[2], c->prd[2],
                                   c->d_dvec.p, c->d_setz.p);
    // get_setq: elesetq = S.d (fix_conp.cpp:1090-1096)
    c->launches += enqueue_matvec(c, s, c->d_dvec.p, c->d_setq.p, nullptr);
    if (c->nranks > 1) {
      if (c->sym) comm_allreduce_sum_f64(c->comm, c->d_setq.p, c->vlen, s);
      else comm_allgather(c->comm, c->d_setq.p + c->r0, c->d_setq.p, sizeof(double) * c->rpr, s);
    }
    std::vector<double> setq(N), setz(N);
    CUDA_CHECK(cudaMemcpyAsync(setq.data(), c->d_setq.p, sizeof(double) * N, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaMemcpyAsync(setz.data(), c->d_setz.p, sizeof(double) * N, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    // with the projection on, e^T S = 0 up to rounding: take that residual off S.d (see charge_epilogue)
    // (only for a matrix projected here: an `inv` file is used exactly as given; not for a single electrode,
    // whose S.d comes from the unprojected inverse and does not sum to zero, fix_conp.cpp:1090-1115)
    c->neutral_polish = nullneutral != 0 && !one_electrode && c->projected_here;
    c->n_left = 0;
    c->sum_setz = 0;
    double tot = 0, zOAz = 0;
    for (int i = 0; i < N; ++i) {
      if (c->h_side[i] == 1) { tot += setq[i]; c->n_left++; }  // :1098-1104
      zOAz += setq[i] * setz[i];                                  // fix_cond.cpp:62
      c->sum_setz += setz[i];
    }
    if (c->neutral_polish) {
      double mean = 0;
      for (int i = 0; i < N; ++i) mean += setq[i];
      mean /= N;
      for (int i = 0; i < N; ++i) setq[i] -= mean;
      CUDA_CHECK(cudaMemcpyAsync(c->d_setq.p, setq.data(), sizeof(double) * N, cudaMemcpyHostToDevice, s));
      CUDA_CHECK(cudaStreamSynchronize(s));
    }
    c->totsetq = tot;
    c->dd = -tot;
    // cond_setup2, fix_cond.cpp:57-68
    {
      const double lz = c->prd[2], Axy = c->prd[0] * c->prd[1];
      double vmult = 4 * MY_PI * zOAz * lz / (evscale * Axy);
      vmult /= 1 + vmult;
      vmult /= zOAz;
      c->vmult = vmult;
    }
    c->have_qinit = q_init != nullptr;
    if (q_init) c->d_qinit.upload(q_init, N, s);
    if (one_electrode && c->d_fullS.p) {  // fix_conp.cpp:1115
      project_full(c, c->d_fullS.p, nullneutral, zneutr);
      store_rows_from_full(c, c->d_fullS.p);
      CUDA_CHECK(cudaStreamSynchronize(s));
      c->d_fullS.release();
    }
    CUDA_CHECK(cudaStreamSynchronize(s));
    if (totsetq_out) *totsetq_out = tot;
    drop_graphs(c);
    c->have_setq = true;
  });
}

// ---- per step --------------------------------------------------------------------

int conp_post_neighbor(conp_ctx *c, int nlocal, const double *q, const int *type, const int *mask, int ele_bits) {
  return guard(c, [&] {
    need(c->have_ele && c->have_pair, "conp_post_neighbor: setup incomplete");
    if (nlocal < 0 || (nlocal > 0 && (!q || !type))) CONP_THROW(CONP_ERR_ARG, "conp_post_neighbor: bad arrays");
    cudaStream_t s = c->stream;
    c->nlocal = nlocal;
    // charged non-electrode atoms; uncharged ones contribute exact zeros to every
    // term of b (km_ewald.cpp:686, pppm_conp.cpp:161, fix_conp.cpp:1339)
    c->h_idx.clear();
    for (int i = 0; i < nlocal; ++i) {
      if (mask && (mask[i] & ele_bits)) continue;
      if (type[i] < 1 || type[i] > c->ntypes) CONP_THROW(CONP_ERR_ARG, "conp_post_neighbor: atom type out of range");
      if (q[i] != 0.0) c->h_idx.push_back(i);
    }
    c->m_local = (int)c->h_idx.size();
    c->m_counts.assign(c->nranks, 0);
    c->m_offsets.assign(c->nranks, 0);
    c->m_counts[c->rank] = c->m_local;
    c->qsum_elyte = 0.0;
    for (int j : c->h_idx) c->qsum_elyte += q[j];
    if (c->nranks > 1) {
      DevBuf<double> cnt;
      cnt.zero(c->nranks + 1, s);
      const double mine[2] = {(double)c->m_local, c->qsum_elyte};
      CUDA_CHECK(cudaMemcpyAsync(cnt.p + c->rank, &mine[0], sizeof(double), cudaMemcpyHostToDevice, s));
      CUDA_CHECK(cudaMemcpyAsync(cnt.p + c->nranks, &mine[1], sizeof(double), cudaMemcpyHostToDevice, s));
      comm_allreduce_sum_f64(c->comm, cnt.p, c->nranks + 1, s);
      std::vector<double> h(c->nranks + 1);
      CUDA_CHECK(cudaMemcpyAsync(h.data(), cnt.p, sizeof(double) * (c->nranks + 1), cudaMemcpyDeviceToHost, s));
      CUDA_CHECK(cudaStreamSynchronize(s));
      for (int r = 0; r < c->nranks; ++r) c->m_counts[r] = (int)std::llround(h[r]);
      c->qsum_elyte = h[c->nranks];
    }
    int tot = 0, cmax = 0;
    for (int r = 0; r < c->nranks; ++r) { tot += c->m_counts[r]; cmax = std::max(cmax, c->m_counts[r]); }
    c->m_total = tot;
    if (c->nranks > 1) {
      // + the slot that carries sum(q z); ~6 % head-room rounded to 1024 slots so that the arena (whose size
      // depends on mpad) survives the small count changes of an ordinary reneighbouring
      c->mpad = (int)round_up((size_t)cmax + 1 + (size_t)cmax / 16, 1024);
      c->m_slots = c->nranks * c->mpad;
      for (int r = 0; r < c->nranks; ++r) c->m_offsets[r] = r * c->mpad;
    } else {
      c->mpad = std::max(c->m_local, 1);
      c->m_slots = c->m_local;
      c->m_offsets[0] = 0;
    }
    c->d_mcounts.upload(c->m_counts, s);
    c->d_qraw.upload(q, nlocal, s);
    c->d_typeraw.upload(type, nlocal, s);
    c->d_idx.upload(c->h_idx, s);
    c->d_xraw.reserve(3 * (size_t)std::max(nlocal, 1));
    const size_t m = std::max(c->m_slots, 1);
    ensure_p2p(c);
    if (c->p2p) CUDA_CHECK(cudaMemsetAsync(c->d_packed.p, 0, sizeof(PosQ) * m, s));
    else c->d_packed.zero(m, s);
    c->d_ptype.zero(m, s);
    c->d_sorted.reserve(m); c->d_sortedf.reserve(m);
    c->d_stype.reserve(m); c->d_ssrc.reserve(m);
    c->d_cellof.reserve(m); c->d_slot.reserve(m);
    c->d_nearlist.reserve(m);
    c->d_sendcnt.zero(std::max(c->nranks, 1), s);
    c->d_own.reserve((size_t)std::max(c->m_local, 1));
    if (c->nranks > 1 && !(c->p2p && c->routed)) {  // static per-charge data: types of every rank's block
      std::vector<int> ht(c->mpad, 0);
      for (int j = 0; j < c->m_local; ++j) ht[j] = type[c->h_idx[j]];
      int *own = c->d_ptype.p + c->m_offsets[c->rank];
      CUDA_CHECK(cudaMemcpyAsync(own, ht.data(), sizeof(int) * c->mpad, cudaMemcpyHostToDevice, s));
      comm_allgather(c->comm, own, c->d_ptype.p, sizeof(int) * (size_t)c->mpad, s);
      CUDA_CHECK(cudaStreamSynchronize(s));
    }
    ensure_static_cells(c);
    // Which spread: measured on a B200 (profiles/r02_spread_kernels.md) the owner-computes tensor-core kernels
    // win from a few 1e5 charges per GPU upward (cfg5 on one GPU: 500 000), the red.global kernel below
    // (cfg4; cfg5 on >= 2 GPUs): fewer fixed costs -- no second sort, no per-step stencil pass.
    c->spread_atomic = c->spread_mode == 0 ||
                       (c->spread_mode < 0 && (long long)c->m_total / c->nranks < 300000LL);
    if (c->have_pppm && (c->splan.use_mma || c->swplan.usable)) {  // per-charge stencil data of the tensor-core spreads
      const size_t mm = (size_t)std::max(c->m_slots, 1);
      c->d_sp_origin.reserve(mm);
      c->d_sp_weights.reserve(mm * 3 * c->pg.order);
      c->splan.origin = c->d_sp_origin.p;
      c->splan.weights = c->d_sp_weights.p;
      c->splan.wstride = mm;
      c->d_sw_binof.reserve(mm);
      c->d_sw_slot.reserve(mm);
      c->d_sw_records.reserve(mm * (size_t)((3 * c->pg.order + 2) & ~1));
      c->swplan.bin_of = c->d_sw_binof.p;
      c->swplan.slot = c->d_sw_slot.p;
      c->swplan.records = c->d_sw_records.p;
    }
    c->d_cellcount.zero((size_t)c->grid_b.ncells + 8, s);
    c->d_cellstart.zero((size_t)c->grid_b.ncells + 8, s);
    drop_graphs(c);
    CUDA_CHECK(cudaStreamSynchronize(s));
    c->have_atoms = true;
    c->eg_valid = false;
  });
}

int conp_solve_device(conp_ctx *c, const double *x_device, int kspace_mode, int variant, double value) {
  return guard(c, [&] { solve_device(c, x_device, kspace_mode, variant, value); });
}

int conp_get_charges(conp_ctx *c, double *q_ele_out, double *scalar_out) {
  return guard(c, [&] {
    need(c->solved, "conp_get_charges: no solve yet");
    double sc[2] = {0, 0};
    if (q_ele_out)
      CUDA_CHECK(cudaMemcpyAsync(q_ele_out, c->d_q.p, sizeof(double) * c->N, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(sc, c->scal(0), sizeof(sc), cudaMemcpyDeviceToHost, c->stream));
    check_range_flag(c);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (scalar_out) *scalar_out = sc[0];
  });
}

int conp_pre_force(conp_ctx *c, const double *x, int kspace_mode, int variant, double value, double *q_ele_out,
                   double *scalar_out) {
  return guard(c, [&] {
    need(c->have_atoms, "conp_pre_force: conp_post_neighbor not called");
    if (c->nlocal > 0 && !x) CONP_THROW(CONP_ERR_ARG, "conp_pre_force: x is NULL");
    if (c->nlocal > 0)
      CUDA_CHECK(cudaMemcpyAsync(c->d_xraw.p, x, sizeof(double) * 3 * (size_t)c->nlocal, cudaMemcpyHostToDevice,
                                 c->stream));
    solve_device(c, c->d_xraw.p, kspace_mode, variant, value);
    double sc[2] = {0, 0};
    if (q_ele_out)
      CUDA_CHECK(cudaMemcpyAsync(q_ele_out, c->d_q.p, sizeof(double) * c->N, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(sc, c->scal(0), sizeof(sc), cudaMemcpyDeviceToHost, c->stream));
    int flag = 0;
    if (kspace_mode == CONP_KSPACE_PPPM)
      CUDA_CHECK(cudaMemcpyAsync(&flag, c->d_flag.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (flag) CONP_THROW(CONP_ERR_RANGE, "Out of range atoms - cannot compute PPPM");
    if (c->p2p && p2p_error(c->p2p)) CONP_THROW(CONP_ERR_COMM, "peer-to-peer exchange timed out (a rank is missing)");
    if (scalar_out) *scalar_out = sc[0];
  });
}

int conp_get_b(conp_ctx *c, double *b_out, double *b_kspace_out) {
  return guard(c, [&] {
    need(c->solved, "conp_get_b: no solve yet");
    cudaStream_t s = c->stream;
    if (c->nranks > 1) comm_allgather(c->comm, c->d_bk.p + c->r0, c->d_bk.p, sizeof(double) * c->rpr, s);
    if (b_out) CUDA_CHECK(cudaMemcpyAsync(b_out, c->d_b.p, sizeof(double) * c->N, cudaMemcpyDeviceToHost, s));
    if (b_kspace_out)
      CUDA_CHECK(cudaMemcpyAsync(b_kspace_out, c->d_bk.p, sizeof(double) * c->N, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
  });
}

int conp_get_density_region(conp_ctx *c, int which, const int lo[3], const int hi[3], double *out);

// whole periodic mesh (diagnostic / single-rank hosts): the region hand-off with lo = 0, hi = n - 1
int conp_get_density(conp_ctx *c, int which, double *brick_out) {
  if (!c) return CONP_ERR_ARG;
  const int lo[3] = {0, 0, 0};
  const int hi[3] = {c->pg.nx - 1, c->pg.ny - 1, c->pg.nz - 1};
  if (!c->have_pppm) {
    c->err = "conp_get_density: no PPPM solve yet";
    return CONP_ERR_STATE;
  }
  return conp_get_density_region(c, which, lo, hi, brick_out);
}

int conp_get_density_region(conp_ctx *c, int which, const int lo[3], const int hi[3], double *out) {
  return guard(c, [&] {
    need(c->have_pppm && c->solved, "conp_get_density_region: no PPPM solve yet");
    if (which < 0 || which > 2 || !lo || !hi || !out) CONP_THROW(CONP_ERR_ARG, "conp_get_density_region: bad arguments");
    const PPPMGeom &g = c->pg;
    const int n[3] = {g.nx, g.ny, g.nz};
    for (int a = 0; a < 3; ++a)
      if (lo[a] < 0 || hi[a] < lo[a] || hi[a] >= n[a])
        CONP_THROW(CONP_ERR_ARG, "conp_get_density_region: region outside the mesh");
    cudaStream_t s = c->stream;
    const double *elyte = c->d_brick.p, *ele = c->d_ebrick.p;
    if (c->nranks > 1) {
      // once per solve: every rank's z-slab of the electrolyte density -> all nzi planes; the electrode
      // density is the sum of the ranks' own rows
      if (!c->density_gathered) {
        c->d_brick_all.reserve((size_t)std::max(g.nzi, 1) * c->plane);
        c->d_ebrick_all.reserve((size_t)std::max(g.nzo, 1) * c->plane);
        std::vector<size_t> bytes(c->nranks), offs(c->nranks);
        for (int r = 0; r < c->nranks; ++r) {
          const int z0 = (int)(((long long)g.nzi * r) / c->nranks), z1 = (int)(((long long)g.nzi * (r + 1)) / c->nranks);
          bytes[r] = (size_t)(z1 - z0) * c->plane * sizeof(double);
          offs[r] = (size_t)z0 * c->plane * sizeof(double);
        }
        comm_allgatherv(c->comm, c->d_brick.p, c->d_brick_all.p, bytes.data(), offs.data(), c->rank, c->nranks, s);
        CUDA_CHECK(cudaMemcpyAsync(c->d_ebrick_all.p, c->d_ebrick.p, sizeof(double) * (size_t)g.nzo * c->plane,
                                   cudaMemcpyDeviceToDevice, s));
        comm_allreduce_sum_f64(c->comm, c->d_ebrick_all.p, (size_t)g.nzo * c->plane, s);
        c->density_gathered = true;
      }
      elyte = c->d_brick_all.p;
      ele = c->d_ebrick_all.p;
    }
    const size_t cnt = (size_t)(hi[0] - lo[0] + 1) * (hi[1] - lo[1] + 1) * (hi[2] - lo[2] + 1);
    c->d_region.reserve(cnt);
    c->launches += launch_region_gather(s, g, which, lo, hi, elyte, ele, c->d_region.p);
    CUDA_CHECK(cudaMemcpyAsync(out, c->d_region.p, sizeof(double) * cnt, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
  });
}

int conp_get_potential_brick(conp_ctx *c, double *brick_out) {
  return guard(c, [&] {
    need(c->have_pppm && c->solved, "conp_get_potential_brick: no PPPM solve yet");
    DevBuf<double> full;
    full_mesh_potential(c, false, full);
    CUDA_CHECK(cudaMemcpyAsync(brick_out, full.p, sizeof(double) * c->ngrid, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
  });
}

int conp_mesh_potential(conp_ctx *c, int n, const double *xyz, double *u_out) {
  return guard(c, [&] {
    if (!c->have_pppm) CONP_THROW(CONP_ERR_STATE, "Compute requires a compatible KSpace provider like pppm/conp");
    need(c->solved, "conp_mesh_potential: no solve yet");
    if (n < 0 || (n > 0 && (!xyz || !u_out))) CONP_THROW(CONP_ERR_ARG, "conp_mesh_potential: bad arguments");
    if (n == 0) return;
    cudaStream_t s = c->stream;
    DevBuf<double> full, dx, dout;
    full_mesh_potential(c, true, full);
    dx.upload(xyz, 3 * (size_t)n, s);
    dout.reserve(n);
    c->launches += launch_mesh_potential(s, c->pg, c->d_rho.p, n, dx.p, full.p, dout.p);
    CUDA_CHECK(cudaMemcpyAsync(u_out, dout.p, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
  });
}

int conp_electrode_potential(conp_ctx *c, int pairflag, int kspaceflag, double eta, int qsumflag, double *phi_out) {
  return guard(c, [&] {
    need(c->solved, "conp_electrode_potential: no solve yet");
    if (!phi_out) CONP_THROW(CONP_ERR_ARG, "conp_electrode_potential: null output");
    if (kspaceflag && !c->have_pppm)  // compute_potential_atom.cpp:111-113
      CONP_THROW(CONP_ERR_STATE, "Compute requires a compatible KSpace provider like pppm/conp");
    cudaStream_t s = c->stream;
    const int N = c->N, nr = c->r1 - c->r0;
    std::vector<double> phi(N, 0.0), q(N);
    CUDA_CHECK(cudaMemcpyAsync(q.data(), c->d_q.p, sizeof(double) * N, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    if (pairflag) {
      // compute_pair_potential (:223-318): every pair inside cutsq and cut_coulsq, erfc(g r)/r minus the
      // Gaussian term erfc(eta r)/r (one electrode atom) or erfc(eta r / sqrt 2)/r (two); no `etypes` filter
      const int n1 = c->ntypes + 1;
      double cut_coulsq = c->cut_coul * c->cut_coul;
      const double cut_erfc = ERFC_MAX * ERFC_MAX / (c->g_ewald * c->g_ewald);
      if (cut_coulsq > cut_erfc) cut_coulsq = cut_erfc;
      std::vector<double> call((size_t)n1 * n1, 0.0);
      double mx = 0.0;
      for (int it = 1; it <= c->ntypes; ++it)
        for (int jt = 1; jt <= c->ntypes; ++jt) {
          call[(size_t)it * n1 + jt] = std::min(c->h_cutsq[(size_t)it * n1 + jt], cut_coulsq);
          mx = std::max(mx, call[(size_t)it * n1 + jt]);
        }
      DevBuf<double> dcut, dp1, dp2;
      dcut.upload(call, s);
      dp1.zero(c->vlen, s);
      dp2.zero(c->vlen, s);
      PairTables pt = pair_tables(c, dcut.p);
      pt.pairmode = CONP_PAIR_ETA;  // the compute knows one Gaussian width only
      pt.eta = eta;
      const double rc = std::sqrt(mx);
      if (c->nranks > 1 && rc > c->rc_b * (1.0 + 1e-12))
        CONP_THROW(CONP_ERR_STATE, "conp_electrode_potential: on several GPUs the pair cut-off of the compute (%g) "
                   "must not exceed the fix's (%g): charges beyond it are not exchanged", rc, c->rc_b);
      if (rc > 0.0 && nr > 0) {
        // electrolyte -> electrode: the step's cell-sorted charges, traversal lists rebuilt for this radius
        if (c->m_total > 0) {
          CellGrid g = c->grid_b;
          g.rc = rc;
          for (int a = 0; a < 3; ++a) g.smax[a] = g.periodic[a] ? (int)std::ceil(g.rc / g.prd[a]) + 1 : 0;
          std::vector<int> run_start;
          std::vector<PairRun> runs;
          build_pair_runs(g, c->r0, c->r1, c->h_xyz.data(), run_start, runs);
          DevBuf<int> drs;
          DevBuf<PairRun> druns;
          drs.upload(run_start, s);
          druns.upload(runs, s);
          c->launches += launch_pair_b(s, g, pt, c->r0, c->r1, c->d_ex.p, c->d_ey.p, c->d_ez.p, c->d_etype.p, drs.p,
                                       druns.p, c->d_sorted.p, c->d_stype.p, c->d_sortedf.p, c->d_cellstart.p, dp1.p);
          CUDA_CHECK(cudaStreamSynchronize(s));
        }
        // electrode -> electrode (all images)
        CellGrid g = make_cell_grid(c->boxlo, c->prd, c->periodic, rc);
        std::vector<EPos> sorted;
        std::vector<int> cs, run_start;
        std::vector<PairRun> runs;
        build_electrode_cells(g, 0, N, c->h_xyz.data(), c->h_type.data(), sorted, cs);
        build_pair_runs(g, c->r0, c->r1, c->h_xyz.data(), run_start, runs);
        DevBuf<EPos> dsorted;
        DevBuf<int> dcs, drs;
        DevBuf<PairRun> druns;
        dsorted.upload(sorted, s);
        dcs.upload(cs, s);
        drs.upload(run_start, s);
        druns.upload(runs, s);
        c->launches += launch_pair_P(s, g, pt, dsorted.p, dcs.p, c->r0, c->r1, c->d_ex.p, c->d_ey.p, c->d_ez.p,
                                     c->d_etype.p, drs.p, druns.p, c->d_q.p, dp2.p);
        CUDA_CHECK(cudaStreamSynchronize(s));
      }
      if (c->nranks > 1) {
        comm_allgather(c->comm, dp1.p + c->r0, dp1.p, sizeof(double) * c->rpr, s);
        comm_allgather(c->comm, dp2.p + c->r0, dp2.p, sizeof(double) * c->rpr, s);
      }
      std::vector<double> p1(N), p2(N);
      CUDA_CHECK(cudaMemcpyAsync(p1.data(), dp1.p, sizeof(double) * N, cudaMemcpyDeviceToHost, s));
      CUDA_CHECK(cudaMemcpyAsync(p2.data(), dp2.p, sizeof(double) * N, cudaMemcpyDeviceToHost, s));
      CUDA_CHECK(cudaStreamSynchronize(s));
      for (int i = 0; i < N; ++i) phi[i] += -p1[i] + p2[i];  // launch_pair_b returns -sum q dudq
    }
    if (kspaceflag) {
      // potential[i] -= compute_particle_potential(i) = -(mesh sum) + 2 g q_i / sqrt(pi)  (pppm_conp.cpp:452-488),
      // + eta q_i sqrt(2)/sqrt(pi) for the Gaussian (electrode) atoms (:168)
      DevBuf<double> full, dout;
      full_mesh_potential(c, true, full);
      dout.reserve(N);
      c->launches += launch_mesh_potential(s, c->pg, c->d_rho.p, N, c->d_exyz.p, full.p, dout.p);
      std::vector<double> u(N);
      double qz = 0.0;
      CUDA_CHECK(cudaMemcpyAsync(u.data(), dout.p, sizeof(double) * N, cudaMemcpyDeviceToHost, s));
      CUDA_CHECK(cudaMemcpyAsync(&qz, c->scal(2), sizeof(double), cudaMemcpyDeviceToHost, s));
      CUDA_CHECK(cudaStreamSynchronize(s));
      for (int i = 0; i < N; ++i) {
        phi[i] += u[i] - 2.0 * c->g_ewald * q[i] / MY_PIS;
        if (eta != 0.0) phi[i] += eta * q[i] * std::sqrt(2.0) / MY_PIS;
      }
      if (c->slabflag) {  // ComputePotentialAtom::slabcorr (:333-358): all atoms, electrode and electrolyte
        const double volume = c->prd[0] * c->prd[1] * c->prd[2] * c->slab_volfactor;
        const double pi2vol = 2.0 * MY_PI / volume;
        double dip = qz, qsum = c->qsum_elyte;
        for (int i = 0; i < N; ++i) { dip += q[i] * c->h_xyz[3 * i + 2]; qsum += q[i]; }
        const double slabcorr = 2.0 * pi2vol * dip;
        for (int i = 0; i < N; ++i) {
          const double z = c->h_xyz[3 * i + 2];
          phi[i] += z * slabcorr;
          if (qsumflag) phi[i] -= pi2vol * qsum * z * z;
        }
      }
    }
    std::memcpy(phi_out, phi.data(), sizeof(double) * N);
  });
}

int conp_post_force(conp_ctx *c, double qqrd2e, double *f_out, double *energies_out) {
  return guard(c, [&] {
    need(c->solved, "conp_post_force: no solve yet");
    cudaStream_t s = c->stream;
    const int N = c->N;
    CUDA_CHECK(cudaMemsetAsync(c->scal(4), 0, sizeof(double) * 8, s));
    c->d_fpacked.zero(3 * (size_t)std::max(c->m_slots, 1), s);
    if (c->rc_f > 0.0 && c->m_total > 0 && c->r1 > c->r0) {
      CellGrid g = c->grid_b;  // same electrode binning, smaller search radius
      g.rc = c->rc_f;
      for (int a = 0; a < 3; ++a) g.smax[a] = g.periodic[a] ? (int)std::ceil(g.rc / g.prd[a]) + 1 : 0;
      CUDA_CHECK(cudaMemsetAsync(c->d_nearcount.p, 0, sizeof(int), s));
      const bool routed = c->p2p && c->routed && c->have_relevant;
      const int *counts = routed ? (const int *)(p2p_local(c->p2p) + c->off_rcnt) : nullptr;
      c->launches += launch_near_list(s, c->grid_b, c->m_slots, c->d_packed.p, c->d_nearmask.p, c->d_nearlist.p,
                                      c->d_nearcount.p, counts, c->mpad);
      c->launches += launch_pair_postforce(s, g, pair_tables(c, c->d_cuteff_b.p), qqrd2e, c->d_esorted.p,
                                           c->d_ecellstart.p, c->d_q.p, c->d_packed.p, c->d_ptype.p,
                                           c->d_nearlist.p, c->d_nearcount.p, c->m_slots, c->d_cutsq_listed.p,
                                           c->d_fpacked.p, c->scal(4), c->num_sms,
                                           routed ? c->d_psrc.p : nullptr, c->mpad);
    }
    if (c->nranks > 1) {
      comm_allreduce_sum_f64(c->comm, c->scal(4), 8, s);
      if (f_out) comm_allreduce_sum_f64(c->comm, c->d_fpacked.p, 3 * (size_t)c->m_slots, s);
    }
    std::vector<double> q(N), en(8), fp;
    CUDA_CHECK(cudaMemcpyAsync(q.data(), c->d_q.p, sizeof(double) * N, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaMemcpyAsync(en.data(), c->scal(4), sizeof(double) * 8, cudaMemcpyDeviceToHost, s));
    if (f_out && c->m_local > 0) {
      fp.resize(3 * (size_t)c->m_local);
      CUDA_CHECK(cudaMemcpyAsync(fp.data(), c->d_fpacked.p + 3 * (size_t)c->m_offsets[c->rank],
                                 sizeof(double) * fp.size(), cudaMemcpyDeviceToHost, s));
    }
    CUDA_CHECK(cudaStreamSynchronize(s));
    // self energy, fix_conp.cpp:1166-1199
    double eself = 0;
    if (c->pairmode == CONP_PAIR_ETA) {
      double eleqsqsum = 0;
      for (int i = 0; i < N; ++i) eleqsqsum += q[i] * q[i];
      eself = qqrd2e * c->eta * eleqsqsum / (std::sqrt(2.0) * MY_PIS);
    } else {
      double u0qsqsum = 0;
      for (int i = 0; i < N; ++i) u0qsqsum += c->h_u0[c->h_type[i]] * q[i] * q[i];
      eself = qqrd2e * u0qsqsum;
    }
    en[1] = eself;
    if (energies_out) std::memcpy(energies_out, en.data(), sizeof(double) * 8);
    if (f_out) {
      std::memset(f_out, 0, sizeof(double) * 3 * (size_t)c->nlocal);
      for (int j = 0; j < c->m_local; ++j) {
        const int i = c->h_idx[j];
        f_out[3 * (size_t)i] = fp[3 * (size_t)j];
        f_out[3 * (size_t)i + 1] = fp[3 * (size_t)j + 1];
        f_out[3 * (size_t)i + 2] = fp[3 * (size_t)j + 2];
      }
    }
  });
}

// ---- instrumentation -------------------------------------------------------------

void *conp_stream(conp_ctx *c) { return c ? (void *)c->stream : nullptr; }

int conp_sync(conp_ctx *c) {
  return guard(c, [&] { CUDA_CHECK(cudaStreamSynchronize(c->stream)); });
}

int conp_timer_record(conp_ctx *c, int slot) {
  return guard(c, [&] {
    if (slot < 0 || slot >= 14) CONP_THROW(CONP_ERR_ARG, "timer slot out of range");
    CUDA_CHECK(cudaEventRecord(c->ev[slot], c->stream));
  });
}

int conp_timer_elapsed_ms(conp_ctx *c, int a, int b, float *ms_out) {
  return guard(c, [&] {
    if (a < 0 || a >= 14 || b < 0 || b >= 14 || !ms_out) CONP_THROW(CONP_ERR_ARG, "timer slot out of range");
    CUDA_CHECK(cudaEventSynchronize(c->ev[b]));
    CUDA_CHECK(cudaEventElapsedTime(ms_out, c->ev[a], c->ev[b]));
  });
}

int conp_stage_times(conp_ctx *c, int enable, double *out8) {
  if (!c) return 0;
  const int n = c->stage_n;
  if (out8)
    for (int i = 0; i < NSTAGE; ++i) out8[i] = n ? c->stage_ms[i] / n : 0.0;
  if (n && c->debug)
    fprintf(stderr, "[conp] rank %d k-space stage (ms): spread %.4f fft %.4f zconv %.4f all-reduce %.4f ifft %.4f\n",
            c->rank, c->kev_ms[0] / n, c->kev_ms[1] / n, c->kev_ms[2] / n, c->kev_ms[3] / n, c->kev_ms[4] / n);
  if (c->trace) {  // the graph replays so far: mean offset of every mark from the step's start
    unsigned long long h[48];
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    CUDA_CHECK(cudaMemcpy(h, c->d_trace.p, sizeof(h), cudaMemcpyDeviceToHost));
    if (h[40]) {
      static const char *name[16] = {"begin", "packed", "binned", "pair", "kspace", "gathered", "b-exchanged",
                                     "matvec", "epilogue", "", "k:begin", "k:spread", "k:fft", "k:zconv",
                                     "k:all-reduce", "k:ifft"};
      char line[1024];
      int len = snprintf(line, sizeof(line), "[conp] rank %d graph timeline over %llu steps (us after begin):", c->rank,
                         h[40]);
      for (int i = 0; i < 16 && len < (int)sizeof(line) - 40; ++i)
        if (name[i][0] && (i == 0 || h[1 + i]))
          len += snprintf(line + len, sizeof(line) - len, " %s %.1f", name[i], 1e-3 * (double)h[1 + i] / (double)h[40]);
      fprintf(stderr, "%s\n", line);  // one write: the ranks' lines must not interleave
    }
    CUDA_CHECK(cudaMemset(c->d_trace.p, 0, sizeof(h)));
  }
  c->stage_timing = enable != 0;
  for (auto &v : c->kev_ms) v = 0;
  for (auto &v : c->stage_ms) v = 0;
  c->stage_n = 0;
  return n;
}

int conp_bench_gemv(conp_ctx *c, int reps, float *ms_per_rep_out) {
  return guard(c, [&] {
    need(c->have_A, "conp_bench_gemv: no matrix resident");
    cudaStream_t s = c->stream;
    // the matvec the step would launch (symmetric or general), into scratch so S.b stays intact
    DevBuf<double> scratch;
    scratch.zero(c->vlen, s);
    enqueue_matvec(c, s, c->d_b.p, scratch.p, nullptr);
    CUDA_CHECK(cudaEventRecord(c->ev[12], s));
    for (int r = 0; r < reps; ++r) c->launches += enqueue_matvec(c, s, c->d_b.p, scratch.p, nullptr);
    CUDA_CHECK(cudaEventRecord(c->ev[13], s));
    CUDA_CHECK(cudaEventSynchronize(c->ev[13]));
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, c->ev[12], c->ev[13]));
    if (ms_per_rep_out) *ms_per_rep_out = ms / std::max(reps, 1);
  });
}

int conp_matvec(conp_ctx *c, const double *v, double *out) {
  return guard(c, [&] {
    need(c->have_A && c->have_ele, "conp_matvec: no matrix resident");
    if (!v || !out) CONP_THROW(CONP_ERR_ARG, "conp_matvec: null vector");
    cudaStream_t s = c->stream;
    DevBuf<double> in, res;
    in.zero(c->vlen, s);
    res.zero(c->vlen, s);
    CUDA_CHECK(cudaMemcpyAsync(in.p, v, sizeof(double) * c->N, cudaMemcpyHostToDevice, s));
    c->launches += enqueue_matvec(c, s, in.p, res.p, nullptr);
    if (c->nranks > 1) {
      if (c->sym) comm_allreduce_sum_f64(c->comm, res.p, c->vlen, s);
      else comm_allgather(c->comm, res.p + c->r0, res.p, sizeof(double) * c->rpr, s);
    }
    CUDA_CHECK(cudaMemcpyAsync(out, res.p, sizeof(double) * c->N, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
  });
}

int conp_plan_symv(int n, int row0, int nrows, int num_sms, int max_strips, int *strips_out, int *nstrips_out,
                   int *slice_len_out) {
  if (!nstrips_out || !slice_len_out || num_sms < 1) return CONP_ERR_ARG;
  std::vector<int2> strips;
  const SymvPlan p = plan_symv(n, row0, nrows, num_sms, strips);
  *nstrips_out = p.usable ? p.nstrips : -1;
  *slice_len_out = p.usable ? p.L : 0;
  if (strips_out)
    for (int s = 0; s < (int)strips.size() && s < max_strips; ++s) {
      strips_out[2 * s] = strips[s].x;
      strips_out[2 * s + 1] = strips[s].y;
    }
  return CONP_OK;
}

int conp_row_block(int n_ele, int nranks, int rank, int *row_begin, int *row_end, int *rows_per_rank) {
  if (n_ele < 0 || nranks < 1 || rank < 0 || rank >= nranks) return CONP_ERR_ARG;
  // equal contiguous row blocks, multiple of 16 rows so every block start is 128-byte aligned
  const int rpr = (int)round_up((size_t)(n_ele + nranks - 1) / nranks, 16);
  const int r0 = std::min(n_ele, rank * rpr);
  if (row_begin) *row_begin = r0;
  if (row_end) *row_end = std::min(n_ele, r0 + rpr);
  if (rows_per_rank) *rows_per_rank = rpr;
  return CONP_OK;
}

int conp_plan_spread(const int mesh[3], int order, double shift, const double boxlo[3], const double prd[3],
                     const int periodic[3], double slab_volfactor, double rc, int zin_lo, int nzi, int zs_lo,
                     int zs_n, int num_sms, int *geom_out, int *run_start_out, int max_tiles, int *runs_out,
                     int max_runs, int *ntiles_out, int *nruns_out) {
  if (!mesh || !boxlo || !prd || !periodic || !geom_out || !ntiles_out || !nruns_out || order < 1 || order > 7)
    return CONP_ERR_ARG;
  PPPMGeom g;
  std::memset(&g, 0, sizeof(g));
  g.nx = mesh[0]; g.ny = mesh[1]; g.nz = mesh[2]; g.order = order;
  g.nlower = -((order - 1) / 2);
  const double prd_slab[3] = {prd[0], prd[1], prd[2] * slab_volfactor};
  for (int a = 0; a < 3; ++a) { g.boxlo[a] = boxlo[a]; g.delinv[a] = mesh[a] / prd_slab[a]; }
  g.shift = shift;
  g.zin_lo = zin_lo; g.nzi = nzi; g.zs_lo = zs_lo; g.zs_n = zs_n;
  const CellGrid cells = make_cell_grid(boxlo, prd, periodic, rc);
  std::vector<int> rs;
  std::vector<int2> rr;
  SpreadPlan sp;
  plan_pppm_spread_tiles(g, cells, num_sms, rs, rr, sp);
  const int geom[12] = {sp.tz, sp.ty, sp.tx, sp.ntz, sp.nty, sp.ntx, sp.halo_z, sp.halo_y, sp.halo_x,
                        cells.nc[0], cells.nc[1], cells.nc[2]};
  std::memcpy(geom_out, geom, sizeof(geom));
  *ntiles_out = sp.ntiles;
  *nruns_out = (int)rr.size();
  if (run_start_out)
    for (int i = 0; i < (int)rs.size() && i <= max_tiles; ++i) run_start_out[i] = rs[i];
  if (runs_out)
    for (int i = 0; i < (int)rr.size() && i < max_runs; ++i) { runs_out[2 * i] = rr[i].x; runs_out[2 * i + 1] = rr[i].y; }
  return CONP_OK;
}

int conp_plan_sweep(const int mesh[3], int order, int nzi, int zs_lo, int zs_n, int num_sms, int *geom_out,
                    int *items_out, int max_items, int *nitems_out) {
  if (!mesh || !geom_out || !nitems_out || order < 1 || order > 7 || nzi < 0 || nzi > mesh[2] || zs_lo < 0 ||
      zs_n < 0 || zs_lo + zs_n > nzi || num_sms < 1)
    return CONP_ERR_ARG;
  PPPMGeom g;
  std::memset(&g, 0, sizeof(g));
  g.nx = mesh[0]; g.ny = mesh[1]; g.nz = mesh[2]; g.order = order;
  g.nzi = nzi; g.zs_lo = zs_lo; g.zs_n = zs_n;
  std::vector<int4> items;
  SweepPlan sp;
  plan_pppm_sweep(g, num_sms, items, sp);
  const int geom[8] = {sp.usable ? 1 : 0, sp.ncolx, sp.ncoly, sp.pz_lo, sp.npz, sp.wrap_z, sp.nbins, sp.grid};
  std::memcpy(geom_out, geom, sizeof(geom));
  *nitems_out = (int)items.size();
  if (items_out)
    for (int i = 0; i < (int)items.size() && i < max_items; ++i) {
      items_out[3 * i] = items[i].x; items_out[3 * i + 1] = items[i].y; items_out[3 * i + 2] = items[i].z;
    }
  return CONP_OK;
}

int conp_plan_zconv(int ncol, int nz, int nzi, int zs_lo, int nzl, int zin_lo, const int *krad, int nzo,
                    const int *zout, int real_kernel, int *groups_out, int max_groups, int *wide_out, int max_wide,
                    int *aout_out, int *caps_out, int *ngroups_out, int *nwide_out) {
  if (ncol <= 0 || nz <= 0 || nzi <= 0 || nzi > nz || zs_lo < 0 || nzl < 0 || zs_lo + nzl > nzi || !krad || nzo < 0 ||
      (nzo > 0 && !zout) || !caps_out || !ngroups_out || !nwide_out)
    return CONP_ERR_ARG;
  static_assert(sizeof(ZconvGroup) == 32 * sizeof(int), "ZconvGroup is handed out as 32 ints");
  std::vector<ZconvGroup> narrow;
  std::vector<int> wide, aout;
  ZconvPlan plan;
  plan_pppm_zconv(std::vector<int>(krad, krad + ncol), ncol, nz, nzi, zs_lo, nzl, zin_lo,
                  std::vector<int>(zout, zout + nzo), real_kernel != 0, narrow, wide, aout, plan);
  *ngroups_out = (int)narrow.size();
  *nwide_out = (int)wide.size();
  caps_out[0] = plan.rcap; caps_out[1] = plan.npcap;
  if (groups_out)
    std::memcpy(groups_out, narrow.data(), sizeof(ZconvGroup) * std::min<size_t>(narrow.size(), (size_t)std::max(max_groups, 0)));
  if (wide_out)
    std::memcpy(wide_out, wide.data(), sizeof(int) * std::min<size_t>(wide.size(), (size_t)std::max(max_wide, 0)));
  if (aout_out)
    for (int i = 0; i < nzo; ++i) aout_out[i] = aout[i];
  return CONP_OK;
}

int conp_plan_pair_runs(const double boxlo[3], const double prd[3], const int periodic[3], double rc, int n,
                        const double *xyz, int *nc_out, int *run_start_out, int *runs_out, int max_runs,
                        int *nruns_out) {
  if (!boxlo || !prd || !periodic || !nc_out || !nruns_out || n < 0 || (n > 0 && !xyz) || !(rc > 0.0))
    return CONP_ERR_ARG;
  const CellGrid g = make_cell_grid(boxlo, prd, periodic, rc);
  std::vector<int> rs;
  std::vector<PairRun> rr;
  build_pair_runs(g, 0, n, xyz, rs, rr);
  for (int a = 0; a < 3; ++a) nc_out[a] = g.nc[a];
  *nruns_out = (int)rr.size();
  if (run_start_out)
    for (int i = 0; i <= n; ++i) run_start_out[i] = rs[i];
  if (runs_out)
    for (int i = 0; i < (int)rr.size() && i < max_runs; ++i) {
      int *o = runs_out + 5 * (size_t)i;
      o[0] = rr[i].c0; o[1] = rr[i].c1; o[2] = rr[i].sx; o[3] = rr[i].sy; o[4] = rr[i].sz;
    }
  return CONP_OK;
}

int conp_bench_dgemm_tflops(conp_ctx *c, int n, double *tflops_out) {
  return guard(c, [&] {
    DevBuf<double> A, B, C;
    A.zero((size_t)n * n, c->stream); B.zero((size_t)n * n, c->stream); C.zero((size_t)n * n, c->stream);
    const double one = 1.0, zero = 0.0;
    CUBLAS_CHECK(cublasDgemm(c->blas, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &one, A.p, n, B.p, n, &zero, C.p, n));
    CUDA_CHECK(cudaEventRecord(c->ev[12], c->stream));
    const int reps = 3;
    for (int r = 0; r < reps; ++r)
      CUBLAS_CHECK(cublasDgemm(c->blas, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &one, A.p, n, B.p, n, &zero, C.p, n));
    CUDA_CHECK(cudaEventRecord(c->ev[13], c->stream));
    CUDA_CHECK(cudaEventSynchronize(c->ev[13]));
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, c->ev[12], c->ev[13]));
    if (tflops_out) *tflops_out = 2.0 * n * (double)n * n * reps / (ms * 1e-3) / 1e12;
  });
}

}  // extern "C"
