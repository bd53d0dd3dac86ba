// Internal declarations shared by the translation units of libconp_b200.so.
// Public boundary: include/conp_b200.h.
#pragma once

#include <cuda_runtime.h>
#include <cufft.h>
#include <cublas_v2.h>
#include <cusolverDn.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include <map>
#include <utility>

#include "../../include/conp_b200.h"
#include "peer.cuh"

namespace conp {

// ---------------------------------------------------------------------------
// error plumbing: nothing throws across the C ABI
// ---------------------------------------------------------------------------
struct Error {
  int code;
  std::string msg;
};

#define CONP_THROW(code_, ...)                                   \
  do {                                                           \
    char buf_[512];                                              \
    snprintf(buf_, sizeof(buf_), __VA_ARGS__);                   \
    throw ::conp::Error{(code_), std::string(buf_)};             \
  } while (0)

#define CUDA_CHECK(expr)                                                                     \
  do {                                                                                       \
    cudaError_t e_ = (expr);                                                                 \
    if (e_ != cudaSuccess)                                                                   \
      CONP_THROW(CONP_ERR_CUDA, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(e_),       \
                 __FILE__, __LINE__, #expr);                                                 \
  } while (0)

#define CUFFT_CHECK(expr)                                                                    \
  do {                                                                                       \
    cufftResult r_ = (expr);                                                                 \
    if (r_ != CUFFT_SUCCESS)                                                                 \
      CONP_THROW(CONP_ERR_CUDA, "cuFFT error %d at %s:%d (%s)", (int)r_, __FILE__, __LINE__, \
                 #expr);                                                                     \
  } while (0)

#define CUSOLVER_CHECK(expr)                                                                 \
  do {                                                                                       \
    cusolverStatus_t r_ = (expr);                                                            \
    if (r_ != CUSOLVER_STATUS_SUCCESS)                                                       \
      CONP_THROW(CONP_ERR_CUDA, "cuSOLVER error %d at %s:%d (%s)", (int)r_, __FILE__,        \
                 __LINE__, #expr);                                                           \
  } while (0)

#define CUBLAS_CHECK(expr)                                                                   \
  do {                                                                                       \
    cublasStatus_t r_ = (expr);                                                              \
    if (r_ != CUBLAS_STATUS_SUCCESS)                                                         \
      CONP_THROW(CONP_ERR_CUDA, "cuBLAS error %d at %s:%d (%s)", (int)r_, __FILE__,          \
                 __LINE__, #expr);                                                           \
  } while (0)

// ---------------------------------------------------------------------------
// device buffer with RAII; grows, never shrinks
// ---------------------------------------------------------------------------
template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t cap = 0;
  bool owned = true;
  ~DevBuf() { release(); }
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  void release() {
    if (p && owned) cudaFree(p);
    p = nullptr;
    cap = 0;
    owned = true;
  }
  // view into memory owned elsewhere (the peer-to-peer exchange arena)
  void attach(T *ptr, size_t n) {
    release();
    p = ptr;
    cap = n;
    owned = false;
  }
  // take over another buffer's allocation
  void adopt(DevBuf &o) {
    release();
    p = o.p; cap = o.cap; owned = o.owned;
    o.p = nullptr; o.cap = 0; o.owned = true;
  }
  void reserve(size_t n) {
    if (n <= cap) return;
    // a view into the peer-to-peer arena must never be silently replaced by a private allocation: the
    // peers keep writing into the arena.  The owner (ctx.cu: drop_p2p / ensure_p2p) re-sizes the arena.
    if (p && !owned)
      CONP_THROW(CONP_ERR_STATE, "internal: buffer attached to the exchange arena asked to grow (%zu > %zu)", n, cap);
    release();
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, n * sizeof(T));
    if (e != cudaSuccess)
      CONP_THROW(CONP_ERR_NOMEM, "cudaMalloc of %zu bytes failed: %s", n * sizeof(T), cudaGetErrorString(e));
    p = (T *)q;
    cap = n;
  }
  void upload(const T *h, size_t n, cudaStream_t s) {
    reserve(n);
    if (n) CUDA_CHECK(cudaMemcpyAsync(p, h, n * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void upload(const std::vector<T> &h, cudaStream_t s) { upload(h.data(), h.size(), s); }
  void zero(size_t n, cudaStream_t s) {
    reserve(n);
    if (n) CUDA_CHECK(cudaMemsetAsync(p, 0, n * sizeof(T), s));
  }
};

template <typename T>
struct PinnedBuf {
  T *p = nullptr;
  size_t cap = 0;
  ~PinnedBuf() {
    if (p) cudaFreeHost(p);
  }
  void reserve(size_t n) {
    if (n <= cap) return;
    if (p) cudaFreeHost(p);
    void *q = nullptr;
    cudaError_t e = cudaMallocHost(&q, n * sizeof(T));
    if (e != cudaSuccess) CONP_THROW(CONP_ERR_NOMEM, "cudaMallocHost of %zu bytes failed", n * sizeof(T));
    p = (T *)q;
    cap = n;
  }
};

// 32-byte packed point charge: position (wrapped into the box along periodic
// dimensions) and charge.  One 2x LDG.128 per atom in the pair/spread kernels.
struct __align__(32) PosQ {
  double x, y, z, q;
};

// electrode atom in the static cell-sorted list (wrapped position, eleall index, type)
struct __align__(32) EPos {
  double x, y, z;
  int idx, type;
};

// one x-contiguous run of cells [c0, c1) seen through the periodic image shift (sx, sy, sz)
struct PairRun {
  int c0, c1;
  short sx, sy, sz, pad;
};

// ---------------------------------------------------------------------------
// geometry of the uniform cell grid used to bin point charges
// ---------------------------------------------------------------------------
struct CellGrid {
  int nc[3];
  int periodic[3];
  int smax[3];  // periodic image shifts searched: -smax..smax
  double lo[3], prd[3], cinv[3];
  double rc;    // search radius the grid was sized for
  int ncells;
};

// per-type-pair tables for the real-space kernels (device pointers)
struct PairTables {
  int ntypes;
  int pairmode;
  double g_ewald, eta;
  const double *cuteff;  // (ntypes+1)^2: min(cutsq, cut_coulsq) if the pair is listed else 0
  const double *eta_ij;  // EHGO
  const double *fo_ij;   // EHGO
};

struct EwaldHost {
  int kxmax = 0, kymax = 0, kzmax = 0, kcount = 0, kcount_flat = 0, kcount_expand = 0;
  int dims[7] = {0, 0, 0, 0, 0, 0, 0};
  double unitk[3] = {0, 0, 0}, gsqmx = 0, volume = 0, ug_tot = 0;
  // device ordering: grouped by (kx, ky) with kz fastest
  std::vector<short> kx, ky, kz;
  std::vector<double> ug;
};

struct PPPMGeom {
  int nx, ny, nz, order, nlower;
  double boxlo[3], delinv[3], delvolinv, shift, shiftone;
  // plane pruning (pppm.cu): compact input planes zi <-> (zin_lo + zi) mod nz, zi < nzi;
  // compact output planes zmap[mz] (device array of nz ints, -1 if the plane is not needed)
  int nzi, zin_lo, nzo;
  const int *zmap;
  // multi-GPU: this rank owns the slab of input planes [zs_lo, zs_lo + zs_n) (all of them on one GPU)
  int zs_lo, zs_n;
};

// ---------------------------------------------------------------------------
// NCCL, loaded at run time (comm.cu)
// ---------------------------------------------------------------------------
struct Comm;
Comm *comm_create(int rank, int nranks, const void *unique_id);
void comm_destroy(Comm *);
int comm_get_unique_id(void *out);
void comm_allgather(Comm *, const void *send, void *recv, size_t bytes_per_rank, cudaStream_t);
void comm_allgatherv(Comm *, const void *send, void *recv, const size_t *bytes, const size_t *offsets,
                     int rank, int nranks, cudaStream_t);
void comm_allreduce_sum_f64(Comm *, double *buf, size_t n, cudaStream_t);

// ---- direct NVLink exchanges over CUDA-IPC mapped arenas (comm.cu) -------------------------
// Every rank allocates one arena of the same size; all ranks map each other's arena.  Buffers
// that are exchanged live at the same offset in every arena.  push = copy my block into every
// peer's arena and raise my flag there; wait = spin until every peer's flag for that channel has
// reached the current epoch.  Replaces the small latency-bound NCCL collectives of the step.
struct PeerArena;
PeerArena *p2p_create(Comm *, size_t bytes, cudaStream_t);          // collective; nullptr if IPC is unavailable
void p2p_destroy(PeerArena *);
char *p2p_local(PeerArena *);
// copy [off, off+bytes) of my arena to the same place in every peer's arena, then signal `chan`
int p2p_push(PeerArena *, size_t off, size_t bytes, int chan, cudaStream_t);
// all-gather with per-rank blocks at off + r*stride: pushes my block and waits for everyone's
int p2p_allgather(PeerArena *, size_t off, size_t stride_bytes, size_t block_bytes, int chan, cudaStream_t);
int p2p_wait(PeerArena *, int chan, cudaStream_t);
// deterministic sum over ranks of n doubles at `off` (result in place on every rank); `stage_off` is
// scratch of nranks * ceil(n/nranks) doubles; uses channels chan and chan+1
int p2p_allreduce_f64(PeerArena *, size_t off, size_t n, size_t stage_off, int chan, cudaStream_t);
// same sum for a vector whose producer kernel raised `chan_ready` itself (peer_block_signal): pull the
// partial slices over NVLink, store the total everywhere, wait on `chan_done`
// raise_ready: the producer only stored; the pull kernel raises this rank's `chan_ready` flags itself
int p2p_allreduce_pull_f64(PeerArena *, size_t off, size_t n, int chan_ready, int chan_done, cudaStream_t,
                           int raise_ready = 0);
size_t p2p_arena_offset(size_t payload_off);  // payload offset -> byte offset from the arena base
int p2p_error(PeerArena *);  // non-zero after a wait timed out
// argument for kernels that do their own exchange (peer.cuh); a null arena gives the no-op value
PeerSync p2p_sync(PeerArena *, int chan);
// flags of `chan` for a producer kernel that did not signal itself; value != nullptr: *value is first
// stored at payload offset value_off of every rank's arena
int p2p_signal(PeerArena *, int chan, cudaStream_t, size_t value_off = 0, const double *value = nullptr,
               size_t count_off = 0, const int *counts = nullptr /* counts[r] -> rank r, slot `rank` at count_off */);
int p2p_wait_sync(const PeerSync &, cudaStream_t);  // stand-alone wait on a PeerSync

// Opt a kernel in to `bytes` of dynamic shared memory.  The attribute is per device (one process may hold
// contexts on several GPUs), so the largest size requested so far is remembered per (kernel, device).
template <class Kernel>
inline void ensure_dynamic_smem(Kernel kernel, size_t bytes) {
  static std::map<std::pair<const void *, int>, size_t> granted;
  int dev = 0;
  CUDA_CHECK(cudaGetDevice(&dev));
  size_t &have = granted[{reinterpret_cast<const void *>(kernel), dev}];
  if (bytes > have) {
    CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    have = bytes;
  }
}

// ---------------------------------------------------------------------------
// kernel launchers (one .cu per group); all take the context's stream and
// return the number of kernels they launched
// ---------------------------------------------------------------------------

// gemv.cu ------------------------------------------------------------------
// update_charge epilogue parameters (device pointers); see charge_epilogue() in gemv.cu
struct ChargeEpilogue {
  int enabled, variant, n, row_offset, one_electrode;
  int neutral, n_left;   // neutral: remove the rounding residual sum(S.b)/n from every charge (projection on)
  double sum_setz;
  double totsetq, lz, vmult;
  const double *value;   // dV | QR | D in device memory
  const double *dipole;  // sum q z of the non-electrode atoms
  const int *side;
  const double *setz, *setq, *qinit, *sb;
  double *q_out, *scalar_out /* [0]=scalar, [1]=potdiff, [13]=mean residual */, *partials /* 3 per block */;
  unsigned int *counter;
};
// out[r] = sum_c S[r*pitch + c] * b[c], r < nrows, c < ncols_pad (pad columns of S and b are zero).
// With ep != nullptr the charge epilogue runs in the tail of the same kernel (single-GPU path).
int launch_gemv(cudaStream_t s, const double *S, size_t pitch, int nrows, int ncols_pad, const double *b,
                double *out, int num_sms, const ChargeEpilogue *ep);
// Symmetric S: out[c] (c < out_len, zero beyond N) = this GPU's share of S.b from the half band of its
// rows [row0, row0+nrows) -- the complete product on one GPU, a partial sum to be all-reduced on
// several.  rowpart: >= N doubles, colpart: nstrips*L doubles of scratch.  See symv_tma_kernel.
struct SymvPlan {
  bool usable = false;
  int grid = 0, nstrips = 0, L = 0;
  const int2 *strips = nullptr;  // device: [a, bnd) global row range of every strip
};
SymvPlan plan_symv(int N, int row0, int nrows, int num_sms, std::vector<int2> &strips);
// wait_b: poll the b exchange in the kernel's prologue; push_parts: store the result into slot `rank` of
// every rank's staging area at off_parts (out_len doubles per slot) instead of `out`, and signal.
int launch_symv(cudaStream_t s, const double *S, size_t pitch, int N, int row0, int nrows, const double *b,
                const SymvPlan &plan, double *rowpart, double *colpart, double *out, int out_len,
                const ChargeEpilogue *ep, const PeerSync &wait_b, const PeerSync &push_parts, size_t off_parts);
// sum of the nranks staged partial vectors (after their flags are up) -> sb_out, then the epilogue
int launch_update_charge_sum(cudaStream_t s, const ChargeEpilogue &ep, const PeerSync &ps, const double *parts,
                             int len, double *sb_out);
int launch_update_charge(cudaStream_t s, const ChargeEpilogue &ep);
int launch_finalize_q(cudaStream_t s, int n, const double *sb, const double *setq, const double *qinit,
                      const double *scal /* [1] = potdiff, [13] = mean residual of S.b */, double *q_out);

// pair.cu ------------------------------------------------------------------
CellGrid make_cell_grid(const double lo[3], const double prd[3], const int periodic[3], double rc);
// host: static electrode structures (electrodes never move)
void build_electrode_cells(const CellGrid &g, int begin, int end, const double *xyz, const int *type,
                           std::vector<EPos> &sorted, std::vector<int> &cell_start);
void build_near_mask(const CellGrid &g, int begin, int end, const double *xyz, std::vector<unsigned char> &mask);
void build_pair_runs(const CellGrid &g, int begin, int end, const double *xyz, std::vector<int> &run_start,
                     std::vector<PairRun> &runs);
// per-step counting sort of the point charges: pack (+histogram, + sum q z), scan, scatter
int launch_pack_count(cudaStream_t s, const CellGrid &g, int m, const double *x_raw, const int *idx,
                      const double *q, const int *type, PosQ *packed, int *packed_type, int *cell_of, int *slot,
                      int *cell_count, double *qz_sum, const PeerSync &ps /* fused all-gather of the block */,
                      size_t off_block /* arena offset of `packed` */, int mpad);
// multi-GPU layout: nranks blocks of mpad slots, counts[r] valid charges each; m = nranks*mpad
int launch_bin_positions(cudaStream_t s, const CellGrid &g, int m, int mpad, const int *counts,
                         const PosQ *packed, int *cell_of, int *slot, int *cell_count,
                         const unsigned char *relevant /* per-cell filter or nullptr */,
                         const PeerSync &ps /* wait for the fused all-gather */);
// qz_sum != nullptr: also sum the per-rank sum(q z) partials stored in the last slot of every block
int launch_cell_scan(cudaStream_t s, int ncells, const int *cell_count, int *cell_start, const PosQ *packed,
                     int mpad, int nranks, double *qz_sum);
int launch_cell_scatter(cudaStream_t s, const CellGrid &g, int m, const PosQ *packed, const int *type,
                        const int *cell_of, const int *slot, const int *cell_start, PosQ *sorted, int *sorted_type,
                        int *sorted_src, float4 *sorted_f);
// charges sitting in a cell within reach of an electrode atom (near_count must be zeroed); post_force only
int launch_near_list(cudaStream_t s, const CellGrid &g, int m, const PosQ *packed, const unsigned char *near_mask,
                     int *near_list, int *near_count, const int *counts = nullptr, int mpad = 1);
// several GPUs, routed exchange: own[j] = wrapped charge j of this rank; a copy goes into the inbox (arena
// regions off_packed / off_ptype / off_psrc, sender block `rank`) of every rank whose rel_all row marks the
// charge's cell; send_count[r] (zeroed by the caller) ends up as the number sent to rank r
int launch_pack_route(cudaStream_t s, const CellGrid &g, int m, const double *x_raw, const int *idx, const double *q,
                      const int *type, PosQ *own, int rank, int nranks, int mpad, const unsigned char *rel_all,
                      const PeerSync &ps, size_t off_packed, size_t off_ptype, size_t off_psrc, int *send_count,
                      double *qz_sum);
// b_real[i] = -sum_j q_j dudq(r_ij), rows [row_begin,row_end), against the cell-sorted point charges
int launch_pair_b(cudaStream_t s, const CellGrid &g, const PairTables &pt, int row_begin, int row_end,
                  const double *ex, const double *ey, const double *ez, const int *etype, const int *run_start,
                  const PairRun *runs, const PosQ *sorted, const int *sorted_type, const float4 *sorted_f,
                  const int *cell_start, double *b_real);
int launch_pair_A(cudaStream_t s, const CellGrid &g, const PairTables &pt, const EPos *esorted,
                  const int *cell_start, int row_begin, int row_end, const double *ex, const double *ey,
                  const double *ez, const int *etype, const int *run_start, const PairRun *runs, double *A_rows,
                  size_t pitch);
// phi[i] = sum_j q_ele[j] dudq_A(r_ij) over the electrode atoms j != i (and the periodic images of i), rows
// [row_begin, row_end): the electrode-electrode pair part of compute potential/atom
int launch_pair_P(cudaStream_t s, const CellGrid &g, const PairTables &pt, const EPos *esorted,
                  const int *cell_start, int row_begin, int row_end, const double *ex, const double *ey,
                  const double *ez, const int *etype, const int *run_start, const PairRun *runs,
                  const double *q_ele, double *phi);
int launch_pair_postforce(cudaStream_t s, const CellGrid &g, const PairTables &pt, double qqrd2e,
                          const EPos *esorted, const int *cell_start, const double *q_ele, const PosQ *packed,
                          const int *packed_type, const int *near_list, const int *near_count, int max_near,
                          const double *cutsq_listed, double *f_packed /* m x 3 */, double *energies /* 8 */,
                          int num_sms, const int *psrc = nullptr /* routed exchange: sender-local index */,
                          int mpad = 1);

// pppm.cu ------------------------------------------------------------------
// spreads the sorted charges of cells [cell_lo, cell_hi) (all of them if cell_start == nullptr) onto the
// rank's slab of input planes; m_bound = upper bound of the number of charges in that range
// inbox_counts != nullptr: the charges are read unsorted, as they arrived -- sender r's block of the inbox starts at
// slot r * mpad and holds inbox_counts[r] charges; `valid` (optional) < 0 marks slots this rank does not read
int launch_pppm_spread(cudaStream_t s, const PPPMGeom &g, const double *rho_coeff, int m_bound, const PosQ *atoms,
                       double *brick, int *range_flag, const int *inbox_counts = nullptr, int nsenders = 0,
                       int mpad = 0, const int *valid = nullptr);
int launch_fill_zero(cudaStream_t s, double *p, size_t n);
// Owner-computes form of the same spread (no atomics): the rank's slab of input planes is cut into tiles of
// tz x ty x tx mesh points, one warp owns one tile at a time in its private shared memory and stores every
// mesh point exactly once (also the zeros: no memset of the brick).  plan_pppm_spread_tiles works out, per
// tile, which x-contiguous runs of cells can hold charges whose stencil reaches the tile.
struct SpreadPlan {
  int ntiles = 0, tz = 0, ty = 0, tx = 0, ntz = 0, nty = 0, ntx = 0;
  int rs = 0, ps = 0;                 // shared-memory row / plane stride in doubles (bank-conflict padding)
  int halo_z = 0, halo_y = 0, halo_x = 0;  // a tile that spans its whole axis keeps order-1 halo entries (folded on store)
  const int *run_start = nullptr;     // device [ntiles + 1]
  const int2 *runs = nullptr;         // device: [c0, c1) ranges of cell indices
  int *counter = nullptr;             // device: tile scheduler, zeroed before every launch
  size_t smem = 0;
  int grid = 0, grid_mma = 0;
  // tensor-core kernel (spread_mma_kernel): in = allowed, out = chosen; per-charge stencil origins / weights
  int use_mma = 0;
  int4 *origin = nullptr;
  double *weights = nullptr;   // [3 order][wstride]
  size_t wstride = 0;
};
void plan_pppm_spread_tiles(const PPPMGeom &g, const CellGrid &cells, int num_sms, std::vector<int> &run_start,
                            std::vector<int2> &runs, SpreadPlan &plan);
int launch_pppm_spread_tiles(cudaStream_t s, const PPPMGeom &g, const SpreadPlan &plan,
                             const double *rho_coeff_host /* order x order, host memory */, const PosQ *atoms,
                             const int *cell_start, int m_bound /* upper bound of the sorted charges */,
                             const int *count_ptr /* device: their actual number, or nullptr */, double *brick,
                             int *range_flag);
// z-sweep spread (pppm.cu): second, mesh-aligned sort of the charges + one warp per (column, z-segment)
struct SweepPlan {
  bool usable = false;
  int ncolx = 0, ncoly = 0, pz_lo = 0, npz = 0, wrap_z = 0, nbins = 0, nitems = 0, grid = 0;
  // device buffers (owned by the context)
  int *bin_of = nullptr, *slot = nullptr, *bin_count = nullptr, *bin_start = nullptr, *counter = nullptr;
  double *records = nullptr;  // [m][(3 order + 2) & ~1]: origin + weights of every charge, in bin order
  const int4 *items = nullptr;
};
void plan_pppm_sweep(const PPPMGeom &g, int num_sms, std::vector<int4> &items, SweepPlan &plan);
// atoms: the packed (unsorted) charges; valid: nullptr, or per slot a value < 0 for slots to skip
int launch_pppm_spread_sweep(cudaStream_t s, const PPPMGeom &g, const SweepPlan &plan, const double *rho_coeff_host,
                             const PosQ *atoms, int m_bound, const int *valid, double *brick, int *range_flag);
int launch_pppm_green_mul(cudaStream_t s, size_t n, cufftDoubleComplex *work, const double *ghalf);
// rhat: spectra of the rank's nzl input planes (compact planes zs_lo..); uhat: (partial) output-plane spectra
// Launch plan of the z-convolution (built once per rank by plan_pppm_zconv): narrow column groups stage
// only the input planes near an output plane; the few small-|k_xy| columns whose kernel spans the mesh
// take the wide path.  See zconv_kernel.
struct ZconvGroup {  // 32 ints
  static constexpr int MAXI = 8;
  int c0, rblock, nint, np;          // first column, window radius, intervals, staged planes
  int lo[MAXI], hi[MAXI], base[MAXI];  // slab-local plane interval [lo, hi) -> compact rows base..
  int pad[4];
};
struct ZconvPlan {
  const ZconvGroup *narrow = nullptr;  // device
  const int *wide = nullptr;           // device list of wide columns
  const int *aout = nullptr;           // device: ring position of every output plane (compact coordinates)
  int n_narrow = 0, n_wide = 0;
  int rcap = 0, npcap = 1;             // narrow path: largest window radius / staged planes it is sized for
};
void plan_pppm_zconv(const std::vector<int> &krad, int ncol, int nz, int nzi, int zs_lo, int nzl, int zin_lo,
                     const std::vector<int> &zout, bool real_k, std::vector<ZconvGroup> &narrow,
                     std::vector<int> &wide, std::vector<int> &aout, ZconvPlan &plan);
int launch_pppm_zconv(cudaStream_t s, int ncol, int nz, int nzl, int zs_lo, int nzo, const int *krad,
                      const ZconvPlan &plan, const cufftDoubleComplex *rhat, const double *Kr,
                      const cufftDoubleComplex *Kc, cufftDoubleComplex *uhat,
                      const PeerSync &ps /* several GPUs: announce the partial spectra */);
int launch_expand_planes(cudaStream_t s, size_t plane, int nplanes, int nz, int lo, const int *list,
                         const double *compact, double *full);
int launch_pppm_ele_stencil(cudaStream_t s, const PPPMGeom &g, const double *rho_coeff, int n, const double *ex,
                            const double *ey, const double *ez, int *part2grid, double *weights);
// wrapped stencil indices [n][3][order] (x, y, compact z plane) of the static electrode atoms; needs g.zmap
int launch_pppm_ele_index(cudaStream_t s, const PPPMGeom &g, int n, const int *part2grid, int *widx);
// flat stencil table [n][order^3]: offset into the compact potential brick, weight product
int launch_pppm_point_table(cudaStream_t s, const PPPMGeom &g, int n, const int *widx, const double *weights,
                            int *poff, double *pw);
int launch_pppm_gather_b(cudaStream_t s, const PPPMGeom &g, int row_begin, int row_end, const int *poff,
                         const double *pw, const double *u_brick, const double *ez, const double *qz_sum,
                         double slab_pref, const double *b_real, double *b_kspace, double *b,
                         const PeerSync &ps /* fused b exchange; null arena: none */, size_t off_b);
// electrode re-spread; forms q_i = sb_i + potdiff*setq_i (+qinit_i) on the fly and stores it to q_out
// (spreads rows [row_begin,row_end) only; all n charges are written to q_out)
int launch_pppm_ele_spread(cudaStream_t s, const PPPMGeom &g, int n, int row_begin, int row_end, const int *widx,
                           const double *weights, const double *sb, const double *setq, const double *qinit,
                           const double *scal, double *q_out, double *brick);
int launch_add_bricks(cudaStream_t s, size_t n, const double *a, const double *b, double *out);
// out[i] = sum over the order^3 stencil of point i (weights from its position, pppm_conp.cpp:452-484) of the
// full-mesh potential u[nz][ny][nx]
int launch_mesh_potential(cudaStream_t s, const PPPMGeom &g, const double *rho_coeff, int n, const double *xyz,
                          const double *u_full, double *out);
// out[z][y][x] over the sub-brick lo..hi (inclusive) of the periodic mesh: which = 0 electrolyte (compact
// input planes, all nzi of them), 1 electrode (compact output planes), 2 sum
int launch_region_gather(cudaStream_t s, const PPPMGeom &g, int which, const int lo[3], const int hi[3],
                         const double *elyte, const double *ele, double *out);

// ewald.cu -----------------------------------------------------------------
void ewald_setup_host(EwaldHost &e, double g_ewald, double accuracy, double q2, long long natoms,
                      const double prd[3], double slab_volfactor);
// axis tables E[a][i][m] = exp(i m unitk_a r_a), m = 0..kmax_a; layout per atom: (kxmax+1)+(kymax+1)+(kzmax+1) double2
int launch_axis_tables(cudaStream_t s, int n, const double *x, const double *y, const double *z,
                       const PosQ *packed /* or null */, const double unitk[3], int kxmax, int kymax, int kzmax,
                       double2 *tab);
int launch_ewald_sfac(cudaStream_t s, int m, const PosQ *atoms, const double2 *tab, int kxmax, int kymax,
                      int kzmax, int kcount, const short *kx, const short *ky, const short *kz,
                      double *sfac /* 2*kcount, zeroed by the launcher */);
int launch_ewald_bextract(cudaStream_t s, int row_begin, int row_end, const double2 *etab, int kxmax, int kymax,
                          int kzmax, int kcount, const short *kx, const short *ky, const short *kz,
                          const double *ug, const double *sfac, const double *ez, const double *qz_sum,
                          double slab_pref, const double *b_real, double *b_kspace, double *b);
// k-major panel Pt[2*kc][ld]: rows kk / kc+kk = sqrt(2 u_k) {cos, sin}(k.r_i), k = k0+kk; segs = runs of
// consecutive k with equal (kx,ky) inside the chunk
// Tensor-core (FP64 DMMA) form of the structure-factor sum and of the b extraction (large systems), see
// ewald.cu: both are one product of k-major operands through launch_tn_gemm (gram.cu)
struct EwaldGemm {
  int nkxy = 0, nkz1 = 0, chunk = 0, ksplit = 1, kz16 = 0;
  size_t np = 0;                           // nkxy * nkz1
  size_t wa = 0, wb = 0, wr = 0;           // padded widths: [Fr | Fi], [Cz | Sz], electrode rows
  size_t pslice = 0;                       // doubles per partial product (2 nkxy x wb)
  DevBuf<short> d_xk, d_yk;                // (kx, ky) of every distinct pair
  DevBuf<int> d_kxyof;                     // pair index of every listed k-vector
  DevBuf<double> d_fa, d_zb;               // operands of one chunk of point charges (one k-row per charge)
  DevBuf<double> d_p, d_a, d_t;            // partial products [ksplit], W combinations (k-major), [Tr | Ti]
  DevBuf<double> d_fx, d_ze;               // static electrode-side operands (own rows): e_xy, E_z k-major
};
void ewald_gemm_plan(EwaldGemm &g, const EwaldHost &e, int m_total, int nrows, int num_sms, cudaStream_t s);
int ewald_gemm_electrodes(cudaStream_t s, EwaldGemm &g, const EwaldHost &e, int row_begin, int row_end,
                          const double2 *etab);
// partial S(k) over the sorted charges [j_begin, j_end)
int ewald_gemm_sfac(cudaStream_t s, EwaldGemm &g, const EwaldHost &e, int j_begin, int j_end, const PosQ *atoms,
                    const double2 *tab, const short *kz, double *sfac);
int ewald_gemm_bextract(cudaStream_t s, EwaldGemm &g, const EwaldHost &e, int row_begin, int row_end,
                        const short *kz, const double *ug, const double *sfac, const double *ez,
                        const double *qz_sum, double slab_pref, const double *b_real, double *b_kspace, double *b);
int launch_ewald_panel(cudaStream_t s, int n, const double2 *etab, int kxmax, int kymax, int kzmax, int k0,
                       int kc, int nseg, const int2 *segs, const short *kx, const short *ky, const short *kz,
                       const double *ug, double *panel, size_t ld);

// gram.cu ------------------------------------------------------------------
// C[i][j] += sum_k Pt[k][row_begin+i] * Pt[k][j] (k-major panel, ld doubles per k-row); FP64 tensor cores
// Only the tiles touching the cyclic half band (j - i) mod n in [0, n/2] of the rows [row0, row0 + nrows) are
// computed (same work for every row block); launch_gram_mirror completes the assembled n x n matrix
int launch_gram_accumulate(cudaStream_t s, int nrows, int n, int kdim, const double *panel_rows,
                           const double *panel_all, size_t ld, double *C, size_t pitch, int row0);
int launch_gram_mirror(cudaStream_t s, int n, double *C, size_t pitch);
// Same DMMA tile kernel as a general product of k-major operands: C[z][i][j] (+)= sum_{k in slice z} A[k][i] B[k][j],
// i < m, j < n.  A (lda) and B (ldb) are zero-padded to multiples of 128 columns and kdim to a multiple of 16;
// ksplit slices of k write separate output matrices slice_stride doubles apart.
int launch_tn_gemm(cudaStream_t s, int m, int n, int kdim, const double *A, size_t lda, const double *B, size_t ldb,
                   double *C, size_t pitch, int ksplit, size_t slice_stride, int accumulate);

// linalg.cu ----------------------------------------------------------------
int launch_a_finish(cudaStream_t s, int row_begin, int row_end, int n, double *A_rows, size_t pitch,
                    double diag_kspace, int pairmode, double self_eta, const double *u0_i, const int *etype,
                    double slab_pref, const double *ez);
int launch_project(cudaStream_t s, int n, double *S, size_t pitch, const int *subset /* or null */,
                   double *rowsum_tmp, double *tot_out /* device scalar */, int apply);
int launch_d_vector(cudaStream_t s, int n, const double *ez, const int *side, int ff_flag, double evscale,
                    double zlo, double zprd, double *d, double *setz);
int launch_pad_identity(cudaStream_t s, int n, double *B, size_t ld);
int launch_asymmetry(cudaStream_t s, int n, const double *S, size_t ld, double *out2);
int launch_symmetrise(cudaStream_t s, int n, double *S, size_t ld);

}  // namespace conp
