// Real-space electrode<->point-charge kernels.
//
//   pair_b         FixConp::blist_coul_cal        fix_conp.cpp:1281-1365
//   pair_A         FixConp::alist_coul_cal        fix_conp.cpp:1209-1279
//   pair_postforce FixConp::blist_coul_cal_post_force fix_conp.cpp:1368-1444
//   erfcr_sqrt / ferfcr_sqrt / eta_* / ehgo_*     fix_conp.cpp:1446-1480, 1561-1573
//
// The pair set is geometric (rsq < cutsq[it][jt] and rsq < cut_coulsq,
// fix_conp.cpp:1333-1334) over *all* periodic images, which is what LAMMPS'
// ghost atoms provide.  Point charges are counting-sorted into a uniform cell
// grid every step (this also gives the PPPM spread its memory locality);
// electrode atoms never move and are sorted once at setup on the host.  One
// warp owns one electrode row, lanes stride the x-contiguous cell runs
// (coalesced 32-byte loads), rows of cells are culled against the cut-off
// sphere, and the candidates that pass the distance test are compacted
// through a per-warp shared-memory queue so the FP64 erfc evaluation always
// runs on full warps.  The row sum stays in registers: no atomics for b.
#include "common.cuh"

#include <algorithm>
#include <cmath>

namespace conp {

namespace {

constexpr double EWALD_F = 1.12837917;
constexpr double EWALD_P = 0.3275911;
constexpr double A1 = 0.254829592;
constexpr double A2 = -0.284496736;
constexpr double A3 = 1.421413741;
constexpr double A4 = -1.453152027;
constexpr double A5 = 1.061405429;
constexpr double ERFC_MAX = 5.8;

// fix_conp.cpp:1446-1454
__device__ __forceinline__ double erfcr_sqrt(double a2_r2) {
  if (a2_r2 < ERFC_MAX * ERFC_MAX) {
    const double a_r = sqrt(a2_r2);
    const double expm2 = exp(-a2_r2);
    const double t = 1.0 / (1.0 + EWALD_P * a_r);
    return t * (A1 + t * (A2 + t * (A3 + t * (A4 + t * A5)))) * expm2 / a_r;
  }
  return 0.0;
}
// fix_conp.cpp:1456-1465
__device__ __forceinline__ double ferfcr_sqrt(double a2_r2) {
  if (a2_r2 < ERFC_MAX * ERFC_MAX) {
    const double a_r = sqrt(a2_r2);
    const double expm2 = exp(-a2_r2);
    const double t = 1.0 / (1.0 + EWALD_P * a_r);
    const double erfcr = t * (A1 + t * (A2 + t * (A3 + t * (A4 + t * A5)))) * expm2 / a_r;
    return erfcr + EWALD_F * expm2;
  }
  return 0.0;
}

enum { MODE_B = 0, MODE_A = 1, MODE_P = 2 };  // MODE_P: MODE_A's pairs, summed with the partners' charges

// erfcr_sqrt(g^2 r^2) g - erfcr_sqrt(eta^2 r^2) eta (the two lines above, Gaussian electrode charges) with the
// shared work done once: one reciprocal square root gives r and 1/r, one reciprocal gives both rational
// arguments t = 1/(1 + p a r).  Same formula, a few ulp apart from evaluating the two terms separately; this is
// the per-step pair kernel's inner loop (~7e6 pairs per update at cfg5).
__device__ __forceinline__ double dudq_two_gauss(double g_ewald, double eta, double rsq) {
  const double rinv = rsqrt(rsq), r = rsq * rinv;
  const double ag = g_ewald * r, ae = eta * r;
  const double u = fma(EWALD_P, ag, 1.0), v = fma(EWALD_P, ae, 1.0);
  const double w = 1.0 / (u * v);
  const double tg = v * w, te = u * w;
  const double ag2 = ag * ag, ae2 = ae * ae;
  const double eg = ag2 < ERFC_MAX * ERFC_MAX ? exp(-ag2) : 0.0;
  const double ee = ae2 < ERFC_MAX * ERFC_MAX ? exp(-ae2) : 0.0;
  const double pg = tg * (A1 + tg * (A2 + tg * (A3 + tg * (A4 + tg * A5))));
  const double pe = te * (A1 + te * (A2 + te * (A3 + te * (A4 + te * A5))));
  return (pg * eg - pe * ee) * rinv;
}

// dudq of fix_conp.cpp:1263-1264 / 1335-1336; (it, jt) = (electrode type, partner type)
template <int MODE>
__device__ __forceinline__ double dudq_pair(const PairTables &pt, double rsq, int it, int jt) {
  if (MODE == MODE_B && pt.pairmode != CONP_PAIR_EHGO) return dudq_two_gauss(pt.g_ewald, pt.eta, rsq);
  double v = erfcr_sqrt(pt.g_ewald * pt.g_ewald * rsq) * pt.g_ewald;
  if (pt.pairmode == CONP_PAIR_EHGO) {  // fix_conp.cpp:1561-1566
    const int ij = it * (pt.ntypes + 1) + jt;
    const double etaij = __ldg(pt.eta_ij + ij);
    const double foij = __ldg(pt.fo_ij + ij);
    const double etarij2 = etaij * etaij * rsq;
    v += foij * exp(-0.5 * etarij2) - erfcr_sqrt(etarij2) * etaij;
  } else if (MODE == MODE_A) {  // eta_potential_A fix_conp.cpp:1467-1470
    const double etarij2 = pt.eta * pt.eta * rsq / 2;
    v += -erfcr_sqrt(etarij2) * pt.eta / sqrt(2.0);
  } else {  // eta_potential fix_conp.cpp:1472-1475
    const double etarij2 = pt.eta * pt.eta * rsq;
    v += -erfcr_sqrt(etarij2) * pt.eta;
  }
  return v;
}

__device__ __forceinline__ int cell_coord(const CellGrid &g, int a, double x) {
  int k = (int)floor((x - g.lo[a]) * g.cinv[a]);
  k = k < 0 ? 0 : k;
  k = k > g.nc[a] - 1 ? g.nc[a] - 1 : k;
  return k;
}

__device__ __forceinline__ double wrap_coord(const CellGrid &g, int a, double x) {
  if (g.periodic[a]) x -= floor((x - g.lo[a]) / g.prd[a]) * g.prd[a];
  return x;
}

// ---------------------------------------------------------------------------
// per-step counting sort of the point charges
// ---------------------------------------------------------------------------
// pack: raw LAMMPS positions -> wrapped PosQ + type, sum(q z), cell histogram.
// Several GPUs, peer-to-peer path (ps.arena != nullptr): `packed` is this rank's block of the gathered
// array; every charge is also stored into the same slot of every peer's array, the last block adds this
// rank's sum(q z) in the block's last (padding) slot and raises the flags -- the position all-gather
// happens inside the kernel that produces the positions.
__global__ void __launch_bounds__(256)
pack_count_kernel(CellGrid g, int m, const double *__restrict__ x_raw, const int *__restrict__ idx,
                  const double *__restrict__ q, const int *__restrict__ type, PosQ *__restrict__ packed,
                  int *__restrict__ packed_type, int *__restrict__ cell_of, int *__restrict__ slot,
                  int *__restrict__ cell_count, double *__restrict__ qz_sum, PeerSync ps, size_t off_block,
                  int mpad) {
  __shared__ __align__(16) PosQ tile[256];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  double qz = 0.0;
  if (j < m) {
    const int src = idx ? idx[j] : j;
    const double xr = x_raw[3 * (size_t)src], yr = x_raw[3 * (size_t)src + 1], zr = x_raw[3 * (size_t)src + 2];
    const double qq = q[src];
    PosQ p;
    p.x = wrap_coord(g, 0, xr);
    p.y = wrap_coord(g, 1, yr);
    p.z = wrap_coord(g, 2, zr);
    p.q = qq;
    packed[j] = p;
    packed_type[j] = type[src];
    qz = qq * zr;  // raw z: km_ewald.cpp:839, fix_cond.cpp:103
    if (cell_of) {
      const int cell = (cell_coord(g, 2, p.z) * g.nc[1] + cell_coord(g, 1, p.y)) * g.nc[0] + cell_coord(g, 0, p.x);
      cell_of[j] = cell;
      slot[j] = atomicAdd(&cell_count[cell], 1);
    }
    if (ps.arena) tile[threadIdx.x] = p;
  }
  __shared__ double sh[8];
  if (ps.arena) {
    // the block's 256 records leave as whole 512-byte warp stores (16 B per lane), one pass per peer:
    // NVLink moves full lines instead of the 8-byte fragments a per-thread struct store would produce
    __syncthreads();
    const int base = blockIdx.x * blockDim.x;
    const int nrec = min((int)blockDim.x, m - base);
    const uint4 *src = reinterpret_cast<const uint4 *>(tile);
    for (int r = 0; r < ps.nranks; ++r) {
      if (r == ps.rank) continue;
      uint4 *dst = reinterpret_cast<uint4 *>(peer_ptr<PosQ>(ps, r, off_block) + base);
      for (int t = threadIdx.x; t < 2 * nrec; t += blockDim.x) dst[t] = src[t];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) qz += __shfl_xor_sync(0xffffffffu, qz, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = qz;
  __syncthreads();
  if (threadIdx.x < 8) {
    double v = sh[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffu, v, o);
    if (threadIdx.x == 0 && v != 0.0) atomicAdd(qz_sum, v);
  }
  peer_block_signal(ps, [&] {
    if ((int)threadIdx.x < ps.nranks) {
      const double tot = __ldcg(qz_sum);  // complete: every block added its share before taking its ticket
      peer_ptr<PosQ>(ps, threadIdx.x, off_block)[mpad - 1].x = tot;
    }
  });
}

// Several GPUs, routed exchange: a charge only travels to the ranks that can use it.  rel_all[r][cell] says
// whether rank r reads charges of that cell (cells within reach of its electrode rows, or feeding its slab
// of PPPM planes -- static, exchanged once).  Every rank r has, in its arena, one inbox of mpad slots per
// sender: the sender appends the charges r needs (position + charge, type, index in the sender's own list)
// with warp-aggregated slot counters, and the p2p_signal kernel behind this one delivers the per-receiver
// counts, the rank's sum(q z) and the flags.  `own` keeps all of this rank's wrapped charges (Ewald
// structure factors are summed by the owner, km_ewald.cpp:668-786).
__global__ void __launch_bounds__(256)
pack_route_kernel(CellGrid g, int m, const double *__restrict__ x_raw, const int *__restrict__ idx,
                  const double *__restrict__ q, const int *__restrict__ type, PosQ *__restrict__ own, int rank,
                  int nranks, int mpad, const unsigned char *__restrict__ rel_all, char *const *__restrict__ arena,
                  size_t off_packed, size_t off_ptype, size_t off_psrc, int *__restrict__ send_count,
                  double *__restrict__ qz_sum) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  double qz = 0.0;
  PosQ p = {0.0, 0.0, 0.0, 0.0};
  int cell = 0, ty = 0;
  if (j < m) {
    const int src = idx ? idx[j] : j;
    const double xr = x_raw[3 * (size_t)src], yr = x_raw[3 * (size_t)src + 1], zr = x_raw[3 * (size_t)src + 2];
    p.x = wrap_coord(g, 0, xr);
    p.y = wrap_coord(g, 1, yr);
    p.z = wrap_coord(g, 2, zr);
    p.q = q[src];
    ty = type[src];
    own[j] = p;
    qz = p.q * zr;  // raw z: km_ewald.cpp:839, fix_cond.cpp:103
    cell = (cell_coord(g, 2, p.z) * g.nc[1] + cell_coord(g, 1, p.y)) * g.nc[0] + cell_coord(g, 0, p.x);
  }
  for (int r = 0; r < nranks; ++r) {
    const bool need = j < m && rel_all[(size_t)r * g.ncells + cell] != 0;
    const unsigned mk = __ballot_sync(0xffffffffu, need);
    if (mk == 0u) continue;
    int base = 0;
    if (lane == __ffs(mk) - 1) base = atomicAdd(send_count + r, __popc(mk));
    base = __shfl_sync(0xffffffffu, base, __ffs(mk) - 1);
    if (need) {
      const size_t slot = (size_t)rank * mpad + base + __popc(mk & lt);
      char *a = arena[r] + P2P_CTRL_BYTES;
      reinterpret_cast<PosQ *>(a + off_packed)[slot] = p;
      reinterpret_cast<int *>(a + off_ptype)[slot] = ty;
      reinterpret_cast<int *>(a + off_psrc)[slot] = j;
    }
  }
  __shared__ double sh[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) qz += __shfl_xor_sync(0xffffffffu, qz, o);
  if (lane == 0) sh[threadIdx.x >> 5] = qz;
  __syncthreads();
  if (threadIdx.x < 8) {
    double v = sh[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffu, v, o);
    if (threadIdx.x == 0 && v != 0.0) atomicAdd(qz_sum, v);
  }
}

// multi-GPU: the gathered charges sit in `nranks` blocks of `mpad` slots; block r holds
// counts[r] charges, the rest is padding (the last slot carries that rank's sum(q z)).  With the fused
// exchange (ps.arena != nullptr) the kernel first waits for every rank's block to have landed.
__global__ void __launch_bounds__(256)
bin_positions_kernel(CellGrid g, int m, int mpad, const int *__restrict__ counts,
                     const PosQ *__restrict__ packed, int *__restrict__ cell_of, int *__restrict__ slot,
                     int *__restrict__ cell_count, const unsigned char *__restrict__ relevant, PeerSync ps) {
  if (ps.arena) {
    peer_block_wait(ps);
    __syncthreads();
  }
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  // (raise_first with counts: this rank's own entry of the delivered counts is written by block 0 of this very
  // kernel -- read it from the source instead)
  const int sender = j / mpad;
  const int cnt = (ps.arena && ps.raise_first && ps.counts && sender == ps.rank) ? __ldcg(ps.counts + sender)
                                                                                   : counts[sender];
  if (j % mpad >= cnt) {
    cell_of[j] = -1;
    return;
  }
  const PosQ p = packed[j];
  const int cell = (cell_coord(g, 2, p.z) * g.nc[1] + cell_coord(g, 1, p.y)) * g.nc[0] + cell_coord(g, 0, p.x);
  if (relevant && !relevant[cell]) {  // nothing on this rank reads charges of that cell
    cell_of[j] = -1;
    return;
  }
  cell_of[j] = cell;
  slot[j] = atomicAdd(&cell_count[cell], 1);
}

// exclusive scan of cell_count -> cell_start[ncells+1] by one block (cell_count is padded with zeros
// to a multiple of 4 and both arrays are 16-byte aligned).  SMEM: the whole histogram is first staged
// in shared memory with coalesced int4 loads that are all in flight at once, each thread then scans
// a contiguous chunk out of shared memory (odd chunk length in words: conflict-free), and the
// result leaves with coalesced stores -- a few microseconds for ~50k cells.  Histograms that do not
// fit take the same steps straight from global memory.
template <bool SMEM>
__global__ void __launch_bounds__(1024, 1)
cell_scan_kernel(int ncells, const int *__restrict__ cell_count, int *__restrict__ cell_start,
                 const PosQ *__restrict__ packed, int mpad, int nranks, double *__restrict__ qz_sum) {
  extern __shared__ __align__(16) int hist[];
  __shared__ int wsum[32];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (qz_sum && t == 0) {  // multi-GPU: sum(q z) partials ride in the last slot of every rank's block
    double v = 0.0;
    for (int r = 0; r < nranks; ++r) v += packed[(size_t)r * mpad + mpad - 1].x;
    *qz_sum = v;
  }
  const int nvec = (ncells + 3) >> 2;  // int4 groups in the histogram
  int per, i0, i1;
  int tsum = 0;
  if (SMEM) {
    const int4 *cnt4 = reinterpret_cast<const int4 *>(cell_count);
    int4 *h4 = reinterpret_cast<int4 *>(hist);
#pragma unroll 4
    for (int v = t; v < nvec; v += 1024) h4[v] = cnt4[v];
    __syncthreads();
    per = ((4 * nvec + 1023) >> 10) | 1;  // words per thread, odd
    i0 = min(t * per, 4 * nvec);
    i1 = min(i0 + per, 4 * nvec);
    for (int i = i0; i < i1; ++i) tsum += hist[i];
  } else {
    per = (nvec + 1023) >> 10;  // int4 groups per thread
    i0 = min(t * per, nvec);
    i1 = min(i0 + per, nvec);
    const int4 *cnt4 = reinterpret_cast<const int4 *>(cell_count);
    for (int v = i0; v < i1; ++v) {
      const int4 c = cnt4[v];
      tsum += (c.x + c.y) + (c.z + c.w);
    }
  }
  int incl = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const int w = wsum[lane];
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += u;
    }
    wsum[lane] = wi - w;
    if (lane == 31) cell_start[ncells] = wi;
  }
  __syncthreads();
  int run = wsum[warp] + incl - tsum;
  if (SMEM) {
    for (int i = i0; i < i1; ++i) {  // in place: count -> exclusive prefix
      const int c = hist[i];
      hist[i] = run;
      run += c;
    }
    __syncthreads();
    for (int i = t; i < ncells; i += 1024) cell_start[i] = hist[i];
  } else {
    const int4 *cnt4 = reinterpret_cast<const int4 *>(cell_count);
    int4 *out4 = reinterpret_cast<int4 *>(cell_start);
    for (int v = i0; v < i1; ++v) {
      const int4 c = cnt4[v];
      const int idx = 4 * v;
      const int4 o = make_int4(run, run + c.x, run + c.x + c.y, run + c.x + c.y + c.z);
      if (idx + 3 < ncells) {
        out4[v] = o;
      } else {  // tail group: cell_start[ncells] belongs to the grand total
        cell_start[idx] = o.x;
        if (idx + 1 < ncells) cell_start[idx + 1] = o.y;
        if (idx + 2 < ncells) cell_start[idx + 2] = o.z;
      }
      run = o.w + c.w;
    }
  }
}

__global__ void __launch_bounds__(256)
cell_scatter_kernel(CellGrid g, int m, const PosQ *__restrict__ packed, const int *__restrict__ type,
                    const int *__restrict__ cell_of, const int *__restrict__ slot,
                    const int *__restrict__ cell_start, PosQ *__restrict__ sorted,
                    int *__restrict__ sorted_type, int *__restrict__ sorted_src,
                    float4 *__restrict__ sorted_f) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  if (cell_of[j] < 0) return;  // padding slot
  const int d = cell_start[cell_of[j]] + slot[j];
  const PosQ p = packed[j];
  sorted[d] = p;
  sorted_type[d] = type[j];
  sorted_src[d] = j;
  // FP32 copy relative to the box corner for the conservative distance prefilter (pair_b)
  sorted_f[d] = make_float4((float)(p.x - g.lo[0]), (float)(p.y - g.lo[1]), (float)(p.z - g.lo[2]), 0.0f);
}

// near list over the gathered charges: warp-aggregated append
__global__ void __launch_bounds__(256)
near_list_kernel(CellGrid g, int m, const PosQ *__restrict__ packed, const unsigned char *__restrict__ near_mask,
                 int *__restrict__ near_list, int *__restrict__ near_count, const int *__restrict__ counts, int mpad) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  bool near = false;
  if (j < m && (!counts || j % mpad < counts[j / mpad])) {  // routed exchange: only the filled inbox slots
    const PosQ p = packed[j];
    const int cell = (cell_coord(g, 2, p.z) * g.nc[1] + cell_coord(g, 1, p.y)) * g.nc[0] + cell_coord(g, 0, p.x);
    near = near_mask[cell] != 0 && p.q != 0.0;
  }
  const unsigned mask = __ballot_sync(0xffffffffu, near);
  if (mask == 0) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == 0) base = atomicAdd(near_count, __popc(mask));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (near) near_list[base + __popc(mask & ((1u << lane) - 1u))] = j;
}

// ---------------------------------------------------------------------------
// warp traversal of the static electrode cell grid around a point
// ---------------------------------------------------------------------------
constexpr int PAIR_WARPS = 8;
constexpr int QCAP = 64;

struct WarpQueue {
  double rsq[QCAP];  // MODE_A: rsq            MODE_B: image shift x
  double aux[QCAP];  //                        MODE_B: image shift y
  double shz[QCAP];  //                        MODE_B: image shift z
  int i[QCAP];       // MODE_A: partner index  MODE_B: index into the sorted charges
  int t[QCAP];       // MODE_A: partner type
};

// Visit, warp-cooperatively, every electrode atom image within sqrt(max cuteff)
// of (xi,yi,zi).  F(lane-candidate) is called for all lanes of each chunk of 32
// consecutive sorted electrode atoms: f(valid, e, shx, shy, shz).
template <class F>
__device__ __forceinline__ void traverse(const CellGrid &g, const int *__restrict__ cell_start, double xi,
                                         double yi, double zi, int lane, F &&f) {
  const double rc = g.rc;
  // FP32 is enough for the conservative "row of cells beyond the cut-off sphere" test
  const float csy = (float)(g.prd[1] / g.nc[1]), csz = (float)(g.prd[2] / g.nc[2]);
  const float rc2f = (float)(rc * rc) * 1.0001f;
  for (int sz = -g.smax[2]; sz <= g.smax[2]; ++sz) {
    const double shz = sz * g.prd[2];
    int lz = (int)floor((zi - shz - rc - g.lo[2]) * g.cinv[2]);
    int hz = (int)floor((zi - shz + rc - g.lo[2]) * g.cinv[2]);
    if (g.periodic[2] && (hz < 0 || lz > g.nc[2] - 1)) continue;
    lz = max(0, min(lz, g.nc[2] - 1));
    hz = max(0, min(hz, g.nc[2] - 1));
    for (int sy = -g.smax[1]; sy <= g.smax[1]; ++sy) {
      const double shy = sy * g.prd[1];
      int ly = (int)floor((yi - shy - rc - g.lo[1]) * g.cinv[1]);
      int hy = (int)floor((yi - shy + rc - g.lo[1]) * g.cinv[1]);
      if (g.periodic[1] && (hy < 0 || ly > g.nc[1] - 1)) continue;
      ly = max(0, min(ly, g.nc[1] - 1));
      hy = max(0, min(hy, g.nc[1] - 1));
      for (int sx = -g.smax[0]; sx <= g.smax[0]; ++sx) {
        const double shx = sx * g.prd[0];
        int lx = (int)floor((xi - shx - rc - g.lo[0]) * g.cinv[0]);
        int hx = (int)floor((xi - shx + rc - g.lo[0]) * g.cinv[0]);
        if (g.periodic[0] && (hx < 0 || lx > g.nc[0] - 1)) continue;
        lx = max(0, min(lx, g.nc[0] - 1));
        hx = max(0, min(hx, g.nc[0] - 1));
        const float fy = (float)(yi - shy - g.lo[1]), fz = (float)(zi - shz - g.lo[2]);
        for (int cz = lz; cz <= hz; ++cz) {
          float dz = fmaxf(fmaxf(cz * csz - fz, fz - (cz + 1) * csz), 0.0f);
          if (!g.periodic[2] && (cz == 0 || cz == g.nc[2] - 1)) dz = 0.0f;  // edge cells hold clamped atoms
          for (int cy = ly; cy <= hy; ++cy) {
            float dy = fmaxf(fmaxf(cy * csy - fy, fy - (cy + 1) * csy), 0.0f);
            if (!g.periodic[1] && (cy == 0 || cy == g.nc[1] - 1)) dy = 0.0f;
            if (dy * dy + dz * dz > rc2f) continue;
            const int base = (cz * g.nc[1] + cy) * g.nc[0];
            const int jb = __ldg(cell_start + base + lx);
            const int je = __ldg(cell_start + base + hx + 1);
            for (int j0 = jb; j0 < je; j0 += 32) f(j0 + lane < je, j0 + lane, shx, shy, shz);
          }
        }
      }
    }
  }
}

// One warp per electrode row i of [row_begin, row_end).  Electrode atoms never move, so the
// x-contiguous runs of cells (and the periodic image shift of each) that the cut-off sphere of
// row i overlaps are enumerated once on the host (build_pair_runs); the kernel walks that flat list.
// MODE_B: targets = cell-sorted point charges; b_real[i] = -sum_j q_j dudq (fix_conp.cpp:1339).
//         Candidates are screened in FP32 (positions relative to the box corner, threshold
//         widened so no true pair is lost); survivors are re-tested exactly in FP64
//         against cutsq / cut_coulsq when they are consumed.
// MODE_A: targets = cell-sorted electrode atoms; A[i][j] += dudq (fix_conp.cpp:1270), self images kept
template <int MODE>
__global__ void __launch_bounds__(PAIR_WARPS * 32, 4)
pair_kernel(CellGrid g, PairTables pt, int row_begin, int row_end, const double *__restrict__ ex,
            const double *__restrict__ ey, const double *__restrict__ ez, const int *__restrict__ etype,
            const int *__restrict__ cell_start, const int *__restrict__ run_start,
            const PairRun *__restrict__ runs,
            const PosQ *__restrict__ sorted, const int *__restrict__ sorted_type,  // MODE_B targets
            const float4 *__restrict__ sorted_f, float cutmax_f,
            const EPos *__restrict__ esorted,                                      // MODE_A / MODE_P targets
            const double *__restrict__ q_ele,                                      // MODE_P: charges (eleall order)
            double *__restrict__ out, size_t pitch) {
  __shared__ WarpQueue queues[PAIR_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = row_begin + blockIdx.x * PAIR_WARPS + warp;
  if (i >= row_end) return;
  WarpQueue &wq = queues[warp];
  const unsigned lt_mask = (1u << lane) - 1u;
  const double xi = ex[i], yi = ey[i], zi = ez[i];
  const int it = etype[i];
  const double *cut_row = pt.cuteff + it * (pt.ntypes + 1);
  double *A_row = (MODE == MODE_A) ? out + (size_t)(i - row_begin) * pitch : nullptr;
  double acc = 0.0;
  int qn = 0;
  auto consume = [&](int e) {
    if (MODE == MODE_B) {
      // exact FP64 distance of the screened candidate
      const int k = wq.i[e];
      const PosQ p = sorted[k];
      const int jt = sorted_type[k];
      const double dx = xi - (p.x + wq.rsq[e]), dy = yi - (p.y + wq.aux[e]), dz = zi - (p.z + wq.shz[e]);
      const double rsq = dx * dx + dy * dy + dz * dz;
      if (rsq < __ldg(cut_row + jt)) acc = fma(p.q, dudq_pair<MODE_B>(pt, rsq, it, jt), acc);
    } else if (MODE == MODE_A) {
      atomicAdd(A_row + wq.i[e], dudq_pair<MODE_A>(pt, wq.rsq[e], it, wq.t[e]));
    } else {  // MODE_P: potential of the other electrode charges (compute_potential_atom.cpp:286-316)
      acc = fma(__ldg(q_ele + wq.i[e]), dudq_pair<MODE_A>(pt, wq.rsq[e], it, wq.t[e]), acc);
    }
  };
  // The runs of a row are walked as one flat sequence of chunks of 32 candidates.  Dependent loads are what this
  // kernel waits for (run -> cell_start -> candidates; ncu: 45 % occupancy, stalls on the load chain), so lane l
  // fetches run l's descriptor and target range up front (two load latencies for up to 32 runs instead of two
  // per run) and the candidates of the next chunk are requested before the current chunk is processed.
  constexpr unsigned FULL = 0xffffffffu;
  const int r0 = run_start[i - row_begin], r1 = run_start[i - row_begin + 1];
  for (int rb = r0; rb < r1; rb += 32) {
    const int nrun = min(32, r1 - rb);
    int4 mr = make_int4(0, 0, 0, 0);  // c0, c1, sx | sy << 16, sz
    int my_jb = 0, my_je = 0;
    if (lane < nrun) {
      mr = __ldg(reinterpret_cast<const int4 *>(runs + rb + lane));
      my_jb = __ldg(cell_start + mr.x);
      my_je = __ldg(cell_start + mr.y);
    }
    // (q, j0): chunk [j0, j0 + 32) of run q; je, sa, sb: the run's end and packed image shifts.  All warp-uniform.
    auto advance = [&](int &q, int &j0, int &je, int &sa, int &sb) {
      j0 += 32;
      while (j0 >= je) {
        if (++q >= nrun) return false;
        j0 = __shfl_sync(FULL, my_jb, q);
        je = __shfl_sync(FULL, my_je, q);
        sa = __shfl_sync(FULL, mr.z, q);
        sb = __shfl_sync(FULL, mr.w, q);
      }
      return true;
    };
    float4 cur_f = make_float4(0.f, 0.f, 0.f, 0.f), nxt_f = cur_f;
    EPos cur_e = {}, nxt_e = {};
    auto fetch = [&](int j0, int je, float4 &pf, EPos &pe) {
      const int k = j0 + lane;
      if (k < je) {
        if (MODE == MODE_B) pf = sorted_f[k];
        else pe = esorted[k];
      }
    };
    int q = -1, j0 = 0, je = 0, sa = 0, sb = 0;
    bool have = advance(q, j0, je, sa, sb);
    if (have) fetch(j0, je, cur_f, cur_e);
    while (have) {
      int nq = q, nj0 = j0, nje = je, nsa = sa, nsb = sb;
      const bool more = advance(nq, nj0, nje, nsa, nsb);
      if (more) fetch(nj0, nje, nxt_f, nxt_e);
      // ---- the current chunk ----
      const int isx = (short)(sa & 0xffff), isy = (short)(sa >> 16), isz = (short)(sb & 0xffff);
      const double shx = isx * g.prd[0], shy = isy * g.prd[1], shz = isz * g.prd[2];
      const int k = j0 + lane;
      bool pass = false;
      double rsq = 0.0;
      int jt = 0, jidx = 0;
      if (k < je) {
        if (MODE == MODE_B) {
          const float fxi = (float)(xi - shx - g.lo[0]), fyi = (float)(yi - shy - g.lo[1]),
                      fzi = (float)(zi - shz - g.lo[2]);
          const float fx = fxi - cur_f.x, fy = fyi - cur_f.y, fz = fzi - cur_f.z;
          pass = fx * fx + fy * fy + fz * fz < cutmax_f;
        } else {
          jt = cur_e.type; jidx = cur_e.idx;
          const double dx = xi - (cur_e.x + shx), dy = yi - (cur_e.y + shy), dz = zi - (cur_e.z + shz);
          rsq = dx * dx + dy * dy + dz * dz;
          pass = rsq < __ldg(cut_row + jt);
          if (jidx == i && isx == 0 && isy == 0 && isz == 0) pass = false;
        }
      }
      const unsigned mask = __ballot_sync(FULL, pass);
      if (pass) {
        const int pos = qn + __popc(mask & lt_mask);
        if (MODE == MODE_B) {
          wq.rsq[pos] = shx; wq.aux[pos] = shy; wq.shz[pos] = shz; wq.i[pos] = k;
        } else {
          wq.rsq[pos] = rsq; wq.i[pos] = jidx; wq.t[pos] = jt;
        }
      }
      qn += __popc(mask);
      if (qn >= 32) {
        __syncwarp();
        consume(qn - 32 + lane);
        qn -= 32;
        __syncwarp();
      }
      cur_f = nxt_f; cur_e = nxt_e;
      q = nq; j0 = nj0; je = nje; sa = nsa; sb = nsb;
      have = more;
    }
  }
  __syncwarp();
  if (lane < qn) consume(lane);
  if (MODE != MODE_A) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[i] = MODE == MODE_B ? -acc : acc;
  }
}

// post-force Gaussian correction; one warp per near charge, hits are rare
// (eta^2 r^2 < 5.8, fix_conp.cpp:1418-1419), so no queue.  The charge gets
// -del*forcecoul with del = electrode - charge (fix_conp.cpp:1425-1434).
__global__ void __launch_bounds__(PAIR_WARPS * 32)
pair_postforce_kernel(CellGrid g, PairTables pt, double qqrd2e, const EPos *__restrict__ esorted,
                      const int *__restrict__ cell_start, const double *__restrict__ q_ele,
                      const PosQ *__restrict__ packed, const int *__restrict__ packed_type,
                      const int *__restrict__ near_list, const int *__restrict__ near_count,
                      const double *__restrict__ cutsq_listed, double *__restrict__ f_packed,
                      double *__restrict__ energies, const int *__restrict__ psrc, int mpad) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps_total = gridDim.x * PAIR_WARPS;
  const int nwork = __ldg(near_count);
  double ecoul = 0.0, v0 = 0, v1 = 0, v2 = 0, v3 = 0, v4 = 0, v5 = 0;
  for (int w = blockIdx.x * PAIR_WARPS + warp; w < nwork; w += nwarps_total) {
    const int j = near_list[w];
    const PosQ p = packed[j];
    const int jt = packed_type[j];
    double fx = 0.0, fy = 0.0, fz = 0.0;
    traverse(g, cell_start, p.x, p.y, p.z, lane, [&](bool valid, int k, double shx, double shy, double shz) {
      if (!valid) return;
      const EPos e = esorted[k];
      // del = x_electrode - x_charge (the reference's delx with i = electrode)
      const double dx = (e.x + shx) - p.x, dy = (e.y + shy) - p.y, dz = (e.z + shz) - p.z;
      const double rsq = dx * dx + dy * dy + dz * dz;
      const int it = e.type;
      if (rsq < __ldg(cutsq_listed + it * (pt.ntypes + 1) + jt)) {  // fix_conp.cpp:1417
        const double etarij2 = pt.eta * pt.eta * rsq;                // :1418
        if (etarij2 < ERFC_MAX) {                                    // :1419 (sic: not squared)
          const double prefactor = qqrd2e * q_ele[e.idx] * p.q;
          double pf, pp;
          if (pt.pairmode == CONP_PAIR_EHGO) {
            const int ij = it * (pt.ntypes + 1) + jt;
            const double etaij = pt.eta_ij[ij], foij = pt.fo_ij[ij];
            const double e2 = etaij * etaij * rsq;
            pf = e2 * foij * exp(-0.5 * e2) - ferfcr_sqrt(e2) * etaij;  // ehgo_force :1568-1573
            pp = foij * exp(-0.5 * e2) - erfcr_sqrt(e2) * etaij;       // ehgo_potential
          } else {
            pf = -ferfcr_sqrt(etarij2) * pt.eta;  // eta_force :1477-1480
            pp = -erfcr_sqrt(etarij2) * pt.eta;   // eta_potential
          }
          const double forcecoul = prefactor * pf;
          const double fpair = forcecoul / rsq;
          fx -= dx * forcecoul; fy -= dy * forcecoul; fz -= dz * forcecoul;
          ecoul += prefactor * pp;  // ev_tally ecoul :1435-1436
          v0 += dx * dx * fpair; v1 += dy * dy * fpair; v2 += dz * dz * fpair;
          v3 += dx * dy * fpair; v4 += dx * dz * fpair; v5 += dy * dz * fpair;
        }
      }
    });
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      fx += __shfl_xor_sync(0xffffffffu, fx, o);
      fy += __shfl_xor_sync(0xffffffffu, fy, o);
      fz += __shfl_xor_sync(0xffffffffu, fz, o);
    }
    if (lane == 0 && (fx != 0.0 || fy != 0.0 || fz != 0.0)) {
      // routed exchange: inbox slot j of sender block j / mpad is that sender's charge psrc[j]
      const size_t fj = psrc ? (size_t)(j / mpad) * mpad + psrc[j] : (size_t)j;
      atomicAdd(f_packed + 3 * fj, fx);  // multi-GPU: each rank adds its own rows' share
      atomicAdd(f_packed + 3 * fj + 1, fy);
      atomicAdd(f_packed + 3 * fj + 2, fz);
    }
  }
  double vals[7] = {ecoul, v0, v1, v2, v3, v4, v5};
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    double v = vals[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && v != 0.0) atomicAdd(energies + (k == 0 ? 0 : k + 1), v);
  }
}

}  // namespace

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
CellGrid make_cell_grid(const double lo[3], const double prd[3], const int periodic[3], double rc) {
  CellGrid g;
  long long tot = 1;
  for (int a = 0; a < 3; ++a) {
    g.lo[a] = lo[a];
    g.prd[a] = prd[a];
    g.periodic[a] = periodic[a];
    int nc = (int)std::floor(prd[a] / (0.5 * rc));
    if (nc < 1) nc = 1;
    if (nc > 512) nc = 512;
    g.nc[a] = nc;
    g.cinv[a] = nc / prd[a];
    g.smax[a] = periodic[a] ? (int)std::ceil(rc / prd[a]) + 1 : 0;
    tot *= nc;
  }
  while (tot > (1LL << 22)) {
    int amax = 0;
    for (int a = 1; a < 3; ++a)
      if (g.nc[a] > g.nc[amax]) amax = a;
    tot /= g.nc[amax];
    g.nc[amax] = (g.nc[amax] + 1) / 2;
    g.cinv[amax] = g.nc[amax] / prd[amax];
    tot *= g.nc[amax];
  }
  g.rc = rc;
  g.ncells = (int)tot;
  return g;
}

namespace {
inline double h_wrap(const CellGrid &g, int a, double x) {
  if (g.periodic[a]) x -= std::floor((x - g.lo[a]) / g.prd[a]) * g.prd[a];
  return x;
}
inline int h_cell(const CellGrid &g, int a, double x) {
  int k = (int)std::floor((x - g.lo[a]) * g.cinv[a]);
  return std::max(0, std::min(k, g.nc[a] - 1));
}
}  // namespace

// Static traversal lists: for every electrode row of [begin, end) the x-contiguous runs of cells
// (as a [c0, c1) range of cell indices) and image shifts its cut-off sphere overlaps -- the same
// enumeration as the device-side traverse(), done once because electrodes never move.  Runs that
// are adjacent in cell order and share a shift are merged.
void build_pair_runs(const CellGrid &g, int begin, int end, const double *xyz, std::vector<int> &run_start,
                     std::vector<PairRun> &runs) {
  run_start.assign((size_t)std::max(end - begin, 0) + 1, 0);
  runs.clear();
  const double rc = g.rc;
  const double csy = g.prd[1] / g.nc[1], csz = g.prd[2] / g.nc[2];
  const double rc2 = rc * rc * 1.0001;
  for (int i = begin; i < end; ++i) {
    const double xi = xyz[3 * (size_t)i], yi = xyz[3 * (size_t)i + 1], zi = xyz[3 * (size_t)i + 2];
    for (int sz = -g.smax[2]; sz <= g.smax[2]; ++sz) {
      const double shz = sz * g.prd[2];
      int lz = (int)std::floor((zi - shz - rc - g.lo[2]) * g.cinv[2]);
      int hz = (int)std::floor((zi - shz + rc - g.lo[2]) * g.cinv[2]);
      if (g.periodic[2] && (hz < 0 || lz > g.nc[2] - 1)) continue;
      lz = std::max(0, std::min(lz, g.nc[2] - 1));
      hz = std::max(0, std::min(hz, g.nc[2] - 1));
      for (int sy = -g.smax[1]; sy <= g.smax[1]; ++sy) {
        const double shy = sy * g.prd[1];
        int ly = (int)std::floor((yi - shy - rc - g.lo[1]) * g.cinv[1]);
        int hy = (int)std::floor((yi - shy + rc - g.lo[1]) * g.cinv[1]);
        if (g.periodic[1] && (hy < 0 || ly > g.nc[1] - 1)) continue;
        ly = std::max(0, std::min(ly, g.nc[1] - 1));
        hy = std::max(0, std::min(hy, g.nc[1] - 1));
        for (int sx = -g.smax[0]; sx <= g.smax[0]; ++sx) {
          const double shx = sx * g.prd[0];
          int lx = (int)std::floor((xi - shx - rc - g.lo[0]) * g.cinv[0]);
          int hx = (int)std::floor((xi - shx + rc - g.lo[0]) * g.cinv[0]);
          if (g.periodic[0] && (hx < 0 || lx > g.nc[0] - 1)) continue;
          lx = std::max(0, std::min(lx, g.nc[0] - 1));
          hx = std::max(0, std::min(hx, g.nc[0] - 1));
          const double fy = yi - shy - g.lo[1], fz = zi - shz - g.lo[2];
          for (int cz = lz; cz <= hz; ++cz) {
            double dz = std::max(std::max(cz * csz - fz, fz - (cz + 1) * csz), 0.0);
            if (!g.periodic[2] && (cz == 0 || cz == g.nc[2] - 1)) dz = 0.0;  // edge cells hold clamped atoms
            for (int cy = ly; cy <= hy; ++cy) {
              double dy = std::max(std::max(cy * csy - fy, fy - (cy + 1) * csy), 0.0);
              if (!g.periodic[1] && (cy == 0 || cy == g.nc[1] - 1)) dy = 0.0;
              if (dy * dy + dz * dz > rc2) continue;
              const int base = (cz * g.nc[1] + cy) * g.nc[0];
              // this row of cells is at least (dy, dz) away: along x only the chord of the cut-off sphere
              int rlx = lx, rhx = hx;
              if (g.periodic[0]) {  // (a non-periodic x keeps its clamped edge cells: they hold far-away atoms too)
                const double xr = std::sqrt(std::max(rc2 - dy * dy - dz * dz, 0.0));
                const int tlx = (int)std::floor((xi - shx - xr - g.lo[0]) * g.cinv[0]);
                const int thx = (int)std::floor((xi - shx + xr - g.lo[0]) * g.cinv[0]);
                if (thx < 0 || tlx > g.nc[0] - 1) continue;
                rlx = std::max(lx, std::max(0, tlx));
                rhx = std::min(hx, std::min(thx, g.nc[0] - 1));
                if (rlx > rhx) continue;
              }
              PairRun run;
              run.c0 = base + rlx; run.c1 = base + rhx + 1;
              run.sx = (short)sx; run.sy = (short)sy; run.sz = (short)sz; run.pad = 0;
              const size_t first = (size_t)run_start[i - begin];
              if (runs.size() > first && runs.back().c1 == run.c0 && runs.back().sx == run.sx &&
                  runs.back().sy == run.sy && runs.back().sz == run.sz)
                runs.back().c1 = run.c1;  // contiguous in cell order: merge
              else
                runs.push_back(run);
            }
          }
        }
      }
    }
    run_start[(size_t)(i - begin) + 1] = (int)runs.size();
  }
  if (runs.empty()) runs.resize(1);
}

// Counting sort of electrode rows [begin, end) into grid g (host; electrodes are static).
void build_electrode_cells(const CellGrid &g, int begin, int end, const double *xyz, const int *type,
                           std::vector<EPos> &sorted, std::vector<int> &cell_start) {
  const int n = end - begin;
  cell_start.assign((size_t)g.ncells + 1, 0);
  std::vector<int> cell(std::max(n, 0));
  std::vector<EPos> tmp(std::max(n, 0));
  for (int k = 0; k < n; ++k) {
    const int i = begin + k;
    EPos e;
    e.x = h_wrap(g, 0, xyz[3 * (size_t)i]);
    e.y = h_wrap(g, 1, xyz[3 * (size_t)i + 1]);
    e.z = h_wrap(g, 2, xyz[3 * (size_t)i + 2]);
    e.idx = i;
    e.type = type[i];
    tmp[k] = e;
    cell[k] = (h_cell(g, 2, e.z) * g.nc[1] + h_cell(g, 1, e.y)) * g.nc[0] + h_cell(g, 0, e.x);
    cell_start[(size_t)cell[k] + 1]++;
  }
  for (int c = 0; c < g.ncells; ++c) cell_start[(size_t)c + 1] += cell_start[c];
  sorted.resize(std::max(n, 0));
  std::vector<int> cursor(cell_start.begin(), cell_start.end() - 1);
  for (int k = 0; k < n; ++k) sorted[cursor[cell[k]]++] = tmp[k];
}

// Cells from which a point charge can reach an electrode atom of rows
// [begin, end): every cell a traversal from inside it could need is covered
// by marking, for each electrode atom, the cells overlapping its rc-sphere's
// bounding box (all periodic shifts), dilated by one cell against rounding.
void build_near_mask(const CellGrid &g, int begin, int end, const double *xyz, std::vector<unsigned char> &mask) {
  mask.assign((size_t)g.ncells, 0);
  for (int i = begin; i < end; ++i) {
    int lo[3], hi[3];
    bool all[3];
    for (int a = 0; a < 3; ++a) {
      const double x = h_wrap(g, a, xyz[3 * (size_t)i + a]);
      lo[a] = (int)std::floor((x - g.rc - g.lo[a]) * g.cinv[a]) - 1;
      hi[a] = (int)std::floor((x + g.rc - g.lo[a]) * g.cinv[a]) + 1;
      all[a] = (hi[a] - lo[a] + 1 >= g.nc[a]);
    }
    for (int cz = lo[2]; cz <= hi[2]; ++cz) {
      int wz = cz;
      if (g.periodic[2]) { wz %= g.nc[2]; if (wz < 0) wz += g.nc[2]; }
      else if (wz < 0 || wz >= g.nc[2]) { if (all[2]) continue; wz = std::max(0, std::min(wz, g.nc[2] - 1)); }
      for (int cy = lo[1]; cy <= hi[1]; ++cy) {
        int wy = cy;
        if (g.periodic[1]) { wy %= g.nc[1]; if (wy < 0) wy += g.nc[1]; }
        else if (wy < 0 || wy >= g.nc[1]) { if (all[1]) continue; wy = std::max(0, std::min(wy, g.nc[1] - 1)); }
        for (int cx = lo[0]; cx <= hi[0]; ++cx) {
          int wx = cx;
          if (g.periodic[0]) { wx %= g.nc[0]; if (wx < 0) wx += g.nc[0]; }
          else if (wx < 0 || wx >= g.nc[0]) { if (all[0]) continue; wx = std::max(0, std::min(wx, g.nc[0] - 1)); }
          mask[((size_t)wz * g.nc[1] + wy) * g.nc[0] + wx] = 1;
        }
      }
    }
  }
}

int launch_pack_count(cudaStream_t s, const CellGrid &g, int m, const double *x_raw, const int *idx,
                      const double *q, const int *type, PosQ *packed, int *packed_type, int *cell_of, int *slot,
                      int *cell_count, double *qz_sum, const PeerSync &ps, size_t off_block, int mpad) {
  if (m <= 0 && !ps.arena) return 0;
  // with the fused exchange a rank without charges still signals: one (idle) block
  pack_count_kernel<<<std::max((m + 255) / 256, 1), 256, 0, s>>>(g, m, x_raw, idx, q, type, packed, packed_type,
                                                                 cell_of, slot, cell_count, qz_sum, ps, off_block, mpad);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_bin_positions(cudaStream_t s, const CellGrid &g, int m, int mpad, const int *counts,
                         const PosQ *packed, int *cell_of, int *slot, int *cell_count,
                         const unsigned char *relevant, const PeerSync &ps) {
  if (m <= 0 && !ps.arena) return 0;
  bin_positions_kernel<<<std::max((m + 255) / 256, 1), 256, 0, s>>>(g, m, mpad, counts, packed, cell_of, slot,
                                                                    cell_count, relevant, ps);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_cell_scan(cudaStream_t s, int ncells, const int *cell_count, int *cell_start, const PosQ *packed,
                     int mpad, int nranks, double *qz_sum) {
  const size_t smem = sizeof(int) * 4 * (size_t)((ncells + 3) / 4);
  if (smem <= 200 * 1024) {
    ensure_dynamic_smem(cell_scan_kernel<true>, 200 * 1024);
    cell_scan_kernel<true><<<1, 1024, smem, s>>>(ncells, cell_count, cell_start, packed, mpad, nranks, qz_sum);
  } else {
    cell_scan_kernel<false><<<1, 1024, 0, s>>>(ncells, cell_count, cell_start, packed, mpad, nranks, qz_sum);
  }
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_cell_scatter(cudaStream_t s, const CellGrid &g, int m, const PosQ *packed, const int *type,
                        const int *cell_of, const int *slot, const int *cell_start, PosQ *sorted, int *sorted_type,
                        int *sorted_src, float4 *sorted_f) {
  if (m <= 0) return 0;
  cell_scatter_kernel<<<(m + 255) / 256, 256, 0, s>>>(g, m, packed, type, cell_of, slot, cell_start, sorted,
                                                      sorted_type, sorted_src, sorted_f);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_pack_route(cudaStream_t s, const CellGrid &g, int m, const double *x_raw, const int *idx, const double *q,
                      const int *type, PosQ *own, int rank, int nranks, int mpad, const unsigned char *rel_all,
                      const PeerSync &ps, size_t off_packed, size_t off_ptype, size_t off_psrc, int *send_count,
                      double *qz_sum) {
  pack_route_kernel<<<std::max((m + 255) / 256, 1), 256, 0, s>>>(g, m, x_raw, idx, q, type, own, rank, nranks, mpad,
                                                                 rel_all, ps.arena, off_packed, off_ptype, off_psrc,
                                                                 send_count, qz_sum);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_near_list(cudaStream_t s, const CellGrid &g, int m, const PosQ *packed, const unsigned char *near_mask,
                     int *near_list, int *near_count, const int *counts, int mpad) {
  if (m <= 0) return 0;
  near_list_kernel<<<(m + 255) / 256, 256, 0, s>>>(g, m, packed, near_mask, near_list, near_count, counts, mpad);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_pair_b(cudaStream_t s, const CellGrid &g, const PairTables &pt, int row_begin, int row_end,
                  const double *ex, const double *ey, const double *ez, const int *etype, const int *run_start,
                  const PairRun *runs, const PosQ *sorted, const int *sorted_type, const float4 *sorted_f,
                  const int *cell_start, double *b_real) {
  const int n = row_end - row_begin;
  if (n <= 0) return 0;
  // FP32 screen radius: the grid's search radius widened so that rounding cannot reject a true pair.
  // Coordinates relative to the box corner carry <= 2^-23 * extent each; the squared distance then
  // errs by <= 2 sqrt(3) rc * (4 * 2^-23 * extent), doubled for safety.
  const double extent = std::max(std::max(g.prd[0], g.prd[1]), g.prd[2]) + 2.0 * g.rc;
  const double slack = 2.0 * (2.0 * std::sqrt(3.0) * g.rc * 4.0 * extent * 1.1920929e-7) + 1e-4 * g.rc * g.rc;
  const float cutmax_f = (float)(g.rc * g.rc + slack);
  // (Capping the blocks per SM with unused dynamic shared memory, to leave room for the k-space chain's blocks
  // beside this kernel, was measured: 2 or 3 blocks per SM cost 3-6 % of the step at cfg4 and on two GPUs.)
  pair_kernel<MODE_B><<<(n + PAIR_WARPS - 1) / PAIR_WARPS, PAIR_WARPS * 32, 0, s>>>(
      g, pt, row_begin, row_end, ex, ey, ez, etype, cell_start, run_start, runs, sorted, sorted_type, sorted_f,
      cutmax_f, nullptr, nullptr, b_real, 0);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_pair_A(cudaStream_t s, const CellGrid &g, const PairTables &pt, const EPos *esorted,
                  const int *cell_start, int row_begin, int row_end, const double *ex, const double *ey,
                  const double *ez, const int *etype, const int *run_start, const PairRun *runs, double *A_rows,
                  size_t pitch) {
  const int n = row_end - row_begin;
  if (n <= 0) return 0;
  pair_kernel<MODE_A><<<(n + PAIR_WARPS - 1) / PAIR_WARPS, PAIR_WARPS * 32, 0, s>>>(
      g, pt, row_begin, row_end, ex, ey, ez, etype, cell_start, run_start, runs, nullptr, nullptr, nullptr, 0.0f,
      esorted, nullptr, A_rows, pitch);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_pair_P(cudaStream_t s, const CellGrid &g, const PairTables &pt, const EPos *esorted,
                  const int *cell_start, int row_begin, int row_end, const double *ex, const double *ey,
                  const double *ez, const int *etype, const int *run_start, const PairRun *runs,
                  const double *q_ele, double *phi) {
  const int n = row_end - row_begin;
  if (n <= 0) return 0;
  pair_kernel<MODE_P><<<(n + PAIR_WARPS - 1) / PAIR_WARPS, PAIR_WARPS * 32, 0, s>>>(
      g, pt, row_begin, row_end, ex, ey, ez, etype, cell_start, run_start, runs, nullptr, nullptr, nullptr, 0.0f,
      esorted, q_ele, phi, 0);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_pair_postforce(cudaStream_t s, const CellGrid &g, const PairTables &pt, double qqrd2e,
                          const EPos *esorted, const int *cell_start, const double *q_ele, const PosQ *packed,
                          const int *packed_type, const int *near_list, const int *near_count, int max_near,
                          const double *cutsq_listed, double *f_packed, double *energies, int num_sms,
                          const int *psrc, int mpad) {
  if (max_near <= 0) return 0;
  int grid = std::min((max_near + PAIR_WARPS - 1) / PAIR_WARPS, num_sms * 8);
  pair_postforce_kernel<<<grid, PAIR_WARPS * 32, 0, s>>>(g, pt, qqrd2e, esorted, cell_start, q_ele, packed,
                                                         packed_type, near_list, near_count, cutsq_listed,
                                                         f_packed, energies, psrc, mpad);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

}  // namespace conp
