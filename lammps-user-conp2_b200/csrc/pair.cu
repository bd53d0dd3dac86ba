// Real-space electrode<->point-charge kernels and the cell binning that
// replaces LAMMPS' neighbour lists on the device.
//
//   pair_b         FixConp::blist_coul_cal        fix_conp.cpp:1281-1365
//   pair_A         FixConp::alist_coul_cal        fix_conp.cpp:1209-1279
//   pair_postforce FixConp::blist_coul_cal_post_force fix_conp.cpp:1368-1444
//   erfcr_sqrt / ferfcr_sqrt / eta_* / ehgo_*     fix_conp.cpp:1446-1480, 1561-1573
//
// The pair set is geometric (rsq < cutsq[it][jt] and rsq < cut_coulsq,
// fix_conp.cpp:1333-1334) over *all* periodic images, which is what LAMMPS'
// ghost atoms provide.  Point charges are counting-sorted into a uniform
// cell grid every step; one warp owns one electrode atom, lanes stride the
// x-contiguous cell runs (coalesced 32-byte PosQ loads), and the candidates
// that pass the distance test are compacted through a per-warp shared-memory
// queue so the expensive FP64 erfc evaluation always runs on full warps.
#include "common.cuh"

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

enum { MODE_B = 0, MODE_A = 1 };

// dudq of fix_conp.cpp:1263-1264 / 1335-1336
template <int MODE>
__device__ __forceinline__ double dudq_pair(const PairTables &pt, double rsq, int it, int jt) {
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
// binning
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_count_kernel(CellGrid g, int m, const double *__restrict__ x_raw, const int *__restrict__ idx,
                  const double *__restrict__ q, const int *__restrict__ type, PosQ *__restrict__ packed,
                  int *__restrict__ packed_type, int *__restrict__ cell_of, int *__restrict__ slot,
                  int *__restrict__ cell_count, double *__restrict__ qz_sum) {
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
  }
  // block reduction of q*z
  __shared__ double sh[8];
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
}

__global__ void __launch_bounds__(256)
bin_positions_kernel(CellGrid g, int m, const PosQ *__restrict__ packed, int *__restrict__ cell_of,
                     int *__restrict__ slot, int *__restrict__ cell_count) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const PosQ p = packed[j];
  const int cell = (cell_coord(g, 2, p.z) * g.nc[1] + cell_coord(g, 1, p.y)) * g.nc[0] + cell_coord(g, 0, p.x);
  cell_of[j] = cell;
  slot[j] = atomicAdd(&cell_count[cell], 1);
}

// exclusive scan of cell_count -> cell_start[ncells+1]; one block, each thread
// owns a contiguous run of cells
__global__ void __launch_bounds__(1024, 1)
cell_scan_kernel(int ncells, const int *__restrict__ cell_count, int *__restrict__ cell_start) {
  __shared__ int sh[33];
  const int t = threadIdx.x;
  const int per = (ncells + 1023) / 1024;
  const int a = min(t * per, ncells), b = min(a + per, ncells);
  int sum = 0;
  for (int c = a; c < b; ++c) sum += cell_count[c];
  // block exclusive scan of `sum`
  const int lane = t & 31, warp = t >> 5;
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) sh[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = sh[lane];
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += v;
    }
    sh[lane] = wi - w;  // exclusive warp offsets
    if (lane == 31) sh[32] = wi;
  }
  __syncthreads();
  int run = sh[warp] + incl - sum;
  for (int c = a; c < b; ++c) {
    cell_start[c] = run;
    run += cell_count[c];
  }
  if (t == 0) cell_start[ncells] = sh[32];
}

__global__ void __launch_bounds__(256)
cell_scatter_kernel(int m, const PosQ *__restrict__ packed, const int *__restrict__ type,
                    const int *__restrict__ cell_of, const int *__restrict__ slot,
                    const int *__restrict__ cell_start, PosQ *__restrict__ sorted,
                    int *__restrict__ sorted_type, int *__restrict__ sorted_src) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const int d = cell_start[cell_of[j]] + slot[j];
  sorted[d] = packed[j];
  sorted_type[d] = type[j];
  if (sorted_src) sorted_src[d] = j;
}

// ---------------------------------------------------------------------------
// warp-per-electrode-atom traversal with candidate compaction
// ---------------------------------------------------------------------------
constexpr int PAIR_WARPS = 8;
constexpr int QCAP = 64;

struct WarpQueue {
  double rsq[QCAP];
  double q[QCAP];
  int j[QCAP];
  int t[QCAP];
};

template <int MODE>
__device__ __forceinline__ void consume(const PairTables &pt, const WarpQueue &wq, int e, int it, int i_global,
                                        double &acc, double *A_row) {
  const double rsq = wq.rsq[e];
  const double d = dudq_pair<MODE>(pt, rsq, it, wq.t[e]);
  if (MODE == MODE_B) {
    acc = fma(wq.q[e], d, acc);
  } else {
    (void)i_global;
    atomicAdd(A_row + wq.j[e], d);
  }
}

template <int MODE>
__global__ void __launch_bounds__(PAIR_WARPS * 32)
pair_kernel(CellGrid g, PairTables pt, int row_begin, int row_end, const double *__restrict__ ex,
            const double *__restrict__ ey, const double *__restrict__ ez, const int *__restrict__ etype,
            const PosQ *__restrict__ sorted, const int *__restrict__ sorted_type,
            const int *__restrict__ sorted_src, const int *__restrict__ cell_start, double *__restrict__ out,
            size_t pitch) {
  __shared__ WarpQueue queues[PAIR_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = row_begin + blockIdx.x * PAIR_WARPS + warp;
  if (i >= row_end) return;
  WarpQueue &wq = queues[warp];
  const double xi = ex[i], yi = ey[i], zi = ez[i];
  const int it = etype[i];
  const double *cut_row = pt.cuteff + it * (pt.ntypes + 1);
  const double rc = g.rc;
  double acc = 0.0;
  int qn = 0;
  double *A_row = (MODE == MODE_A) ? out + (size_t)(i - row_begin) * pitch : nullptr;
  const unsigned lt_mask = (1u << lane) - 1u;

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
        const bool zero_shift = (sx == 0 && sy == 0 && sz == 0);
        for (int cz = lz; cz <= hz; ++cz) {
          for (int cy = ly; cy <= hy; ++cy) {
            const int base = (cz * g.nc[1] + cy) * g.nc[0];
            const int jb = __ldg(cell_start + base + lx);
            const int je = __ldg(cell_start + base + hx + 1);
            for (int j0 = jb; j0 < je; j0 += 32) {
              const int j = j0 + lane;
              bool pass = false;
              double rsq = 0.0, qj = 0.0;
              int jt = 0, jsrc = 0;
              if (j < je) {
                const PosQ p = sorted[j];
                jt = sorted_type[j];
                const double dx = xi - (p.x + shx);
                const double dy = yi - (p.y + shy);
                const double dz = zi - (p.z + shz);
                rsq = dx * dx + dy * dy + dz * dz;
                qj = p.q;
                pass = rsq < __ldg(cut_row + jt);
                if (MODE == MODE_A) {
                  jsrc = sorted_src[j];
                  if (zero_shift && jsrc == i) pass = false;  // no self pair; self images are kept
                }
              }
              const unsigned mask = __ballot_sync(0xffffffffu, pass);
              if (pass) {
                const int pos = qn + __popc(mask & lt_mask);
                wq.rsq[pos] = rsq;
                wq.q[pos] = qj;
                wq.j[pos] = jsrc;
                wq.t[pos] = jt;
              }
              qn += __popc(mask);
              if (qn >= 32) {
                __syncwarp();
                consume<MODE>(pt, wq, qn - 32 + lane, it, i, acc, A_row);
                qn -= 32;
                __syncwarp();
              }
            }
          }
        }
      }
    }
  }
  __syncwarp();
  if (lane < qn) consume<MODE>(pt, wq, lane, it, i, acc, A_row);
  if (MODE == MODE_B) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[i] = -acc;  // m[elei] -= q[j]*dudq, fix_conp.cpp:1339
  }
}

// post-force Gaussian correction; hits are rare (eta^2 r^2 < 5.8), so no queue
__global__ void __launch_bounds__(PAIR_WARPS * 32)
pair_postforce_kernel(CellGrid g, PairTables pt, double qqrd2e, int row_begin, int row_end,
                      const double *__restrict__ ex, const double *__restrict__ ey,
                      const double *__restrict__ ez, const int *__restrict__ etype,
                      const double *__restrict__ q_ele, const PosQ *__restrict__ sorted,
                      const int *__restrict__ sorted_type, const int *__restrict__ sorted_src,
                      const int *__restrict__ cell_start, const double *__restrict__ cutsq_listed,
                      double *__restrict__ f_packed, double *__restrict__ energies) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = row_begin + blockIdx.x * PAIR_WARPS + warp;
  double ecoul = 0.0, v0 = 0, v1 = 0, v2 = 0, v3 = 0, v4 = 0, v5 = 0;
  if (i < row_end) {
    const double xi = ex[i], yi = ey[i], zi = ez[i], qi = q_ele[i];
    const int it = etype[i];
    const double *cut_row = cutsq_listed + it * (pt.ntypes + 1);
    const double rc = g.rc;
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
          for (int cz = lz; cz <= hz; ++cz)
            for (int cy = ly; cy <= hy; ++cy) {
              const int base = (cz * g.nc[1] + cy) * g.nc[0];
              const int jb = cell_start[base + lx], je = cell_start[base + hx + 1];
              for (int j = jb + lane; j < je; j += 32) {
                const PosQ p = sorted[j];
                const int jt = sorted_type[j];
                const double dx = xi - (p.x + shx), dy = yi - (p.y + shy), dz = zi - (p.z + shz);
                const double rsq = dx * dx + dy * dy + dz * dz;
                if (rsq < cut_row[jt]) {                      // fix_conp.cpp:1417
                  const double etarij2 = pt.eta * pt.eta * rsq;  // :1418
                  if (etarij2 < ERFC_MAX) {                    // :1419 (sic: not squared)
                    const double prefactor = qqrd2e * qi * p.q;
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
                    // del = electrode - electrolyte; the electrolyte atom gets -del*forcecoul (:1425-1434)
                    const int src = sorted_src[j];
                    atomicAdd(f_packed + 3 * (size_t)src, -dx * forcecoul);
                    atomicAdd(f_packed + 3 * (size_t)src + 1, -dy * forcecoul);
                    atomicAdd(f_packed + 3 * (size_t)src + 2, -dz * forcecoul);
                    ecoul += prefactor * pp;  // ev_tally ecoul :1435-1436
                    v0 += dx * dx * fpair; v1 += dy * dy * fpair; v2 += dz * dz * fpair;
                    v3 += dx * dy * fpair; v4 += dx * dz * fpair; v5 += dy * dz * fpair;
                  }
                }
              }
            }
        }
      }
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
  // bound the cell count (memory for cell_start and the single-block scan)
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

int launch_pack_count(cudaStream_t s, const CellGrid &g, int m, const double *x_raw, const int *idx,
                      const double *q, const int *type, PosQ *packed, int *packed_type, int *cell_of, int *slot,
                      int *cell_count, double *qz_sum) {
  if (m <= 0) return 0;
  pack_count_kernel<<<(m + 255) / 256, 256, 0, s>>>(g, m, x_raw, idx, q, type, packed, packed_type, cell_of, slot,
                                                    cell_count, qz_sum);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_bin_positions(cudaStream_t s, const CellGrid &g, int m, const PosQ *packed, int *cell_of, int *slot,
                         int *cell_count) {
  if (m <= 0) return 0;
  bin_positions_kernel<<<(m + 255) / 256, 256, 0, s>>>(g, m, packed, cell_of, slot, cell_count);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_cell_scan(cudaStream_t s, int ncells, const int *cell_count, int *cell_start) {
  cell_scan_kernel<<<1, 1024, 0, s>>>(ncells, cell_count, cell_start);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_cell_scatter(cudaStream_t s, int m, const PosQ *packed, const int *type, const int *cell_of,
                        const int *slot, const int *cell_start, PosQ *sorted, int *sorted_type, int *sorted_src) {
  if (m <= 0) return 0;
  cell_scatter_kernel<<<(m + 255) / 256, 256, 0, s>>>(m, packed, type, cell_of, slot, cell_start, sorted,
                                                      sorted_type, sorted_src);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_pair_b(cudaStream_t s, const CellGrid &g, const PairTables &pt, int row_begin, int row_end,
                  const double *ex, const double *ey, const double *ez, const int *etype, const PosQ *sorted,
                  const int *sorted_type, const int *cell_start, double *b_real) {
  const int n = row_end - row_begin;
  if (n <= 0) return 0;
  pair_kernel<MODE_B><<<(n + PAIR_WARPS - 1) / PAIR_WARPS, PAIR_WARPS * 32, 0, s>>>(
      g, pt, row_begin, row_end, ex, ey, ez, etype, sorted, sorted_type, nullptr, cell_start, b_real, 0);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_pair_A(cudaStream_t s, const CellGrid &g, const PairTables &pt, int row_begin, int row_end,
                  const double *ex, const double *ey, const double *ez, const int *etype, const PosQ *sorted,
                  const int *sorted_type, const int *sorted_src, const int *cell_start, double *A_rows,
                  size_t pitch) {
  const int n = row_end - row_begin;
  if (n <= 0) return 0;
  pair_kernel<MODE_A><<<(n + PAIR_WARPS - 1) / PAIR_WARPS, PAIR_WARPS * 32, 0, s>>>(
      g, pt, row_begin, row_end, ex, ey, ez, etype, sorted, sorted_type, sorted_src, cell_start, A_rows, pitch);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_pair_postforce(cudaStream_t s, const CellGrid &g, const PairTables &pt, double qqrd2e, int row_begin,
                          int row_end, const double *ex, const double *ey, const double *ez, const int *etype,
                          const double *q_ele, const PosQ *sorted, const int *sorted_type, const int *sorted_src,
                          const int *cell_start, const double *cutsq_listed, double *f_packed, double *energies) {
  const int n = row_end - row_begin;
  if (n <= 0) return 0;
  pair_postforce_kernel<<<(n + PAIR_WARPS - 1) / PAIR_WARPS, PAIR_WARPS * 32, 0, s>>>(
      g, pt, qqrd2e, row_begin, row_end, ex, ey, ez, etype, q_ele, sorted, sorted_type, sorted_src, cell_start,
      cutsq_listed, f_packed, energies);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

}  // namespace conp
