// PPPM-mode k-space kernels (kspace_style pppm/conp):
//   pppm_spread       PPPMCONP::elyte_particle_map + elyte_make_rho  pppm_conp.cpp:126-228
//   pppm_green_mul    PPPMCONP::elyte_poisson (the pointwise part)   pppm_conp.cpp:242-249
//   pppm_ele_stencil  PPPMCONP::aaa_map_rho                          pppm_conp.cpp:318-344
//   pppm_gather_b     PPPMCONP::b_cal gather + slab term             pppm_conp.cpp:278-314
//   pppm_ele_spread   PPPMCONP::ele_make_rho                         pppm_conp.cpp:385-426
//   pppm_zconv        PPPMCONP::elyte_poisson along z                pppm_conp.cpp:230-267
// The mesh is one global periodic brick [nz][ny][nx] (x fastest); periodic
// wrap of the stencil index replaces LAMMPS' ghost-cell exchange
// (gc->reverse_comm_kspace / forward_comm_kspace, pppm_conp.cpp:114-123).
//
// Plane pruning.  The reference transforms the whole mesh (3-D FFT, multiply
// by greensfn, inverse 3-D FFT).  Here only the z-planes that can hold charge
// ("input planes": the box plus the stencil reach; in slab geometry 1/slab of
// the mesh) are transformed in (x,y), and the potential is produced only on
// the planes the cached electrode stencils read ("output planes", static).
// Along z the FFT -> greensfn -> inverse FFT chain is applied exactly as the
// circular convolution it is:
//     u^(kx,ky;zo) = sum_zi K(kx,ky;(zo-zi) mod nz) rho^(kx,ky;zi),
//     K(kx,ky;d)   = sum_kz greensfn(kx,ky,kz)/(nx ny nz) exp(+2 pi i kz d/nz)
// (K is tabulated once at setup).  The result equals the full transform up to
// rounding; zero planes contribute exact zeros.  Bricks are stored compactly:
// density [nzi][ny][nx] with plane zi <-> (zin_lo + zi) mod nz, potential and
// electrode density [nzo][ny][nx] with plane index zmap[mz].
// The 2-D FFTs are cuFFT (library); everything else is hand-written.
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace conp {

namespace {

constexpr int OFFSET = 16384;  // pppm_conp.cpp:32
constexpr int SPREAD_BLOCKS_PER_SM = 0;  // 0: no cap (see launch_pppm_spread)
constexpr int MAXORDER = 7;    // LAMMPS PPPM MAXORDER

__device__ __forceinline__ int wrapi(int m, int n) {
  m %= n;
  return m < 0 ? m + n : m;
}

template <int BYTES>
__device__ __forceinline__ void cp_async(void *dst_smem, const void *src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(src), "n"(BYTES) : "memory");
}

// LAMMPS PPPM::compute_rho1d for one axis (Horner in rho_coeff, same loop as
// pppm_conp_intel.cpp:290-301); rc is [order][order] with k shifted by -nlower
__device__ __forceinline__ double rho1d(const double *__restrict__ rc, int order, int k, double d) {
  double r = 0.0;
  for (int l = order - 1; l >= 0; --l) r = rc[l * order + k] + r * d;
  return r;
}

// Clears of the step's bricks.  (A cudaMemsetAsync node inside the captured step ran at ~0.4 TB/s -- 45 MB of
// slab brick held the spread back by ~50 us on two GPUs; this streams at HBM store speed.)
__global__ void __launch_bounds__(256)
fill_zero_kernel(double *__restrict__ p, size_t n) {
  const size_t head = ((reinterpret_cast<uintptr_t>(p) & 15) && n) ? 1 : 0;  // up to the first 16-byte boundary
  const size_t n2 = (n - head) / 2;
  double2 *p2 = reinterpret_cast<double2 *>(p + head);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x)
    p2[i] = make_double2(0.0, 0.0);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (head) p[0] = 0.0;
    if ((n - head) & 1) p[n - 1] = 0.0;
  }
}

// one thread per (atom, z-plane n, y-row m); the thread adds its `order`
// x-consecutive mesh points with red.global.add.f64.  (A point-per-thread
// mapping that puts the `order` x-neighbours in adjacent lanes was measured 3x
// slower: same-sector atomics of one warp instruction serialise in L2.)
__global__ void __launch_bounds__(256)
spread_kernel(PPPMGeom g, const double *__restrict__ rho_coeff, int m_atoms, const PosQ *__restrict__ atoms,
              const int *__restrict__ inbox_counts, int nsenders, int mpad, const int *__restrict__ valid,
              double *__restrict__ brick, int *__restrict__ range_flag) {
  __shared__ double rc[MAXORDER * MAXORDER];
  __shared__ int cnt_end[17];  // inbox mode: running totals of the per-sender counts
  const int order = g.order;
  for (int t = threadIdx.x; t < order * order; t += blockDim.x) rc[t] = rho_coeff[t];
  if (inbox_counts && threadIdx.x == 0) {
    int tot = 0;
    for (int r = 0; r < nsenders; ++r) { tot += min(inbox_counts[r], mpad); cnt_end[r] = tot; }
  }
  __syncthreads();
  // one GPU: atoms[0 .. m_atoms); several (inbox mode): the charges as they arrived, sender r's block starts at
  // slot r * mpad and holds inbox_counts[r] (planes outside this rank's slab are skipped below)
  const int jb = 0;
  const int je = inbox_counts ? cnt_end[nsenders - 1] : m_atoms;
  const int per_atom = order * order;
  const long long work = (long long)(je - jb) * per_atom;
  for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < work;
       gid += (long long)gridDim.x * blockDim.x) {
    int j = jb + (int)(gid / per_atom);
    const int nm = (int)(gid % per_atom);
    const int n = nm / order, m = nm - n * order;
    if (inbox_counts) {
      int r = 0;
      while (j >= cnt_end[r]) ++r;
      j = r * mpad + j - (r ? cnt_end[r - 1] : 0);
      if (valid && valid[j] < 0) continue;  // not in a cell this rank reads
    }
    const PosQ p = atoms[j];
    if (p.q == 0.0) continue;  // pppm_conp.cpp:161
    const double fx = (p.x - g.boxlo[0]) * g.delinv[0];
    const double fy = (p.y - g.boxlo[1]) * g.delinv[1];
    const double fz = (p.z - g.boxlo[2]) * g.delinv[2];
    if (!(fabs(fx) < OFFSET / 2 && fabs(fy) < OFFSET / 2 && fabs(fz) < OFFSET / 2)) {
      *range_flag = 1;  // "Out of range atoms - cannot compute PPPM", pppm_conp.cpp:167
      continue;
    }
    const int nx = (int)(fx + g.shift) - OFFSET;  // :146-148
    const int ny = (int)(fy + g.shift) - OFFSET;
    const int nz = (int)(fz + g.shift) - OFFSET;
    const int zi = wrapi(n + g.nlower + nz - g.zin_lo, g.nz);  // compact input plane
    if (zi >= g.nzi) {
      *range_flag = 1;  // outside the planes the box can reach: "Out of range atoms"
      continue;
    }
    const int t = zi - g.zs_lo;  // plane inside this rank's slab?
    if (t < 0 || t >= g.zs_n) continue;
    const double dx = nx + g.shiftone - fx;  // :199-201
    const double dy = ny + g.shiftone - fy;
    const double dz = nz + g.shiftone - fz;
    const double z0 = g.delvolinv * p.q;  // :205
    const double y0 = z0 * rho1d(rc, order, n, dz);
    const double x0 = y0 * rho1d(rc, order, m, dy);
    const int my = wrapi(m + g.nlower + ny, g.ny);
    double *row = brick + ((size_t)t * g.ny + my) * g.nx;
    int mx = wrapi(g.nlower + nx, g.nx);
    for (int l = 0; l < order; ++l) {
      atomicAdd(row + mx, x0 * rho1d(rc, order, l, dx));
      mx = (mx + 1 == g.nx) ? 0 : mx + 1;
    }
  }
}

// ---------------------------------------------------------------------------
// Owner-computes spread, tile form (CONP_SPREAD=smem): no atomics, every mesh point stored exactly once.
//
// The rank's slab of input planes is cut into tiles of tz x ty x tx points.  One warp (= one CTA) owns a
// tile at a time, accumulated in its private shared memory: a warp never shares a tile, so plain
// load / FMA / store replaces the 125 red.global per charge of spread_kernel (the SM's red issue rate,
// ~1.3 cycles per lane, was that kernel's bound; shared-memory atomics are slower still).  For every tile the
// host has listed the x-contiguous runs of cells whose charges can reach it (plan_pppm_spread_tiles); the
// warp reads those charges 32 at a time, each lane works out its charge's stencil origin relative to the
// tile and its 3 x order weights (into a per-lane scratch row), and the charges that overlap are then
// visited one after another by the whole warp: lane (m, l) owns the point (y0 + m, x0 + l) of the
// stencil footprint and walks the valid z-planes (a charge straddling a tile border is visited by both
// tiles, each adding only its own points).  Positions inside the tile are affine in the stencil index:
// a tile that spans a whole periodic axis carries order-1 halo entries that are folded back before the
// store, any other tile is so much shorter than the mesh that the valid indices form one interval.
// Summation order is the order of the sorted charges: deterministic for a given sort.
// ---------------------------------------------------------------------------
// RS_ / PS_: compile-time row / plane strides of the default tile shape (0: run-time strides from the plan),
// so that the five plane accesses of a visit are immediate offsets from one address.
struct RhoCoeff {  // by-value kernel argument: the Horner coefficients come from the constant bank
  double c[MAXORDER * MAXORDER];
};

template <int P>
__device__ __forceinline__ double rho1d_c(const RhoCoeff &rc, int k, double d) {
  double r = 0.0;
#pragma unroll
  for (int l = P - 1; l >= 0; --l) r = rc.c[l * P + k] + r * d;
  return r;
}

// v in [-2n, 2n) -> [0, n) without an integer division
__device__ __forceinline__ int wrap2(int v, int n) {
  v += (v < 0) ? n : 0;
  v += (v < 0) ? n : 0;
  v -= (v >= n) ? n : 0;
  return v;
}

template <int P, int RS_, int PS_>
__global__ void __launch_bounds__(32)
spread_tile_kernel(PPPMGeom g, SpreadPlan sp, RhoCoeff rc, const PosQ *__restrict__ atoms,
                   const int *__restrict__ cell_start, double *__restrict__ brick, int *__restrict__ range_flag) {
  constexpr int P2 = P * P;
  constexpr int NR = (P2 + 31) / 32;
  constexpr int WS = (3 * P) | 1;  // odd stride: the 32 scratch rows fall into distinct banks
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char spt_smem[];
  double *tile = reinterpret_cast<double *>(spt_smem);
  const int az = sp.tz + sp.halo_z, ay = sp.ty + sp.halo_y;  // allocated planes / rows
  const int rs = RS_ ? RS_ : sp.rs, ps = PS_ ? PS_ : sp.ps;
  const int tile_len = az * ps;
  double *wsc = tile + tile_len;  // [32][WS]: z0*wz[P] | wy[P] | wx[P] of the lane's charge
  const int lane = threadIdx.x;
  int lm[NR], ll[NR], lconst[NR];
  bool lin[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const int item = lane + 32 * r;
    lin[r] = item < P2;
    lm[r] = item / P;
    ll[r] = item - lm[r] * P;
    lconst[r] = 8 * (lm[r] * rs + ll[r]);  // bytes
  }
  char *const tile_b = reinterpret_cast<char *>(tile);
  const int NX = g.nx, NY = g.ny, NZ = g.nz;

  for (;;) {
    int tile_id = 0;
    if (lane == 0) tile_id = atomicAdd(sp.counter, 1);
    tile_id = __shfl_sync(FULL, tile_id, 0);
    if (tile_id >= sp.ntiles) break;
    const int ix = tile_id % sp.ntx, iyz = tile_id / sp.ntx;
    const int iy = iyz % sp.nty, iz = iyz / sp.nty;
    const int t0 = iz * sp.tz, y0 = iy * sp.ty, x0 = ix * sp.tx;
    const int ez = min(sp.tz, g.zs_n - t0), ey = min(sp.ty, NY - y0), ex = min(sp.tx, NX - x0);
    const int ezv = ez + sp.halo_z, eyv = ey + sp.halo_y, exv = ex + sp.halo_x;  // valid positions incl. halo
    const int rb = sp.run_start[tile_id], re = sp.run_start[tile_id + 1];
    {
      double2 *t2 = reinterpret_cast<double2 *>(tile);
      for (int i = lane; i < (tile_len >> 1); i += 32) t2[i] = make_double2(0.0, 0.0);
      if ((tile_len & 1) && lane == 0) tile[tile_len - 1] = 0.0;
    }
    __syncwarp();

    for (int rbase = rb; rbase < re; rbase += 32) {
      // the charges of up to 32 cell runs, as one flat sequence of 32-charge chunks: lane i keeps run i's
      // range of sorted charges, an inclusive scan of the chunk counts maps a chunk number to its run
      int my_jb = 0, my_je = 0;
      if (rbase + lane < re) {
        const int2 run = sp.runs[rbase + lane];
        my_jb = cell_start[run.x];
        my_je = cell_start[run.y];
      }
      const int my_nch = (my_je - my_jb + 31) >> 5;
      int incl = my_nch;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += u;
      }
      const int nchunks = __shfl_sync(FULL, incl, 31);
      // chunk c -> (first charge, end of its run)
      auto locate = [&](int c, int &j0, int &je) {
        const unsigned mk = __ballot_sync(FULL, incl > c);
        const int i = __ffs(mk) - 1;
        const int jb_i = __shfl_sync(FULL, my_jb, i), excl_i = __shfl_sync(FULL, incl - my_nch, i);
        je = __shfl_sync(FULL, my_je, i);
        j0 = jb_i + ((c - excl_i) << 5);
      };
      PosQ nxt = {0.0, 0.0, 0.0, 0.0};
      if (nchunks > 0) {
        int j0, je;
        locate(0, j0, je);
        if (j0 + lane < je) nxt = atoms[j0 + lane];
      }
      for (int c = 0; c < nchunks; ++c) {
        const PosQ p = nxt;  // q == 0 also marks "no charge in this lane"
        nxt.q = 0.0;
        if (c + 1 < nchunks) {  // the next chunk's loads fly while this one is spread
          int j0, je;
          locate(c + 1, j0, je);
          if (j0 + lane < je) nxt = atoms[j0 + lane];
        }
        bool hit = false;
        unsigned my_lo = 0, my_hi = 0;
        int my_ob = 0;
        if (p.q != 0.0) {  // pppm_conp.cpp:161
          const double fx = (p.x - g.boxlo[0]) * g.delinv[0];
          const double fy = (p.y - g.boxlo[1]) * g.delinv[1];
          const double fz = (p.z - g.boxlo[2]) * g.delinv[2];
          if (!(fabs(fx) < OFFSET / 2 && fabs(fy) < OFFSET / 2 && fabs(fz) < OFFSET / 2)) {
            *range_flag = 1;  // "Out of range atoms - cannot compute PPPM", pppm_conp.cpp:167
          } else {
            const int nx = (int)(fx + g.shift) - OFFSET;  // :146-148
            const int ny = (int)(fy + g.shift) - OFFSET;
            const int nz = (int)(fz + g.shift) - OFFSET;
            const int uz = nz + g.nlower - g.zin_lo, uy = ny + g.nlower - y0, ux = nx + g.nlower - x0;
            // positions are wrapped into the box along periodic axes, so |u| < 2 n there; anything else is a
            // charge far outside a non-periodic box
            if (uz < -NZ || uz >= 2 * NZ || uy < -2 * NY || uy >= 2 * NY || ux < -2 * NX || ux >= 2 * NX) {
              *range_flag = 1;
            } else {
              const int zi0 = wrap2(uz, NZ);  // compact input plane of the first stencil plane
              if (g.nzi < NZ && zi0 + P > g.nzi) {
                *range_flag = 1;  // a plane outside what the box can reach: "Out of range atoms"
              } else {
                // signed start of the stencil relative to the tile, per axis (see header comment)
                int rz = wrap2(zi0 - g.zs_lo - t0, NZ), ry = wrap2(uy, NY), rx = wrap2(ux, NX);
                if (!sp.halo_z && rz >= ez) rz -= NZ;
                if (!sp.halo_y && ry >= ey) ry -= NY;
                if (!sp.halo_x && rx >= ex) rx -= NX;
                const int ka = max(0, -rz), kb = min(P, ezv - rz);
                hit = ka < kb && ry > -P && ry < eyv && rx > -P && rx < exv;
                if (hit) {
                  const double dx = nx + g.shiftone - fx;  // :199-201
                  const double dy = ny + g.shiftone - fy;
                  const double dz = nz + g.shiftone - fz;
                  const double z0 = g.delvolinv * p.q;  // :205
                  double *w = wsc + lane * WS;
#pragma unroll
                  for (int k = 0; k < P; ++k) {
                    w[k] = z0 * rho1d_c<P>(rc, k, dz);
                    w[P + k] = rho1d_c<P>(rc, k, dy);
                    w[2 * P + k] = rho1d_c<P>(rc, k, dx);
                  }
                  // which of the order x order footprint points (item m * order + l) and which of the `order`
                  // planes lie in the tile, and the byte offset of the stencil origin in the tile (may be
                  // negative: the valid points are not)
                  // (valid rows and columns are intervals: the footprint mask is the column mask repeated per row)
                  const int la = max(0, -rx), lb = min(P, exv - rx), ma = max(0, -ry), mb = min(P, eyv - ry);
                  const unsigned long long colmask = ((1u << lb) - 1u) & ~((1u << la) - 1u);
                  unsigned long long items = 0ull;
#pragma unroll
                  for (int m = 0; m < P; ++m)
                    if (m >= ma && m < mb) items |= colmask << (m * P);
                  my_lo = (unsigned)items;
                  my_hi = (unsigned)(items >> 32) | ((((1u << kb) - 1u) & ~((1u << ka) - 1u)) << 24);
                  my_ob = 8 * (rz * ps + ry * rs + rx);
                }
              }
            }
          }
        }
        unsigned mask = __ballot_sync(FULL, hit);
        __syncwarp();
        while (mask) {
          const int src = __ffs(mask) - 1;
          mask &= mask - 1;
          const unsigned lo = __shfl_sync(FULL, my_lo, src);
          const unsigned hi = __shfl_sync(FULL, my_hi, src);  // items 32.. in bits 0..16, plane mask in bits 24..30
          const int ob = __shfl_sync(FULL, my_ob, src);
          const unsigned pm = hi >> 24;
          const double *w = wsc + src * WS;
#pragma unroll
          for (int r2 = 0; r2 < NR; ++r2) {
            if (lin[r2] && (((r2 == 0 ? lo : hi) >> lane) & 1u)) {
              const double wyx = w[P + lm[r2]] * w[2 * P + ll[r2]];
              char *t = tile_b + (ob + lconst[r2]);
              double v[P];
#pragma unroll
              for (int n = 0; n < P; ++n)
                if (pm & (1u << n)) v[n] = *reinterpret_cast<double *>(t + n * 8 * ps);
#pragma unroll
              for (int n = 0; n < P; ++n)
                if (pm & (1u << n)) v[n] = fma(w[n], wyx, v[n]);
#pragma unroll
              for (int n = 0; n < P; ++n)
                if (pm & (1u << n)) *reinterpret_cast<double *>(t + n * 8 * ps) = v[n];
            }
          }
          __syncwarp();
        }
      }
    }
    __syncwarp();
    // fold the halo of a whole-axis tile back onto the periodic images (x, then y, then z: corners add up)
    if (sp.halo_x) {
      for (int i = lane; i < az * ay * sp.halo_x; i += 32) {
        const int h = i % sp.halo_x, pr = i / sp.halo_x;
        const int row = pr % ay, pl = pr / ay;
        double *q = tile + pl * ps + row * rs;
        q[h] += q[ex + h];
      }
      __syncwarp();
    }
    if (sp.halo_y) {
      for (int i = lane; i < az * sp.halo_y * ex; i += 32) {
        const int col = i % ex, pr = i / ex;
        const int h = pr % sp.halo_y, pl = pr / sp.halo_y;
        double *q = tile + pl * ps + col;
        q[h * rs] += q[(ey + h) * rs];
      }
      __syncwarp();
    }
    if (sp.halo_z) {
      for (int i = lane; i < sp.halo_z * ey * ex; i += 32) {
        const int col = i % ex, pr = i / ex;
        const int row = pr % ey, h = pr / ey;
        double *q = tile + row * rs + col;
        q[h * ps] += q[(ez + h) * ps];
      }
      __syncwarp();
    }
    // one plain store per mesh point of the tile
    for (int pl = 0; pl < ez; ++pl) {
      const double *q = tile + pl * ps + lane;
      double *dst = brick + ((size_t)(t0 + pl) * NY + y0) * NX + x0 + lane;
      for (int row = 0; row < ey; ++row, q += rs, dst += NX)
        for (int col = lane; col < ex; col += 32) dst[col - lane] = q[col - lane];
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------
// The same owner-computes spread with the accumulation on the FP64 tensor cores.
//
// On a z-plane the charge assignment of a group of charges is a small matrix product,
//     rho_n[y][x] += sum_a (zw_a[n] wy_a[y]) * wx_a[x],
// with the stencil weights zero outside the `order` points a charge touches: C (8 x 8) += A (8 x 4) . B (4 x 8)
// is one mma.sync.m8n8k4.f64 (SASS DMMA) for 4 charges, 8 mesh rows and 8 mesh columns.  A warp owns a tile
// of TZ planes x 8 rows x 32 columns as TZ x 4 accumulator fragments IN REGISTERS: no shared-memory
// read-modify-write and no dependent load/store chain per charge, only independent DMMAs.
//
// Two kernels.  stencil_prepass_kernel runs once per step over the cell-sorted charges: stencil origin
// (pppm_conp.cpp:146-148), the 3 x order weights (:199-203, z weights pre-multiplied by q/dV) and the
// "Out of range atoms" check (:167), so no charge can be dropped silently and nothing is recomputed per
// tile.  spread_mma_kernel then reads, per tile, the 16-byte origins of the charges in the tile's candidate
// cell runs (plan_pppm_spread_tiles), 32 at a time; the ones that overlap get a zero-padded copy of their
// weights in tile coordinates in a shared scratch row (zw over the TZ planes, wy over the 8 rows, wx over
// the 32 columns) and are taken in groups of 4 consecutive charges -- neighbours in the cell-sorted order,
// so a group touches few planes and column blocks, and only those (plane, block) fragments get a DMMA
// (warp-uniform masks).  Products and sums are exact FP64 FMAs; the summation order is the order of the
// sorted charges.  Used when no tile spans a whole periodic axis (spread_tile_kernel keeps those meshes).
// ---------------------------------------------------------------------------
constexpr int SM_TY = 8, SM_TX = 32, SM_NXB = SM_TX / 8;

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// origin[j] = (nx, ny, zi0, ok): mesh index of the stencil's first point along x and y (wrapped), compact
// input plane of its first plane, ok = 0 for q == 0 or an out-of-range charge; weights[3 P][wstride] (weight
// k of charge j at k * wstride + j: coalesced stores here, one 8-byte gather per weight in the tile kernel)
template <int P>
__global__ void __launch_bounds__(256)
stencil_prepass_kernel(PPPMGeom g, RhoCoeff rc, int m_bound, const int *__restrict__ count_ptr,
                       const PosQ *__restrict__ atoms, int4 *__restrict__ origin, double *__restrict__ weights,
                       size_t wstride, int *__restrict__ range_flag) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = count_ptr ? min(m_bound, *count_ptr) : m_bound;
  if (j >= m) return;
  const PosQ p = atoms[j];
  int4 o = make_int4(0, 0, 0, 0);
  if (p.q != 0.0) {  // pppm_conp.cpp:161
    const double fx = (p.x - g.boxlo[0]) * g.delinv[0];
    const double fy = (p.y - g.boxlo[1]) * g.delinv[1];
    const double fz = (p.z - g.boxlo[2]) * g.delinv[2];
    if (!(fabs(fx) < OFFSET / 2 && fabs(fy) < OFFSET / 2 && fabs(fz) < OFFSET / 2)) {
      *range_flag = 1;  // "Out of range atoms - cannot compute PPPM", pppm_conp.cpp:167
    } else {
      const int nx = (int)(fx + g.shift) - OFFSET;  // :146-148
      const int ny = (int)(fy + g.shift) - OFFSET;
      const int nz = (int)(fz + g.shift) - OFFSET;
      const int zi0 = wrapi(nz + g.nlower - g.zin_lo, g.nz);
      if (g.nzi < g.nz && zi0 + P > g.nzi) {
        *range_flag = 1;  // a plane outside what the box can reach: "Out of range atoms"
      } else {
        o = make_int4(wrapi(nx + g.nlower, g.nx), wrapi(ny + g.nlower, g.ny), zi0, 1);
        const double dx = nx + g.shiftone - fx;  // :199-201
        const double dy = ny + g.shiftone - fy;
        const double dz = nz + g.shiftone - fz;
        const double z0 = g.delvolinv * p.q;  // :205
        double *w = weights + j;
#pragma unroll
        for (int k = 0; k < P; ++k) {
          w[(size_t)k * wstride] = z0 * rho1d_c<P>(rc, k, dz);
          w[(size_t)(P + k) * wstride] = rho1d_c<P>(rc, k, dy);
          w[(size_t)(2 * P + k) * wstride] = rho1d_c<P>(rc, k, dx);
        }
      }
    }
  }
  origin[j] = o;
}

template <int P, int TZ>
__global__ void __launch_bounds__(32)
spread_mma_kernel(PPPMGeom g, SpreadPlan sp, const int4 *__restrict__ origin, const double *__restrict__ weights,
                  size_t wstride, const int *__restrict__ cell_start, double *__restrict__ brick) {
  constexpr int WS = (TZ + SM_TY + SM_TX) | 1;  // scratch row: zw[TZ] | wy[8] | wx[32], odd stride
  constexpr unsigned FULL = 0xffffffffu;
  __shared__ double wsc[32 * WS];
  __shared__ unsigned masks[32];  // per overlapping charge: plane mask | column-block mask << 16
  const int lane = threadIdx.x;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int fy = lane >> 2, fa = lane & 3;  // fragment row (A: mesh row, B: mesh column) / charge of the group
  const int NX = g.nx, NY = g.ny, NZ = g.nz;

  for (;;) {
    int tile_id = 0;
    if (lane == 0) tile_id = atomicAdd(sp.counter, 1);
    tile_id = __shfl_sync(FULL, tile_id, 0);
    if (tile_id >= sp.ntiles) break;
    const int ix = tile_id % sp.ntx, iyz = tile_id / sp.ntx;
    const int iy = iyz % sp.nty, iz = iyz / sp.nty;
    const int t0 = iz * TZ, y0 = iy * SM_TY, x0 = ix * SM_TX;
    const int ez = min(TZ, g.zs_n - t0), ey = min(SM_TY, NY - y0), ex = min(SM_TX, NX - x0);
    const int rb = sp.run_start[tile_id], re = sp.run_start[tile_id + 1];
    double acc[TZ][SM_NXB][2];
#pragma unroll
    for (int n = 0; n < TZ; ++n)
#pragma unroll
      for (int b = 0; b < SM_NXB; ++b) acc[n][b][0] = acc[n][b][1] = 0.0;

    for (int rbase = rb; rbase < re; rbase += 32) {
      // flat sequence of 32-charge chunks over up to 32 cell runs (see spread_tile_kernel)
      int my_jb = 0, my_je = 0;
      if (rbase + lane < re) {
        const int2 run = sp.runs[rbase + lane];
        my_jb = cell_start[run.x];
        my_je = cell_start[run.y];
      }
      const int my_nch = (my_je - my_jb + 31) >> 5;
      int incl = my_nch;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += u;
      }
      const int nchunks = __shfl_sync(FULL, incl, 31);
      auto locate = [&](int c) {  // index of this lane's charge in chunk c, or -1
        const unsigned mk = __ballot_sync(FULL, incl > c);
        const int i = __ffs(mk) - 1;
        const int jb_i = __shfl_sync(FULL, my_jb, i), excl_i = __shfl_sync(FULL, incl - my_nch, i);
        const int je = __shfl_sync(FULL, my_je, i);
        const int j = jb_i + ((c - excl_i) << 5) + lane;
        return j < je ? j : -1;
      };
      int jn = nchunks > 0 ? locate(0) : -1;
      int4 on = make_int4(0, 0, 0, 0);
      if (jn >= 0) on = origin[jn];
      for (int c = 0; c < nchunks; ++c) {
        const int j = jn;
        const int4 o = on;
        jn = -1;
        on.w = 0;
        if (c + 1 < nchunks) {  // the next chunk's origins fly while this one is spread
          jn = locate(c + 1);
          if (jn >= 0) on = origin[jn];
        }
        // signed start of the stencil relative to the tile (tiles are shorter than the mesh by at least the
        // stencil, so the touched points form one interval per axis)
        int rz = o.z - g.zs_lo - t0, ry = o.y - y0, rx = o.x - x0;
        rz += rz < 0 ? NZ : 0;   // o.z in [0, NZ), zs_lo + t0 in [0, NZ)
        ry += ry < 0 ? NY : 0;
        rx += rx < 0 ? NX : 0;
        if (rz >= ez) rz -= NZ;
        if (ry >= ey) ry -= NY;
        if (rx >= ex) rx -= NX;
        const bool hit = o.w != 0 && rz > -P && rz < ez && ry > -P && ry < ey && rx > -P && rx < ex;
        const unsigned hits = __ballot_sync(FULL, hit);
        const int nh = __popc(hits);
        // clear the scratch rows that will be used, then drop the weights at their tile coordinates
        for (int t = lane; t < nh * WS; t += 32) wsc[t] = 0.0;
        __syncwarp();
        if (hit) {
          const int r = __popc(hits & lt_mask);
          const double *w = weights + j;
          double *row = wsc + r * WS;
          unsigned planes = 0u, blocks = 0u;
#pragma unroll
          for (int k = 0; k < P; ++k) {
            const double wz = w[(size_t)k * wstride], wy = w[(size_t)(P + k) * wstride],
                         wx = w[(size_t)(2 * P + k) * wstride];
            if ((unsigned)(rz + k) < (unsigned)ez) { row[rz + k] = wz; planes |= 1u << (rz + k); }
            if ((unsigned)(ry + k) < (unsigned)ey) row[TZ + ry + k] = wy;
            if ((unsigned)(rx + k) < (unsigned)ex) { row[TZ + SM_TY + rx + k] = wx; blocks |= 1u << ((rx + k) >> 3); }
          }
          masks[r] = planes | (blocks << 16);
        }
        __syncwarp();
        for (int g0 = 0; g0 < nh; g0 += 4) {
          const int row = g0 + fa;
          const bool have = row < nh;
          unsigned pbg = have ? masks[row] : 0u;
          pbg |= __shfl_xor_sync(FULL, pbg, 1);
          pbg |= __shfl_xor_sync(FULL, pbg, 2);  // union over the 4 charges: identical in every lane
          const double *w = wsc + row * WS;
          const double ay = have ? w[TZ + fy] : 0.0;  // A fragment: row fy of the tile, charge fa
          double bx[SM_NXB];
#pragma unroll
          for (int b = 0; b < SM_NXB; ++b)  // B fragment: column 8 b + fy of the tile, charge fa
            bx[b] = (have && (pbg & (0x10000u << b))) ? w[TZ + SM_TY + 8 * b + fy] : 0.0;
#pragma unroll
          for (int n = 0; n < TZ; ++n) {
            if (pbg & (1u << n)) {
              const double a = ay * (have ? w[n] : 0.0);
#pragma unroll
              for (int b = 0; b < SM_NXB; ++b)
                if (pbg & (0x10000u << b)) dmma884(acc[n][b][0], acc[n][b][1], a, bx[b]);
            }
          }
        }
        __syncwarp();
      }
    }
    // C fragment: row = lane / 4, columns 2 (lane % 4) + {0, 1} of the 8 x 8 block: one plain store per point
#pragma unroll
    for (int n = 0; n < TZ; ++n) {
      if (n < ez && fy < ey) {
        double *dst = brick + ((size_t)(t0 + n) * NY + (y0 + fy)) * NX + x0;
#pragma unroll
        for (int b = 0; b < SM_NXB; ++b) {
          const int col = 8 * b + 2 * fa;
          if (col < ex) dst[col] = acc[n][b][0];
          if (col + 1 < ex) dst[col + 1] = acc[n][b][1];
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// z-sweep form of the tensor-core spread (the default for large systems).
//
// What the tile kernels above pay for is finding and clipping the charges of a tile: the sort cells are
// physical (rc/2 wide, shared with the pair kernels), so every tile re-reads several times the charges it
// uses, and a 5-point stencil straddles a 4-8 point tile in most directions.  Here the charges get a second,
// mesh-aligned counting sort per step -- key = (column, origin plane), a column being a footprint of 8 mesh
// rows x 32 mesh columns, plane fastest -- and a warp sweeps one column upwards through z:
//   * all charges whose stencil starts on plane p touch exactly the planes p .. p+order-1, so the warp keeps
//     a window of `order` planes x 8 rows x 32 columns as accumulator fragments in registers; after the
//     charges of origin plane p, plane p is complete: it is stored (once, plain stores) and its registers
//     become plane p+order (the window rotates by renaming inside an unrolled loop, no register moves);
//   * candidates of plane p are the four bins (this column and its -y, -x, -xy neighbours, whose stencils
//     can reach over the border) of that plane: no scan of unrelated charges, no clipping in z at all;
//   * the accumulation is the same block-sparse DMMA as in spread_mma_kernel (4 charges x 8 rows x 8 columns
//     per mma.sync.m8n8k4.f64), always over all `order` planes.
// Work items are (column, z-segment); a segment first runs the order-1 origin planes below it (warm-up).
// mesh_bin_kernel / mesh_scatter_kernel do the sort: origin + bin, exclusive scan (cell_scan_kernel), then the
// weights are computed once and written straight to their sorted place.
// ---------------------------------------------------------------------------
constexpr int SW_FY = 8, SW_FX = 32, SW_NXB = SW_FX / 8;

struct SweepGeom {
  int ncolx, ncoly;   // columns along x / y
  int pz_lo, npz;     // origin planes binned: compact planes pz_lo .. pz_lo + npz - 1 (mod nz when periodic)
  int wrap_z;         // the slab is the whole periodic mesh: origin planes wrap
};

// origin plane -> bin plane index, or -1
__device__ __forceinline__ int sweep_plane(const SweepGeom &sg, int zi0, int nz) {
  int t = zi0 - sg.pz_lo;
  if (sg.wrap_z) { t += t < 0 ? nz : 0; t -= t >= nz ? nz : 0; }
  return ((unsigned)t < (unsigned)sg.npz) ? t : -1;
}

template <int P>
__device__ __forceinline__ bool stencil_origin(const PPPMGeom &g, const PosQ &p, int &ox, int &oy, int &zi0, double &dx,
                                               double &dy, double &dz, int *range_flag) {
  if (p.q == 0.0) return false;  // pppm_conp.cpp:161
  const double fx = (p.x - g.boxlo[0]) * g.delinv[0];
  const double fy = (p.y - g.boxlo[1]) * g.delinv[1];
  const double fz = (p.z - g.boxlo[2]) * g.delinv[2];
  if (!(fabs(fx) < OFFSET / 2 && fabs(fy) < OFFSET / 2 && fabs(fz) < OFFSET / 2)) {
    *range_flag = 1;  // "Out of range atoms - cannot compute PPPM", pppm_conp.cpp:167
    return false;
  }
  const int nx = (int)(fx + g.shift) - OFFSET;  // :146-148
  const int ny = (int)(fy + g.shift) - OFFSET;
  const int nz = (int)(fz + g.shift) - OFFSET;
  zi0 = wrapi(nz + g.nlower - g.zin_lo, g.nz);
  if (g.nzi < g.nz && zi0 + P > g.nzi) {
    *range_flag = 1;  // a plane outside what the box can reach: "Out of range atoms"
    return false;
  }
  ox = wrapi(nx + g.nlower, g.nx);
  oy = wrapi(ny + g.nlower, g.ny);
  dx = nx + g.shiftone - fx;  // :199-201
  dy = ny + g.shiftone - fy;
  dz = nz + g.shiftone - fz;
  return true;
}

template <int P>
__global__ void __launch_bounds__(256)
mesh_bin_kernel(PPPMGeom g, SweepGeom sg, int m, const int *__restrict__ valid, const PosQ *__restrict__ atoms,
                int *__restrict__ bin_of, int *__restrict__ slot, int *__restrict__ bin_count,
                int *__restrict__ range_flag) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  int ox, oy, zi0;
  double dx, dy, dz;
  int bin = -1;
  // several GPUs: `valid` (the cell index bin_positions assigned, < 0 = empty or irrelevant inbox slot)
  if ((!valid || valid[j] >= 0) && stencil_origin<P>(g, atoms[j], ox, oy, zi0, dx, dy, dz, range_flag)) {
    const int t = sweep_plane(sg, zi0, g.nz);
    if (t >= 0) bin = ((oy / SW_FY) * sg.ncolx + ox / SW_FX) * sg.npz + t;
  }
  bin_of[j] = bin;
  if (bin >= 0) slot[j] = atomicAdd(bin_count + bin, 1);
}

// record of charge j at its sorted position d: SW_ND(P) doubles = one or two cache lines,
//   [0] = (ox, oy) as two ints | [1 .. P] z weights x q/dV | [1+P .. 2P] y weights | [1+2P .. 3P] x weights
template <int P>
struct SweepRec {
  static constexpr int ND = (3 * P + 2) & ~1;  // 1 + 3 P, rounded up to a multiple of 16 bytes
};

template <int P>
__global__ void __launch_bounds__(256)
mesh_scatter_kernel(PPPMGeom g, RhoCoeff rc, int m, const PosQ *__restrict__ atoms, const int *__restrict__ bin_of,
                    const int *__restrict__ slot, const int *__restrict__ bin_start, double *__restrict__ records) {
  constexpr int ND = SweepRec<P>::ND;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const int bin = bin_of[j];
  if (bin < 0) return;
  const PosQ p = atoms[j];
  int ox, oy, zi0, dummy = 0;
  double dx, dy, dz;
  stencil_origin<P>(g, p, ox, oy, zi0, dx, dy, dz, &dummy);
  const size_t d = (size_t)bin_start[bin] + slot[j];
  const double z0 = g.delvolinv * p.q;  // :205
  double r[ND];
  r[0] = __longlong_as_double(((long long)(unsigned)oy << 32) | (unsigned)ox);
  r[ND - 1] = 0.0;
#pragma unroll
  for (int k = 0; k < P; ++k) {
    r[1 + k] = z0 * rho1d_c<P>(rc, k, dz);
    r[1 + P + k] = rho1d_c<P>(rc, k, dy);
    r[1 + 2 * P + k] = rho1d_c<P>(rc, k, dx);
  }
  double2 *dst = reinterpret_cast<double2 *>(records + d * ND);  // one 128-byte line per charge (order 5)
#pragma unroll
  for (int k = 0; k < ND / 2; ++k) dst[k] = make_double2(r[2 * k], r[2 * k + 1]);
}

constexpr int SW_MAXS = 64;  // plane-steps per work item (segment + warm-up), see plan_pppm_sweep

template <int P>
__global__ void __launch_bounds__(32, 12)
spread_sweep_kernel(PPPMGeom g, SweepGeom sg, int nitems, const int4 *__restrict__ items, int *__restrict__ counter,
                    const int *__restrict__ bin_start, const double *__restrict__ records,
                    double *__restrict__ brick) {
  constexpr int ND = SweepRec<P>::ND;
  constexpr int STR = ND + 2;  // staging row stride: 16-byte aligned, rows spread over the banks
  constexpr unsigned FULL = 0xffffffffu;
  __shared__ __align__(16) double stage[2][32 * STR];  // the records of one chunk of 32 candidates, double-buffered
  __shared__ int order[32];                            // r-th overlapping candidate of the chunk -> its lane
  __shared__ int rngb[4][SW_MAXS], rnge[4][SW_MAXS];   // per plane-step: record ranges of the four bins
  __shared__ int stot[SW_MAXS];                        // ... and their total length
  const int lane = threadIdx.x;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int fy = lane >> 2, fa = lane & 3;
  const int NX = g.nx, NY = g.ny, NZ = g.nz;

  for (;;) {
    int it = 0;
    if (lane == 0) it = atomicAdd(counter, 1);
    it = __shfl_sync(FULL, it, 0);
    if (it >= nitems) break;
    const int4 item = items[it];  // column, first / end plane of the segment (slab-local t), unused
    const int col = item.x, t_lo = item.y, t_hi = item.z;
    const int cy = col / sg.ncolx, cx = col - cy * sg.ncolx;
    const int y0 = cy * SW_FY, x0 = cx * SW_FX;
    const int ey = min(SW_FY, NY - y0), ex = min(SW_FX, NX - x0);
    // the four columns whose charges can reach this one (stencils run towards +y / +x, periodic mesh)
    const int cym = cy == 0 ? sg.ncoly - 1 : cy - 1, cxm = cx == 0 ? sg.ncolx - 1 : cx - 1;
    const int ccol[4] = {col, cy * sg.ncolx + cxm, cym * sg.ncolx + cx, cym * sg.ncolx + cxm};  // >= 2 columns per axis
    const int ts = t_lo - (P - 1);  // origin plane of step 0 (warm-up: the planes below the segment)
    const int ns = t_hi - ts;       // plane-steps; step s handles origin plane ts + s and completes that plane
    __syncwarp();
    for (int u = lane; u < 4 * ns; u += 32) {
      const int c = u / ns, sidx = u - c * ns;
      const int cc = c == 0 ? ccol[0] : c == 1 ? ccol[1] : c == 2 ? ccol[2] : ccol[3];
      int zi = g.zs_lo + ts + sidx;  // compact plane of the step's origin plane
      if (sg.wrap_z) { zi += zi < 0 ? NZ : 0; zi -= zi >= NZ ? NZ : 0; }
      const int bp = zi >= 0 ? sweep_plane(sg, zi, NZ) : -1;
      int b0 = 0, b1 = 0;
      if (bp >= 0) {
        const int b = cc * sg.npz + bp;
        b0 = bin_start[b];
        b1 = bin_start[b + 1];
      }
      rngb[c][sidx] = b0;
      rnge[c][sidx] = b1;
    }
    __syncwarp();
    for (int u = lane; u < ns; u += 32)
      stot[u] = (rnge[0][u] - rngb[0][u]) + (rnge[1][u] - rngb[1][u]) + (rnge[2][u] - rngb[2][u]) +
                (rnge[3][u] - rngb[3][u]);
    __syncwarp();

    double acc[P][SW_NXB][2];  // window: acc[n] = plane (current origin plane + n)
#pragma unroll
    for (int n = 0; n < P; ++n)
#pragma unroll
      for (int b = 0; b < SW_NXB; ++b) acc[n][b][0] = acc[n][b][1] = 0.0;

    // candidates of step sidx: the four ranges back to back; total count
    auto step_total = [&](int sidx) { return stot[sidx]; };
    // first chunk position at or after (sidx, c0)
    auto settle = [&](int &sidx, int &c0) {
      while (sidx < ns && c0 >= step_total(sidx)) { ++sidx; c0 = 0; }
    };
    // start the copy of chunk (sidx, c0) into stage[buf]: lane l takes candidate c0 + l
    auto issue = [&](int sidx, int c0, int buf) {
      int q = c0 + lane;
      long long j = -1;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int n = rnge[c][sidx] - rngb[c][sidx];
        if (j < 0 && q >= 0 && q < n) j = rngb[c][sidx] + q;
        q -= n;
      }
      if (j >= 0) {
        const double *src = records + (size_t)j * ND;
        double *dst = &stage[buf][lane * STR];
#pragma unroll
        for (int k = 0; k < ND / 2; ++k) cp_async<16>(dst + 2 * k, src + 2 * k);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // plane (origin plane of step `sidx`) is complete: one plain store per point, rotate the window
    auto finish_step = [&](int sidx) {
      const int t = ts + sidx;
      if (t >= t_lo && fy < ey) {
        double *dst = brick + ((size_t)t * NY + (y0 + fy)) * NX + x0;
#pragma unroll
        for (int b = 0; b < SW_NXB; ++b) {
          const int cc = 8 * b + 2 * fa;
          if (cc < ex) dst[cc] = acc[0][b][0];
          if (cc + 1 < ex) dst[cc + 1] = acc[0][b][1];
        }
      }
#pragma unroll
      for (int n = 0; n + 1 < P; ++n)
#pragma unroll
        for (int b = 0; b < SW_NXB; ++b) { acc[n][b][0] = acc[n + 1][b][0]; acc[n][b][1] = acc[n + 1][b][1]; }
#pragma unroll
      for (int b = 0; b < SW_NXB; ++b) acc[P - 1][b][0] = acc[P - 1][b][1] = 0.0;
    };

    int s_cur = 0, c_cur = 0, buf = 0, s_done = 0;  // s_done: first step whose plane is not finished yet
    settle(s_cur, c_cur);
    if (s_cur < ns) issue(s_cur, c_cur, buf);
    while (s_cur < ns) {
      int s_nxt = s_cur, c_nxt = c_cur + 32;
      settle(s_nxt, c_nxt);
      if (s_nxt < ns) {  // the next chunk's records fly while this one is spread
        issue(s_nxt, c_nxt, buf ^ 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncwarp();
      while (s_done < s_cur) finish_step(s_done++);  // steps without candidates in between
      const int nvalid = min(32, step_total(s_cur) - c_cur);
      const double *st = stage[buf];
      bool hit = false;
      if (lane < nvalid) {
        const long long oo = __double_as_longlong(st[lane * STR]);
        int ry = (int)(oo >> 32) - y0, rx = (int)(oo & 0xffffffffll) - x0;
        // (o - 0) mod n, then the last order-1 values are stencils that wrap in from below
        ry += ry < 0 ? NY : 0; ry -= ry > NY - P ? NY : 0;
        rx += rx < 0 ? NX : 0; rx -= rx > NX - P ? NX : 0;
        hit = ry > -P && ry < ey && rx > -P && rx < ex;
      }
      const unsigned hits = __ballot_sync(FULL, hit);
      const int nh = __popc(hits);
      if (hit) order[__popc(hits & lt_mask)] = lane;
      __syncwarp();
      for (int g0 = 0; g0 < nh; g0 += 4) {
        const int row = g0 + fa;
        const bool have = row < nh;
        const double *rec = st + (have ? order[row] : 0) * STR;
        const long long oo = __double_as_longlong(rec[0]);
        int ry = (int)(oo >> 32) - y0, rx = (int)(oo & 0xffffffffll) - x0;
        ry += ry < 0 ? NY : 0; ry -= ry > NY - P ? NY : 0;
        rx += rx < 0 ? NX : 0; rx -= rx > NX - P ? NX : 0;
        // column blocks this charge reaches
        const int c_lo = max(rx, 0), c_hi = min(rx + P - 1, SW_FX - 1);
        unsigned bg = have ? (((1u << ((c_hi >> 3) + 1)) - 1u) & ~((1u << (c_lo >> 3)) - 1u)) : 0u;
        bg |= __shfl_xor_sync(FULL, bg, 1);
        bg |= __shfl_xor_sync(FULL, bg, 2);  // union over the 4 charges: identical in every lane
        const int my = fy - ry;  // A fragment: row fy of the column, charge fa
        const double ay = (have && (unsigned)my < (unsigned)P) ? rec[1 + P + my] : 0.0;
        // B fragment of block b: column 8 b + fy, charge fa.  The usual block patterns of four neighbouring
        // charges (one block, or two adjacent ones) get straight-line code: no predicate per DMMA.
#define CONP_SW_BLOCKS(MASK)                                                                           \
  {                                                                                                    \
    double bx[SW_NXB];                                                                                 \
    _Pragma("unroll") for (int b = 0; b < SW_NXB; ++b) {                                               \
      bx[b] = 0.0;                                                                                     \
      if ((MASK) & (1u << b)) {                                                                        \
        const int lx = 8 * b + fy - rx;                                                                \
        bx[b] = (have && (unsigned)lx < (unsigned)P) ? rec[1 + 2 * P + lx] : 0.0;                      \
      }                                                                                                \
    }                                                                                                  \
    _Pragma("unroll") for (int n = 0; n < P; ++n) {                                                    \
      const double a = ay * (have ? rec[1 + n] : 0.0);                                                 \
      _Pragma("unroll") for (int b = 0; b < SW_NXB; ++b)                                               \
        if ((MASK) & (1u << b)) dmma884(acc[n][b][0], acc[n][b][1], a, bx[b]);                         \
    }                                                                                                  \
  }
        switch (bg) {
          case 1u: CONP_SW_BLOCKS(1u) break;
          case 2u: CONP_SW_BLOCKS(2u) break;
          case 4u: CONP_SW_BLOCKS(4u) break;
          case 8u: CONP_SW_BLOCKS(8u) break;
          case 3u: CONP_SW_BLOCKS(3u) break;
          case 6u: CONP_SW_BLOCKS(6u) break;
          case 12u: CONP_SW_BLOCKS(12u) break;
          default: CONP_SW_BLOCKS(bg) break;
        }
#undef CONP_SW_BLOCKS
      }
      __syncwarp();
      s_cur = s_nxt; c_cur = c_nxt; buf ^= 1;
    }
    while (s_done < ns) finish_step(s_done++);
  }
}

// z-convolution with the tabulated kernel, one launch for all (kx,ky) columns.
//
// For k_xy != 0 the kernel decays like the Ewald Gaussian / exp(-|k_xy| |dz|): krad[col] bounds the
// circular |d| beyond which |K| is below ~1e-18 of the column maximum (measured on the table at
// setup, see ctx.cu); the dropped tail is below the rounding level of the reference's own FFTs.
// The output planes sit in thin groups at the electrodes, so a column only needs the input planes
// within its radius of an output plane.
//
//  * "Narrow" blocks own ZC_COLS = 8 consecutive columns (128 B of every plane of the plane-major
//    spectra).  The host has worked out which planes of this rank's slab the group's window touches
//    (ZconvGroup: a few intervals, compacted row numbers); the block stages those planes of rho^ and
//    the 2 rblock + 1 table entries K[-rblock .. rblock] with cp.async (everything in flight at
//    once), then each thread owns one (column, output plane) pair and runs its window out of
//    shared memory, ascending input plane.  ~15 KB per block: several blocks per SM, so the loads of
//    one hide behind the sums of another.
//  * "Wide" blocks take one of the few small-|k_xy| columns whose kernel reaches across the whole
//    mesh: each warp owns output planes, its lanes stride over all slab planes straight from L2 and
//    the partial sums are combined with a fixed shuffle tree.
constexpr int ZC_THREADS = 256;
constexpr int ZC_COLS = 8;

template <bool REALK>
__global__ void __launch_bounds__(ZC_THREADS)
zconv_kernel(int ncol, int nz, int nzl, int zs_lo, int nzo, const int *__restrict__ aout,
             const int *__restrict__ krad, const ZconvGroup *__restrict__ groups, int n_narrow,
             const int *__restrict__ wide_cols, int rcap, int npcap, const double2 *__restrict__ rhat,
             const double *__restrict__ Kr, const double2 *__restrict__ Kc, double2 *__restrict__ uhat,
             PeerSync ps) {
  // rhat: spectra of this rank's nzl input planes (compact planes zs_lo .. zs_lo+nzl-1); on several
  // GPUs uhat is this rank's partial sum and is all-reduced afterwards.  aout[zo] = position of output
  // plane zo on the ring in compact input-plane coordinates.
  if ((int)blockIdx.x >= n_narrow) {
    // ---------------- wide column ----------------
    const int c = wide_cols[blockIdx.x - n_narrow];
    const int R = krad[c];
    const bool all = 2 * R + 1 >= nz;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int zo = warp; zo < nzo; zo += ZC_THREADS / 32) {
      const int a = aout[zo];
      double ar = 0.0, ai = 0.0;
      for (int t = lane; t < nzl; t += 32) {
        int d = a - (zs_lo + t);
        d += (d < 0) ? nz : 0;  // table index (a - zi) mod nz
        if (all || min(d, nz - d) <= R) {
          const double2 r = rhat[(size_t)t * ncol + c];
          if (REALK) {
            const double k = Kr[(size_t)c * nz + d];
            ar = fma(k, r.x, ar);
            ai = fma(k, r.y, ai);
          } else {
            const double2 k = Kc[(size_t)c * nz + d];
            ar = fma(k.x, r.x, ar); ar = fma(-k.y, r.y, ar);
            ai = fma(k.x, r.y, ai); ai = fma(k.y, r.x, ai);
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        ar += __shfl_xor_sync(0xffffffffu, ar, o);
        ai += __shfl_xor_sync(0xffffffffu, ai, o);
      }
      if (lane == 0) uhat[(size_t)zo * ncol + c] = make_double2(ar, ai);
    }
    peer_block_signal(ps);  // several GPUs: this rank's partial spectra are complete when every block is through
    return;
  }
  // ---------------- narrow group ----------------
  extern __shared__ __align__(16) unsigned char zc_smem[];
  double2 *rh = reinterpret_cast<double2 *>(zc_smem);                      // [npcap][ZC_COLS]
  const int kst = 2 * rcap + 1;
  double *ksr = reinterpret_cast<double *>(rh + (size_t)npcap * ZC_COLS);  // REALK: [ZC_COLS][kst]
  double2 *ksc = reinterpret_cast<double2 *>(ksr);                          // else:  [ZC_COLS][kst]
  __shared__ ZconvGroup gd;
  if (threadIdx.x < sizeof(ZconvGroup) / sizeof(int))
    reinterpret_cast<int *>(&gd)[threadIdx.x] = reinterpret_cast<const int *>(groups + blockIdx.x)[threadIdx.x];
  __syncthreads();
  const int c0 = gd.c0, rblock = gd.rblock;
  // compact row of slab plane t (t must lie in one of the group's intervals)
  auto rowof = [&](int t) {
    int row = 0;
#pragma unroll
    for (int i = 0; i < ZconvGroup::MAXI; ++i)
      if (i < gd.nint && t >= gd.lo[i] && t < gd.hi[i]) row = gd.base[i] + t - gd.lo[i];
    return row;
  };
  for (int i = 0; i < gd.nint; ++i) {
    const int n = (gd.hi[i] - gd.lo[i]) * ZC_COLS;
    for (int idx = threadIdx.x; idx < n; idx += ZC_THREADS) {
      const int t = gd.lo[i] + idx / ZC_COLS, cc = idx % ZC_COLS;
      cp_async<16>(&rh[(gd.base[i] + t - gd.lo[i]) * ZC_COLS + cc], &rhat[(size_t)t * ncol + min(c0 + cc, ncol - 1)]);
    }
  }
  for (int idx = threadIdx.x; idx < ZC_COLS * (2 * rblock + 1); idx += ZC_THREADS) {
    const int cc = idx / (2 * rblock + 1), j = idx - cc * (2 * rblock + 1);
    const int c = min(c0 + cc, ncol - 1);
    int d = j - rblock;  // signed ring distance
    d += (d < 0) ? nz : 0;
    if (REALK)
      cp_async<8>(&ksr[cc * kst + j], Kr + (size_t)c * nz + d);
    else
      cp_async<16>(&ksc[cc * kst + j], Kc + (size_t)c * nz + d);
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  for (int item = threadIdx.x; item < ZC_COLS * nzo; item += ZC_THREADS) {
    const int zo = item / ZC_COLS, cc = item - zo * ZC_COLS;
    const int c = c0 + cc;
    if (c >= ncol) continue;
    const int R = krad[c];
    const int a = aout[zo];
    double ar = 0.0, ai = 0.0;
    // input planes z0..z1 clipped to the slab; signed distance of plane zi is a - zi + shift
    auto run = [&](int z0, int z1, int shift) {
      z0 = max(z0, zs_lo);
      z1 = min(z1, zs_lo + nzl - 1);
      if (z0 > z1) return;
      int row = rowof(z0 - zs_lo);  // the planes of one run are consecutive compact rows
      int j = a - z0 + shift + rblock;
      for (int zi = z0; zi <= z1; ++zi, ++row, --j) {
        const double2 r = rh[row * ZC_COLS + cc];
        if (REALK) {
          const double k = ksr[cc * kst + j];
          ar = fma(k, r.x, ar);
          ai = fma(k, r.y, ai);
        } else {
          const double2 k = ksc[cc * kst + j];
          ar = fma(k.x, r.x, ar); ar = fma(-k.y, r.y, ar);
          ai = fma(k.x, r.y, ai); ai = fma(k.y, r.x, ai);
        }
      }
    };
    run(max(a - R, 0), min(a + R, nz - 1), 0);   // main window
    if (a - R < 0) run(a - R + nz, nz - 1, nz);   // wrapped from below
    if (a + R >= nz) run(0, a + R - nz, -nz);     // wrapped from above
    uhat[(size_t)zo * ncol + c] = make_double2(ar, ai);
  }
  peer_block_signal(ps);
}

// compact <-> full brick copies for the on-demand outputs
__global__ void __launch_bounds__(256)
expand_planes_kernel(size_t plane, int nplanes, int nz, int lo, const int *__restrict__ list,
                     const double *__restrict__ compact, double *__restrict__ full) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= plane * (size_t)nplanes) return;
  const int p = (int)(i / plane);
  const size_t r = i - (size_t)p * plane;
  int mz = list ? list[p] : (lo + p) % nz;
  if (mz < 0) mz += nz;
  full[(size_t)mz * plane + r] = compact[i];
}

// sub-brick [lo, hi] of the full periodic mesh out of the compact bricks: electrolyte planes zi <->
// (zin_lo + zi) mod nz, electrode planes zmap[mz]; a plane that is not stored is zero
__global__ void __launch_bounds__(256)
region_gather_kernel(PPPMGeom g, int which, int lx, int ly, int lz, int ex, int ey, int ez,
                     const double *__restrict__ elyte, const double *__restrict__ ele, double *__restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)ex * ey * ez) return;
  const int x = lx + (int)(i % ex);
  const size_t r = i / ex;
  const int y = ly + (int)(r % ey), z = lz + (int)(r / ey);
  const size_t inplane = (size_t)y * g.nx + x;
  double v = 0.0;
  if (which != 1) {
    const int zi = wrapi(z - g.zin_lo, g.nz);
    if (zi < g.nzi) v += elyte[(size_t)zi * g.ny * g.nx + inplane];
  }
  if (which != 0) {
    const int zo = g.zmap[z];
    if (zo >= 0) v += ele[(size_t)zo * g.ny * g.nx + inplane];
  }
  out[i] = v;
}

__global__ void __launch_bounds__(256)
green_mul_kernel(size_t n, cufftDoubleComplex *__restrict__ work, const double *__restrict__ ghalf) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double gq = ghalf[i];
  cufftDoubleComplex w = work[i];
  w.x *= gq;
  w.y *= gq;
  work[i] = w;
}

__global__ void __launch_bounds__(128)
ele_stencil_kernel(PPPMGeom g, const double *__restrict__ rho_coeff, int n_ele, const double *__restrict__ ex,
                   const double *__restrict__ ey, const double *__restrict__ ez, int *__restrict__ part2grid,
                   double *__restrict__ weights) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_ele) return;
  const double xs[3] = {ex[i], ey[i], ez[i]};
  for (int ic = 0; ic < 3; ++ic) {
    const double xlo = xs[ic] - g.boxlo[ic];
    const int nn = (int)(xlo * g.delinv[ic] + g.shift) - OFFSET;  // :333
    part2grid[3 * i + ic] = nn;
    const double d = nn + g.shiftone - xlo * g.delinv[ic];  // :335
    for (int l = 0; l < g.order; ++l) weights[((size_t)i * 3 + ic) * g.order + l] = rho1d(rho_coeff, g.order, l, d);
  }
}

// wrapped stencil indices of the (static) electrode atoms: widx[i][0][l] = mx, [1][m] = my,
// [2][n] = compact output plane; removes all integer modulo work from gather / re-spread
__global__ void __launch_bounds__(128)
ele_index_kernel(PPPMGeom g, int n_ele, const int *__restrict__ part2grid, int *__restrict__ widx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_ele) return;
  const int dims[3] = {g.nx, g.ny, g.nz};
  for (int ic = 0; ic < 3; ++ic)
    for (int l = 0; l < g.order; ++l) {
      int m = wrapi(l + g.nlower + part2grid[3 * i + ic], dims[ic]);
      if (ic == 2) m = g.zmap[m];
      widx[((size_t)i * 3 + ic) * g.order + l] = m;
    }
}

// flat per-point stencil table of the static electrode atoms: offset into the compact potential
// brick and the weight product w_z w_y w_x (same multiplication order as pppm_conp.cpp:287-293)
__global__ void __launch_bounds__(128)
ele_point_table_kernel(PPPMGeom g, int n_ele, const int *__restrict__ widx, const double *__restrict__ weights,
                       int *__restrict__ poff, double *__restrict__ pw) {
  const int order = g.order, npts = order * order * order;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int i = (int)(gid / npts);
  if (i >= n_ele) return;
  const int t = (int)(gid - (long long)i * npts);
  const int n = t / (order * order);
  const int r = t - n * order * order;
  const int m = r / order, l = r - m * order;
  const double *w = weights + (size_t)i * 3 * order;
  const int *wi = widx + (size_t)i * 3 * order;
  poff[gid] = (wi[2 * order + n] * g.ny + wi[order + m]) * g.nx + wi[l];
  pw[gid] = w[2 * order + n] * w[order + m] * w[l];
}

// one warp per electrode row: b_k = -sum w u (pppm_conp.cpp:285-298), slab
// term (:301-313), then b = b_k + b_real.  On several GPUs (peer-to-peer path) the kernel is its own
// b_comm (fix_conp.cpp:641-648): the row's b goes straight into every peer's copy of the vector and
// the last block raises the flags the matvec kernel polls.
constexpr int GB_ROWS = 32;  // rows per block on several GPUs: one 256-byte line of b per peer

__global__ void __launch_bounds__(256)
gather_b_kernel(PPPMGeom g, int row_begin, int row_end, const int *__restrict__ poff,
                const double *__restrict__ pw, const double *__restrict__ u_brick,
                const double *__restrict__ ez, const double *__restrict__ qz_sum, double slab_pref,
                const double *__restrict__ b_real, double *__restrict__ b_kspace, double *__restrict__ b,
                PeerSync ps, size_t off_b, int rows_per_block) {
  __shared__ double sb[GB_ROWS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int base = row_begin + blockIdx.x * rows_per_block;
  const int npts = g.order * g.order * g.order;
  for (int k = warp; k < rows_per_block; k += 8) {
    const int i = base + k;
    if (i >= row_end) break;
    const int *po = poff + (size_t)i * npts;
    const double *w = pw + (size_t)i * npts;
    double acc = 0.0;
    for (int t = lane; t < npts; t += 32) acc = fma(w[t], u_brick[po[t]], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      double bk = -acc;
      if (slab_pref != 0.0) bk -= ez[i] * (slab_pref * qz_sum[0]);
      const double bi = bk + b_real[i];
      b_kspace[i] = bk;
      b[i] = bi;
      sb[k] = bi;
    }
  }
  if (ps.arena) {  // warp r sends the block's rows to rank r: one contiguous line per peer
    __syncthreads();
    for (int r = warp; r < ps.nranks; r += 8)
      if (r != ps.rank && base + lane < row_end) peer_ptr<double>(ps, r, off_b)[base + lane] = sb[lane];
  }
  peer_block_signal(ps);
}

// electrode re-spread with the cached weights; one thread per (atom, n, m) row.
// The new charge q_i = (S.b)_i + potdiff*setq_i (+qinit_i) (fix_conp.cpp:1153-1158)
// is formed here from the epilogue scalars and stored by the atom's first thread.
__global__ void __launch_bounds__(256)
ele_spread_kernel(PPPMGeom g, int n_ele, int row_begin, int row_end, const int *__restrict__ widx,
                  const double *__restrict__ weights, const double *__restrict__ sb,
                  const double *__restrict__ setq, const double *__restrict__ qinit,
                  const double *__restrict__ scal, double *__restrict__ q_out, double *__restrict__ brick) {
  const int order = g.order;
  const int per_atom = order * order;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int i = (int)(gid / per_atom);
  if (i >= n_ele) return;
  // every rank forms all charges (the host scatters them to local + ghost atoms) but spreads only
  // its own rows; the bricks are summed across ranks when the host asks for the density
  const bool mine = i >= row_begin && i < row_end;
  const int nm = (int)(gid - (long long)i * per_atom);
  const int n = nm / order, m = nm - n * order;
  const double *w = weights + (size_t)i * 3 * order;
  const int *wi = widx + (size_t)i * 3 * order;
  double qi = (sb[i] - scal[13]) + scal[1] * setq[i];
  if (qinit) qi += qinit[i];
  if (nm == 0) q_out[i] = qi;
  if (!mine) return;
  const double z0 = g.delvolinv * qi;  // pppm_conp.cpp:411
  const double x0 = z0 * w[2 * order + n] * w[order + m];
  double *row = brick + ((size_t)wi[2 * order + n] * g.ny + wi[order + m]) * g.nx;
  for (int l = 0; l < order; ++l) atomicAdd(row + wi[l], x0 * w[l]);
}

// PPPMCONP::compute_particle_potential, the mesh sum (pppm_conp.cpp:452-484): one warp per point, lanes over
// the order^3 stencil, weights by Horner from the point's position; u is the full periodic mesh
__global__ void __launch_bounds__(256)
mesh_potential_kernel(PPPMGeom g, const double *__restrict__ rho_coeff, int n, const double *__restrict__ xyz,
                      const double *__restrict__ u, double *__restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + warp;
  if (i >= n) return;
  const int order = g.order, npts = order * order * order;
  const double fx = (xyz[3 * i] - g.boxlo[0]) * g.delinv[0];
  const double fy = (xyz[3 * i + 1] - g.boxlo[1]) * g.delinv[1];
  const double fz = (xyz[3 * i + 2] - g.boxlo[2]) * g.delinv[2];
  const int nx = (int)(fx + g.shift) - OFFSET, ny = (int)(fy + g.shift) - OFFSET, nz = (int)(fz + g.shift) - OFFSET;
  const double dx = nx + g.shiftone - fx, dy = ny + g.shiftone - fy, dz = nz + g.shiftone - fz;
  double acc = 0.0;
  for (int t = lane; t < npts; t += 32) {
    const int nn = t / (order * order), r = t - nn * order * order;
    const int m = r / order, l = r - m * order;
    const double z0 = rho1d(rho_coeff, order, nn, dz);
    const double y0 = z0 * rho1d(rho_coeff, order, m, dy);
    const double x0 = y0 * rho1d(rho_coeff, order, l, dx);
    const int mz = wrapi(nn + g.nlower + nz, g.nz), my = wrapi(m + g.nlower + ny, g.ny), mx = wrapi(l + g.nlower + nx, g.nx);
    acc = fma(x0, u[((size_t)mz * g.ny + my) * g.nx + mx], acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[i] = acc;
}

__global__ void __launch_bounds__(256)
add_bricks_kernel(size_t n, const double *__restrict__ a, const double *__restrict__ b, double *__restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}

}  // namespace

int launch_fill_zero(cudaStream_t s, double *p, size_t n) {
  if (n == 0) return 0;
  const unsigned grid = (unsigned)std::min<size_t>((n / 2 + 255) / 256 + 1, (size_t)148 * 8);
  fill_zero_kernel<<<grid, 256, 0, s>>>(p, n);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_pppm_spread(cudaStream_t s, const PPPMGeom &g, const double *rho_coeff, int m_bound, const PosQ *atoms,
                       double *brick, int *range_flag, const int *inbox_counts, int nsenders, int mpad,
                       const int *valid) {
  if (m_bound <= 0 || g.zs_n <= 0) return 0;
  const long long threads = (long long)m_bound * g.order * g.order;
  const unsigned grid = (unsigned)((threads + 255) / 256);
  // The kernel is bound by the SM's red.global issue rate, not by occupancy.  Asking for a slice of
  // (unused) dynamic shared memory caps it at SPREAD_BLOCKS_PER_SM blocks per SM, which leaves
  // registers and warp slots for the real-space pair kernel that runs beside it on the side stream.
  static int per_sm = -1;
  static size_t smem = 0;
  if (per_sm < 0) {
    const char *e = getenv("CONP_SPREAD_BLOCKS_PER_SM");
    per_sm = e ? atoi(e) : SPREAD_BLOCKS_PER_SM;
    if (per_sm > 0 && per_sm < 8) smem = ((size_t)(227 * 1024) / per_sm - 1024) & ~(size_t)127;
  }
  if (smem > 0) ensure_dynamic_smem(spread_kernel, smem);
  spread_kernel<<<grid, 256, smem, s>>>(g, rho_coeff, m_bound, atoms, inbox_counts, nsenders, mpad, valid, brick,
                                        range_flag);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

// Tile decomposition of the rank's slab of input planes and, per tile, the x-contiguous runs of cells
// whose charges can reach it.  A charge with stencil origin n (per axis, n = int(f + shift) - OFFSET, f the
// position in mesh units) touches the mesh indices n + nlower .. n + nlower + order - 1 (mod the mesh), so
// the tile [g0, g0 + e) is reached from f in [g0 - nlower - (order-1) - s, g0 + e - nlower - s), s = 0.5
// for odd orders and 0 for even ones; the interval is widened by 0.01 mesh cells against rounding, mapped
// to cell indices (floor(f * cinv / delinv), the same expression the sort uses up to that margin) and,
// along periodic axes, also taken one period up and down.  The kernel re-tests every charge exactly.
void plan_pppm_spread_tiles(const PPPMGeom &g, const CellGrid &cells, int num_sms, std::vector<int> &run_start,
                            std::vector<int2> &runs, SpreadPlan &plan) {
  const int P = g.order;
  int T[3] = {32, 8, 4};  // x, y, z
  if (const char *e = getenv("CONP_SPREAD_TILE")) {
    int a = 0, b = 0, c = 0;
    if (sscanf(e, "%d,%d,%d", &a, &b, &c) == 3 && a > 0 && b > 0 && c > 0) { T[2] = a; T[1] = b; T[0] = c; }
  }
  // tensor-core kernel: 8 rows x 32 columns x (4 | 8) planes in registers, no whole-axis tiles
  const bool mma_shape = T[0] == SM_TX && T[1] == SM_TY && (T[2] == 4 || T[2] == 8);
  const int len[3] = {g.nx, g.ny, g.zs_n};
  const int mod[3] = {g.nx, g.ny, g.nz};
  // can a stencil wrap around inside this rank's index range?  x, y: always (periodic mesh); z: only if the
  // slab is the whole periodic mesh
  const bool wraps[3] = {true, true, g.zs_n == g.nz};
  int nt[3], te[3], halo[3];
  for (int a = 0; a < 3; ++a) {
    if (len[a] <= T[a] + P - 1) {
      nt[a] = 1; te[a] = std::max(len[a], 1); halo[a] = wraps[a] ? P - 1 : 0;
    } else {
      nt[a] = (len[a] + T[a] - 1) / T[a]; te[a] = T[a]; halo[a] = 0;
    }
  }
  const int want_mma = plan.use_mma;  // set by the caller (an input): may the tensor-core kernel be used?
  plan = SpreadPlan();
  run_start.clear();
  runs.clear();
  if (g.zs_n <= 0) { run_start.push_back(0); return; }
  plan.tx = te[0]; plan.ty = te[1]; plan.tz = te[2];
  plan.ntx = nt[0]; plan.nty = nt[1]; plan.ntz = nt[2];
  plan.halo_x = halo[0]; plan.halo_y = halo[1]; plan.halo_z = halo[2];
  plan.ntiles = nt[0] * nt[1] * nt[2];
  // row stride == order (mod 16) doubles: the order x order footprint of one charge then covers distinct
  // 8-byte bank pairs (two wavefronts for 25 lanes, the minimum)
  const int ax = te[0] + halo[0], ay = te[1] + halo[1], az = te[2] + halo[2];
  plan.rs = ax + (((P - ax) % 16) + 16) % 16;
  plan.ps = ay * plan.rs;
  plan.smem = sizeof(double) * ((size_t)az * plan.ps + 32 * (size_t)((3 * P) | 1));
  const double s = g.shift - 16384.0;
  auto axis_cells = [&](int a, int g0, int e, std::vector<int> &out) {
    out.clear();
    const int nc = cells.nc[a];
    const double ratio = cells.cinv[a] / g.delinv[a];
    const double fa = (double)(g0 - g.nlower - (P - 1)) - s - 0.01, fb = (double)(g0 + e - g.nlower) - s + 0.01;
    std::vector<char> mark(nc, 0);
    if (cells.periodic[a]) {
      const double period = cells.prd[a] * g.delinv[a];  // positions are wrapped into [0, period) mesh units
      if (fb - fa >= mod[a]) {
        std::fill(mark.begin(), mark.end(), 1);
      } else {
        for (int j = -1; j <= 1; ++j) {
          const double a0 = fa + j * (double)mod[a], b0 = fb + j * (double)mod[a];
          if (b0 < 0.0 || a0 > period) continue;
          const int lo = std::max(0, (int)std::floor(a0 * ratio)), hi = std::min(nc - 1, (int)std::floor(b0 * ratio));
          for (int c = lo; c <= hi; ++c) mark[c] = 1;
        }
      }
    } else {  // charges beyond the box sit in the edge cells (cell_coord clamps)
      const int lo = std::max(0, std::min(nc - 1, (int)std::floor(fa * ratio)));
      const int hi = std::max(0, std::min(nc - 1, (int)std::floor(fb * ratio)));
      for (int c = lo; c <= hi; ++c) mark[c] = 1;
    }
    for (int c = 0; c < nc; ++c)
      if (mark[c]) out.push_back(c);
  };
  std::vector<std::vector<int>> cx(nt[0]), cy(nt[1]), cz(nt[2]);
  for (int i = 0; i < nt[0]; ++i) axis_cells(0, i * te[0], std::min(te[0], len[0] - i * te[0]), cx[i]);
  for (int i = 0; i < nt[1]; ++i) axis_cells(1, i * te[1], std::min(te[1], len[1] - i * te[1]), cy[i]);
  for (int i = 0; i < nt[2]; ++i)
    axis_cells(2, g.zin_lo + g.zs_lo + i * te[2], std::min(te[2], len[2] - i * te[2]), cz[i]);
  for (int iz = 0; iz < nt[2]; ++iz)
    for (int iy = 0; iy < nt[1]; ++iy)
      for (int ix = 0; ix < nt[0]; ++ix) {
        run_start.push_back((int)runs.size());
        const size_t first = runs.size();
        for (int z : cz[iz])
          for (int y : cy[iy]) {
            const int base = (z * cells.nc[1] + y) * cells.nc[0];
            const std::vector<int> &xs = cx[ix];
            for (size_t k = 0; k < xs.size();) {
              size_t k2 = k + 1;
              while (k2 < xs.size() && xs[k2] == xs[k2 - 1] + 1) ++k2;
              const int c0 = base + xs[k], c1 = base + xs[k2 - 1] + 1;
              if (runs.size() > first && runs.back().y == c0) runs.back().y = c1;  // contiguous in cell order
              else runs.push_back(make_int2(c0, c1));
              k = k2;
            }
          }
      }
  run_start.push_back((int)runs.size());
  const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(32, (size_t)(227 * 1024 - 1024) / (plan.smem + 1024)));
  plan.grid = std::min(plan.ntiles, num_sms * per_sm);
  plan.use_mma = want_mma && mma_shape && !halo[0] && !halo[1] && !halo[2] && nt[0] * te[0] >= len[0] &&
                 te[0] == SM_TX && te[1] == SM_TY && (te[2] == 4 || te[2] == 8);
  // the tensor-core kernel keeps its tile in registers: one-warp CTAs, as many as the register file holds
  plan.grid_mma = std::min(plan.ntiles, num_sms * (plan.tz == 4 ? 16 : 9));
}

int launch_pppm_spread_tiles(cudaStream_t s, const PPPMGeom &g, const SpreadPlan &plan, const double *rho_coeff_host,
                             const PosQ *atoms, const int *cell_start, int m_bound, const int *count_ptr,
                             double *brick, int *range_flag) {
  if (plan.ntiles <= 0 || g.zs_n <= 0) return 0;
  CUDA_CHECK(cudaMemsetAsync(plan.counter, 0, sizeof(int), s));
  RhoCoeff rho_coeff;
  std::memset(&rho_coeff, 0, sizeof(rho_coeff));
  std::memcpy(rho_coeff.c, rho_coeff_host, sizeof(double) * g.order * g.order);
  // no whole-axis tiles, default shape: accumulate on the FP64 tensor cores (CONP_SPREAD_SMEM=1: A/B run of
  // the shared-memory tile kernel)
  if (plan.use_mma) {
    const int mb = m_bound;
    if (mb <= 0) {
      CUDA_CHECK(cudaMemsetAsync(brick, 0, sizeof(double) * (size_t)g.zs_n * g.ny * g.nx, s));
      return 0;
    }
#define CONP_SPM_CASE(P_)                                                                                          \
  case P_:                                                                                                         \
    stencil_prepass_kernel<P_><<<(mb + 255) / 256, 256, 0, s>>>(g, rho_coeff, mb, count_ptr, atoms, plan.origin,  \
                                                                plan.weights, plan.wstride, range_flag);          \
    if (plan.tz == 4)                                                                                              \
      spread_mma_kernel<P_, 4><<<plan.grid_mma, 32, 0, s>>>(g, plan, plan.origin, plan.weights, plan.wstride,     \
                                                            cell_start, brick);                                   \
    else                                                                                                           \
      spread_mma_kernel<P_, 8><<<plan.grid_mma, 32, 0, s>>>(g, plan, plan.origin, plan.weights, plan.wstride,     \
                                                            cell_start, brick);                                   \
    break;
    switch (g.order) {
      CONP_SPM_CASE(1) CONP_SPM_CASE(2) CONP_SPM_CASE(3) CONP_SPM_CASE(4) CONP_SPM_CASE(5) CONP_SPM_CASE(6)
      CONP_SPM_CASE(7)
      default: CONP_THROW(CONP_ERR_ARG, "PPPM order %d not supported", g.order);
    }
#undef CONP_SPM_CASE
    CUDA_CHECK(cudaGetLastError());
    return 2;
  }
  // LAMMPS' default order with the default tile shape: strides known at compile time
  constexpr int RS5 = 37, PS5 = 8 * 37;
  if (g.order == 5 && plan.rs == RS5 && plan.ps == PS5) {
    ensure_dynamic_smem(spread_tile_kernel<5, RS5, PS5>, plan.smem);
    spread_tile_kernel<5, RS5, PS5><<<plan.grid, 32, plan.smem, s>>>(g, plan, rho_coeff, atoms, cell_start, brick,
                                                                     range_flag);
    CUDA_CHECK(cudaGetLastError());
    return 1;
  }
#define CONP_SPT_CASE(P_)                                                                                        \
  case P_:                                                                                                       \
    ensure_dynamic_smem(spread_tile_kernel<P_, 0, 0>, plan.smem);                                                \
    spread_tile_kernel<P_, 0, 0><<<plan.grid, 32, plan.smem, s>>>(g, plan, rho_coeff, atoms, cell_start, brick, \
                                                                  range_flag);                                  \
    break;
  switch (g.order) {
    CONP_SPT_CASE(1) CONP_SPT_CASE(2) CONP_SPT_CASE(3) CONP_SPT_CASE(4) CONP_SPT_CASE(5) CONP_SPT_CASE(6)
    CONP_SPT_CASE(7)
    default: CONP_THROW(CONP_ERR_ARG, "PPPM order %d not supported", g.order);
  }
#undef CONP_SPT_CASE
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

// z-sweep spread: columns of 8 x 32 mesh points, origin-plane bins, (column, segment) work items
void plan_pppm_sweep(const PPPMGeom &g, int num_sms, std::vector<int4> &items, SweepPlan &plan) {
  plan = SweepPlan();
  items.clear();
  const int P = g.order;
  if (g.zs_n <= 0) return;
  const int ncx = (g.nx + SW_FX - 1) / SW_FX, ncy = (g.ny + SW_FY - 1) / SW_FY;
  // a stencil must not wrap inside a column nor jump over a narrow last column
  if (ncx < 2 || ncy < 2 || g.nx - (ncx - 1) * SW_FX < P - 1 || g.ny - (ncy - 1) * SW_FY < P - 1) return;
  if (g.nx <= SW_FX + P - 1 || g.ny <= SW_FY + P - 1) return;
  plan.ncolx = ncx; plan.ncoly = ncy;
  plan.wrap_z = g.nzi == g.nz;
  if (plan.wrap_z) {  // periodic z: origin planes below the slab wrap around the ring
    plan.pz_lo = g.zs_n == g.nz ? 0 : g.zs_lo - (P - 1);
    plan.npz = g.zs_n == g.nz ? g.nz : g.zs_n + (P - 1);
    if (plan.npz > g.nz) return;
  } else {
    plan.pz_lo = std::max(0, g.zs_lo - (P - 1));
    plan.npz = g.zs_lo + g.zs_n - plan.pz_lo;
  }
  plan.nbins = ncx * ncy * plan.npz;
  const int ncols = ncx * ncy;
  const int per_sm = 12;  // register-limited (~165 per lane) one-warp CTAs per SM
  const int target = num_sms * per_sm;
  int nseg = std::max(1, std::min((target + ncols - 1) / ncols, std::max(1, g.zs_n / 16)));
  nseg = std::max(nseg, (g.zs_n + (SW_MAXS - P) - 1) / (SW_MAXS - P));  // segment + warm-up <= SW_MAXS plane-steps
  const int seg = (g.zs_n + nseg - 1) / nseg;
  for (int t0 = 0; t0 < g.zs_n; t0 += seg)  // segment-major: neighbouring CTAs work on the same planes (L2 locality)
    for (int c = 0; c < ncols; ++c) items.push_back(make_int4(c, t0, std::min(t0 + seg, g.zs_n), 0));
  plan.nitems = (int)items.size();
  plan.grid = std::min(plan.nitems, target);
  plan.usable = true;
}

int launch_pppm_spread_sweep(cudaStream_t s, const PPPMGeom &g, const SweepPlan &sp, const double *rho_coeff_host,
                             const PosQ *atoms, int m_bound, const int *valid, double *brick, int *range_flag) {
  if (!sp.usable || g.zs_n <= 0) return 0;
  if (m_bound <= 0) {
    CUDA_CHECK(cudaMemsetAsync(brick, 0, sizeof(double) * (size_t)g.zs_n * g.ny * g.nx, s));
    return 0;
  }
  RhoCoeff rc;
  std::memset(&rc, 0, sizeof(rc));
  std::memcpy(rc.c, rho_coeff_host, sizeof(double) * g.order * g.order);
  SweepGeom sg;
  sg.ncolx = sp.ncolx; sg.ncoly = sp.ncoly; sg.pz_lo = sp.pz_lo; sg.npz = sp.npz; sg.wrap_z = sp.wrap_z;
  CUDA_CHECK(cudaMemsetAsync(sp.bin_count, 0, sizeof(int) * ((size_t)sp.nbins + 8), s));
  CUDA_CHECK(cudaMemsetAsync(sp.counter, 0, sizeof(int), s));
  const unsigned gb = (unsigned)((m_bound + 255) / 256);
#define CONP_SWEEP_CASE(P_)                                                                                          \
  case P_:                                                                                                           \
    mesh_bin_kernel<P_><<<gb, 256, 0, s>>>(g, sg, m_bound, valid, atoms, sp.bin_of, sp.slot, sp.bin_count,          \
                                           range_flag);                                                             \
    launch_cell_scan(s, sp.nbins, sp.bin_count, sp.bin_start, nullptr, 1, 1, nullptr);                               \
    mesh_scatter_kernel<P_><<<gb, 256, 0, s>>>(g, rc, m_bound, atoms, sp.bin_of, sp.slot, sp.bin_start,             \
                                               sp.records);                                                         \
    spread_sweep_kernel<P_><<<sp.grid, 32, 0, s>>>(g, sg, sp.nitems, sp.items, sp.counter, sp.bin_start, sp.records, \
                                                   brick);                                                          \
    break;
  switch (g.order) {
    CONP_SWEEP_CASE(1) CONP_SWEEP_CASE(2) CONP_SWEEP_CASE(3) CONP_SWEEP_CASE(4) CONP_SWEEP_CASE(5)
    CONP_SWEEP_CASE(6) CONP_SWEEP_CASE(7)
    default: CONP_THROW(CONP_ERR_ARG, "PPPM order %d not supported", g.order);
  }
#undef CONP_SWEEP_CASE
  CUDA_CHECK(cudaGetLastError());
  return 4;
}

int launch_region_gather(cudaStream_t s, const PPPMGeom &g, int which, const int lo[3], const int hi[3],
                         const double *elyte, const double *ele, double *out) {
  const int ex = hi[0] - lo[0] + 1, ey = hi[1] - lo[1] + 1, ez = hi[2] - lo[2] + 1;
  const size_t n = (size_t)ex * ey * ez;
  region_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(g, which, lo[0], lo[1], lo[2], ex, ey, ez, elyte,
                                                                  ele, out);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_pppm_green_mul(cudaStream_t s, size_t n, cufftDoubleComplex *work, const double *ghalf) {
  green_mul_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, work, ghalf);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

namespace {
size_t zc_narrow_smem(int npcap, int rcap, bool real_k) {
  return sizeof(double2) * (size_t)npcap * ZC_COLS +
         (real_k ? sizeof(double) : sizeof(double2)) * (size_t)ZC_COLS * (2 * rcap + 1);
}
}  // namespace

// Split the column groups between the narrow and the wide path and describe, for every narrow
// group, which planes of this rank's slab [zs_lo, zs_lo + nzl) it stages.  The narrow path is sized for
// the largest window radius whose staging still fits ZC_NARROW_SMEM; np(R) = planes of [0, nzi)
// within R (on the ring) of an output plane.
void plan_pppm_zconv(const std::vector<int> &krad, int ncol, int nz, int nzi, int zs_lo, int nzl, int zin_lo,
                     const std::vector<int> &zout, bool real_k, std::vector<ZconvGroup> &narrow,
                     std::vector<int> &wide, std::vector<int> &aout, ZconvPlan &plan) {
  constexpr size_t ZC_NARROW_SMEM = 44 * 1024;
  narrow.clear();
  wide.clear();
  aout.clear();
  plan = ZconvPlan();
  // distance of every input plane to the nearest output plane (on the ring)
  std::vector<int> dist(std::max(nzi, 1), nz);
  for (int zo : zout) {
    int a = (zo - zin_lo) % nz;
    if (a < 0) a += nz;
    aout.push_back(a);
    for (int zi = 0; zi < nzi; ++zi) {
      const int d = std::abs(a - zi);
      dist[zi] = std::min(dist[zi], std::min(d, nz - d));
    }
  }
  std::vector<int> npof(nz + 1, 0);  // np(R), cumulative histogram of dist
  for (int zi = 0; zi < nzi; ++zi) npof[std::min(dist[zi], nz)]++;
  for (int r = 1; r <= nz; ++r) npof[r] += npof[r - 1];
  int rcap = -1;
  for (int r = 0; 2 * r + 1 < nz; ++r) {
    if (zc_narrow_smem(npof[r], r, real_k) > ZC_NARROW_SMEM) break;
    rcap = r;
  }
  const int ngroups = (ncol + ZC_COLS - 1) / ZC_COLS;
  int npmax = 1;
  for (int gi = 0; gi < ngroups; ++gi) {
    int rb = 0;
    for (int cc = 0; cc < ZC_COLS; ++cc) rb = std::max(rb, krad[std::min(gi * ZC_COLS + cc, ncol - 1)]);
    ZconvGroup g;
    std::memset(&g, 0, sizeof(g));
    g.c0 = gi * ZC_COLS;
    g.rblock = rb;
    bool ok = rb <= rcap;
    if (ok) {  // intervals of slab planes within rb of an output plane
      int t = 0;
      while (t < nzl && ok) {
        if (dist[zs_lo + t] > rb) { ++t; continue; }
        int e = t;
        while (e < nzl && dist[zs_lo + e] <= rb) ++e;
        if (g.nint == ZconvGroup::MAXI) { ok = false; break; }
        g.lo[g.nint] = t;
        g.hi[g.nint] = e;
        g.base[g.nint] = g.np;
        g.np += e - t;
        ++g.nint;
        t = e;
      }
    }
    if (ok) {
      narrow.push_back(g);
      npmax = std::max(npmax, g.np);
    } else {
      for (int cc = 0; cc < ZC_COLS && gi * ZC_COLS + cc < ncol; ++cc) wide.push_back(gi * ZC_COLS + cc);
    }
  }
  plan.n_narrow = (int)narrow.size();
  plan.n_wide = (int)wide.size();
  plan.rcap = std::max(rcap, 0);
  plan.npcap = npmax;
}

int launch_pppm_zconv(cudaStream_t s, int ncol, int nz, int nzl, int zs_lo, int nzo, const int *krad,
                      const ZconvPlan &plan, const cufftDoubleComplex *rhat, const double *Kr,
                      const cufftDoubleComplex *Kc, cufftDoubleComplex *uhat, const PeerSync &ps) {
  const int grid = plan.n_narrow + plan.n_wide;
  if (grid <= 0) return 0;
  const size_t smem = zc_narrow_smem(plan.npcap, plan.rcap, Kr != nullptr);
  if (Kr) {
    ensure_dynamic_smem(zconv_kernel<true>, smem);
    zconv_kernel<true><<<grid, ZC_THREADS, smem, s>>>(ncol, nz, nzl, zs_lo, nzo, plan.aout, krad, plan.narrow,
                                                      plan.n_narrow, plan.wide, plan.rcap, plan.npcap,
                                                      (const double2 *)rhat, Kr, nullptr, (double2 *)uhat, ps);
  } else {
    ensure_dynamic_smem(zconv_kernel<false>, smem);
    zconv_kernel<false><<<grid, ZC_THREADS, smem, s>>>(ncol, nz, nzl, zs_lo, nzo, plan.aout, krad, plan.narrow,
                                                       plan.n_narrow, plan.wide, plan.rcap, plan.npcap,
                                                       (const double2 *)rhat, nullptr, (const double2 *)Kc,
                                                       (double2 *)uhat, ps);
  }
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_expand_planes(cudaStream_t s, size_t plane, int nplanes, int nz, int lo, const int *list,
                         const double *compact, double *full) {
  const size_t tot = plane * (size_t)nplanes;
  expand_planes_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(plane, nplanes, nz, lo, list, compact, full);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_pppm_ele_stencil(cudaStream_t s, const PPPMGeom &g, const double *rho_coeff, int n, const double *ex,
                            const double *ey, const double *ez, int *part2grid, double *weights) {
  if (n <= 0) return 0;
  ele_stencil_kernel<<<(n + 127) / 128, 128, 0, s>>>(g, rho_coeff, n, ex, ey, ez, part2grid, weights);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_pppm_ele_index(cudaStream_t s, const PPPMGeom &g, int n, const int *part2grid, int *widx) {
  if (n <= 0) return 0;
  ele_index_kernel<<<(n + 127) / 128, 128, 0, s>>>(g, n, part2grid, widx);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_pppm_point_table(cudaStream_t s, const PPPMGeom &g, int n, const int *widx, const double *weights,
                            int *poff, double *pw) {
  if (n <= 0) return 0;
  const long long threads = (long long)n * g.order * g.order * g.order;
  ele_point_table_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(g, n, widx, weights, poff, pw);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_pppm_gather_b(cudaStream_t s, const PPPMGeom &g, int row_begin, int row_end, const int *poff,
                         const double *pw, const double *u_brick, const double *ez, const double *qz_sum,
                         double slab_pref, const double *b_real, double *b_kspace, double *b, const PeerSync &ps,
                         size_t off_b) {
  const int n = row_end - row_begin;
  if (n <= 0 && !ps.arena) return 0;
  // a rank without rows still takes part in the exchange: one block that only signals
  const int rpb = ps.arena ? GB_ROWS : 8;  // one GPU: a row per warp, as many blocks as possible
  gather_b_kernel<<<std::max((n + rpb - 1) / rpb, 1), 256, 0, s>>>(g, row_begin, row_end, poff, pw, u_brick, ez, qz_sum,
                                                                   slab_pref, b_real, b_kspace, b, ps, off_b, rpb);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_pppm_ele_spread(cudaStream_t s, const PPPMGeom &g, int n, int row_begin, int row_end, const int *widx,
                           const double *weights, const double *sb, const double *setq, const double *qinit,
                           const double *scal, double *q_out, double *brick) {
  if (n <= 0) return 0;
  const long long threads = (long long)n * g.order * g.order;
  ele_spread_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(g, n, row_begin, row_end, widx, weights, sb,
                                                                     setq, qinit, scal, q_out, brick);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_mesh_potential(cudaStream_t s, const PPPMGeom &g, const double *rho_coeff, int n, const double *xyz,
                          const double *u_full, double *out) {
  if (n <= 0) return 0;
  mesh_potential_kernel<<<(n + 7) / 8, 256, 0, s>>>(g, rho_coeff, n, xyz, u_full, out);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_add_bricks(cudaStream_t s, size_t n, const double *a, const double *b, double *out) {
  add_bricks_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, a, b, out);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

}  // namespace conp
