// q = S.b : FP64 GEMV streaming the row block of S from HBM with TMA bulk
// copies (cp.async.bulk -> SASS UBLKCP) through an mbarrier ring, plus the
// conp/conq/cond charge epilogue.
//
// Replaces the ddot_ loop of FixConp::update_charge (fix_conp.cpp:1135-1139),
// FixConq::update_charge (fix_conq.cpp:57-66) and FixCond::update_charge
// (fix_cond.cpp:86-95), and the epilogues at fix_conp.cpp:1143-1159,
// fix_conq.cpp:74-86, fix_cond.cpp:99-123.
//
// Roofline: HBM.  Algorithmic bytes per launch = 8*nrows*N (S) + 8*N (b) +
// 8*nrows (out).  One persistent CTA per SM; each CTA owns a contiguous strip
// of rows and streams it as [R rows x C cols] stages, 36 KB per stage
// (32 KB of S + the 4 KB slice of b from L2), STAGES deep => ~216 KB of loads
// in flight per SM, enough to cover HBM latency at 6.5 TB/s / 148 SMs.
#include "common.cuh"

#include <algorithm>
#include <cstring>

namespace conp {

namespace {

constexpr int R = 8;            // rows per stage
constexpr int C = 512;          // columns per stage
constexpr int STAGES = 6;
constexpr int CONSUMER_WARPS = 8;
constexpr int CONSUMERS = CONSUMER_WARPS * 32;  // one double2 column pair per thread
constexpr int THREADS = CONSUMERS + 32;         // + producer warp
static_assert(C == 2 * CONSUMERS, "each consumer thread owns one double2 per row");

struct __align__(128) Stage {
  double tile[R][C];
  double bs[C];
};

struct Smem {
  Stage st[STAGES];
  double red[2][CONSUMER_WARPS][R];
  unsigned long long full[STAGES];
  unsigned long long empty[STAGES];
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
  uint32_t done;
  uint32_t addr = smem_u32(bar);
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
// TMA 1-D bulk copy global -> shared, completion counted on an mbarrier
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar,
                                             uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}

// ---------------------------------------------------------------------------
// update_charge epilogue (fix_conp.cpp:1143-1159, fix_conq.cpp:74-86,
// fix_cond.cpp:99-123), shared by the fused GEMV tail and the stand-alone kernel.
// Every block sums sum_{left} (S.b)_i and sum_i setz_i (S.b)_i over its own rows,
// publishes the two partials and takes a ticket; the last block to arrive
// reduces the partials with a fixed-shape tree (deterministic) and fixes the
// potential difference.  The charges q_i = (S.b)_i + potdiff*setq_i (+qinit_i)
// are formed by the consumer kernels (ele_spread / finalize_q).
// ---------------------------------------------------------------------------
__device__ __forceinline__ double block_reduce_ordered(double v, double *sh, int tid, int nthreads, int bar_id) {
  // fixed-shape tree over `nthreads` (power of two) threads
  sh[tid] = v;
  for (int o = nthreads >> 1; o > 0; o >>= 1) {
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthreads) : "memory");
    if (tid < o) sh[tid] += sh[tid + o];
  }
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthreads) : "memory");
  const double r = sh[0];
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthreads) : "memory");
  return r;
}

// called by `nthreads` threads (tid 0..nthreads-1) of every block; a, z are the
// thread's partial sums over rows of this block
__device__ void charge_epilogue(const ChargeEpilogue &ep, double a, double z, double e, int tid, int nthreads,
                                int bar_id, double *sh /* >= nthreads doubles */, int *sh_flag) {
  const double pl = block_reduce_ordered(a, sh, tid, nthreads, bar_id);
  const double pz = block_reduce_ordered(z, sh, tid, nthreads, bar_id);
  const double pe = block_reduce_ordered(e, sh, tid, nthreads, bar_id);
  if (tid == 0) {
    ep.partials[3 * blockIdx.x] = pl;
    ep.partials[3 * blockIdx.x + 1] = pz;
    ep.partials[3 * blockIdx.x + 2] = pe;
    __threadfence();
    const unsigned ticket = atomicAdd(ep.counter, 1u);
    *sh_flag = (ticket == gridDim.x - 1);
  }
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthreads) : "memory");
  if (!*sh_flag) return;
  __threadfence();
  double ta = 0.0, tz = 0.0, te = 0.0;
  for (unsigned k = tid; k < gridDim.x; k += nthreads) {
    ta += __ldcg(ep.partials + 3 * k);
    tz += __ldcg(ep.partials + 3 * k + 1);
    te += __ldcg(ep.partials + 3 * k + 2);
  }
  double tot_left = block_reduce_ordered(ta, sh, tid, nthreads, bar_id);
  double tot_z = block_reduce_ordered(tz, sh, tid, nthreads, bar_id);
  const double tot_all = block_reduce_ordered(te, sh, tid, nthreads, bar_id);
  if (tid == 0) {
    // e^T S = 0 holds only to the rounding of the projection (fix_conp.cpp:1011-1020) and of the N-term
    // sums: at N = 40 000 the charges add up to a few 1e-12 e.  With the projection on, that residual
    // is taken off every charge evenly (~1e-16 e per atom), which keeps the total at the 1e-14 level.
    const double mean = ep.neutral ? tot_all / ep.n : 0.0;
    tot_left -= mean * ep.n_left;
    tot_z -= mean * ep.sum_setz;
    ep.scalar_out[13] = mean;
    const double value = __ldcg(ep.value);
    double potdiff, scalar;
    if (ep.variant == CONP_VARIANT_CONP) {  // fix_conp.cpp:1149-1159
      potdiff = value;
      scalar = potdiff * ep.totsetq + tot_left;
    } else if (ep.variant == CONP_VARIANT_CONQ) {  // fix_conq.cpp:74-80
      const double netcharge_right = -tot_left;
      scalar = -(value - netcharge_right) / ep.totsetq;
      if (ep.one_electrode) scalar += 2 * value / ep.totsetq;
      potdiff = scalar;
    } else {  // fix_cond.cpp:101-115; dipole[0] = sum q z over non-electrode atoms
      const double dipole_all = -__ldcg(ep.dipole);
      potdiff = value - dipole_all / ep.lz;
      potdiff -= tot_z;
      potdiff *= ep.vmult;
      scalar = potdiff;
    }
    ep.scalar_out[0] = scalar;
    ep.scalar_out[1] = potdiff;
    *ep.counter = 0u;  // ready for the next launch
  }
}

__global__ void __launch_bounds__(THREADS, 1)
gemv_tma_kernel(const double *__restrict__ S, size_t pitch, int nrows, int ncols_pad,
                const double *__restrict__ b, double *__restrict__ out, ChargeEpilogue ep) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem &sm = *reinterpret_cast<Smem *>(smem_raw);

  // contiguous strip of rows for this CTA
  const int row_a = (int)(((long long)nrows * blockIdx.x) / gridDim.x);
  const int row_b = (int)(((long long)nrows * (blockIdx.x + 1)) / gridDim.x);
  if (row_a >= row_b) return;
  const int ngroups = (row_b - row_a + R - 1) / R;
  const int nchunks = (ncols_pad + C - 1) / C;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], CONSUMER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (warp == CONSUMER_WARPS) {
    // ===== producer: one elected lane issues all bulk copies =====
    if (lane == 0) {
      const uint64_t pol_s = policy_evict_first();  // S is streamed once
      const uint64_t pol_b = policy_evict_last();   // b is re-read by every CTA
      int stage = 0;
      uint32_t phase = 0;
      for (int g = 0; g < ngroups; ++g) {
        const int r0 = row_a + g * R;
        const int nr = min(R, row_b - r0);
        for (int c = 0; c < nchunks; ++c) {
          const int c0 = c * C;
          const int w = min(C, ncols_pad - c0);
          const uint32_t row_bytes = (uint32_t)w * 8u;
          mbar_wait(&sm.empty[stage], phase ^ 1);
          mbar_expect_tx(&sm.full[stage], row_bytes * (uint32_t)(nr + 1));
          tma_bulk_g2s(sm.st[stage].bs, b + c0, row_bytes, &sm.full[stage], pol_b);
          for (int r = 0; r < nr; ++r)
            tma_bulk_g2s(sm.st[stage].tile[r], S + (size_t)(r0 + r) * pitch + c0, row_bytes, &sm.full[stage],
                         pol_s);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    return;
  }

  // ===== consumers =====
  int stage = 0;
  uint32_t phase = 0;
  for (int g = 0; g < ngroups; ++g) {
    const int r0 = row_a + g * R;
    const int nr = min(R, row_b - r0);
    double acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.0;
    for (int c = 0; c < nchunks; ++c) {
      const int w = min(C, ncols_pad - c * C);
      mbar_wait(&sm.full[stage], phase);
      if (2 * tid < w) {
        const double2 bv = *reinterpret_cast<const double2 *>(&sm.st[stage].bs[2 * tid]);
        if (nr == R) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const double2 sv = *reinterpret_cast<const double2 *>(&sm.st[stage].tile[r][2 * tid]);
            acc[r] = fma(sv.x, bv.x, acc[r]);
            acc[r] = fma(sv.y, bv.y, acc[r]);
          }
        } else {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            if (r < nr) {
              const double2 sv = *reinterpret_cast<const double2 *>(&sm.st[stage].tile[r][2 * tid]);
              acc[r] = fma(sv.x, bv.x, acc[r]);
              acc[r] = fma(sv.y, bv.y, acc[r]);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.empty[stage]);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    // reduce the R partial dot products over the 256 consumer threads
#pragma unroll
    for (int r = 0; r < R; ++r) {
      double v = acc[r];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) sm.red[g & 1][warp][r] = v;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(CONSUMERS) : "memory");
    if (tid < nr) {
      double v = 0.0;
#pragma unroll
      for (int wv = 0; wv < CONSUMER_WARPS; ++wv) v += sm.red[g & 1][wv][tid];
      out[r0 + tid] = v;
    }
  }
  if (ep.enabled) {
    // the tile ring is idle now: reuse stage 0 as reduction scratch.  The barrier also makes this
    // block's rows of `out` visible to all of its threads.
    double *scratch = &sm.st[0].tile[0][0];
    int *flag = reinterpret_cast<int *>(&sm.red[0][0][0]);
    asm volatile("bar.sync 1, %0;" ::"n"(CONSUMERS) : "memory");
    double a = 0.0, z = 0.0, e = 0.0;
    for (int r = row_a + tid; r < row_b; r += CONSUMERS) {
      const double v = out[r];
      const int gi = ep.row_offset + r;
      if (ep.side[gi] == 1) a += v;
      z = fma(ep.setz[gi], v, z);
      e += v;
    }
    charge_epilogue(ep, a, z, e, tid, CONSUMERS, 1, scratch, flag);
  }
}

// stand-alone epilogue over the full (gathered) S.b vector: the multi-GPU path
constexpr int UC_THREADS = 256;

__global__ void __launch_bounds__(UC_THREADS)
update_charge_kernel(ChargeEpilogue ep) {
  __shared__ double sh[UC_THREADS];
  __shared__ int flag;
  double a = 0.0, z = 0.0, e = 0.0;
  for (int i = blockIdx.x * UC_THREADS + threadIdx.x; i < ep.n; i += gridDim.x * UC_THREADS) {
    const double v = ep.sb[i];
    if (ep.side[i] == 1) a += v;
    z = fma(ep.setz[i], v, z);
    e += v;
  }
  charge_epilogue(ep, a, z, e, threadIdx.x, UC_THREADS, 0, sh, &flag);
}

// Several GPUs, symmetric matvec, peer-to-peer path: every rank's partial S.b sits in slot r of the
// local staging area once the flags are up.  Sum the slots in rank order (the same order on every
// rank: bitwise identical charges everywhere), store S.b and run the epilogue -- the all-reduce of
// fix_conp.cpp:1140 and the epilogue in one kernel.
__global__ void __launch_bounds__(UC_THREADS)
update_charge_sum_kernel(ChargeEpilogue ep, PeerSync ps, const double *__restrict__ parts, int len,
                         double *__restrict__ sb_out) {
  __shared__ double sh[UC_THREADS];
  __shared__ int flag;
  peer_block_wait(ps);
  __syncthreads();
  double a = 0.0, z = 0.0, e = 0.0;
  for (int i = blockIdx.x * UC_THREADS + threadIdx.x; i < len; i += gridDim.x * UC_THREADS) {
    double v = 0.0;
    for (int r = 0; r < ps.nranks; ++r) v += __ldcg(parts + (size_t)r * len + i);
    sb_out[i] = v;
    if (i < ep.n) {
      if (ep.side[i] == 1) a += v;
      z = fma(ep.setz[i], v, z);
      e += v;
    }
  }
  charge_epilogue(ep, a, z, e, threadIdx.x, UC_THREADS, 0, sh, &flag);
}

// q_i = (S.b)_i + potdiff * setq_i (+ qinit_i): fix_conp.cpp:1153-1158
__global__ void __launch_bounds__(256)
finalize_q_kernel(int n, const double *__restrict__ sb, const double *__restrict__ setq,
                  const double *__restrict__ qinit, const double *__restrict__ scal, double *__restrict__ q_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double q = (sb[i] - scal[13]) + scal[1] * setq[i];
  if (qinit) q += qinit[i];
  q_out[i] = q;
}


// ---------------------------------------------------------------------------
// Symmetric variant.  S = A^-1 - w w^T/(e^T w) is symmetric (fix_conp.cpp:982-1020 applies the
// projection to the inverse of the symmetric A), so q = S.b needs every unordered pair {i,j} only
// once: streaming S_ij gives both S_ij b_j (to row i) and S_ij b_i (to row j), and the kernel reads
// half of the matrix.  Pair {i,j} is taken from row i iff (j - i) mod N is in [1, H], H = N/2 (for
// even N the distance-H pairs come from the smaller index): a cyclic half band, so every row has
// the same amount of work for any row partition -- one GPU or a row block per GPU alike.
//
// A CTA owns strips of consecutive rows.  For a strip [a, bnd) the band is the unwrapped column
// range [a & ~1, bnd + H); it is cut into chunks of C columns (two segments when it wraps past N).
// Loop order: chunk outer, row groups (R rows) inner, so the two column sums a thread owns stay in
// registers for the whole strip and are stored once per chunk into the strip's private slice of
// `colpart` (no atomics, deterministic).  Row sums are reduced per stage inside the warp and
// accumulated in a per-warp shared array.  symv_reduce_kernel then adds, per column, the strip
// slices in a fixed order, and runs the charge epilogue in its tail.
// Same TMA ring as gemv_tma_kernel; only the interior of the band takes the unmasked fast path.
// ---------------------------------------------------------------------------
constexpr int SY_STAGES = 5;
constexpr int SY_HMAX = 512;  // rows per strip (per-warp row sums live in shared memory)

struct __align__(128) SyStage {
  double tile[R][C];
};
struct SySmem {
  SyStage st[SY_STAGES];
  double yrow[CONSUMER_WARPS][SY_HMAX];
  double bstrip[SY_HMAX];
  unsigned long long full[SY_STAGES];
  unsigned long long empty[SY_STAGES];
};

struct SyChunk {
  int cb;    // first column, unwrapped (segment 2: N + actual column)
  int cev;   // end of the columns any row of the strip can use (exclusive, unwrapped)
  int w;     // columns copied (even)
  int cact;  // actual first column in S and b
  int joff;  // offset of the chunk in the strip's colpart slice
  bool seg2;
};

__device__ __forceinline__ bool sy_chunk(int k, int a, int bnd, int N, int H, SyChunk &ch) {
  const int CS = a & ~1, CE = bnd + H;
  const int end1 = min(CE, N), len1 = end1 - CS;
  const int n1 = (len1 + C - 1) / C;
  const int len2 = max(CE - N, 0);
  const int n2 = (len2 + C - 1) / C;
  if (k < n1) {
    ch.cb = CS + k * C;
    ch.cev = min(ch.cb + C, end1);
    ch.seg2 = false;
    ch.cact = ch.cb;
    ch.joff = k * C;
  } else if (k < n1 + n2) {
    const int kk = k - n1;
    ch.cb = N + kk * C;
    ch.cev = min(ch.cb + C, CE);
    ch.seg2 = true;
    ch.cact = kk * C;
    ch.joff = ((len1 + 1) & ~1) + kk * C;
  } else {
    return false;
  }
  ch.w = (ch.cev - ch.cb + 1) & ~1;
  return true;
}
// does any row of the group [rg, rg+nr) use a column of the chunk?
__device__ __forceinline__ bool sy_in_band(int rg, int nr, const SyChunk &ch, int H) {
  return rg <= ch.cev - 1 && rg + nr - 1 + H >= ch.cb;
}

__global__ void __launch_bounds__(THREADS, 1)
symv_tma_kernel(const double *__restrict__ S, size_t pitch, int N, int row0, const int2 *__restrict__ strips,
                int nstrips, int L, const double *__restrict__ b, double *__restrict__ rowpart,
                double *__restrict__ colpart, PeerSync ps_b) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SySmem &sm = *reinterpret_cast<SySmem *>(smem_raw);
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int H = N / 2;
  const bool tie = (N & 1) == 0;

  if (tid == 0) {
    for (int s = 0; s < SY_STAGES; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], CONSUMER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  int stage = 0;
  uint32_t phase = 0;

  if (warp == CONSUMER_WARPS) {
    // ===== producer =====
    if (lane == 0) {
      const uint64_t pol_s = policy_evict_first();
      for (int s = blockIdx.x; s < nstrips; s += gridDim.x) {
        const int a = strips[s].x, bnd = strips[s].y;
        if (a >= bnd) continue;
        const int ngroups = (bnd - a + R - 1) / R;
        SyChunk ch;
        for (int k = 0; sy_chunk(k, a, bnd, N, H, ch); ++k) {
          const uint32_t row_bytes = (uint32_t)ch.w * 8u;
          for (int g = 0; g < ngroups; ++g) {
            const int rg = a + g * R;
            const int nr = min(R, bnd - rg);
            if (!sy_in_band(rg, nr, ch, H)) continue;
            mbar_wait(&sm.empty[stage], phase ^ 1);
            mbar_expect_tx(&sm.full[stage], row_bytes * (uint32_t)nr);
            for (int r = 0; r < nr; ++r)
              tma_bulk_g2s(sm.st[stage].tile[r], S + (size_t)(rg - row0 + r) * pitch + ch.cact, row_bytes,
                           &sm.full[stage], pol_s);
            if (++stage == SY_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    return;
  }

  // ===== consumers =====
  // several GPUs: b is being written by the peers' gather kernels.  The producer warp is already
  // streaming S (which is local); the consumers poll the flags before their first read of b.
  if (ps_b.arena) {
    peer_block_wait(ps_b);
    asm volatile("bar.sync 1, %0;" ::"n"(CONSUMERS) : "memory");
  }
  for (int s = blockIdx.x; s < nstrips; s += gridDim.x) {
    const int a = strips[s].x, bnd = strips[s].y;
    if (a >= bnd) continue;
    const int h = bnd - a;
    const int ngroups = (h + R - 1) / R;
    for (int t = lane; t < h; t += 32) sm.yrow[warp][t] = 0.0;
    for (int t = tid; t < h; t += CONSUMERS) sm.bstrip[t] = b[a + t];
    asm volatile("bar.sync 1, %0;" ::"n"(CONSUMERS) : "memory");
    double *cp = colpart + (size_t)s * L;
    SyChunk ch, nxt;
    bool have = sy_chunk(0, a, bnd, N, H, ch);
    double2 bv = make_double2(0.0, 0.0);
    if (have && 2 * tid < ch.w) bv = *reinterpret_cast<const double2 *>(b + ch.cact + 2 * tid);
    for (int k = 0; have; ++k) {
      const bool active = 2 * tid < ch.w;
      // b of the next chunk is fetched now so that its L2 latency hides behind this chunk's stages
      const bool have_next = sy_chunk(k + 1, a, bnd, N, H, nxt);
      double2 bv_next = make_double2(0.0, 0.0);
      if (have_next && 2 * tid < nxt.w) bv_next = *reinterpret_cast<const double2 *>(b + nxt.cact + 2 * tid);
      const int cu0 = ch.cb + 2 * tid;  // unwrapped column of this thread's first element
      double col0 = 0.0, col1 = 0.0;
      for (int g = 0; g < ngroups; ++g) {
        const int rg = a + g * R;
        const int nr = min(R, bnd - rg);
        if (!sy_in_band(rg, nr, ch, H)) continue;
        mbar_wait(&sm.full[stage], phase);
        double acc[R];
#pragma unroll
        for (int i = 0; i < R; ++i) acc[i] = 0.0;
        if (active) {
          // interior of the band for all R rows: every element is a distinct pair with 1 <= d < H
          const bool fast = ch.cb - (rg + nr - 1) >= 1 && (ch.cb + ch.w - 1) - rg <= H - 1 &&
                            (ch.seg2 || ch.cb + ch.w <= N);
          if (fast) {
            // two interleaved column-sum chains per element so the FP64 latency of one hides the other
            double c0b = 0.0, c1b = 0.0;
#pragma unroll
            for (int i = 0; i < R; i += 2) {
              if (i < nr) {
                const double2 sv = *reinterpret_cast<const double2 *>(&sm.st[stage].tile[i][2 * tid]);
                const double br = sm.bstrip[rg - a + i];
                acc[i] = fma(sv.y, bv.y, sv.x * bv.x);
                col0 = fma(sv.x, br, col0);
                col1 = fma(sv.y, br, col1);
              }
              if (i + 1 < nr) {
                const double2 sv = *reinterpret_cast<const double2 *>(&sm.st[stage].tile[i + 1][2 * tid]);
                const double br = sm.bstrip[rg - a + i + 1];
                acc[i + 1] = fma(sv.y, bv.y, sv.x * bv.x);
                c0b = fma(sv.x, br, c0b);
                c1b = fma(sv.y, br, c1b);
              }
            }
            col0 += c0b;
            col1 += c1b;
          } else {
#pragma unroll
            for (int i = 0; i < R; ++i) {
              if (i < nr) {
                const int r = rg + i;
                const double2 sv = *reinterpret_cast<const double2 *>(&sm.st[stage].tile[i][2 * tid]);
                const double br = sm.bstrip[r - a];
                const int d0 = cu0 - r, d1 = d0 + 1;
                const bool ok0 = d0 >= 0 && d0 <= H && !(tie && d0 == H && r >= H) && (ch.seg2 || cu0 < N);
                const bool ok1 = d1 >= 0 && d1 <= H && !(tie && d1 == H && r >= H) && (ch.seg2 || cu0 + 1 < N);
                const double x0 = ok0 ? sv.x : 0.0, x1 = ok1 ? sv.y : 0.0;
                acc[i] = fma(x1, bv.y, x0 * bv.x);
                if (d0 >= 1) col0 = fma(x0, br, col0);  // d == 0 is the diagonal: row part only
                if (d1 >= 1) col1 = fma(x1, br, col1);
              }
            }
          }
        }
        // recursive-halving reduction of the R = 8 row partials over the warp: afterwards lane
        // 4*j holds the warp total of row (bit4, bit3, bit2 of the lane)
        {
          const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const double send = h16 ? acc[i] : acc[i + 4];
            const double keep = h16 ? acc[i + 4] : acc[i];
            acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
          }
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const double send = h8 ? acc[i] : acc[i + 2];
            const double keep = h8 ? acc[i + 2] : acc[i];
            acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
          }
          {
            const double send = h4 ? acc[0] : acc[1];
            const double keep = h4 ? acc[1] : acc[0];
            acc[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
          }
          acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], 2);
          acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], 1);
          const int row = (h16 ? 4 : 0) + (h8 ? 2 : 0) + (h4 ? 1 : 0);
          if ((lane & 3) == 0 && row < nr) sm.yrow[warp][rg - a + row] += acc[0];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[stage]);
        if (++stage == SY_STAGES) { stage = 0; phase ^= 1; }
      }
      if (active) *reinterpret_cast<double2 *>(cp + ch.joff + 2 * tid) = make_double2(col0, col1);
      ch = nxt;
      bv = bv_next;
      have = have_next;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(CONSUMERS) : "memory");
    for (int t = tid; t < h; t += CONSUMERS) {
      double v = 0.0;
#pragma unroll
      for (int wv = 0; wv < CONSUMER_WARPS; ++wv) v += sm.yrow[wv][t];
      rowpart[a + t] = v;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(CONSUMERS) : "memory");
  }
}

// out[c] = row part (own rows) + sum over strips of the strip's column slice.  A block owns tiles of
// 32 columns; warp w adds the strips s = w, w+8, ... for its lane's column, the eight partial sums
// are then added in warp order: a fixed summation shape, so the product is reproducible bit for
// bit.  The charge epilogue (single GPU) runs in the tail.
constexpr int SR_THREADS = 256;

__global__ void __launch_bounds__(SR_THREADS)
symv_reduce_kernel(int N, int out_len, int row0, int nrows, const int2 *__restrict__ strips, int nstrips, int L,
                   const double *__restrict__ rowpart, const double *__restrict__ colpart,
                   double *__restrict__ out, ChargeEpilogue ep, PeerSync ps, size_t off_parts) {
  __shared__ double sh[SR_THREADS];
  __shared__ double part[SR_THREADS / 32][32];
  __shared__ int flag;
  const int H = N / 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = (out_len + 31) / 32;
  double pa = 0.0, pz = 0.0, pe = 0.0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int c = tile * 32 + lane;
    double v = 0.0;
    if (c < N) {
#pragma unroll 4
      for (int s = warp; s < nstrips; s += SR_THREADS / 32) {
        const int2 ab = strips[s];
        const int CS = ab.x & ~1, CE = ab.y + H;
        const int end1 = min(CE, N);
        const bool in1 = c >= CS && c < end1;
        const bool in2 = !in1 && c < CE - N;
        const int j = in1 ? c - CS : ((end1 - CS + 1) & ~1) + c;
        if (ab.x < ab.y && (in1 || in2)) v += colpart[(size_t)s * L + j];
      }
    }
    part[warp][lane] = v;
    __syncthreads();
    if (warp == 0 && c < out_len) {
      double t = 0.0;
      if (c < N) {
        t = (c >= row0 && c < row0 + nrows) ? rowpart[c] : 0.0;
#pragma unroll
        for (int wv = 0; wv < SR_THREADS / 32; ++wv) t += part[wv][lane];
        if (ep.enabled) {
          if (ep.side[c] == 1) pa += t;
          pz = fma(ep.setz[c], t, pz);
          pe += t;
        }
      }
      if (ps.arena) {  // this rank's partial sum goes into slot `rank` of every rank's staging area
        for (int r = 0; r < ps.nranks; ++r) peer_ptr<double>(ps, r, off_parts)[(size_t)ps.rank * out_len + c] = t;
      } else {
        out[c] = t;
      }
    }
    __syncthreads();
  }
  if (ep.enabled) charge_epilogue(ep, pa, pz, pe, threadIdx.x, SR_THREADS, 0, sh, &flag);
  peer_block_signal(ps);
}

}  // namespace

int launch_gemv(cudaStream_t s, const double *S, size_t pitch, int nrows, int ncols_pad, const double *b,
                double *out, int num_sms, const ChargeEpilogue *ep) {
  if (nrows <= 0) return 0;
  const size_t smem = sizeof(Smem);
  ensure_dynamic_smem(gemv_tma_kernel, smem);
  int grid = num_sms < nrows ? num_sms : nrows;
  ChargeEpilogue e;
  if (ep) e = *ep;
  else { memset(&e, 0, sizeof(e)); }
  gemv_tma_kernel<<<grid, THREADS, smem, s>>>(S, pitch, nrows, ncols_pad, b, out, e);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

SymvPlan plan_symv(int N, int row0, int nrows, int num_sms, std::vector<int2> &strips) {
  SymvPlan p;
  strips.clear();
  if (N < 64) return p;
  if (nrows <= 0) {  // a rank without rows still takes part in the exchanges: no strips, zero partial sum
    p.usable = true;
    p.L = 2;
    return p;
  }
  const int grid = std::min(num_sms, nrows);
  const int nstrips = std::max(grid, (nrows + SY_HMAX - 1) / SY_HMAX);
  const int hmax = (nrows + nstrips - 1) / nstrips;
  const int H = N / 2;
  if (hmax > SY_HMAX || hmax + H + 2 > N) return p;  // the unwrapped band of a strip must not lap itself
  for (int s = 0; s < nstrips; ++s)
    strips.push_back(make_int2(row0 + (int)(((long long)nrows * s) / nstrips),
                               row0 + (int)(((long long)nrows * (s + 1)) / nstrips)));
  p.usable = true;
  p.grid = grid;
  p.nstrips = nstrips;
  p.L = (hmax + H + 8 + 1) & ~1;
  return p;
}

int launch_symv(cudaStream_t s, const double *S, size_t pitch, int N, int row0, int nrows, const double *b,
                const SymvPlan &plan, double *rowpart, double *colpart, double *out, int out_len,
                const ChargeEpilogue *ep, const PeerSync &wait_b, const PeerSync &push_parts, size_t off_parts) {
  if (!plan.usable || !plan.strips) CONP_THROW(CONP_ERR_STATE, "launch_symv: no plan");
  const size_t smem = sizeof(SySmem);
  ensure_dynamic_smem(symv_tma_kernel, smem);
  int launched = 1;
  if (plan.grid > 0) {
    symv_tma_kernel<<<plan.grid, THREADS, smem, s>>>(S, pitch, N, row0, plan.strips, plan.nstrips, plan.L, b,
                                                    rowpart, colpart, wait_b);
    CUDA_CHECK(cudaGetLastError());
    ++launched;
  } else if (wait_b.arena) {
    launched += p2p_wait_sync(wait_b, s);  // nobody here reads b, but the exchange must be consumed
  }
  ChargeEpilogue e;
  if (ep) e = *ep;
  else memset(&e, 0, sizeof(e));
  int grid = (out_len + 31) / 32;
  grid = grid < 1 ? 1 : (grid > 1024 ? 1024 : grid);  // epilogue partials: 3 per block, 1024 blocks max
  symv_reduce_kernel<<<grid, SR_THREADS, 0, s>>>(N, out_len, row0, nrows, plan.strips, plan.nstrips, plan.L, rowpart,
                                                 colpart, out, e, push_parts, off_parts);
  CUDA_CHECK(cudaGetLastError());
  return launched;
}

int launch_update_charge(cudaStream_t s, const ChargeEpilogue &ep) {
  int grid = (ep.n + UC_THREADS * 4 - 1) / (UC_THREADS * 4);
  grid = grid < 1 ? 1 : (grid > 128 ? 128 : grid);
  update_charge_kernel<<<grid, UC_THREADS, 0, s>>>(ep);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_update_charge_sum(cudaStream_t s, const ChargeEpilogue &ep, const PeerSync &ps, const double *parts,
                             int len, double *sb_out) {
  int grid = (len + UC_THREADS * 4 - 1) / (UC_THREADS * 4);
  grid = grid < 1 ? 1 : (grid > 128 ? 128 : grid);
  update_charge_sum_kernel<<<grid, UC_THREADS, 0, s>>>(ep, ps, parts, len, sb_out);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_finalize_q(cudaStream_t s, int n, const double *sb, const double *setq, const double *qinit,
                      const double *scal, double *q_out) {
  finalize_q_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, sb, setq, qinit, scal, q_out);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

}  // namespace conp
