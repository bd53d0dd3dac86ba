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
__device__ void charge_epilogue(const ChargeEpilogue &ep, double a, double z, int tid, int nthreads, int bar_id,
                                double *sh /* >= nthreads doubles */, int *sh_flag) {
  const double pl = block_reduce_ordered(a, sh, tid, nthreads, bar_id);
  const double pz = block_reduce_ordered(z, sh, tid, nthreads, bar_id);
  if (tid == 0) {
    ep.partials[2 * blockIdx.x] = pl;
    ep.partials[2 * blockIdx.x + 1] = pz;
    __threadfence();
    const unsigned ticket = atomicAdd(ep.counter, 1u);
    *sh_flag = (ticket == gridDim.x - 1);
  }
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthreads) : "memory");
  if (!*sh_flag) return;
  __threadfence();
  double ta = 0.0, tz = 0.0;
  for (unsigned k = tid; k < gridDim.x; k += nthreads) {
    ta += __ldcg(ep.partials + 2 * k);
    tz += __ldcg(ep.partials + 2 * k + 1);
  }
  const double tot_left = block_reduce_ordered(ta, sh, tid, nthreads, bar_id);
  const double tot_z = block_reduce_ordered(tz, sh, tid, nthreads, bar_id);
  if (tid == 0) {
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
    double a = 0.0, z = 0.0;
    for (int r = row_a + tid; r < row_b; r += CONSUMERS) {
      const double v = out[r];
      const int gi = ep.row_offset + r;
      if (ep.side[gi] == 1) a += v;
      z = fma(ep.setz[gi], v, z);
    }
    charge_epilogue(ep, a, z, tid, CONSUMERS, 1, scratch, flag);
  }
}

// stand-alone epilogue over the full (gathered) S.b vector: the multi-GPU path
constexpr int UC_THREADS = 256;

__global__ void __launch_bounds__(UC_THREADS)
update_charge_kernel(ChargeEpilogue ep) {
  __shared__ double sh[UC_THREADS];
  __shared__ int flag;
  double a = 0.0, z = 0.0;
  for (int i = blockIdx.x * UC_THREADS + threadIdx.x; i < ep.n; i += gridDim.x * UC_THREADS) {
    const double v = ep.sb[i];
    if (ep.side[i] == 1) a += v;
    z = fma(ep.setz[i], v, z);
  }
  charge_epilogue(ep, a, z, threadIdx.x, UC_THREADS, 0, sh, &flag);
}

// q_i = (S.b)_i + potdiff * setq_i (+ qinit_i): fix_conp.cpp:1153-1158
__global__ void __launch_bounds__(256)
finalize_q_kernel(int n, const double *__restrict__ sb, const double *__restrict__ setq,
                  const double *__restrict__ qinit, const double *__restrict__ scal, double *__restrict__ q_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double q = sb[i] + scal[1] * setq[i];
  if (qinit) q += qinit[i];
  q_out[i] = q;
}

}  // namespace

int launch_gemv(cudaStream_t s, const double *S, size_t pitch, int nrows, int ncols_pad, const double *b,
                double *out, int num_sms, const ChargeEpilogue *ep) {
  if (nrows <= 0) return 0;
  static bool attr_set = false;
  const size_t smem = sizeof(Smem);
  if (!attr_set) {
    CUDA_CHECK(cudaFuncSetAttribute(gemv_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  int grid = num_sms < nrows ? num_sms : nrows;
  ChargeEpilogue e;
  if (ep) e = *ep;
  else { memset(&e, 0, sizeof(e)); }
  gemv_tma_kernel<<<grid, THREADS, smem, s>>>(S, pitch, nrows, ncols_pad, b, out, e);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_update_charge(cudaStream_t s, const ChargeEpilogue &ep) {
  int grid = (ep.n + UC_THREADS * 4 - 1) / (UC_THREADS * 4);
  grid = grid < 1 ? 1 : (grid > 128 ? 128 : grid);
  update_charge_kernel<<<grid, UC_THREADS, 0, s>>>(ep);
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
