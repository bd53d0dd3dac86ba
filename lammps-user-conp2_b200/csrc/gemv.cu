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

__global__ void __launch_bounds__(THREADS, 1)
gemv_tma_kernel(const double *__restrict__ S, size_t pitch, int nrows, int ncols_pad,
                const double *__restrict__ b, double *__restrict__ out) {
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
}

// ---------------------------------------------------------------------------
// update_charge epilogue: one block, deterministic reductions
// ---------------------------------------------------------------------------
__device__ double block_sum_1024(double v, double *sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (warp == 0) {
    t = (lane < (int)(blockDim.x >> 5)) ? sh[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) sh[32] = t;
  }
  __syncthreads();
  return sh[32];
}

__global__ void __launch_bounds__(1024, 1)
update_charge_kernel(int variant, int n, const double *__restrict__ sb, const double *__restrict__ setq,
                     const double *__restrict__ qinit, const int *__restrict__ side,
                     const double *__restrict__ setz, double totsetq, double value, int one_electrode,
                     const double *__restrict__ dipole_dev, double lz, double vmult, double *__restrict__ q_out,
                     double *__restrict__ scalar_out) {
  __shared__ double sh[33];
  double part = 0.0;
  if (variant == CONP_VARIANT_COND) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) part += setz[i] * sb[i];
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      if (side[i] == 1) part += sb[i];
  }
  const double tot = block_sum_1024(part, sh);
  double potdiff, scalar;
  if (variant == CONP_VARIANT_CONP) {  // fix_conp.cpp:1149-1159
    potdiff = value;
    scalar = potdiff * totsetq + tot;
  } else if (variant == CONP_VARIANT_CONQ) {  // fix_conq.cpp:74-80
    const double netcharge_right = -tot;
    scalar = -(value - netcharge_right) / totsetq;
    if (one_electrode) scalar += 2 * value / totsetq;
    potdiff = scalar;
  } else {  // fix_cond.cpp:101-115; dipole_dev[0] = sum q z over non-electrode atoms
    const double dipole_all = -dipole_dev[0];
    potdiff = value - dipole_all / lz;
    potdiff -= tot;
    potdiff *= vmult;
    scalar = potdiff;
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double q = sb[i] + potdiff * setq[i];
    if (qinit) q += qinit[i];
    q_out[i] = q;
  }
  if (threadIdx.x == 0) {
    scalar_out[0] = scalar;
    scalar_out[1] = potdiff;
  }
}

}  // namespace

int launch_gemv(cudaStream_t s, const double *S, size_t pitch, int nrows, int ncols_pad, const double *b,
                double *out, int num_sms) {
  if (nrows <= 0) return 0;
  static bool attr_set = false;
  const size_t smem = sizeof(Smem);
  if (!attr_set) {
    CUDA_CHECK(cudaFuncSetAttribute(gemv_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  int grid = num_sms < nrows ? num_sms : nrows;
  gemv_tma_kernel<<<grid, THREADS, smem, s>>>(S, pitch, nrows, ncols_pad, b, out);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_update_charge(cudaStream_t s, int variant, int n, const double *sb, const double *setq,
                         const double *qinit, const int *side, const double *setz, double totsetq, double value,
                         int one_electrode, const double *dipole_dev, double lz, double vmult, double *q_out,
                         double *scalar_out) {
  update_charge_kernel<<<1, 1024, 0, s>>>(variant, n, sb, setq, qinit, side, setz, totsetq, value, one_electrode,
                                          dipole_dev, lz, vmult, q_out, scalar_out);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

}  // namespace conp
