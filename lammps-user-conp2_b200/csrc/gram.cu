// k-space part of the A matrix as an FP64 tensor-core Gram contraction:
//     A_ij += sum_k 2 u_k (cos_ik cos_jk + sin_ik sin_jk)
// i.e. C += Pr^T . Pa with the k-major panel Pt[2*kc][ld] written by
// ewald_panel (row k holds sqrt(2 u_k) cos / sin of every electrode atom).
// Replaces KSpaceModuleEwald::aaa_from_sincos_a / ewald_dot_ij
// (km_ewald.cpp:560-666), which evaluates the same sum pair by pair.
//
// tcgen05 has no FP64 kind; the FP64 tensor path on sm_100a is the DMMA
// instruction (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4).  CTA tile 128x128,
// BK = 16, 8 warps each owning a 64x32 sub-tile (32 DMMA accumulators),
// 3-stage cp.async pipeline.  Bound: FP64 tensor pipe; flops = 2*nrows*n*kdim.
#include "common.cuh"

#include <algorithm>

namespace conp {

namespace {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int LDS_ = BM + 4;  // k-row stride in doubles: 4-double pad => conflict-free fragment loads
constexpr int STAGES = 3;
constexpr int THREADS = 256;

struct GramSmem {
  double a[STAGES][BK][LDS_];
  double b[STAGES][BK][LDS_];
};

__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// C[i][j] (+)= sum_k A[k][i] B[k][j] for k-major operands (row k of A holds column k of A^T): the general
// form of the Gram contraction.  panel_r / panel_a: the two operands (for the Gram both point into the same
// panel: its columns of this rank's rows, and all columns), lda / ldb doubles per k-row.  Both are padded
// with zeros to multiples of the tile, so loads need no bounds checks.  gridDim.z splits the k range:
// slice z handles a contiguous range of k-tiles and writes its own output matrix Cmat + z * slice_stride
// (partial sums, added up by the consumer in a fixed order).  accumulate == 0 overwrites C.
__global__ void __launch_bounds__(THREADS, 1)
gram_kernel(int nrows, int n, int kdim, const double *__restrict__ panel_r, const double *__restrict__ panel_a,
            size_t lda, size_t ldb, double *__restrict__ Cmat, size_t pitch, size_t slice_stride, int band_n,
            int band_row0, int accumulate) {
  // Symmetric product (the Gram): only the tiles that touch the cyclic half band (j - i) mod N in [0, N/2] of
  // their rows are computed -- the same number for every row, so row blocks of any rank are balanced; the
  // other half is the mirror image (launch_gram_band_mirror, after the row blocks have been assembled).
  if (band_n > 0) {
    const int gi0 = band_row0 + blockIdx.y * BM, gj0 = blockIdx.x * BN;
    const int dlo = gj0 - (gi0 + BM - 1), span = BM + BN - 2, H = band_n / 2;
    if (span + 1 < band_n) {
      int t = dlo % band_n;
      t += t < 0 ? band_n : 0;
      if (!(t <= H || t + span >= band_n)) return;
    }
  }
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GramSmem &sm = *reinterpret_cast<GramSmem *>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp >> 2, wn = warp & 3;  // 2 x 4 warps -> 64 x 32 per warp
  const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
  const int nk_all = kdim / BK;
  const int kt0 = (int)(((long long)nk_all * blockIdx.z) / gridDim.z);
  const int nk = (int)(((long long)nk_all * (blockIdx.z + 1)) / gridDim.z) - kt0;
  Cmat += (size_t)blockIdx.z * slice_stride;

  double acc[8][4][2];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

  auto load_stage = [&](int st, int kt) {
    // 16 k-rows x 128 doubles per operand = 1024 16-byte chunks per operand
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int chunk = tid + c * THREADS;  // 0..1023
      const int kr = chunk >> 6;            // 64 chunks per k-row
      const int col = (chunk & 63) * 2;
      const size_t krow = (size_t)((kt0 + kt) * BK + kr);
      cp_async16(&sm.a[st][kr][col], panel_r + krow * lda + i0 + col);
      cp_async16(&sm.b[st][kr][col], panel_a + krow * ldb + j0 + col);
    }
  };

  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) load_stage(s, s);
    cp_async_commit();
  }

  const int fr = lane >> 2, fk = lane & 3;  // fragment row / k within the quad
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    const int nxt = kt + STAGES - 1;
    if (nxt < nk) load_stage(nxt % STAGES, nxt);
    cp_async_commit();
    const int st = kt % STAGES;
#pragma unroll
    for (int k4 = 0; k4 < BK / 4; ++k4) {
      double af[8], bf[4];
#pragma unroll
      for (int a = 0; a < 8; ++a) af[a] = sm.a[st][k4 * 4 + fk][wm * 64 + a * 8 + fr];
#pragma unroll
      for (int b = 0; b < 4; ++b) bf[b] = sm.b[st][k4 * 4 + fk][wn * 32 + b * 8 + fr];
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) dmma(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
    }
  }
  cp_async_wait<0>();

  // C fragment: row = lane/4, cols = 2*(lane%4) + {0,1}
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int i = i0 + wm * 64 + a * 8 + fr;
    if (i >= nrows) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int j = j0 + wn * 32 + b * 8 + 2 * fk;
      double *p = Cmat + (size_t)i * pitch + j;
      if (j + 1 < n) {
        double2 v = accumulate ? *reinterpret_cast<double2 *>(p) : make_double2(0.0, 0.0);
        v.x += acc[a][b][0];
        v.y += acc[a][b][1];
        *reinterpret_cast<double2 *>(p) = v;
      } else if (j < n) {
        p[0] = (accumulate ? p[0] : 0.0) + acc[a][b][0];
      }
    }
  }
}

// C[i][j] = C[j][i] for the elements outside the cyclic half band, (j - i) mod n > n/2 (square matrix held
// whole on this GPU; their mirror images lie inside the band and were computed).  32 x 32 tiles through
// shared memory so that both reads and writes are unit stride.
__global__ void __launch_bounds__(256)
mirror_kernel(int n, double *__restrict__ C, size_t pitch) {
  __shared__ double tile[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int H = n / 2;
  for (int r = ty; r < 32; r += 8) {
    const int i = bj * 32 + r, j = bi * 32 + tx;  // transposed tile (bj, bi)
    tile[r][tx] = (i < n && j < n) ? C[(size_t)i * pitch + j] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int i = bi * 32 + r, j = bj * 32 + tx;
    if (i < n && j < n) {
      int d = j - i;
      d += d < 0 ? n : 0;
      if (d > H) C[(size_t)i * pitch + j] = tile[tx][r];
    }
  }
}

}  // namespace

int launch_gram_mirror(cudaStream_t s, int n, double *C, size_t pitch) {
  dim3 grid((n + 31) / 32, (n + 31) / 32);
  mirror_kernel<<<grid, 256, 0, s>>>(n, C, pitch);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_gram_accumulate(cudaStream_t s, int nrows, int n, int kdim, const double *panel_rows,
                           const double *panel_all, size_t ld, double *C, size_t pitch, int row0) {
  if (nrows <= 0 || n <= 0 || kdim <= 0) return 0;
  if (kdim % BK) CONP_THROW(CONP_ERR_ARG, "gram: kdim must be a multiple of %d", BK);
  ensure_dynamic_smem(gram_kernel, sizeof(GramSmem));
  dim3 grid((n + BN - 1) / BN, (nrows + BM - 1) / BM);
  gram_kernel<<<grid, THREADS, sizeof(GramSmem), s>>>(nrows, n, kdim, panel_rows, panel_all, ld, ld, C, pitch, 0, n,
                                                      row0, 1);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_tn_gemm(cudaStream_t s, int m, int n, int kdim, const double *A, size_t lda, const double *B, size_t ldb,
                   double *C, size_t pitch, int ksplit, size_t slice_stride, int accumulate) {
  if (m <= 0 || n <= 0 || kdim <= 0) return 0;
  if (kdim % BK) CONP_THROW(CONP_ERR_ARG, "tn_gemm: kdim must be a multiple of %d", BK);
  if (pitch % 2) CONP_THROW(CONP_ERR_ARG, "tn_gemm: pitch must be even");
  ensure_dynamic_smem(gram_kernel, sizeof(GramSmem));
  dim3 grid((n + BN - 1) / BN, (m + BM - 1) / BM, std::max(ksplit, 1));
  gram_kernel<<<grid, THREADS, sizeof(GramSmem), s>>>(m, n, kdim, A, B, lda, ldb, C, pitch, slice_stride, 0, 0,
                                                      accumulate);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

}  // namespace conp
