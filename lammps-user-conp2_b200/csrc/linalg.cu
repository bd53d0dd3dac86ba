// Setup-time dense kernels around the inverse:
//   a_finish   diagonal / self / slab terms of the A matrix
//              km_ewald.cpp:632-634, 647-665; fix_conp.cpp:796-810
//   project    FixConp::inv_project                 fix_conp.cpp:982-1067
//   d_vector   FixConp::b_setq_cal + cond_setup     fix_conp.cpp:609-637, fix_cond.cpp:46-55
// The O(N^3) inverse itself (fix_conp.cpp:947-949, LAPACK dgetrf_/dgetri_) is
// cuSOLVER getrf/getrs, driven from ctx.cu and timed separately.
#include "common.cuh"

namespace conp {

namespace {


__global__ void __launch_bounds__(256)
a_finish_kernel(int row_begin, int row_end, int n, double *__restrict__ A, size_t pitch, double diag_kspace,
                int pairmode, double self_eta, const double *__restrict__ u0_i, const int *__restrict__ etype,
                double slab_pref, const double *__restrict__ ez) {
  const int i = row_begin + blockIdx.y;
  if (i >= row_end) return;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  double *p = A + (size_t)(i - row_begin) * pitch + j;
  double v = *p;
  if (j == i) {
    // aaa[idx1d] = ug_tot - 2 g/sqrt(pi) (km_ewald.cpp:632-634) + self (fix_conp.cpp:796-810)
    v = diag_kspace + (pairmode == CONP_PAIR_ETA ? self_eta : u0_i[etype[i]]);
  }
  if (slab_pref != 0.0) v += slab_pref * ez[i] * ez[j];  // km_ewald.cpp:647-665 after symmetrisation
  *p = v;
}

// w_i = sum_{j in subset} S_ij : one warp per row
__global__ void __launch_bounds__(256)
rowsum_kernel(int n, const double *__restrict__ S, size_t pitch, const int *__restrict__ subset,
              double *__restrict__ w) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + warp;
  if (i >= n) return;
  const double *row = S + (size_t)i * pitch;
  double acc = 0.0;
  for (int j = lane; j < n; j += 32)
    if (!subset || subset[j]) acc += row[j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) w[i] = acc;
}

// tot = sum_{i in subset} w_i (single block, deterministic)
__global__ void __launch_bounds__(1024, 1)
total_kernel(int n, const double *__restrict__ w, const int *__restrict__ subset, double *__restrict__ tot) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    if (!subset || subset[i]) acc += w[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = sh[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) *tot = v;
  }
}

// S_ij -= w_i w_j / tot  if tot^2 > 1e-8  (fix_conp.cpp:1012-1019, 1052-1059)
__global__ void __launch_bounds__(256)
rank1_kernel(int n, double *__restrict__ S, size_t pitch, const double *__restrict__ w,
             const double *__restrict__ tot, int apply) {
  const double t = *tot;
  if (!apply || !(t * t > 1e-8)) return;
  const int i = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  S[(size_t)i * pitch + j] -= w[i] * w[j] / t;
}

__global__ void __launch_bounds__(256)
d_vector_kernel(int n, const double *__restrict__ ez, const int *__restrict__ side, int ff_flag, double evscale,
                double zlo, double zprd, double *__restrict__ d, double *__restrict__ setz) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double zhalf = 0.5 * zprd + zlo;
  const int eci = side[i];
  const double z = ez[i];
  double v;
  if (ff_flag == CONP_FF_FFIELD) {
    if (eci == 1 && z < zhalf) v = -evscale * (z / zprd + 1);  // fix_conp.cpp:625-627
    else v = -evscale * z / zprd;
  } else {
    v = -0.5 * evscale * eci;  // :630
  }
  d[i] = v;
  setz[i] = v / evscale;  // fix_cond.cpp:50-52
}

__global__ void __launch_bounds__(256)
identity_kernel(int n, double *__restrict__ B, size_t ld) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n * ld) return;
  const size_t i = idx / ld, j = idx - i * ld;
  B[idx] = (i == j) ? 1.0 : 0.0;
}

// Symmetry of the (projected) inverse.  asym_kernel: out[0] = max |S_ij - S_ji|, out[1] = max |S_ij|
// (non-negative doubles order like their bit patterns, so atomicMax on the bits is exact);
// symmetrise_kernel: S_ij = S_ji = (S_ij + S_ji)/2.  32x32 tiles through shared memory so that both
// triangles are read and written with unit stride.
__global__ void __launch_bounds__(256)
asym_kernel(int n, const double *__restrict__ S, size_t ld, unsigned long long *__restrict__ out) {
  __shared__ double t[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj < bi) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int i = bj * 32 + r, j = bi * 32 + tx;  // tile (bj, bi), transposed on the way out
    t[r][tx] = (i < n && j < n) ? S[(size_t)i * ld + j] : 0.0;
  }
  __syncthreads();
  double da = 0.0, ma = 0.0;
  for (int r = ty; r < 32; r += 8) {
    const int i = bi * 32 + r, j = bj * 32 + tx;
    if (i < n && j < n) {
      const double v = S[(size_t)i * ld + j];
      da = fmax(da, fabs(v - t[tx][r]));
      ma = fmax(ma, fmax(fabs(v), fabs(t[tx][r])));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    da = fmax(da, __shfl_xor_sync(0xffffffffu, da, o));
    ma = fmax(ma, __shfl_xor_sync(0xffffffffu, ma, o));
  }
  if (tx == 0) {
    atomicMax(out, (unsigned long long)__double_as_longlong(da));
    atomicMax(out + 1, (unsigned long long)__double_as_longlong(ma));
  }
}

__global__ void __launch_bounds__(256)
symmetrise_kernel(int n, double *__restrict__ S, size_t ld) {
  __shared__ double t[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj < bi) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int i = bj * 32 + r, j = bi * 32 + tx;
    t[r][tx] = (i < n && j < n) ? S[(size_t)i * ld + j] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int i = bi * 32 + r, j = bj * 32 + tx;
    if (i < n && j < n) {
      const double v = 0.5 * (S[(size_t)i * ld + j] + t[tx][r]);
      S[(size_t)i * ld + j] = v;
      t[tx][r] = v;
    }
  }
  __syncthreads();
  if (bj == bi) return;
  for (int r = ty; r < 32; r += 8) {
    const int i = bj * 32 + r, j = bi * 32 + tx;
    if (i < n && j < n) S[(size_t)i * ld + j] = t[r][tx];
  }
}

}  // namespace

int launch_a_finish(cudaStream_t s, int row_begin, int row_end, int n, double *A_rows, size_t pitch,
                    double diag_kspace, int pairmode, double self_eta, const double *u0_i, const int *etype,
                    double slab_pref, const double *ez) {
  const int nr = row_end - row_begin;
  if (nr <= 0) return 0;
  dim3 grid((n + 255) / 256, nr);
  a_finish_kernel<<<grid, 256, 0, s>>>(row_begin, row_end, n, A_rows, pitch, diag_kspace, pairmode, self_eta,
                                       u0_i, etype, slab_pref, ez);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

// one electroneutrality projection on the full matrix; tot_out (device) keeps
// e^T S e for the "<e,e>" log line
int launch_project(cudaStream_t s, int n, double *S, size_t pitch, const int *subset, double *rowsum_tmp,
                   double *tot_out, int apply) {
  rowsum_kernel<<<(n + 7) / 8, 256, 0, s>>>(n, S, pitch, subset, rowsum_tmp);
  CUDA_CHECK(cudaGetLastError());
  total_kernel<<<1, 1024, 0, s>>>(n, rowsum_tmp, subset, tot_out);
  CUDA_CHECK(cudaGetLastError());
  dim3 grid((n + 255) / 256, n);
  rank1_kernel<<<grid, 256, 0, s>>>(n, S, pitch, rowsum_tmp, tot_out, apply);
  CUDA_CHECK(cudaGetLastError());
  return 3;
}

int launch_d_vector(cudaStream_t s, int n, const double *ez, const int *side, int ff_flag, double evscale,
                    double zlo, double zprd, double *d, double *setz) {
  d_vector_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, ez, side, ff_flag, evscale, zlo, zprd, d, setz);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

// out2 (device, 2 doubles): max |S - S^T| and max |S| of the n x n matrix
int launch_asymmetry(cudaStream_t s, int n, const double *S, size_t ld, double *out2) {
  CUDA_CHECK(cudaMemsetAsync(out2, 0, 2 * sizeof(double), s));
  if (n <= 0) return 0;
  const unsigned nb = (unsigned)((n + 31) / 32);
  asym_kernel<<<dim3(nb, nb), 256, 0, s>>>(n, S, ld, reinterpret_cast<unsigned long long *>(out2));
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_symmetrise(cudaStream_t s, int n, double *S, size_t ld) {
  if (n <= 0) return 0;
  const unsigned nb = (unsigned)((n + 31) / 32);
  symmetrise_kernel<<<dim3(nb, nb), 256, 0, s>>>(n, S, ld);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_pad_identity(cudaStream_t s, int n, double *B, size_t ld) {
  const size_t tot = (size_t)n * ld;
  identity_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(n, B, ld);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

}  // namespace conp
