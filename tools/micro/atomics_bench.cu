// Microbenchmark: L2 atomic throughput by operand type with the PPPM spread's access pattern
// (each thread adds 5 consecutive elements of a row; rows of neighbouring threads are nearby).
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#include <random>
#include <array>
#include <algorithm>

template <typename T> __device__ void add(T* p, T v) { atomicAdd(p, v); }

template <typename T>
__global__ void k_spread(T* brick, const int* base, int n, int nx, int ny, int nz) {
  long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int j = gid / 25; if (j >= n) return;
  int nm = gid % 25; int nn = nm / 5, m = nm % 5;
  int bx = base[3*j], by = base[3*j+1], bz = base[3*j+2];
  int z = (bz + nn) % nz, y = (by + m) % ny;
  T* row = brick + ((size_t)z * ny + y) * nx;
  int x = bx;
  for (int l = 0; l < 5; ++l) { add<T>(row + x, (T)1); x = (x + 1 == nx) ? 0 : x + 1; }
}

template <typename T> float run(const std::vector<int>& hb, int n, int nx, int ny, int nz) {
  T* d; int* db; size_t g = (size_t)nx*ny*nz;
  cudaMalloc(&d, g*sizeof(T)); cudaMemset(d, 0, g*sizeof(T));
  cudaMalloc(&db, hb.size()*sizeof(int)); cudaMemcpy(db, hb.data(), hb.size()*sizeof(int), cudaMemcpyHostToDevice);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  long long th = (long long)n*25; int grid = (th + 255)/256;
  for (int i = 0; i < 3; ++i) k_spread<T><<<grid,256>>>(d, db, n, nx, ny, nz);
  cudaEventRecord(a);
  for (int i = 0; i < 20; ++i) k_spread<T><<<grid,256>>>(d, db, n, nx, ny, nz);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  cudaFree(d); cudaFree(db);
  return ms/20*1000;
}

int main() {
  const int nx = 64, ny = 108, nz = 342, n = 100000;
  std::mt19937 rng(1);
  std::vector<int> hb(3*n);
  // cell-sorted-like order: sort by (z/6, y/6, x/6)
  std::vector<std::array<int,3>> a(n);
  for (auto& p : a) { p[0] = rng()%nx; p[1] = rng()%ny; p[2] = 8 + rng()%(nz-16); }
  auto key = [&](const std::array<int,3>& p){ return ((p[2]/6)*100 + p[1]/6)*100 + p[0]/6; };
  std::vector<std::array<int,3>> s = a;
  std::sort(s.begin(), s.end(), [&](auto& u, auto& v){ return key(u) < key(v); });
  for (int pass = 0; pass < 2; ++pass) {
    auto& src = pass ? a : s;
    for (int i = 0; i < n; ++i) { hb[3*i]=src[i][0]; hb[3*i+1]=src[i][1]; hb[3*i+2]=src[i][2]; }
    printf("%s: double %.1f us | u64 %.1f us | float %.1f us | u32 %.1f us\n", pass ? "random order" : "cell-sorted ",
           run<double>(hb,n,nx,ny,nz), run<unsigned long long>(hb,n,nx,ny,nz), run<float>(hb,n,nx,ny,nz), run<unsigned int>(hb,n,nx,ny,nz));
  }
  return 0;
}
