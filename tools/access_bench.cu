// access_bench.cu -- measures achievable HBM bandwidth of the memory access patterns the column / row
// passes can use at 4096^2 complex128 (design experiment, not part of the library).
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
typedef double2 cplx;
#define NY 4096
#define NK 2048   // columns used (even)
#define LD 2050

__global__ void k_copy(const cplx* __restrict__ a, cplx* __restrict__ b, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
// natural layout [y][LD]; CTA = C columns (C=2: 512 thr, lane pairs), rows j + r*256
template <int C>
__global__ void __launch_bounds__(512) k_col_nat(const cplx* __restrict__ a, cplx* __restrict__ b, int rd_nat, int wr_nat) {
  const int T = 512 / C;  // threads per column
  const int PTS = NY / T;
  int c = threadIdx.x % C, j = threadIdx.x / C;
  int col = blockIdx.x * C + c;
  cplx v[32];
#pragma unroll
  for (int r = 0; r < PTS; ++r) {
    int y = j + r * T;
    v[r] = rd_nat ? a[(size_t)y * LD + col] : a[((size_t)blockIdx.x * NY + y) * C + c];
  }
#pragma unroll
  for (int r = 0; r < PTS; ++r) {
    int y = j + r * T;
    v[r].x += 1.0;
    if (wr_nat) b[(size_t)y * LD + col] = v[r]; else b[((size_t)blockIdx.x * NY + y) * C + c] = v[r];
  }
}
// row kernel: CTA (256 thr) handles R rows; tiled layout T[kt][y][C] (C columns per tile) or natural
template <int C, int R>
__global__ void __launch_bounds__(256) k_row(const cplx* __restrict__ a, cplx* __restrict__ b, int rd_nat, int wr_nat) {
  int y0 = blockIdx.x * R;
  // each thread moves elements e = tid + i*256 over R*NK elements; element -> (kt, ry, c)
  for (int e = threadIdx.x; e < R * NK; e += 256) {
    int c = e % C, ry = (e / C) % R, kt = e / (C * R);
    int k = kt * C + c, y = y0 + ry;
    cplx v = rd_nat ? a[(size_t)y * LD + k] : a[((size_t)kt * NY + y) * C + c];
    v.x += 1.0;
    if (wr_nat) b[(size_t)y * LD + k] = v; else b[((size_t)kt * NY + y) * C + c] = v;
  }
}
template <class F>
float timeit(F f, int reps = 10) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); f();
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps;
}
int main() {
  size_t n = (size_t)NY * LD;
  cplx *a, *b;
  cudaMalloc(&a, n * sizeof(cplx) * 2); cudaMalloc(&b, n * sizeof(cplx) * 2);
  cudaMemset(a, 0, n * sizeof(cplx) * 2); cudaMemset(b, 0, n * sizeof(cplx) * 2);
  // alternate between two halves so consecutive reps do not hit L2 (each half 134 MB > L2 126MB)
  int flip = 0;
  double bytes = 2.0 * NY * NK * 16;
  auto rep = [&](const char* name, auto f) {
    float ms = timeit([&] { f(a + (flip & 1) * n, b + (flip & 1) * n); flip++; });
    printf("%-46s %7.3f ms  %7.1f GB/s\n", name, ms, bytes / ms / 1e6);
  };
  rep("copy contiguous", [&](cplx* x, cplx* y) { k_copy<<<148 * 8, 512>>>(x, y, (size_t)NY * NK); });
  rep("col C=2: read natural, write natural", [&](cplx* x, cplx* y) { k_col_nat<2><<<NK / 2, 512>>>(x, y, 1, 1); });
  rep("col C=2: read natural, write tiled", [&](cplx* x, cplx* y) { k_col_nat<2><<<NK / 2, 512>>>(x, y, 1, 0); });
  rep("col C=2: read tiled,   write natural", [&](cplx* x, cplx* y) { k_col_nat<2><<<NK / 2, 512>>>(x, y, 0, 1); });
  rep("col C=2: read tiled,   write tiled", [&](cplx* x, cplx* y) { k_col_nat<2><<<NK / 2, 512>>>(x, y, 0, 0); });
  rep("col C=1: read natural, write natural", [&](cplx* x, cplx* y) { k_col_nat<1><<<NK, 512>>>(x, y, 1, 1); });
  rep("col C=4(16pts/thr,128thr/col): nat->nat", [&](cplx* x, cplx* y) { k_col_nat<4><<<NK / 4, 512>>>(x, y, 1, 1); });
  rep("row R=2: natural -> natural", [&](cplx* x, cplx* y) { k_row<2, 2><<<NY / 2, 256>>>(x, y, 1, 1); });
  rep("row R=2: tiled C=2 (64B) -> natural", [&](cplx* x, cplx* y) { k_row<2, 2><<<NY / 2, 256>>>(x, y, 0, 1); });
  rep("row R=2: natural -> tiled C=2 (64B)", [&](cplx* x, cplx* y) { k_row<2, 2><<<NY / 2, 256>>>(x, y, 1, 0); });
  rep("row R=2: tiled C=2 -> tiled C=2", [&](cplx* x, cplx* y) { k_row<2, 2><<<NY / 2, 256>>>(x, y, 0, 0); });
  rep("row R=2: tiled C=4 (128B) -> tiled C=4", [&](cplx* x, cplx* y) { k_row<4, 2><<<NY / 2, 256>>>(x, y, 0, 0); });
  rep("row R=4: tiled C=2 (128B) -> tiled C=2", [&](cplx* x, cplx* y) { k_row<2, 4><<<NY / 4, 256>>>(x, y, 0, 0); });
  rep("row R=1: tiled C=2 (32B) -> tiled C=2", [&](cplx* x, cplx* y) { k_row<2, 1><<<NY, 256>>>(x, y, 0, 0); });
  rep("row R=2: tiled C=8 (256B) -> tiled C=8", [&](cplx* x, cplx* y) { k_row<8, 2><<<NY / 2, 256>>>(x, y, 0, 0); });
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
