// tile_bench.cu -- design experiment: memory phases of a column-pair kernel (K1-like: read pair, write 3 pairs) and
// of a row-pair kernel (K2-like: read 3 row pairs, write 2) as a function of the tile height H of the layout
//   off(row, kp) = ((row/H)*NKP + kp)*2H + (row%H)*2      [complex elements]
#include <cuda_runtime.h>
#include <cstdio>
typedef double2 cplx;
#define NY 4096
#define NKP 1025
#define KCP 683  // live column pairs (kcut/2)
template <int H> __device__ __forceinline__ size_t off(int row, int kp) { return ((size_t)(row / H) * NKP + kp) * (2 * H) + (row % H) * 2; }

template <int H>
__global__ void __launch_bounds__(256, 2) k_col(const cplx* __restrict__ in, cplx* __restrict__ out, size_t fs) {
  extern __shared__ cplx sm[];
  const int tid = threadIdx.x, kp = blockIdx.x, f = blockIdx.y;
  cplx a[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const cplx* p = in + off<H>(tid + 256 * r, kp);
    a[r] = p[0];
    sm[r * 256 + tid] = p[1];
  }
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    cplx* o = out + f * fs + off<H>(tid + 256 * r, kp);
    a[r].x += 1.0;
    o[0] = a[r];
    o[1] = sm[r * 256 + tid];
  }
}
// row pair kernel: thread k handles columns k + 256 r (r < 8 => up to 2048); 3 input fields, 2 outputs
template <int H>
__global__ void __launch_bounds__(256, 2) k_rowp(const cplx* __restrict__ in, cplx* __restrict__ out, size_t fs) {
  const int tid = threadIdx.x, ya = 2 * blockIdx.x;
  cplx acc[6];
#pragma unroll
  for (int q = 0; q < 6; ++q) acc[q] = make_double2(0, 0);
  for (int f = 0; f < 3; ++f) {
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      const int k = tid + 256 * r;
      if (k < 2 * KCP) {
        const cplx x = in[f * fs + off<H>(ya, k >> 1) + (k & 1)];
        const cplx y = in[f * fs + off<H>(ya + 1, k >> 1) + (k & 1)];
        acc[r].x += x.x + y.y;
        acc[r].y += x.y - y.x;
      }
    }
  }
  for (int f = 0; f < 2; ++f) {
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      const int k = tid + 256 * r;
      if (k < 2 * KCP) {
        out[f * fs + off<H>(ya, k >> 1) + (k & 1)] = acc[r];
        out[f * fs + off<H>(ya + 1, k >> 1) + (k & 1)] = acc[(r + 1) % 6];
      }
    }
  }
}
template <int H>
void run(cplx* a, cplx* b, size_t fs) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaFuncSetAttribute(k_col<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k_col<H>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  float ms;
  for (int i = 0; i < 2; ++i) k_col<H><<<dim3(KCP, 3), 256, 100 * 1024>>>(a, b, fs);
  cudaEventRecord(e0);
  for (int i = 0; i < 10; ++i) k_col<H><<<dim3(KCP, 3), 256, 100 * 1024>>>(a, b, fs);
  cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
  double bytes_c = 3.0 * KCP * NY * 32 * 2;  // each of 3 field-CTAs reads the pair and writes a pair
  printf("H=%2d  col kernel %.3f ms (%.0f GB/s)", H, ms / 10, bytes_c / (ms / 10) / 1e6);
  for (int i = 0; i < 2; ++i) k_rowp<H><<<NY / 2, 256>>>(b, a, fs);
  cudaEventRecord(e0);
  for (int i = 0; i < 10; ++i) k_rowp<H><<<NY / 2, 256>>>(b, a, fs);
  cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
  double bytes_r = 5.0 * KCP * 2 * NY * 16;
  printf("   row kernel %.3f ms (%.0f GB/s)   %s\n", ms / 10, bytes_r / (ms / 10) / 1e6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  size_t fs = (size_t)NY * NKP * 2;
  cplx *a, *b;
  cudaMalloc(&a, fs * 16 * 3); cudaMalloc(&b, fs * 16 * 3);
  cudaMemset(a, 0, fs * 16 * 3); cudaMemset(b, 0, fs * 16 * 3);
  run<1>(a, b, fs); run<2>(a, b, fs); run<4>(a, b, fs); run<8>(a, b, fs); run<16>(a, b, fs); run<32>(a, b, fs);
  return 0;
}
