// tmem_bench.cu -- design experiment (round 2): cost of moving a thread's 16 complex values (256 B) through TMEM as
// 16 x tcgen05.{ld,st}.32x32b.x4 versus 1 x .x64, for 8 and 16 active warps per SM.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/tmem_bench tools/tmem_bench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ void st4(uint32_t a, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld4(uint32_t a) {
  uint32_t x, y, z, w;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(a) : "memory");
  return x ^ y ^ z ^ w;
}
__device__ __forceinline__ void st64(uint32_t a, const uint32_t (&r)[64]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x64.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
      "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63, %64};" ::"r"(a),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
      "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
      "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]), "r"(r[32]), "r"(r[33]),
      "r"(r[34]), "r"(r[35]), "r"(r[36]), "r"(r[37]), "r"(r[38]), "r"(r[39]), "r"(r[40]), "r"(r[41]), "r"(r[42]), "r"(r[43]), "r"(r[44]),
      "r"(r[45]), "r"(r[46]), "r"(r[47]), "r"(r[48]), "r"(r[49]), "r"(r[50]), "r"(r[51]), "r"(r[52]), "r"(r[53]), "r"(r[54]), "r"(r[55]),
      "r"(r[56]), "r"(r[57]), "r"(r[58]), "r"(r[59]), "r"(r[60]), "r"(r[61]), "r"(r[62]), "r"(r[63])
      : "memory");
}
__device__ __forceinline__ void ld64(uint32_t a, uint32_t (&r)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
      "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),
        "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]),
        "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]),
        "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(a)
      : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MODE 0: 16 x ld.x4 + wait ; 1: 1 x ld.x64 + wait ; 2: 16 x st.x4 + wait ; 3: 1 x st.x64 + wait
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(int iters, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t tbase_s;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tbase_s;
  const int warp = threadIdx.x / 32;
  const uint32_t t = base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);  // 64 columns per warp, <= 8 warps/quadrant... (warp>>2) < 4
  uint32_t r[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) r[i] = threadIdx.x + i;
  st64(t, r);
  wait_st();
  __syncthreads();
  uint32_t acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc ^= ld4(t + 4 * i);
      wait_ld();
    } else if (MODE == 1) {
      ld64(t, r);
      wait_ld();
#pragma unroll
      for (int i = 0; i < 64; ++i) acc ^= r[i];
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 16; ++i) st4(t + 4 * i, acc + it + i);
      wait_st();
    } else {
#pragma unroll
      for (int i = 0; i < 64; ++i) r[i] += it;
      st64(t, r);
      wait_st();
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc + r[5];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}

template <int MODE>
void run(const char* name, int threads) {
  long long* cyc;
  uint32_t* sink;
  cudaMalloc(&cyc, 148 * 8);
  cudaMalloc(&sink, 64);
  const int iters = 2000;
  k<MODE><<<148, threads>>>(iters, cyc, sink);
  k<MODE><<<148, threads>>>(iters, cyc, sink);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148 * (double)iters;
  printf("%-18s %3d threads/SM: %8.1f cycles per 256 B/thread round  (%.1f B/cycle/SM)  %s\n", name, threads, avg, threads * 256.0 / avg,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  for (int threads : {256, 512}) {
    run<0>("16 x ld.x4", threads);
    run<1>("1 x ld.x64", threads);
    run<2>("16 x st.x4", threads);
    run<3>("1 x st.x64", threads);
  }
  return 0;
}
