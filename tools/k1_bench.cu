// k1_bench.cu -- design experiment: where does K1 (k_col_inv<4096, PrepSingle>) spend its time?
// variants: full | no FFT (memory phases only) | no global memory (FFT + smem only)
#include <cstdio>
#include <vector>
#include "kernels.cuh"
using namespace gfb;

__device__ int g_tok[256];
__device__ __forceinline__ unsigned smid() { unsigned r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
template <int NY, int MODE>  // MODE 0 full, 1 no-fft, 2 no-global, 3 full + one-CTA-per-SM compute token
__global__ void __launch_bounds__(GroupCfg<NY>::THREADS, GroupCfg<NY>::MINB)
    k_var(const GridDesc g, const PrepSingle prep, cplx* __restrict__ W, const int ldw, const long long wfs,
          const int ncols, const cplx* __restrict__ tw, int never) {
  using GC = GroupCfg<NY>;
  constexpr int T = GC::T, THREADS = GC::THREADS;
  GFB_DYN_SMEM(double, smd);
  const int tid = threadIdx.x;
  const int gi = tid / T, j = tid % T;
  const int colA = 2 * (blockIdx.x * GC::G + gi);
  const int field = blockIdx.y;
  double* xb = smd + gi * GC::GSD;
  cplx* X = reinterpret_cast<cplx*>(smd + GC::XB_DOUBLES);
  cplx v[PT];
#pragma unroll
  for (int c0 = 0; c0 < PT; c0 += 4) {
    cplx bb[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (MODE == 2) { v[c0 + q] = cmake(tid * 1e-3, c0 + q); bb[q] = cmake(q, tid); }
      else prep.load_pair(g, field, j + (c0 + q) * T, colA, true, toff(j + (c0 + q) * T, colA, prep.ld), v[c0 + q], bb[q]);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) X[(c0 + q) * THREADS + tid] = bb[q];
    compiler_fence();
  }
  if (MODE == 3) {
    if (tid == 0) { while (atomicCAS(&g_tok[smid()], 0, 1) != 0) __nanosleep(100); }
    __syncthreads();
  }
  if (MODE != 1) block_fft_split<NY, INV>(v, j, xb, tw);
  swap_stash(v, X, tid, THREADS);
  if (MODE != 1) block_fft_split<NY, INV>(v, j, xb, tw);
  if (MODE == 3) {
    __syncthreads();
    if (tid == 0) atomicExch(&g_tok[smid()], 0);
  }
  if (MODE != 2 || never) {
    cplx* out = W + field * wfs;
#pragma unroll
    for (int r = 0; r < PT; ++r) stpair(out + toff(j + r * T, colA, ldw), X[r * THREADS + tid], v[r]);
  }
}

template <int MODE>
float run(const GridDesc& g, const PrepSingle& ps, cplx* W, int ldw, long long wfs, int ncols, const cplx* tw, int threads_dummy) {
  using GC = GroupCfg<4096>;
  const size_t smem = (size_t)GC::XB_DOUBLES * sizeof(double) + (size_t)GC::STASH * sizeof(cplx);
  cudaFuncSetAttribute(k_var<4096, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_var<4096, MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  dim3 grid((ncols + 1) / 2, 3);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) k_var<4096, MODE><<<grid, GC::THREADS, smem>>>(g, ps, W, ldw, wfs, ncols, tw, 0);
  cudaEventRecord(e0);
  for (int i = 0; i < 10; ++i) k_var<4096, MODE><<<grid, GC::THREADS, smem>>>(g, ps, W, ldw, wfs, ncols, tw, 0);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("mode %d: %.3f ms per launch  (%s)\n", MODE, ms / 10, cudaGetErrorString(cudaGetLastError()));
  return ms / 10;
}

int main() {
  const int n = 4096, nkr = n / 2 + 1, ld = 2050;
  GridDesc g; g.nx = n; g.ny = n; g.nkr = nkr; g.kcut = 1365; g.la0 = 1365; g.la1 = 2731; g.dkx = 1; g.dky = 1; g.invN = 1.0 / n / n;
  cplx *sol, *W, *tw; double* G;
  cudaMalloc(&sol, (size_t)n * ld * 16); cudaMalloc(&W, (size_t)3 * n * ld * 16); cudaMalloc(&G, (size_t)n * ld * 8); cudaMalloc(&tw, n * 16);
  cudaMemset(sol, 0, (size_t)n * ld * 16); cudaMemset(G, 0, (size_t)n * ld * 8);
  std::vector<cplx> h(n); for (int i = 0; i < n; ++i) h[i] = make_double2(cos(-2 * M_PI * i / n), sin(-2 * M_PI * i / n));
  cudaMemcpy(tw, h.data(), n * 16, cudaMemcpyHostToDevice);
  PrepSingle ps; ps.sol = sol; ps.ld = ld; ps.G = G; ps.ldg = ld;
  run<0>(g, ps, W, ld, (long long)n * ld, g.kcut, tw, 0);
  run<1>(g, ps, W, ld, (long long)n * ld, g.kcut, tw, 0);
  run<2>(g, ps, W, ld, (long long)n * ld, g.kcut, tw, 0);
  run<3>(g, ps, W, ld, (long long)n * ld, g.kcut, tw, 0);
  return 0;
}
