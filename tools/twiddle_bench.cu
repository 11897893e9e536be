// twiddle_bench.cu -- design experiment: the inter-pass twiddles w^1 .. w^15 of a thread are the same for every transform it
// ever does (they depend on the thread index only).  fft_core.cuh rebuilds them per pass from one table load by 14 complex
// products (112 of the ~740 FP64 instructions of a 4096-point transform).  Does it pay to keep them instead -- pass 3's
// fifteen in the thread's TMEM lane, pass 2's (16 distinct sets per CTA) in a 3.8 KB shared-memory table?
//   V0 block_fft_split (product chain in both passes)          V1 pass 2 from shared memory, pass 3 from TMEM
//   V2 pass 3 from TMEM only                                   V3 pass 2 from shared memory only
// 256-thread CTAs, two per SM, split re/im exchange: the regime of the row kernel (K2).
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -I geophysicalflows.jl_b200/csrc tools/twiddle_bench.cu -o tools/twiddle_bench
#include <cstdio>
#include <cmath>
#include <vector>
#include "fft_core.cuh"
using namespace gfb;

// pass with Ns > 1 and R = 16, M = 1 (N = 4096): twiddles handed in
template <int DIR, int Ns, bool LAST>
__device__ __forceinline__ void pass_given(cplx (&v)[PT], const int j, double* __restrict__ xb, const cplx (&w)[16]) {
  constexpr int N = 4096, T = 256, R = 16;
  cplx x[R];
#pragma unroll
  for (int s = 0; s < R; ++s) x[s] = v[s];
#pragma unroll
  for (int s = 1; s < R; ++s) x[s] = cmul(x[s], DIR == INV ? cconj(w[s]) : w[s]);
  Dft<DIR, R>::run(x);
#pragma unroll
  for (int s = 0; s < R; ++s) v[s] = x[s];
  if (!LAST) {
    const int k = j & (Ns - 1), base = (j - k) * R + k;
#pragma unroll
    for (int s = 0; s < R; ++s) xb[xpad(base + s * Ns)] = v[s].x;
    __syncthreads();
    double re[PT];
#pragma unroll
    for (int r = 0; r < PT; ++r) re[r] = xb[xpad(j + r * T)];
    __syncthreads();
#pragma unroll
    for (int s = 0; s < R; ++s) xb[xpad(base + s * Ns)] = v[s].y;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < PT; ++r) v[r] = cmake(re[r], xb[xpad(j + r * T)]);
    __syncthreads();
  }
}

template <int DIR, int VAR>
__device__ __forceinline__ void fft_var(cplx (&v)[PT], const int j, double* xb, const cplx* tw, const cplx* __restrict__ tab2, const uint32_t tm) {
  fft_pass_split<4096, DIR, 16, 1, false>(v, j, xb, tw);
  if (VAR == 1 || VAR == 3) {
    cplx w[16];
#pragma unroll
    for (int s = 1; s < 16; ++s) w[s] = tab2[(j & 15) * 16 + s];
    pass_given<DIR, 16, false>(v, j, xb, w);
  } else {
    fft_pass_split<4096, DIR, 16, 16, false>(v, j, xb, tw);
  }
  if (VAR == 1 || VAR == 2) {
    cplx w[16];
#pragma unroll
    for (int s = 1; s < 16; ++s) w[s] = tmem_load(tm + 4 * s);
    tmem_wait_ld();
    pass_given<DIR, 256, true>(v, j, xb, w);
  } else {
    fft_pass_split<4096, DIR, 16, 256, true>(v, j, xb, tw);
  }
}

template <int VAR>
__global__ void __launch_bounds__(256, 2) k_fft(const cplx* __restrict__ tw, cplx* __restrict__ io, const int reps, const int store) {
  GFB_DYN_SMEM(double, xb);
  cplx* tab2 = reinterpret_cast<cplx*>(xb + ((xbuf_elems(4096) + 1) & ~1));   // 16 x 16 complex
  uint32_t* word = reinterpret_cast<uint32_t*>(tab2 + 256);
  const int t = threadIdx.x;
  const uint32_t tbase = tmem_setup(word, 256);
  const uint32_t tm = tmem_thread_base(tbase, 128);
  // tables: pass 2 uses W_N^{16 k s}, k = j & 15; pass 3 uses W_N^{k s}, k = j & 255  (forward sign)
  {
    const int k = t >> 4, s = t & 15;
    tab2[k * 16 + s] = tw[(16 * k * s) & 4095];
    for (int s3 = 1; s3 < 16; ++s3) tmem_store(tm + 4 * s3, tw[((t & 255) * s3) & 4095]);
    tmem_wait_st();
  }
  __syncthreads();
  cplx v[PT];
  const cplx* in = io + (size_t)(blockIdx.x % 8) * 4096;
#pragma unroll
  for (int r = 0; r < PT; ++r) v[r] = in[t + 256 * r];
  for (int i = 0; i < reps; ++i) {
    fft_var<INV, VAR>(v, t, xb, tw, tab2, tm);
    fft_var<FWD, VAR>(v, t, xb, tw, tab2, tm);
#pragma unroll
    for (int r = 0; r < PT; ++r) v[r] = cscale(v[r], 1.0 / 4096.0);
  }
  if (store) {
    cplx* out = io + (size_t)(8 + blockIdx.x) * 4096;
#pragma unroll
    for (int r = 0; r < PT; ++r) out[t + 256 * r] = v[r];
  }
  tmem_release(tbase, 256);
}

template <int VAR>
static void bench(const cplx* tw, cplx* io, const char* name) {
  const size_t smem = (size_t)((xbuf_elems(4096) + 1) & ~1) * 8 + 256 * 16 + 16;
  cudaFuncSetAttribute(k_fft<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_fft<VAR>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_fft<VAR>, 256, smem);
  const int reps = 200, grid = 148 * 2;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k_fft<VAR><<<grid, 256, smem>>>(tw, io, 2, 1);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k_fft<VAR><<<grid, 256, smem>>>(tw, io, reps, 1);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  // round-trip error of CTA 0 against its input
  std::vector<cplx> a(4096), b(4096);
  cudaMemcpy(a.data(), io, 4096 * sizeof(cplx), cudaMemcpyDeviceToHost);
  cudaMemcpy(b.data(), io + 8 * 4096, 4096 * sizeof(cplx), cudaMemcpyDeviceToHost);
  double err = 0, nrm = 0;
  for (int i = 0; i < 4096; ++i) {
    err += (a[i].x - b[i].x) * (a[i].x - b[i].x) + (a[i].y - b[i].y) * (a[i].y - b[i].y);
    nrm += a[i].x * a[i].x + a[i].y * a[i].y;
  }
  printf("%-62s occupancy %d  %.3f us per 4096-FFT per SM   round-trip error after %d pairs %.2e  (%s)\n", name, occ,
         ms * 1e3 / (2.0 * reps * 2), reps, std::sqrt(err / nrm), cudaGetErrorString(cudaGetLastError()));
}

int main() {
  std::vector<cplx> tw(4096), x(8 * 4096);
  for (int i = 0; i < 4096; ++i) tw[i] = make_double2(std::cos(2 * M_PI * i / 4096.0), -std::sin(2 * M_PI * i / 4096.0));
  for (size_t i = 0; i < x.size(); ++i) x[i] = make_double2(std::sin(0.37 * i) + 0.1, std::cos(0.11 * i * i));
  cplx *dtw, *io;
  cudaMalloc(&dtw, 4096 * sizeof(cplx));
  cudaMalloc(&io, (size_t)(8 + 296) * 4096 * sizeof(cplx));
  cudaMemcpy(dtw, tw.data(), 4096 * sizeof(cplx), cudaMemcpyHostToDevice);
  cudaMemcpy(io, x.data(), x.size() * sizeof(cplx), cudaMemcpyHostToDevice);
  bench<0>(dtw, io, "V0 product chain in both passes (block_fft_split)");
  bench<1>(dtw, io, "V1 pass 2 from shared memory, pass 3 from TMEM");
  bench<2>(dtw, io, "V2 pass 3 from TMEM only");
  bench<3>(dtw, io, "V3 pass 2 from shared memory only");
  return 0;
}
