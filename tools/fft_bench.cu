// fft_bench.cu -- design experiment: what does one 4096-point FP64 transform cost per SM in the FFT-only
// regime (registers + shared memory, no global traffic), for the exchange variants of fft_core.cuh?
//   V0 split exchange, two CTA-wide exchanges (block_fft_split, the round-1 kernels)
//   V1 un-split, two CTA-wide exchanges (block_fft)
//   V2 split, one CTA-wide + one warp-local exchange (fft4096_nat2tr / tr2nat)
//   V3 un-split, one CTA-wide + one warp-local exchange
// Also checks V2/V3 against V0 element by element.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -I geophysicalflows.jl_b200/csrc -I tools tools/fft_bench.cu -o tools/fft_bench
#include <cstdio>
#include <vector>
#include "fft_variants.cuh"
using namespace gfb;

template <int VAR>
__global__ void __launch_bounds__(256, 2) k_fft(const cplx* __restrict__ tw, cplx* __restrict__ io, const int reps, const int store) {
  GFB_DYN_SMEM(double, xb);
  const int t = threadIdx.x;
  cplx v[PT];
  const cplx* in = io + (size_t)(blockIdx.x % 8) * 4096;
#pragma unroll
  for (int r = 0; r < PT; ++r) v[r] = in[t + 256 * r];
  for (int i = 0; i < reps; ++i) {
    if (VAR == 0) {
      block_fft_split<4096, INV>(v, t, xb, tw);
      block_fft_split<4096, FWD>(v, t, xb, tw);
    } else if (VAR == 1) {
      block_fft<4096, INV>(v, t, reinterpret_cast<cplx*>(xb), tw);
      block_fft<4096, FWD>(v, t, reinterpret_cast<cplx*>(xb), tw);
    } else if (VAR == 2) {
      fft4096_nat2tr<INV, true>(v, t, xb, tw);
      fft4096_tr2nat<FWD, true>(v, t, xb, tw);
    } else {
      fft4096_nat2tr<INV, false>(v, t, xb, tw);
      fft4096_tr2nat<FWD, false>(v, t, xb, tw);
    }
#pragma unroll
    for (int r = 0; r < PT; ++r) v[r] = cscale(v[r], 1.0 / 4096.0);
  }
  if (store) {
    cplx* out = io + (size_t)(8 + blockIdx.x) * 4096;
#pragma unroll
    for (int r = 0; r < PT; ++r) out[t + 256 * r] = v[r];
  }
}

// single inverse transform, result written in NATURAL order whatever the variant's thread mapping
template <int VAR>
__global__ void __launch_bounds__(256, 2) k_one(const cplx* __restrict__ tw, const cplx* __restrict__ in, cplx* __restrict__ out) {
  GFB_DYN_SMEM(double, xb);
  const int t = threadIdx.x;
  cplx v[PT];
#pragma unroll
  for (int r = 0; r < PT; ++r) v[r] = in[t + 256 * r];
  int j = t;
  if (VAR == 0) block_fft_split<4096, INV>(v, t, xb, tw);
  else if (VAR == 2) { fft4096_nat2tr<INV, true>(v, t, xb, tw); j = jT4096(t); }
  else if (VAR == 3) { fft4096_nat2tr<INV, false>(v, t, xb, tw); j = jT4096(t); }
  else if (VAR == 4) {  // forward through tr2nat: feed it the transposed mapping
#pragma unroll
    for (int r = 0; r < PT; ++r) v[r] = in[jT4096(t) + 256 * r];
    fft4096_tr2nat<FWD, true>(v, t, xb, tw);
  } else if (VAR == 5) {
#pragma unroll
    for (int r = 0; r < PT; ++r) v[r] = in[jT4096(t) + 256 * r];
    fft4096_tr2nat<FWD, false>(v, t, xb, tw);
  } else if (VAR == 6) block_fft_split<4096, FWD>(v, t, xb, tw);
#pragma unroll
  for (int r = 0; r < PT; ++r) out[j + 256 * r] = v[r];
}

template <int VAR>
static void bench(const cplx* tw, cplx* io, size_t smem, int ctas_per_sm) {
  cudaFuncSetAttribute(k_fft<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_fft<VAR>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_fft<VAR>, 256, smem);
  const int reps = 200, grid = 148 * ctas_per_sm;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k_fft<VAR><<<grid, 256, smem>>>(tw, io, 10, 0);
  cudaEventRecord(e0);
  k_fft<VAR><<<grid, 256, smem>>>(tw, io, reps, 0);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double per_sm_us = ms * 1e3 / (2.0 * reps * ctas_per_sm);
  printf("V%d smem %6zu B occupancy %d CTA/SM, %d CTAs/SM launched: %.3f us per 4096-FFT per SM  (%s)\n", VAR, smem, occ,
         ctas_per_sm, per_sm_us, cudaGetErrorString(cudaGetLastError()));
}

template <int VAR>
static std::vector<cplx> one(const cplx* tw, const cplx* in, cplx* out, size_t smem) {
  cudaFuncSetAttribute(k_one<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_one<VAR><<<1, 256, smem>>>(tw, in, out);
  std::vector<cplx> h(4096);
  cudaMemcpy(h.data(), out, 4096 * 16, cudaMemcpyDeviceToHost);
  return h;
}

static double maxdiff(const std::vector<cplx>& a, const std::vector<cplx>& b) {
  double m = 0, s = 0;
  for (size_t i = 0; i < a.size(); ++i) {
    m = fmax(m, fmax(fabs(a[i].x - b[i].x), fabs(a[i].y - b[i].y)));
    s = fmax(s, fmax(fabs(a[i].x), fabs(a[i].y)));
  }
  return m / s;
}

int main() {
  const int n = 4096;
  cplx *tw, *io;
  cudaMalloc(&tw, n * 16);
  cudaMalloc(&io, (size_t)(8 + 148 * 2) * n * 16);
  std::vector<cplx> h(n), x(8 * n);
  for (int i = 0; i < n; ++i) h[i] = make_double2(cos(-2 * M_PI * i / n), sin(-2 * M_PI * i / n));
  for (int i = 0; i < 8 * n; ++i) x[i] = make_double2(sin(0.37 * i) + 0.1 * (i % 7), cos(0.11 * i * i));
  cudaMemcpy(tw, h.data(), n * 16, cudaMemcpyHostToDevice);
  cudaMemcpy(io, x.data(), 8 * n * 16, cudaMemcpyHostToDevice);
  const size_t split = (size_t)xbuf_elems(n) * 8 + 64, full = (size_t)xbuf_elems(n) * 16 + 64;
  cplx* out = io + 8 * n;
  auto r0 = one<0>(tw, io, out, split), r2 = one<2>(tw, io, out, split), r3 = one<3>(tw, io, out, full);
  auto f0 = one<6>(tw, io, out, split), f4 = one<4>(tw, io, out, split), f5 = one<5>(tw, io, out, full);
  printf("nat2tr split vs block_fft_split: %.2e   unsplit: %.2e\n", maxdiff(r2, r0), maxdiff(r3, r0));
  printf("tr2nat split vs block_fft_split: %.2e   unsplit: %.2e\n", maxdiff(f4, f0), maxdiff(f5, f0));
  for (int c = 1; c <= 2; ++c) {
    bench<0>(tw, io, split, c);
    bench<1>(tw, io, full, c);
    bench<2>(tw, io, split, c);
    bench<3>(tw, io, full, c);
  }
  // the same with the shared-memory footprint of the real kernels (stash included): 2 CTAs/SM need <= ~113 KB each
  bench<0>(tw, io, split + 65536, 2);
  bench<2>(tw, io, split + 65536, 2);
  bench<3>(tw, io, full + 44 * 1024, 2);
  printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
