// fft_variants.cuh -- EXPERIMENT, not part of the product library: 4096-point transforms with one CTA-wide and one
// warp-local exchange, measured against fft_core.cuh's two CTA-wide exchanges by tools/fft_bench.cu on B200
// (2.16 us vs 2.09 us per transform per SM: not faster -- the transform phases are FP64-bound).
#pragma once
#include "fft_core.cuh"

namespace gfb {

// ------------------------------------------------------------------------------------------------
// 4096-point transforms with ONE CTA-wide exchange instead of two.
//
// Index algebra (n = n0 + 16 n1 + 256 n2, k = k0 + 16 k1 + 256 k2, all digits radix 16):
//   X[k] = sum_n0 W16^{n0 k2} W4096^{n0 (k0 + 16 k1)}  sum_n1 W16^{n1 k1} W256^{n1 k0}  sum_n2 W16^{n2 k0} x[n]
// Three radix-16 passes; the pass over n2 is in registers on entry (v[r] = x[j + 256 r]).  Between two passes
// the 16 threads that share the untouched digit transpose a 16 x 16 block.  If those 16 threads are 16
// consecutive lanes, the transpose needs no CTA barrier -- only __syncwarp -- and it runs inside the slots of the
// exchange buffer that only this warp reads in the CTA-wide exchange (xpad(t + 256 rho), t in the warp), skewed
// along the diagonal so that both sides are bank-conflict free.  One of the two exchanges can be made warp-local
// at the price of a permuted thread <-> element mapping on ONE side of the transform:
//
//   fft4096_nat2tr : natural in  (thread t holds x[t + 256 r])  ->  "transposed" out: X[jT(t) + 256 r],
//                    jT(t) = (t >> 4) + 16 (t & 15)              [CTA-wide exchange first, warp-local second]
//   fft4096_tr2nat : transposed in (x[jT(t) + 256 r])  ->  natural out              [warp-local first]
//
// An inverse transform followed by pointwise work and a forward transform (the row pass) uses nat2tr then
// tr2nat: the physical-space points of a thread are permuted, which pointwise products do not care about.
// Barrier contract: every CTA-wide WRITE is preceded by a __syncthreads(); CTA-wide READS touch only the reader's
// own slots; warp-local phases touch only the warp's own slots.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int jT4096(const int t) { return (t >> 4) + ((t & 15) << 4); }

// CTA-wide exchange: this thread writes v[s] to position wbase + 16 s, then reads v[r] from t + 256 r
template <bool SPLIT>
__device__ __forceinline__ void xchg_cta(cplx (&v)[PT], const int wbase, const int t, double* __restrict__ xb) {
  if constexpr (SPLIT) {
    __syncthreads();
#pragma unroll
    for (int s = 0; s < PT; ++s) xb[xpad(wbase + 16 * s)] = v[s].x;
    __syncthreads();
    double re[PT];
#pragma unroll
    for (int r = 0; r < PT; ++r) re[r] = xb[xpad(t + 256 * r)];
    __syncthreads();
#pragma unroll
    for (int s = 0; s < PT; ++s) xb[xpad(wbase + 16 * s)] = v[s].y;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < PT; ++r) v[r] = cmake(re[r], xb[xpad(t + 256 * r)]);
  } else {
    cplx* xc = reinterpret_cast<cplx*>(xb);
    __syncthreads();
#pragma unroll
    for (int s = 0; s < PT; ++s) xc[xpad(wbase + 16 * s)] = v[s];
    __syncthreads();
#pragma unroll
    for (int r = 0; r < PT; ++r) v[r] = xc[xpad(t + 256 * r)];
  }
}

// 16 x 16 transpose among the 16 lanes of a half-warp: lane a's v[b] becomes lane b's v[a].  Slot (row, col) is
// the warp-owned position xpad((t - lane16 + col) + 256 row); element (a, b) sits in row b, column (a + b) & 15.
template <bool SPLIT>
__device__ __forceinline__ void xchg_warp(cplx (&v)[PT], const int t, double* __restrict__ xb) {
  const int l16 = t & 15, t0 = t - l16;
  __syncwarp();
  if constexpr (SPLIT) {
#pragma unroll
    for (int b = 0; b < PT; ++b) xb[xpad(t0 + ((l16 + b) & 15) + 256 * b)] = v[b].x;
    __syncwarp();
    double re[PT];
#pragma unroll
    for (int a = 0; a < PT; ++a) re[a] = xb[xpad(t0 + ((a + l16) & 15) + 256 * l16)];
    __syncwarp();
#pragma unroll
    for (int b = 0; b < PT; ++b) xb[xpad(t0 + ((l16 + b) & 15) + 256 * b)] = v[b].y;
    __syncwarp();
#pragma unroll
    for (int a = 0; a < PT; ++a) v[a] = cmake(re[a], xb[xpad(t0 + ((a + l16) & 15) + 256 * l16)]);
  } else {
    cplx* xc = reinterpret_cast<cplx*>(xb);
#pragma unroll
    for (int b = 0; b < PT; ++b) xc[xpad(t0 + ((l16 + b) & 15) + 256 * b)] = v[b];
    __syncwarp();
#pragma unroll
    for (int a = 0; a < PT; ++a) v[a] = xc[xpad(t0 + ((a + l16) & 15) + 256 * l16)];
  }
  __syncwarp();
}

template <int DIR>
__device__ __forceinline__ void tw16(cplx (&v)[PT], const cplx* __restrict__ tw, const int idx) {
  cplx w1 = __ldg(&tw[idx]);
  if (DIR == INV) w1.y = -w1.y;
  twiddle_apply<16>(v, w1);
}

template <int DIR, bool SPLIT>
__device__ __forceinline__ void fft4096_nat2tr(cplx (&v)[PT], const int t, double* __restrict__ xb,
                                               const cplx* __restrict__ tw) {
  const int lo = t & 15, hi = t >> 4;
  Dft<DIR, 16>::run(v);                       // over n2; thread = (n0 = lo, n1 = hi), v[k0]
  tw16<DIR>(v, tw, 16 * hi);                  // W256^{n1 k0}
  xchg_cta<SPLIT>(v, lo + 256 * hi, t, xb);   // (n0, n1, k0) at n0 + 16 k0 + 256 n1 -> thread (n0 = lo, k0 = hi), v[n1]
  Dft<DIR, 16>::run(v);                       // over n1 -> v[k1]
  xchg_warp<SPLIT>(v, t, xb);                 // n0 <-> k1 inside the half-warp -> thread (k1 = lo, k0 = hi), v[n0]
  tw16<DIR>(v, tw, hi + 16 * lo);             // W4096^{n0 (k0 + 16 k1)}
  Dft<DIR, 16>::run(v);                       // over n0 -> v[k2] = X[(k0 + 16 k1) + 256 k2], k0 + 16 k1 = jT(t)
}

template <int DIR, bool SPLIT>
__device__ __forceinline__ void fft4096_tr2nat(cplx (&v)[PT], const int t, double* __restrict__ xb,
                                               const cplx* __restrict__ tw) {
  const int lo = t & 15, hi = t >> 4;
  Dft<DIR, 16>::run(v);                       // over n2; thread = (n0 = hi, n1 = lo), v[k0]
  tw16<DIR>(v, tw, 16 * lo);                  // W256^{n1 k0}
  xchg_warp<SPLIT>(v, t, xb);                 // n1 <-> k0 inside the half-warp -> thread (n0 = hi, k0 = lo), v[n1]
  Dft<DIR, 16>::run(v);                       // over n1 -> v[k1]
  xchg_cta<SPLIT>(v, lo + 256 * hi, t, xb);   // (n0, k0, k1) at k0 + 16 k1 + 256 n0 -> thread T = k0 + 16 k1, v[n0]
  tw16<DIR>(v, tw, t);                        // W4096^{n0 T}
  Dft<DIR, 16>::run(v);                       // over n0 -> v[k2] = X[T + 256 k2]
}

}  // namespace gfb
