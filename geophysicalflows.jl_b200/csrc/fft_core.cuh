// fft_core.cuh -- register-resident Stockham FFT for one thread group ("FFT group") of a CTA.
//
// Replaces (for this path) the FFTW / cuFFT plans the reference reaches through
// `grid.rfftplan` (/root/reference/src/twodnavierstokes.jl:258-270) and the layered plan of
// /root/reference/src/multilayerqg.jl:287.  Float64, power-of-two lengths 16 .. 8192.
//
// Data model.  A length-N complex FFT is executed by T = N/16 threads; thread j (0 <= j < T) holds
// 16 points in registers, `v[r]` <-> element `j + r*T` ("natural strided layout") -- on input AND on
// output.  That layout makes the global loads of the first pass and the global stores after the last
// pass coalesced for row transforms (consecutive j = consecutive addresses) and lets column
// transforms assign (column, j) to lanes freely.  The passes are radix-16 (then one radix 2/4/8
// remainder pass, placed last where it needs the fewest twiddle multiplies); between passes the
// points are exchanged through shared memory with the Stockham auto-sort index map, so no bit
// reversal is ever needed.  Twiddles: one table load (W_N^k) per butterfly per pass, the powers
// W^2k .. W^15k are formed in registers by a log-depth product chain (no sincos, no per-point loads).
#pragma once
#include "gfb_platform.h"

namespace gfb {

constexpr int FWD = -1;  // exp(-i...)  : rfft / mul!(.., plan, ..)
constexpr int INV = +1;  // exp(+i...)  : unnormalised backward, the 1/(nx*ny) is folded elsewhere

constexpr int PT = 16;  // points per thread

__host__ __device__ constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n >> 1); }

// multiply by -i (forward) or +i (inverse): the radix-4 "W_4" rotation
template <int DIR>
__device__ __forceinline__ cplx rot4(cplx a) {
  return DIR == FWD ? cmul_mi(a) : cmul_i(a);
}
// multiply by W_8 = (1 -/+ i)/sqrt2 and W_8^3 = (-1 -/+ i)/sqrt2
template <int DIR>
__device__ __forceinline__ cplx rot8_1(cplx a) {
  const double h = 0.70710678118654752440;
  return DIR == FWD ? cmake((a.x + a.y) * h, (a.y - a.x) * h) : cmake((a.x - a.y) * h, (a.y + a.x) * h);
}
template <int DIR>
__device__ __forceinline__ cplx rot8_3(cplx a) {
  const double h = 0.70710678118654752440;
  return DIR == FWD ? cmake((a.y - a.x) * h, -(a.x + a.y) * h) : cmake(-(a.x + a.y) * h, (a.x - a.y) * h);
}
// multiply by W_16^e for e = 1, 3 (forward: cos - i sin)
template <int DIR, int E>
__device__ __forceinline__ cplx rot16(cplx a) {
  const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173;  // cos, sin(pi/8)
  const double c = (E == 1) ? c1 : s1, s = (E == 1) ? s1 : c1;
  // forward: (c - i s) a ; inverse: (c + i s) a
  return DIR == FWD ? cmake(a.x * c + a.y * s, a.y * c - a.x * s) : cmake(a.x * c - a.y * s, a.y * c + a.x * s);
}

template <int DIR>
__device__ __forceinline__ void dft2(cplx& a, cplx& b) {
  cplx t = csub(a, b);
  a = cadd(a, b);
  b = t;
}

// in: a0..a3, out natural order X0..X3
template <int DIR>
__device__ __forceinline__ void dft4(cplx& a0, cplx& a1, cplx& a2, cplx& a3) {
  cplx t0 = cadd(a0, a2), t1 = csub(a0, a2);
  cplx t2 = cadd(a1, a3), t3 = rot4<DIR>(csub(a1, a3));
  a0 = cadd(t0, t2);
  a2 = csub(t0, t2);
  a1 = cadd(t1, t3);
  a3 = csub(t1, t3);
}

// Radix-R DFT on x[0..R-1] (elements picked by the caller), natural order in and out.
template <int DIR, int R>
struct Dft;

template <int DIR>
struct Dft<DIR, 2> {
  __device__ __forceinline__ static void run(cplx* x) { dft2<DIR>(x[0], x[1]); }
};
template <int DIR>
struct Dft<DIR, 4> {
  __device__ __forceinline__ static void run(cplx* x) { dft4<DIR>(x[0], x[1], x[2], x[3]); }
};
template <int DIR>
struct Dft<DIR, 8> {
  // n = 2*n1 + n2 (n1<4, n2<2), k = k1 + 4*k2
  __device__ __forceinline__ static void run(cplx* x) {
    dft4<DIR>(x[0], x[2], x[4], x[6]);  // n2 = 0 -> A[k1][0] at 2*k1
    dft4<DIR>(x[1], x[3], x[5], x[7]);  // n2 = 1 -> A[k1][1] at 2*k1+1
    x[3] = rot8_1<DIR>(x[3]);           // W_8^{k1}, k1 = 1
    x[5] = rot4<DIR>(x[5]);             // k1 = 2
    x[7] = rot8_3<DIR>(x[7]);           // k1 = 3
    dft2<DIR>(x[0], x[1]);
    dft2<DIR>(x[2], x[3]);
    dft2<DIR>(x[4], x[5]);
    dft2<DIR>(x[6], x[7]);
    // position 2*k1 + k2 holds X[k1 + 4*k2] -> natural order
    cplx y1 = x[2], y2 = x[4], y3 = x[6], y4 = x[1], y5 = x[3], y6 = x[5];
    x[1] = y1;
    x[2] = y2;
    x[3] = y3;
    x[4] = y4;
    x[5] = y5;
    x[6] = y6;
  }
};
template <int DIR>
struct Dft<DIR, 16> {
  // n = 4*n1 + n2, k = k1 + 4*k2 ; W_16^{n2*k1} between the two radix-4 layers
  __device__ __forceinline__ static void run(cplx* x) {
    dft4<DIR>(x[0], x[4], x[8], x[12]);  // n2 = 0 : A[k1][0] -> x[4*k1 + 0]
    dft4<DIR>(x[1], x[5], x[9], x[13]);
    dft4<DIR>(x[2], x[6], x[10], x[14]);
    dft4<DIR>(x[3], x[7], x[11], x[15]);
    // twiddle A[k1][n2] (at 4*k1 + n2) by W_16^{n2*k1}
    x[5] = rot16<DIR, 1>(x[5]);                  // e = 1
    x[6] = rot8_1<DIR>(x[6]);                    // e = 2
    x[7] = rot16<DIR, 3>(x[7]);                  // e = 3
    x[9] = rot8_1<DIR>(x[9]);                    // e = 2
    x[10] = rot4<DIR>(x[10]);                    // e = 4
    x[11] = rot8_3<DIR>(x[11]);                  // e = 6
    x[13] = rot16<DIR, 3>(x[13]);                // e = 3
    x[14] = rot8_3<DIR>(x[14]);                  // e = 6
    x[15] = rot4<DIR>(rot4<DIR>(rot16<DIR, 1>(x[15])));  // e = 9 = 8 + 1 -> -(W^1)
    dft4<DIR>(x[0], x[1], x[2], x[3]);
    dft4<DIR>(x[4], x[5], x[6], x[7]);
    dft4<DIR>(x[8], x[9], x[10], x[11]);
    dft4<DIR>(x[12], x[13], x[14], x[15]);
    // position 4*k1 + k2 holds X[k1 + 4*k2] -> transpose the 4x4 to natural order
    cplx t;
#define GFB_SWAP(a, b) \
  t = x[a];            \
  x[a] = x[b];         \
  x[b] = t;
    GFB_SWAP(1, 4)
    GFB_SWAP(2, 8)
    GFB_SWAP(3, 12)
    GFB_SWAP(6, 9)
    GFB_SWAP(7, 13)
    GFB_SWAP(11, 14)
#undef GFB_SWAP
  }
};

// padded shared-memory index: one pad element per 16 keeps both the strided Stockham writes
// (stride R) and the unit-stride reads conflict-free for 16-byte accesses
__host__ __device__ __forceinline__ constexpr int xpad(int p) { return p + (p >> 4); }
__host__ __device__ constexpr int xbuf_elems(int N) { return N + (N >> 4); }

// radices for length N: as many 16s as fit, then the remainder (2, 4 or 8) LAST
template <int N>
struct Plan {
  static constexpr int lg = ilog2(N);
  static constexpr int n16 = lg / 4;
  static constexpr int rem = 1 << (lg % 4);
  static constexpr int npass = n16 + (rem > 1 ? 1 : 0);
  static constexpr int T = N / PT;
  static_assert((1 << lg) == N && N >= 16 && N <= 8192, "N must be a power of two in [16, 8192]");
};

// x[s] *= w1^s for s = 1..R-1.  Powers are built from the binary decomposition of s
// (w[s] = w[s - lowbit(s)] * w[lowbit(s)], powers of two by squaring): log-depth rounding error, the same
// 14 complex products as a plain chain for R = 16, and only w1, w2, w4, w8 plus one running product alive.
template <int R>
__device__ __forceinline__ void twiddle_apply(cplx* x, const cplx w1) {
  cplx w[R];
  w[1] = w1;
  x[1] = cmul(x[1], w1);
#pragma unroll
  for (int s = 2; s < R; ++s) {
    const int low = s & (-s);
    if (low == s) {
      const cplx h = w[s / 2];
      w[s] = cmake(h.x * h.x - h.y * h.y, 2.0 * h.x * h.y);
    } else {
      w[s] = cmul(w[s - low], w[low]);
    }
    x[s] = cmul(x[s], w[s]);
  }
}

// Barrier between the exchange phases of a transform: BAR = 0 -> __syncthreads() (the whole CTA takes part in the
// transform), BAR = k > 0 -> named barrier k over the N/16 threads of the group (warp-specialised kernels: the other
// warps of the CTA are busy moving data and never join).
template <int BAR, int T>
__device__ __forceinline__ void fft_sync() {
  if (BAR == 0) __syncthreads();
  else bar_sync(BAR, T);
}

// One Stockham pass of radix R with Ns = product of earlier radices.
//   tw: W_N^i (forward sign), i < N.   xb: this group's exchange buffer (xbuf_elems(N) cplx),
// LAST pass leaves the result in registers in natural strided layout; other passes exchange.
template <int N, int DIR, int R, int Ns, bool LAST, int BAR = 0>
__device__ __forceinline__ void fft_pass(cplx (&v)[PT], const int j, cplx* __restrict__ xb,
                                         const cplx* __restrict__ tw) {
  constexpr int T = N / PT;
  constexpr int M = PT / R;  // butterflies per thread in this pass
#pragma unroll
  for (int i = 0; i < M; ++i) {
    const int b = j + i * T;
    cplx x[R];
#pragma unroll
    for (int s = 0; s < R; ++s) x[s] = v[i + s * M];
    if (Ns > 1) {
      const int k = b & (Ns - 1);
      cplx w1 = __ldg(&tw[k * (N / (Ns * R))]);
      if (DIR == INV) w1.y = -w1.y;
      // powers by a log-depth chain: w[2a] = w[a]^2, w[2a+1] = w[2a]*w1
      twiddle_apply<R>(x, w1);
    }
    Dft<DIR, R>::run(x);
    if (LAST) {
#pragma unroll
      for (int s = 0; s < R; ++s) v[i + s * M] = x[s];
    } else {
      const int k = b & (Ns - 1);
      const int base = (b - k) * R + k;
#pragma unroll
      for (int s = 0; s < R; ++s) xb[xpad(base + s * Ns)] = x[s];
    }
  }
  if (!LAST) {
    fft_sync<BAR, T>();
#pragma unroll
    for (int r = 0; r < PT; ++r) v[r] = xb[xpad(j + r * T)];
    fft_sync<BAR, T>();
  }
}

// Full length-N transform of the group's registers.  All threads of the CTA must call it together
// (it contains __syncthreads); idle groups simply transform garbage in their own buffer.
template <int N, int DIR, int BAR = 0>
__device__ __forceinline__ void block_fft(cplx (&v)[PT], const int j, cplx* __restrict__ xb,
                                          const cplx* __restrict__ tw) {
  using P = Plan<N>;
  constexpr int rem = P::rem;
  constexpr int RR = rem > 1 ? rem : 2;
  if constexpr (P::n16 == 1) {
    if constexpr (rem == 1) {
      fft_pass<N, DIR, 16, 1, true, BAR>(v, j, xb, tw);
    } else {
      fft_pass<N, DIR, 16, 1, false, BAR>(v, j, xb, tw);
      fft_pass<N, DIR, RR, 16, true, BAR>(v, j, xb, tw);
    }
  } else if constexpr (P::n16 == 2) {
    fft_pass<N, DIR, 16, 1, false, BAR>(v, j, xb, tw);
    if constexpr (rem == 1) {
      fft_pass<N, DIR, 16, 16, true, BAR>(v, j, xb, tw);
    } else {
      fft_pass<N, DIR, 16, 16, false, BAR>(v, j, xb, tw);
      fft_pass<N, DIR, RR, 256, true, BAR>(v, j, xb, tw);
    }
  } else {  // n16 == 3
    fft_pass<N, DIR, 16, 1, false, BAR>(v, j, xb, tw);
    fft_pass<N, DIR, 16, 16, false, BAR>(v, j, xb, tw);
    if constexpr (rem == 1) {
      fft_pass<N, DIR, 16, 256, true, BAR>(v, j, xb, tw);
    } else {
      fft_pass<N, DIR, 16, 256, false, BAR>(v, j, xb, tw);
      fft_pass<N, DIR, RR, 4096, true, BAR>(v, j, xb, tw);
    }
  }
}

// Register-resident twiddles (N = 16^3 = 4096): with one radix-16 butterfly per thread and pass, the twiddle base of a
// thread is the same for every transform it ever does -- W_N^{16 (j & 15)} in pass 2, W_N^{j} in pass 3 -- so it is
// loaded ONCE per kernel instead of once per pass.  (The per-pass __ldg sat on the critical path of every pass: with
// ~200 KB of the SM's 256 KB carved out as shared memory the 64 KB table does not stay in L1, and the load -- an L2
// round trip -- was the top stall of the warp-specialised kernels' FFT group.)
struct Tw4096 {
  cplx w2, w3;  // forward sign
};
__device__ __forceinline__ Tw4096 load_tw4096(const int j, const cplx* __restrict__ tw) {
  Tw4096 t;
  t.w2 = __ldg(&tw[(j & 15) * 16]);
  t.w3 = __ldg(&tw[j & 255]);
  return t;
}
template <int DIR, int Ns, bool LAST, int BAR>
__device__ __forceinline__ void fft_pass4096(cplx (&v)[PT], const int j, cplx* __restrict__ xb, cplx w1) {
  constexpr int N = 4096, T = N / PT, R = 16;
  cplx x[R];
#pragma unroll
  for (int s = 0; s < R; ++s) x[s] = v[s];
  if (Ns > 1) {
    if (DIR == INV) w1.y = -w1.y;
    twiddle_apply<R>(x, w1);
  }
  Dft<DIR, R>::run(x);
  if (LAST) {
#pragma unroll
    for (int s = 0; s < R; ++s) v[s] = x[s];
  } else {
    const int k = j & (Ns - 1);
    const int base = (j - k) * R + k;
#pragma unroll
    for (int s = 0; s < R; ++s) xb[xpad(base + s * Ns)] = x[s];
    fft_sync<BAR, T>();
#pragma unroll
    for (int r = 0; r < PT; ++r) v[r] = xb[xpad(j + r * T)];
    fft_sync<BAR, T>();
  }
}
// 4096-point transform of the group's registers, un-split exchange, twiddles from registers
template <int DIR, int BAR>
__device__ __forceinline__ void block_fft4096(cplx (&v)[PT], const int j, cplx* __restrict__ xb, const Tw4096& t) {
  fft_pass4096<DIR, 1, false, BAR>(v, j, xb, t.w2);
  fft_pass4096<DIR, 16, false, BAR>(v, j, xb, t.w2);
  fft_pass4096<DIR, 256, true, BAR>(v, j, xb, t.w3);
}

// Split (re / im) exchange variant: the exchange buffer holds N doubles (+pad) instead of N complex, so it
// costs half the shared memory (two CTAs of a 4096-point kernel fit one SM) for twice the barriers.
template <int N, int DIR, int R, int Ns, bool LAST>
__device__ __forceinline__ void fft_pass_split(cplx (&v)[PT], const int j, double* __restrict__ xb,
                                               const cplx* __restrict__ tw) {
  constexpr int T = N / PT;
  constexpr int M = PT / R;
  int base[M];
#pragma unroll
  for (int i = 0; i < M; ++i) {
    const int b = j + i * T;
    cplx x[R];
#pragma unroll
    for (int s = 0; s < R; ++s) x[s] = v[i + s * M];
    if (Ns > 1) {
      const int k = b & (Ns - 1);
      cplx w1 = __ldg(&tw[k * (N / (Ns * R))]);
      if (DIR == INV) w1.y = -w1.y;
      twiddle_apply<R>(x, w1);
    }
    Dft<DIR, R>::run(x);
#pragma unroll
    for (int s = 0; s < R; ++s) v[i + s * M] = x[s];
    const int k = b & (Ns - 1);
    base[i] = (b - k) * R + k;
  }
  if (!LAST) {
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int s = 0; s < R; ++s) xb[xpad(base[i] + s * Ns)] = v[i + s * M].x;
    __syncthreads();
    double re[PT];
#pragma unroll
    for (int r = 0; r < PT; ++r) re[r] = xb[xpad(j + r * T)];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int s = 0; s < R; ++s) xb[xpad(base[i] + s * Ns)] = v[i + s * M].y;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < PT; ++r) v[r] = cmake(re[r], xb[xpad(j + r * T)]);
    __syncthreads();
  }
}

// fft_pass_split with the butterfly twiddle handed in by a functor, w1(i) = W_N^{k (N / (Ns R))} of butterfly i (forward
// sign): kernels that run ONE CTA per SM cannot hide the L2 round trip of the table load at the head of every pass behind
// another CTA, so they keep these per-thread constants on chip (shared-memory tables, TMEM) and pass them in.
template <int N, int DIR, int R, int Ns, bool LAST, class W1>
__device__ __forceinline__ void fft_pass_split_w1(cplx (&v)[PT], const int j, double* __restrict__ xb, const W1& w1of) {
  constexpr int T = N / PT;
  constexpr int M = PT / R;
  int base[M];
#pragma unroll
  for (int i = 0; i < M; ++i) {
    const int b = j + i * T;
    cplx x[R];
#pragma unroll
    for (int s = 0; s < R; ++s) x[s] = v[i + s * M];
    cplx w1 = w1of(i);
    if (DIR == INV) w1.y = -w1.y;
    twiddle_apply<R>(x, w1);
    Dft<DIR, R>::run(x);
#pragma unroll
    for (int s = 0; s < R; ++s) v[i + s * M] = x[s];
    const int k = b & (Ns - 1);
    base[i] = (b - k) * R + k;
  }
  if (!LAST) {
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int s = 0; s < R; ++s) xb[xpad(base[i] + s * Ns)] = v[i + s * M].x;
    __syncthreads();
    double re[PT];
#pragma unroll
    for (int r = 0; r < PT; ++r) re[r] = xb[xpad(j + r * T)];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int s = 0; s < R; ++s) xb[xpad(base[i] + s * Ns)] = v[i + s * M].y;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < PT; ++r) v[r] = cmake(re[r], xb[xpad(j + r * T)]);
    __syncthreads();
  }
}

template <int N, int DIR>
__device__ __forceinline__ void block_fft_split(cplx (&v)[PT], const int j, double* __restrict__ xb,
                                                const cplx* __restrict__ tw) {
  using P = Plan<N>;
  constexpr int rem = P::rem;
  constexpr int RR = rem > 1 ? rem : 2;
  if constexpr (P::n16 == 1) {
    if constexpr (rem == 1) {
      fft_pass_split<N, DIR, 16, 1, true>(v, j, xb, tw);
    } else {
      fft_pass_split<N, DIR, 16, 1, false>(v, j, xb, tw);
      fft_pass_split<N, DIR, RR, 16, true>(v, j, xb, tw);
    }
  } else if constexpr (P::n16 == 2) {
    fft_pass_split<N, DIR, 16, 1, false>(v, j, xb, tw);
    if constexpr (rem == 1) {
      fft_pass_split<N, DIR, 16, 16, true>(v, j, xb, tw);
    } else {
      fft_pass_split<N, DIR, 16, 16, false>(v, j, xb, tw);
      fft_pass_split<N, DIR, RR, 256, true>(v, j, xb, tw);
    }
  } else {
    fft_pass_split<N, DIR, 16, 1, false>(v, j, xb, tw);
    fft_pass_split<N, DIR, 16, 16, false>(v, j, xb, tw);
    if constexpr (rem == 1) {
      fft_pass_split<N, DIR, 16, 256, true>(v, j, xb, tw);
    } else {
      fft_pass_split<N, DIR, 16, 256, false>(v, j, xb, tw);
      fft_pass_split<N, DIR, RR, 4096, true>(v, j, xb, tw);
    }
  }
}

}  // namespace gfb
