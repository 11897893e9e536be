// k2_row_half.cuh -- K2 for nx = 16384: real rows as half-length complex transforms (k_row_product_half, k_row_half)
// (part of kernels.cuh: include that, not this file)
#pragma once
#include "layout.cuh"

namespace gfb {

// ------------------------------------------------------------------------------------------------
// K2 for rows of 2*NH points (nx = 16384 with NH = 8192): every real row is ONE half-length complex transform.
//
// A 2*NH-point complex line does not fit the register file, but a real row does not need one:
//   forward  z[n] = x[2n] + i x[2n+1],  Z = FFT_NH(z),  E = (Z[k] + conj Z[NH-k])/2,  O = (Z[k] - conj Z[NH-k])/(2i),
//            X[k] = E + W_N^k O   and   X[NH-k] = conj(E - W_N^k O)            (both from the SAME thread's E, O)
//   inverse  Zin[k] = (X[k] + conj X[NH-k]) + i W_N^{-k} (X[k] - conj X[NH-k]),  z = backward FFT_NH(Zin)
// (FFTW's c2r ignores Im X[0] and Im X[N/2]; so does the load below).  The two-for-one split E, O is the existing
// unpack_pair; the inverse pre-pass reads the half-spectrum row from the line buffer with a mirrored index, so
// nothing new crosses threads.  Same persistent, prefetched, TMEM-stashed structure as k_row_product, one ROW per
// unit and three phases (q, u, v): 5 transforms of NH per row = the flops of 5 transforms of 2*NH per row pair.
// A thread owns the physical points x = 2n, 2n+1 (n = j + r*T), so the products stay thread-local.
// ------------------------------------------------------------------------------------------------
template <int T>
__device__ __forceinline__ void row_prefetch1(cplx* __restrict__ lb, const cplx* __restrict__ W, const RowLayout& wl, const int f,
                                              const int row, const int kc, const int j) {
  const cplx* __restrict__ p = W + wl.rowbase(f, row);
  if (wl.nranks == 1) {
    const cplx* __restrict__ pa = p + (j >> 1) * 16 + (j & 1);
    cplx* __restrict__ da = lb + j;
#pragma unroll
    for (int i = 0; i <= PT; ++i)  // kc <= 16 T + 1
      if (j + i * T < kc) cp_async16_cg(da + i * T, pa + i * (T / 2) * 16);
    return;
  }
#pragma unroll
  for (int i = 0; i <= PT; ++i) {
    const int k = j + i * T;
    if (k < kc) cp_async16_cg(lb + k, p + wl.coloff(k));
  }
}

// Zin[k], k = j + r*T, from the half-spectrum row in the line buffer; tw2[i] = W_{2NH}^i
template <int NH>
__device__ __forceinline__ void assemble_half(cplx (&v)[PT], const int j, const cplx* __restrict__ lb, const int kc,
                                              const cplx* __restrict__ tw2) {
  constexpr int T = NH / PT;
#pragma unroll
  for (int r = 0; r < PT; ++r) {
    const int k = j + r * T, km = NH - k;
    cplx a = cmake(0.0, 0.0), b = cmake(0.0, 0.0);
    if (k < kc) a = lb[k];
    if (km < kc) b = lb[km];
    if (k == 0) {  // c2r semantics: the imaginary parts of X[0] and X[NH] are ignored
      a.y = 0.0;
      b.y = 0.0;
    }
    const cplx s = cmake(a.x + b.x, a.y - b.y), d = cmake(a.x - b.x, a.y + b.y);  // a + conj b, a - conj b
    const cplx w = cconj(__ldg(&tw2[k]));
    const cplx t = cmul(w, d);
    v[r] = cmake(s.x - t.y, s.y + t.x);  // s + i t
  }
}

// The twiddles W_{2NH}^k, k = j + r*T, a thread needs in assemble_half and again when it separates the output spectra are the
// same for every row it ever processes.  Without the Q-gradient term a thread owns 64 spare TMEM columns: they are parked
// there once per kernel (HTW = true) instead of being fetched from L2 twice per phase, a round trip that nothing hides in a
// one-CTA-per-SM kernel.  ttw: TMEM address of the thread's 16 values (forward sign).
template <int NH, bool HTW>
__device__ __forceinline__ void assemble_half_tw(cplx (&v)[PT], const int j, const cplx* __restrict__ lb, const int kc,
                                                 const cplx* __restrict__ tw2, const uint32_t ttw) {
  if (!HTW) {
    assemble_half<NH>(v, j, lb, kc, tw2);
    return;
  }
  constexpr int T = NH / PT;
#pragma unroll
  for (int r0 = 0; r0 < PT; r0 += 2) {
    cplx w4[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) w4[i] = tmem_load(ttw + 4 * (r0 + i));
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = r0 + i, k = j + r * T, km = NH - k;
      cplx a = cmake(0.0, 0.0), b = cmake(0.0, 0.0);
      if (k < kc) a = lb[k];
      if (km < kc) b = lb[km];
      if (k == 0) {
        a.y = 0.0;
        b.y = 0.0;
      }
      const cplx s = cmake(a.x + b.x, a.y - b.y), d = cmake(a.x - b.x, a.y + b.y);
      const cplx t = cmul(cconj(w4[i]), d);
      v[r] = cmake(s.x - t.y, s.y + t.x);
    }
  }
}

// 8192-point transform of the group's registers with every pass twiddle taken from on-chip copies: pass 2 (16 distinct
// values per CTA) and pass 3 (256) from the shared-memory tables t2 / t3, the radix-2 pass 4 from the thread's row twiddles
// in TMEM -- W_8192^{j + 512 i} is the square of W_16384^{j + 512 i}.
template <int DIR>
__device__ __forceinline__ void block_fft_split_8192_onchip(cplx (&v)[PT], const int j, double* __restrict__ xb, const cplx* __restrict__ tw,
                                                            const cplx* __restrict__ t2, const cplx* __restrict__ t3, const uint32_t ttw) {
  fft_pass_split<8192, DIR, 16, 1, false>(v, j, xb, tw);
  fft_pass_split_w1<8192, DIR, 16, 16, false>(v, j, xb, [&](int) { return t2[j & 15]; });
  fft_pass_split_w1<8192, DIR, 16, 256, false>(v, j, xb, [&](int) { return t3[j & 255]; });
  cplx w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = tmem_load(ttw + 4 * i);
  tmem_wait_ld();
  fft_pass_split_w1<8192, DIR, 2, 4096, true>(v, j, xb, [&](int i) { return cmake(w[i].x * w[i].x - w[i].y * w[i].y, 2.0 * w[i].x * w[i].y); });
}

template <int NH, bool WITHQ>
__global__ void __launch_bounds__(GroupCfg<NH>::THREADS, GroupCfg<NH>::MINB)
    k_row_product_half(const __grid_constant__ RowArgs a, const cplx* __restrict__ tw, const cplx* __restrict__ tw2, const int nunits,
                       const int kcp) {
  using RC = GroupCfg<NH>;
  constexpr int T = RC::T, G = RC::G;
  GFB_DYN_SMEM(double, smd);
  const int tid = threadIdx.x;
  const int gi = tid / T, j = tid % T;
  double* xb = smd + gi * RC::GSD;
  cplx* lb = reinterpret_cast<cplx*>(smd + RC::XB_DOUBLES) + 1 + (size_t)gi * kcp;  // [0] holds the TMEM base word
  constexpr bool HTW = !WITHQ;  // the second 64-column slot holds the row twiddles when the Q-gradient term does not need it
  constexpr int tmcols = (RC::THREADS / 128) * 128;
  const uint32_t tbase = tmem_setup(reinterpret_cast<uint32_t*>(smd + RC::XB_DOUBLES), tmcols);
  const uint32_t tq = tmem_thread_base(tbase, 128);
  const uint32_t tg = tq + 64;  // WITHQ: (u + U) Qx + v Qy of this row at the thread's even / odd points ; else: twiddles
  cplx* t2 = reinterpret_cast<cplx*>(smd + RC::XB_DOUBLES) + 1 + (size_t)G * kcp;  // pass-2 / pass-3 twiddle tables (HTW)
  cplx* t3 = t2 + 16;
  if (HTW) {
    static_assert(!HTW || NH == 8192, "on-chip twiddles: 8192-point half-length rows");
#pragma unroll
    for (int r = 0; r < PT; ++r) tmem_store(tg + 4 * r, __ldg(&tw2[j + r * T]));
    tmem_wait_st();
    if (tid < 16) t2[tid] = tw[tid * 32];
    if (tid < 256) t3[tid] = tw[tid * 2];
    __syncthreads();
  }
  const int stride = (int)gridDim.x * G, unit0 = (int)blockIdx.x * G + gi;
  const int niter = (nunits - (int)blockIdx.x * G + stride - 1) / stride;  // same for every group of the CTA
  cplx v[PT];
  cplx A[9], B[9];
  // phase p of unit u: field p == 0 ? q : (p == 1 ? u : v) of row (u % ny) of layer (u / ny)
  if (unit0 < nunits) row_prefetch1<T>(lb, a.W, a.wl, a.wstride * (unit0 / a.ny) + 2, unit0 % a.ny, a.kc, j);
#pragma unroll 1
  for (int it = 0; it < niter; ++it) {
    const int unit = unit0 + it * stride;
    const bool live = unit < nunits;
    const int layer = live ? unit / a.ny : 0, y = live ? unit % a.ny : 0;
#pragma unroll 1
    for (int ph = 0; ph < 3; ++ph) {
      cp_async_wait_all();
      __syncthreads();  // the line buffer is complete and visible
      if (live) assemble_half_tw<NH, HTW>(v, j, lb, a.kc, tw2, tg);
      else {
#pragma unroll
        for (int r = 0; r < PT; ++r) v[r] = cmake(0.0, 0.0);
      }
      __syncthreads();  // everyone has consumed the line buffer: refill it behind the transforms
      if (live && ph < 2) row_prefetch1<T>(lb, a.W, a.wl, a.wstride * layer + (ph == 0 ? 0 : 1), y, a.kc, j);
      else if (ph == 2 && unit + stride < nunits)
        row_prefetch1<T>(lb, a.W, a.wl, a.wstride * ((unit + stride) / a.ny) + 2, (unit + stride) % a.ny, a.kc, j);
      if constexpr (HTW) block_fft_split_8192_onchip<INV>(v, j, xb, tw, t2, t3, tg);
      else block_fft_split<NH, INV>(v, j, xb, tw);  // v[r] = (x[2n], x[2n+1]), n = j + r*T
      if (ph == 0) {  // q (+ eta) of this row -> TMEM
        if (a.eta != nullptr && live) {
          const double* ea = a.eta + (long long)y * (2 * NH);
#pragma unroll
          for (int r = 0; r < PT; ++r) {
            const double2 e2 = *reinterpret_cast<const double2*>(ea + 2 * (j + r * T));
            v[r].x += e2.x;
            v[r].y += e2.y;
          }
        }
#pragma unroll
        for (int r = 0; r < PT; ++r) tmem_store(tq + 4 * r, v[r]);
        tmem_wait_st();
        continue;
      }
      if (WITHQ) {  // u += U(y); accumulate (u + U) Qx (phase u) + v Qy (phase v) in TMEM
        const double Uy = (ph == 1 && a.U && live) ? a.U[(long long)layer * a.ny + y] : 0.0;
        const double* qg = (ph == 1 ? a.Qx : a.Qy) + ((long long)layer * a.ny + y) * (2 * NH);
#pragma unroll
        for (int c0 = 0; c0 < PT; c0 += 4) {
          cplx gs[4];
          if (ph == 2) {
#pragma unroll
            for (int q = 0; q < 4; ++q) gs[q] = tmem_load(tg + 4 * (c0 + q));
            tmem_wait_ld();
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int r = c0 + q;
            v[r].x += Uy;
            v[r].y += Uy;
            double2 g2 = make_double2(0.0, 0.0);
            if (live) g2 = *reinterpret_cast<const double2*>(qg + 2 * (j + r * T));
            if (ph == 1) gs[q] = cmake(v[r].x * g2.x, v[r].y * g2.y);
            else gs[q] = cmake(gs[q].x + v[r].x * g2.x, gs[q].y + v[r].y * g2.y);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) tmem_store(tg + 4 * (c0 + q), gs[q]);
        }
        tmem_wait_st();
        if (a.linear) {  // calcN_linearadvection!: U q in the u phase, nothing in the v phase
#pragma unroll
          for (int r = 0; r < PT; ++r) v[r] = cmake(Uy, Uy);
        }
      }
#pragma unroll
      for (int c0 = 0; c0 < PT; c0 += 4) {
        cplx q2[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) q2[q] = tmem_load(tq + 4 * (c0 + q));
        tmem_wait_ld();
#pragma unroll
        for (int q = 0; q < 4; ++q) {  // u q (or v q) at the even and the odd point
          v[c0 + q].x *= q2[q].x;
          v[c0 + q].y *= q2[q].y;
        }
      }
      if constexpr (HTW) block_fft_split_8192_onchip<FWD>(v, j, xb, tw, t2, t3, tg);
      else block_fft_split<NH, FWD>(v, j, xb, tw);
      unpack_pair<NH>(v, j, xb, A, B);  // A = E, B = O for k = j + r*T <= NH/2
      if (live) {  // `live` is uniform over the CTA (one group per CTA): the TMEM reads below are warp-uniform
        const long long rb = a.pl.rowbase(2 * layer + (ph - 1), y);
#pragma unroll
        for (int r = 0; r < 9; ++r) {
          const int k = j + r * T, km = NH - k;
          cplx wk;
          if (HTW) {
            wk = tmem_load(tg + 4 * r);
            tmem_wait_ld();
          }
          if (r < 8 || j == 0) {
            const cplx t = cmul(HTW ? wk : __ldg(&tw2[k]), B[r]);
            cplx X = cadd(A[r], t), Xm = cconj(csub(A[r], t));  // X[k], X[NH - k]
            if (ph == 1) {  // the d field carries -i kr (uq)_h
              const double kr = k * a.dkx, krm = km * a.dkx;
              X = cmake(kr * X.y, -kr * X.x);
              Xm = cmake(krm * Xm.y, -krm * Xm.x);
            }
            if (k < a.kout) *prow_ptr(a.pout, rb, k) = X;
            if (km != k && km < a.kout) *prow_ptr(a.pout, rb, km) = Xm;
          }
          compiler_fence();  // one wavenumber pair at a time: bounded registers
        }
      }
    }
    if (WITHQ) {  // d -= ((u + U) Qx + v Qy)_h : one more half-length transform of the accumulated row
#pragma unroll
      for (int c0 = 0; c0 < PT; c0 += 4) {
#pragma unroll
        for (int q = 0; q < 4; ++q) v[c0 + q] = tmem_load(tg + 4 * (c0 + q));
        tmem_wait_ld();
      }
      block_fft_split<NH, FWD>(v, j, xb, tw);
      unpack_pair<NH>(v, j, xb, A, B);
      if (live) {
        const long long rb = a.pl.rowbase(2 * layer, y);
#pragma unroll
        for (int r = 0; r < 9; ++r) {
          const int k = j + r * T, km = NH - k;
          if (r < 8 || j == 0) {
            const cplx t = cmul(__ldg(&tw2[k]), B[r]);
            if (k < a.kout) {
              cplx& e = *prow_ptr(a.pout, rb, k);  // same thread wrote it in the u phase
              e = csub(e, cadd(A[r], t));
            }
            if (km != k && km < a.kout) {
              cplx& e = *prow_ptr(a.pout, rb, km);
              e = csub(e, cconj(csub(A[r], t)));
            }
          }
        }
      }
    }
  }
  tmem_release(tbase, tmcols);
}

}  // namespace gfb
