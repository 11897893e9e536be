// cols16k.cuh -- K1 / K3 for ny = 16384: columns cut 4 x 4096 (k_col_inv16k, k_col_fwd16k, k_col_fwd16k_plain)
// (part of kernels.cuh: include that, not this file)
#pragma once
#include "layout.cuh"

namespace gfb {

// ------------------------------------------------------------------------------------------------
// Columns of 4*NS points (ny = 16384 with NS = 4096).  A 16384-point complex line is 256 KB -- the whole register
// file -- so it is cut by one radix-4 decimation-in-frequency step done in registers while loading:
//     y_q[n] = W_N^{s n q} * sum_m (s i)^{m q} x[n + NS m],      X[4k + q] = FFT_NS(y_q)[k]      (s = -1 fwd, +1 inv)
// Sub-transform q of a column pair is an independent unit on the ordinary 256-thread structure; its outputs are the
// rows 4k + q.  The four quarter-rows a unit combines are read by all four q units of the pair, which are launch
// neighbours: DRAM sees them once, L2 serves the rest.
// ------------------------------------------------------------------------------------------------
template <int DIR>
__device__ __forceinline__ cplx dif4(const cplx x0, const cplx x1, const cplx x2, const cplx x3, const int q) {
  if ((q & 1) == 0) {
    const cplx e = cadd(x0, x2), f = cadd(x1, x3);
    return q == 0 ? cadd(e, f) : csub(e, f);
  }
  const cplx o = csub(x0, x2), d = csub(x1, x3);
  const cplx h = (DIR == INV) ? cmul_i(d) : cmul_mi(d);  // (s i) (x1 - x3)
  return q == 1 ? cadd(o, h) : csub(o, h);
}

// K1 for ny = 4*NS (single-layer pre-pass or plain inverse transform): unit = (column pair, field, q)
template <int NS, class Prep>
__global__ void __launch_bounds__(GroupCfg<NS>::THREADS, GroupCfg<NS>::MINB)
    k_col_inv16k(const GridDesc g, const Prep prep, const __grid_constant__ PeerOut W, const int ldw, const long long wfs,
                 const int ncols, const int nfields, const cplx* __restrict__ tw, const cplx* __restrict__ tw2) {
  using GC = GroupCfg<NS>;
  constexpr int T = GC::T, THREADS = GC::THREADS;
  static_assert(GC::G == 1 && T % 8 == 0, "one 256-thread group per CTA");
  GFB_DYN_SMEM(double, smd);
  const int tid = threadIdx.x, j = tid;
  const int q = blockIdx.x & 3, field = (blockIdx.x >> 2) % nfields;
  const int colA = 2 * ((blockIdx.x >> 2) / nfields);
  const bool liveB = colA + 1 < ncols;
  double* xb = smd;
  cplx* X = reinterpret_cast<cplx*>(smd + GC::XB_DOUBLES);
  cplx v[PT];
#pragma unroll
  for (int r = 0; r < PT; ++r) {
    const int n = j + r * T;
    cplx a[4], b[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int l = n + NS * m;
      prep.load_pair(g, field, l, colA, liveB, toff(l, colA, prep.ld), a[m], b[m]);
    }
    const cplx w = __ldg(&tw2[(n * q) & (4 * NS - 1)]);  // inverse: conj(W_N^{n q})
    v[r] = cmul(cconj(w), dif4<INV>(a[0], a[1], a[2], a[3], q));
    X[r * THREADS + tid] = cmul(cconj(w), dif4<INV>(b[0], b[1], b[2], b[3], q));
    if ((r & 1) == 1) compiler_fence();
  }
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
    block_fft_split<NS, INV>(v, j, xb, tw);
    if (c == 0) swap_stash(v, X, tid, THREADS);
  }
  // rows y = 4 (j + r T) + q
  const int y0 = 4 * j + q, ooff = toff(y0, colA, ldw);
#pragma unroll
  for (int r = 0; r < PT; ++r) {
    const int d = (W.nranks == 1) ? 0 : (y0 + 4 * T * r) >> W.lg_rows;
    cplx* out = W.base[d] + (W.rank - d) * W.chunk + field * wfs;
    stpair(out + ooff + (long long)r * (4 * T) * ldw, X[r * THREADS + tid], v[r]);  // 64-bit: 5 layers x 16384^2 exceed 2^31 elements
  }
}

// both columns of one tiled P field, radix-4 combined for sub-transform q: column A -> v, column B -> the stash
template <int NS>
__device__ __forceinline__ void k3_load16k(cplx (&v)[PT], cplx* __restrict__ X, const cplx* __restrict__ p, const int ldp, const int colA,
                                           const bool liveA, const int q, const int j, const int tid, const cplx* __restrict__ tw2) {
  constexpr int T = NS / PT, THREADS = GroupCfg<NS>::THREADS;
#pragma unroll
  for (int r = 0; r < PT; ++r) {
    const int n = j + r * T;
    cpair x[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) x[m] = liveA ? ldpair(p + toff(n + NS * m, colA, ldp)) : zero_pair();
    const cplx w = __ldg(&tw2[(n * q) & (4 * NS - 1)]);  // forward: W_N^{n q}
    v[r] = cmul(w, dif4<FWD>(x[0].a, x[1].a, x[2].a, x[3].a, q));
    X[r * THREADS + tid] = cmul(w, dif4<FWD>(x[0].b, x[1].b, x[2].b, x[3].b, q));
    if ((r & 1) == 1) compiler_fence();
  }
}

// K3 for ny = 4*NS: unit = (column pair, q); N(l, kr) for the rows l = 4k + q, then the ordinary fused epilogue
template <int NS>
__global__ void __launch_bounds__(GroupCfg<NS>::THREADS, GroupCfg<NS>::MINB)
    k_col_fwd16k(const __grid_constant__ GridDesc g, const cplx* __restrict__ P, const int ldp, const long long pfs,
                 const __grid_constant__ EpiArgs e, const int ncols, const cplx* __restrict__ tw, const cplx* __restrict__ tw2) {
  using GC = GroupCfg<NS>;
  constexpr int T = GC::T, THREADS = GC::THREADS;
  static_assert(GC::G == 1 && T % 8 == 0, "one 256-thread group per CTA");
  GFB_DYN_SMEM(double, smd);
  const int tid = threadIdx.x, j = tid;
  const int q = blockIdx.x & 3, colA = 2 * (blockIdx.x >> 2);
  const int layer = blockIdx.y;
  const bool liveA = colA < ncols;
  double* xb = smd;
  cplx* X = reinterpret_cast<cplx*>(smd + GC::XB_DOUBLES);
  const cplx* pd = P + (2 * layer) * pfs;
  const cplx* pv = P + (2 * layer + 1) * pfs;
  cplx v[PT];
  const uint32_t tbase = tmem_setup(reinterpret_cast<uint32_t*>(X + GC::STASH), TmemCfg<NS>::COLS);
  const uint32_t tA = tmem_thread_base(tbase, TmemCfg<NS>::PER_WG), tB = tA + 64;
  const int l0 = 4 * j + q;  // this thread's rows: l0 + r * 4T
#pragma unroll 1
  for (int ci = 0; ci < 4; ++ci) {
    if ((ci & 1) == 0) k3_load16k<NS>(v, X, ci == 0 ? pv : pd, ldp, colA, liveA, q, j, tid, tw2);
    block_fft_split<NS, FWD>(v, j, xb, tw);
    if (ci < 2) {  // -i l (vq)_h of column A / B -> TMEM
      const uint32_t tdst = (ci == 0) ? tA : tB;
#pragma unroll
      for (int r = 0; r < PT; ++r) {
        const double ll = wav_l(g, l0 + r * 4 * T);
        tmem_store(tdst + 4 * r, cmake(ll * v[r].y, -ll * v[r].x));
      }
      if (ci == 0) {
#pragma unroll
        for (int r = 0; r < PT; ++r) v[r] = X[r * THREADS + tid];
      }
      tmem_wait_st();
    } else if (ci == 2) {  // N of column A -> stash, column B of the d field comes out
#pragma unroll
      for (int c0 = 0; c0 < PT; c0 += 4) {
        cplx a[4];
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) a[qq] = tmem_load(tA + 4 * (c0 + qq));
        tmem_wait_ld();
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          const cplx t = X[(c0 + qq) * THREADS + tid];
          X[(c0 + qq) * THREADS + tid] = cadd(v[c0 + qq], a[qq]);
          v[c0 + qq] = t;
        }
      }
    }
  }
  // N of column B -> TMEM slot A, then the kind-specialised epilogue on the rows l0 + r * 4T
#pragma unroll
  for (int c0 = 0; c0 < PT; c0 += 4) {
    cplx b[4];
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) b[qq] = tmem_load(tB + 4 * (c0 + qq));
    tmem_wait_ld();
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) tmem_store(tA + 4 * (c0 + qq), cadd(v[c0 + qq], b[qq]));
  }
  tmem_wait_st();
  if (e.cplxL) epilogue_dispatch<NS, true, 4 * T>(g, e, layer, colA, liveA, X, tA, l0, tid);
  else epilogue_dispatch<NS, false, 4 * T>(g, e, layer, colA, liveA, X, tA, l0, tid);
  tmem_release(tbase, TmemCfg<NS>::COLS);
}

// plain forward transform of one tiled field along 4*NS-point columns -> dense user array (the mul! seam at ny = 16384)
template <int NS>
__global__ void __launch_bounds__(GroupCfg<NS>::THREADS, GroupCfg<NS>::MINB)
    k_col_fwd16k_plain(const GridDesc g, const cplx* __restrict__ P, const int ldp, const long long pfs, cplx* __restrict__ out,
                       const int ldo, const long long los, const int ncols, const cplx* __restrict__ tw, const cplx* __restrict__ tw2) {
  using GC = GroupCfg<NS>;
  constexpr int T = GC::T, THREADS = GC::THREADS;
  GFB_DYN_SMEM(double, smd);
  const int tid = threadIdx.x, j = tid;
  const int q = blockIdx.x & 3, colA = 2 * (blockIdx.x >> 2);
  const int field = blockIdx.y;
  const bool liveA = colA < ncols;
  double* xb = smd;
  cplx* X = reinterpret_cast<cplx*>(smd + GC::XB_DOUBLES);
  cplx v[PT];
  k3_load16k<NS>(v, X, P + field * pfs, ldp, colA, liveA, q, j, tid, tw2);
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
    block_fft_split<NS, FWD>(v, j, xb, tw);
    if (c == 0) swap_stash(v, X, tid, THREADS);
  }
  if (liveA) {
#pragma unroll
    for (int r = 0; r < PT; ++r) {
      const int l = 4 * (j + r * T) + q;
      cplx* o = out + field * los + (long long)l * ldo + colA;
      o[0] = X[r * THREADS + tid];
      if (colA + 1 < g.nkl) o[1] = v[r];
    }
  }
}

// plain real transforms of rows of 2*NH points at half length (the ldiv! / mul! seams at nx = 16384); one row per CTA
// group; MODE = ROW_C2R or ROW_R2C, arguments as for k_row
template <int NH, int MODE>
__global__ void __launch_bounds__(GroupCfg<NH>::THREADS, GroupCfg<NH>::MINB)
    k_row_half(const __grid_constant__ RowArgs a, const cplx* __restrict__ tw, const cplx* __restrict__ tw2, const int kcp) {
  using RC = GroupCfg<NH>;
  constexpr int T = RC::T;
  static_assert(RC::G == 1, "one group per CTA");
  GFB_DYN_SMEM(double, smd);
  const int j = threadIdx.x;
  const int y = blockIdx.x, layer = blockIdx.y;
  double* xb = smd;
  cplx v[PT];
  if (MODE == ROW_C2R) {
    cplx* lb = reinterpret_cast<cplx*>(smd + RC::XB_DOUBLES);
    row_prefetch1<T>(lb, a.W, a.wl, a.wstride * layer + a.fsel, y, a.kc, j);
    cp_async_wait_all();
    __syncthreads();
    assemble_half<NH>(v, j, lb, a.kc, tw2);
    block_fft_split<NH, INV>(v, j, xb, tw);
    double* o = a.phys + ((long long)layer * a.ny + y) * (2 * NH);
#pragma unroll
    for (int r = 0; r < PT; ++r) *reinterpret_cast<double2*>(o + 2 * (j + r * T)) = v[r];
    return;
  }
  const double* in = a.phys + ((long long)layer * a.ny + y) * (2 * NH);
#pragma unroll
  for (int r = 0; r < PT; ++r) v[r] = *reinterpret_cast<const double2*>(in + 2 * (j + r * T));
  block_fft_split<NH, FWD>(v, j, xb, tw);
  cplx A[9], B[9];
  unpack_pair<NH>(v, j, xb, A, B);
  const long long rb = a.pl.rowbase(layer, y);
#pragma unroll
  for (int r = 0; r < 9; ++r) {
    const int k = j + r * T, km = NH - k;
    if (r < 8 || j == 0) {
      const cplx t = cmul(__ldg(&tw2[k]), B[r]);
      if (k < a.kout) *prow_ptr(a.pout, rb, k) = cadd(A[r], t);
      if (km != k && km < a.kout) *prow_ptr(a.pout, rb, km) = cconj(csub(A[r], t));
    }
  }
}

}  // namespace gfb
