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

// ------------------------------------------------------------------------------------------------
// Lane-paired columns of 4*NS points, decimated in TIME (the hot K1 at ny = 16384).
// k_col_inv16k above pays for its independence: each of the four q units of a column pair re-reads all 4*NS rows (and
// redoes the pre-pass arithmetic), so a launch pulls 12x the state through L2, and its outputs (rows 4k + q) are isolated
// 32-byte sectors -- the worst case for the NVLink stores of the sharded transport.  Here ONE 512-thread CTA owns a
// column pair: lane 2j + c holds column c, so the two 16-byte halves of every 32-byte sector are touched by adjacent
// lanes and merged by the LSU (no half-sector traffic, no shuffles).  The four sub-lines x[4n + q] are read once each,
// transformed, and parked -- F_0, F_1 in the thread's two TMEM slots (all 512 columns), F_2 in the shared-memory stash,
// F_3 in registers -- then combined by the radix-4 butterfly
//     X[k + NS m] = sum_q (s i)^{m q} W_N^{s k q} F_q[k]                                   (s = +1: inverse)
// whose outputs are in natural order: a warp stores 16 consecutive rows = two full 256-byte tile lines.
// Persistent (one CTA per SM walks over (column pair, field) units); the state of the unit a CTA takes NEXT is pulled
// into L2 by bulk prefetches while the current one is transformed.
// ------------------------------------------------------------------------------------------------
template <int NS>
struct Pair16kCfg {
  static constexpr int T = NS / PT;       // threads per column
  static constexpr int THREADS = 2 * T;   // lane 2j + c <-> (thread j of column c)
  // doubles per column exchange buffer, == 8 (mod 16): the 8 + 8 lanes of a half-warp fall into disjoint banks
  static constexpr int XBS = ((xbuf_elems(NS) + 15) & ~15) + 8;
  static constexpr int STASH = PT * THREADS;
  static constexpr size_t SMEM = (size_t)2 * XBS * sizeof(double) + (size_t)STASH * sizeof(cplx) + 16;
};

template <int NS, class Prep>
__global__ void __launch_bounds__(Pair16kCfg<NS>::THREADS, 1)
    k_col_inv16k_dit(const __grid_constant__ GridDesc g, const __grid_constant__ Prep prep, const __grid_constant__ PeerOut W,
                     const int ldw, const long long wfs, const int ncols, const int nfields, const int nunits, const int l2pf,
                     const cplx* __restrict__ tw, const cplx* __restrict__ tw2) {
  using PC = Pair16kCfg<NS>;
  constexpr int T = PC::T, THREADS = PC::THREADS;
  GFB_DYN_SMEM(double, smd);
  const int tid = threadIdx.x, c = tid & 1, j = tid >> 1;
  double* xb = smd + c * PC::XBS;
  cplx* X = reinterpret_cast<cplx*>(smd + 2 * PC::XBS);
  const uint32_t tbase = tmem_setup(reinterpret_cast<uint32_t*>(X + PC::STASH), 512);
  const uint32_t tA = tmem_thread_base(tbase, 128);
  cplx v[PT];
#pragma unroll 1
  for (int unit = blockIdx.x; unit < nunits; unit += gridDim.x) {
    const int field = unit % nfields, col = 2 * (unit / nfields) + c;
    const bool live = col < ncols;
    if (l2pf) {
      const int nu = unit + gridDim.x;
      if (nu < nunits && (l2pf == 1 || nu % nfields == 0)) prep.prefetch_pair(g, nu % nfields, 2 * (nu / nfields), tid, THREADS);
    }
#pragma unroll 1
    for (int q = 0; q < 4; ++q) {
      // sub-line q: x[4 n + q], n = j + r T; rows advance by 4T (a multiple of 8): the tiled offset is affine in r
      const int l0 = 4 * j + q, off0 = prep.off_one(l0, col);
#pragma unroll
      for (int r = 0; r < PT; ++r) {
        v[r] = live ? prep.load_one(g, field, l0 + 4 * T * r, col, off0 + r * 4 * T * prep.ld) : cmake(0.0, 0.0);
        if ((r & 7) == 7) compiler_fence();
      }
      block_fft_split<NS, INV>(v, j, xb, tw);
      if (q < 2) {
        const uint32_t t = tA + 64 * q;
#pragma unroll
        for (int r = 0; r < PT; ++r) tmem_store(t + 4 * r, v[r]);
        tmem_wait_st();
      } else if (q == 2) {
#pragma unroll
        for (int r = 0; r < PT; ++r) X[r * THREADS + tid] = v[r];
      }
    }
    // radix-4 combination, rows k + NS m with k = j + r T
    const int ooff = toff(j, col, ldw);
#pragma unroll
    for (int r0 = 0; r0 < PT; r0 += 2) {
      cplx f0[2], f1[2], w1[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        f0[i] = tmem_load(tA + 4 * (r0 + i));
        f1[i] = tmem_load(tA + 64 + 4 * (r0 + i));
        w1[i] = cconj(__ldg(&tw2[j + (r0 + i) * T]));  // inverse: conj(W_N^k)
      }
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int r = r0 + i;
        const cplx w2 = cmake(w1[i].x * w1[i].x - w1[i].y * w1[i].y, 2.0 * w1[i].x * w1[i].y), w3 = cmul(w2, w1[i]);
        cplx a0 = f0[i], a1 = cmul(w1[i], f1[i]), a2 = cmul(w2, X[r * THREADS + tid]), a3 = cmul(w3, v[r]);
        dft4<INV>(a0, a1, a2, a3);
        const cplx a[4] = {a0, a1, a2, a3};
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int row = j + r * T + NS * m;
          const int d = (W.nranks == 1) ? 0 : row >> W.lg_rows;
          cplx* out = W.base[d] + (W.rank - d) * W.chunk + field * wfs;
          out[ooff + (long long)(r * T + NS * m) * ldw] = a[m];  // T, NS multiples of 8: affine in the tiled layout
        }
      }
    }
  }
  tmem_release(tbase, 512);
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

// ------------------------------------------------------------------------------------------------
// Lane-paired K3 at ny = 4*NS (the hot path at ny = 16384).  Same thread arrangement as k_col_inv16k_dit: one 512-thread
// CTA per column pair, lane 2j + c = column c.  The forward transform keeps the decimation in FREQUENCY (its outputs,
// rows 4k + q, are final, so the epilogue needs no combination step), but one CTA now produces two sub-transforms
// (q, q + 2) of both fields from one read of the four quarter-rows -- they share the partial sums x0 +- x2, x1 +- x3 -- and
// the two halves (q = 0, 2 and q = 1, 3) run back to back on the same CTA: every row is pulled through L2 twice instead
// of four times, as two adjacent lanes' 16-byte halves of one sector.  Line storage per thread (16 points each):
//     registers  y^v_q      shared-memory stash  y^d_q      TMEM slot 0  y^v_{q+2}      TMEM slot 1  y^d_{q+2}
// After the four transforms N_q sits in TMEM slot 1 and N_{q+2} in slot 0, exchanged between the two lanes of a pair so
// that the even lane holds both columns of the thread's rows r = 0..7 and the odd lane those of r = 8..15: the
// kind-specialised epilogue (epi_fetch / epi_finish of k3_col_fwd.cuh) then runs unchanged on complete 32-byte sectors.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ cplx shfl_xor_cplx(const cplx a, const int m) {
  return cmake(__shfl_xor_sync(0xffffffffu, a.x, m), __shfl_xor_sync(0xffffffffu, a.y, m));
}
// lo / hi = this lane's column at rows r = i and r = i + 8 -> the (A, B) pair of row r = i + 8c, written to TMEM entries
// 2i, 2i + 1 of `slot` (the even lane keeps its low rows and receives the partner's, the odd lane the high rows)
__device__ __forceinline__ void pair_exchange_store(const uint32_t slot, const int i, const cplx lo, const cplx hi, const int c) {
  const cplx keep = c ? hi : lo;
  const cplx got = shfl_xor_cplx(c ? lo : hi, 1);
  tmem_store(slot + 4 * (2 * i), c ? got : keep);
  tmem_store(slot + 4 * (2 * i + 1), c ? keep : got);
}

template <int T, int KIND, bool CL>
__device__ __forceinline__ void epilogue_lp(const GridDesc& g, const EpiArgs& e, const int layer, const int colA, const bool liveA,
                                            const uint32_t tA, const int j, const int c, const int qa, const int qb) {
  constexpr int CH = 2, H = PT / 2;
  // entry i0 (0..15) of the thread: slot 1 (N_qa pairs) for i0 < 8, slot 0 (N_qb pairs) after; row r = (i0 & 7) + 8c
  if (e.Nconst.p || e.F.p || (e.mu != 0.0 && layer == e.nlayers - 1)) {
#pragma unroll 1
    for (int i0 = 0; i0 < PT; ++i0) {
      const uint32_t at = ((i0 < H) ? tA + 64 : tA) + 8 * (i0 & (H - 1));
      cplx NA = tmem_load(at), NB = tmem_load(at + 4);
      tmem_wait_ld();
      const int l = 4 * (j + ((i0 & (H - 1)) + H * c) * T) + ((i0 < H) ? qa : qb);
      if (liveA && !(e.prune && l_aliased(g, l))) epi_addends(g, e, layer, l, colA, toff(l, colA, e.out.ld), NA, NB);
      tmem_store(at, NA);
      tmem_store(at + 4, NB);
    }
    tmem_wait_st();
  }
#pragma unroll 1
  for (int i0 = 0; i0 < PT; i0 += CH) {
    EpiIn<KIND, CL> in[CH];
    bool act[CH];
    int off[CH];
    const uint32_t slot = (i0 < H) ? tA + 64 : tA;
    const int q = (i0 < H) ? qa : qb;
#pragma unroll
    for (int u = 0; u < CH; ++u) {
      const int l = 4 * (j + (((i0 + u) & (H - 1)) + H * c) * T) + q;
      act[u] = liveA && !(e.prune && l_aliased(g, l));
      off[u] = toff(l, colA, e.out.ld);
      if (act[u]) epi_fetch<KIND, CL>(g, e, layer, l, colA, off[u], in[u]);
    }
#pragma unroll
    for (int u = 0; u < CH; ++u) {
      const uint32_t at = slot + 8 * ((i0 + u) & (H - 1));
      const cplx NA = tmem_load(at), NB = tmem_load(at + 4);
      tmem_wait_ld();
      if (act[u]) epi_finish<KIND, CL>(g, e, layer, off[u], NA, NB, in[u]);
    }
  }
}

template <int T, bool CL>
__device__ __forceinline__ void epilogue_lp_dispatch(const GridDesc& g, const EpiArgs& e, int layer, int colA, bool liveA, uint32_t tA,
                                                     int j, int c, int qa, int qb) {
  switch (e.kind) {
#define GFB_EPI_CASE(K) case K: epilogue_lp<T, K, CL>(g, e, layer, colA, liveA, tA, j, c, qa, qb); break;
    GFB_EPI_CASE(EPI_STORE)
    GFB_EPI_CASE(EPI_RK_FIRST)
    GFB_EPI_CASE(EPI_RK_MID)
    GFB_EPI_CASE(EPI_RK_LAST)
    GFB_EPI_CASE(EPI_RK_ONLY)
    GFB_EPI_CASE(EPI_ETD_1)
    GFB_EPI_CASE(EPI_ETD_2)
    GFB_EPI_CASE(EPI_ETD_3)
    GFB_EPI_CASE(EPI_ETD_4)
    GFB_EPI_CASE(EPI_AB3)
    GFB_EPI_CASE(EPI_LSRK)
#undef GFB_EPI_CASE
    default: break;
  }
}

template <int NS>
__global__ void __launch_bounds__(Pair16kCfg<NS>::THREADS, 1)
    k_col_fwd16k_lp(const __grid_constant__ GridDesc g, const cplx* __restrict__ P, const int ldp, const long long pfs,
                    const __grid_constant__ EpiArgs e, const int ncols, const int npairs, const cplx* __restrict__ tw,
                    const cplx* __restrict__ tw2) {
  using PC = Pair16kCfg<NS>;
  constexpr int T = PC::T, THREADS = PC::THREADS;
  GFB_DYN_SMEM(double, smd);
  const int tid = threadIdx.x, c = tid & 1, j = tid >> 1;
  const int layer = blockIdx.y;
  double* xb = smd + c * PC::XBS;
  cplx* X = reinterpret_cast<cplx*>(smd + 2 * PC::XBS);
  const uint32_t tbase = tmem_setup(reinterpret_cast<uint32_t*>(X + PC::STASH), 512);
  const uint32_t tA = tmem_thread_base(tbase, 128), tB = tA + 64;
  const cplx* pd = P + (2 * layer) * pfs;
  const cplx* pv = P + (2 * layer + 1) * pfs;
  cplx v[PT];
  // one CTA per (column pair, half): the two halves of a pair are launch neighbours (the quarter-rows they both read come
  // from DRAM once); not persistent, so that the epilogue at the end of the kernel carries no loop state in registers
  {
    const int unit = blockIdx.x >> 1, h = blockIdx.x & 1;
    const int colA = 2 * unit, col = colA + c;
    const bool liveA = colA < ncols;
    {
      const int qa = h, qb = h + 2;
      // load: the four quarter-rows of a field at n = j + r T, radix-4 partial sums, twiddles W_N^{n q}; one field at a
      // time (bounded registers): v -> registers + slot 0, d -> stash + slot 1
      const int off0 = toff(j, col, ldp);
#pragma unroll 1
      for (int f = 0; f < 2; ++f) {
        const cplx* p = (f == 0 ? pv : pd) + off0;
        const uint32_t slot = (f == 0) ? tA : tB;
#pragma unroll
        for (int r = 0; r < PT; ++r) {
          const int n = j + r * T;
          cplx x[4];
#pragma unroll
          for (int m = 0; m < 4; ++m) x[m] = liveA ? ldcplx(p + (long long)(r * T + m * NS) * ldp) : cmake(0.0, 0.0);
          cplx ya, yb;
          if (h == 0) {
            const cplx wb = __ldg(&tw2[(2 * n) & (4 * NS - 1)]);
            const cplx ee = cadd(x[0], x[2]), ff = cadd(x[1], x[3]);
            ya = cadd(ee, ff);
            yb = cmul(wb, csub(ee, ff));
          } else {
            const cplx wa = __ldg(&tw2[n]), wb = __ldg(&tw2[(3 * n) & (4 * NS - 1)]);
            const cplx oo = csub(x[0], x[2]), hh = cmul_mi(csub(x[1], x[3]));  // forward: (-i)(x1 - x3)
            ya = cmul(wa, cadd(oo, hh));
            yb = cmul(wb, csub(oo, hh));
          }
          tmem_store(slot + 4 * r, yb);
          if (f == 0) v[r] = ya;
          else X[r * THREADS + tid] = ya;
          if ((r & 1) == 1) compiler_fence();
        }
      }
      tmem_wait_st();
      // ONE transform call site (several inlined copies make ptxas spill): ci = 0 y^v_qa -> (-i l) FFT -> stash, y^d_qa comes
      // out; ci = 1 FFT(y^d_qa) + stash = N_qa -> stash, y^v_qb comes out of slot 0; ci = 2 (-i l) FFT -> slot 0, y^d_qb comes
      // out of slot 1; ci = 3 FFT(y^d_qb) + slot 0 = N_qb
#pragma unroll 1
      for (int ci = 0; ci < 4; ++ci) {
        block_fft_split<NS, FWD>(v, j, xb, tw);
        if (ci == 0) {
#pragma unroll
          for (int r = 0; r < PT; ++r) {
            const double ll = wav_l(g, 4 * (j + r * T) + qa);
            const cplx t = X[r * THREADS + tid];
            X[r * THREADS + tid] = cmake(ll * v[r].y, -ll * v[r].x);
            v[r] = t;
          }
        } else {
          const uint32_t src = (ci == 2) ? tB : tA;
#pragma unroll
          for (int c0 = 0; c0 < PT; c0 += 4) {
            cplx a[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = tmem_load(src + 4 * (c0 + i));
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = c0 + i;
              if (ci == 1) {
                X[r * THREADS + tid] = cadd(v[r], X[r * THREADS + tid]);  // N_qa
                v[r] = a[i];                                              // y^v_qb
              } else if (ci == 2) {
                const double ll = wav_l(g, 4 * (j + r * T) + qb);
                tmem_store(tA + 4 * r, cmake(ll * v[r].y, -ll * v[r].x));
                v[r] = a[i];                                              // y^d_qb
              } else {
                v[r] = cadd(v[r], a[i]);                                  // N_qb
              }
            }
          }
          if (ci == 2) tmem_wait_st();
        }
      }
      // both columns of a row into one lane: N_qb pairs -> slot 0, N_qa pairs -> slot 1
#pragma unroll
      for (int i = 0; i < PT / 2; ++i) pair_exchange_store(tA, i, v[i], v[PT / 2 + i], c);
#pragma unroll
      for (int i = 0; i < PT / 2; ++i) pair_exchange_store(tB, i, X[i * THREADS + tid], X[(PT / 2 + i) * THREADS + tid], c);
      tmem_wait_st();
      if (e.cplxL) epilogue_lp_dispatch<T, true>(g, e, layer, colA, liveA, tA, j, c, qa, qb);
      else epilogue_lp_dispatch<T, false>(g, e, layer, colA, liveA, tA, j, c, qa, qb);
    }
  }
  tmem_release(tbase, 512);
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
