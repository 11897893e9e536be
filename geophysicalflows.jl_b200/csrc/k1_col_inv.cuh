// k1_col_inv.cuh -- K1: spectral pre-pass + inverse FFT along y (k_col_inv, k_col_inv_single)
// (part of kernels.cuh: include that, not this file)
#pragma once
#include "layout.cuh"

namespace gfb {

// ------------------------------------------------------------------------------------------------
// K1 pre-pass functors: value of `field` at (l, col) ready for the inverse y-FFT
// ------------------------------------------------------------------------------------------------
// load_pair(g, field, l, colA, liveB, a, b): values of `field` at (l, colA) and (l, colA+1), ready for the
// inverse y-FFT.  Sources that are ctx-owned (kAligned) are read with one 256-bit access per row.
struct PrepPlain {  // irfft of arbitrary (user, dense) spectra: field = batch index
  const cplx* in;
  int ld;  // dense (user) leading dimension; offsets are computed inside
  long long ls;
  double scale;
  __device__ __forceinline__ void load_pair(const GridDesc&, int field, int l, int colA, bool liveB, int, cplx& a, cplx& b) const {
    const cplx* p = in + field * ls + (long long)l * ld + colA;
    a = cscale(p[0], scale);
    b = liveB ? cscale(p[1], scale) : cmake(0.0, 0.0);
  }
  // one column (the lane-paired kernels: adjacent lanes hold the two columns of a pair, the LSU merges their halves)
  __device__ __forceinline__ cplx load_one(const GridDesc&, int field, int l, int col, int) const {
    return cscale(in[field * ls + (long long)l * ld + col], scale);
  }
  __device__ __forceinline__ int off_one(int, int) const { return 0; }
  __device__ __forceinline__ void prefetch_pair(const GridDesc&, int, int, int, int) const {}
};

struct PrepSingle {  // single-layer modules: psi_h = -G*sol ; uh = i l G sol ; vh = -i kr G sol ; qh = sol
  const cplx* sol;   // ctx-owned, ld even
  int ld;
  const double* G;   // ctx-owned, ld = ldg even
  int ldg;
  __device__ __forceinline__ cplx apply(const GridDesc& g, int field, double ll, double kr, double gg, cplx raw) const {
    const cplx s = cscale(raw, g.invN);
    if (field == 2) return s;
    if (field == 0) return cscale(cmul_i(s), ll * gg);
    if (field == 1) return cscale(cmul_mi(s), kr * gg);
    return cscale(s, -gg);  // field 3: psi_h (updatevars! only)
  }
  __device__ __forceinline__ void load_pair(const GridDesc& g, int field, int l, int colA, bool, int off, cplx& a, cplx& b) const {
    if (l_aliased(g, l)) {
      a = b = cmake(0.0, 0.0);
      return;
    }
    const cpair p = ldpair(sol + off);
    double2 gg = make_double2(0.0, 0.0);
    if (field != 2) gg = ldpair_d(G + off);  // G shares the state's leading dimension
    const double ll = wav_l(g, l);
    a = apply(g, field, ll, wav_kr(g, colA), gg.x, p.a);
    b = apply(g, field, ll, wav_kr(g, colA + 1), gg.y, p.b);
  }
  __device__ __forceinline__ cplx load_one(const GridDesc& g, int field, int l, int col, int off) const {
    if (l_aliased(g, l)) return cmake(0.0, 0.0);
    const cplx raw = ldcplx(sol + off);
    const double gg = (field != 2) ? lddouble(G + off) : 0.0;
    return apply(g, field, wav_l(g, l), wav_kr(g, col), gg, raw);
  }
  __device__ __forceinline__ int off_one(int l, int col) const { return toff(l, col, ld); }
  // pull the live 8-row tile lines of a column pair (256 bytes of the state, 128 bytes of G each) into L2
  __device__ __forceinline__ void prefetch_pair(const GridDesc& g, int field, int colA, int tid, int nthreads) const {
    for (int t = tid; t < g.ny / 8; t += nthreads) {
      if (8 * t >= g.la0 && 8 * t + 8 <= g.la1) continue;  // the whole tile is zeroed by dealias!
      const int off = toff(8 * t, colA, ld);
      prefetch_l2_bulk(sol + off, 256);
      if (field != 2) prefetch_l2_bulk(G + off, 128);
    }
  }
};

struct PrepMulti {  // MultiLayerQG: psi_h = S^-1 q_h ; uh = -i l psi_h ; vh = i kr psi_h ; field = nf*layer + {u,v,q[,psi]}
  const cplx* sol;     // ctx-owned, ld even
  int ld;
  long long ls;
  const double* Sinv;  // [l][kr][row][col] row-major nl x nl, dense ld = ldS
  int ldS;
  int nf;              // fields per layer: 3 (u, v, q) on the hot path, 4 (+psi) for updatevars!
  int nl;              // layers
  __device__ __forceinline__ void load_pair(const GridDesc& g, int field, int l, int colA, bool liveB, int off, cplx& a, cplx& b) const {
    if (l_aliased(g, l)) {
      a = b = cmake(0.0, 0.0);
      return;
    }
    const int layer = field / nf, what = field - nf * layer;
    if (what == 2) {
      const cpair p = ldpair(sol + layer * ls + off);
      a = cscale(p.a, g.invN);
      b = cscale(p.b, g.invN);
      return;
    }
    const double* MA = Sinv + ((long long)l * ldS + colA) * (nl * nl) + layer * nl;
    const double* MB = MA + nl * nl;
    cplx pa = cmake(0.0, 0.0), pb = cmake(0.0, 0.0);
    for (int m = 0; m < nl; ++m) {
      const cpair q = ldpair(sol + m * ls + off);
      const double ca = __ldg(&MA[m]);
      const double cb = liveB ? __ldg(&MB[m]) : 0.0;
      pa.x += ca * q.a.x;
      pa.y += ca * q.a.y;
      pb.x += cb * q.b.x;
      pb.y += cb * q.b.y;
    }
    pa = cscale(pa, g.invN);
    pb = cscale(pb, g.invN);
    if (what == 0) {
      const double ll = wav_l(g, l);
      a = cscale(cmul_mi(pa), ll);
      b = cscale(cmul_mi(pb), ll);
    } else if (what == 1) {
      a = cscale(cmul_i(pa), wav_kr(g, colA));
      b = cscale(cmul_i(pb), wav_kr(g, colA + 1));
    } else {
      a = pa;
      b = pb;
    }
  }
  __device__ __forceinline__ cplx load_one(const GridDesc& g, int field, int l, int col, int off) const {
    if (l_aliased(g, l)) return cmake(0.0, 0.0);
    const int layer = field / nf, what = field - nf * layer;
    if (what == 2) return cscale(ldcplx(sol + layer * ls + off), g.invN);
    const double* M = Sinv + ((long long)l * ldS + col) * (nl * nl) + layer * nl;
    cplx p = cmake(0.0, 0.0);
    for (int m = 0; m < nl; ++m) {
      const cplx q = ldcplx(sol + m * ls + off);
      const double cm = __ldg(&M[m]);
      p.x += cm * q.x;
      p.y += cm * q.y;
    }
    p = cscale(p, g.invN);
    if (what == 0) return cscale(cmul_mi(p), wav_l(g, l));
    if (what == 1) return cscale(cmul_i(p), wav_kr(g, col));
    return p;
  }
  __device__ __forceinline__ int off_one(int l, int col) const { return toff(l, col, ld); }
  __device__ __forceinline__ void prefetch_pair(const GridDesc&, int, int, int, int) const {}
};

// K1 (generic: MultiLayerQG pre-pass and plain transforms; the single-layer hot path is k_col_inv_single below).
// Each group takes a PAIR of adjacent kr columns: one 256-bit access per row on the spectral side and
// on the W side.  The two columns are transformed one after the other in the same registers; the other
// column waits in the thread-private stash.  Loads are issued four rows at a time (bounded registers,
// ~200 bytes in flight per thread).
template <int NY, class Prep>
__global__ void __launch_bounds__(GroupCfg<NY>::THREADS, GroupCfg<NY>::MINB)
    k_col_inv(const GridDesc g, const Prep prep, const __grid_constant__ PeerOut W, const int ldw, const long long wfs,
              const int ncols, const cplx* __restrict__ tw) {
  using GC = GroupCfg<NY>;
  constexpr int T = GC::T, THREADS = GC::THREADS;
  GFB_DYN_SMEM(double, smd);
  const int tid = threadIdx.x;
  const int gi = tid / T, j = tid % T;
  // the CTAs of one column pair (one per field) are neighbours in launch order: they run at the same time, so
  // the state columns they all read come from DRAM once and from L2 afterwards
  const int nfields = gridDim.y, lin = blockIdx.y * gridDim.x + blockIdx.x;
  const int field = lin % nfields;
  const int colA = 2 * ((lin / nfields) * GC::G + gi), colB = colA + 1;
  const bool liveA = colA < ncols, liveB = colB < ncols;
  double* xb = smd + gi * GC::GSD;
  cplx* X = reinterpret_cast<cplx*>(smd + GC::XB_DOUBLES);
  cplx v[PT];
  const int inoff = toff(j, colA, prep.ld);
#pragma unroll
  for (int c0 = 0; c0 < PT; c0 += 4) {
    cplx bb[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (liveA) prep.load_pair(g, field, j + (c0 + q) * T, colA, liveB, toff_row<T>(inoff, j, c0 + q, colA, prep.ld), v[c0 + q], bb[q]);
      else v[c0 + q] = bb[q] = cmake(0.0, 0.0);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) X[(c0 + q) * THREADS + tid] = bb[q];
    compiler_fence();
  }
  // one transform call site for both columns (two inlined copies made ptxas spill ~0.5 KB per thread)
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
    block_fft_split<NY, INV>(v, j, xb, tw);
    if (c == 0) swap_stash(v, X, tid, THREADS);
  }
  if (liveA) {
    const int ooff = toff(j, colA, ldw);
    if (W.nranks == 1) {
      cplx* out = W.base[0] + field * wfs;
#pragma unroll
      for (int r = 0; r < PT; ++r) stpair(out + toff_row<T>(ooff, j, r, colA, ldw), X[r * THREADS + tid], v[r]);
    } else {
      // row y belongs to rank d = y / (rows per rank): it lands in chunk `rank` of d's buffer, at the offset it
      // has inside chunk d of a local send buffer
#pragma unroll
      for (int r = 0; r < PT; ++r) {
        const int d = (j + r * T) >> W.lg_rows;
        cplx* out = W.base[d] + (W.rank - d) * W.chunk + field * wfs;
        stpair(out + toff_row<T>(ooff, j, r, colA, ldw), X[r * THREADS + tid], v[r]);
      }
    }
  }
}

// K1 for the single-layer modules (the hot path): persistent and software-pipelined.  A CTA group walks over
// (column pair, field) units; the raw state column it transforms NEXT (and the matching column of G) is always
// in flight behind the transform it is doing NOW:
//     FFT(A_u) | park A_u in TMEM | take B_u from the slots, refill them with A_{u+1} | FFT(B_u) | store the pair
//     | take A_{u+1}, refill with B_{u+1} | ...
// The slots are thread-private shared memory filled by cp.async (no registers, no barrier -- a thread only reads
// what it copied itself) and hold only the rows that survive the dealias mask (<= 11 of a thread's 16 rows with
// the 2/3 rule), which is what lets state + G slots fit beside the exchange buffer with two CTAs per SM.
// rows r of thread j (l = j + r*T) that survive dealias!, as a bit mask; recomputed where it is used (a handful of
// integer compares) instead of being carried across the transforms in a register
template <int T>
__device__ __forceinline__ unsigned k1_live_mask(const GridDesc& g, const int j) {
  unsigned m = 0;
#pragma unroll
  for (int r = 0; r < PT; ++r)
    if (!l_aliased(g, j + r * T)) m |= 1u << r;
  return m;
}

// start fetching column `which` (0 = A, 1 = B) of the unit's pair into the slots.  S / SG: [slot][tid]
template <int T, int THREADS>
__device__ __forceinline__ void k1_issue(cplx* __restrict__ S, double* __restrict__ SG, const PrepSingle& prep, const GridDesc& g,
                                         const int unit, const int which, const int nunits, const int nfields, const int ncols,
                                         const int j, const int tid) {
  if (unit >= nunits) return;
  const int field = unit % nfields, colA = 2 * (unit / nfields);
  if (colA + which >= ncols) return;
  const unsigned lmask = k1_live_mask<T>(g, j);
  const int off0 = toff(j, colA, prep.ld) + which;
  int slot = tid;
#pragma unroll
  for (int r = 0; r < PT; ++r)
    if (lmask >> r & 1) {
      const int off = (T % 8 == 0) ? off0 + r * T * prep.ld : toff(j + r * T, colA, prep.ld) + which;
      cp_async16_cg(&S[slot], prep.sol + off);
      if (field != 2) cp_async8(&SG[slot], prep.G + off);
      slot += THREADS;
      compiler_fence();  // addresses are formed one row at a time: v[] is live here
    }
}

// v <- the prepared column (zeros for dead units / columns / rows); branch-free: dead rows read a valid slot
template <int T, int THREADS>
__device__ __forceinline__ void k1_take(cplx (&v)[PT], const cplx* __restrict__ S, const double* __restrict__ SG, const PrepSingle& prep,
                                        const GridDesc& g, const int unit, const int which, const int nunits, const int nfields,
                                        const int ncols, const int nslots, const int j, const int tid) {
  const bool ok = unit < nunits && 2 * (unit / nfields) + which < ncols;
  const int field = ok ? unit % nfields : 2, col = ok ? 2 * (unit / nfields) + which : 0;
  const double kr = wav_kr(g, col);
  const unsigned lmask = k1_live_mask<T>(g, j);
  const unsigned m = ok ? lmask : 0u;
  const int last = (nslots - 1) * THREADS + tid;
  int slot = tid;
#pragma unroll
  for (int r = 0; r < PT; ++r) {
    const int sl = slot < last ? slot : last;
    const double gg = (field != 2) ? SG[sl] : 0.0;
    const cplx t = prep.apply(g, field, wav_l(g, j + r * T), kr, gg, S[sl]);
    const bool on = (m >> r) & 1u;
    v[r] = cmake(on ? t.x : 0.0, on ? t.y : 0.0);
    if (lmask >> r & 1) slot += THREADS;
    if ((r & 3) == 3) compiler_fence();
  }
}

template <int NY>
__global__ void __launch_bounds__(GroupCfg<NY>::THREADS, GroupCfg<NY>::MINB)
    k_col_inv_single(const __grid_constant__ GridDesc g, const __grid_constant__ PrepSingle prep, const __grid_constant__ PeerOut W,
                     const int ldw, const long long wfs, const int ncols, const int nfields, const int nslots, const int nunits,
                     const cplx* __restrict__ tw) {
  using GC = GroupCfg<NY>;
  constexpr int T = GC::T, THREADS = GC::THREADS, G = GC::G;
  GFB_DYN_SMEM(double, smd);
  const int tid = threadIdx.x;
  const int gi = tid / T, j = tid % T;
  double* xb = smd + gi * GC::GSD;
  cplx* S = reinterpret_cast<cplx*>(smd + GC::XB_DOUBLES) + 1;  // element [0] of the region = TMEM base word
  double* SG = reinterpret_cast<double*>(S + (size_t)nslots * THREADS);
  constexpr int tmcols = (THREADS / 128) * 64;
  const uint32_t tbase = tmem_setup(reinterpret_cast<uint32_t*>(smd + GC::XB_DOUBLES), tmcols);
  const uint32_t tA = tmem_thread_base(tbase, 64);
  const int stride = (int)gridDim.x * G, unit0 = (int)blockIdx.x * G + gi;
  const int ncolumns = 2 * ((nunits - (int)blockIdx.x * G + stride - 1) / stride);  // same for every group of the CTA
  cplx v[PT];

  // column ci of this group's stream = column (ci & 1) of unit unit0 + (ci >> 1) * stride
  k1_issue<T, THREADS>(S, SG, prep, g, unit0, 0, nunits, nfields, ncols, j, tid);
  cp_async_wait_all();
  k1_take<T, THREADS>(v, S, SG, prep, g, unit0, 0, nunits, nfields, ncols, nslots, j, tid);
  k1_issue<T, THREADS>(S, SG, prep, g, unit0, 1, nunits, nfields, ncols, j, tid);
#pragma unroll 1
  for (int ci = 0; ci < ncolumns; ++ci) {
    const int unit = unit0 + (ci >> 1) * stride, which = ci & 1;
    block_fft_split<NY, INV>(v, j, xb, tw);
    if (which == 0) {  // column A waits in TMEM for its partner
#pragma unroll
      for (int r = 0; r < PT; ++r) tmem_store(tA + 4 * r, v[r]);
      tmem_wait_st();
    } else {
      // `unit` depends on the group index and, for NY < 512 (T < 32), one warp spans several groups: the
      // warp-collective TMEM reads are therefore issued by EVERY thread (the slot is always valid) and only the
      // global stores are predicated on the unit being live (gfb_platform.h: no tcgen05 under divergence)
      const bool live = unit < nunits;
      const int field = live ? unit % nfields : 0, colA = live ? 2 * (unit / nfields) : 0;
      const int ooff = toff(j, colA, ldw);
#pragma unroll
      for (int c0 = 0; c0 < PT; c0 += 4) {
        cplx ra[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) ra[q] = tmem_load(tA + 4 * (c0 + q));
        tmem_wait_ld();
        if (live) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int r = c0 + q;
            const int d = (W.nranks == 1) ? 0 : (j + r * T) >> W.lg_rows;
            cplx* out = W.base[d] + (W.rank - d) * W.chunk + field * wfs;
            stpair(out + toff_row<T>(ooff, j, r, colA, ldw), ra[q], v[r]);
          }
        }
      }
    }
    cp_async_wait_all();
    k1_take<T, THREADS>(v, S, SG, prep, g, unit0 + ((ci + 1) >> 1) * stride, (ci + 1) & 1, nunits, nfields, ncols, nslots, j, tid);
    k1_issue<T, THREADS>(S, SG, prep, g, unit0 + ((ci + 2) >> 1) * stride, ci & 1, nunits, nfields, ncols, j, tid);
  }
  tmem_release(tbase, tmcols);
}

}  // namespace gfb
