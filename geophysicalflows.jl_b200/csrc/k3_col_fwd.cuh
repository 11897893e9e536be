// k3_col_fwd.cuh -- K3: forward FFT along y + divergence + fused time-stepper epilogue (k_col_fwd)
// (part of kernels.cuh: include that, not this file)
#pragma once
#include "layout.cuh"

namespace gfb {

// ------------------------------------------------------------------------------------------------
// K3 epilogues: what happens to N(l, kr) = FFTy(d) - i l FFTy(v)
// ------------------------------------------------------------------------------------------------
enum EpiKind {
  EPI_STORE = 0,   // N (+F +Nconst + drag) -> out                (calcN! seam)
  EPI_RK_FIRST,    // RK4 stage 1 / first stage of any explicit RK
  EPI_RK_MID,
  EPI_RK_LAST,
  EPI_RK_ONLY,     // ForwardEuler
  EPI_ETD_1,
  EPI_ETD_2,
  EPI_ETD_3,
  EPI_ETD_4,
  EPI_AB3,
  EPI_LSRK,        // one stage of the low-storage RK (LSRK54): S2 = a S2 + d (N + L sol); sol += b S2 (in place)
  EPI_PLAIN        // plain forward transform: single field, store
};

// All spectral arrays the epilogue touches are ctx-owned (even leading dimension): every access below is a
// 256-bit column-pair access.
struct Coef {  // real or complex coefficient table [layer][l][kr]
  const double* r;
  const cplx* c;
  int ld;
  long long ls;
};
// a pair of coefficients: complex tables keep 8 registers, real tables only 4
template <bool CL>
struct CoefPair;
template <>
struct CoefPair<true> {
  cpair v;
  __device__ __forceinline__ void load(const Coef& t, int layer, int off) { v = ldpair(t.c + layer * t.ls + off); }
  __device__ __forceinline__ cplx mulA(cplx x) const { return cmul(v.a, x); }
  __device__ __forceinline__ cplx mulB(cplx x) const { return cmul(v.b, x); }
};
template <>
struct CoefPair<false> {
  double2 v;
  __device__ __forceinline__ void load(const Coef& t, int layer, int off) { v = ldpair_d(t.r + layer * t.ls + off); }
  __device__ __forceinline__ cplx mulA(cplx x) const { return cscale(x, v.x); }
  __device__ __forceinline__ cplx mulB(cplx x) const { return cscale(x, v.y); }
};

struct EpiArgs {
  int kind;
  int nlayers;
  int prune;       // 1: rows in the aliased l band are neither read nor written (non-final steps)
  int cplxL;       // coefficient tables are complex
  Spec out;        // EPI_STORE / EPI_PLAIN: destination of N ; stage kinds: the state written this stage
  Spec sol;        // s0 source (the problem state), masked on load
  Spec sin;        // stage input (== sol for stage 1), masked on load
  Spec acc;        // accumulator (AB3: rhs of the previous step)
  Spec n1;         // saved N1 (ETDRK4) / AB3: rhs two steps back
  Spec rhs;        // AB3: where this step's rhs is stored
  Spec F;          // forcing or null
  Spec Nconst;     // constant addition to N or null (SingleLayerQG: -i kr U eta_h)
  Spec prevsol;    // if non-null at a first stage: receives the dealiased state
  Coef L;          // linear operator
  Coef c1, c2, c3, c4;  // ETD: exp(L dt/2), zeta, exp(L dt) / alpha, beta or Gamma  (see gfb.cu)
  SpecR filter;    // null p -> no filter (layer stride 0: shared by layers)
  double a, b, d;  // RK: a = c_i*dt (sub-step), b = weight*dt ; AB3: a, b, d = dt*(23/12, -16/12, 5/12) ; LSRK: A_i, B_i, dt
  // MultiLayerQG bottom drag: N[nlayers-1] += mu * Krsq * psi_h[nlayers-1], psi = S^-1 q recomputed
  double mu;
  const double* Sinv;
  int ldS;
};

__device__ __forceinline__ cpair zero_pair() {
  cpair z;
  z.a = z.b = cmake(0.0, 0.0);
  return z;
}

// dealias mask applied on load: rows in the aliased l band and columns >= kcut read as zero (and are not fetched)
__device__ __forceinline__ cpair masked_pair(const GridDesc& g, const Spec& s, int layer, int l, int colA, int off) {
  const int ga = gcol(g, colA);
  if (ga >= g.kcut || l_aliased(g, l)) return zero_pair();
  cpair p = ldpair(s.p + layer * s.ls + off);
  if (ga + 1 >= g.kcut) p.b = cmake(0.0, 0.0);
  return p;
}

template <int KIND, bool CL>
struct EpiIn {
  cpair s0, si, acc, n1;
  CoefPair<CL> c1, c2, c3, c4, L;
  double2 filt;
};

// N += Nconst + drag + F for one row of the column pair (the optional additive terms of calcN!), done in a
// separate light pass so the stage-update pass below carries only what its kind needs
__device__ __forceinline__ void epi_addends(const GridDesc& g, const EpiArgs& e, int layer, int l, int colA, int off, cplx& NA,
                                            cplx& NB) {
  if (e.Nconst.p) {
    const cpair t = ldpair(e.Nconst.p + layer * e.Nconst.ls + off);
    NA = cadd(NA, t.a);
    NB = cadd(NB, t.b);
  }
  const int NL = e.nlayers;
  if (e.mu != 0.0 && layer == NL - 1) {
    // bottom drag on the (dealiased) stage input: psi_n = sum_m Sinv[n-1][m] q_m
    cplx pa = cmake(0.0, 0.0), pb = cmake(0.0, 0.0);
    const double* MA = e.Sinv + ((long long)l * e.ldS + colA) * (NL * NL) + (NL - 1) * NL;
    const double* MB = MA + NL * NL;
    const bool hasB = colA + 1 < g.nkl;
    for (int m = 0; m < NL; ++m) {
      const cpair q = masked_pair(g, e.sin, m, l, colA, off);
      const double ca = MA[m], cb = hasB ? MB[m] : 0.0;
      pa.x += ca * q.a.x;
      pa.y += ca * q.a.y;
      pb.x += cb * q.b.x;
      pb.y += cb * q.b.y;
    }
    const double ll = wav_l(g, l), ka = wav_kr(g, colA), kb = wav_kr(g, colA + 1);
    NA = cadd(NA, cscale(pa, e.mu * (ka * ka + ll * ll)));
    NB = cadd(NB, cscale(pb, e.mu * (kb * kb + ll * ll)));
  }
  if (e.F.p) {
    const cpair t = ldpair(e.F.p + layer * e.F.ls + off);
    NA = cadd(NA, t.a);
    NB = cadd(NB, t.b);
  }
}

// all loads of one row (column pair) ...
template <int KIND, bool CL>
__device__ __forceinline__ void epi_fetch(const GridDesc& g, const EpiArgs& e, int layer, int l, int colA, int off,
                                          EpiIn<KIND, CL>& in) {
  if (KIND == EPI_RK_LAST || KIND == EPI_RK_ONLY || KIND == EPI_ETD_4 || KIND == EPI_AB3) {
    in.filt = e.filter.p ? ldpair_d(e.filter.p + off) : make_double2(1.0, 1.0);
  }
  if (KIND == EPI_RK_FIRST || KIND == EPI_RK_MID || KIND == EPI_RK_ONLY || KIND == EPI_ETD_1 || KIND == EPI_ETD_2 ||
      KIND == EPI_ETD_3 || KIND == EPI_AB3)
    in.s0 = masked_pair(g, e.sol, layer, l, colA, off);
  if (KIND == EPI_RK_MID || KIND == EPI_RK_LAST) in.si = masked_pair(g, e.sin, layer, l, colA, off);
  if (KIND == EPI_RK_FIRST || KIND == EPI_RK_MID || KIND == EPI_RK_LAST || KIND == EPI_RK_ONLY || KIND == EPI_AB3)
    in.L.load(e.L, layer, off);
  if (KIND == EPI_RK_MID || KIND == EPI_RK_LAST || KIND == EPI_ETD_2 || KIND == EPI_ETD_3 || KIND == EPI_ETD_4)
    in.acc = ldpair(e.acc.p + layer * e.acc.ls + off);
  if (KIND == EPI_AB3 && (e.b != 0.0 || e.d != 0.0)) {
    in.acc = ldpair(e.acc.p + layer * e.acc.ls + off);
    in.n1 = ldpair(e.n1.p + layer * e.n1.ls + off);
  }
  if (KIND == EPI_LSRK) {
    in.filt = e.filter.p ? ldpair_d(e.filter.p + off) : make_double2(1.0, 1.0);  // filter set on the last stage only
    in.s0 = masked_pair(g, e.sol, layer, l, colA, off);
    in.L.load(e.L, layer, off);
    in.acc = (e.a != 0.0) ? ldpair(e.acc.p + layer * e.acc.ls + off) : zero_pair();
  }
  if (KIND == EPI_ETD_3) in.n1 = ldpair(e.n1.p + layer * e.n1.ls + off);
  if (KIND == EPI_ETD_1 || KIND == EPI_ETD_2 || KIND == EPI_ETD_3) {
    in.c1.load(e.c1, layer, off);
    in.c2.load(e.c2, layer, off);
  }
  if (KIND == EPI_ETD_1) in.c3.load(e.c3, layer, off);
  if (KIND == EPI_ETD_1 || KIND == EPI_ETD_2 || KIND == EPI_ETD_3 || KIND == EPI_ETD_4) in.c4.load(e.c4, layer, off);
}

// Start moving one 8-row tile (rows l .. l + 7 of the column pair at colA) of every array the epilogue of this kind READS
// from DRAM into L2 (same selection as epi_fetch; the masks of masked_pair are honoured so nothing useless moves).
template <int KIND, bool CL>
__device__ __forceinline__ void epi_prefetch_tile(const GridDesc& g, const EpiArgs& e, const int layer, const int colA, const int off) {
  auto pc = [&](const Spec& a) {
    if (a.p) prefetch_l2_bulk(a.p + layer * a.ls + off, 256);
  };
  auto pk = [&](const Coef& t) {
    if (CL) prefetch_l2_bulk(t.c + layer * t.ls + off, 256);
    else prefetch_l2_bulk(t.r + layer * t.ls + off, 128);
  };
  const bool live_col = gcol(g, colA) < g.kcut;
  if (KIND == EPI_RK_LAST || KIND == EPI_RK_ONLY || KIND == EPI_ETD_4 || KIND == EPI_AB3) {
    if (e.filter.p) prefetch_l2_bulk(e.filter.p + off, 128);
  }
  if ((KIND == EPI_RK_FIRST || KIND == EPI_RK_MID || KIND == EPI_RK_ONLY || KIND == EPI_ETD_1 || KIND == EPI_ETD_2 || KIND == EPI_ETD_3 ||
       KIND == EPI_AB3) && live_col)
    pc(e.sol);
  if ((KIND == EPI_RK_MID || KIND == EPI_RK_LAST) && live_col) pc(e.sin);
  if (KIND == EPI_RK_FIRST || KIND == EPI_RK_MID || KIND == EPI_RK_LAST || KIND == EPI_RK_ONLY || KIND == EPI_AB3) pk(e.L);
  if (KIND == EPI_RK_MID || KIND == EPI_RK_LAST || KIND == EPI_ETD_2 || KIND == EPI_ETD_3 || KIND == EPI_ETD_4) pc(e.acc);
  if (KIND == EPI_AB3 && (e.b != 0.0 || e.d != 0.0)) {
    pc(e.acc);
    pc(e.n1);
  }
  if (KIND == EPI_ETD_3) pc(e.n1);
  if (KIND == EPI_LSRK) {
    if (e.filter.p) prefetch_l2_bulk(e.filter.p + off, 128);
    if (live_col) pc(e.sol);
    pk(e.L);
    if (e.a != 0.0) pc(e.acc);
  }
  if (KIND == EPI_ETD_1 || KIND == EPI_ETD_2 || KIND == EPI_ETD_3) {
    pk(e.c1);
    pk(e.c2);
  }
  if (KIND == EPI_ETD_1) pk(e.c3);
  if (KIND == EPI_ETD_1 || KIND == EPI_ETD_2 || KIND == EPI_ETD_3 || KIND == EPI_ETD_4) pk(e.c4);
  pc(e.Nconst);
  pc(e.F);
}

// prefetch (DRAM -> L2) of the epilogue's input tiles of rows [r_begin, r_end) of this thread -- issued by one thread in
// eight (a tile = the rows of eight neighbouring threads), kind-dispatched like the epilogue itself
template <int T, bool CL>
__device__ __forceinline__ void epi_prefetch_rows(const GridDesc& g, const EpiArgs& e, const int layer, const int colA, const int j,
                                                  const int r_begin, const int r_end) {
  if ((j & 7) != 0) return;
  const int poff = toff(j, colA, e.out.ld);
#pragma unroll 1
  for (int r = r_begin; r < r_end; ++r) {
    const int l = j + r * T;
    if (e.prune && l_aliased(g, l) && l_aliased(g, l + 7)) continue;
    const int off = toff_row<T>(poff, j, r, colA, e.out.ld);
    switch (e.kind) {
#define GFB_EPI_CASE(K) case K: epi_prefetch_tile<K, CL>(g, e, layer, colA, off); break;
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
}

__device__ __forceinline__ cplx axpy(cplx y, double a, cplx x) { return cmake(y.x + a * x.x, y.y + a * x.y); }

// ... then the arithmetic and all stores of that row
template <int KIND, bool CL>
__device__ __forceinline__ void epi_finish(const GridDesc& g, const EpiArgs& e, int layer, int off, cplx NA,
                                           cplx NB, const EpiIn<KIND, CL>& in) {
  cplx* out = e.out.p + layer * e.out.ls + off;
  if (KIND == EPI_STORE) {
    stpair(out, NA, NB);
  } else if (KIND == EPI_RK_FIRST) {
    const cplx ra = cadd(NA, in.L.mulA(in.s0.a)), rb = cadd(NB, in.L.mulB(in.s0.b));
    stpair(out, axpy(in.s0.a, e.a, ra), axpy(in.s0.b, e.a, rb));
    stpair(e.acc.p + layer * e.acc.ls + off, axpy(in.s0.a, e.b, ra), axpy(in.s0.b, e.b, rb));
    if (e.prevsol.p) stpair(e.prevsol.p + layer * e.prevsol.ls + off, in.s0.a, in.s0.b);
  } else if (KIND == EPI_RK_MID) {
    const cplx ra = cadd(NA, in.L.mulA(in.si.a)), rb = cadd(NB, in.L.mulB(in.si.b));
    stpair(out, axpy(in.s0.a, e.a, ra), axpy(in.s0.b, e.a, rb));
    stpair(e.acc.p + layer * e.acc.ls + off, axpy(in.acc.a, e.b, ra), axpy(in.acc.b, e.b, rb));
  } else if (KIND == EPI_RK_LAST) {
    const cplx ra = cadd(NA, in.L.mulA(in.si.a)), rb = cadd(NB, in.L.mulB(in.si.b));
    stpair(out, cscale(axpy(in.acc.a, e.b, ra), in.filt.x), cscale(axpy(in.acc.b, e.b, rb), in.filt.y));
  } else if (KIND == EPI_RK_ONLY) {
    const cplx ra = cadd(NA, in.L.mulA(in.s0.a)), rb = cadd(NB, in.L.mulB(in.s0.b));
    if (e.prevsol.p) stpair(e.prevsol.p + layer * e.prevsol.ls + off, in.s0.a, in.s0.b);
    stpair(out, cscale(axpy(in.s0.a, e.b, ra), in.filt.x), cscale(axpy(in.s0.b, e.b, rb), in.filt.y));
  } else if (KIND == EPI_ETD_1) {  // c1 = exp(L dt/2), c2 = zeta, c3 = exp(L dt), c4 = alpha
    stpair(out, cadd(in.c1.mulA(in.s0.a), in.c2.mulA(NA)),
           cadd(in.c1.mulB(in.s0.b), in.c2.mulB(NB)));
    stpair(e.acc.p + layer * e.acc.ls + off, cadd(in.c3.mulA(in.s0.a), in.c4.mulA(NA)),
           cadd(in.c3.mulB(in.s0.b), in.c4.mulB(NB)));
    stpair(e.n1.p + layer * e.n1.ls + off, NA, NB);
    if (e.prevsol.p) stpair(e.prevsol.p + layer * e.prevsol.ls + off, in.s0.a, in.s0.b);
  } else if (KIND == EPI_ETD_2) {  // c4 = beta
    stpair(out, cadd(in.c1.mulA(in.s0.a), in.c2.mulA(NA)),
           cadd(in.c1.mulB(in.s0.b), in.c2.mulB(NB)));
    stpair(e.acc.p + layer * e.acc.ls + off, cadd(in.acc.a, in.c4.mulA(cscale(NA, 2.0))),
           cadd(in.acc.b, in.c4.mulB(cscale(NB, 2.0))));
  } else if (KIND == EPI_ETD_3) {  // e.sol = sol1 (dealiased on load), out = next stage input
    const cplx ta = cmake(2.0 * NA.x - in.n1.a.x, 2.0 * NA.y - in.n1.a.y);
    const cplx tb = cmake(2.0 * NB.x - in.n1.b.x, 2.0 * NB.y - in.n1.b.y);
    stpair(out, cadd(in.c1.mulA(in.s0.a), in.c2.mulA(ta)),
           cadd(in.c1.mulB(in.s0.b), in.c2.mulB(tb)));
    stpair(e.acc.p + layer * e.acc.ls + off, cadd(in.acc.a, in.c4.mulA(cscale(NA, 2.0))),
           cadd(in.acc.b, in.c4.mulB(cscale(NB, 2.0))));
  } else if (KIND == EPI_ETD_4) {  // c4 = Gamma
    stpair(out, cscale(cadd(in.acc.a, in.c4.mulA(NA)), in.filt.x),
           cscale(cadd(in.acc.b, in.c4.mulB(NB)), in.filt.y));
  } else if (KIND == EPI_AB3) {
    const cplx ra = cadd(NA, in.L.mulA(in.s0.a)), rb = cadd(NB, in.L.mulB(in.s0.b));
    cplx ua = cscale(ra, e.a), ub = cscale(rb, e.a);
    if (e.b != 0.0 || e.d != 0.0) {
      ua = axpy(axpy(ua, e.b, in.acc.a), e.d, in.n1.a);
      ub = axpy(axpy(ub, e.b, in.acc.b), e.d, in.n1.b);
    }
    if (e.prevsol.p) stpair(e.prevsol.p + layer * e.prevsol.ls + off, in.s0.a, in.s0.b);
    stpair(e.rhs.p + layer * e.rhs.ls + off, ra, rb);
    stpair(out, cscale(cadd(in.s0.a, ua), in.filt.x), cscale(cadd(in.s0.b, ub), in.filt.y));
  } else if (KIND == EPI_LSRK) {  // a = A_i, b = B_i, d = dt
    const cplx ra = cadd(NA, in.L.mulA(in.s0.a)), rb = cadd(NB, in.L.mulB(in.s0.b));
    const cplx sa = axpy(cscale(in.acc.a, e.a), e.d, ra), sb = axpy(cscale(in.acc.b, e.a), e.d, rb);
    if (e.prevsol.p) stpair(e.prevsol.p + layer * e.prevsol.ls + off, in.s0.a, in.s0.b);
    stpair(e.acc.p + layer * e.acc.ls + off, sa, sb);
    stpair(out, cscale(axpy(in.s0.a, e.b, sa), in.filt.x), cscale(axpy(in.s0.b, e.b, sb), in.filt.y));
  }
}

// Epilogue of one column-pair tile: N of column A waits in the stash X, of column B in the scratch slot.
// Two rows (four wavenumbers) are fetched before anything is stored, so ~10 independent 256-bit loads per
// thread are in flight; the kind is a template parameter so only the arrays a stage needs are touched.
// N of column A waits in the shared-memory stash X, N of column B in the thread's TMEM slot `tb`.
// TMEM accesses are warp-collective, so they sit outside every row-dependent branch.
// The thread's 16 wavenumber rows are l = l0 + r*S: (l0, S) = (j, T) for a register-resident line, (4j + q, 4T) for
// the q-th sub-transform of a 4T*16-point line (k_col_fwd16k).
template <int NY, int KIND, bool CL, int S = GroupCfg<NY>::T>
__device__ __forceinline__ void epilogue_tile(const GridDesc& g, const EpiArgs& e, const int layer, const int colA,
                                              const bool liveA, cplx* __restrict__ X, const uint32_t tb, const int l0,
                                              const int tid, const int pf = 0) {
  using GC = GroupCfg<NY>;
  constexpr int T = S, THREADS = GC::THREADS;
  const int j = l0;
  // rows fetched before anything is stored.  Four rows of the kinds that read five or six arrays per wavenumber do not
  // fit the register file next to N (ptxas spilled ~2 KB per thread): those take two rows at a time.
  constexpr int CH = (KIND == EPI_STORE || KIND == EPI_RK_LAST || KIND == EPI_RK_ONLY || KIND == EPI_ETD_4) ? 4 : 2;
  const int off0 = toff(j, colA, e.out.ld);  // every ctx-owned array shares this leading dimension
  if (e.Nconst.p || e.F.p || (e.mu != 0.0 && layer == e.nlayers - 1)) {
#pragma unroll 1
    for (int r0 = 0; r0 < PT; r0 += CH) {
      cplx nb[CH];
#pragma unroll
      for (int q = 0; q < CH; ++q) nb[q] = tmem_load(tb + 4 * (r0 + q));
      tmem_wait_ld();
#pragma unroll
      for (int q = 0; q < CH; ++q) {
        const int r = r0 + q, l = j + r * T;
        if (liveA && !(e.prune && l_aliased(g, l))) {
          cplx NA = X[r * THREADS + tid];
          epi_addends(g, e, layer, l, colA, toff_row<T>(off0, j, r, colA, e.out.ld), NA, nb[q]);
          X[r * THREADS + tid] = NA;
        }
      }
#pragma unroll
      for (int q = 0; q < CH; ++q) tmem_store(tb + 4 * (r0 + q), nb[q]);
    }
    tmem_wait_st();
  }
  // pf == 2: rolling L2 prefetch -- the input tiles of the chunk two iterations ahead start moving DRAM -> L2 while this
  // chunk is fetched and finished (one thread in eight issues: a tile = the rows of eight neighbouring threads); nothing
  // is prefetched that is not consumed a few microseconds later, so no extra DRAM traffic (unlike the bulk variant, pf == 1)
  constexpr int PFD = 2 * CH;
  auto pf_rows = [&](const int rb) {
    if (T % 8 != 0 || (j & 7) != 0 || !liveA) return;
#pragma unroll
    for (int q = 0; q < CH; ++q) {
      const int r = rb + q;
      if (r < PT) {
        const int l = j + r * T;
        if (!(e.prune && l_aliased(g, l) && l_aliased(g, l + 7))) epi_prefetch_tile<KIND, CL>(g, e, layer, colA, toff_row<T>(off0, j, r, colA, e.out.ld));
      }
    }
  };
  if (pf == 2) {
    pf_rows(0);
    pf_rows(CH);
  }
#pragma unroll 1
  for (int r0 = 0; r0 < PT; r0 += CH) {
    if (pf == 2) pf_rows(r0 + PFD);
    EpiIn<KIND, CL> in[CH];
    bool act[CH];
    int off[CH];
#pragma unroll
    for (int q = 0; q < CH; ++q) {
      const int l = j + (r0 + q) * T;
      act[q] = liveA && !(e.prune && l_aliased(g, l));
      off[q] = toff_row<T>(off0, j, r0 + q, colA, e.out.ld);
      if (act[q]) epi_fetch<KIND, CL>(g, e, layer, l, colA, off[q], in[q]);
    }
#pragma unroll
    for (int q = 0; q < CH; ++q) {
      const cplx nb = tmem_load(tb + 4 * (r0 + q));  // TMEM latency is tiny next to the global loads above
      tmem_wait_ld();
      if (act[q]) epi_finish<KIND, CL>(g, e, layer, off[q], X[(r0 + q) * THREADS + tid], nb, in[q]);
    }
  }
}

template <int NY, bool CL, int S = GroupCfg<NY>::T>
__device__ __forceinline__ void epilogue_dispatch(const GridDesc& g, const EpiArgs& e, int layer, int colA, bool liveA, cplx* X,
                                                  uint32_t tb, int l0, int tid, int pf = 0) {
  switch (e.kind) {
#define GFB_EPI_CASE(K) case K: epilogue_tile<NY, K, CL, S>(g, e, layer, colA, liveA, X, tb, l0, tid, pf); break;
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

// K3.  TWO = true: N = FFTy(Pd) - i l FFTy(Pv) ; TWO = false: plain forward transform of one field.
// One CTA group per column pair.  The other column waits in the shared-memory stash X and the two (-i l)-weighted
// v-transforms are parked in the thread's TMEM slots (2 x 64 columns per thread, 256 of the SM's 512 columns per
// CTA -- two CTAs per SM use all of it).  The epilogue then touches every spectral array with complete 32-byte
// sectors (both columns of a row).
template <int NY>
struct TmemCfg {
  static constexpr int PER_WG = 128;                                  // two 64-column slots per thread
  static constexpr int COLS = (GroupCfg<NY>::THREADS / 128) * PER_WG;  // 256 (512 for 512-thread CTAs)
};

// start fetching one column (16 B per row) of a tiled P field into the thread-private stash
template <int NY>
__device__ __forceinline__ void k3_issue(cplx* __restrict__ X, const cplx* __restrict__ p, const int ldp, const int colA, const int which,
                                         const int j, const int tid) {
  using GC = GroupCfg<NY>;
  constexpr int T = GC::T, THREADS = GC::THREADS;
  const int off0 = toff(j, colA, ldp);
#pragma unroll
  for (int r = 0; r < PT; ++r) cp_async16_cg(&X[r * THREADS + tid], p + toff_row<T>(off0, j, r, colA, ldp) + which);
}

// ONE transform call site in a loop over the columns (four inlined copies of the transform made ptxas spill
// kilobytes): ci = 0, 1 are the v-field columns A, B (their -i l weighted transforms wait in TMEM), ci = 2, 3 the
// d-field columns.  Every column arrives by cp.async in the thread-private stash X while the previous one is
// being transformed.
template <int NY, bool TWO>
__global__ void __launch_bounds__(GroupCfg<NY>::THREADS, GroupCfg<NY>::MINB)
    k_col_fwd(const __grid_constant__ GridDesc g, const cplx* __restrict__ P, const int ldp, const long long pfs,
              const __grid_constant__ EpiArgs e, const int ncols, const cplx* __restrict__ tw, const int l2pf) {
  using GC = GroupCfg<NY>;
  constexpr int T = GC::T, THREADS = GC::THREADS;
  GFB_DYN_SMEM(double, smd);
  const int tid = threadIdx.x;
  const int gi = tid / T, j = tid % T;
  const int layer = blockIdx.y;
  double* xb = smd + gi * GC::GSD;
  cplx* X = reinterpret_cast<cplx*>(smd + GC::XB_DOUBLES);
  const int colA = 2 * (blockIdx.x * GC::G + gi);
  const bool liveA = colA < ncols;
  const cplx* pd = P + (TWO ? 2 * layer : layer) * pfs;
  const cplx* pv = P + (2 * layer + 1) * pfs;  // TWO only
  cplx v[PT];
  uint32_t tbase = 0, tA = 0, tB = 0;
  if (TWO) {
    tbase = tmem_setup(reinterpret_cast<uint32_t*>(X + GC::STASH), TmemCfg<NY>::COLS);
    tA = tmem_thread_base(tbase, TmemCfg<NY>::PER_WG);
    tB = tA + 64;
  }
  constexpr int NCI = TWO ? 4 : 2;
  if (liveA) k3_issue<NY>(X, TWO ? pv : pd, ldp, colA, 0, j, tid);
  // The epilogue streams 5-6 arrays one or two rows at a time and used to wait a DRAM round trip per chunk (39 % of the
  // warp samples).  Its input tiles of this thread's first eight rows start moving DRAM -> L2 NOW, four transforms before
  // they are needed; the second eight follow when the epilogue starts (bounded L2 footprint: ~50 MB over all CTAs).
  if (TWO && l2pf == 1 && liveA && T % 8 == 0) {
    if (e.cplxL) epi_prefetch_rows<T, true>(g, e, layer, colA, j, 0, PT / 2);
    else epi_prefetch_rows<T, false>(g, e, layer, colA, j, 0, PT / 2);
  }
  cp_async_wait_all();
#pragma unroll
  for (int r = 0; r < PT; ++r) v[r] = liveA ? X[r * THREADS + tid] : cmake(0.0, 0.0);
  if (liveA) k3_issue<NY>(X, TWO ? pv : pd, ldp, colA, 1, j, tid);
#pragma unroll 1
  for (int ci = 0; ci < NCI; ++ci) {
    block_fft_split<NY, FWD>(v, j, xb, tw);
    if (TWO && ci < 2) {
      // -i l (vq)_h of column A / B -> TMEM, then take the next column and start fetching the one after it
      const uint32_t tdst = (ci == 0) ? tA : tB;
#pragma unroll
      for (int r = 0; r < PT; ++r) {
        const double ll = wav_l(g, j + r * T);
        tmem_store(tdst + 4 * r, cmake(ll * v[r].y, -ll * v[r].x));
      }
      cp_async_wait_all();
#pragma unroll
      for (int r = 0; r < PT; ++r) v[r] = liveA ? X[r * THREADS + tid] : cmake(0.0, 0.0);
      if (liveA) k3_issue<NY>(X, pd, ldp, colA, ci, j, tid);
      tmem_wait_st();
    } else if (ci == NCI - 2) {
      // first column of the d field: N_A = FFTy(d_A) (+ its parked v partner) goes to the stash, column B comes out
      cp_async_wait_all();
#pragma unroll
      for (int c0 = 0; c0 < PT; c0 += 4) {
        cplx a[4];
        if (TWO) {
#pragma unroll
          for (int q = 0; q < 4; ++q) a[q] = tmem_load(tA + 4 * (c0 + q));
          tmem_wait_ld();
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const cplx t = liveA ? X[(c0 + q) * THREADS + tid] : cmake(0.0, 0.0);
          X[(c0 + q) * THREADS + tid] = TWO ? cadd(v[c0 + q], a[q]) : v[c0 + q];
          v[c0 + q] = t;
        }
      }
    }
  }
  if constexpr (!TWO) {
    if (liveA) {
#pragma unroll
      for (int r = 0; r < PT; ++r) {
        const int l = j + r * T;
        cplx* o = e.out.p + layer * e.out.ls + (long long)l * e.out.ld + colA;  // EPI_PLAIN: dense user array
        o[0] = X[r * THREADS + tid];
        if (colA + 1 < g.nkl) o[1] = v[r];
      }
    }
  } else {
    // N of column B -> TMEM slot A (its phase-V partner is consumed), then the kind-specialised epilogue
#pragma unroll
    for (int c0 = 0; c0 < PT; c0 += 4) {
      cplx b[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) b[q] = tmem_load(tB + 4 * (c0 + q));
      tmem_wait_ld();
#pragma unroll
      for (int q = 0; q < 4; ++q) tmem_store(tA + 4 * (c0 + q), cadd(v[c0 + q], b[q]));
    }
    tmem_wait_st();
    if (l2pf == 1 && liveA && T % 8 == 0) {
      if (e.cplxL) epi_prefetch_rows<T, true>(g, e, layer, colA, j, PT / 2, PT);
      else epi_prefetch_rows<T, false>(g, e, layer, colA, j, PT / 2, PT);
    }
    if (e.cplxL) epilogue_dispatch<NY, true>(g, e, layer, colA, liveA, X, tA, j, tid, l2pf);
    else epilogue_dispatch<NY, false>(g, e, layer, colA, liveA, X, tA, j, tid, l2pf);
    tmem_release(tbase, TmemCfg<NY>::COLS);
  }
}

}  // namespace gfb
