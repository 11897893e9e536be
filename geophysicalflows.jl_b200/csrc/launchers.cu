// launchers.cu -- compiled once per FFT length: -DGFB_N=<16..8192>
#include "launch.h"

#ifndef GFB_N
#error "compile with -DGFB_N=<fft length>"
#endif

namespace gfb {

#define GFB_CAT2(a, b) a##b
#define GFB_CAT(a, b) GFB_CAT2(a, b)

template <class K>
static inline int set_smem(K kernel, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  return e == cudaSuccess ? 0 : -2;
}

template <class Prep>
static int do_col_inv(const ColInvLaunch& L, const Prep& prep) {
  using CC = ColCfg<GFB_N>;
  const size_t smem = (size_t)CC::C * CC::GS * sizeof(cplx);
  static bool attr = false;
  if (!attr) {
    if (set_smem(k_col_inv<GFB_N, Prep>, smem)) return -2;
    attr = true;
  }
  dim3 grid((L.ncols + CC::C - 1) / CC::C, L.nfields);
  GFB_LAUNCH((k_col_inv<GFB_N, Prep>), grid, dim3(CC::THREADS), smem, L.st, L.g, prep, L.W, L.ldw, L.wfs, L.ncols, L.tw);
  return 0;
}

template <int NL>
static int do_col_inv_multi(const ColInvLaunch& L) {
  PrepMulti<NL> p;
  p.sol = L.pm.sol;
  p.ld = L.pm.ld;
  p.ls = L.pm.ls;
  p.Sinv = L.pm.Sinv;
  p.ldS = L.pm.ldS;
  p.nf = L.pm.nf;
  return do_col_inv(L, p);
}

int GFB_CAT(launch_col_inv_, GFB_N)(const ColInvLaunch& L) {
  switch (L.prep) {
    case PREP_PLAIN: return do_col_inv(L, L.pp);
    case PREP_SINGLE: return do_col_inv(L, L.ps);
    case PREP_MULTI:
      switch (L.nl) {
        case 1: return do_col_inv_multi<1>(L);
        case 2: return do_col_inv_multi<2>(L);
        case 3: return do_col_inv_multi<3>(L);
        case 4: return do_col_inv_multi<4>(L);
        case 5: return do_col_inv_multi<5>(L);
      }
  }
  return -1;
}

template <int MODE>
static int do_row(const RowLaunch& L) {
  using RC = RowCfg<GFB_N>;
  const int nst = (MODE == ROW_PRODUCT) ? 1 : (MODE == ROW_PRODUCT_Q ? 2 : 0);
  const size_t smem = ((size_t)RC::G * RC::GS + (size_t)nst * PT * RC::THREADS) * sizeof(cplx);
  static bool attr = false;
  if (!attr) {
    if (set_smem(k_row<GFB_N, MODE>, smem)) return -2;
    attr = true;
  }
  const int nunits = L.a.ny / 2;
  dim3 grid((nunits + RC::G - 1) / RC::G, L.nlayers);
  GFB_LAUNCH((k_row<GFB_N, MODE>), grid, dim3(RC::THREADS), smem, L.st, L.a, L.tw);
  return 0;
}

int GFB_CAT(launch_row_, GFB_N)(const RowLaunch& L) {
  switch (L.mode) {
    case ROW_PRODUCT: return do_row<ROW_PRODUCT>(L);
    case ROW_C2R: return do_row<ROW_C2R>(L);
    case ROW_R2C: return do_row<ROW_R2C>(L);
    case ROW_PRODUCT_Q: return do_row<ROW_PRODUCT_Q>(L);
  }
  return -1;
}

template <int NL, bool TWO>
static int do_col_fwd(const ColFwdLaunch& L) {
  using CC = ColCfg<GFB_N>;
  const size_t xch = ((size_t)CC::C * ColSplitCfg<GFB_N>::GS + 1) & ~(size_t)1;
  const size_t smem = xch * sizeof(double) + (TWO ? (size_t)PT * CC::THREADS * sizeof(cplx) : 0);
  static bool attr = false;
  if (!attr) {
    if (set_smem(k_col_fwd<GFB_N, NL, TWO>, smem)) return -2;
    attr = true;
  }
  dim3 grid((L.ncols + CC::C - 1) / CC::C, L.nlayers);
  GFB_LAUNCH((k_col_fwd<GFB_N, NL, TWO>), grid, dim3(CC::THREADS), smem, L.st, L.g, L.P, L.ldp, L.pfs, L.e, L.ncols, L.tw);
  return 0;
}

int GFB_CAT(launch_col_fwd_, GFB_N)(const ColFwdLaunch& L) {
  if (!L.two) return do_col_fwd<1, false>(L);
  switch (L.nl) {
    case 1: return do_col_fwd<1, true>(L);
    case 2: return do_col_fwd<2, true>(L);
    case 3: return do_col_fwd<3, true>(L);
    case 4: return do_col_fwd<4, true>(L);
    case 5: return do_col_fwd<5, true>(L);
  }
  return -1;
}

}  // namespace gfb
