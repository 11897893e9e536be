// cols_dft.cu -- K1 / K3 for column lengths that are NOT a power of two.
//
// The reference's own tests run ny = 126 (test/test_twodnavierstokes.jl:259-317, test/test_multilayerqg.jl:567-655),
// ny = 66 (test/test_multilayerqg.jl:225-400) and ny = 34 (:663-760) next to power-of-two nx.  The Stockham kernels are
// radix-16 and need ny = 2^k; for any other even ny <= 1024 the two column passes run here as direct O(ny^2) transforms
// out of shared memory -- same pre-pass functors, same W / P layouts, same fused epilogue functions (epi_addends /
// epi_fetch / epi_finish of k3_col_fwd.cuh), so the row pass and everything else is shared with the fast path.  These
// sizes are test-sized problems (a 128 x 126 step moves ~1 MB); nothing here is tuned, it has to be exact and complete.
// One CTA per (column pair, field | layer); twiddles W_ny^i come from the same long-double table as the fast path and
// are indexed (l*y) mod ny incrementally, so no angle is ever evaluated on the device.
#include "launch.h"

namespace gfb {

constexpr int DFT_THREADS = 128;

template <class Prep>
__global__ void __launch_bounds__(DFT_THREADS)
    k_col_inv_dft(const GridDesc g, const Prep prep, const __grid_constant__ PeerOut W, const int ldw, const long long wfs, const int ncols,
                  const cplx* __restrict__ tw) {
  GFB_DYN_SMEM(double, smd);
  const int ny = g.ny, tid = threadIdx.x;
  cplx* A = reinterpret_cast<cplx*>(smd);
  cplx* B = A + ny;
  cplx* w = B + ny;
  const int colA = 2 * (int)blockIdx.x, field = blockIdx.y;
  if (colA >= ncols) return;
  const bool liveB = colA + 1 < ncols;
  for (int l = tid; l < ny; l += DFT_THREADS) {
    cplx a, b;
    prep.load_pair(g, field, l, colA, liveB, toff(l, colA, prep.ld), a, b);
    A[l] = a;
    B[l] = b;
    w[l] = tw[l];
  }
  __syncthreads();
  for (int y = tid; y < ny; y += DFT_THREADS) {
    cplx sa = cmake(0.0, 0.0), sb = cmake(0.0, 0.0);
    int idx = 0;  // (l*y) mod ny
    for (int l = 0; l < ny; ++l) {
      const cplx t = cconj(w[idx]);  // e^{+2 pi i l y / ny}
      sa = cadd(sa, cmul(A[l], t));
      sb = cadd(sb, cmul(B[l], t));
      idx += y;
      if (idx >= ny) idx -= ny;
    }
    const int d = (W.nranks == 1) ? 0 : (y >> W.lg_rows);
    cplx* out = W.base[d] + (W.rank - d) * W.chunk + field * wfs;
    stpair(out + toff(y, colA, ldw), sa, sb);
  }
}

template <int KIND, bool CL>
__device__ __forceinline__ void dft_epilogue_row(const GridDesc& g, const EpiArgs& e, int layer, int l, int colA, int off, cplx NA, cplx NB) {
  EpiIn<KIND, CL> in;
  epi_fetch<KIND, CL>(g, e, layer, l, colA, off, in);
  epi_finish<KIND, CL>(g, e, layer, off, NA, NB, in);
}

template <bool CL>
__device__ __forceinline__ void dft_epilogue_dispatch(const GridDesc& g, const EpiArgs& e, int layer, int l, int colA, int off, cplx NA, cplx NB) {
  switch (e.kind) {
#define GFB_EPI_CASE(K) case K: dft_epilogue_row<K, CL>(g, e, layer, l, colA, off, NA, NB); break;
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

// TWO: N = FFTy(P_d) - i l FFTy(P_v) + the fused epilogue ; !TWO: plain forward transform of one field -> dense user array
template <bool TWO>
__global__ void __launch_bounds__(DFT_THREADS)
    k_col_fwd_dft(const __grid_constant__ GridDesc g, const cplx* __restrict__ P, const int ldp, const long long pfs,
                  const __grid_constant__ EpiArgs e, const int ncols, const cplx* __restrict__ tw) {
  GFB_DYN_SMEM(double, smd);
  const int ny = g.ny, tid = threadIdx.x;
  cplx* w = reinterpret_cast<cplx*>(smd);
  cpair* D = reinterpret_cast<cpair*>(w + ny);
  cpair* V = D + ny;  // TWO only
  const int colA = 2 * (int)blockIdx.x, layer = blockIdx.y;
  if (colA >= ncols) return;
  const cplx* pd = P + (TWO ? 2 * layer : layer) * pfs;
  const cplx* pv = P + (2 * layer + 1) * pfs;
  for (int y = tid; y < ny; y += DFT_THREADS) {
    w[y] = tw[y];
    D[y] = ldpair(pd + toff(y, colA, ldp));
    if (TWO) V[y] = ldpair(pv + toff(y, colA, ldp));
  }
  __syncthreads();
  for (int l = tid; l < ny; l += DFT_THREADS) {
    cplx da = cmake(0.0, 0.0), db = da, va = da, vb = da;
    int idx = 0;
    for (int y = 0; y < ny; ++y) {
      const cplx t = w[idx];  // e^{-2 pi i l y / ny}
      const cpair d = D[y];
      da = cadd(da, cmul(d.a, t));
      db = cadd(db, cmul(d.b, t));
      if (TWO) {
        const cpair v = V[y];
        va = cadd(va, cmul(v.a, t));
        vb = cadd(vb, cmul(v.b, t));
      }
      idx += l;
      if (idx >= ny) idx -= ny;
    }
    if (!TWO) {
      cplx* o = e.out.p + layer * e.out.ls + (long long)l * e.out.ld + colA;
      o[0] = da;
      if (colA + 1 < g.nkl) o[1] = db;
      continue;
    }
    if (e.prune && l_aliased(g, l)) continue;
    const double ll = wav_l(g, l);
    cplx NA = cmake(da.x + ll * va.y, da.y - ll * va.x), NB = cmake(db.x + ll * vb.y, db.y - ll * vb.x);
    const int off = toff(l, colA, e.out.ld);
    epi_addends(g, e, layer, l, colA, off, NA, NB);
    if (e.cplxL) dft_epilogue_dispatch<true>(g, e, layer, l, colA, off, NA, NB);
    else dft_epilogue_dispatch<false>(g, e, layer, l, colA, off, NA, NB);
  }
}

template <class K>
static int dft_smem(K kernel, size_t bytes) {
  if (bytes <= 48 * 1024) return 0;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) == cudaSuccess ? 0 : -2;
}

template <class Prep>
static int do_col_inv_dft(const ColInvLaunch& L, const Prep& prep) {
  const size_t smem = (size_t)3 * L.g.ny * sizeof(cplx);
  if (dft_smem(k_col_inv_dft<Prep>, smem)) return -2;
  dim3 grid((L.ncols + 1) / 2, L.nfields);
  GFB_LAUNCH((k_col_inv_dft<Prep>), grid, dim3(DFT_THREADS), smem, L.st, L.g, prep, L.W, L.ldw, L.wfs, L.ncols, L.tw);
  return 0;
}

int launch_col_inv_dft(const ColInvLaunch& L) {
  if (L.ncols <= 0) return 0;
  switch (L.prep) {
    case PREP_PLAIN: return do_col_inv_dft(L, L.pp);
    case PREP_SINGLE: return do_col_inv_dft(L, L.ps);
    default: return do_col_inv_dft(L, L.pm);
  }
}

int launch_col_fwd_dft(const ColFwdLaunch& L) {
  if (L.ncols <= 0) return 0;
  const size_t smem = (size_t)L.g.ny * (sizeof(cplx) + (L.two ? 2 : 1) * sizeof(cpair));
  dim3 grid((L.ncols + 1) / 2, L.nlayers);
  if (L.two) {
    if (dft_smem(k_col_fwd_dft<true>, smem)) return -2;
    GFB_LAUNCH((k_col_fwd_dft<true>), grid, dim3(DFT_THREADS), smem, L.st, L.g, L.P, L.ldp, L.pfs, L.e, L.ncols, L.tw);
  } else {
    if (dft_smem(k_col_fwd_dft<false>, smem)) return -2;
    GFB_LAUNCH((k_col_fwd_dft<false>), grid, dim3(DFT_THREADS), smem, L.st, L.g, L.P, L.ldp, L.pfs, L.e, L.ncols, L.tw);
  }
  return 0;
}

}  // namespace gfb
