// k3_col_fwd_ws.cuh -- K3 at ny = 4096, warp-specialised (k_col_fwd_ws)
// (part of kernels.cuh: include that, not this file)
#pragma once
#include "k1_col_inv_ws.cuh"
#include "k3_col_fwd.cuh"

namespace gfb {

// ------------------------------------------------------------------------------------------------
// K3, warp-specialised: forward y-FFT of the two product fields + divergence + stage update, one 512-thread CTA per SM.
//
//   FFT group (warps 0..7)     per column pair: the four columns v_A, v_B, d_A, d_B arrive in a double-buffered shared
//                              memory stash; transform each, accumulate N = FFTy(d) - i l FFTy(v) of both columns in the
//                              thread's TMEM slots (2 x 2 x 64 columns: double-buffered by column pair).
//   helper group (warps 8..15) (1) streams the columns of the NEXT pairs from P into the stash with cp.async, always one
//                              column ahead of the transforms; (2) runs the time-stepper epilogue of the PREVIOUS pair --
//                              reads N out of TMEM (same lanes as its FFT partner thread), loads the state / coefficient
//                              tables, writes the stage arrays -- in four slices interleaved with the column loads.
// The epilogue is a latency-bound streaming pass over 5 or 6 arrays and the transforms are FP64/shared-memory bound: in
// the round-1 kernel they alternated inside each CTA (39 % of the warp samples waited on the epilogue's loads); here
// they run side by side on different warps, and the memory system stays busy during the transforms.
//
// Schedule of the helper group, one step per column n = 4 * unit + ci (ci = 0, 1: v_A, v_B; 2, 3: d_A, d_B):
//     step n:  wait until the FFT group has taken column n - 2 out of stash slot n & 1   [bar X_EMPTY]
//              issue the cp.async of column n into that slot
//              run epilogue slice (n - 5) & 3 (rows 4s .. 4s + 3 of every thread) of unit (n - 5) >> 2
//                   first slice of a unit: wait for its N                                [bar N_FULL]
//                   last slice: hand the TMEM buffer back                                [bar N_EMPTY]
//              wait for the cp.async, publish the column                                 [bar X_FULL]
// Named barriers: 1 = FFT group exchanges; 2,3 = X_FULL; 4,5 = X_EMPTY; 6,7 = N_FULL; 8,9 = N_EMPTY.
// ------------------------------------------------------------------------------------------------
constexpr int WS_BAR_NFULL = 6;
constexpr int WS_BAR_NEMPTY = 8;

// rows r_begin .. r_begin + 3 of this thread's 16 rows of the column pair, N of both columns in TMEM (ta, tb).
// (nlayer, ncolA, nr_begin): the slice this thread works on NEXT (nr_begin < 0: none) -- its tiles are prefetched into
// L2 first, by one thread in eight (a tile = the rows of eight neighbouring threads), so that the one-row-at-a-time loop
// below waits for an L2 hit instead of a DRAM round trip per row.
template <int T, int KIND, bool CL>
__device__ __forceinline__ void epilogue_rows_ws(const GridDesc& g, const EpiArgs& e, const int layer, const int colA,
                                                 const uint32_t ta, const uint32_t tb, const int j, const int r_begin,
                                                 const int nlayer, const int ncolA, const int nr_begin) {
  if (nr_begin >= 0 && (j & 7) == 0) {
    const int poff = toff(j, ncolA, e.out.ld);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int l = j + (nr_begin + q) * T;
      if (!(e.prune && l_aliased(g, l) && l_aliased(g, l + 7)))
        epi_prefetch_tile<KIND, CL>(g, e, nlayer, ncolA, toff_row<T>(poff, j, nr_begin + q, ncolA, e.out.ld));
    }
  }
  // one row (two wavenumbers) per iteration: the helper group shares the 128-register budget of the FFT group's
  // transform, and two rows of the six-array kinds spill; 256 threads x ~6 x 32 bytes in flight still cover the latency
  constexpr int CH = 1;
  const int off0 = toff(j, colA, e.out.ld);  // every ctx-owned array shares this leading dimension
  const bool addends = e.Nconst.p || e.F.p || (e.mu != 0.0 && layer == e.nlayers - 1);
#pragma unroll 1
  for (int r0 = r_begin; r0 < r_begin + 4; r0 += CH) {
    bool act[CH];
    int off[CH];
    cplx na[CH], nb[CH];
#pragma unroll
    for (int q = 0; q < CH; ++q) {
      na[q] = tmem_load(ta + 4 * (r0 + q));
      nb[q] = tmem_load(tb + 4 * (r0 + q));
    }
    tmem_wait_ld();
#pragma unroll
    for (int q = 0; q < CH; ++q) {
      const int l = j + (r0 + q) * T;
      act[q] = !(e.prune && l_aliased(g, l));
      off[q] = toff_row<T>(off0, j, r0 + q, colA, e.out.ld);
      // the optional additive terms first, so that their temporaries are dead before the stage arrays are fetched
      if (addends && act[q]) epi_addends(g, e, layer, l, colA, off[q], na[q], nb[q]);
    }
    EpiIn<KIND, CL> in[CH];
#pragma unroll
    for (int q = 0; q < CH; ++q)
      if (act[q]) epi_fetch<KIND, CL>(g, e, layer, j + (r0 + q) * T, colA, off[q], in[q]);
#pragma unroll
    for (int q = 0; q < CH; ++q)
      if (act[q]) epi_finish<KIND, CL>(g, e, layer, off[q], na[q], nb[q], in[q]);
  }
}

template <int T, bool CL>
__device__ __forceinline__ void epilogue_rows_ws_dispatch(const GridDesc& g, const EpiArgs& e, int layer, int colA, uint32_t ta,
                                                          uint32_t tb, int j, int r_begin, int nlayer, int ncolA, int nr_begin) {
  switch (e.kind) {
#define GFB_EPI_CASE(K) case K: epilogue_rows_ws<T, K, CL>(g, e, layer, colA, ta, tb, j, r_begin, nlayer, ncolA, nr_begin); break;
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

template <int NY>
__global__ void __launch_bounds__(WS_THREADS, 1)
    k_col_fwd_ws(const __grid_constant__ GridDesc g, const cplx* __restrict__ P, const int ldp, const long long pfs,
                 const __grid_constant__ EpiArgs e, const int npairs, const int nunits, const cplx* __restrict__ tw) {
  constexpr int T = NY / PT;
  static_assert(T == WS_FFT_THREADS, "one 256-thread FFT group: ny = 4096");
  GFB_DYN_SMEM(double, smd);
  cplx* xb = reinterpret_cast<cplx*>(smd);            // un-split exchange buffer of the FFT group
  cplx* X = xb + xbuf_elems(NY);                      // column stash, two slots of 16 x 256 values: X[slot][r * 256 + j]
  uint32_t* tmem_word = reinterpret_cast<uint32_t*>(X + 2 * PT * T);
  const int tid = threadIdx.x;
  const bool is_fft = tid < WS_FFT_THREADS;
  const int j = is_fft ? tid : tid - WS_FFT_THREADS;
  const uint32_t tbase = tmem_setup(tmem_word, 512);
  const uint32_t tm = ws_tmem_base(tbase);  // N buffer nb: column A at tm + 128 * nb, column B 64 columns further
  const int nmine = (nunits - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int ncolumns = 4 * nmine;
  // unit i of this CTA -> (layer, first column of the pair)
  auto unit_of = [&](int i, int& layer, int& colA) {
    const int u = (int)blockIdx.x + i * (int)gridDim.x;
    layer = u / npairs;
    colA = 2 * (u - layer * npairs);
  };

  if (!is_fft) {
    // ------------------------------------------ helper group ---------------------------------------------------------
#pragma unroll 1
    for (int n = 0; n < ncolumns + 5; ++n) {
      const int s = n & 1;
      if (n < ncolumns) {
        if (n >= 2) bar_sync(WS_BAR_EMPTY + s, WS_THREADS);  // the FFT group has column n - 2 in its registers
        int layer, colA;
        unit_of(n >> 2, layer, colA);
        const int ci = n & 3;
        const cplx* p = P + (ci < 2 ? 2 * layer + 1 : 2 * layer) * pfs;  // v field first, then d
        const int off0 = toff(j, colA, ldp) + (ci & 1);
        cplx* dst = X + s * (PT * T) + j;
#pragma unroll
        for (int r = 0; r < PT; ++r) cp_async16_cg(dst + r * T, p + off0 + r * T * ldp);
      }
      if (n >= 5 && n - 5 < ncolumns) {
        const int i = (n - 5) >> 2, slice = (n - 5) & 3, nb = i & 1;
        int layer, colA;
        unit_of(i, layer, colA);
        if (slice == 0) {  // N of unit i is complete
          bar_sync(WS_BAR_NFULL + nb, WS_THREADS);
          tmem_fence_after_sync();
        }
        const uint32_t ta = tm + 128 * nb, tb = ta + 64;
        // the slice after this one: same unit, or the first slice of this CTA's next unit
        int nlayer = layer, ncolA = colA, nr_begin = 4 * (slice + 1);
        if (slice == 3) {
          nr_begin = (i + 1 < nmine) ? 0 : -1;
          if (nr_begin == 0) unit_of(i + 1, nlayer, ncolA);
        }
        if (e.cplxL) epilogue_rows_ws_dispatch<T, true>(g, e, layer, colA, ta, tb, j, 4 * slice, nlayer, ncolA, nr_begin);
        else epilogue_rows_ws_dispatch<T, false>(g, e, layer, colA, ta, tb, j, 4 * slice, nlayer, ncolA, nr_begin);
        if (slice == 3 && i + 2 < nmine) {  // the FFT group may overwrite this buffer with unit i + 2
          tmem_fence_before_sync();
          bar_arrive(WS_BAR_NEMPTY + nb, WS_THREADS);
        }
      }
      if (n < ncolumns) {
        cp_async_wait_all();
        bar_arrive(WS_BAR_FULL + s, WS_THREADS);
      }
    }
  } else {
    // ------------------------------------------ FFT group ------------------------------------------------------------
    cplx v[PT];
    const Tw4096 twr = load_tw4096(j, tw);
#pragma unroll 1
    for (int n = 0; n < ncolumns; ++n) {
      const int s = n & 1, i = n >> 2, ci = n & 3, nb = i & 1;
      bar_sync(WS_BAR_FULL + s, WS_THREADS);
      {
        const cplx* src = X + s * (PT * T) + j;
#pragma unroll
        for (int r = 0; r < PT; ++r) v[r] = src[r * T];
      }
      if (n + 2 < ncolumns) bar_arrive(WS_BAR_EMPTY + s, WS_THREADS);
      block_fft4096<FWD, WS_BAR_FFT>(v, j, xb, twr);
      const uint32_t tdst = tm + 128 * nb + 64 * (ci & 1);
      if (ci < 2) {
        if (ci == 0 && i >= 2) {  // the epilogue of unit i - 2 has read this buffer
          bar_sync(WS_BAR_NEMPTY + nb, WS_THREADS);
          tmem_fence_after_sync();
        }
#pragma unroll
        for (int r = 0; r < PT; ++r) {
          const double ll = wav_l(g, j + r * T);
          v[r] = cmake(ll * v[r].y, -ll * v[r].x);  // -i l (vq)_h
        }
#pragma unroll
        for (int r = 0; r < PT; ++r) tmem_store(tdst + 4 * r, v[r]);  // 16 x .x4: one .x64 store is 5x slower (tools/tmem_bench)
        tmem_wait_st();
      } else {
#pragma unroll
        for (int c0 = 0; c0 < PT; c0 += 4) {
          cplx a[4];
          tmem_load4(tdst + 4 * c0, a);
          tmem_wait_ld();
#pragma unroll
          for (int q = 0; q < 4; ++q) a[q] = cadd(v[c0 + q], a[q]);
#pragma unroll
          for (int q = 0; q < 4; ++q) tmem_store(tdst + 4 * (c0 + q), a[q]);
        }
        tmem_wait_st();
        if (ci == 3) {
          tmem_fence_before_sync();
          bar_arrive(WS_BAR_NFULL + nb, WS_THREADS);
        }
      }
    }
  }
  tmem_release(tbase, 512);
}

}  // namespace gfb
