// k2_row_ws.cuh -- K2 product mode of the single-layer modules at nx = 4096, warp-specialised (k_row_product_ws)
// (part of kernels.cuh: include that, not this file)
#pragma once
#include "k1_col_inv_ws.cuh"
#include "k2_row.cuh"

namespace gfb {

// ------------------------------------------------------------------------------------------------
// K2, warp-specialised: one 512-thread CTA per SM, persistent over row pairs.  Three phases per row pair, as in
// k_row_product: (q_a + i q_b) ; (u_a + i v_a) ; (u_b + i v_b).
//
//   FFT group (warps 0..7)      phase input (16 points per thread, Hermitian-assembled) out of its TMEM in-slot -> inverse
//                               transform -> phase 0: park q of both rows in TMEM; phases 1, 2: multiply by q, forward
//                               transform, hand the raw spectrum Z = (uq)_h + i (vq)_h to the helper group through the
//                               TMEM out-slot.  Five transforms per row pair and nothing else.
//   helper group (warps 8..15)  input side: the two half-spectrum rows of a phase are gathered out of the tiled W by TMA
//                               tensor loads (one thread, mbarrier) into a double-buffered line buffer, TWO phases ahead;
//                               the Hermitian extension Z[k] = A[k] + i B[k], Z[N-k] = conj A[k] + i conj B[k] is
//                               assembled by the helper threads and written into the partner FFT thread's in-slot.
//                               output side: takes Z out of the out-slot, separates the two real rows' spectra (the
//                               Z[k] <-> Z[N-k] exchange through shared memory, helper-group barrier), applies -i kr,
//                               lays the d and v rows out densely and scatters them into the tiled P by TMA tensor stores.
// TMEM per FFT thread: [in 0: 64][in 1: 64][q: 64][out: 64] = 256 columns (all 512 for the two warps of a quadrant).
// Named barriers: 1 FFT group, 10 helper group, 2/3 IN_FULL, 4/5 IN_EMPTY, 6 OUT_FULL, 7 OUT_EMPTY.
// Shared memory: FFT exchange buffer (un-split, 68 KB) | control | line buffer 2 slots x 2 rows | staging (d row, v row;
// doubles as the exchange buffer of the separation step).
// ------------------------------------------------------------------------------------------------
constexpr int WS_BAR_HELP = 10;
constexpr int WS_BAR_OUT_FULL = 6;
constexpr int WS_BAR_OUT_EMPTY = 7;

template <int NX>
__global__ void __launch_bounds__(WS_THREADS, 1)
    k_row_product_ws(const __grid_constant__ RowArgs a, const __grid_constant__ RowTma tm, const cplx* __restrict__ tw, const int nunits) {
  constexpr int T = NX / PT;
  static_assert(T == WS_FFT_THREADS, "one 256-thread FFT group: nx = 4096");
  GFB_DYN_SMEM(double, smd);
  cplx* xb = reinterpret_cast<cplx*>(smd);
  unsigned char* ctl = reinterpret_cast<unsigned char*>(xb + xbuf_elems(NX));  // [tmem word | 2 mbarriers]
  Mbar* full = reinterpret_cast<Mbar*>(ctl + 16);
  const int kcp = 2 * tm.nbox_in * tm.boxp_in;    // complex elements per staged input row
  const int kop = 2 * tm.nbox_out * tm.boxp_out;  // ... per staged output row
  cplx* lb = reinterpret_cast<cplx*>(ctl + 128);  // slot s: rows lb + s * 2 * kcp, lb + s * 2 * kcp + kcp
  cplx* stage = lb + 4 * (size_t)kcp;             // d row, then v row (kop each); >= xbuf doubles for the separation step
  const int tid = threadIdx.x;
  const bool is_fft = tid < WS_FFT_THREADS;
  const int j = is_fft ? tid : tid - WS_FFT_THREADS;
  if (tid == WS_FFT_THREADS) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_fence_init();
  }
  const uint32_t tbase = tmem_setup(reinterpret_cast<uint32_t*>(ctl), 512);
  const uint32_t tmb = ws_tmem_base(tbase);
  const uint32_t t_q = tmb + 128, t_out = tmb + 192;
  const int npairs = a.ny / 2;
  const int nmine = (nunits - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int nphases = 3 * nmine;

  if (!is_fft) {
    // ------------------------------------------ helper group ---------------------------------------------------------
    // one elected thread gathers the two rows of phase m out of W into line-buffer slot m & 1
    auto prefetch = [&](int m) {
      const int unit = (int)blockIdx.x + (m / 3) * (int)gridDim.x, ph = m % 3;
      const int layer = unit / npairs, ya = 2 * (unit - layer * npairs);
      const int fu = a.wstride * layer, fv = fu + 1, fq = fu + 2;
      const int fA = ph == 0 ? fq : fu, fB = ph == 0 ? fq : fv;
      const int rowA = ph == 0 ? ya : ya + ph - 1, rowB = ph == 0 ? ya + 1 : ya + ph - 1;
      cplx* dst = lb + (size_t)(m & 1) * 2 * kcp;
      mbar_expect_tx(&full[m & 1], (unsigned)(2 * kcp * sizeof(cplx)));
      for (int b = 0; b < tm.nbox_in; ++b) {
        tma_load_5d(dst + 2 * b * tm.boxp_in, &tm.w, &full[m & 1], 0, rowA & 7, b * tm.boxp_in, fA, rowA >> 3);
        tma_load_5d(dst + kcp + 2 * b * tm.boxp_in, &tm.w, &full[m & 1], 0, rowB & 7, b * tm.boxp_in, fB, rowB >> 3);
      }
    };
    if (j == 0) {
      if (nphases > 0) prefetch(0);
      if (nphases > 1) prefetch(1);
    }
    cplx v[PT];
    cplx A[9], B[9];
#pragma unroll 1
    for (int m = 0; m <= nphases; ++m) {
      if (m < nphases) {
        // ---- input side of phase m ----
        const int s = m & 1;
        mbar_wait(&full[s], (unsigned)((m >> 1) & 1));
        assemble_hermitian<NX>(v, j, lb + (size_t)s * 2 * kcp, kcp, a.kc);
        bar_sync(WS_BAR_HELP, WS_FFT_THREADS);  // every helper thread has read the slot: refill it two phases ahead
        if (j == 0 && m + 2 < nphases) prefetch(m + 2);
        if (m >= 2) {
          bar_sync(WS_BAR_EMPTY + s, WS_THREADS);  // the FFT group has taken phase m - 2 out of the in-slot
          tmem_fence_after_sync();
        }
#pragma unroll
        for (int r = 0; r < PT; ++r) tmem_store(tmb + 64 * s + 4 * r, v[r]);  // 16 x .x4: one .x64 store is 5x slower (tools/tmem_bench)
        tmem_wait_st();
        tmem_fence_before_sync();
        bar_arrive(WS_BAR_FULL + s, WS_THREADS);
      }
      // ---- output side of phase m - 1 (phases 1 and 2 of a row pair produce a d row and a v row) ----
      const int mo = m - 1;
      if (mo >= 0 && mo % 3 != 0) {
        const int unit = (int)blockIdx.x + (mo / 3) * (int)gridDim.x, h = mo % 3 - 1;
        const int layer = unit / npairs, y = 2 * (unit - layer * npairs) + h;
        bar_sync(WS_BAR_OUT_FULL, WS_THREADS);
        tmem_fence_after_sync();
        tmem_load16(t_out, v);
        tmem_wait_ld();
        if (mo + 1 < nphases) {  // another output follows (phase mo + 1 or mo + 2): hand the out-slot back
          tmem_fence_before_sync();
          bar_arrive(WS_BAR_OUT_EMPTY, WS_THREADS);
        }
        if (j == 0) tma_wait_read<0>();  // the previous rows have left the staging buffer
        bar_sync(WS_BAR_HELP, WS_FFT_THREADS);
        unpack_pair<NX, WS_BAR_HELP>(v, j, reinterpret_cast<double*>(stage), A, B);  // A = (uq)_h, B = (vq)_h
        cplx* drow = stage;
        cplx* vrow = stage + kop;
#pragma unroll
        for (int r = 0; r < 9; ++r) {
          const int k = j + r * T;
          if ((r < 8 || j == 0) && k < a.kout) {
            const double kr = k * a.dkx;
            drow[k] = cmake(kr * A[r].y, -kr * A[r].x);  // -i kr (uq)_h
            vrow[k] = B[r];
          }
        }
        fence_proxy_async();
        bar_sync(WS_BAR_HELP, WS_FFT_THREADS);
        if (j == 0) {
          for (int b = 0; b < tm.nbox_out; ++b) {
            tma_store_5d(&tm.p, drow + 2 * b * tm.boxp_out, 0, y & 7, b * tm.boxp_out, 2 * layer, y >> 3);
            tma_store_5d(&tm.p, vrow + 2 * b * tm.boxp_out, 0, y & 7, b * tm.boxp_out, 2 * layer + 1, y >> 3);
          }
          tma_commit();
        }
      }
    }
    if (j == 0) tma_wait_all<0>();  // every result row is in P before the CTA (and its shared memory) goes away
  } else {
    // ------------------------------------------ FFT group ------------------------------------------------------------
    cplx v[PT];
    const Tw4096 twr = load_tw4096(j, tw);
    int nout = 0;
#pragma unroll 1
    for (int m = 0; m < nphases; ++m) {
      const int s = m & 1, ph = m % 3;
      bar_sync(WS_BAR_FULL + s, WS_THREADS);
      tmem_fence_after_sync();
      tmem_load16(tmb + 64 * s, v);
      tmem_wait_ld();
      if (m + 2 < nphases) {
        tmem_fence_before_sync();
        bar_arrive(WS_BAR_EMPTY + s, WS_THREADS);
      }
      block_fft4096<INV, WS_BAR_FFT>(v, j, xb, twr);
      if (ph == 0) {  // q of both rows (+ eta) -> TMEM
        if (a.eta != nullptr) {
          const int unit = (int)blockIdx.x + (m / 3) * (int)gridDim.x;
          const int layer = unit / npairs, ya = 2 * (unit - layer * npairs);
          const double* ea = a.eta + (long long)ya * NX;
          const double* eb = ea + NX;
#pragma unroll
          for (int r = 0; r < PT; ++r) {
            v[r].x += ea[j + r * T];
            v[r].y += eb[j + r * T];
          }
        }
#pragma unroll
        for (int r = 0; r < PT; ++r) tmem_store(t_q + 4 * r, v[r]);
        tmem_wait_st();
        continue;
      }
#pragma unroll
      for (int c0 = 0; c0 < PT; c0 += 4) {
        cplx q2[4];
        tmem_load4(t_q + 4 * c0, q2);
        tmem_wait_ld();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double qq = (ph == 1) ? q2[q].x : q2[q].y;
          v[c0 + q].x *= qq;
          v[c0 + q].y *= qq;
        }
      }
      block_fft4096<FWD, WS_BAR_FFT>(v, j, xb, twr);
      if (nout > 0) {  // the helper group has taken the previous spectrum out of the out-slot
        bar_sync(WS_BAR_OUT_EMPTY, WS_THREADS);
        tmem_fence_after_sync();
      }
#pragma unroll
      for (int r = 0; r < PT; ++r) tmem_store(t_out + 4 * r, v[r]);
      tmem_wait_st();
      tmem_fence_before_sync();
      bar_arrive(WS_BAR_OUT_FULL, WS_THREADS);
      ++nout;
    }
  }
  tmem_release(tbase, 512);
}

}  // namespace gfb
