// k1_col_inv_ws.cuh -- K1 for the single-layer modules at ny = 4096, warp-specialised (k_col_inv_ws)
// (part of kernels.cuh: include that, not this file)
#pragma once
#include "k1_col_inv.cuh"

namespace gfb {

// ------------------------------------------------------------------------------------------------
// Warp-specialised column kernels: one CTA per SM, 512 threads = an FFT group (warps 0..7) + a helper group (warps
// 8..15).  The FFT group does nothing but transforms: it takes a prepared column out of its TMEM input slot, transforms
// it (exchanges through shared memory, synchronised by a named barrier over its own 256 threads) and stores / parks the
// result.  The helper group does all the data movement and pointwise arithmetic around the transforms -- here: the
// global loads of the raw state and of G, the dealias mask, the inversion and the i*k factors -- and runs AHEAD of the
// FFT group through a double-buffered hand-off, so global-memory latency never sits on the transforms' critical path
// (in the round-1 kernels the load / prepare / store phases and the transform phases of a CTA alternated, and two CTAs
// per SM overlapped them only by chance: measured time = sum of the phases).
//
// Hand-off = Tensor Memory.  A TMEM lane belongs to a warp quadrant (warp % 4), so FFT warp w and helper warp w + 8 see
// the same 32 lanes: helper thread 256 + j writes the 16 points of FFT thread j into "its" lane with tcgen05.st, FFT
// thread j reads them with tcgen05.ld -- no shared memory, no LSU traffic, no bank conflicts.  Synchronisation: named
// barriers, producer bar.arrive / consumer bar.sync, one full/empty pair per slot, tcgen05 fences on both sides.
//
// TMEM map per FFT thread (columns, 4 per complex value): [in slot 0: 64][in slot 1: 64][parked column A: 64]; the two
// FFT warps of a quadrant use column ranges 0..255 / 256..511.
// ------------------------------------------------------------------------------------------------
constexpr int WS_THREADS = 512;      // 8 FFT warps + 8 helper warps
constexpr int WS_FFT_THREADS = 256;  // = T of a 4096-point line
constexpr int WS_BAR_FFT = 1;        // named barrier of the FFT group's exchanges
constexpr int WS_BAR_FULL = 2;       // + slot (2, 3): helper -> FFT "slot holds a column"
constexpr int WS_BAR_EMPTY = 4;      // + slot (4, 5): FFT -> helper "slot is free again"

// TMEM address of this thread's private columns: lane quadrant of the warp, column range of the warp pair
__device__ __forceinline__ uint32_t ws_tmem_base(uint32_t base) {
  const int warp = (int)threadIdx.x / 32;
  return base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(((warp >> 2) & 1) * 256);
}

__device__ __forceinline__ cpair zero_pair_k1() {
  cpair z;
  z.a = z.b = cmake(0.0, 0.0);
  return z;
}

// VAR: experiment switches (tools/, GFB_K1WS_VAR): bit 0 = twiddles in registers, bit 1 = L2 bulk prefetch of the next
// unit, bit 2 = all 16 rows in one load batch (else two batches of 8), bit 3 = .x64 TMEM load of the input slot,
// bit 4 = the helper group reads whole sectors and prepares both columns of a unit in one pass
template <int NY, int VAR>
__global__ void __launch_bounds__(WS_THREADS, 1)
    k_col_inv_ws(const __grid_constant__ GridDesc g, const __grid_constant__ PrepSingle prep, const __grid_constant__ PeerOut W,
                 const int ldw, const long long wfs, const int ncols, const int nfields, const int nunits,
                 const cplx* __restrict__ tw, const int stagger_ns) {
  constexpr int T = NY / PT;
  static_assert(T == WS_FFT_THREADS, "one 256-thread FFT group: ny = 4096");
  cta_stagger(stagger_ns);
  GFB_DYN_SMEM(double, smd);
  cplx* xb = reinterpret_cast<cplx*>(smd);  // un-split exchange buffer of the FFT group
  uint32_t* tmem_word = reinterpret_cast<uint32_t*>(xb + xbuf_elems(NY));
  const int tid = threadIdx.x;
  const bool is_fft = tid < WS_FFT_THREADS;
  const int j = is_fft ? tid : tid - WS_FFT_THREADS;
  const uint32_t tbase = tmem_setup(tmem_word, 512);
  const uint32_t tm = ws_tmem_base(tbase);
  const uint32_t t_in[2] = {tm, tm + 64};  // (indexed with b = n & 1)
  const uint32_t t_park = tm + 128;
  // unit u = (column pair, field); this CTA transforms units blockIdx.x + i * gridDim.x, two columns each
  const int nmine = (nunits - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int ncolumns = 2 * nmine;

  if (!is_fft && (VAR & 16)) {
    // ---------------- helper group, pair form: BOTH columns of a unit from one pass over the rows -----------------------
    // The per-column form below reads every 32-byte sector twice (once per column, 16 bytes each time) and spends two
    // memory round trips per column; ncu showed the FFT group waiting on the helper more than half of the time.  Here a
    // thread reads whole sectors (state pair + G pair, eight rows = 96 registers in flight per batch), prepares the values
    // of both columns and streams them straight into the two input slots: half the load instructions, two round trips per
    // UNIT.  Same barrier protocol: one EMPTY wait and one FULL arrive per column.
#pragma unroll 1
    for (int u = 0; u < nmine; ++u) {
      const int unit = (int)blockIdx.x + u * (int)gridDim.x;
      const int field = unit % nfields, colA = 2 * (unit / nfields);
      const bool liveA = colA < ncols, liveB = colA + 1 < ncols;
      const double krA = wav_kr(g, colA), krB = wav_kr(g, colA + 1);
      const int off0 = toff(j, colA, prep.ld);
#pragma unroll
      for (int h = 0; h < PT; h += 8) {
        cpair raw[8];
        double2 gg[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int r = h + q;
          const bool on = liveA && !l_aliased(g, j + r * T);
          const int off = off0 + r * T * prep.ld;  // T % 8 == 0: the tiled layout is affine in r
          raw[q] = on ? ldpair(prep.sol + off) : zero_pair_k1();
          gg[q] = (on && field != 2) ? ldpair_d(prep.G + off) : make_double2(0.0, 0.0);
        }
        if (h == 0 && u >= 1) {  // the FFT group has taken both columns of the previous unit out of the slots
          bar_sync(WS_BAR_EMPTY + 0, WS_THREADS);
          bar_sync(WS_BAR_EMPTY + 1, WS_THREADS);
          tmem_fence_after_sync();
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int r = h + q, l = j + r * T;
          const bool on = liveA && !l_aliased(g, l);
          const double ll = wav_l(g, l);
          const cplx ta = prep.apply(g, field, ll, krA, gg[q].x, raw[q].a);
          const cplx tb = prep.apply(g, field, ll, krB, gg[q].y, raw[q].b);
          tmem_store(t_in[0] + 4 * r, cmake(on ? ta.x : 0.0, on ? ta.y : 0.0));
          tmem_store(t_in[1] + 4 * r, cmake(on && liveB ? tb.x : 0.0, on && liveB ? tb.y : 0.0));
        }
      }
      tmem_wait_st();
      tmem_fence_before_sync();
      bar_arrive(WS_BAR_FULL + 0, WS_THREADS);
      bar_arrive(WS_BAR_FULL + 1, WS_THREADS);
    }
  } else if (!is_fft) {
    // ---------------- helper group: raw state + G -> masked, inverted, differentiated column -> TMEM slot -------------
#pragma unroll 1
    for (int n = 0; n < ncolumns; ++n) {
      const int unit = (int)blockIdx.x + (n >> 1) * (int)gridDim.x, which = n & 1, b = n & 1;
      const int field = unit % nfields, colA = 2 * (unit / nfields), col = colA + which;
      const bool live_col = col < ncols;
      const double kr = wav_kr(g, col);
      const int off0 = toff(j, colA, prep.ld) + which;
      // the NEXT unit's column pair (state + G tiles of this thread's live rows) starts moving DRAM -> L2 now: one
      // thread in eight issues, for the 256-byte tile its row shares with its seven neighbours
      if ((VAR & 2) && which == 0 && (n >> 1) + 1 < nmine && (j & 7) == 0) {
        const int un = unit + (int)gridDim.x, fn = un % nfields, cn = 2 * (un / nfields);
        const int pf0 = toff(j, cn, prep.ld);
#pragma unroll
        for (int r = 0; r < PT; ++r) {
          if (!l_aliased(g, j + r * T) || !l_aliased(g, j + r * T + 7)) {
            prefetch_l2_bulk(prep.sol + pf0 + r * T * prep.ld, 256);
            if (fn != 2) prefetch_l2_bulk(prep.G + pf0 + r * T * prep.ld, 128);
          }
        }
      }
      cplx val[PT];
      if (VAR & 4) {
        // all 16 rows in one batch: 16 x (16 + 8) bytes in flight per thread, one memory round trip per column
        double gg[PT];
#pragma unroll
        for (int r = 0; r < PT; ++r) {
          const bool on = live_col && !l_aliased(g, j + r * T);
          const int off = off0 + r * T * prep.ld;  // T % 8 == 0: the tiled layout is affine in r
          val[r] = on ? prep.sol[off] : cmake(0.0, 0.0);
          gg[r] = (on && field != 2) ? prep.G[off] : 0.0;
        }
#pragma unroll
        for (int r = 0; r < PT; ++r) {
          const int l = j + r * T;
          const bool on = live_col && !l_aliased(g, l);
          const cplx t = prep.apply(g, field, wav_l(g, l), kr, gg[r], val[r]);
          val[r] = cmake(on ? t.x : 0.0, on ? t.y : 0.0);
        }
      } else {
#pragma unroll
        for (int h = 0; h < PT; h += 8) {  // two batches of eight rows: 8 x (16 + 8) bytes in flight per thread
          cplx raw[8];
          double gg[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int r = h + q, l = j + r * T;
            const bool on = live_col && !l_aliased(g, l);
            const int off = off0 + r * T * prep.ld;  // T % 8 == 0: the tiled layout is affine in r
            raw[q] = on ? prep.sol[off] : cmake(0.0, 0.0);
            gg[q] = (on && field != 2) ? prep.G[off] : 0.0;
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int r = h + q, l = j + r * T;
            const bool on = live_col && !l_aliased(g, l);
            const cplx t = prep.apply(g, field, wav_l(g, l), kr, gg[q], raw[q]);
            val[r] = cmake(on ? t.x : 0.0, on ? t.y : 0.0);
          }
        }
      }
      if (n >= 2) {  // the FFT group has taken column n - 2 out of this slot
        bar_sync(WS_BAR_EMPTY + b, WS_THREADS);
        tmem_fence_after_sync();
      }
#pragma unroll
      for (int r = 0; r < PT; ++r) tmem_store(t_in[b] + 4 * r, val[r]);  // 16 x .x4: one .x64 store is 5x slower (tools/tmem_bench)
      tmem_wait_st();
      tmem_fence_before_sync();
      bar_arrive(WS_BAR_FULL + b, WS_THREADS);
    }
  } else {
    // ---------------- FFT group: slot -> registers -> transform -> park (column A) / store the pair (column B) --------
    cplx v[PT];
    const Tw4096 twr = load_tw4096(j, tw);
#pragma unroll 1
    for (int n = 0; n < ncolumns; ++n) {
      const int unit = (int)blockIdx.x + (n >> 1) * (int)gridDim.x, which = n & 1, b = n & 1;
      bar_sync(WS_BAR_FULL + b, WS_THREADS);
      tmem_fence_after_sync();
      if (VAR & 8) {
        tmem_load16(t_in[b], v);
      } else {
#pragma unroll
        for (int r = 0; r < PT; ++r) v[r] = tmem_load(t_in[b] + 4 * r);
      }
      tmem_wait_ld();
      if (n + 2 < ncolumns) {  // hand the slot back (only if somebody will wait for it)
        tmem_fence_before_sync();
        bar_arrive(WS_BAR_EMPTY + b, WS_THREADS);
      }
      if (VAR & 1) block_fft4096<INV, WS_BAR_FFT>(v, j, xb, twr);
      else block_fft<NY, INV, WS_BAR_FFT>(v, j, xb, tw);
      if (which == 0) {  // column A waits in TMEM for its partner
#pragma unroll
        for (int r = 0; r < PT; ++r) tmem_store(t_park + 4 * r, v[r]);
        tmem_wait_st();
      } else {
        const int field = unit % nfields, colA = 2 * (unit / nfields);
        const int ooff = toff(j, colA, ldw);
#pragma unroll
        for (int c0 = 0; c0 < PT; c0 += 4) {
          cplx ra[4];
          tmem_load4(t_park + 4 * c0, ra);
          tmem_wait_ld();
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
  }
  tmem_release(tbase, 512);
}

}  // namespace gfb
