// k2_row.cuh -- K2: row pass (k_row plain transforms, k_row_product hot path)
// (part of kernels.cuh: include that, not this file)
#pragma once
#include "layout.cuh"

namespace gfb {

// ------------------------------------------------------------------------------------------------
// K2 row pass
// ------------------------------------------------------------------------------------------------
// W and P as the row kernel sees them: LOCAL rows, ALL (global) kr.  Both are tiled, row-block major with the
// fields interleaved per row block, and -- on several GPUs -- split into one chunk per peer rank, which is
// exactly the send/receive buffer of the all-to-all (no packing):
//   off(field, y, k) = (kp % nranks)*chunk + (y/8)*rbs + field*fstr + (kp / nranks)*16 + (y%8)*2 + k%2,  kp = k/2
struct RowLayout {
  int nranks;
  long long chunk;  // elements per peer chunk = (local rows / 8) * rbs
  int rbs;          // row-block stride = 8 * nfields * ldl
  int fstr;         // field stride inside a row block = 8 * ldl
  // off(field, y, k) = rowbase(field, y) + coloff(k): the row part is hoisted out of the per-column loops
  __device__ __forceinline__ long long rowbase(int field, int y) const {
    return (long long)(y >> 3) * rbs + field * fstr + ((y & 7) << 1);
  }
  __device__ __forceinline__ long long coloff(int k) const {
    const int kp = k >> 1;
    if (nranks == 1) return kp * 16 + (k & 1);
    return (kp & (nranks - 1)) * chunk + (kp / nranks) * 16 + (k & 1);  // nranks is a power of two
  }
  __device__ __forceinline__ long long off(int field, int y, int k) const { return rowbase(field, y) + coloff(k); }
};

// Z[k] = A[k] + i B[k] with Hermitian extension, A/B half-spectra rows known for k < kc (else 0).
// c2r semantics of FFTW/cuFFT/pocketfft: imaginary parts of the k = 0 and k = N/2 entries are ignored.
template <int N>
__device__ __forceinline__ void load_hermitian(cplx (&v)[PT], const int j, const cplx* __restrict__ W, const RowLayout& wl,
                                               const int fA, const int rowA, const int fB, const int rowB, const int kc) {
  constexpr int T = N / PT;
  const cplx* __restrict__ pa = W + wl.rowbase(fA, rowA);
  const cplx* __restrict__ pb = W + wl.rowbase(fB, rowB);
#pragma unroll
  for (int r = 0; r < PT; ++r) {
    const int k = j + r * T;
    const int kk = (k <= N / 2) ? k : N - k;
    cplx a = cmake(0.0, 0.0), b = cmake(0.0, 0.0);
    if (kk < kc) {
      const long long co = wl.coloff(kk);
      a = pa[co];
      b = pb[co];
    }
    if (k == 0 || k == N / 2) {
      v[r] = cmake(a.x, b.x);
    } else if (k < N / 2) {
      v[r] = cmake(a.x - b.y, a.y + b.x);
    } else {
      v[r] = cmake(a.x + b.y, b.x - a.y);
    }
  }
}

// After a forward FFT of z = a + i b (a, b real rows): thread holds Z[k], k = j + r*T.  Exchanges
// through shared memory (re then im) to fetch Z[N-k] and returns A[k], B[k] for r = 0..8 (r = 8 only valid
// for j == 0, k = N/2).
template <int N, int BAR = 0>
__device__ __forceinline__ void unpack_pair(const cplx (&v)[PT], const int j, double* __restrict__ xb,
                                            cplx (&A)[9], cplx (&B)[9]) {
  constexpr int T = N / PT;
  double wre[9];
#pragma unroll
  for (int r = 0; r < PT; ++r) xb[xpad(j + r * T)] = v[r].x;
  fft_sync<BAR, T>();
#pragma unroll
  for (int r = 0; r < 9; ++r) {
    const int m = (N - (j + r * T)) & (N - 1);
    wre[r] = (r < 8 || j == 0) ? xb[xpad(m)] : v[r].x;
  }
  fft_sync<BAR, T>();
#pragma unroll
  for (int r = 0; r < PT; ++r) xb[xpad(j + r * T)] = v[r].y;
  fft_sync<BAR, T>();
#pragma unroll
  for (int r = 0; r < 9; ++r) {
    const int m = (N - (j + r * T)) & (N - 1);
    const double wim = (r < 8 || j == 0) ? xb[xpad(m)] : v[r].y;
    const cplx o = v[r];
    A[r] = cmake(0.5 * (o.x + wre[r]), 0.5 * (o.y - wim));
    B[r] = cmake(0.5 * (o.y + wim), -0.5 * (o.x - wre[r]));
  }
  fft_sync<BAR, T>();
}

// address of P(row, k) given the row part of the offset: column pair kp belongs to rank kp % nranks
__device__ __forceinline__ cplx* prow_ptr(const PeerOut& po, const long long rowbase, const int k) {
  const int kp = k >> 1;
  if (po.nranks == 1) return po.base[0] + rowbase + kp * 16 + (k & 1);
  const int d = kp & (po.nranks - 1);
  return po.base[d] + po.rank * po.chunk + rowbase + (kp / po.nranks) * 16 + (k & 1);
}

enum RowMode { ROW_PRODUCT = 0, ROW_C2R = 1, ROW_R2C = 2, ROW_PRODUCT_Q = 3 };

struct RowArgs {
  const cplx* W;  // field = 3*layer + {u,v,q} (or the batch index for plain transforms)
  RowLayout wl;
  int kc;  // live input columns (kcut, or nkr for plain transforms)
  // outputs of the forward part: field = 2*layer + {d, v}
  PeerOut pout;  // P as the destination of the stores (see PeerOut)
  RowLayout pl;
  int kout;  // columns to write (nkr)
  // physical-space operands
  const double* eta;  // [y][x] or null              (SingleLayerQG: q + eta)
  const double* U;    // [layer][y] or null          (imposed zonal flow added to u)
  const double* Qx;   // [layer][y][x] or null
  const double* Qy;   // [layer][y][x] or null
  double* phys;       // ROW_C2R output / ROW_R2C input, [batch][y][x]
  double dkx;
  int ny;       // LOCAL rows handled by this launch
  int wstride;  // fields per layer in W (3, or 4 when updatevars! also wants psi; 1 for plain transforms)
  int fsel;     // ROW_C2R: which of the wstride fields of each layer to transform
  int linear;   // ROW_PRODUCT_Q: MultiLayerQG.calcNlinear! -- the self-advection (u q, v q) is dropped, U q stays
};

template <int NX, int MODE>
__global__ void __launch_bounds__(GroupCfg<NX>::THREADS, GroupCfg<NX>::MINB)
    k_row(const __grid_constant__ RowArgs a, const cplx* __restrict__ tw) {
  using RC = GroupCfg<NX>;
  constexpr int T = RC::T, THREADS = RC::THREADS;
  GFB_DYN_SMEM(double, smd);
  const int gi = threadIdx.x / T, j = threadIdx.x % T;
  const int unit = blockIdx.x * RC::G + gi;  // row pair
  const int layer = blockIdx.y;
  const int ya = 2 * unit, yb = ya + 1;
  const bool live = yb < a.ny;
  double* xb = smd + gi * RC::GSD;
  cplx* stash = reinterpret_cast<cplx*>(smd + RC::XB_DOUBLES);  // thread-private: stash[r*THREADS + tid]
  const int tid = threadIdx.x;
  cplx v[PT];
  cplx A[9], B[9];

  if (MODE == ROW_C2R) {
    const int fc = a.wstride * layer + a.fsel;
    if (live) load_hermitian<NX>(v, j, a.W, a.wl, fc, ya, fc, yb, a.kc);
    else {
#pragma unroll
      for (int r = 0; r < PT; ++r) v[r] = cmake(0.0, 0.0);
    }
    block_fft_split<NX, INV>(v, j, xb, tw);
    if (live) {
      double* oa = a.phys + ((long long)layer * a.ny + ya) * NX;
      double* ob = oa + NX;
#pragma unroll
      for (int r = 0; r < PT; ++r) {
        oa[j + r * T] = v[r].x;
        ob[j + r * T] = v[r].y;
      }
    }
    return;
  }
  if (MODE == ROW_R2C) {
    if (live) {
      const double* ia = a.phys + ((long long)layer * a.ny + ya) * NX;
      const double* ib = ia + NX;
#pragma unroll
      for (int r = 0; r < PT; ++r) v[r] = cmake(ia[j + r * T], ib[j + r * T]);
    } else {
#pragma unroll
      for (int r = 0; r < PT; ++r) v[r] = cmake(0.0, 0.0);
    }
    block_fft_split<NX, FWD>(v, j, xb, tw);
    unpack_pair<NX>(v, j, xb, A, B);
    if (live) {
      const long long ra = a.pl.rowbase(layer, ya), rb = a.pl.rowbase(layer, yb);
#pragma unroll
      for (int r = 0; r < 9; ++r) {
        const int k = j + r * T;
        if ((r < 8 || j == 0) && k < a.kout) {
          *prow_ptr(a.pout, ra, k) = A[r];
          *prow_ptr(a.pout, rb, k) = B[r];
        }
      }
    }
    return;
  }

  // the product modes (ROW_PRODUCT, ROW_PRODUCT_Q) live in k_row_product below
}

// ------------------------------------------------------------------------------------------------
// K2, product modes (the hot path): persistent, software-pipelined.
//
// A CTA loops over row pairs.  Every row pair has three input phases -- (q_a, q_b), (u_a, v_a), (u_b, v_b) -- and the
// two half-spectra rows of the NEXT phase are fetched with cp.async into a shared-memory line buffer while the
// transforms of the current phase run (the last phase of a pair prefetches the first phase of the CTA's next
// pair), so no thread ever waits on a global load in steady state.  q of both rows (and the (u+U)Qx + vQy
// accumulator) is parked in the thread's TMEM slot instead of shared memory, which is what makes room for the
// line buffer with two CTAs per SM.
// ------------------------------------------------------------------------------------------------
template <int N>
struct RowPipeCfg {
  using RC = GroupCfg<N>;
  static constexpr int TM_PER_WG(bool withq) { return withq ? 128 : 64; }               // TMEM columns per thread slot set
  static constexpr int TM_COLS(bool withq) { return (RC::THREADS / 128) * TM_PER_WG(withq); }  // 128 / 256 (256 / 512 at 512 threads)
};

// fetch the two half-spectrum rows of one phase into the group's line buffer: lb[0..kc) = A, lb[kcp..kcp+kc) = B
template <int T>
__device__ __forceinline__ void row_prefetch(cplx* __restrict__ lb, const int kcp, const cplx* __restrict__ W, const RowLayout& wl,
                                             const int fA, const int rowA, const int fB, const int rowB, const int kc, const int j) {
  // fully unrolled with one base address per row and immediate offsets: an LDGSTS holds its address registers until
  // the LSU has taken it, so a rolled loop that bumps them stalls on every iteration
  if (wl.nranks == 1) {
    const cplx* __restrict__ pa = W + wl.rowbase(fA, rowA) + (j >> 1) * 16 + (j & 1);
    const cplx* __restrict__ pb = W + wl.rowbase(fB, rowB) + (j >> 1) * 16 + (j & 1);
    cplx* __restrict__ da = lb + j;
    cplx* __restrict__ db = lb + kcp + j;
#pragma unroll
    for (int i = 0; i <= PT / 2; ++i)  // kc <= N/2 + 1 = 8 T + 1
      if (j + i * T < kc) {
        if (T % 2 == 0) {
          cp_async16_cg(da + i * T, pa + i * (T / 2) * 16);
          cp_async16_cg(db + i * T, pb + i * (T / 2) * 16);
        } else {
          const int k = j + i * T;
          cp_async16_cg(da + i * T, pa - ((j >> 1) * 16 + (j & 1)) + (k >> 1) * 16 + (k & 1));
          cp_async16_cg(db + i * T, pb - ((j >> 1) * 16 + (j & 1)) + (k >> 1) * 16 + (k & 1));
        }
      }
    return;
  }
  const cplx* __restrict__ pa = W + wl.rowbase(fA, rowA);
  const cplx* __restrict__ pb = W + wl.rowbase(fB, rowB);
#pragma unroll
  for (int i = 0; i <= PT / 2; ++i) {  // kc <= N/2 + 1 = 8 T + 1
    const int k = j + i * T;
    if (k < kc) {
      const long long co = wl.coloff(k);
      cp_async16_cg(lb + k, pa + co);
      cp_async16_cg(lb + kcp + k, pb + co);
    }
  }
}

// Z[k] = A[k] + i B[k] with Hermitian extension, from the line buffer (same semantics as load_hermitian)
template <int N>
__device__ __forceinline__ void assemble_hermitian(cplx (&v)[PT], const int j, const cplx* __restrict__ lb, const int kcp, const int kc) {
  constexpr int T = N / PT;
#pragma unroll
  for (int r = 0; r < PT; ++r) {
    const int k = j + r * T;
    const int kk = (k <= N / 2) ? k : N - k;
    cplx a = cmake(0.0, 0.0), b = cmake(0.0, 0.0);
    if (kk < kc) {
      a = lb[kk];
      b = lb[kcp + kk];
    }
    if (k == 0 || k == N / 2) {
      v[r] = cmake(a.x, b.x);
    } else if (k < N / 2) {
      v[r] = cmake(a.x - b.y, a.y + b.x);
    } else {
      v[r] = cmake(a.x + b.y, b.x - a.y);
    }
  }
}

template <int NX, bool WITHQ>
__global__ void __launch_bounds__(GroupCfg<NX>::THREADS, GroupCfg<NX>::MINB)
    k_row_product(const __grid_constant__ RowArgs a, const cplx* __restrict__ tw, const int nunits, const int nlayers, const int kcp) {
  using RC = GroupCfg<NX>;
  constexpr int T = RC::T, G = RC::G;
  GFB_DYN_SMEM(double, smd);
  const int tid = threadIdx.x;
  const int gi = tid / T, j = tid % T;
  double* xb = smd + gi * RC::GSD;
  cplx* lb = reinterpret_cast<cplx*>(smd + RC::XB_DOUBLES) + 1 + (size_t)gi * 2 * kcp;  // [0] holds the TMEM base word
  const int tmcols = RowPipeCfg<NX>::TM_COLS(WITHQ);
  const uint32_t tbase = tmem_setup(reinterpret_cast<uint32_t*>(smd + RC::XB_DOUBLES), tmcols);
  const uint32_t tq = tmem_thread_base(tbase, RowPipeCfg<NX>::TM_PER_WG(WITHQ));
  const uint32_t tg = tq + 64;
  const int npairs = a.ny / 2;  // row pairs per layer
  const int niter = (nunits + (int)gridDim.x * G - 1) / ((int)gridDim.x * G);
  cplx v[PT];
  cplx A[9], B[9];

  // unit -> (layer, first row); phase -> (fields, rows)
  auto unit_of = [&](int it) { return (it * (int)gridDim.x + (int)blockIdx.x) * G + gi; };
  auto prefetch = [&](int unit, int ph) {
    if (unit >= nunits) return;
    const int layer = unit / npairs, ya = 2 * (unit - layer * npairs);
    const int fu = a.wstride * layer, fv = fu + 1, fq = fu + 2;
    if (ph == 0) row_prefetch<T>(lb, kcp, a.W, a.wl, fq, ya, fq, ya + 1, a.kc, j);
    else row_prefetch<T>(lb, kcp, a.W, a.wl, fu, ya + ph - 1, fv, ya + ph - 1, a.kc, j);
  };
  prefetch(unit_of(0), 0);
#pragma unroll 1
  for (int it = 0; it < niter; ++it) {
    const int unit = unit_of(it);
    const bool live = unit < nunits;
    const int layer = live ? unit / npairs : 0;
    const int ya = live ? 2 * (unit - layer * npairs) : 0;
#pragma unroll 1
    for (int ph = 0; ph < 3; ++ph) {
      cp_async_wait_all();
      __syncthreads();  // the line buffer is complete and visible
      if (live) assemble_hermitian<NX>(v, j, lb, kcp, a.kc);
      else {
#pragma unroll
        for (int r = 0; r < PT; ++r) v[r] = cmake(0.0, 0.0);
      }
      __syncthreads();  // everyone has consumed the line buffer: refill it behind the transforms
      if (ph < 2) prefetch(unit, ph + 1);
      else if (it + 1 < niter) prefetch(unit_of(it + 1), 0);
      block_fft_split<NX, INV>(v, j, xb, tw);
      if (ph == 0) {  // q of both rows (+ eta) -> TMEM
        if (a.eta != nullptr && live) {
          const double* ea = a.eta + (long long)ya * NX;
          const double* eb = ea + NX;
#pragma unroll
          for (int r = 0; r < PT; ++r) {
            v[r].x += ea[j + r * T];
            v[r].y += eb[j + r * T];
          }
        }
#pragma unroll
        for (int r = 0; r < PT; ++r) tmem_store(tq + 4 * r, v[r]);
        tmem_wait_st();
        continue;
      }
      const int h = ph - 1, y = ya + h;
      if (WITHQ) {  // (u + U) Qx + v Qy of this row -> its half of the TMEM accumulator
        const double Uy = (a.U && live) ? a.U[(long long)layer * a.ny + y] : 0.0;
        const double* qx = a.Qx + ((long long)layer * a.ny + y) * NX;
        const double* qy = a.Qy + ((long long)layer * a.ny + y) * NX;
#pragma unroll
        for (int c0 = 0; c0 < PT; c0 += 4) {
          cplx gs[4];
          if (h == 1) {
#pragma unroll
            for (int q = 0; q < 4; ++q) gs[q] = tmem_load(tg + 4 * (c0 + q));
            tmem_wait_ld();
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int r = c0 + q;
            v[r].x += Uy;
            const double gq = live ? v[r].x * qx[j + r * T] + v[r].y * qy[j + r * T] : 0.0;
            if (h == 0) gs[q] = cmake(gq, 0.0);
            else gs[q].y = gq;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) tmem_store(tg + 4 * (c0 + q), gs[q]);
        }
        tmem_wait_st();
        if (a.linear) {  // calcN_linearadvection! (src/multilayerqg.jl:744-779): only -d/dx (U q) of the q-advection remains
#pragma unroll
          for (int r = 0; r < PT; ++r) v[r] = cmake(Uy, 0.0);
        }
      }
#pragma unroll
      for (int c0 = 0; c0 < PT; c0 += 4) {
        cplx q2[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) q2[q] = tmem_load(tq + 4 * (c0 + q));
        tmem_wait_ld();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double qq = (h == 0) ? q2[q].x : q2[q].y;
          v[c0 + q].x *= qq;
          v[c0 + q].y *= qq;
        }
      }
      block_fft_split<NX, FWD>(v, j, xb, tw);
      unpack_pair<NX>(v, j, xb, A, B);  // A = (uq)_h, B = (vq)_h
      if (live) {
        const long long rd = a.pl.rowbase(2 * layer, y), rv = a.pl.rowbase(2 * layer + 1, y);
#pragma unroll
        for (int r = 0; r < 9; ++r) {
          const int k = j + r * T;
          if ((r < 8 || j == 0) && k < a.kout) {
            const double kr = k * a.dkx;
            cplx* qd = prow_ptr(a.pout, rd, k);
            qd[0] = cmake(kr * A[r].y, -kr * A[r].x);  // -i kr (uq)_h
            qd[rv - rd] = B[r];                         // same row block, next field
          }
        }
      }
    }
    if (WITHQ) {  // the Q-gradient term of both rows with one more transform: d -= ((u+U) Qx + v Qy)_h
#pragma unroll
      for (int c0 = 0; c0 < PT; c0 += 4) {
#pragma unroll
        for (int q = 0; q < 4; ++q) v[c0 + q] = tmem_load(tg + 4 * (c0 + q));
        tmem_wait_ld();
      }
      block_fft_split<NX, FWD>(v, j, xb, tw);
      unpack_pair<NX>(v, j, xb, A, B);
      if (live) {
#pragma unroll
        for (int r = 0; r < 9; ++r) {
          const int k = j + r * T;
          if ((r < 8 || j == 0) && k < a.kout) {
            cplx& ea = *prow_ptr(a.pout, a.pl.rowbase(2 * layer, ya), k);  // same thread wrote these above
            cplx& eb = *prow_ptr(a.pout, a.pl.rowbase(2 * layer, ya + 1), k);
            ea = csub(ea, A[r]);
            eb = csub(eb, B[r]);
          }
        }
      }
    }
  }
  tmem_release(tbase, tmcols);
}

}  // namespace gfb

namespace gfb {

// ------------------------------------------------------------------------------------------------
// K2, product mode of the single-layer modules, TMA-staged (single GPU, one group per CTA: NX >= 4096).
//
// Same arithmetic and phase structure as k_row_product, but no thread ever issues a global load or store for the
// spectral rows:
//   * the two half-spectrum rows of the next phase are gathered out of the 8-row tiled W by TMA tensor loads
//     (cp.async.bulk.tensor.5d, box = 1 row x <= 256 column pairs of 32 bytes) issued by ONE thread and signalled on
//     an mbarrier -- instead of 22 LDGSTS per thread and phase, each of which cost the LSU 16 sector wavefronts;
//   * the two result rows of a phase are assembled densely in shared memory (the d row in the FFT exchange buffer,
//     which is idle between the unpack and the next transform; the v row behind the line buffer) and scattered into
//     the tiled P by TMA tensor stores -- instead of 18 STG.128 per thread and phase hitting 16 lines each.
// Geometry of a staged row: nbox boxes of boxp column pairs (boxp a multiple of 4, so every box starts 128-byte
// aligned in shared memory); columns beyond the live ones inside the last box are real (aliased) columns of the
// tensor or out of bounds (zero-filled on load, dropped on store) and never used.
// ------------------------------------------------------------------------------------------------
struct RowTma {
  TmaDesc w, p;            // tensor maps of W (3 fields per layer) and P (2 fields per layer); gfb.cu make_row_tma
  int nbox_in, boxp_in;    // input rows: boxes per row, column pairs per box
  int nbox_out, boxp_out;  // output rows (depends on kout)
};

template <int NX>
__global__ void __launch_bounds__(GroupCfg<NX>::THREADS, GroupCfg<NX>::MINB)
    k_row_product_tma(const __grid_constant__ RowArgs a, const __grid_constant__ RowTma tm, const cplx* __restrict__ tw, const int nunits,
                      const int stagger_ns) {
  cta_stagger(stagger_ns);
  using RC = GroupCfg<NX>;
  constexpr int T = RC::T;
  static_assert(RC::G == 1, "one FFT group per CTA");
  GFB_DYN_SMEM(double, smd);
  const int j = threadIdx.x;
  double* xb = smd;
  const int kcp = 2 * tm.nbox_in * tm.boxp_in;  // complex elements per staged input row
  const int kop = 2 * tm.nbox_out * tm.boxp_out;
  // [xb | 128 B: TMEM base word, mbarrier | line buffer: 2 input rows | v row staging]
  unsigned char* ctl = reinterpret_cast<unsigned char*>(smd + RC::XB_DOUBLES);
  Mbar* full = reinterpret_cast<Mbar*>(ctl + 16);
  cplx* lb = reinterpret_cast<cplx*>(ctl + 128);
  cplx* vst = lb + 2 * (size_t)kcp;
  cplx* dst = reinterpret_cast<cplx*>(xb);  // d row staging: kop complex <= XB_DOUBLES / 2 (checked by the launcher)
  (void)kop;
  const int tmcols = RowPipeCfg<NX>::TM_COLS(false);
  if (j == 0) {
    mbar_init(full, 1);
    mbar_fence_init();
  }
  const uint32_t tbase = tmem_setup(reinterpret_cast<uint32_t*>(ctl), tmcols);  // contains the __syncthreads
  const uint32_t tq = tmem_thread_base(tbase, RowPipeCfg<NX>::TM_PER_WG(false));
  const int npairs = a.ny / 2;  // row pairs per layer
  const int niter = (nunits - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  cplx v[PT];
  cplx A[9], B[9];
  unsigned parity = 0;

  // one elected thread gathers the two rows of a phase out of W into the line buffer
  auto prefetch = [&](int unit, int ph) {
    if (j != 0) return;
    const int layer = unit / npairs, ya = 2 * (unit - layer * npairs);
    const int fu = a.wstride * layer, fv = fu + 1, fq = fu + 2;
    const int fA = ph == 0 ? fq : fu, fB = ph == 0 ? fq : fv;
    const int rowA = ph == 0 ? ya : ya + ph - 1, rowB = ph == 0 ? ya + 1 : ya + ph - 1;
    mbar_expect_tx(full, (unsigned)(2 * kcp * sizeof(cplx)));
    for (int b = 0; b < tm.nbox_in; ++b) {
      tma_load_5d(lb + 2 * b * tm.boxp_in, &tm.w, full, 0, rowA & 7, b * tm.boxp_in, fA, rowA >> 3);
      tma_load_5d(lb + kcp + 2 * b * tm.boxp_in, &tm.w, full, 0, rowB & 7, b * tm.boxp_in, fB, rowB >> 3);
    }
  };
  if (niter > 0) prefetch((int)blockIdx.x, 0);
#pragma unroll 1
  for (int it = 0; it < niter; ++it) {
    const int unit = (int)blockIdx.x + it * (int)gridDim.x;
    const int layer = unit / npairs;
    const int ya = 2 * (unit - layer * npairs);
#pragma unroll 1
    for (int ph = 0; ph < 3; ++ph) {
      mbar_wait(full, parity);  // the line buffer holds this phase's two rows
      parity ^= 1u;
      assemble_hermitian<NX>(v, j, lb, kcp, a.kc);
      if (j == 0) tma_wait_read<0>();  // the result rows of the previous phase have left the staging buffers (xb is one)
      __syncthreads();                 // everyone has consumed the line buffer: refill it behind the transforms
      if (ph < 2) prefetch(unit, ph + 1);
      else if (it + 1 < niter) prefetch(unit + (int)gridDim.x, 0);
      block_fft_split<NX, INV>(v, j, xb, tw);
      if (ph == 0) {  // q of both rows (+ eta) -> TMEM
        if (a.eta != nullptr) {
          const double* ea = a.eta + (long long)ya * NX;
          const double* eb = ea + NX;
#pragma unroll
          for (int r = 0; r < PT; ++r) {
            v[r].x += ea[j + r * T];
            v[r].y += eb[j + r * T];
          }
        }
#pragma unroll
        for (int r = 0; r < PT; ++r) tmem_store(tq + 4 * r, v[r]);
        tmem_wait_st();
        continue;
      }
      const int h = ph - 1, y = ya + h;
#pragma unroll
      for (int c0 = 0; c0 < PT; c0 += 4) {
        cplx q2[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) q2[q] = tmem_load(tq + 4 * (c0 + q));
        tmem_wait_ld();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double qq = (h == 0) ? q2[q].x : q2[q].y;
          v[c0 + q].x *= qq;
          v[c0 + q].y *= qq;
        }
      }
      block_fft_split<NX, FWD>(v, j, xb, tw);
      unpack_pair<NX>(v, j, xb, A, B);  // A = (uq)_h, B = (vq)_h ; ends with a barrier: xb is free
      // result rows, dense, into the staging buffers: d = -i kr (uq)_h in xb, v = (vq)_h behind the line buffer
#pragma unroll
      for (int r = 0; r < 9; ++r) {
        const int k = j + r * T;
        if ((r < 8 || j == 0) && k < a.kout) {
          const double kr = k * a.dkx;
          dst[k] = cmake(kr * A[r].y, -kr * A[r].x);
          vst[k] = B[r];
        }
      }
      fence_proxy_async();  // generic-proxy writes above -> visible to the TMA engine
      __syncthreads();
      if (j == 0) {
        for (int b = 0; b < tm.nbox_out; ++b) {
          tma_store_5d(&tm.p, dst + 2 * b * tm.boxp_out, 0, y & 7, b * tm.boxp_out, 2 * layer, y >> 3);
          tma_store_5d(&tm.p, vst + 2 * b * tm.boxp_out, 0, y & 7, b * tm.boxp_out, 2 * layer + 1, y >> 3);
        }
        tma_commit();
      }
    }
  }
  if (j == 0) tma_wait_all<0>();  // every result row is in P before the CTA (and its shared memory) goes away
  tmem_release(tbase, tmcols);
}

}  // namespace gfb
