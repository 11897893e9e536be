// kernels.cuh -- the three fused passes of one calcN! + stage update, and the plain 2-D transforms.
//
//   K1  spectral pre-pass (dealias mask on load, psi inversion / S^-1 matvec, i*k derivatives, 1/(nx*ny)) fused
//       with the inverse FFT along y (columns)
//         k_col_inv_single   single-layer modules, the hot path: persistent, rolling cp.async prefetch, TMEM stash
//         k_col_inv          MultiLayerQG pre-pass and plain inverse transforms
//         k_col_inv16k       ny = 16384: 4 x 4096 by a radix-4 DIF step while loading
//   K2  Hermitian c2r along x of (u, v) and q, physical products, r2c along x -- physical fields never touch HBM
//         k_row_product      the hot path: persistent, line-buffer prefetch of the next phase, q parked in TMEM
//         k_row              plain c2r / r2c (ldiv! / mul! seams, updatevars!)
//         k_row_product_half / k_row_half   nx = 16384: every real row as one 8192-point half-length transform
//   K3  forward FFT along y of the two product fields fused with the divergence (-i kr, -i l), forcing and the
//       time-stepper stage update
//         k_col_fwd          one transform call site over the four columns, TMEM-parked partners, fused epilogue
//         k_col_fwd16k (+ _plain)   ny = 16384
//
// Reference sequence being replaced: calcN_advection! /root/reference/src/twodnavierstokes.jl:253-275,
// src/surfaceqg.jl:257-278, src/singlelayerqg.jl:356-432, src/multilayerqg.jl:695-731; dealias!
// (FourierFlows) called at src/twodnavierstokes.jl:287; stage updates of FourierFlows' RK4/ETDRK4/...
// steppers.
//
// Layouts: spectral `cplx a[layer][l][kr]` (kr contiguous), physical
// `double a[layer][y][x]`.  Work arrays between the passes: W[field][y][kr] (ld = ldw) and
// P[field][y][kr] (ld = ldp), ctx-owned.
#pragma once
#include "fft_core.cuh"

namespace gfb {

struct GridDesc {
  int nx, ny, nkr; // global grid (nkr = nx/2 + 1)
  int kcut;        // dealias!: kr columns [kcut, nkr) are zeroed (global index)
  int la0, la1;    // dealias!: l rows [la0, la1) are zeroed
  double dkx, dky; // kr = m*dkx ; l = (n < ny/2 ? n : n - ny)*dky   (FourierFlows TwoDGrid)
  double invN;     // 1/(nx*ny): ldiv!(., rfftplan, .) normalisation
  // kr-sharding of the spectral side (multi-GPU): column PAIRS are dealt round-robin, local pair p of rank r
  // is global pair p*nranks + r (balanced for the live AND the aliased columns).  Single GPU: rank 0 of 1.
  int rank, nranks;
  int nkl;         // local dense columns (= nkr on one GPU)
};

// global kr index of local column `col`
__device__ __forceinline__ int gcol(const GridDesc& g, int col) { return ((col >> 1) * g.nranks + g.rank) * 2 + (col & 1); }

__device__ __forceinline__ double wav_l(const GridDesc& g, int n) { return (n < g.ny / 2 ? n : n - g.ny) * g.dky; }
__device__ __forceinline__ double wav_kr(const GridDesc& g, int m) { return gcol(g, m) * g.dkx; }  // m: LOCAL column
__device__ __forceinline__ bool l_aliased(const GridDesc& g, int n) { return n >= g.la0 && n < g.la1; }

// Ctx-owned arrays (states, tables, W, P) use an 8-row TILED layout: element (row, col) of an array with
// even leading dimension ld lives at
//     toff = ((row/8)*(ld/2) + col/2)*16 + (row%8)*2 + col%2
// i.e. [row/8][column pair][row%8][2].  Same footprint as row-major [rows][ld] (rows % 8 == 0), but a column
// kernel whose lanes hold consecutive rows of one column pair now touches 256 contiguous bytes per 8 lanes
// (DRAM-page friendly even when CTAs drift apart), while the row kernel still reads/writes whole 32-byte
// sectors.  User arrays at the C ABI stay dense row-major; k_repack converts.
// (32-bit: one field of the largest supported grid, 8192 x 4098, has 3.4e7 elements)
__host__ __device__ __forceinline__ int toff(int row, int col, int ld) {
  return ((row >> 3) * (ld >> 1) + (col >> 1)) * 16 + ((row & 7) << 1) + (col & 1);
}

// offset of row j + r*T given the offset of row j: when T is a multiple of 8 the tiled layout is affine in r
// with the same stride as a row-major array (T*ld) -- one base register plus immediates
template <int T>
__device__ __forceinline__ int toff_row(const int off_j, const int j, const int r, const int col, const int ld) {
  if (T % 8 == 0) return off_j + r * T * ld;
  return toff(j + r * T, col, ld);
}

// a ctx-owned spectral array view: element (layer, l, kr) = p[layer*ls + toff(l, kr, ld)]
struct Spec {
  cplx* p;
  int ld;
  long long ls;
  __device__ __forceinline__ cplx& at(int layer, int l, int kr) const { return p[layer * ls + toff(l, kr, ld)]; }
};
struct SpecR {  // real-valued table with the same indexing
  const double* p;
  int ld;
  long long ls;
  __device__ __forceinline__ double at(int layer, int l, int kr) const { return p[layer * ls + toff(l, kr, ld)]; }
};

// Destination of the W (K1) and P (K2) stores.  On several GPUs both buffers are split into one chunk per
// rank, and the chunk a store falls into names the rank that consumes it (K1: the owner of row y; K2: the owner
// of column pair kp).  `base[d]` is where rank d's chunk lives, pre-biased by the host so that
//     address = base[d] + rank*chunk + (offset inside the chunk)
// holds for both transports: "nccl" -- base[d] = send buffer + (d - rank)*chunk, i.e. chunk d of the local send
// buffer, exchanged afterwards by an all-to-all; "p2p" -- base[d] = rank d's RECEIVE buffer mapped into this
// process (CUDA IPC over NVLink): the kernel's own stores are the all-to-all, no collective is launched.
// One GPU: nranks = 1, base[0] = the ctx-owned array.
constexpr int MAX_RANKS = 8;
struct PeerOut {
  cplx* base[MAX_RANKS];
  long long chunk;  // elements per chunk
  int rank, nranks;
  int lg_rows;      // log2(rows per rank): K1 finds the destination of row y as y >> lg_rows
};

// One "FFT group" = T = N/16 threads transforming one length-N line at a time.  A CTA has G groups
// (256 threads, or T when T > 256).  Shared memory per CTA: G split-exchange buffers of N(+pad) doubles
// and thread-private stashes of PT complex per thread (`stash[r*THREADS + tid]`, conflict-free, no
// barrier needed) -- 34 KB + 64 KB at N = 4096, so two CTAs share an SM and one CTA's global-memory
// phases overlap the other's butterflies.
template <int N>
struct GroupCfg {
  static constexpr int T = N / PT;
  static constexpr int THREADS = (T >= 256) ? T : 256;
  static constexpr int G = THREADS / T;
  static constexpr int GSD = (xbuf_elems(N) + 1) & ~1;  // doubles per group buffer; 17*T keeps groups in distinct banks
  static constexpr int XB_DOUBLES = G * GSD;
  static constexpr int STASH = PT * THREADS;           // complex elements per stash
  static constexpr int MINB = (THREADS == 256) ? 2 : 1;
};

__device__ __forceinline__ void swap_stash(cplx (&v)[PT], cplx* __restrict__ X, const int tid, const int threads) {
#pragma unroll
  for (int r = 0; r < PT; ++r) {
    const cplx t = X[r * threads + tid];
    X[r * threads + tid] = v[r];
    v[r] = t;
  }
}

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
    } else if (unit < nunits) {
      const int field = unit % nfields, colA = 2 * (unit / nfields);
      const int ooff = toff(j, colA, ldw);
#pragma unroll
      for (int c0 = 0; c0 < PT; c0 += 4) {
        cplx ra[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) ra[q] = tmem_load(tA + 4 * (c0 + q));
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
    cp_async_wait_all();
    k1_take<T, THREADS>(v, S, SG, prep, g, unit0 + ((ci + 1) >> 1) * stride, (ci + 1) & 1, nunits, nfields, ncols, nslots, j, tid);
    k1_issue<T, THREADS>(S, SG, prep, g, unit0 + ((ci + 2) >> 1) * stride, ci & 1, nunits, nfields, ncols, j, tid);
  }
  tmem_release(tbase, tmcols);
}

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
template <int N>
__device__ __forceinline__ void unpack_pair(const cplx (&v)[PT], const int j, double* __restrict__ xb,
                                            cplx (&A)[9], cplx (&B)[9]) {
  constexpr int T = N / PT;
  double wre[9];
#pragma unroll
  for (int r = 0; r < PT; ++r) xb[xpad(j + r * T)] = v[r].x;
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 9; ++r) {
    const int m = (N - (j + r * T)) & (N - 1);
    wre[r] = (r < 8 || j == 0) ? xb[xpad(m)] : v[r].x;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < PT; ++r) xb[xpad(j + r * T)] = v[r].y;
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 9; ++r) {
    const int m = (N - (j + r * T)) & (N - 1);
    const double wim = (r < 8 || j == 0) ? xb[xpad(m)] : v[r].y;
    const cplx o = v[r];
    A[r] = cmake(0.5 * (o.x + wre[r]), 0.5 * (o.y - wim));
    B[r] = cmake(0.5 * (o.y + wim), -0.5 * (o.x - wre[r]));
  }
  __syncthreads();
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

// ------------------------------------------------------------------------------------------------
// K2 for rows of 2*NH points (nx = 16384 with NH = 8192): every real row is ONE half-length complex transform.
//
// A 2*NH-point complex line does not fit the register file, but a real row does not need one:
//   forward  z[n] = x[2n] + i x[2n+1],  Z = FFT_NH(z),  E = (Z[k] + conj Z[NH-k])/2,  O = (Z[k] - conj Z[NH-k])/(2i),
//            X[k] = E + W_N^k O   and   X[NH-k] = conj(E - W_N^k O)            (both from the SAME thread's E, O)
//   inverse  Zin[k] = (X[k] + conj X[NH-k]) + i W_N^{-k} (X[k] - conj X[NH-k]),  z = backward FFT_NH(Zin)
// (FFTW's c2r ignores Im X[0] and Im X[N/2]; so does the load below).  The two-for-one split E, O is the existing
// unpack_pair; the inverse pre-pass reads the half-spectrum row from the line buffer with a mirrored index, so
// nothing new crosses threads.  Same persistent, prefetched, TMEM-stashed structure as k_row_product, one ROW per
// unit and three phases (q, u, v): 5 transforms of NH per row = the flops of 5 transforms of 2*NH per row pair.
// A thread owns the physical points x = 2n, 2n+1 (n = j + r*T), so the products stay thread-local.
// ------------------------------------------------------------------------------------------------
template <int T>
__device__ __forceinline__ void row_prefetch1(cplx* __restrict__ lb, const cplx* __restrict__ W, const RowLayout& wl, const int f,
                                              const int row, const int kc, const int j) {
  const cplx* __restrict__ p = W + wl.rowbase(f, row);
  if (wl.nranks == 1) {
    const cplx* __restrict__ pa = p + (j >> 1) * 16 + (j & 1);
    cplx* __restrict__ da = lb + j;
#pragma unroll
    for (int i = 0; i <= PT; ++i)  // kc <= 16 T + 1
      if (j + i * T < kc) cp_async16_cg(da + i * T, pa + i * (T / 2) * 16);
    return;
  }
#pragma unroll
  for (int i = 0; i <= PT; ++i) {
    const int k = j + i * T;
    if (k < kc) cp_async16_cg(lb + k, p + wl.coloff(k));
  }
}

// Zin[k], k = j + r*T, from the half-spectrum row in the line buffer; tw2[i] = W_{2NH}^i
template <int NH>
__device__ __forceinline__ void assemble_half(cplx (&v)[PT], const int j, const cplx* __restrict__ lb, const int kc,
                                              const cplx* __restrict__ tw2) {
  constexpr int T = NH / PT;
#pragma unroll
  for (int r = 0; r < PT; ++r) {
    const int k = j + r * T, km = NH - k;
    cplx a = cmake(0.0, 0.0), b = cmake(0.0, 0.0);
    if (k < kc) a = lb[k];
    if (km < kc) b = lb[km];
    if (k == 0) {  // c2r semantics: the imaginary parts of X[0] and X[NH] are ignored
      a.y = 0.0;
      b.y = 0.0;
    }
    const cplx s = cmake(a.x + b.x, a.y - b.y), d = cmake(a.x - b.x, a.y + b.y);  // a + conj b, a - conj b
    const cplx w = cconj(__ldg(&tw2[k]));
    const cplx t = cmul(w, d);
    v[r] = cmake(s.x - t.y, s.y + t.x);  // s + i t
  }
}

template <int NH>
__global__ void __launch_bounds__(GroupCfg<NH>::THREADS, GroupCfg<NH>::MINB)
    k_row_product_half(const __grid_constant__ RowArgs a, const cplx* __restrict__ tw, const cplx* __restrict__ tw2, const int nunits,
                       const int kcp) {
  using RC = GroupCfg<NH>;
  constexpr int T = RC::T, G = RC::G;
  GFB_DYN_SMEM(double, smd);
  const int tid = threadIdx.x;
  const int gi = tid / T, j = tid % T;
  double* xb = smd + gi * RC::GSD;
  cplx* lb = reinterpret_cast<cplx*>(smd + RC::XB_DOUBLES) + 1 + (size_t)gi * kcp;  // [0] holds the TMEM base word
  constexpr int tmcols = (RC::THREADS / 128) * 64;
  const uint32_t tbase = tmem_setup(reinterpret_cast<uint32_t*>(smd + RC::XB_DOUBLES), tmcols);
  const uint32_t tq = tmem_thread_base(tbase, 64);
  const int stride = (int)gridDim.x * G, unit0 = (int)blockIdx.x * G + gi;
  const int niter = (nunits - (int)blockIdx.x * G + stride - 1) / stride;  // same for every group of the CTA
  cplx v[PT];
  cplx A[9], B[9];
  // phase p of unit u: field p == 0 ? q : (p == 1 ? u : v) of row (u % ny) of layer (u / ny)
  if (unit0 < nunits) row_prefetch1<T>(lb, a.W, a.wl, a.wstride * (unit0 / a.ny) + 2, unit0 % a.ny, a.kc, j);
#pragma unroll 1
  for (int it = 0; it < niter; ++it) {
    const int unit = unit0 + it * stride;
    const bool live = unit < nunits;
    const int layer = live ? unit / a.ny : 0, y = live ? unit % a.ny : 0;
#pragma unroll 1
    for (int ph = 0; ph < 3; ++ph) {
      cp_async_wait_all();
      __syncthreads();  // the line buffer is complete and visible
      if (live) assemble_half<NH>(v, j, lb, a.kc, tw2);
      else {
#pragma unroll
        for (int r = 0; r < PT; ++r) v[r] = cmake(0.0, 0.0);
      }
      __syncthreads();  // everyone has consumed the line buffer: refill it behind the transforms
      if (live && ph < 2) row_prefetch1<T>(lb, a.W, a.wl, a.wstride * layer + (ph == 0 ? 0 : 1), y, a.kc, j);
      else if (ph == 2 && unit + stride < nunits)
        row_prefetch1<T>(lb, a.W, a.wl, a.wstride * ((unit + stride) / a.ny) + 2, (unit + stride) % a.ny, a.kc, j);
      block_fft_split<NH, INV>(v, j, xb, tw);  // v[r] = (x[2n], x[2n+1]), n = j + r*T
      if (ph == 0) {  // q (+ eta) of this row -> TMEM
        if (a.eta != nullptr && live) {
          const double* ea = a.eta + (long long)y * (2 * NH);
#pragma unroll
          for (int r = 0; r < PT; ++r) {
            const double2 e2 = *reinterpret_cast<const double2*>(ea + 2 * (j + r * T));
            v[r].x += e2.x;
            v[r].y += e2.y;
          }
        }
#pragma unroll
        for (int r = 0; r < PT; ++r) tmem_store(tq + 4 * r, v[r]);
        tmem_wait_st();
        continue;
      }
#pragma unroll
      for (int c0 = 0; c0 < PT; c0 += 4) {
        cplx q2[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) q2[q] = tmem_load(tq + 4 * (c0 + q));
        tmem_wait_ld();
#pragma unroll
        for (int q = 0; q < 4; ++q) {  // u q (or v q) at the even and the odd point
          v[c0 + q].x *= q2[q].x;
          v[c0 + q].y *= q2[q].y;
        }
      }
      block_fft_split<NH, FWD>(v, j, xb, tw);
      unpack_pair<NH>(v, j, xb, A, B);  // A = E, B = O for k = j + r*T <= NH/2
      if (live) {
        const long long rb = a.pl.rowbase(2 * layer + (ph - 1), y);
#pragma unroll
        for (int r = 0; r < 9; ++r) {
          const int k = j + r * T, km = NH - k;
          if (r < 8 || j == 0) {
            const cplx t = cmul(__ldg(&tw2[k]), B[r]);
            cplx X = cadd(A[r], t), Xm = cconj(csub(A[r], t));  // X[k], X[NH - k]
            if (ph == 1) {  // the d field carries -i kr (uq)_h
              const double kr = k * a.dkx, krm = km * a.dkx;
              X = cmake(kr * X.y, -kr * X.x);
              Xm = cmake(krm * Xm.y, -krm * Xm.x);
            }
            if (k < a.kout) *prow_ptr(a.pout, rb, k) = X;
            if (km != k && km < a.kout) *prow_ptr(a.pout, rb, km) = Xm;
          }
          compiler_fence();  // one wavenumber pair at a time: bounded registers
        }
      }
    }
  }
  tmem_release(tbase, tmcols);
}

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
  double a, b, d;  // RK: a = c_i*dt (sub-step), b = weight*dt ; AB3: a, b, d = dt*(23/12, -16/12, 5/12)
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
  if (KIND == EPI_ETD_3) in.n1 = ldpair(e.n1.p + layer * e.n1.ls + off);
  if (KIND == EPI_ETD_1 || KIND == EPI_ETD_2 || KIND == EPI_ETD_3) {
    in.c1.load(e.c1, layer, off);
    in.c2.load(e.c2, layer, off);
  }
  if (KIND == EPI_ETD_1) in.c3.load(e.c3, layer, off);
  if (KIND == EPI_ETD_1 || KIND == EPI_ETD_2 || KIND == EPI_ETD_3 || KIND == EPI_ETD_4) in.c4.load(e.c4, layer, off);
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
                                              const int tid) {
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
#pragma unroll 1
  for (int r0 = 0; r0 < PT; r0 += CH) {
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
                                                  uint32_t tb, int l0, int tid) {
  switch (e.kind) {
#define GFB_EPI_CASE(K) case K: epilogue_tile<NY, K, CL, S>(g, e, layer, colA, liveA, X, tb, l0, tid); break;
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
              const __grid_constant__ EpiArgs e, const int ncols, const cplx* __restrict__ tw) {
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
    if (e.cplxL) epilogue_dispatch<NY, true>(g, e, layer, colA, liveA, X, tA, j, tid);
    else epilogue_dispatch<NY, false>(g, e, layer, colA, liveA, X, tA, j, tid);
    tmem_release(tbase, TmemCfg<NY>::COLS);
  }
}

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
    stpair(out + ooff + r * (4 * T) * ldw, X[r * THREADS + tid], v[r]);
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

// ------------------------------------------------------------------------------------------------
// small elementwise kernels
// ------------------------------------------------------------------------------------------------
// dense user array [layer][l][nkr] <-> ctx-owned tiled array (toff); to_tiled selects the direction
static __global__ void k_repack(const int nl, const int nkr, const int nlayers, cplx* __restrict__ dst, const int ldd,
                                const long long lsd, const cplx* __restrict__ src, const int lds, const long long lss,
                                const int to_tiled) {
  const long long n = (long long)nl * nkr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n * nlayers; i += (long long)gridDim.x * blockDim.x) {
    const int layer = (int)(i / n);
    const long long rr = i - layer * n;
    const int l = (int)(rr / nkr), k = (int)(rr - (long long)l * nkr);
    if (to_tiled) dst[layer * lsd + toff(l, k, ldd)] = src[layer * lss + (long long)l * lds + k];
    else dst[layer * lsd + (long long)l * ldd + k] = src[layer * lss + toff(l, k, lds)];
  }
}

static __global__ void k_dealias(const GridDesc g, cplx* __restrict__ a, const int ld, const long long ls, const int nlayers) {
  const long long n = (long long)g.ny * g.nkl;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n * nlayers; i += (long long)gridDim.x * blockDim.x) {
    const int layer = (int)(i / n);
    const long long rr = i - layer * n;
    const int l = (int)(rr / g.nkl), k = (int)(rr - (long long)l * g.nkl);
    if (gcol(g, k) >= g.kcut || l_aliased(g, l)) a[layer * ls + (long long)l * ld + k] = cmake(0.0, 0.0);
  }
}

// y[:, l, k] = M[l, k] (NL x NL, row-major) * x[:, l, k]   -- pv_streamfunction_kernel!,
// /root/reference/src/multilayerqg.jl:495-509.  One thread per wavenumber, vectors and the matrix row
// live in registers (register-blocked small matvec); consecutive threads = consecutive kr.
template <int NL>
__global__ void k_matvec(const int nl, const int nkr, cplx* __restrict__ y, const cplx* __restrict__ x,
                         const double* __restrict__ M) {
  const long long n = (long long)nl * nkr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    cplx xv[NL];
#pragma unroll
    for (int m = 0; m < NL; ++m) xv[m] = x[m * n + i];
    const double* Mi = M + i * (NL * NL);
#pragma unroll
    for (int r = 0; r < NL; ++r) {
      cplx acc = cmake(0.0, 0.0);
#pragma unroll
      for (int m = 0; m < NL; ++m) {
        const double a = __ldg(&Mi[r * NL + m]);
        acc.x += a * xv[m].x;
        acc.y += a * xv[m].y;
      }
      y[r * n + i] = acc;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// diagnostics: weighted Parseval sums (parsevalsum / parsevalsum2 of FourierFlows; call sites
// /root/reference/src/twodnavierstokes.jl:374-393, src/surfaceqg.jl:390-416,
// src/singlelayerqg.jl:548-599, src/multilayerqg.jl:918-951)
// ------------------------------------------------------------------------------------------------
enum DiagKind {
  DIAG_ABS2 = 0,        // sum w |s|^2
  DIAG_INVK_ABS2,       // sum w invKrsq |s|^2          ([0,0] -> 0)
  DIAG_K_ABS2,          // sum w Krsq |s|^2
  DIAG_KP_ABS2,         // sum w Krsq^p |s|^2 with [0,0] excluded
  DIAG_G_ABS2,          // sum w G |s|^2
  DIAG_G2_ABS2,         // sum w G^2 |s|^2              (|psi_h|^2)
  DIAG_KG2_ABS2,        // sum w Krsq G^2 |s|^2         (|k|^2 |psi_h|^2)
  DIAG_ABS2_PLUS,       // sum w |s + aux|^2            (enstrophy with eta_h)
  DIAG_WORK,            // Re sum w s conj(F)           ; aux2 = prevsol -> (prevsol + s)/2
  DIAG_WORK_INVK,       // Re sum w invKrsq s conj(F)
  DIAG_MLQG_ENERGIES    // per layer: sum w Krsq |psi_j|^2 (NL values) then sum w |psi_j - psi_j+1|^2 (NL-1)
};

struct DiagArgs {
  int kind;
  const cplx* s;
  const cplx* aux;   // eta_h or F
  const cplx* aux2;  // prevsol
  const double* G;   // [l][ldg]
  int ldg;
  const double* Sinv;
  double p;
  double scale;      // multiplies every term (e.g. 0.5)
};

constexpr int DIAG_MAXOUT = 16;

template <int NL>
__global__ void __launch_bounds__(256) k_diag_partial(const GridDesc g, const DiagArgs d, double* __restrict__ partial) {
  const long long n = (long long)g.ny * g.nkl;
  constexpr int NOUT = (2 * NL - 1) > 1 ? (2 * NL - 1) : 1;
  double acc[NOUT];
#pragma unroll
  for (int q = 0; q < NOUT; ++q) acc[q] = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int l = (int)(i / g.nkl), k = (int)(i - (long long)l * g.nkl);
    const int kg = gcol(g, k);
    const bool zero_mode = (l == 0 && kg == 0);
    const double w = (kg == 0 || kg == g.nkr - 1) ? 1.0 : 2.0;
    const double kr = wav_kr(g, k), ll = wav_l(g, l);
    const double K2 = kr * kr + ll * ll;
    if (d.kind == DIAG_MLQG_ENERGIES) {
      cplx psi[NL];
      const double* Mi = d.Sinv + i * (NL * NL);
#pragma unroll
      for (int r = 0; r < NL; ++r) {
        psi[r] = cmake(0.0, 0.0);
#pragma unroll
        for (int m = 0; m < NL; ++m) {
          const cplx q = d.s[m * n + i];
          psi[r].x += Mi[r * NL + m] * q.x;
          psi[r].y += Mi[r * NL + m] * q.y;
        }
      }
#pragma unroll
      for (int r = 0; r < NL; ++r) acc[r] += w * K2 * cabs2(psi[r]);
#pragma unroll
      for (int r = 0; r + 1 < NL; ++r) acc[NL + r] += w * cabs2(csub(psi[r], psi[r + 1]));
      continue;
    }
    const cplx s = d.s[i];
    double t = 0.0;
    switch (d.kind) {
      case DIAG_ABS2: t = cabs2(s); break;
      case DIAG_INVK_ABS2: t = zero_mode ? 0.0 : cabs2(s) / K2; break;
      case DIAG_K_ABS2: t = K2 * cabs2(s); break;
      case DIAG_KP_ABS2: t = zero_mode ? 0.0 : pow(K2, d.p) * cabs2(s); break;
      case DIAG_G_ABS2: t = d.G[toff(l, k, d.ldg)] * cabs2(s); break;
      case DIAG_G2_ABS2: { const double gg = d.G[toff(l, k, d.ldg)]; t = gg * gg * cabs2(s); } break;
      case DIAG_KG2_ABS2: { const double gg = d.G[toff(l, k, d.ldg)]; t = K2 * gg * gg * cabs2(s); } break;
      case DIAG_ABS2_PLUS: t = cabs2(cadd(s, d.aux[i])); break;
      case DIAG_WORK:
      case DIAG_WORK_INVK: {
        cplx ss = s;
        if (d.aux2) ss = cscale(cadd(ss, d.aux2[i]), 0.5);
        const cplx f = d.aux[i];
        t = ss.x * f.x + ss.y * f.y;  // Re(s conj f)
        if (d.kind == DIAG_WORK_INVK) t = zero_mode ? 0.0 : t / K2;
      } break;
      default: break;
    }
    acc[0] += w * t * d.scale;
  }
  // warp-shuffle tree, then one partial per warp through shared memory
  __shared__ double red[8][DIAG_MAXOUT];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < NOUT; ++q) {
    double x = acc[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (lane == 0) red[warp][q] = x;
  }
  __syncthreads();
  if (threadIdx.x < NOUT) {
    double x = 0.0;
    for (int wv = 0; wv < 8; ++wv) x += red[wv][threadIdx.x];
    partial[(long long)blockIdx.x * DIAG_MAXOUT + threadIdx.x] = x;
  }
}

// deterministic second stage: one block sums the per-block partials in a fixed order
static __global__ void __launch_bounds__(256) k_diag_final(const double* __restrict__ partial, const int nblocks, const int nout,
                                                    double* __restrict__ out) {
  __shared__ double red[8];
  for (int q = 0; q < nout; ++q) {
    double x = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) x += partial[(long long)b * DIAG_MAXOUT + q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int wv = 0; wv < 8; ++wv) t += red[wv];
      out[q] = t;
    }
    __syncthreads();
  }
}

}  // namespace gfb
