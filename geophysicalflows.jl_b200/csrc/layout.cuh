// layout.cuh -- grid description, the tiled layouts of the ctx-owned arrays and of the work buffers between the
// passes (W, P), the destinations of the exchange stores (PeerOut) and the thread-group configuration.
// Overview of the kernels that use them: kernels.cuh.
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

}  // namespace gfb
