// launchers.cu -- compiled once per FFT length: -DGFB_N=<16..8192>
#include <cstdlib>

#include "launch.h"

#ifndef GFB_N
#error "compile with -DGFB_N=<fft length>"
#endif

namespace gfb {

#define GFB_CAT2(a, b) a##b
#define GFB_CAT(a, b) GFB_CAT2(a, b)

// cudaFuncSetAttribute and the SM count are PER DEVICE: the "already configured" caches below are indexed by the
// current device, so several contexts on different devices of one process (gfb_comm_connect_ptrs) each opt in
constexpr int MAX_DEVICES = 64;
static inline int cur_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) dev = 0;
  return dev;
}
struct AttrCache {
  size_t bytes[MAX_DEVICES] = {};
  bool done[MAX_DEVICES] = {};
};

template <class K>
static inline int set_smem(K kernel, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return -2;
  // ask for the largest shared-memory carveout: the kernels are sized so that TWO CTAs (2 x ~100 KB at
  // N = 4096) share an SM; with the default carveout the driver only guarantees room for one
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  return e == cudaSuccess ? 0 : -2;
}

template <class Prep>
static int do_col_inv(const ColInvLaunch& L, const Prep& prep) {
  using GC = GroupCfg<GFB_N>;
  const size_t smem = (size_t)GC::XB_DOUBLES * sizeof(double) + (size_t)GC::STASH * sizeof(cplx);
  static AttrCache attr;
  if (!attr.done[cur_device()]) {
    if (set_smem(k_col_inv<GFB_N, Prep>, smem)) return -2;
    attr.done[cur_device()] = true;
  }
  const int npairs = (L.ncols + 1) / 2;
  dim3 grid((npairs + GC::G - 1) / GC::G, L.nfields);
  GFB_LAUNCH((k_col_inv<GFB_N, Prep>), grid, dim3(GC::THREADS), smem, L.st, L.g, prep, L.W, L.ldw, L.wfs, L.ncols, L.tw);
  return 0;
}

// GFB_STAGGER_K1 / GFB_STAGGER_K2: start stagger (ns) of the persistent 4096 kernels (cta_stagger), 0 = off
static int env_ns(const char* name) {
  const char* e = getenv(name);
  return e ? atoi(e) : 0;
}

static int sm_count() {
  static int n[MAX_DEVICES] = {};
  const int dev = cur_device();
  if (n[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess) n[dev] = v;
    if (n[dev] <= 0) n[dev] = 148;
  }
  return n[dev];
}

// the persistent, software-pipelined pre-pass of the single-layer modules (k_col_inv_single)
static int do_col_inv_single(const ColInvLaunch& L) {
  using GC = GroupCfg<GFB_N>;
  // most rows any thread keeps after the dealias mask (thread j owns rows j + r*T)
  int nslots = 1;
  for (int j = 0; j < GC::T; ++j) {
    int n = 0;
    for (int r = 0; r < PT; ++r) {
      const int l = j + r * GC::T;
      if (!(l >= L.g.la0 && l < L.g.la1)) ++n;
    }
    if (n > nslots) nslots = n;
  }
  const size_t smem = (size_t)GC::XB_DOUBLES * sizeof(double) + 16 + (size_t)nslots * GC::THREADS * (sizeof(cplx) + sizeof(double));
  static AttrCache attr;
  if (smem > attr.bytes[cur_device()]) {
    if (set_smem(k_col_inv_single<GFB_N>, smem)) return -2;
    attr.bytes[cur_device()] = smem;
  }
  const int nunits = ((L.ncols + 1) / 2) * L.nfields;
  const int want = (nunits + GC::G - 1) / GC::G, cap = sm_count() * GC::MINB;
  dim3 grid(want < cap ? want : cap);
  GFB_LAUNCH((k_col_inv_single<GFB_N>), grid, dim3(GC::THREADS), smem, L.st, L.g, L.ps, L.W, L.ldw, L.wfs, L.ncols, L.nfields,
             nslots, nunits, L.tw);
  return 0;
}

#if GFB_N == 4096
// the warp-specialised pre-pass (k_col_inv_ws): one 512-thread CTA per SM
template <int VAR>
static int do_col_inv_ws_var(const ColInvLaunch& L) {
  const size_t smem = (size_t)xbuf_elems(GFB_N) * sizeof(cplx) + 128;
  static AttrCache attr;
  if (!attr.done[cur_device()]) {
    if (set_smem(k_col_inv_ws<GFB_N, VAR>, smem)) return -2;
    attr.done[cur_device()] = true;
  }
  const int nunits = ((L.ncols + 1) / 2) * L.nfields;
  const int cap = sm_count();
  dim3 grid(nunits < cap ? nunits : cap);
  static const int stagger = env_ns("GFB_STAGGER_K1");
  GFB_LAUNCH((k_col_inv_ws<GFB_N, VAR>), grid, dim3(WS_THREADS), smem, L.st, L.g, L.ps, L.W, L.ldw, L.wfs, L.ncols, L.nfields, nunits, L.tw,
             stagger);
  return 0;
}
static int do_col_inv_ws(const ColInvLaunch& L) {
  static int var = -1;
  if (var < 0) {
    const char* e = getenv("GFB_K1WS_VAR");
    var = e ? atoi(e) : 16;  // default: the pair-form helper (r2v: K1 0.575 -> 0.545 ms/step at 4096^2)
  }
  switch (var) {
    case 1: return do_col_inv_ws_var<1>(L);
    case 2: return do_col_inv_ws_var<2>(L);
    case 4: return do_col_inv_ws_var<4>(L);
    case 8: return do_col_inv_ws_var<8>(L);
    case 0: return do_col_inv_ws_var<0>(L);
    default: return do_col_inv_ws_var<16>(L);
  }
}
#endif

int GFB_CAT(launch_col_inv_, GFB_N)(const ColInvLaunch& L) {
#if GFB_N == 4096
  if (L.prep == PREP_SINGLE && L.ws) return do_col_inv_ws(L);
#endif
  switch (L.prep) {
    case PREP_PLAIN: return do_col_inv(L, L.pp);
    case PREP_SINGLE: return do_col_inv_single(L);
    case PREP_MULTI: return do_col_inv(L, L.pm);
  }
  return -1;
}

template <int MODE>
static int do_row(const RowLaunch& L) {
  using RC = GroupCfg<GFB_N>;
  const int nst = (MODE == ROW_PRODUCT) ? 1 : (MODE == ROW_PRODUCT_Q ? 2 : 0);
  const size_t smem = (size_t)RC::XB_DOUBLES * sizeof(double) + (size_t)nst * RC::STASH * sizeof(cplx);
  static AttrCache attr;
  if (!attr.done[cur_device()]) {
    if (set_smem(k_row<GFB_N, MODE>, smem)) return -2;
    attr.done[cur_device()] = true;
  }
  const int nunits = L.a.ny / 2;
  dim3 grid((nunits + RC::G - 1) / RC::G, L.nlayers);
  GFB_LAUNCH((k_row<GFB_N, MODE>), grid, dim3(RC::THREADS), smem, L.st, L.a, L.tw);
  return 0;
}

// the persistent, software-pipelined product pass (k_row_product)
template <bool WITHQ>
static int do_row_product(const RowLaunch& L) {
  using RC = GroupCfg<GFB_N>;
  const int kcp = (L.a.kc + 1) & ~1;  // line buffer rows stay 32-byte aligned
  const size_t smem = (size_t)RC::XB_DOUBLES * sizeof(double) + 16 + (size_t)RC::G * 2 * kcp * sizeof(cplx);
  static AttrCache attr;
  if (smem > attr.bytes[cur_device()]) {
    if (set_smem(k_row_product<GFB_N, WITHQ>, smem)) return -2;
    attr.bytes[cur_device()] = smem;
  }
  const int nunits = (L.a.ny / 2) * L.nlayers;
  const int want = (nunits + RC::G - 1) / RC::G, cap = sm_count() * RC::MINB;
  dim3 grid(want < cap ? want : cap);
  GFB_LAUNCH((k_row_product<GFB_N, WITHQ>), grid, dim3(RC::THREADS), smem, L.st, L.a, L.tw, nunits, L.nlayers, kcp);
  return 0;
}

#if GFB_N >= 4096
// the TMA-staged product pass (k_row_product_tma): rows gathered / scattered by cp.async.bulk.tensor
static int do_row_product_tma(const RowLaunch& L) {
  using RC = GroupCfg<GFB_N>;
  const RowTma& t = *L.tma;
  const size_t in_row = (size_t)2 * t.nbox_in * t.boxp_in * sizeof(cplx), out_row = (size_t)2 * t.nbox_out * t.boxp_out * sizeof(cplx);
  if (out_row > (size_t)RC::XB_DOUBLES * sizeof(double)) return -1;  // the d row is staged in the exchange buffer
  const size_t smem = (size_t)RC::XB_DOUBLES * sizeof(double) + 128 + 2 * in_row + out_row;
  static AttrCache attr;
  if (smem > attr.bytes[cur_device()]) {
    if (set_smem(k_row_product_tma<GFB_N>, smem)) return -2;
    attr.bytes[cur_device()] = smem;
  }
  const int nunits = (L.a.ny / 2) * L.nlayers;
  const int cap = sm_count() * RC::MINB;
  dim3 grid(nunits < cap ? nunits : cap);
  static const int stagger = env_ns("GFB_STAGGER_K2");
  GFB_LAUNCH((k_row_product_tma<GFB_N>), grid, dim3(RC::THREADS), smem, L.st, L.a, t, L.tw, nunits, stagger);
  return 0;
}
#endif

#if GFB_N == 4096
// the warp-specialised product pass (k_row_product_ws): one 512-thread CTA per SM
static int do_row_product_ws(const RowLaunch& L) {
  const RowTma& t = *L.tma;
  const size_t in_row = (size_t)2 * t.nbox_in * t.boxp_in * sizeof(cplx), out_row = (size_t)2 * t.nbox_out * t.boxp_out * sizeof(cplx);
  const size_t smem = (size_t)xbuf_elems(GFB_N) * sizeof(cplx) + 128 + 4 * in_row + 2 * out_row;
  if (2 * out_row < (size_t)xbuf_elems(GFB_N) * sizeof(double) || smem > 227 * 1024) return do_row_product_tma(L);
  static AttrCache attr;
  if (smem > attr.bytes[cur_device()]) {
    if (set_smem(k_row_product_ws<GFB_N>, smem)) return -2;
    attr.bytes[cur_device()] = smem;
  }
  const int nunits = (L.a.ny / 2) * L.nlayers;
  const int cap = sm_count();
  dim3 grid(nunits < cap ? nunits : cap);
  GFB_LAUNCH((k_row_product_ws<GFB_N>), grid, dim3(WS_THREADS), smem, L.st, L.a, t, L.tw, nunits);
  return 0;
}
#endif

int GFB_CAT(launch_row_, GFB_N)(const RowLaunch& L) {
#if GFB_N == 4096
  if (L.mode == ROW_PRODUCT && L.tma != nullptr && L.ws) return do_row_product_ws(L);
#endif
#if GFB_N >= 4096
  if (L.mode == ROW_PRODUCT && L.tma != nullptr) return do_row_product_tma(L);
#endif
  switch (L.mode) {
    case ROW_PRODUCT: return do_row_product<false>(L);
    case ROW_PRODUCT_Q: return do_row_product<true>(L);
    case ROW_C2R: return do_row<ROW_C2R>(L);
    case ROW_R2C: return do_row<ROW_R2C>(L);
  }
  return -1;
}

#if GFB_N == 4096
// ny = 16384: column passes as 4 x 4096 (k_col_inv16k / k_col_fwd16k); stepping of the single-layer modules only
template <class Prep>
static int do_col_inv_16384(const ColInvLaunch& L, const Prep& prep) {
  using GC = GroupCfg<4096>;
  const size_t smem = (size_t)GC::XB_DOUBLES * sizeof(double) + (size_t)GC::STASH * sizeof(cplx);
  static AttrCache attr;
  if (!attr.done[cur_device()]) {
    if (set_smem(k_col_inv16k<4096, Prep>, smem)) return -2;
    attr.done[cur_device()] = true;
  }
  const int npairs = (L.ncols + 1) / 2;
  // default: the lane-paired decimation-in-time kernel (one CTA per SM, every row read once, natural-order stores);
  // GFB_K1_16K_DIT=0 selects the four independent q units per pair (kept for comparison)
  static int dit = -1, l2pf = 0;
  if (dit < 0) {
    const char* e = getenv("GFB_K1_16K_DIT");
    dit = (e && e[0] == '0') ? 0 : 1;
    const char* f = getenv("GFB_K1_16K_PF");
    if (f) l2pf = atoi(f);
  }
  if (dit) {
    using PC = Pair16kCfg<4096>;
    static AttrCache attr2;
    if (!attr2.done[cur_device()]) {
      if (set_smem(k_col_inv16k_dit<4096, Prep>, PC::SMEM)) return -2;
      attr2.done[cur_device()] = true;
    }
    const int nunits = npairs * L.nfields;
    const int nsm = sm_count();
    GFB_LAUNCH((k_col_inv16k_dit<4096, Prep>), dim3(nunits < nsm ? nunits : nsm), dim3(PC::THREADS), PC::SMEM, L.st, L.g, prep, L.W, L.ldw,
               L.wfs, L.ncols, L.nfields, nunits, l2pf, L.tw, L.tw2);
    return 0;
  }
  dim3 grid(npairs * L.nfields * 4);
  GFB_LAUNCH((k_col_inv16k<4096, Prep>), grid, dim3(GC::THREADS), smem, L.st, L.g, prep, L.W, L.ldw, L.wfs, L.ncols, L.nfields, L.tw,
             L.tw2);
  return 0;
}

int launch_col_inv_16384(const ColInvLaunch& L) {
  switch (L.prep) {
    case PREP_PLAIN: return do_col_inv_16384(L, L.pp);
    case PREP_SINGLE: return do_col_inv_16384(L, L.ps);
    case PREP_MULTI: return do_col_inv_16384(L, L.pm);
  }
  return -1;
}

int launch_col_fwd_16384(const ColFwdLaunch& L) {
  using GC = GroupCfg<4096>;
  const dim3 grid(((L.ncols + 1) / 2) * 4, L.nlayers);
  if (!L.two) {  // plain forward transform -> dense user array
    const size_t smem = (size_t)GC::XB_DOUBLES * sizeof(double) + (size_t)GC::STASH * sizeof(cplx);
    static AttrCache attr;
    if (!attr.done[cur_device()]) {
      if (set_smem(k_col_fwd16k_plain<4096>, smem)) return -2;
      attr.done[cur_device()] = true;
    }
    GFB_LAUNCH((k_col_fwd16k_plain<4096>), grid, dim3(GC::THREADS), smem, L.st, L.g, L.P, L.ldp, L.pfs, L.e.out.p, L.e.out.ld, L.e.out.ls,
               L.ncols, L.tw, L.tw2);
    return 0;
  }
  // GFB_K3_16K_LP=1 selects the lane-paired kernel (one 512-thread CTA per column pair and half, rows pulled through L2
  // twice instead of four times).  Measured on B200 (profiles/r2q_*): correct, but slower than the four independent q
  // units per pair (20.6 vs 17.7 ms/step at 16384^2) -- one CTA per SM in lock step exposes every load phase (65 % of the
  // warp samples on long-scoreboard stalls) -- so it stays opt-in
  static int lp = -1;
  if (lp < 0) {
    const char* e = getenv("GFB_K3_16K_LP");
    lp = (e && e[0] == '1') ? 1 : 0;
  }
  if (lp) {
    using PC = Pair16kCfg<4096>;
    static AttrCache attr3;
    if (!attr3.done[cur_device()]) {
      if (set_smem(k_col_fwd16k_lp<4096>, PC::SMEM)) return -2;
      attr3.done[cur_device()] = true;
    }
    const int npairs = (L.ncols + 1) / 2;
    GFB_LAUNCH((k_col_fwd16k_lp<4096>), dim3(2 * npairs, L.nlayers), dim3(PC::THREADS), PC::SMEM, L.st, L.g,
               L.P, L.ldp, L.pfs, L.e, L.ncols, npairs, L.tw, L.tw2);
    return 0;
  }
  const size_t smem = (size_t)GC::XB_DOUBLES * sizeof(double) + (size_t)GC::STASH * sizeof(cplx) + 16;
  static AttrCache attr2;
  if (!attr2.done[cur_device()]) {
    if (set_smem(k_col_fwd16k<4096>, smem)) return -2;
    attr2.done[cur_device()] = true;
  }
  GFB_LAUNCH((k_col_fwd16k<4096>), grid, dim3(GC::THREADS), smem, L.st, L.g, L.P, L.ldp, L.pfs, L.e, L.ncols, L.tw, L.tw2);
  return 0;
}
#endif

#if GFB_N == 8192
// nx = 16384: the row pass at half length (k_row_product_half); product mode of the single-layer modules only
template <int MODE>
static int do_row_half_plain(const RowLaunch& L) {
  using RC = GroupCfg<8192>;
  const int kcp = (L.a.kc + 1) & ~1;
  const size_t smem = (size_t)RC::XB_DOUBLES * sizeof(double) + (MODE == ROW_C2R ? (size_t)kcp * sizeof(cplx) : 0);
  static AttrCache attr;
  if (smem > attr.bytes[cur_device()]) {
    if (set_smem(k_row_half<8192, MODE>, smem)) return -2;
    attr.bytes[cur_device()] = smem;
  }
  GFB_LAUNCH((k_row_half<8192, MODE>), dim3(L.a.ny, L.nlayers), dim3(RC::THREADS), smem, L.st, L.a, L.tw, L.tw2, kcp);
  return 0;
}

template <bool WITHQ>
static int do_row_half_product(const RowLaunch& L) {
  using RC = GroupCfg<8192>;
  const int kcp = (L.a.kc + 1) & ~1;
  const size_t smem = (size_t)RC::XB_DOUBLES * sizeof(double) + 16 + (size_t)RC::G * kcp * sizeof(cplx) + (16 + 256) * sizeof(cplx);  // + twiddle tables
  static AttrCache attr;
  if (smem > attr.bytes[cur_device()]) {
    if (set_smem(k_row_product_half<8192, WITHQ>, smem)) return -2;
    attr.bytes[cur_device()] = smem;
  }
  const int nunits = L.a.ny * L.nlayers;
  const int want = (nunits + RC::G - 1) / RC::G, cap = sm_count() * RC::MINB;
  dim3 grid(want < cap ? want : cap);
  GFB_LAUNCH((k_row_product_half<8192, WITHQ>), grid, dim3(RC::THREADS), smem, L.st, L.a, L.tw, L.tw2, nunits, kcp);
  return 0;
}

int launch_row_half_8192(const RowLaunch& L) {
  switch (L.mode) {
    case ROW_C2R: return do_row_half_plain<ROW_C2R>(L);
    case ROW_R2C: return do_row_half_plain<ROW_R2C>(L);
    case ROW_PRODUCT: return do_row_half_product<false>(L);
    case ROW_PRODUCT_Q: return do_row_half_product<true>(L);
  }
  return -1;
}
#endif

template <bool TWO>
static int do_col_fwd(const ColFwdLaunch& L) {
  using GC = GroupCfg<GFB_N>;
  const size_t smem = (size_t)GC::XB_DOUBLES * sizeof(double) + (size_t)GC::STASH * sizeof(cplx) + 16;  // + TMEM base word
  static AttrCache attr;
  if (!attr.done[cur_device()]) {
    if (set_smem(k_col_fwd<GFB_N, TWO>, smem)) return -2;
    attr.done[cur_device()] = true;
  }
  const int ntiles = ((L.ncols + 1) / 2 + GC::G - 1) / GC::G;
  dim3 grid(ntiles, L.nlayers);
  static int l2pf = -1;
  if (l2pf < 0) {
    // opt-in experiment (GFB_K3_L2PF=1): measured slower on B200 (K3 0.645 vs 0.576 ms/step, DRAM reads 491 vs 359 MB/launch)
    // GFB_K3_L2PF=1: bulk prefetch of half a unit (slower), =2: rolling prefetch two chunks ahead inside the epilogue loop
    const char* e = getenv("GFB_K3_L2PF");
    l2pf = e ? atoi(e) : 0;
  }
  GFB_LAUNCH((k_col_fwd<GFB_N, TWO>), grid, dim3(GC::THREADS), smem, L.st, L.g, L.P, L.ldp, L.pfs, L.e, L.ncols, L.tw, l2pf);
  return 0;
}

#if GFB_N == 4096
// the warp-specialised forward column pass + epilogue (k_col_fwd_ws): one 512-thread CTA per SM
static int do_col_fwd_ws(const ColFwdLaunch& L) {
  const size_t smem = (size_t)xbuf_elems(GFB_N) * sizeof(cplx) + (size_t)2 * PT * WS_FFT_THREADS * sizeof(cplx) + 128;
  static AttrCache attr;
  if (!attr.done[cur_device()]) {
    if (set_smem(k_col_fwd_ws<GFB_N>, smem)) return -2;
    attr.done[cur_device()] = true;
  }
  const int npairs = (L.ncols + 1) / 2, nunits = npairs * L.nlayers;
  const int cap = sm_count();
  dim3 grid(nunits < cap ? nunits : cap);
  GFB_LAUNCH((k_col_fwd_ws<GFB_N>), grid, dim3(WS_THREADS), smem, L.st, L.g, L.P, L.ldp, L.pfs, L.e, npairs, nunits, L.tw);
  return 0;
}
#endif

int GFB_CAT(launch_col_fwd_, GFB_N)(const ColFwdLaunch& L) {
#if GFB_N == 4096
  if (L.two && L.ws) return do_col_fwd_ws(L);
#endif
  return L.two ? do_col_fwd<true>(L) : do_col_fwd<false>(L);
}

}  // namespace gfb
