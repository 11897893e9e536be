// gfb.cu -- context and the extern "C" ABI declared in include/gfb.h.
//
// Nothing here allocates on the step path: all tables, work arrays and stage states are created in
// gfb_create and live in the context.  Every entry point only enqueues kernels on the caller's
// stream (gfb_diag additionally copies a few doubles back and synchronises that stream).
#include "../../include/gfb.h"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "launch.h"

using namespace gfb;

static thread_local std::string g_create_error;

struct gfb_ctx {
  gfb_config cfg;  // host pointers inside are NOT retained (nulled after create)
  GridDesc g;
  GridDesc gd;     // = g with the true alias ranges: what dealias! zeroes (differs from g only for calcNlinear!, which never dealiases)
  bool linear = false;  // MultiLayerQG.calcNlinear! (config.no_dealias_in_calcN)
  int nlayers = 1, kind = 0, stepper = 0;
  bool ql = false;  // GFB_KIND_QL (kind is GFB_KIND_SINGLE then)
  bool filtered = false, cplxL = false;
  double dt = 0, t = 0;
  long long step = 0;
  std::string err;
  long long bytes = 0;
  std::vector<void*> allocs;
  // sharding (single GPU: rank 0 of 1, ny_loc = ny)
  int ny_loc = 0;   // physical rows owned by this rank (row pass)
  int nyp = 0;      // ny rounded up to the 8-row tiles of the ctx-owned layouts (= ny unless ny is not a power of two)
  int nkpl = 0;     // column pairs per rank (ldi = 2 * nkpl)
  // tables
  cplx *tw_x = nullptr, *tw_y = nullptr;
  cplx *tw_x2 = nullptr, *tw_y2 = nullptr;  // nx = 16384: W_nx^i for the half-length row scheme (tw_x then holds W_{nx/2}^i); ny = 16384: W_ny^i (tw_y holds W_{ny/4}^i)
  double* G = nullptr;
  double *Sinv = nullptr, *S = nullptr;
  double *eta = nullptr, *U = nullptr, *Qx = nullptr, *Qy = nullptr;
  cplx* Nconst = nullptr;
  void* L = nullptr;
  double* filter = nullptr;
  void *E = nullptr, *Eh = nullptr, *zeta = nullptr, *alpha = nullptr, *beta = nullptr, *Gamma = nullptr;
  // work (ctx-owned W/P only on a single GPU; sharded runs pass the all-to-all buffers in)
  cplx *W = nullptr, *P = nullptr;
  int ldi = 0;
  long long ils = 0;  // layer stride of ctx-owned spectral arrays
  cplx* st[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // stage states / accumulators / histories
  cplx *S0 = nullptr, *Fint = nullptr, *PV = nullptr, *Cin = nullptr, *Nint = nullptr;  // tiled copies of user arrays
  cplx* ALT = nullptr;  // layered kind with a one-stage stepper (ForwardEuler, AB3): the state ping-pongs S0 <-> ALT (plan_stage)
  double *dpart = nullptr, *dout = nullptr;
  // diagnostics come back through a mapped pinned buffer the final reduction kernel writes itself: no copy engine is
  // involved (a cudaMemcpy to the caller's pageable memory serialised against in-flight PCIe transfers of other streams,
  // profiles/r2l_e2e_probe_modes.txt)
  double *dres_host = nullptr, *dres_dev = nullptr;
  int dblocks = 0;
  // "p2p" transport: one IPC-shareable buffer per rank = [flags | W receive | P receive]; peer[d] is rank d's
  // buffer mapped into this process (peer[rank] = comm)
  char* comm = nullptr;
  size_t comm_w_off = 0, comm_p_off = 0, comm_bytes = 0;
  char* peer[MAX_RANKS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool peer_ipc[MAX_RANKS] = {false, false, false, false, false, false, false, false};
  bool connected = false;
  unsigned long long epoch = 0;
  // sticky barrier failure: a mapped pinned word a barrier that gave up writes its epoch into; every later barrier
  // launch, gfb_step_p2p and the sharded transforms read it on the HOST (no synchronisation) and fail
  unsigned long long* tmo_host = nullptr;
  unsigned long long* tmo_dev = nullptr;
  bool defer_alias = false;  // gfb_set_defer_alias_band
  // TMA-staged row pass (single GPU, single-layer kinds, nx = 4096 / 8192): tensor maps of W and P
  bool row_tma = false;
  bool row_ws = false;  // nx = 4096: the warp-specialised row kernel on top of the TMA staging
  long long row_tma_launches = 0;
  bool col_ws = false;  // warp-specialised column pre-pass (ny = 4096, single-layer kinds)
  bool col_ws3 = false; // warp-specialised forward column pass + epilogue (ny = 4096, any kind)
  long long col_ws_launches = 0;
  RowTma rt_pruned, rt_full;  // output boxes sized for kout = kcut / nkr
  // optional per-kernel event timing
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_ev;   // pairs (start, stop)
  std::vector<int> prof_cls;
  size_t prof_used = 0;
  long long prof_launches[4] = {0, 0, 0, 0};
};

// RAII: records an event pair around one launch when profiling is on
struct ProfScope {
  gfb_ctx* c;
  cudaStream_t st;
  cudaEvent_t stop = nullptr;
  ProfScope(gfb_ctx* c_, int cls, cudaStream_t st_) : c(c_), st(st_) {
    if (!c->prof_on) return;
    c->prof_launches[cls]++;
    if (c->prof_used + 2 > c->prof_ev.size()) return;
    cudaEventRecord(c->prof_ev[c->prof_used], st);
    stop = c->prof_ev[c->prof_used + 1];
    c->prof_cls.push_back(cls);
    c->prof_used += 2;
  }
  ~ProfScope() {
    if (stop) cudaEventRecord(stop, st);
  }
};

static int fail(gfb_ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf; else g_create_error = buf;
  return code;
}

#define GFB_CUDA(c, call)                                                                  \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) return fail((c), GFB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
  } while (0)

static int check_launch(gfb_ctx* c, int rc, const char* what) {
  if (rc == -1) return fail(c, GFB_ERR_UNSUPPORTED, "%s: unsupported size/layer count", what);
  if (rc != 0) return fail(c, GFB_ERR_CUDA, "%s: cudaFuncSetAttribute failed", what);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(c, GFB_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return 0;
}

template <class T>
static int dev_alloc(gfb_ctx* c, T** p, size_t n);
static int nstages_of_stepper(int stepper) {
  switch (stepper) {
    case GFB_RK4: case GFB_FILTERED_RK4: case GFB_ETDRK4: case GFB_FILTERED_ETDRK4: return 4;
    case GFB_LSRK54: case GFB_FILTERED_LSRK54: return 5;
    default: return 1;
  }
}

template <class T>
static int dev_alloc(gfb_ctx* c, T** p, size_t n) {
  void* q = nullptr;
  if (cudaMalloc(&q, n * sizeof(T)) != cudaSuccess) return fail(c, GFB_ERR_ALLOC, "cudaMalloc of %zu bytes failed", n * sizeof(T));
  c->allocs.push_back(q);
  c->bytes += (long long)(n * sizeof(T));
  *p = (T*)q;
  return 0;
}

template <class T>
static int dev_upload(gfb_ctx* c, T** p, const void* host, size_t n) {
  int rc = dev_alloc(c, p, n);
  if (rc) return rc;
  GFB_CUDA(c, cudaMemcpy(*p, host, n * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

// upload a dense host table [nlayers][nl][nkr] (double or cplx) into the ctx-owned tiled layout (toff, ld = ldi)
static int upload_padded(gfb_ctx* c, void** out, const void* host, int nlay, bool is_cplx) {
  const int nl = c->g.ny, nkr = c->g.nkl, ldi = c->ldi;
  const size_t es = is_cplx ? sizeof(cplx) : sizeof(double);
  const size_t nlp = (size_t)c->nyp;  // tiled: whole 8-row blocks
  std::vector<char> tmp((size_t)nlay * nlp * ldi * es, 0);
  for (int a = 0; a < nlay; ++a)
    for (int l = 0; l < nl; ++l)
      for (int k = 0; k < nkr; ++k)
        std::memcpy(&tmp[((size_t)a * nlp * ldi + (size_t)toff(l, k, ldi)) * es],
                    (const char*)host + ((size_t)(a * nl + l) * nkr + k) * es, es);
  char* d = nullptr;
  int rc = dev_alloc(c, &d, tmp.size());
  if (rc) return rc;
  GFB_CUDA(c, cudaMemcpy(d, tmp.data(), tmp.size(), cudaMemcpyHostToDevice));
  *out = d;
  return 0;
}

static bool pow2_ok(int n) { return n >= 32 && n <= 8192 && (n & (n - 1)) == 0; }

static int make_twiddles(gfb_ctx* c, cplx** out, int N) {
  std::vector<cplx> h(N);
  for (int i = 0; i < N; ++i) {
    // exact octant symmetries keep |w| = 1 to the last bit where it matters
    const long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)i / (long double)N;
    h[i] = make_double2((double)cosl(a), (double)sinl(a));
  }
  return dev_upload(c, out, h.data(), (size_t)N);
}

// ---- TMA tensor maps ---------------------------------------------------------------------------------------
// W / P as the row kernel sees them (k2_row.cuh RowLayout, one GPU): element (field, y, k) at
//   (y/8)*rbs + field*fstr + (k/2)*16 + (y%8)*2 + k%2   complex elements, rbs = 8*nf*ld, fstr = 8*ld,
// i.e. a 5-D tensor of doubles [y/8][field][k/2][y%8][4] (slowest first).  A box of {4, 1, boxp, 1, 1} is one row's
// run of boxp column pairs: boxp 32-byte pieces 256 bytes apart in global memory, dense in shared memory.
#ifndef GFB_EMULATE
#include <cuda.h>
typedef CUresult (*gfb_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
#endif
static int make_tma_5d(gfb_ctx* c, TmaDesc* out, void* base, const unsigned long long dims[5], const unsigned long long strides_bytes[5],
                       const unsigned box[5]) {
#ifdef GFB_EMULATE
  out->base = (char*)base;
  out->rank = 5;
  out->esize = 8;
  for (int i = 0; i < 5; ++i) {
    out->dims[i] = dims[i];
    out->strides[i] = strides_bytes[i];
    out->box[i] = box[i];
  }
  return 0;
#else
  static gfb_encode_fn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn ||
        q != cudaDriverEntryPointSuccess)
      return fail(c, GFB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    encode = (gfb_encode_fn)fn;
  }
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < 5; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    if (i > 0) gs[i - 1] = strides_bytes[i];
  }
  static_assert(sizeof(TmaDesc) == sizeof(CUtensorMap), "TmaDesc is the opaque CUtensorMap");
  const CUresult r = encode(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, base, gd, gs, bx, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(c, GFB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
#endif
}
// boxes per row and column pairs per box (a multiple of 4 = 128 bytes) covering ncols columns
static void row_boxes(int ncols, int* nbox, int* boxp) {
  const int npairs = (ncols + 1) / 2;
  *nbox = (npairs + 255) / 256;
  *boxp = (((npairs + *nbox - 1) / *nbox) + 3) & ~3;
}
static int make_row_tensor(gfb_ctx* c, TmaDesc* out, cplx* base, int nf, int boxp) {
  const unsigned long long ld = (unsigned long long)c->ldi;
  const unsigned long long dims[5] = {4, 8, ld / 2, (unsigned long long)nf, (unsigned long long)(c->ny_loc / 8)};
  const unsigned long long strides[5] = {8, 32, 256, 8 * ld * 16, 8 * (unsigned long long)nf * ld * 16};
  const unsigned box[5] = {4, 1, (unsigned)boxp, 1, 1};
  return make_tma_5d(c, out, base, dims, strides, box);
}

namespace gfb {
int launch_col_inv(int NY, const ColInvLaunch& L) {
  if (NY == 16384) return launch_col_inv_16384(L);
  switch (NY) {
#define GFB_CASE(N) case N: return launch_col_inv_##N(L);
    GFB_CASE(32) GFB_CASE(64) GFB_CASE(128) GFB_CASE(256) GFB_CASE(512) GFB_CASE(1024) GFB_CASE(2048) GFB_CASE(4096) GFB_CASE(8192)
#undef GFB_CASE
  }
  if (NY >= 4 && NY <= DFT_MAX_NY && NY % 2 == 0) return launch_col_inv_dft(L);
  return -1;
}
int launch_row(int NX, const RowLaunch& L) {
  if (NX == 16384) return launch_row_half_8192(L);
  switch (NX) {
#define GFB_CASE(N) case N: return launch_row_##N(L);
    GFB_CASE(32) GFB_CASE(64) GFB_CASE(128) GFB_CASE(256) GFB_CASE(512) GFB_CASE(1024) GFB_CASE(2048) GFB_CASE(4096) GFB_CASE(8192)
#undef GFB_CASE
  }
  return -1;
}
int launch_col_fwd(int NY, const ColFwdLaunch& L) {
  if (NY == 16384) return launch_col_fwd_16384(L);
  switch (NY) {
#define GFB_CASE(N) case N: return launch_col_fwd_##N(L);
    GFB_CASE(32) GFB_CASE(64) GFB_CASE(128) GFB_CASE(256) GFB_CASE(512) GFB_CASE(1024) GFB_CASE(2048) GFB_CASE(4096) GFB_CASE(8192)
#undef GFB_CASE
  }
  if (NY >= 4 && NY <= DFT_MAX_NY && NY % 2 == 0) return launch_col_fwd_dft(L);
  return -1;
}
}  // namespace gfb


// ---------------------------------------------------------------------------------------------
// geometry helpers
// ---------------------------------------------------------------------------------------------
static int local_cols_below(const gfb_ctx* c, int gbound) {  // local columns whose global index is < gbound
  int n = 0;
  for (int col = 0; col < 2 * c->nkpl; ++col) {
    const int gc = ((col >> 1) * c->g.nranks + c->g.rank) * 2 + (col & 1);
    if (gc < gbound) n = col + 1;  // live columns form a prefix of the local ones
  }
  return n;
}
static long long w_elems(const gfb_ctx* c, int nf) { return (long long)c->nyp * nf * c->ldi; }
static RowLayout row_layout(const gfb_ctx* c, int nf) {
  RowLayout r;
  r.nranks = c->g.nranks;
  r.rbs = 8 * nf * c->ldi;
  r.fstr = 8 * c->ldi;
  r.chunk = (long long)(c->ny_loc / 8) * r.rbs;
  return r;
}

static int ilog2i(int n) {
  int l = 0;
  while ((1 << l) < n) ++l;
  return l;
}
// destinations of the K1 / K2 stores (layout.cuh PeerOut)
static PeerOut out_local(cplx* p) {  // a ctx-owned array on one GPU
  PeerOut o;
  std::memset(&o, 0, sizeof o);
  o.base[0] = p;
  o.nranks = 1;
  return o;
}
static PeerOut out_send(const gfb_ctx* c, cplx* send, int nf) {  // "nccl": chunk d of a local send buffer
  PeerOut o;
  std::memset(&o, 0, sizeof o);
  o.chunk = row_layout(c, nf).chunk;
  o.rank = c->g.rank;
  o.nranks = c->g.nranks;
  o.lg_rows = ilog2i(c->ny_loc);
  for (int d = 0; d < o.nranks; ++d) o.base[d] = send + (long long)(d - o.rank) * o.chunk;
  return o;
}
static PeerOut out_peers(const gfb_ctx* c, size_t byte_off, int nf) {  // "p2p": straight into rank d's receive buffer
  PeerOut o;
  std::memset(&o, 0, sizeof o);
  o.chunk = row_layout(c, nf).chunk;
  o.rank = c->g.rank;
  o.nranks = c->g.nranks;
  o.lg_rows = ilog2i(c->ny_loc);
  for (int d = 0; d < o.nranks; ++d) o.base[d] = reinterpret_cast<cplx*>(c->peer[d] + byte_off);
  return o;
}

extern "C" {

const char* gfb_version(void) { return "gfb 0.3 (sm_100a)"; }

int gfb_config_size(void) { return (int)sizeof(gfb_config); }

const char* gfb_last_error(const gfb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

long long gfb_device_bytes(const gfb_ctx* ctx) { return ctx ? ctx->bytes : 0; }

int gfb_destroy(gfb_ctx* c) {
  if (!c) return GFB_OK;
#ifndef GFB_EMULATE
  for (int d = 0; d < MAX_RANKS; ++d)
    if (c->peer_ipc[d] && c->peer[d]) cudaIpcCloseMemHandle(c->peer[d]);
#endif
  for (void* p : c->allocs) cudaFree(p);
  if (c->tmo_host) cudaFreeHost(c->tmo_host);
  if (c->dres_host) cudaFreeHost(c->dres_host);
  for (auto& e : c->prof_ev) cudaEventDestroy(e);
  delete c;
  return GFB_OK;
}

int gfb_create(gfb_ctx** out, const gfb_config* cfg) {
  if (!out || !cfg) return fail(nullptr, GFB_ERR_INVALID, "null argument");
  *out = nullptr;
  if (cfg->struct_size != (int)sizeof(gfb_config))
    return fail(nullptr, GFB_ERR_INVALID, "gfb_config.struct_size = %d, the library expects %d: the binding's struct mirror does not match include/gfb.h",
                cfg->struct_size, (int)sizeof(gfb_config));
  // 16384-point lines: rows as 8192-point half-length transforms, columns as 4 x 4096 
  const bool half_rows = cfg->nx == 16384, quad_cols = cfg->ny == 16384;
  // any other even ny <= DFT_MAX_NY: direct column transforms (cols_dft.cu; the reference's tests use ny = 126, 66, 34), one GPU
  const bool dft_cols = !pow2_ok(cfg->ny) && !quad_cols && cfg->ny >= 8 && cfg->ny <= DFT_MAX_NY && cfg->ny % 2 == 0 && cfg->nranks <= 1;
  if (!(pow2_ok(cfg->nx) || half_rows) || !(pow2_ok(cfg->ny) || quad_cols || dft_cols))
    return fail(nullptr, GFB_ERR_UNSUPPORTED, "nx=%d, ny=%d: nx must be a power of two in [32, 16384]; ny a power of two in [32, 16384] or any "
                "even number in [8, %d] (one GPU)", cfg->nx, cfg->ny, DFT_MAX_NY);
  if (cfg->nlayers < 1 || cfg->nlayers > MAX_LAYERS) return fail(nullptr, GFB_ERR_UNSUPPORTED, "nlayers=%d not in [1, %d]", cfg->nlayers, MAX_LAYERS);
  if ((cfg->kind == GFB_KIND_SINGLE || cfg->kind == GFB_KIND_QL) && cfg->nlayers != 1)
    return fail(nullptr, GFB_ERR_INVALID, "GFB_KIND_SINGLE / GFB_KIND_QL need nlayers == 1");
  if (cfg->kind != GFB_KIND_SINGLE && cfg->kind != GFB_KIND_LAYERED && cfg->kind != GFB_KIND_QL)
    return fail(nullptr, GFB_ERR_INVALID, "unknown kind %d", cfg->kind);
  if (cfg->kind == GFB_KIND_QL && (cfg->eta || cfg->Nconst)) return fail(nullptr, GFB_ERR_INVALID, "GFB_KIND_QL takes no eta / Nconst");
  if (cfg->stepper < GFB_FORWARD_EULER || cfg->stepper > GFB_FILTERED_LSRK54) return fail(nullptr, GFB_ERR_INVALID, "unknown stepper %d", cfg->stepper);
  if (cfg->no_dealias_in_calcN && cfg->kind != GFB_KIND_LAYERED)
    return fail(nullptr, GFB_ERR_INVALID, "no_dealias_in_calcN (MultiLayerQG.calcNlinear!) needs GFB_KIND_LAYERED");
  if (!cfg->L) return fail(nullptr, GFB_ERR_INVALID, "config.L is null");
  const int nkr = cfg->nx / 2 + 1;
  if (cfg->kr_alias0 < 1 || cfg->kr_alias0 > nkr || cfg->l_alias0 < 0 || cfg->l_alias1 > cfg->ny || cfg->l_alias0 > cfg->l_alias1)
    return fail(nullptr, GFB_ERR_INVALID, "bad alias ranges");
  const int nranks = cfg->nranks > 0 ? cfg->nranks : 1;
  const int rank = cfg->nranks > 0 ? cfg->rank : 0;
  if (rank < 0 || rank >= nranks || nranks > MAX_RANKS || (nranks & (nranks - 1)) != 0 || (cfg->ny % (8 * nranks) != 0 && !dft_cols))
    return fail(nullptr, GFB_ERR_INVALID, "bad sharding: rank %d of %d (power of two, ny divisible by 8*nranks)", rank, nranks);
  gfb_ctx* c = new gfb_ctx();
  c->cfg = *cfg;
  c->nlayers = cfg->nlayers;
  // BarotropicQGQL is a single-layer kind whose row pass is k_row_ql; everything else treats it as GFB_KIND_SINGLE
  c->ql = cfg->kind == GFB_KIND_QL;
  c->kind = c->ql ? (int)GFB_KIND_SINGLE : cfg->kind;
  c->stepper = cfg->stepper;
  c->filtered = (cfg->stepper & 1) != 0;
  c->cplxL = cfg->L_is_complex != 0;
  c->dt = cfg->dt;
  GridDesc& g = c->g;
  g.nx = cfg->nx;
  g.ny = cfg->ny;
  g.nkr = nkr;
  g.kcut = cfg->kr_alias0;
  g.la0 = cfg->l_alias0;
  g.la1 = cfg->l_alias1;
  // wavenumber multipliers exactly as FourierFlows forms them: (2pi/L*n)/n
  g.dkx = (2.0 * M_PI / cfg->Lx * cfg->nx) / cfg->nx;
  g.dky = (2.0 * M_PI / cfg->Ly * cfg->ny) / cfg->ny;
  g.invN = 1.0 / ((double)cfg->nx * (double)cfg->ny);
  g.rank = rank;
  g.nranks = nranks;
  const int nkp_total = (nkr + 1) / 2;
  c->nkpl = (nkp_total + nranks - 1) / nranks;
  c->ldi = 2 * c->nkpl;
  c->ny_loc = cfg->ny / nranks;
  g.nkl = local_cols_below(c, nkr);
  c->gd = g;
  c->linear = cfg->no_dealias_in_calcN != 0;
  if (c->linear) {  // calcNlinear! (src/multilayerqg.jl:675-683) does not call dealias!: the stepping kernels mask nothing
    g.kcut = nkr;
    g.la0 = g.la1 = 0;
  }
  const int n = c->nlayers, nl = g.ny;
  c->nyp = (g.ny + 7) & ~7;
  c->ils = (long long)c->nyp * c->ldi;
  int rc = 0;
#define TRY(x)             \
  do {                     \
    rc = (x);              \
    if (rc) {              \
      g_create_error = c->err; \
      gfb_destroy(c);      \
      return rc;           \
    }                      \
  } while (0)
  TRY(make_twiddles(c, &c->tw_x, half_rows ? g.nx / 2 : g.nx));
  if (half_rows) TRY(make_twiddles(c, &c->tw_x2, g.nx));
  TRY(make_twiddles(c, &c->tw_y, quad_cols ? g.ny / 4 : g.ny));
  if (quad_cols) TRY(make_twiddles(c, &c->tw_y2, g.ny));
  const size_t nspec = (size_t)nl * g.nkl, nphys = (size_t)c->ny_loc * g.nx;
  if (cfg->G) {
    void* gp = nullptr;
    TRY(upload_padded(c, &gp, cfg->G, 1, false));
    c->G = (double*)gp;
  }
  if (c->kind == GFB_KIND_SINGLE) {
    if (!cfg->G) { fail(c, GFB_ERR_INVALID, "GFB_KIND_SINGLE needs config.G"); TRY(GFB_ERR_INVALID); }
    if (cfg->eta) TRY(dev_upload(c, &c->eta, cfg->eta, nphys));
  } else {
    if (!cfg->Sinv || !cfg->Qx || !cfg->Qy) { fail(c, GFB_ERR_INVALID, "GFB_KIND_LAYERED needs config.Sinv, Qx, Qy"); TRY(GFB_ERR_INVALID); }
    TRY(dev_upload(c, &c->Sinv, cfg->Sinv, nspec * n * n));
    if (cfg->S) TRY(dev_upload(c, &c->S, cfg->S, nspec * n * n));
    if (cfg->U) TRY(dev_upload(c, &c->U, cfg->U, (size_t)n * c->ny_loc));
    TRY(dev_upload(c, &c->Qx, cfg->Qx, nphys * n));
    TRY(dev_upload(c, &c->Qy, cfg->Qy, nphys * n));
  }
  if (cfg->Nconst) {
    void* np = nullptr;
    TRY(upload_padded(c, &np, cfg->Nconst, n, true));
    c->Nconst = (cplx*)np;
  }
  TRY(upload_padded(c, &c->L, cfg->L, n, c->cplxL));
  if (c->filtered) {
    if (!cfg->filter) { fail(c, GFB_ERR_INVALID, "filtered stepper needs config.filter"); TRY(GFB_ERR_INVALID); }
    void* f = nullptr;
    TRY(upload_padded(c, &f, cfg->filter, 1, false));
    c->filter = (double*)f;
  }
  const bool etd = cfg->stepper == GFB_ETDRK4 || cfg->stepper == GFB_FILTERED_ETDRK4;
  if (etd) {
    if (!cfg->expLdt || !cfg->exphalfLdt || !cfg->zeta || !cfg->alpha || !cfg->beta || !cfg->Gamma) {
      fail(c, GFB_ERR_INVALID, "ETDRK4 needs expLdt, exphalfLdt, zeta, alpha, beta, Gamma");
      TRY(GFB_ERR_INVALID);
    }
    TRY(upload_padded(c, &c->E, cfg->expLdt, n, c->cplxL));
    TRY(upload_padded(c, &c->Eh, cfg->exphalfLdt, n, c->cplxL));
    TRY(upload_padded(c, &c->zeta, cfg->zeta, n, c->cplxL));
    TRY(upload_padded(c, &c->alpha, cfg->alpha, n, c->cplxL));
    TRY(upload_padded(c, &c->beta, cfg->beta, n, c->cplxL));
    TRY(upload_padded(c, &c->Gamma, cfg->Gamma, n, c->cplxL));
  }
  if (nranks == 1) {
    TRY(dev_alloc(c, &c->W, (size_t)w_elems(c, 4 * n)));
    TRY(dev_alloc(c, &c->P, (size_t)w_elems(c, 2 * n)));
  }
  if (nranks == 1 && c->kind == GFB_KIND_SINGLE && !c->ql && (g.nx == 4096 || g.nx == 8192) && (g.ny % 8) == 0) {
    // TMA-staged row pass: tensor maps of W (3 fields) and P (2 fields); the output boxes come in two sizes
    int nbi, bpi, nbo, bpo;
    row_boxes(g.kcut, &nbi, &bpi);
    for (int full = 0; full < 2; ++full) {
      RowTma& rt = full ? c->rt_full : c->rt_pruned;
      row_boxes(full ? g.nkr : g.kcut, &nbo, &bpo);
      rt.nbox_in = nbi;
      rt.boxp_in = bpi;
      rt.nbox_out = nbo;
      rt.boxp_out = bpo;
      TRY(make_row_tensor(c, &rt.w, c->W, 3 * n, bpi));
      TRY(make_row_tensor(c, &rt.p, c->P, 2 * n, bpo));
    }
    const char* off = getenv("GFB_NO_ROW_TMA");
    c->row_tma = !(off && off[0] == '1');
    const char* onw = getenv("GFB_ROW_WS");
    c->row_ws = c->row_tma && g.nx == 4096 && onw && onw[0] == '1';
  }
  {
    const char* off = getenv("GFB_NO_COL_WS");
    c->col_ws = c->kind == GFB_KIND_SINGLE && g.ny == 4096 && !(off && off[0] == '1');
    // the warp-specialised K3 / K2 are opt-in experiments (GFB_COL_WS3=1 / GFB_ROW_WS=1): correct on hardware, but so
    // far slower than the kernels they would replace
    const char* on3 = getenv("GFB_COL_WS3");
    c->col_ws3 = g.ny == 4096 && on3 && on3[0] == '1';
  }
  int nstate = 0;
  switch (cfg->stepper) {
    case GFB_FORWARD_EULER: case GFB_FILTERED_FORWARD_EULER: nstate = 0; break;
    case GFB_RK4: case GFB_FILTERED_RK4: nstate = 3; break;        // s1, s2 (ping-pong), acc
    case GFB_ETDRK4: case GFB_FILTERED_ETDRK4: nstate = 5; break;  // s1, s2, acc, N1, s3
    case GFB_LSRK54: case GFB_FILTERED_LSRK54: nstate = (cfg->kind == GFB_KIND_LAYERED) ? 3 : 1; break;  // S2 (+ two stage states)
    default: nstate = 3; break;                                    // AB3: rhs, rhs-1, rhs-2
  }
  for (int i = 0; i < nstate; ++i) {
    TRY(dev_alloc(c, &c->st[i], (size_t)n * c->ils));
    GFB_CUDA(c, cudaMemset(c->st[i], 0, (size_t)n * c->ils * sizeof(cplx)));
  }
  if (cfg->kind == GFB_KIND_LAYERED && nstages_of_stepper(cfg->stepper) == 1) {
    TRY(dev_alloc(c, &c->ALT, (size_t)n * c->ils));
    GFB_CUDA(c, cudaMemset(c->ALT, 0, (size_t)n * c->ils * sizeof(cplx)));
  }
  {
    cplx** arrs[5] = {&c->S0, &c->Fint, &c->PV, &c->Cin, &c->Nint};
    for (auto a : arrs) {
      TRY(dev_alloc(c, a, (size_t)n * c->ils));
      GFB_CUDA(c, cudaMemset(*a, 0, (size_t)n * c->ils * sizeof(cplx)));
    }
  }
  c->dblocks = 592;
  TRY(dev_alloc(c, &c->dpart, (size_t)c->dblocks * DIAG_MAXOUT));
  TRY(dev_alloc(c, &c->dout, (size_t)DIAG_MAXOUT));
  if (cudaHostAlloc((void**)&c->dres_host, DIAG_MAXOUT * sizeof(double), cudaHostAllocMapped) != cudaSuccess ||
      cudaHostGetDevicePointer((void**)&c->dres_dev, c->dres_host, 0) != cudaSuccess) {
    fail(c, GFB_ERR_ALLOC, "cudaHostAlloc of the diagnostics buffer failed");
    TRY(GFB_ERR_ALLOC);
  }
#undef TRY
  c->cfg.L = c->cfg.filter = nullptr;  // host pointers are not retained
  *out = c;
  return GFB_OK;
}

int gfb_shard_info(const gfb_ctx* c, long long* info8) {
  if (!c || !info8) return GFB_ERR_INVALID;
  info8[0] = c->g.rank;
  info8[1] = c->g.nranks;
  info8[2] = c->g.nkl;                       // local dense spectral columns
  info8[3] = c->ldi;                         // padded local columns (2 * pairs per rank)
  info8[4] = c->ny_loc;                      // local physical rows
  info8[5] = w_elems(c, 3 * c->nlayers);     // complex elements of the W all-to-all buffers
  info8[6] = w_elems(c, 2 * c->nlayers);     // complex elements of the P all-to-all buffers
  info8[7] = c->nkpl;
  return GFB_OK;
}

int gfb_profile_begin(gfb_ctx* c) {
  if (!c) return GFB_ERR_INVALID;
  if (c->prof_ev.empty()) {
    c->prof_ev.resize(2 * 8192);
    for (auto& e : c->prof_ev) GFB_CUDA(c, cudaEventCreate(&e));
  }
  c->prof_used = 0;
  c->prof_cls.clear();
  for (int i = 0; i < 4; ++i) c->prof_launches[i] = 0;
  c->prof_on = true;
  return GFB_OK;
}

int gfb_profile_end(gfb_ctx* c, double* ms4, long long* launches4) {
  if (!c || !ms4 || !launches4) return GFB_ERR_INVALID;
  c->prof_on = false;
  double ms[4] = {0, 0, 0, 0};
  long long timed[4] = {0, 0, 0, 0};
  for (size_t i = 0; i + 1 < c->prof_used; i += 2) {
    GFB_CUDA(c, cudaEventSynchronize(c->prof_ev[i + 1]));
    float t = 0.f;
    GFB_CUDA(c, cudaEventElapsedTime(&t, c->prof_ev[i], c->prof_ev[i + 1]));
    const int cls = c->prof_cls[i / 2];
    ms[cls] += t;
    timed[cls]++;
  }
  for (int k = 0; k < 4; ++k) {
    // launches beyond the event pool are extrapolated from the timed ones of the same class
    ms4[k] = timed[k] ? ms[k] * (double)c->prof_launches[k] / (double)timed[k] : 0.0;
    launches4[k] = c->prof_launches[k];
  }
  return GFB_OK;
}

int gfb_path_info(const gfb_ctx* c, long long* info8) {
  if (!c || !info8) return GFB_ERR_INVALID;
  info8[0] = c->row_tma ? 1 : 0;
  info8[1] = c->row_tma_launches;
  info8[2] = c->col_ws ? 1 : 0;
  info8[3] = c->col_ws_launches;
  info8[4] = c->row_ws ? 1 : 0;
  info8[5] = c->col_ws3 ? 1 : 0;
  info8[6] = info8[7] = 0;
  return GFB_OK;
}

int gfb_get_clock(const gfb_ctx* c, double* t, long long* step) {
  if (!c) return GFB_ERR_INVALID;
  if (t) *t = c->t;
  if (step) *step = c->step;
  return GFB_OK;
}
int gfb_set_clock(gfb_ctx* c, double t, long long step) {
  if (!c) return GFB_ERR_INVALID;
  c->t = t;
  c->step = step;
  return GFB_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// internal building blocks
// ---------------------------------------------------------------------------------------------
static Spec int_spec(const gfb_ctx* c, cplx* p) {
  Spec s;
  s.p = p;
  s.ld = c->ldi;
  s.ls = c->ils;
  return s;
}
static Spec null_spec() {
  Spec s;
  s.p = nullptr;
  s.ld = 0;
  s.ls = 0;
  return s;
}
static Coef coef(const gfb_ctx* c, const void* p) {
  Coef k;
  k.r = c->cplxL ? nullptr : (const double*)p;
  k.c = c->cplxL ? (const cplx*)p : nullptr;
  k.ld = c->ldi;
  k.ls = c->ils;
  return k;
}

// dense local user array [layer][l][nkl] <-> ctx-owned tiled array
static int repack(gfb_ctx* c, cplx* dst, int ldd, long long lsd, const cplx* src, int lds, long long lss, int to_tiled,
                  cudaStream_t st) {
  ProfScope ps(c, 3, st);
  GFB_LAUNCH(k_repack, dim3(1184), dim3(256), 0, st, c->g.ny, c->g.nkl, c->nlayers, dst, ldd, lsd, src, lds, lss, to_tiled);
  return check_launch(c, 0, "k_repack");
}
static int import_spec(gfb_ctx* c, cplx* internal, const void* user, cudaStream_t st) {
  return repack(c, internal, c->ldi, c->ils, (const cplx*)user, c->g.nkl, (long long)c->g.ny * c->g.nkl, 1, st);
}
static int export_spec(gfb_ctx* c, void* user, const cplx* internal, cudaStream_t st) {
  return repack(c, (cplx*)user, c->g.nkl, (long long)c->g.ny * c->g.nkl, internal, c->ldi, c->ils, 0, st);
}

// K1 on a stage input: fields u, v, q (+psi if nf == 4) for every layer of the LOCAL columns -> W
// (tiled, row-block major, fields interleaved per row block: ld = nf*n*ldi, field stride = 8*ldi)
static int run_col_inv_state(gfb_ctx* c, const Spec& sin, int nf, const PeerOut& W, cudaStream_t st) {
  ColInvLaunch L;
  std::memset(&L, 0, sizeof L);
  L.g = c->g;
  L.W = W;
  L.ldw = nf * c->nlayers * c->ldi;
  L.wfs = 8LL * c->ldi;
  L.ncols = local_cols_below(c, c->g.kcut);
  L.tw = c->tw_y;
  L.tw2 = c->tw_y2;
  L.st = st;
  L.nl = c->nlayers;
  if (L.ncols == 0) return 0;
  if (c->kind == GFB_KIND_SINGLE) {
    L.prep = PREP_SINGLE;
    L.ps.sol = sin.p;
    L.ps.ld = sin.ld;
    L.ps.G = c->G;
    L.ps.ldg = c->ldi;
    L.nfields = nf;
    L.ws = c->col_ws;
    if (L.ws) c->col_ws_launches += 1;
  } else {
    L.prep = PREP_MULTI;
    L.pm.sol = sin.p;
    L.pm.ld = sin.ld;
    L.pm.ls = sin.ls;
    L.pm.Sinv = c->Sinv;
    L.pm.ldS = c->g.nkl;
    L.pm.nf = nf;
    L.pm.nl = c->nlayers;
    L.nfields = nf * c->nlayers;
  }
  ProfScope ps(c, 0, st);
  return check_launch(c, launch_col_inv(c->g.ny, L), "k_col_inv");
}

// ---- BarotropicQGQL row pass (GFB_KIND_QL) -----------------------------------------------------------------------
// calcN_advection! of /root/reference/src/barotropicqgql.jl:293-333 between the two column passes.  After K1 the work
// array holds, per row y, the x-spectra W_f(kr, y) of u = i l G sol, v = -i kr G sol and q = sol.  The zonal parts are
// the kr = 0 entries -- U(y) = Re W_u(0, y), Zeta(y) = Re W_q(0, y) -- and do not depend on x, so the reference's
// eddy-mean products U zeta' + u' Zeta and v' Zeta are LINEAR in the eddy x-spectra: no x-transform is needed,
//     P_d(kr, y) = -i kr nx (U W_q + Zeta W_u)(kr, y),   P_v(kr, y) = nx Zeta W_v(kr, y)        0 < kr < kcut,
// and the only eddy-eddy term the quasi-linear closure keeps, the kr = 0 entry of (v' zeta')_h, is the x-sum of the
// product, i.e. (Parseval) P_v(0, y) = nx sum_{kr != 0} W_v conj(W_q).  Columns >= kcut of both products are exactly
// zero in the reference as well (the inputs are dealiased and a kr = 0 factor does not spread in kr).
// One CTA per 8-row block of the tiled W/P (contiguous 256-byte tiles on one GPU); the row sums are reduced in a fixed
// order (deterministic).
static __global__ void __launch_bounds__(256) k_row_ql(const __grid_constant__ RowArgs a, const int nx) {
  __shared__ double part[256];
  __shared__ double UZ[16];
  const int t = threadIdx.x, yy = (t & 15) >> 1, kk = t & 1;
  const int y = (int)blockIdx.x * 8 + yy;
  if (t < 16) {  // U and Zeta of the block's eight rows
    const int r = (int)blockIdx.x * 8 + (t >> 1);
    UZ[t] = r < a.ny ? a.W[a.wl.off((t & 1) ? 2 : 0, r, 0)].x : 0.0;
  }
  __syncthreads();
  const double U = UZ[2 * yy], Z = UZ[2 * yy + 1], sx = (double)nx;
  const long long rd = a.pl.rowbase(0, y), rv = a.pl.rowbase(1, y);
  double acc = 0.0;
  const int npairs = (a.kout + 1) / 2;
  for (int kp = t >> 4; kp < npairs && y < a.ny; kp += 16) {
    const int k = 2 * kp + kk;
    if (k >= a.kout) continue;
    cplx d = cmake(0.0, 0.0), v = cmake(0.0, 0.0);
    if (k > 0 && k < a.kc) {
      cplx wu = a.W[a.wl.off(0, y, k)], wv = a.W[a.wl.off(1, y, k)], wq = a.W[a.wl.off(2, y, k)];
      const bool nyq = 2 * k == nx;  // c2r ignores the imaginary part of the Nyquist entry and counts it once
      if (nyq) wu.y = wv.y = wq.y = 0.0;
      const double kr = k * a.dkx;
      const cplx s = cmake(sx * (U * wq.x + Z * wu.x), sx * (U * wq.y + Z * wu.y));
      d = cmake(kr * s.y, -kr * s.x);
      v = cmake(sx * Z * wv.x, sx * Z * wv.y);
      acc += (nyq ? 1.0 : 2.0) * (wv.x * wq.x + wv.y * wq.y);
    }
    if (k > 0) {
      *prow_ptr(a.pout, rd, k) = d;
      *prow_ptr(a.pout, rv, k) = v;
    }
  }
  part[t] = acc;
  __syncthreads();
  if (t < 8 && (int)blockIdx.x * 8 + t < a.ny) {  // row t of the block: its 32 partial sums live at threads with ((tid & 15) >> 1) == t
    double ssum = 0.0;
    for (int i = 0; i < 16; ++i) ssum += part[i * 16 + 2 * t] + part[i * 16 + 2 * t + 1];
    const int r = (int)blockIdx.x * 8 + t;
    *prow_ptr(a.pout, a.pl.rowbase(0, r), 0) = cmake(0.0, 0.0);
    *prow_ptr(a.pout, a.pl.rowbase(1, r), 0) = cmake(sx * ssum, 0.0);
  }
}

static int run_row_product(gfb_ctx* c, const cplx* Wrecv, const PeerOut& Psend, bool prune, cudaStream_t st) {
  RowLaunch R;
  std::memset(&R, 0, sizeof R);
  R.mode = (c->kind == GFB_KIND_LAYERED) ? ROW_PRODUCT_Q : ROW_PRODUCT;
  R.a.W = Wrecv;
  R.a.wl = row_layout(c, 3 * c->nlayers);
  R.a.wstride = 3;
  R.a.kc = c->g.kcut;
  R.a.pout = Psend;
  R.a.pl = row_layout(c, 2 * c->nlayers);
  R.a.kout = prune ? c->g.kcut : c->g.nkr;
  R.a.eta = c->eta;
  R.a.U = c->U;
  R.a.Qx = c->Qx;
  R.a.Qy = c->Qy;
  R.a.linear = c->linear ? 1 : 0;
  R.a.dkx = c->g.dkx;
  R.a.ny = c->ny_loc;
  R.nlayers = c->nlayers;
  R.tw = c->tw_x;
  R.tw2 = c->tw_x2;
  R.st = st;
  // the TMA-staged kernel reads the ctx-owned W and writes the ctx-owned P (the tensor maps are bound to them)
  if (c->row_tma && Wrecv == c->W && Psend.nranks == 1 && Psend.base[0] == c->P) {
    R.tma = prune ? &c->rt_pruned : &c->rt_full;
    R.ws = c->row_ws;
    c->row_tma_launches += 1;
  }
  ProfScope ps(c, 1, st);
  if (c->ql) {
    GFB_LAUNCH(k_row_ql, dim3((c->ny_loc + 7) / 8), dim3(256), 0, st, R.a, c->g.nx);
    return check_launch(c, 0, "k_row_ql");
  }
  return check_launch(c, launch_row(c->g.nx, R), "k_row");
}

static int run_col_fwd(gfb_ctx* c, const EpiArgs& e, const cplx* Precv, cudaStream_t st) {
  ColFwdLaunch F;
  std::memset(&F, 0, sizeof F);
  F.g = c->g;
  F.P = Precv;
  F.ldp = 2 * c->nlayers * c->ldi;
  F.pfs = 8LL * c->ldi;
  F.e = e;
  F.ncols = e.prune ? local_cols_below(c, c->g.kcut) : c->g.nkl;
  F.nlayers = c->nlayers;
  F.nl = c->nlayers;
  F.two = true;
  F.ws = c->col_ws3;
  F.tw = c->tw_y;
  F.tw2 = c->tw_y2;
  F.st = st;
  if (F.ncols == 0) return 0;
  ProfScope ps(c, 2, st);
  return check_launch(c, launch_col_fwd(c->g.ny, F), "k_col_fwd");
}

static EpiArgs base_epi(const gfb_ctx* c, bool forced) {
  EpiArgs e;
  std::memset(&e, 0, sizeof e);
  e.nlayers = c->nlayers;
  e.F = forced ? int_spec(c, c->Fint) : null_spec();  // callers import the user's Fh into Fint first
  e.Nconst = c->Nconst ? int_spec(c, c->Nconst) : null_spec();
  e.cplxL = c->cplxL ? 1 : 0;
  e.L = coef(c, c->L);
  e.mu = (c->kind == GFB_KIND_LAYERED) ? c->cfg.mu : 0.0;
  e.Sinv = c->Sinv;
  e.ldS = c->g.nkl;
  e.filter.p = nullptr;
  return e;
}

static int nstages_of(const gfb_ctx* c) { return nstages_of_stepper(c->stepper); }

// Which array feeds stage `istage` and what its epilogue does.
// prune = 1 ("lazy alias band"): the aliased band of N -- which only ever lands in the aliased band of the next
// state, where the next stage-1 dealias! wipes it -- is not computed.
static int plan_stage(gfb_ctx* c, int istage, bool forced, bool want_prevsol, int prune, Spec& sin, EpiArgs& e) {
  const int ns = nstages_of(c);
  if (istage < 0 || istage >= ns) return fail(c, GFB_ERR_INVALID, "stage %d out of range", istage);
  e = base_epi(c, forced);
  e.prune = (prune && c->g.kcut < c->g.nkr) ? 1 : 0;
  const Spec S0 = int_spec(c, c->S0);
  const double dt = c->dt;
  if (c->filtered) {
    e.filter.p = c->filter;
    e.filter.ld = c->ldi;
    e.filter.ls = 0;
  }
  const Spec pv = want_prevsol ? int_spec(c, c->PV) : null_spec();
  sin = S0;
  e.sol = S0;
  switch (c->stepper) {
    case GFB_FORWARD_EULER: case GFB_FILTERED_FORWARD_EULER:
      e.kind = EPI_RK_ONLY;
      e.b = dt;
      e.out = S0;
      e.prevsol = pv;
      break;
    case GFB_RK4: case GFB_FILTERED_RK4: {
      const Spec s1 = int_spec(c, c->st[0]), s2 = int_spec(c, c->st[1]), acc = int_spec(c, c->st[2]);
      e.acc = acc;
      if (istage == 0) {
        e.kind = EPI_RK_FIRST; e.a = dt / 2; e.b = dt / 6; sin = S0; e.out = s1; e.prevsol = pv;
      } else if (istage == 1) {
        e.kind = EPI_RK_MID; e.a = dt / 2; e.b = dt / 3; sin = s1; e.out = s2;
      } else if (istage == 2) {
        e.kind = EPI_RK_MID; e.a = dt; e.b = dt / 3; sin = s2; e.out = s1;
      } else {
        e.kind = EPI_RK_LAST; e.b = dt / 6; sin = s1; e.out = S0;
      }
    } break;
    case GFB_ETDRK4: case GFB_FILTERED_ETDRK4: {
      const Spec s1 = int_spec(c, c->st[0]), s2 = int_spec(c, c->st[1]), acc = int_spec(c, c->st[2]);
      const Spec n1 = int_spec(c, c->st[3]), s3 = int_spec(c, c->st[4]);
      e.acc = acc;
      e.n1 = n1;
      e.c1 = coef(c, c->Eh);
      e.c2 = coef(c, c->zeta);
      if (istage == 0) {
        e.kind = EPI_ETD_1; sin = S0; e.out = s1; e.c3 = coef(c, c->E); e.c4 = coef(c, c->alpha); e.prevsol = pv;
      } else if (istage == 1) {
        e.kind = EPI_ETD_2; sin = s1; e.out = s2; e.c4 = coef(c, c->beta);
      } else if (istage == 2) {
        e.kind = EPI_ETD_3; sin = s2; e.sol = s1; e.out = s3; e.c4 = coef(c, c->beta);
      } else {
        e.kind = EPI_ETD_4; sin = s3; e.out = S0; e.c4 = coef(c, c->Gamma);
      }
    } break;
    case GFB_LSRK54: case GFB_FILTERED_LSRK54: {
      // [FF] LSRK54TimeStepper (Carpenter & Kennedy 1994, 2N-storage RK4(5)): per stage S2 = A_i S2 + dt (N + L sol),
      // sol += B_i S2 -- the state is updated IN PLACE (every element is read and written by the same thread)
      static const double A[5] = {0.0, -567301805773.0 / 1357537059087.0, -2404267990393.0 / 2016746695238.0,
                                  -3550918686646.0 / 2091501179385.0, -1275806237668.0 / 842570457699.0};
      static const double B[5] = {1432997174477.0 / 9575080441755.0, 5161836677717.0 / 13612068292357.0,
                                  1720146321549.0 / 2090206949498.0, 3134564353537.0 / 4481467310338.0,
                                  2277821191437.0 / 14882151754819.0};
      e.kind = EPI_LSRK;
      e.a = A[istage];
      e.b = B[istage];
      e.d = dt;
      e.acc = int_spec(c, c->st[0]);
      e.out = S0;
      if (c->kind == GFB_KIND_LAYERED) {
        // the bottom-drag term of a layer reads the stage input of ALL layers, which other CTAs would be overwriting:
        // the layered kind steps out of place, S0 -> A -> B -> A -> B -> S0
        const Spec A2 = int_spec(c, c->st[1]), B2 = int_spec(c, c->st[2]);
        sin = (istage == 0) ? S0 : ((istage & 1) ? A2 : B2);
        e.out = (istage == ns - 1) ? S0 : ((istage & 1) ? B2 : A2);
        e.sol = sin;
      }
      if (istage == 0) e.prevsol = pv;
      if (istage != ns - 1) e.filter.p = nullptr;  // the filter closes the step
    } break;
    default: {  // AB3
      const long long k = c->step;
      e.kind = EPI_AB3;
      if (k < 3) { e.a = dt; e.b = 0; e.d = 0; }
      else { e.a = dt * 23.0 / 12.0; e.b = -dt * 16.0 / 12.0; e.d = dt * 5.0 / 12.0; }
      e.out = S0;
      e.rhs = int_spec(c, c->st[k % 3]);
      e.acc = int_spec(c, c->st[(k + 2) % 3]);
      e.n1 = int_spec(c, c->st[(k + 1) % 3]);
      e.prevsol = pv;
    } break;
  }
  // Layered kind: the bottom-drag term of a layer reads the stage input of ALL layers (epi_addends), which the CTAs of the
  // other layers would be overwriting if a one-stage stepper updated S0 in place (seen on hardware as a run-to-run
  // varying 1e-9 error of the bottom layer under AB3).  ForwardEuler / AB3 therefore write the new state to ALT and
  // commit_stage swaps the two pointers once the stage is enqueued.
  if (c->ALT && e.out.p == S0.p) e.out = int_spec(c, c->ALT);
  e.sin = sin;
  return 0;
}
static void commit_stage(gfb_ctx* c, const EpiArgs& e) {
  if (c->ALT && e.out.p == c->ALT) std::swap(c->S0, c->ALT);
}

static bool is_ab3(const gfb_ctx* c) { return c->stepper == GFB_AB3 || c->stepper == GFB_FILTERED_AB3; }

// one stage on ctx-owned W/P (single GPU)
static int do_stage(gfb_ctx* c, int istage, bool forced, bool want_prevsol, int prune, cudaStream_t st) {
  Spec sin;
  EpiArgs e;
  int rc = plan_stage(c, istage, forced, want_prevsol, prune, sin, e);
  if (rc) return rc;
  if ((rc = run_col_inv_state(c, sin, 3, out_local(c->W), st))) return rc;
  if ((rc = run_row_product(c, c->W, out_local(c->P), e.prune != 0, st))) return rc;
  if ((rc = run_col_fwd(c, e, c->P, st))) return rc;
  commit_stage(c, e);
  if (istage == nstages_of(c) - 1) {
    c->t += c->dt;
    c->step += 1;
  }
  return 0;
}

// ---- "p2p" transport: the all-to-all is done by the kernels' own stores over NVLink -------------------------
// Flags: slot [s] of rank r's buffer is written by rank s.  A barrier = every rank stores the epoch into its slot
// on every peer (system-scope release after a system fence), then waits until all of its own slots carry it.
struct PeerFlags {
  unsigned long long* flags[MAX_RANKS];
  int rank, nranks;
};
constexpr int FLAG_SLOTS = 32;       // 256 bytes: [0..7] epochs, [16] timeout marker
constexpr int FLAG_TIMEOUT_SLOT = 16;

__device__ __forceinline__ void flag_store_release(unsigned long long* p, unsigned long long v) {
#ifdef GFB_EMULATE
  __atomic_store_n(p, v, __ATOMIC_RELEASE);
#else
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#endif
}
__device__ __forceinline__ unsigned long long flag_load_acquire(const unsigned long long* p) {
#ifdef GFB_EMULATE
  return __atomic_load_n(p, __ATOMIC_ACQUIRE);
#else
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
#endif
}

static __global__ void k_peer_barrier(const PeerFlags f, const unsigned long long epoch, const long long max_spins,
                                      unsigned long long* tmo) {
  const int d = threadIdx.x;
  if (d >= f.nranks) return;
#ifdef GFB_EMULATE
  __atomic_thread_fence(__ATOMIC_SEQ_CST);
#else
  __threadfence_system();
#endif
  flag_store_release(f.flags[d] + f.rank, epoch);
  const unsigned long long* mine = f.flags[f.rank] + d;
  long long spins = 0;
  while (flag_load_acquire(mine) < epoch) {
#ifdef GFB_EMULATE
    std::this_thread::yield();
#else
    __nanosleep(64);
#endif
    if (++spins > max_spins) {  // a peer died or never launched: give up instead of hanging the GPU
      flag_store_release(f.flags[f.rank] + FLAG_TIMEOUT_SLOT, epoch);
      if (tmo) flag_store_release(tmo, epoch);
      break;
    }
  }
}

static int peer_barrier(gfb_ctx* c, cudaStream_t st) {
  PeerFlags f;
  std::memset(&f, 0, sizeof f);
  f.rank = c->g.rank;
  f.nranks = c->g.nranks;
  for (int d = 0; d < f.nranks; ++d) f.flags[d] = reinterpret_cast<unsigned long long*>(c->peer[d]);
  if (c->tmo_host && *(volatile unsigned long long*)c->tmo_host != 0)
    return fail(c, GFB_ERR_CUDA, "peer barrier timed out at epoch %llu (a peer died or never launched); the state is invalid",
                *(volatile unsigned long long*)c->tmo_host);
  c->epoch += 1;
  ProfScope ps(c, 3, st);
  static long long spins = -1;
  if (spins < 0) {
    const char* e = getenv("GFB_BARRIER_SPINS");
    spins = e ? atoll(e) : 40000000LL;
  }
  GFB_LAUNCH(k_peer_barrier, dim3(1), dim3(32), 0, st, f, c->epoch, spins, c->tmo_dev);
  return check_launch(c, 0, "k_peer_barrier");
}

extern "C" {

// Allocate this rank's exchange buffer.  handle64_out (may be null) receives the cudaIpcMemHandle_t other
// PROCESSES open with gfb_comm_connect; local_ptr_out (may be null) the raw pointer for same-process peers.
int gfb_comm_create(gfb_ctx* c, void* handle64_out, void** local_ptr_out) {
  if (!c) return GFB_ERR_INVALID;
  if (c->g.nranks < 2) return fail(c, GFB_ERR_INVALID, "gfb_comm_create: context is not sharded");
  if (!c->comm) {
    // W: u, v, q (+ psi for updatevars!) ; P: d, v
    const size_t wb = (size_t)w_elems(c, 4 * c->nlayers) * sizeof(cplx), pb = (size_t)w_elems(c, 2 * c->nlayers) * sizeof(cplx);
    c->comm_w_off = 256;
    c->comm_p_off = c->comm_w_off + wb;
    c->comm_bytes = c->comm_p_off + pb;
    int rc = dev_alloc(c, &c->comm, c->comm_bytes);
    if (rc) return rc;
    GFB_CUDA(c, cudaMemset(c->comm, 0, c->comm_bytes));
    GFB_CUDA(c, cudaDeviceSynchronize());
    if (!c->tmo_host) {
      GFB_CUDA(c, cudaHostAlloc((void**)&c->tmo_host, sizeof(unsigned long long), cudaHostAllocMapped));
      *c->tmo_host = 0;
      GFB_CUDA(c, cudaHostGetDevicePointer((void**)&c->tmo_dev, c->tmo_host, 0));
    }
  }
  if (local_ptr_out) *local_ptr_out = c->comm;
  if (handle64_out) {
#ifdef GFB_EMULATE
    std::memset(handle64_out, 0, 64);
#else
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    GFB_CUDA(c, cudaIpcGetMemHandle(&h, c->comm));
    std::memcpy(handle64_out, &h, 64);
#endif
  }
  return GFB_OK;
}

// handles: nranks x 64 bytes, entry d = rank d's handle from gfb_comm_create (one process per GPU)
int gfb_comm_connect(gfb_ctx* c, const void* handles) {
  if (!c || !handles || !c->comm) return fail(c, GFB_ERR_INVALID, "gfb_comm_connect: bad argument (call gfb_comm_create first)");
#ifdef GFB_EMULATE
  return fail(c, GFB_ERR_UNSUPPORTED, "gfb_comm_connect: CUDA IPC only");
#else
  for (int d = 0; d < c->g.nranks; ++d) {
    if (d == c->g.rank) {
      c->peer[d] = c->comm;
      continue;
    }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, (const char*)handles + 64 * d, 64);
    void* p = nullptr;
    GFB_CUDA(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer[d] = (char*)p;
    c->peer_ipc[d] = true;
  }
  c->connected = true;
  return GFB_OK;
#endif
}

// same-process variant (several contexts in one process, peer access already enabled by the caller)
int gfb_comm_connect_ptrs(gfb_ctx* c, void* const* ptrs) {
  if (!c || !ptrs || !c->comm) return fail(c, GFB_ERR_INVALID, "gfb_comm_connect_ptrs: bad argument (call gfb_comm_create first)");
  for (int d = 0; d < c->g.nranks; ++d) c->peer[d] = (d == c->g.rank) ? c->comm : (char*)ptrs[d];
  c->connected = true;
  return GFB_OK;
}

// 0 = fine; GFB_ERR_CUDA if a barrier gave up waiting for a peer (synchronises the device)
int gfb_comm_status(gfb_ctx* c) {
  if (!c || !c->comm) return GFB_ERR_INVALID;
  unsigned long long v = 0;
  GFB_CUDA(c, cudaDeviceSynchronize());
  GFB_CUDA(c, cudaMemcpy(&v, c->comm + 8 * FLAG_TIMEOUT_SLOT, 8, cudaMemcpyDeviceToHost));
  if (v != 0) return fail(c, GFB_ERR_CUDA, "peer barrier timed out at epoch %llu", v);
  return GFB_OK;
}

// stepforward! on a slab-sharded problem with the p2p transport: per stage
//   K1 (stores W rows into their owners' buffers) -> barrier -> K2 (stores P column pairs into their owners'
//   buffers) -> barrier -> K3.
// One receive buffer each suffices: K1 of the next stage starts after this rank's K3, i.e. after the barrier that
// proves every rank finished reading W; likewise for P.
int gfb_step_p2p(gfb_ctx* c, void* sol_local, int nsteps, const void* Fh_local, void* prevsol_local, void* stream) {
  if (!c || !sol_local || nsteps < 0) return fail(c, GFB_ERR_INVALID, "gfb_step_p2p: bad argument");
  if (!c->connected) return fail(c, GFB_ERR_INVALID, "gfb_step_p2p: peers not connected");
  if (nsteps == 0) return GFB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int ns = nstages_of(c), n = c->nlayers;
  int rc = import_spec(c, c->S0, sol_local, st);
  if (rc) return rc;
  if (Fh_local && (rc = import_spec(c, c->Fint, Fh_local, st))) return rc;
  const PeerOut wout = out_peers(c, c->comm_w_off, 3 * n), pout = out_peers(c, c->comm_p_off, 2 * n);
  const cplx* wrecv = reinterpret_cast<const cplx*>(c->comm + c->comm_w_off);
  const cplx* precv = reinterpret_cast<const cplx*>(c->comm + c->comm_p_off);
  for (int s = 0; s < nsteps; ++s) {
    const bool last = (s == nsteps - 1);
    const int prune = ((!last || c->defer_alias) && !is_ab3(c)) ? 1 : 0;
    for (int i = 0; i < ns; ++i) {
      Spec sin;
      EpiArgs e;
      if ((rc = plan_stage(c, i, Fh_local != nullptr, last && prevsol_local != nullptr, prune, sin, e))) return rc;
      if ((rc = run_col_inv_state(c, sin, 3, wout, st))) return rc;
      if ((rc = peer_barrier(c, st))) return rc;
      if ((rc = run_row_product(c, wrecv, pout, e.prune != 0, st))) return rc;
      if ((rc = peer_barrier(c, st))) return rc;
      if ((rc = run_col_fwd(c, e, precv, st))) return rc;
      commit_stage(c, e);
    }
    c->t += c->dt;
    c->step += 1;
  }
  if (prevsol_local && (rc = export_spec(c, prevsol_local, c->PV, st))) return rc;
  if ((rc = export_spec(c, sol_local, c->S0, st))) return rc;
  // asynchronous error model: a barrier of THIS call that gives up is reported by the next call that launches a barrier
  // (or by gfb_comm_status, which synchronises); one that already gave up is reported here
  if (c->tmo_host && *(volatile unsigned long long*)c->tmo_host != 0)
    return fail(c, GFB_ERR_CUDA, "peer barrier timed out at epoch %llu; the state is invalid", *(volatile unsigned long long*)c->tmo_host);
  return GFB_OK;
}

}  // extern "C"

template <int NL>
static void launch_matvec(const gfb_ctx* c, cplx* y, const cplx* x, const double* M, cudaStream_t st) {
  GFB_LAUNCH((k_matvec<NL>), dim3(592), dim3(256), 0, st, c->g.ny, c->g.nkl, y, x, M);
}

static int matvec(gfb_ctx* c, void* y, const void* x, const double* M, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  switch (c->nlayers) {
    case 1: launch_matvec<1>(c, (cplx*)y, (const cplx*)x, M, st); break;
    case 2: launch_matvec<2>(c, (cplx*)y, (const cplx*)x, M, st); break;
    case 3: launch_matvec<3>(c, (cplx*)y, (const cplx*)x, M, st); break;
    case 4: launch_matvec<4>(c, (cplx*)y, (const cplx*)x, M, st); break;
    case 5: launch_matvec<5>(c, (cplx*)y, (const cplx*)x, M, st); break;
    default: return fail(c, GFB_ERR_UNSUPPORTED, "matvec: nlayers");
  }
  return check_launch(c, 0, "k_matvec");
}

static __global__ void k_scale_by_table(const int nl, const int nkr, cplx* __restrict__ y, const cplx* __restrict__ x,
                                 const double* __restrict__ G, const int ldg, const double sign) {
  const long long n = (long long)nl * nkr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int l = (int)(i / nkr), k = (int)(i - (long long)l * nkr);
    y[i] = cscale(x[i], sign * G[toff(l, k, ldg)]);
  }
}

template <int NL>
static void launch_diag(const gfb_ctx* c, const DiagArgs& d, cudaStream_t st) {
  GFB_LAUNCH((k_diag_partial<NL>), dim3(c->dblocks), dim3(256), 0, st, c->g, d, c->dpart);
}

#define NEED_SINGLE_GPU(c, what) \
  if ((c)->g.nranks != 1) return fail((c), GFB_ERR_UNSUPPORTED, what ": not available on a sharded context (use the gfb_stage_pre/mid/post path)")

extern "C" {

int gfb_dealias(gfb_ctx* c, void* sol, void* stream) {
  if (!c || !sol) return fail(c, GFB_ERR_INVALID, "gfb_dealias: null argument");
  GFB_LAUNCH(k_dealias, dim3(592), dim3(256), 0, (cudaStream_t)stream, c->gd, (cplx*)sol, c->g.nkl,
             (long long)c->g.ny * c->g.nkl, c->nlayers);
  return check_launch(c, 0, "k_dealias");
}

int gfb_calcN(gfb_ctx* c, void* N, void* sol, const void* Fh, void* stream) {
  if (!c || !N || !sol) return fail(c, GFB_ERR_INVALID, "gfb_calcN: null argument");
  NEED_SINGLE_GPU(c, "gfb_calcN");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = c->linear ? 0 : gfb_dealias(c, sol, stream);  // calcNlinear! leaves sol alone
  if (rc) return rc;
  if ((rc = import_spec(c, c->Cin, sol, st))) return rc;
  if (Fh && (rc = import_spec(c, c->Fint, Fh, st))) return rc;
  EpiArgs e = base_epi(c, Fh != nullptr);
  e.kind = EPI_STORE;
  e.out = int_spec(c, c->Nint);
  const Spec sin = int_spec(c, c->Cin);
  e.sin = sin;
  if ((rc = run_col_inv_state(c, sin, 3, out_local(c->W), st))) return rc;
  if ((rc = run_row_product(c, c->W, out_local(c->P), false, st))) return rc;
  if ((rc = run_col_fwd(c, e, c->P, st))) return rc;
  return export_spec(c, N, c->Nint, st);
}

int gfb_nstages(const gfb_ctx* c) { return c ? nstages_of(c) : GFB_ERR_INVALID; }

int gfb_set_defer_alias_band(gfb_ctx* c, int on) {
  if (!c) return GFB_ERR_INVALID;
  c->defer_alias = on != 0;
  return GFB_OK;
}

int gfb_stage(gfb_ctx* c, void* sol, int istage, const void* Fh, void* prevsol, void* stream) {
  if (!c || !sol) return fail(c, GFB_ERR_INVALID, "gfb_stage: null argument");
  NEED_SINGLE_GPU(c, "gfb_stage");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = 0;
  if (istage == 0 && (rc = import_spec(c, c->S0, sol, st))) return rc;  // later stages continue from the ctx copy
  if (Fh && (rc = import_spec(c, c->Fint, Fh, st))) return rc;
  if ((rc = do_stage(c, istage, Fh != nullptr, prevsol != nullptr, 0, st))) return rc;
  if (istage == 0 && prevsol && (rc = export_spec(c, prevsol, c->PV, st))) return rc;
  if (istage == nstages_of(c) - 1) return export_spec(c, sol, c->S0, st);
  return GFB_OK;
}

int gfb_step(gfb_ctx* c, void* sol, int nsteps, const void* Fh, void* prevsol, void* stream) {
  if (!c || !sol || nsteps < 0) return fail(c, GFB_ERR_INVALID, "gfb_step: bad argument");
  NEED_SINGLE_GPU(c, "gfb_step");
  if (nsteps == 0) return GFB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int ns = nstages_of(c);
  int rc = import_spec(c, c->S0, sol, st);
  if (rc) return rc;
  if (Fh && (rc = import_spec(c, c->Fint, Fh, st))) return rc;
  for (int s = 0; s < nsteps; ++s) {
    const bool last = (s == nsteps - 1);
    for (int i = 0; i < ns; ++i)
      if ((rc = do_stage(c, i, Fh != nullptr, last && prevsol != nullptr, ((!last || c->defer_alias) && !is_ab3(c)) ? 1 : 0, st))) return rc;
  }
  if (prevsol && (rc = export_spec(c, prevsol, c->PV, st))) return rc;
  return export_spec(c, sol, c->S0, st);
}

// ---- sharded (multi-GPU) stepping: the caller owns the all-to-all buffers and runs the exchange -------------
int gfb_state_import(gfb_ctx* c, const void* sol_local, void* stream) {
  if (!c || !sol_local) return fail(c, GFB_ERR_INVALID, "gfb_state_import: null argument");
  return import_spec(c, c->S0, sol_local, (cudaStream_t)stream);
}
int gfb_state_export(gfb_ctx* c, void* sol_local, void* stream) {
  if (!c || !sol_local) return fail(c, GFB_ERR_INVALID, "gfb_state_export: null argument");
  return export_spec(c, sol_local, c->S0, (cudaStream_t)stream);
}
int gfb_stage_pre(gfb_ctx* c, int istage, void* wsend, void* stream) {
  if (!c || !wsend) return fail(c, GFB_ERR_INVALID, "gfb_stage_pre: null argument");
  Spec sin;
  EpiArgs e;
  int rc = plan_stage(c, istage, false, false, 0, sin, e);
  if (rc) return rc;
  return run_col_inv_state(c, sin, 3, out_send(c, (cplx*)wsend, 3 * c->nlayers), (cudaStream_t)stream);
}
int gfb_stage_mid(gfb_ctx* c, const void* wrecv, void* psend, int prune, void* stream) {
  if (!c || !wrecv || !psend) return fail(c, GFB_ERR_INVALID, "gfb_stage_mid: null argument");
  return run_row_product(c, (const cplx*)wrecv, out_send(c, (cplx*)psend, 2 * c->nlayers), prune && c->g.kcut < c->g.nkr,
                         (cudaStream_t)stream);
}
int gfb_stage_post(gfb_ctx* c, int istage, const void* precv, const void* Fh_local, void* prevsol_local, int prune,
                   void* stream) {
  if (!c || !precv) return fail(c, GFB_ERR_INVALID, "gfb_stage_post: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = 0;
  if (Fh_local && (rc = import_spec(c, c->Fint, Fh_local, st))) return rc;
  Spec sin;
  EpiArgs e;
  if ((rc = plan_stage(c, istage, Fh_local != nullptr, prevsol_local != nullptr, prune, sin, e))) return rc;
  if ((rc = run_col_fwd(c, e, (const cplx*)precv, st))) return rc;
  commit_stage(c, e);
  if (istage == 0 && prevsol_local && (rc = export_spec(c, prevsol_local, c->PV, st))) return rc;
  if (istage == nstages_of(c) - 1) {
    c->t += c->dt;
    c->step += 1;
  }
  return GFB_OK;
}

// Sharded context (p2p transport connected): in_spec is the LOCAL dense [batch][nl][nkl] slice, out_phys the LOCAL rows
// [batch][ny_loc][nx]; K1's stores are the exchange, a flag barrier before K2 and one after it (the receive buffer is
// shared with the stepping passes).
int gfb_irfft(gfb_ctx* c, double* out_phys, const void* in_spec, int batch, void* stream) {
  if (!c || !out_phys || !in_spec || batch < 1) return fail(c, GFB_ERR_INVALID, "gfb_irfft: bad argument");
  const bool sharded = c->g.nranks != 1;
  if (sharded && !c->connected)
    return fail(c, GFB_ERR_UNSUPPORTED, "gfb_irfft on a sharded context needs the p2p transport (gfb_comm_create / gfb_comm_connect)");
  cudaStream_t st = (cudaStream_t)stream;
  const int maxb = (sharded ? 3 : 4) * c->nlayers;
  const long long spec_ls = (long long)c->g.ny * c->g.nkl, phys_ls = (long long)c->ny_loc * c->g.nx;
  for (int b0 = 0; b0 < batch; b0 += maxb) {
    const int nb = (batch - b0 < maxb) ? batch - b0 : maxb;
    ColInvLaunch L;
    std::memset(&L, 0, sizeof L);
    L.g = c->g;
    L.prep = PREP_PLAIN;
    L.pp.in = (const cplx*)in_spec + (long long)b0 * spec_ls;
    L.pp.ld = c->g.nkl;
    L.pp.ls = spec_ls;
    L.pp.scale = c->g.invN;
    L.W = sharded ? out_peers(c, c->comm_w_off, nb) : out_local(c->W);
    L.ldw = nb * c->ldi;
    L.wfs = 8LL * c->ldi;
    L.ncols = c->g.nkl;
    L.nfields = nb;
    L.tw = c->tw_y;
    L.tw2 = c->tw_y2;
    L.st = st;
    L.nl = 1;
    int rc = check_launch(c, launch_col_inv(c->g.ny, L), "k_col_inv(plain)");
    if (rc) return rc;
    if (sharded && (rc = peer_barrier(c, st))) return rc;
    RowLaunch R;
    std::memset(&R, 0, sizeof R);
    R.mode = ROW_C2R;
    R.a.W = sharded ? reinterpret_cast<const cplx*>(c->comm + c->comm_w_off) : c->W;
    R.a.wl = row_layout(c, nb);
    R.a.wstride = 1;
    R.a.fsel = 0;
    R.a.kc = c->g.nkr;
    R.a.phys = out_phys + (long long)b0 * phys_ls;
    R.a.ny = c->ny_loc;
    R.nlayers = nb;
    R.tw = c->tw_x;
    R.tw2 = c->tw_x2;
    R.st = st;
    rc = check_launch(c, launch_row(c->g.nx, R), "k_row(c2r)");
    if (rc) return rc;
    if (sharded && (rc = peer_barrier(c, st))) return rc;
  }
  return GFB_OK;
}

// Sharded context: in_phys is the LOCAL rows [batch][ny_loc][nx], out_spec the LOCAL dense [batch][nl][nkl] slice.
int gfb_rfft(gfb_ctx* c, void* out_spec, const double* in_phys, int batch, void* stream) {
  if (!c || !out_spec || !in_phys || batch < 1) return fail(c, GFB_ERR_INVALID, "gfb_rfft: bad argument");
  const bool sharded = c->g.nranks != 1;
  if (sharded && !c->connected)
    return fail(c, GFB_ERR_UNSUPPORTED, "gfb_rfft on a sharded context needs the p2p transport (gfb_comm_create / gfb_comm_connect)");
  cudaStream_t st = (cudaStream_t)stream;
  const int maxb = 2 * c->nlayers;
  const long long spec_ls = (long long)c->g.ny * c->g.nkl, phys_ls = (long long)c->ny_loc * c->g.nx;
  for (int b0 = 0; b0 < batch; b0 += maxb) {
    const int nb = (batch - b0 < maxb) ? batch - b0 : maxb;
    RowLaunch R;
    std::memset(&R, 0, sizeof R);
    R.mode = ROW_R2C;
    R.a.pout = sharded ? out_peers(c, c->comm_p_off, nb) : out_local(c->P);
    R.a.pl = row_layout(c, nb);
    R.a.kout = c->g.nkr;
    R.a.phys = const_cast<double*>(in_phys) + (long long)b0 * phys_ls;
    R.a.ny = c->ny_loc;
    R.nlayers = nb;
    R.tw = c->tw_x;
    R.tw2 = c->tw_x2;
    R.st = st;
    int rc = check_launch(c, launch_row(c->g.nx, R), "k_row(r2c)");
    if (rc) return rc;
    if (sharded && (rc = peer_barrier(c, st))) return rc;
    ColFwdLaunch F;
    std::memset(&F, 0, sizeof F);
    F.g = c->g;
    F.P = sharded ? reinterpret_cast<const cplx*>(c->comm + c->comm_p_off) : c->P;
    F.ldp = nb * c->ldi;
    F.pfs = 8LL * c->ldi;
    F.e.kind = EPI_PLAIN;
    F.e.out.p = (cplx*)out_spec + (long long)b0 * spec_ls;
    F.e.out.ld = c->g.nkl;
    F.e.out.ls = spec_ls;
    F.ncols = c->g.nkl;
    F.nlayers = nb;
    F.nl = 1;
    F.two = false;
    F.tw = c->tw_y;
    F.tw2 = c->tw_y2;
    F.st = st;
    rc = check_launch(c, launch_col_fwd(c->g.ny, F), "k_col_fwd(plain)");
    if (rc) return rc;
    if (sharded && (rc = peer_barrier(c, st))) return rc;
  }
  return GFB_OK;
}

// Sharded context (p2p transport connected): `sol` is the local spectral slice, the outputs are this rank's rows.
int gfb_updatevars(gfb_ctx* c, void* sol, double* q, double* psi, double* u, double* v, void* stream) {
  if (!c || !sol) return fail(c, GFB_ERR_INVALID, "gfb_updatevars: null argument");
  const bool sharded = c->g.nranks != 1;
  if (sharded && !c->connected)
    return fail(c, GFB_ERR_UNSUPPORTED, "gfb_updatevars on a sharded context needs the p2p transport (gfb_comm_create / gfb_comm_connect)");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = gfb_dealias(c, sol, stream);
  if (rc) return rc;
  const int nf = 4;
  if ((rc = import_spec(c, c->Cin, sol, st))) return rc;
  rc = run_col_inv_state(c, int_spec(c, c->Cin), nf, sharded ? out_peers(c, c->comm_w_off, nf * c->nlayers) : out_local(c->W), st);
  if (rc) return rc;
  if (sharded && (rc = peer_barrier(c, st))) return rc;
  double* outs[4] = {u, v, q, psi};
  for (int f = 0; f < 4; ++f) {
    if (!outs[f]) continue;
    RowLaunch R;
    std::memset(&R, 0, sizeof R);
    R.mode = ROW_C2R;
    R.a.W = sharded ? reinterpret_cast<const cplx*>(c->comm + c->comm_w_off) : c->W;
    R.a.wl = row_layout(c, nf * c->nlayers);
    R.a.wstride = nf;
    R.a.fsel = f;
    R.a.kc = c->g.kcut;
    R.a.phys = outs[f];
    R.a.ny = c->ny_loc;
    R.nlayers = c->nlayers;
    R.tw = c->tw_x;
    R.tw2 = c->tw_x2;
    R.st = st;
    rc = check_launch(c, launch_row(c->g.nx, R), "k_row(c2r,updatevars)");
    if (rc) return rc;
  }
  if (sharded && (rc = peer_barrier(c, st))) return rc;
  return GFB_OK;
}

int gfb_streamfunction_from_pv(gfb_ctx* c, void* psih, const void* qh, void* stream) {
  if (!c || !psih || !qh) return fail(c, GFB_ERR_INVALID, "gfb_streamfunction_from_pv: null argument");
  if (c->kind == GFB_KIND_SINGLE) {
    GFB_LAUNCH(k_scale_by_table, dim3(592), dim3(256), 0, (cudaStream_t)stream, c->g.ny, c->g.nkl,
               (cplx*)psih, (const cplx*)qh, (const double*)c->G, c->ldi, -1.0);
    return check_launch(c, 0, "k_scale_by_table");
  }
  return matvec(c, psih, qh, c->Sinv, stream);
}

int gfb_pv_from_streamfunction(gfb_ctx* c, void* qh, const void* psih, void* stream) {
  if (!c || !qh || !psih) return fail(c, GFB_ERR_INVALID, "gfb_pv_from_streamfunction: null argument");
  if (c->kind == GFB_KIND_SINGLE || !c->S) return fail(c, GFB_ERR_UNSUPPORTED, "gfb_pv_from_streamfunction needs GFB_KIND_LAYERED with config.S");
  return matvec(c, qh, psih, c->S, stream);
}

// On a sharded context the sums are LOCAL partial sums: the caller all-reduces them.
int gfb_diag(gfb_ctx* c, int kind, const void* sol, const void* aux, const void* aux2, double p,
             double* out_host, int nout, void* stream) {
  if (!c || !sol || !out_host || nout < 1 || nout > DIAG_MAXOUT) return fail(c, GFB_ERR_INVALID, "gfb_diag: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  DiagArgs d;
  std::memset(&d, 0, sizeof d);
  d.kind = kind;
  d.s = (const cplx*)sol;
  d.aux = (const cplx*)aux;
  d.aux2 = (const cplx*)aux2;
  d.G = c->G;
  d.ldg = c->ldi;
  d.Sinv = c->Sinv;
  d.p = p;
  d.scale = 1.0;
  int need = 1;
  if (kind == GFB_DIAG_LAYERED_ENERGIES) {
    if (c->kind != GFB_KIND_LAYERED) return fail(c, GFB_ERR_INVALID, "gfb_diag: layered energies need GFB_KIND_LAYERED");
    need = 2 * c->nlayers - 1;
  }
  if ((kind == GFB_DIAG_G_ABS2 || kind == GFB_DIAG_G2_ABS2 || kind == GFB_DIAG_KG2_ABS2) && !c->G)
    return fail(c, GFB_ERR_INVALID, "gfb_diag: kind %d needs the G table", kind);
  if ((kind == GFB_DIAG_ABS2_PLUS || kind == GFB_DIAG_WORK || kind == GFB_DIAG_WORK_INVK) && !aux)
    return fail(c, GFB_ERR_INVALID, "gfb_diag: kind %d needs aux", kind);
  if (kind < 0 || kind > GFB_DIAG_LAYERED_ENERGIES) return fail(c, GFB_ERR_INVALID, "gfb_diag: unknown kind %d", kind);
  if (nout < need) return fail(c, GFB_ERR_INVALID, "gfb_diag: nout=%d < %d", nout, need);
  if (kind == GFB_DIAG_LAYERED_ENERGIES) {
    switch (c->nlayers) {
      case 1: launch_diag<1>(c, d, st); break;
      case 2: launch_diag<2>(c, d, st); break;
      case 3: launch_diag<3>(c, d, st); break;
      case 4: launch_diag<4>(c, d, st); break;
      case 5: launch_diag<5>(c, d, st); break;
    }
  } else {
    launch_diag<1>(c, d, st);
  }
  int rc = check_launch(c, 0, "k_diag_partial");
  if (rc) return rc;
  GFB_LAUNCH(k_diag_final, dim3(1), dim3(256), 0, st, (const double*)c->dpart, c->dblocks, need, c->dres_dev);
  rc = check_launch(c, 0, "k_diag_final");
  if (rc) return rc;
  GFB_CUDA(c, cudaStreamSynchronize(st));  // kernel completion makes its stores to the mapped buffer visible to the host
  std::memcpy(out_host, c->dres_host, need * sizeof(double));
  return GFB_OK;
}

}  // extern "C"
