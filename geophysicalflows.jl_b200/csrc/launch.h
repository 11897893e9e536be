// launch.h -- size-dispatched launchers.  Each supported FFT length has its own translation unit
// (launchers.cu compiled with -DGFB_N=<length>) so the library builds in parallel.
#pragma once
#include "kernels.cuh"

namespace gfb {

enum PrepKind { PREP_PLAIN = 0, PREP_SINGLE = 1, PREP_MULTI = 2 };

struct ColInvLaunch {
  GridDesc g;
  int prep;            // PrepKind
  PrepPlain pp;
  PrepSingle ps;
  PrepMulti pm;
  int nl;              // layers (PREP_MULTI template argument)
  PeerOut W;
  int ldw;
  long long wfs;
  int ncols;
  int nfields;         // grid.y
  bool ws;             // PREP_SINGLE at ny = 4096: the warp-specialised kernel (k_col_inv_ws)
  const cplx* tw;
  const cplx* tw2;     // W_{ny}^i when the columns are cut 4 x (ny/4) (ny = 16384), else unused
  cudaStream_t st;
};

struct RowLaunch {
  int mode;            // RowMode
  RowArgs a;
  const RowTma* tma;   // non-null: TMA-staged product kernel (single GPU, single-layer kinds, NX >= 4096)
  bool ws;             // with tma at nx = 4096: the warp-specialised kernel (k_row_product_ws)
  int nlayers;         // grid.y
  const cplx* tw;
  const cplx* tw2;     // W_{nx}^i when the row transforms run at half length (nx = 16384), else unused
  cudaStream_t st;
};

struct ColFwdLaunch {
  GridDesc g;
  const cplx* P;
  int ldp;
  long long pfs;
  EpiArgs e;
  int ncols;
  int nlayers;         // grid.y
  int nl;              // template NL (layers coupled in the epilogue)
  bool two;
  bool ws;             // ny = 4096, two = true: the warp-specialised kernel (k_col_fwd_ws)
  const cplx* tw;
  const cplx* tw2;     // as in ColInvLaunch
  cudaStream_t st;
};

constexpr int MAX_LAYERS = 5;

#define GFB_DECL_SIZE(N)                            \
  int launch_col_inv_##N(const ColInvLaunch& L);    \
  int launch_row_##N(const RowLaunch& L);           \
  int launch_col_fwd_##N(const ColFwdLaunch& L);
GFB_DECL_SIZE(32)
GFB_DECL_SIZE(64)
GFB_DECL_SIZE(128)
GFB_DECL_SIZE(256)
GFB_DECL_SIZE(512)
GFB_DECL_SIZE(1024)
GFB_DECL_SIZE(2048)
GFB_DECL_SIZE(4096)
GFB_DECL_SIZE(8192)
#undef GFB_DECL_SIZE

int launch_row_half_8192(const RowLaunch& L);
int launch_col_inv_16384(const ColInvLaunch& L);  // columns of 16384 points as 4 x 4096 (radix-4 DIF while loading)
int launch_col_fwd_16384(const ColFwdLaunch& L);  // rows of 16384 points as 8192-point half-length transforms

int launch_col_inv_dft(const ColInvLaunch& L);    // any even ny <= DFT_MAX_NY: direct transforms (cols_dft.cu)
int launch_col_fwd_dft(const ColFwdLaunch& L);
constexpr int DFT_MAX_NY = 1024;

int launch_col_inv(int NY, const ColInvLaunch& L);
int launch_row(int NX, const RowLaunch& L);
int launch_col_fwd(int NY, const ColFwdLaunch& L);

}  // namespace gfb

