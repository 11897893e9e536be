// aux_kernels.cuh -- repack / dealias / matvec / diagnostics kernels
// (part of kernels.cuh: include that, not this file)
#pragma once
#include "layout.cuh"

namespace gfb {

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
