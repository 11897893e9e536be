/* test_abi.c -- the C ABI exercised from plain C through include/gfb.h itself (no ctypes mirror in between).
 *
 *   test_abi          CPU box: the header compiles as C99, the struct the compiler lays out has the size the library
 *                     reports, argument errors come back as codes, and a valid config fails only at the first CUDA
 *                     call (no GPU here) -- never a crash.
 *   test_abi --gpu    B200: create a TwoDNavierStokes-like problem (RK4, 64 x 64), take steps through gfb_step and
 *                     check that the zero state stays zero and a single Fourier mode decays by exp(L dt) to 1e-12.
 */
#define _USE_MATH_DEFINES
#define _GNU_SOURCE
#include <math.h>
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gfb.h"

/* libgfb.so brings the CUDA runtime into the process; the three runtime calls the GPU leg needs are looked up there */
#include <dlfcn.h>
typedef int (*malloc_fn)(void**, size_t);
typedef int (*memcpy_fn)(void*, const void*, size_t, int);
typedef int (*free_fn)(void*);

#define CHECK(cond)                                                          \
  do {                                                                       \
    if (!(cond)) {                                                           \
      printf("FAILED %s:%d: %s (%s)\n", __FILE__, __LINE__, #cond, gfb_last_error(NULL)); \
      return 1;                                                              \
    }                                                                        \
  } while (0)

int main(int argc, char** argv) {
  const int gpu = argc > 1 && strcmp(argv[1], "--gpu") == 0;
  const int n = 64, nkr = n / 2 + 1;
  gfb_config cfg;
  gfb_ctx* ctx = NULL;
  memset(&cfg, 0, sizeof cfg);
  CHECK(gfb_config_size() == (int)sizeof(gfb_config));
  CHECK(strstr(gfb_version(), "sm_100a") != NULL);
  cfg.nx = cfg.ny = n;
  cfg.nlayers = 1;
  CHECK(gfb_create(&ctx, &cfg) == GFB_ERR_INVALID && ctx == NULL); /* struct_size missing */
  cfg.struct_size = (int)sizeof cfg;
  cfg.nx = 100;
  CHECK(gfb_create(&ctx, &cfg) == GFB_ERR_UNSUPPORTED);
  cfg.nx = n;
  cfg.kind = GFB_KIND_SINGLE;
  cfg.Lx = cfg.Ly = 2.0 * M_PI;
  cfg.kr_alias0 = nkr; /* no dealiasing */
  cfg.l_alias0 = cfg.l_alias1 = 0;
  cfg.stepper = GFB_RK4;
  cfg.dt = 1e-3;
  CHECK(gfb_create(&ctx, &cfg) == GFB_ERR_INVALID); /* L missing */
  {
    double* L = (double*)malloc(sizeof(double) * n * nkr);
    double* G = (double*)malloc(sizeof(double) * n * nkr);
    int l, k, rc;
    for (l = 0; l < n; ++l)
      for (k = 0; k < nkr; ++k) {
        const double ll = l < n / 2 ? l : l - n, K2 = (double)k * k + ll * ll;
        L[l * nkr + k] = -0.5 * K2;            /* nu = 0.5, nnu = 1 */
        G[l * nkr + k] = K2 > 0 ? 1.0 / K2 : 0; /* invKrsq */
      }
    cfg.L = L;
    cfg.G = G;
    rc = gfb_create(&ctx, &cfg);
    if (!gpu) {
      /* no device in this container: the first CUDA call fails and is reported, nothing crashes */
      CHECK(rc == GFB_ERR_CUDA || rc == GFB_ERR_ALLOC);
      CHECK(ctx == NULL && strlen(gfb_last_error(NULL)) > 0);
    } else {
      double* h = (double*)calloc((size_t)2 * n * nkr, sizeof(double));
      void* d = NULL;
      double t = 0, amp, want;
      long long step = 0;
      malloc_fn cudaMalloc = (malloc_fn)dlsym(RTLD_DEFAULT, "cudaMalloc");
      memcpy_fn cudaMemcpy = (memcpy_fn)dlsym(RTLD_DEFAULT, "cudaMemcpy");
      free_fn cudaFree = (free_fn)dlsym(RTLD_DEFAULT, "cudaFree");
      CHECK(cudaMalloc && cudaMemcpy && cudaFree);
      CHECK(rc == GFB_OK && ctx != NULL);
      CHECK(gfb_nstages(ctx) == 4);
      CHECK(cudaMalloc(&d, sizeof(double) * 2 * n * nkr) == 0);
      h[2 * (3 * nkr + 2)] = 1.0; /* one mode (l = 3, kr = 2): the nonlinear term of a single mode vanishes */
      h[2 * ((n - 3) * nkr + 2)] = 0.0;
      CHECK(cudaMemcpy(d, h, sizeof(double) * 2 * n * nkr, 1) == 0);
      CHECK(gfb_step(ctx, d, 10, NULL, NULL, NULL) == GFB_OK);
      CHECK(cudaMemcpy(h, d, sizeof(double) * 2 * n * nkr, 2) == 0);
      CHECK(gfb_get_clock(ctx, &t, &step) == GFB_OK && step == 10 && fabs(t - 10 * cfg.dt) < 1e-15);
      amp = h[2 * (3 * nkr + 2)];
      { /* RK4 amplification of a linear mode: (1 + z + z^2/2 + z^3/6 + z^4/24)^10, z = L dt */
        const double z = -0.5 * 13.0 * cfg.dt, r = 1 + z + z * z / 2 + z * z * z / 6 + z * z * z * z / 24;
        want = pow(r, 10);
      }
      CHECK(fabs(amp / want - 1) < 1e-12);
      CHECK(gfb_destroy(ctx) == GFB_OK);
      cudaFree(d);
      free(h);
    }
    free(L);
    free(G);
  }
  CHECK(gfb_destroy(NULL) == GFB_OK);
  CHECK(gfb_nstages(NULL) == GFB_ERR_INVALID);
  printf("abi ok (%s)\n", gpu ? "gpu" : "no device");
  return 0;
}
