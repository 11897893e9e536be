// tlb_probe.cu -- does the column pass at 16384^2 pay for address translation?  (design experiment, not part of the library)
// A column-pair CTA of the 8-row tiled layout touches one 256-byte run per 8 rows at a stride of 8*ld*16 bytes: at
// ld = 2050 the 4096 rows of a column span 134 MB (64 pages of 2 MB), at ld = 8194 the 16384 rows span 2.1 GB (1024
// pages).  Same kernel, same bytes per thread, three layouts:
//   tiled    [row/8][pair][row%8][2]                      (the library's layout)
//   blocked  [pair/PB][row/8][pair%PB][row%8][2]          (a column pair's rows span rows*PB*32 bytes)
// and the row pattern (one CTA walks along rows) on both, because the blocked layout moves the page span to the rows.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
typedef double4 sec;  // one 32-byte sector = (row, column pair)
__device__ __forceinline__ sec ldsec(const sec* p) {
  sec r;
  asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stsec(sec* p, const sec v) {
  asm volatile("st.global.v4.f64 [%4], {%0,%1,%2,%3};" ::"d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w), "l"(p) : "memory");
}

struct Lay {
  long long rows, pairs, pb;  // pb = 0: tiled
  __host__ __device__ long long off(long long row, long long pair) const {
    if (pb == 0) return ((row >> 3) * pairs + pair) * 8 + (row & 7);
    const long long b = pair / pb, pi = pair % pb;
    return ((b * (rows >> 3) + (row >> 3)) * pb + pi) * 8 + (row & 7);
  }
};

// column pattern: CTA = one column pair, thread t reads rows t + 256 i (16 in flight), adds, writes to `out` (same layout)
__global__ void __launch_bounds__(256, 2) k_col(const sec* __restrict__ in, sec* __restrict__ out, const Lay L, const int write) {
  const long long pair = blockIdx.x;
  for (long long r0 = 0; r0 < L.rows; r0 += 4096) {
    sec v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = ldsec(&in[L.off(r0 + threadIdx.x + 256 * i, pair)]);
    if (write) {
#pragma unroll
      for (int i = 0; i < 16; ++i) { v[i].x += 1.0; stsec(&out[L.off(r0 + threadIdx.x + 256 * i, pair)], v[i]); }
    } else {
      double s = 0;
#pragma unroll
      for (int i = 0; i < 16; ++i) s += v[i].x;
      if (s == 1.2345) out[0].x = s;
    }
  }
}
// row pattern: CTA = one row pair (rows 2b, 2b+1), thread t reads pairs t + 256 i of both rows
__global__ void __launch_bounds__(256, 2) k_row(const sec* __restrict__ in, sec* __restrict__ out, const Lay L, const int write) {
  const long long row = 2LL * blockIdx.x;
  for (long long p0 = 0; p0 < L.pairs; p0 += 2048) {
    sec v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const long long p = p0 + threadIdx.x + 256 * (i >> 1);
      v[i] = p < L.pairs ? ldsec(&in[L.off(row + (i & 1), p)]) : make_double4(0, 0, 0, 0);
    }
    if (write) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const long long p = p0 + threadIdx.x + 256 * (i >> 1);
        v[i].x += 1.0;
        if (p < L.pairs) stsec(&out[L.off(row + (i & 1), p)], v[i]);
      }
    } else {
      double s = 0;
#pragma unroll
      for (int i = 0; i < 16; ++i) s += v[i].x;
      if (s == 1.2345) out[0].x = s;
    }
  }
}

template <class F>
static float timeit(F f, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps;
}

int main() {
  struct Case { const char* name; long long rows, pairs, pb; };
  const Case cases[] = {
      {"4096 rows x 1024 pairs, tiled", 4096, 1024, 0},
      {"4096 rows x 4096 pairs, tiled", 4096, 4096, 0},
      {"16384 rows x 4096 pairs, tiled", 16384, 4096, 0},
      {"16384 rows x 4096 pairs, blocked PB=16", 16384, 4096, 16},
      {"16384 rows x 4096 pairs, blocked PB=64", 16384, 4096, 64},
      {"16384 rows x 4096 pairs, blocked PB=256", 16384, 4096, 256},
  };
  sec *a, *b;
  const size_t maxn = (size_t)16384 * 4096;
  cudaMalloc(&a, maxn * sizeof(sec));
  cudaMalloc(&b, maxn * sizeof(sec));
  cudaMemset(a, 0, maxn * sizeof(sec));
  cudaMemset(b, 0, maxn * sizeof(sec));
  for (const Case& c : cases) {
    const Lay L{c.rows, c.pairs, c.pb};
    const double bytes = (double)c.rows * c.pairs * 32;
    const int reps = c.rows == 4096 ? 20 : 3;
    for (int write = 0; write < 2; ++write) {
      const float tc = timeit([&] { k_col<<<(unsigned)c.pairs, 256>>>(a, b, L, write); }, reps);
      const float tr = timeit([&] { k_row<<<(unsigned)(c.rows / 2), 256>>>(a, b, L, write); }, reps);
      printf("%-44s %-10s  column pass %8.3f ms %7.1f GB/s   row pass %8.3f ms %7.1f GB/s\n", c.name, write ? "read+write" : "read",
             tc, bytes * (1 + write) / tc * 1e-6, tr, bytes * (1 + write) / tr * 1e-6);
    }
  }
  printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
