// gfb_platform.h -- the one place that knows whether we are compiled by nvcc (product) or by g++
// through the test-only emulation layer (tests/emu/cuda_emu.h, -DGFB_EMULATE).
#pragma once

#ifdef GFB_EMULATE
#include "cuda_emu.h"  // found via -I tests/emu (test builds only)
#define GFB_HD
#define __noinline__ __attribute__((noinline))
#define __grid_constant__
#else
#include <cuda_runtime.h>
#define GFB_HD __host__ __device__
#define GFB_DYN_SMEM(type, name)                           \
  extern __shared__ __align__(16) unsigned char gfb_smem_raw[]; \
  type* name = reinterpret_cast<type*>(gfb_smem_raw)
#define GFB_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

#include <cstdint>

namespace gfb {

// 16-byte asynchronous global->shared copy (LDGSTS): no destination registers, so a whole column can be
// in flight while the registers hold another one.  The emulation build copies synchronously.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
#ifdef GFB_EMULATE
  *reinterpret_cast<double2*>(smem_dst) = *reinterpret_cast<const double2*>(gmem_src);
#else
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
#endif
}
// same, L2 only (no L1 allocation): for streams that are read exactly once by this SM
__device__ __forceinline__ void cp_async16_cg(void* smem_dst, const void* gmem_src) {
#ifdef GFB_EMULATE
  *reinterpret_cast<double2*>(smem_dst) = *reinterpret_cast<const double2*>(gmem_src);
#else
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
#endif
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
#ifdef GFB_EMULATE
  *reinterpret_cast<double*>(smem_dst) = *reinterpret_cast<const double*>(gmem_src);
#else
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gmem_src) : "memory");
#endif
}
// Two adjacent complex values (a column pair of one row) = one 32-byte sector.  On sm_100a this is a
// single LDG.E.256 / STG.E.256; `p` must be 32-byte aligned (all ctx-owned arrays have an even leading
// dimension, so a pair starting at an even column always is).
struct cpair {
  double2 a, b;
};
__device__ __forceinline__ cpair ldpair(const double2* p) {
  cpair r;
#ifdef GFB_EMULATE
  r.a = p[0];
  r.b = p[1];
#else
  asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.a.x), "=d"(r.a.y), "=d"(r.b.x), "=d"(r.b.y) : "l"(p));
#endif
  return r;
}
__device__ __forceinline__ void stpair(double2* p, double2 a, double2 b) {
#ifdef GFB_EMULATE
  p[0] = a;
  p[1] = b;
#else
  asm volatile("st.global.v4.f64 [%4], {%0,%1,%2,%3};" ::"d"(a.x), "d"(a.y), "d"(b.x), "d"(b.y), "l"(p) : "memory");
#endif
}
__device__ __forceinline__ double2 ldpair_d(const double* p) {  // two adjacent doubles, 16-byte aligned
#ifdef GFB_EMULATE
  return *reinterpret_cast<const double2*>(p);
#else
  double2 r;
  asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
#endif
}
// one complex value / one double, L2 only (no L1 allocation): the lane-paired column kernels read a 32-byte sector as two
// adjacent lanes' halves, merged by the LSU
__device__ __forceinline__ double2 ldcplx(const double2* p) {
#ifdef GFB_EMULATE
  return *p;
#else
  double2 r;
  asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
#endif
}
__device__ __forceinline__ double lddouble(const double* p) {
#ifdef GFB_EMULATE
  return *p;
#else
  double r;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(r) : "l"(p));
  return r;
#endif
}
// predicated forms: zero when `on` is false, WITHOUT a branch (a branch per row puts every row's loads in their own basic
// block and the rows are then fetched one round trip after the other)
__device__ __forceinline__ double2 ldcplx_if(const double2* p, const bool on) {
#ifdef GFB_EMULATE
  return on ? *p : make_double2(0.0, 0.0);
#else
  double2 r;
  asm volatile(
      "{\n.reg .pred q;\nsetp.ne.s32 q, %3, 0;\nmov.f64 %0, 0d0000000000000000;\nmov.f64 %1, 0d0000000000000000;\n"
      "@q ld.global.cg.v2.f64 {%0,%1}, [%2];\n}\n"
      : "=d"(r.x), "=d"(r.y)
      : "l"(p), "r"((int)on));
  return r;
#endif
}
__device__ __forceinline__ double lddouble_if(const double* p, const bool on) {
#ifdef GFB_EMULATE
  return on ? *p : 0.0;
#else
  double r;
  asm volatile("{\n.reg .pred q;\nsetp.ne.s32 q, %2, 0;\nmov.f64 %0, 0d0000000000000000;\n@q ld.global.cg.f64 %0, [%1];\n}\n"
               : "=d"(r)
               : "l"(p), "r"((int)on));
  return r;
#endif
}
__device__ __forceinline__ void compiler_fence() { asm volatile("" ::: "memory"); }
// Persistent kernels start all their CTAs at once on identical work: without help the whole chip loads, transforms and
// stores in lock step (HBM bursts above the peak, then silence).  A pseudo-random per-CTA start delay in [0, ns) keeps the
// phases of different SMs interleaved for the rest of the launch.
__device__ __forceinline__ void cta_stagger(const int ns) {
#ifndef GFB_EMULATE
  if (ns > 0) __nanosleep((unsigned)(((blockIdx.x * 2654435761u) >> 16) % (unsigned)ns));
#else
  (void)ns;
#endif
}

// ---- Tensor Memory (TMEM) as thread-private parking space ------------------------------------------------
// 256 KB per SM, 512 columns x 128 lanes x 32 bit, reachable only through tcgen05.st / tcgen05.ld.  With the
// 32x32b shape every thread reads and writes its OWN lane (warp w owns lanes 32*(w%4) ..), which is exactly
// what a thread-private stash needs -- and it costs neither shared memory nor LSU bandwidth.  Warps w and
// w+4 share lanes, so they use different column ranges.  All tcgen05 calls are warp-collective
// (.sync.aligned): never call them under divergent control flow.
#ifdef GFB_EMULATE
__device__ __forceinline__ uint32_t tmem_setup(uint32_t*, int) {
  __syncthreads();
  return 0u;
}
__device__ __forceinline__ void tmem_release(uint32_t, int) { __syncthreads(); }
__device__ __forceinline__ uint32_t tmem_thread_base(uint32_t base, int cols_per_warpgroup) {
  const int warp = (int)threadIdx.x / 32;
  return base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * cols_per_warpgroup);
}
__device__ __forceinline__ uint32_t* emu_tmem_cell(uint32_t taddr) {
  const int lane = (int)(taddr >> 16) + ((int)threadIdx.x & 31);
  const int col = (int)(taddr & 0xffffu);
  return emu::tmem_storage() + (size_t)lane * 512 + col;
}
// every emulated TMEM operation is logged (emu::tmem_note): the emulation aborts when the lanes of a warp disagree on
// the sequence of these warp-collective operations, i.e. when one of them sits under divergent control flow
__device__ __forceinline__ void tmem_store(uint32_t taddr, double2 v) {
  emu::tmem_note(1, taddr & 0xffffu);
  std::memcpy(emu_tmem_cell(taddr), &v, 16);
}
__device__ __forceinline__ double2 tmem_load(uint32_t taddr) {
  emu::tmem_note(2, taddr & 0xffffu);
  double2 v;
  std::memcpy(&v, emu_tmem_cell(taddr), 16);
  return v;
}
__device__ __forceinline__ void tmem_wait_st() { emu::tmem_note(3, 0); }
__device__ __forceinline__ void tmem_wait_ld() { emu::tmem_note(4, 0); }
#else
// all threads of the CTA call this; warp 0 allocates `cols` (power of two >= 32) columns
__device__ __forceinline__ uint32_t tmem_setup(uint32_t* smem_word, int cols) {
  if (threadIdx.x < 32) {
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_word);
    if (cols == 512) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(dst) : "memory");
    else if (cols == 256) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(dst) : "memory");
    else asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(dst) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  return *reinterpret_cast<volatile uint32_t*>(smem_word);
}
__device__ __forceinline__ void tmem_release(uint32_t base, int cols) {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    if (cols == 512) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
    else if (cols == 256) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(base) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(base) : "memory");
  }
}
__device__ __forceinline__ uint32_t tmem_thread_base(uint32_t base, int cols_per_warpgroup) {
  const int warp = (int)threadIdx.x / 32;
  return base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * cols_per_warpgroup);
}
__device__ __forceinline__ void tmem_store(uint32_t taddr, double2 v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(__double2loint(v.x)),
               "r"(__double2hiint(v.x)), "r"(__double2loint(v.y)), "r"(__double2hiint(v.y))
               : "memory");
}
__device__ __forceinline__ double2 tmem_load(uint32_t taddr) {  // result valid after tmem_wait_ld()
  uint32_t a, b, c, d;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(taddr) : "memory");
  return make_double2(__hiloint2double((int)b, (int)a), __hiloint2double((int)d, (int)c));
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
#endif

// ---- named barriers: synchronise a SUBSET of the CTA's warps (bar.sync / bar.arrive id, count) ----------------
// The warp-specialised kernels split a CTA into an FFT group and a helper group: the FFT group's internal exchanges use
// bar_sync(id, 256) instead of __syncthreads(), and the two groups hand buffers to each other with the classic
// producer/consumer pair: the producer bar_arrive()s (does not wait), the consumer bar_sync()s; `count` is the total of
// both.  id 0 is __syncthreads' barrier: ids 1..15 are free.
#ifdef GFB_EMULATE
__device__ __forceinline__ void bar_sync(int id, int count) { emu::named_bar(id, count, true); }
__device__ __forceinline__ void bar_arrive(int id, int count) { emu::named_bar(id, count, false); }
__device__ __forceinline__ void tmem_fence_before_sync() {}
__device__ __forceinline__ void tmem_fence_after_sync() {}
#else
__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
// order this thread's TMEM accesses with a following / preceding thread synchronisation (hand-off to other warps)
__device__ __forceinline__ void tmem_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
#endif

// Bulk L2 prefetch (cp.async.bulk.prefetch.L2, issued through the TMA unit): `bytes` contiguous bytes (a multiple of 16,
// 16-byte aligned) start moving from DRAM into L2; costs no register and nothing waits on it.  The helper groups of the
// warp-specialised kernels use it to turn the DRAM latency of the NEXT block of rows into an L2 hit.
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, unsigned bytes) {
#ifndef GFB_EMULATE
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
#else
  (void)p;
  (void)bytes;
#endif
}

__device__ __forceinline__ void cp_async_wait_all() {
#ifndef GFB_EMULATE
  asm volatile("cp.async.wait_all;\n" ::: "memory");
#endif
}

typedef double2 cplx;

__device__ __forceinline__ cplx cmake(double x, double y) { return make_double2(x, y); }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cplx cscale(cplx a, double s) { return make_double2(a.x * s, a.y * s); }
__device__ __forceinline__ cplx cconj(cplx a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ cplx cmul_i(cplx a) { return make_double2(-a.y, a.x); }    //  i * a
__device__ __forceinline__ cplx cmul_mi(cplx a) { return make_double2(a.y, -a.x); }   // -i * a
__device__ __forceinline__ cplx cfma(cplx a, cplx b, cplx c) { return cadd(cmul(a, b), c); }
__device__ __forceinline__ double cabs2(cplx a) { return a.x * a.x + a.y * a.y; }

}  // namespace gfb

// ---- TMA (tensor memory accelerator) + mbarrier ------------------------------------------------------------
// Bulk asynchronous tensor copies between global and shared memory (cp.async.bulk.tensor, SASS UTMALDG / UTMASTG):
// ONE instruction of ONE thread moves a whole box of a <= 5-D tensor; the LSU never sees the individual sectors.
// The row kernel uses them to gather a row out of the 8-row tiled work arrays (a train of 32-byte pieces 256 bytes
// apart) into a dense line buffer and to scatter result rows back.  Completion of loads is signalled on an mbarrier
// in shared memory (transaction bytes), of stores through bulk async-groups.
//
// TmaDesc is the 128-byte CUtensorMap on the device; the emulation build carries the tensor geometry in the clear and
// copies synchronously.
namespace gfb {

constexpr int TMA_MAXRANK = 5;

#ifdef GFB_EMULATE
struct TmaDesc {
  char* base;
  int rank, esize;
  unsigned long long dims[TMA_MAXRANK];     // elements, fastest first
  unsigned long long strides[TMA_MAXRANK];  // bytes; strides[0] = esize
  unsigned box[TMA_MAXRANK];
};
struct Mbar {
  int arrivals_left, arrivals_init;
  long long tx;
  unsigned phase;
};
__device__ __forceinline__ void mbar_init(Mbar* b, int count) {
  b->arrivals_left = b->arrivals_init = count;
  b->tx = 0;
  b->phase = 0;
}
__device__ __forceinline__ void mbar_fence_init() {}
__device__ __forceinline__ void emu_mbar_check(Mbar* b) {
  if (b->arrivals_left == 0 && b->tx == 0) {
    b->arrivals_left = b->arrivals_init;
    b->phase ^= 1u;
  }
}
__device__ __forceinline__ void mbar_expect_tx(Mbar* b, unsigned bytes) {  // arrive + expect_tx
  b->tx += bytes;
  b->arrivals_left -= 1;
  emu_mbar_check(b);
}
__device__ __forceinline__ void mbar_wait(Mbar* b, unsigned parity) {
  while ((b->phase & 1u) == (parity & 1u)) emu::yield_to_sched();
}
// copy one box; dir = 0: global -> smem (out-of-bounds elements read as zero), 1: smem -> global (dropped)
__device__ __forceinline__ void emu_tma_copy(const TmaDesc& d, char* smem, const int* c, int dir) {
  unsigned idx[TMA_MAXRANK] = {0, 0, 0, 0, 0};
  unsigned long long total = 1;
  for (int i = 0; i < d.rank; ++i) total *= d.box[i];
  for (unsigned long long n = 0; n < total; ++n) {
    bool inb = true;
    unsigned long long off = 0;
    for (int i = 0; i < d.rank; ++i) {
      const long long g = (long long)c[i] + idx[i];
      if (g < 0 || (unsigned long long)g >= d.dims[i]) inb = false;
      off += (unsigned long long)g * d.strides[i];
    }
    char* s = smem + n * d.esize;
    if (dir == 0) {
      if (inb) std::memcpy(s, d.base + off, d.esize);
      else std::memset(s, 0, d.esize);
    } else if (inb) {
      std::memcpy(d.base + off, s, d.esize);
    }
    for (int i = 0; i < d.rank; ++i) {
      if (++idx[i] < d.box[i]) break;
      idx[i] = 0;
    }
  }
}
__device__ __forceinline__ void tma_load_5d(void* dst, const TmaDesc* d, Mbar* bar, int c0, int c1, int c2, int c3, int c4) {
  const int c[5] = {c0, c1, c2, c3, c4};
  emu_tma_copy(*d, (char*)dst, c, 0);
  unsigned long long bytes = d->esize;
  for (int i = 0; i < d->rank; ++i) bytes *= d->box[i];
  bar->tx -= (long long)bytes;
  emu_mbar_check(bar);
}
__device__ __forceinline__ void tma_store_5d(const TmaDesc* d, const void* src, int c0, int c1, int c2, int c3, int c4) {
  const int c[5] = {c0, c1, c2, c3, c4};
  emu_tma_copy(*d, (char*)const_cast<void*>(src), c, 1);
}
__device__ __forceinline__ void tma_commit() {}
template <int N>
__device__ __forceinline__ void tma_wait_read() {}
template <int N>
__device__ __forceinline__ void tma_wait_all() {}
__device__ __forceinline__ void fence_proxy_async() {}
#else
struct alignas(64) TmaDesc {
  unsigned char opaque[128];
};
typedef unsigned long long Mbar;
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(Mbar* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(Mbar* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(Mbar* b, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_addr(b)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_5d(void* dst, const TmaDesc* d, Mbar* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
          smem_addr(dst)),
      "l"(d), "r"(smem_addr(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_store_5d(const TmaDesc* d, const void* src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(d),
               "r"(smem_addr(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_read() {  // the sources of all but the N youngest store groups have been read
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_wait_all() {  // all but the N youngest store groups are complete (globally visible)
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// generic-proxy writes to shared memory -> visible to the async proxy (TMA store) that reads them next
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

}  // namespace gfb

// ---- wide TMEM transfers: a thread's 16 (or 4) complex values in ONE tcgen05 instruction ---------------------------
// tcgen05.ld / .st are issued per warp and carry a fixed cost per instruction; the .x64 shape moves 64 columns (16
// complex doubles) of the thread's lane at once, the .x16 shape 4 complex values.
namespace gfb {
#ifdef GFB_EMULATE
__device__ __forceinline__ void tmem_store16(uint32_t taddr, const cplx (&v)[16]) {
  emu::tmem_note(5, taddr & 0xffffu);
  for (int r = 0; r < 16; ++r) std::memcpy(emu_tmem_cell(taddr + 4 * r), &v[r], 16);
}
__device__ __forceinline__ void tmem_load16(uint32_t taddr, cplx (&v)[16]) {
  emu::tmem_note(6, taddr & 0xffffu);
  for (int r = 0; r < 16; ++r) std::memcpy(&v[r], emu_tmem_cell(taddr + 4 * r), 16);
}
__device__ __forceinline__ void tmem_load4(uint32_t taddr, cplx (&v)[4]) {
  emu::tmem_note(7, taddr & 0xffffu);
  for (int r = 0; r < 4; ++r) std::memcpy(&v[r], emu_tmem_cell(taddr + 4 * r), 16);
}
__device__ __forceinline__ void tmem_store4(uint32_t taddr, const cplx (&v)[4]) {
  emu::tmem_note(8, taddr & 0xffffu);
  for (int r = 0; r < 4; ++r) std::memcpy(emu_tmem_cell(taddr + 4 * r), &v[r], 16);
}
#else
__device__ __forceinline__ void tmem_store16(uint32_t taddr, const cplx (&v)[16]) {
  uint32_t w[64];
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    w[4 * r + 0] = (uint32_t)__double2loint(v[r].x);
    w[4 * r + 1] = (uint32_t)__double2hiint(v[r].x);
    w[4 * r + 2] = (uint32_t)__double2loint(v[r].y);
    w[4 * r + 3] = (uint32_t)__double2hiint(v[r].y);
  }
  asm volatile("tcgen05.st.sync.aligned.32x32b.x64.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63, %64};" ::"r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]), "r"(w[9]), "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]), "r"(w[16]), "r"(w[17]), "r"(w[18]), "r"(w[19]), "r"(w[20]), "r"(w[21]), "r"(w[22]), "r"(w[23]), "r"(w[24]), "r"(w[25]), "r"(w[26]), "r"(w[27]), "r"(w[28]), "r"(w[29]), "r"(w[30]), "r"(w[31]), "r"(w[32]), "r"(w[33]), "r"(w[34]), "r"(w[35]), "r"(w[36]), "r"(w[37]), "r"(w[38]), "r"(w[39]), "r"(w[40]), "r"(w[41]), "r"(w[42]), "r"(w[43]), "r"(w[44]), "r"(w[45]), "r"(w[46]), "r"(w[47]), "r"(w[48]), "r"(w[49]), "r"(w[50]), "r"(w[51]), "r"(w[52]), "r"(w[53]), "r"(w[54]), "r"(w[55]), "r"(w[56]), "r"(w[57]), "r"(w[58]), "r"(w[59]), "r"(w[60]), "r"(w[61]), "r"(w[62]), "r"(w[63]) : "memory");
}
__device__ __forceinline__ void tmem_load16(uint32_t taddr, cplx (&v)[16]) {  // valid after tmem_wait_ld()
  uint32_t w[64];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]), "=r"(w[16]), "=r"(w[17]), "=r"(w[18]), "=r"(w[19]), "=r"(w[20]), "=r"(w[21]), "=r"(w[22]), "=r"(w[23]), "=r"(w[24]), "=r"(w[25]), "=r"(w[26]), "=r"(w[27]), "=r"(w[28]), "=r"(w[29]), "=r"(w[30]), "=r"(w[31]), "=r"(w[32]), "=r"(w[33]), "=r"(w[34]), "=r"(w[35]), "=r"(w[36]), "=r"(w[37]), "=r"(w[38]), "=r"(w[39]), "=r"(w[40]), "=r"(w[41]), "=r"(w[42]), "=r"(w[43]), "=r"(w[44]), "=r"(w[45]), "=r"(w[46]), "=r"(w[47]), "=r"(w[48]), "=r"(w[49]), "=r"(w[50]), "=r"(w[51]), "=r"(w[52]), "=r"(w[53]), "=r"(w[54]), "=r"(w[55]), "=r"(w[56]), "=r"(w[57]), "=r"(w[58]), "=r"(w[59]), "=r"(w[60]), "=r"(w[61]), "=r"(w[62]), "=r"(w[63]) : "r"(taddr) : "memory");
#pragma unroll
  for (int r = 0; r < 16; ++r)
    v[r] = make_double2(__hiloint2double((int)w[4 * r + 1], (int)w[4 * r + 0]), __hiloint2double((int)w[4 * r + 3], (int)w[4 * r + 2]));
}
__device__ __forceinline__ void tmem_store4(uint32_t taddr, const cplx (&v)[4]) {
  uint32_t w[16];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    w[4 * r + 0] = (uint32_t)__double2loint(v[r].x);
    w[4 * r + 1] = (uint32_t)__double2hiint(v[r].x);
    w[4 * r + 2] = (uint32_t)__double2loint(v[r].y);
    w[4 * r + 3] = (uint32_t)__double2hiint(v[r].y);
  }
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]), "r"(w[9]), "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]) : "memory");
}
__device__ __forceinline__ void tmem_load4(uint32_t taddr, cplx (&v)[4]) {  // valid after tmem_wait_ld()
  uint32_t w[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]) : "r"(taddr) : "memory");
#pragma unroll
  for (int r = 0; r < 4; ++r)
    v[r] = make_double2(__hiloint2double((int)w[4 * r + 1], (int)w[4 * r + 0]), __hiloint2double((int)w[4 * r + 3], (int)w[4 * r + 2]));
}
#endif
}  // namespace gfb
