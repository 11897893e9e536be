// cuda_emu.h -- TEST INFRASTRUCTURE ONLY.
//
// A tiny CUDA-on-CPU emulation layer so that the *same* kernel sources under
// geophysicalflows.jl_b200/csrc can be compiled with g++ and executed in this GPU-less container by
// the `-m "not gpu"` tests (tests/test_emu_*.py).  One CUDA thread = one ucontext fiber; one CTA runs
// on one OS thread (CTAs are spread over OS threads); __syncthreads()/__shfl_*_sync are real barriers
// between fibers, so races and wrong barrier placement show up here.  It is never built into, loaded
// by, or shipped with the product library (`libgfb.so`, nvcc only): the product has no CPU path.
#pragma once
#include <ucontext.h>

#include <atomic>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#define GFB_EMULATE 1

struct double2 {
  double x, y;
};
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
struct uint3 {
  unsigned x, y, z;
};
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __constant__ static
#define __shared__ static thread_local

typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };

static inline cudaError_t cudaMalloc(void** p, size_t n) {
  *p = std::malloc(n ? n : 1);
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
static inline cudaError_t cudaFree(void* p) {
  std::free(p);
  return cudaSuccess;
}
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) {
  std::memmove(d, s, n);
  return cudaSuccess;
}
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) {
  std::memmove(d, s, n);
  return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) {
  std::memset(d, v, n);
  return cudaSuccess;
}
static inline cudaError_t cudaMemset(void* d, int v, size_t n) {
  std::memset(d, v, n);
  return cudaSuccess;
}
enum { cudaHostAllocMapped = 2 };
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) {
  *p = std::malloc(n ? n : 1);
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
static inline cudaError_t cudaHostGetDevicePointer(void** d, void* h, unsigned) {
  *d = h;
  return cudaSuccess;
}
static inline cudaError_t cudaFreeHost(void* p) {
  std::free(p);
  return cudaSuccess;
}
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t e) { return e == 0 ? "no error" : "emulated error"; }
template <class F>
static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
typedef struct emuEvent* cudaEvent_t;
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) {
  *e = nullptr;
  return cudaSuccess;
}
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) {
  *ms = 0.f;
  return cudaSuccess;
}
static inline cudaError_t cudaGetDevice(int* d) {
  *d = 0;
  return cudaSuccess;
}
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) {
  *v = 148;
  return cudaSuccess;
}
struct cudaDeviceProp {
  int multiProcessorCount;
  size_t sharedMemPerBlockOptin;
};
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
  p->multiProcessorCount = 148;
  p->sharedMemPerBlockOptin = 227 * 1024;
  return cudaSuccess;
}

namespace emu {

struct Barrier {
  int count = 0, expected = 0;
  unsigned generation = 0;
};

struct Fiber {
  ucontext_t ctx;
  bool done = false;
  unsigned tid = 0;
};

struct BlockState {
  std::vector<Fiber> fibers;
  std::vector<char*> stacks;
  ucontext_t sched;
  Barrier block_bar;
  std::vector<Barrier> warp_bar;
  Barrier named_bar[16];  // bar.sync / bar.arrive id, count
  std::vector<uint64_t> shfl_slot;
  int current = -1;
  const std::function<void()>* body = nullptr;
  char* smem = nullptr;
  size_t smem_cap = 0;
  uint32_t* tmem = nullptr;  // emulated tensor memory: 128 lanes x 512 columns
  // tcgen05.* are warp-collective (.sync.aligned): every lane of a warp must execute the same sequence of them.  Each
  // fiber keeps a running hash of its TMEM operations; the hashes of a warp are compared whenever the CTA barrier
  // completes and when the CTA ends, so a TMEM access under divergent control flow aborts the emulated kernel.
  std::vector<uint64_t> tm_hash;
  ~BlockState() {
    for (char* s : stacks) std::free(s);
    std::free(smem);
    std::free(tmem);
  }
};

extern thread_local BlockState* tl_block;
extern thread_local uint3 tl_threadIdx, tl_blockIdx;
extern thread_local dim3 tl_blockDim, tl_gridDim;

static const size_t kStack = 256 * 1024;

inline uint32_t* tmem_storage();

inline void yield_to_sched() {
  BlockState* b = tl_block;
  Fiber& f = b->fibers[b->current];
  swapcontext(&f.ctx, &b->sched);
}

inline void barrier_wait(Barrier& bar) {
  unsigned gen = bar.generation;
  if (++bar.count >= bar.expected) {
    bar.count = 0;
    bar.generation++;
  } else {
    while (bar.generation == gen) yield_to_sched();
  }
}

inline void barrier_drop(Barrier& bar) {  // a thread exits: it no longer participates
  bar.expected--;
  if (bar.expected > 0 && bar.count >= bar.expected) {
    bar.count = 0;
    bar.generation++;
  }
}

inline void tmem_note(unsigned kind, unsigned column) {
  BlockState* b = tl_block;
  uint64_t& h = b->tm_hash[b->current];
  h = (h ^ (((uint64_t)kind << 32) | column)) * 0x100000001b3ull + 0x9e3779b97f4a7c15ull;
}
inline void tmem_check_uniform(const char* where) {
  BlockState* b = tl_block;
  const size_t nt = b->tm_hash.size();
  for (size_t w = 0; w * 32 < nt; ++w) {
    bool have = false;
    uint64_t ref = 0;
    for (size_t t = w * 32; t < nt && t < (w + 1) * 32; ++t) {
      if (b->fibers[t].done) continue;
      if (!have) { ref = b->tm_hash[t]; have = true; }
      else if (b->tm_hash[t] != ref) {
        std::fprintf(stderr, "cuda_emu: warp %zu of CTA (%u,%u) executed tcgen05 (TMEM) operations under divergent control flow "
                             "(detected at %s): lanes disagree on the sequence of warp-collective operations\n",
                     w, tl_blockIdx.x, tl_blockIdx.y, where);
        std::abort();
      }
    }
  }
}
inline void sync_block() {
  Barrier& bar = tl_block->block_bar;
  if (bar.count + 1 >= bar.expected) tmem_check_uniform("__syncthreads");  // last arrival: everybody else is waiting here
  barrier_wait(bar);
}
// bar.sync (wait = true) / bar.arrive (wait = false) on named barrier `id` expecting `count` threads in total
inline void named_bar(int id, int count, bool wait) {
  Barrier& bar = tl_block->named_bar[id & 15];
  const unsigned gen = bar.generation;
  if (++bar.count >= count) {
    bar.count = 0;
    bar.generation++;
  } else if (wait) {
    while (bar.generation == gen) yield_to_sched();
  }
}
inline void sync_warp() { barrier_wait(tl_block->warp_bar[tl_block->current / 32]); }

template <class T>
inline T shfl_idx(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shfl payload");
  BlockState* b = tl_block;
  int tid = b->current;
  uint64_t raw = 0;
  std::memcpy(&raw, &v, sizeof(T));
  b->shfl_slot[tid] = raw;
  sync_warp();
  int base = (tid / 32) * 32;
  int src = base + (src_lane & 31);
  if (src >= (int)b->fibers.size()) src = tid;
  uint64_t got = b->shfl_slot[src];
  sync_warp();
  T out;
  std::memcpy(&out, &got, sizeof(T));
  return out;
}

inline uint32_t* tmem_storage() {
  BlockState* b = tl_block;
  if (!b->tmem) b->tmem = (uint32_t*)std::calloc((size_t)128 * 512, sizeof(uint32_t));
  return b->tmem;
}

void fiber_entry();

inline void run_block(BlockState& b, dim3 grid, dim3 block, uint3 bidx, size_t smem_bytes, const std::function<void()>& body) {
  unsigned nt = block.x * block.y * block.z;
  tl_block = &b;
  tl_blockIdx = bidx;
  tl_blockDim = block;
  tl_gridDim = grid;
  if (b.smem_cap < smem_bytes + 64) {
    std::free(b.smem);
    b.smem = (char*)std::aligned_alloc(128, ((smem_bytes + 64 + 127) / 128) * 128);
    b.smem_cap = smem_bytes + 64;
  }
  // poison shared memory so that reads of never-written smem are visible as NaNs
  std::memset(b.smem, 0xFF, smem_bytes);
  b.body = &body;
  if (b.fibers.size() < nt) b.fibers.resize(nt);
  while (b.stacks.size() < nt) b.stacks.push_back((char*)std::malloc(kStack));
  b.block_bar = Barrier();
  for (auto& nb : b.named_bar) nb = Barrier();
  b.block_bar.expected = (int)nt;
  b.warp_bar.assign((nt + 31) / 32, Barrier());
  for (unsigned w = 0; w < b.warp_bar.size(); ++w) b.warp_bar[w].expected = (int)std::min(32u, nt - 32 * w);
  b.shfl_slot.assign(nt, 0);
  b.tm_hash.assign(nt, 0);
  for (unsigned t = 0; t < nt; ++t) {
    Fiber& f = b.fibers[t];
    f.done = false;
    f.tid = t;
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = b.stacks[t];
    f.ctx.uc_stack.ss_size = kStack;
    f.ctx.uc_link = &b.sched;
    makecontext(&f.ctx, (void (*)())fiber_entry, 0);
  }
  unsigned alive = nt;
  while (alive) {
    for (unsigned t = 0; t < nt; ++t) {
      Fiber& f = b.fibers[t];
      if (f.done) continue;
      b.current = (int)t;
      tl_threadIdx = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
      swapcontext(&b.sched, &f.ctx);
      if (f.done) --alive;
    }
  }
  for (unsigned t = 0; t < nt; ++t) b.fibers[t].done = false;  // compare the final TMEM histories of all lanes
  tmem_check_uniform("kernel exit");
  b.current = -1;
}

inline int num_workers() {
  const char* e = std::getenv("GFB_EMU_THREADS");
  int n = e ? std::atoi(e) : (int)std::thread::hardware_concurrency();
  return n < 1 ? 1 : n;
}

inline void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body) {
  size_t nblocks = (size_t)grid.x * grid.y * grid.z;
  if (nblocks == 0) return;
  int nw = (int)std::min<size_t>(num_workers(), nblocks);
  std::atomic<size_t> next{0};
  auto worker = [&]() {
    static thread_local BlockState state;
    for (;;) {
      size_t i = next.fetch_add(1);
      if (i >= nblocks) break;
      uint3 bidx{(unsigned)(i % grid.x), (unsigned)((i / grid.x) % grid.y), (unsigned)(i / ((size_t)grid.x * grid.y))};
      run_block(state, grid, block, bidx, smem_bytes, body);
    }
  };
  if (nw == 1) {
    worker();
  } else {
    std::vector<std::thread> th;
    for (int w = 0; w < nw; ++w) th.emplace_back(worker);
    for (auto& t : th) t.join();
  }
}

}  // namespace emu

#define threadIdx (emu::tl_threadIdx)
#define blockIdx (emu::tl_blockIdx)
#define blockDim (emu::tl_blockDim)
#define gridDim (emu::tl_gridDim)

static inline void __syncthreads() { emu::sync_block(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::sync_warp(); }
template <class T>
static inline T __ldg(const T* p) { return *p; }
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int lanemask) { return emu::shfl_idx(v, (emu::tl_block->current & 31) ^ lanemask); }
template <class T>
static inline T __shfl_down_sync(unsigned, T v, int delta) {
  int lane = emu::tl_block->current & 31;
  return emu::shfl_idx(v, lane + delta > 31 ? lane : lane + delta);
}
template <class T>
static inline T __shfl_sync(unsigned, T v, int src) { return emu::shfl_idx(v, src); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }

// dynamic shared memory of the running CTA
#define GFB_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::tl_block->smem)

// kernel launch: same call shape as the CUDA build
#define GFB_LAUNCH(kernel, grid, block, smem, stream, ...) \
  emu::launch((grid), (block), (smem), [=]() { kernel(__VA_ARGS__); })

#ifdef GFB_EMU_IMPLEMENTATION
namespace emu {
thread_local BlockState* tl_block = nullptr;
thread_local uint3 tl_threadIdx, tl_blockIdx;
thread_local dim3 tl_blockDim, tl_gridDim;
void fiber_entry() {
  BlockState* b = tl_block;
  int tid = b->current;
  (*b->body)();
  b = tl_block;
  b->fibers[tid].done = true;
  barrier_drop(b->block_bar);
  barrier_drop(b->warp_bar[tid / 32]);
  // returning switches to uc_link (the scheduler)
}
}  // namespace emu
#endif
