// TEST INFRASTRUCTURE — not product code.
//
// A minimal CUDA execution-model emulator so that the *same* kernel sources under
// vaesne-dev_b200/csrc (the portable, non-tcgen05 ones) can be compiled by g++ and
// exercised on a CPU-only box.  One OS worker per block-in-flight; the threads of a
// block are ucontext fibers with real block barriers, warp-synchronous shuffles and
// atomics, so data races on shared memory or missing barriers behave as on a GPU
// scheduler that runs one thread at a time.  It is never loaded by the product: the
// product's loader only opens the nvcc-built library and refuses CPU tensors.
#pragma once
#include <ucontext.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>
#include <atomic>
#include <algorithm>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define __align__(n) __attribute__((aligned(n)))

struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct uint3_ { unsigned x, y, z; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct int4 { int x, y, z, w; };
static inline float4 make_float4(float a, float b, float c, float d) { return {a, b, c, d}; }
static inline float2 make_float2(float a, float b) { return {a, b}; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return {a, b}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
static const int cudaSuccess = 0;
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }

namespace emu {

struct Fiber {
  ucontext_t ctx;
  char* stack = nullptr;
  int state = 0;        // 0 runnable, 1 waiting block barrier, 2 waiting warp barrier, 3 done
  unsigned gen = 0;     // generation waited for
};

struct Block {
  uint3_ tid[1024];
  std::vector<Fiber> fibers;
  ucontext_t sched;
  int cur = -1;
  int nthreads = 0;
  int live = 0;
  int bar_count = 0; unsigned bar_gen = 0;
  int warp_count[32]; unsigned warp_gen[32]; int warp_live[32];
  uint64_t shfl[32][2][32];
  unsigned shfl_n[1024];
  char* dyn = nullptr;
  const std::function<void()>* body = nullptr;
};

extern thread_local Block* g_blk;
extern thread_local uint3_ g_blockIdx;
extern thread_local dim3 g_blockDim, g_gridDim;

static const size_t kStack = 256 * 1024;

inline void yield_to_sched() { Block* b = g_blk; swapcontext(&b->fibers[b->cur].ctx, &b->sched); }

inline void fiber_entry() {
  Block* b = g_blk;
  (*b->body)();
  Fiber& f = b->fibers[b->cur];
  f.state = 3;
  b->live--;
  b->warp_live[b->cur / 32]--;
  // a finished thread must not block others waiting at a barrier
  if (b->live > 0 && b->bar_count == b->live) { b->bar_count = 0; b->bar_gen++; }
  int w = b->cur / 32;
  if (b->warp_live[w] > 0 && b->warp_count[w] == b->warp_live[w]) { b->warp_count[w] = 0; b->warp_gen[w]++; }
  swapcontext(&f.ctx, &b->sched);
}

inline void run_block(Block* b, const std::function<void()>& body, dim3 bd) {
  g_blk = b;
  b->body = &body;
  b->nthreads = bd.x * bd.y * bd.z;
  if ((int)b->fibers.size() < b->nthreads) {
    size_t old = b->fibers.size();
    b->fibers.resize(b->nthreads);
    for (size_t i = old; i < b->fibers.size(); ++i) b->fibers[i].stack = (char*)malloc(kStack);
  }
  b->live = b->nthreads; b->bar_count = 0; b->bar_gen = 0;
  for (int w = 0; w < 32; ++w) { b->warp_count[w] = 0; b->warp_gen[w] = 0; b->warp_live[w] = 0; }
  for (int t = 0; t < b->nthreads; ++t) {
    b->tid[t] = {t % bd.x, (t / bd.x) % bd.y, t / (bd.x * bd.y)};
    b->shfl_n[t] = 0;
    b->warp_live[t / 32]++;
    Fiber& f = b->fibers[t];
    f.state = 0; f.gen = 0;
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack; f.ctx.uc_stack.ss_size = kStack; f.ctx.uc_link = nullptr;
    makecontext(&f.ctx, (void (*)())fiber_entry, 0);
  }
  while (b->live > 0) {
    bool progressed = false;
    for (int t = 0; t < b->nthreads; ++t) {
      Fiber& f = b->fibers[t];
      if (f.state == 3) continue;
      if (f.state == 1 && b->bar_gen == f.gen) continue;
      if (f.state == 2 && b->warp_gen[t / 32] == f.gen) continue;
      f.state = 0;
      b->cur = t;
      swapcontext(&b->sched, &f.ctx);
      progressed = true;
    }
    if (!progressed) { fprintf(stderr, "emu: deadlock (divergent barrier?)\n"); abort(); }
  }
}

inline void sync_block() {
  Block* b = g_blk; Fiber& f = b->fibers[b->cur];
  f.state = 1; f.gen = b->bar_gen;
  if (++b->bar_count == b->live) { b->bar_count = 0; b->bar_gen++; }
  yield_to_sched();
}
inline void sync_warp() {
  Block* b = g_blk; Fiber& f = b->fibers[b->cur]; int w = b->cur / 32;
  f.state = 2; f.gen = b->warp_gen[w];
  if (++b->warp_count[w] == b->warp_live[w]) { b->warp_count[w] = 0; b->warp_gen[w]++; }
  yield_to_sched();
}

template <class T> inline T shfl_idx(T v, int src) {
  static_assert(sizeof(T) <= 8, "shfl payload");
  Block* b = g_blk; int t = b->cur, w = t / 32, l = t % 32;
  unsigned par = (b->shfl_n[t]++) & 1u;
  uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
  b->shfl[w][par][l] = raw;
  sync_warp();
  uint64_t got = b->shfl[w][par][src & 31];
  T out; memcpy(&out, &got, sizeof(T));
  return out;
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
char* dyn_smem();
}  // namespace emu

#define threadIdx (emu::g_blk->tid[emu::g_blk->cur])
#define blockIdx (emu::g_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)

static inline void __syncthreads() { emu::sync_block(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::sync_warp(); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu::shfl_idx(v, (emu::g_blk->cur % 32) ^ m); }
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return emu::shfl_idx(v, src); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d) { int l = emu::g_blk->cur % 32; return emu::shfl_idx(v, l + d < 32 ? l + d : l); }

static inline float atomicAdd(float* p, float v) {
  uint32_t* ip = (uint32_t*)p; uint32_t old = __atomic_load_n(ip, __ATOMIC_RELAXED), nw; float f;
  do { memcpy(&f, &old, 4); f += v; memcpy(&nw, &f, 4); } while (!__atomic_compare_exchange_n(ip, &old, nw, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
  memcpy(&f, &old, 4); return f;
}
static inline double atomicAdd(double* p, double v) {
  uint64_t* ip = (uint64_t*)p; uint64_t old = __atomic_load_n(ip, __ATOMIC_RELAXED), nw; double f;
  do { memcpy(&f, &old, 8); f += v; memcpy(&nw, &f, 8); } while (!__atomic_compare_exchange_n(ip, &old, nw, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
  memcpy(&f, &old, 8); return f;
}
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

template <class T> static inline T __ldg(const T* p) { return *p; }
#define __expf(x) expf(x)
#define __logf(x) logf(x)
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline void sincosf_(float x, float* s, float* c) { *s = sinf(x); *c = cosf(x); }
