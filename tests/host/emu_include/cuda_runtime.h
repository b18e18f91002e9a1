// Test infrastructure, never shipped: a stand-in for <cuda_runtime.h> that lets g++ compile the DEVICE functions of
// kimimaro_b200/csrc/*.cu for the CPU.  One CUDA block is emulated at a time with one FIBER per CUDA thread, all on the
// calling OS thread and switched cooperatively at the synchronisation points (simt_impl.h: a 20-instruction context
// switch, ~100x faster than OS threads meeting in pthread barriers, and deterministic):
//   __syncthreads                      -> a barrier over the block's fibers
//   __ballot_sync / __shfl*_sync       -> an exchange buffer and a barrier per warp (all 32 lanes must take part, which
//                                         is also what the full masks in the sources promise)
//   atomics                            -> GCC __atomic builtins (atomicMin by compare-and-swap)
//   __ldg / __ldcg                     -> relaxed atomic loads (so that the compiler cannot cache them across barriers)
//   __fmul_rn & co                     -> plain operators (compile with -ffp-contract=off)
//   __shared__                         -> static (shared by all threads because only one block runs at a time)
// What this checks is the LOGIC of a device function -- list handling, reductions, the order of barriers, claim words --
// against the oracle; what it cannot check is the GPU's memory model (cache coherence of plain loads, divergence).
#pragma once
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x = 1, y = 1, z = 1; };
typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }

#define __device__
#define __global__
#define __host__
#define __constant__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __shared__ static
#define __launch_bounds__(...)

namespace simt {
constexpr int kMaxThreads = 1024;
struct Fiber;
struct Barrier {
  int expected = 0, arrived = 0;
  Fiber* head = nullptr;   // fibers parked here until the last one arrives
  Fiber* tail = nullptr;
};
struct Warp {
  Barrier bar;
  unsigned long long v[32];
};
struct Block {
  Barrier bar;
  Warp warps[kMaxThreads / 32];
  int n_threads = 0;
};
extern Block g_block;
void barrier_wait(Barrier& b);
}  // namespace simt

extern thread_local uint3 threadIdx;
extern thread_local uint3 blockIdx;
extern uint3 blockDim, gridDim;

static inline void __syncthreads() { simt::barrier_wait(simt::g_block.bar); }
static inline void __syncwarp(unsigned = 0xffffffffu) { simt::barrier_wait(simt::g_block.warps[threadIdx.x >> 5].bar); }
static inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

namespace simt {
template <typename T> inline unsigned long long to_bits(T x) { unsigned long long b = 0; memcpy(&b, &x, sizeof(T)); return b; }
template <typename T> inline T from_bits(unsigned long long b) { T x; memcpy(&x, &b, sizeof(T)); return x; }
// every lane publishes a value, then reads the one it wants; two barriers so that the buffer can be reused at once
template <typename T, typename F> inline T exchange(T mine, F pick) {
  Warp& w = g_block.warps[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  w.v[lane] = to_bits(mine);
  barrier_wait(w.bar);
  const T r = from_bits<T>(w.v[pick(lane) & 31]);
  barrier_wait(w.bar);
  return r;
}
}  // namespace simt

static inline unsigned __ballot_sync(unsigned, int pred) {
  simt::Warp& w = simt::g_block.warps[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  w.v[lane] = pred ? 1ull : 0ull;
  simt::barrier_wait(w.bar);
  unsigned m = 0;
  for (int l = 0; l < 32; l++) m |= (unsigned)(w.v[l] & 1ull) << l;
  simt::barrier_wait(w.bar);
  return m;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
// all 32 lanes take part (full mask): the lanes holding the same value
static inline unsigned __match_any_sync(unsigned, unsigned value) {
  simt::Warp& w = simt::g_block.warps[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  w.v[lane] = value;
  simt::barrier_wait(w.bar);
  unsigned m = 0;
  for (int l = 0; l < 32; l++) m |= (unsigned)(w.v[l] == (unsigned long long)value) << l;
  simt::barrier_wait(w.bar);
  return m;
}
// redux.sync over the lanes of `mask`.  On the GPU only those lanes execute it; here they are the lanes of one
// __match_any_sync group calling from inside a branch, so the exchange uses a barrier sized to the group: every member
// publishes, the members meet, every member folds the group's values.
namespace simt { unsigned reduce_group(unsigned mask, unsigned v, int op); }
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) { return simt::reduce_group(mask, v, 0); }
static inline unsigned __reduce_min_sync(unsigned mask, unsigned v) { return simt::reduce_group(mask, v, 1); }
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v) { return simt::reduce_group(mask, v, 2); }
template <typename T> static inline T __shfl_sync(unsigned, T v, int src) { return simt::exchange(v, [src](int) { return src; }); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int lm) { return simt::exchange(v, [lm](int l) { return l ^ lm; }); }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, int d) {
  const int lane = threadIdx.x & 31;
  const T r = simt::exchange(v, [d](int l) { return l - d < 0 ? l : l - d; });
  (void)lane;
  return r;
}
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int d) { return simt::exchange(v, [d](int l) { return l + d > 31 ? l : l + d; }); }

static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline unsigned __brev(unsigned x) {
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
  return r;
}

template <typename T> static inline T __ldg(const T* p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
template <typename T> static inline T __ldcg(const T* p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
template <typename T> static inline T __ldcs(const T* p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
static inline float __ldg(const float* p) { uint32_t b = __atomic_load_n((const uint32_t*)p, __ATOMIC_RELAXED); float f; memcpy(&f, &b, 4); return f; }
static inline float __ldcg(const float* p) { return __ldg(p); }

namespace simt { template <typename T> struct ident { using type = T; }; }
#define SIMT_V(T) typename simt::ident<T>::type   /* the value argument converts to the pointee type, like CUDA's overloads */
template <typename T> static inline T atomicAdd(T* p, SIMT_V(T) v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <typename T> static inline T atomicExch(T* p, SIMT_V(T) v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template <typename T> static inline T atomicOr(T* p, SIMT_V(T) v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
template <typename T> static inline T atomicCAS(T* p, SIMT_V(T) cmp, SIMT_V(T) v) {
  __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
  return cmp;   // the old value, like CUDA
}
template <typename T> static inline T atomicMin(T* p, SIMT_V(T) v) {
  T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
template <typename T> static inline T atomicMax(T* p, SIMT_V(T) v) {
  T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}

static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }

static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
static inline unsigned long min(unsigned long a, unsigned long b) { return a < b ? a : b; }
static inline unsigned long max(unsigned long a, unsigned long b) { return a > b ? a : b; }
static inline float min(float a, float b) { return fminf(a, b); }
static inline float max(float a, float b) { return fmaxf(a, b); }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }

namespace simt {
// run fn(arg) on n_threads fibers as ONE block (threadIdx.x = 0 .. n_threads-1, blockIdx.x = block)
void run_block(int n_threads, unsigned block, unsigned grid, void (*fn)(void*), void* arg);

// kernels WITHOUT block-level synchronisation: a launch is a loop over blocks and threads on the calling thread (one
// valid interleaving of a grid whose threads only meet in atomics)
template <typename F> struct SeqLaunch {
  unsigned g, b;
  F f;
  template <typename... A> void operator()(A... a) const {
    gridDim = uint3{g, 1, 1};
    blockDim = uint3{b, 1, 1};
    for (unsigned bx = 0; bx < g; bx++)
      for (unsigned tx = 0; tx < b; tx++) {
        blockIdx = uint3{bx, 0, 0};
        threadIdx = uint3{tx, 0, 0};
        f(a...);
      }
  }
};
template <typename F> SeqLaunch<F> seq_launch(unsigned long long g, unsigned b, F f) { return SeqLaunch<F>{(unsigned)g, b, f}; }

// kernels WITH block-level synchronisation: every block, one after the other, on b fibers
template <typename F> struct BlockLaunch {
  unsigned g, b;
  F f;
  template <typename... A> void operator()(A... a) const {
    auto call = [&]() { f(a...); };
    using C = decltype(call);
    for (unsigned bx = 0; bx < g; bx++) run_block((int)b, bx, g, [](void* p) { (*(C*)p)(); }, &call);
  }
};
template <typename F> BlockLaunch<F> block_launch(unsigned long long g, unsigned b, F f) { return BlockLaunch<F>{(unsigned)g, b, f}; }
}  // namespace simt

static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
