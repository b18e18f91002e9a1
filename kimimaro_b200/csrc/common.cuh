// Shared helpers for libb2t.so (sm_100a only).  Not a public header: the C ABI is include/b2t.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b2t.h"

#define B2T_EXPORT extern "C" __attribute__((visibility("default")))

void b2t_set_error(const char* fmt, ...);
void b2t_count_launches(int n);  // bookkeeping for b2t_launch_count()
int b2t_coop_limit();    // max blocks per SM for cooperative kernels (0 = whatever fits), see b2t_set_launch_limits
int b2t_trace_limit();   // max blocks per SM for the path-loop kernel (0 = whatever fits)

#define B2T_CUDA_TRY(expr)                                                                  \
  do {                                                                                      \
    cudaError_t e_ = (expr);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      b2t_set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(e_)); \
      return B2T_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

#define B2T_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      b2t_set_error(__VA_ARGS__);              \
      return B2T_ERR_ARG;                      \
    }                                          \
  } while (0)

// Kernel launches of preamble.cu and field.cu go through macros so that the CPU suite can compile those files with g++
// against the SIMT emulation of tests/host/emu_include/cuda_runtime.h (B2T_HOST_EMU) and run the very kernels on host
// arrays against the oracle.  There B2T_LAUNCH (kernels without block-level synchronisation) is a loop over blocks and
// threads, B2T_LAUNCH_SYNC (kernels that use __syncthreads or warp intrinsics) runs every block on OS threads.
#ifdef B2T_HOST_EMU
#define B2T_LAUNCH(kernel_, grid_, block_, stream_) simt::seq_launch((grid_), (block_), [](auto... a_) { kernel_(a_...); })
#define B2T_LAUNCH_SYNC(kernel_, grid_, block_, stream_) simt::block_launch((grid_), (block_), [](auto... a_) { kernel_(a_...); })
#else
#define B2T_LAUNCH(kernel_, grid_, block_, stream_) kernel_<<<(grid_), (block_), 0, (stream_)>>>
#define B2T_LAUNCH_SYNC(kernel_, grid_, block_, stream_) kernel_<<<(grid_), (block_), 0, (stream_)>>>
#endif

// 26-neighbourhood in the enumeration order the reference uses
// (ext/skeletontricks/dijkstra_invalidation.hpp:60-124): -x,+x,-y,+y,-z,+z, xy, yz, xz diagonals, corners.
__device__ __constant__ static const int8_t kDX[26] = {-1, 1, 0, 0, 0, 0, -1, -1, 1, 1, 0, 0, 0, 0, -1, -1, 1, 1, -1, 1, -1, -1, 1, 1, -1, 1};
__device__ __constant__ static const int8_t kDY[26] = {0, 0, -1, 1, 0, 0, -1, 1, -1, 1, -1, -1, 1, 1, 0, 0, 0, 0, -1, -1, 1, -1, 1, -1, 1, 1};
__device__ __constant__ static const int8_t kDZ[26] = {0, 0, 0, 0, -1, 1, 0, 0, 0, 0, -1, 1, -1, 1, -1, 1, -1, 1, -1, -1, -1, 1, -1, 1, 1, 1};

static inline int b2t_ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
