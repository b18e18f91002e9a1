// Test infrastructure: the part of <cooperative_groups.h> the kernels use, for the SIMT emulation.  A cooperative
// kernel is emulated as ONE block, so the grid barrier is the block barrier.
#pragma once
#include <cuda_runtime.h>
namespace cooperative_groups {
struct grid_group {
  void sync() const { __syncthreads(); }
};
inline grid_group this_grid() { return grid_group{}; }
}  // namespace cooperative_groups
