// Test infrastructure, never shipped: kimimaro_b200/csrc/field.cu (label statistics, the multi-source distance-field
// sweep, per-label arg-max, PDRF + target buckets, the grid-wide ball invalidation) compiled for the CPU against the
// SIMT emulation; the exported entry points take HOST pointers here.  The cooperative kernels run as one block of 1024
// emulated threads.  tests/test_field_emu_cpu.py runs them against the oracle.
#define B2T_HOST_EMU 1
#ifdef B2T_EMU_COMBINED
#include <cuda_runtime.h>
#else
#include "emu_include/simt_impl.h"
#endif

#include "../../kimimaro_b200/csrc/field.cu"

#ifndef B2T_EMU_COMBINED
// b2t_invalidate_ball_single calls the connected-components entry point of preamble.cu; the stand-alone harness of this
// file has no preamble.cu, so that one call is only exercised through the combined library (the product tests' soma case)
extern "C" __attribute__((visibility("default"))) int b2t_ccl26_roots(const void*, int, int64_t, int64_t, int64_t, uint32_t*,
                                                                      uint8_t*, void*) {
  return B2T_ERR_ARG;
}
#endif
