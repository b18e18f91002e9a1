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
