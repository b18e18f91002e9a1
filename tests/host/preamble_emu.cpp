// Test infrastructure, never shipped: kimimaro_b200/csrc/preamble.cu (connected components, hole filling, relabel,
// sequential segment sums, path gathering) compiled for the CPU against the SIMT emulation.  Its kernels have no
// block-level synchronisation, so B2T_LAUNCH runs a launch as a loop over blocks and threads, and the exported entry
// points (b2t_ccl26_roots, b2t_fill_voids, ...) take HOST pointers here.  tests/test_preamble_emu_cpu.py runs them
// against the oracle.
#define B2T_HOST_EMU 1
#ifdef B2T_EMU_COMBINED
#include <cuda_runtime.h>
#else
#include "emu_include/simt_impl.h"
#endif

#include "../../kimimaro_b200/csrc/preamble.cu"
