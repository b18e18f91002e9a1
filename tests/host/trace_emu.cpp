// Test infrastructure, never shipped: the device functions of kimimaro_b200/csrc/trace.cu compiled for the CPU
// against the SIMT emulation of emu_include/cuda_runtime.h, so that tests/test_trace_emu_cpu.py can run the engine's
// invalidation -- key-ordered rounds (the default), the literal heap of the strict mode and the hop-synchronous rounds --
// on the very source text the GPU runs and compare it with the oracle (orc_invalidate_window / _heap / _rounds).
#define B2T_HOST_EMU 1
#ifndef B2T_EMU_COMBINED        /* the stand-alone harness shrinks the strict heap's own region so that small labels spill */
#define B2T_HEAP_PER_VOXEL 1
#define B2T_HEAP_SLACK 64
#endif
#ifdef B2T_EMU_COMBINED
#include <cuda_runtime.h>
#else
#include "emu_include/simt_impl.h"
#endif   // emu_include/ comes first on the include path: <cuda_runtime.h> is the emulation

#include "../../kimimaro_b200/csrc/trace.cu"

namespace {
struct InvArgs {
  Arena A; LabelDesc L; Pools P; const uint32_t* seeds; uint32_t n_seeds; float scale, konst, delta;
  uint32_t *r0, *r1, *r2, *r3; int mode; uint32_t result; uint32_t overflow; int team;
};
Shared g_S;
Local g_Lc;
template <bool TEAM> void inv_run(InvArgs* a) {
  Team T{0u, 0u};
  uint32_t n;
  if (a->mode == B2T_INVALIDATE_STRICT) {
    n = invalidate_strict<TEAM>(a->A, a->L, a->P, 0, a->seeds, a->n_seeds, a->scale, a->konst, g_S, T);
    if (threadIdx.x == 0) a->overflow = g_S.n_proc;
  } else if (a->mode == B2T_INVALIDATE_WINDOW) {
    n = invalidate_window<TEAM>(a->A, a->L, a->seeds, a->n_seeds, a->scale, a->konst, a->delta, a->r0, a->r1, a->r2, a->r3, g_S,
                                g_Lc, T);
  } else {
    n = invalidate<TEAM>(a->A, a->L, a->seeds, a->n_seeds, a->scale, a->konst, a->r0, a->r1, a->r2, a->r3, g_S, T);
  }
  if (threadIdx.x == 0) a->result = n;
}
void inv_thread(void* p) {
  InvArgs* a = (InvArgs*)p;
  if (a->team) inv_run<true>(a); else inv_run<false>(a);
}
}  // namespace

// One block of the engine's 512 threads runs roll_invalidation_ball_inside_component on a label of the dense arena.
// claim: one 64-bit word per voxel, ~0 = valid (the engine's kValid), 0 = invalid; edited in place.
// team = 1 runs the team instantiation of the function (a cluster of one emulated CTA).
// mode: B2T_INVALIDATE_ROUNDS (hop rounds), _WINDOW (invalidate_window with `delta`, already in physical units), _STRICT
// (the literal heap; spill_words = size of the spill arena behind the label's own heap region, so that a test can make
// the heap outgrow its region).  Returns -4 (B2T_ERR_CAPACITY) when the strict heap fits nowhere.
extern "C" long emu_invalidate(const uint32_t* cc, const float* dbf, unsigned long long* claim, int sx, int sy, int sz,
                               float wx, float wy, float wz, uint32_t segid, uint32_t n_fg, const uint32_t* seeds,
                               uint32_t n_seeds, float scale, float konst, float delta, int mode, long spill_words,
                               int team) {
  InvArgs a;
  memset(&a, 0, sizeof(a));
  a.A.cc = cc; a.A.dbf = dbf; a.A.claim = claim;
  a.A.d = Dims{sx, sy, sz, (uint32_t)(sx * sy)};
  a.A.wx = wx; a.A.wy = wy; a.A.wz = wz;
  a.L.segid = segid; a.L.n_fg = n_fg; a.L.bbox_x0 = 0; a.L.bbox_x1 = (uint32_t)(sx - 1);
  a.seeds = seeds; a.n_seeds = n_seeds; a.scale = scale; a.konst = konst; a.delta = delta; a.mode = mode; a.team = team;
  uint32_t* scratch = (uint32_t*)malloc(sizeof(uint32_t) * 4 * (size_t)(n_fg + 1));
  a.r0 = scratch; a.r1 = scratch + n_fg; a.r2 = scratch + 2 * (size_t)n_fg; a.r3 = scratch + 3 * (size_t)n_fg;
  unsigned long long bump = 0;
  uint32_t* heap = nullptr;
  g_S.heap_cap = 0;
  if (mode == B2T_INVALIDATE_STRICT) {
    const uint64_t stat = b2t_trace_heap_words(n_fg, 1) - 2;
    heap = (uint32_t*)malloc(sizeof(uint32_t) * (stat + (uint64_t)spill_words));
    a.P.heap = heap; a.P.heap_static = stat; a.P.heap_words = stat + (uint64_t)spill_words; a.P.heap_bump = &bump;
  }
  simt::run_block(kThreads, 0, 1, inv_thread, &a);
  free(scratch);
  free(heap);
  return a.overflow ? -4L : (long)a.result;
}

// The whole path loop (trace_kernel: find_target -> railroad -> invalidate -> rail, trace.py:196-267) for a batch of
// labels on ONE emulated block, which pulls the labels one after another from the work counter like a CTA on the GPU.
// Arguments as b2t_trace_batch (host arrays instead of device arrays); claim_window in physical units.
namespace {
struct KArgs { Arena A; const LabelDesc* descs; Pools P; Params prm; uint32_t n_team; Shared* team; };
void kernel_thread(void* p) {
  KArgs* k = (KArgs*)p;
  trace_kernel(k->A, k->descs, k->P, k->prm, k->n_team, k->team);
}
}  // namespace

extern "C" int emu_trace_batch(const uint32_t* cc, const float* dbf, float* pdrf, float* dist, unsigned long long* claim,
                               uint32_t* stamp, int sx, int sy, int sz, float wx, float wy, float wz, const void* desc,
                               int n_desc, float scale, float konst, float soma_scale, float soma_const, int fix_branching,
                               int nbuckets, const unsigned long long* keys, const uint32_t* hist, const uint32_t* cursor,
                               uint32_t* scratch, uint32_t* paths, const uint32_t* targets, uint32_t* out_len,
                               uint32_t* out_npaths, int32_t* out_status, uint32_t* out_stats, uint32_t* work_counter,
                               int inval_mode, float claim_window, uint32_t* heap, uint64_t heap_words,
                               uint64_t heap_static_words, int n_team) {
  static_assert(sizeof(LabelDesc) == 80, "LabelDesc layout");
  KArgs k;
  memset(&k, 0, sizeof(k));
  k.A.cc = cc; k.A.dbf = dbf; k.A.pdrf = pdrf; k.A.dist = dist; k.A.claim = claim; k.A.stamp = stamp;
  k.A.d = Dims{sx, sy, sz, (uint32_t)(sx * sy)};
  k.A.wx = wx; k.A.wy = wy; k.A.wz = wz;
  k.descs = (const LabelDesc*)desc;
  k.P.keys = keys; k.P.hist = hist; k.P.cursor = cursor; k.P.scratch = scratch; k.P.paths = paths; k.P.targets = targets;
  k.P.out_len = out_len; k.P.out_npaths = out_npaths; k.P.out_status = out_status; k.P.out_stats = out_stats;
  k.P.work_counter = work_counter;
  k.prm = Params{scale, konst, soma_scale, soma_const, nbuckets, n_desc, fix_branching ? 1 : 0, inval_mode, claim_window};
  unsigned long long bump = 0;
  if (inval_mode == B2T_INVALIDATE_STRICT) {
    if (!heap || heap_static_words < 2 || heap_words < heap_static_words) return -1;
    k.P.heap = heap + 2; k.P.heap_words = heap_words - 2; k.P.heap_static = heap_static_words - 2; k.P.heap_bump = &bump;
  }
  *work_counter = 0;
  // the first n_team jobs go to teams (emulated clusters have ONE CTA: block b < n_team is the team of job b), then one
  // more block pulls the remaining jobs from the work counter like a solo CTA on the GPU
  k.n_team = (uint32_t)n_team;
  k.team = (Shared*)calloc((size_t)n_team + 1, sizeof(Shared));
  for (int b = 0; b < n_team; b++) simt::run_block(kThreads, (unsigned)b, (unsigned)n_team + 1, kernel_thread, &k);
  if (n_desc > n_team) simt::run_block(kThreads, (unsigned)n_team, (unsigned)n_team + 1, kernel_thread, &k);
  free(k.team);
  return 0;
}

#ifdef B2T_EMU_COMBINED
// b2t_trace_batch of include/b2t.h on host arrays (the real launcher is not compiled under B2T_HOST_EMU)
extern "C" __attribute__((visibility("default"))) int b2t_trace_batch(
    const uint32_t* d_cc, const float* d_dbf, float* d_pdrf, float* d_dist, uint64_t* d_claim, uint32_t* d_stamp, int64_t sx,
    int64_t sy, int64_t sz, float wx, float wy, float wz, const void* d_desc, int n_desc, float scale, float konst,
    float soma_scale, float soma_const, int fix_branching, int nbuckets, const uint64_t* d_keys, const uint32_t* d_hist,
    const uint32_t* d_cursor, uint32_t* d_scratch, uint32_t* d_paths, const uint32_t* d_targets, uint32_t* d_out_len,
    uint32_t* d_out_npaths, int32_t* d_out_status, uint32_t* d_out_stats, uint32_t* d_work_counter, int invalidation_mode,
    float claim_window_voxels, uint32_t* d_heap, uint64_t heap_words, uint64_t heap_static_words, int n_team, void*, void*) {
  if (n_desc <= 0) return 0;
  const float wmin = wx < wy ? (wx < wz ? wx : wz) : (wy < wz ? wy : wz);
  return emu_trace_batch(d_cc, d_dbf, d_pdrf, d_dist, (unsigned long long*)d_claim, d_stamp, (int)sx, (int)sy, (int)sz, wx, wy,
                         wz, d_desc, n_desc, scale, konst, soma_scale, soma_const, fix_branching, nbuckets,
                         (const unsigned long long*)d_keys, d_hist, d_cursor, d_scratch, d_paths, d_targets, d_out_len,
                         d_out_npaths, d_out_status, d_out_stats, d_work_counter, invalidation_mode,
                         claim_window_voxels * wmin, d_heap, heap_words, heap_static_words, n_team);
}
#endif
