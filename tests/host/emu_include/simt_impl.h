// Definitions behind emu_include/cuda_runtime.h; included by exactly one translation unit per harness library.
#pragma once
#include <cuda_runtime.h>

thread_local uint3 threadIdx, blockIdx;
uint3 blockDim, gridDim;
namespace simt {
Block g_block;
struct Start { void (*fn)(void*); void* arg; unsigned tid, block; };
static void* entry(void* p) {
  Start* s = (Start*)p;
  threadIdx = uint3{s->tid, 0, 0};
  blockIdx = uint3{s->block, 0, 0};
  s->fn(s->arg);
  return nullptr;
}
void run_block(int n_threads, unsigned block, unsigned grid, void (*fn)(void*), void* arg) {
  blockDim = uint3{(unsigned)n_threads, 1, 1};
  gridDim = uint3{grid, 1, 1};
  g_block.n_threads = n_threads;
  pthread_barrier_init(&g_block.bar, nullptr, n_threads);
  for (int w = 0; w < n_threads / 32; w++) pthread_barrier_init(&g_block.warps[w].bar, nullptr, 32);
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
  Start* st = (Start*)malloc(sizeof(Start) * n_threads);
  pthread_attr_t attr;
  pthread_attr_init(&attr);
  pthread_attr_setstacksize(&attr, 256 * 1024);
  for (int t = 0; t < n_threads; t++) {
    st[t] = Start{fn, arg, (unsigned)t, block};
    if (pthread_create(&th[t], &attr, entry, &st[t]) != 0) { fprintf(stderr, "pthread_create failed\n"); abort(); }
  }
  for (int t = 0; t < n_threads; t++) pthread_join(th[t], nullptr);
  pthread_attr_destroy(&attr);
  pthread_barrier_destroy(&g_block.bar);
  for (int w = 0; w < n_threads / 32; w++) pthread_barrier_destroy(&g_block.warps[w].bar);
  free(th); free(st);
}
}  // namespace simt

// what common.cuh declares and capi.cu defines in the real library
static char g_emu_err[512];
void b2t_set_error(const char* fmt, ...) { snprintf(g_emu_err, sizeof(g_emu_err), "%s", fmt); }
extern "C" const char* emu_last_error() { return g_emu_err; }
void b2t_count_launches(int) {}
int b2t_coop_limit() { return 0; }
int b2t_trace_limit() { return 0; }
bool b2t_claim_window_built() { return true; }
float g_emu_claim_window = 0.0f;
float b2t_claim_window() { return g_emu_claim_window; }
