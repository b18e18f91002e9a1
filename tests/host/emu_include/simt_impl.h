// Definitions behind emu_include/cuda_runtime.h; included by exactly one translation unit per harness library.
#pragma once
#include <cuda_runtime.h>

#include <sys/mman.h>

thread_local uint3 threadIdx, blockIdx;
uint3 blockDim, gridDim;

// simt_switch(&from_sp, to_sp): save the callee-saved registers of the running fiber on its stack, switch stacks, restore
// the other fiber's (x86-64 System V; the image's only CPU architecture)
extern "C" void simt_switch(void** from_sp, void* to_sp);
asm(R"(
.text
.globl simt_switch
.type simt_switch,@function
simt_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size simt_switch,.-simt_switch
)");

namespace simt {
Block g_block;
struct Fiber {
  void* sp = nullptr;
  char* stack = nullptr;
  unsigned tid = 0, block = 0;
  bool done = false;
  Fiber* next = nullptr;      // ready queue / barrier wait list
  void (*fn)(void*) = nullptr;
  void* arg = nullptr;
};
constexpr size_t kStack = 256 * 1024;
static Fiber* g_cur = nullptr;
static Fiber g_main;                     // the OS thread's own context
static Fiber *g_ready_head = nullptr, *g_ready_tail = nullptr;
static int g_alive = 0;

static void ready_push(Fiber* f) {
  f->next = nullptr;
  if (g_ready_tail) g_ready_tail->next = f; else g_ready_head = f;
  g_ready_tail = f;
}
static Fiber* ready_pop() {
  Fiber* f = g_ready_head;
  if (f) { g_ready_head = f->next; if (!g_ready_head) g_ready_tail = nullptr; f->next = nullptr; }
  return f;
}
static void switch_to(Fiber* to) {
  Fiber* from = g_cur;
  g_cur = to;
  simt_switch(&from->sp, to->sp);
  // resumed: g_cur is this fiber again
  threadIdx = uint3{g_cur->tid, 0, 0};
  blockIdx = uint3{g_cur->block, 0, 0};
}
// the running fiber cannot continue: run the next ready one (or go back to the OS thread when none is left)
static void park() {
  Fiber* n = ready_pop();
  if (!n) {
    if (g_alive > 0) { fprintf(stderr, "simt: deadlock -- %d fibers wait at barriers that can never fill\n", g_alive); abort(); }
    n = &g_main;
  }
  switch_to(n);
}
void barrier_wait(Barrier& b) {
  if (++b.arrived == b.expected) {       // last one in: release the others, keep running
    b.arrived = 0;
    Fiber* f = b.head;
    b.head = b.tail = nullptr;
    while (f) { Fiber* nx = f->next; ready_push(f); f = nx; }
    return;
  }
  Fiber* me = g_cur;
  me->next = nullptr;
  if (b.tail) b.tail->next = me; else b.head = me;
  b.tail = me;
  park();
}
// Group reduction for lanes inside a divergent branch.  Fibers run one at a time and only switch at barriers, so a group
// can meet through a per-warp table keyed by its member mask: every member adds itself to its group's entry and parks
// until the last one has arrived.
struct GroupSlot { unsigned mask = 0, arrived = 0, phase = 0; unsigned vals[32]; Barrier bar; };
static GroupSlot g_groups[kMaxThreads / 32][32];
unsigned reduce_group(unsigned mask, unsigned v, int op) {
  const int lane = threadIdx.x & 31;
  GroupSlot& s = g_groups[threadIdx.x >> 5][__builtin_ctz(mask)];     // the group's lowest lane names its slot
  if (s.bar.expected != __builtin_popcount(mask)) { s.bar = Barrier(); s.bar.expected = __builtin_popcount(mask); }
  s.vals[lane] = v;
  barrier_wait(s.bar);
  unsigned r = op == 0 ? 0u : (op == 1 ? 0xffffffffu : 0u);
  for (int l = 0; l < 32; l++)
    if (mask >> l & 1u) r = op == 0 ? r + s.vals[l] : (op == 1 ? (s.vals[l] < r ? s.vals[l] : r) : (s.vals[l] > r ? s.vals[l] : r));
  barrier_wait(s.bar);
  return r;
}
static void fiber_entry() {
  Fiber* me = g_cur;
  threadIdx = uint3{me->tid, 0, 0};
  blockIdx = uint3{me->block, 0, 0};
  me->fn(me->arg);
  me->done = true;
  g_alive--;
  park();                                // never comes back
  abort();
}
void run_block(int n_threads, unsigned block, unsigned grid, void (*fn)(void*), void* arg) {
  blockDim = uint3{(unsigned)n_threads, 1, 1};
  gridDim = uint3{grid, 1, 1};
  g_block.n_threads = n_threads;
  g_block.bar = Barrier();
  g_block.bar.expected = n_threads;
  for (int w = 0; w < (n_threads + 31) / 32; w++) {
    g_block.warps[w].bar = Barrier();
    g_block.warps[w].bar.expected = (w + 1) * 32 <= n_threads ? 32 : n_threads - w * 32;
  }
  const size_t total = kStack * (size_t)n_threads;
  char* stacks = (char*)mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  if (stacks == (char*)MAP_FAILED) { fprintf(stderr, "simt: mmap of fiber stacks failed\n"); abort(); }
  Fiber* fb = new Fiber[n_threads];
  g_ready_head = g_ready_tail = nullptr;
  for (int t = 0; t < n_threads; t++) {
    Fiber& f = fb[t];
    f.tid = (unsigned)t; f.block = block; f.fn = fn; f.arg = arg;
    f.stack = stacks + kStack * (size_t)t;
    void** top = (void**)(f.stack + kStack);          // 16-byte aligned (mmap + multiple of 16)
    top[-1] = nullptr;                                // fake return address of fiber_entry
    top[-2] = (void*)&fiber_entry;                    // where simt_switch's ret lands
    for (int r = 3; r <= 8; r++) top[-r] = nullptr;   // rbp rbx r12 r13 r14 r15
    f.sp = (void*)(top - 8);
    ready_push(&f);
  }
  g_alive = n_threads;
  const uint3 t_save = threadIdx, b_save = blockIdx;
  g_cur = &g_main;
  park();                                             // returns when every fiber has finished
  threadIdx = t_save; blockIdx = b_save;
  delete[] fb;
  munmap(stacks, total);
}
}  // namespace simt

// what common.cuh declares and capi.cu defines in the real library
static char g_emu_err[512];
void b2t_set_error(const char* fmt, ...) { snprintf(g_emu_err, sizeof(g_emu_err), "%s", fmt); }
extern "C" const char* emu_last_error() { return g_emu_err; }
void b2t_count_launches(int) {}
int b2t_coop_limit() { return 0; }
int b2t_trace_limit() { return 0; }
