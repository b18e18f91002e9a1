// Phase B of the TEASAR trace: the per-label path loop of kimimaro/trace.py:196-267 as ONE
// device-resident kernel (sm_100a).  One CTA owns one label for the whole loop -- no host round
// trip per path -- and CTAs pull labels (largest first) from a device work counter: two 512-thread
// CTAs per SM, 296 labels in flight on the 148 SMs (the few labels above 100 k voxels get a cluster).
//
//   per path:  target   manual_targets_before (LIFO) | CachedTargetFinder | manual_targets_after
//                       (trace.py:225-230; pyx:1008-1045).  The finder is a block arg-max over the
//                       current DAF bucket (keys built by field.cu), ties by largest index (rule T5).
//              road     dijkstra3d.railroad(PDRF, target) (trace.py:240-242): delta-stepping over
//                       (distance, voxel) pairs with lazy deletion from the target (near lists in shared
//                       memory for a solo CTA), atomicMin on float bits, stops once every voxel at least
//                       as close as the best rail-adjacent voxel is final; parents by rule T3.
//              cull     soma only (trace.py:246-251), float64 with the reference's uint32 wrap.
//              erase    roll_invalidation_ball_inside_component (pyx:373-418 ->
//                       dijkstra_invalidation.hpp:239-332), three claim orders (b2t_set_invalidation_mode):
//                         window  parallel rounds ordered by the reference's heap key (distance to the seed) in
//                                 windows of one voxel edge: packed (dist, seed) 64-bit atomicMin per candidate (default)
//                         strict  the reference's std::priority_queue LITERALLY (libstdc++ push_heap / pop_heap order
//                                 for equal keys included), one warp per label: identical to the compiled reference
//                         rounds  hop-synchronous rounds (round 1's order; kept for A/B)
//              rail     PDRF[path] = 0 (trace.py:261-263)
//
// Per-voxel fields are the dense arrays of field.cu (cc, dbf, pdrf, dist, claim, stamp); per-label
// queues live in a scratch pool indexed by the label's foreground-count prefix sum.
#ifndef B2T_HOST_EMU
#include <cooperative_groups.h>
#endif
#include "common.cuh"

// B2T_HOST_EMU: tests/host/trace_emu.cpp compiles this file with g++ against a SIMT emulation (one OS thread per CUDA
// thread of ONE block, barriers for __syncthreads and the warp intrinsics) so that the CPU suite can run the device
// functions below against the oracle; only the PTX timer read and the host launcher are left out of that build.
#ifdef B2T_HOST_EMU
#define B2T_GLOBALTIMER(v_) v_ = 0
#else
#define B2T_GLOBALTIMER(v_) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v_))
#endif

// B2T_TRACE_PROF (variant build "prof"): thread 0 of a label's team accumulates SM cycles per phase of the path loop;
// b2t_trace_prof_read copies the table of the first 64 jobs (the largest labels) to the host.  Off in the shipped build.
#ifdef B2T_TRACE_PROF
#define PROF_START() long long pt_ = 0; if (tid == 0) pt_ = clock64()
#define PROF_LAP(k_) do { if (tid == 0) { const long long n_ = clock64(); S.prof[k_] += (unsigned long long)(n_ - pt_); pt_ = n_; } } while (0)
#define PROF_COUNT(k_) do { if (tid == 0) S.prof[k_] += 1ull; } while (0)
#define PROF_RESET() do { if (tid == 0) pt_ = clock64(); } while (0)
#else
#define PROF_RESET() do { } while (0)
#define PROF_START() do { } while (0)
#define PROF_LAP(k_) do { } while (0)
#define PROF_COUNT(k_) do { } while (0)
#endif

namespace {

#ifdef B2T_TRACE_PROF
__device__ unsigned long long g_prof[64][16];
#endif

constexpr uint32_t kInfBits = 0x7f800000u;
constexpr unsigned long long kValid = ~0ull;
#ifndef B2T_TRACE_THREADS
#define B2T_TRACE_THREADS 512   // 16 warps: a whole narrow batch expands in one pass
#endif
constexpr int kThreads = B2T_TRACE_THREADS;
#ifndef B2T_HEAP_PER_VOXEL
#define B2T_HEAP_PER_VOXEL 4         // (the CPU harness builds with smaller numbers to exercise the spill path)
#define B2T_HEAP_SLACK 4096
#endif
constexpr int kHeapPerVoxel = B2T_HEAP_PER_VOXEL;   // strict mode: static heap nodes per foreground voxel of a label ...
constexpr int kHeapSlack = B2T_HEAP_SLACK;          // ... plus this many; a heap that outgrows it moves to the spill arena
#ifndef B2T_TRACE_MINB
#define B2T_TRACE_MINB 2        // resident CTAs per SM: 64 registers per thread.  Measured on synthetic-512 (profiles/
                                // r02_trace_occupancy_ab.jsonl): 3 CTAs (42 registers, spills in the 9-neighbour tasks) 47 ms,
                                // 2 CTAs 27 ms, 1 CTA (128 registers, no spills) 44 ms -- the loop is latency-bound per label
                                // and throughput-bound over labels, so resident labels and spill-free code both count
#endif
constexpr int kWarps = kThreads / 32;
constexpr int kScratchPerVoxel = 22;   // u32 of queue scratch per foreground voxel of a label (see trace_label); even
#ifndef B2T_RR_BATCH
#define B2T_RR_BATCH 1                 // railroad's batch target in units of (2 .. 8) voxels per warp and round
#endif
#ifndef B2T_RR_CAP_SHIFT
#define B2T_RR_CAP_SHIFT 0             // (the CPU harness shrinks railroad's lists to exercise the overflow rebuild)
#endif

struct Dims {
  int sx, sy, sz;
  uint32_t sxy;
};

struct Arena {
  const uint32_t* cc;
  const float* dbf;
  float* pdrf;
  float* dist;
  unsigned long long* claim;
  uint32_t* stamp;
  Dims d;
  float wx, wy, wz;
};

// One label to trace.  Mirrors the arguments kimimaro/intake.py:494-504 hands to trace().
struct LabelDesc {
  uint32_t segid;       // value of this label in cc
  uint32_t root;        // linear index of the root voxel
  uint32_t n_fg;        // foreground voxels (np.count_nonzero(labels), trace.py:211)
  uint32_t region_off;  // prefix sum of n_fg over the batch: scratch region starts at kScratchPerVoxel * region_off
  uint32_t path_off;    // first slot of this label in the path pool
  uint32_t path_cap;    // slots available
  uint32_t tb_off, tb_n;  // manual_targets_before: targets[tb_off .. tb_off+tb_n), popped from the end
  uint32_t ta_off, ta_n;  // manual_targets_after
  uint32_t max_paths;   // 0xffffffff = None
  uint32_t soma_mode;   // trace.py:118-127
  float soma_radius;    // dbf_max * soma_invalidation_scale + soma_invalidation_const (float32)
  uint32_t bucket_row;  // row of this label in the (label x bucket) tables = its cc id
  uint32_t soma_done;   // 1: the one-off soma invalidation already ran grid-wide (b2t_invalidate_ball)
  uint32_t pre_invalid; // voxels it invalidated
  uint32_t bbox_x0, bbox_x1;  // x extent of the label's bounding box (inclusive): the reference runs on the bbox crop
                              // (intake.py:463-466), and the x faces of THAT array are where the corner entries of its
                              // neighbour table alias (strict mode reproduces the duplicate pushes)
  uint32_t single_path; // 1: one path to the first manual target and nothing else (point_to_point, trace.py:358-390)
  uint32_t reserved1;
};

struct Params {
  float scale, konst;            // teasar_params scale / const
  float soma_scale, soma_const;  // soma_invalidation_scale / _const
  int nbuckets;
  int n_desc;
  int fix_branching;             // 0: paths come from the parental field already in A.dist (trace.py:154-158, 244)
  int inval_mode;                // B2T_INVALIDATE_ROUNDS / _WINDOW / _STRICT
  float claim_window;            // _WINDOW: width of a key-ordered round in physical units
};

struct Pools {
  const unsigned long long* keys;  // bucket-partitioned (daf_bits << 32 | index)
  const uint32_t* hist;            // bucket sizes
  const uint32_t* cursor;          // bucket ends
  uint32_t* scratch;               // kScratchPerVoxel * sum(n_fg) u32
  uint32_t* paths;                 // path pool: voxel indices, each path terminated by 0xffffffff
  const uint32_t* targets;         // manual targets (linear indices)
  uint32_t* out_len;               // per desc: slots written
  uint32_t* out_npaths;            // per desc
  int32_t* out_status;             // per desc: 0 ok, <0 error
  uint32_t* out_stats;             // per desc x 4: relaxations, rounds, invalidated, elapsed microseconds
  uint32_t* work_counter;
  uint32_t* heap;                  // strict mode: 3 u32 per heap node (key | voxel | seed, one array each per label)
  unsigned long long heap_words;   // size of `heap`; words past the static regions are the spill arena
  unsigned long long heap_static;  // words taken by the static regions
  unsigned long long* heap_bump;   // spill arena: words handed out so far
};

__device__ __forceinline__ void unravel(uint32_t loc, const Dims& d, int& x, int& y, int& z) {
  z = loc / d.sxy;
  const uint32_t r = loc - (uint32_t)z * d.sxy;
  y = r / (uint32_t)d.sx;
  x = r - (uint32_t)y * d.sx;
}

// ---- the team that traces one label -----------------------------------------------------------------
// solo: one CTA (what every label but the largest few gets); team: the kCluster CTAs of a thread-block cluster, for the
// labels whose searches keep 128 warps busy -- the path loop of a label is sequential, so the largest label IS the tail of
// the kernel (45 ms of a 45 ms launch on the benchmark volume when one CTA has it to itself).  The code below is written
// against the team: strides, barriers and reductions go through these helpers, the per-label state `Shared` lives in
// shared memory (solo) or in a global-memory slot all CTAs of the cluster see (team), the work lists are in global
// scratch either way.
#ifdef B2T_HOST_EMU
constexpr int kCluster = 1;        // the CPU harness runs one CTA at a time: a "cluster" of one exercises the team code path
#else
constexpr int kCluster = 8;
#endif

struct Team {
  uint32_t rank;      // this CTA within the team
  uint32_t par;       // parity of the team reductions' exchange slots
};

template <bool TEAM> __device__ __forceinline__ uint32_t t_threads() { return TEAM ? kCluster * kThreads : kThreads; }
template <bool TEAM> __device__ __forceinline__ uint32_t t_warps() { return TEAM ? kCluster * kWarps : kWarps; }
template <bool TEAM> __device__ __forceinline__ uint32_t t_tid(const Team& T) { return TEAM ? T.rank * kThreads + threadIdx.x : threadIdx.x; }
template <bool TEAM> __device__ __forceinline__ uint32_t t_warp(const Team& T) { return TEAM ? T.rank * kWarps + (threadIdx.x >> 5) : (threadIdx.x >> 5); }

template <bool TEAM> __device__ __forceinline__ void team_sync() {
#ifdef B2T_HOST_EMU
  __syncthreads();
#else
  if (TEAM) cooperative_groups::this_cluster().sync(); else __syncthreads();
#endif
}
// a value of the team's shared state that other threads change between barriers
template <bool TEAM, typename V> __device__ __forceinline__ V t_peek(const V* p) { return TEAM ? __ldcg(p) : *p; }

// ---- block-wide reductions (all threads must call) ------------------------------------------------
__device__ __forceinline__ uint32_t block_min_u32(uint32_t v, uint32_t* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  uint32_t r = s_red[0];
#pragma unroll
  for (int i = 1; i < kWarps; i++) r = min(r, s_red[i]);
  return r;
}

__device__ __forceinline__ unsigned long long block_max_u64(unsigned long long v, unsigned long long* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t > v ? t : v;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  unsigned long long r = s_red[0];
#pragma unroll
  for (int i = 1; i < kWarps; i++) r = s_red[i] > r ? s_red[i] : r;
  return r;
}

__device__ __forceinline__ uint32_t block_sum_u32(uint32_t v, uint32_t* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < kWarps; i++) r += s_red[i];
  return r;
}

// Per-CTA scratch of the block reductions (always shared memory).
struct Local {
  uint32_t red32[kWarps];
  unsigned long long red64[kWarps];
  unsigned long long bcast;
};

// Per-label state: shared memory (solo) or a global-memory slot of the team.
// railroad_solo: the counters of one round (two sets, used alternately)
struct RoundCounters {
  uint32_t n_keep, n_proc, n_next, min_next, min_far, overflow;
};

struct Shared {
  unsigned long long best;   // railroad: (dist_bits << 32) | voxel of the best rail-adjacent voxel
  RoundCounters rc[2];
  unsigned long long x64[2][kCluster];   // team reductions: one value per CTA, two parities
  uint32_t n_keep, n_proc, n_next, n_touched;
  uint32_t r32[2];           // small results handed from one thread to the team
  uint32_t job;
  int bucket;
  uint32_t relax, rounds, invalidated;
  uint32_t* heap_k;          // strict mode: this label's heap (its own region, or the spill arena once it outgrew that)
  uint32_t heap_cap;         // 0 = not set up yet
#ifdef B2T_TRACE_PROF
  unsigned long long prof[16];
#endif
};

// team-wide reductions (all threads of the team must call): block reduction, then one value per CTA through `S`
template <bool TEAM, int OP>   // OP 0: min u32, 1: max u64, 2: sum u32
__device__ __forceinline__ unsigned long long team_reduce(unsigned long long v, Shared& S, Local& Lc, Team& T) {
  unsigned long long r;
  if (OP == 0) r = block_min_u32((uint32_t)v, Lc.red32);
  else if (OP == 1) r = block_max_u64(v, Lc.red64);
  else r = block_sum_u32((uint32_t)v, Lc.red32);
  if (!TEAM) return r;
  T.par ^= 1u;
  if (threadIdx.x == 0) S.x64[T.par][T.rank] = r;
  team_sync<TEAM>();
  if (threadIdx.x < 32) {
    unsigned long long x = threadIdx.x < kCluster ? __ldcg(&S.x64[T.par][threadIdx.x]) : (OP == 0 ? 0xffffffffull : 0ull);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long y = __shfl_xor_sync(0xffffffffu, x, o);
      x = OP == 0 ? (y < x ? y : x) : (OP == 1 ? (y > x ? y : x) : x + y);
    }
    if (threadIdx.x == 0) Lc.bcast = x;
  }
  __syncthreads();
  r = Lc.bcast;
  __syncthreads();
  return r;
}
template <bool TEAM> __device__ __forceinline__ uint32_t team_min_u32(uint32_t v, Shared& S, Local& Lc, Team& T) {
  return (uint32_t)team_reduce<TEAM, 0>(v, S, Lc, T);
}
template <bool TEAM> __device__ __forceinline__ unsigned long long team_max_u64(unsigned long long v, Shared& S, Local& Lc, Team& T) {
  return team_reduce<TEAM, 1>(v, S, Lc, T);
}
template <bool TEAM> __device__ __forceinline__ uint32_t team_sum_u32(uint32_t v, Shared& S, Local& Lc, Team& T) {
  return (uint32_t)team_reduce<TEAM, 2>(v, S, Lc, T);
}

// ---- CachedTargetFinder.find_target ------------------------------------------------------------------
template <bool TEAM>
__device__ uint32_t find_target(const Arena& A, const LabelDesc& L, const Pools& P, const Params& prm, Shared& S, Local& Lc,
                                Team& T) {
  for (;;) {
    const int b = S.bucket;
    if (b < 0) return 0xffffffffu;
    const size_t row = (size_t)L.bucket_row * prm.nbuckets + b;
    const uint32_t end = P.cursor[row], n = P.hist[row];
    const unsigned long long* k = P.keys + (end - n);
    unsigned long long best = 0;
    for (uint32_t i = t_tid<TEAM>(T); i < n; i += t_threads<TEAM>()) {
      const unsigned long long key = k[i];
      const uint32_t v = (uint32_t)key;
      if (__ldcg(&A.claim[v]) == kValid && key + 1 > best) best = key + 1;  // +1 so that key 0 is distinguishable from "none"
    }
    best = team_max_u64<TEAM>(best, S, Lc, T);
    if (best != 0) return (uint32_t)(best - 1);
    team_sync<TEAM>();
    if (t_tid<TEAM>(T) == 0) S.bucket = b - 1;
    team_sync<TEAM>();
  }
}

// (d) + (e) of railroad: walk back from the best rail-adjacent voxel (rule T4) along parents by rule T3, then put the
// distance field back to +inf on every voxel the search touched.  Shared by the team form and the solo form below.
template <bool TEAM>
__device__ __forceinline__ uint32_t railroad_finish(const Arena& A, const LabelDesc& L, uint32_t target, const uint32_t* touched,
                                                    uint32_t* out, uint32_t out_cap, uint32_t relax, uint32_t rounds, Shared& S,
                                                    Local& Lc, Team& T) {
  const int lane = threadIdx.x & 31;
  const uint32_t tid = t_tid<TEAM>(T), nth = t_threads<TEAM>(), tw = t_warp<TEAM>(T);
  const uint32_t seg = L.segid;
  int dx = 0, dy = 0, dz = 0;
  if (lane < 26) { dx = kDX[lane]; dy = kDY[lane]; dz = kDZ[lane]; }
  const int64_t off = (int64_t)dx + (int64_t)dy * A.d.sx + (int64_t)dz * A.d.sxy;
  PROF_START();
  // (d) walk back: rail voxel, then parents by rule T3
  uint32_t len = 0;
  const unsigned long long best = S.best;
  if (tw == 0) {
    if (best == ~0ull) {
      if (lane == 0 && out_cap > 0) out[0] = target;
      len = 1;
    } else {
      uint32_t loc = (uint32_t)best;
      {
        int x, y, z;
        unravel(loc, A.d, x, y, z);
        const int nx = x + dx, ny = y + dy, nz = z + dz;
        bool israil = false;
        uint32_t v = 0;
        if (lane < 26 && nx >= 0 && ny >= 0 && nz >= 0 && nx < A.d.sx && ny < A.d.sy && nz < A.d.sz) {
          v = (uint32_t)((int64_t)loc + off);
          israil = (__ldg(&A.cc[v]) == seg) && (__ldcg(&A.pdrf[v]) == 0.0f);
        }
        const uint32_t m = __ballot_sync(0xffffffffu, israil);
        const int first = __ffs(m) - 1;
        const uint32_t rail = __shfl_sync(0xffffffffu, v, first < 0 ? 0 : first);
        if (lane == 0 && len < out_cap) out[len] = rail;
        len++;
      }
      uint32_t guard = 0;
      while (loc != target && guard <= L.n_fg) {
        if (lane == 0 && len < out_cap) out[len] = loc;
        len++;
        int x, y, z;
        unravel(loc, A.d, x, y, z);
        const int nx = x + dx, ny = y + dy, nz = z + dz;
        unsigned long long key = ~0ull;
        if (lane < 26 && nx >= 0 && ny >= 0 && nz >= 0 && nx < A.d.sx && ny < A.d.sy && nz < A.d.sz) {
          const uint32_t v = (uint32_t)((int64_t)loc + off);
          if (__ldg(&A.cc[v]) == seg) {
            const uint32_t dv = __float_as_uint(__ldcg(&A.dist[v]));
            if (dv < kInfBits) key = ((unsigned long long)dv << 32) | (unsigned long long)lane;
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const unsigned long long t = __shfl_xor_sync(0xffffffffu, key, o);
          key = t < key ? t : key;
        }
        if (key == ~0ull) break;  // cannot happen for a settled voxel
        const int dir = (int)(key & 31u);
        loc = (uint32_t)((int64_t)loc + (int64_t)kDX[dir] + (int64_t)kDY[dir] * A.d.sx + (int64_t)kDZ[dir] * A.d.sxy);
        guard++;
      }
      if (lane == 0 && len < out_cap) out[len] = target;
      len++;
    }
    if (lane == 0) S.r32[0] = len;
  }
  team_sync<TEAM>();
  PROF_LAP(6);
  len = S.r32[0];
  // (e) reset the distance field on every voxel this search touched
  const uint32_t nt = S.n_touched;
  for (uint32_t i = tid; i < nt; i += nth) A.dist[touched[i]] = __int_as_float(kInfBits);
  relax = team_sum_u32<TEAM>(relax, S, Lc, T);
  if (tid == 0) { S.relax += relax; S.rounds += rounds; }
  team_sync<TEAM>();
  PROF_LAP(7);
  return len;
}

// ---- dijkstra3d.railroad ----------------------------------------------------------------------------
// Returns the path length written to out[0..): out[0] = rail voxel ... out[len-1] = target.
//
// Delta-stepping with a three-level pile of (tentative distance, voxel) PAIRS and lazy deletion:
//   proc   the batch expanded this round: every entry at most `delta` above the smallest open distance
//   mid    a bounded band (<= thr_mid) that is split every round into proc / keep
//   far    everything beyond; scanned only when the band drains
// A relaxation that improves a voxel appends a NEW pair; the pair it supersedes stays where it is and is recognised as
// stale when its turn comes (its key is no longer the voxel's distance).  Selecting a batch therefore streams 8-byte pairs
// instead of gathering one 32-byte sector per candidate and round from the distance field -- that gather was 25 of the
// 31 GB of DRAM traffic of this kernel in round 1 -- and no "is queued" flag per voxel is needed.  Relaxation order never
// changes the result (the distances are the least fixed point of d[v] = min_u fl(d[u] + w[v])).
// A list that would overflow (stale pairs pile up; live ones are at most one per voxel) raises a flag, and the pile is
// rebuilt from the list of touched voxels: every voxel whose distance is not yet final gets one fresh pair.
#ifdef B2T_HOST_EMU
extern "C" { int g_emu_rr_rebuilds = 0; }      // how often the CPU harness saw the pile rebuilt
#endif
__device__ __forceinline__ unsigned long long rr_pair(uint32_t dbits, uint32_t v) { return ((unsigned long long)dbits << 32) | v; }

template <bool TEAM>
__device__ uint32_t railroad(const Arena& A, const LabelDesc& L, uint32_t target, unsigned long long* midA,
                             unsigned long long* midB, unsigned long long* proc, unsigned long long* farA,
                             unsigned long long* farB, uint32_t* touched, uint32_t* out, uint32_t out_cap, Shared& S,
                             Local& Lc, Team& T) {
  const int lane = threadIdx.x & 31;
  const uint32_t tid = t_tid<TEAM>(T), nth = t_threads<TEAM>(), tw = t_warp<TEAM>(T), ntw = t_warps<TEAM>();
  const uint32_t seg = L.segid;
  // entries per list: twice the voxel count, so that a rebuilt pile (one pair per voxel at most) leaves room to go on
  const uint32_t cap = B2T_RR_CAP_SHIFT ? max(64u, (2u * L.n_fg) >> B2T_RR_CAP_SHIFT) : 2u * L.n_fg;
  if (__ldcg(&A.pdrf[target]) == 0.0f) {
    if (tid == 0 && out_cap > 0) out[0] = target;
    team_sync<TEAM>();
    return out_cap > 0 ? 1 : 0;
  }
  int dx = 0, dy = 0, dz = 0;
  if (lane < 26) { dx = kDX[lane]; dy = kDY[lane]; dz = kDZ[lane]; }
  const int64_t off = (int64_t)dx + (int64_t)dy * A.d.sx + (int64_t)dz * A.d.sxy;
  const uint32_t ltmask = (1u << lane) - 1u;

  unsigned long long* mid = midA;
  unsigned long long* mid2 = midB;
  unsigned long long* far = farA;
  unsigned long long* far2 = farB;
  if (tid == 0) {
    A.dist[target] = 0.0f;
    mid[0] = rr_pair(0u, target);
    touched[0] = target;
    S.n_touched = 1;
    S.best = ~0ull;
    S.n_keep = 0; S.n_proc = 0; S.n_next = 0; S.r32[1] = 0;
  }
  team_sync<TEAM>();
  PROF_START();
  uint32_t n_mid = 1, n_far = 0;
  float delta = __ldcg(&A.pdrf[target]);                      // width of the near batch
  if (!(delta > 0.0f) || __float_as_uint(delta) >= kInfBits) delta = 1.0f;
  float delta_mid = __fmul_rn(delta, 16.0f);                   // width of the mid band
  uint32_t thr_mid = __float_as_uint(delta_mid);
  uint32_t relax = 0, rounds = 0;
  const uint32_t lo_proc = (TEAM ? 2u * kWarps * kCluster : 2u * kWarps) * B2T_RR_BATCH, hi_proc = 4u * lo_proc;   // batch-size feedback
  const uint32_t lo_mid = TEAM ? 256u * kCluster : 256u, hi_mid = 4u * lo_mid;

  for (;;) {
    if (n_mid == 0) {
      // ---- refill the mid band from the far pile ----
      if (n_far == 0) break;
      const uint32_t bound = (uint32_t)(S.best >> 32);
      uint32_t mn = 0xffffffffu;
      for (uint32_t i = tid; i < n_far; i += nth) mn = min(mn, (uint32_t)(far[i] >> 32));
      mn = team_min_u32<TEAM>(mn, S, Lc, T);
      if (mn > bound) break;                                  // nothing left that could beat the rail we have
      thr_mid = __float_as_uint(__fadd_rn(__uint_as_float(mn), delta_mid));
      if (thr_mid > bound) thr_mid = bound;
      for (uint32_t i0 = 0; i0 < n_far; i0 += nth) {
        const uint32_t i = i0 + tid;
        bool tomid = false, tokeep = false;
        unsigned long long e = 0;
        if (i < n_far) {
          e = far[i];
          const uint32_t du = (uint32_t)(e >> 32);
          tomid = du <= thr_mid;
          tokeep = !tomid && du <= bound;                     // beyond the bound: dropped for good
        }
        const uint32_t mp = __ballot_sync(0xffffffffu, tomid), mk = __ballot_sync(0xffffffffu, tokeep);
        uint32_t bp = 0, bk = 0;
        if (lane == 0) {
          if (mp) bp = atomicAdd(&S.n_keep, __popc(mp));
          if (mk) bk = atomicAdd(&S.n_next, __popc(mk));
        }
        bp = __shfl_sync(0xffffffffu, bp, 0);
        bk = __shfl_sync(0xffffffffu, bk, 0);
        if (tomid) mid[bp + __popc(mp & ltmask)] = e;
        if (tokeep) far2[bk + __popc(mk & ltmask)] = e;
      }
      team_sync<TEAM>();
      n_mid = S.n_keep;
      n_far = S.n_next;
      { unsigned long long* t = far; far = far2; far2 = t; }
      if (n_mid < lo_mid) delta_mid = __fmul_rn(delta_mid, 2.0f);                  // aim at 256..1024 candidates per CTA
      else if (n_mid > hi_mid) delta_mid = __fmul_rn(delta_mid, 0.5f);
      team_sync<TEAM>();
      if (tid == 0) { S.n_keep = 0; S.n_next = 0; }
      team_sync<TEAM>();
      PROF_LAP(5); PROF_COUNT(13);
      continue;
    }
    // ---- (a) smallest tentative distance in the mid band ----
    uint32_t mn = 0xffffffffu;
    for (uint32_t i = tid; i < n_mid; i += nth) mn = min(mn, (uint32_t)(mid[i] >> 32));
    mn = team_min_u32<TEAM>(mn, S, Lc, T);
    PROF_LAP(1);
    const uint32_t bound = (uint32_t)(S.best >> 32);
    uint32_t thr = __float_as_uint(__fadd_rn(__uint_as_float(mn), delta));
    if (thr > bound) thr = bound;
    // ---- (b) split the band: <= thr expand now, <= bound keep, else drop ----
    for (uint32_t i0 = 0; i0 < n_mid; i0 += nth) {
      const uint32_t i = i0 + tid;
      bool toproc = false, tokeep = false;
      unsigned long long e = 0;
      if (i < n_mid) {
        e = mid[i];
        const uint32_t du = (uint32_t)(e >> 32);
        toproc = du <= thr;
        tokeep = !toproc && du <= bound;
      }
      const uint32_t mp = __ballot_sync(0xffffffffu, toproc), mk = __ballot_sync(0xffffffffu, tokeep);
      uint32_t bp = 0, bk = 0;
      if (lane == 0) {
        if (mp) bp = atomicAdd(&S.n_proc, __popc(mp));
        if (mk) bk = atomicAdd(&S.n_keep, __popc(mk));
      }
      bp = __shfl_sync(0xffffffffu, bp, 0);
      bk = __shfl_sync(0xffffffffu, bk, 0);
      if (toproc) proc[bp + __popc(mp & ltmask)] = e;
      if (tokeep) mid2[bk + __popc(mk & ltmask)] = e;
    }
    team_sync<TEAM>();
    PROF_LAP(2);
    const uint32_t n_proc = S.n_proc;
    // ---- (c) expand the batch: lane per neighbour, TWO voxels per warp iteration so that their chains of
    //          dependent global accesses (dist of u -> label/weight of v -> atomicMin) overlap ----
    for (uint32_t it = tw; it < n_proc; it += 2 * ntw) {
      uint32_t u[2], v[2], lv[2], nd[2], old[2], dq[2];
      float du[2], c[2];
      bool ok[2], relaxed[2], push_mid[2], push_far[2];
#pragma unroll
      for (int e = 0; e < 2; e++) {
        ok[e] = it + e * ntw < n_proc;
        const unsigned long long pe = ok[e] ? proc[it + e * ntw] : 0ull;
        u[e] = (uint32_t)pe;
        dq[e] = (uint32_t)(pe >> 32);
      }
#pragma unroll
      for (int e = 0; e < 2; e++) {
        du[e] = ok[e] ? __ldcg(&A.dist[u[e]]) : 0.0f;
        ok[e] = ok[e] && __float_as_uint(du[e]) == dq[e];        // a superseded pair: its voxel has (had) a closer one
      }
#pragma unroll
      for (int e = 0; e < 2; e++) {
        int x, y, z;
        unravel(u[e], A.d, x, y, z);
        const int nx = x + dx, ny = y + dy, nz = z + dz;
        ok[e] = ok[e] && lane < 26 && nx >= 0 && ny >= 0 && nz >= 0 && nx < A.d.sx && ny < A.d.sy && nz < A.d.sz;
        v[e] = ok[e] ? (uint32_t)((int64_t)u[e] + off) : 0u;
        lv[e] = ok[e] ? __ldg(&A.cc[v[e]]) : 0xffffffffu;       // independent loads, issued together
        c[e] = ok[e] ? __ldcg(&A.pdrf[v[e]]) : 0.0f;
      }
#pragma unroll
      for (int e = 0; e < 2; e++) {
        relaxed[e] = false; push_mid[e] = false; push_far[e] = false; nd[e] = 0; old[e] = 0;
        if (ok[e] && lv[e] == seg) {
          if (c[e] == 0.0f) {
            atomicMin(&S.best, ((unsigned long long)__float_as_uint(du[e]) << 32) | u[e]);   // rule T4 candidate
          } else {
            nd[e] = __float_as_uint(__fadd_rn(du[e], c[e]));
            if (nd[e] <= (uint32_t)(t_peek<TEAM>(&S.best) >> 32)) {
              old[e] = atomicMin(reinterpret_cast<uint32_t*>(&A.dist[v[e]]), nd[e]);
              relaxed[e] = nd[e] < old[e];
            }
          }
        }
      }
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const bool fresh = relaxed[e] && old[e] == kInfBits;            // first time this search reaches v
        const uint32_t mt = __ballot_sync(0xffffffffu, fresh);
        if (mt) {
          uint32_t bt = 0;
          if (lane == 0) bt = atomicAdd(&S.n_touched, __popc(mt));
          bt = __shfl_sync(0xffffffffu, bt, 0);
          if (fresh) touched[bt + __popc(mt & ltmask)] = v[e];
        }
        if (relaxed[e]) {
          relax++;
          push_mid[e] = nd[e] <= thr_mid;
          push_far[e] = !push_mid[e];
        }
        const uint32_t mm = __ballot_sync(0xffffffffu, push_mid[e]), mf = __ballot_sync(0xffffffffu, push_far[e]);
        if (mm | mf) {
          uint32_t bm = 0, bf = 0;
          if (lane == 0) {
            if (mm) bm = atomicAdd(&S.n_keep, __popc(mm));
            if (mf) bf = atomicAdd(&S.n_next, __popc(mf));
          }
          bm = __shfl_sync(0xffffffffu, bm, 0);
          bf = __shfl_sync(0xffffffffu, bf, 0);
          const uint32_t pm = bm + __popc(mm & ltmask), pf = n_far + bf + __popc(mf & ltmask);
          if (push_mid[e]) { if (pm < cap) mid2[pm] = rr_pair(nd[e], v[e]); else S.r32[1] = 1u; }
          if (push_far[e]) { if (pf < cap) far[pf] = rr_pair(nd[e], v[e]); else S.r32[1] = 1u; }
        }
      }
    }
    team_sync<TEAM>();
    PROF_LAP(3);
    n_mid = S.n_keep;
    n_far += S.n_next;
    const bool overflow = S.r32[1] != 0u;
    { unsigned long long* t = mid; mid = mid2; mid2 = t; }
    // batch-size feedback: keep roughly 2..8 voxels per warp in flight
    if (n_proc < lo_proc) delta = __fmul_rn(delta, 2.0f);
    else if (n_proc > hi_proc) delta = __fmul_rn(delta, 0.5f);
    rounds++;
    team_sync<TEAM>();
    if (tid == 0) { S.n_keep = 0; S.n_proc = 0; S.n_next = 0; S.r32[1] = 0; }
    team_sync<TEAM>();
    PROF_LAP(4);
    if (overflow) {
      // a list filled up with superseded pairs (and lost some new ones): rebuild the pile from the voxels this search has
      // touched -- one fresh pair for every voxel that is not final yet (everything below this round's smallest open
      // distance `mn` is; expanding a final voxel again would change nothing anyway)
#ifdef B2T_HOST_EMU
      if (tid == 0) g_emu_rr_rebuilds++;
#endif
      const uint32_t nt = S.n_touched;
      for (uint32_t i0 = 0; i0 < nt; i0 += nth) {
        const uint32_t i = i0 + tid;
        bool keep = false;
        uint32_t v = 0, dv = 0;
        if (i < nt) {
          v = touched[i];
          dv = __float_as_uint(__ldcg(&A.dist[v]));
          keep = dv >= mn && dv <= (uint32_t)(S.best >> 32);
        }
        const uint32_t mk = __ballot_sync(0xffffffffu, keep);
        uint32_t bk = 0;
        if (lane == 0 && mk) bk = atomicAdd(&S.n_next, __popc(mk));
        bk = __shfl_sync(0xffffffffu, bk, 0);
        if (keep) far[bk + __popc(mk & ltmask)] = rr_pair(dv, v);
      }
      team_sync<TEAM>();
      n_mid = 0;
      n_far = S.n_next;
      team_sync<TEAM>();
      if (tid == 0) S.n_next = 0;
      team_sync<TEAM>();
    }
  }

  return railroad_finish<TEAM>(A, L, target, touched, out, out_cap, relax, rounds, S, Lc, T);
}

// ---- railroad for a solo CTA: the near lists live in SHARED memory ------------------------------------------------
// Same search, same result (the least fixed point does not depend on the order of the relaxations); what changes is
// where a round's time goes.  A label's path loop is a chain of ~100 rounds per path and the profile of the global-
// memory form (profiles/r02_trace_prof_before.jsonl) shows ~20 us per round, spent on dependent round trips to L2:
// the band is streamed twice (min, split), the batch is read back, then dist[u] -> cc/pdrf[v] -> atomicMin.  Here
//   * the mid band (both halves) and the batch are shared-memory arrays; entries that do not fit go to the far pile,
//   * the smallest key of the next band is tracked while the band is written (split and push), so no min pass,
//   * the counters of a round come in two sets used alternately, so nothing has to be reset between two barriers,
//   * dist[u] is loaded together with the neighbours' label / weight instead of before them,
// which leaves two block barriers and two global round trips (loads, atomicMin) per round.
#ifndef B2T_RR_SOLO_CAP
#define B2T_RR_SOLO_CAP 2048           // entries per shared-memory list (three lists of 8-byte pairs: 48 KB per CTA)
#endif
#ifndef B2T_RR_SOLO
#define B2T_RR_SOLO 1                  // solo CTAs use railroad_solo (0: the global-memory form everywhere)
#endif
// railroad_solo's neighbour gather of the PDRF reads through L1 (B2T_PDRF_CA=1, the default): a label's PDRF only changes
// between two searches and only by stores of the CTA that owns the label (the rail zeroing), which write through and keep
// that SM's L1 line current; a sector may also hold voxels of OTHER labels whose owners zero them on other SMs, but a
// neighbour whose cc is not this label's is never used.  Measured: path loop 27.15 -> 26.19 ms, digests identical.
#ifndef B2T_PDRF_CA
#define B2T_PDRF_CA 1
#endif
#if B2T_PDRF_CA && !defined(B2T_HOST_EMU)
#define B2T_PDRF_LD(p_) __ldca(p_)
#else
#define B2T_PDRF_LD(p_) __ldcg(p_)
#endif
constexpr uint32_t kSoloCap = B2T_RR_SOLO_CAP;
struct RrLists {
  unsigned long long mid[2][kSoloCap];
  unsigned long long proc[kSoloCap];
};

__device__ uint32_t railroad_solo(const Arena& A, const LabelDesc& L, uint32_t target, RrLists& R, unsigned long long* farA,
                                  unsigned long long* farB, uint32_t* touched, uint32_t* out, uint32_t out_cap, Shared& S,
                                  Local& Lc, Team& T) {
  const int lane = threadIdx.x & 31;
  const uint32_t tid = threadIdx.x, nth = kThreads;
  const uint32_t seg = L.segid;
  const uint32_t cap = B2T_RR_CAP_SHIFT ? max(64u, (2u * L.n_fg) >> B2T_RR_CAP_SHIFT) : 2u * L.n_fg;   // far lists (global)
#ifndef B2T_RR_SCAP_SHIFT
#define B2T_RR_SCAP_SHIFT B2T_RR_CAP_SHIFT
#endif
  const uint32_t scap = B2T_RR_SCAP_SHIFT ? max(64u, kSoloCap >> B2T_RR_SCAP_SHIFT) : kSoloCap;         // shared lists
  if (__ldcg(&A.pdrf[target]) == 0.0f) {
    if (tid == 0 && out_cap > 0) out[0] = target;
    __syncthreads();
    return out_cap > 0 ? 1 : 0;
  }
  const uint32_t ltmask = (1u << lane) - 1u;

  unsigned long long* far = farA;
  unsigned long long* far2 = farB;
  int cur = 0;                                   // R.mid[cur] is the band of this round, R.mid[cur ^ 1] the next one
  uint32_t par = 0;                              // S.rc[par]: the counters of this round
  if (tid == 0) {
    A.dist[target] = 0.0f;
    R.mid[0][0] = rr_pair(0u, target);
    touched[0] = target;
    S.n_touched = 1;
    S.best = ~0ull;
    S.rc[0] = RoundCounters{0u, 0u, 0u, 0xffffffffu, 0xffffffffu, 0u};
    S.rc[1] = RoundCounters{0u, 0u, 0u, 0xffffffffu, 0xffffffffu, 0u};
  }
  __syncthreads();
  PROF_START();
  uint32_t n_mid = 1, n_far = 0, mn = 0u;        // mn: smallest key in the band
  uint32_t far_low = 0xffffffffu;                // a lower bound of every key in the far pile
  float delta = __ldcg(&A.pdrf[target]);
  if (!(delta > 0.0f) || __float_as_uint(delta) >= kInfBits) delta = 1.0f;
  float delta_mid = __fmul_rn(delta, 16.0f);
  uint32_t thr_mid = __float_as_uint(delta_mid);
  uint32_t relax = 0, rounds = 0;
  const uint32_t lo_proc = 32u * B2T_RR_BATCH, hi_proc = 4u * lo_proc;    // voxels per round: up to 170 expand in one pass of the CTA
  const uint32_t lo_mid = 256u, hi_mid = 1024u;

  for (;;) {
    RoundCounters& C = S.rc[par];
    if (n_mid == 0) {
      // ---- refill the band from the far pile (global memory; rare) ----
      if (n_far == 0) break;
      const uint32_t bound = (uint32_t)(S.best >> 32);
      uint32_t fm = 0xffffffffu;
      for (uint32_t i = tid; i < n_far; i += nth) fm = min(fm, (uint32_t)(far[i] >> 32));
      fm = block_min_u32(fm, Lc.red32);
      if (fm > bound) break;                                  // nothing left that could beat the rail we have
      thr_mid = __float_as_uint(__fadd_rn(__uint_as_float(fm), delta_mid));
      if (thr_mid > bound) thr_mid = bound;
      for (uint32_t i0 = 0; i0 < n_far; i0 += nth) {
        const uint32_t i = i0 + tid;
        bool tomid = false, tokeep = false;
        unsigned long long e = 0;
        if (i < n_far) {
          e = far[i];
          const uint32_t du = (uint32_t)(e >> 32);
          tomid = du <= thr_mid;
          tokeep = !tomid && du <= bound;                     // beyond the bound: dropped for good
        }
        const uint32_t mp = __ballot_sync(0xffffffffu, tomid);
        uint32_t bp = 0;
        if (lane == 0 && mp) bp = atomicAdd(&C.n_keep, __popc(mp));
        bp = __shfl_sync(0xffffffffu, bp, 0);
        if (tomid) {
          const uint32_t pm = bp + __popc(mp & ltmask);
          if (pm < scap) R.mid[cur][pm] = e; else { tomid = false; tokeep = true; }   // band full: stays in the pile
        }
        const uint32_t mk = __ballot_sync(0xffffffffu, tokeep);
        uint32_t bk = 0;
        if (lane == 0 && mk) bk = atomicAdd(&C.n_next, __popc(mk));
        bk = __shfl_sync(0xffffffffu, bk, 0);
        if (tokeep) far2[bk + __popc(mk & ltmask)] = e;
      }
      __syncthreads();
      n_mid = min(C.n_keep, scap);
      n_far = C.n_next;
      mn = fm;                                                // the smallest key moved with the others
      far_low = fm;
      { unsigned long long* t = far; far = far2; far2 = t; }
      if (n_mid < lo_mid) delta_mid = __fmul_rn(delta_mid, 2.0f);
      else if (n_mid > hi_mid) delta_mid = __fmul_rn(delta_mid, 0.5f);
      __syncthreads();
      if (tid == 0) { C.n_keep = 0; C.n_next = 0; }
      __syncthreads();
      PROF_LAP(5); PROF_COUNT(13);
      continue;
    }
    const uint32_t bound = (uint32_t)(S.best >> 32);
    uint32_t thr = __float_as_uint(__fadd_rn(__uint_as_float(mn), delta));
    if (thr > bound) thr = bound;
    // ---- (b) split the band: <= thr expand now, <= thr_mid keep, <= bound to the far pile, else drop ----
    uint32_t pmin = 0xffffffffu, fmin = 0xffffffffu;
    {
      const unsigned long long* midc = R.mid[cur];
      unsigned long long* midn = R.mid[cur ^ 1];
      const bool can_shed = n_far + n_mid <= cap;              // shedding never overflows the pile: only a relaxation may
      for (uint32_t i0 = 0; i0 < n_mid; i0 += nth) {
        const uint32_t i = i0 + tid;
        bool toproc = false, tokeep = false, tofar = false;
        unsigned long long e = 0;
        if (i < n_mid) {
          e = midc[i];
          const uint32_t du = (uint32_t)(e >> 32);
          toproc = du <= thr;
          tokeep = !toproc && (du <= thr_mid || !can_shed) && du <= bound;
          tofar = !toproc && !tokeep && du <= bound;            // the band was narrowed after this pair came in
          if (tokeep) pmin = min(pmin, du);
        }
        const uint32_t mp = __ballot_sync(0xffffffffu, toproc), mk = __ballot_sync(0xffffffffu, tokeep);
        uint32_t bp = 0, bk = 0;
        if (lane == 0) {
          if (mp) bp = atomicAdd(&C.n_proc, __popc(mp));
          if (mk) bk = atomicAdd(&C.n_keep, __popc(mk));
        }
        bp = __shfl_sync(0xffffffffu, bp, 0);
        bk = __shfl_sync(0xffffffffu, bk, 0);
        if (toproc) R.proc[bp + __popc(mp & ltmask)] = e;      // at most n_mid <= scap entries
        if (tokeep) midn[bk + __popc(mk & ltmask)] = e;        // likewise
        if (tofar) {
          far[n_far + atomicAdd(&C.n_next, 1u)] = e;
          fmin = min(fmin, (uint32_t)(e >> 32));
        }
      }
    }
    __syncthreads();
    PROF_LAP(2);
    if (tid == 0) S.rc[par ^ 1] = RoundCounters{0u, 0u, 0u, 0xffffffffu, 0xffffffffu, 0u};   // the next round's set: idle until the barrier below
    const uint32_t n_proc = C.n_proc;
    // ---- (c) expand the batch: one thread per (voxel, z-plane of its neighbourhood), nine neighbours each.  The order of
    //          the relaxations does not matter (least fixed point; rule T4 is an atomicMin), so the neighbours are taken in
    //          memory order: three rows of three per thread, all loads in flight before the first atomic ----
    const uint32_t bound_now = (uint32_t)(S.best >> 32);
    for (uint32_t task = tid; task < 3u * n_proc; task += nth) {
      const uint32_t i = task / 3u, q = task - 3u * i;
      const unsigned long long pe = R.proc[i];
      const uint32_t u = (uint32_t)pe, dq = (uint32_t)(pe >> 32);
      int x, y, z;
      unravel(u, A.d, x, y, z);
      const int nz = z + (int)q - 1;
      const float du = __ldcg(&A.dist[u]);
      const bool planeok = nz >= 0 && nz < A.d.sz;
      const int64_t base = (int64_t)u + ((int64_t)q - 1) * (int64_t)A.d.sxy;
      uint32_t lv[9];
      float c[9];
#pragma unroll
      for (int j = 0; j < 9; j++) {
        const int ddx = j % 3 - 1, ddy = j / 3 - 1;
        const int nx = x + ddx, ny = y + ddy;
        const bool ok = planeok && nx >= 0 && nx < A.d.sx && ny >= 0 && ny < A.d.sy && !(j == 4 && q == 1u);
        const uint32_t v = (uint32_t)(base + (int64_t)ddy * A.d.sx + ddx);
        lv[j] = ok ? __ldg(&A.cc[v]) : seg + 1u;                          // never equal to seg
        c[j] = ok ? B2T_PDRF_LD(&A.pdrf[v]) : 0.0f;
      }
      if (__float_as_uint(du) != dq) continue;                            // a superseded pair: its voxel has (had) a closer one
      uint32_t nd[9], old[9];
      bool rail = false;
#pragma unroll
      for (int j = 0; j < 9; j++) {
        nd[j] = 0u; old[j] = 0u;
        if (lv[j] == seg) {
          if (c[j] == 0.0f) rail = true;
          else {
            nd[j] = __float_as_uint(__fadd_rn(du, c[j]));
            if (nd[j] <= bound_now) {
              const int ddx = j % 3 - 1, ddy = j / 3 - 1;
              const uint32_t v = (uint32_t)(base + (int64_t)ddy * A.d.sx + ddx);
              old[j] = atomicMin(reinterpret_cast<uint32_t*>(&A.dist[v]), nd[j]);
            }
          }
        }
      }
      if (rail) atomicMin(&S.best, ((unsigned long long)dq << 32) | u);  // rule T4 candidate
#pragma unroll
      for (int j = 0; j < 9; j++) {
        if (nd[j] < old[j]) {                                             // relaxed (old stays 0 where nothing was tried)
          const int ddx = j % 3 - 1, ddy = j / 3 - 1;
          const uint32_t v = (uint32_t)(base + (int64_t)ddy * A.d.sx + ddx);
          relax++;
          if (old[j] == kInfBits) touched[atomicAdd(&S.n_touched, 1u)] = v;   // first time this search reaches v
          bool tofar = nd[j] > thr_mid;
          if (!tofar) {
            const uint32_t pm = atomicAdd(&C.n_keep, 1u);
            if (pm < scap) { R.mid[cur ^ 1][pm] = rr_pair(nd[j], v); pmin = min(pmin, nd[j]); }
            else tofar = true;                                            // band full: to the pile
          }
          if (tofar) {
            const uint32_t pf = n_far + atomicAdd(&C.n_next, 1u);
            if (pf < cap) { far[pf] = rr_pair(nd[j], v); fmin = min(fmin, nd[j]); } else C.overflow = 1u;
          }
        }
      }
    }
    if (pmin != 0xffffffffu) atomicMin(&C.min_next, pmin);
    if (fmin != 0xffffffffu) atomicMin(&C.min_far, fmin);
    __syncthreads();
    PROF_LAP(3);
    n_mid = min(C.n_keep, scap);
    n_far += C.n_next;                                        // exact unless a push was lost (then the pile is rebuilt below)
    const uint32_t mn_round = mn;                             // every open key of this round was >= min(mn_round, far_low)
    const uint32_t low = min(mn_round, far_low);
    mn = C.min_next;
    far_low = min(far_low, C.min_far);
    const bool overflow = C.overflow != 0u;
    cur ^= 1;
    par ^= 1u;
    if (n_proc < lo_proc) delta = __fmul_rn(delta, 2.0f);
    else if (n_proc > hi_proc) delta = __fmul_rn(delta, 0.5f);
    if (n_mid > hi_mid || C.n_keep > scap) {                  // the band outgrew its list: narrow it, the next split sheds the rest
      delta_mid = __fmul_rn(delta_mid, 0.5f);
      const uint32_t t = __float_as_uint(__fadd_rn(__uint_as_float(mn), delta_mid));
      if (t < thr_mid) thr_mid = t;
    }
    rounds++;
    if (overflow) {
      // the far pile filled up with superseded pairs (and lost some new ones): rebuild it from the voxels this search has
      // touched -- one fresh pair for every voxel that may not be final yet; the band is dropped, its live pairs come back
      // with the others
#ifdef B2T_HOST_EMU
      if (tid == 0) g_emu_rr_rebuilds++;
#endif
      RoundCounters& C2 = S.rc[par];                          // reset after the split barrier of the round that just ended
      const uint32_t nt = S.n_touched;                        // below `low`: expanded with its final distance already
      for (uint32_t i0 = 0; i0 < nt; i0 += nth) {
        const uint32_t i = i0 + tid;
        bool keep = false;
        uint32_t v = 0, dv = 0;
        if (i < nt) {
          v = touched[i];
          dv = __float_as_uint(__ldcg(&A.dist[v]));
          keep = dv >= low && dv <= (uint32_t)(S.best >> 32);
        }
        const uint32_t mk = __ballot_sync(0xffffffffu, keep);
        uint32_t bk = 0;
        if (lane == 0 && mk) bk = atomicAdd(&C2.n_next, __popc(mk));
        bk = __shfl_sync(0xffffffffu, bk, 0);
        if (keep) far[bk + __popc(mk & ltmask)] = rr_pair(dv, v);
      }
      __syncthreads();
      n_mid = 0;
      n_far = C2.n_next;
      far_low = low;
      __syncthreads();
      if (tid == 0) { C2.n_next = 0; }
      __syncthreads();
    }
    PROF_LAP(4);
  }
  return railroad_finish<false>(A, L, target, touched, out, out_cap, relax, rounds, S, Lc, T);
}

// ---- roll_invalidation_ball_inside_component ---------------------------------------------------------
// seeds[0..n_seeds) are path voxels in path order; radius_i = fl32(fl32(scale * DBF[seed]) + const)
// (skeletontricks.pyx:393-395 under NumPy-2 scalar rules).  Returns the number of voxels invalidated.
template <bool TEAM>
__device__ uint32_t invalidate(const Arena& A, const LabelDesc& L, const uint32_t* seeds, uint32_t n_seeds, float scale,
                               float konst, uint32_t* fvA, uint32_t* fsA, uint32_t* fvB, uint32_t* fsB, Shared& S,
                               Team& T) {
  const int lane = threadIdx.x & 31;
  const uint32_t tid = t_tid<TEAM>(T), nth = t_threads<TEAM>(), tw = t_warp<TEAM>(T), ntw = t_warps<TEAM>();
  const uint32_t seg = L.segid;
  int dx = 0, dy = 0, dz = 0;
  if (lane < 26) { dx = kDX[lane]; dy = kDY[lane]; dz = kDZ[lane]; }
  const int64_t off = (int64_t)dx + (int64_t)dy * A.d.sx + (int64_t)dz * A.d.sxy;
  const uint32_t ltmask = (1u << lane) - 1u;
  if (tid == 0) { S.n_next = 0; }
  team_sync<TEAM>();
  // round 0: every still-valid seed claims itself
  for (uint32_t i0 = 0; i0 < n_seeds; i0 += nth) {
    const uint32_t i = i0 + tid;
    bool won = false;
    uint32_t v = 0;
    if (i < n_seeds) {
      v = seeds[i];
      if (__ldg(&A.cc[v]) == seg) won = atomicCAS(&A.claim[v], kValid, 0ull) == kValid;
    }
    const uint32_t m = __ballot_sync(0xffffffffu, won);
    uint32_t base = 0;
    if (lane == 0 && m) base = atomicAdd(&S.n_next, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (won) { const uint32_t p = base + __popc(m & ltmask); fvA[p] = v; fsA[p] = i; }
  }
  team_sync<TEAM>();
  uint32_t n_cur = S.n_next, total = n_cur;
  uint32_t *fv = fvA, *fs = fsA, *nv = fvB, *ns = fsB;
  while (n_cur > 0) {
    team_sync<TEAM>();
    if (tid == 0) S.n_next = 0;
    team_sync<TEAM>();
    for (uint32_t it = tw; it < n_cur; it += ntw) {
      const uint32_t u = fv[it], s = fs[it];
      const uint32_t o = seeds[s];
      const float r = __fadd_rn(__fmul_rn(scale, __ldg(&A.dbf[o])), konst);
      int x, y, z, ox, oy, oz;
      unravel(u, A.d, x, y, z);
      unravel(o, A.d, ox, oy, oz);
      const int nx = x + dx, ny = y + dy, nz = z + dz;
      bool push = false;
      uint32_t v = 0;
      if (lane < 26 && nx >= 0 && ny >= 0 && nz >= 0 && nx < A.d.sx && ny < A.d.sy && nz < A.d.sz) {
        v = (uint32_t)((int64_t)u + off);
        const uint32_t lv = __ldg(&A.cc[v]);
        const unsigned long long cl = __ldcg(&A.claim[v]);
        if (lv == seg && cl != 0ull) {
          // float32 expression of dijkstra_invalidation.hpp:49-52,319-323
          const float a = __fmul_rn(A.wx, (float)(nx - ox)), b = __fmul_rn(A.wy, (float)(ny - oy)),
                      c = __fmul_rn(A.wz, (float)(nz - oz));
          const float dd = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c)));
          if (dd < r) {
            const unsigned long long cand = ((unsigned long long)__float_as_uint(dd) << 32) | s;
            const unsigned long long old = atomicMin(&A.claim[v], cand);
            push = old == kValid;
          }
        }
      }
      const uint32_t m = __ballot_sync(0xffffffffu, push);
      if (m) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&S.n_next, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (push) nv[base + __popc(m & ltmask)] = v;
      }
    }
    team_sync<TEAM>();
    const uint32_t n_next = S.n_next;
    // end of round: winners become claimed (0) and carry their owner into the next frontier
    for (uint32_t i = tid; i < n_next; i += nth) {
      const uint32_t v = nv[i];
      const unsigned long long c = __ldcg(&A.claim[v]);
      ns[i] = (uint32_t)c;
      A.claim[v] = 0ull;
    }
    total += n_next;
    n_cur = n_next;
    { uint32_t* t = fv; fv = nv; nv = t; t = fs; fs = ns; ns = t; }
    team_sync<TEAM>();
  }
  return total;
}

// ---- the same, ordered by KEY instead of by hop count (the default: b2t_set_invalidation_mode) -----------
// The reference pops its heap in order of ||w.(v - seed)|| (dijkstra_invalidation.hpp:233-237); the hop-synchronous
// rounds above hand a voxel to whichever seed reaches it in the fewest steps.  Where balls of different radii overlap
// that changes owners, and with them how far the claim spreads: on the CPU (oracle/oracle.c: orc_invalidate_heap is the
// compiled reference voxel for voxel, orc_invalidate_window is this function) hop rounds give 126 of 176 skeletons
// identical to the reference, key rounds of one voxel's width 174 of 176 at the same number of rounds.
// Every valid voxel next to a claimed one holds its best candidate (key << 32 | seed order) in A.claim; a round claims
// all open candidates whose key is the smallest one or below (smallest + delta), then the new owners push candidates
// to their neighbours with atomicMin -- a min-reduction, so the order inside a round does not matter.
// act / act2: open candidates (ping-pong), fv / fs: the voxels claimed in this round and their seeds.
template <bool TEAM>
__device__ __noinline__ uint32_t invalidate_window(const Arena& A, const LabelDesc& L, const uint32_t* seeds, uint32_t n_seeds,
                                                   float scale, float konst, float delta, uint32_t* act, uint32_t* fv,
                                                   uint32_t* fs, uint32_t* act2, Shared& S, Local& Lc, Team& T) {
  const int lane = threadIdx.x & 31;
  const uint32_t tid = t_tid<TEAM>(T), nth = t_threads<TEAM>();
  const uint32_t seg = L.segid;
  const uint32_t ltmask = (1u << lane) - 1u;
  if (tid == 0) { S.n_next = 0; S.n_keep = 0; }
  team_sync<TEAM>();
  PROF_START();
  // round 0: every still-valid seed claims itself (key 0)
  for (uint32_t i0 = 0; i0 < n_seeds; i0 += nth) {
    const uint32_t i = i0 + tid;
    bool won = false;
    uint32_t v = 0;
    if (i < n_seeds) {
      v = seeds[i];
      if (__ldg(&A.cc[v]) == seg) won = atomicCAS(&A.claim[v], kValid, 0ull) == kValid;
    }
    const uint32_t m = __ballot_sync(0xffffffffu, won);
    uint32_t base = 0;
    if (lane == 0 && m) base = atomicAdd(&S.n_next, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (won) { const uint32_t p = base + __popc(m & ltmask); fv[p] = v; fs[p] = i; }
  }
  team_sync<TEAM>();
  uint32_t n_cur = S.n_next, n_act = 0, total = n_cur;
  while (n_cur > 0) {
    // the voxels claimed in the last round push candidates; a voxel that gets its first one joins the open list.
    // One thread per (voxel, z-plane of its neighbourhood): the winner of a voxel is the smallest (distance, seed) pair
    // whatever the order of the pushes, so the nine neighbours are taken in memory order with all loads in flight together.
    for (uint32_t task = tid; task < 3u * n_cur; task += nth) {
      const uint32_t i = task / 3u, q = task - 3u * i;
      const uint32_t u = fv[i], s = fs[i];
      const uint32_t o = seeds[s];
      const float r = __fadd_rn(__fmul_rn(scale, __ldg(&A.dbf[o])), konst);
      int x, y, z, ox, oy, oz;
      unravel(u, A.d, x, y, z);
      unravel(o, A.d, ox, oy, oz);
      const int nz = z + (int)q - 1;
      if (nz < 0 || nz >= A.d.sz) continue;
      const int64_t base = (int64_t)u + ((int64_t)q - 1) * (int64_t)A.d.sxy;
      uint32_t lv[9];
      unsigned long long cl[9];
#pragma unroll
      for (int j = 0; j < 9; j++) {
        const int ddx = j % 3 - 1, ddy = j / 3 - 1;
        const int nx = x + ddx, ny = y + ddy;
        const bool ok = nx >= 0 && nx < A.d.sx && ny >= 0 && ny < A.d.sy && !(j == 4 && q == 1u);
        const uint32_t v = (uint32_t)(base + (int64_t)ddy * A.d.sx + ddx);
        lv[j] = ok ? __ldg(&A.cc[v]) : seg + 1u;                          // never equal to seg
        cl[j] = ok ? __ldcg(&A.claim[v]) : 0ull;
      }
      const float c = __fmul_rn(A.wz, (float)(nz - oz));
      const float cc2 = __fmul_rn(c, c);
#pragma unroll
      for (int j = 0; j < 9; j++) {
        if (lv[j] == seg && cl[j] != 0ull) {
          const int ddx = j % 3 - 1, ddy = j / 3 - 1;
          const float a = __fmul_rn(A.wx, (float)(x + ddx - ox)), b = __fmul_rn(A.wy, (float)(y + ddy - oy));
          const float dd = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), cc2));
          if (dd < r) {
            const uint32_t v = (uint32_t)(base + (int64_t)ddy * A.d.sx + ddx);
            const unsigned long long cand = ((unsigned long long)__float_as_uint(dd) << 32) | s;
            if (cand < cl[j]) {                                           // cannot win otherwise: claims only ever decrease
              const unsigned long long old = atomicMin(&A.claim[v], cand);
              if (old == kValid) act[n_act + atomicAdd(&S.n_keep, 1u)] = v;
            }
          }
        }
      }
    }
    team_sync<TEAM>();
    n_act += S.n_keep;
    team_sync<TEAM>();
    if (tid == 0) { S.n_next = 0; S.n_keep = 0; S.n_proc = 0; }
    PROF_LAP(8);
    if (n_act == 0) break;
    // smallest open key (non-negative floats order like their bit patterns)
    uint32_t kmin = 0xffffffffu;
    for (uint32_t i = tid; i < n_act; i += nth) kmin = min(kmin, (uint32_t)(__ldcg(&A.claim[act[i]]) >> 32));
    kmin = team_min_u32<TEAM>(kmin, S, Lc, T);     // also orders the counter reset above before the appends below
    PROF_LAP(9);
    const float lim = __fadd_rn(__uint_as_float(kmin), delta);
    // claim what is inside the window, keep the rest open
    for (uint32_t i0 = 0; i0 < n_act; i0 += nth) {
      const uint32_t i = i0 + tid;
      bool take = false, keep = false;
      uint32_t v = 0, owner = 0;
      if (i < n_act) {
        v = act[i];
        const unsigned long long c = __ldcg(&A.claim[v]);
        const uint32_t kb = (uint32_t)(c >> 32);
        take = kb == kmin || __uint_as_float(kb) < lim;
        keep = !take;
        owner = (uint32_t)c;
        if (take) A.claim[v] = 0ull;
      }
      const uint32_t mt = __ballot_sync(0xffffffffu, take), mk = __ballot_sync(0xffffffffu, keep);
      uint32_t bt = 0, bk = 0;
      if (lane == 0 && mt) bt = atomicAdd(&S.n_next, __popc(mt));
      if (lane == 0 && mk) bk = atomicAdd(&S.n_proc, __popc(mk));
      bt = __shfl_sync(0xffffffffu, bt, 0);
      bk = __shfl_sync(0xffffffffu, bk, 0);
      if (take) { const uint32_t p = bt + __popc(mt & ltmask); fv[p] = v; fs[p] = owner; }
      if (keep) act2[bk + __popc(mk & ltmask)] = v;
    }
    team_sync<TEAM>();
    n_cur = S.n_next;
    n_act = S.n_proc;
    total += n_cur;
    { uint32_t* t = act; act = act2; act2 = t; }
    team_sync<TEAM>();
    if (tid == 0) { S.n_keep = 0; }
    team_sync<TEAM>();
    PROF_LAP(10); PROF_COUNT(12);
  }
  team_sync<TEAM>();
  return total;
}

// ---- the same, LITERALLY: std::priority_queue<HeapDistanceNode, vector, HeapDistanceNodeCompare> on one warp ----------
// dijkstra_invalidation.hpp:233-237, 291-329.  The reference's comparator `t1.dist >= t2.dist` is not a strict weak
// order, so which of two entries with equal keys is popped first is whatever libstdc++'s push_heap / pop_heap
// (bits/stl_heap.h: __push_heap, __adjust_heap) make of the array -- and with it which seed owns a voxel that two seeds
// reach at the same distance, whose radius then governs how far the claim spreads.  Only the literal process reproduces
// that, so this mode keeps the very array: three u32 planes (key = distance bits, voxel, seed order) in the label's heap
// region.  It is sequential per label by nature; the warp is used for what is independent inside ONE heap operation:
//   push   the ancestors of slot n are known in advance: every lane loads one, a ballot finds how far the new entry
//          rises, the lanes shift their entries down together                       (one round trip instead of log n)
//   pop    __adjust_heap walks the hole to the bottom along the smaller children: the 30 descendants of the next four
//          levels are loaded at once and the four choices made by shuffles         (log n / 4 round trips); the entries
//          on that path then move up together and the displaced last entry drops in where __push_heap would leave it
//   visit  lanes = the 26 neighbours in the reference's order, with its aliasing at the x faces (a corner entry is gated
//          on y and z only, hpp:116-123: at x = 0 / sx-1 it repeats the yz diagonal and that voxel is pushed twice)
// Non-negative floats order like their bit patterns, so keys are compared as u32.
struct WarpHeap {
  uint32_t* k;
  uint32_t* v;
  uint32_t* s;
  uint32_t n, cap;
};

__device__ __forceinline__ void heap_push(WarpHeap& H, uint32_t key, uint32_t vox, uint32_t seed, int lane) {
  const uint32_t idx = H.n;
  H.n = idx + 1;
  // lane j: the ancestor j+1 levels above slot idx (1-based heap numbering: (idx+1) >> (j+1))
  const uint32_t a1 = lane < 31 ? ((idx + 1u) >> (lane + 1)) : 0u;
  const bool have = a1 != 0u;
  const uint32_t anc = a1 - 1u;
  const uint32_t ka = have ? H.k[anc] : 0u;
  // __push_heap: while (hole > top && comp(parent, value)) -- comp(parent, value) = parent.dist >= value.dist
  const uint32_t m = __ballot_sync(0xffffffffu, have && ka >= key);
  const int t = __ffs((int)~m) - 1;                  // entries that move down: lanes 0 .. t-1
  const uint32_t up = __shfl_up_sync(0xffffffffu, anc, 1);
  const uint32_t hole = __shfl_sync(0xffffffffu, anc, t > 0 ? t - 1 : 0);
  if (t > 0) {
    uint32_t pv = 0, ps = 0;
    if (lane < t) { pv = H.v[anc]; ps = H.s[anc]; }
    __syncwarp();
    if (lane < t) {
      const uint32_t dest = lane == 0 ? idx : up;
      H.k[dest] = ka; H.v[dest] = pv; H.s[dest] = ps;
    }
  }
  if (lane == 0) {
    const uint32_t at = t > 0 ? hole : idx;
    H.k[at] = key; H.v[at] = vox; H.s[at] = seed;
  }
  __syncwarp();
}

// queue.top() + queue.pop(): returns the entry in (tk, tv, ts)
__device__ __forceinline__ void heap_pop(WarpHeap& H, uint32_t& tk, uint32_t& tv, uint32_t& ts, int lane) {
  tk = H.k[0]; tv = H.v[0]; ts = H.s[0];
  const uint32_t len = H.n - 1u;                     // std::pop_heap: the last entry is re-inserted into [0, len)
  H.n = len;
  if (len == 0u) return;
  const uint32_t vk = H.k[len], vv = H.v[len], vs = H.s[len];
  __syncwarp();
  uint32_t c = 0;                                    // the hole
  int depth = 0;
  uint32_t myidx = 0, mykey = 0;                     // lane d (>= 1): the entry on the hole's way down at depth d
  const uint32_t lim = (len - 1u) / 2u;              // __adjust_heap: while (second < (len - 1) / 2)
  // lane l < 30: descendant of the hole k levels down (k = 1..4), o-th of its level
  const int kk = lane < 2 ? 1 : (lane < 6 ? 2 : (lane < 14 ? 3 : 4));
  const uint32_t oo = (uint32_t)lane + 2u - (1u << kk);
  while (c < lim) {
    const unsigned long long di = (((unsigned long long)c + 1ull) << kk) - 1ull + oo;
    const uint32_t key = (lane < 30 && di < (unsigned long long)len) ? H.k[di] : 0xffffffffu;
    uint32_t r = 0;
#pragma unroll
    for (int k = 1; k <= 4; k++) {
      if (c < lim) {                                 // uniform: both children exist
        const int base = (1 << k) - 2;
        const uint32_t kl = __shfl_sync(0xffffffffu, key, base + 2 * (int)r);
        const uint32_t kr = __shfl_sync(0xffffffffu, key, base + 2 * (int)r + 1);
        const bool left = kr >= kl;                  // comp(first + second, first + (second - 1)): take the left one
        r = 2u * r + (left ? 0u : 1u);
        c = 2u * c + (left ? 1u : 2u);
        depth++;
        if (lane == depth) { myidx = c; mykey = left ? kl : kr; }
      }
    }
  }
  if ((len & 1u) == 0u && c == (len - 2u) / 2u) {    // a last parent with a left child only
    c = 2u * c + 1u;
    depth++;
    if (lane == depth) { myidx = c; mykey = H.k[c]; }
  }
  // __push_heap(first, hole, top = 0, value) along the same path: the entries at depth 1 .. j move up one level, where
  // j is the deepest one whose key is strictly smaller than the value's; the value lands at depth j
  const uint32_t lt = __ballot_sync(0xffffffffu, lane >= 1 && lane <= depth && mykey < vk);
  const int j = lt ? 31 - __clz((int)lt) : 0;
  const uint32_t up = __shfl_up_sync(0xffffffffu, myidx, 1);
  const uint32_t hole = __shfl_sync(0xffffffffu, myidx, j);
  uint32_t pv = 0, ps = 0;
  const bool mover = lane >= 1 && lane <= j;
  if (mover) { pv = H.v[myidx]; ps = H.s[myidx]; }
  __syncwarp();
  if (mover) { H.k[up] = mykey; H.v[up] = pv; H.s[up] = ps; }
  if (lane == 0) { H.k[hole] = vk; H.v[hole] = vv; H.s[hole] = vs; }
  __syncwarp();
}

// Returns the number of voxels invalidated; S.n_proc = 1 when the heap outgrew both its region and the spill arena.
template <bool TEAM>
__device__ __noinline__ uint32_t invalidate_strict(const Arena& A, const LabelDesc& L, const Pools& P, uint32_t job,
                                                   const uint32_t* seeds, uint32_t n_seeds, float scale, float konst,
                                                   Shared& S, Team& T) {
  const int lane = threadIdx.x & 31;
  if (t_tid<TEAM>(T) == 0) { S.n_next = 0; S.n_proc = 0; }
  team_sync<TEAM>();
  if (t_warp<TEAM>(T) == 0) {
    const uint32_t seg = L.segid;
    WarpHeap H;
    bool overflow = false;
    if (S.heap_cap == 0) {                           // first call for this label: its own region
      H.cap = (uint32_t)kHeapPerVoxel * L.n_fg + (uint32_t)kHeapSlack;
      H.k = P.heap + 3ull * ((unsigned long long)kHeapPerVoxel * L.region_off + (unsigned long long)kHeapSlack * job);
      overflow = (unsigned long long)(H.k + 3ull * H.cap - P.heap) > P.heap_static;   // the caller's buffer is for another batch
    } else {
      H.cap = S.heap_cap;
      H.k = S.heap_k;
    }
    H.v = H.k + H.cap;
    H.s = H.v + H.cap;
    H.n = 0;
    uint32_t total = 0;
    auto grow = [&]() {     // the heap can hold 26 entries per claimed voxel plus the seeds: move it to the spill arena
      const unsigned long long want = 27ull * L.n_fg + n_seeds + 64ull;
      if (want <= H.cap || want >= 0x7fffffffull) { overflow = true; return; }
      unsigned long long at = 0;
      if (lane == 0) at = atomicAdd(P.heap_bump, 3ull * want);
      at = __shfl_sync(0xffffffffu, at, 0);
      if (P.heap_static + at + 3ull * want > P.heap_words) { overflow = true; return; }
      uint32_t* nk = P.heap + P.heap_static + at;
      uint32_t* nv = nk + want;
      uint32_t* ns = nv + want;
      for (uint32_t i = lane; i < H.n; i += 32) { nk[i] = H.k[i]; nv[i] = H.v[i]; ns[i] = H.s[i]; }
      __syncwarp();
      H.k = nk; H.v = nv; H.s = ns; H.cap = (uint32_t)want;
    };
    for (uint32_t i = 0; i < n_seeds && !overflow; i++) {           // queue.emplace(0.0, sources[i], sources[i], max_distances[i])
      if (H.n == H.cap) grow();
      if (!overflow) heap_push(H, 0u, seeds[i], i, lane);
    }
    const int fdx = lane < 26 ? kDX[lane] : 0, fdy = lane < 26 ? kDY[lane] : 0, fdz = lane < 26 ? kDZ[lane] : 0;
    while (H.n > 0 && !overflow) {
      uint32_t tk, loc, sd;
      heap_pop(H, tk, loc, sd, lane);
      if (__ldcg(&A.claim[loc]) != kValid || __ldg(&A.cc[loc]) != seg) continue;   // if (!field[loc]) continue;
      __syncwarp();
      if (lane == 0) A.claim[loc] = 0ull;
      total++;
      const uint32_t o = seeds[sd];
      const float r = __fadd_rn(__fmul_rn(scale, __ldg(&A.dbf[o])), konst);
      int x, y, z, ox, oy, oz;
      unravel(loc, A.d, x, y, z);
      unravel(o, A.d, ox, oy, oz);
      // compute_neighborhood (hpp:60-124): face terms are zero at the volume's border; an edge entry needs both of its
      // terms, a corner entry only its y and z terms
      // (the reference's array is the label's bounding-box crop: its x faces are the box's; in y and z the volume's
      // faces do, what lies between them and the box is not this label)
      const int tx = fdx < 0 ? (x > (int)L.bbox_x0 ? -1 : 0) : (fdx > 0 ? (x < (int)L.bbox_x1 ? 1 : 0) : 0);
      const int ty = fdy < 0 ? (y > 0 ? -1 : 0) : (fdy > 0 ? (y < A.d.sy - 1 ? 1 : 0) : 0);
      const int tz = fdz < 0 ? (z > 0 ? -1 : 0) : (fdz > 0 ? (z < A.d.sz - 1 ? 1 : 0) : 0);
      bool ok;
      if (lane < 6) ok = (tx | ty | tz) != 0;
      else if (lane < 18) ok = (fdx == 0 || tx != 0) && (fdy == 0 || ty != 0) && (fdz == 0 || tz != 0);
      else ok = lane < 26 && ty != 0 && tz != 0;
      bool push = false;
      uint32_t nb = 0, dbits = 0;
      if (ok) {
        const int nx = x + tx, ny = y + ty, nz = z + tz;
        nb = (uint32_t)((int64_t)loc + (int64_t)tx + (int64_t)ty * A.d.sx + (int64_t)tz * A.d.sxy);
        if (__ldg(&A.cc[nb]) == seg && __ldcg(&A.claim[nb]) == kValid) {
          const float a = __fmul_rn(A.wx, (float)(nx - ox)), b = __fmul_rn(A.wy, (float)(ny - oy)),
                      c = __fmul_rn(A.wz, (float)(nz - oz));
          const float dd = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c)));
          dbits = __float_as_uint(dd);
          push = dd < r;
        }
      }
      uint32_t m = __ballot_sync(0xffffffffu, push);
      while (m && !overflow) {                                      // in the reference's neighbour order
        const int i = __ffs((int)m) - 1;
        m &= m - 1u;
        const uint32_t pk = __shfl_sync(0xffffffffu, dbits, i), pvx = __shfl_sync(0xffffffffu, nb, i);
        if (H.n == H.cap) grow();
        if (!overflow) heap_push(H, pk, pvx, sd, lane);
      }
    }
    if (lane == 0) { S.n_next = total; S.n_proc = overflow ? 1u : 0u; S.heap_k = H.k; S.heap_cap = H.cap; }
  }
  team_sync<TEAM>();
  return S.n_next;
}

// ---- dijkstra3d.path_from_parents on the parental field held in A.dist (fix_branching=False) ---------
// parents follow rule T3 (neighbour with the smallest (dist, direction)); the path is returned in
// source -> target order like the library does: out[0] = root ... out[len-1] = target.
template <bool TEAM>
__device__ uint32_t path_from_parents(const Arena& A, const LabelDesc& L, uint32_t target, uint32_t* out,
                                      uint32_t out_cap, Shared& S, Team& T) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp = t_warp<TEAM>(T);
  const uint32_t seg = L.segid;
  int dx = 0, dy = 0, dz = 0;
  if (lane < 26) { dx = kDX[lane]; dy = kDY[lane]; dz = kDZ[lane]; }
  const int64_t off = (int64_t)dx + (int64_t)dy * A.d.sx + (int64_t)dz * A.d.sxy;
  if (warp == 0) {
    uint32_t len = 0, loc = target, guard = 0;
    while (guard <= L.n_fg) {
      if (lane == 0 && len < out_cap) out[len] = loc;
      len++;
      if (loc == L.root) break;
      const uint32_t dloc = __float_as_uint(__ldcg(&A.dist[loc]));
      int x, y, z;
      unravel(loc, A.d, x, y, z);
      const int nx = x + dx, ny = y + dy, nz = z + dz;
      unsigned long long key = ~0ull;
      if (lane < 26 && nx >= 0 && ny >= 0 && nz >= 0 && nx < A.d.sx && ny < A.d.sy && nz < A.d.sz) {
        const uint32_t v = (uint32_t)((int64_t)loc + off);
        if (__ldg(&A.cc[v]) == seg) {
          const uint32_t dv = __float_as_uint(__ldcg(&A.dist[v]));
          if (dv < kInfBits) key = ((unsigned long long)dv << 32) | (unsigned long long)lane;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = __shfl_xor_sync(0xffffffffu, key, o);
        key = t < key ? t : key;
      }
      if (key == ~0ull || dloc >= kInfBits) break;            // unreachable target: the walk stops here
      const int dir = (int)(key & 31u);
      loc = (uint32_t)((int64_t)loc + (int64_t)kDX[dir] + (int64_t)kDY[dir] * A.d.sx + (int64_t)kDZ[dir] * A.d.sxy);
      guard++;
    }
    if (lane == 0) S.r32[0] = len;
  }
  team_sync<TEAM>();
  const uint32_t len = S.r32[0];
  const uint32_t n = len < out_cap ? len : out_cap;
  team_sync<TEAM>();
  for (uint32_t i = t_tid<TEAM>(T); i < n / 2; i += t_threads<TEAM>()) {   // reverse: source first
    const uint32_t a = out[i], b = out[n - 1 - i];
    out[i] = b; out[n - 1 - i] = a;
  }
  team_sync<TEAM>();
  return len;
}

// ---- the per-label conductor (trace.py:196-267) ----------------------------------------------------
template <bool TEAM>
__device__ void trace_label(const Arena& A, const LabelDesc& L, const Pools& P, const Params& prm, Shared& S, Local& Lc,
                            Team& T, uint32_t job, RrLists* R) {
  const uint32_t tid = t_tid<TEAM>(T), nth = t_threads<TEAM>();
  // scratch of the label: kScratchPerVoxel u32 per voxel.  Four pair lists and the batch of railroad (2 n_fg entries of
  // 2 words each), the touched list; the invalidation's four voxel lists alias the first pair list between searches.
  uint32_t* scr = P.scratch + (unsigned long long)kScratchPerVoxel * L.region_off;
  const unsigned long long n = L.n_fg;
  unsigned long long* pl0 = reinterpret_cast<unsigned long long*>(scr);
  unsigned long long* pl1 = reinterpret_cast<unsigned long long*>(scr + 4 * n);
  unsigned long long* pl2 = reinterpret_cast<unsigned long long*>(scr + 8 * n);
  unsigned long long* pl3 = reinterpret_cast<unsigned long long*>(scr + 12 * n);
  unsigned long long* pl4 = reinterpret_cast<unsigned long long*>(scr + 16 * n);
  uint32_t* touched = scr + 20 * n;
  uint32_t* r0 = scr;
  uint32_t* r1 = scr + n;
  uint32_t* r2 = scr + 2 * n;
  uint32_t* r3 = scr + 3 * n;
  uint32_t* out = P.paths + L.path_off;
  unsigned long long t_start = 0;
  if (tid == 0) {
    B2T_GLOBALTIMER(t_start);
    S.bucket = prm.nbuckets - 1;
    S.relax = 0; S.rounds = 0; S.invalidated = 0;
    S.heap_cap = 0;
    A.pdrf[L.root] = 0.0f;      // parents[root] = 0: the first rail (trace.py:220)
#ifdef B2T_TRACE_PROF
    for (int i = 0; i < 16; i++) S.prof[i] = 0ull;
#endif
  }
  team_sync<TEAM>();
  PROF_START();
  uint32_t valid = L.n_fg;
  int32_t status = 0;
  if (L.soma_mode) {            // one-off soma invalidation around the root (trace.py:160-168)
    // a single seed has no competitor: every claim order gives the same set
    const uint32_t n = L.soma_done ? L.pre_invalid
                                   : invalidate<TEAM>(A, L, &L.root, 1, prm.soma_scale, prm.soma_const, r0, r1, r2, r3, S, T);
    valid -= min(valid, n);
  }
  uint32_t tb_n = L.tb_n, ta_n = L.ta_n;
  uint32_t max_paths = (L.max_paths == 0xffffffffu) ? valid : L.max_paths;
  uint32_t npaths = 0, used = 0;
  const bool single = L.single_path == 1u;
  if (single) max_paths = 1;
  else if ((unsigned long long)tb_n + ta_n >= max_paths) max_paths = 0;  // trace.py:217-218: return []
  while ((valid > 0 || tb_n > 0 || ta_n > 0) && npaths < max_paths) {
    uint32_t target;
    if (tb_n > 0) target = P.targets[L.tb_off + (--tb_n)];
    else if (valid == 0) target = P.targets[L.ta_off + (--ta_n)];
    else {
      PROF_LAP(11);
      target = find_target<TEAM>(A, L, P, prm, S, Lc, T);
      PROF_LAP(0);
      if (target == 0xffffffffu) { status = -10; break; }   // bookkeeping mismatch: valid > 0 but no valid voxel
    }
    if (used + 2 > L.path_cap) { status = B2T_ERR_CAPACITY; break; }
    uint32_t* pout = out + used;
    const uint32_t cap = L.path_cap - used - 1;
    uint32_t len;
    if (!prm.fix_branching) len = path_from_parents<TEAM>(A, L, target, pout, cap, S, T);
    else if (!TEAM && B2T_RR_SOLO) len = railroad_solo(A, L, target, *R, pl2, pl3, touched, pout, cap, S, Lc, T);
    else len = railroad<TEAM>(A, L, target, pl0, pl1, pl4, pl2, pl3, touched, pout, cap, S, Lc, T);
    PROF_RESET();
    if (len > cap) { status = B2T_ERR_CAPACITY; break; }
    if (L.soma_mode) {
      // keep path[:1] + points farther than soma_radius from the root; float64, uint32 wrap
      int rx, ry, rz;
      unravel(L.root, A.d, rx, ry, rz);
      if (tid == 0) {
        uint32_t k = 0;
        r0[k++] = pout[0];
        for (uint32_t i = 0; i < len; i++) {
          int x, y, z;
          unravel(pout[i], A.d, x, y, z);
          const double ddx = (double)A.wx * (double)(uint32_t)(x - rx), ddy = (double)A.wy * (double)(uint32_t)(y - ry),
                       ddz = (double)A.wz * (double)(uint32_t)(z - rz);
          const double dist = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy)), __dmul_rn(ddz, ddz)));
          if (dist > (double)L.soma_radius) {
            if (k < L.n_fg) r0[k] = pout[i];
            k++;
          }
        }
        S.r32[1] = k;
      }
      team_sync<TEAM>();
      const uint32_t k = S.r32[1];
      if (k > cap || k > L.n_fg) { status = B2T_ERR_CAPACITY; break; }
      for (uint32_t i = tid; i < k; i += nth) pout[i] = r0[i];
      len = k;
      team_sync<TEAM>();
    }
    if (valid > 0 && !single) {
      uint32_t n;
      if (prm.inval_mode == B2T_INVALIDATE_STRICT) {
        n = invalidate_strict<TEAM>(A, L, P, job, pout, len, prm.scale, prm.konst, S, T);
        if (S.n_proc) { status = B2T_ERR_CAPACITY; break; }
      } else if (prm.inval_mode == B2T_INVALIDATE_WINDOW) {
        n = invalidate_window<TEAM>(A, L, pout, len, prm.scale, prm.konst, prm.claim_window, r0, r1, r2, r3, S, Lc, T);
      } else {
        n = invalidate<TEAM>(A, L, pout, len, prm.scale, prm.konst, r0, r1, r2, r3, S, T);
      }
      valid -= min(valid, n);
      if (tid == 0) S.invalidated += n;
      PROF_RESET();
    }
    if (prm.fix_branching)
      for (uint32_t i = tid; i < len; i += nth) A.pdrf[pout[i]] = 0.0f;   // trace.py:261-263
    if (tid == 0) pout[len] = 0xffffffffu;
    used += len + 1;
    npaths++;
    team_sync<TEAM>();
  }
  if (tid == 0) {
    P.out_len[job] = used;
    P.out_npaths[job] = npaths;
    P.out_status[job] = status;
    P.out_stats[4 * job + 0] = S.relax;
    P.out_stats[4 * job + 1] = S.rounds;
    P.out_stats[4 * job + 2] = S.invalidated;
    unsigned long long t_end;
    B2T_GLOBALTIMER(t_end);
    P.out_stats[4 * job + 3] = (uint32_t)((t_end - t_start) / 1000ull);   // microseconds this label held its team
#ifdef B2T_TRACE_PROF
    if (job < 64) for (int i = 0; i < 16; i++) g_prof[job][i] = S.prof[i];
#endif
  }
  team_sync<TEAM>();
}

// Grid: n_team clusters, each a team for job c (the caller sorts the jobs by size, largest first), then clusters whose
// CTAs work alone and pull the remaining jobs from the work counter.  d_team: one Shared slot per team in global memory.
// The argument structs are __grid_constant__: device functions below take them by reference, and a reference to an ordinary
// by-value kernel parameter makes the compiler copy all of them to local memory at kernel entry (25 64-bit stores) and
// read every field back from there (the pointers of cc / pdrf / dist were re-loaded before each of the nine neighbour
// loads of an expansion task).  A grid constant may have its address taken: the fields stay in the constant bank.
#ifdef B2T_HOST_EMU
#define B2T_GRID_CONSTANT
#else
#define B2T_GRID_CONSTANT __grid_constant__
#endif
__global__ void __launch_bounds__(kThreads, B2T_TRACE_MINB) trace_kernel(const B2T_GRID_CONSTANT Arena A,
                                                                         const LabelDesc* __restrict__ descs,
                                                                         const B2T_GRID_CONSTANT Pools P,
                                                                         const B2T_GRID_CONSTANT Params prm, uint32_t n_team,
                                                                         Shared* d_team) {
  __shared__ Shared S;
  __shared__ Local Lc;
  __shared__ LabelDesc L;
#ifdef B2T_HOST_EMU
  static RrLists rr_lists;                       // one emulated block at a time
  RrLists* R = &rr_lists;
#else
  extern __shared__ __align__(16) unsigned char b2t_trace_dyn_smem[];
  RrLists* R = reinterpret_cast<RrLists*>(b2t_trace_dyn_smem);
#endif
  const uint32_t cid = blockIdx.x / kCluster;
  Team T{blockIdx.x % kCluster, 0u};
  if (cid < n_team) {
    if (threadIdx.x == 0) L = descs[cid];
    __syncthreads();
    trace_label<true>(A, L, P, prm, d_team[cid], Lc, T, cid, R);
    return;
  }
  T.rank = 0;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) S.job = n_team + atomicAdd(P.work_counter, 1u);
    __syncthreads();
    const uint32_t job = S.job;
    if (job >= (uint32_t)prm.n_desc) return;
    if (threadIdx.x == 0) L = descs[job];
    __syncthreads();
    trace_label<false>(A, L, P, prm, S, Lc, T, job, R);
  }
}

}  // namespace

// =================================================================================================
// C ABI: the whole path loop for a batch of labels in one launch.
// Replaces the body of kimimaro/trace.py:compute_paths (trace.py:196-267) and the native calls in
// it: dijkstra3d.railroad, CachedTargetFinder.find_target, roll_invalidation_ball_inside_component.
//   d_desc       n_desc records of 20 x u32/f32 (struct LabelDesc above, same field order)
//   d_scratch    b2t_trace_scratch_words(sum(n_fg)) u32;  d_paths: path pool;  d_targets: manual targets (linear indices)
//   d_out_len / d_out_npaths / d_out_status: n_desc each; d_out_stats: 4 * n_desc; d_work_counter: 1 u32 (zeroed here)
// =================================================================================================
// Words (u32) of the strict mode's static heap regions for a batch: b2t_trace_batch wants at least this many in d_heap
// (after the two control words); everything beyond is the spill arena for heaps that outgrow their region.
B2T_EXPORT uint64_t b2t_trace_heap_words(uint64_t sum_n_fg, uint64_t n_desc) {
  return 2ull + 3ull * ((uint64_t)kHeapPerVoxel * sum_n_fg + (uint64_t)kHeapSlack * n_desc);
}

// u32 words of d_scratch for a batch whose labels have sum_n_fg voxels together (region_off of a label = the prefix sum
// of n_fg before it; a label's lists start at an even word, so pairs are 8-byte aligned)
B2T_EXPORT uint64_t b2t_trace_scratch_words(uint64_t sum_n_fg) { return (uint64_t)kScratchPerVoxel * sum_n_fg + 16; }

// bytes of one team slot (d_team of b2t_trace_batch holds n_team of them)
B2T_EXPORT uint64_t b2t_trace_team_bytes(void) { return sizeof(Shared); }

#if defined(B2T_TRACE_PROF) && !defined(B2T_HOST_EMU)
// variant build only: 64 x 16 u64 (cycles per phase, rows = the first 64 jobs of the last batch)
B2T_EXPORT int b2t_trace_prof_read(unsigned long long* h_out) {
  B2T_CUDA_TRY(cudaDeviceSynchronize());
  B2T_CUDA_TRY(cudaMemcpyFromSymbol(h_out, g_prof, sizeof(unsigned long long) * 64 * 16));
  return B2T_OK;
}
#endif

#ifndef B2T_HOST_EMU
B2T_EXPORT int b2t_trace_batch(const uint32_t* d_cc, const float* d_dbf, float* d_pdrf, float* d_dist, uint64_t* d_claim,
                               uint32_t* d_stamp, int64_t sx, int64_t sy, int64_t sz, float wx, float wy, float wz,
                               const void* d_desc, int n_desc, float scale, float konst, float soma_scale,
                               float soma_const, int fix_branching, int nbuckets, const uint64_t* d_keys,
                               const uint32_t* d_hist,
                               const uint32_t* d_cursor, uint32_t* d_scratch, uint32_t* d_paths,
                               const uint32_t* d_targets, uint32_t* d_out_len, uint32_t* d_out_npaths,
                               int32_t* d_out_status, uint32_t* d_out_stats, uint32_t* d_work_counter,
                               int invalidation_mode, float claim_window_voxels, uint32_t* d_heap, uint64_t heap_words,
                               uint64_t heap_static_words, int n_team, void* d_team, void* stream) {
  static_assert(sizeof(LabelDesc) == 80, "LabelDesc must stay 20 x 4 bytes (mirrored in kimimaro_b200/engine.py)");
  B2T_REQUIRE(sx > 0 && sy > 0 && sz > 0 && (double)sx * sy * sz < 4294967295.0, "bad volume shape");
  B2T_REQUIRE(invalidation_mode == B2T_INVALIDATE_ROUNDS || invalidation_mode == B2T_INVALIDATE_WINDOW ||
              invalidation_mode == B2T_INVALIDATE_STRICT, "b2t_trace_batch: unknown invalidation mode");
  B2T_REQUIRE(invalidation_mode != B2T_INVALIDATE_WINDOW || claim_window_voxels > 0.0f,
              "b2t_trace_batch: the window mode needs a positive window");
  B2T_REQUIRE(invalidation_mode != B2T_INVALIDATE_STRICT ||
                  (d_heap != nullptr && heap_static_words >= 2 && heap_words >= heap_static_words),
              "b2t_trace_batch: the strict mode needs a heap buffer (b2t_trace_heap_words)");
  B2T_REQUIRE(n_team >= 0 && n_team <= n_desc && (n_team == 0 || d_team != nullptr),
              "b2t_trace_batch: n_team out of range or no d_team (b2t_trace_team_bytes)");
  if (n_desc <= 0) return B2T_OK;
  cudaStream_t st = (cudaStream_t)stream;
  Arena A;
  A.cc = d_cc; A.dbf = d_dbf; A.pdrf = d_pdrf; A.dist = d_dist;
  A.claim = reinterpret_cast<unsigned long long*>(d_claim); A.stamp = d_stamp;
  A.d = Dims{(int)sx, (int)sy, (int)sz, (uint32_t)(sx * sy)};
  A.wx = wx; A.wy = wy; A.wz = wz;
  Pools P;
  P.keys = reinterpret_cast<const unsigned long long*>(d_keys); P.hist = d_hist; P.cursor = d_cursor;
  P.scratch = d_scratch; P.paths = d_paths; P.targets = d_targets; P.out_len = d_out_len; P.out_npaths = d_out_npaths;
  P.out_status = d_out_status; P.out_stats = d_out_stats; P.work_counter = d_work_counter;
  P.heap = nullptr; P.heap_words = 0; P.heap_static = 0; P.heap_bump = nullptr;
  if (invalidation_mode == B2T_INVALIDATE_STRICT) {
    // the caller sized d_heap from the batch (b2t_trace_heap_words); the kernel checks every region against heap_words
    P.heap_bump = reinterpret_cast<unsigned long long*>(d_heap);
    P.heap = d_heap + 2;
    P.heap_words = heap_words - 2;
    P.heap_static = heap_static_words - 2;
    B2T_CUDA_TRY(cudaMemsetAsync(d_heap, 0, 2 * sizeof(uint32_t), st));
  }
  // width of the key-ordered invalidation rounds: in units of the smallest voxel edge
  const float wmin = wx < wy ? (wx < wz ? wx : wz) : (wy < wz ? wy : wz);
  Params prm{scale, konst, soma_scale, soma_const, nbuckets, n_desc, fix_branching ? 1 : 0, invalidation_mode,
             claim_window_voxels * wmin};
  B2T_CUDA_TRY(cudaMemsetAsync(d_work_counter, 0, sizeof(uint32_t), st));
  int dev = 0, sms = 0, per_sm = 0;
  B2T_CUDA_TRY(cudaGetDevice(&dev));
  B2T_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t dyn_smem = B2T_RR_SOLO ? sizeof(RrLists) : 0;
  static bool smem_set = false;
  if (!smem_set && dyn_smem) {
    B2T_CUDA_TRY(cudaFuncSetAttribute(trace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
    smem_set = true;
  }
  B2T_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trace_kernel, kThreads, dyn_smem));
  if (per_sm < 1) per_sm = 1;
  if (b2t_trace_limit() > 0 && per_sm > b2t_trace_limit()) per_sm = b2t_trace_limit();
  int solo = sms * per_sm;
  if (solo > n_desc - n_team) solo = n_desc - n_team;
  const int clusters = n_team + (solo + kCluster - 1) / kCluster;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(clusters * kCluster));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = dyn_smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  B2T_CUDA_TRY(cudaLaunchKernelEx(&cfg, trace_kernel, A, reinterpret_cast<const LabelDesc*>(d_desc), P, prm,
                                  (uint32_t)n_team, reinterpret_cast<Shared*>(d_team)));
  b2t_count_launches(1);
  return B2T_OK;
}
#endif  // B2T_HOST_EMU
