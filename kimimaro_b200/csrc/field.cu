// Phase A of the TEASAR trace: whole-volume, all-labels-at-once kernels (sm_100a).
//
// The reference traces one label at a time on bounding-box crops (kimimaro/intake.py:445-515).
// On a B200 the label volume, its DBF and every per-voxel work field live in HBM as dense
// [sx,sy,sz] arrays shared by all labels (labels are disjoint, so they never touch each other's
// voxels) and the steps that precede the sequential path loop run for ALL labels in one launch:
//
//   label_stats      per-label voxel count, bounding box, max DBF, first voxel   (intake.py:195-203, trace.py:100, pyx:307-326)
//   edf_multi        26-connected geometric distance field from one source per label, all labels
//                    at once: frontier label-correcting sweep, warp per frontier voxel, lanes =
//                    the 26 neighbours, atomicMin on the float bit pattern, ballot-compacted push.
//                    Replaces dijkstra3d.euclidean_distance_field (trace.py:139-145, 302-307).
//                    Distances are the least fixed point of d[v] = min_u fl(d[u] + w_uv), i.e.
//                    bit-identical to a sequential Dijkstra regardless of relaxation order.
//   field_argmax     per-label location of the largest finite distance (return_max_location)
//   pdrf             fused zero2inf / inf2zero / compute_pdrf (trace.py:138,146,315-356) plus the
//                    initialisation of the path-loop work fields and the target-bucket histogram
//   bucket_scatter   counting sort of every label's voxels into DAF buckets (CachedTargetFinder,
//                    pyx:995-1006, without a full sort: the finder only ever needs the maximum)
//
// All of it is HBM / L2-atomic bound integer and float32 work; no tensor cores.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr uint32_t kInfBits = 0x7f800000u;
constexpr uint32_t kFrozen = 0xffffffffu;

struct Dims {
  int sx, sy, sz;
  uint32_t sxy;
};

__device__ __forceinline__ void unravel(uint32_t loc, const Dims& d, int& x, int& y, int& z) {
  z = loc / d.sxy;
  const uint32_t r = loc - (uint32_t)z * d.sxy;
  y = r / (uint32_t)d.sx;
  x = r - (uint32_t)y * d.sx;
}

// ------------------------------------------------------------------------------------------------
// label_stats: each thread walks a 16-voxel x segment and merges equal-label runs before touching
// the per-label tables, which cuts the atomic traffic by the mean run length.
// ------------------------------------------------------------------------------------------------
constexpr int kSeg = 16;

__global__ void label_stats_kernel(const uint32_t* __restrict__ cc, const float* __restrict__ dbf, Dims d,
                                   uint32_t nseg_x, uint64_t nsegs, uint32_t n_labels, uint32_t* __restrict__ count,
                                   int* __restrict__ bbox, uint32_t* __restrict__ dbfmax_bits,
                                   uint32_t* __restrict__ first) {
  const int lane = threadIdx.x & 31;
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;     // blocks are whole warps: t - lane is warp-uniform
  const bool have = t < nsegs;
  const uint32_t row = have ? (uint32_t)(t / nseg_x) : 0u;
  const int x0 = have ? (int)(t - (uint64_t)row * nseg_x) * kSeg : 0;
  const int y = row % (uint32_t)d.sy, z = row / (uint32_t)d.sy;
  const uint32_t base = row * (uint32_t)d.sx;
  const int x1 = have ? min(x0 + kSeg, d.sx) : 0;
  uint32_t cur = 0, cnt = 0;
  int xa = 0, xb = 0;
  float mx = 0.0f;
  auto flush = [&]() {
    if (cur != 0 && cur <= n_labels) {
      atomicAdd(&count[cur], cnt);
      int* b = bbox + 6 * (size_t)cur;
      atomicMin(&b[0], xa); atomicMax(&b[3], xb);
      atomicMin(&b[1], y); atomicMax(&b[4], y);
      atomicMin(&b[2], z); atomicMax(&b[5], z);
      if (dbf) atomicMax(&dbfmax_bits[cur], __float_as_uint(mx));
      atomicMin(&first[cur], base + (uint32_t)xa);
    }
  };
  int nruns = 0;
  for (int x = x0; x < x1; x++) {
    const uint32_t l = cc[base + x];
    if (l != cur) {
      flush();
      cur = l; cnt = 0; xa = x; mx = 0.0f; nruns++;
    }
    cnt++; xb = x;
    if (dbf && l) mx = fmaxf(mx, dbf[base + x]);
  }
  // The last run of every segment is settled together with the other lanes': the lanes of a warp are neighbouring segments
  // (a whole 512-voxel row when sx = 512), so inside a large label they all carry the same label and one lane does the ten
  // table atomics for the group instead of all of them.
  const bool pending = have && cur != 0 && cur <= n_labels;
  const uint32_t g = __match_any_sync(0xffffffffu, pending ? cur : 0xffffffffu);
  if (pending) {
    const uint32_t c_sum = __reduce_add_sync(g, cnt);
    const int xa_min = (int)__reduce_min_sync(g, (uint32_t)xa), xb_max = (int)__reduce_max_sync(g, (uint32_t)xb);
    const int y_min = (int)__reduce_min_sync(g, (uint32_t)y), y_max = (int)__reduce_max_sync(g, (uint32_t)y);
    const int z_min = (int)__reduce_min_sync(g, (uint32_t)z), z_max = (int)__reduce_max_sync(g, (uint32_t)z);
    const uint32_t mx_max = __reduce_max_sync(g, __float_as_uint(mx));      // non-negative floats order like their bits
    const uint32_t f_min = __reduce_min_sync(g, base + (uint32_t)xa);
    if (lane == __ffs((int)g) - 1) {
      atomicAdd(&count[cur], c_sum);
      int* b = bbox + 6 * (size_t)cur;
      atomicMin(&b[0], xa_min); atomicMax(&b[3], xb_max);
      atomicMin(&b[1], y_min); atomicMax(&b[4], y_max);
      atomicMin(&b[2], z_min); atomicMax(&b[5], z_max);
      if (dbf) atomicMax(&dbfmax_bits[cur], mx_max);
      atomicMin(&first[cur], f_min);
    }
  }
  (void)nruns;
}

__global__ void bbox_init_kernel(int* bbox, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bbox[6 * i + 0] = bbox[6 * i + 1] = bbox[6 * i + 2] = 0x7fffffff;
  bbox[6 * i + 3] = bbox[6 * i + 4] = bbox[6 * i + 5] = -1;
}

// ------------------------------------------------------------------------------------------------
// edf_multi: persistent cooperative kernel, one grid barrier per relaxation round.
// ctrl[0..2] = rotating frontier counters (round r reads ctrl[r%3], pushes to ctrl[(r+1)%3]);
// ctrl[3] = rounds executed, ctrl[4] = total relaxations that improved a voxel.
// ------------------------------------------------------------------------------------------------
struct EdfParams {
  const uint32_t* cc;
  const float* node_w;   // NULL: edge lengths (euclidean_distance_field); else cost of ENTERING a voxel (parental_field)
  float* dist;
  uint32_t* stamp;
  uint32_t* queue;
  uint32_t* ctrl;
  uint64_t cap;
  Dims d;
  float w[26];
  float wc[8];           // the same edge lengths by class: bit 0 = steps in x, bit 1 = y, bit 2 = z
};

__global__ void edf_seed_kernel(EdfParams p, const uint32_t* __restrict__ src, uint32_t n_src) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) { p.ctrl[1] = n_src; p.ctrl[0] = 0; p.ctrl[2] = 0; p.ctrl[3] = 0; p.ctrl[4] = 0; }
  if (i < n_src) {
    const uint32_t s = src[i];
    p.dist[s] = 0.0f;
    p.queue[p.cap + i] = s;  // round 1 reads buffer (1 & 1) = 1
  }
}

template <bool HAS_FROZEN, bool NODE_W>
__global__ void __launch_bounds__(1024, 2) edf_multi_kernel(EdfParams p) {
  cg::grid_group grid = cg::this_grid();
  const int lane = threadIdx.x & 31;
  const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  int dx = 0, dy = 0, dz = 0;
  float w = 0.0f;
  if (lane < 26) { dx = kDX[lane]; dy = kDY[lane]; dz = kDZ[lane]; w = p.w[lane]; }
  const int64_t off = (int64_t)dx + (int64_t)dy * p.d.sx + (int64_t)dz * p.d.sxy;
  const uint32_t ltmask = (1u << lane) - 1u;
  uint32_t improved = 0;
  uint32_t round = 1;
  for (;;) {
    const uint32_t n = __ldcg(&p.ctrl[round % 3]);
    if (n == 0) break;
    const uint32_t* qin = p.queue + (uint64_t)(round & 1) * p.cap;
    uint32_t* qout = p.queue + (uint64_t)((round + 1) & 1) * p.cap;
    uint32_t* cnt_out = &p.ctrl[(round + 1) % 3];
    // two frontier voxels per warp iteration: the sweep is bound by the chain of dependent global accesses
    // per voxel (queue -> dist/label -> neighbour label -> atomicMin -> stamp), so two independent chains
    // in flight per warp nearly halve the time of a round
    for (uint32_t it = gwarp * 2; it < n; it += nwarps * 2) {
      uint32_t u[2], lab[2], v[2], lv[2];
      float du[2];
      bool ok[2], push[2];
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const bool have = it + e < n;
        u[e] = have ? __ldcg(&qin[it + e]) : 0u;
        ok[e] = have;
      }
#pragma unroll
      for (int e = 0; e < 2; e++) {
        du[e] = ok[e] ? __ldcg(&p.dist[u[e]]) : 0.0f;
        lab[e] = ok[e] ? __ldg(&p.cc[u[e]]) : 0u;
      }
#pragma unroll
      for (int e = 0; e < 2; e++) {
        int x, y, z;
        unravel(u[e], p.d, x, y, z);
        const int nx = x + dx, ny = y + dy, nz = z + dz;
        ok[e] = ok[e] && lane < 26 && nx >= 0 && ny >= 0 && nz >= 0 && nx < p.d.sx && ny < p.d.sy && nz < p.d.sz;
        v[e] = ok[e] ? (uint32_t)((int64_t)u[e] + off) : 0u;
        lv[e] = ok[e] ? __ldg(&p.cc[v[e]]) : 0xffffffffu;
      }
#pragma unroll
      for (int e = 0; e < 2; e++) {
        push[e] = false;
        if (ok[e] && lv[e] == lab[e]) {
          bool frozen = false;
          if (HAS_FROZEN) frozen = __ldcg(&p.stamp[v[e]]) == kFrozen;
          if (!frozen) {
            const uint32_t nd = __float_as_uint(__fadd_rn(du[e], NODE_W ? __ldg(&p.node_w[v[e]]) : w));
            const uint32_t old = atomicMin(reinterpret_cast<uint32_t*>(&p.dist[v[e]]), nd);
            if (nd < old) {
              improved++;
              push[e] = atomicExch(&p.stamp[v[e]], round) != round;
            }
          }
        }
      }
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const uint32_t m = __ballot_sync(0xffffffffu, push[e]);
        if (m) {
          uint32_t base = 0;
          if (lane == 0) base = atomicAdd(cnt_out, __popc(m));
          base = __shfl_sync(0xffffffffu, base, 0);
          if (push[e]) qout[base + __popc(m & ltmask)] = v[e];
        }
      }
    }
    grid.sync();
    if (gwarp == 0 && lane == 0) { p.ctrl[round % 3] = 0; p.ctrl[3] = round; }
    round++;
  }
  if (improved) atomicAdd(&p.ctrl[4], improved);
}

// ------------------------------------------------------------------------------------------------
// edf_label: the same sweep, but one TEAM per label instead of one grid for all of them.  A round of the grid-wide
// kernel above costs a grid barrier (~9 us) and there are as many rounds as the longest geodesic of ANY label has hops
// (~1500 on a 512^3 volume), whatever the label: 14-20 ms per sweep of which almost all is barrier latency.  Here a
// label's rounds are private to its team:
//   solo   one 512-thread CTA, __syncthreads() per round, frontier counters in shared memory      (labels below the team threshold:
//          the median label has 10^4 voxels and a frontier of a few dozen)
//   team   a thread-block cluster of kEdfCluster CTAs, one hardware cluster barrier per round, counters in global memory
//          (the few labels whose frontier keeps 128 warps busy)
// and every label stops after its OWN number of rounds.  One launch: cluster c < n_team is a team for job c, the CTAs of
// the other clusters take one job each.  Same relaxation, same least fixed point, bit for bit.
// ------------------------------------------------------------------------------------------------
constexpr int kEdfCluster = 8;
constexpr int kEdfThreads = 512;
#ifndef B2T_EDF_MINB
#define B2T_EDF_MINB (B2T_EDF_SOLO ? 2 : 4)   // resident CTAs per SM (the solo form's nine-neighbour tasks want 64 registers)
#endif

struct EdfJob {
  uint32_t source;      // linear index of the source voxel
  uint32_t segid;       // (informative: relaxation follows cc[source])
  uint32_t n_fg;        // voxels of the label: its two queue buffers hold n_fg entries each
  uint32_t region_off;  // prefix sum of n_fg over the jobs: the label's queue starts at 2 * region_off
};

template <bool TEAM>
__device__ __forceinline__ void edf_team_barrier() {
#ifdef B2T_HOST_EMU
  __syncthreads();
#else
  if (TEAM) cg::this_cluster().sync(); else __syncthreads();
#endif
}

template <bool TEAM, bool NODE_W>
__device__ void edf_label_run(const EdfParams& p, const EdfJob J, uint32_t* cnt, uint32_t rank, uint32_t nranks) {
  const int lane = threadIdx.x & 31;
  const uint32_t wpb = blockDim.x >> 5;
  const uint32_t tw = rank * wpb + (threadIdx.x >> 5), ntw = nranks * wpb;
  uint32_t* q = p.queue + 2ull * J.region_off;
  int dx = 0, dy = 0, dz = 0;
  float w = 0.0f;
  if (lane < 26) { dx = kDX[lane]; dy = kDY[lane]; dz = kDZ[lane]; w = p.w[lane]; }
  const int64_t off = (int64_t)dx + (int64_t)dy * p.d.sx + (int64_t)dz * p.d.sxy;
  const uint32_t ltmask = (1u << lane) - 1u;
  if (rank == 0 && threadIdx.x == 0) {
    p.dist[J.source] = 0.0f;
    q[J.n_fg] = J.source;          // round 1 reads buffer 1
    cnt[0] = 0; cnt[1] = 1; cnt[2] = 0;
  }
  const uint32_t lab = __ldg(&p.cc[J.source]);
  edf_team_barrier<TEAM>();
  for (uint32_t round = 1;; round++) {
    const uint32_t n = TEAM ? __ldcg(&cnt[round % 3]) : cnt[round % 3];
    if (n == 0) break;
    const uint32_t* qin = q + (uint64_t)(round & 1) * J.n_fg;
    uint32_t* qout = q + (uint64_t)((round + 1) & 1) * J.n_fg;
    uint32_t* cnt_out = &cnt[(round + 1) % 3];
    for (uint32_t it = tw * 2; it < n; it += ntw * 2) {
      uint32_t u[2], v[2], lv[2];
      float du[2];
      bool ok[2], push[2];
#pragma unroll
      for (int e = 0; e < 2; e++) {
        ok[e] = it + e < n;
        u[e] = ok[e] ? __ldcg(&qin[it + e]) : 0u;
      }
#pragma unroll
      for (int e = 0; e < 2; e++) du[e] = ok[e] ? __ldcg(&p.dist[u[e]]) : 0.0f;
#pragma unroll
      for (int e = 0; e < 2; e++) {
        int x, y, z;
        unravel(u[e], p.d, x, y, z);
        const int nx = x + dx, ny = y + dy, nz = z + dz;
        ok[e] = ok[e] && lane < 26 && nx >= 0 && ny >= 0 && nz >= 0 && nx < p.d.sx && ny < p.d.sy && nz < p.d.sz;
        v[e] = ok[e] ? (uint32_t)((int64_t)u[e] + off) : 0u;
        lv[e] = ok[e] ? __ldg(&p.cc[v[e]]) : 0xffffffffu;
      }
#pragma unroll
      for (int e = 0; e < 2; e++) {
        push[e] = false;
        if (ok[e] && lv[e] == lab) {
          const uint32_t nd = __float_as_uint(__fadd_rn(du[e], NODE_W ? __ldg(&p.node_w[v[e]]) : w));
          const uint32_t old = atomicMin(reinterpret_cast<uint32_t*>(&p.dist[v[e]]), nd);
          if (nd < old) push[e] = atomicExch(&p.stamp[v[e]], round) != round;
        }
      }
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const uint32_t m = __ballot_sync(0xffffffffu, push[e]);
        if (m) {
          uint32_t base = 0;
          if (lane == 0) base = atomicAdd(cnt_out, __popc(m));
          base = __shfl_sync(0xffffffffu, base, 0);
          if (push[e]) qout[base + __popc(m & ltmask)] = v[e];
        }
      }
    }
    edf_team_barrier<TEAM>();
    if (rank == 0 && threadIdx.x == 0) cnt[round % 3] = 0;   // read again in round + 3; two barriers lie in between
  }
}

// ---- the solo form: frontier in SHARED memory, (distance, voxel) pairs, no stamps ---------------------------------
// A round of the form above is a chain of five dependent global accesses (queue -> dist[u] -> cc[v] -> atomicMin ->
// stamp exchange -> queue) behind one barrier, ~5 us, and a label has as many rounds as its longest geodesic has hops.
// Here the frontier of a CTA's label lives in two shared-memory lists of (tentative distance, voxel) pairs: a pair whose
// distance is no longer the voxel's is stale and skipped, so no stamp is needed to keep a voxel from entering a round
// twice, and dist[u] is loaded together with the neighbours' labels -- two global round trips per round (loads,
// atomicMin).  One thread per (voxel, z-plane of its neighbourhood) takes nine neighbours in memory order.  A frontier
// that outgrows the list spills into the label's global queue, voxel by voxel with the stamp rule of the form above
// (capacity n_fg: a voxel enters a round's spill at most once).  Same relaxation, same least fixed point.
#ifndef B2T_EDF_SOLO_CAP
#define B2T_EDF_SOLO_CAP 2048
#endif
#ifndef B2T_EDF_SOLO
#define B2T_EDF_SOLO 0      // off: measured on synthetic-512 (calls 31/32, profiles/r02_edf_solo_ab.jsonl) the solo form is SLOWER,
#endif                      // find_root 7.7 vs 5.8 ms, DAF 11.1 vs 8.6 ms -- with ~2000 labels the sweep is bound by how many
                            // labels are resident (4 CTAs per SM at 32 registers, no shared memory), not by a label's round
constexpr uint32_t kEdfSoloCap = B2T_EDF_SOLO_CAP;
struct EdfSoloLists {
  unsigned long long q[2][kEdfSoloCap];
  uint32_t cs[3], cg[3];         // entries of a round in the shared list / in the global spill, rotating like cnt above
};

template <bool NODE_W>
__device__ void edf_label_solo(const EdfParams& p, const EdfJob J, EdfSoloLists& R) {
  const uint32_t tid = threadIdx.x, nth = blockDim.x;
  uint32_t* gq = p.queue + 2ull * J.region_off;
  const uint32_t lab = __ldg(&p.cc[J.source]);
  if (tid == 0) {
    p.dist[J.source] = 0.0f;
    R.q[1][0] = (unsigned long long)J.source;          // key 0: round 1 reads list 1
    R.cs[0] = 0; R.cs[1] = 1; R.cs[2] = 0;
    R.cg[0] = 0; R.cg[1] = 0; R.cg[2] = 0;
  }
  __syncthreads();
  for (uint32_t round = 1;; round++) {
    const uint32_t ci = round % 3, co = (round + 1) % 3;
    const uint32_t ns = min(R.cs[ci], kEdfSoloCap), ng = R.cg[ci];
    if (ns + ng == 0) break;
    const unsigned long long* qin = R.q[round & 1];
    unsigned long long* qout = R.q[(round + 1) & 1];
    const uint32_t* gin = gq + (uint64_t)(round & 1) * J.n_fg;
    uint32_t* gout = gq + (uint64_t)((round + 1) & 1) * J.n_fg;
    for (uint32_t task = tid; task < 3u * (ns + ng); task += nth) {
      const uint32_t i = task / 3u, q = task - 3u * i;
      uint32_t u, dq = 0;
      const bool paired = i < ns;
      if (paired) { const unsigned long long pe = qin[i]; u = (uint32_t)pe; dq = (uint32_t)(pe >> 32); }
      else u = __ldcg(&gin[i - ns]);
      int x, y, z;
      unravel(u, p.d, x, y, z);
      const int nz = z + (int)q - 1;
      const float du = __ldcg(&p.dist[u]);
      const bool planeok = nz >= 0 && nz < p.d.sz;
      const int64_t base = (int64_t)u + ((int64_t)q - 1) * (int64_t)p.d.sxy;
      uint32_t lv[9];
      float wv[9];
#pragma unroll
      for (int j = 0; j < 9; j++) {
        const int ddx = j % 3 - 1, ddy = j / 3 - 1;
        const int nx = x + ddx, ny = y + ddy;
        const bool ok = planeok && nx >= 0 && nx < p.d.sx && ny >= 0 && ny < p.d.sy && !(j == 4 && q == 1u);
        const uint32_t v = (uint32_t)(base + (int64_t)ddy * p.d.sx + ddx);
        lv[j] = ok ? __ldg(&p.cc[v]) : lab + 1u;                            // never equal to lab
        if (NODE_W) wv[j] = ok ? __ldg(&p.node_w[v]) : 0.0f;
        else wv[j] = p.wc[(ddx != 0 ? 1 : 0) | (ddy != 0 ? 2 : 0) | (q != 1u ? 4 : 0)];
      }
      if (paired && __float_as_uint(du) != dq) continue;                    // a superseded pair
      uint32_t nd[9], old[9];
#pragma unroll
      for (int j = 0; j < 9; j++) {                                         // all nine atomics in flight before any result is used
        nd[j] = 0u; old[j] = 0u;
        if (lv[j] == lab) {
          const int ddx = j % 3 - 1, ddy = j / 3 - 1;
          const uint32_t v = (uint32_t)(base + (int64_t)ddy * p.d.sx + ddx);
          nd[j] = __float_as_uint(__fadd_rn(du, wv[j]));
          old[j] = atomicMin(reinterpret_cast<uint32_t*>(&p.dist[v]), nd[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 9; j++) {
        if (nd[j] < old[j]) {                                               // relaxed (old stays 0 where nothing was tried)
          const int ddx = j % 3 - 1, ddy = j / 3 - 1;
          const uint32_t v = (uint32_t)(base + (int64_t)ddy * p.d.sx + ddx);
          const uint32_t pos = atomicAdd(&R.cs[co], 1u);
          if (pos < kEdfSoloCap) qout[pos] = ((unsigned long long)nd[j] << 32) | v;
          else if (atomicExch(&p.stamp[v], round) != round) gout[atomicAdd(&R.cg[co], 1u)] = v;
        }
      }
    }
    __syncthreads();
    if (tid == 0) { R.cs[ci] = 0; R.cg[ci] = 0; }   // written again in round + 2; the barrier of round + 1 lies in between
  }
}

template <bool NODE_W>
__global__ void __launch_bounds__(kEdfThreads, B2T_EDF_MINB) edf_label_kernel(EdfParams p, const EdfJob* __restrict__ jobs,
                                                                    uint32_t n_jobs, uint32_t n_team) {
#if B2T_EDF_SOLO
  __shared__ EdfSoloLists s_lists;
#else
  __shared__ uint32_t s_cnt[4];
#endif
#ifdef B2T_HOST_EMU
  const uint32_t cid = blockIdx.x, rank = 0, csize = 1;      // emulated clusters have one CTA
#else
  const uint32_t csize = kEdfCluster;
  const uint32_t cid = blockIdx.x / csize, rank = blockIdx.x % csize;
#endif
  if (cid < n_team) {
    edf_label_run<true, NODE_W>(p, jobs[cid], p.ctrl + 4ull * cid, rank, csize);
  } else {
    const uint32_t job = n_team + (cid - n_team) * csize + rank;
#if B2T_EDF_SOLO
    if (job < n_jobs) edf_label_solo<NODE_W>(p, jobs[job], s_lists);
#else
    if (job < n_jobs) edf_label_run<false, NODE_W>(p, jobs[job], s_cnt, 0, 1);
#endif
  }
}

// soma free-space box (dijkstra3d free_space_radius): closed-form distances inside the box
// inscribed in the sphere, interior frozen, shell voxels become the frontier.  One label, one source.
__global__ void edf_freespace_seed_kernel(EdfParams p, uint32_t src, float radius, float wx, float wy, float wz) {
  int cx, cy, cz;
  unravel(src, p.d, cx, cy, cz);
  const float half = radius / sqrtf(3.0f);
  const int rx = (int)(half / wx), ry = (int)(half / wy), rz = (int)(half / wz);
  const int bx = 2 * rx + 1, by = 2 * ry + 1, bz = 2 * rz + 1;
  const uint64_t nbox = (uint64_t)bx * by * bz;
  const uint32_t lab = p.cc[src];
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbox; i += (uint64_t)gridDim.x * blockDim.x) {
    const int ix = (int)(i % bx), iy = (int)((i / bx) % by), iz = (int)(i / ((uint64_t)bx * by));
    const int x = cx - rx + ix, y = cy - ry + iy, z = cz - rz + iz;
    if (x < 0 || y < 0 || z < 0 || x >= p.d.sx || y >= p.d.sy || z >= p.d.sz) continue;
    const uint32_t loc = (uint32_t)x + (uint32_t)p.d.sx * ((uint32_t)y + (uint32_t)p.d.sy * (uint32_t)z);
    if (p.cc[loc] != lab) continue;
    const float ax = __fmul_rn(wx, (float)(x - cx)), ay = __fmul_rn(wy, (float)(y - cy)), az = __fmul_rn(wz, (float)(z - cz));
    const float dd = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)), __fmul_rn(az, az)));
    p.dist[loc] = dd;
    const bool shell = ix == 0 || iy == 0 || iz == 0 || ix == bx - 1 || iy == by - 1 || iz == bz - 1;
    if (shell) {
      const uint32_t pos = atomicAdd(&p.ctrl[1], 1u);
      p.queue[p.cap + pos] = loc;
    } else {
      p.stamp[loc] = kFrozen;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// field_argmax: per label max of (dist, smallest index) packed as (dist_bits << 32) | ~index
// ------------------------------------------------------------------------------------------------
__global__ void field_argmax_kernel(const uint32_t* __restrict__ cc, const float* __restrict__ dist, Dims d,
                                    uint32_t nseg_x, uint64_t nsegs, uint32_t n_labels,
                                    unsigned long long* __restrict__ best) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nsegs) return;
  const uint32_t row = (uint32_t)(t / nseg_x);
  const int x0 = (int)(t - (uint64_t)row * nseg_x) * kSeg;
  const uint32_t base = row * (uint32_t)d.sx;
  const int x1 = min(x0 + kSeg, d.sx);
  uint32_t cur = 0;
  unsigned long long key = 0;
  auto flush = [&]() {
    if (cur != 0 && cur <= n_labels && key != 0) {
      if (key > best[cur]) atomicMax(&best[cur], key);
    }
  };
  for (int x = x0; x < x1; x++) {
    const uint32_t l = cc[base + x];
    if (l != cur) { flush(); cur = l; key = 0; }
    if (l) {
      const uint32_t b = __float_as_uint(dist[base + x]);
      if (b < kInfBits) {
        const unsigned long long k = ((unsigned long long)b << 32) | (unsigned long long)(0xffffffffu - (base + (uint32_t)x));
        if (k > key) key = k;
      }
    }
  }
  flush();
}

// ------------------------------------------------------------------------------------------------
// pdrf (+ work-field init + bucket histogram)
// ------------------------------------------------------------------------------------------------
struct PdrfParams {
  const uint32_t* cc;
  const float* dbf;
  const float* daf;          // distance-from-root field (inf where unreachable)
  float* pdrf;
  unsigned long long* claim; // set to ~0 (valid) on participating voxels
  const float* M;            // per label: f32(1 / dbf_max^1.01)     (trace.py:336)
  const float* inv_maxdaf;   // per label: 1 / DAF[target], 0 when max_daf == 0 (trace.py:352-354)
  const uint32_t* row;       // per label: its row in the (label x bucket) tables, 0xffffffff = does not take part
  uint32_t* hist;            // [ n_rows * nbuckets ]
  uint32_t n_labels;
  int nbuckets;
  float pdrf_scale;
  float exponent;
  int n_squarings;           // >= 0: exponent is 2^n (repeated squaring, trace.py:343-345); -1: powf
  uint64_t V;
};

__device__ __forceinline__ int daf_bucket(float daf, float inv, int nb) {
  // monotone in daf; the maximum lands in bucket nb-1
  const float t = __fmul_rn(__fmul_rn(daf, inv), (float)(nb - 1));
  int b = (int)t;
  return b < 0 ? 0 : (b > nb - 1 ? nb - 1 : b);
}

// Lanes of a warp that hit the same table slot (key) elect a leader and tell every lane its rank and the group's size:
// one atomic per distinct slot and warp instead of one per voxel.  The inside of a large label is thousands of
// consecutive voxels with the same (label, bucket): without this, a soma's 19 M voxels queue up on 256 addresses.
__device__ __forceinline__ void warp_group(bool valid, uint32_t key, int lane, bool& leader, uint32_t& rank, uint32_t& size,
                                           int& leader_lane) {
  const uint32_t g = __match_any_sync(0xffffffffu, valid ? key : 0xffffffffu);
  leader_lane = __ffs((int)g) - 1;
  leader = valid && lane == leader_lane;
  rank = __popc(g & ((1u << lane) - 1u));
  size = __popc(g);
}

__global__ void pdrf_kernel(PdrfParams p) {
  const int lane = threadIdx.x & 31;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t i_first = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (uint64_t base = i_first - lane; base < p.V; base += stride) {      // warp-uniform loop: the group vote needs all lanes
    const uint64_t i = base + lane;
    uint32_t l = 0;
    bool valid = false;
    if (i < p.V) {
      l = p.cc[i];
      valid = l != 0 && l <= p.n_labels && p.row[l] != 0xffffffffu;
    }
    uint32_t key = 0;
    if (valid) {
      float dbf = p.dbf[i];
      if (dbf == 0.0f) dbf = __int_as_float(kInfBits);           // zero2inf (trace.py:138)
      float daf = p.daf[i];
      if (__float_as_uint(daf) >= kInfBits) daf = 0.0f;           // inf2zero (trace.py:146)
      float P = __fsub_rn(1.0f, __fmul_rn(dbf, p.M[l]));
      if (p.n_squarings >= 0) {
        for (int k = 0; k < p.n_squarings; k++) P = __fmul_rn(P, P);
      } else {
        P = powf(P, p.exponent);
      }
      P = __fmul_rn(P, p.pdrf_scale);
      const float inv = p.inv_maxdaf[l];
      if (inv != 0.0f) P = __fadd_rn(P, __fmul_rn(daf, inv));
      p.pdrf[i] = P;
      p.claim[i] = ~0ull;
      key = p.row[l] * (uint32_t)p.nbuckets + (uint32_t)daf_bucket(daf, inv, p.nbuckets);
    }
    bool leader; uint32_t rank, size; int ll;
    warp_group(valid, key, lane, leader, rank, size, ll);
    if (leader) atomicAdd(&p.hist[key], size);
  }
}

struct ScatterParams {
  const uint32_t* cc;
  float* dist;               // holds DAF on entry; reset to +inf on exit for participating voxels
  const float* inv_maxdaf;
  const uint32_t* row;
  uint32_t* cursor;          // exclusive-scanned histogram, advanced by the scatter
  unsigned long long* keys;  // (daf_bits << 32) | linear index, bucket-partitioned per label
  uint32_t n_labels;
  int nbuckets;
  uint64_t V;
};

__global__ void bucket_scatter_kernel(ScatterParams p) {
  const int lane = threadIdx.x & 31;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t i_first = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (uint64_t base = i_first - lane; base < p.V; base += stride) {
    const uint64_t i = base + lane;
    uint32_t l = 0;
    bool valid = false;
    if (i < p.V) {
      l = p.cc[i];
      valid = l != 0 && l <= p.n_labels && p.row[l] != 0xffffffffu;
    }
    float daf = 0.0f;
    uint32_t key = 0;
    if (valid) {
      daf = p.dist[i];
      if (__float_as_uint(daf) >= kInfBits) daf = 0.0f;
      key = p.row[l] * (uint32_t)p.nbuckets + (uint32_t)daf_bucket(daf, p.inv_maxdaf[l], p.nbuckets);
    }
    bool leader; uint32_t rank, size; int ll;
    warp_group(valid, key, lane, leader, rank, size, ll);
    uint32_t pos = 0;
    if (leader) pos = atomicAdd(&p.cursor[key], size);
    pos = __shfl_sync(0xffffffffu, pos, ll);
    if (valid) {
      p.keys[pos + rank] = ((unsigned long long)__float_as_uint(daf) << 32) | (unsigned long long)i;
      p.dist[i] = __int_as_float(kInfBits);
    }
  }
}

// single-CTA exclusive scan (n up to a few million): good enough for the (label x bucket) table
__global__ void __launch_bounds__(1024) exclusive_scan_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint64_t n) {
  __shared__ uint32_t s_part[1024];
  const uint32_t t = threadIdx.x;
  const uint64_t per = (n + 1023) / 1024;
  const uint64_t a = (uint64_t)t * per, b = min(n, a + per);
  uint32_t sum = 0;
  for (uint64_t i = a; i < b; i++) sum += in[i];
  s_part[t] = sum;
  __syncthreads();
  // Hillis-Steele inclusive scan of the 1024 partial sums
  for (int ofs = 1; ofs < 1024; ofs <<= 1) {
    uint32_t v = (t >= (uint32_t)ofs) ? s_part[t - ofs] : 0;
    __syncthreads();
    s_part[t] += v;
    __syncthreads();
  }
  uint32_t run = (t == 0) ? 0 : s_part[t - 1];
  for (uint64_t i = a; i < b; i++) { const uint32_t v = in[i]; out[i] = run; run += v; }
  if (t == 1023) out[n] = s_part[1023];
}

// The (label x bucket) table of a 512^3 volume has 1.3 M entries: the single-CTA scan above walks them with one thread per
// 1240-entry stretch (every load its own sector) and took 1.4 of K3's 3.4 ms.  Three small launches instead: per-block
// sums of 4096 coalesced entries, the single-CTA scan over those few hundred sums, block-local scans plus the offsets.
constexpr int kScanThreads = 1024, kScanPer = 4, kScanTile = kScanThreads * kScanPer;

__device__ __forceinline__ uint32_t scan_block_inclusive(uint32_t v, uint32_t* s_w) {   // all kScanThreads threads call
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
  if (lane == 31) s_w[warp] = v;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = s_w[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
    s_w[lane] = w;
  }
  __syncthreads();
  const uint32_t r = v + (warp ? s_w[warp - 1] : 0u);
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_sums_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ sums,
                                                                       uint64_t n) {
  __shared__ uint32_t s_w[32];
  const uint64_t base = (uint64_t)blockIdx.x * kScanTile;
  uint32_t v = 0;
#pragma unroll
  for (int k = 0; k < kScanPer; k++) { const uint64_t i = base + (uint64_t)k * kScanThreads + threadIdx.x; if (i < n) v += in[i]; }
  const uint32_t r = scan_block_inclusive(v, s_w);
  if (threadIdx.x == kScanThreads - 1) sums[blockIdx.x] = r;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_apply_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                                        const uint32_t* __restrict__ offs, uint64_t n) {
  __shared__ uint32_t s_w[32];
  const uint64_t i0 = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanPer;   // kScanPer consecutive entries
  uint32_t v[kScanPer], sum = 0;
#pragma unroll
  for (int k = 0; k < kScanPer; k++) { v[k] = (i0 + k < n) ? in[i0 + k] : 0u; sum += v[k]; }
  uint32_t run = scan_block_inclusive(sum, s_w) - sum + offs[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kScanPer; k++) { if (i0 + k < n) out[i0 + k] = run; run += v[k]; }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == kScanThreads - 1) out[n] = offs[gridDim.x];
}

#ifndef B2T_HOST_EMU
int coop_grid(const void* kernel, int threads, size_t smem, int* blocks_out) {
  int dev = 0, sms = 0, per_sm = 0;
  B2T_CUDA_TRY(cudaGetDevice(&dev));
  B2T_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  B2T_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
  if (per_sm < 1) { b2t_set_error("cooperative kernel does not fit on an SM"); return B2T_ERR_CUDA; }
  if (b2t_coop_limit() > 0 && per_sm > b2t_coop_limit()) per_sm = b2t_coop_limit();
  *blocks_out = sms * per_sm;
  return B2T_OK;
}
#endif

void fill_weights(float wx, float wy, float wz, float* w) {
  // float32 expressions of ext/skeletontricks/dijkstra_invalidation.hpp:45-52 (_s, _c)
  const float sxy = sqrtf(wx * wx + wy * wy), syz = sqrtf(wy * wy + wz * wz), sxz = sqrtf(wx * wx + wz * wz);
  const float c = sqrtf(wx * wx + wy * wy + wz * wz);
  w[0] = w[1] = wx; w[2] = w[3] = wy; w[4] = w[5] = wz;
  for (int i = 6; i < 10; i++) w[i] = sxy;
  for (int i = 10; i < 14; i++) w[i] = syz;
  for (int i = 14; i < 18; i++) w[i] = sxz;
  for (int i = 18; i < 26; i++) w[i] = c;
}

void fill_weights(float wx, float wy, float wz, EdfParams& p) {
  fill_weights(wx, wy, wz, p.w);
  p.wc[0] = 0.0f; p.wc[1] = p.w[0]; p.wc[2] = p.w[2]; p.wc[4] = p.w[4];
  p.wc[3] = p.w[6]; p.wc[6] = p.w[10]; p.wc[5] = p.w[14]; p.wc[7] = p.w[18];
}

int check_dims(int64_t sx, int64_t sy, int64_t sz) {
  B2T_REQUIRE(sx > 0 && sy > 0 && sz > 0, "empty volume");
  B2T_REQUIRE((double)sx * (double)sy * (double)sz < 4294967295.0, "volumes of 2^32 voxels or more are not supported");
  return B2T_OK;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
B2T_EXPORT int b2t_label_stats(const uint32_t* d_cc, const float* d_dbf, int64_t sx, int64_t sy, int64_t sz,
                               uint32_t n_labels, uint32_t* d_count, int32_t* d_bbox, float* d_dbfmax,
                               uint32_t* d_first, void* stream) {
  if (int rc = check_dims(sx, sy, sz)) return rc;
  B2T_REQUIRE(d_cc && d_count && d_bbox && d_first, "b2t_label_stats: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  Dims d{(int)sx, (int)sy, (int)sz, (uint32_t)(sx * sy)};
  const size_t n1 = (size_t)n_labels + 1;
  B2T_CUDA_TRY(cudaMemsetAsync(d_count, 0, n1 * sizeof(uint32_t), st));
  B2T_CUDA_TRY(cudaMemsetAsync(d_first, 0xff, n1 * sizeof(uint32_t), st));
  if (d_dbfmax) B2T_CUDA_TRY(cudaMemsetAsync(d_dbfmax, 0, n1 * sizeof(float), st));
  B2T_LAUNCH(bbox_init_kernel, (unsigned)((n1 + 255) / 256), 256, st)(d_bbox, (uint32_t)n1);
  const uint32_t nseg_x = (uint32_t)((sx + kSeg - 1) / kSeg);
  const uint64_t nsegs = (uint64_t)nseg_x * sy * sz;
  const unsigned blocks = (unsigned)((nsegs + 255) / 256);
  B2T_LAUNCH_SYNC(label_stats_kernel, blocks, 256, st)(d_cc, d_dbf, d, nseg_x, nsegs, n_labels, d_count, d_bbox,
                                             reinterpret_cast<uint32_t*>(d_dbfmax), d_first);
  B2T_CUDA_TRY(cudaGetLastError());
  b2t_count_launches(2);
  return B2T_OK;
}

// d_dist must be pre-filled with +inf, d_stamp with 0 (both [V]); d_queue holds 2*queue_cap u32 with
// queue_cap >= number of foreground voxels of the participating labels; d_ctrl holds >= 8 u32.
B2T_EXPORT int b2t_edf_multi(const uint32_t* d_cc, int64_t sx, int64_t sy, int64_t sz, float wx, float wy, float wz,
                             const uint32_t* d_sources, uint32_t n_sources, float free_space_radius,
                             uint32_t h_free_space_source, const float* d_node_weights, float* d_dist,
                             uint32_t* d_stamp, uint32_t* d_queue, uint64_t queue_cap, uint32_t* d_ctrl, void* stream) {
  if (int rc = check_dims(sx, sy, sz)) return rc;
  B2T_REQUIRE(d_cc && d_dist && d_stamp && d_queue && d_ctrl, "b2t_edf_multi: null pointer");
  B2T_REQUIRE(queue_cap >= n_sources, "b2t_edf_multi: queue capacity below source count");
  cudaStream_t st = (cudaStream_t)stream;
  EdfParams p;
  p.cc = d_cc; p.node_w = d_node_weights; p.dist = d_dist; p.stamp = d_stamp; p.queue = d_queue; p.ctrl = d_ctrl;
  p.cap = queue_cap;
  p.d = Dims{(int)sx, (int)sy, (int)sz, (uint32_t)(sx * sy)};
  fill_weights(wx, wy, wz, p);
  const bool frozen = free_space_radius > 0.0f;
  if (frozen) {
    B2T_REQUIRE(n_sources == 1, "free_space_radius needs exactly one source");
    B2T_REQUIRE(d_node_weights == nullptr, "free_space_radius and node weights are exclusive");
    const uint32_t zero = 0;
    B2T_LAUNCH(edf_seed_kernel, 1, 32, st)(p, &zero, 0);  // clears ctrl
    B2T_LAUNCH(edf_freespace_seed_kernel, 256, 256, st)(p, h_free_space_source, free_space_radius, wx, wy, wz);
  } else {
    if (n_sources == 0) return B2T_OK;
    B2T_LAUNCH(edf_seed_kernel, (n_sources + 255) / 256, 256, st)(p, d_sources, n_sources);
  }
  B2T_CUDA_TRY(cudaGetLastError());
#ifdef B2T_HOST_EMU   // one block of 1024 emulated threads: the grid barrier is the block barrier
  if (frozen) simt::block_launch(1, 1024, [](auto a_) { edf_multi_kernel<true, false>(a_); })(p);
  else if (d_node_weights) simt::block_launch(1, 1024, [](auto a_) { edf_multi_kernel<false, true>(a_); })(p);
  else simt::block_launch(1, 1024, [](auto a_) { edf_multi_kernel<false, false>(a_); })(p);
#else
  const void* kern = frozen ? (const void*)edf_multi_kernel<true, false>
                            : (d_node_weights ? (const void*)edf_multi_kernel<false, true> : (const void*)edf_multi_kernel<false, false>);
  int blocks = 0;
  if (int rc = coop_grid(kern, 1024, 0, &blocks)) return rc;   // few large blocks: the grid barrier costs per block
  void* args[] = {&p};
  B2T_CUDA_TRY(cudaLaunchCooperativeKernel(kern, dim3(blocks), dim3(1024), args, 0, st));
#endif
  b2t_count_launches(frozen ? 3 : 2);
  return B2T_OK;
}

// The same field, label by label: job i = (source voxel, cc id, voxel count, prefix sum of the counts before it).
// The first n_team jobs -- the caller sorts the jobs by size, largest first -- get a thread-block cluster each, the
// others one CTA each; every label runs its own rounds (see edf_label above).  d_dist: +inf, d_stamp: 0 on entry;
// d_queue: 2 * sum(n_fg) u32; d_ctrl: 4 * n_team u32.  d_node_weights as in b2t_edf_multi.
B2T_EXPORT int b2t_edf_labels(const uint32_t* d_cc, int64_t sx, int64_t sy, int64_t sz, float wx, float wy, float wz,
                              const uint32_t* d_jobs, uint32_t n_jobs, uint32_t n_team, const float* d_node_weights,
                              float* d_dist, uint32_t* d_stamp, uint32_t* d_queue, uint32_t* d_ctrl, void* stream) {
  if (int rc = check_dims(sx, sy, sz)) return rc;
  B2T_REQUIRE(d_cc && d_dist && d_stamp && d_queue && (d_jobs || n_jobs == 0), "b2t_edf_labels: null pointer");
  B2T_REQUIRE(n_team <= n_jobs && (n_team == 0 || d_ctrl), "b2t_edf_labels: n_team out of range or no d_ctrl");
  if (n_jobs == 0) return B2T_OK;
  cudaStream_t st = (cudaStream_t)stream;
  EdfParams p;
  p.cc = d_cc; p.node_w = d_node_weights; p.dist = d_dist; p.stamp = d_stamp; p.queue = d_queue; p.ctrl = d_ctrl;
  p.cap = 0;
  p.d = Dims{(int)sx, (int)sy, (int)sz, (uint32_t)(sx * sy)};
  fill_weights(wx, wy, wz, p);
  const EdfJob* jobs = reinterpret_cast<const EdfJob*>(d_jobs);
#ifdef B2T_HOST_EMU
  if (d_node_weights) simt::block_launch(n_jobs, kEdfThreads, [](auto... a_) { edf_label_kernel<true>(a_...); })(p, jobs, n_jobs, n_team);
  else simt::block_launch(n_jobs, kEdfThreads, [](auto... a_) { edf_label_kernel<false>(a_...); })(p, jobs, n_jobs, n_team);
#else
  const uint32_t n_solo = n_jobs - n_team;
  const uint32_t clusters = n_team + (n_solo + kEdfCluster - 1) / kEdfCluster;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * kEdfCluster);
  cfg.blockDim = dim3(kEdfThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kEdfCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (d_node_weights) B2T_CUDA_TRY(cudaLaunchKernelEx(&cfg, edf_label_kernel<true>, p, jobs, n_jobs, n_team));
  else B2T_CUDA_TRY(cudaLaunchKernelEx(&cfg, edf_label_kernel<false>, p, jobs, n_jobs, n_team));
#endif
  b2t_count_launches(1);
  return B2T_OK;
}

B2T_EXPORT int b2t_field_argmax(const uint32_t* d_cc, const float* d_dist, int64_t sx, int64_t sy, int64_t sz,
                                uint32_t n_labels, uint64_t* d_best, void* stream) {
  if (int rc = check_dims(sx, sy, sz)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  Dims d{(int)sx, (int)sy, (int)sz, (uint32_t)(sx * sy)};
  B2T_CUDA_TRY(cudaMemsetAsync(d_best, 0, ((size_t)n_labels + 1) * sizeof(uint64_t), st));
  const uint32_t nseg_x = (uint32_t)((sx + kSeg - 1) / kSeg);
  const uint64_t nsegs = (uint64_t)nseg_x * sy * sz;
  B2T_LAUNCH(field_argmax_kernel, (unsigned)((nsegs + 255) / 256), 256, st)(d_cc, d_dist, d, nseg_x, nsegs, n_labels,
                                                                      reinterpret_cast<unsigned long long*>(d_best));
  B2T_CUDA_TRY(cudaGetLastError());
  b2t_count_launches(1);
  return B2T_OK;
}

// Fused compute_pdrf + work-field initialisation + target-bucket build.
//   d_hist / d_cursor: (n_labels+1)*nbuckets + 1 u32 each; d_keys: u64 per foreground voxel.
//   On return d_cursor[l*nbuckets + b] is the END of bucket b of label l inside d_keys (the scatter
//   advances the exclusive offsets), d_hist holds the counts, and d_dist is +inf on every
//   participating voxel, ready for the path loop.
B2T_EXPORT int b2t_pdrf_and_buckets(const uint32_t* d_cc, const float* d_dbf, float* d_dist, float* d_pdrf,
                                    uint64_t* d_claim, int64_t sx, int64_t sy, int64_t sz,
                                    uint32_t n_labels, const float* d_M, const float* d_inv_maxdaf,
                                    const uint32_t* d_row, uint32_t n_rows, float pdrf_scale, float pdrf_exponent, int nbuckets,
                                    uint32_t* d_hist, uint32_t* d_cursor, uint64_t* d_keys, void* stream) {
  if (int rc = check_dims(sx, sy, sz)) return rc;
  B2T_REQUIRE(nbuckets >= 1 && nbuckets <= 4096, "nbuckets out of range");
  cudaStream_t st = (cudaStream_t)stream;
  const uint64_t V = (uint64_t)sx * sy * sz;
  B2T_REQUIRE(d_row != nullptr, "b2t_pdrf_and_buckets: null row table");
  const uint64_t ntab = (uint64_t)n_rows * nbuckets;   // rows = participating labels, not cc ids: a chunk with 10^6 dust
                                                       // components keeps tables of its few thousand traced labels
  B2T_CUDA_TRY(cudaMemsetAsync(d_hist, 0, (ntab + 1) * sizeof(uint32_t), st));
  PdrfParams p;
  p.cc = d_cc; p.dbf = d_dbf; p.daf = d_dist; p.pdrf = d_pdrf;
  p.claim = reinterpret_cast<unsigned long long*>(d_claim);
  p.M = d_M; p.inv_maxdaf = d_inv_maxdaf; p.row = d_row; p.hist = d_hist;
  p.n_labels = n_labels; p.nbuckets = nbuckets; p.pdrf_scale = pdrf_scale; p.exponent = pdrf_exponent;
  p.n_squarings = -1;
  {
    // is_power_of_two(pdrf_exponent) and pdrf_exponent < 2^16  (trace.py:343)
    const float e = pdrf_exponent;
    if (e == (float)(int)e && e > 0 && e < 65536.0f) {
      const int ei = (int)e;
      if ((ei & (ei - 1)) == 0) { int n = 0; while ((1 << n) < ei) n++; p.n_squarings = n; }
    }
  }
  p.V = V;
  const uint64_t want = (V + 255) / 256;
  const unsigned blocks = (unsigned)(want < 148ull * 32 ? want : 148ull * 32);
  B2T_REQUIRE(ntab < 0xffffffffull, "b2t_pdrf_and_buckets: rows * nbuckets must stay below 2^32");
  B2T_LAUNCH_SYNC(pdrf_kernel, blocks, 256, st)(p);
  if (ntab <= (uint64_t)kScanTile) {
    B2T_LAUNCH_SYNC(exclusive_scan_kernel, 1, 1024, st)(d_hist, d_cursor, ntab);
  } else {
    const unsigned tiles = (unsigned)((ntab + kScanTile - 1) / kScanTile);
    // tile sums, then their exclusive scan: a scratch buffer the library keeps (grown on demand; a first call allocates)
    const size_t want_words = 2 * (size_t)tiles + 1;
#ifdef B2T_HOST_EMU
    uint32_t* d_sums = (uint32_t*)malloc(want_words * sizeof(uint32_t));
#else
    static uint32_t* s_sums = nullptr;
    static size_t s_sums_words = 0;
    if (want_words > s_sums_words) {
      if (s_sums) { B2T_CUDA_TRY(cudaStreamSynchronize(st)); B2T_CUDA_TRY(cudaFree(s_sums)); s_sums = nullptr; s_sums_words = 0; }
      B2T_CUDA_TRY(cudaMalloc((void**)&s_sums, 2 * want_words * sizeof(uint32_t)));
      s_sums_words = 2 * want_words;
    }
    uint32_t* d_sums = s_sums;
#endif
    B2T_LAUNCH_SYNC(scan_tile_sums_kernel, tiles, kScanThreads, st)(d_hist, d_sums, ntab);
    B2T_LAUNCH_SYNC(exclusive_scan_kernel, 1, 1024, st)(d_sums, d_sums + tiles, (uint64_t)tiles);
    B2T_LAUNCH_SYNC(scan_tile_apply_kernel, tiles, kScanThreads, st)(d_hist, d_cursor, d_sums + tiles, ntab);
#ifdef B2T_HOST_EMU
    free(d_sums);
#endif
    b2t_count_launches(2);
  }
  ScatterParams s;
  s.cc = d_cc; s.dist = d_dist; s.inv_maxdaf = d_inv_maxdaf; s.row = d_row; s.cursor = d_cursor;
  s.keys = reinterpret_cast<unsigned long long*>(d_keys); s.n_labels = n_labels; s.nbuckets = nbuckets; s.V = V;
  B2T_LAUNCH_SYNC(bucket_scatter_kernel, blocks, 256, st)(s);
  B2T_CUDA_TRY(cudaGetLastError());
  b2t_count_launches(3);
  return B2T_OK;
}

// =================================================================================================
// Grid-wide rolling-ball invalidation (same round-synchronous claim semantics as trace.cu's
// in-CTA version) for seeds whose ball is too large for one CTA: the one-off soma invalidation
// (kimimaro/trace.py:160-168 -> skeletontricks.pyx:373-418 -> dijkstra_invalidation.hpp:239-332).
// All SMs expand the frontier; two grid barriers per round (claim, then finalise owners).
//   d_seeds[n_seeds]: linear indices; radius_i = fl32(fl32(scale * dbf[seed_i]) + konst)
//   d_fv / d_fs: 2 * cap u32 each (frontier voxels / owning seed), cap >= label's voxel count
//   d_ctrl: >= 8 u32; the number of invalidated voxels is left in d_ctrl[6]
// =================================================================================================
namespace {

struct BallParams {
  const uint32_t* cc;
  const float* dbf;
  unsigned long long* claim;
  const uint32_t* seeds;
  uint32_t n_seeds;
  uint32_t* fv;
  uint32_t* fs;
  uint32_t* ctrl;
  uint64_t cap;
  Dims d;
  float wx, wy, wz, scale, konst;
};

__global__ void ball_seed_kernel(BallParams p) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n_seeds) return;
  const uint32_t v = p.seeds[i];
  if (atomicCAS(&p.claim[v], ~0ull, 0ull) == ~0ull) {
    const uint32_t pos = atomicAdd(&p.ctrl[1], 1u);
    p.fv[p.cap + pos] = v;   // round 1 reads buffer 1
    p.fs[p.cap + pos] = i;
    atomicAdd(&p.ctrl[6], 1u);
  }
}

// single seed: no ownership to arbitrate, so a voxel is claimed the moment it is first reached (same result as
// the round-synchronous claim with one candidate seed) -- one grid barrier per round, no finalise pass
__global__ void __launch_bounds__(1024, 2) ball_flood_single_kernel(BallParams p) {
  cg::grid_group grid = cg::this_grid();
  const int lane = threadIdx.x & 31;
  const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  int dx = 0, dy = 0, dz = 0;
  if (lane < 26) { dx = kDX[lane]; dy = kDY[lane]; dz = kDZ[lane]; }
  const int64_t off = (int64_t)dx + (int64_t)dy * p.d.sx + (int64_t)dz * p.d.sxy;
  const uint32_t ltmask = (1u << lane) - 1u;
  const uint32_t o = p.seeds[0];
  const uint32_t seg = __ldg(&p.cc[o]);
  const float r = __fadd_rn(__fmul_rn(p.scale, __ldg(&p.dbf[o])), p.konst);
  int ox, oy, oz;
  unravel(o, p.d, ox, oy, oz);
  uint32_t round = 1;
  for (;;) {
    const uint32_t n = __ldcg(&p.ctrl[round % 3]);
    if (n == 0) break;
    const uint32_t* qv = p.fv + (uint64_t)(round & 1) * p.cap;
    uint32_t* nv = p.fv + (uint64_t)((round + 1) & 1) * p.cap;
    uint32_t* cnt_out = &p.ctrl[(round + 1) % 3];
    for (uint32_t it = gwarp; it < n; it += nwarps) {
      const uint32_t u = __ldcg(&qv[it]);
      int x, y, z;
      unravel(u, p.d, x, y, z);
      const int nx = x + dx, ny = y + dy, nz = z + dz;
      bool push = false;
      uint32_t v = 0;
      if (lane < 26 && nx >= 0 && ny >= 0 && nz >= 0 && nx < p.d.sx && ny < p.d.sy && nz < p.d.sz) {
        v = (uint32_t)((int64_t)u + off);
        const uint32_t lv = __ldg(&p.cc[v]);
        const unsigned long long cl = __ldcg(&p.claim[v]);
        if (lv == seg && cl == ~0ull) {
          const float a = __fmul_rn(p.wx, (float)(nx - ox)), b = __fmul_rn(p.wy, (float)(ny - oy)),
                      c = __fmul_rn(p.wz, (float)(nz - oz));
          const float dd = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c)));
          if (dd < r) push = atomicCAS(&p.claim[v], ~0ull, 0ull) == ~0ull;
        }
      }
      const uint32_t m = __ballot_sync(0xffffffffu, push);
      if (m) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(cnt_out, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (push) nv[base + __popc(m & ltmask)] = v;
      }
    }
    grid.sync();
    if (gwarp == 0 && lane == 0) { p.ctrl[round % 3] = 0; p.ctrl[6] += __ldcg(cnt_out); }
    round++;
  }
}

__global__ void __launch_bounds__(1024, 2) ball_flood_kernel(BallParams p) {
  cg::grid_group grid = cg::this_grid();
  const int lane = threadIdx.x & 31;
  const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t nthreads = gridDim.x * blockDim.x;
  int dx = 0, dy = 0, dz = 0;
  if (lane < 26) { dx = kDX[lane]; dy = kDY[lane]; dz = kDZ[lane]; }
  const int64_t off = (int64_t)dx + (int64_t)dy * p.d.sx + (int64_t)dz * p.d.sxy;
  const uint32_t ltmask = (1u << lane) - 1u;
  uint32_t round = 1;
  for (;;) {
    const uint32_t n = __ldcg(&p.ctrl[round % 3]);
    if (n == 0) break;
    const uint32_t* qv = p.fv + (uint64_t)(round & 1) * p.cap;
    const uint32_t* qs = p.fs + (uint64_t)(round & 1) * p.cap;
    uint32_t* nv = p.fv + (uint64_t)((round + 1) & 1) * p.cap;
    uint32_t* ns = p.fs + (uint64_t)((round + 1) & 1) * p.cap;
    uint32_t* cnt_out = &p.ctrl[(round + 1) % 3];
    for (uint32_t it = gwarp; it < n; it += nwarps) {
      const uint32_t u = __ldcg(&qv[it]), s = __ldcg(&qs[it]);
      const uint32_t o = p.seeds[s];
      const uint32_t seg = __ldg(&p.cc[o]);
      const float r = __fadd_rn(__fmul_rn(p.scale, __ldg(&p.dbf[o])), p.konst);
      int x, y, z, ox, oy, oz;
      unravel(u, p.d, x, y, z);
      unravel(o, p.d, ox, oy, oz);
      const int nx = x + dx, ny = y + dy, nz = z + dz;
      bool push = false;
      uint32_t v = 0;
      if (lane < 26 && nx >= 0 && ny >= 0 && nz >= 0 && nx < p.d.sx && ny < p.d.sy && nz < p.d.sz) {
        v = (uint32_t)((int64_t)u + off);
        if (__ldg(&p.cc[v]) == seg && __ldcg(&p.claim[v]) != 0ull) {
          const float a = __fmul_rn(p.wx, (float)(nx - ox)), b = __fmul_rn(p.wy, (float)(ny - oy)),
                      c = __fmul_rn(p.wz, (float)(nz - oz));
          const float dd = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c)));
          if (dd < r) {
            const unsigned long long cand = ((unsigned long long)__float_as_uint(dd) << 32) | s;
            push = atomicMin(&p.claim[v], cand) == ~0ull;
          }
        }
      }
      const uint32_t m = __ballot_sync(0xffffffffu, push);
      if (m) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(cnt_out, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (push) nv[base + __popc(m & ltmask)] = v;
      }
    }
    grid.sync();
    const uint32_t n_next = __ldcg(cnt_out);
    for (uint32_t i = tid; i < n_next; i += nthreads) {
      const uint32_t v = nv[i];
      ns[i] = (uint32_t)__ldcg(&p.claim[v]);
      p.claim[v] = 0ull;
    }
    if (tid == 0) { p.ctrl[round % 3] = 0; p.ctrl[6] += n_next; }
    grid.sync();
    round++;
  }
}

}  // namespace

// The same call with ONE seed (the soma's one-off ball, trace.py:160-168) without a frontier: with a single seed the
// claimed set does not depend on any order -- it is the 26-connected component, inside {voxels of the label closer to the
// seed than its radius}, that holds the seed -- so it is computed as a connected-components problem: mark the set,
// union-find over it (b2t_ccl26_roots, lock-free, no rounds), claim the seed's component.  The frontier version needs
// one grid barrier per hop of the ball's radius (11 ms for a 19 M voxel soma); this one is three streaming passes.
//   d_mark: [V] u8 scratch; d_parent: [V] u32 scratch; d_is_root: [V] u8 scratch; the count is left in d_ctrl[6]
namespace {
__global__ void ball_mark_kernel(BallParams p, uint32_t seed, uint8_t* __restrict__ mark, uint64_t V) {
  const uint32_t seg = __ldg(&p.cc[seed]);
  const float r = __fadd_rn(__fmul_rn(p.scale, __ldg(&p.dbf[seed])), p.konst);
  int ox, oy, oz;
  unravel(seed, p.d, ox, oy, oz);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < V; i += (uint64_t)gridDim.x * blockDim.x) {
    uint8_t m = 0;
    if (__ldg(&p.cc[i]) == seg && p.claim[i] == ~0ull) {
      int x, y, z;
      unravel((uint32_t)i, p.d, x, y, z);
      const float a = __fmul_rn(p.wx, (float)(x - ox)), b = __fmul_rn(p.wy, (float)(y - oy)), c = __fmul_rn(p.wz, (float)(z - oz));
      const float dd = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c)));
      m = (dd < r || i == seed) ? 1 : 0;            // the seed itself is claimed whatever its radius (hpp:297-303)
    }
    mark[i] = m;
  }
}

__global__ void ball_claim_kernel(unsigned long long* __restrict__ claim, const uint32_t* __restrict__ parent, uint32_t seed,
                                  uint64_t V, uint32_t* __restrict__ count) {
  const uint32_t want = parent[seed];
  uint32_t n = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < V; i += (uint64_t)gridDim.x * blockDim.x) {
    if (parent[i] == want && want != 0xffffffffu) { claim[i] = 0ull; n++; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  if ((threadIdx.x & 31) == 0 && n) atomicAdd(count, n);
}
}  // namespace

B2T_EXPORT int b2t_invalidate_ball_single(const uint32_t* d_cc, const float* d_dbf, uint64_t* d_claim, int64_t sx, int64_t sy,
                                          int64_t sz, float wx, float wy, float wz, uint32_t h_seed, float scale,
                                          float konst, uint8_t* d_mark, uint32_t* d_parent, uint8_t* d_is_root,
                                          uint32_t* d_ctrl, void* stream) {
  if (int rc = check_dims(sx, sy, sz)) return rc;
  B2T_REQUIRE(d_cc && d_dbf && d_claim && d_mark && d_parent && d_is_root && d_ctrl, "b2t_invalidate_ball_single: null pointer");
  const uint64_t V = (uint64_t)sx * sy * sz;
  B2T_REQUIRE(h_seed < V, "b2t_invalidate_ball_single: seed outside the volume");
  cudaStream_t st = (cudaStream_t)stream;
  B2T_CUDA_TRY(cudaMemsetAsync(d_ctrl, 0, 8 * sizeof(uint32_t), st));
  BallParams p;
  p.cc = d_cc; p.dbf = d_dbf; p.claim = reinterpret_cast<unsigned long long*>(d_claim); p.seeds = nullptr;
  p.n_seeds = 1; p.fv = nullptr; p.fs = nullptr; p.ctrl = d_ctrl; p.cap = 0;
  p.d = Dims{(int)sx, (int)sy, (int)sz, (uint32_t)(sx * sy)};
  p.wx = wx; p.wy = wy; p.wz = wz; p.scale = scale; p.konst = konst;
  const uint64_t want = (V + 255) / 256;
  const unsigned blocks = (unsigned)(want < 148ull * 16 ? want : 148ull * 16);
  B2T_LAUNCH(ball_mark_kernel, blocks, 256, st)(p, h_seed, d_mark, V);
  B2T_CUDA_TRY(cudaGetLastError());
  if (int rc = b2t_ccl26_roots(d_mark, 1, sx, sy, sz, d_parent, d_is_root, stream)) return rc;
  B2T_LAUNCH_SYNC(ball_claim_kernel, blocks, 256, st)(reinterpret_cast<unsigned long long*>(d_claim), d_parent, h_seed, V,
                                                      d_ctrl + 6);
  B2T_CUDA_TRY(cudaGetLastError());
  b2t_count_launches(2);
  return B2T_OK;
}

B2T_EXPORT int b2t_invalidate_ball(const uint32_t* d_cc, const float* d_dbf, uint64_t* d_claim, int64_t sx, int64_t sy,
                                   int64_t sz, float wx, float wy, float wz, const uint32_t* d_seeds, uint32_t n_seeds,
                                   float scale, float konst, uint32_t* d_fv, uint32_t* d_fs, uint64_t cap,
                                   uint32_t* d_ctrl, void* stream) {
  if (int rc = check_dims(sx, sy, sz)) return rc;
  B2T_REQUIRE(d_cc && d_dbf && d_claim && d_seeds && d_fv && d_fs && d_ctrl, "b2t_invalidate_ball: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  B2T_CUDA_TRY(cudaMemsetAsync(d_ctrl, 0, 8 * sizeof(uint32_t), st));
  if (n_seeds == 0) return B2T_OK;
  BallParams p;
  p.cc = d_cc; p.dbf = d_dbf; p.claim = reinterpret_cast<unsigned long long*>(d_claim); p.seeds = d_seeds;
  p.n_seeds = n_seeds; p.fv = d_fv; p.fs = d_fs; p.ctrl = d_ctrl; p.cap = cap;
  p.d = Dims{(int)sx, (int)sy, (int)sz, (uint32_t)(sx * sy)};
  p.wx = wx; p.wy = wy; p.wz = wz; p.scale = scale; p.konst = konst;
  B2T_LAUNCH(ball_seed_kernel, (n_seeds + 255) / 256, 256, st)(p);
#ifdef B2T_HOST_EMU
  if (n_seeds == 1) simt::block_launch(1, 1024, [](auto a_) { ball_flood_single_kernel(a_); })(p);
  else simt::block_launch(1, 1024, [](auto a_) { ball_flood_kernel(a_); })(p);
#else
  int blocks = 0;
  const void* kern = (n_seeds == 1) ? (const void*)ball_flood_single_kernel : (const void*)ball_flood_kernel;
  if (int rc = coop_grid(kern, 1024, 0, &blocks)) return rc;
  void* args[] = {&p};
  B2T_CUDA_TRY(cudaLaunchCooperativeKernel(kern, dim3(blocks), dim3(1024), args, 0, st));
#endif
  b2t_count_launches(2);
  return B2T_OK;
}
