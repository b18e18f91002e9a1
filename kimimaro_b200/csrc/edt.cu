// K1 -- anisotropic multi-label Euclidean distance transform for sm_100a.
//
// Replaces edt.edt() on kimimaro's hot path (kimimaro/intake.py:174-185, trace.py:112-117,
// intake.py:565).  Semantics: for every non-zero voxel, distance to the
// nearest voxel with a different label (optionally also to the virtual voxels outside the array:
// black_border), separable over the axes, squared distances carried in float32, sqrt at the end.
//
// B200-first design, three launches, HBM-bound integer/byte work (no tensor cores):
//   pass x   one warp per row, coalesced loads, run boundaries found with __ballot_sync / clz / ffs
//            (no shared-memory scan, no sequential dependency along the row);
//            R L + W 4 bytes per voxel.
//   pass y/z one CTA per [32 columns x full column] tile staged in shared memory, 128-byte
//            coalesced row segments; per voxel an exact *windowed* lower-envelope search
//            min_j f[j] + w^2 (i-j)^2 that walks outward from i and stops as soon as w^2 d^2 can no
//            longer win -- the work per voxel is proportional to its own distance value, which is
//            small for the thin processes a connectomics volume is made of.  Run membership is
//            carried in the sign bit of the staged value (f >= 0), so one LDS yields both the
//            parabola height and the "label changes here" flag.
//            R L + R 4 + W 4 bytes per voxel; sqrt fused into the last pass.
// Algorithmic traffic: (3L + 20) bytes per voxel; 32 B/voxel for uint32 labels.
#include <stdlib.h>

#include "common.cuh"
#include "edt_fh3.cuh"
#include "edt_xtma.cuh"

namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kMaxGroups = 128;  // pass x supports rows up to 32*128 = 4096 voxels

// ------------------------------------------------------------------------------------------------
// pass x
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
edt_pass_x_kernel(const T* __restrict__ labels, float* __restrict__ out, int sx, int64_t nrows, float w,
                  int black_border) {
  __shared__ uint32_t s_brk[kWarpsPerBlock][kMaxGroups];
  __shared__ uint32_t s_nz[kWarpsPerBlock][kMaxGroups];
  __shared__ int s_next[kWarpsPerBlock][kMaxGroups];

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  if (row >= nrows) return;
  const T* lrow = labels + row * sx;
  float* orow = out + row * sx;
  const int K = (sx + 31) >> 5;

  // sweep 1: label-change and non-zero bit masks, one ballot per group of 32 voxels.
  // Loads are issued 16 groups (2 KB of uint32 labels per warp) at a time before any of them is consumed.
  T carry = 0;
  for (int k0 = 0; k0 < K; k0 += 16) {
    T lab16[16];
#pragma unroll
    for (int j = 0; j < 16; j++) {
      const int p = ((k0 + j) << 5) + lane;
      lab16[j] = (p < sx) ? lrow[p] : T(0);
    }
#pragma unroll
    for (int j = 0; j < 16; j++) {
      const int k = k0 + j;
      if (k < K) {
        const int p = (k << 5) + lane;
        const bool valid = p < sx;
        const T lab = lab16[j];
        T prev = __shfl_up_sync(0xffffffffu, lab, 1);
        if (lane == 0) prev = carry;
        const bool brk = valid && (p == 0 || lab != prev);
        const uint32_t mb = __ballot_sync(0xffffffffu, brk);
        const uint32_t mn = __ballot_sync(0xffffffffu, valid && lab != T(0));
        carry = __shfl_sync(0xffffffffu, lab, 31);
        if (lane == 0) { s_brk[warp][k] = mb; s_nz[warp][k] = mn; }
      }
    }
  }
  __syncwarp();
  // sweep 2 (uniform): position of the first label change after each group
  if (lane == 0) {
    int nxt = sx;
    for (int k = K - 1; k >= 0; k--) {
      s_next[warp][k] = nxt;
      const uint32_t m = s_brk[warp][k];
      if (m) nxt = (k << 5) + __ffs(m) - 1;
    }
  }
  __syncwarp();
  // sweep 3: distances from the masks alone
  int last = 0;  // last label change before the current group (position 0 always is one)
  const uint32_t le = 0xffffffffu >> (31 - lane);
  const uint32_t gt = (lane == 31) ? 0u : (0xffffffffu << (lane + 1));
#pragma unroll 4
  for (int k = 0; k < K; k++) {
    const int p = (k << 5) + lane;
    const uint32_t m = s_brk[warp][k];
    const uint32_t nz = s_nz[warp][k];
    const uint32_t mle = m & le, mgt = m & gt;
    const int s = mle ? ((k << 5) + 31 - __clz(mle)) : last;
    const int nx = mgt ? ((k << 5) + __ffs(mgt) - 1) : s_next[warp][k];
    if (p < sx) {
      float v = 0.0f;
      if ((nz >> lane) & 1u) {
        const bool lok = (s > 0) || black_border;
        const bool rok = (nx < sx) || black_border;
        const int dl = p - s + 1, dr = nx - p;
        if (lok || rok) {
          const int d = lok ? (rok ? min(dl, dr) : dl) : dr;
          const float fd = __fmul_rn((float)d, w);
          v = __fmul_rn(fd, fd);
        } else {
          v = __int_as_float(0x7f800000);  // +inf: no label change along this row
        }
      }
      orow[p] = v;
    }
    if (m) last = (k << 5) + 31 - __clz(m);
  }
}

// ------------------------------------------------------------------------------------------------
// pass y / z: columns of length n with element stride cstride; lanes run along the contiguous
// axis.  grid.x = tiles of 32 columns, grid.y = index along the remaining axis.
// ------------------------------------------------------------------------------------------------
constexpr int kBlk = 8;   // rows per summary block of the column pass

// One candidate row of the windowed search.  Returns true when the walk in this direction is over.
// u = staged value of the row (sign bit = label change between this row and the previous one),
// d = distance in rows from the voxel being solved.
template <bool UP>
__device__ __forceinline__ bool edt_visit(uint32_t u, float d, float w2, float& best, bool edge_counts) {
  const float wd = __fmul_rn(__fmul_rn(w2, d), d);
  if (wd >= best) return true;                       // no farther row can win: w^2 d^2 only grows
  if (UP) {
    // walking towards row 0: this row is still inside the run; a flag on it closes the run above it
    best = fminf(best, __fadd_rn(__uint_as_float(u & 0x7fffffffu), wd));
    if (u >> 31) {
      if (edge_counts) { const float e = d + 1.0f; best = fminf(best, __fmul_rn(__fmul_rn(w2, e), e)); }
      return true;
    }
    return false;
  } else {
    // walking towards the end: a flag on this row means the run ended on the previous one
    if (u >> 31) { best = fminf(best, wd); return true; }
    best = fminf(best, __fadd_rn(__uint_as_float(u), wd));
    return false;
  }
}

template <typename T, bool SKIP>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
edt_pass_col_kernel(const T* __restrict__ labels, float* __restrict__ f, int n, int64_t cstride, int nx,
                    int64_t ostride, float w, int black_border, int do_sqrt) {
  // g[n][32]: staged column values, sign bit = "label differs from the previous voxel of the column"
  // sm[nb][32]: per 8-row block: minimum |value| of the block, sign bit = "some row of the block is flagged"
  extern __shared__ float g[];
  const int nb = (n + kBlk - 1) / kBlk;
  float* sm = g + (size_t)n * 32;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int x = blockIdx.x * 32 + lane;
  const bool xin = x < nx;
  const int64_t base = (int64_t)blockIdx.y * ostride + x;

  // stage the tile: every warp owns a contiguous band of rows and keeps 16 independent 128-byte
  // loads in flight (8 rows of f + 8 rows of labels) before touching any of them
  {
    const int R = (n + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const int y0 = warp * R, y1 = min(n, y0 + R);
    T prev = T(0);
    if (xin && y0 > 0 && y0 < n) prev = labels[base + (int64_t)(y0 - 1) * cstride];
    for (int yb = y0; yb < y1; yb += 8) {
      float fv[8];
      T lv[8];
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int y = yb + k;
        fv[k] = 0.0f; lv[k] = T(0);
        if (xin && y < y1) {
          const int64_t idx = base + (int64_t)y * cstride;
          fv[k] = f[idx];
          lv[k] = labels[idx];
        }
      }
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int y = yb + k;
        if (y < y1) {
          float v = fv[k];
          if (xin && (y == 0 || lv[k] != prev)) v = __uint_as_float(__float_as_uint(v) | 0x80000000u);
          g[y * 32 + lane] = v;
          prev = lv[k];
        }
      }
    }
  }
  __syncthreads();
  if (SKIP) {
    for (int b = warp; b < nb; b += kWarpsPerBlock) {
      uint32_t mn = 0x7f800000u, flag = 0;
      const int r1 = min(n, (b + 1) * kBlk);
      for (int r = b * kBlk; r < r1; r++) {
        const uint32_t u = __float_as_uint(g[r * 32 + lane]);
        mn = min(mn, u & 0x7fffffffu);   // non-negative floats order like their bit patterns
        flag |= u & 0x80000000u;
      }
      sm[b * 32 + lane] = __uint_as_float(mn | flag);
    }
    __syncthreads();
  }
  if (!xin) return;

  const float w2 = __fmul_rn(w, w);
  const bool bb = black_border != 0;
  for (int i = warp; i < n; i += kWarpsPerBlock) {
    const uint32_t gi = __float_as_uint(g[i * 32 + lane]);
    float best = __uint_as_float(gi & 0x7fffffffu);
    if (best != 0.0f) {  // background stays 0
      const int bi = i / kBlk;
      // ---------------- towards row 0 ----------------
      {
        bool done = false;
        if (gi >> 31) {                                  // the run starts on this very row
          if (i > 0 || bb) best = fminf(best, w2);
          done = true;
        }
        int r = i - 1;
        for (; !done && r >= bi * kBlk; r--)              // rest of the own block, row by row
          done = edt_visit<true>(__float_as_uint(g[r * 32 + lane]), (float)(i - r), w2, best, (r > 0) || bb);
        for (int b = bi - 1; !done && b >= 0; b--) {     // whole blocks above
          const float dmin = (float)(i - (b * kBlk + kBlk - 1));
          const float wd = __fmul_rn(__fmul_rn(w2, dmin), dmin);
          if (wd >= best) break;
          if (SKIP) {
            const uint32_t s = __float_as_uint(sm[b * 32 + lane]);
            if (!(s >> 31) && __fadd_rn(__uint_as_float(s), wd) >= best) continue;   // nothing in it can win
          }
          uint32_t u[kBlk];
#pragma unroll
          for (int k = 0; k < kBlk; k++) u[k] = __float_as_uint(g[(b * kBlk + kBlk - 1 - k) * 32 + lane]);
#pragma unroll
          for (int k = 0; k < kBlk; k++) {
            if (!done) {
              const int rr = b * kBlk + kBlk - 1 - k;
              done = edt_visit<true>(u[k], (float)(i - rr), w2, best, (rr > 0) || bb);
            }
          }
        }
      }
      // ---------------- towards the end of the column ----------------
      {
        bool done = false;
        int r = i + 1;
        const int own_end = min(n, (bi + 1) * kBlk);
        for (; !done && r < own_end; r++)
          done = edt_visit<false>(__float_as_uint(g[r * 32 + lane]), (float)(r - i), w2, best, false);
        for (int b = bi + 1; !done && b < nb; b++) {
          const float dmin = (float)(b * kBlk - i);
          const float wd = __fmul_rn(__fmul_rn(w2, dmin), dmin);
          if (wd >= best) { done = true; break; }
          if (SKIP) {
            const uint32_t s = __float_as_uint(sm[b * 32 + lane]);
            if (!(s >> 31) && __fadd_rn(__uint_as_float(s), wd) >= best) continue;
          }
          const int r1 = min(n, (b + 1) * kBlk);
          if (r1 - b * kBlk == kBlk) {
            uint32_t u[kBlk];
#pragma unroll
            for (int k = 0; k < kBlk; k++) u[k] = __float_as_uint(g[(b * kBlk + k) * 32 + lane]);
#pragma unroll
            for (int k = 0; k < kBlk; k++)
              if (!done) done = edt_visit<false>(u[k], (float)(b * kBlk + k - i), w2, best, false);
          } else {
            for (int rr = b * kBlk; !done && rr < r1; rr++)
              done = edt_visit<false>(__float_as_uint(g[rr * 32 + lane]), (float)(rr - i), w2, best, false);
          }
        }
        if (!done && bb) {                               // ran off the end of the array inside the run
          const float e = (float)(n - i);
          best = fminf(best, __fmul_rn(__fmul_rn(w2, e), e));
        }
      }
      if (do_sqrt) best = sqrtf(best);
    }
    f[base + (int64_t)i * cstride] = best;
  }
}

// ================================================================================================
// v2 kernels (ncu on the v1 kernels above showed them ISSUE-bound, not memory-bound: 1293 warp
// instructions per 512-voxel row in pass x, 513 per 32-voxel warp-row in pass y; DRAM traffic was
// already equal to the algorithmic bytes).  The v2 kernels cut the instruction count:
//   pass x  each lane keeps SEG consecutive labels of the row in registers (uint4 loads), finds the
//           run boundaries of its own segment sequentially and gets the rest from ONE inclusive
//           warp max-scan (last label change to the left) and ONE min-scan (next change to the right).
//   pass y/z one THREAD per column runs the O(n) lower-envelope algorithm (Felzenszwalb & Huttenlocher)
//           with its parabola stack in local memory (hot top in registers); lanes are 32 adjacent
//           columns, so every row access of the warp is one coalesced 128-byte segment and there is no
//           shared-memory tile at all.  Float operations mirror oracle/oracle.c (edt_parabolic_run)
//           one for one -- run-relative indices, the same intersection formula, FLT_MAX in place of
//           +inf inside the envelope arithmetic -- so the result is bit-identical to the CPU restatement.
// ================================================================================================
constexpr float kFltMax = 3.4028234664e38f;

// PRED (the "roles" form of the hybrid pass): also tell the y pass which 32-row blocks of which 32-column tiles hold a
// value above its stencil threshold -- a necessary condition for a voxel the y stencil cannot finish -- by setting
// bit (y >> 5) of pred[z * ntx + tile].  A hint only: it decides who computes a voxel, never what the result is.
template <int SEG, bool PRED = false>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
edt_pass_x_v2_kernel(const uint32_t* __restrict__ labels, float* __restrict__ out, int sx, int64_t nrows, float w,
                     int black_border, unsigned long long* __restrict__ pred = nullptr, float thr = 0.0f, int sy = 1,
                     int ntx = 0) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= nrows) return;
  const uint32_t* lrow = labels + row * sx;
  float* orow = out + row * sx;
  const int p0 = lane * SEG;
  uint32_t lab[SEG];
#pragma unroll
  for (int q = 0; q < SEG / 4; q++) {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (p0 + 4 * q < sx) v = *reinterpret_cast<const uint4*>(lrow + p0 + 4 * q);
    lab[4 * q] = v.x; lab[4 * q + 1] = v.y; lab[4 * q + 2] = v.z; lab[4 * q + 3] = v.w;
  }
  uint32_t prev = __shfl_up_sync(0xffffffffu, lab[SEG - 1], 1);
  // label changes inside the segment: bit j set <=> position p0+j starts a run
  uint32_t brk = 0;
#pragma unroll
  for (int j = 0; j < SEG; j++) {
    const int p = p0 + j;
    const uint32_t pl = (j == 0) ? prev : lab[j - 1];
    if (p < sx && (p == 0 || lab[j] != pl)) brk |= 1u << j;
  }
  // last change at or before the end of my segment / first change at or after its start, across lanes
  int last_in = brk ? p0 + 31 - __clz(brk) : -1;
  int first_in = brk ? p0 + __ffs(brk) - 1 : sx;
  int lastb = last_in, firstb = first_in;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int a = __shfl_up_sync(0xffffffffu, lastb, o);
    const int b = __shfl_down_sync(0xffffffffu, firstb, o);
    if (lane >= o) lastb = max(lastb, a);
    if (lane + o < 32) firstb = min(firstb, b);
  }
  int left_in = __shfl_up_sync(0xffffffffu, lastb, 1);      // last change strictly before my segment
  int right_in = __shfl_down_sync(0xffffffffu, firstb, 1);  // first change strictly after my segment
  if (lane == 0) left_in = 0;
  if (lane == 31) right_in = sx;
  float val[SEG];
#pragma unroll
  for (int j = 0; j < SEG; j++) {
    const int p = p0 + j;
    const uint32_t le = brk & (0xffffffffu >> (31 - j));
    const uint32_t gt = (j == 31) ? 0u : (brk & (0xffffffffu << (j + 1)));
    const int s = le ? (p0 + 31 - __clz(le)) : left_in;
    const int nx = gt ? (p0 + __ffs(gt) - 1) : right_in;
    float v = 0.0f;
    if (lab[j] != 0) {
      const bool lok = (s > 0) || black_border, rok = (nx < sx) || black_border;
      if (lok || rok) {
        const int dl = p - s + 1, dr = nx - p;
        const int d = lok ? (rok ? min(dl, dr) : dl) : dr;
        const float fd = __fmul_rn((float)d, w);
        v = __fmul_rn(fd, fd);
      } else {
        v = __int_as_float(0x7f800000);
      }
    }
    val[j] = v;
  }
#pragma unroll
  for (int q = 0; q < SEG / 4; q++)
    if (p0 + 4 * q < sx)
      *reinterpret_cast<float4*>(orow + p0 + 4 * q) = make_float4(val[4 * q], val[4 * q + 1], val[4 * q + 2], val[4 * q + 3]);
  if (PRED) {
    bool hot = false;
#pragma unroll
    for (int j = 0; j < SEG; j++) hot |= val[j] > thr;      // positions past the row end hold 0
    const uint32_t hb = __ballot_sync(0xffffffffu, hot);
    constexpr int LPT = 32 / SEG;                            // lanes per 32-column tile
    if ((lane % LPT) == 0 && p0 < sx) {
      const uint32_t tm = (LPT == 32 ? 0xffffffffu : ((1u << LPT) - 1u)) << lane;
      if (hb & tm) {
        const int64_t z = row / sy;
        const int y = (int)(row - z * sy);
        atomicOr(pred + z * ntx + (p0 >> 5), 1ull << (y >> 5));
      }
    }
  }
}

// intersection of the parabola rooted at run-relative row i (height fi) with the one at v (height h):
// same expression as oracle.c / the edt library: (f[i] - f[v] + (i-v) w^2 (i+v)) / (2 (i-v) w^2)
__device__ __forceinline__ float fh_intersect(float fi, int i, float h, int v, float w2) {
  const float f1 = __fmul_rn((float)(i - v), w2);
  const float f2 = (float)(i + v);
  return __fdiv_rn(__fadd_rn(__fsub_rn(fi, h), __fmul_rn(f1, f2)), __fmul_rn(2.0f, f1));
}

#ifndef B2T_FH_STREAM
#define B2T_FH_STREAM 1
#endif
#ifndef B2T_FH_PEND
#define B2T_FH_PEND 1
#endif
// The column pass touches every f / label exactly once: with B2T_FH_STREAM the accesses are marked evict-first
// so that L1/L2 keep the local-memory parabola stacks instead.
template <typename U> __device__ __forceinline__ U fh_ld(const U* p) {
#if B2T_FH_STREAM
  return __ldcs(p);
#else
  return *p;
#endif
}
__device__ __forceinline__ void fh_st(float* p, float v) {
#if B2T_FH_STREAM
  __stcs(p, v);
#else
  *p = v;
#endif
}

template <typename T, int NMAX, int MINB>
__global__ void __launch_bounds__(128, MINB)
edt_pass_col_fh_kernel(const T* __restrict__ labels, float* __restrict__ f, int n, int64_t cstride, int nx,
                       int64_t ostride, float w, int black_border, int last_pass) {
  const int x = blockIdx.x * 128 + threadIdx.x;
  const bool active = x < nx;                      // no early return: the warp reduces loop bounds together
  const int64_t base = (int64_t)blockIdx.y * ostride + (active ? x : 0);
  const float w2 = __fmul_rn(w, w);
  const float kInf = __int_as_float(0x7f800000);
  constexpr int kChunk = 32;                       // rows between flushes
  // Envelope entries of the runs that are not flushed yet, concatenated: row (ev), height (eh), left end
  // of the parabola's reign (ez, run-relative).  Entry 0 of a run always sits on the run's first row (F&H
  // never pops it), so ev[first] is the run start; its ez slot (the -inf boundary, never read as a
  // number) stores (run end row << 16 | number of entries).  Completed runs are queried every kChunk rows
  // and the stack restarts from 0 whenever no run is open, so for the thin processes of a connectomics
  // volume only a few dozen entries are ever live (cache resident); a blob keeps its run open and simply
  // grows the stack up to the run length.
  uint16_t ev[NMAX];
  float eh[NMAX];
  float ez[NMAX];
#define LD_EV(k_) ((int)ev[k_])
#define LD_EH(k_) (eh[k_])
#define LD_EZ(k_) (ez[k_])
#define ST_EV(k_, v_) ev[k_] = (uint16_t)(v_)
#define ST_EH(k_, v_) eh[k_] = (v_)
#define ST_EZ(k_, v_) ez[k_] = (v_)

  int ktot = 0;          // entries in the stack
  int k_lo = 0, k = -1;  // first / top entry of the open run
  int a = 0;             // first row of the open run
  T run_lab = T(0);
  int tv = 0; float th = 0.0f, tz = 0.0f;   // top entry, cached
  // query cursor over the closed runs; the NEXT envelope entry (row, height, left end) is always already in
  // registers so that stepping to it never waits for local memory
  int q_next = 0;                            // first row not written yet
  int r_lo = 0, r_cnt = 0, r_a = 0x7fffffff, r_b = 0;
  int kk = 0, cv = 0;
  float ch = 0.0f;
  int nv = 0; float nh = 0.0f, nz = kInf;    // entry kk+1 (nz = +inf when there is none)

  for (int c0 = 0; c0 < n; c0 += kChunk) {
    const int c1 = min(n, c0 + kChunk);
    // ---------------- build: rows [c0, c1), 8-row batches, the next batch is loaded while this one is used ----
    float fv[8], fnx[8];
    T lv[8], lnx[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      fnx[j] = 0.0f; lnx[j] = T(0);
      if (active && c0 + j < c1) {
        const int64_t idx = base + (int64_t)(c0 + j) * cstride;
        fnx[j] = fh_ld(f + idx);
        lnx[j] = fh_ld(labels + idx);
      }
    }
    for (int i0 = c0; i0 < c1; i0 += 8) {
#pragma unroll
      for (int j = 0; j < 8; j++) { fv[j] = fnx[j]; lv[j] = lnx[j]; }
#pragma unroll
      for (int j = 0; j < 8; j++) {
        fnx[j] = 0.0f; lnx[j] = T(0);
        if (active && i0 + 8 + j < c1) {
          const int64_t idx = base + (int64_t)(i0 + 8 + j) * cstride;
          fnx[j] = fh_ld(f + idx);
          lnx[j] = fh_ld(labels + idx);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int i = i0 + j;
        if (i < c1) {
          const T lab = lv[j];
          float fi = fv[j];
          if (fi > kFltMax) fi = kFltMax;                      // tofinite()
          if (lab != run_lab) {
            if (run_lab != T(0)) ST_EZ(k_lo, __uint_as_float(((uint32_t)i << 16) | (uint32_t)(k - k_lo + 1)));
            run_lab = lab;
            if (lab != T(0)) {
              a = i; k_lo = ktot; k = ktot; ktot++;
              ST_EV(k, i); ST_EH(k, fi);
              tv = 0; th = fi; tz = -kInf;
            }
          } else if (lab != T(0)) {
            const int ir = i - a;
            float s = fh_intersect(fi, ir, th, tv, w2);
            while (k > k_lo && s <= tz) {
              k--;
              tv = LD_EV(k) - a; th = LD_EH(k); tz = (k > k_lo) ? LD_EZ(k) : -kInf;
              s = fh_intersect(fi, ir, th, tv, w2);
            }
            k++;
            ktot = k + 1;
            ST_EV(k, i); ST_EH(k, fi); ST_EZ(k, s);
            tv = ir; th = fi; tz = s;
          }
        }
      }
    }
    const bool open = run_lab != T(0) && c1 < n;
    if (run_lab != T(0) && c1 == n) {                        // the column ends inside a run
      ST_EZ(k_lo, __uint_as_float(((uint32_t)n << 16) | (uint32_t)(k - k_lo + 1)));
      run_lab = T(0);
    }
    // ---------------- query: rows [q_next, q_end) of the closed runs ----------------
    const int q_end = open ? a : c1;                         // rows of a still open run wait for its end
    const int kdone = open ? k_lo : ktot;                    // entries that belong to closed runs
#if B2T_FH_PEND
    // a lane whose closed runs are all written has nothing to do here (most lanes of a sparse volume)
    const bool todo = active && q_next < q_end && ((r_cnt > 0 && q_next < r_b) || r_lo + r_cnt < kdone);
#else
    const bool todo = active && q_next < q_end;
#endif
    const int lo = __reduce_min_sync(0xffffffffu, todo ? q_next : 0x7fffffff);
    const int hi = __reduce_max_sync(0xffffffffu, todo ? q_end : 0);
    for (int i = lo; i < hi; i++) {
      if (todo && i >= q_next && i < q_end) {
        if (i >= r_b) {                                       // move to the next closed run
          r_lo += r_cnt;
          if (r_lo < kdone) {
            const uint32_t pk = __float_as_uint(LD_EZ(r_lo));
            r_a = LD_EV(r_lo); r_b = (int)(pk >> 16); r_cnt = (int)(pk & 0xffffu);
            kk = r_lo; cv = 0; ch = LD_EH(kk);
            nz = kInf;
            if (r_cnt > 1) { nv = LD_EV(kk + 1); nh = LD_EH(kk + 1); nz = LD_EZ(kk + 1); }
          } else {
            r_a = 0x7fffffff; r_b = q_end; r_cnt = 0;        // only background left before q_end
          }
        }
        if (i >= r_a && i < r_b) {
          const int ir = i - r_a;
          while (nz < (float)ir) {
            kk++;
            cv = nv - r_a; ch = nh;
            nz = kInf;
            if (kk + 1 < r_lo + r_cnt) { nv = LD_EV(kk + 1); nh = LD_EH(kk + 1); nz = LD_EZ(kk + 1); }
          }
          const float di = (float)(ir - cv);
          float val = __fadd_rn(__fmul_rn(__fmul_rn(w2, di), di), ch);
          if (r_a > 0 || black_border) { const float ee = (float)(ir + 1); val = fminf(val, __fmul_rn(__fmul_rn(w2, ee), ee)); }
          if (r_b < n || black_border) { const float ee = (float)(r_b - i); val = fminf(val, __fmul_rn(__fmul_rn(w2, ee), ee)); }
          if (last_pass) val = (val >= kFltMax) ? kInf : sqrtf(val);
          fh_st(f + base + (int64_t)i * cstride, val);
        }
      }
    }
    if (q_end > q_next) q_next = q_end;
    if (!open) {                                             // everything flushed: restart the stack
      ktot = 0; k_lo = 0; k = -1;
      r_lo = 0; r_cnt = 0; r_a = 0x7fffffff; r_b = 0;
    }
  }
}
#undef LD_EV
#undef LD_EH
#undef LD_EZ
#undef ST_EV
#undef ST_EH
#undef ST_EZ

// ================================================================================================
// v3 column pass: the body is edt_fh3.cuh (shared with the CPU test harness); this is its device context.
// Shared memory: three [C][128] float planes (apex row, height, left end): 12*C*128 bytes per CTA.
// ================================================================================================
template <int C, int NMAX, int NT = 128>
struct FhDevCtx {
  uint32_t sbase;   // shared-state-space address of s_plane[0][0][threadIdx.x]: slot stride NT*4 B, plane stride C*NT*4 B
  float* lv; float* lh; float* lz;      // local-memory backing, indexed by entry
  unsigned long long* nflag = nullptr;  // NEXT: prediction word of (row 0, this tile) for the next pass, rows nstride apart
  int64_t nstride = 0;
  unsigned long long nbit = 0;
  __device__ __forceinline__ bool any(bool p) const { return __any_sync(0xffffffffu, p) != 0; }
  __device__ __forceinline__ void note_next(int row) const {   // called by all lanes of the warp together
    if ((threadIdx.x & 31) == 0) atomicOr(nflag + (int64_t)row * nstride, nbit);
  }
  __device__ __forceinline__ float mul(float a, float b) const { return __fmul_rn(a, b); }
  __device__ __forceinline__ float add(float a, float b) const { return __fadd_rn(a, b); }
  __device__ __forceinline__ float sub(float a, float b) const { return __fsub_rn(a, b); }
  __device__ __forceinline__ float div(float a, float b) const { return __fdiv_rn(a, b); }
  __device__ __forceinline__ float sqrt(float a) const { return __fsqrt_rn(a); }
  __device__ __forceinline__ float fmin(float a, float b) const { return fminf(a, b); }
  template <typename U> __device__ __forceinline__ U ld_label(const U* p) const { return __ldcs(p); }
  __device__ __forceinline__ float ld_f(const float* p) const { return __ldcs(p); }
  __device__ __forceinline__ void st_f(float* p, float v) const { __stcs(p, v); }
  // explicit ld/st.shared on a 32-bit address (one address register, the plane is an immediate offset); a thread
  // only ever reads what it wrote itself, and volatile asm statements keep their order
  template <int PLANE> __device__ __forceinline__ float lds(int s) const {
    float r;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(r) : "r"(sbase + (uint32_t)s * (NT * 4u)), "n"(PLANE * C * NT * 4));
    return r;
  }
  template <int PLANE> __device__ __forceinline__ void sts(int s, float v) const {
    asm volatile("st.shared.f32 [%0+%1], %2;" ::"r"(sbase + (uint32_t)s * (NT * 4u)), "n"(PLANE * C * NT * 4), "f"(v));
  }
  __device__ __forceinline__ void s_st(int s, float v, float h, float z) { sts<0>(s, v); sts<1>(s, h); sts<2>(s, z); }
  __device__ __forceinline__ void s_st_z(int s, float z) { sts<2>(s, z); }
  __device__ __forceinline__ float s_ld_v(int s) const { return lds<0>(s); }
  __device__ __forceinline__ float s_ld_h(int s) const { return lds<1>(s); }
  __device__ __forceinline__ float s_ld_z(int s) const { return lds<2>(s); }
  __device__ __forceinline__ void l_st(int k, float v, float h, float z) { lv[k] = v; lh[k] = h; lz[k] = z; }
  __device__ __forceinline__ void l_st_z(int k, float z) { lz[k] = z; }
  __device__ __forceinline__ float l_ld_v(int k) const { return lv[k]; }
  __device__ __forceinline__ float l_ld_h(int k) const { return lh[k]; }
  __device__ __forceinline__ float l_ld_z(int k) const { return lz[k]; }
  __device__ __forceinline__ int wmin(int x) const { return __reduce_min_sync(0xffffffffu, x); }
  __device__ __forceinline__ int wmax(int x) const { return __reduce_max_sync(0xffffffffu, x); }
};

// device context of the stencil half of the hybrid pass (no stack): arithmetic, streaming accesses, warp votes,
// and the prefetch rings: rows go global -> shared with cp.async (no register staging, so the prefetch depth is free)
struct StDevCtx {
  unsigned long long* flag;   // flag word of this warp's tile (bit b: rows [32b, 32b+32) need the envelope kernel)
  int last_blk;
  uint32_t fbase, lbase;      // shared-state-space addresses of s_f[0][threadIdx.x], s_l[0][threadIdx.x]; slots 512 B apart
  unsigned long long* nflag = nullptr;  // NEXT: see FhDevCtx
  int64_t nstride = 0;
  unsigned long long nbit = 0;
  __device__ __forceinline__ void note_next(int row) const {
    if ((threadIdx.x & 31) == 0) atomicOr(nflag + (int64_t)row * nstride, nbit);
  }
  __device__ __forceinline__ float mul(float a, float b) const { return __fmul_rn(a, b); }
  __device__ __forceinline__ float add(float a, float b) const { return __fadd_rn(a, b); }
  __device__ __forceinline__ float sqrt(float a) const { return __fsqrt_rn(a); }
  __device__ __forceinline__ float fmin(float a, float b) const { return fminf(a, b); }
  __device__ __forceinline__ void st_f(float* p, float v) const { __stcs(p, v); }
  __device__ __forceinline__ bool any(bool p) const { return __any_sync(0xffffffffu, p) != 0; }
  // v >= 0 (or +inf): unsigned order of the bit patterns == order of the values
  __device__ __forceinline__ float wmaxf(float v) const {
    return __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(v)));
  }
  __device__ __forceinline__ void note_row(int row) {   // called by all lanes of the warp together
    const int blk = row >> 5;
    if (blk != last_blk) {
      last_blk = blk;
      if ((threadIdx.x & 31) == 0) atomicOr(flag, 1ull << blk);
    }
  }
  template <typename U> __device__ __forceinline__ U ld_label(const U* p) const { return __ldcs(p); }
  __device__ __forceinline__ void ring_fetch_f(int foff, const float* fp) const {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(fbase + (uint32_t)foff), "l"(fp) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  template <typename U> __device__ __forceinline__ void ring_fetch(int foff, int loff, const U* lp, const float* fp) const {
    static_assert(sizeof(U) == 4, "the ring holds 32-bit labels");
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(fbase + (uint32_t)foff), "l"(fp) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(lbase + (uint32_t)loff), "l"(lp) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  __device__ __forceinline__ void ring_put(int foff, float f) const {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(fbase + (uint32_t)foff), "f"(f) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  template <int N> __device__ __forceinline__ void ring_wait() const {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
  }
  __device__ __forceinline__ float ring_f(int foff) const {
    float r;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(fbase + (uint32_t)foff) : "memory");
    return r;
  }
  template <typename U> __device__ __forceinline__ U ring_l(int loff) const {
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(lbase + (uint32_t)loff) : "memory");
    return (U)r;
  }
};

template <typename T, int W, int WR, int D, int MINB, bool WRITE_BG, int OPT = 0>
__global__ void __launch_bounds__(128, MINB)
edt_pass_col_stencil_kernel(const T* __restrict__ labels, const float* __restrict__ fin, float* __restrict__ fout, int n,
                            int64_t cstride, int nx, int64_t ostride, float w, int black_border, int last_pass,
                            unsigned long long* __restrict__ flags, int ntx) {
  static_assert(fh3::kRingSlotBytes == 128 * 4, "one ring slot = one float per thread of the CTA");
  __shared__ float s_f[fh3::kRingF][128];
  __shared__ uint32_t s_l[fh3::kRingL][128];
  const int x = blockIdx.x * 128 + threadIdx.x;
  const int tile = (blockIdx.x * 128 + (threadIdx.x & ~31)) >> 5;
  if (tile >= ntx) return;                         // the whole warp is outside the volume
  const bool active = x < nx;
  const int64_t base = (int64_t)blockIdx.y * ostride + (active ? x : nx - 1);   // a lane outside shadows the last column
  StDevCtx cx{flags + (int64_t)blockIdx.y * ntx + tile, -1, (uint32_t)__cvta_generic_to_shared(&s_f[0][threadIdx.x]),
              (uint32_t)__cvta_generic_to_shared(&s_l[0][threadIdx.x])};
  if (OPT)
    fh3::stencil_column_v2<T, W, WR, D, WRITE_BG>(cx, labels + base, fin + base, fout + base, n, cstride, w,
                                                  black_border != 0, last_pass != 0, active);
  else
    fh3::stencil_column<T, W, WR, D, WRITE_BG>(cx, labels + base, fin + base, fout + base, n, cstride, w, black_border != 0,
                                               last_pass != 0, active);
}

// envelope half of the hybrid pass: only the flagged 32-row blocks of a tile, extended to complete runs.
// One WARP per CTA: most tiles have nothing flagged and leave at once, and a resident CTA that is one busy warp
// plus three finished ones would waste three quarters of its slot.
template <typename T, int C, int NMAX, int MINB, int R, int B, int QP = 1, int PP = 0>
__global__ void __launch_bounds__(32, MINB)
edt_pass_col_fh3_range_kernel(const T* __restrict__ labels, const float* __restrict__ fin, float* __restrict__ fout, int n,
                              int64_t cstride, int nx, int64_t ostride, float w, int black_border, int last_pass,
                              const unsigned long long* __restrict__ flags, int ntx) {
  __shared__ float s_plane[3][C][32];
  const int tile = blockIdx.x;
  const int x = tile * 32 + threadIdx.x;
  unsigned long long m = flags[(int64_t)blockIdx.y * ntx + tile];   // warp-uniform
  if (m == 0ull) return;
  float lv[NMAX];
  float lh[NMAX];
  float lz[NMAX];
  const bool active = x < nx;
  const int64_t base = (int64_t)blockIdx.y * ostride + (active ? x : 0);
  FhDevCtx<C, NMAX, 32> cx{(uint32_t)__cvta_generic_to_shared(&s_plane[0][0][threadIdx.x]), lv, lh, lz};
  while (m) {                                      // maximal groups of consecutive flagged blocks
    const int b0 = __ffsll((long long)m) - 1;
    const unsigned long long rest = ~(m >> b0);    // first clear bit above b0 ends the group
    const int len = rest ? (__ffsll((long long)rest) - 1) : 64;
    const int b1 = b0 + len - 1;
    m = (b1 >= 63) ? 0ull : (m & (~0ull << (b1 + 1)));
    const int rlo = 32 * b0, rhi = min(n - 1, 32 * b1 + 31);
    int own_lo, own_hi;
    fh3::extend_to_runs<T>(cx, labels + base, n, cstride, active, rlo, rhi, own_lo, own_hi);
    const int rb = cx.wmin(own_lo), re = cx.wmax(own_hi);
    if (rb < re)
      fh3::column_range<T, C, R, B, true, false, QP, PP>(cx, labels + base, fin + base, fout + base, n, cstride, w,
                                                     black_border != 0, last_pass != 0, active, rb, re, own_lo, own_hi);
  }
}

// "Roles" form of the hybrid pass.  ncu on the two-kernel form: the envelope kernel is a latency-bound chain (one warp
// per blob tile walks ~450 rows at ~0.1 instructions per cycle, 18 such warps per SM, 40 % issue) that only STARTS when
// the stencil kernel has ended.  Here the PREVIOUS pass predicts the blocks the stencil cannot finish (the x pass for y,
// the y pass for z: a value above the threshold going in is necessary for one coming out), and one launch runs both
// roles side by side: the CTAs with blockIdx.z == 0 are dispatched first and run the envelope over the predicted blocks
// of their four tiles (most have none and leave at once), the CTAs with blockIdx.z == 1 run the stencil, skipping the
// predicted blocks (PRED in edt_fh3.cuh) and flagging what the prediction missed for a residual envelope launch.  The
// chain starts at t = 0 and the stencil warps fill the issue slots it leaves.  Both roles report rows above the next
// pass's threshold (NEXT).  Shared memory: the 48 ring slots of the stencil role are the 3 x 16 plane slots of the
// envelope role; a thread only ever touches column threadIdx.x of either.
template <typename T, int W, int WR, int D, int MINB, bool WRITE_BG, bool NEXT, int NMAX, int R, int B, int OPT>
__global__ void __launch_bounds__(128, MINB)
edt_pass_col_roles_kernel(const T* __restrict__ labels, const float* __restrict__ fin, float* __restrict__ fout, int n,
                          int64_t cstride, int nx, int64_t ostride, float w, int black_border, int last_pass,
                          const unsigned long long* __restrict__ pred, unsigned long long* __restrict__ resid,
                          unsigned long long* __restrict__ next, float thr_next, int ntx) {
  static_assert(fh3::kRingSlotBytes == 128 * 4, "one ring slot = one float per thread of the CTA");
  static_assert(fh3::kRingF + fh3::kRingL == 48, "48 slots: f ring + label ring == 3 planes of 16 entries");
  constexpr int C = 16;
  __shared__ float s_raw[fh3::kRingF + fh3::kRingL][128];
  const int x = blockIdx.x * 128 + threadIdx.x;
  const int tile = (blockIdx.x * 128 + (threadIdx.x & ~31)) >> 5;
  if (tile >= ntx) return;                         // the whole warp is outside the volume
  const bool active = x < nx;
  const int outer = blockIdx.y;
  unsigned long long m = pred[(int64_t)outer * ntx + tile];   // warp-uniform
  const unsigned long long nbit = 1ull << (outer >> 5);
  if (blockIdx.z == 0) {
    if (m == 0ull) return;
    float lv[NMAX];
    float lh[NMAX];
    float lz[NMAX];
    const int64_t base = (int64_t)outer * ostride + (active ? x : 0);
    FhDevCtx<C, NMAX, 128> cx{(uint32_t)__cvta_generic_to_shared(&s_raw[0][threadIdx.x]), lv, lh, lz, next + tile, ntx, nbit};
    while (m) {                                    // maximal groups of consecutive predicted blocks
      const int b0 = __ffsll((long long)m) - 1;
      const unsigned long long rest = ~(m >> b0);
      const int len = rest ? (__ffsll((long long)rest) - 1) : 64;
      const int b1 = b0 + len - 1;
      m = (b1 >= 63) ? 0ull : (m & (~0ull << (b1 + 1)));
      const int rlo = 32 * b0, rhi = min(n - 1, 32 * b1 + 31);
      if (rlo >= n) break;                         // a prediction is a hint: ignore bits past the column end
      int own_lo, own_hi;
      fh3::extend_to_runs<T>(cx, labels + base, n, cstride, active, rlo, rhi, own_lo, own_hi);
      const int rb = cx.wmin(own_lo), re = cx.wmax(own_hi);
      if (rb < re)
        fh3::column_range<T, C, R, B, true, NEXT>(cx, labels + base, fin + base, fout + base, n, cstride, w, black_border != 0,
                                                  last_pass != 0, active, rb, re, own_lo, own_hi, thr_next);
    }
  } else {
    const int64_t base = (int64_t)outer * ostride + (active ? x : nx - 1);   // a lane outside shadows the last column
    StDevCtx cx{resid + (int64_t)outer * ntx + tile, -1, (uint32_t)__cvta_generic_to_shared(&s_raw[0][threadIdx.x]),
                (uint32_t)__cvta_generic_to_shared(&s_raw[fh3::kRingF][threadIdx.x]), next + tile, ntx, nbit};
    if (OPT)
      fh3::stencil_column_v2<T, W, WR, D, WRITE_BG, true, NEXT>(cx, labels + base, fin + base, fout + base, n, cstride, w,
                                                                black_border != 0, last_pass != 0, active, m, thr_next);
    else
      fh3::stencil_column<T, W, WR, D, WRITE_BG, true, NEXT>(cx, labels + base, fin + base, fout + base, n, cstride, w,
                                                             black_border != 0, last_pass != 0, active, m, thr_next);
  }
}

template <typename T, int C, int NMAX, int MINB, int R, int B>
__global__ void __launch_bounds__(128, MINB)
edt_pass_col_fh3_kernel(const T* __restrict__ labels, float* __restrict__ f, int n, int64_t cstride, int nx,
                        int64_t ostride, float w, int black_border, int last_pass) {
  __shared__ float s_plane[3][C][128];
  float lv[NMAX];
  float lh[NMAX];
  float lz[NMAX];
  const int x = blockIdx.x * 128 + threadIdx.x;
  const bool active = x < nx;                      // no early return: the warp reduces loop bounds together
  const int64_t base = (int64_t)blockIdx.y * ostride + (active ? x : 0);
  FhDevCtx<C, NMAX> cx{(uint32_t)__cvta_generic_to_shared(&s_plane[0][0][threadIdx.x]), lv, lh, lz};
  fh3::column<T, C, R, B>(cx, labels + base, f + base, n, cstride, w, black_border != 0, last_pass != 0, active);
}

// Kernel selection (experiments and A/B timing; results are identical whatever is chosen).  Initialised from the
// environment -- B2T_EDT_ALGO: 3 = shared-memory-ring F-H (default), 2 = local-memory F-H, w = windowed search;
// B2T_FH3 = "C,MINB,R,B": one of the compiled instantiations -- and changeable at run time with b2t_edt_config().
struct EdtCfg { int algo, c, minb, r, b; int hybrid, hwy, hwz, hwr, hpf, hminb; int ec, eminb, er, eb; int roles, sopt; float pscale; int eqp, epp; int xtma; };
static EdtCfg& edt_cfg() {
  static EdtCfg cfg = []() {
    // hwy = 0: tap radii chosen from the anisotropy; stencil_column_v2 on; envelope write-out with four entries of lookahead
    // (round 2, call 49: bit-identical on all seven inputs of scripts/k1_shot.py, 2.388 -> 2.324 ms on synthetic-512)
    EdtCfg c{3, 16, 6, 32, 4, 1, 0, 0, 4, 11, 8, 32, 4, 32, 8, 0, 1, 1.0f, 4, 0, 0};
    const char* a = getenv("B2T_EDT_ALGO");
    if (a) c.algo = (a[0] == 'w') ? 1 : (a[0] == '2' ? 2 : 3);
    const char* e = getenv("B2T_FH3");
    if (e) sscanf(e, "%d,%d,%d,%d", &c.c, &c.minb, &c.r, &c.b);
    const char* h = getenv("B2T_EDT_HYBRID");   // "0" = off, or "WY,WZ,WR,PF,MINB"
    if (h) { if (h[0] == '0' && h[1] == 0) c.hybrid = 0; else sscanf(h, "%d,%d,%d,%d,%d", &c.hwy, &c.hwz, &c.hwr, &c.hpf, &c.hminb); }
    const char* r = getenv("B2T_EDT_ROLES");    // "1": predicted envelope + stencil in one launch per pass
    if (r) c.roles = atoi(r);
    const char* q = getenv("B2T_EDT_QP");       // "4": the envelope kernel's write-out keeps 4 entries ahead in registers
    if (q) c.eqp = (atoi(q) == 4) ? 4 : 1;
    const char* pp = getenv("B2T_EDT_PP");      // "1": the envelope kernel's build keeps the entry below the top in registers
    if (pp) c.epp = atoi(pp) ? 1 : 0;
    const char* xt = getenv("B2T_EDT_XTMA");    // "1": the x pass stages its tiles with TMA (edt_xtma.cuh)
    if (xt) c.xtma = atoi(xt) == 2 ? 2 : (atoi(xt) ? 1 : 0);
    const char* o = getenv("B2T_EDT_STENCIL");  // "1": the first stencil body; "2": stencil_column_v2 (default)
    if (o) c.sopt = (atoi(o) == 1) ? 0 : 1;
    return c;
  }();
  return cfg;
}

template <typename T, int NMAX>
bool edt_launch_fh3_passes(const T* labels, int64_t sx, int64_t sy, int64_t sz, float wy, float wz, int black_border,
                           int ndim, float* out, cudaStream_t st) {
  const dim3 gy((unsigned)b2t_ceil_div(sx, 128), (unsigned)sz), gz((unsigned)b2t_ceil_div(sx, 128), (unsigned)sy);
  const EdtCfg c = edt_cfg();
#define B2T_FH3_GO(C_, MB_, R_, B_)                                                                                    \
  if (c.c == C_ && c.minb == MB_ && c.r == R_ && c.b == B_) {                                                          \
    edt_pass_col_fh3_kernel<T, C_, NMAX, MB_, R_, B_><<<gy, 128, 0, st>>>(labels, out, (int)sy, sx, (int)sx, sx * sy, \
                                                                         wy, black_border, ndim == 2);                \
    if (ndim == 3)                                                                                                     \
      edt_pass_col_fh3_kernel<T, C_, NMAX, MB_, R_, B_><<<gz, 128, 0, st>>>(labels, out, (int)sz, sx * sy, (int)sx,   \
                                                                           sx, wz, black_border, 1);                  \
    return true;                                                                                                       \
  }
  B2T_FH3_GO(16, 6, 32, 4)
  B2T_FH3_GO(16, 8, 32, 4)
  B2T_FH3_GO(32, 4, 32, 8)
#undef B2T_FH3_GO
  return false;
}

template <typename T>
int edt_launch_v2(const T* labels, int64_t sx, int64_t sy, int64_t sz, float wx, float wy, float wz, int black_border,
                  int ndim, float* out, cudaStream_t st, bool* handled) {
  *handled = false;
  const int64_t nmax = sy > sz ? sy : sz;
  if (nmax > 2048 || nmax >= 32768) return B2T_OK;
  *handled = true;
  const int64_t nrows = sy * sz;
  bool x_done = false;
  if (sizeof(T) == 4 && (sx % 4) == 0 && sx <= 1024 && ((uintptr_t)labels % 16) == 0 && ((uintptr_t)out % 16) == 0) {
    const unsigned blocks = (unsigned)((nrows + kWarpsPerBlock - 1) / kWarpsPerBlock);
    const uint32_t* l32 = reinterpret_cast<const uint32_t*>(labels);
    if (sx <= 128) edt_pass_x_v2_kernel<4><<<blocks, kWarpsPerBlock * 32, 0, st>>>(l32, out, (int)sx, nrows, wx, black_border);
    else if (sx <= 256) edt_pass_x_v2_kernel<8><<<blocks, kWarpsPerBlock * 32, 0, st>>>(l32, out, (int)sx, nrows, wx, black_border);
    else if (sx <= 512) edt_pass_x_v2_kernel<16><<<blocks, kWarpsPerBlock * 32, 0, st>>>(l32, out, (int)sx, nrows, wx, black_border);
    else edt_pass_x_v2_kernel<32><<<blocks, kWarpsPerBlock * 32, 0, st>>>(l32, out, (int)sx, nrows, wx, black_border);
    x_done = true;
  }
  if (!x_done) {
    const int64_t blocks = (nrows + kWarpsPerBlock - 1) / kWarpsPerBlock;
    edt_pass_x_kernel<T><<<(unsigned)blocks, kWarpsPerBlock * 32, 0, st>>>(labels, out, (int)sx, nrows, wx, black_border);
  }
  // column passes: v3 (shared-memory ring, edt_fh3.cuh) unless B2T_EDT_ALGO=2 asks for the v2 kernel
  if (edt_cfg().algo == 3 && nmax <= fh3::kMaxN) {
    bool ok;
    if (nmax <= 256) ok = edt_launch_fh3_passes<T, 256>(labels, sx, sy, sz, wy, wz, black_border, ndim, out, st);
    else if (nmax <= 512) ok = edt_launch_fh3_passes<T, 512>(labels, sx, sy, sz, wy, wz, black_border, ndim, out, st);
    else if (nmax <= 1024) ok = edt_launch_fh3_passes<T, 1024>(labels, sx, sy, sz, wy, wz, black_border, ndim, out, st);
    else ok = edt_launch_fh3_passes<T, 2048>(labels, sx, sy, sz, wy, wz, black_border, ndim, out, st);
    B2T_REQUIRE(ok, "b2t_edt: B2T_FH3 names a variant that is not compiled in");
    B2T_CUDA_TRY(cudaGetLastError());
    b2t_count_launches(ndim == 3 ? 3 : 2);
    return B2T_OK;
  }
  const dim3 gy((unsigned)b2t_ceil_div(sx, 128), (unsigned)sz), gz((unsigned)b2t_ceil_div(sx, 128), (unsigned)sy);
  static const int minb = []() { const char* e = getenv("B2T_FH_MINB"); return e ? atoi(e) : 8; }();
#define B2T_FH2(NM, MB)                                                                                               \
  do {                                                                                                                \
    edt_pass_col_fh_kernel<T, NM, MB><<<gy, 128, 0, st>>>(labels, out, (int)sy, sx, (int)sx, sx * sy, wy,            \
                                                         black_border, ndim == 2);                                    \
    if (ndim == 3)                                                                                                    \
      edt_pass_col_fh_kernel<T, NM, MB><<<gz, 128, 0, st>>>(labels, out, (int)sz, sx * sy, (int)sx, sx, wz,          \
                                                           black_border, 1);                                          \
  } while (0)
#define B2T_FH(NM)                          \
  do {                                      \
    if (minb >= 16) B2T_FH2(NM, 16);        \
    else if (minb <= 8) B2T_FH2(NM, 8);     \
    else B2T_FH2(NM, 12);                   \
  } while (0)
  if (nmax <= 256) B2T_FH(256);
  else if (nmax <= 512) B2T_FH(512);
  else if (nmax <= 1024) B2T_FH(1024);
  else B2T_FH(2048);
#undef B2T_FH
#undef B2T_FH2
  B2T_CUDA_TRY(cudaGetLastError());
  b2t_count_launches(ndim == 3 ? 3 : 2);
  return B2T_OK;
}

template <typename T>
int edt_launch(const T* labels, int64_t sx, int64_t sy, int64_t sz, float wx, float wy, float wz, int black_border,
               int ndim, float* out, cudaStream_t st) {
  const int64_t nrows = sy * sz;
  static const bool skip = []() { const char* e = getenv("B2T_EDT_SKIP"); return e ? (e[0] != '0') : true; }();
  {
    const int64_t blocks = (nrows + kWarpsPerBlock - 1) / kWarpsPerBlock;
    edt_pass_x_kernel<T><<<(unsigned)blocks, kWarpsPerBlock * 32, 0, st>>>(labels, out, (int)sx, nrows, wx,
                                                                          black_border);
  }
  {
    const size_t smem = ((size_t)sy + (sy + kBlk - 1) / kBlk) * 32 * sizeof(float);
    B2T_CUDA_TRY(cudaFuncSetAttribute(edt_pass_col_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B2T_CUDA_TRY(cudaFuncSetAttribute(edt_pass_col_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    dim3 grid((unsigned)b2t_ceil_div(sx, 32), (unsigned)sz);
    if (skip)
      edt_pass_col_kernel<T, true><<<grid, kWarpsPerBlock * 32, smem, st>>>(labels, out, (int)sy, sx, (int)sx, sx * sy, wy,
                                                                           black_border, ndim == 2);
    else
      edt_pass_col_kernel<T, false><<<grid, kWarpsPerBlock * 32, smem, st>>>(labels, out, (int)sy, sx, (int)sx, sx * sy, wy,
                                                                            black_border, ndim == 2);
  }
  if (ndim == 3) {
    const size_t smem = ((size_t)sz + (sz + kBlk - 1) / kBlk) * 32 * sizeof(float);
    dim3 grid((unsigned)b2t_ceil_div(sx, 32), (unsigned)sy);
    if (skip)
      edt_pass_col_kernel<T, true><<<grid, kWarpsPerBlock * 32, smem, st>>>(labels, out, (int)sz, sx * sy, (int)sx, sx, wz,
                                                                           black_border, 1);
    else
      edt_pass_col_kernel<T, false><<<grid, kWarpsPerBlock * 32, smem, st>>>(labels, out, (int)sz, sx * sy, (int)sx, sx, wz,
                                                                            black_border, 1);
  }
  B2T_CUDA_TRY(cudaGetLastError());
  b2t_count_launches(ndim == 3 ? 3 : 2);
  return B2T_OK;
}

}  // namespace

B2T_EXPORT int b2t_edt_config(int algo, int c, int minb, int r, int b) {
  B2T_REQUIRE(algo >= 0 && algo <= 3, "b2t_edt_config: algo must be 0 (keep), 1 (windowed), 2 (local-memory F-H) or 3 (ring F-H)");
  EdtCfg& cfg = edt_cfg();
  if (algo) cfg.algo = algo;
  if (c > 0) { cfg.c = c; cfg.minb = minb; cfg.r = r; cfg.b = b; cfg.ec = c; cfg.eminb = minb; cfg.er = r; cfg.eb = b; }
  return B2T_OK;
}

B2T_EXPORT int b2t_edt(const void* d_labels, int label_bytes, int64_t sx, int64_t sy, int64_t sz, float wx, float wy,
                       float wz, int black_border, int ndim, float* d_out, void* stream) {
  B2T_REQUIRE(d_labels && d_out, "b2t_edt: null pointer");
  B2T_REQUIRE(sx > 0 && sy > 0 && sz > 0, "b2t_edt: empty volume %lld x %lld x %lld", (long long)sx, (long long)sy,
              (long long)sz);
  B2T_REQUIRE(ndim == 2 || ndim == 3, "b2t_edt: ndim must be 2 or 3");
  B2T_REQUIRE(ndim == 3 || sz == 1, "b2t_edt: ndim=2 requires sz == 1");
  B2T_REQUIRE(sx <= 32 * kMaxGroups, "b2t_edt: sx > %d not supported", 32 * kMaxGroups);
  B2T_REQUIRE(sy * 144 <= 227 * 1024 && sz * 144 <= 227 * 1024, "b2t_edt: sy, sz > 1614 not supported");
  B2T_REQUIRE(sz <= 65535 && sy <= 65535, "b2t_edt: extent too large for grid.y");
  cudaStream_t st = (cudaStream_t)stream;
  black_border = black_border ? 1 : 0;
  if (edt_cfg().algo != 1) {
    bool handled = false;
    int rc = B2T_OK;
    switch (label_bytes) {
      case 1: rc = edt_launch_v2<uint8_t>((const uint8_t*)d_labels, sx, sy, sz, wx, wy, wz, black_border, ndim, d_out, st, &handled); break;
      case 2: rc = edt_launch_v2<uint16_t>((const uint16_t*)d_labels, sx, sy, sz, wx, wy, wz, black_border, ndim, d_out, st, &handled); break;
      case 4: rc = edt_launch_v2<uint32_t>((const uint32_t*)d_labels, sx, sy, sz, wx, wy, wz, black_border, ndim, d_out, st, &handled); break;
      case 8: rc = edt_launch_v2<unsigned long long>((const unsigned long long*)d_labels, sx, sy, sz, wx, wy, wz, black_border, ndim, d_out, st, &handled); break;
      default: b2t_set_error("b2t_edt: label_bytes must be 1, 2, 4 or 8 (got %d)", label_bytes); return B2T_ERR_ARG;
    }
    if (handled || rc != B2T_OK) return rc;
  }
  switch (label_bytes) {
    case 1: return edt_launch<uint8_t>((const uint8_t*)d_labels, sx, sy, sz, wx, wy, wz, black_border, ndim, d_out, st);
    case 2: return edt_launch<uint16_t>((const uint16_t*)d_labels, sx, sy, sz, wx, wy, wz, black_border, ndim, d_out, st);
    case 4: return edt_launch<uint32_t>((const uint32_t*)d_labels, sx, sy, sz, wx, wy, wz, black_border, ndim, d_out, st);
    case 8: return edt_launch<unsigned long long>((const unsigned long long*)d_labels, sx, sy, sz, wx, wy, wz, black_border, ndim, d_out, st);
    default: b2t_set_error("b2t_edt: label_bytes must be 1, 2, 4 or 8 (got %d)", label_bytes); return B2T_ERR_ARG;
  }
}

// ------------------------------------------------------------------------------------------------
// hybrid pass launcher (uint32 labels, integer anisotropy): stencil kernel over all columns, then the envelope
// kernel over the blocks the stencil flagged; out of place (fin -> fout).
// ------------------------------------------------------------------------------------------------
static bool edt_launch_hybrid_pass(const uint32_t* labels, const float* fin, float* fout, int n, int64_t cstride, int64_t sx,
                                   int64_t nouter, int64_t ostride, float w, int black_border, int last, int W,
                                   bool write_bg, unsigned long long* flags, cudaStream_t st) {
  const EdtCfg c = edt_cfg();
  const dim3 grid((unsigned)b2t_ceil_div(sx, 128), (unsigned)nouter);
  const int ntx = b2t_ceil_div(sx, 32);
  bool done = false;
#define B2T_ST_GO(W_, WR_, PF_, MB_)                                                                                    \
  if (!done && W == W_ && c.hwr == WR_ && c.hpf == PF_ && c.hminb == MB_) {                                              \
    if (write_bg)                                                                                                         \
      edt_pass_col_stencil_kernel<uint32_t, W_, (WR_ < W_ ? WR_ : W_), PF_, MB_, true><<<grid, 128, 0, st>>>(            \
          labels, fin, fout, n, cstride, (int)sx, ostride, w, black_border, last, flags, ntx);                           \
    else                                                                                                                  \
      edt_pass_col_stencil_kernel<uint32_t, W_, (WR_ < W_ ? WR_ : W_), PF_, MB_, false><<<grid, 128, 0, st>>>(           \
          labels, fin, fout, n, cstride, (int)sx, ostride, w, black_border, last, flags, ntx);                           \
    done = true;                                                                                                          \
  }
  // stencil_column_v2 (b2t_edt_config_roles(., 1)): the two radii the anisotropies in use ask for
  const bool v2ok = c.sopt && c.hwr == 4 && c.hpf == 11 && c.hminb == 8 && (W == 4 || W == 10) &&
                    ((int64_t)n + 64) * cstride < (int64_t)0xffffffffll;
#define B2T_ST2_GO(W_)                                                                                                    \
  if (!done && v2ok && W == W_) {                                                                                         \
    if (write_bg)                                                                                                         \
      edt_pass_col_stencil_kernel<uint32_t, W_, 4, 11, 8, true, 1><<<grid, 128, 0, st>>>(                                 \
          labels, fin, fout, n, cstride, (int)sx, ostride, w, black_border, last, flags, ntx);                           \
    else                                                                                                                  \
      edt_pass_col_stencil_kernel<uint32_t, W_, 4, 11, 8, false, 1><<<grid, 128, 0, st>>>(                                \
          labels, fin, fout, n, cstride, (int)sx, ostride, w, black_border, last, flags, ntx);                           \
    done = true;                                                                                                          \
  }
  B2T_ST2_GO(4) B2T_ST2_GO(10)
#undef B2T_ST2_GO
  // tap radius (4 .. 12), register-window radius, prefetch depth, min blocks per SM
  B2T_ST_GO(4, 4, 11, 8) B2T_ST_GO(6, 4, 11, 8) B2T_ST_GO(8, 4, 11, 8) B2T_ST_GO(10, 4, 11, 8)
  B2T_ST_GO(4, 4, 7, 8) B2T_ST_GO(10, 4, 7, 8) B2T_ST_GO(12, 4, 7, 8)
#undef B2T_ST_GO
  if (!done) return false;
  done = false;
  const dim3 rgrid((unsigned)ntx, (unsigned)nouter);
#define B2T_FR_GO(NM_, C_, MB_, R_, B_)                                                                                   \
  if (!done && n <= NM_ && c.ec == C_ && c.eminb == MB_ && c.er == R_ && c.eb == B_) {                                    \
    edt_pass_col_fh3_range_kernel<uint32_t, C_, NM_, 4 * MB_, R_, B_><<<rgrid, 32, 0, st>>>(                             \
        labels, fin, fout, n, cstride, (int)sx, ostride, w, black_border, last, flags, ntx);                             \
    done = true;                                                                                                          \
  }
#define B2T_FR4_GO(NM_, QP_, PP_)                                                                                         \
  if (!done && n <= NM_ && c.eqp == QP_ && c.epp == PP_ && c.ec == 32 && c.eminb == 4 && c.er == 32 && c.eb == 8) {        \
    edt_pass_col_fh3_range_kernel<uint32_t, 32, NM_, 16, 32, 8, QP_, PP_><<<rgrid, 32, 0, st>>>(                          \
        labels, fin, fout, n, cstride, (int)sx, ostride, w, black_border, last, flags, ntx);                             \
    done = true;                                                                                                          \
  }
#define B2T_FR4_ALL(QP_, PP_) B2T_FR4_GO(256, QP_, PP_) B2T_FR4_GO(512, QP_, PP_) B2T_FR4_GO(1024, QP_, PP_) B2T_FR4_GO(2048, QP_, PP_)
  B2T_FR4_ALL(4, 0) B2T_FR4_ALL(1, 1) B2T_FR4_ALL(4, 1)
#undef B2T_FR4_ALL
#undef B2T_FR4_GO
#define B2T_FR_ALL(C_, MB_, R_, B_) B2T_FR_GO(256, C_, MB_, R_, B_) B2T_FR_GO(512, C_, MB_, R_, B_) B2T_FR_GO(1024, C_, MB_, R_, B_) B2T_FR_GO(2048, C_, MB_, R_, B_)
  B2T_FR_ALL(32, 4, 32, 8) B2T_FR_ALL(16, 6, 32, 4) B2T_FR_ALL(16, 8, 32, 4)
#undef B2T_FR_ALL
#undef B2T_FR_GO
  return done;
}

// roles form of one column pass: the roles kernel (envelope over the predicted blocks next to the stencil over the
// rest), then the envelope kernel over what the stencil flagged outside the prediction.
static bool edt_launch_roles_pass(const uint32_t* labels, const float* fin, float* fout, int n, int64_t cstride, int64_t sx,
                                  int64_t nouter, int64_t ostride, float w, int black_border, int last, int W, bool write_bg,
                                  const unsigned long long* pred, unsigned long long* resid, unsigned long long* next,
                                  float thr_next, cudaStream_t st) {
  const EdtCfg c = edt_cfg();
  const dim3 grid((unsigned)b2t_ceil_div(sx, 128), (unsigned)nouter, 2);
  const int ntx = b2t_ceil_div(sx, 32);
  const bool nx_ = next != nullptr;
  bool done = false;
  const bool v2 = c.sopt && ((int64_t)n + 64) * cstride < (int64_t)0xffffffffll;
#define B2T_RO_GO(W_, BG_, NX_, NM_)                                                                                      \
  if (!done && W == W_ && write_bg == BG_ && nx_ == NX_ && n <= NM_) {                                                    \
    if (v2)                                                                                                               \
      edt_pass_col_roles_kernel<uint32_t, W_, 4, 11, 8, BG_, NX_, NM_, 32, 4, 1><<<grid, 128, 0, st>>>(                   \
          labels, fin, fout, n, cstride, (int)sx, ostride, w, black_border, last, pred, resid, next, thr_next, ntx);      \
    else                                                                                                                  \
      edt_pass_col_roles_kernel<uint32_t, W_, 4, 11, 8, BG_, NX_, NM_, 32, 4, 0><<<grid, 128, 0, st>>>(                   \
          labels, fin, fout, n, cstride, (int)sx, ostride, w, black_border, last, pred, resid, next, thr_next, ntx);      \
    done = true;                                                                                                          \
  }
#define B2T_RO_ALL(W_) B2T_RO_GO(W_, true, true, 512) B2T_RO_GO(W_, true, true, 2048) B2T_RO_GO(W_, false, false, 512) \
  B2T_RO_GO(W_, false, false, 2048) B2T_RO_GO(W_, true, false, 512) B2T_RO_GO(W_, true, false, 2048)
  B2T_RO_ALL(4) B2T_RO_ALL(10)
#undef B2T_RO_ALL
#undef B2T_RO_GO
  if (!done) return false;
  done = false;
  const dim3 rgrid((unsigned)ntx, (unsigned)nouter);
#define B2T_FR_GO(NM_, C_, MB_, R_, B_)                                                                                   \
  if (!done && n <= NM_ && c.ec == C_ && c.eminb == MB_ && c.er == R_ && c.eb == B_) {                                    \
    edt_pass_col_fh3_range_kernel<uint32_t, C_, NM_, 4 * MB_, R_, B_><<<rgrid, 32, 0, st>>>(                             \
        labels, fin, fout, n, cstride, (int)sx, ostride, w, black_border, last, resid, ntx);                             \
    done = true;                                                                                                          \
  }
#define B2T_FR4_GO(NM_, QP_, PP_)                                                                                         \
  if (!done && n <= NM_ && c.eqp == QP_ && c.epp == PP_ && c.ec == 32 && c.eminb == 4 && c.er == 32 && c.eb == 8) {        \
    edt_pass_col_fh3_range_kernel<uint32_t, 32, NM_, 16, 32, 8, QP_, PP_><<<rgrid, 32, 0, st>>>(                          \
        labels, fin, fout, n, cstride, (int)sx, ostride, w, black_border, last, resid, ntx);                             \
    done = true;                                                                                                          \
  }
#define B2T_FR4_ALL(QP_, PP_) B2T_FR4_GO(256, QP_, PP_) B2T_FR4_GO(512, QP_, PP_) B2T_FR4_GO(1024, QP_, PP_) B2T_FR4_GO(2048, QP_, PP_)
  B2T_FR4_ALL(4, 0) B2T_FR4_ALL(1, 1) B2T_FR4_ALL(4, 1)
#undef B2T_FR4_ALL
#undef B2T_FR4_GO
#define B2T_FR_ALL(C_, MB_, R_, B_) B2T_FR_GO(256, C_, MB_, R_, B_) B2T_FR_GO(512, C_, MB_, R_, B_) B2T_FR_GO(1024, C_, MB_, R_, B_) B2T_FR_GO(2048, C_, MB_, R_, B_)
  B2T_FR_ALL(32, 4, 32, 8) B2T_FR_ALL(16, 6, 32, 4) B2T_FR_ALL(16, 8, 32, 4)
#undef B2T_FR_ALL
#undef B2T_FR_GO
  return done;
}

static bool edt_is_small_int(float w) { return w >= 1.0f && w <= 2048.0f && w == rintf(w); }

B2T_EXPORT size_t b2t_edt_workspace_bytes(int64_t sx, int64_t sy, int64_t sz) {
  if (sx <= 0 || sy <= 0 || sz <= 0) return 0;
  const size_t ntx = (size_t)b2t_ceil_div(sx, 32);
  // scratch volume + per pass one flag word per (32-column tile, outer index), twice: flagged by the stencil, predicted
  return (size_t)sx * sy * sz * sizeof(float) + 2 * (ntx * (size_t)sz + ntx * (size_t)sy) * sizeof(unsigned long long);
}

B2T_EXPORT int b2t_edt_ws(const void* d_labels, int label_bytes, int64_t sx, int64_t sy, int64_t sz, float wx, float wy,
                          float wz, int black_border, int ndim, float* d_out, void* d_workspace, size_t workspace_bytes,
                          void* stream) {
  const EdtCfg c = edt_cfg();
  const int64_t nmax = sy > sz ? sy : sz;
  // Tap radius of a pass: a process r voxels thick along x needs taps out to r * wx / w rows, so the radius scales
  // with wx / w (10 rows at w = wx; 4 in the z pass of a 16 x 16 x 40 nm volume).  Any radius gives the same result.
  auto auto_radius = [&](float w) { const int r = 2 * (int)lrintf(5.0f * wx / w); return r < 4 ? 4 : (r > 10 ? 10 : r); };
  const int hwy = c.hwy > 0 ? c.hwy : auto_radius(wy), hwz = c.hwy > 0 ? c.hwz : auto_radius(wz);
  // The stencil is exact -- and therefore bit-identical to the envelope -- when every value a thin voxel can take is an
  // integer below 2^24: integer anisotropy, and thresholds w^2 (W+1)^2 below 2^24.
  const bool hybrid_ok =
      c.hybrid && c.algo == 3 && d_labels && d_out && d_workspace && label_bytes == 4 && (ndim == 2 || ndim == 3) &&
      (ndim == 3 || sz == 1) && sx > 0 && sy > 0 && sz > 0 && (sx % 4) == 0 && sx <= 1024 && nmax <= fh3::kMaxN &&
      sy <= 65535 && sz <= 65535 && ((uintptr_t)d_labels % 16) == 0 && ((uintptr_t)d_out % 16) == 0 &&
      ((uintptr_t)d_workspace % 16) == 0 && workspace_bytes >= b2t_edt_workspace_bytes(sx, sy, sz) &&
      edt_is_small_int(wx) && edt_is_small_int(wy) && edt_is_small_int(wz) && wy * (float)(hwy + 1) < 4096.0f &&
      wz * (float)(hwz + 1) < 4096.0f;
  if (!hybrid_ok) return b2t_edt(d_labels, label_bytes, sx, sy, sz, wx, wy, wz, black_border, ndim, d_out, stream);
  cudaStream_t st = (cudaStream_t)stream;
  black_border = black_border ? 1 : 0;
  const uint32_t* labels = (const uint32_t*)d_labels;
  const int64_t V = sx * sy * sz;
  const size_t ntx = (size_t)b2t_ceil_div(sx, 32);
  float* ws_f = (float*)d_workspace;
  unsigned long long* flags_y = (unsigned long long*)(ws_f + V);
  unsigned long long* flags_z = flags_y + ntx * (size_t)sz;
  unsigned long long* pred_y = flags_z + ntx * (size_t)sy;
  unsigned long long* pred_z = pred_y + ntx * (size_t)sz;
  const bool roles = c.roles != 0 && (hwy == 4 || hwy == 10) && (ndim == 2 || hwz == 4 || hwz == 10) && sx <= 512;
  B2T_CUDA_TRY(cudaMemsetAsync(flags_y, 0, (roles ? 2 : 1) * (ntx * (size_t)sz + ntx * (size_t)sy) * sizeof(unsigned long long), st));
  float* a = (ndim == 3) ? d_out : ws_f;   // x -> a, y -> b, z -> a: the result ends in d_out either way
  float* b = (ndim == 3) ? ws_f : d_out;
  if (roles) {
    const int64_t nrows = sy * sz;
    const unsigned blocks = (unsigned)((nrows + kWarpsPerBlock - 1) / kWarpsPerBlock);
    // pscale > 1 predicts fewer blocks (experiments; what it misses is flagged by the stencil and done by the residual launch)
    const float thr_y = c.pscale * wy * wy * (float)((hwy + 1) * (hwy + 1)), thr_z = c.pscale * wz * wz * (float)((hwz + 1) * (hwz + 1));
    if (sx <= 128) edt_pass_x_v2_kernel<4, true><<<blocks, kWarpsPerBlock * 32, 0, st>>>(labels, a, (int)sx, nrows, wx, black_border, pred_y, thr_y, (int)sy, (int)ntx);
    else if (sx <= 256) edt_pass_x_v2_kernel<8, true><<<blocks, kWarpsPerBlock * 32, 0, st>>>(labels, a, (int)sx, nrows, wx, black_border, pred_y, thr_y, (int)sy, (int)ntx);
    else edt_pass_x_v2_kernel<16, true><<<blocks, kWarpsPerBlock * 32, 0, st>>>(labels, a, (int)sx, nrows, wx, black_border, pred_y, thr_y, (int)sy, (int)ntx);
    bool ok = edt_launch_roles_pass(labels, a, b, (int)sy, sx, sx, sz, sx * sy, wy, black_border, ndim == 2, hwy, true, pred_y,
                                    flags_y, ndim == 3 ? pred_z : nullptr, thr_z, st);
    if (ok && ndim == 3)
      ok = edt_launch_roles_pass(labels, b, a, (int)sz, sx * sy, sx, sy, sx, wz, black_border, 1, hwz, false, pred_z, flags_z,
                                 nullptr, 0.0f, st);
    B2T_REQUIRE(ok, "b2t_edt_ws: the configured roles / envelope variant is not compiled in");
    B2T_CUDA_TRY(cudaGetLastError());
    b2t_count_launches(ndim == 3 ? 5 : 3);
    return B2T_OK;
  }
  {
    const int64_t nrows = sy * sz;
    const unsigned blocks = (unsigned)((nrows + kWarpsPerBlock - 1) / kWarpsPerBlock);
    // x pass with TMA-staged tiles (edt_xtma.cuh) when asked for and the row length is one or two boxes; else the v2 kernel
    if (c.xtma && xtma::launch(labels, a, sx, nrows, wx, black_border, st, c.xtma)) { /* launched */ }
    else if (sx <= 128) edt_pass_x_v2_kernel<4><<<blocks, kWarpsPerBlock * 32, 0, st>>>(labels, a, (int)sx, nrows, wx, black_border);
    else if (sx <= 256) edt_pass_x_v2_kernel<8><<<blocks, kWarpsPerBlock * 32, 0, st>>>(labels, a, (int)sx, nrows, wx, black_border);
    else if (sx <= 512) edt_pass_x_v2_kernel<16><<<blocks, kWarpsPerBlock * 32, 0, st>>>(labels, a, (int)sx, nrows, wx, black_border);
    else edt_pass_x_v2_kernel<32><<<blocks, kWarpsPerBlock * 32, 0, st>>>(labels, a, (int)sx, nrows, wx, black_border);
  }
  bool ok = edt_launch_hybrid_pass(labels, a, b, (int)sy, sx, sx, sz, sx * sy, wy, black_border, ndim == 2, hwy, true, flags_y, st);
  if (ok && ndim == 3)
    ok = edt_launch_hybrid_pass(labels, b, a, (int)sz, sx * sy, sx, sy, sx, wz, black_border, 1, hwz, false, flags_z, st);
  B2T_REQUIRE(ok, "b2t_edt_ws: the configured stencil / envelope variant is not compiled in");
  B2T_CUDA_TRY(cudaGetLastError());
  b2t_count_launches(ndim == 3 ? 5 : 3);
  return B2T_OK;
}

// x pass of b2t_edt_ws: 1 = TMA-staged tiles (edt_xtma.cuh; rows of 256 or 512 labels), 0 = the register-only v2 kernel.
// Returns the number of TMA x-pass launches so far (tma < 0: query only).
B2T_EXPORT long long b2t_edt_config_xpass(int tma) {
  if (tma >= 0) edt_cfg().xtma = tma > 2 ? 1 : tma;      // 1: a tile per CTA; 2: persistent CTAs with a two-stage ring
  return (long long)xtma::launches();
}

B2T_EXPORT int b2t_edt_config_envelope(int query_prefetch, int pop_ahead) {
  B2T_REQUIRE(query_prefetch == 1 || query_prefetch == 4, "b2t_edt_config_envelope: query_prefetch must be 1 or 4");
  edt_cfg().eqp = query_prefetch;
  edt_cfg().epp = pop_ahead ? 1 : 0;
  return B2T_OK;
}

B2T_EXPORT int b2t_edt_config_roles(int enable, int stencil_v2, float predict_scale) {
  edt_cfg().roles = enable ? 1 : 0;
  edt_cfg().sopt = stencil_v2 ? 1 : 0;
  edt_cfg().pscale = predict_scale >= 1.0f ? predict_scale : 1.0f;
  return B2T_OK;
}

B2T_EXPORT int b2t_edt_config_hybrid(int enable, int wy, int wz, int wr, int pf, int minb) {
  EdtCfg& cfg = edt_cfg();
  cfg.hybrid = enable ? 1 : 0;
  if (wy > 0) { cfg.hwy = wy; cfg.hwz = wz; }
  else if (wy == 0) { cfg.hwy = 0; cfg.hwz = 0; }   // radii from the anisotropy
  if (wr > 0) { cfg.hwr = wr; cfg.hpf = pf; cfg.hminb = minb; }
  return B2T_OK;
}

