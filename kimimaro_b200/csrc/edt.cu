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
  switch (label_bytes) {
    case 1: return edt_launch<uint8_t>((const uint8_t*)d_labels, sx, sy, sz, wx, wy, wz, black_border, ndim, d_out, st);
    case 2: return edt_launch<uint16_t>((const uint16_t*)d_labels, sx, sy, sz, wx, wy, wz, black_border, ndim, d_out, st);
    case 4: return edt_launch<uint32_t>((const uint32_t*)d_labels, sx, sy, sz, wx, wy, wz, black_border, ndim, d_out, st);
    case 8: return edt_launch<unsigned long long>((const unsigned long long*)d_labels, sx, sy, sz, wx, wy, wz, black_border, ndim, d_out, st);
    default: b2t_set_error("b2t_edt: label_bytes must be 1, 2, 4 or 8 (got %d)", label_bytes); return B2T_ERR_ARG;
  }
}
