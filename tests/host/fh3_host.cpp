// Test infrastructure: runs the K1 v3 column-pass body (kimimaro_b200/csrc/edt_fh3.cuh, the very code the
// CUDA kernel instantiates) on the CPU, one emulated thread per column, so that the CPU suite can check it
// bit for bit against the oracle's EDT.  Compiled by tests/test_edt_fh3_host.py with g++; never shipped.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../kimimaro_b200/csrc/edt_fh3.cuh"

static int g_stencil_v2 = 0;   // 1: the stencil passes run fh3::stencil_column_v2
extern "C" void fh3_host_set_stencil(int v2) { g_stencil_v2 = v2; }
static int g_query_prefetch = 1;   // 4: write-out lookahead 4 + pop-ahead (QP = 4, PP = 1 of fh3::column_range); 5: pop-ahead only
extern "C" void fh3_host_set_query_prefetch(int qp) { g_query_prefetch = qp; }

namespace {

struct HostCtx {
  // ring slots [0, C) emulate shared memory, l_* the local-memory backing (indexed by entry)
  float sv[64]; float sh[64]; float sz[64];
  float* lv; float* lh; float* lz;
  long spills = 0, lloads = 0, maxk = 0;
  float mul(float a, float b) const { return a * b; }
  float add(float a, float b) const { return a + b; }
  float sub(float a, float b) const { return a - b; }
  float div(float a, float b) const { return a / b; }
  float sqrt(float a) const { return sqrtf(a); }
  float fmin(float a, float b) const { return fminf(a, b); }
  template <typename T> T ld_label(const T* p) const { return *p; }
  float ld_f(const float* p) const { return *p; }
  void st_f(float* p, float v) const { *p = v; }
  void s_st(int s, float v, float h, float z) { sv[s] = v; sh[s] = h; sz[s] = z; }
  void s_st_z(int s, float z) { sz[s] = z; }
  float s_ld_v(int s) const { return sv[s]; }
  float s_ld_h(int s) const { return sh[s]; }
  float s_ld_z(int s) const { return sz[s]; }
  void l_st(int k, float v, float h, float z) { spills++; if (k > maxk) maxk = k; lv[k] = v; lh[k] = h; lz[k] = z; }
  void l_st_z(int k, float z) { lz[k] = z; }
  float l_ld_v(int k) { lloads++; return lv[k]; }
  float l_ld_h(int k) { return lh[k]; }
  float l_ld_z(int k) { return lz[k]; }
  int wmin(int x) const { return x; }
  int wmax(int x) const { return x; }
  float wmaxf(float x) const { return x; }
  bool any(bool p) const { return p; }
  // hybrid pass: 32-row blocks of the tile of 32 columns that the stencil could not finish
  // prefetch rings of the stencil (cp.async on the device): copies complete immediately here
  float rf[fh3::kRingF]; uint32_t rl[fh3::kRingL];
  template <typename T> void ring_fetch(int foff, int loff, const T* lp, const float* fp) {
    rl[loff / fh3::kRingSlotBytes] = (uint32_t)*lp; rf[foff / fh3::kRingSlotBytes] = *fp;
  }
  void ring_fetch_f(int foff, const float* fp) { rf[foff / fh3::kRingSlotBytes] = *fp; }
  void ring_put(int foff, float f) { rf[foff / fh3::kRingSlotBytes] = f; }
  template <int N> void ring_wait() const {}
  float ring_f(int foff) const { return rf[foff / fh3::kRingSlotBytes]; }
  template <typename T> T ring_l(int loff) const { return (T)rl[loff / fh3::kRingSlotBytes]; }
  uint64_t* flag = nullptr;
  void note_row(int row) { if (flag) *flag |= 1ull << (row >> 5); }
  // roles form: prediction words of the next pass
  uint64_t* nflag = nullptr; int64_t nstride = 0; uint64_t nbit = 0;
  void note_next(int row) { if (nflag) nflag[(int64_t)row * nstride] |= nbit; }
};

template <typename T, int C, int R, int B>
void pass(const T* labels, float* f, int n, int64_t cstride, int64_t ncols_inner, int64_t inner_stride,
          int64_t ncols_outer, int64_t outer_stride, float w, int bb, int last, long* stats) {
  static_assert(C <= 64, "host ring");
  HostCtx cx;
  cx.lv = (float*)malloc(sizeof(float) * (n + 4)); cx.lh = (float*)malloc(sizeof(float) * (n + 4)); cx.lz = (float*)malloc(sizeof(float) * (n + 4));
  // poison: a read of an entry that was never written must not go unnoticed
  for (int i = 0; i < n + 4; i++) { cx.lv[i] = NAN; cx.lh[i] = NAN; cx.lz[i] = NAN; }
  for (int64_t o = 0; o < ncols_outer; o++)
    for (int64_t x = 0; x < ncols_inner; x++) {
      const int64_t base = o * outer_stride + x * inner_stride;
      for (int s = 0; s < 64; s++) { cx.sv[s] = NAN; cx.sh[s] = NAN; cx.sz[s] = NAN; }
      fh3::column<T, C, R, B>(cx, labels + base, f + base, n, cstride, w, bb != 0, last != 0, true);
    }
  if (stats) { stats[0] += cx.spills; if (cx.maxk > stats[1]) stats[1] = cx.maxk; stats[2] += cx.lloads; }
  free(cx.lv); free(cx.lh); free(cx.lz);
}

// x pass exactly like the CUDA x-pass kernels (edt.cu): nearest label change on either side, (d*w)^2
template <typename T>
void pass_x(const T* lab, float* out, int64_t sx, int64_t nrows, float w, int bb) {
  for (int64_t r = 0; r < nrows; r++) {
    const T* l = lab + r * sx;
    float* o = out + r * sx;
    int64_t s = 0;
    for (int64_t p = 0; p < sx; p++) {
      if (p > 0 && l[p] != l[p - 1]) s = p;
      int64_t nx = p + 1;
      while (nx < sx && l[nx] == l[p]) nx++;
      float v = 0.0f;
      if (l[p] != 0) {
        const bool lok = (s > 0) || bb, rok = (nx < sx) || bb;
        if (lok || rok) {
          const int64_t dl = p - s + 1, dr = nx - p;
          const int64_t d = lok ? (rok ? (dl < dr ? dl : dr) : dl) : dr;
          const float fd = (float)d * w;
          v = fd * fd;
        } else {
          v = INFINITY;
        }
      }
      o[p] = v;
    }
  }
}

template <int C, int R, int B>
void run(const uint32_t* labels, int64_t sx, int64_t sy, int64_t sz, float wx, float wy, float wz, int bb, int ndim,
         float* out, long* stats) {
  pass_x<uint32_t>(labels, out, sx, sy * sz, wx, bb);
  pass<uint32_t, C, R, B>(labels, out, (int)sy, sx, sx, 1, sz, sx * sy, wy, bb, ndim == 2, stats);
  if (ndim == 3) pass<uint32_t, C, R, B>(labels, out, (int)sz, sx * sy, sx, 1, sy, sx, wz, bb, 1, stats);
}

}  // namespace

// variant = (C, R, B): 0 = (16,32,4)  1 = (4,16,4)  2 = (2,8,4)  3 = (1,32,8)  4 = (32,32,8)  5 = (8,32,4)  6 = (16,16,4)
extern "C" int fh3_host_edt(const uint32_t* labels, int64_t sx, int64_t sy, int64_t sz, float wx, float wy, float wz,
                            int bb, int ndim, int variant, float* out, long* stats) {
  if (sy > fh3::kMaxN || sz > fh3::kMaxN) return -1;
  switch (variant) {
    case 0: run<16, 32, 4>(labels, sx, sy, sz, wx, wy, wz, bb, ndim, out, stats); break;
    case 1: run<4, 16, 4>(labels, sx, sy, sz, wx, wy, wz, bb, ndim, out, stats); break;
    case 2: run<2, 8, 4>(labels, sx, sy, sz, wx, wy, wz, bb, ndim, out, stats); break;
    case 3: run<1, 32, 8>(labels, sx, sy, sz, wx, wy, wz, bb, ndim, out, stats); break;
    case 4: run<32, 32, 8>(labels, sx, sy, sz, wx, wy, wz, bb, ndim, out, stats); break;
    case 5: run<8, 32, 4>(labels, sx, sy, sz, wx, wy, wz, bb, ndim, out, stats); break;
    case 6: run<16, 16, 4>(labels, sx, sy, sz, wx, wy, wz, bb, ndim, out, stats); break;
    default: return -2;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Hybrid pipeline (b2t_edt_ws): plain x pass, then per pass the stencil over every column (it flags the
// 32-row blocks of each 32-column tile it could not finish) and the envelope over the flagged blocks,
// extended to complete runs, out of place.
// ---------------------------------------------------------------------------------------------------------
namespace {

template <int W, int WR, int PF, int C, int R, int B, bool WRITE_BG>
void hybrid_pass(const uint32_t* labels, const float* fin, float* fout, int n, int64_t cstride, int64_t sx, int64_t nouter,
                 int64_t ostride, float w, int bb, int last, int64_t ntx, long* stats) {
  HostCtx cx;
  cx.lv = (float*)malloc(sizeof(float) * (n + 4)); cx.lh = (float*)malloc(sizeof(float) * (n + 4)); cx.lz = (float*)malloc(sizeof(float) * (n + 4));
  uint64_t* flags = (uint64_t*)calloc(ntx * nouter, sizeof(uint64_t));
  for (int64_t o = 0; o < nouter; o++)
    for (int64_t x = 0; x < sx; x++) {
      const int64_t base = o * ostride + x;
      cx.flag = flags + o * ntx + (x >> 5);
      if (g_stencil_v2)
        fh3::stencil_column_v2<uint32_t, W, WR, PF, WRITE_BG>(cx, labels + base, fin + base, fout + base, n, cstride, w, bb != 0, last != 0, true);
      else
        fh3::stencil_column<uint32_t, W, WR, PF, WRITE_BG>(cx, labels + base, fin + base, fout + base, n, cstride, w, bb != 0, last != 0, true);
    }
  cx.flag = nullptr;
  if (stats) for (int64_t i = 0; i < ntx * nouter; i++) stats[WRITE_BG ? 6 : 7] += __builtin_popcountll(flags[i]);
  for (int64_t o = 0; o < nouter; o++)
    for (int64_t x = 0; x < sx; x++) {
      const int64_t base = o * ostride + x;
      uint64_t m = flags[o * ntx + (x >> 5)];
      while (m) {                                   // maximal groups of consecutive flagged blocks
        const int b0 = __builtin_ctzll(m);
        int b1 = b0;
        while (b1 + 1 < 64 && ((m >> (b1 + 1)) & 1)) b1++;
        m &= (b1 == 63) ? 0ull : (~0ull << (b1 + 1));
        const int rlo = 32 * b0, rhi = (32 * b1 + 31 < n - 1) ? 32 * b1 + 31 : n - 1;
        int own_lo, own_hi;
        fh3::extend_to_runs<uint32_t>(cx, labels + base, n, cstride, true, rlo, rhi, own_lo, own_hi);
        for (int q = 0; q < 64; q++) { cx.sv[q] = NAN; cx.sh[q] = NAN; cx.sz[q] = NAN; }
        if (g_query_prefetch == 4)
          fh3::column_range<uint32_t, C, R, B, true, false, 4, 1>(cx, labels + base, fin + base, fout + base, n, cstride, w, bb != 0, last != 0, true,
                                                                  own_lo, own_hi, own_lo, own_hi);
        else if (g_query_prefetch == 5)
          fh3::column_range<uint32_t, C, R, B, true, false, 1, 1>(cx, labels + base, fin + base, fout + base, n, cstride, w, bb != 0, last != 0, true,
                                                                  own_lo, own_hi, own_lo, own_hi);
        else
          fh3::column_range<uint32_t, C, R, B, true>(cx, labels + base, fin + base, fout + base, n, cstride, w, bb != 0, last != 0, true,
                                                     own_lo, own_hi, own_lo, own_hi);
        if (stats) stats[3] += rhi - rlo + 1;
      }
    }
  if (stats) stats[0] += cx.spills;
  free(flags); free(cx.lv); free(cx.lh); free(cx.lz);
}

template <int WY, int WZ, int WR, int PF>
void run_hybrid(const uint32_t* labels, int64_t sx, int64_t sy, int64_t sz, float wx, float wy, float wz, int bb, int ndim,
                float* out, long* stats) {
  const int64_t V = sx * sy * sz, ntx = (sx + 31) / 32;
  float* ws = (float*)malloc(sizeof(float) * V);
  for (int64_t i = 0; i < V; i++) ws[i] = NAN;           // every voxel must be written by someone
  for (int64_t i = 0; i < V; i++) out[i] = NAN;
  float* a = (ndim == 3) ? out : ws;     // x -> a, y -> b, z -> a
  float* b = (ndim == 3) ? ws : out;
  pass_x<uint32_t>(labels, a, sx, sy * sz, wx, bb);
  hybrid_pass<WY, (WR < WY ? WR : WY), PF, 4, 16, 4, true>(labels, a, b, (int)sy, sx, sx, sz, sx * sy, wy, bb, ndim == 2, ntx, stats);
  if (ndim == 3) hybrid_pass<WZ, (WR < WZ ? WR : WZ), PF, 4, 16, 4, false>(labels, b, a, (int)sz, sx * sy, sx, sy, sx, wz, bb, 1, ntx, stats);
  free(ws);
}

}  // namespace

// variant = tap radii (y, z), register-window radius, prefetch depth:
//   0 = (10, 4) wr 4 pf 11   1 = (4, 4) wr 4 pf 4   2 = (12, 8) wr 5 pf 6   3 = (2, 1) wr 1 pf 1   4 = (8, 6) wr 8 pf 15
extern "C" int fh3_host_edt_hybrid(const uint32_t* labels, int64_t sx, int64_t sy, int64_t sz, float wx, float wy, float wz,
                                   int bb, int ndim, int variant, float* out, long* stats) {
  if (sy > fh3::kMaxN || sz > fh3::kMaxN) return -1;
  switch (variant) {
    case 0: run_hybrid<10, 4, 4, 11>(labels, sx, sy, sz, wx, wy, wz, bb, ndim, out, stats); break;
    case 1: run_hybrid<4, 4, 4, 4>(labels, sx, sy, sz, wx, wy, wz, bb, ndim, out, stats); break;
    case 2: run_hybrid<12, 8, 5, 6>(labels, sx, sy, sz, wx, wy, wz, bb, ndim, out, stats); break;
    case 3: run_hybrid<2, 1, 1, 1>(labels, sx, sy, sz, wx, wy, wz, bb, ndim, out, stats); break;
    case 4: run_hybrid<8, 6, 8, 15>(labels, sx, sy, sz, wx, wy, wz, bb, ndim, out, stats); break;
    default: return -2;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Roles form of the hybrid pipeline (b2t_edt_config_roles): the previous pass predicts the blocks the stencil
// cannot finish; per pass an envelope role over the predicted blocks and a stencil role that skips them run
// concurrently on the device (here: in either order, `order`), then the envelope over what the stencil flagged.
// ---------------------------------------------------------------------------------------------------------
namespace {

// like edt_pass_x_v2_kernel<SEG, true>: bit (y >> 5) of pred[z * ntx + tile] when a value of the tile's row exceeds thr
void predict_from_x(const float* a, int64_t sx, int64_t sy, int64_t sz, float thr, uint64_t* pred, int64_t ntx) {
  for (int64_t z = 0; z < sz; z++)
    for (int64_t y = 0; y < sy; y++)
      for (int64_t x = 0; x < sx; x++)
        if (a[x + sx * (y + sy * z)] > thr) pred[z * ntx + (x >> 5)] |= 1ull << (y >> 5);
}

template <int C, int R, int B, bool NEXT>
void envelope_over(HostCtx& cx, const uint32_t* labels, const float* fin, float* fout, int n, int64_t cstride, int64_t sx,
                   int64_t nouter, int64_t ostride, float w, int bb, int last, int64_t ntx, const uint64_t* flags,
                   uint64_t* next, float thr_next, long* stats) {
  for (int64_t o = 0; o < nouter; o++)
    for (int64_t t = 0; t < ntx; t++) {
      uint64_t m = flags[o * ntx + t];
      while (m) {                                   // maximal groups of consecutive blocks, like the kernels
        const int b0 = __builtin_ctzll(m);
        int b1 = b0;
        while (b1 + 1 < 64 && ((m >> (b1 + 1)) & 1)) b1++;
        m &= (b1 == 63) ? 0ull : (~0ull << (b1 + 1));
        const int rlo = 32 * b0, rhi = (32 * b1 + 31 < n - 1) ? 32 * b1 + 31 : n - 1;
        if (rlo >= n) break;
        int lo[32], hi[32], rb = fh3::kBig, re = 0;
        for (int l = 0; l < 32; l++) {              // the warp's common row range
          const int64_t x = t * 32 + l;
          lo[l] = fh3::kBig; hi[l] = 0;
          if (x >= sx) continue;
          fh3::extend_to_runs<uint32_t>(cx, labels + o * ostride + x, n, cstride, true, rlo, rhi, lo[l], hi[l]);
          if (lo[l] < rb) rb = lo[l];
          if (hi[l] > re) re = hi[l];
        }
        for (int l = 0; l < 32; l++) {
          const int64_t x = t * 32 + l;
          if (x >= sx) continue;
          const int64_t base = o * ostride + x;
          for (int q = 0; q < 64; q++) { cx.sv[q] = NAN; cx.sh[q] = NAN; cx.sz[q] = NAN; }
          cx.nflag = next ? next + t : nullptr; cx.nstride = ntx; cx.nbit = 1ull << (o >> 5);
          if (g_query_prefetch == 4)
            fh3::column_range<uint32_t, C, R, B, true, NEXT, 4, 1>(cx, labels + base, fin + base, fout + base, n, cstride, w, bb != 0,
                                                                   last != 0, true, rb, re, lo[l], hi[l], thr_next);
          else if (g_query_prefetch == 5)
            fh3::column_range<uint32_t, C, R, B, true, NEXT, 1, 1>(cx, labels + base, fin + base, fout + base, n, cstride, w, bb != 0,
                                                                   last != 0, true, rb, re, lo[l], hi[l], thr_next);
          else
            fh3::column_range<uint32_t, C, R, B, true, NEXT>(cx, labels + base, fin + base, fout + base, n, cstride, w, bb != 0,
                                                             last != 0, true, rb, re, lo[l], hi[l], thr_next);
        }
        if (stats) stats[3] += rhi - rlo + 1;
      }
    }
}

template <int W, int WR, int PF, int C, int R, int B, bool WRITE_BG, bool NEXT>
void roles_pass(const uint32_t* labels, const float* fin, float* fout, int n, int64_t cstride, int64_t sx, int64_t nouter,
                int64_t ostride, float w, int bb, int last, int64_t ntx, const uint64_t* pred, uint64_t* next, float thr_next,
                int order, long* stats) {
  HostCtx cx;
  cx.lv = (float*)malloc(sizeof(float) * (n + 4)); cx.lh = (float*)malloc(sizeof(float) * (n + 4)); cx.lz = (float*)malloc(sizeof(float) * (n + 4));
  uint64_t* resid = (uint64_t*)calloc(ntx * nouter, sizeof(uint64_t));
  for (int step = 0; step < 2; step++) {
    if ((step == 0) == (order == 0)) {
      cx.flag = nullptr;
      envelope_over<C, R, B, NEXT>(cx, labels, fin, fout, n, cstride, sx, nouter, ostride, w, bb, last, ntx, pred, next, thr_next, stats);
    } else {
      for (int64_t o = 0; o < nouter; o++)
        for (int64_t x = 0; x < sx; x++) {
          const int64_t base = o * ostride + x;
          cx.flag = resid + o * ntx + (x >> 5);
          cx.nflag = next ? next + (x >> 5) : nullptr; cx.nstride = ntx; cx.nbit = 1ull << (o >> 5);
          if (g_stencil_v2)
            fh3::stencil_column_v2<uint32_t, W, WR, PF, WRITE_BG, true, NEXT>(cx, labels + base, fin + base, fout + base, n, cstride, w,
                                                                              bb != 0, last != 0, true, pred[o * ntx + (x >> 5)], thr_next);
          else
            fh3::stencil_column<uint32_t, W, WR, PF, WRITE_BG, true, NEXT>(cx, labels + base, fin + base, fout + base, n, cstride, w,
                                                                           bb != 0, last != 0, true, pred[o * ntx + (x >> 5)], thr_next);
        }
    }
  }
  cx.flag = nullptr;
  if (stats) for (int64_t i = 0; i < ntx * nouter; i++) stats[4] += __builtin_popcountll(resid[i]);
  envelope_over<C, R, B, false>(cx, labels, fin, fout, n, cstride, sx, nouter, ostride, w, bb, last, ntx, resid, nullptr, 0.0f, stats);
  if (stats) stats[0] += cx.spills;
  free(resid); free(cx.lv); free(cx.lh); free(cx.lz);
}

template <int WY, int WZ, int WR, int PF>
void run_roles(const uint32_t* labels, int64_t sx, int64_t sy, int64_t sz, float wx, float wy, float wz, int bb, int ndim,
               int order, int garbage, float* out, long* stats) {
  const int64_t V = sx * sy * sz, ntx = (sx + 31) / 32;
  float* ws = (float*)malloc(sizeof(float) * V);
  for (int64_t i = 0; i < V; i++) ws[i] = NAN;           // every voxel must be written by someone
  for (int64_t i = 0; i < V; i++) out[i] = NAN;
  float* a = (ndim == 3) ? out : ws;     // x -> a, y -> b, z -> a
  float* b = (ndim == 3) ? ws : out;
  uint64_t* pred_y = (uint64_t*)calloc(ntx * sz, sizeof(uint64_t));
  uint64_t* pred_z = (uint64_t*)calloc(ntx * sy, sizeof(uint64_t));
  pass_x<uint32_t>(labels, a, sx, sy * sz, wx, bb);
  const float thr_y = wy * wy * (float)((WY + 1) * (WY + 1)), thr_z = wz * wz * (float)((WZ + 1) * (WZ + 1));
  predict_from_x(a, sx, sy, sz, thr_y, pred_y, ntx);
  // a prediction is only a hint: any other prediction must give the same result (garbage: 1 = none, 2 = all,
  // 3 = pseudo-random words)
  if (garbage) {
    uint64_t h = 0x9e3779b97f4a7c15ull;
    for (int64_t i = 0; i < ntx * sz; i++) { h = h * 6364136223846793005ull + 1442695040888963407ull; pred_y[i] = garbage == 1 ? 0ull : (garbage == 2 ? ~0ull : (h & (h >> 7) & (h << 9))); }
  }
  if (stats) for (int64_t i = 0; i < ntx * sz; i++) { stats[5] += __builtin_popcountll(pred_y[i]); stats[6] += __builtin_popcountll(pred_y[i]); }
  roles_pass<WY, (WR < WY ? WR : WY), PF, 4, 16, 4, true, true>(labels, a, b, (int)sy, sx, sx, sz, sx * sy, wy, bb, ndim == 2, ntx, pred_y,
                                                               ndim == 3 ? pred_z : nullptr, thr_z, order, stats);
  if (ndim == 3) {
    if (garbage) {
      uint64_t h = 0x2545f4914f6cdd1dull;
      for (int64_t i = 0; i < ntx * sy; i++) { h = h * 6364136223846793005ull + 1442695040888963407ull; pred_z[i] = garbage == 1 ? 0ull : (garbage == 2 ? ~0ull : (h & (h >> 5) & (h << 11))); }
    }
    if (stats) for (int64_t i = 0; i < ntx * sy; i++) { stats[5] += __builtin_popcountll(pred_z[i]); stats[7] += __builtin_popcountll(pred_z[i]); }
    roles_pass<WZ, (WR < WZ ? WR : WZ), PF, 4, 16, 4, false, false>(labels, b, a, (int)sz, sx * sy, sx, sy, sx, wz, bb, 1, ntx, pred_z,
                                                                   nullptr, 0.0f, order, stats);
  }
  free(pred_y); free(pred_z); free(ws);
}

}  // namespace

// variant as in fh3_host_edt_hybrid; order: 0 = envelope role first, 1 = stencil role first; garbage: see run_roles.
// stats[3] rows given to envelope warps, stats[4] blocks the stencil flagged outside the prediction, stats[5] predicted blocks
extern "C" int fh3_host_edt_roles(const uint32_t* labels, int64_t sx, int64_t sy, int64_t sz, float wx, float wy, float wz,
                                  int bb, int ndim, int variant, int order, int garbage, float* out, long* stats) {
  if (sy > fh3::kMaxN || sz > fh3::kMaxN) return -1;
  switch (variant) {
    case 0: run_roles<10, 4, 4, 11>(labels, sx, sy, sz, wx, wy, wz, bb, ndim, order, garbage, out, stats); break;
    case 1: run_roles<4, 4, 4, 4>(labels, sx, sy, sz, wx, wy, wz, bb, ndim, order, garbage, out, stats); break;
    case 2: run_roles<12, 8, 5, 6>(labels, sx, sy, sz, wx, wy, wz, bb, ndim, order, garbage, out, stats); break;
    case 3: run_roles<2, 1, 1, 1>(labels, sx, sy, sz, wx, wy, wz, bb, ndim, order, garbage, out, stats); break;
    case 4: run_roles<8, 6, 8, 15>(labels, sx, sy, sz, wx, wy, wz, bb, ndim, order, garbage, out, stats); break;
    default: return -2;
  }
  return 0;
}
