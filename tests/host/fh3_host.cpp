// Test infrastructure: runs the K1 v3 column-pass body (kimimaro_b200/csrc/edt_fh3.cuh, the very code the
// CUDA kernel instantiates) on the CPU, one emulated thread per column, so that the CPU suite can check it
// bit for bit against the oracle's EDT.  Compiled by tests/test_edt_fh3_host.py with g++; never shipped.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../kimimaro_b200/csrc/edt_fh3.cuh"

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
