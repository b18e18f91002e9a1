// K1 column pass, v3: one thread per column, Felzenszwalb-Huttenlocher lower envelope with the parabola
// stack in SHARED memory (a ring of C entries per thread; older entries of a deep stack -- blobs only --
// spill to local memory).
//
// Same arithmetic as oracle/oracle.c edt_parabolic_run (which restates the PyPI `edt` library called at
// kimimaro/intake.py:174-185, trace.py:112-117, intake.py:565): run-relative indices, the same
// intersection formula with IEEE division, FLT_MAX in place of +inf, no fused multiply-add -- the
// result is bit-identical to the CPU restatement.  What changed against the v2 kernel (edt.cu) is only
// how the work is laid out on the machine (ncu: v2 was latency-bound on its local-memory stack and spent
// 80 % of its 144 warp instructions per 32-voxel warp-row on integer bookkeeping and divergent control
// flow):
//   * stack entries live in shared memory, [entry][thread] so that any mix of depths across the lanes of a
//     warp is bank-conflict free: three float planes (apex row, height, left end) behind one slot address;
//   * a run's first entry is tagged with a NaN in its left-end slot: `s <= NaN` is false, so the pop loop
//     needs no depth check, and the NaN's payload carries (run end row, number of entries) for the query;
//   * completed runs are written every R rows by all lanes together (coalesced stores); the shared-memory
//     part of the stack is a RING over an ever-growing entry index, so a dense segmentation (a run is always
//     open, the stack never empties) stays inside shared memory as well, and a blob keeps the top C entries
//     -- the ones pushes and pops touch -- there and spills only the older ones to local memory;
//   * labels are loaded two batches ahead and f one batch ahead, f only where the label is non-zero
//     (background needs neither its value nor a store: it stays 0 from the x pass);
//   * the body is a __host__ __device__ template over a context type, so that tests/ can run the very same
//     code on the CPU against the oracle (tests/test_edt_fh3_host.py); the device context is below.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define FH3_HD __host__ __device__ __forceinline__
#else
#define FH3_HD inline
#endif

namespace fh3 {

constexpr float kFltMax = 3.4028234664e38f;
constexpr int kBig = 0x3fffffff;
constexpr int kMaxN = 2047;  // 11-bit row / count fields in the NaN payload

union FI { float f; uint32_t u; };
FH3_HD float u2f(uint32_t u) { FI x; x.u = u; return x.f; }
FH3_HD uint32_t f2u(float f) { FI x; x.f = f; return x.u; }

FH3_HD uint32_t clz32(uint32_t x) {
#ifdef __CUDA_ARCH__
  return (uint32_t)__clz((int)x);
#else
  return x ? (uint32_t)__builtin_clz(x) : 32u;
#endif
}
FH3_HD uint32_t brev32(uint32_t x) {
#ifdef __CUDA_ARCH__
  return __brev(x);
#else
  x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
  x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
  x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
  x = ((x >> 8) & 0x00ff00ffu) | ((x & 0x00ff00ffu) << 8);
  return (x >> 16) | (x << 16);
#endif
}

// Context interface (device: FhDevCtx in edt.cu; host: tests/host/fh3_host.cpp):
//   float mul(a,b), add(a,b), sub(a,b), div(a,b), sqrt(a)   round-to-nearest, never contracted
//   float fmin(a,b)
//   T     ld_label(const T*), float ld_f(const float*), void st_f(float*, float)
//   shared-memory ring, slot in [0, C):  s_st(slot, v, h, z), s_st_z(slot, z), s_ld_v(slot), s_ld_h(slot), s_ld_z(slot)
//   local-memory backing, entry index:   l_st(k, v, h, z), l_st_z(k, z), l_ld_v(k), l_ld_h(k), l_ld_z(k)
//   int   wmin(int), wmax(int), float wmaxf(float >= 0), bool any(bool)   warp-wide (identity on the host)
//   void  note_row(int row)         hybrid pass only: "this row of my tile holds a voxel the stencil could not
//                                   finish" (sets the bit of the row's 32-row block in the tile's flag word)
//   void  note_next(int row)        NEXT only: "this row of my tile holds a value above the NEXT pass's threshold"
//                                   (a hint: sets the bit of this column's outer index in the next pass's prediction
//                                   word of (row, tile)); called by all lanes of the warp together
// An entry is (v, h, z): row of the parabola's apex (absolute, kept as a float: rows < 2^24 are exact and the
// hot loop then needs no int->float conversion), its height, and the left end of its reign (run-relative).

// the entry below the top of the stack (PP of column_range); empty for the original loop
template <int PP> struct PopAhead { float v, h, z; };
template <> struct PopAhead<0> {};

// query lookahead of column_range (QP > 1); empty for the original loop so that it compiles to the original code
template <int QP> struct Lookahead { float v[QP], h[QP], z[QP]; };
template <> struct Lookahead<1> {};

// intersection of the parabola rooted at run-relative row i (height fi) with the one at v (height h):
// (f[i] - f[v] + (i-v) w^2 (i+v)) / (2 (i-v) w^2), oracle.c:137-139
template <typename Ctx>
FH3_HD float intersect(Ctx& cx, float fi, float ir, float h, float v, float w2) {
  const float f1 = cx.mul(cx.sub(ir, v), w2);
  const float f2 = cx.add(ir, v);
  return cx.div(cx.add(cx.sub(fi, h), cx.mul(f1, f2)), cx.mul(2.0f, f1));
}

// The stack is addressed by an ever-growing entry index k.  Entries [ring_lo, k] are in the shared-memory
// ring (slot = k mod C), entries below ring_lo that are still needed are in local memory.  For the thin
// processes of a connectomics volume and for a dense segmentation the live window [kn, k] never exceeds C
// and local memory is never touched; a blob keeps its top C entries -- the ones pushes and pops work on --
// in the ring.
#define FH3_SLOT(k_) ((k_) & (C - 1))
#define FH3_LD_V(k_) (((k_) >= ring_lo) ? cx.s_ld_v(FH3_SLOT(k_)) : cx.l_ld_v(k_))
#define FH3_LD_H(k_) (((k_) >= ring_lo) ? cx.s_ld_h(FH3_SLOT(k_)) : cx.l_ld_h(k_))
#define FH3_LD_Z(k_) (((k_) >= ring_lo) ? cx.s_ld_z(FH3_SLOT(k_)) : cx.l_ld_z(k_))

// C ring entries (power of two), R rows between flushes, B rows per load batch (R % B == 0).
// RANGE = false: the whole column [0, n), in place (fin == fout), background rows are never touched.
// RANGE = true (the blob half of the hybrid pass, see below): rows [rb, re) only, out of place; a lane takes
//   part with the rows [own_lo, own_hi) (complete runs by construction); the rest of the range is treated as
//   background.  Background is never written.
// NEXT: rows whose (squared) result exceeds thr_next are reported with cx.note_next (prediction for the next pass).
// QP: entries the query keeps in registers AHEAD of the one it is evaluating.  The per-SASS-line stall samples of the
// envelope kernel (profiles/r01_edt_hybrid_sass_envelope_y.txt) put 21 % of all samples on one compare: the query's
// `while (next left end < row)` waiting for the left end it has just asked local memory for -- a blob's stack is hundreds
// of entries deep, far beyond the ring, and every row of the write-out walks one entry further, one L2 round trip at a
// time.  With QP > 1 the next QP entries (apex, height, left end) are already in registers and each step only issues the
// load of the entry QP steps ahead.  QP = 1 is the original code.
// PP: the entry BELOW the top of the stack is kept in registers too, so that a pop -- in a blob about one per row -- needs no
// load before the next intersection; the entry below the new top is fetched after the pop, off the critical path.
template <typename T, int C, int R, int B, bool RANGE, bool NEXT = false, int QP = 1, int PP = 0, typename Ctx>
FH3_HD void column_range(Ctx& cx, const T* lp, const float* fin, float* fout, int n, int64_t cstride, float w,
                         bool black_border, bool last_pass, bool active, int rb, int re, int own_lo, int own_hi,
                         float thr_next = 0.0f) {
  static_assert((C & (C - 1)) == 0, "ring size must be a power of two");
  static_assert(R % B == 0, "flush period must be a multiple of the batch");
  const float w2 = cx.mul(w, w);
  const float kNaN = u2f(0x7fc00000u);
  const float kInf = u2f(0x7f800000u);

  // ---- build state: the open run ----
  T run_lab = T(0);
  float af = 0.0f;       // first row of the open run
  int k = -1;            // top entry
  int k_lo = 0;          // first entry of the open run
  int ring_lo = 0;       // entries below this index are in local memory (if still needed at all)
  float tv = 0.0f, th = 0.0f, tz = kNaN;   // top entry, cached (tv run-relative)
  PopAhead<PP> pa;                         // PP: the entry below it, cached (v run-relative); only read when the top can pop
  int last_b = 0;        // end row of the most recent closed run
  // ---- query state: closed runs not written yet occupy entries [kn, kdone) ----
  int kn = 0;

  // software pipeline: labels of batch b+2 and f of batch b+1 are in flight while batch b is consumed
  T l0[B], l1[B], l2[B];
  float f0[B], f1[B];
  const T* lq = lp + rb * cstride;        // next label batch to load
  const float* fq = fin + rb * cstride;   // next f batch to load
  const int64_t bstride = (int64_t)B * cstride;
#define FH3_ROW_OK(r_) (RANGE ? ((r_) >= own_lo && (r_) < own_hi) : (active && (r_) < re))
#pragma unroll
  for (int j = 0; j < B; j++) l1[j] = FH3_ROW_OK(rb + j) ? cx.ld_label(lq + j * cstride) : T(0);
  lq += bstride;
#pragma unroll
  for (int j = 0; j < B; j++) l2[j] = FH3_ROW_OK(rb + B + j) ? cx.ld_label(lq + j * cstride) : T(0);
  lq += bstride;
#pragma unroll
  for (int j = 0; j < B; j++) f1[j] = (l1[j] != T(0)) ? cx.ld_f(fq + j * cstride) : 0.0f;
  fq += bstride;

  for (int i0 = rb; i0 < re; i0 += B) {
#pragma unroll
    for (int j = 0; j < B; j++) { l0[j] = l1[j]; f0[j] = f1[j]; l1[j] = l2[j]; }
#pragma unroll
    for (int j = 0; j < B; j++)   // l1 != 0 implies that the row exists and the thread takes part
      f1[j] = (l1[j] != T(0)) ? cx.ld_f(fq + j * cstride) : 0.0f;
    fq += bstride;
#pragma unroll
    for (int j = 0; j < B; j++) l2[j] = FH3_ROW_OK(i0 + 2 * B + j) ? cx.ld_label(lq + j * cstride) : T(0);
    lq += bstride;
    // ---------------- build: rows [i0, i0 + B); rows >= re were loaded as background ----------------
    const float i0f = (float)i0;
#pragma unroll
    for (int j = 0; j < B; j++) {
      const int i = i0 + j;
      const T lab = l0[j];
      const bool same = lab == run_lab;
      if (!same && run_lab != T(0)) {            // close: tag the run's first entry with (end row, entries)
        const float tag = u2f(0x7fc00000u | ((uint32_t)i << 11) | (uint32_t)(k - k_lo + 1));
        if (k_lo >= ring_lo) cx.s_st_z(FH3_SLOT(k_lo), tag); else cx.l_st_z(k_lo, tag);
        last_b = i;
      }
      run_lab = lab;
      if (lab != T(0)) {
        const float fi = cx.fmin(f0[j], kFltMax);  // the envelope arithmetic stays finite (oracle.c:182-186)
        const float vf = i0f + (float)j;
        float s = kNaN;
        float ir = 0.0f;
        if (same) {
          ir = vf - af;
          s = intersect(cx, fi, ir, th, tv, w2);
          while (s <= tz) {                      // false on the run's first entry (its tz is a NaN)
            k--;
            if constexpr (PP == 0) {
              if (k < ring_lo) {
                tv = cx.l_ld_v(k) - af; th = cx.l_ld_h(k); tz = cx.l_ld_z(k);
                ring_lo = k + 1;
              } else {
                tv = cx.s_ld_v(FH3_SLOT(k)) - af; th = cx.s_ld_h(FH3_SLOT(k)); tz = cx.s_ld_z(FH3_SLOT(k));
              }
            } else {
              tv = pa.v; th = pa.h; tz = pa.z;   // entry k, from registers
              if (k < ring_lo) ring_lo = k + 1;  // the bookkeeping of the load that did not happen
              if (k > k_lo) {                    // entry k-1 of the same run, for the pop after this one
                const int b = k - 1;
                if (b < ring_lo) { pa.v = cx.l_ld_v(b) - af; pa.h = cx.l_ld_h(b); pa.z = cx.l_ld_z(b); }
                else { pa.v = cx.s_ld_v(FH3_SLOT(b)) - af; pa.h = cx.s_ld_h(FH3_SLOT(b)); pa.z = cx.s_ld_z(FH3_SLOT(b)); }
              }
            }
            s = intersect(cx, fi, ir, th, tv, w2);
          }
        } else {
          af = vf; k_lo = k + 1;
        }
        k++;
        const int e = k - C;                     // the entry whose slot is about to be reused
        if (e >= ring_lo) {
          if (e >= kn)                           // still needed: spill it
            cx.l_st(e, cx.s_ld_v(FH3_SLOT(e)), cx.s_ld_h(FH3_SLOT(e)), cx.s_ld_z(FH3_SLOT(e)));
          ring_lo = e + 1;
        }
        cx.s_st(FH3_SLOT(k), vf, fi, s);
        if constexpr (PP != 0) { pa.v = tv; pa.h = th; pa.z = tz; }   // the old top is the entry below the new one
        tv = ir; th = fi; tz = s;
      }
    }
    const int c1 = i0 + B;
    if (((c1 - rb) % R) != 0 && c1 < re) continue;
    // ---------------- flush: write the rows of all closed runs ----------------
    if (c1 >= re && run_lab != T(0)) {           // the range ends inside a run (at the array end, or at a run end)
      const float tag = u2f(0x7fc00000u | ((uint32_t)re << 11) | (uint32_t)(k - k_lo + 1));
      if (k_lo >= ring_lo) cx.s_st_z(FH3_SLOT(k_lo), tag); else cx.l_st_z(k_lo, tag);
      last_b = re;
      run_lab = T(0);
    }
    const int kdone = (run_lab != T(0)) ? k_lo : k + 1;
    const bool todo = kn < kdone;
    const int lo = cx.wmin(todo ? (int)FH3_LD_V(kn) : kBig);
    const int hi = cx.wmax(todo ? last_b : 0);
    if (lo < hi) {
      int qa = todo ? -1 : kBig, qb = qa;        // rows [qa, qb) of the run being written; qb <= i: fetch the next
      int kq = 0, kend = 0;
      float qaf = 0.0f, qbf = 0.0f, cv = 0.0f, ch = 0.0f, nz = kInf;
      Lookahead<QP> la;                          // QP > 1: entries kq+1 .. kq+QP (left end +inf past the run's last one)
      if constexpr (QP > 1) {
#pragma unroll
        for (int j = 0; j < QP; j++) { la.v[j] = 0.0f; la.h[j] = 0.0f; la.z[j] = kInf; }
      }
      bool bl = false, br = false;
      float* fw = fout + lo * cstride;
      float iqf = (float)lo;
      for (int i = lo; i < hi; i++, fw += cstride, iqf += 1.0f) {
        bool hot = false;
        if (i >= qb) {                           // step to the next closed run (or to "nothing left")
          if (kn < kdone) {
            const uint32_t pk = f2u(FH3_LD_Z(kn));
            qaf = FH3_LD_V(kn);
            qa = (int)qaf;
            qb = (int)((pk >> 11) & 0x7ffu);
            qbf = (float)qb;
            kq = kn; kend = kn + (int)(pk & 0x7ffu); kn = kend;
            cv = 0.0f; ch = FH3_LD_H(kq);
            if constexpr (QP == 1) {
              nz = (kq + 1 < kend) ? FH3_LD_Z(kq + 1) : kInf;
            } else {
#pragma unroll
              for (int j = 0; j < QP; j++) {
                const int e = kq + 1 + j;
                const bool in = e < kend;
                la.z[j] = in ? FH3_LD_Z(e) : kInf;
                la.v[j] = in ? FH3_LD_V(e) : 0.0f;
                la.h[j] = in ? FH3_LD_H(e) : 0.0f;
              }
              nz = la.z[0];
            }
            bl = (qa > 0) || black_border;
            br = (qb < n) || black_border;
          } else {
            qa = kBig; qb = kBig;
          }
        }
        if (i >= qa) {                           // qa <= i < qb
          const float ir = iqf - qaf;
          while (nz < ir) {
            kq++;
            if constexpr (QP == 1) {
              cv = FH3_LD_V(kq) - qaf;
              ch = FH3_LD_H(kq);
              nz = (kq + 1 < kend) ? FH3_LD_Z(kq + 1) : kInf;
            } else {
              cv = la.v[0] - qaf;
              ch = la.h[0];
#pragma unroll
              for (int j = 0; j + 1 < QP; j++) { la.v[j] = la.v[j + 1]; la.h[j] = la.h[j + 1]; la.z[j] = la.z[j + 1]; }
              const int e = kq + QP;
              const bool in = e < kend;
              la.z[QP - 1] = in ? FH3_LD_Z(e) : kInf;
              la.v[QP - 1] = in ? FH3_LD_V(e) : 0.0f;
              la.h[QP - 1] = in ? FH3_LD_H(e) : 0.0f;
              nz = la.z[0];
            }
          }
          const float di = cx.sub(ir, cv);
          float val = cx.add(cx.mul(cx.mul(w2, di), di), ch);
          if (bl) { const float e = cx.add(ir, 1.0f); val = cx.fmin(val, cx.mul(cx.mul(w2, e), e)); }
          if (br) { const float e = qbf - iqf; val = cx.fmin(val, cx.mul(cx.mul(w2, e), e)); }
          if (NEXT) hot = val > thr_next;
          if (last_pass) val = (val >= kFltMax) ? kInf : cx.sqrt(val);
          cx.st_f(fw, val);
        }
        if (NEXT) { if (cx.any(hot)) cx.note_next(i); }   // the row loop is warp-uniform
      }
    }
    // every closed run is written: kn == kdone, and the slots below it are free for reuse
  }
#undef FH3_ROW_OK
}

template <typename T, int C, int R, int B, typename Ctx>
FH3_HD void column(Ctx& cx, const T* lp, float* fp, int n, int64_t cstride, float w, bool black_border,
                   bool last_pass, bool active) {
  column_range<T, C, R, B, false>(cx, lp, fp, fp, n, cstride, w, black_border, last_pass, active, 0, n, 0, n);
}

// ---------------------------------------------------------------------------------------------------------
// Hybrid pass (b2t_edt_ws, integer anisotropies): a register stencil everywhere, the envelope only for blobs.
//
// Rows farther than W from row i contribute candidates >= thr = w^2 (W+1)^2 (a parabola's value is at least
// w^2 d^2, and so is a run-end clamp at distance d).  Hence if the minimum v over the 2W+1 rows around i is
// at most thr, v is the exact result.  The stencil pass computes v for every voxel from a sliding register
// window -- no stack, no division, no divergence:
//     v[i] = min_d ( (rows i..i+d all in i's run ? f[i+d] : 0) + w^2 d^2 )
// Masking an unreachable row to 0 turns the tap at the first foreign row into the run-end clamp w^2 d^2, and
// the taps beyond it cannot undercut that clamp.  Run membership is one bit per row ("same label as the row
// before") kept in a 32-bit shift register, like the foreground bits; the clamps inside the window follow from
// those bits alone and bound the tap loop before a single tap is evaluated.  Rows outside the array are virtual:
// with black_border foreign with f = 0 (tap = clamp), without it same-run with f = +inf (tap never wins).
// The tap loop ends as soon as w^2 d^2 reaches the warp-wide maximum of v (uniform branches), so thin
// processes cost a few taps whatever W is.  Voxels with v > thr -- the inside of blobs -- are not final: the
// pass records which 32-row blocks of a 32-column tile hold such voxels (cx.note_row), and a second kernel
// runs column_range<RANGE> over those blocks, extended to complete runs, and overwrites them with the exact
// envelope values.  With integer anisotropy all values a thin voxel can take are integers below 2^24, every
// operation is exact, and the stencil's minimum IS the envelope algorithm's value bit for bit
// (tests/test_edt_fh3_host.py checks the composition against the oracle).
// ---------------------------------------------------------------------------------------------------------

// W tap radius, WR <= W radius of the register window, D rows of load prefetch.  Rows travel global -> shared-memory
// ring (cx.ring_fetch: cp.async on the device, so the prefetch depth costs no registers) -> register window of 2WR+1
// rows with static register indices (the row loop is unrolled 2WR+1 times).  Taps at distance <= WR read registers;
// the rarely needed taps at WR < d <= W (thick processes only) read the ring, which still holds those rows, in a
// rolled loop -- that keeps the unrolled code inside the instruction cache (a 21-phase body at W = 10 ran at a
// quarter of the issue rate).  WRITE_BG: also store the zeros of the background (needed when fout is uninitialised
// scratch; the volume the x pass wrote already has them).
// Ring interface of the context (f slots kRingSlotBytes apart, 32 of them; label slots likewise, 16 of them):
//   ring_fetch(foff, loff, lp, fp) starts the copy of one real row (label and f) and closes a group;
//   ring_put(foff, f) stores a virtual row's f and closes an (empty) group;
//   ring_wait<N>() returns when at most N groups are still in flight; ring_f(foff) / ring_l<T>(loff) read a landed row.
// Foreground is recognised by f > 0: the x pass gives every foreground voxel at least wx^2 and the passes keep it
// positive, background is exactly 0.  A lane outside the volume shadows the last column of its tile (the kernel
// points it there) and merely does not store.
constexpr int kRingSlotBytes = 512;
constexpr int kRingF = 32, kRingL = 16;

// PRED: `pred` (warp-uniform) names the 32-row blocks of this tile that an envelope warp is already redoing, extended
// to complete runs, at the same time (the "roles" kernel in edt.cu): the stencil evaluates no taps there and stores no
// foreground -- every foreground voxel of such a block belongs to a run that meets the block, so the envelope warp
// writes it -- but still the background zeros (WRITE_BG).  Rows of such runs OUTSIDE the predicted blocks are written
// by both warps: with the same bits when the stencil's value is final, and otherwise the block is flagged and the
// residual envelope launch, which runs after both, has the last word.  NEXT: see column_range.
template <typename T, int W, int WR, int D, bool WRITE_BG, bool PRED = false, bool NEXT = false, typename Ctx>
FH3_HD void stencil_column(Ctx& cx, const T* lp, const float* fin, float* fout, int n, int64_t cstride, float w,
                           bool black_border, bool last_pass, bool active, uint64_t pred = 0, float thr_next = 0.0f) {
  constexpr int S = 2 * WR + 1;        // register window: rows i-WR .. i+WR
  constexpr int FMASK = kRingF * kRingSlotBytes - 1, LMASK = kRingL * kRingSlotBytes - 1;
  static_assert(WR >= 1 && WR <= W, "register window inside the tap radius");
  static_assert(D >= 1 && D < kRingL, "rows in flight must fit the label ring");
  static_assert(2 * W + D + 1 <= kRingF, "rows i-W .. i+W+D must fit the f ring");
  const float w2 = cx.mul(w, w);
  const float kInf = u2f(0x7f800000u);
  float cd[WR + 1];
#pragma unroll
  for (int d = 0; d <= WR; d++) cd[d] = cx.mul(cx.mul(w2, (float)d), (float)d);
  const float thr = cx.mul(cx.mul(w2, (float)(W + 1)), (float)(W + 1));
  const float edge_f = black_border ? 0.0f : kInf;
  const uint32_t edge_link = black_border ? 0u : 1u;
  float wf[S];
  uint32_t em = 0xffffffffu;   // bit t: rows (j-t-1, j-t) carry the same label, j = i + W (newest linked row)
  T lprev = T(0);
  const T* lq = lp;            // rows are fetched in order: running pointers instead of 64-bit multiplies
  const float* fq = fin;
  float* fw = fout;
  // three fronts, all advancing one row per step: fetch (row i+W+D), link (row i+W), admit (row i+WR)
  int jf = -W, jl = -W, ja = -W;
  int fo = 0, lo = 0;          // ring offsets of the fetch front (f ring, label ring)
  int ko = 0;                  // label-ring offset of the link front
  int ao = 0;                  // f-ring offset of the admit front
  // generic steps: rows outside the array are virtual
#define FH3_FETCH()                                                                  \
  do {                                                                               \
    if (jf >= 0 && jf < n) { cx.ring_fetch(fo, lo, lq, fq); lq += cstride; fq += cstride; } \
    else cx.ring_put(fo, edge_f);                                                    \
    fo = (fo + kRingSlotBytes) & FMASK; lo = (lo + kRingSlotBytes) & LMASK; jf++;    \
  } while (0)
#define FH3_LINK()                                                                   \
  do {                                                                               \
    uint32_t link = 1u;                                                              \
    if (jl >= 0 && jl < n) {                                                         \
      const T lj = cx.template ring_l<T>(ko);                                        \
      link = (jl == 0) ? edge_link : (uint32_t)(lj == lprev);                        \
      lprev = lj;                                                                    \
    } else if (jl == n) {                                                            \
      link = edge_link;                                                              \
    }                                                                                \
    em = (em << 1) | link;                                                           \
    ko = (ko + kRingSlotBytes) & LMASK; jl++;                                        \
  } while (0)
#define FH3_ADMIT(slot_)                                                             \
  do {                                                                               \
    wf[slot_] = cx.ring_f(ao);                                                       \
    ao = (ao + kRingSlotBytes) & FMASK; ja++;                                        \
  } while (0)
  // steady state: the fetched row i+W+D and the linked row i+W are real rows, and not the first one
#define FH3_FETCH_STEADY()                                                           \
  do {                                                                               \
    cx.ring_fetch(fo, lo, lq, fq); lq += cstride; fq += cstride;                     \
    fo = (fo + kRingSlotBytes) & FMASK; lo = (lo + kRingSlotBytes) & LMASK; jf++;    \
  } while (0)
#define FH3_LINK_STEADY()                                                            \
  do {                                                                               \
    const T lj = cx.template ring_l<T>(ko);                                          \
    em = (em << 1) | (uint32_t)(lj == lprev);                                        \
    lprev = lj;                                                                      \
    ko = (ko + kRingSlotBytes) & LMASK; jl++;                                        \
  } while (0)
  // one row: centre i = window slot (ph + WR) % S; its f-ring offset is WR + 1 slots behind the admit front
#define FH3_ROW(ph_)                                                                 \
  do {                                                                               \
    const int c = ((ph_) + WR) % S;                                                  \
    float v = wf[c];                                                                 \
    const bool fg = v > 0.0f;                                                        \
    float out = 0.0f;                                                                \
    const bool skip = PRED && ((pred >> (i >> 5)) & 1ull) != 0;                      \
    if (!skip && cx.any(fg)) {                                                       \
      /* rows reachable inside the run: consecutive set links above / below the centre */ \
      const int rr = (int)clz32(~(em << (32 - W)));        /* links (i,i+1), (i+1,i+2), ...: bits W-1, W-2, ... */ \
      const int ll = (int)clz32(brev32(~(em >> W)));       /* links (i-1,i), (i-2,i-1), ...: bits W, W+1, ...   */ \
      /* the run-end clamps inside the window follow from the links alone: a tight bound for the tap loop */ \
      if (rr < W) { const float e = (float)(rr + 1); v = cx.fmin(v, cx.mul(cx.mul(w2, e), e)); } \
      if (ll < W) { const float e = (float)(ll + 1); v = cx.fmin(v, cx.mul(cx.mul(w2, e), e)); } \
      const float mx = cx.wmaxf(fg ? v : 0.0f);                                      \
      bool far = W > WR;                                                             \
      _Pragma("unroll")                                                              \
      for (int d = 1; d <= WR; d++) {                                                \
        if (!(cd[d] < mx)) { far = false; break; }         /* uniform: no farther row can improve any lane */ \
        const int cp = ((ph_) + WR + d) % S, cm = ((ph_) + WR - d + S) % S;          \
        const float fp_ = (d <= rr) ? wf[cp] : 0.0f;                                 \
        const float fm_ = (d <= ll) ? wf[cm] : 0.0f;                                 \
        v = cx.fmin(v, cx.add(cx.fmin(fp_, fm_), cd[d]));                            \
      }                                                                              \
      if (W > WR && far) {                                 /* thick processes: the taps beyond the registers */ \
        const int co = ao - (WR + 1) * kRingSlotBytes;                               \
        _Pragma("unroll 1")                                                          \
        for (int d = WR + 1; d <= W; d++) {                                          \
          const float cdd = cx.mul(cx.mul(w2, (float)d), (float)d);                  \
          if (!(cdd < mx)) break;                                                    \
          const float fp_ = (d <= rr) ? cx.ring_f((co + d * kRingSlotBytes) & FMASK) : 0.0f; \
          const float fm_ = (d <= ll) ? cx.ring_f((co - d * kRingSlotBytes) & FMASK) : 0.0f; \
          v = cx.fmin(v, cx.add(cx.fmin(fp_, fm_), cdd));                            \
        }                                                                            \
      }                                                                              \
      if (fg) out = last_pass ? cx.sqrt(v) : v;                                      \
      if (cx.any(fg && v > thr)) cx.note_row(i);           /* not final: the envelope kernel redoes this block */ \
      if (NEXT) { if (cx.any(fg && v > thr_next)) cx.note_next(i); }                 \
    }                                                                                \
    /* !WRITE_BG: fout already holds 0 on background; skip: the foreground is the envelope warp's */ \
    if (active && (skip ? (WRITE_BG && !fg) : (WRITE_BG || fg))) cx.st_f(fw, out);   \
    fw += cstride;                                                                   \
  } while (0)
  // prologue: D rows in flight; the links of rows -W .. W-1; the window rows -WR .. WR-1 (row j -> slot (j + WR) mod S)
  for (int t = 0; t < D; t++) FH3_FETCH();
  for (int t = 0; t < 2 * W; t++) {
    FH3_FETCH();
    cx.template ring_wait<D>();
    FH3_LINK();
  }
  ao = ((W - WR) * kRingSlotBytes) & FMASK; ja = -WR;      // the rows -W .. -WR-1 stay in the ring only
#pragma unroll
  for (int t = 0; t < S - 1; t++) FH3_ADMIT(t);
  for (int base = 0; base < n; base += S) {
    const bool steady = base + S - 1 + W + D < n;                // every row this block touches is an ordinary one
#pragma unroll
    for (int ph = 0; ph < S; ph++) {
      const int i = base + ph;
      if (steady || i < n) {
        // row i + W + D is fetched, row i + W (landed) is linked, row i + WR replaces row i - WR - 1 in the registers
        if (steady) { FH3_FETCH_STEADY(); cx.template ring_wait<D>(); FH3_LINK_STEADY(); }
        else { FH3_FETCH(); cx.template ring_wait<D>(); FH3_LINK(); }
        FH3_ADMIT((ph + S - 1) % S);
        FH3_ROW(ph);
      }
    }
  }
  cx.template ring_wait<0>();
#undef FH3_FETCH
#undef FH3_LINK
#undef FH3_ADMIT
#undef FH3_FETCH_STEADY
#undef FH3_LINK_STEADY
#undef FH3_ROW
}

// ---------------------------------------------------------------------------------------------------------
// stencil_column_v2: the same pass with a leaner steady state.  The per-SASS-line counts of the shipped stencil
// (profiles/r01_edt_hybrid_sass_stencil_z.txt) put ~50 warp instructions of FIXED cost on every row, background
// included: a 7-instruction "is this block steady" test per row, two 64-bit pointer bumps plus register-pair copies
// per stream (three streams), a second shared-memory ring for the labels with its own address, offset and mask
// upkeep, and three row counters that only the column ends need.  Here
//   * the steady loop is its own loop (the column ends run a generic copy of the same phases);
//   * labels travel global -> REGISTERS: a ring lr[S] with static indices, phase ph consumes the label of row i+W that
//     it loaded S rows earlier and reloads the slot, so the label ring in shared memory and its upkeep are gone;
//   * every global address of a row is base pointer + ONE shared 32-bit element offset (one add per row, one
//     multiply-add-wide per stream);
//   * the row counters exist only in the generic loop.
// Same arithmetic, same order of operations on the values: results are bit-identical to stencil_column.  The caller
// guarantees (n + W + D + S) * cstride < 2^32.
// Extra context call: ring_fetch_f(foff, fp) starts the copy of one real row's f and closes a group.
// ---------------------------------------------------------------------------------------------------------
template <typename T, int W, int WR, int D, bool WRITE_BG, bool PRED = false, bool NEXT = false, typename Ctx>
FH3_HD void stencil_column_v2(Ctx& cx, const T* lp, const float* fin, float* fout, int n, int64_t cstride, float w,
                              bool black_border, bool last_pass, bool active, uint64_t pred = 0, float thr_next = 0.0f) {
  constexpr int S = 2 * WR + 1;        // register window: rows i-WR .. i+WR; also the depth of the label ring
  constexpr int FMASK = kRingF * kRingSlotBytes - 1;
  static_assert(WR >= 1 && WR <= W, "register window inside the tap radius");
  static_assert(D >= 1 && 2 * W + D + 1 <= kRingF, "rows i-W .. i+W+D must fit the f ring");
  const float w2 = cx.mul(w, w);
  const float kInf = u2f(0x7f800000u);
  float cd[WR + 1];
#pragma unroll
  for (int d = 0; d <= WR; d++) cd[d] = cx.mul(cx.mul(w2, (float)d), (float)d);
  const float thr = cx.mul(cx.mul(w2, (float)(W + 1)), (float)(W + 1));
  const float edge_f = black_border ? 0.0f : kInf;
  const uint32_t edge_link = black_border ? 0u : 1u;
  const uint32_t cs = (uint32_t)cstride;
  float wf[S];
  T lr[S];                     // lr[ph]: label of row base + ph + W (loaded one block of S rows earlier)
  uint32_t em = 0xffffffffu;   // bit t: rows (j-t-1, j-t) carry the same label, j = i + W (newest linked row)
  T lprev = T(0);
  int fo = 0;                  // f-ring offset of the fetch front (row i + W + D)
  int ao = 0;                  // f-ring offset of the admit front (row i + WR)
  uint32_t eo = 0;             // element offset of row i: i * cstride
  // generic steps (prologue and column end): rows outside the array are virtual
#define FH3_FETCH2(jf_)                                                              \
  do {                                                                               \
    const int jf = (jf_);                                                            \
    if (jf >= 0 && jf < n) cx.ring_fetch_f(fo, fin + (int64_t)jf * cstride);         \
    else cx.ring_put(fo, edge_f);                                                    \
    fo = (fo + kRingSlotBytes) & FMASK;                                              \
  } while (0)
#define FH3_LINK2(jl_, lab_)                                                         \
  do {                                                                               \
    const int jl = (jl_);                                                            \
    uint32_t link = 1u;                                                              \
    if (jl >= 0 && jl < n) {                                                         \
      const T lj = (lab_);                                                           \
      link = (jl == 0) ? edge_link : (uint32_t)(lj == lprev);                        \
      lprev = lj;                                                                    \
    } else if (jl == n) {                                                            \
      link = edge_link;                                                              \
    }                                                                                \
    em = (em << 1) | link;                                                           \
  } while (0)
#define FH3_ADMIT2(slot_)                                                            \
  do {                                                                               \
    wf[slot_] = cx.ring_f(ao);                                                       \
    ao = (ao + kRingSlotBytes) & FMASK;                                              \
  } while (0)
  // one row: centre i = window slot (ph + WR) % S; its f-ring offset is WR + 1 slots behind the admit front
#define FH3_ROW2(ph_)                                                                \
  do {                                                                               \
    const int c = ((ph_) + WR) % S;                                                  \
    float v = wf[c];                                                                 \
    const bool fg = v > 0.0f;                                                        \
    float out = 0.0f;                                                                \
    const bool skip = PRED && ((pred >> (i >> 5)) & 1ull) != 0;                      \
    if (!skip && cx.any(fg)) {                                                       \
      const int rr = (int)clz32(~(em << (32 - W)));                                  \
      const int ll = (int)clz32(brev32(~(em >> W)));                                 \
      if (rr < W) { const float e = (float)(rr + 1); v = cx.fmin(v, cx.mul(cx.mul(w2, e), e)); } \
      if (ll < W) { const float e = (float)(ll + 1); v = cx.fmin(v, cx.mul(cx.mul(w2, e), e)); } \
      const float mx = cx.wmaxf(fg ? v : 0.0f);                                      \
      bool far = W > WR;                                                             \
      _Pragma("unroll")                                                              \
      for (int d = 1; d <= WR; d++) {                                                \
        if (!(cd[d] < mx)) { far = false; break; }                                   \
        const int cp = ((ph_) + WR + d) % S, cm = ((ph_) + WR - d + S) % S;          \
        const float fp_ = (d <= rr) ? wf[cp] : 0.0f;                                 \
        const float fm_ = (d <= ll) ? wf[cm] : 0.0f;                                 \
        v = cx.fmin(v, cx.add(cx.fmin(fp_, fm_), cd[d]));                            \
      }                                                                              \
      if (W > WR && far) {                                                           \
        const int co = ao - (WR + 1) * kRingSlotBytes;                               \
        _Pragma("unroll 1")                                                          \
        for (int d = WR + 1; d <= W; d++) {                                          \
          const float cdd = cx.mul(cx.mul(w2, (float)d), (float)d);                  \
          if (!(cdd < mx)) break;                                                    \
          const float fp_ = (d <= rr) ? cx.ring_f((co + d * kRingSlotBytes) & FMASK) : 0.0f; \
          const float fm_ = (d <= ll) ? cx.ring_f((co - d * kRingSlotBytes) & FMASK) : 0.0f; \
          v = cx.fmin(v, cx.add(cx.fmin(fp_, fm_), cdd));                            \
        }                                                                            \
      }                                                                              \
      if (fg) out = last_pass ? cx.sqrt(v) : v;                                      \
      if (cx.any(fg && v > thr)) cx.note_row(i);                                     \
      if (NEXT) { if (cx.any(fg && v > thr_next)) cx.note_next(i); }                 \
    }                                                                                \
    if (active && (skip ? (WRITE_BG && !fg) : (WRITE_BG || fg))) cx.st_f(fout + eo, out); \
    eo += cs;                                                                        \
  } while (0)
  // prologue: D rows in flight; the links of rows -W .. W-1 (labels read directly); the label ring for rows W .. W+S-1;
  // the window rows -WR .. WR-1 (row j -> slot (j + WR) mod S)
  for (int t = 0; t < D; t++) FH3_FETCH2(-W + t);
  for (int t = 0; t < 2 * W; t++) {
    FH3_FETCH2(-W + D + t);
    cx.template ring_wait<D>();
    const int jp = -W + t;
    FH3_LINK2(jp, cx.ld_label(lp + (int64_t)jp * cstride));
  }
#pragma unroll
  for (int t = 0; t < S; t++) lr[t] = (W + t < n) ? cx.ld_label(lp + (int64_t)(W + t) * cstride) : T(0);
  ao = ((W - WR) * kRingSlotBytes) & FMASK;                // the rows -W .. -WR-1 stay in the ring only
#pragma unroll
  for (int t = 0; t < S - 1; t++) FH3_ADMIT2(t);
  const float* finF = fin + (int64_t)(W + D) * cstride;    // row i + W + D
  const T* lpS = lp + (int64_t)(W + S) * cstride;          // row i + W + S
  int base = 0;
  // steady state: the fetched row i+W+D, the linked row i+W and the label row i+W+S are ordinary rows
  for (; base + S - 1 + W + (D > S ? D : S) < n; base += S) {
#pragma unroll
    for (int ph = 0; ph < S; ph++) {
      const int i = base + ph;
      cx.ring_fetch_f(fo, finF + eo);
      fo = (fo + kRingSlotBytes) & FMASK;
      cx.template ring_wait<D>();
      const T lj = lr[ph];
      em = (em << 1) | (uint32_t)(lj == lprev);
      lprev = lj;
      lr[ph] = cx.ld_label(lpS + eo);
      FH3_ADMIT2((ph + S - 1) % S);
      FH3_ROW2(ph);
    }
  }
  for (; base < n; base += S) {
#pragma unroll
    for (int ph = 0; ph < S; ph++) {
      const int i = base + ph;
      if (i < n) {
        FH3_FETCH2(i + W + D);
        cx.template ring_wait<D>();
        FH3_LINK2(i + W, lr[ph]);
        if (i + W + S < n) lr[ph] = cx.ld_label(lpS + eo);
        FH3_ADMIT2((ph + S - 1) % S);
        FH3_ROW2(ph);
      }
    }
  }
  cx.template ring_wait<0>();
#undef FH3_FETCH2
#undef FH3_LINK2
#undef FH3_ADMIT2
#undef FH3_ROW2
}

// the rows a lane contributes to the envelope half: the complete runs that meet [rlo, rhi]
template <typename T, typename Ctx>
FH3_HD void extend_to_runs(Ctx& cx, const T* lp, int n, int64_t cstride, bool active, int rlo, int rhi, int& own_lo,
                           int& own_hi) {
  own_lo = rlo; own_hi = rhi + 1;
  if (!active) { own_lo = kBig; own_hi = 0; return; }   // takes no part, and does not widen the warp's range
  const T a = cx.ld_label(lp + rlo * cstride);
  if (a != T(0)) while (own_lo > 0 && cx.ld_label(lp + (own_lo - 1) * cstride) == a) own_lo--;
  const T b = cx.ld_label(lp + rhi * cstride);
  if (b != T(0)) while (own_hi < n && cx.ld_label(lp + own_hi * cstride) == b) own_hi++;
}

#undef FH3_SLOT
#undef FH3_LD_V
#undef FH3_LD_H
#undef FH3_LD_Z

}  // namespace fh3
