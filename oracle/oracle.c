/*
 * TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 * CPU restatement ("oracle") of the arithmetic on kimimaro's per-label TEASAR hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load this library; the product (kimimaro_b200) never does.
 *
 * Most of the arithmetic on the path lives in un-vendored PyPI dependencies that are NOT
 * in /root/reference (requirements.txt:2-7): edt>=3.0.0, dijkstra3d>=1.15.0,
 * fill-voids>=2.0.0, connected-components-3d>=3.16.0.  For those, this file restates the
 * published algorithm and anchors on the reference's call sites:
 *   edt.edt                            kimimaro/intake.py:174-185, trace.py:112-117
 *   dijkstra3d.euclidean_distance_field trace.py:139-145, 302-307
 *   dijkstra3d.parental_field          trace.py:155
 *   dijkstra3d.railroad                trace.py:240-242
 *   fill_voids.fill                    trace.py:109
 *   cc3d.connected_components          kimimaro/utility.py:77
 * PARITY FOR THOSE IS "restated, unpinned vs the binary" except through the reference's
 * own known-answer tests (automated_test.py:48-102, 104-199) and brute-force definitions.
 * The in-tree half (ext/skeletontricks/dijkstra_invalidation.hpp:239-332,
 * skeletontricks.pyx:373-418, 995-1045) IS pinned against the compiled reference
 * extension (oracle/_ref, see oracle/build_ref.py).
 *
 * Conventions (kimimaro/intake.py:320-322, skeletontricks.pyx:398): Fortran order,
 * loc = x + sx*(y + sy*z).  26-neighbour enumeration order follows
 * dijkstra_invalidation.hpp:60-124: -x,+x,-y,+y,-z,+z, 4 xy diagonals, 4 yz, 4 xz, 8 corners.
 *
 * Canonical tie rules (the reference leaves these to heap/sort internals):
 *   T1  Dijkstra pops are ordered by (dist, linear index) lexicographically.
 *   T2  arg-max of a distance field = smallest linear index among the maxima.
 *   T3  parent(v) = the 26-neighbour u with the smallest (dist[u], direction index).
 *   T4  railroad stops at the first popped voxel (T1 order) that has a zero-valued
 *       neighbour; the rail voxel is that voxel's first zero neighbour in direction order.
 *   T5  target order = DAF descending, ties by descending linear index.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

static const int DX[26] = {-1, 1, 0, 0, 0, 0, -1, -1, 1, 1, 0, 0, 0, 0, -1, -1, 1, 1, -1, 1, -1, -1, 1, 1, -1, 1};
static const int DY[26] = {0, 0, -1, 1, 0, 0, -1, 1, -1, 1, -1, -1, 1, 1, 0, 0, 0, 0, -1, -1, 1, -1, 1, -1, 1, 1};
static const int DZ[26] = {0, 0, 0, 0, -1, 1, 0, 0, 0, 0, -1, 1, -1, 1, -1, 1, -1, 1, -1, -1, -1, 1, -1, 1, 1, 1};

/* edge lengths in float32, same expressions as dijkstra_invalidation.hpp:45-52 (_s, _c) */
static void edge_weights(float wx, float wy, float wz, float* w) {
  const float sxy = sqrtf(wx * wx + wy * wy);
  const float syz = sqrtf(wy * wy + wz * wz);
  const float sxz = sqrtf(wx * wx + wz * wz);
  const float c = sqrtf(wx * wx + wy * wy + wz * wz);
  w[0] = w[1] = wx; w[2] = w[3] = wy; w[4] = w[5] = wz;
  for (int i = 6; i < 10; i++) w[i] = sxy;
  for (int i = 10; i < 14; i++) w[i] = syz;
  for (int i = 14; i < 18; i++) w[i] = sxz;
  for (int i = 18; i < 26; i++) w[i] = c;
}

/* ------------------------------------------------------------------------------------------
 * Binary min-heap keyed on (float key, int64 value), strict lexicographic order (rule T1).
 * ---------------------------------------------------------------------------------------- */
typedef struct { float key; int64_t val; } hnode;
typedef struct { hnode* a; int64_t n, cap; } heap;

static void heap_init(heap* h) { h->cap = 1024; h->n = 0; h->a = (hnode*)malloc(sizeof(hnode) * h->cap); }
static void heap_free(heap* h) { free(h->a); }
static inline int hless(hnode p, hnode q) { return p.key < q.key || (p.key == q.key && p.val < q.val); }
static void heap_push(heap* h, float key, int64_t val) {
  if (h->n == h->cap) { h->cap *= 2; h->a = (hnode*)realloc(h->a, sizeof(hnode) * h->cap); }
  int64_t i = h->n++;
  hnode x = {key, val};
  while (i > 0) {
    int64_t p = (i - 1) >> 1;
    if (!hless(x, h->a[p])) break;
    h->a[i] = h->a[p];
    i = p;
  }
  h->a[i] = x;
}
static hnode heap_pop(heap* h) {
  hnode top = h->a[0];
  hnode x = h->a[--h->n];
  int64_t i = 0;
  for (;;) {
    int64_t c = 2 * i + 1;
    if (c >= h->n) break;
    if (c + 1 < h->n && hless(h->a[c + 1], h->a[c])) c++;
    if (!hless(h->a[c], x)) break;
    h->a[i] = h->a[c];
    i = c;
  }
  if (h->n > 0) h->a[i] = x;
  return top;
}

/* ------------------------------------------------------------------------------------------
 * K1  multi-label anisotropic Euclidean distance transform (restates PyPI `edt`)
 *     reference call sites: kimimaro/intake.py:174-185, trace.py:112-117, intake.py:565
 *
 * Pass along x: two-sided run scan (distance, in units of wx, to the nearest voxel whose label
 * differs; array edge counts only when black_border); squared.  Passes along y and z: within
 * each maximal run of one non-zero label, lower envelope of the parabolas f[j] + w^2 (i-j)^2
 * (Felzenszwalb & Huttenlocher 2012), clamped by the run's own ends w^2 (i-a+1)^2, w^2 (b-i)^2
 * where an end is a label change (always) or the array edge (only when black_border).
 * ---------------------------------------------------------------------------------------- */
static void edt_pass_first(const uint32_t* seg, float* d, int64_t n, int64_t stride, float w, int black_border) {
  uint32_t working = seg[0];
  if (black_border) d[0] = (working != 0) ? w : 0.0f;
  else d[0] = (working == 0) ? 0.0f : INFINITY;
  for (int64_t i = 1; i < n; i++) {
    const uint32_t s = seg[i * stride];
    if (s == 0) d[i * stride] = 0.0f;
    else if (s == working) d[i * stride] = d[(i - 1) * stride] + w;
    else {
      d[i * stride] = w;
      d[(i - 1) * stride] = (seg[(i - 1) * stride] != 0) ? w : 0.0f;
    }
    working = s;
  }
  int64_t lo = 0;
  if (black_border) { d[(n - 1) * stride] = (seg[(n - 1) * stride] != 0) ? w : 0.0f; lo = 1; }
  for (int64_t i = n - 2; i >= lo; i--) d[i * stride] = fminf(d[i * stride], d[(i + 1) * stride] + w);
  for (int64_t i = 0; i < n; i++) d[i * stride] *= d[i * stride];
}

static void edt_parabolic_run(float* f, int64_t n, int64_t stride, float w, int bl, int br,
                              int* v, float* ff, float* ranges) {
  if (n <= 0) return;
  const float w2 = w * w;
  int k = 0;
  for (int64_t i = 0; i < n; i++) ff[i] = f[i * stride];
  v[0] = 0; ranges[0] = -INFINITY; ranges[1] = INFINITY;
  for (int64_t i = 1; i < n; i++) {
    float s;
    for (;;) {
      const float f1 = (float)(i - v[k]) * w2;
      const float f2 = (float)(i + v[k]);
      s = (ff[i] - ff[v[k]] + f1 * f2) / (2.0f * f1);
      if (k > 0 && s <= ranges[k]) k--; else break;
    }
    k++; v[k] = (int)i; ranges[k] = s; ranges[k + 1] = INFINITY;
  }
  k = 0;
  for (int64_t i = 0; i < n; i++) {
    while (ranges[k + 1] < (float)i) k++;
    const float di = (float)(i - v[k]);
    float val = w2 * di * di + ff[v[k]];
    if (bl) { const float e = (float)(i + 1); val = fminf(val, w2 * e * e); }
    if (br) { const float e = (float)(n - i); val = fminf(val, w2 * e * e); }
    f[i * stride] = val;
  }
}

static void edt_pass_parabolic(const uint32_t* seg, float* f, int64_t n, int64_t stride, float w, int black_border,
                               int* v, float* ff, float* ranges) {
  uint32_t working = seg[0];
  int64_t last = 0;
  for (int64_t i = 1; i < n; i++) {
    const uint32_t s = seg[i * stride];
    if (s != working) {
      if (working != 0)
        edt_parabolic_run(f + last * stride, i - last, stride, w, (black_border || last > 0), 1, v, ff, ranges);
      working = s; last = i;
    }
  }
  if (working != 0 && last < n)
    edt_parabolic_run(f + last * stride, n - last, stride, w, (black_border || last > 0), black_border, v, ff, ranges);
}

/* ndim = 2 runs the x and y passes only (edt.edt on a 2-D array, intake.py:565). */
ORC_API void orc_edt(const uint32_t* labels, int64_t sx, int64_t sy, int64_t sz, float wx, float wy, float wz,
                     int black_border, int ndim, float* out) {
  const int64_t sxy = sx * sy;
  int64_t nmax = sx; if (sy > nmax) nmax = sy; if (sz > nmax) nmax = sz;
  int* v = (int*)malloc(sizeof(int) * (nmax + 2));
  float* ff = (float*)malloc(sizeof(float) * (nmax + 2));
  float* ranges = (float*)malloc(sizeof(float) * (nmax + 2));
  for (int64_t z = 0; z < sz; z++)
    for (int64_t y = 0; y < sy; y++)
      edt_pass_first(labels + sx * y + sxy * z, out + sx * y + sxy * z, sx, 1, wx, black_border);
  /* the library keeps the envelope arithmetic finite: +inf (a row without any label change and no
   * black border) is carried as FLT_MAX through the parabolic passes and restored afterwards, so
   * infinite parabolas never win and never produce NaN intersections. */
  const int64_t V0 = sxy * sz;
  for (int64_t i = 0; i < V0; i++) if (out[i] == INFINITY) out[i] = 3.4028234664e38f;
  for (int64_t z = 0; z < sz; z++)
    for (int64_t x = 0; x < sx; x++)
      edt_pass_parabolic(labels + x + sxy * z, out + x + sxy * z, sy, sx, wy, black_border, v, ff, ranges);
  if (ndim >= 3)
    for (int64_t y = 0; y < sy; y++)
      for (int64_t x = 0; x < sx; x++)
        edt_pass_parabolic(labels + x + sx * y, out + x + sx * y, sz, sxy, wz, black_border, v, ff, ranges);
  const int64_t V = sxy * sz;
  for (int64_t i = 0; i < V; i++) out[i] = (out[i] >= 3.4028234664e38f) ? INFINITY : sqrtf(out[i]);
  free(v); free(ff); free(ranges);
}

/* Closed-form definition, O(N^2): only for tiny volumes, pins orc_edt.
 * Distance from voxel p to a voxel q of different label is measured to the nearer FACE-side
 * position exactly as the separable algorithm does: per axis |p_a - q_a| with q the first
 * differing voxel, which for a separable squared metric equals sqrt(sum w_a^2 (p_a-q_a)^2). */
ORC_API void orc_edt_bruteforce(const uint32_t* labels, int64_t sx, int64_t sy, int64_t sz, float wx, float wy,
                                float wz, int black_border, int ndim, float* out) {
  const int64_t sxy = sx * sy;
  const int64_t z0 = (black_border && ndim >= 3) ? -1 : 0, z1 = (black_border && ndim >= 3) ? sz + 1 : sz;
  const int64_t y0 = black_border ? -1 : 0, y1 = black_border ? sy + 1 : sy;
  const int64_t x0 = black_border ? -1 : 0, x1 = black_border ? sx + 1 : sx;
  for (int64_t pz = 0; pz < sz; pz++) for (int64_t py = 0; py < sy; py++) for (int64_t px = 0; px < sx; px++) {
    const uint32_t lp = labels[px + sx * py + sxy * pz];
    double best = INFINITY;
    if (lp == 0) { out[px + sx * py + sxy * pz] = 0.0f; continue; }
    for (int64_t qz = z0; qz < z1; qz++) for (int64_t qy = y0; qy < y1; qy++) for (int64_t qx = x0; qx < x1; qx++) {
      const int inside = qx >= 0 && qx < sx && qy >= 0 && qy < sy && qz >= 0 && qz < sz;
      const uint32_t lq = inside ? labels[qx + sx * qy + sxy * qz] : 0xFFFFFFFFu - lp; /* virtual: differs */
      if (inside && lq == lp) continue;
      const double ddx = (double)wx * (double)(px - qx), ddy = (double)wy * (double)(py - qy),
                   ddz = (double)wz * (double)(pz - qz);
      const double d2 = ddx * ddx + ddy * ddy + ddz * ddz;
      if (d2 < best) best = d2;
    }
    out[px + sx * py + sxy * pz] = (float)sqrt(best);
  }
}

/* ------------------------------------------------------------------------------------------
 * K2  dijkstra3d.euclidean_distance_field; call sites trace.py:139-145, 302-307.
 *     field != 0 is foreground; edge lengths per direction; background/unreachable = +inf.
 *     free_space_radius > 0 (soma only, trace.py:134): voxels of the axis-aligned box inscribed
 *     in the sphere of that radius (half-side r/sqrt(3) in physical units) take the closed-form
 *     anisotropic Euclidean distance and are frozen; the box's outer shell seeds the frontier.
 *     (The exact seeding rule of dijkstra3d is not recoverable here: parity UNPINNED for r>0.)
 *     max_loc: rule T2.  n_src sources allowed (utility.py:613).
 * ---------------------------------------------------------------------------------------- */
ORC_API int64_t orc_edf(const uint8_t* field, int64_t sx, int64_t sy, int64_t sz, float wx, float wy, float wz,
                        const int64_t* sources, int64_t n_src, float free_space_radius, float* dist) {
  const int64_t sxy = sx * sy, V = sxy * sz;
  float w[26]; edge_weights(wx, wy, wz, w);
  uint8_t* done = (uint8_t*)calloc(V, 1);
  for (int64_t i = 0; i < V; i++) dist[i] = INFINITY;
  heap h; heap_init(&h);
  if (free_space_radius > 0 && n_src == 1) {
    const int64_t s = sources[0];
    const int64_t cz = s / sxy, cy = (s - cz * sxy) / sx, cx = s - sx * (cy + sy * cz);
    const float half = free_space_radius / sqrtf(3.0f);
    const int64_t rx = (int64_t)(half / wx), ry = (int64_t)(half / wy), rz = (int64_t)(half / wz);
    for (int64_t z = cz - rz; z <= cz + rz; z++) for (int64_t y = cy - ry; y <= cy + ry; y++)
      for (int64_t x = cx - rx; x <= cx + rx; x++) {
        if (x < 0 || y < 0 || z < 0 || x >= sx || y >= sy || z >= sz) continue;
        const int64_t loc = x + sx * (y + sy * z);
        if (!field[loc]) continue;
        const float ax = wx * (float)(x - cx), ay = wy * (float)(y - cy), az = wz * (float)(z - cz);
        dist[loc] = sqrtf(ax * ax + ay * ay + az * az);
        const int shell = (x == cx - rx || x == cx + rx || y == cy - ry || y == cy + ry || z == cz - rz || z == cz + rz);
        if (shell) heap_push(&h, dist[loc], loc); else done[loc] = 2; /* frozen interior */
      }
    if (done[s] != 2 && !(dist[s] == 0.0f)) { dist[s] = 0.0f; heap_push(&h, 0.0f, s); }
  } else {
    for (int64_t i = 0; i < n_src; i++) { dist[sources[i]] = 0.0f; heap_push(&h, 0.0f, sources[i]); }
  }
  int64_t max_loc = (n_src > 0) ? sources[0] : -1;
  float max_d = -1.0f;
  while (h.n > 0) {
    const hnode t = heap_pop(&h);
    const int64_t loc = t.val;
    if (done[loc] == 1) continue;
    if (t.key > dist[loc]) continue; /* stale */
    done[loc] = 1;
    const int64_t z = loc / sxy, y = (loc - z * sxy) / sx, x = loc - sx * (y + sy * z);
    for (int i = 0; i < 26; i++) {
      const int64_t nx = x + DX[i], ny = y + DY[i], nz = z + DZ[i];
      if (nx < 0 || ny < 0 || nz < 0 || nx >= sx || ny >= sy || nz >= sz) continue;
      const int64_t n = nx + sx * (ny + sy * nz);
      if (!field[n] || done[n]) continue;
      const float nd = dist[loc] + w[i];
      if (nd < dist[n]) { dist[n] = nd; heap_push(&h, nd, n); }
    }
  }
  /* rule T2: smallest linear index among the maxima of the finite distances */
  for (int64_t i = 0; i < V; i++) if (dist[i] < INFINITY && dist[i] > max_d) { max_d = dist[i]; max_loc = i; }
  heap_free(&h); free(done);
  return max_loc;
}

/* ------------------------------------------------------------------------------------------
 * K4  node-weighted Dijkstra family: cost of ENTERING voxel v is field[v].
 *     Distances are computed with a real heap Dijkstra (rule T1); parents by rule T3 from the
 *     final distances, which equals "first strict improvement" whenever no exact ties occur.
 * ---------------------------------------------------------------------------------------- */
static int64_t best_parent(const float* dist, int64_t loc, int64_t sx, int64_t sy, int64_t sz) {
  const int64_t sxy = sx * sy;
  const int64_t z = loc / sxy, y = (loc - z * sxy) / sx, x = loc - sx * (y + sy * z);
  int64_t best = -1; float bd = INFINITY;
  for (int i = 0; i < 26; i++) {
    const int64_t nx = x + DX[i], ny = y + DY[i], nz = z + DZ[i];
    if (nx < 0 || ny < 0 || nz < 0 || nx >= sx || ny >= sy || nz >= sz) continue;
    const int64_t n = nx + sx * (ny + sy * nz);
    if (dist[n] < bd) { bd = dist[n]; best = n; }
  }
  return best;
}

/* dijkstra3d.railroad(field, source) -- trace.py:240-242.  Returns path length; path[0] is the
 * rail voxel, path[len-1] is `source` (kimimaro's target).  scratch dist must hold V floats. */
ORC_API int64_t orc_railroad(const float* field, int64_t sx, int64_t sy, int64_t sz, int64_t source,
                             float* dist, int64_t* path) {
  const int64_t sxy = sx * sy, V = sxy * sz;
  if (field[source] == 0.0f) { path[0] = source; return 1; }
  for (int64_t i = 0; i < V; i++) dist[i] = INFINITY;
  uint8_t* done = (uint8_t*)calloc(V, 1);
  heap h; heap_init(&h);
  dist[source] = 0.0f; heap_push(&h, 0.0f, source);
  int64_t rail = -1, last = -1;
  while (h.n > 0 && rail < 0) {
    const hnode t = heap_pop(&h);
    const int64_t loc = t.val;
    if (done[loc] || t.key > dist[loc]) continue;
    done[loc] = 1;
    const int64_t z = loc / sxy, y = (loc - z * sxy) / sx, x = loc - sx * (y + sy * z);
    for (int i = 0; i < 26; i++) {
      const int64_t nx = x + DX[i], ny = y + DY[i], nz = z + DZ[i];
      if (nx < 0 || ny < 0 || nz < 0 || nx >= sx || ny >= sy || nz >= sz) continue;
      const int64_t n = nx + sx * (ny + sy * nz);
      if (field[n] == 0.0f) { rail = n; last = loc; break; }   /* rule T4 */
      if (done[n]) continue;
      const float nd = dist[loc] + field[n];
      if (nd < dist[n]) { dist[n] = nd; heap_push(&h, nd, n); }
    }
  }
  int64_t len = 0;
  if (rail < 0) { path[0] = source; len = 1; }
  else {
    path[len++] = rail;
    int64_t loc = last;
    while (loc != source && len < V) {
      path[len++] = loc;
      /* only settled voxels carry final distances; unsettled ones are >= dist[loc] */
      loc = best_parent(dist, loc, sx, sy, sz);
    }
    path[len++] = source;
  }
  heap_free(&h); free(done);
  return len;
}

/* dijkstra3d.parental_field(field, source) -- trace.py:155 (fix_branching=False only).
 * parents[v] = parent linear index + 1, 0 = none. */
ORC_API void orc_parental_field(const float* field, int64_t sx, int64_t sy, int64_t sz, int64_t source,
                                float* dist, uint32_t* parents) {
  const int64_t sxy = sx * sy, V = sxy * sz;
  for (int64_t i = 0; i < V; i++) { dist[i] = INFINITY; parents[i] = 0; }
  uint8_t* done = (uint8_t*)calloc(V, 1);
  heap h; heap_init(&h);
  dist[source] = 0.0f; heap_push(&h, 0.0f, source);
  while (h.n > 0) {
    const hnode t = heap_pop(&h);
    const int64_t loc = t.val;
    if (done[loc] || t.key > dist[loc]) continue;
    done[loc] = 1;
    const int64_t z = loc / sxy, y = (loc - z * sxy) / sx, x = loc - sx * (y + sy * z);
    for (int i = 0; i < 26; i++) {
      const int64_t nx = x + DX[i], ny = y + DY[i], nz = z + DZ[i];
      if (nx < 0 || ny < 0 || nz < 0 || nx >= sx || ny >= sy || nz >= sz) continue;
      const int64_t n = nx + sx * (ny + sy * nz);
      if (done[n]) continue;
      const float nd = dist[loc] + field[n];
      if (nd < dist[n]) { dist[n] = nd; heap_push(&h, nd, n); }
    }
  }
  for (int64_t i = 0; i < V; i++) {
    if (i == source || !(dist[i] < INFINITY)) continue;
    parents[i] = (uint32_t)(best_parent(dist, i, sx, sy, sz) + 1);
  }
  heap_free(&h); free(done);
}

/* ------------------------------------------------------------------------------------------
 * K5  in-component rolling-ball invalidation.
 *     Reference: skeletontricks.pyx:373-418 -> dijkstra_invalidation.hpp:239-332 (in tree; the
 *     compiled reference in oracle/_ref is the real thing).  Three restatements:
 *
 *     orc_invalidate_seq    the reference's ordered best-first claim process with a canonical
 *                           heap order (dist, seed order in path, voxel index) instead of
 *                           libstdc++'s unspecified equal-key order ("strict" semantics).
 *     orc_invalidate_rounds round-synchronous parallel claim (the engine's semantics): round 0
 *                           claims every still-valid seed for itself; in round k every unclaimed
 *                           valid voxel adjacent to a voxel claimed in round k-1 collects the
 *                           candidates (||w.(v - seed)||, seed order) of those neighbours' owners
 *                           with dist < r_seed (strict, hpp:325) and is claimed by the minimum.
 *     Both return the number of voxels zeroed and edit `mask` in place.
 *     Distance expression in float32 exactly as hpp:45-52,319-323.
 * ---------------------------------------------------------------------------------------- */
static inline float seed_dist(float wx, float wy, float wz, int64_t x, int64_t y, int64_t z,
                              int64_t ox, int64_t oy, int64_t oz) {
  const float a = wx * (float)(x - ox), b = wy * (float)(y - oy), c = wz * (float)(z - oz);
  return sqrtf(a * a + b * b + c * c);
}

typedef struct { float key; int32_t seed; int64_t val; } inode;
static inline int iless(inode p, inode q) {
  if (p.key != q.key) return p.key < q.key;
  if (p.seed != q.seed) return p.seed < q.seed;
  return p.val < q.val;
}

ORC_API int64_t orc_invalidate_seq(uint8_t* mask, int64_t sx, int64_t sy, int64_t sz, float wx, float wy, float wz,
                                   const int64_t* seeds, const float* radii, int64_t n_seeds) {
  const int64_t sxy = sx * sy;
  int64_t cap = 4096, n = 0;
  inode* a = (inode*)malloc(sizeof(inode) * cap);
#define IPUSH(K, S, VV) do { \
    if (n == cap) { cap *= 2; a = (inode*)realloc(a, sizeof(inode) * cap); } \
    inode xx = {(K), (S), (VV)}; int64_t ii = n++; \
    while (ii > 0) { int64_t pp = (ii - 1) >> 1; if (!iless(xx, a[pp])) break; a[ii] = a[pp]; ii = pp; } \
    a[ii] = xx; } while (0)
  for (int64_t i = 0; i < n_seeds; i++) IPUSH(0.0f, (int32_t)i, seeds[i]);
  int64_t invalidated = 0;
  while (n > 0) {
    inode top = a[0];
    inode xx = a[--n];
    int64_t ii = 0;
    for (;;) {
      int64_t c = 2 * ii + 1;
      if (c >= n) break;
      if (c + 1 < n && iless(a[c + 1], a[c])) c++;
      if (!iless(a[c], xx)) break;
      a[ii] = a[c]; ii = c;
    }
    if (n > 0) a[ii] = xx;
    const int64_t loc = top.val;
    if (!mask[loc]) continue;
    mask[loc] = 0; invalidated++;
    const int64_t o = seeds[top.seed];
    const int64_t oz = o / sxy, oy = (o - oz * sxy) / sx, ox = o - sx * (oy + sy * oz);
    const int64_t z = loc / sxy, y = (loc - z * sxy) / sx, x = loc - sx * (y + sy * z);
    const float r = radii[top.seed];
    for (int i = 0; i < 26; i++) {
      const int64_t nx = x + DX[i], ny = y + DY[i], nz = z + DZ[i];
      if (nx < 0 || ny < 0 || nz < 0 || nx >= sx || ny >= sy || nz >= sz) continue;
      const int64_t nb = nx + sx * (ny + sy * nz);
      if (!mask[nb]) continue;
      const float d = seed_dist(wx, wy, wz, nx, ny, nz, ox, oy, oz);
      if (d < r) IPUSH(d, top.seed, nb);
    }
  }
#undef IPUSH
  free(a);
  return invalidated;
}

/* orc_invalidate_window: a parallel claim process ordered by KEY instead of by hop count (design study for the
 * engine's K5): every valid voxel holds its best candidate (||w.(v - seed)||, seed order); a round
 * claims ALL candidates whose key is below (smallest open key + delta) at once, then the new owners push candidates to
 * their valid neighbours (min-reduction, so the order inside a round does not matter).  delta = 0 claims exact ties
 * only and is the ordered process of orc_invalidate_seq up to pushes that undercut the current key; a delta of about
 * one voxel needs about as many rounds as the hop-synchronous orc_invalidate_rounds. */
ORC_API int64_t orc_invalidate_window(uint8_t* mask, int64_t sx, int64_t sy, int64_t sz, float wx, float wy, float wz,
                                      const int64_t* seeds, const float* radii, int64_t n_seeds, float delta,
                                      int64_t* n_rounds, int tie_high) {
  const int64_t sxy = sx * sy, V = sxy * sz;
  float* ck = (float*)malloc(sizeof(float) * V);
  int32_t* cs = (int32_t*)malloc(sizeof(int32_t) * V);
  int64_t* act = (int64_t*)malloc(sizeof(int64_t) * (V + n_seeds));
  int64_t* claim = (int64_t*)malloc(sizeof(int64_t) * (V + n_seeds));
  for (int64_t i = 0; i < V; i++) cs[i] = -1;
  int64_t nact = 0, invalidated = 0, rounds = 0;
  for (int64_t i = 0; i < n_seeds; i++) {
    const int64_t s = seeds[i];
    if (!mask[s]) continue;
    if (cs[s] < 0) { cs[s] = (int32_t)i; ck[s] = 0.0f; act[nact++] = s; }
  }
  while (nact > 0) {
    float kmin = ck[act[0]];
    for (int64_t q = 1; q < nact; q++) if (ck[act[q]] < kmin) kmin = ck[act[q]];
    const float lim = kmin + delta;
    int64_t nc = 0, keep = 0;
    for (int64_t q = 0; q < nact; q++) {
      const int64_t v = act[q];
      if (ck[v] == kmin || ck[v] < lim) claim[nc++] = v; else act[keep++] = v;
    }
    nact = keep;
    for (int64_t q = 0; q < nc; q++) { mask[claim[q]] = 0; invalidated++; }
    for (int64_t q = 0; q < nc; q++) {
      const int64_t loc = claim[q];
      const int32_t sd = cs[loc];
      const int64_t o = seeds[sd];
      const int64_t oz = o / sxy, oy = (o - oz * sxy) / sx, ox = o - sx * (oy + sy * oz);
      const int64_t z = loc / sxy, y = (loc - z * sxy) / sx, x = loc - sx * (y + sy * z);
      const float r = radii[sd];
      for (int i = 0; i < 26; i++) {
        const int64_t nx = x + DX[i], ny = y + DY[i], nz = z + DZ[i];
        if (nx < 0 || ny < 0 || nz < 0 || nx >= sx || ny >= sy || nz >= sz) continue;
        const int64_t nb = nx + sx * (ny + sy * nz);
        if (!mask[nb]) continue;
        const float d = seed_dist(wx, wy, wz, nx, ny, nz, ox, oy, oz);
        if (!(d < r)) continue;
        if (cs[nb] < 0) { cs[nb] = sd; ck[nb] = d; act[nact++] = nb; }
        else if (d < ck[nb] || (d == ck[nb] && (tie_high ? sd > cs[nb] : sd < cs[nb]))) { cs[nb] = sd; ck[nb] = d; }
      }
    }
    rounds++;
  }
  if (n_rounds) *n_rounds = rounds;
  free(ck); free(cs); free(act); free(claim);
  return invalidated;
}

/* orc_invalidate_heap: the reference's loop LITERALLY, including what the C++ standard leaves to the library: the
 * order of equal keys in std::priority_queue.  libstdc++'s push_heap / pop_heap (bits/stl_heap.h: __push_heap,
 * __adjust_heap) are restated below and driven with the reference's comparator `t1.dist >= t2.dist`
 * (dijkstra_invalidation.hpp:233-237), the neighbours are visited in the reference's order with its aliasing at the
 * x faces (a corner entry is gated on y and z only, hpp:116-123, so at x = 0 / sx-1 it repeats the yz diagonal and that
 * voxel is pushed twice).  With this the restatement reproduces the compiled reference voxel for voxel
 * (tests/test_oracle_cpu.py::test_invalidation_vs_reference_ext), i.e. every difference of the other two forms is a
 * tie-order effect and nothing else. */
typedef struct { float dist; int64_t orig; int64_t value; float maxd; } rnode;
static inline int rcomp(const rnode* a, const rnode* b) { return a->dist >= b->dist; }
static void r_push_heap(rnode* first, int64_t hole, int64_t top, rnode value) {
  int64_t parent = (hole - 1) / 2;
  while (hole > top && rcomp(&first[parent], &value)) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}
static void r_adjust_heap(rnode* first, int64_t hole, int64_t len, rnode value) {
  const int64_t top = hole;
  int64_t second = hole;
  while (second < (len - 1) / 2) {
    second = 2 * (second + 1);
    if (rcomp(&first[second], &first[second - 1])) second--;
    first[hole] = first[second];
    hole = second;
  }
  if ((len & 1) == 0 && second == (len - 2) / 2) {
    second = 2 * (second + 1);
    first[hole] = first[second - 1];
    hole = second - 1;
  }
  r_push_heap(first, hole, top, value);
}

/* statistics of the last orc_invalidate_heap call: peak heap size, pushes (sizing study for the engine's strict mode) */
ORC_API int64_t orc_heap_stats[2] = {0, 0};
ORC_API int64_t orc_invalidate_heap(uint8_t* mask, int64_t sx, int64_t sy, int64_t sz, float wx, float wy, float wz,
                                    const int64_t* seeds, const float* radii, int64_t n_seeds) {
  const int64_t sxy = sx * sy;
  int64_t cap = 4096, n = 0, peak = 0, pushes = 0;
  rnode* a = (rnode*)malloc(sizeof(rnode) * cap);
#define HPUSH(D, O, VV, M) do { \
    if (n == cap) { cap *= 2; a = (rnode*)realloc(a, sizeof(rnode) * cap); } \
    rnode xx = {(D), (O), (VV), (M)}; a[n++] = xx; r_push_heap(a, n - 1, 0, xx); \
    pushes++; if (n > peak) peak = n; } while (0)
  for (int64_t i = 0; i < n_seeds; i++) HPUSH(0.0f, seeds[i], seeds[i], radii[i]);
  int64_t invalidated = 0;
  while (n > 0) {
    const rnode top = a[0];
    if (n > 1) {                       /* std::pop_heap */
      const rnode value = a[n - 1];
      a[n - 1] = a[0];
      r_adjust_heap(a, 0, n - 1, value);
    }
    n--;                               /* pop_back */
    const int64_t loc = top.value;
    if (!mask[loc]) continue;
    mask[loc] = 0; invalidated++;
    const int64_t o = top.orig;
    const int64_t oz = o / sxy, oy = (o - oz * sxy) / sx, ox = o - sx * (oy + sy * oz);
    const int64_t z = loc / sxy, y = (loc - z * sxy) / sx, x = loc - sx * (y + sy * z);
    int64_t nh[26];
    nh[0] = -1 * (x > 0); nh[1] = (x < sx - 1); nh[2] = -sx * (y > 0); nh[3] = sx * (y < sy - 1);
    nh[4] = -sxy * (z > 0); nh[5] = sxy * (z < sz - 1);
    nh[6] = (nh[0] + nh[2]) * (nh[0] && nh[2]); nh[7] = (nh[0] + nh[3]) * (nh[0] && nh[3]);
    nh[8] = (nh[1] + nh[2]) * (nh[1] && nh[2]); nh[9] = (nh[1] + nh[3]) * (nh[1] && nh[3]);
    nh[10] = (nh[2] + nh[4]) * (nh[2] && nh[4]); nh[11] = (nh[2] + nh[5]) * (nh[2] && nh[5]);
    nh[12] = (nh[3] + nh[4]) * (nh[3] && nh[4]); nh[13] = (nh[3] + nh[5]) * (nh[3] && nh[5]);
    nh[14] = (nh[0] + nh[4]) * (nh[0] && nh[4]); nh[15] = (nh[0] + nh[5]) * (nh[0] && nh[5]);
    nh[16] = (nh[1] + nh[4]) * (nh[1] && nh[4]); nh[17] = (nh[1] + nh[5]) * (nh[1] && nh[5]);
    nh[18] = (nh[0] + nh[2] + nh[4]) * (nh[2] && nh[4]); nh[19] = (nh[1] + nh[2] + nh[4]) * (nh[2] && nh[4]);
    nh[20] = (nh[0] + nh[3] + nh[4]) * (nh[3] && nh[4]); nh[21] = (nh[0] + nh[2] + nh[5]) * (nh[2] && nh[5]);
    nh[22] = (nh[1] + nh[3] + nh[4]) * (nh[3] && nh[4]); nh[23] = (nh[1] + nh[2] + nh[5]) * (nh[2] && nh[5]);
    nh[24] = (nh[0] + nh[3] + nh[5]) * (nh[3] && nh[5]); nh[25] = (nh[1] + nh[3] + nh[5]) * (nh[3] && nh[5]);
    for (int i = 0; i < 26; i++) {
      if (nh[i] == 0) continue;
      const int64_t nb = loc + nh[i];
      if (!mask[nb]) continue;
      const int64_t nz = nb / sxy, ny = (nb - nz * sxy) / sx, nx = nb - sx * (ny + sy * nz);
      const float d = seed_dist(wx, wy, wz, nx, ny, nz, ox, oy, oz);
      if (d < top.maxd) HPUSH(d, o, nb, top.maxd);
    }
  }
#undef HPUSH
  free(a);
  orc_heap_stats[0] = peak; orc_heap_stats[1] = pushes;
  return invalidated;
}

ORC_API int64_t orc_invalidate_rounds(uint8_t* mask, int64_t sx, int64_t sy, int64_t sz, float wx, float wy,
                                      float wz, const int64_t* seeds, const float* radii, int64_t n_seeds) {
  const int64_t sxy = sx * sy, V = sxy * sz;
  int64_t* cur = (int64_t*)malloc(sizeof(int64_t) * (V + n_seeds));
  int64_t* nxt = (int64_t*)malloc(sizeof(int64_t) * (V + n_seeds));
  int32_t* owner = (int32_t*)malloc(sizeof(int32_t) * V);
  float* cd = (float*)malloc(sizeof(float) * V);   /* candidate dist of this round */
  int32_t* cs = (int32_t*)malloc(sizeof(int32_t) * V); /* candidate seed of this round, -1 = none */
  for (int64_t i = 0; i < V; i++) cs[i] = -1;
  int64_t ncur = 0, invalidated = 0;
  for (int64_t i = 0; i < n_seeds; i++) {
    const int64_t s = seeds[i];
    if (!mask[s]) continue;            /* already invalid seeds never expand (hpp:297-299) */
    mask[s] = 0; owner[s] = (int32_t)i; cur[ncur++] = s; invalidated++;
  }
  while (ncur > 0) {
    int64_t nn = 0;
    for (int64_t q = 0; q < ncur; q++) {
      const int64_t loc = cur[q];
      const int32_t sd = owner[loc];
      const int64_t o = seeds[sd];
      const int64_t oz = o / sxy, oy = (o - oz * sxy) / sx, ox = o - sx * (oy + sy * oz);
      const int64_t z = loc / sxy, y = (loc - z * sxy) / sx, x = loc - sx * (y + sy * z);
      const float r = radii[sd];
      for (int i = 0; i < 26; i++) {
        const int64_t nx = x + DX[i], ny = y + DY[i], nz = z + DZ[i];
        if (nx < 0 || ny < 0 || nz < 0 || nx >= sx || ny >= sy || nz >= sz) continue;
        const int64_t nb = nx + sx * (ny + sy * nz);
        if (!mask[nb]) continue;
        const float d = seed_dist(wx, wy, wz, nx, ny, nz, ox, oy, oz);
        if (!(d < r)) continue;
        if (cs[nb] < 0) { cs[nb] = sd; cd[nb] = d; nxt[nn++] = nb; }
        else if (d < cd[nb] || (d == cd[nb] && sd < cs[nb])) { cs[nb] = sd; cd[nb] = d; }
      }
    }
    for (int64_t q = 0; q < nn; q++) {
      const int64_t nb = nxt[q];
      mask[nb] = 0; owner[nb] = cs[nb]; cs[nb] = -1; invalidated++;
    }
    int64_t* t = cur; cur = nxt; nxt = t; ncur = nn;
  }
  free(cur); free(nxt); free(owner); free(cd); free(cs);
  return invalidated;
}

/* ------------------------------------------------------------------------------------------
 * K6  fill_voids.fill -- trace.py:109.  6-connected flood of the background from
 *     the six faces; everything not reached becomes foreground.  Returns number filled.
 * ---------------------------------------------------------------------------------------- */
ORC_API int64_t orc_fill_voids(uint8_t* mask, int64_t sx, int64_t sy, int64_t sz) {
  const int64_t sxy = sx * sy, V = sxy * sz;
  uint8_t* reach = (uint8_t*)calloc(V, 1);
  int64_t* stack = (int64_t*)malloc(sizeof(int64_t) * (V + 1));
  int64_t sp = 0;
#define SEED(L) do { int64_t l_ = (L); if (!mask[l_] && !reach[l_]) { reach[l_] = 1; stack[sp++] = l_; } } while (0)
  for (int64_t z = 0; z < sz; z++) for (int64_t y = 0; y < sy; y++) { SEED(0 + sx * (y + sy * z)); SEED(sx - 1 + sx * (y + sy * z)); }
  for (int64_t z = 0; z < sz; z++) for (int64_t x = 0; x < sx; x++) { SEED(x + sx * (0 + sy * z)); SEED(x + sx * (sy - 1 + sy * z)); }
  for (int64_t y = 0; y < sy; y++) for (int64_t x = 0; x < sx; x++) { SEED(x + sx * y); SEED(x + sx * (y + sy * (sz - 1))); }
  while (sp > 0) {
    const int64_t loc = stack[--sp];
    const int64_t z = loc / sxy, y = (loc - z * sxy) / sx, x = loc - sx * (y + sy * z);
    if (x > 0) SEED(loc - 1);
    if (x < sx - 1) SEED(loc + 1);
    if (y > 0) SEED(loc - sx);
    if (y < sy - 1) SEED(loc + sx);
    if (z > 0) SEED(loc - sxy);
    if (z < sz - 1) SEED(loc + sxy);
  }
#undef SEED
  int64_t filled = 0;
  for (int64_t i = 0; i < V; i++) if (!mask[i] && !reach[i]) { mask[i] = 1; filled++; }
  free(reach); free(stack);
  return filled;
}

/* ------------------------------------------------------------------------------------------
 * N1  cc3d.connected_components -- utility.py:77.  26-connected (8 in 2-D, which
 *     is the same code with sz=1), multi-label: different non-zero values never merge.
 *     Output ids 1..M in order of first appearance in a Fortran-order raster scan.
 * ---------------------------------------------------------------------------------------- */
static int64_t uf_find(int64_t* p, int64_t i) { while (p[i] != i) { p[i] = p[p[i]]; i = p[i]; } return i; }

ORC_API int64_t orc_ccl26(const uint32_t* labels, int64_t sx, int64_t sy, int64_t sz, uint32_t* out) {
  const int64_t sxy = sx * sy, V = sxy * sz;
  int64_t* parent = (int64_t*)malloc(sizeof(int64_t) * V);
  for (int64_t i = 0; i < V; i++) parent[i] = i;
  /* the 13 already-visited neighbours of a raster scan */
  for (int64_t z = 0; z < sz; z++) for (int64_t y = 0; y < sy; y++) for (int64_t x = 0; x < sx; x++) {
    const int64_t loc = x + sx * (y + sy * z);
    const uint32_t l = labels[loc];
    if (!l) continue;
    for (int dz = -1; dz <= 0; dz++) for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) {
      if (dz == 0 && (dy > 0 || (dy == 0 && dx >= 0))) continue;
      const int64_t nx = x + dx, ny = y + dy, nz = z + dz;
      if (nx < 0 || ny < 0 || nz < 0 || nx >= sx || ny >= sy) continue;
      const int64_t n = nx + sx * (ny + sy * nz);
      if (labels[n] != l) continue;
      int64_t a = uf_find(parent, loc), b = uf_find(parent, n);
      if (a != b) { if (a < b) parent[b] = a; else parent[a] = b; }
    }
  }
  int64_t next = 0;
  for (int64_t i = 0; i < V; i++) {
    if (!labels[i]) { out[i] = 0; continue; }
    const int64_t r = uf_find(parent, i);
    if (r == i) out[i] = (uint32_t)(++next);      /* root = min index of its component = first appearance */
    else out[i] = out[r];
  }
  free(parent);
  return next;
}
