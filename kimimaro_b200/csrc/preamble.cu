// Preamble / epilogue kernels around the trace, sm_100a.
//
//   ccl26        26-connected multi-label connected components (replaces cc3d.connected_components,
//                kimimaro/utility.py:77): lock-free union-find over the dense volume, roots are the
//                smallest linear index of each component, i.e. first appearance in a Fortran raster
//                scan -- the same numbering order cc3d produces.
//   relabel      cc[v] = rank[root[v]] after the host side has ranked the roots.
//   gather_paths compact the per-label path segments of the trace pool into one contiguous buffer
//                and fetch the radii (DBF at the vertex, trace.py:186-187) in the same pass.
#include "common.cuh"

namespace {

struct Dims {
  int sx, sy, sz;
  uint32_t sxy;
};

constexpr uint32_t kNone = 0xffffffffu;

__device__ __forceinline__ uint32_t uf_find(uint32_t* parent, uint32_t i) {
  uint32_t p = __ldcg(&parent[i]);
  while (p != i) {
    const uint32_t gp = __ldcg(&parent[p]);
    if (gp != p) parent[i] = gp;  // path halving (benign race: only ever points further up)
    i = p;
    p = gp;
  }
  return i;
}

// read-only traversal: used by the flatten pass, where a concurrent path-halving store could otherwise
// overwrite an already flattened parent[i] = root with a non-root ancestor
__device__ __forceinline__ uint32_t uf_find_ro(const uint32_t* parent, uint32_t i) {
  uint32_t p = __ldcg(&parent[i]);
  while (p != i) {
    i = p;
    p = __ldcg(&parent[i]);
  }
  return i;
}

__device__ __forceinline__ void uf_union(uint32_t* parent, uint32_t a, uint32_t b) {
  for (;;) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a > b) { const uint32_t t = a; a = b; b = t; }
    const uint32_t old = atomicMin(&parent[b], a);   // hook the larger root under the smaller
    if (old == b) return;
    b = old;
  }
}

// ---- run-based union-find -------------------------------------------------------------------------------------------
// A voxel's first parent is the start of its x-run (consecutive voxels of one label in a row), cut at every 32nd linear
// index so that a warp finds it with one ballot.  The merge then needs ONE union per pair of touching runs instead of
// one per pair of touching voxels (a tube of radius 5 has ~10 voxels per run; the all-pairs form spent 8.5 of the CCL's
// 11.5 ms chasing parents on synthetic-512): for a run [s, e] and a run [s', e'] of the same label in one of the four
// backward rows
//   direct overlap        the voxel at max(s, s') does the union: it is a head itself or sees a head straight across,
//   diagonal only, left   (e' = s - 1) the head at s,
//   diagonal only, right  (s' = e + 1) the tail at e,
// and the pieces of a run that a 32-voxel cut separated are joined by the cut's first voxel.  Roots stay the smallest
// linear index of a component (the larger root is always hooked under the smaller), so the numbering is unchanged.
template <typename T>
__global__ void ccl_init_kernel(const T* __restrict__ labels, uint32_t* __restrict__ parent, Dims d, uint64_t V) {
  const uint64_t span = ((V + 31) / 32) * 32;        // whole warps stay together for the ballot
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < span; i += (uint64_t)gridDim.x * blockDim.x) {
    const bool in = i < V;
    const T l = in ? labels[i] : T(0);
    const uint32_t x = (uint32_t)(i % (uint64_t)d.sx);
    const bool head = l != T(0) && ((i & 31u) == 0 || x == 0 || labels[i - 1] != l);
#ifdef B2T_HOST_EMU
    uint64_t s = i;                                    // sequential emulation: walk back to the head (at most 31 steps)
    if (l != T(0)) while ((s & 31u) != 0 && (uint32_t)(s % (uint64_t)d.sx) != 0 && labels[s - 1] == l) s--;
    (void)head;
    if (in) parent[i] = l != T(0) ? (uint32_t)s : kNone;
#else
    const uint32_t heads = __ballot_sync(0xffffffffu, head);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t upto = heads & (0xffffffffu >> (31u - lane));          // heads at or before this lane
    if (in) parent[i] = l != T(0) ? (uint32_t)(i - lane + (31u - (uint32_t)__clz((int)upto))) : kNone;
#endif
  }
}

// The unions themselves are chains of dependent L2 round trips (two finds, one atomicMin), and only the few lanes that
// sit on a head or a tail have any: done in place, a warp would wait for its busiest lane while the other lanes idle
// (call 27: 10.5 of the CCL's 13 ms, against 2.6 ms for everything else).  So the voxels of a block only QUEUE their
// pairs in shared memory, and whenever the queue holds enough of them every thread of the block takes one: the union
// work is spread evenly over all lanes whatever the shape of the labels.
constexpr int kMergeThreads = 256;
constexpr int kMergeDrain = 1024;                       // drain the queue once it holds this many pairs
constexpr int kMergeCap = kMergeDrain + 5 * kMergeThreads;

template <typename T>
__global__ void __launch_bounds__(kMergeThreads) ccl_merge_kernel(const T* __restrict__ labels, uint32_t* __restrict__ parent,
                                                                  Dims d, uint64_t V) {
  __shared__ uint32_t qa[kMergeCap], qb[kMergeCap];
  __shared__ uint32_t qn;
  if (threadIdx.x == 0) qn = 0;
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * kMergeThreads;
  for (uint64_t base = (uint64_t)blockIdx.x * kMergeThreads; base < V; base += stride) {
    const uint64_t i = base + threadIdx.x;
    const T l = i < V ? labels[i] : T(0);
    if (l != T(0)) {
      const uint32_t loc = (uint32_t)i;
      const int z = loc / d.sxy;
      const uint32_t r = loc - (uint32_t)z * d.sxy;
      const int y = r / (uint32_t)d.sx;
      const int x = r - (uint32_t)y * d.sx;
      const bool xm = x > 0, xp = x < d.sx - 1;
      const bool left = xm && labels[loc - 1] == l, right = xp && labels[loc + 1] == l;
      auto push = [&](uint32_t other) { const uint32_t k = atomicAdd(&qn, 1u); qa[k] = loc; qb[k] = other; };
      if (left && (loc & 31u) == 0) push(loc - 1);                           // a run cut at a warp boundary
      const bool head = !left, tail = !right;
      const int64_t sx = d.sx, sxy = d.sxy;
      // the four rows that precede this one in raster order: (y-1, z), (y-1, z-1), (y, z-1), (y+1, z-1)
      const bool row_ok[4] = {y > 0, y > 0 && z > 0, z > 0, y < d.sy - 1 && z > 0};
      const int64_t row_off[4] = {-sx, -sx - sxy, -sxy, sx - sxy};
#pragma unroll
      for (int k = 0; k < 4; k++) {
        if (!row_ok[k]) continue;
        const uint32_t c = (uint32_t)((int64_t)loc + row_off[k]);
        const bool a = xm && labels[c - 1] == l, b = labels[c] == l, e = xp && labels[c + 1] == l;
        if (b) {
          if (head || !a) push(c);
        } else {
          if (a && head) push(c - 1);
          if (e && tail) push(c + 1);
        }
      }
    }
    __syncthreads();
    const uint32_t n = qn;
    if (n >= (uint32_t)kMergeDrain || base + stride >= V) {                  // uniform over the block
      for (uint32_t k = threadIdx.x; k < n; k += kMergeThreads) uf_union(parent, qa[k], qb[k]);
      __syncthreads();
      if (threadIdx.x == 0) qn = 0;
      __syncthreads();
    }
  }
}

__global__ void ccl_flatten_kernel(uint32_t* __restrict__ parent, uint8_t* __restrict__ is_root, uint64_t V) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < V; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t p = parent[i];
    if (p == kNone) { is_root[i] = 0; continue; }
    const uint32_t r = uf_find_ro(parent, (uint32_t)i);
    parent[i] = r;   // every store of this kernel writes a root, so concurrent readers stay correct
    is_root[i] = r == (uint32_t)i;
  }
}

__global__ void ccl_relabel_kernel(uint32_t* __restrict__ parent_to_cc, const int32_t* __restrict__ rank, uint64_t V) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < V; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t p = parent_to_cc[i];
    parent_to_cc[i] = (p == kNone) ? 0u : (uint32_t)rank[p];
  }
}

__global__ void gather_paths_kernel(const uint32_t* __restrict__ pool, const uint32_t* __restrict__ src_off,
                                    const uint32_t* __restrict__ len, const uint64_t* __restrict__ dst_off,
                                    const float* __restrict__ dbf, uint32_t* __restrict__ dst_vox,
                                    float* __restrict__ dst_radius) {
  const uint32_t j = blockIdx.x;
  const uint32_t n = len[j];
  const uint32_t* s = pool + src_off[j];
  const uint64_t o = dst_off[j];
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    const uint32_t v = s[i];
    dst_vox[o + i] = v;
    dst_radius[o + i] = (v == kNone) ? 0.0f : dbf[v];
  }
}

unsigned grid_for(uint64_t V) {
  const uint64_t want = (V + 255) / 256;
  return (unsigned)(want < 148ull * 64 ? want : 148ull * 64);
}

template <typename T>
int ccl_launch(const T* labels, Dims d, uint64_t V, uint32_t* parent, uint8_t* is_root, cudaStream_t st) {
  B2T_LAUNCH(ccl_init_kernel<T>, grid_for(V), 256, st)(labels, parent, d, V);
  B2T_LAUNCH_SYNC(ccl_merge_kernel<T>, grid_for(V), kMergeThreads, st)(labels, parent, d, V);
  B2T_LAUNCH(ccl_flatten_kernel, grid_for(V), 256, st)(parent, is_root, V);
  B2T_CUDA_TRY(cudaGetLastError());
  b2t_count_launches(3);
  return B2T_OK;
}

}  // namespace

// Step 1 of the CCL: d_parent[v] = smallest linear index of v's component (0xffffffff on background),
// d_is_root[v] = 1 on exactly one voxel per component.  The caller ranks the roots (an exclusive
// prefix sum over d_is_root, in raster order) and calls b2t_ccl_relabel.
B2T_EXPORT int b2t_ccl26_roots(const void* d_labels, int label_bytes, int64_t sx, int64_t sy, int64_t sz,
                               uint32_t* d_parent, uint8_t* d_is_root, void* stream) {
  B2T_REQUIRE(d_labels && d_parent && d_is_root, "b2t_ccl26_roots: null pointer");
  B2T_REQUIRE(sx > 0 && sy > 0 && sz > 0 && (double)sx * sy * sz < 4294967295.0, "bad volume shape");
  cudaStream_t st = (cudaStream_t)stream;
  Dims d{(int)sx, (int)sy, (int)sz, (uint32_t)(sx * sy)};
  const uint64_t V = (uint64_t)sx * sy * sz;
  switch (label_bytes) {
    case 1: return ccl_launch<uint8_t>((const uint8_t*)d_labels, d, V, d_parent, d_is_root, st);
    case 2: return ccl_launch<uint16_t>((const uint16_t*)d_labels, d, V, d_parent, d_is_root, st);
    case 4: return ccl_launch<uint32_t>((const uint32_t*)d_labels, d, V, d_parent, d_is_root, st);
    case 8: return ccl_launch<unsigned long long>((const unsigned long long*)d_labels, d, V, d_parent, d_is_root, st);
    default: b2t_set_error("b2t_ccl26_roots: label_bytes must be 1, 2, 4 or 8"); return B2T_ERR_ARG;
  }
}

// Step 2: d_parent (in place) becomes the cc label volume: cc[v] = d_rank[parent[v]], 0 on background.
B2T_EXPORT int b2t_ccl_relabel(uint32_t* d_parent, const int32_t* d_rank, uint64_t n_voxels, void* stream) {
  B2T_REQUIRE(d_parent && d_rank, "b2t_ccl_relabel: null pointer");
  B2T_LAUNCH(ccl_relabel_kernel, grid_for(n_voxels), 256, (cudaStream_t)stream)(d_parent, d_rank, n_voxels);
  B2T_CUDA_TRY(cudaGetLastError());
  b2t_count_launches(1);
  return B2T_OK;
}

// =================================================================================================
// K6  fill_voids.fill (kimimaro/trace.py:109, soma labels only): every background voxel
// that is not 6-connected to a face of the array becomes foreground.  Same lock-free union-find as the
// CCL, restricted to background voxels and the 3 backward face neighbours; components that own a voxel
// on a face of the array are "outside", everything else is a void and gets filled.  (A frontier flood
// needs one grid barrier per voxel of depth -- 23 ms on the 450x450x180 soma crop; this takes passes
// over the crop instead.)
// d_mask: uint8 [V] edited in place; d_reach: uint32 [V] scratch (outside flags); d_queue: >= V u32
// scratch (parents); the number of filled voxels is left in d_ctrl[5].
// =================================================================================================
namespace {

// run-based like the CCL above: a background voxel starts under the head of its x-run (cut every 32 voxels), and two
// runs that face each other in y or z are joined once, by the voxel at the larger of the two starts
__global__ void fill_init_kernel(const uint8_t* __restrict__ mask, uint32_t* __restrict__ parent,
                                 uint32_t* __restrict__ outside, Dims d, uint64_t V) {
  const uint64_t span = ((V + 31) / 32) * 32;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < span; i += (uint64_t)gridDim.x * blockDim.x) {
    const bool in = i < V;
    const bool bg = in && !mask[i];
    const uint32_t x = (uint32_t)(i % (uint64_t)d.sx);
    const bool head = bg && ((i & 31u) == 0 || x == 0 || mask[i - 1]);
#ifdef B2T_HOST_EMU
    uint64_t s = i;
    if (bg) while ((s & 31u) != 0 && (uint32_t)(s % (uint64_t)d.sx) != 0 && !mask[s - 1]) s--;
    (void)head;
    if (in) { parent[i] = bg ? (uint32_t)s : kNone; outside[i] = 0; }
#else
    const uint32_t heads = __ballot_sync(0xffffffffu, head);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t upto = heads & (0xffffffffu >> (31u - lane));
    if (in) { parent[i] = bg ? (uint32_t)(i - lane + (31u - (uint32_t)__clz((int)upto))) : kNone; outside[i] = 0; }
#endif
  }
}

__global__ void fill_merge_kernel(const uint8_t* __restrict__ mask, uint32_t* __restrict__ parent, Dims d, uint64_t V) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < V; i += (uint64_t)gridDim.x * blockDim.x) {
    if (mask[i]) continue;
    const uint32_t loc = (uint32_t)i;
    const int z = loc / d.sxy;
    const uint32_t r = loc - (uint32_t)z * d.sxy;
    const int y = r / (uint32_t)d.sx;
    const int x = r - (uint32_t)y * d.sx;
    const bool left = x > 0 && !mask[loc - 1];
    if (left && (loc & 31u) == 0) uf_union(parent, loc, loc - 1);          // a run cut at a warp boundary
    const bool head = !left;
    if (y > 0) {
      const uint32_t c = loc - (uint32_t)d.sx;
      if (!mask[c] && (head || x == 0 || mask[c - 1])) uf_union(parent, loc, c);
    }
    if (z > 0) {
      const uint32_t c = loc - d.sxy;
      if (!mask[c] && (head || x == 0 || mask[c - 1])) uf_union(parent, loc, c);
    }
  }
}

__global__ void fill_flatten_kernel(uint32_t* __restrict__ parent, uint32_t* __restrict__ outside, Dims d, uint64_t V) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < V; i += (uint64_t)gridDim.x * blockDim.x) {
    if (parent[i] == kNone) continue;
    const uint32_t root = uf_find_ro(parent, (uint32_t)i);
    parent[i] = root;
    const uint32_t loc = (uint32_t)i;
    const int z = loc / d.sxy;
    const uint32_t r = loc - (uint32_t)z * d.sxy;
    const int y = r / (uint32_t)d.sx;
    const int x = r - (uint32_t)y * d.sx;
    if (x == 0 || y == 0 || z == 0 || x == d.sx - 1 || y == d.sy - 1 || z == d.sz - 1) outside[root] = 1;
  }
}

__global__ void fill_apply_kernel(uint8_t* __restrict__ mask, const uint32_t* __restrict__ parent,
                                  const uint32_t* __restrict__ outside, uint32_t* __restrict__ ctrl, uint64_t V) {
  uint32_t filled = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < V; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t p = parent[i];
    if (p != kNone && !outside[p]) { mask[i] = 1; filled++; }
  }
  if (filled) atomicAdd(&ctrl[5], filled);
}

}  // namespace

B2T_EXPORT int b2t_fill_voids(uint8_t* d_mask, int64_t sx, int64_t sy, int64_t sz, uint32_t* d_reach,
                              uint32_t* d_queue, uint64_t queue_cap, uint32_t* d_ctrl, void* stream) {
  B2T_REQUIRE(d_mask && d_reach && d_queue && d_ctrl, "b2t_fill_voids: null pointer");
  B2T_REQUIRE(sx > 0 && sy > 0 && sz > 0 && (double)sx * sy * sz < 4294967295.0, "bad volume shape");
  const uint64_t V = (uint64_t)sx * sy * sz;
  B2T_REQUIRE(queue_cap >= V, "b2t_fill_voids: d_queue must hold at least one u32 per voxel");
  cudaStream_t st = (cudaStream_t)stream;
  Dims d{(int)sx, (int)sy, (int)sz, (uint32_t)(sx * sy)};
  B2T_CUDA_TRY(cudaMemsetAsync(d_ctrl, 0, 8 * sizeof(uint32_t), st));
  B2T_LAUNCH(fill_init_kernel, grid_for(V), 256, st)(d_mask, d_queue, d_reach, d, V);
  B2T_LAUNCH(fill_merge_kernel, grid_for(V), 256, st)(d_mask, d_queue, d, V);
  B2T_LAUNCH(fill_flatten_kernel, grid_for(V), 256, st)(d_queue, d_reach, d, V);
  B2T_LAUNCH(fill_apply_kernel, grid_for(V), 256, st)(d_mask, d_queue, d_reach, d_ctrl, V);
  B2T_CUDA_TRY(cudaGetLastError());
  b2t_count_launches(4);
  return B2T_OK;
}

// Sequential float32 sums over segments: out[s] = ((0 + v[a]) + v[a+1]) + ... for a = off[s] .. off[s+1]-1.
// One thread per segment, strictly left to right, because the reference accumulates label centroids in
// float32 in scan order (compute_centroids, ext/skeletontricks/skeletontricks.pyx:528-588: `xsum[label] += x`)
// and a tree reduction would round differently once a sum passes 2^24.
__global__ void segment_seqsum_kernel(const float* __restrict__ xs, const float* __restrict__ ys,
                                      const int64_t* __restrict__ off, uint32_t n_seg, float* __restrict__ outx,
                                      float* __restrict__ outy) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seg) return;
  float ax = 0.0f, ay = 0.0f;
  for (int64_t i = off[s]; i < off[s + 1]; i++) {
    ax = __fadd_rn(ax, xs[i]);
    ay = __fadd_rn(ay, ys[i]);
  }
  outx[s] = ax;
  outy[s] = ay;
}

B2T_EXPORT int b2t_segment_seqsum(const float* d_xs, const float* d_ys, const int64_t* d_off, uint32_t n_seg,
                                  float* d_outx, float* d_outy, void* stream) {
  if (n_seg == 0) return B2T_OK;
  B2T_REQUIRE(d_xs && d_ys && d_off && d_outx && d_outy, "b2t_segment_seqsum: null pointer");
  B2T_LAUNCH(segment_seqsum_kernel, (n_seg + 127) / 128, 128, (cudaStream_t)stream)(d_xs, d_ys, d_off, n_seg, d_outx, d_outy);
  B2T_CUDA_TRY(cudaGetLastError());
  b2t_count_launches(1);
  return B2T_OK;
}

// ------------------------------------------------------------------------------------------------
// fix_borders, per face (kimimaro/intake.py:544-585): the per-component reductions of find_border_targets
// (pyx:591-648: DT maximum, every voxel that attains it, first raster position), compute_centroids (pyx:528-588:
// coordinate sums, voxel count) and get_mapping (pyx:490-525: the volume label under a face component) in three small
// launches, compacted on the device so that the host reads a few thousand numbers per face in one go.
//   d_tab    6 * (P + 1) u32 scratch: max DT bits | first position | count | sum x | sum y | volume label
//   d_cand   2 * P u32: (position, face component) of every voxel that attains its component's DT maximum, any order
//   d_rec    7 * P u32: per face component present: id, max DT bits, first position, count, sum x, sum y, volume label
//   d_count  2 u32: number of candidates, number of records (zeroed here)
// Coordinate sums are integers: as float32 they equal the reference's sequential float32 sums whenever they stay below
// 2^24 (every partial sum is then exact); the host redoes the rare larger component in the reference's order.
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void face_reduce_kernel(const uint32_t* __restrict__ ccp, const float* __restrict__ dt,
                                   const uint32_t* __restrict__ plane, uint32_t p0, uint32_t P, uint32_t* __restrict__ tab) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const uint32_t l = ccp[i];
  if (l == 0 || l > P) return;
  const size_t T = (size_t)P + 1;
  const float d = dt[i];
  if (d != 0.0f) {
    atomicMax(&tab[l], __float_as_uint(d));
    atomicMin(&tab[T + l], i);
  }
  atomicAdd(&tab[2 * T + l], 1u);
  atomicAdd(&tab[3 * T + l], i % p0);
  atomicAdd(&tab[4 * T + l], i / p0);
  tab[5 * T + l] = plane[i];          // every voxel of a face component lies in the same volume component
}

__global__ void face_candidates_kernel(const uint32_t* __restrict__ ccp, const float* __restrict__ dt, uint32_t P,
                                       const uint32_t* __restrict__ tab, uint32_t* __restrict__ cand,
                                       uint32_t* __restrict__ count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const uint32_t l = ccp[i];
  if (l == 0 || l > P) return;
  const float d = dt[i];
  if (d != 0.0f && __float_as_uint(d) == tab[l]) {
    const uint32_t k = atomicAdd(&count[0], 1u);
    cand[2 * (size_t)k] = i;
    cand[2 * (size_t)k + 1] = l;
  }
}

__global__ void face_records_kernel(uint32_t P, const uint32_t* __restrict__ tab, uint32_t* __restrict__ rec,
                                    uint32_t* __restrict__ count) {
  const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (l > P) return;
  const size_t T = (size_t)P + 1;
  if (tab[2 * T + l] == 0 || tab[l] == 0) return;     // absent, or no voxel with a non-zero DT
  const uint32_t k = atomicAdd(&count[1], 1u);
  uint32_t* r = rec + 7 * (size_t)k;
  r[0] = l; r[1] = tab[l]; r[2] = tab[T + l]; r[3] = tab[2 * T + l]; r[4] = tab[3 * T + l]; r[5] = tab[4 * T + l];
  r[6] = tab[5 * T + l];
}
}  // namespace

B2T_EXPORT int b2t_face_stats(const uint32_t* d_cc_plane, const float* d_dt, const uint32_t* d_plane, int64_t p0, int64_t p1,
                              uint32_t* d_tab, uint32_t* d_cand, uint32_t* d_rec, uint32_t* d_count, void* stream) {
  B2T_REQUIRE(p0 > 0 && p1 > 0 && (double)p0 * (double)p1 < 2147483647.0, "b2t_face_stats: bad plane shape");
  B2T_REQUIRE(d_cc_plane && d_dt && d_plane && d_tab && d_cand && d_rec && d_count, "b2t_face_stats: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const uint32_t P = (uint32_t)(p0 * p1);
  const size_t T = (size_t)P + 1;
  B2T_CUDA_TRY(cudaMemsetAsync(d_tab, 0, 6 * T * sizeof(uint32_t), st));
  B2T_CUDA_TRY(cudaMemsetAsync(d_tab + T, 0xff, T * sizeof(uint32_t), st));     // first position: minimum
  B2T_CUDA_TRY(cudaMemsetAsync(d_count, 0, 2 * sizeof(uint32_t), st));
  const unsigned blocks = (P + 255) / 256;
  B2T_LAUNCH(face_reduce_kernel, blocks, 256, st)(d_cc_plane, d_dt, d_plane, (uint32_t)p0, P, d_tab);
  B2T_LAUNCH(face_candidates_kernel, blocks, 256, st)(d_cc_plane, d_dt, P, d_tab, d_cand, d_count);
  B2T_LAUNCH(face_records_kernel, blocks, 256, st)(P, d_tab, d_rec, d_count);
  B2T_CUDA_TRY(cudaGetLastError());
  b2t_count_launches(3);
  return B2T_OK;
}

// Compact n_seg path segments (pool[src_off[j] .. +len[j])) to dst[dst_off[j] ..) and fetch DBF at each vertex.
B2T_EXPORT int b2t_gather_paths(const uint32_t* d_pool, const uint32_t* d_src_off, const uint32_t* d_len,
                                const uint64_t* d_dst_off, uint32_t n_seg, const float* d_dbf, uint32_t* d_dst_vox,
                                float* d_dst_radius, void* stream) {
  if (n_seg == 0) return B2T_OK;
  B2T_LAUNCH(gather_paths_kernel, n_seg, 128, (cudaStream_t)stream)(d_pool, d_src_off, d_len, d_dst_off, d_dbf, d_dst_vox,
                                                              d_dst_radius);
  B2T_CUDA_TRY(cudaGetLastError());
  b2t_count_launches(1);
  return B2T_OK;
}

// =================================================================================================
// Skeleton assembly (kimimaro/trace.py:182-192: Skeleton.from_path per path, simple_merge, consolidate; intake.py:509-517,
// 587-593: the components of one original label merged and consolidated again) for one GROUP of path segments per CTA:
//   vertices   the distinct path voxels of the group in lexicographic (x, y, z) order (np.unique(vertices, axis=0)),
//              radius = DBF at the first occurrence in the group's path buffer (consolidate keeps the first)
//   edges      consecutive path voxels, as (smaller, larger) vertex ranks, distinct, sorted, no self loops
//   then vertices without an edge are dropped and the ranks renumbered (the order of both lists is kept).
// A dense u32 scratch volume (0xffffffff everywhere on entry, restored on exit) is the hash set: atomicMin of the entry
// position finds the first occurrence of a voxel, later it holds the voxel's vertex rank for the edge pass.  Both sorts
// are shared-memory bitonic sorts over the next power of two.  Groups with more than kAsmCap entries (the caller sends
// those down its general path) report n_vertices = 0xffffffff and are left alone.
//   d_vox / d_rad      the path buffer: voxel indices (0xffffffff = end of a path) and DBF per entry
//   d_seg_start/_len   per segment: first entry and number of entries; segments of a group are consecutive
//   d_grp_seg          [n_grp + 1] first segment of every group;  d_grp_out [n_grp + 1] first output slot of every group
//                      (prefix sum of the groups' entry counts: a group never has more vertices or edges than entries)
//   d_out_verts [3 * N] f32 physical coordinates, d_out_rad [N], d_out_edges [2 * N] u32, d_out_count [2 * n_grp]
// =================================================================================================
namespace {

constexpr uint32_t kAsmCap = 8192;
constexpr int kAsmThreads = 256;

struct AsmShared {
  unsigned long long vk[kAsmCap];     // (lexicographic key << 32) | position of the first occurrence
  uint32_t ek[kAsmCap];               // edge keys (lo rank << 13) | hi rank ... ranks < kAsmCap = 2^13
  uint32_t n_v, n_e, n_used;
};

template <typename K>
__device__ void asm_bitonic(K* a, uint32_t n_pow2) {          // ascending; all kAsmThreads threads call
  for (uint32_t k = 2; k <= n_pow2; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t i = threadIdx.x; i < n_pow2; i += kAsmThreads) {
        const uint32_t l = i ^ j;
        if (l > i) {
          const K x = a[i], y = a[l];
          const bool up = (i & k) == 0;
          if ((x > y) == up) { a[i] = y; a[l] = x; }
        }
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ uint32_t asm_pow2(uint32_t n) { uint32_t p = 1; while (p < n) p <<= 1; return p; }

__global__ void __launch_bounds__(kAsmThreads) assemble_kernel(
    const uint32_t* __restrict__ vox, const float* __restrict__ rad, const uint32_t* __restrict__ seg_start,
    const uint32_t* __restrict__ seg_len, const uint32_t* __restrict__ grp_seg, const uint32_t* __restrict__ grp_out,
    const uint32_t* __restrict__ grp_list, uint32_t* __restrict__ stamp, Dims d, float ax, float ay, float az, float ox, float oy,
    float oz, float* __restrict__ out_verts, float* __restrict__ out_rad, uint32_t* __restrict__ out_edges,
    uint32_t* __restrict__ out_count) {
#ifdef B2T_HOST_EMU
  static AsmShared S;
#else
  extern __shared__ __align__(16) unsigned char b2t_asm_smem[];
  AsmShared& S = *reinterpret_cast<AsmShared*>(b2t_asm_smem);
#endif
  const uint32_t g = grp_list[blockIdx.x];
  const uint32_t s0 = grp_seg[g], s1 = grp_seg[g + 1];
  const uint32_t o0 = grp_out[g], n_ent = grp_out[g + 1] - o0;
  if (n_ent > kAsmCap) {
    if (threadIdx.x == 0) { out_count[2 * g] = 0xffffffffu; out_count[2 * g + 1] = 0; }
    return;
  }
  if (threadIdx.x == 0) { S.n_v = 0; S.n_e = 0; S.n_used = 0; }
  __syncthreads();
  const uint32_t sxy = d.sxy, sx = (uint32_t)d.sx, sy = (uint32_t)d.sy, sz = (uint32_t)d.sz;
  // 1. first occurrence of every voxel: smallest entry position wins
  for (uint32_t s = s0; s < s1; s++) {
    const uint32_t b = seg_start[s], n = seg_len[s];
    for (uint32_t i = threadIdx.x; i < n; i += kAsmThreads) {
      const uint32_t v = vox[b + i];
      if (v != kNone) atomicMin(&stamp[v], b + i);
    }
  }
  __syncthreads();
  for (uint32_t s = s0; s < s1; s++) {
    const uint32_t b = seg_start[s], n = seg_len[s];
    for (uint32_t i = threadIdx.x; i < n; i += kAsmThreads) {
      const uint32_t v = vox[b + i];
      if (v != kNone && stamp[v] == b + i) {
        const uint32_t z = v / sxy, r = v - z * sxy, y = r / sx, x = r - y * sx;
        const uint32_t key = (x * sy + y) * sz + z;                    // < V < 2^32
        S.vk[atomicAdd(&S.n_v, 1u)] = ((unsigned long long)key << 32) | (unsigned long long)(b + i);
      }
    }
  }
  __syncthreads();
  const uint32_t nv = S.n_v;
  const uint32_t pv = asm_pow2(nv);
  for (uint32_t i = nv + threadIdx.x; i < pv; i += kAsmThreads) S.vk[i] = ~0ull;
  __syncthreads();
  asm_bitonic(S.vk, pv);
  // 2. the dense volume now holds the vertex rank of every voxel of the group
  for (uint32_t p = threadIdx.x; p < nv; p += kAsmThreads) stamp[vox[(uint32_t)S.vk[p]]] = p;
  __syncthreads();
  // 3. edges between consecutive path entries
  for (uint32_t s = s0; s < s1; s++) {
    const uint32_t b = seg_start[s], n = seg_len[s];
    for (uint32_t i = threadIdx.x; i + 1 < n; i += kAsmThreads) {
      const uint32_t va = vox[b + i], vb = vox[b + i + 1];
      if (va != kNone && vb != kNone) {
        const uint32_t ra = stamp[va], rb = stamp[vb];
        if (ra != rb) S.ek[atomicAdd(&S.n_e, 1u)] = (min(ra, rb) << 13) | max(ra, rb);
      }
    }
  }
  __syncthreads();
  const uint32_t ne_all = S.n_e;
  const uint32_t pe = asm_pow2(ne_all);
  for (uint32_t i = ne_all + threadIdx.x; i < pe; i += kAsmThreads) S.ek[i] = 0xffffffffu;
  __syncthreads();
  asm_bitonic(S.ek, pe);
  // 4. distinct edges (compacted in place, order kept) and a bitset of the vertices they use
  __shared__ uint32_t s_scan[kAsmThreads];
  __shared__ uint32_t s_used[kAsmCap / 32];
  for (uint32_t i = threadIdx.x; i < kAsmCap / 32; i += kAsmThreads) s_used[i] = 0;
  __syncthreads();
  // unique count + compaction by a block scan over chunks of kAsmThreads sorted edges
  uint32_t ne = 0;
  for (uint32_t base = 0; base < ne_all; base += kAsmThreads) {
    const uint32_t i = base + threadIdx.x;
    uint32_t e = 0;
    bool keep = false;
    if (i < ne_all) {
      e = S.ek[i];
      keep = i == 0 || S.ek[i - 1] != e;
    }
    __syncthreads();                                                 // every thread has read its neighbours before the writes
    s_scan[threadIdx.x] = keep ? 1u : 0u;
    __syncthreads();
    for (int o = 1; o < kAsmThreads; o <<= 1) {
      const uint32_t t = threadIdx.x >= (uint32_t)o ? s_scan[threadIdx.x - o] : 0u;
      __syncthreads();
      s_scan[threadIdx.x] += t;
      __syncthreads();
    }
    if (keep) {
      S.ek[ne + s_scan[threadIdx.x] - 1] = e;                        // ne + rank <= i: never overtakes an unread entry
      atomicOr(&s_used[(e >> 13) >> 5], 1u << ((e >> 13) & 31u));
      atomicOr(&s_used[(e & 0x1fffu) >> 5], 1u << (e & 31u));
    }
    ne += s_scan[kAsmThreads - 1];
    __syncthreads();
  }
  // 5. new rank of a used vertex = number of used vertices before it (exclusive scan over the bitset words)
  __shared__ uint32_t s_wpre[kAsmCap / 32 + 1];
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (uint32_t w = 0; w < kAsmCap / 32; w++) { s_wpre[w] = run; run += __popc(s_used[w]); }
    s_wpre[kAsmCap / 32] = run;
  }
  __syncthreads();
  const uint32_t n_used = s_wpre[kAsmCap / 32];
  auto newrank = [&](uint32_t r) { return s_wpre[r >> 5] + __popc(s_used[r >> 5] & ((1u << (r & 31u)) - 1u)); };
  // 6. outputs; the dense volume goes back to 0xffffffff
  for (uint32_t p = threadIdx.x; p < nv; p += kAsmThreads) {
    const unsigned long long e = S.vk[p];
    const uint32_t pos = (uint32_t)e, key = (uint32_t)(e >> 32);
    const uint32_t v = vox[pos];
    if ((s_used[p >> 5] >> (p & 31u)) & 1u) {
      const uint32_t q = o0 + newrank(p);
      const uint32_t z = key % sz, xy = key / sz, y = xy % sy, x = xy / sy;
      out_verts[3ull * q + 0] = __fmul_rn(__fadd_rn((float)x, ox), ax);   // float32 like intake.py:509-513
      out_verts[3ull * q + 1] = __fmul_rn(__fadd_rn((float)y, oy), ay);
      out_verts[3ull * q + 2] = __fmul_rn(__fadd_rn((float)z, oz), az);
      out_rad[q] = rad[pos];
    }
    stamp[v] = kNone;
  }
  for (uint32_t k = threadIdx.x; k < ne; k += kAsmThreads) {
    const uint32_t e = S.ek[k];
    out_edges[2ull * (o0 + k) + 0] = newrank(e >> 13);
    out_edges[2ull * (o0 + k) + 1] = newrank(e & 0x1fffu);
  }
  if (threadIdx.x == 0) { out_count[2 * g] = n_used; out_count[2 * g + 1] = ne; }
}

}  // namespace

B2T_EXPORT uint32_t b2t_assemble_group_cap(void) { return kAsmCap; }

// d_grp_list: the n_list groups to assemble in this launch (groups of one launch must not share voxels: see above).
B2T_EXPORT int b2t_assemble(const uint32_t* d_vox, const float* d_rad, const uint32_t* d_seg_start, const uint32_t* d_seg_len,
                            const uint32_t* d_grp_seg, const uint32_t* d_grp_out, const uint32_t* d_grp_list, uint32_t n_list,
                            uint32_t* d_stamp, int64_t sx, int64_t sy, int64_t sz, float ax, float ay, float az, float ox,
                            float oy, float oz, float* d_out_verts, float* d_out_rad, uint32_t* d_out_edges,
                            uint32_t* d_out_count, void* stream) {
  if (n_list == 0) return B2T_OK;
  B2T_REQUIRE(d_vox && d_rad && d_seg_start && d_seg_len && d_grp_seg && d_grp_out && d_grp_list && d_stamp && d_out_verts &&
              d_out_rad && d_out_edges && d_out_count, "b2t_assemble: null pointer");
  B2T_REQUIRE(sx > 0 && sy > 0 && sz > 0 && (double)sx * sy * sz < 4294967295.0, "bad volume shape");
  Dims d{(int)sx, (int)sy, (int)sz, (uint32_t)(sx * sy)};
#ifdef B2T_HOST_EMU
  simt::block_launch(n_list, kAsmThreads, [](auto... a_) { assemble_kernel(a_...); })(
      d_vox, d_rad, d_seg_start, d_seg_len, d_grp_seg, d_grp_out, d_grp_list, d_stamp, d, ax, ay, az, ox, oy, oz, d_out_verts,
      d_out_rad, d_out_edges, d_out_count);
#else
  static bool attr_set = false;
  if (!attr_set) {
    B2T_CUDA_TRY(cudaFuncSetAttribute(assemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AsmShared)));
    attr_set = true;
  }
  assemble_kernel<<<n_list, kAsmThreads, sizeof(AsmShared), (cudaStream_t)stream>>>(
      d_vox, d_rad, d_seg_start, d_seg_len, d_grp_seg, d_grp_out, d_grp_list, d_stamp, d, ax, ay, az, ox, oy, oz, d_out_verts,
      d_out_rad, d_out_edges, d_out_count);
  B2T_CUDA_TRY(cudaGetLastError());
#endif
  b2t_count_launches(1);
  return B2T_OK;
}
