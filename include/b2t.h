/*
 * b2t.h -- C ABI of libb2t.so, the B200-native (sm_100a) TEASAR hot path.
 *
 * Drop-in boundary for the per-label trace of seung-lab/kimimaro (kimimaro/trace.py) and the
 * EDT that feeds it.  Every entry point replaces one native call the reference makes on that
 * path; the reference-side binding a maintainer would add is a ctypes stub.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / numpy types.
 *   - pointers named d_* are DEVICE pointers (caller-owned, e.g. tensor.data_ptr());
 *     pointers named h_* are HOST pointers.  The library allocates nothing persistent.
 *   - volumes are Fortran-ordered: loc = x + sx*(y + sy*z)  (kimimaro/intake.py:320-322,
 *     ext/skeletontricks/skeletontricks.pyx:398).
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *   - every function returns B2T_OK (0) or a negative status; b2t_last_error() describes it.
 *   - there is NO CPU fallback: without an sm_100 device every call fails with B2T_ERR_DEVICE.
 */
#ifndef B2T_H
#define B2T_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2T_OK 0
#define B2T_ERR_ARG (-1)       /* bad argument (shape, dtype width, null pointer) */
#define B2T_ERR_DEVICE (-2)    /* no CUDA device / wrong architecture */
#define B2T_ERR_CUDA (-3)      /* a CUDA runtime call failed, see b2t_last_error() */
#define B2T_ERR_CAPACITY (-4)  /* caller-provided buffer too small */

/* library / device ------------------------------------------------------------------------- */
int b2t_version(void);                 /* e.g. 100 = 0.1.0 */
const char* b2t_last_error(void);      /* thread-local message of the last failure */
int b2t_device_check(void);            /* B2T_OK iff current device is compute capability 10.x */
unsigned long long b2t_launch_count(int reset); /* kernels launched by this library so far */
/* cap resident blocks per SM of the cooperative sweeps / of the path-loop kernel (0 = no cap) so that two
 * arenas can be traced concurrently on two streams */
int b2t_set_launch_limits(int coop_blocks_per_sm, int trace_blocks_per_sm);
/* Claim order of roll_invalidation_ball_inside_component inside the path loop (b2t_trace_batch), DESIGN.md 4:
 *   WINDOW  parallel rounds ordered by the reference's heap key -- distance to the seed, dijkstra_invalidation.hpp:233-237
 *           -- in windows of `claim_window_voxels` smallest-voxel-edges (the default of kimimaro_b200: 1)
 *   STRICT  the reference's std::priority_queue literally, libstdc++'s order of equal keys included: identical to the
 *           compiled reference voxel for voxel; sequential per label (one warp), needs the heap buffer
 *   ROUNDS  hop-synchronous rounds (round 1's order) */
#define B2T_INVALIDATE_ROUNDS 0
#define B2T_INVALIDATE_WINDOW 1
#define B2T_INVALIDATE_STRICT 2

/* K1  anisotropic multi-label Euclidean distance transform ------------------------------------
 * replaces  edt.edt(labels, anisotropy, black_border)            kimimaro/intake.py:174-185
 *           edt.edt(labels, anisotropy, black_border=np.all())   kimimaro/trace.py:112-117
 *           edt.edt(cc_plane, black_border=True, anisotropy=(wx,wy))  kimimaro/intake.py:565
 * d_labels: [sx,sy,sz] unsigned ints of label_bytes in {1,2,4,8}; d_out: float32 [sx,sy,sz].
 * ndim = 2 runs the x and y passes only (sz must be 1), ndim = 3 all three passes.
 * Three launches: pass x (run scan), pass y, pass z (lower envelope of parabolas per run, one thread per
 * column); sqrt fused into the last one.  d_out is used in place between passes; no workspace. */
int b2t_edt(const void* d_labels, int label_bytes, int64_t sx, int64_t sy, int64_t sz,
            float wx, float wy, float wz, int black_border, int ndim, float* d_out, void* stream);
/* Tuning hook, no reference counterpart: picks the column-pass kernel of b2t_edt (results are identical).
 * algo: 0 = keep, 1 = windowed search, 2 = F-H with a local-memory stack, 3 = F-H with a shared-memory ring
 * (default).  c > 0 selects the compiled instantiation (ring entries, min blocks per SM, rows between flushes,
 * rows per load batch) of algo 3; an instantiation that is not compiled in makes the next b2t_edt fail. */
int b2t_edt_config(int algo, int c, int minb, int r, int b);
/* K1 with a caller-provided workspace (the form SURVEY 8b proposes): same result as b2t_edt, bit for bit.
 * With uint32 labels and integer anisotropy the column passes run as a hybrid: a register-window min-plus
 * stencil over every column (exact wherever the result is at most w^2 (W+1)^2, i.e. on the thin processes a
 * connectomics volume is made of) and the envelope kernel only over the 32-row blocks of 32-column tiles
 * the stencil flagged (the inside of blobs).  The passes ping-pong between d_out and the workspace:
 *   workspace = [sx*sy*sz float32][per-tile 64-bit block flags of the y and z passes][the same again: predictions],
 * b2t_edt_workspace_bytes() says how much.  Any other input, or a workspace that is null / too small, takes the
 * b2t_edt path (d_out in place).  Launches: memset, pass x, (stencil, envelope) x 2. */
size_t b2t_edt_workspace_bytes(int64_t sx, int64_t sy, int64_t sz);
int b2t_edt_ws(const void* d_labels, int label_bytes, int64_t sx, int64_t sy, int64_t sz,
               float wx, float wy, float wz, int black_border, int ndim, float* d_out,
               void* d_workspace, size_t workspace_bytes, void* stream);
/* Tuning hook of the hybrid: enable = 0 sends b2t_edt_ws down the b2t_edt path; wy, wz = tap radius of the stencil
 * in the y / z pass, wr = radius of its register window (farther taps read the shared-memory ring), pf = rows of
 * load prefetch, minb = min blocks per SM.  wy > 0 fixes the radii, wy = 0 leaves them to the library (they follow
 * the anisotropy), wy < 0 keeps them; wr > 0 selects a compiled (wr, pf, minb) instantiation. */
int b2t_edt_config_hybrid(int enable, int wy, int wz, int wr, int pf, int minb);
/* Tuning hook: enable = 1 runs every column pass of the hybrid as ONE "roles" launch plus a residual one: the previous
 * pass predicts the blocks the stencil cannot finish (x pass -> y, y pass -> z), envelope warps start on them at once
 * while stencil warps do the rest and flag what the prediction missed.  Same result bit for bit; the workspace of
 * b2t_edt_workspace_bytes() already has room for the prediction words.  stencil_v2 = 1 runs the stencil (roles or not)
 * with the leaner steady-state loop (labels through a register ring, one shared 32-bit row offset); same result.
 * predict_scale >= 1 scales the thresholds of the prediction (1 = every block that can need the envelope). */
int b2t_edt_config_roles(int enable, int stencil_v2, float predict_scale);
/* Tuning hook of the hybrid's envelope kernel: query_prefetch = 4 keeps the next four stack entries of the write-out in
 * registers (the stack of a blob lives in local memory; its walk was one L2 round trip per row), 1 = the original loop.
 * pop_ahead = 1 keeps the entry below the top of the stack in registers during the build (a pop then needs no load before
 * the next intersection).  Same result. */
/* x pass of b2t_edt_ws: 1 = TMA-staged tiles (cp.async.bulk.tensor loads / stores of 256 x 8 boxes; rows of 256 or 512
 * labels), 0 = the register-only kernel; same bits.  Returns the number of TMA x-pass launches so far (tma < 0: query). */
long long b2t_edt_config_xpass(int tma);
int b2t_edt_config_envelope(int query_prefetch, int pop_ahead);


/* N1  connected components ------------------------------------------------------------------------
 * replaces  cc3d.connected_components(labels) (26-connected, multi-label; 2-D: 8-connected)
 *           kimimaro/utility.py:77, kimimaro/intake.py:564
 * Two steps so that the caller ranks the component roots (a prefix sum in raster order):
 *   b2t_ccl26_roots: d_parent[v] = smallest linear index of v's component (0xffffffff on background),
 *                    d_is_root[v] = 1 on that voxel.  Numbering by first appearance == cc3d's order.
 *   b2t_ccl_relabel: d_parent[v] <- d_rank[d_parent[v]] (0 on background): the cc label volume. */
int b2t_ccl26_roots(const void* d_labels, int label_bytes, int64_t sx, int64_t sy, int64_t sz,
                    uint32_t* d_parent, uint8_t* d_is_root, void* stream);
int b2t_ccl_relabel(uint32_t* d_parent, const int32_t* d_rank, uint64_t n_voxels, void* stream);

/* per-label statistics in one pass -----------------------------------------------------------------
 * replaces  fastremap.unique(cc_labels, return_counts=True)      kimimaro/intake.py:198
 *           scipy.ndimage.find_objects (bounding boxes)          kimimaro/utility.py:85-102
 *           np.max(DBF) per cropped label                        kimimaro/trace.py:100
 *           skeletontricks.first_label                           ext/skeletontricks/skeletontricks.pyx:307-326
 * tables have n_labels+1 entries (row 0 = background, unused); d_bbox rows are x0,y0,z0,x1,y1,z1 inclusive. */
int b2t_label_stats(const uint32_t* d_cc, const float* d_dbf, int64_t sx, int64_t sy, int64_t sz,
                    uint32_t n_labels, uint32_t* d_count, int32_t* d_bbox, float* d_dbfmax,
                    uint32_t* d_first, void* stream);

/* K2  geometric distance field, all labels at once ----------------------------------------------------
 * replaces  dijkstra3d.euclidean_distance_field(labels, source, anisotropy, free_space_radius,
 *           return_max_location=True)           kimimaro/trace.py:139-145 (DAF), :302-307 (find_root)
 * One source voxel per participating label (d_sources, n_sources); relaxation only between voxels of
 * equal d_cc value.  d_dist [V] must hold +inf and d_stamp [V] zero on entry; d_queue holds
 * 2*queue_cap u32 (queue_cap >= foreground voxels of the participating labels); d_ctrl >= 8 u32.
 * free_space_radius > 0 (soma labels, trace.py:134) needs n_sources == 1 and the source's linear
 * index in h_free_space_source.
 * d_node_weights != NULL turns the sweep into dijkstra3d.parental_field(PDRF, root) (kimimaro/trace.py:155,
 * fix_branching=False): the cost of entering voxel v is d_node_weights[v]; parents follow from the distances. */
int b2t_edf_multi(const uint32_t* d_cc, int64_t sx, int64_t sy, int64_t sz, float wx, float wy, float wz,
                  const uint32_t* d_sources, uint32_t n_sources, float free_space_radius,
                  uint32_t h_free_space_source, const float* d_node_weights, float* d_dist, uint32_t* d_stamp,
                  uint32_t* d_queue, uint64_t queue_cap, uint32_t* d_ctrl, void* stream);

/* The same field, label by label instead of one grid-wide sweep: d_jobs holds n_jobs records of four u32
 * (source voxel, cc id, voxel count n_fg, prefix sum of n_fg over the preceding records), sorted by n_fg descending.
 * The first n_team jobs get a thread-block cluster (8 CTAs, hardware cluster barrier per relaxation round), the others one
 * CTA each (block barrier); every label runs exactly its own number of rounds instead of the volume's maximum.
 * d_queue: 2*sum(n_fg) u32; d_ctrl: 4*n_team u32; d_dist / d_stamp / d_node_weights as above.  Bit-identical result. */
int b2t_edf_labels(const uint32_t* d_cc, int64_t sx, int64_t sy, int64_t sz, float wx, float wy, float wz,
                   const uint32_t* d_jobs, uint32_t n_jobs, uint32_t n_team, const float* d_node_weights,
                   float* d_dist, uint32_t* d_stamp, uint32_t* d_queue, uint32_t* d_ctrl, void* stream);

/* return_max_location of the call above: d_best[l] = (dist_bits << 32) | (0xffffffff - index) of the
 * largest finite distance of label l, smallest index on ties; 0 if the label has none. */
int b2t_field_argmax(const uint32_t* d_cc, const float* d_dist, int64_t sx, int64_t sy, int64_t sz,
                     uint32_t n_labels, uint64_t* d_best, void* stream);

/* K3  penalised distance field + target cache ---------------------------------------------------------
 * replaces  skeletontricks.zero2inf / inf2zero     kimimaro/trace.py:138,146 (pyx:177-224)
 *           compute_pdrf                          kimimaro/trace.py:315-356
 *           CachedTargetFinder.__init__           ext/skeletontricks/skeletontricks.pyx:995-1006
 * d_dist holds DAF on entry and +inf on exit; d_pdrf / d_claim are initialised on every voxel of
 * an active label; d_keys receives (daf_bits << 32 | index) partitioned into nbuckets DAF buckets per
 * label: bucket b of the label in table row r is d_keys[d_cursor[r*nb+b] - d_hist[r*nb+b] .. d_cursor[r*nb+b]).
 * d_row[l] (l = 0 .. n_labels) is the table row of cc label l, 0xffffffff for a label that does not take part; the
 * tables have n_rows rows (the traced labels, not the cc ids: dust components cost nothing).
 * d_M[l] = float32(1 / dbf_max**1.01), d_inv_maxdaf[l] = 1 / DAF[target] (0 if that is 0): computed by
 * the host with the reference's numpy expressions (trace.py:336, 353). */
int b2t_pdrf_and_buckets(const uint32_t* d_cc, const float* d_dbf, float* d_dist, float* d_pdrf,
                         uint64_t* d_claim, int64_t sx, int64_t sy, int64_t sz,
                         uint32_t n_labels, const float* d_M, const float* d_inv_maxdaf,
                         const uint32_t* d_row, uint32_t n_rows, float pdrf_scale, float pdrf_exponent, int nbuckets,
                         uint32_t* d_hist, uint32_t* d_cursor, uint64_t* d_keys, void* stream);

/* K4 + K5  the path loop ----------------------------------------------------------------------------------
 * replaces  compute_paths                                   kimimaro/trace.py:196-267, including
 *           CachedTargetFinder.find_target                  ext/skeletontricks/skeletontricks.pyx:1008-1045
 *           dijkstra3d.railroad(PDRF, target)               kimimaro/trace.py:240-242
 *           roll_invalidation_ball_inside_component         pyx:373-418 -> dijkstra_invalidation.hpp:239-332
 *           the soma cull and soma invalidation             kimimaro/trace.py:160-168, 246-251
 *           dijkstra3d.path_from_parents (fix_branching == 0: d_dist must hold the parental field) trace.py:244
 * d_desc: n_desc records of 80 bytes (20 little-endian 32-bit fields):
 *   segid, root, n_fg, region_off, path_off, path_cap, tb_off, tb_n, ta_off, ta_n, max_paths (0xffffffff =
 *   None), soma_mode, soma_radius (float32), bucket_row, soma_done, pre_invalid, bbox_x0, bbox_x1 (x extent of the
 *   label's bounding box: the reference runs on that crop, intake.py:463-466, and STRICT reproduces the duplicate
 *   pushes of its neighbour table at the crop's x faces), single_path (1: exactly one path, to the first manual target,
 *   nothing invalidated -- dijkstra3d.dijkstra(PDRF, end, start) of trace.point_to_point, kimimaro/trace.py:358-390, with
 *   fix_branching == 0, root = end and d_dist = the node-weighted field grown from end), one reserved word
 * d_scratch: b2t_trace_scratch_words(sum(n_fg)) u32; d_paths: pool of voxel indices, each path [rail ... target] terminated by
 * 0xffffffff; d_out_len / d_out_npaths / d_out_status: n_desc; d_out_stats: 4*n_desc; d_work_counter: 1 u32.
 * invalidation_mode: B2T_INVALIDATE_*; claim_window_voxels: width of a WINDOW round.  STRICT only: d_heap holds heap_words
 * u32, of which the first heap_static_words = b2t_trace_heap_words(sum(n_fg), n_desc) are the labels' own heap regions
 * (4 entries per voxel) and the rest a spill arena for heaps that outgrow theirs (up to 27 entries per voxel of the
 * label); a label whose heap fits nowhere reports B2T_ERR_CAPACITY in d_out_status.  Other modes: NULL, 0, 0.
 * n_team: the first n_team records (the caller sorts by n_fg, largest first) are traced by a thread-block cluster of
 * 8 CTAs each instead of one CTA -- a label's paths are sequential, so the largest label is the tail of the launch;
 * d_team: n_team * b2t_trace_team_bytes() bytes of device memory for the teams' shared state. */
uint64_t b2t_trace_heap_words(uint64_t sum_n_fg, uint64_t n_desc);
int b2t_trace_batch(const uint32_t* d_cc, const float* d_dbf, float* d_pdrf, float* d_dist, uint64_t* d_claim,
                    uint32_t* d_stamp, int64_t sx, int64_t sy, int64_t sz, float wx, float wy, float wz,
                    const void* d_desc, int n_desc, float scale, float konst, float soma_scale,
                    float soma_const, int fix_branching, int nbuckets, const uint64_t* d_keys, const uint32_t* d_hist,
                    const uint32_t* d_cursor, uint32_t* d_scratch, uint32_t* d_paths,
                    const uint32_t* d_targets, uint32_t* d_out_len, uint32_t* d_out_npaths,
                    int32_t* d_out_status, uint32_t* d_out_stats, uint32_t* d_work_counter,
                    int invalidation_mode, float claim_window_voxels, uint32_t* d_heap, uint64_t heap_words,
                    uint64_t heap_static_words, int n_team, void* d_team, void* stream);
uint64_t b2t_trace_team_bytes(void);
uint64_t b2t_trace_scratch_words(uint64_t sum_n_fg);

/* the same rolling-ball invalidation, grid-wide, for balls too large for one CTA (the one-off soma
 * invalidation, kimimaro/trace.py:160-168).  d_fv / d_fs: 2*cap u32 each; count left in d_ctrl[6]. */
int b2t_invalidate_ball(const uint32_t* d_cc, const float* d_dbf, uint64_t* d_claim, int64_t sx, int64_t sy,
                        int64_t sz, float wx, float wy, float wz, const uint32_t* d_seeds, uint32_t n_seeds,
                        float scale, float konst, uint32_t* d_fv, uint32_t* d_fs, uint64_t cap,
                        uint32_t* d_ctrl, void* stream);

/* the same for ONE seed (what the soma call is), as a connected-components problem instead of a frontier sweep: the
 * claimed set is the component of {label voxels closer to the seed than its radius} that holds the seed.
 * d_mark [V] u8, d_parent [V] u32, d_is_root [V] u8: scratch; count left in d_ctrl[6] (d_ctrl >= 8 u32). */
int b2t_invalidate_ball_single(const uint32_t* d_cc, const float* d_dbf, uint64_t* d_claim, int64_t sx, int64_t sy,
                               int64_t sz, float wx, float wy, float wz, uint32_t h_seed, float scale, float konst,
                               uint8_t* d_mark, uint32_t* d_parent, uint8_t* d_is_root, uint32_t* d_ctrl, void* stream);

/* N2  fix_borders, the per-component reductions of one face -------------------------------------------------------
 * replaces  skeletontricks.find_border_targets (DT maximum per face component, the voxels that attain it, first raster
 *           position; pyx:591-648), compute_centroids (coordinate sums and counts; pyx:528-588) and get_mapping (the
 *           volume label under a face component; pyx:490-525) as called from kimimaro/intake.py:544-585.
 * d_cc_plane / d_dt / d_plane: [p0*p1] face components (b2t_ccl26_roots with sz = 1), their 2-D EDT, the volume's cc labels
 * on the face.  d_tab: 6*(P+1) u32 scratch; d_cand: 2*P u32 (position, component) pairs; d_rec: 7*P u32 records
 * (component, max DT bits, first position, count, sum x, sum y, volume label); d_count: candidates, records.  The tie-break
 * among a component's candidates stays on the host (float32 / float64 expression order of pyx:650-760). */
int b2t_face_stats(const uint32_t* d_cc_plane, const float* d_dt, const uint32_t* d_plane, int64_t p0, int64_t p1,
                   uint32_t* d_tab, uint32_t* d_cand, uint32_t* d_rec, uint32_t* d_count, void* stream);

/* N2 helper: strictly sequential float32 sums per segment (label centroids of a face in the reference's
 * accumulation order, ext/skeletontricks/skeletontricks.pyx:528-588 compute_centroids). */
int b2t_segment_seqsum(const float* d_xs, const float* d_ys, const int64_t* d_off, uint32_t n_seg,
                       float* d_outx, float* d_outy, void* stream);

/* skeleton buffers: compact the path segments and fetch radii = DBF[vertex] (kimimaro/trace.py:186-187) */
int b2t_gather_paths(const uint32_t* d_pool, const uint32_t* d_src_off, const uint32_t* d_len,
                     const uint64_t* d_dst_off, uint32_t n_seg, const float* d_dbf, uint32_t* d_dst_vox,
                     float* d_dst_radius, void* stream);

/* Skeleton assembly for groups of path segments (kimimaro/trace.py:182-192 Skeleton.from_path / simple_merge /
 * consolidate, kimimaro/intake.py:509-517, 587-593): per group the distinct path voxels in lexicographic (x, y, z)
 * order as float32 physical coordinates ((voxel + offset) * anisotropy), their radii (DBF at the first occurrence),
 * the distinct sorted edges between consecutive path voxels, vertices without an edge dropped.  d_stamp: one u32 per
 * voxel of the volume, 0xffffffff on entry and on return.  Groups of one call must not share voxels.  A group with more
 * than b2t_assemble_group_cap() entries reports n_vertices = 0xffffffff and is left to the caller. */
uint32_t b2t_assemble_group_cap(void);
int b2t_assemble(const uint32_t* d_vox, const float* d_rad, const uint32_t* d_seg_start, const uint32_t* d_seg_len,
                 const uint32_t* d_grp_seg, const uint32_t* d_grp_out, const uint32_t* d_grp_list, uint32_t n_list,
                 uint32_t* d_stamp, int64_t sx, int64_t sy, int64_t sz, float ax, float ay, float az, float ox,
                 float oy, float oz, float* d_out_verts, float* d_out_rad, uint32_t* d_out_edges,
                 uint32_t* d_out_count, void* stream);

/* K6  hole filling (soma labels only) -----------------------------------------------------------------------
 * replaces  fill_voids.fill(labels, in_place=True, return_fill_count=True)    kimimaro/trace.py:109
 * d_mask uint8 [V] edited in place; d_reach [V] u32 scratch; d_queue >= V u32 scratch (queue_cap >= V);
 * the fill count is left in d_ctrl[5]. */
int b2t_fill_voids(uint8_t* d_mask, int64_t sx, int64_t sy, int64_t sz, uint32_t* d_reach,
                   uint32_t* d_queue, uint64_t queue_cap, uint32_t* d_ctrl, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B2T_H */
