/*
 * graphem_b200 -- C ABI of the B200 (sm_100a) implementation of GraphEm's force-directed
 * layout iteration.
 *
 * The reference (sashakolpakov/graphem-rapids v0.2.0) is pure Python and has no FFI: the
 * boundary a caller sees is the class GraphEmbedderPyTorch
 * (graphem_rapids/backends/embedder_pytorch.py).  Each entry point below replaces the
 * torch-op body of one of its methods; the Python host graphem_rapids_b200.embedder binds
 * them with ctypes (see INTEGRATION.md for the stub a maintainer of the reference would add).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; the library never
 *     allocates, frees or synchronises: scratch comes from the caller-provided workspace,
 *     every launch goes to `stream` (a cudaStream_t passed as void*), so every call is
 *     CUDA-graph capturable;
 *   - return value: 0 on success, otherwise a negative GEM_E_* code or the positive
 *     cudaError_t of the failed launch; gem_error_string() describes both;
 *   - vertex positions / forces are row-major (n, ld) fp32 with ld = gem_row_pitch(d):
 *     2 for d=2, 4 for d=3 (lane 3 is zero padding so one 128-bit access moves a row),
 *     d otherwise;
 *   - midpoints are row-major (e, gem_mid_pitch(d)) fp32: d=2 -> (x, y);
 *     d=3 -> (x, y, z, |m|^2); other d -> d coordinates followed by |m|^2;
 *   - edge ids are int32 pairs (i<j, sorted by (i,j)) -- n < 2^31 at every target size; the
 *     public `edges` attribute of the Python class stays int64;
 *   - neighbour lists are int64 (global edge ids), ascending by (distance, index).
 */
#ifndef GRAPHEM_B200_H
#define GRAPHEM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GEM_ABI_VERSION 3

#define GEM_OK 0
#define GEM_E_BADARG (-1)      /* null pointer, negative size, unsupported d */
#define GEM_E_WORKSPACE (-2)   /* workspace too small / misaligned */
#define GEM_E_KRANGE (-3)      /* k+1 > number of candidates (the reference's topk RuntimeError) */
#define GEM_E_NODEVICE (-4)    /* no sm_100 device / kernel image not loadable */
#define GEM_E_BUSY (-5)        /* gem_coef_slot_acquire: every coefficient slot of the device is owned */

int gem_abi_version(void);
/* sizeof of the four parameter structs below, in declaration order (gem_plan, gem_knn_prep_args, gem_knn_publish,
 * gem_merge_publish): lets a foreign-language binding verify its mirror of the layouts at load time. */
int gem_abi_struct_sizes(size_t *out4);
/* Once per process and device, before the first launch (sets kernel attributes, creates the side
 * stream of gem_layout_step; not capturable). */
int gem_init(void);
const char *gem_error_string(int code);
/* row pitch (floats) of positions/forces and of midpoints for n_components = d */
int gem_row_pitch(int d);
int gem_mid_pitch(int d);

/* (a) spring forces + midpoints.  Replaces _compute_spring_forces (embedder_pytorch.py:595-636)
 * and the midpoint expression of update_positions (:785).
 * force (n, ld) is zeroed inside, then F[i] += f, F[j] -= f per edge.  mid may be NULL.
 * `edges` points at the first edge of the (shard of the) list, `e` edges long. */
int gem_spring_midpoints(const float *pos, const int32_t *edges, int64_t n, int64_t e, int d,
                         float k_attr, float l_min, float *force, float *mid, void *stream);

/* Same stage in vertex-parallel ("pull") form over the symmetric CSR of the graph -- no atomics,
 * no zero-fill, deterministic: a group of lanes per vertex v sums the terms of its incident edges
 * and writes force[v - v_begin]; the row entries w > v are the edges (v,w) with ids
 * up_ptr[v] .. up_ptr[v+1]-1 (edge list sorted by (i,j)), whose midpoints are written to
 * mid[(id - mid_base)].  Only d = 2, 3.  The multi-GPU path gives each rank a vertex range
 * [v_begin, v_end): it produces COMPLETE forces for its vertices and the midpoints of the edges
 * whose first endpoint it owns, so no force reduction across ranks is needed.
 *   row_ptr (n+1) int64, col (2e) int32 ascending per row, up_ptr (n+1) int64 (#edges with first
 *   endpoint < v); hubs: ids (within the range) of the vertices with degree > gem_hub_degree(),
 *   handled by one CTA each; force -> row v_begin of the accumulator; mid -> row of edge mid_base. */
int gem_hub_degree(void);
int gem_spring_midpoints_csr(const float *pos, const int64_t *row_ptr, const int32_t *col, const int64_t *up_ptr,
                             int64_t v_begin, int64_t v_end, const int32_t *hubs, int64_t n_hubs, int d,
                             float k_attr, float l_min, float *force, float *mid, int64_t mid_base, void *stream);

/* The same kernel in its fused spring+update form: `newpos` (row 0 = vertex v_begin) receives
 * pos + F_spring -- the add of the update's pass 1 (:799) without materialising F.  Follow with
 * gem_update_positions(newpos, ..., phase = 3) for the column sums, add the intersection forces with
 * the sum-correcting form (gem_topk_merge_intersect, or the select kernel inside gem_layout_step) and
 * finish with gem_update_normalise_push or pass 2. */
int gem_spring_update_csr(const float *pos, const int64_t *row_ptr, const int32_t *col, const int64_t *up_ptr,
                          int64_t v_begin, int64_t v_end, const int32_t *hubs, int64_t n_hubs, int d,
                          float k_attr, float l_min, float *newpos, float *mid, int64_t mid_base, void *stream);
/* Multi-GPU form of gem_spring_update_csr, fused with the exchange of the updated positions: the new row
 * pos + F_spring of every vertex v in [v_begin, v_end) is stored at row v of EACH of the `world` raw
 * (unnormalised) position buffers peer_raw_host[r] (host array of device pointers: the rank's own buffer and the
 * peers' buffers mapped into this process, e.g. torch.distributed._symmetric_memory buffer_ptrs).  The stores to
 * the peers travel over NVLink P2P while the rank goes on with its KNN scan: the all-gather of the positions is
 * the store phase of the FIRST kernel of the iteration, not a collective behind the last one.  Rows that later
 * receive intersection forces are re-published by gem_topk_merge_intersect; after one cross-rank barrier every rank
 * normalises all rows locally (gem_update_normalise_push with world = 1).  The caller double-buffers the raw
 * buffers by iteration parity (a fast rank may start iteration t+1 while a slow one still normalises t).
 * multicast_raw (optional): the NVSwitch multicast mapping of the same raw buffers (symmetric-memory multicast_ptr):
 * one multimem.st per row is replicated by the switch into every replica (the rank's own, peer_raw_host[rank], is
 * also written directly) instead of world-1 unicast stores.
 * work (optional): 2 x uint32, zero-initialised once by the caller -- the kernel then claims its vertex ranges
 * dynamically (it shares the SMs with the KNN preparation kernel; a static split leaves late CTAs a full share). */
int gem_spring_update_csr_push(const float *pos, const int64_t *row_ptr, const int32_t *col, const int64_t *up_ptr,
                               int64_t v_begin, int64_t v_end, const int32_t *hubs, int64_t n_hubs, int d,
                               float k_attr, float l_min, float *const *peer_raw_host, int world, int rank,
                               float *multicast_raw, float *mid, int64_t mid_base, void *work, void *stream);

/* Sampling of the query edges.  Replaces `torch.randperm(E, device)[:S]` / `arange(E)`
 * (_locate_knn_midpoints, :404-413) by a keyed bijection of [0,e) evaluated at 0..s-1
 * (Feistel network with cycle walking, key = (seed, *iter_counter)); s distinct ids, no host
 * sync, replayable from a CUDA graph.  When s >= e the result is arange(e).
 * If bump_counter != 0 the kernel increments *iter_counter afterwards. */
int gem_sample_edges(uint64_t seed, int64_t *iter_counter, int bump_counter, int64_t e, int64_t s,
                     int64_t *samp, void *stream);

/* Midpoints of the s sampled edges, computed from positions (identical bits to what
 * gem_spring_midpoints writes): replaces `midpoints[sampled_indices]` (:410). */
int gem_query_midpoints(const float *pos, const int32_t *edges, const int64_t *samp, int64_t s, int d,
                        float *qmid, void *stream);

/* (b) brute-force KNN over midpoints.  Replaces _compute_knn_chunked / _compute_knn_torch
 * (:426-483, :543-593: torch.cdist + torch.topk(k+1, largest=False)).
 * Distances reproduce torch.cdist bit for bit: matmul-mode FMA chain when s>25 or e_total>25,
 * direct mode otherwise (mm_mode: 1 / 0; pass -1 to choose like torch does from s and e).
 * out_idx (s, kp1) int64 = idx_offset + local candidate id, out_dist (s, kp1) fp32, each row
 * ascending by (distance, index).  Column 0 is NOT dropped here (:421 is applied by the consumer).
 * gem_knn_workspace_bytes sizes `ws` (256-byte aligned). */
int gem_knn_workspace_bytes(int64_t e, int d, int64_t s, int kp1, size_t *bytes);
int gem_knn_midpoints(const float *mid, int64_t e, int64_t idx_offset, int d, const float *qmid,
                      int64_t s, int kp1, int mm_mode, const float *tau_hint, int64_t *out_idx,
                      float *out_dist, void *ws, size_t ws_bytes, int coef_slot, void *stream);
/* Shard-local form for the multi-GPU path: `mid` holds e of the e_total candidates of the whole
 * problem (ids idx_offset .. idx_offset+e-1).  The k-range check and torch.cdist's mode choice use
 * e_total; rows with fewer than kp1 local candidates (or none: e == 0, mid may be NULL) are padded
 * with (distance +inf, index -1), which gem_topk_merge orders last. */
int gem_knn_midpoints_shard(const float *mid, int64_t e, int64_t e_total, int64_t idx_offset, int d,
                            const float *qmid, int64_t s, int kp1, int mm_mode, const float *tau_hint,
                            int64_t *out_idx, float *out_dist, void *ws, size_t ws_bytes, int coef_slot, void *stream);
/* Diagnostics: kernel time stamps.  With a device buffer of gem_debug_stamp_words() uint64 registered (NULL: off,
 * the default), every CTA of the iteration's kernels folds %globaltimer (ns) into [2*id] (atomicMin: first CTA start)
 * and [2*id+1] (atomicMax: last CTA end); ids: 0 knn_prep, 1 spring_csr, 2 column sums, 3 knn_scan, 4 knn_select,
 * 5 topk_merge_intersect, 6 normalise.  The caller resets the buffer ([2*id] = ~0, [2*id+1] = 0) before the step it
 * wants to see.  This is how the per-kernel timeline inside a multi-rank CUDA-graph replay is measured.
 * Synchronous (cudaMemcpyToSymbol): call outside capture. */
int gem_debug_stamps(unsigned long long *buffer);
int gem_debug_stamp_count(void);
/* size of the buffer in uint64 words (2 per kernel id).  (Per-CTA stamps inside the scan kernel were tried and removed:
 * one extra global store in its prologue made ptxas drop the uniform-register operands of the whole main loop.) */
int gem_debug_stamp_words(void);

/* Coefficient slots.  The scan reads the query coefficients (-2q) as uniform-register operands from a __constant__
 * table; the table has gem_coef_slots() independent slots per device, and every KNN entry point below takes the
 * slot it may use (`coef_slot`).  An object that issues KNN work owns one slot for its lifetime
 * (gem_coef_slot_acquire on the current device -> 0 .. slots-1, or GEM_E_BUSY; gem_coef_slot_release), so two
 * embedders (or two captured CUDA graphs) on different streams of one GPU never share filter coefficients.
 * (Round 1 had ONE table per device: concurrent KNNs silently filtered with each other's queries.) */
int gem_coef_slots(void);
int gem_coef_slot_acquire(int *slot);
int gem_coef_slot_release(int slot);

/* The two phases of the fast path of gem_knn_midpoints[_shard] as separate calls, for callers that overlap them
 * with other work (gem_layout_step does so internally; the multi-GPU host does it with its own side stream).
 *
 * gem_knn_prep -- ONE kernel launch + one device-to-device copy: draws the sample (draw != 0: the keyed bijection of
 *   gem_sample_edges from (seed, *iter_counter); bump != 0: *iter_counter += 1 at the end), computes the query
 *   midpoints from (pos, edges, samp) (or takes them from qmid_in), the line-graph bound of every query (row_ptr /
 *   col given: see gem_knn_linegraph_hint), a stratified bound pass over `bound_samples` of the e_bound bound
 *   candidates (0: 15*sqrt((k+1)*e_bound), the size that balances the pass against the scan's slow path) -- taken
 *   from bound_mid, or recomputed from (pos, bound_edges) so that the call depends on the positions only and can
 *   run while the spring kernel is still producing the midpoints --, the filter thresholds and the coefficient
 *   pairs of slot coef_slot, and zeroes the scan's counters.  The bound candidates need not be the scan's
 *   candidates: any k+1 distinct candidates of the WHOLE problem bound the (k+1)-th neighbour distance, so the
 *   multi-GPU path bounds with the full edge list on every rank (identical thresholds everywhere) and scans its
 *   shard.  `e` = the candidate count of the gem_knn_scan that follows (it fixes the workspace layout).
 * gem_knn_scan -- all-pairs scan + select -> (s, kp1) lists (short rows padded with +inf / -1).  pub (optional):
 *   remap = table local candidate number -> global edge id (NULL: idx_offset + number); world > 0: every row is also
 *   stored into each rank's exchange buffer (peer-mapped base pointers, the rank's own included) at
 *   idx_offset_bytes (int64 rows) / dist_offset_bytes (fp32 rows): the rank's partial list is published by the
 *   kernel that produces it.
 * Valid only when gem_knn_fast_path(e, e_total, d, s, kp1) returns 1 (matmul-mode cdist, d in {2,3},
 * e >= 2048, k+1 <= 56 and <= e, s <= 1024); otherwise use gem_knn_midpoints_shard. */
int gem_knn_fast_path(int64_t e, int64_t e_total, int d, int64_t s, int kp1);
typedef struct gem_knn_prep_args {
    int32_t d, kp1;
    int64_t s;                  /* queries of this batch (<= 1024) */
    int64_t e;                  /* candidates the following gem_knn_scan sees */
    const float *pos;           /* (n, ld); may be NULL when qmid_in and bound_mid are given */
    const int32_t *edges;       /* (e_total, 2): the edge list the query ids refer to */
    int64_t e_total;
    int64_t *samp;              /* (s) in (draw == 0) / out (draw != 0); unused with qmid_in */
    int32_t draw, bump;
    uint64_t seed;
    int64_t *iter_counter;
    const float *qmid_in;       /* (s, mld) precomputed query midpoints, or NULL */
    float *qmid_out;            /* (s, mld) out, when qmid_in == NULL */
    const int64_t *row_ptr;     /* symmetric CSR for the line-graph bound, or NULL */
    const int32_t *col;
    const float *tau_hint_in;   /* (s) caller-provided bound, or NULL */
    float *tau_hint_out;        /* (s) scratch of the line-graph bound (required with row_ptr) */
    const float *bound_mid;     /* (e_bound, mld) bound candidates, or NULL: recomputed from (pos, bound_edges) */
    const int32_t *bound_edges; /* (e_bound, 2) */
    int64_t e_bound;
    int64_t bound_samples;      /* 0 = auto */
    int32_t coef_slot;
    void *ws; size_t ws_bytes;  /* gem_knn_workspace_bytes(e, d, s, kp1) */
} gem_knn_prep_args;
int gem_knn_prep(const gem_knn_prep_args *args_host, void *stream);
typedef struct gem_knn_publish {
    const int64_t *remap;
    void *const *peer_base_host; int32_t world;
    size_t idx_offset_bytes, dist_offset_bytes;
} gem_knn_publish;
int gem_knn_scan(const float *mid, int64_t e, int64_t idx_offset, int d, const float *qmid, int64_t s, int kp1,
                 int64_t *out_idx, float *out_dist, void *ws, size_t ws_bytes, int coef_slot,
                 const gem_knn_publish *pub_host, void *stream);
/* Optional search radius per query for gem_knn_midpoints (tau_hint, may be NULL): only
 * neighbours with distance <= tau_hint[q] are required.  gem_knn_linegraph_hint computes a valid
 * one: the (k+1)-th smallest exact distance among the edges incident to the endpoints of the query
 * edge (row_ptr (n+1) int64 / col int32 = symmetric CSR of the graph).  In a force-directed layout
 * those are the true neighbours, and they are clustered in index space where a strided sample of
 * the candidates cannot see them.  With a hint a shard-local search (multi-GPU) may return fewer
 * than k+1 neighbours: short rows are padded with (distance +inf, index -1). */
int gem_knn_linegraph_hint(const float *pos, const int64_t *row_ptr, const int32_t *col, const int32_t *edges,
                           const int64_t *samp, int64_t s, int d, int kp1, float *tau_hint, void *stream);
/* Same contract, single exact streaming kernel (one CTA per query); the slow, simple
 * implementation used as in-library cross-check and as overflow fallback. */
int gem_knn_midpoints_exact(const float *mid, int64_t e, int64_t idx_offset, int d, const float *qmid,
                            int64_t s, int kp1, int mm_mode, int64_t *out_idx, float *out_dist,
                            void *stream);
/* Diagnostics: enable/disable the scan's event counters and report where they live inside the KNN
 * workspace (byte offsets): stats = 8 x uint64 {0: exact re-checks rejected by the key bound,
 * 1: accepted for insertion, 2: list inserts, 3: warp-level slow-path entries}; counts = uint32 per
 * query (survivors published); tau = fp32 per query.  The caller zeroes the stats words. */
int gem_knn_debug_stats(int enable, int64_t e, int d, int64_t s, int kp1, size_t *stats_offset,
                        size_t *counts_offset, size_t *tau_offset, int *cap, int *g);

/* Merge `parts` partial lists (parts, s, kp1) into (s, kp1) by (distance, index): the
 * exchange step of the edge-sharded multi-GPU KNN (after an all-gather of the partial lists). */
int gem_topk_merge(const float *dists, const int64_t *idxs, int parts, int64_t s, int kp1,
                   int64_t *out_idx, float *out_dist, void *stream);

/* Same, with the partial lists of part p at dists + p*dist_stride / idxs + p*idx_stride (elements):
 * lets one all-gather of a packed {dist | idx} block per rank feed the merge without repacking. */
int gem_topk_merge_strided(const float *dists, const int64_t *idxs, int64_t dist_stride, int64_t idx_stride,
                           int parts, int64_t s, int kp1, int64_t *out_idx, float *out_dist, void *stream);

/* gem_topk_merge_strided with the intersection stage fused into its tail (multi-GPU form of what the
 * select kernel does inside gem_layout_step): the CTA that merged query q's list evaluates its k
 * candidate pairs and adds the repulsion to the rows of `newpos` (fused spring+update form, row 0 =
 * vertex v_begin, only vertices in [v_begin, v_end)) with atomics that return the old row; the exact
 * change of the column sums is added to `sums` (2*ld doubles: sum | sum of squares).
 * pub (optional, multi-GPU publication by the last CTA of the launch): the rows of this rank that received a
 * repulsion term are stored again into every peer's raw buffer (row = padded vertex id; `newpos` must be row
 * v_begin of peer_raw_host[rank]), and the rank's corrected column sums (2*ld doubles at `sums`) go to slot `rank`
 * of the statistics area at stats_offset_bytes of every rank's exchange buffer.  touched: scratch of 4*s*(kp1-1)
 * int32; counters: 2 x uint32, zero-initialised once by the caller (left at zero by the kernel). */
typedef struct gem_merge_publish {
    float *const *peer_raw_host; void *const *peer_xchg_host;
    int32_t world, rank;
    size_t stats_offset_bytes;
    int32_t *touched; uint32_t *counters;
} gem_merge_publish;
int gem_topk_merge_intersect(const float *dists, const int64_t *idxs, int64_t dist_stride, int64_t idx_stride,
                             int parts, int64_t s, int kp1, int64_t *out_idx, float *out_dist, const float *pos,
                             const int32_t *edges, const int64_t *samp, int d, float k_inter, int64_t v_begin,
                             int64_t v_end, float *newpos, double *sums, const gem_merge_publish *pub_host,
                             void *stream);

/* (c) intersection repulsion.  Replaces _compute_intersection_forces (:638-736) and
 * _check_line_intersections (:738-774).  knn_full is the (s, kp1) list INCLUDING column 0,
 * which the kernel skips (:421).  Accumulates into `force` (n, ld) (not zeroed). */
int gem_intersection_forces(const float *pos, const int32_t *edges, int64_t n, int d,
                            const int64_t *samp, const int64_t *knn_full, int64_t s, int kp1,
                            float k_inter, float *force, void *stream);
/* Vertex-sliced variant for the multi-GPU path (every rank evaluates all <= s*k pairs from the
 * replicated positions but accumulates only into the vertex range [v_begin, v_end) it owns):
 * `force` points at row v_begin of the accumulator, i.e. it is a (v_end - v_begin, ld) slice. */
int gem_intersection_forces_range(const float *pos, const int32_t *edges, int64_t n, int d,
                                  const int64_t *samp, const int64_t *knn_full, int64_t s, int kp1,
                                  float k_inter, int64_t v_begin, int64_t v_end, float *force, void *stream);

/* (d) position update.  Replaces the tail of update_positions (:796-804):
 * new = pos + (f_spring + f_inter); new -= mean; new /= (unbiased std + 1e-6).
 * f_inter may be NULL.  stats_ws: 256-byte aligned scratch of gem_update_workspace_bytes(n,d),
 * zero-initialised once by the caller.
 * phase 0 = both passes; phase 1 = pass 1 only (writes unnormalised positions and the
 * column sums {sum, sum of squares} as 2*ld doubles at the start of stats_ws, for a
 * cross-rank all-reduce); phase 2 = pass 2 only (reads the reduced sums; n_total = global n);
 * phase 3 = column sums of `pos` only (nothing added or written; f_spring ignored) -- used by
 * gem_layout_step, whose spring kernel already wrote pos + F_spring. */
int gem_update_workspace_bytes(int64_t n, int d, size_t *bytes);
int gem_update_positions(float *pos, const float *f_spring, const float *f_inter, int64_t n,
                         int64_t n_total, int d, void *stats_ws, int phase, void *stream);

/* Multi-GPU: pass 2 fused with the exchange of the updated positions.  The rank normalises its n own
 * rows (read from `src`, row 0 = its first row) with the reduced column sums in stats_ws and stores
 * every result row at row_begin + i of EACH of the `world` position buffers peer_pos_host[r]
 * (host array of device pointers: the rank's own buffer and the peers' buffers mapped into this
 * process, e.g. torch.distributed._symmetric_memory buffer_ptrs).  The stores to the peers travel
 * over NVLink P2P inside this kernel; the caller separates iterations with a cross-rank barrier.
 * d in {2, 3}, world <= 16.
 * rank_sums (optional): world slots of 2*ld doubles, one per rank (its partial column sums, delivered by
 * gem_push_bytes); when given they are added in rank order inside the kernel and stats_ws is not read,
 * which replaces the all-reduce of the statistics and makes them bit-identical on every rank. */
int gem_update_normalise_push(float *const *peer_pos_host, int world, const float *src, int64_t row_begin,
                              int64_t n, int64_t n_total, int d, void *stats_ws, const double *rank_sums,
                              void *stream);
/* Multi-GPU, early-push flow: every rank normalises ALL n_rows rows of its replica -- raw (unnormalised new
 * positions: pushed by the owners' spring kernels, patched by their merge kernels) -> pos -- with the column sums
 * that arrived as `slots` per-rank blocks of 2*ld doubles (added in rank order: bit-identical on every rank). */
int gem_update_normalise_all(float *pos, const float *raw, int64_t n_rows, int64_t n_total, int d,
                             const double *rank_sums, int slots, void *stream);
/* idx[i] = table[idx[i]] for idx[i] >= 0 (the -1 padding of short lists is kept): a rank's KNN runs over
 * its own edges in the order its spring kernel produced their midpoints; with strided vertex ownership
 * that numbering is not the original one, and the partial lists are mapped back before the merge (which
 * breaks distance ties by ORIGINAL index). */
int gem_remap_indices(int64_t *idx, int64_t n, const int64_t *table, void *stream);
/* Copy nbytes (multiple of 16) from `src` to byte offset dst_offset of EVERY rank's exchange buffer
 * (peer_base_host[r]: peer-mapped base pointers): the all-gather of a small per-rank block (partial
 * top-(k+1) lists, column sums) as P2P stores. */
int gem_push_bytes(void *const *peer_base_host, int world, size_t dst_offset, const void *src, size_t nbytes,
                   void *stream);

/* Positions cross the public API as (n, d) row-major arrays in the caller's vertex numbering; on the device they are
 * (n_pad, ld) rows in padded numbering.  gem_rows_scatter: rows [row0, row0+cnt) of the public array (`src`: cnt x d,
 * already on the device) -> row pad_index[row0+i] (NULL: row0+i) of every buffer in peer_pos_host[0..world), pad
 * lanes zeroed -- the positions setter; on several GPUs each rank uploads 1/world of the rows over its own PCIe link
 * and this kernel fans them out over NVLink.  gem_rows_gather: the inverse for one buffer (the positions getter). */
int gem_rows_scatter(const float *src, int64_t row0, int64_t cnt, int d, const int64_t *pad_index,
                     float *const *peer_pos_host, int world, void *stream);
int gem_rows_gather(const float *pos, int64_t row0, int64_t cnt, int d, const int64_t *pad_index, float *out,
                    void *stream);

/* One whole iteration (update_positions, :776-806) on one GPU:
 *   side stream : KNN preparation (sample, query midpoints, line-graph + stratified bound, thresholds: one launch)
 *   `stream`    : spring forces + midpoints  ==join==>  KNN scan -> select + intersection forces -> update
 * The side stream and its two events belong to the library (created by gem_init, one set per
 * device); the fork/join is expressed with events, so the call stays CUDA-graph capturable.
 * Problems outside the KNN fast path run all stages in series on `stream`. */
typedef struct gem_plan {
    int64_t n, e, s;          /* vertices, edges, sample size (already min(sample_size, e)) */
    int32_t d, kp1;           /* n_components, n_neighbors + 1 */
    float k_attr, l_min, k_inter;
    uint64_t seed;
    float *pos;               /* (n, ld)   in/out */
    const int32_t *edges;     /* (e, 2) */
    const int64_t *row_ptr;   /* (n+1) symmetric CSR offsets, or NULL (no line-graph bound) */
    const int32_t *col;       /* (2e)  symmetric CSR columns, or NULL */
    const int64_t *up_ptr;    /* (n+1) #edges with first endpoint < v, or NULL (edge-parallel spring kernel) */
    const int32_t *hubs;      /* (n_hubs) vertices with degree > gem_hub_degree(), or NULL */
    int64_t n_hubs;
    float *force;             /* (n, ld)   scratch */
    float *mid;               /* (e, mld)  scratch */
    float *qmid;              /* (s, mld)  scratch */
    float *tau_hint;          /* (s)       scratch, or NULL */
    int64_t *samp;            /* (s)       out: the sample used (or in, when external_sample) */
    int64_t *knn_idx;         /* (s, kp1)  out */
    float *knn_dist;          /* (s, kp1)  out */
    int64_t *iter_counter;    /* (1)       device iteration counter */
    void *knn_ws; size_t knn_ws_bytes;
    void *stats_ws;
    int32_t external_sample;  /* 1: samp was filled by the caller (torch.randperm parity mode) */
    int32_t mm_mode;          /* -1 auto */
    int32_t coef_slot;        /* constant-bank coefficient slot owned by the caller (gem_coef_slot_acquire) */
} gem_plan;

int gem_layout_step(const gem_plan *plan_host, void *stream);

/* Profiling variant (bench.py roofline): the same kernels launched IN SERIES on `stream` (no side
 * stream, separate sample / query-midpoint / hint launches) with a CUDA event after every stage;
 * SYNCHRONISES the stream and writes the GEM_NUM_STAGES stage durations (ms) to ms_host.
 * Only valid when the KNN needs a single query batch (s <= 1024). */
#define GEM_STAGE_SAMPLE 0         /* fast path: empty (inside the fused preparation) */
#define GEM_STAGE_SPRING 1         /* spring/midpoint kernel (writes pos+F in the fused form) + column-sum pass */
#define GEM_STAGE_QUERY_MID 2      /* fast path: empty (inside the fused preparation) */
#define GEM_STAGE_KNN_BOUND 3      /* the fused KNN preparation kernel */
#define GEM_STAGE_KNN_THRESHOLD 4  /* coefficient copy to the constant bank */
#define GEM_STAGE_KNN_SCAN 5       /* the dominant kernel */
#define GEM_STAGE_KNN_SELECT 6     /* select kernel (with the fused intersection forces in gem_layout_step) */
#define GEM_STAGE_KNN_FALLBACK 7   /* exact kernel: the whole KNN for tiny / generic-d / k+1 > 64 inputs */
#define GEM_STAGE_INTERSECT 8      /* stand-alone intersection kernel (general path only) */
#define GEM_STAGE_UPDATE 9         /* normalisation pass (fused form) or both update passes */
#define GEM_NUM_STAGES 10
int gem_profile_step(const gem_plan *plan_host, void *stream, float *ms_host);

/* SURVEY 8(f).1, initial embedding.  Replaces the operator inside ARPACK's eigsh(L, k, which='SM') of
 * _compute_laplacian_embedding (embedder_pytorch.py:337-379): the eigenvectors of the normalised
 * Laplacian L = I - M with the smallest eigenvalues are those of M = D^-1/2 A D^-1/2 with the largest.
 *   y = alpha * (M x) + beta * z + gamma * x     x, y, z row-major (n, gem_spmv_cols() = 8) fp32, y != x, y != z,
 * z may be NULL (one pass = one step of the Chebyshev three-term recurrence), pull form over the symmetric CSR
 * (row_ptr (n+1) int64, col int32), dinv_sqrt[v] = deg(v)^-1/2 (0 for isolated vertices). */
int gem_spmv_cols(void);
int gem_spmv_normalized_adjacency(const int64_t *row_ptr, const int32_t *col, const float *dinv_sqrt, const float *x,
                                  float *y, int64_t n, float alpha, float beta, const float *z, float gamma,
                                  void *stream);

/* Single right-hand side form (the PageRank power iteration of the correlation harness): y = alpha * (M x), x, y (n) fp32. */
int gem_spmv_normalized_adjacency_vec(const int64_t *row_ptr, const int32_t *col, const float *dinv_sqrt, const float *x,
                                      float *y, int64_t n, float alpha, void *stream);

/* SURVEY 8(f).3, the caller graphem_seed_selection (graphem_rapids/influence.py:28-37):
 *   radial = np.linalg.norm(positions, axis=1); seeds = np.argsort(-radial)[:k]
 * on the device: out_idx[0..k) = the k vertex ids with the largest radius, ordered by (radius descending, id ascending),
 * out_radius (optional) their radii.  radius = sqrt of the left-to-right fp32 sum of squares (numpy's value bit for bit).
 * pos: (n_pad, ld) position buffer; pad_index (n) int64: vertex id -> row, or NULL (identity).  A 64-bit
 * most-significant-digit radix select over the keys (radius bits, ~id) recomputed from the positions in every pass:
 * 14 launches, no host synchronisation, nothing of size n materialised.  k <= gem_seed_select_max_k(), n < 2^32;
 * ws: gem_seed_select_workspace_bytes(k) bytes, 256-byte aligned. */
int gem_seed_select_max_k(void);
int gem_seed_select_workspace_bytes(int64_t k, size_t *bytes);
int gem_seed_select(const float *pos, int64_t n, int d, const int64_t *pad_index, int64_t k, int64_t *out_idx,
                    float *out_radius, void *ws, size_t ws_bytes, void *stream);

/* SURVEY 8(f).2, graph arrays on the device.  Replaces the host work of _validate_adjacency +
 * _extract_edges_from_adjacency (embedder_pytorch.py:182-245: adjacency.nonzero(), rows < cols, column_stack,
 * H2D of the int64 list) and the host build of the symmetric CSR of the pull kernels, for an adjacency whose CSR
 * is canonical (every row strictly ascending: sorted, no duplicates), has no stored zeros (caller's check) and a
 * symmetric pattern.  Then the reference's edge list is "the entries with col > row in storage order" and the
 * symmetric CSR is the adjacency minus its diagonal.
 *   gem_graph_count: row_ptr / up_ptr (n+1, int64) = offsets of the off-diagonal entries / of the entries above
 *     the diagonal per row (row_ptr[n] = 2E, up_ptr[n] = E); *flags (device int32) = 0 or a combination of
 *     GEM_GRAPH_* naming the precondition the input violates (the arrays are then not to be used);
 *     ws: gem_graph_workspace_bytes(n) bytes, 8-byte aligned.
 *   gem_graph_fill: col (2E int32) and edges (E x 2 int32, 8-byte aligned) from the offsets of gem_graph_count. */
#define GEM_GRAPH_NOT_SYMMETRIC 1  /* an off-diagonal entry (r, c) has no mirror entry (c, r) */
#define GEM_GRAPH_NOT_CANONICAL 2  /* a row is not strictly ascending, or a column index is out of range */
int gem_graph_workspace_bytes(int64_t n, size_t *bytes);
int gem_graph_count(const int64_t *indptr, const int32_t *indices, int64_t n, int64_t *row_ptr, int64_t *up_ptr,
                    int32_t *flags, void *ws, size_t ws_bytes, void *stream);
int gem_graph_fill(const int64_t *indptr, const int32_t *indices, int64_t n, const int64_t *row_ptr,
                   const int64_t *up_ptr, int32_t *col, int32_t *edges, void *stream);

/* Helpers behind the reference's private, unit-tested methods:
 * gem_pack_points: arbitrary (n,d) row-major points -> midpoint layout, for
 *   _compute_knn_chunked(query, reference, k) / _compute_knn_torch (:426-483, :543-593);
 * gem_check_line_intersections: _check_line_intersections(p1,p2,q1,q2) (:738-774) on (p,d)
 *   row-major inputs, out[i] in {0,1}. */
int gem_pack_points(const float *pts, int64_t n, int d, float *out, void *stream);
int gem_check_line_intersections(const float *p1, const float *p2, const float *q1, const float *q2, int64_t p, int d,
                                 uint8_t *out, void *stream);

/* Measured FP32 FMA throughput of the device (dependent-chain-free FFMA loop), for the
 * roofline denominator of the KNN kernel: writes flop/s of scalar FFMA to *flops_host and of the
 * packed FFMA2 (fma.rn.f32x2) form to *flops2_host (may be NULL).  Synchronises. */
int gem_fp32_peak_probe(double *flops_host, double *flops2_host, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GRAPHEM_B200_H */
