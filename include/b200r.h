/* b200r.h -- C ABI of the B200-native retrieval hot path (libb200r.so).
 *
 * The reference (AndyUkJ/A-Modular-RAG-Framework) is pure Python; nothing in it is C.  Each entry
 * point below therefore cites the PYTHON interface it replaces (paths relative to the reference
 * root) -- this is what a maintainer binds with ctypes/cffi from
 * app/modules/retrieval (see INTEGRATION.md for the stub).
 *
 * Conventions
 *   - every function returns 0 on success, a negative B200R_E* code on failure;
 *     b200r_last_error() returns the message of the calling thread's last failure;
 *   - all pointers are DEVICE pointers unless the name ends in _host;
 *   - the caller owns every buffer it passes in; the library borrows corpus / graph pointers for
 *     the lifetime of the handle and allocates only its own small derived state (row norms, tensor maps);
 *   - `stream` is a cudaStream_t passed as void*; no call synchronises the device except the
 *     *_host convenience calls and b200r_index_create (which returns with the norms computed);
 *   - handles are not thread-safe; distinct handles are independent;
 *   - there is NO CPU fallback: without a CUDA device every call fails with B200R_ECUDA.
 */
#ifndef B200R_H_
#define B200R_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200R_VERSION 100 /* 0.1.0 */

/* error codes */
#define B200R_OK 0
#define B200R_EINVAL -1   /* bad argument (message says which) */
#define B200R_ECUDA -2    /* CUDA runtime/driver error */
#define B200R_ENOMEM -3   /* workspace too small / allocation failed */
#define B200R_EOVERFLOW -4 /* a bounded on-chip structure overflowed (graph frontier) */
#define B200R_ESTATE -5   /* bad handle / call out of order */
#define B200R_ETIMEOUT -6 /* a device-side exchange wait expired (a rank never arrived) */

/* element types of the corpus matrix (queries use the same type, except I8 -> F32 queries) */
#define B200R_BF16 0
#define B200R_F16 1
#define B200R_F32 2
#define B200R_I8 3

/* metric.  COS == DenseReranker._cosine (retrieval_backend.py:192-197): dot/(|q||x|), 0.0 when a norm is 0 */
#define B200R_IP 0
#define B200R_COS 1

/* search path */
#define B200R_PATH_AUTO 0
#define B200R_PATH_GEMM 1 /* tcgen05/TMEM/TMA tiles with the top-k fused into the epilogue (bf16/f16) */
#define B200R_PATH_SCAN 2 /* CUDA-core streaming scan, 128-bit loads (all types; meant for B <= 4) */

#define B200R_MAX_K 2048
#define B200R_MAX_PEERS 8
#define B200R_COMM_BLOB_BYTES 128

int b200r_version(void);
/* copies the calling thread's last error message (NUL terminated) into buf; returns its length */
int b200r_last_error(char* buf, size_t n);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
int64_t b200r_launch_count(void);

/* ---- corpus index ------------------------------------------------------------------------
 * Replaces the per-query embedding fetch + candidate list of DenseReranker.score
 * (retrieval_backend.py:215-247): X[N, D] row-major is the passage-embedding matrix, row n being
 * line n of docs.jsonl (BM25LiteIndex._build, text_index.py:35-53).  Reported ids are id_base + n.
 * row_scale (I8 only, may be NULL = 1.0): x[n, d] = row_scale[n] * X[n, d].
 */
int b200r_index_create(int device, const void* X, int64_t N, int32_t D, int32_t dtype, int32_t metric,
                       int64_t id_base, const float* row_scale, void* stream, uint64_t* out_index);
int b200r_index_destroy(uint64_t index);
int b200r_index_info(uint64_t index, int64_t* N, int32_t* D, int32_t* dtype, int32_t* metric, int64_t* id_base);

/* ---- exact top-k search --------------------------------------------------------------------
 * Replaces `{id: _cosine(qv, v)}` + `sort(key=score, reverse=True)[:k]`
 * (retrieval_backend.py:245, 371-372; retrieval_adapter.py:129-131) for a batch of B queries over
 * the WHOLE corpus.  Q[B, D] row-major in the corpus dtype (F32 for an I8 corpus).
 * out_scores[B, k] f32, out_ids[B, k] i64, ordered by (score desc, id asc); when N < k the tail is
 * (-inf, -1).  workspace: device scratch of at least b200r_search_workspace_bytes().
 */
size_t b200r_search_workspace_bytes(uint64_t index, int32_t B, int32_t k, int32_t path);
int b200r_search(uint64_t index, const void* Q, int32_t B, int32_t k, int32_t path, float* out_scores,
                 int64_t* out_ids, void* workspace, size_t workspace_bytes, void* stream);
/* Same call with HOST query / result buffers (what the Python plug-in hands over): copies Q to the
 * device, searches, copies the results back and synchronises `stream`.  Uses library-owned staging
 * buffers that grow on demand. */
int b200r_search_host(uint64_t index, const void* Q_host, int32_t B, int32_t k, int32_t path,
                      float* out_scores_host, int64_t* out_ids_host, void* stream);
/* The same with the queries as FP32, whatever the corpus dtype -- what an embedding provider returns
 * (router.embed yields Python floats, retrieval_backend.py:227-231).  The fp32 rows are copied to the device and
 * rounded there to the index's query dtype (round to nearest even, the conversion torch / numpy apply), so the host
 * never touches the 2-byte types. */
int b200r_search_host_f32(uint64_t index, const float* Q_host_f32, int32_t B, int32_t k, int32_t path,
                          float* out_scores_host, int64_t* out_ids_host, void* stream);

/* fp32 -> bf16 / f16 on the device (round to nearest even): how fp32 query embeddings become the query dtype of a
 * bf16 / f16 index without a host-side conversion.  src / dst are device pointers on `device`; dtype = B200R_BF16 / B200R_F16. */
int b200r_convert_f32(int device, const float* src, void* dst, int64_t n, int32_t dtype, void* stream);

/* Similarity floor: only results with score >= min_score are reported (at most the k best per query, the rest
 * of the row is (-inf, -1)).  With Q = rows of the corpus itself this is the all-pairs `sim >= semantic_threshold`
 * test of EdgeBuilder.build step 4 (app/modules/graph_construction/edge_builder.py:146-169) at corpus scale:
 * the floor seeds every query's filter threshold, so the tile kernels reject almost everything from the start. */
int b200r_search_min_score(uint64_t index, const void* Q, int32_t B, int32_t k, int32_t path, float min_score,
                           float* out_scores, int64_t* out_ids, void* workspace, size_t workspace_bytes,
                           void* stream);

/* ---- range search: EVERY row at or above a similarity threshold --------------------------------------------------
 * The all-pairs `sim >= semantic_threshold` test of EdgeBuilder.build step 4 (app/modules/graph_construction/
 * edge_builder.py:146-169) at corpus scale, without any per-row cap: the tile kernel's epilogue tests the FINAL score
 * (dot * row multiplier * 1/|q|, 0.0 for zero norms like _cosine) against min_score instead of keeping a top-k.
 * Two passes over the corpus: b200r_range_count counts the hits of every (row chunk, query), turns them into offsets
 * (out_offsets[B+1] on the device, exclusive; *total_host = out_offsets[B]; synchronises `stream`) and leaves the
 * per-chunk offsets in `workspace`; the caller allocates exactly `total` ids / scores and b200r_range_fill (same Q,
 * min_score, id_gt and workspace) writes query q's hits to [out_offsets[q], out_offsets[q+1]) in ASCENDING row id.
 * id_gt (nullable, i64 [B]): only rows whose global id is > id_gt[q] -- with Q = rows of the corpus and id_gt[q] = q's own
 * id every unordered pair is reported once (itertools.combinations order, edge_builder.py:153) and never a self match.
 * bf16 / f16 corpora with D % 8 == 0 (the tensor path). */
size_t b200r_range_workspace_bytes(uint64_t index, int32_t B);
int b200r_range_count(uint64_t index, const void* Q, int32_t B, float min_score, const int64_t* id_gt, int64_t* out_offsets,
                      int64_t* total_host, void* workspace, size_t workspace_bytes, void* stream);
int b200r_range_fill(uint64_t index, const void* Q, int32_t B, float min_score, const int64_t* id_gt, int64_t total,
                     int64_t* out_ids, float* out_scores, void* workspace, size_t workspace_bytes, void* stream);

/* ---- shard exchange (row-sharded corpus) ------------------------------------------------------
 * The reference keeps one index in one process (BM25LiteIndex / DenseReranker, retrieval_backend.py:252-281); at
 * 8 GPUs the corpus is row-sharded and the per-shard top-k lists must meet.  The exchange lives HERE, under the C
 * ABI, and uses neither NCCL nor torch.distributed: every rank owns one window of device memory that all ranks map
 * (CUDA IPC between processes, peer access inside one process) and the search kernels store into the peers' windows
 * over NVLink themselves, ordered by device-side epoch flags (csrc/comm.cuh).
 *
 * Bring-up, one process per GPU:   b200r_comm_create on every rank -> b200r_comm_export (a 128-byte descriptor) ->
 *   the host swaps the descriptors through ANY channel it has (file, socket, MPI, torch.distributed object gather) ->
 *   b200r_comm_connect(all descriptors, rank order).
 * Bring-up, one process, G GPUs:   b200r_comm_init_all(devices, G, ...) does all of it.
 * max_B / max_k / max_graph_k bound the (B, k) of later exchanges (they size the window: ~24*max_B*max_k*world bytes
 * for the dense channel, ~56*max_B*max_graph_k*world for the graph stage's; max_graph_k = 0 leaves those out).
 * Every rank must issue the same sequence of exchange calls with the same (B, k).  Waits on the device are bounded
 * (20 s): b200r_comm_status returns B200R_ETIMEOUT afterwards instead of the GPU hanging. */
int b200r_comm_create(int device, int32_t rank, int32_t world, int32_t max_B, int32_t max_k, int32_t max_graph_k,
                      uint64_t* out_comm);
int b200r_comm_export(uint64_t comm, void* blob_host /* B200R_COMM_BLOB_BYTES */);
int b200r_comm_connect(uint64_t comm, const void* blobs_host /* world x B200R_COMM_BLOB_BYTES, rank order */);
int b200r_comm_init_all(const int32_t* devices, int32_t G, int32_t max_B, int32_t max_k, int32_t max_graph_k,
                        uint64_t* out_comms /* [G] */);
int b200r_comm_destroy(uint64_t comm);
int b200r_comm_status(uint64_t comm, void* stream);   /* synchronises `stream`; B200R_OK or B200R_ETIMEOUT */
int b200r_comm_info(uint64_t comm, int32_t* rank, int32_t* world, int32_t* device, size_t* window_bytes);

/* Row-sharded exact top-k: `index` holds this rank's rows (id_base = first global row), every rank passes the SAME
 * Q[B, D] and gets the same global result -- DenseReranker.score + sort/slice (retrieval_backend.py:215-247, 371-372)
 * over the union of the shards.  _begin = local search whose final select kernel stores this rank's packed record
 * [ids i64 [B,k] | scores f32 [B,k]] into every rank's window (the all-gather IS those stores); _finish = one kernel that
 * waits for all ranks' records and merges the G sorted lists by (score desc, global id asc).  Splitting the two lets a
 * throughput loop issue begin(i+1) before finish(i), so no rank ever waits for a slower peer.  Stream-ordered, no host
 * synchronisation; safe to capture in a CUDA graph (epochs live on the device). */
int b200r_search_sharded_begin(uint64_t index, uint64_t comm, const void* Q, int32_t B, int32_t k, int32_t path,
                               void* workspace, size_t workspace_bytes, void* stream);
int b200r_search_sharded_finish(uint64_t comm, int32_t B, int32_t k, float* out_scores, int64_t* out_ids, void* stream);
int b200r_search_sharded(uint64_t index, uint64_t comm, const void* Q, int32_t B, int32_t k, int32_t path,
                         float* out_scores, int64_t* out_ids, void* workspace, size_t workspace_bytes, void* stream);
/* One process driving G GPUs (comms from b200r_comm_init_all, indexes[g] on comms[g]'s device): HOST queries in, HOST
 * results out -- H2D to every GPU, sharded search, exchange, merge on GPU 0, D2H, all streams synchronised.  This is the
 * call behind B200RetrievalBackend(devices=[...]).  streams may be NULL (default streams). */
int b200r_search_sharded_host_all(const uint64_t* indexes, const uint64_t* comms, int32_t G, const void* Q_host, int32_t B,
                                  int32_t k, int32_t path, float* out_scores_host, int64_t* out_ids_host,
                                  void* const* streams);
/* ... with FP32 host queries, converted to the indexes' query dtype on every device (see b200r_search_host_f32). */
int b200r_search_sharded_host_all_f32(const uint64_t* indexes, const uint64_t* comms, int32_t G, const float* Q_host_f32,
                                      int32_t B, int32_t k, int32_t path, float* out_scores_host, int64_t* out_ids_host,
                                      void* const* streams);

/* ---- MMR diversification (optional post-filter) ----------------------------------------------------------------
 * Replaces mmr_diversify (app/utils/similarity.py:33-62) for B candidate lists at once: cand_ids / cand_scores [B, P]
 * (global ids, < 0 = padding; list order matters: ties go to the earlier item), the item vectors are the corpus rows
 * behind the ids (an id whose row is not in this index has no vector: similarity 0, like a None vector).  Greedy:
 * first pick = best score, then argmax of lambda*score - (1-lambda)*max(0, max cos to the picked ones).
 * out_ids / out_scores / out_pos [B, top_k]: the picks in order, their original scores and positions ((-1, -inf, -1)
 * once a list is exhausted). */
int b200r_mmr_diversify(uint64_t index, const int64_t* cand_ids, const float* cand_scores, int32_t B, int32_t P, int32_t top_k,
                        float lambda_weight, int64_t* out_ids, float* out_scores, int32_t* out_pos, void* stream);

/* ---- merge of per-shard top-k lists ----------------------------------------------------------
 * The same stable-sort rule applied to the union of G lists (one per GPU, however the host gathered them -- the
 * library's own exchange, b200r_search_sharded*, merges without this call):
 * cand_scores/cand_ids [G, B, kin] -> out [B, k], (score desc, id asc); ids < 0 are padding.
 */
int b200r_merge_topk(int device, const float* cand_scores, const int64_t* cand_ids, int32_t G, int32_t B,
                     int32_t kin, int32_t k, float* out_scores, int64_t* out_ids, void* stream);

/* Same merge with explicit shard strides (in elements) and the provenance of every winner: out_src[B, k] =
 * g*kin + j, the winner's position in shard g's list (-1 for padding), so that a caller that gathered more
 * per-candidate fields than (score, id) -- the graph stage's hop / normalised channel scores / cosine
 * (retrieval_backend.py:365-369 meta keys) -- can pick them without a search. */
int b200r_merge_topk_src(int device, const float* cand_scores, size_t shard_stride_scores, const int64_t* cand_ids,
                         size_t shard_stride_ids, int32_t G, int32_t B, int32_t kin, int32_t k, float* out_scores,
                         int64_t* out_ids, int32_t* out_src, void* stream);

/* Same merge over the layout the all-gather of one packed buffer per rank produces:
 * packed = G records, each [ids i64 [B, kin] | scores f32 [B, kin]] (12*B*kin bytes). */
int b200r_merge_topk_packed(int device, const void* packed, int32_t G, int32_t B, int32_t kin, int32_t k,
                            float* out_scores, int64_t* out_ids, void* stream);

/* ---- kernel timing (bench.py's roofline leg) --------------------------------------------------
 * When enabled, the library brackets every launch of its DOMINANT search kernel (gemm_topk or
 * scan_topk) with CUDA events on the caller's stream.  b200r_profile_read synchronises those events and
 * returns the summed device time in ms and the number of bracketed launches (and clears them). */
int b200r_profile_enable(int on);
int b200r_profile_read(double* kernel_ms, int64_t* launches);
/* b200r_profile_enable(2): EVERY launch site of the library is bracketed by CUDA events under its name; b200r_profile_dump
 * synchronises them, writes {"site": [total_ms, launches], ...} (JSON, NUL terminated, truncated to n) into buf, clears the
 * records and returns the untruncated length.  The timeline behind profiles/*_timeline.json. */
int b200r_profile_dump(char* buf, size_t n);

/* ---- passage/entity graph --------------------------------------------------------------------
 * Replaces load_graph_json + build_index (graph_utils.py:13-55), done once instead of per query:
 * CSR of the UNDIRECTED next_in_doc adjacency (neighbours(s) = next_forward[s] + next_backward[s],
 * graph_utils.py:123).  row_ptr[V+1] i64, col_idx[E] i32, node id == corpus row id.
 */
int b200r_graph_create(int device, const int64_t* row_ptr, const int32_t* col_idx, int64_t V, int64_t E,
                       uint64_t* out_graph);
/* Graph whose node ids are LOCAL (0..V-1) with node_rows[V] i64 mapping every node to its global corpus row (< 0 = the
 * node has no embedding row and scores 0.0): the shape of the reference's per-question graphs
 * (data/graph/<graph_id>/graph.json, load_graph_json graph_utils.py:13-22), and of a block-diagonal union of many of
 * them so that a batch of requests with DIFFERENT graph_ids expands in one launch.  Seeds passed to expand_rescore are
 * local node ids; reported ids are corpus rows.  Both creators validate the CSR on the device (row_ptr monotone from
 * 0 to E, 0 <= col_idx < V) and fail with B200R_EINVAL on a malformed one. */
int b200r_graph_create_mapped(int device, const int64_t* row_ptr, const int32_t* col_idx, int64_t V, int64_t E,
                              const int64_t* node_rows, uint64_t* out_graph);
int b200r_graph_destroy(uint64_t graph);

/* Replaces expand_qmatch_neighbors (graph_utils.py:58-129) + _cosine re-score + _minmax_norm and the
 * weighted fusion / sort / slice of HybridRetrievalBackend.run (retrieval_backend.py:296-301, 353-372)
 * for B queries at once.  seeds[B, S] i64 are q_match targets (< 0 = padding); BFS depth <= window;
 * graph score = decay(hop) (graph_utils.py:87-94); fused = alpha_graph*minmax(graph) + alpha_dense*minmax(cos).
 * Outputs [B, k]: fused score, id, hop distance, the two normalised channel scores (the
 * score_graph_norm / score_dense_norm meta keys, retrieval_backend.py:368) and the raw cosine; tail
 * entries beyond the candidate count are (-inf, -1, -1, 0, 0, 0).  out_ncand[B] = candidates found.
 * Fails with B200R_EOVERFLOW if a query reaches more than b200r_expand_max_candidates() nodes.
 * Three kernels run on `stream`: BFS (candidate lists parked in `workspace`), a grid-wide balanced gather of the
 * candidate rows (the HBM-bound part), and the per-query fusion + top-k.  `workspace`: caller-owned device scratch
 * of at least b200r_expand_workspace_bytes(B) bytes, 256-byte aligned, private to this call until it completes.
 */
int32_t b200r_expand_max_candidates(void);
size_t b200r_expand_workspace_bytes(int32_t B);
int b200r_expand_rescore(uint64_t index, uint64_t graph, const void* Q, const int64_t* seeds, int32_t B,
                         int32_t S, int32_t window, float alpha_dense, float alpha_graph, int32_t k,
                         void* workspace, size_t workspace_bytes,
                         float* out_scores, int64_t* out_ids, int32_t* out_hops, float* out_graph_norm,
                         float* out_dense_norm, float* out_cos, int32_t* out_ncand, int32_t* out_status,
                         void* stream);

/* Unbounded frontier: ONE query whose expansion exceeds b200r_expand_max_candidates() (out_status == B200R_EOVERFLOW
 * above) redone with everything in global memory -- a visited bitmap over the graph's nodes, one grid-wide kernel per
 * BFS level, the same gather / fusion / top-k -- so that, like the reference's BFS (graph_utils.py:109-127), no
 * neighbourhood is too large.  Q: that query's vector, seeds [S]; outputs [k] (+ out_ncand [1]), same meaning as above.
 * workspace: b200r_expand_big_workspace_bytes(graph) bytes (about 21 bytes per graph node), 256-byte aligned. */
size_t b200r_expand_big_workspace_bytes(uint64_t graph);
int b200r_expand_rescore_big(uint64_t index, uint64_t graph, const void* Q, const int64_t* seeds, int32_t S, int32_t window,
                             float alpha_dense, float alpha_graph, int32_t k, void* workspace, size_t workspace_bytes, float* out_scores,
                             int64_t* out_ids, int32_t* out_hops, float* out_graph_norm, float* out_dense_norm, float* out_cos,
                             int32_t* out_ncand, void* stream);

/* Row-sharded corpora (one rank per GPU, CSR replicated, embeddings sharded by rows; SURVEY 8e "owner
 * computes").  pass 1: BFS + re-score of the candidates whose global id lies in [own_lo, own_hi), parked in
 * `workspace` (b200r_expand_workspace_bytes(B)), and minmax[B][2] = this rank's (min, max) cosine.  The caller
 * all-reduces minmax (MIN on [:,0], MAX on [:,1]) across ranks, then pass 2 fuses with the global range and
 * returns this rank's top-k among the candidates it owns; the per-rank lists are merged with
 * b200r_merge_topk.  Seeds and reported ids are GLOBAL row ids (the index's id_base maps them to local rows). */
int b200r_expand_rescore_sharded(uint64_t index, uint64_t graph, const void* Q, const int64_t* seeds, int32_t B,
                                 int32_t S, int32_t window, float alpha_dense, float alpha_graph, int32_t k,
                                 int32_t pass, int64_t own_lo, int64_t own_hi, void* workspace,
                                 size_t workspace_bytes, float* minmax, float* out_scores, int64_t* out_ids,
                                 int32_t* out_hops, float* out_graph_norm, float* out_dense_norm, float* out_cos,
                                 int32_t* out_ncand, int32_t* out_status, void* stream);

/* The same stage with the QUERIES sharded over the ranks as well, entirely over the exchange (no NCCL collective, no
 * host reduction): every rank passes the same Q / seeds and receives the same [B, k] results.
 *   phase 1  the HOME rank of a query (q / ceil(B/world)) runs its BFS and stores, for every rank, the candidates whose
 *            corpus rows that rank OWNS (row / ceil(n_total_rows/world)) into that rank's window;
 *   phase 2  every rank gathers + re-scores the candidates it received (the HBM-bound part, now 1/world of the work per
 *            rank) and stores each cosine into the home rank's window, at the candidate's position in the home's list;
 *   phase 3  the home rank min-max-normalises, fuses and ranks its queries (it now holds every cosine, so the
 *            reference's per-query min/max, retrieval_backend.py:296-301, needs no reduction) and stores the results
 *            into every rank's window;
 *   phase 4  every rank copies the results into the caller's arrays.
 * phase = 0 runs all four (one process per GPU); a host thread that drives several ranks issues phase p on every rank
 * before phase p+1 on any.  The comm must have been created with max_graph_k >= k and max_B >= B.  Node ids reported
 * are corpus rows; `workspace` as for b200r_expand_rescore. */
int b200r_expand_rescore_exchange(uint64_t index, uint64_t graph, uint64_t comm, const void* Q, const int64_t* seeds, int32_t B,
                                  int32_t S, int32_t window, float alpha_dense, float alpha_graph, int32_t k, int64_t n_total_rows,
                                  int32_t phase, void* workspace, size_t workspace_bytes, float* out_scores, int64_t* out_ids,
                                  int32_t* out_hops, float* out_graph_norm, float* out_dense_norm, float* out_cos, int32_t* out_ncand,
                                  int32_t* out_status, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200R_H_ */
