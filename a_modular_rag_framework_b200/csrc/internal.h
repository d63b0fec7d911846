// internal.h -- host-side declarations shared by the translation units of libb200r.so
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string>

#include "../../include/b200r.h"
#include "comm.cuh"

namespace b200r {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define B200R_CUDA_CHECK(expr)                                                            \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::b200r::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return B200R_ECUDA;                                                                 \
    }                                                                                     \
  } while (0)

struct Index {
  uint32_t magic;
  int device;
  const void* X;
  int64_t N;
  int32_t D;
  int32_t dtype;
  int32_t metric;
  int64_t id_base;
  float* row_mult;      // [N] or nullptr: cos -> 1/|x_n| (0 when |x_n| == 0); ip+i8 -> row_scale; ip -> nullptr
  bool has_tmap;        // X tensor map valid (bf16/f16, D % 8 == 0)
  CUtensorMap tmap_x;        // box 64 x 256 rows (one CTA per tile)
  CUtensorMap tmap_x_half;   // box 64 x 128 rows (CTA pair, each CTA loads half of the tile's rows)
  bool has_tmap_i8;          // int8 corpus: tensor map for the kind::i8 rows-mode kernel
  CUtensorMap tmap_x_i8;     // box 128 bytes x 128 rows
  bool has_tmap_f32;         // fp32 corpus: tensor map for the kind::tf32 split-precision rows-mode kernel (rows_f32.cu)
  CUtensorMap tmap_x_f32;    // box 32 floats x 128 rows
  // staging for b200r_search_host
  void* h_q_dev; size_t q_dev_bytes;
  void* h_q32_dev; size_t q32_dev_bytes;   // fp32 host queries before their on-device conversion (b200r_search_host_f32)
  float* h_s_dev; int64_t* h_i_dev; size_t out_dev_elems;
  void* h_ws; size_t ws_bytes;
};

struct Graph {
  uint32_t magic;
  int device;
  const int64_t* row_ptr;
  const int32_t* col_idx;
  const int64_t* node_rows;   // nullable: node -> global corpus row (b200r_graph_create_mapped)
  int64_t V, E;
};
int validate_csr(const long long* row_ptr, const int* col_idx, long long V, long long E);

// SMs of the CURRENT device (148 on B200), queried once per device: grids and the plan are sized from it
int num_sms();
constexpr uint32_t kIndexMagic = 0xB200A001u;
constexpr uint32_t kGraphMagic = 0xB200A002u;

inline size_t elem_size(int dtype) { return dtype == B200R_F32 ? 4 : (dtype == B200R_I8 ? 1 : 2); }
inline int list_cap_soft(int k) { return k <= 256 ? 512 : 2 * k; }
inline int list_cap(int k) { return list_cap_soft(k) + 32; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
// keys one select CTA stages in shared memory when it reduces a query's lists (also the capacity of the carry list)
inline int select_stage_limit(int k, int B) {
  if (B <= 148) return 24576;             // a handful of queries: one CTA per SM can afford a 192 KB stage
  int limit = 3 * k + 2048;
  if (limit < 2560) limit = 2560;
  return limit > 24576 ? 24576 : limit;
}

constexpr int kMaxPhases = 16;
struct GemmPhase { int tile_begin, tile_end, num_chunks, tiles_per_chunk; };

// ---- plan of one search: how the corpus is cut into units and where the partial lists live -------
struct SearchPlan {
  int path;             // B200R_PATH_GEMM / B200R_PATH_SCAN
  int B, k, cap, cap_soft;
  int Bpad;             // list slots per segment (>= B)
  int num_segments;     // lists per query
  // gemm: row tiles are processed in phases [tile_begin, tile_end) of geometrically growing size; after each
  // phase the exact k-th best so far becomes every query's filter threshold for the next one
  int cta_group;         // 1 or 2 (CTA pair, cta_group::2)
  int rows_i8;           // > 0: int8 tensor path (rows_i8.cu) with this many groups of rows_i8_nq queries (ganged in one launch); 128-row tiles, 4 lists per unit
  int rows_i8_nq;        // 16 / 32 / 64
  int rows_f32;          // 1: the groups above are served by the fp32 split-precision kernel (rows_f32.cu, 16 queries each)
  int q_replicas;        // 1: single query tile replicated in the workspace (off_qrep) against L2 hot-spotting
  int num_q_tiles, num_row_tiles, num_chunks;   // num_q_tiles counts tiles of cta_group*128 queries
  int num_phases;
  GemmPhase phases[kMaxPhases];
  // scan
  int scan_bq, scan_warps, scan_blocks;
  int scan_bulk;         // 1: bulk-copy-fed scan (scan_bulk.cu), else the register-fed one (scan_topk.cu)
  // workspace layout (byte offsets)
  int num_groups;        // > 0: two-level final select (lists2/counts2 hold the first level's output)
  int carry_cap;         // > 0: phased search with one carry list per query (select_merge.cu), this many keys each
  size_t off_lists, off_counts, off_gthr, off_qinv, off_lists2, off_counts2, off_qrep, off_carry, off_end;
};

int make_plan(const Index* ix, int B, int k, int path, SearchPlan* plan);

// kernels (each returns a B200R_* code and records the launch)
int launch_row_mult(const Index* ix, const float* row_scale, cudaStream_t st);
int launch_query_inv_norm(const void* Q, int B, int D, int qdtype, float* qinv, cudaStream_t st);
int launch_init_search(const SearchPlan& p, void* ws, cudaStream_t st);
int launch_min_score(const SearchPlan& p, void* ws, bool cosine, float min_score, cudaStream_t st);
int encode_query_tmap(const Index* ix, const void* Q, int B, CUtensorMap* tmap_q);
// range mode of the tile kernel (gemm_topk.cu): mode 1 = count per (chunk, query) into the plan's counts, 2 = fill
struct RangeArgs { int mode; float tau; const float* qinv; const int64_t* id_gt; const int64_t* offsets; void* out; };
int launch_gemm_topk(const Index* ix, const CUtensorMap* tmap_q, const SearchPlan& p, const GemmPhase& ph, void* ws, cudaStream_t st,
                     const RangeArgs* rg = nullptr);
int launch_range_offsets(const SearchPlan& pl, void* ws, int64_t* offsets, int64_t* slot_off, cudaStream_t st);
int launch_range_unpack(const void* keys, int64_t total, const int64_t* offsets, int B, float unused, int64_t id_base, int64_t* out_ids,
                        float* out_scores, cudaStream_t st);
size_t query_replica_bytes(int D);
int launch_replicate_queries(const void* Q, void* rep, int B, int D, cudaStream_t st);
int launch_search_prologue(const SearchPlan& p, void* ws, const void* Q, int D, int qdtype, bool cosine, cudaStream_t st);   // K1: init + 1/|q| + query replicas in one launch
int launch_convert_queries(const float* src, void* dst, size_t n, int dtype, cudaStream_t st);   // f32 -> bf16 / f16, RNE
int encode_query_tmap_replicated(const Index* ix, const void* rep, CUtensorMap* tmap_q);
int launch_scan_topk(const Index* ix, const void* Q, const SearchPlan& p, void* ws, cudaStream_t st);
int rows_i8_nq(int B, int D);                 // queries per corpus pass (16 / 32 / 64), 0 = int8 tensor path not available
int rows_i8_tile_rows();
size_t rows_i8_query_bytes(int B, int nq, int D);
int launch_quantize_queries_i8(const void* Q, int B, int nq, int D, void* qws, cudaStream_t st);
int encode_query_tmap_i8(const void* qws, int groups, int nq, int D, CUtensorMap* tmap_q);   // one map over every group's limb rows
int rows_i8_ctas_per_group(int groups);      // CTAs per query group when `groups` groups are ganged in one launch
int launch_rows_i8(const Index* ix, const CUtensorMap* tmap_q, const void* qws, const SearchPlan& p, const GemmPhase& ph, void* ws,
                   cudaStream_t st);        // one phase of ALL the plan's query groups
int rows_f32_nq(int D);                       // queries per corpus pass of the fp32 tensor path (16), 0 = not available
int rows_f32_tile_rows();
size_t rows_f32_query_bytes(int B, int D);
int launch_split_queries_tf32(const void* Q, int B, int D, void* qws, cudaStream_t st);
int encode_query_tmap_f32(const void* qws, int group, int D, CUtensorMap* tmap_q);
int launch_rows_f32(const Index* ix, const CUtensorMap* tmap_q, int group, const SearchPlan& p, const GemmPhase& ph, void* ws, cudaStream_t st);
int launch_mmr(const Index* ix, const int64_t* ids, const float* scores, int B, int P, int top_k, float lambda, int64_t* out_ids,
               float* out_scores, int32_t* out_pos, cudaStream_t st);
int scan_bulk_warps(const Index* ix);        // consumer warps per CTA of the bulk-fed scan, 0 if it does not apply
int scan_bulk_tile_rows(const Index* ix);
int launch_scan_bulk(const Index* ix, const void* Q, const SearchPlan& p, void* ws, cudaStream_t st);
// out_scores == nullptr: threshold mode (publish the exact k-th best so far into gthr); else final output
// comm != nullptr: the final select stores this rank's packed record into every rank's exchange window (comm.cuh)
int launch_select_lists(const Index* ix, const SearchPlan& p, void* ws, float* out_scores, int64_t* out_ids, cudaStream_t st,
                        const CommDev* comm = nullptr, size_t rec_stride = 0, float min_score = -__builtin_huge_valf());

// ---- shard exchange (comm.cu) ----------------------------------------------------------------------------------
constexpr uint32_t kCommMagic = 0xB200A003u;
constexpr int kChanDense = 0, kChanGraphReq = 1, kChanGraphCos = 2, kChanGraphOut = 3;
struct Comm {
  uint32_t magic;
  int device, rank, world;
  int max_B, max_k, max_graph_k;
  void* window;
  size_t window_bytes;
  unsigned long long peer[B200R_MAX_PEERS];     // every rank's window as mapped in this process
  bool ipc_opened[B200R_MAX_PEERS];
  bool connected;
  size_t data_off[kCommChannels], rec_cap[kCommChannels];
};
Comm* get_comm(uint64_t h);
// device-side view of one channel; fails if rec_stride exceeds the channel's record capacity
int comm_dev(const Comm* c, int channel, size_t rec_stride, CommDev* out);
inline size_t dense_record_stride(int B, int k) { return align_up((size_t)12 * B * k, 16); }
int launch_exchange_merge(const Comm* c, int B, int kin, int k, float* out_scores, int64_t* out_ids, cudaStream_t st);
int launch_exchange_skip(const Comm* c, cudaStream_t st);
int launch_scatter_record(const CommDev& cd, const void* src, size_t bytes, size_t rec_stride, cudaStream_t st);
int launch_merge_topk(const float* cs, const int64_t* ci, size_t shard_stride_s, size_t shard_stride_i, int G, int B, int kin, int k,
                      float* os, int64_t* oi, cudaStream_t st, int32_t* osrc = nullptr);
void profile_begin(cudaStream_t st);
void profile_end(cudaStream_t st);
// b200r_profile_enable(2): every launch site is bracketed by CUDA events under its name (b200r_profile_dump)
struct ProfScope {
  cudaStream_t st; const char* name; cudaEvent_t a, b; bool on;
  ProfScope(cudaStream_t stream, const char* name);
  ~ProfScope();
};
int launch_expand_rescore(const Index* ix, const Graph* g, const void* Q, const int64_t* seeds, int B, int S, int window,
                          float alpha_dense, float alpha_graph, int k, float* out_scores, int64_t* out_ids, int32_t* out_hops,
                          float* out_gn, float* out_dn, float* out_cos, int32_t* out_ncand, int32_t* out_status, int mode,
                          int64_t own_lo, int64_t own_hi, void* ws, float* minmax, cudaStream_t st);
size_t expand_workspace_bytes(int B);
size_t expand_big_workspace_bytes(long long V);
int launch_expand_rescore_big(const Index* ix, const Graph* g, const void* Q, const int64_t* seeds, int S, int window, float alpha_dense,
                              float alpha_graph, int k, float* out_scores, int64_t* out_ids, int32_t* out_hops, float* out_gn, float* out_dn,
                              float* out_cos, int32_t* out_ncand, void* ws, cudaStream_t st);
// graph stage over the exchange (csr_expand.cu, mode 3): slot sizes of its three channels and the phased launcher
size_t expand_xchg_req_bytes(int per);          // one home rank's request record for `per` queries
size_t expand_xchg_cos_bytes(int per);          // a home rank's cosine slot
size_t expand_xchg_out_bytes(int B, int k);     // the result slot
struct Comm;
int launch_expand_rescore_exchange(const Index* ix, const Graph* g, const Comm* comm, const void* Q, const int64_t* seeds, int B, int S,
                                   int window, float alpha_dense, float alpha_graph, int k, int64_t n_total_rows, float* out_scores,
                                   int64_t* out_ids, int32_t* out_hops, float* out_gn, float* out_dn, float* out_cos, int32_t* out_ncand,
                                   int32_t* out_status, void* ws, int phase, cudaStream_t st);
int encode_tmap_2d(CUtensorMap* tm, int dtype, const void* base, uint64_t inner, uint64_t outer, uint32_t box_inner, uint32_t box_outer);

}  // namespace b200r
