// api.cu -- the C ABI of libb200r.so (declared in include/b200r.h): handles, planning, dispatch.
#include "internal.h"

#include <atomic>
#include <mutex>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <new>
#include <cstdint>
#include <vector>
#include <string>

namespace b200r {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms() {
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cache[dev] == 0) {
    int n = 0;
    cache[dev] = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
  }
  return cache[dev];
}

// ---- optional event bracketing of the dominant kernel (b200r_profile_*) ----
// Host threads may search concurrently (one per device in a multi-GPU host): the switches are atomics, every thread
// keeps its own open bracket, and finished (start, stop) pairs are appended under one mutex.
static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mu;
struct ProfPair { cudaEvent_t a, b; };
static std::vector<ProfPair> g_prof_events;
static thread_local cudaEvent_t t_prof_open = nullptr;
void profile_begin(cudaStream_t st) {
  if (!g_prof_on.load(std::memory_order_relaxed) || t_prof_open) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) == cudaSuccess) { cudaEventRecord(e, st); t_prof_open = e; }
}
void profile_end(cudaStream_t st) {
  if (!t_prof_open) return;
  cudaEvent_t a = t_prof_open, e;
  t_prof_open = nullptr;
  if (cudaEventCreate(&e) != cudaSuccess) { cudaEventDestroy(a); return; }
  cudaEventRecord(e, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_events.push_back({a, e});
}

// ---- per-launch-site timeline (b200r_profile_enable(2) / b200r_profile_dump) ----
struct ProfRec { const char* name; cudaEvent_t a, b; };
static std::vector<ProfRec> g_prof_recs;
static std::atomic<bool> g_prof_all{false};
ProfScope::ProfScope(cudaStream_t stream, const char* nm) : st(stream), name(nm), a(nullptr), b(nullptr), on(false) {
  if (!g_prof_all.load(std::memory_order_relaxed)) return;
  if (cudaEventCreate(&a) != cudaSuccess) return;
  if (cudaEventCreate(&b) != cudaSuccess) { cudaEventDestroy(a); return; }
  cudaEventRecord(a, st);
  on = true;
}
ProfScope::~ProfScope() {
  if (!on) return;
  cudaEventRecord(b, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_recs.push_back({name, a, b});
}

// B200R_CTA_GROUP=1 forces the single-CTA tile kernel, B200R_Q_REPLICAS=0 disables the query-tile replication (A/B measurements)
static int g_q_replicas = [] { const char* e = getenv("B200R_Q_REPLICAS"); return (e && e[0] == '0') ? 0 : 1; }();
static int g_cta_group_max = [] { const char* e = getenv("B200R_CTA_GROUP"); return (e && e[0] == '1') ? 1 : 2; }();
int select_group_size();
int select_max_flat_segments();
int gemm_tile_rows();
int gemm_tile_queries();
int32_t expand_max_candidates();

// cuTensorMapEncodeTiled through the runtime's driver entry point: no link-time libcuda dependency
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_tmap_2d(CUtensorMap* tm, int dtype, const void* base, uint64_t inner, uint64_t outer, uint32_t box_inner, uint32_t box_outer) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !sym) {
      set_error("cuTensorMapEncodeTiled entry point unavailable: %s", cudaGetErrorString(e));
      return B200R_ECUDA;
    }
    fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  const cuuint64_t dims[2] = {inner, outer};
  if (dtype < B200R_BF16 || dtype > B200R_I8) { set_error("no tensor map for dtype %d", dtype); return B200R_EINVAL; }
  const cuuint64_t strides[1] = {inner * (dtype == B200R_I8 ? 1 : (dtype == B200R_F32 ? 4 : 2))};  // bytes
  const cuuint32_t box[2] = {box_inner, box_outer};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = (dtype == B200R_BF16) ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                 : (dtype == B200R_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                    : (dtype == B200R_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8));
  // 256-byte L2 promotion measured best for the streamed corpus tiles (5.22 vs 5.10 TB/s with 128 B / none)
  CUresult r = fn(tm, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (base=%p inner=%llu outer=%llu)", (int)r, base,
              (unsigned long long)inner, (unsigned long long)outer);
    return B200R_ECUDA;
  }
  return B200R_OK;
}

static Index* get_index(uint64_t h) {
  Index* ix = reinterpret_cast<Index*>(h);
  if (!ix || ix->magic != kIndexMagic) { set_error("bad index handle"); return nullptr; }
  return ix;
}
static Graph* get_graph(uint64_t h) {
  Graph* g = reinterpret_cast<Graph*>(h);
  if (!g || g->magic != kGraphMagic) { set_error("bad graph handle"); return nullptr; }
  return g;
}

int make_plan(const Index* ix, int B, int k, int path, SearchPlan* pl) {
  if (B <= 0) { set_error("B must be > 0 (got %d)", B); return B200R_EINVAL; }
  if (k <= 0 || k > B200R_MAX_K) { set_error("k must be in [1, %d] (got %d)", B200R_MAX_K, k); return B200R_EINVAL; }
  const bool gemm_ok = ix->has_tmap;
  // measured on B200 (10M x 768 bf16): with the query tile replicated in global memory the tcgen05 tile kernel streams
  // the corpus at 6.9-7.0 TB/s (105% of the measured copy bandwidth) for every B from 1 to 128 and is tensor-bound
  // beyond; the CUDA-core scan (6.7 TB/s at B=1, 5.6 at B=2) serves fp32 / int8 / odd-D corpora
  // int8 corpora: tcgen05 kind::i8 with the queries split into signed-byte limbs, 16-64 queries per corpus pass (rows_i8.cu)
  // fp32 corpora: tcgen05 kind::tf32 with both operands split into two exact tf32 pieces, 16 queries per corpus pass (rows_f32.cu)
  const int f32_nq = ix->has_tmap_f32 ? rows_f32_nq(ix->D) : 0;
  const int i8_nq = ix->has_tmap_i8 ? rows_i8_nq(B, ix->D) : f32_nq;
  const bool i8_ok = i8_nq > 0;
  if (path == B200R_PATH_AUTO) path = (gemm_ok || i8_ok) ? B200R_PATH_GEMM : B200R_PATH_SCAN;
  if (path == B200R_PATH_GEMM && !gemm_ok && !i8_ok) {
    set_error("tensor path needs a bf16/f16 corpus with D %% 8 == 0, an int8 corpus or an fp32 corpus (dtype=%d D=%d)", ix->dtype, ix->D);
    return B200R_EINVAL;
  }
  if (path != B200R_PATH_GEMM && path != B200R_PATH_SCAN) { set_error("bad path %d", path); return B200R_EINVAL; }
  memset(pl, 0, sizeof(*pl));
  pl->path = path; pl->B = B; pl->k = k;
  pl->cap_soft = list_cap_soft(k); pl->cap = list_cap(k);
  if (path == B200R_PATH_GEMM) {
    // more than one 128-query tile: pair the CTAs (cta_group::2) so each loads half of every corpus tile
    // int8: rows-mode kernel (corpus rows on the MMA's M, 128-row tiles, one list per epilogue warp)
    pl->rows_i8_nq = (ix->dtype == B200R_I8 || ix->dtype == B200R_F32) ? i8_nq : 0;
    pl->rows_f32 = (ix->dtype == B200R_F32) ? 1 : 0;
    pl->rows_i8 = pl->rows_i8_nq ? (B + i8_nq - 1) / i8_nq : 0;
    pl->cta_group = (!pl->rows_i8 && B > gemm_tile_queries() && g_cta_group_max >= 2) ? 2 : 1;
    const int BMq = pl->rows_i8 ? pl->rows_i8_nq : gemm_tile_queries() * pl->cta_group;
    const int BNr = pl->rows_i8 ? rows_i8_tile_rows() : gemm_tile_rows();
    const int list_rows_per_tile = pl->rows_i8 ? 32 : BNr;      // rows one list can receive from one tile
    pl->num_q_tiles = pl->rows_i8 ? 1 : (B + BMq - 1) / BMq;      // int8: every group of queries is its own launch
    pl->num_row_tiles = (int)((ix->N + BNr - 1) / BNr);
    pl->Bpad = (pl->rows_i8 ? pl->rows_i8 : pl->num_q_tiles) * BMq;
    // units = (query tile, chunk of row tiles).  More units per CTA = better balance, fewer = less list
    // traffic; pick the chunk count (near a target) that wastes the fewest CTA slots in the last wave.
    // int8: the query groups of a slice share one launch, num_sms / groups CTAs each (rows_i8.cu), so the chunks are
    // sized for that many CTAs per group
    const int grid_max = (pl->rows_i8 && !pl->rows_f32) ? rows_i8_ctas_per_group(pl->rows_i8) : num_sms() / pl->cta_group;
    const int m = pl->num_q_tiles >= 16 ? 8 : (pl->num_q_tiles >= 4 ? 4 : (pl->num_q_tiles >= 2 ? 2 : 1));
    long long target = ((long long)grid_max * m + pl->num_q_tiles - 1) / pl->num_q_tiles;
    if (target < 1) target = 1;
    if (target > pl->num_row_tiles) target = pl->num_row_tiles;
    int best = (int)target; double best_waste = 1e30;
    for (long long c = target; c <= target + 40 && c <= pl->num_row_tiles; ++c) {
      const int tpc = (int)((pl->num_row_tiles + c - 1) / c);
      const long long cc = (pl->num_row_tiles + tpc - 1) / tpc;
      const long long units = cc * pl->num_q_tiles;
      const long long waves = (units + grid_max - 1) / grid_max;
      const double waste = (double)(waves * grid_max) / (double)units * (double)(tpc * cc) / (double)pl->num_row_tiles;
      if (waste < best_waste - 1e-9) { best_waste = waste; best = (int)cc; }
    }
    {
      const int tpc = (pl->num_row_tiles + best - 1) / best;
      pl->num_chunks = (pl->num_row_tiles + tpc - 1) / tpc;
    }
    // Phase schedule.  Phase 0 starts with no threshold, so every score is a candidate: keep it tiny and
    // spread it over enough lists that none overflows (rows per list <= cap_soft - 32).  Each later phase
    // covers ~8x the rows seen so far, which bounds the survivors per query per phase by ~7k.
    const int T = pl->num_row_tiles;
    const int C = pl->num_chunks;
    int seen = 0, np = 0;
    // The slow-path work of a phase is ~ B * k * ratio appends, so small batches afford a much coarser schedule
    // (fewer launches and threshold selects, which is what matters for batch-1 latency on small shards).
    static const int env_ratio = [] { const char* e = getenv("B200R_PHASE_RATIO"); return e ? atoi(e) : 0; }();   // A/B knob
    const int ratio = env_ratio >= 2 ? env_ratio : (B <= 16 ? 64 : (B <= 128 ? 16 : 8));
    int want0 = (10 * k + BNr - 1) / BNr;
    if (want0 < 4) want0 = 4;
    // a handful of queries: make the unfiltered first phase 1/64 of the corpus (one tile per chunk, all chunks busy) so
    // that ONE more phase finishes the scan -- two launches and one select per search instead of three or four
    if (B <= 16 && want0 < pl->num_row_tiles / 64) want0 = pl->num_row_tiles / 64;
    int tpc0 = (pl->cap_soft - 32) / list_rows_per_tile;
    if (tpc0 < 1) tpc0 = 1;
    if (want0 > C * tpc0) want0 = C * tpc0;
    while (seen < T && np < kMaxPhases) {
      int end;
      if (np == 0) end = want0 < T ? want0 : T;
      else {
        long long e = (long long)seen * ratio;
        end = e < T ? (int)e : T;
        if (T - end < end / 2 || np == kMaxPhases - 1) end = T;   // fold a short tail into this phase
      }
      GemmPhase& ph = pl->phases[np];
      ph.tile_begin = seen; ph.tile_end = end;
      const int tiles = end - seen;
      int chunks = tiles < C ? tiles : C;
      if (np == 0) { const int need = (tiles + tpc0 - 1) / tpc0; chunks = need < C ? need : C; }
      ph.tiles_per_chunk = (tiles + chunks - 1) / chunks;
      ph.num_chunks = (tiles + ph.tiles_per_chunk - 1) / ph.tiles_per_chunk;
      seen = end; ++np;
    }
    pl->num_phases = np;
    pl->num_segments = pl->num_chunks * (pl->rows_i8 ? 4 : 1);
  } else {
    pl->Bpad = B;
    const int bulk_warps = scan_bulk_warps(ix);
    if (bulk_warps) {
      // fp32 / int8: bulk-copy-fed scan, one persistent CTA per SM, every consumer warp keeps its own lists
      pl->scan_bulk = 1;
      const long long need = (ix->N + scan_bulk_tile_rows(ix) - 1) / scan_bulk_tile_rows(ix);
      long long blocks = num_sms();
      if (blocks > need) blocks = need < 1 ? 1 : need;
      pl->scan_blocks = (int)blocks;
      pl->scan_warps = (int)blocks * bulk_warps;
    } else {
      const int R = 4;
      const int warps_per_block = 8;
      long long blocks = (long long)num_sms() * 3;   // scan kernels are compiled for 3 CTAs/SM: one full wave
      const long long groups = (ix->N + R - 1) / R;
      const long long need = (groups + warps_per_block - 1) / warps_per_block;
      if (blocks > need) blocks = need < 1 ? 1 : need;
      pl->scan_blocks = (int)blocks;
      pl->scan_warps = (int)blocks * warps_per_block;
    }
    pl->num_segments = pl->scan_warps;
  }
  size_t off = 0;
  pl->off_lists = off; off += align_up((size_t)pl->num_segments * pl->Bpad * pl->cap * 8, 256);
  // counts[num_segments][Bpad] followed by the carry lists' counts [Bpad] (zeroed together by init_search)
  pl->off_counts = off; off += align_up((size_t)(pl->num_segments + 1) * pl->Bpad * 4, 256);
  if (path == B200R_PATH_GEMM && pl->num_phases > 1 && pl->num_segments + 1 <= select_max_flat_segments()) {
    pl->carry_cap = select_stage_limit(k, B);
    pl->off_carry = off; off += align_up((size_t)pl->Bpad * pl->carry_cap * 8, 256);
  }
  pl->off_gthr = off; off += align_up((size_t)pl->Bpad * 4, 256);
  pl->off_qinv = off; off += align_up((size_t)pl->Bpad * 4, 256);
  pl->num_groups = 0;
  // Two-level final select for the thousands of short lists of the scan paths.  A handful of queries on the tile kernel
  // (<= 148 pruned lists each) take ONE wide CTA per query instead: their candidates (~k + the last phase's survivors) fit
  // its 192 KB stage, and the second launch costs more than it saves (batch 1: 21 + 19 us against ~22 us).
  const bool one_level = path == B200R_PATH_GEMM && !pl->rows_i8 && B <= 16 && k <= 256 && pl->num_phases > 1 && pl->num_segments + 1 <= select_max_flat_segments();
  if (pl->num_segments > 2 * select_group_size() && !one_level) {
    pl->num_groups = (pl->num_segments + select_group_size() - 1) / select_group_size();
    pl->off_lists2 = off; off += align_up((size_t)pl->num_groups * pl->Bpad * pl->k * 8, 256);
    pl->off_counts2 = off; off += align_up((size_t)pl->num_groups * pl->Bpad * 4, 256);
  }
  pl->q_replicas = (path == B200R_PATH_GEMM && !pl->rows_i8 && pl->cta_group == 1 && pl->num_q_tiles == 1 && g_q_replicas) ? 1 : 0;
  if (pl->q_replicas) { pl->off_qrep = off; off += align_up(query_replica_bytes(ix->D), 256); }
  if (pl->rows_i8 && pl->rows_f32) { pl->off_qrep = off; off += rows_f32_query_bytes(B, ix->D); }                   // hi / mid query planes
  else if (pl->rows_i8) { pl->off_qrep = off; off += align_up(rows_i8_query_bytes(B, pl->rows_i8_nq, ix->D), 256); }   // quantised query limbs + scales
  pl->off_end = off;
  return B200R_OK;
}

static int query_dtype(const Index* ix) { return ix->dtype == B200R_I8 ? B200R_F32 : ix->dtype; }

// The path `auto` resolves to (make_plan applies the same rule).
static int resolve_path(const Index* ix, int B, int path) {
  if (path != B200R_PATH_AUTO) return path;
  const bool i8_ok = ix->has_tmap_i8 && rows_i8_nq(B, ix->D) > 0;
  const bool f32_ok = ix->has_tmap_f32 && rows_f32_nq(ix->D) > 0;
  return (ix->has_tmap || i8_ok || f32_ok) ? B200R_PATH_GEMM : B200R_PATH_SCAN;
}

// Queries one planned search may carry.  The scan kernels and the int8 tensor kernel keep one candidate list per
// (consumer warp | epilogue warp, query), i.e. megabytes of workspace per query: bigger batches run as consecutive
// slices over the same workspace (the corpus is streamed once per <= 8 / <= 64 queries either way).
static int slice_queries(const Index* ix, int B, int k, int path) {
  const int resolved = resolve_path(ix, B, path);
  if (resolved == B200R_PATH_SCAN) return k <= 256 ? 32 : 16;
  if (ix->dtype == B200R_I8 || ix->dtype == B200R_F32) return 256;
  return 1 << 30;
}

static int do_search_one(Index* ix, const void* Q, int B, int k, int path, float* out_scores, int64_t* out_ids, void* ws,
                         size_t ws_bytes, cudaStream_t st, const CommDev* comm, size_t rec_stride, float min_score) {
  SearchPlan pl;
  int rc = make_plan(ix, B, k, path, &pl);
  if (rc) return rc;
  if (!ws || ws_bytes < pl.off_end) {
    set_error("workspace too small: need %zu bytes, got %zu", pl.off_end, ws_bytes);
    return B200R_ENOMEM;
  }
  const bool k1 = pl.path == B200R_PATH_GEMM && !pl.rows_i8;      // bf16 / fp16 tile kernel: one fused prologue launch
  if (k1) {
    if ((rc = launch_search_prologue(pl, ws, Q, ix->D, query_dtype(ix), ix->metric == B200R_COS, st))) return rc;
  } else {
    if ((rc = launch_init_search(pl, ws, st))) return rc;
    if (ix->metric == B200R_COS) {
      float* qinv = reinterpret_cast<float*>((char*)ws + pl.off_qinv);
      if ((rc = launch_query_inv_norm(Q, B, ix->D, query_dtype(ix), qinv, st))) return rc;
    }
  }
  const bool floor_on = min_score > -HUGE_VALF;
  if (floor_on && (rc = launch_min_score(pl, ws, ix->metric == B200R_COS, min_score, st))) return rc;
  if (pl.path == B200R_PATH_GEMM && pl.rows_i8 && pl.rows_f32) {
    void* qws = (char*)ws + pl.off_qrep;
    if ((rc = launch_split_queries_tf32(Q, B, ix->D, qws, st))) return rc;
    // one launch per group of 16 queries per phase: ganging the groups like the int8 kernel was measured and dropped -- this
    // kernel is bound by its per-SM converter / MMA pipeline, not by HBM, so sharing tiles through L2 gains nothing and the
    // extra index arithmetic cost 11 % (profiles/r2_f32_mid_buffers_and_gang_ab.jsonl)
    std::vector<CUtensorMap> tmaps(pl.rows_i8);
    for (int g = 0; g < pl.rows_i8; ++g)
      if ((rc = encode_query_tmap_f32(qws, g, ix->D, &tmaps[g]))) return rc;
    for (int i = 0; i < pl.num_phases; ++i) {
      profile_begin(st);
      for (int g = 0; g < pl.rows_i8 && !rc; ++g) rc = launch_rows_f32(ix, &tmaps[g], g, pl, pl.phases[i], ws, st);
      profile_end(st);
      if (rc) return rc;
      if (i + 1 < pl.num_phases && (rc = launch_select_lists(ix, pl, ws, nullptr, nullptr, st))) return rc;
    }
  } else if (pl.path == B200R_PATH_GEMM && pl.rows_i8) {
    void* qws = (char*)ws + pl.off_qrep;
    if ((rc = launch_quantize_queries_i8(Q, B, pl.rows_i8_nq, ix->D, qws, st))) return rc;
    CUtensorMap tmap_q;
    if ((rc = encode_query_tmap_i8(qws, pl.rows_i8, pl.rows_i8_nq, ix->D, &tmap_q))) return rc;
    for (int i = 0; i < pl.num_phases; ++i) {
      profile_begin(st);
      rc = launch_rows_i8(ix, &tmap_q, qws, pl, pl.phases[i], ws, st);
      profile_end(st);
      if (rc) return rc;
      if (i + 1 < pl.num_phases && (rc = launch_select_lists(ix, pl, ws, nullptr, nullptr, st))) return rc;
    }
  } else if (pl.path == B200R_PATH_GEMM) {
    CUtensorMap tmap_q;
    if (pl.q_replicas) rc = encode_query_tmap_replicated(ix, (char*)ws + pl.off_qrep, &tmap_q);   // written by the prologue
    else rc = encode_query_tmap(ix, Q, B, &tmap_q);
    if (rc) return rc;
    for (int i = 0; i < pl.num_phases; ++i) {
      profile_begin(st);
      rc = launch_gemm_topk(ix, &tmap_q, pl, pl.phases[i], ws, st);
      profile_end(st);
      if (rc) return rc;
      if (i + 1 < pl.num_phases && (rc = launch_select_lists(ix, pl, ws, nullptr, nullptr, st))) return rc;
    }
  } else {
    profile_begin(st);
    rc = pl.scan_bulk ? launch_scan_bulk(ix, Q, pl, ws, st) : launch_scan_topk(ix, Q, pl, ws, st);
    profile_end(st);
    if (rc) return rc;
  }
  return launch_select_lists(ix, pl, ws, out_scores, out_ids, st, comm, rec_stride, min_score);
}

// workspace of a whole (possibly sliced) search; *temp_off = where the packed [ids | scores] record of a sliced
// search under an exchange is staged before it is scattered to the peers
static size_t search_workspace(const Index* ix, int B, int k, int path, size_t* temp_off) {
  const int S = slice_queries(ix, B, k, path);
  SearchPlan pl;
  if (make_plan(ix, B < S ? B : S, k, resolve_path(ix, B, path), &pl)) return 0;
  size_t need = pl.off_end;
  if (B > S && B % S) {          // a short last slice may plan a different layout: cover both
    SearchPlan pt;
    if (make_plan(ix, B % S, k, resolve_path(ix, B, path), &pt)) return 0;
    if (pt.off_end > need) need = pt.off_end;
  }
  if (temp_off) *temp_off = need;
  if (B > S) need += align_up((size_t)12 * B * k, 256);
  return need;
}

// comm != nullptr: instead of a local result, this rank's packed record goes to every rank's exchange window
static int do_search(Index* ix, const void* Q, int B, int k, int path, float* out_scores, int64_t* out_ids, void* ws,
                     size_t ws_bytes, cudaStream_t st, const Comm* comm = nullptr, float min_score = -HUGE_VALF) {
  if (B <= 0) { set_error("B must be > 0 (got %d)", B); return B200R_EINVAL; }
  if (k <= 0 || k > B200R_MAX_K) { set_error("k must be in [1, %d] (got %d)", B200R_MAX_K, k); return B200R_EINVAL; }
  if (!Q || (!comm && (!out_scores || !out_ids))) { set_error("null Q / output pointer"); return B200R_EINVAL; }
  if ((reinterpret_cast<uintptr_t>(Q) & 15) || (reinterpret_cast<uintptr_t>(ws) & 255)) {
    set_error("Q must be 16-byte aligned and the workspace 256-byte aligned");
    return B200R_EINVAL;
  }
  size_t temp_off = 0;
  const size_t need = search_workspace(ix, B, k, path, &temp_off);
  if (need == 0) return B200R_EINVAL;
  if (!ws || ws_bytes < need) { set_error("workspace too small: need %zu bytes, got %zu", need, ws_bytes); return B200R_ENOMEM; }
  CommDev cd;
  const size_t stride = dense_record_stride(B, k);
  int rc;
  if (comm && (rc = comm_dev(comm, kChanDense, stride, &cd))) return rc;
  const int S = slice_queries(ix, B, k, path);
  const int resolved = resolve_path(ix, B, path);
  if (B <= S) return do_search_one(ix, Q, B, k, resolved, out_scores, out_ids, ws, ws_bytes, st, comm ? &cd : nullptr, stride, min_score);
  float* os = out_scores; int64_t* oi = out_ids;
  if (comm) {
    oi = reinterpret_cast<int64_t*>((char*)ws + temp_off);
    os = reinterpret_cast<float*>((char*)ws + temp_off + (size_t)8 * B * k);
  }
  const size_t qrow = (size_t)ix->D * elem_size(query_dtype(ix));
  for (int b0 = 0; b0 < B; b0 += S) {
    const int nb = B - b0 < S ? B - b0 : S;
    rc = do_search_one(ix, (const char*)Q + (size_t)b0 * qrow, nb, k, resolved, os + (size_t)b0 * k, oi + (size_t)b0 * k, ws, temp_off, st,
                       nullptr, 0, min_score);
    if (rc) return rc;
  }
  if (comm) return launch_scatter_record(cd, (char*)ws + temp_off, (size_t)12 * B * k, stride, st);
  return B200R_OK;
}

// (re)size the library-owned device staging of the *_host calls; the stream is drained before a buffer is replaced
static int ensure_staging(Index* ix, size_t qbytes, size_t oelems, size_t wsbytes, cudaStream_t st, size_t q32bytes = 0) {
  if (ix->q_dev_bytes >= qbytes && ix->out_dev_elems >= oelems && ix->ws_bytes >= wsbytes && ix->q32_dev_bytes >= q32bytes) return B200R_OK;
  B200R_CUDA_CHECK(cudaStreamSynchronize(st));
  if (ix->q32_dev_bytes < q32bytes) {
    if (ix->h_q32_dev) cudaFree(ix->h_q32_dev);
    ix->h_q32_dev = nullptr; ix->q32_dev_bytes = 0;
    B200R_CUDA_CHECK(cudaMalloc(&ix->h_q32_dev, q32bytes));
    ix->q32_dev_bytes = q32bytes;
  }
  if (ix->q_dev_bytes < qbytes) {
    if (ix->h_q_dev) cudaFree(ix->h_q_dev);
    ix->h_q_dev = nullptr; ix->q_dev_bytes = 0;
    B200R_CUDA_CHECK(cudaMalloc(&ix->h_q_dev, qbytes));
    ix->q_dev_bytes = qbytes;
  }
  if (ix->out_dev_elems < oelems) {
    if (ix->h_s_dev) cudaFree(ix->h_s_dev);
    if (ix->h_i_dev) cudaFree(ix->h_i_dev);
    ix->h_s_dev = nullptr; ix->h_i_dev = nullptr; ix->out_dev_elems = 0;
    B200R_CUDA_CHECK(cudaMalloc(&ix->h_s_dev, oelems * sizeof(float)));
    B200R_CUDA_CHECK(cudaMalloc(&ix->h_i_dev, oelems * sizeof(int64_t)));
    ix->out_dev_elems = oelems;
  }
  if (ix->ws_bytes < wsbytes) {
    if (ix->h_ws) cudaFree(ix->h_ws);
    ix->h_ws = nullptr; ix->ws_bytes = 0;
    B200R_CUDA_CHECK(cudaMalloc(&ix->h_ws, wsbytes));
    ix->ws_bytes = wsbytes;
  }
  return B200R_OK;
}

// Host queries -> ix->h_q_dev on `st`.  from_f32: the host rows are fp32 whatever the index's query dtype; a 2-byte query
// dtype is produced on the device (RNE), so the host never converts.  The staging must have been sized by ensure_staging.
static int stage_queries(Index* ix, const void* Q_host, int B, bool from_f32, cudaStream_t st) {
  const size_t n = (size_t)B * ix->D;
  const int qd = query_dtype(ix);
  if (!from_f32 || qd == B200R_F32) {
    B200R_CUDA_CHECK(cudaMemcpyAsync(ix->h_q_dev, Q_host, n * elem_size(qd), cudaMemcpyHostToDevice, st));
    return B200R_OK;
  }
  B200R_CUDA_CHECK(cudaMemcpyAsync(ix->h_q32_dev, Q_host, n * sizeof(float), cudaMemcpyHostToDevice, st));
  return launch_convert_queries(reinterpret_cast<const float*>(ix->h_q32_dev), ix->h_q_dev, n, qd, st);
}
static size_t q32_staging_bytes(const Index* ix, int B, bool from_f32) {
  return (from_f32 && query_dtype(ix) != B200R_F32) ? (size_t)B * ix->D * sizeof(float) : 0;
}

}  // namespace b200r

using namespace b200r;

extern "C" {

int b200r_version(void) { return B200R_VERSION; }

int b200r_last_error(char* buf, size_t n) {
  const size_t len = strlen(g_err);
  if (buf && n) { strncpy(buf, g_err, n - 1); buf[n - 1] = 0; }
  return (int)len;
}

int64_t b200r_launch_count(void) { return (int64_t)g_launches.load(); }

int b200r_index_create(int device, const void* X, int64_t N, int32_t D, int32_t dtype, int32_t metric, int64_t id_base,
                       const float* row_scale, void* stream, uint64_t* out_index) {
  if (!out_index) { set_error("out_index is null"); return B200R_EINVAL; }
  *out_index = 0;
  if (!X || N <= 0 || D <= 0) { set_error("X must be non-null and N, D > 0 (N=%lld D=%d)", (long long)N, D); return B200R_EINVAL; }
  if (N > 0xffffffffll) { set_error("N=%lld exceeds the 2^32-1 rows one shard can hold", (long long)N); return B200R_EINVAL; }
  if (dtype < B200R_BF16 || dtype > B200R_I8) { set_error("bad dtype %d", dtype); return B200R_EINVAL; }
  if (metric != B200R_IP && metric != B200R_COS) { set_error("bad metric %d", metric); return B200R_EINVAL; }
  if (((size_t)D * elem_size(dtype)) % 16 != 0 || (reinterpret_cast<uintptr_t>(X) & 15)) {
    set_error("rows must be 16-byte aligned: D*sizeof(elem) %% 16 == 0 and X 16-byte aligned (D=%d dtype=%d)", D, dtype);
    return B200R_EINVAL;
  }
  B200R_CUDA_CHECK(cudaSetDevice(device));
  Index* ix = new (std::nothrow) Index();
  if (!ix) { set_error("out of host memory"); return B200R_ENOMEM; }
  memset(ix, 0, sizeof(*ix));
  ix->magic = kIndexMagic; ix->device = device; ix->X = X; ix->N = N; ix->D = D; ix->dtype = dtype; ix->metric = metric;
  ix->id_base = id_base;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool need_mult = (metric == B200R_COS) || (dtype == B200R_I8 && row_scale);
  if (need_mult) {
    cudaError_t e = cudaMalloc(&ix->row_mult, (size_t)N * sizeof(float));
    if (e != cudaSuccess) { set_error("cudaMalloc(row_mult) failed: %s", cudaGetErrorString(e)); delete ix; return B200R_ENOMEM; }
    int rc = launch_row_mult(ix, row_scale, st);
    if (rc) { cudaFree(ix->row_mult); delete ix; return rc; }
  }
  if ((dtype == B200R_BF16 || dtype == B200R_F16) && D % 8 == 0) {
    int rc = encode_tmap_2d(&ix->tmap_x, dtype, X, (uint64_t)D, (uint64_t)N, 64, (uint32_t)gemm_tile_rows());
    if (!rc) rc = encode_tmap_2d(&ix->tmap_x_half, dtype, X, (uint64_t)D, (uint64_t)N, 64, (uint32_t)gemm_tile_rows() / 2);
    if (rc) { if (ix->row_mult) cudaFree(ix->row_mult); delete ix; return rc; }
    ix->has_tmap = true;
  }
  if (dtype == B200R_F32 && rows_f32_nq(D) > 0) {   // fp32 tensor path (rows_f32.cu): tiles of 128 rows x 32 floats
    int rc = encode_tmap_2d(&ix->tmap_x_f32, dtype, X, (uint64_t)D, (uint64_t)N, 32, (uint32_t)rows_f32_tile_rows());
    if (rc) { if (ix->row_mult) cudaFree(ix->row_mult); delete ix; return rc; }
    ix->has_tmap_f32 = true;
  }
  if (dtype == B200R_I8) {   // int8 tensor path (rows_i8.cu): tiles of 128 rows x 128 bytes
    int rc = encode_tmap_2d(&ix->tmap_x_i8, dtype, X, (uint64_t)D, (uint64_t)N, 128, (uint32_t)rows_i8_tile_rows());
    if (rc) { if (ix->row_mult) cudaFree(ix->row_mult); delete ix; return rc; }
    ix->has_tmap_i8 = true;
  }
  cudaError_t e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    set_error("index build failed: %s", cudaGetErrorString(e));
    if (ix->row_mult) cudaFree(ix->row_mult);
    delete ix;
    return B200R_ECUDA;
  }
  *out_index = reinterpret_cast<uint64_t>(ix);
  return B200R_OK;
}

int b200r_index_destroy(uint64_t index) {
  Index* ix = get_index(index);
  if (!ix) return B200R_ESTATE;
  cudaSetDevice(ix->device);
  if (ix->row_mult) cudaFree(ix->row_mult);
  if (ix->h_q_dev) cudaFree(ix->h_q_dev);
  if (ix->h_q32_dev) cudaFree(ix->h_q32_dev);
  if (ix->h_s_dev) cudaFree(ix->h_s_dev);
  if (ix->h_i_dev) cudaFree(ix->h_i_dev);
  if (ix->h_ws) cudaFree(ix->h_ws);
  ix->magic = 0;
  delete ix;
  return B200R_OK;
}

int b200r_index_info(uint64_t index, int64_t* N, int32_t* D, int32_t* dtype, int32_t* metric, int64_t* id_base) {
  Index* ix = get_index(index);
  if (!ix) return B200R_ESTATE;
  if (N) *N = ix->N;
  if (D) *D = ix->D;
  if (dtype) *dtype = ix->dtype;
  if (metric) *metric = ix->metric;
  if (id_base) *id_base = ix->id_base;
  return B200R_OK;
}

size_t b200r_search_workspace_bytes(uint64_t index, int32_t B, int32_t k, int32_t path) {
  Index* ix = get_index(index);
  if (!ix) return 0;
  if (B <= 0) { set_error("B must be > 0 (got %d)", B); return 0; }
  if (k <= 0 || k > B200R_MAX_K) { set_error("k must be in [1, %d] (got %d)", B200R_MAX_K, k); return 0; }
  return search_workspace(ix, B, k, path, nullptr);
}

int b200r_search(uint64_t index, const void* Q, int32_t B, int32_t k, int32_t path, float* out_scores, int64_t* out_ids,
                 void* workspace, size_t workspace_bytes, void* stream) {
  Index* ix = get_index(index);
  if (!ix) return B200R_ESTATE;
  B200R_CUDA_CHECK(cudaSetDevice(ix->device));
  return do_search(ix, Q, B, k, path, out_scores, out_ids, workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}

int b200r_search_min_score(uint64_t index, const void* Q, int32_t B, int32_t k, int32_t path, float min_score, float* out_scores,
                           int64_t* out_ids, void* workspace, size_t workspace_bytes, void* stream) {
  Index* ix = get_index(index);
  if (!ix) return B200R_ESTATE;
  if (!(min_score == min_score)) { set_error("min_score is NaN"); return B200R_EINVAL; }
  B200R_CUDA_CHECK(cudaSetDevice(ix->device));
  return do_search(ix, Q, B, k, path, out_scores, out_ids, workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream),
                   nullptr, min_score);
}

int b200r_convert_f32(int device, const float* src, void* dst, int64_t n, int32_t dtype, void* stream) {
  if (!src || !dst || n < 0) { set_error("null pointer / negative n"); return B200R_EINVAL; }
  B200R_CUDA_CHECK(cudaSetDevice(device));
  return launch_convert_queries(src, dst, (size_t)n, dtype, reinterpret_cast<cudaStream_t>(stream));
}

static int search_host_impl(uint64_t index, const void* Q_host, bool from_f32, int32_t B, int32_t k, int32_t path, float* out_scores_host,
                            int64_t* out_ids_host, void* stream) {
  Index* ix = get_index(index);
  if (!ix) return B200R_ESTATE;
  if (!Q_host || !out_scores_host || !out_ids_host) { set_error("null host pointer"); return B200R_EINVAL; }
  if (B <= 0 || k <= 0 || k > B200R_MAX_K) { set_error("bad B / k (%d / %d)", B, k); return B200R_EINVAL; }
  B200R_CUDA_CHECK(cudaSetDevice(ix->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t need = search_workspace(ix, B, k, path, nullptr);
  if (!need) return B200R_EINVAL;
  const size_t qbytes = (size_t)B * ix->D * elem_size(query_dtype(ix));
  const size_t oelems = (size_t)B * k;
  int rc = ensure_staging(ix, qbytes, oelems, need, st, q32_staging_bytes(ix, B, from_f32));
  if (rc) return rc;
  if ((rc = stage_queries(ix, Q_host, B, from_f32, st))) return rc;
  rc = do_search(ix, ix->h_q_dev, B, k, path, ix->h_s_dev, ix->h_i_dev, ix->h_ws, ix->ws_bytes, st);
  if (rc) return rc;
  B200R_CUDA_CHECK(cudaMemcpyAsync(out_scores_host, ix->h_s_dev, oelems * sizeof(float), cudaMemcpyDeviceToHost, st));
  B200R_CUDA_CHECK(cudaMemcpyAsync(out_ids_host, ix->h_i_dev, oelems * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  B200R_CUDA_CHECK(cudaStreamSynchronize(st));
  return B200R_OK;
}

int b200r_search_host(uint64_t index, const void* Q_host, int32_t B, int32_t k, int32_t path, float* out_scores_host,
                      int64_t* out_ids_host, void* stream) {
  return search_host_impl(index, Q_host, false, B, k, path, out_scores_host, out_ids_host, stream);
}

int b200r_search_host_f32(uint64_t index, const float* Q_host_f32, int32_t B, int32_t k, int32_t path, float* out_scores_host,
                          int64_t* out_ids_host, void* stream) {
  return search_host_impl(index, Q_host_f32, true, B, k, path, out_scores_host, out_ids_host, stream);
}

// ---- range search: every row with score >= min_score (gemm_topk.cu range mode) ----------------------------------------
// One pass over the corpus per call, single phase.  Workspace: counts, per-(chunk, query) offsets, query norms, the
// (unused) threshold words and, for one query tile, the replicated queries.
static int range_plan(const Index* ix, int B, SearchPlan* pl, size_t* off_slot) {
  if (!ix->has_tmap) { set_error("range search runs on the tensor path: bf16/f16 corpus with D %% 8 == 0 (dtype=%d D=%d)", ix->dtype, ix->D); return B200R_EINVAL; }
  int rc = make_plan(ix, B, 1, B200R_PATH_GEMM, pl);
  if (rc) return rc;
  pl->num_phases = 1;
  pl->phases[0].tile_begin = 0; pl->phases[0].tile_end = pl->num_row_tiles;
  pl->phases[0].tiles_per_chunk = (pl->num_row_tiles + pl->num_chunks - 1) / pl->num_chunks;
  pl->phases[0].num_chunks = (pl->num_row_tiles + pl->phases[0].tiles_per_chunk - 1) / pl->phases[0].tiles_per_chunk;
  size_t off = 0;
  pl->off_lists = 0;
  pl->off_counts = off; off += align_up((size_t)(pl->num_chunks + 1) * pl->Bpad * 4, 256);
  *off_slot = off; off += align_up((size_t)pl->num_chunks * pl->Bpad * 8, 256);
  pl->off_gthr = off; off += align_up((size_t)pl->Bpad * 4, 256);
  pl->off_qinv = off; off += align_up((size_t)pl->Bpad * 4, 256);
  pl->carry_cap = 0; pl->num_groups = 0;
  if (pl->q_replicas) { pl->off_qrep = off; off += align_up(query_replica_bytes(ix->D), 256); }
  pl->off_end = off;
  return B200R_OK;
}

static int range_pass(Index* ix, const void* Q, int B, float min_score, const int64_t* id_gt, int mode, const int64_t* slot_off, void* out,
                      const SearchPlan& pl, void* ws, cudaStream_t st) {
  int rc;
  float* qinv = nullptr;
  if (ix->metric == B200R_COS) {
    qinv = reinterpret_cast<float*>((char*)ws + pl.off_qinv);
    if ((rc = launch_query_inv_norm(Q, B, ix->D, query_dtype(ix), qinv, st))) return rc;
  }
  CUtensorMap tmap_q;
  if (pl.q_replicas) {
    void* rep = (char*)ws + pl.off_qrep;
    if ((rc = launch_replicate_queries(Q, rep, B, ix->D, st))) return rc;
    rc = encode_query_tmap_replicated(ix, rep, &tmap_q);
  } else {
    rc = encode_query_tmap(ix, Q, B, &tmap_q);
  }
  if (rc) return rc;
  RangeArgs rg;
  rg.mode = mode; rg.tau = min_score; rg.qinv = qinv; rg.id_gt = id_gt; rg.offsets = slot_off; rg.out = out;
  return launch_gemm_topk(ix, &tmap_q, pl, pl.phases[0], ws, st, &rg);
}

// ---- row-sharded search over the exchange (comm.cu) ----------------------------------------------------------------
int b200r_search_sharded_begin(uint64_t index, uint64_t comm, const void* Q, int32_t B, int32_t k, int32_t path, void* workspace,
                               size_t workspace_bytes, void* stream) {
  Index* ix = get_index(index);
  Comm* c = get_comm(comm);
  if (!ix || !c) return B200R_ESTATE;
  if (ix->device != c->device) { set_error("index and comm live on different devices"); return B200R_EINVAL; }
  B200R_CUDA_CHECK(cudaSetDevice(ix->device));
  return do_search(ix, Q, B, k, path, nullptr, nullptr, workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream), c);
}

int b200r_search_sharded_finish(uint64_t comm, int32_t B, int32_t k, float* out_scores, int64_t* out_ids, void* stream) {
  Comm* c = get_comm(comm);
  if (!c) return B200R_ESTATE;
  if (!out_scores || !out_ids || B <= 0 || k <= 0 || k > B200R_MAX_K) { set_error("bad arguments"); return B200R_EINVAL; }
  B200R_CUDA_CHECK(cudaSetDevice(c->device));
  return launch_exchange_merge(c, B, k, k, out_scores, out_ids, reinterpret_cast<cudaStream_t>(stream));
}

int b200r_search_sharded(uint64_t index, uint64_t comm, const void* Q, int32_t B, int32_t k, int32_t path, float* out_scores,
                         int64_t* out_ids, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = b200r_search_sharded_begin(index, comm, Q, B, k, path, workspace, workspace_bytes, stream);
  if (rc) return rc;
  return b200r_search_sharded_finish(comm, B, k, out_scores, out_ids, stream);
}

// One host thread driving G GPUs (the plug-in's single-process mode): host queries in, host results out.  Every device
// gets the queries, searches its row shard and scatters its record (begin on ALL devices before any finish, so no
// device's merge kernel waits for a kernel that has not been launched yet); the merged result is read from device 0.
static int search_sharded_host_all_impl(const uint64_t* indexes, const uint64_t* comms, int32_t G, const void* Q_host, bool from_f32,
                                        int32_t B, int32_t k, int32_t path, float* out_scores_host, int64_t* out_ids_host,
                                        void* const* streams) {
  if (!indexes || !comms || G < 1 || G > B200R_MAX_PEERS) { set_error("bad indexes / comms / G"); return B200R_EINVAL; }
  if (!Q_host || !out_scores_host || !out_ids_host) { set_error("null host pointer"); return B200R_EINVAL; }
  if (B <= 0 || k <= 0 || k > B200R_MAX_K) { set_error("bad B / k (%d / %d)", B, k); return B200R_EINVAL; }
  Index* ix[B200R_MAX_PEERS]; Comm* cm[B200R_MAX_PEERS]; cudaStream_t st[B200R_MAX_PEERS];
  for (int g = 0; g < G; ++g) {
    ix[g] = get_index(indexes[g]); cm[g] = get_comm(comms[g]);
    if (!ix[g] || !cm[g]) return B200R_ESTATE;
    if (cm[g]->rank != g || cm[g]->world != G || ix[g]->device != cm[g]->device) { set_error("comm %d does not match (rank/world/device)", g); return B200R_EINVAL; }
    st[g] = streams ? reinterpret_cast<cudaStream_t>(streams[g]) : nullptr;
  }
  const size_t oelems = (size_t)B * k;
  int rc;
  // H2D + local search (+ peer stores) of every device, issued from the calling thread.  One launch thread per device was
  // measured on 8 GPUs (profiles/r2_plugin_devices_host_threads_ab.jsonl): batch 1 0.645 -> 0.607 ms, but batch 4096 7.05 ->
  // 9.74 ms (the eight H2D copies of the queries then start together and contend, instead of each hiding behind the
  // previous devices' searches) -- not kept.
  for (int g = 0; g < G; ++g) {
    B200R_CUDA_CHECK(cudaSetDevice(ix[g]->device));
    const size_t need = search_workspace(ix[g], B, k, path, nullptr);
    if (!need) return B200R_EINVAL;
    const size_t qbytes = (size_t)B * ix[g]->D * elem_size(query_dtype(ix[g]));
    if ((rc = ensure_staging(ix[g], qbytes, g == 0 ? oelems : 1, need, st[g], q32_staging_bytes(ix[g], B, from_f32)))) return rc;
    if ((rc = stage_queries(ix[g], Q_host, B, from_f32, st[g]))) return rc;
    if ((rc = do_search(ix[g], ix[g]->h_q_dev, B, k, path, nullptr, nullptr, ix[g]->h_ws, ix[g]->ws_bytes, st[g], cm[g]))) return rc;
  }
  // only rank 0's merged copy is needed on the host; the other ranks still consume the exchange (epochs advance together)
  for (int g = 0; g < G; ++g) {
    B200R_CUDA_CHECK(cudaSetDevice(ix[g]->device));
    if (g == 0) rc = launch_exchange_merge(cm[0], B, k, k, ix[0]->h_s_dev, ix[0]->h_i_dev, st[0]);
    else rc = launch_exchange_skip(cm[g], st[g]);
    if (rc) return rc;
  }
  B200R_CUDA_CHECK(cudaSetDevice(ix[0]->device));
  B200R_CUDA_CHECK(cudaMemcpyAsync(out_scores_host, ix[0]->h_s_dev, oelems * sizeof(float), cudaMemcpyDeviceToHost, st[0]));
  B200R_CUDA_CHECK(cudaMemcpyAsync(out_ids_host, ix[0]->h_i_dev, oelems * sizeof(int64_t), cudaMemcpyDeviceToHost, st[0]));
  for (int g = 0; g < G; ++g) {
    B200R_CUDA_CHECK(cudaSetDevice(ix[g]->device));
    B200R_CUDA_CHECK(cudaStreamSynchronize(st[g]));
  }
  return B200R_OK;
}

int b200r_search_sharded_host_all(const uint64_t* indexes, const uint64_t* comms, int32_t G, const void* Q_host, int32_t B, int32_t k,
                                  int32_t path, float* out_scores_host, int64_t* out_ids_host, void* const* streams) {
  return search_sharded_host_all_impl(indexes, comms, G, Q_host, false, B, k, path, out_scores_host, out_ids_host, streams);
}

int b200r_search_sharded_host_all_f32(const uint64_t* indexes, const uint64_t* comms, int32_t G, const float* Q_host_f32, int32_t B,
                                      int32_t k, int32_t path, float* out_scores_host, int64_t* out_ids_host, void* const* streams) {
  return search_sharded_host_all_impl(indexes, comms, G, Q_host_f32, true, B, k, path, out_scores_host, out_ids_host, streams);
}

size_t b200r_range_workspace_bytes(uint64_t index, int32_t B) {
  Index* ix = get_index(index);
  if (!ix || B <= 0) return 0;
  SearchPlan pl; size_t off_slot;
  if (range_plan(ix, B, &pl, &off_slot)) return 0;
  return pl.off_end;
}

int b200r_range_count(uint64_t index, const void* Q, int32_t B, float min_score, const int64_t* id_gt, int64_t* out_offsets,
                      int64_t* total_host, void* workspace, size_t workspace_bytes, void* stream) {
  Index* ix = get_index(index);
  if (!ix) return B200R_ESTATE;
  if (!Q || !out_offsets || !total_host || B <= 0) { set_error("null pointer / bad B"); return B200R_EINVAL; }
  if (!(min_score == min_score)) { set_error("min_score is NaN"); return B200R_EINVAL; }
  B200R_CUDA_CHECK(cudaSetDevice(ix->device));
  SearchPlan pl; size_t off_slot;
  int rc = range_plan(ix, B, &pl, &off_slot);
  if (rc) return rc;
  if (!workspace || workspace_bytes < pl.off_end || (reinterpret_cast<uintptr_t>(workspace) & 255)) {
    set_error("range workspace too small or misaligned: need %zu bytes, 256-byte aligned", pl.off_end);
    return B200R_ENOMEM;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if ((rc = range_pass(ix, Q, B, min_score, id_gt, 1, nullptr, nullptr, pl, workspace, st))) return rc;
  int64_t* slot_off = reinterpret_cast<int64_t*>((char*)workspace + off_slot);
  if ((rc = launch_range_offsets(pl, workspace, out_offsets, slot_off, st))) return rc;
  B200R_CUDA_CHECK(cudaMemcpyAsync(total_host, out_offsets + B, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  B200R_CUDA_CHECK(cudaStreamSynchronize(st));
  return B200R_OK;
}

int b200r_range_fill(uint64_t index, const void* Q, int32_t B, float min_score, const int64_t* id_gt, int64_t total,
                     int64_t* out_ids, float* out_scores, void* workspace, size_t workspace_bytes, void* stream) {
  Index* ix = get_index(index);
  if (!ix) return B200R_ESTATE;
  if (!Q || B <= 0 || total < 0 || (total > 0 && (!out_ids || !out_scores))) { set_error("null pointer / bad B / total"); return B200R_EINVAL; }
  B200R_CUDA_CHECK(cudaSetDevice(ix->device));
  SearchPlan pl; size_t off_slot;
  int rc = range_plan(ix, B, &pl, &off_slot);
  if (rc) return rc;
  if (!workspace || workspace_bytes < pl.off_end) { set_error("range workspace too small: need %zu bytes", pl.off_end); return B200R_ENOMEM; }
  if (total == 0) return B200R_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t* slot_off = reinterpret_cast<const int64_t*>((char*)workspace + off_slot);      // left there by b200r_range_count
  if ((rc = range_pass(ix, Q, B, min_score, id_gt, 2, slot_off, out_ids, pl, workspace, st))) return rc;
  return launch_range_unpack(nullptr, total, nullptr, B, 0.f, ix->id_base, out_ids, out_scores, st);
}

int b200r_mmr_diversify(uint64_t index, const int64_t* cand_ids, const float* cand_scores, int32_t B, int32_t P, int32_t top_k,
                        float lambda_weight, int64_t* out_ids, float* out_scores, int32_t* out_pos, void* stream) {
  Index* ix = get_index(index);
  if (!ix) return B200R_ESTATE;
  if (!cand_ids || !cand_scores || !out_ids || !out_scores || !out_pos) { set_error("null pointer"); return B200R_EINVAL; }
  if (B <= 0 || P <= 0 || top_k <= 0 || P > 16384) { set_error("bad B / P / top_k (%d / %d / %d)", B, P, top_k); return B200R_EINVAL; }
  B200R_CUDA_CHECK(cudaSetDevice(ix->device));
  return launch_mmr(ix, cand_ids, cand_scores, B, P, top_k, lambda_weight, out_ids, out_scores, out_pos, reinterpret_cast<cudaStream_t>(stream));
}

int b200r_merge_topk(int device, const float* cand_scores, const int64_t* cand_ids, int32_t G, int32_t B, int32_t kin,
                     int32_t k, float* out_scores, int64_t* out_ids, void* stream) {
  if (!cand_scores || !cand_ids || !out_scores || !out_ids) { set_error("null pointer"); return B200R_EINVAL; }
  if (G <= 0 || B <= 0 || kin <= 0 || k <= 0) { set_error("G, B, kin, k must be > 0"); return B200R_EINVAL; }
  B200R_CUDA_CHECK(cudaSetDevice(device));
  return launch_merge_topk(cand_scores, cand_ids, (size_t)B * kin, (size_t)B * kin, G, B, kin, k, out_scores, out_ids,
                           reinterpret_cast<cudaStream_t>(stream));
}

int b200r_merge_topk_src(int device, const float* cand_scores, size_t shard_stride_scores, const int64_t* cand_ids,
                         size_t shard_stride_ids, int32_t G, int32_t B, int32_t kin, int32_t k, float* out_scores, int64_t* out_ids,
                         int32_t* out_src, void* stream) {
  if (!cand_scores || !cand_ids || !out_scores || !out_ids || !out_src) { set_error("null pointer"); return B200R_EINVAL; }
  if (G <= 0 || B <= 0 || kin <= 0 || k <= 0) { set_error("G, B, kin, k must be > 0"); return B200R_EINVAL; }
  if (shard_stride_scores < (size_t)B * kin || shard_stride_ids < (size_t)B * kin) { set_error("shard strides smaller than B*kin"); return B200R_EINVAL; }
  B200R_CUDA_CHECK(cudaSetDevice(device));
  return launch_merge_topk(cand_scores, cand_ids, shard_stride_scores, shard_stride_ids, G, B, kin, k, out_scores, out_ids,
                           reinterpret_cast<cudaStream_t>(stream), out_src);
}

int b200r_merge_topk_packed(int device, const void* packed, int32_t G, int32_t B, int32_t kin, int32_t k, float* out_scores,
                            int64_t* out_ids, void* stream) {
  if (!packed || !out_scores || !out_ids) { set_error("null pointer"); return B200R_EINVAL; }
  if (G <= 0 || B <= 0 || kin <= 0 || k <= 0) { set_error("G, B, kin, k must be > 0"); return B200R_EINVAL; }
  B200R_CUDA_CHECK(cudaSetDevice(device));
  const size_t n = (size_t)B * kin;
  const int64_t* ids = reinterpret_cast<const int64_t*>(packed);
  const float* scores = reinterpret_cast<const float*>(reinterpret_cast<const char*>(packed) + n * 8);
  // one record = n*8 + n*4 bytes: 1.5*n int64 elements, 3*n float elements
  if (n % 2) { set_error("merge_topk_packed needs B*kin even (record stride must be a whole number of int64)"); return B200R_EINVAL; }
  return launch_merge_topk(scores, ids, 3 * n, n + n / 2, G, B, kin, k, out_scores, out_ids, reinterpret_cast<cudaStream_t>(stream));
}

int b200r_profile_dump(char* buf, size_t n) {
  // {"name": [total_ms, launches], ...} in first-seen order; clears the records
  std::vector<ProfRec> recs;
  { std::lock_guard<std::mutex> lk(g_prof_mu); recs.swap(g_prof_recs); }
  std::vector<const char*> names; std::vector<double> ms; std::vector<int> cnt;
  for (const ProfRec& r : recs) {
    float t = 0.f;
    if (cudaEventSynchronize(r.b) == cudaSuccess) cudaEventElapsedTime(&t, r.a, r.b);
    size_t i = 0;
    for (; i < names.size(); ++i) if (strcmp(names[i], r.name) == 0) break;
    if (i == names.size()) { names.push_back(r.name); ms.push_back(0.0); cnt.push_back(0); }
    ms[i] += t; cnt[i] += 1;
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  std::string out = "{";
  for (size_t i = 0; i < names.size(); ++i) {
    char tmp[160];
    snprintf(tmp, sizeof(tmp), "%s\"%s\": [%.6f, %d]", i ? ", " : "", names[i], ms[i], cnt[i]);
    out += tmp;
  }
  out += "}";
  if (buf && n) { strncpy(buf, out.c_str(), n - 1); buf[n - 1] = 0; }
  return (int)out.size();
}

int b200r_profile_enable(int on) {
  g_prof_all.store(on == 2);
  g_prof_on.store(on == 1);
  std::vector<ProfRec> recs;
  std::vector<ProfPair> pairs;
  {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (on != 2) recs.swap(g_prof_recs);
    if (on != 1) pairs.swap(g_prof_events);
  }
  for (const ProfRec& r : recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (const ProfPair& e : pairs) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  return B200R_OK;
}

int b200r_profile_read(double* kernel_ms, int64_t* launches) {
  std::vector<ProfPair> pairs;
  { std::lock_guard<std::mutex> lk(g_prof_mu); pairs.swap(g_prof_events); }
  double total = 0.0;
  int64_t n = 0;
  int rc = B200R_OK;
  for (const ProfPair& e : pairs) {
    float ms = 0.f;
    if (rc == B200R_OK && (cudaEventSynchronize(e.b) != cudaSuccess || cudaEventElapsedTime(&ms, e.a, e.b) != cudaSuccess)) {
      set_error("profile_read: %s", cudaGetErrorString(cudaGetLastError()));
      rc = B200R_ECUDA;
    }
    total += ms;
    ++n;
    cudaEventDestroy(e.a); cudaEventDestroy(e.b);
  }
  if (kernel_ms) *kernel_ms = total;
  if (launches) *launches = n;
  return rc;
}

static int graph_create(int device, const int64_t* row_ptr, const int32_t* col_idx, int64_t V, int64_t E, const int64_t* node_rows,
                        uint64_t* out_graph) {
  if (!out_graph) { set_error("out_graph is null"); return B200R_EINVAL; }
  *out_graph = 0;
  if (!row_ptr || V <= 0 || E < 0 || (E > 0 && !col_idx)) { set_error("bad graph arguments"); return B200R_EINVAL; }
  if (V > 0xfffffffell) { set_error("V too large"); return B200R_EINVAL; }
  B200R_CUDA_CHECK(cudaSetDevice(device));
  int rc = validate_csr(reinterpret_cast<const long long*>(row_ptr), col_idx, V, E);
  if (rc) return rc;
  Graph* g = new (std::nothrow) Graph();
  if (!g) { set_error("out of host memory"); return B200R_ENOMEM; }
  g->magic = kGraphMagic; g->device = device; g->row_ptr = row_ptr; g->col_idx = col_idx; g->node_rows = node_rows; g->V = V; g->E = E;
  *out_graph = reinterpret_cast<uint64_t>(g);
  return B200R_OK;
}

int b200r_graph_create(int device, const int64_t* row_ptr, const int32_t* col_idx, int64_t V, int64_t E, uint64_t* out_graph) {
  return graph_create(device, row_ptr, col_idx, V, E, nullptr, out_graph);
}

int b200r_graph_create_mapped(int device, const int64_t* row_ptr, const int32_t* col_idx, int64_t V, int64_t E,
                              const int64_t* node_rows, uint64_t* out_graph) {
  if (!node_rows) { set_error("node_rows is null"); return B200R_EINVAL; }
  return graph_create(device, row_ptr, col_idx, V, E, node_rows, out_graph);
}

int b200r_graph_destroy(uint64_t graph) {
  Graph* g = get_graph(graph);
  if (!g) return B200R_ESTATE;
  g->magic = 0;
  delete g;
  return B200R_OK;
}

int32_t b200r_expand_max_candidates(void) { return expand_max_candidates(); }

int b200r_expand_rescore(uint64_t index, uint64_t graph, const void* Q, const int64_t* seeds, int32_t B, int32_t S,
                         int32_t window, float alpha_dense, float alpha_graph, int32_t k, void* workspace, size_t workspace_bytes,
                         float* out_scores, int64_t* out_ids, int32_t* out_hops, float* out_graph_norm, float* out_dense_norm,
                         float* out_cos, int32_t* out_ncand, int32_t* out_status, void* stream) {
  Index* ix = get_index(index);
  Graph* g = get_graph(graph);
  if (!ix || !g) return B200R_ESTATE;
  if (ix->device != g->device) { set_error("index and graph live on different devices"); return B200R_EINVAL; }
  if (!Q || !seeds || !workspace || !out_scores || !out_ids || !out_hops || !out_graph_norm || !out_dense_norm || !out_cos || !out_ncand || !out_status) {
    set_error("null pointer");
    return B200R_EINVAL;
  }
  if (B <= 0 || S <= 0 || k <= 0 || k > B200R_MAX_K || window < 0 || window > 255) {
    set_error("bad B/S/k/window (%d/%d/%d/%d)", B, S, k, window);
    return B200R_EINVAL;
  }
  if (workspace_bytes < expand_workspace_bytes(B) || (reinterpret_cast<uintptr_t>(workspace) & 255)) {
    set_error("expand workspace too small or misaligned: need %zu bytes, 256-byte aligned", expand_workspace_bytes(B));
    return B200R_ENOMEM;
  }
  B200R_CUDA_CHECK(cudaSetDevice(ix->device));
  return launch_expand_rescore(ix, g, Q, seeds, B, S, window, alpha_dense, alpha_graph, k, out_scores, out_ids, out_hops,
                               out_graph_norm, out_dense_norm, out_cos, out_ncand, out_status, 0, INT64_MIN, INT64_MAX, workspace,
                               nullptr, reinterpret_cast<cudaStream_t>(stream));
}

size_t b200r_expand_workspace_bytes(int32_t B) { return B > 0 ? expand_workspace_bytes(B) : 0; }

size_t b200r_expand_big_workspace_bytes(uint64_t graph) {
  Graph* g = get_graph(graph);
  return g ? expand_big_workspace_bytes(g->V) : 0;
}

int b200r_expand_rescore_big(uint64_t index, uint64_t graph, const void* Q, const int64_t* seeds, int32_t S, int32_t window,
                             float alpha_dense, float alpha_graph, int32_t k, void* workspace, size_t workspace_bytes, float* out_scores,
                             int64_t* out_ids, int32_t* out_hops, float* out_graph_norm, float* out_dense_norm, float* out_cos,
                             int32_t* out_ncand, void* stream) {
  Index* ix = get_index(index);
  Graph* g = get_graph(graph);
  if (!ix || !g) return B200R_ESTATE;
  if (ix->device != g->device) { set_error("index and graph live on different devices"); return B200R_EINVAL; }
  if (!Q || !seeds || !workspace || !out_scores || !out_ids || !out_hops || !out_graph_norm || !out_dense_norm || !out_cos || !out_ncand) {
    set_error("null pointer");
    return B200R_EINVAL;
  }
  if (S <= 0 || k <= 0 || k > B200R_MAX_K || window < 0 || window > 255) { set_error("bad S/k/window (%d/%d/%d)", S, k, window); return B200R_EINVAL; }
  if (g->V > 0x7fffffffll) { set_error("unbounded frontier path: V must be < 2^31"); return B200R_EINVAL; }
  if (workspace_bytes < expand_big_workspace_bytes(g->V) || (reinterpret_cast<uintptr_t>(workspace) & 255)) {
    set_error("workspace too small or misaligned: need %zu bytes, 256-byte aligned", expand_big_workspace_bytes(g->V));
    return B200R_ENOMEM;
  }
  B200R_CUDA_CHECK(cudaSetDevice(ix->device));
  return launch_expand_rescore_big(ix, g, Q, seeds, S, window, alpha_dense, alpha_graph, k, out_scores, out_ids, out_hops, out_graph_norm,
                                   out_dense_norm, out_cos, out_ncand, workspace, reinterpret_cast<cudaStream_t>(stream));
}

int b200r_expand_rescore_exchange(uint64_t index, uint64_t graph, uint64_t comm, const void* Q, const int64_t* seeds, int32_t B,
                                  int32_t S, int32_t window, float alpha_dense, float alpha_graph, int32_t k, int64_t n_total_rows,
                                  int32_t phase, void* workspace, size_t workspace_bytes, float* out_scores, int64_t* out_ids,
                                  int32_t* out_hops, float* out_graph_norm, float* out_dense_norm, float* out_cos, int32_t* out_ncand,
                                  int32_t* out_status, void* stream) {
  Index* ix = get_index(index);
  Graph* g = get_graph(graph);
  Comm* c = get_comm(comm);
  if (!ix || !g || !c) return B200R_ESTATE;
  if (ix->device != g->device || ix->device != c->device) { set_error("index, graph and comm must live on the same device"); return B200R_EINVAL; }
  if (!Q || !seeds || !workspace || !out_scores || !out_ids || !out_hops || !out_graph_norm || !out_dense_norm || !out_cos || !out_ncand ||
      !out_status) { set_error("null pointer"); return B200R_EINVAL; }
  if (B <= 0 || S <= 0 || k <= 0 || k > B200R_MAX_K || window < 0 || window > 255 || phase < 0 || phase > 4) {
    set_error("bad B/S/k/window/phase (%d/%d/%d/%d/%d)", B, S, k, window, phase);
    return B200R_EINVAL;
  }
  if (workspace_bytes < expand_workspace_bytes(B) || (reinterpret_cast<uintptr_t>(workspace) & 255)) {
    set_error("expand workspace too small or misaligned: need %zu bytes, 256-byte aligned", expand_workspace_bytes(B));
    return B200R_ENOMEM;
  }
  B200R_CUDA_CHECK(cudaSetDevice(ix->device));
  for (int ph = (phase ? phase : 1); ph <= (phase ? phase : 4); ++ph) {
    int rc = launch_expand_rescore_exchange(ix, g, c, Q, seeds, B, S, window, alpha_dense, alpha_graph, k, n_total_rows, out_scores, out_ids,
                                            out_hops, out_graph_norm, out_dense_norm, out_cos, out_ncand, out_status, workspace, ph,
                                            reinterpret_cast<cudaStream_t>(stream));
    if (rc) return rc;
  }
  return B200R_OK;
}

int b200r_expand_rescore_sharded(uint64_t index, uint64_t graph, const void* Q, const int64_t* seeds, int32_t B, int32_t S,
                                 int32_t window, float alpha_dense, float alpha_graph, int32_t k, int32_t pass, int64_t own_lo,
                                 int64_t own_hi, void* workspace, size_t workspace_bytes, float* minmax, float* out_scores,
                                 int64_t* out_ids, int32_t* out_hops, float* out_graph_norm, float* out_dense_norm, float* out_cos,
                                 int32_t* out_ncand, int32_t* out_status, void* stream) {
  Index* ix = get_index(index);
  Graph* g = get_graph(graph);
  if (!ix || !g) return B200R_ESTATE;
  if (pass != 1 && pass != 2) { set_error("pass must be 1 or 2"); return B200R_EINVAL; }
  if (!Q || !seeds || !workspace || !minmax || !out_scores || !out_ids || !out_hops || !out_graph_norm || !out_dense_norm ||
      !out_cos || !out_ncand || !out_status) { set_error("null pointer"); return B200R_EINVAL; }
  if (B <= 0 || S <= 0 || k <= 0 || k > B200R_MAX_K || window < 0 || window > 255) { set_error("bad B/S/k/window"); return B200R_EINVAL; }
  if (workspace_bytes < expand_workspace_bytes(B) || (reinterpret_cast<uintptr_t>(workspace) & 255)) {
    set_error("expand workspace too small or misaligned: need %zu bytes, 256-byte aligned", expand_workspace_bytes(B));
    return B200R_ENOMEM;
  }
  B200R_CUDA_CHECK(cudaSetDevice(ix->device));
  return launch_expand_rescore(ix, g, Q, seeds, B, S, window, alpha_dense, alpha_graph, k, out_scores, out_ids, out_hops,
                               out_graph_norm, out_dense_norm, out_cos, out_ncand, out_status, pass, own_lo, own_hi, workspace,
                               minmax, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
