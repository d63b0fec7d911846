// csr_expand.cu -- K3: graph-enhanced second stage, split so that its HBM-bound part streams on its own.
//
//   K3a bfs_kernel           one CTA (256 threads) per query: level-synchronous BFS from the seed passages over the
//                            CSR adjacency, depth <= window.  `visited` is a shared-memory hash set (first arrival ==
//                            minimum hop distance).  A level is expanded edge-parallel: the adjacency ranges of the
//                            whole level are read at once, a block scan turns degrees into edge offsets and the
//                            flattened edge list is spread over all threads (coalesced col_idx reads, several in
//                            flight per thread; hubs need no special case), with a warp-level __match_any_sync
//                            pre-dedup so each distinct neighbour costs one atomic per warp.  The candidate list, the
//                            hop distances and the query's inverse norm are parked in the workspace -- the candidates
//                            this rank owns first -- and the query reserves a contiguous range of 4-row groups in a
//                            flat work list (one 64-bit atomic).
//   K3b rescore_gather_kernel persistent, grid-wide: every warp owns a contiguous slice of the flat group list (so the
//                            load is balanced whatever the per-query candidate counts are), keeps the current query
//                            as fp32 in its own shared-memory slice and gathers 4 candidate rows at a time with all
//                            their 128-bit loads in flight.  This is the HBM-bound part (random 1.5 KB rows;
//                            scripts/gather_bw.cu measures 6.9 TB/s for the bare pattern).  Per-query cosine min/max
//                            are folded in with two order-preserving integer atomics per (warp, query).
//   K3c fuse_select_kernel   one CTA per query: min/max normalisation of both channels (retrieval_backend.py:296-301),
//                            fused = a_g*minmax(decay(hop)) + a_d*minmax(cos), exact top-k by (fused desc, id asc):
//                            radix select of the k-th key in shared memory (stops as soon as a digit bucket is wholly
//                            inside the top-k) + bitonic sort of the k winners.
//
// Candidate capacity: K3a/K3c run a 4096-candidate instantiation (4 CTAs/SM) for every query and a persistent
// 8192-candidate one that only picks up the queries that overflowed the first; beyond that B200R_EOVERFLOW.
//
// Replaces expand_qmatch_neighbors (reference app/modules/retrieval/graph_utils.py:58-129: BFS :109-127,
// decay :87-94), DenseReranker._cosine (retrieval_backend.py:192-197), _minmax_norm (:296-301) and the
// fusion / sort / slice of HybridRetrievalBackend.run (:353-372).
#include "common.cuh"
#include "comm.cuh"
#include "internal.h"
#include <math_constants.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace b200r {

constexpr int EXP_MAXC_SMALL = 4096;
constexpr int EXP_MAXC_BIG = 8192;    // row stride of the parked per-query arrays
constexpr int BFS_THREADS = 256;
constexpr int BFS_WARPS = BFS_THREADS / 32;
constexpr int EXP_FB = 2 * BFS_THREADS;   // frontier nodes expanded per batch (2 per thread)
constexpr int EXP_EU = 4;                 // col_idx loads in flight per thread
constexpr int GATHER_ROWS = 4;            // candidate rows a warp keeps in flight (2: 0.89x, 8: 0.68x in gather_bw.cu)
constexpr int GATHER_MAX_WARPS = 16;
constexpr int GATHER_CTAS_PER_SM = 2;
constexpr int SEL_THREADS = 256;
constexpr int SEL_WARPS = SEL_THREADS / 32;
static int exp_persistent_ctas() { return 2 * num_sms(); }
constexpr u32 EXP_EMPTY = 0xffffffffu;
constexpr int kStatusRetry = 1;       // internal: overflowed the small instantiation, the big one redoes it

// per-call scratch, carved out of the caller's workspace (b200r_expand_workspace_bytes)
struct ExpandWs {
  u32* cand;            // [B][EXP_MAXC_BIG] node ids in BFS order
  float* cos;           // [B][EXP_MAXC_BIG] cosine of the candidates this rank owns (0 elsewhere)
  uint8_t* dist;        // [B][EXP_MAXC_BIG] hop distance
  int* ncand;           // [B] (0 for a query that overflowed)
  int* nown;            // [B] candidates this rank owns: they are parked FIRST in cand/dist, so the gather reads a dense prefix
  float* qinv;          // [B] 1/|q| (cosine) or 1
  u32* mm;              // [B][2] order-preserving images of the owned candidates' (min, max) cosine
  u32* seg_start;       // [2B] first flat group of the i-th reservation (monotonic in i)
  u32* seg_q;           // [2B] its query
  u64* counter;         // reservations << 32 | groups reserved
};

size_t expand_workspace_bytes(int B) {
  const size_t b = (size_t)B;
  return align_up(b * EXP_MAXC_BIG * 4, 256) * 2 + align_up(b * EXP_MAXC_BIG, 256) + align_up(b * 4, 256) * 3 + align_up(b * 8, 256) * 3 + 256;
}

static ExpandWs carve_ws(void* ws, int B) {
  const size_t b = (size_t)B;
  char* w = reinterpret_cast<char*>(ws);
  ExpandWs s;
  s.cand = reinterpret_cast<u32*>(w); w += align_up(b * EXP_MAXC_BIG * 4, 256);
  s.cos = reinterpret_cast<float*>(w); w += align_up(b * EXP_MAXC_BIG * 4, 256);
  s.dist = reinterpret_cast<uint8_t*>(w); w += align_up(b * EXP_MAXC_BIG, 256);
  s.ncand = reinterpret_cast<int*>(w); w += align_up(b * 4, 256);
  s.nown = reinterpret_cast<int*>(w); w += align_up(b * 4, 256);
  s.qinv = reinterpret_cast<float*>(w); w += align_up(b * 4, 256);
  s.mm = reinterpret_cast<u32*>(w); w += align_up(b * 8, 256);
  s.seg_start = reinterpret_cast<u32*>(w); w += align_up(b * 8, 256);
  s.seg_q = reinterpret_cast<u32*>(w); w += align_up(b * 8, 256);
  s.counter = reinterpret_cast<u64*>(w);
  return s;
}

struct ExpandParams {
  const void* X;
  const void* Q;
  const float* row_mult;
  const long long* row_ptr;
  const int* col_idx;
  const long long* node_rows;     // nullable: graph node -> global corpus row (per-graph_id graphs keep LOCAL node ids)
  const long long* seeds;
  long long N, V, id_base;
  int D, B, S, window, k, kp2, metric;
  float alpha_dense, alpha_graph;
  float* out_scores; long long* out_ids; int* out_hops; float* out_gn; float* out_dn; float* out_cos;
  int* out_ncand; int* out_status;
  // row-sharded corpora (one rank per GPU): mode 0 = everything in one call; 1 = BFS + re-score of the candidates this
  // rank owns + their (min, max) cosine; 2 = fuse with the all-reduced (min, max) and rank the owned candidates
  int mode;
  long long own_lo, own_hi;       // global node ids this rank scores: [own_lo, own_hi)
  float* minmax;                  // [B][2] float (min, max): written by mode 1, read by mode 2
  ExpandWs ws;
  // mode 3 (b200r_expand_rescore_exchange): queries are sharded over the ranks as well.  The HOME rank of a query runs its
  // BFS and its fusion / top-k; the candidates travel to the ranks that OWN their rows (c_req), the cosines travel back
  // (c_cos), and every home rank stores its queries' results into all ranks' output slots (c_out) -- all by peer stores
  // from inside the kernels (comm.cuh), no collective.
  CommDev c_req, c_cos, c_out;
  int per;                        // queries per home rank: home(q) = q / per
  int q_lo, q_hi;                 // this rank's home queries
  long long rows_per;             // corpus rows per owner rank: owner(row) = min(row / rows_per, world - 1)
};

// ---- layout of the exchange slots (mode 3) ----------------------------------------------------------------------
// c_req slot (in the OWNER's window): one record per sending (home) rank: hdr[per][2] u32 (count, offset of this owner's
// sublist inside the home's candidate list) followed by ids[per][EXP_MAXC_BIG] u32 (corpus rows).
__host__ __device__ inline size_t xreq_record_bytes(int per) { return (size_t)per * 8 + (size_t)per * EXP_MAXC_BIG * 4; }
__device__ __forceinline__ u32* xreq_hdr(const CommDev& c, int owner, u32 E, int sender, int lq, int per) {
  return reinterpret_cast<u32*>(comm_slot(c, owner, E) + (size_t)sender * xreq_record_bytes(per)) + 2 * lq;
}
__device__ __forceinline__ u32* xreq_ids(const CommDev& c, int owner, u32 E, int sender, int lq, int per) {
  return reinterpret_cast<u32*>(comm_slot(c, owner, E) + (size_t)sender * xreq_record_bytes(per) + (size_t)per * 8) + (size_t)lq * EXP_MAXC_BIG;
}
// c_cos slot (in the HOME's window): cos[per][EXP_MAXC_BIG] f32, same positions as the home's parked candidate list
__device__ __forceinline__ float* xcos(const CommDev& c, int home, u32 E, int lq) {
  return reinterpret_cast<float*>(comm_slot(c, home, E)) + (size_t)lq * EXP_MAXC_BIG;
}
// c_out slot (in EVERY window): ids i64 [B][k] | score f32 | hop i32 | graph_norm f32 | dense_norm f32 | cos f32 | ncand i32 [B] | status i32 [B]
__host__ __device__ inline size_t xout_bytes(int B, int k) { return (size_t)28 * B * k + (size_t)8 * B; }
struct XOut { long long* ids; float* score; int* hop; float* gn; float* dn; float* cos; int* ncand; int* status; };
__device__ __forceinline__ XOut xout(char* base, int B, int k) {
  const size_t n = (size_t)B * k;
  XOut o;
  o.ids = reinterpret_cast<long long*>(base); o.score = reinterpret_cast<float*>(base + 8 * n);
  o.hop = reinterpret_cast<int*>(base + 12 * n); o.gn = reinterpret_cast<float*>(base + 16 * n);
  o.dn = reinterpret_cast<float*>(base + 20 * n); o.cos = reinterpret_cast<float*>(base + 24 * n);
  o.ncand = reinterpret_cast<int*>(base + 28 * n); o.status = o.ncand + B;
  return o;
}
__device__ __forceinline__ int row_owner(const ExpandParams& p, u32 row) {
  const long long o = (long long)row / p.rows_per;
  return (int)(o < p.c_req.world ? o : p.c_req.world - 1);
}

__device__ __forceinline__ float hop_decay(int d) {
  // graph_utils.py:87-94
  if (d == 0) return 1.0f;
  if (d == 1) return 0.7f;
  if (d == 2) return 0.5f;
  return fmaxf(0.5f - 0.1f * (float)(d - 2), 0.1f);
}

// insert `v` into the visited set; returns true if this call inserted it
template <int HASH>
__device__ __forceinline__ bool hash_insert(u32* table, u32 v) {
  u32 h = (v * 2654435761u) & (HASH - 1);
  while (true) {
    const u32 prev = atomicCAS(&table[h], EXP_EMPTY, v);
    if (prev == EXP_EMPTY) return true;
    if (prev == v) return false;
    h = (h + 1) & (HASH - 1);
  }
}

__device__ __forceinline__ void fill_no_result(const ExpandParams& p, int q) {
  for (int i = threadIdx.x; i < p.k; i += blockDim.x) {
    const size_t o = (size_t)q * p.k + i;
    p.out_scores[o] = -CUDART_INF_F; p.out_ids[o] = -1; p.out_hops[o] = -1; p.out_gn[o] = 0; p.out_dn[o] = 0; p.out_cos[o] = 0;
  }
}

// ------------------------------------------------------------------------------------------------ K3a
static size_t bfs_smem_bytes(int maxc) { return (size_t)maxc * 8 + (size_t)maxc * 4 + (size_t)EXP_FB * 8 + (size_t)(EXP_FB + 1) * 4; }

template <typename QT, int MAXC>
__global__ void __launch_bounds__(BFS_THREADS) bfs_kernel(const ExpandParams p) {
  constexpr bool RETRY_PASS = (MAXC == EXP_MAXC_BIG);
  constexpr int HASH = 2 * MAXC;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  u32* table = reinterpret_cast<u32*>(smem_raw);                                  // [HASH]
  u32* cand = table + HASH;                                                       // [MAXC] node ids in BFS order
  long long* f_e0 = reinterpret_cast<long long*>(cand + MAXC);                    // [EXP_FB] first edge of a frontier node
  u32* f_pre = reinterpret_cast<u32*>(f_e0 + EXP_FB);                             // [EXP_FB + 1] exclusive edge offsets
  __shared__ int s_count, s_overflow, s_nown, s_nother, s_nlevels;
  __shared__ int s_level_end[256];        // s_level_end[d] = candidates at hop distance <= d (window <= 255)
  __shared__ u32 s_wsum[BFS_WARPS];
  __shared__ unsigned long long s_wwide[BFS_WARPS];
  __shared__ double s_qn[BFS_WARPS];
  __shared__ int s_ocnt[B200R_MAX_PEERS], s_ooff[B200R_MAX_PEERS + 1], s_ofill[B200R_MAX_PEERS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool xchg = p.mode == 3;
  u32 epoch = 0;
  if (xchg) epoch = comm_send_begin(p.c_req);         // the owners have consumed the requests this slot held two exchanges ago
  const int q_first = xchg ? p.q_lo : 0, q_last = xchg ? p.q_hi : p.B;

  for (int q = q_first + blockIdx.x; q < q_last; q += gridDim.x) {
    if (RETRY_PASS && p.out_status[q] != kStatusRetry) continue;      // block-uniform
    __syncthreads();                                                  // the previous query of this CTA is done with smem
    for (int i = tid; i < HASH; i += BFS_THREADS) table[i] = EXP_EMPTY;
    if (tid == 0) { s_count = 0; s_overflow = 0; s_nown = 0; s_nother = 0; s_nlevels = 1; }
    // the query's inverse norm (the re-score kernel multiplies by it)
    const QT* Qp = reinterpret_cast<const QT*>(p.Q) + (size_t)q * p.D;
    double qn = 0.0;
    for (int d = tid; d < p.D; d += BFS_THREADS) {
      const float v = q_to_float<QT>(Qp[d]);
      qn += (double)v * (double)v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) qn += __shfl_xor_sync(kFullMask, qn, o);
    if (lane == 0) s_qn[warp] = qn;
    __syncthreads();
    if (tid == 0) {
      double qn_tot = 0.0;
      for (int w = 0; w < BFS_WARPS; ++w) qn_tot += s_qn[w];
      p.ws.qinv[q] = (p.metric == B200R_COS) ? (qn_tot > 0.0 ? (float)(1.0 / sqrt(qn_tot)) : 0.0f) : 1.0f;
    }

    // ---- level 0: seeds ---------------------------------------------------------------------------
    for (int i = tid; i < p.S; i += BFS_THREADS) {
      const long long s = p.seeds[(size_t)q * p.S + i];   // node id == global corpus row id
      if (s >= 0 && s < p.V) {
        if (hash_insert<HASH>(table, (u32)s)) {
          const int pos = atomicAdd(&s_count, 1);
          if (pos < MAXC) cand[pos] = (u32)s; else s_overflow = 1;
        }
      }
    }
    __syncthreads();
    if (tid == 0) s_level_end[0] = min(s_count, MAXC);

    // ---- levels 1..window: edge-parallel expansion, EXP_FB frontier nodes per batch ------------------
    int level_begin = 0;
    for (int d = 1; d <= p.window; ++d) {
      const int level_end = min(s_count, MAXC);
      __syncthreads();                      // everyone has read s_count before the level appends to it
      if (level_end == level_begin || s_overflow) break;  // block-uniform
      for (int b0 = level_begin; b0 < level_end; b0 += EXP_FB) {
        const int nb = min(EXP_FB, level_end - b0);
        // (1) adjacency ranges, two consecutive frontier nodes per thread
        u32 deg[2] = {0u, 0u};
        unsigned long long wide = 0;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int f = 2 * tid + r;
          if (f < nb) {
            const u32 u = cand[b0 + f];
            const long long e0 = __ldg(p.row_ptr + u), e1 = __ldg(p.row_ptr + u + 1);
            const long long dl = e1 > e0 ? e1 - e0 : 0;
            f_e0[f] = e0;
            deg[r] = (u32)min(dl, (long long)0x7fffffff);
            wide += (unsigned long long)dl;
          }
        }
        // (2) block exclusive scan of the degrees (and a 64-bit total to catch a batch with >= 2^31 edges)
        u32 incl = deg[0] + deg[1];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const u32 t = __shfl_up_sync(kFullMask, incl, o);
          if (lane >= o) incl += t;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wide += __shfl_xor_sync(kFullMask, wide, o);
        if (lane == 31) s_wsum[warp] = incl;
        if (lane == 0) s_wwide[warp] = wide;
        __syncthreads();
        u32 base = 0, total = 0;
        unsigned long long total_wide = 0;
#pragma unroll
        for (int w = 0; w < BFS_WARPS; ++w) {
          const u32 ws = s_wsum[w];
          if (w < warp) base += ws;
          total += ws;
          total_wide += s_wwide[w];
        }
        if (total_wide > 0x7fffffffull) { if (tid == 0) s_overflow = 1; total = 0; }   // block-uniform
        const u32 excl = base + incl - (deg[0] + deg[1]);
        if (2 * tid < nb) f_pre[2 * tid] = excl;
        if (2 * tid + 1 < nb) f_pre[2 * tid + 1] = excl + deg[0];
        if (tid == 0) f_pre[nb] = total;
        __syncthreads();
        // (3) the batch's edges, flattened over the CTA
        for (u32 eb = 0; eb < total; eb += BFS_THREADS * EXP_EU) {
          u32 v[EXP_EU];
#pragma unroll
          for (int r = 0; r < EXP_EU; ++r) {
            const u32 e = eb + (u32)r * BFS_THREADS + (u32)tid;
            v[r] = EXP_EMPTY;
            if (e < total) {
              int lo = 0, hi = nb;              // largest node with f_pre[node] <= e
              while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (f_pre[mid] <= e) lo = mid; else hi = mid;
              }
              const int nb_id = __ldg(p.col_idx + f_e0[lo] + (long long)(e - f_pre[lo]));
              v[r] = (nb_id >= 0 && (long long)nb_id < p.V) ? (u32)nb_id : EXP_EMPTY;     // malformed CSR entries are dropped
            }
          }
#pragma unroll
          for (int r = 0; r < EXP_EU; ++r) {
            if (eb + (u32)r * BFS_THREADS >= total) break;        // block-uniform
            const u32 peers = __match_any_sync(kFullMask, v[r]);
            const bool leader = v[r] != EXP_EMPTY && (lane == __ffs(peers) - 1);
            if (leader && !*(volatile int*)&s_overflow && hash_insert<HASH>(table, v[r])) {
              const int pos = atomicAdd(&s_count, 1);
              if (pos < MAXC) cand[pos] = v[r]; else s_overflow = 1;
            }
          }
        }
        __syncthreads();                    // the scratch is reused by the next batch
      }
      // the nodes this level appended are at hop distance d
      if (tid == 0) { s_level_end[d] = min(s_count, MAXC); s_nlevels = d + 1; }
      level_begin = level_end;
    }
    __syncthreads();
    const int ncand = min(s_count, MAXC);
    if (s_overflow) {  // block-uniform
      if (tid == 0) {
        p.out_status[q] = (MAXC < EXP_MAXC_BIG) ? kStatusRetry : B200R_EOVERFLOW;
        p.out_ncand[q] = s_count;
        p.ws.ncand[q] = 0;
        p.ws.nown[q] = 0;
      }
      if (!xchg) fill_no_result(p, q);
      else if (MAXC == EXP_MAXC_BIG && tid < p.c_req.world) {       // final overflow: every owner learns "no candidates"
        u32* h = xreq_hdr(p.c_req, tid, epoch, p.c_req.rank, q - p.q_lo, p.per);
        h[0] = 0u; h[1] = 0u;
      }
      continue;
    }
    // park the list with the candidates this rank owns FIRST (the order inside a list carries no meaning: ranking breaks
    // ties by id), so the gather kernel reads a dense prefix however the corpus is sharded
    u32* cand_out = p.ws.cand + (size_t)q * EXP_MAXC_BIG;
    uint8_t* dist_out = p.ws.dist + (size_t)q * EXP_MAXC_BIG;
    const int nlevels = s_nlevels;
    if (xchg) {
      // exchange mode: the list is parked grouped by the rank that owns each candidate's row; every owner receives its
      // sublist (corpus rows) and where it sits in this list, so that its cosines come back to the same positions
      const int G = p.c_req.world, lq = q - p.q_lo;
      if (tid < G) { s_ocnt[tid] = 0; s_ofill[tid] = 0; }
      __syncthreads();
      for (int c = tid; c < ncand; c += BFS_THREADS) {
        u32 node = cand[c];
        if (p.node_rows) {
          const long long r = __ldg(p.node_rows + node);
          node = (r >= 0 && r < 0xffffffffll) ? (u32)r : 0xfffffffeu;
          cand[c] = node;                                           // from here on the list holds corpus rows
        }
        atomicAdd(&s_ocnt[row_owner(p, node)], 1);
      }
      __syncthreads();
      if (tid == 0) {
        int run = 0;
        for (int o = 0; o < G; ++o) { s_ooff[o] = run; run += s_ocnt[o]; }
        s_ooff[G] = run;
      }
      __syncthreads();
      for (int c = tid; c < ncand; c += BFS_THREADS) {
        const u32 row = cand[c];
        int d = 0;
        while (d + 1 < nlevels && c >= s_level_end[d]) ++d;
        const int o = row_owner(p, row);
        const int j = atomicAdd(&s_ofill[o], 1);
        cand_out[s_ooff[o] + j] = row;
        dist_out[s_ooff[o] + j] = (uint8_t)min(d, 255);
        xreq_ids(p.c_req, o, epoch, p.c_req.rank, lq, p.per)[j] = row;
      }
      if (tid < G) {
        u32* h = xreq_hdr(p.c_req, tid, epoch, p.c_req.rank, lq, p.per);
        h[0] = (u32)s_ocnt[tid]; h[1] = (u32)s_ooff[tid];
      }
      if (tid == 0) { p.ws.ncand[q] = ncand; p.ws.nown[q] = ncand; p.out_ncand[q] = ncand; p.out_status[q] = 0; }
      continue;
    }
    for (int c = tid; c < ncand; c += BFS_THREADS) {
      u32 node = cand[c];
      int d = 0;
      while (d + 1 < nlevels && c >= s_level_end[d]) ++d;
      if (p.node_rows) {            // mapped graph: report (and score) the corpus row behind the node; < 0 = no row
        const long long r = __ldg(p.node_rows + node);
        node = (r >= 0 && r < 0xffffffffll) ? (u32)r : 0xfffffffeu;
      }
      const bool own = (long long)node >= p.own_lo && (long long)node < p.own_hi;
      const int pos = own ? atomicAdd(&s_nown, 1) : ncand - 1 - atomicAdd(&s_nother, 1);
      cand_out[pos] = node;
      dist_out[pos] = (uint8_t)min(d, 255);
    }
    __syncthreads();
    if (tid == 0) {
      p.ws.ncand[q] = ncand;
      p.ws.nown[q] = s_nown;
      p.out_ncand[q] = ncand;
      p.out_status[q] = 0;
      p.ws.mm[2 * q] = 0xffffffffu;       // "no owned candidate yet": min image > max image
      p.ws.mm[2 * q + 1] = 0u;
      const u32 groups = (u32)((s_nown + GATHER_ROWS - 1) / GATHER_ROWS);
      const u64 old = atomicAdd(p.ws.counter, (1ull << 32) | (u64)groups);
      p.ws.seg_start[old >> 32] = (u32)old;
      p.ws.seg_q[old >> 32] = (u32)q;
    }
  }
  if (xchg) {
    // the small instantiation only makes its stores visible; the persistent big one (always launched after it) signals
    if (RETRY_PASS) comm_send_end(p.c_req, epoch, gridDim.x);
    else __threadfence_system();
  }
}

// ------------------------------------------------------------------------------------------------ K3b
template <typename T>
__global__ void __launch_bounds__(GATHER_MAX_WARPS * 32, GATHER_CTAS_PER_SM) rescore_gather_kernel(const ExpandParams p, const int dpad) {
  typedef typename QueryType<T>::type QT;
  constexpr int VEC = ElemTraits<T>::kVec;
  constexpr int ROWS = GATHER_ROWS;
  extern __shared__ __align__(16) float qs_all[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps_per_cta = blockDim.x >> 5;
  float* qs = qs_all + (size_t)warp * dpad;          // this warp's current query, fp32

  const bool xchg = p.mode == 3;
  u32 e_req = 0, e_cos = 0;
  if (xchg) { e_cos = comm_send_begin(p.c_cos); e_req = comm_recv_peek(p.c_req); }   // the plan kernel already waited for the requests
  const u64 packed = *p.ws.counter;
  const int nseg = (int)(packed >> 32);
  const u32 total = (u32)packed;
  const u64 W = (u64)gridDim.x * warps_per_cta, w = (u64)blockIdx.x * warps_per_cta + warp;
  const u32 g_lo = (u32)((u64)total * w / W), g_hi = (u32)((u64)total * (w + 1) / W);
  do {
  if (g_lo >= g_hi) break;

  // the reservation that contains group g_lo: the last one that starts at or before it (empty ones share a start)
  int seg = 0;
  {
    int lo = 0, hi = nseg;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(p.ws.seg_start + mid) <= g_lo) lo = mid; else hi = mid;
    }
    seg = lo;
  }
  u32 seg_begin = __ldg(p.ws.seg_start + seg);
  u32 seg_end = (seg + 1 < nseg) ? __ldg(p.ws.seg_start + seg + 1) : total;

  const uint8_t* Xb = reinterpret_cast<const uint8_t*>(p.X);
  const int chunks = p.D / VEC;
  const size_t row_bytes = (size_t)p.D * sizeof(T);
  int q = -1, ncand = 0;
  float qinv = 1.0f, cmin = CUDART_INF_F, cmax = -CUDART_INF_F;
  const u32* cand_src = nullptr;      // this query's candidate rows (parked list, or the request received from its home rank)
  float* cos_dst = nullptr;           // where their cosines go (workspace, or the home rank's window)

  auto flush_minmax = [&]() {
    if (xchg) return;                 // the home rank sees every cosine and takes the range itself
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      cmin = fminf(cmin, __shfl_xor_sync(kFullMask, cmin, o));
      cmax = fmaxf(cmax, __shfl_xor_sync(kFullMask, cmax, o));
    }
    if (lane == 0 && cmin <= cmax) {
      atomicMin(p.ws.mm + 2 * q, f32_to_ordered(cmin));
      atomicMax(p.ws.mm + 2 * q + 1, f32_to_ordered(cmax));
    }
    cmin = CUDART_INF_F; cmax = -CUDART_INF_F;
  };
  auto load_query = [&]() {
    q = (int)__ldg(p.ws.seg_q + seg);
    ncand = p.ws.nown[q];          // the owned candidates are the first nown entries of the parked list
    qinv = p.ws.qinv[q];
    if (xchg) {
      const int home = q / p.per, lq = q - home * p.per;
      cand_src = xreq_ids(p.c_req, p.c_req.rank, e_req, home, lq, p.per);
      cos_dst = xcos(p.c_cos, home, e_cos, lq) + xreq_hdr(p.c_req, p.c_req.rank, e_req, home, lq, p.per)[1];
    } else {
      cand_src = p.ws.cand + (size_t)q * EXP_MAXC_BIG;
      cos_dst = p.ws.cos + (size_t)q * EXP_MAXC_BIG;
    }
    const QT* Qp = reinterpret_cast<const QT*>(p.Q) + (size_t)q * p.D;
    __syncwarp();
    // float4 "part" layout: the 4-element part `e/4` of 16-byte chunk `c` lives at q4[(e/4)*chunks + c], so a warp's 32
    // consecutive chunks read 32 consecutive float4 (conflict-free LDS.128)
    for (int d = lane; d < p.D; d += 32) {
      const int c = d / VEC, e = d % VEC;
      qs[((size_t)(e >> 2) * chunks + c) * 4 + (e & 3)] = q_to_float<QT>(Qp[d]);
    }
    __syncwarp();
  };
  load_query();

  for (u32 g = g_lo; g < g_hi; ++g) {
    if (g >= seg_end) {                       // warp-uniform: next (non-empty) reservation
      flush_minmax();
      do {
        ++seg;
        seg_begin = seg_end;
        seg_end = (seg + 1 < nseg) ? __ldg(p.ws.seg_start + seg + 1) : total;
      } while (g >= seg_end);
      load_query();
    }
    const int c0 = (int)(g - seg_begin) * ROWS;
    // lane r < ROWS owns row r of the group: its node id, ownership and row multiplier (fetched with the rows, so its
    // latency is hidden behind them)
    long long my_row = 0;
    bool my_valid = false, my_have = false;
    float my_mult = 0.0f;
    if (lane < ROWS && c0 + lane < ncand) {
      const long long node = (long long)cand_src[c0 + lane];
      const long long row = node - p.id_base;
      my_valid = true;
      my_have = row >= 0 && row < p.N;       // a graph node that is not a corpus row scores 0.0
      if (my_have) { my_row = row; my_mult = p.row_mult ? __ldg(p.row_mult + row) : 1.0f; }
    }
    const uint8_t* rp[ROWS];
    float acc[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      rp[r] = Xb + (size_t)__shfl_sync(kFullMask, my_row, r) * row_bytes;     // (padding lanes of a query's last group read row 0)
      acc[r] = 0.0f;
    }
    for (int chunk = lane; chunk < chunks; chunk += 32) {
      uint4 xv[ROWS];
#pragma unroll
      for (int r = 0; r < ROWS; ++r) xv[r] = ld_stream_16B(rp[r] + (size_t)chunk * 16);
#pragma unroll
      for (int part = 0; part < VEC / 4; ++part) {
        const float4 qv = reinterpret_cast<const float4*>(qs)[(size_t)part * chunks + chunk];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
          float xf[4];
          ElemTraits<T>::unpack4(xv[r], part, xf);
          acc[r] = fmaf(xf[0], qv.x, acc[r]);
          acc[r] = fmaf(xf[1], qv.y, acc[r]);
          acc[r] = fmaf(xf[2], qv.z, acc[r]);
          acc[r] = fmaf(xf[3], qv.w, acc[r]);
        }
      }
    }
    float mine = 0.0f;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      float a = acc[r];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(kFullMask, a, o);
      if (lane == r) mine = a;
    }
    if (my_valid) {
      const float v = my_have ? (mine * my_mult) * qinv + 0.0f : 0.0f;
      cos_dst[c0 + lane] = v;
      cmin = fminf(cmin, v); cmax = fmaxf(cmax, v);
    }
  }
  flush_minmax();
  } while (0);
  if (xchg) {
    comm_recv_end(p.c_req, e_req, gridDim.x);       // the requests have been read: their slot may be reused
    comm_send_end(p.c_cos, e_cos, gridDim.x);       // every cosine of this rank is in its home's window
  }
}

// exchange mode, owner side: wait for every home rank's requests, then lay the flat work list of the gather out over ALL
// queries (query q's groups follow query q-1's) and note how many candidates of each query this rank owns.  One CTA.
__global__ void __launch_bounds__(1024) gather_plan_kernel(const ExpandParams p) {
  __shared__ u32 s_warp[32];
  __shared__ u32 s_base;
  const u32 E = comm_recv_begin(p.c_req);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int q0 = 0; q0 < p.B; q0 += 1024) {
    const int q = q0 + tid;
    u32 cnt = 0;
    if (q < p.B) {
      const int home = q / p.per, lq = q - home * p.per;
      cnt = xreq_hdr(p.c_req, p.c_req.rank, E, home, lq, p.per)[0];
      if (cnt > (u32)EXP_MAXC_BIG) cnt = 0;          // cannot happen; never index beyond a record
      p.ws.nown[q] = (int)cnt;
      if (p.metric != B200R_COS) p.ws.qinv[q] = 1.0f;
    }
    const u32 groups = (cnt + GATHER_ROWS - 1) / GATHER_ROWS;
    u32 incl = groups;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(kFullMask, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    u32 wbase = 0, tot = 0;
    for (int w = 0; w < 32; ++w) { if (w < warp) wbase += s_warp[w]; tot += s_warp[w]; }
    const u32 base = s_base;
    if (q < p.B) { p.ws.seg_start[q] = base + wbase + incl - groups; p.ws.seg_q[q] = (u32)q; }
    __syncthreads();
    if (tid == 0) s_base = base + tot;
    __syncthreads();
  }
  if (tid == 0) *p.ws.counter = ((u64)(u32)p.B << 32) | (u64)s_base;
}

// exchange mode, every rank: wait for all home ranks' results and hand them to the caller's arrays
__global__ void __launch_bounds__(256) collect_outputs_kernel(const ExpandParams p) {
  const u32 E = comm_recv_begin(p.c_out);
  const XOut o = xout(comm_slot(p.c_out, p.c_out.rank, E), p.B, p.k);
  const size_t n = (size_t)p.B * p.k, i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
  for (size_t i = i0; i < n; i += step) {
    p.out_ids[i] = o.ids[i]; p.out_scores[i] = o.score[i]; p.out_hops[i] = o.hop[i];
    p.out_gn[i] = o.gn[i]; p.out_dn[i] = o.dn[i]; p.out_cos[i] = o.cos[i];
  }
  for (size_t q = i0; q < (size_t)p.B; q += step) { p.out_ncand[q] = o.ncand[q]; p.out_status[q] = o.status[q]; }
  comm_recv_end(p.c_out, E, gridDim.x);
}

// mode 1: the host all-reduces float (min, max) pairs
__global__ void minmax_decode_kernel(const u32* __restrict__ mm, float* __restrict__ minmax, const int* __restrict__ status, int B) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= B) return;
  const bool any = status[q] == 0 && mm[2 * q] <= mm[2 * q + 1];
  minmax[2 * q] = any ? ordered_to_f32(mm[2 * q]) : CUDART_INF_F;
  minmax[2 * q + 1] = any ? ordered_to_f32(mm[2 * q + 1]) : -CUDART_INF_F;
}

// ------------------------------------------------------------------------------------------------ K3c
constexpr int FUSE_BINS = 1024;        // linear bins of the histogram select in fuse_select_kernel
static size_t select_smem_bytes(int maxc, int kp2) { return (size_t)maxc * 8 + (size_t)maxc * 4 + (size_t)kp2 * 8 + (size_t)kp2 * 4; }

template <int MAXC>
__global__ void __launch_bounds__(SEL_THREADS) fuse_select_kernel(const ExpandParams p) {
  constexpr bool BIG = (MAXC == EXP_MAXC_BIG);
  extern __shared__ __align__(16) uint8_t smem_raw[];
  u64* keys = reinterpret_cast<u64*>(smem_raw);                    // [MAXC]
  u64* sk = keys + MAXC;                                           // [kp2]
  u32* cand = reinterpret_cast<u32*>(sk + p.kp2);                  // [MAXC]
  int* wslot = reinterpret_cast<int*>(cand + MAXC);                // [kp2] position of every winner in the candidate list
  __shared__ int s_remaining, s_nout, s_done;
  __shared__ float s_red[4][SEL_WARPS];
  __shared__ __align__(16) u32 hist[FUSE_BINS];          // the radix passes use the first 256 words
  __shared__ u64 s_prefix;
  __shared__ u32 s_lo32, s_hi32, s_t32;                  // histogram select: range of the fused scores, threshold
  __shared__ int s_valid, s_bstar, s_reach;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool xchg = p.mode == 3;
  u32 e_cos = 0, e_out = 0;
  if (xchg) { e_cos = comm_recv_begin(p.c_cos); e_out = comm_send_begin(p.c_out); }   // every owner's cosines have arrived
  const int q_first = xchg ? p.q_lo : 0, q_last = xchg ? p.q_hi : p.B;

  for (int q = q_first + blockIdx.x; q < q_last; q += gridDim.x) {
    if (p.out_status[q] != 0) {                                    // overflow was reported by the BFS (block-uniform)
      if (xchg && !BIG) {                                          // ... and every rank must learn it
        const int st = p.out_status[q], nc = p.out_ncand[q];
        for (int g = 0; g < p.c_out.world; ++g) {
          const XOut o = xout(comm_slot(p.c_out, g, e_out), p.B, p.k);
          for (int i = tid; i < p.k; i += SEL_THREADS) {
            const size_t oi = (size_t)q * p.k + i;
            o.score[oi] = -CUDART_INF_F; o.ids[oi] = -1; o.hop[oi] = -1; o.gn[oi] = 0; o.dn[oi] = 0; o.cos[oi] = 0;
          }
          if (tid == 0) { o.ncand[q] = nc; o.status[q] = st; }
        }
      }
      continue;
    }
    const int ncand = p.ws.ncand[q];
    if (BIG != (ncand > EXP_MAXC_SMALL)) continue;                 // the other instantiation ranks this query
    __syncthreads();
    const u32* cand_in = p.ws.cand + (size_t)q * EXP_MAXC_BIG;
    const float* cos_in = xchg ? xcos(p.c_cos, p.c_cos.rank, e_cos, q - p.q_lo) : p.ws.cos + (size_t)q * EXP_MAXC_BIG;
    const uint8_t* dist_in = p.ws.dist + (size_t)q * EXP_MAXC_BIG;

    // ---- min/max of both channels (retrieval_backend.py:296-301) ----------------------------------
    float gmin = CUDART_INF_F, gmax = -CUDART_INF_F, xmin = CUDART_INF_F, xmax = -CUDART_INF_F;
    for (int c = tid; c < ncand; c += SEL_THREADS) {
      cand[c] = cand_in[c];
      const float gv = hop_decay(dist_in[c]);
      gmin = fminf(gmin, gv); gmax = fmaxf(gmax, gv);
      if (xchg) { const float cv = cos_in[c]; xmin = fminf(xmin, cv); xmax = fmaxf(xmax, cv); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      gmin = fminf(gmin, __shfl_xor_sync(kFullMask, gmin, o)); gmax = fmaxf(gmax, __shfl_xor_sync(kFullMask, gmax, o));
      xmin = fminf(xmin, __shfl_xor_sync(kFullMask, xmin, o)); xmax = fmaxf(xmax, __shfl_xor_sync(kFullMask, xmax, o));
    }
    if (lane == 0) { s_red[0][warp] = gmin; s_red[1][warp] = gmax; s_red[2][warp] = xmin; s_red[3][warp] = xmax; }
    if (tid == 0) { s_prefix = 0; s_remaining = p.k; s_nout = 0; s_done = 0; s_lo32 = 0xffffffffu; s_hi32 = 0u; s_t32 = 0xffffffffu; s_valid = 0; }
    for (int i = tid; i < p.kp2; i += SEL_THREADS) sk[i] = 0;
    __syncthreads();
    for (int w = 0; w < SEL_WARPS; ++w) {
      gmin = fminf(gmin, s_red[0][w]); gmax = fmaxf(gmax, s_red[1][w]);
      xmin = fminf(xmin, s_red[2][w]); xmax = fmaxf(xmax, s_red[3][w]);
    }
    float cmin, cmax;
    if (xchg) {
      cmin = xmin; cmax = xmax;                                    // this rank holds every candidate's cosine
    } else if (p.mode == 2) {
      cmin = p.minmax[2 * q]; cmax = p.minmax[2 * q + 1];          // all-reduced across the ranks
    } else {
      const u32 lo = p.ws.mm[2 * q], hi = p.ws.mm[2 * q + 1];
      cmin = lo <= hi ? ordered_to_f32(lo) : CUDART_INF_F;
      cmax = lo <= hi ? ordered_to_f32(hi) : -CUDART_INF_F;
    }
    const bool c_flat = !(cmax > cmin), g_flat = !(gmax > gmin);   // "vmax <= vmin -> all 0.0" (:300)
    const float c_den = c_flat ? 1.0f : (cmax - cmin), g_den = g_flat ? 1.0f : (gmax - gmin);

    // ---- fused keys (and their score range / count, for the histogram select) ------------------------------
    {
      u32 lo32 = 0xffffffffu, hi32 = 0u; int nv = 0;
      for (int c = tid; c < ncand; c += SEL_THREADS) {
        const float dn = c_flat ? 0.0f : (cos_in[c] - cmin) / c_den;
        const float gn = g_flat ? 0.0f : (hop_decay(dist_in[c]) - gmin) / g_den;
        const long long node = (long long)cand[c];
        const bool own = node >= p.own_lo && node < p.own_hi;
        const u64 key = own ? make_key(p.alpha_graph * gn + p.alpha_dense * dn, cand[c]) : 0ull;   // ties on the fused score -> lower node id
        keys[c] = key;
        if (key != 0) { lo32 = min(lo32, (u32)(key >> 32)); hi32 = max(hi32, (u32)(key >> 32)); ++nv; }
      }
      lo32 = __reduce_min_sync(kFullMask, lo32); hi32 = __reduce_max_sync(kFullMask, hi32); nv = __reduce_add_sync(kFullMask, nv);
      if (lane == 0 && nv) { atomicMin(&s_lo32, lo32); atomicMax(&s_hi32, hi32); atomicAdd(&s_valid, nv); }
    }
    __syncthreads();

    // ---- k-th key.  Histogram select (as select_merge.cu): ONE pass into 1 024 bins linear in the fused score, the bin b*
    // of the k-th best from a suffix scan, T = the smallest score in bins >= b*; the `reach` >= k keys with score >= T are
    // sorted below and the best k emitted.  Mass ties (reach > kp2) take the exact radix select. ----
    u64 kth = 0;
    bool have_kth = s_valid <= p.k;                 // every valid key wins (kth = 0)
    if (!have_kth) {
      for (int i = tid; i < FUSE_BINS; i += SEL_THREADS) hist[i] = 0;
      __syncthreads();
      const float flo = ordered_to_f32(s_lo32), fhi = ordered_to_f32(s_hi32);
      const float range = fhi - flo;
      const float scale = (range > 0.0f && range < CUDART_INF_F) ? (float)(FUSE_BINS - 1) / range : 0.0f;
      auto bin_of = [&](u64 key) -> int {
        const int b = (int)((key_score(key) - flo) * scale);
        return b < FUSE_BINS - 1 ? b : FUSE_BINS - 1;
      };
      for (int c = tid; c < ncand; c += SEL_THREADS) { const u64 key = keys[c]; if (key != 0) atomicAdd(&hist[bin_of(key)], 1u); }
      __syncthreads();
      if (tid < 32) {                               // lane L owns bins [32 L, 32 L + 32)
        u32 w[32]; u32 lane_sum = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint4 t = reinterpret_cast<const uint4*>(hist + lane * 32)[j];
          w[4 * j] = t.x; w[4 * j + 1] = t.y; w[4 * j + 2] = t.z; w[4 * j + 3] = t.w;
          lane_sum += t.x + t.y + t.z + t.w;
        }
        u32 suffix = lane_sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_down_sync(kFullMask, suffix, o); if (lane + o < 32) suffix += t; }
        const u32 above = suffix - lane_sum;
        if (above < (u32)p.k && (u32)p.k <= suffix) {
          u32 acc = above;
#pragma unroll
          for (int j = 31; j >= 0; --j) {
            acc += w[j];
            if (acc >= (u32)p.k) { s_bstar = lane * 32 + j; s_reach = (int)acc; break; }
          }
        }
      }
      __syncthreads();
      const int bstar = s_bstar;
      u32 t32 = 0xffffffffu;
      for (int c = tid; c < ncand; c += SEL_THREADS) {
        const u64 key = keys[c];
        if (key != 0 && bin_of(key) >= bstar) t32 = min(t32, (u32)(key >> 32));
      }
      t32 = __reduce_min_sync(kFullMask, t32);
      if (lane == 0 && t32 != 0xffffffffu) atomicMin(&s_t32, t32);
      __syncthreads();
      if (s_reach <= p.kp2) { kth = (u64)s_t32 << 32; have_kth = true; }
    }
    // ---- exact k-th key by radix select (mass ties only), then sort the winners ------------------------------------------
    if (!have_kth) {
      for (int pass = 7; pass >= 0; --pass) {
        const int shift = pass * 8;
        for (int i = tid; i < 256; i += SEL_THREADS) hist[i] = 0;
        __syncthreads();
        const u64 prefix = s_prefix;
        const u64 hi_mask = (pass == 7) ? 0ull : (~0ull << (shift + 8));
        for (int base = warp * 32; base < ncand; base += SEL_THREADS) {
          const int i = base + lane;
          const bool active = i < ncand;
          const u64 key = active ? keys[i] : 0ull;
          const bool m = active && (key & hi_mask) == prefix;
          const u32 dgt = m ? ((u32)(key >> shift) & 255u) : 0xffffffffu;
          const u32 peers = __match_any_sync(kFullMask, dgt);
          if (m && lane == __ffs(peers) - 1) atomicAdd(&hist[dgt], (u32)__popc(peers));
        }
        __syncthreads();
        if (tid < 32) {
          const int remaining = s_remaining;
          u32 c[8]; u32 lane_sum = 0;
#pragma unroll
          for (int b = 0; b < 8; ++b) { c[b] = hist[lane * 8 + b]; lane_sum += c[b]; }
          u32 suffix = lane_sum;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const u32 t = __shfl_down_sync(kFullMask, suffix, o);
            if (lane + o < 32) suffix += t;
          }
          const u32 above = suffix - lane_sum;
          if (above < (u32)remaining && (u32)remaining <= suffix) {
            u32 run = above;
#pragma unroll
            for (int b = 7; b >= 0; --b) {
              if (run + c[b] >= (u32)remaining) {
                s_prefix = prefix | ((u64)(u32)(lane * 8 + b) << shift);
                s_remaining = remaining - (int)run;
                // every key of this bucket is a winner: the bucket's lowest possible key is a valid k-th key, stop refining
                if (c[b] == (u32)remaining - run) s_done = 1;
                break;
              }
              run += c[b];
            }
          }
        }
        __syncthreads();
        if (s_done) break;      // block-uniform
      }
      kth = s_prefix;
    }
    for (int i = tid; i < ncand; i += SEL_THREADS) {
      const u64 key = keys[i];
      if (key >= kth && key != 0) {
        const int pos = atomicAdd(&s_nout, 1);
        if (pos < p.kp2) sk[pos] = key;
      }
    }
    __syncthreads();
    block_bitonic_desc(sk, p.kp2);

    // ---- emit: every winning candidate finds its rank by binary search of its key in the sorted winners and leaves its
    // position there; then one thread per winner writes the record.  (A warp-wide scan of the candidate list per winner
    // was 45 % of this kernel's instructions at ~2 200 candidates.)
    const int nwin = min(min(ncand, p.k), s_nout);
    for (int c = tid; c < ncand; c += SEL_THREADS) {
      const u64 key = keys[c];
      if (key >= kth && key != 0) {
        int lo = 0, hi = nwin;                     // sk[0, nwin) is strictly descending (keys are unique)
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (sk[mid] > key) lo = mid + 1; else hi = mid; }
        if (lo < nwin && sk[lo] == key) wslot[lo] = c;
      }
    }
    __syncthreads();
    for (int i = tid; i < p.k; i += SEL_THREADS) {
      const size_t o = (size_t)q * p.k + i;
      float v_score = -CUDART_INF_F, v_gn = 0.0f, v_dn = 0.0f, v_cos = 0.0f;
      long long v_id = -1;
      int v_hop = -1;
      if (i < nwin) {
        const u64 key = sk[i];
        const int slot = wslot[i];
        const float cv = cos_in[slot];
        const int hd = dist_in[slot];
        v_score = key_score(key); v_id = (long long)key_id(key); v_hop = hd;
        v_gn = g_flat ? 0.0f : (hop_decay(hd) - gmin) / g_den;
        v_dn = c_flat ? 0.0f : (cv - cmin) / c_den;
        v_cos = cv;
      }
      if (!xchg) {
        p.out_scores[o] = v_score; p.out_ids[o] = v_id; p.out_hops[o] = v_hop; p.out_gn[o] = v_gn; p.out_dn[o] = v_dn; p.out_cos[o] = v_cos;
      } else {
        for (int g = 0; g < p.c_out.world; ++g) {    // the winner goes into every rank's output slot
          const XOut x = xout(comm_slot(p.c_out, g, e_out), p.B, p.k);
          x.score[o] = v_score; x.ids[o] = v_id; x.hop[o] = v_hop; x.gn[o] = v_gn; x.dn[o] = v_dn; x.cos[o] = v_cos;
        }
      }
    }
    if (xchg && tid < p.c_out.world) {
      const XOut x = xout(comm_slot(p.c_out, tid, e_out), p.B, p.k);
      x.ncand[q] = ncand; x.status[q] = 0;
    }
  }
  if (xchg) {
    if (BIG) { comm_recv_end(p.c_cos, e_cos, gridDim.x); comm_send_end(p.c_out, e_out, gridDim.x); }
    else __threadfence_system();
  }
}

int32_t expand_max_candidates() { return EXP_MAXC_BIG; }

// CSR sanity check at graph creation: row_ptr[0] == 0, non-decreasing, row_ptr[V] == E, 0 <= col_idx < V
__global__ void validate_csr_kernel(const long long* __restrict__ row_ptr, const int* __restrict__ col_idx, long long V, long long E,
                                    int* __restrict__ bad) {
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, step = (long long)gridDim.x * blockDim.x;
  for (long long i = i0; i <= V; i += step) {
    const long long a = row_ptr[i];
    if ((i == 0 && a != 0) || (i == V && a != E) || a < 0 || a > E || (i < V && row_ptr[i + 1] < a)) atomicOr(bad, 1);
  }
  for (long long e = i0; e < E; e += step) {
    const int c = col_idx[e];
    if (c < 0 || (long long)c >= V) atomicOr(bad, 2);
  }
}

int validate_csr(const long long* row_ptr, const int* col_idx, long long V, long long E) {
  int* d_bad = nullptr;
  B200R_CUDA_CHECK(cudaMalloc(&d_bad, sizeof(int)));
  B200R_CUDA_CHECK(cudaMemset(d_bad, 0, sizeof(int)));
  validate_csr_kernel<<<2 * num_sms(), 256>>>(row_ptr, col_idx, V, E, d_bad);
  count_launch();
  int bad = 0;
  cudaError_t e = cudaMemcpy(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost);
  cudaFree(d_bad);
  if (e != cudaSuccess) { set_error("CSR validation failed to run: %s", cudaGetErrorString(e)); return B200R_ECUDA; }
  if (bad) {
    set_error("malformed CSR:%s%s", (bad & 1) ? " row_ptr must start at 0, be non-decreasing and end at E;" : "",
              (bad & 2) ? " col_idx entries must lie in [0, V)" : "");
    return B200R_EINVAL;
  }
  return B200R_OK;
}

// ------------------------------------------------------------------------------------------------ host side
// queries this rank's per-query kernels (BFS, fusion) work on: all of them, or its home queries in exchange mode
static int work_queries(const ExpandParams& p) { return p.mode == 3 ? std::max(p.q_hi - p.q_lo, 0) : p.B; }

template <typename QT>
static int launch_bfs(const ExpandParams& p, cudaStream_t st) {
  ProfScope _ps(st, "graph_bfs");
  const int nq = std::max(work_queries(p), 1);       // exchange mode: a rank without home queries still signals its peers
  B200R_CUDA_CHECK(cudaFuncSetAttribute(bfs_kernel<QT, EXP_MAXC_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bfs_smem_bytes(EXP_MAXC_SMALL)));
  bfs_kernel<QT, EXP_MAXC_SMALL><<<nq, BFS_THREADS, bfs_smem_bytes(EXP_MAXC_SMALL), st>>>(p);
  count_launch();
  // persistent: only the queries that overflowed the small instantiation do any work
  B200R_CUDA_CHECK(cudaFuncSetAttribute(bfs_kernel<QT, EXP_MAXC_BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bfs_smem_bytes(EXP_MAXC_BIG)));
  bfs_kernel<QT, EXP_MAXC_BIG><<<std::min(nq, exp_persistent_ctas()), BFS_THREADS, bfs_smem_bytes(EXP_MAXC_BIG), st>>>(p);
  count_launch();
  return B200R_OK;
}

template <typename T>
static int launch_gather(const ExpandParams& p, cudaStream_t st) {
  ProfScope _ps(st, "graph_gather");
  const int dpad = (p.D + 3) & ~3;
  // A/B knobs (measurements only): B200R_GATHER_WARPS = warps per CTA, B200R_GATHER_CTAS = CTAs per SM
  static const int env_warps = [] { const char* e = getenv("B200R_GATHER_WARPS"); return e ? atoi(e) : 0; }();
  static const int env_ctas = [] { const char* e = getenv("B200R_GATHER_CTAS"); return e ? atoi(e) : 0; }();
  const int max_warps = (env_warps >= 1 && env_warps <= GATHER_MAX_WARPS) ? env_warps : GATHER_MAX_WARPS;
  const int ctas_per_sm = (env_ctas >= 1 && env_ctas <= 8) ? env_ctas : GATHER_CTAS_PER_SM;
  int warps = (int)std::min<size_t>(max_warps, (96 * 1024) / ((size_t)dpad * 4));
  if (warps < 1) {
    if ((size_t)dpad * 4 > 200 * 1024) { set_error("expand_rescore: D=%d too large for shared memory", p.D); return B200R_EINVAL; }
    warps = 1;
  }
  const size_t smem = (size_t)warps * dpad * 4;
  B200R_CUDA_CHECK(cudaFuncSetAttribute(rescore_gather_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  profile_begin(st);
  rescore_gather_kernel<T><<<num_sms() * ctas_per_sm, warps * 32, smem, st>>>(p, dpad);
  profile_end(st);
  count_launch();
  return B200R_OK;
}

static int launch_select(const ExpandParams& p, cudaStream_t st) {
  ProfScope _ps(st, "graph_fuse_select");
  if (select_smem_bytes(EXP_MAXC_BIG, p.kp2) > 220 * 1024) { set_error("expand_rescore: k=%d too large for shared memory", p.k); return B200R_EINVAL; }
  const int nq = std::max(work_queries(p), 1);
  B200R_CUDA_CHECK(cudaFuncSetAttribute(fuse_select_kernel<EXP_MAXC_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)select_smem_bytes(EXP_MAXC_SMALL, p.kp2)));
  fuse_select_kernel<EXP_MAXC_SMALL><<<nq, SEL_THREADS, select_smem_bytes(EXP_MAXC_SMALL, p.kp2), st>>>(p);
  count_launch();
  B200R_CUDA_CHECK(cudaFuncSetAttribute(fuse_select_kernel<EXP_MAXC_BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)select_smem_bytes(EXP_MAXC_BIG, p.kp2)));
  fuse_select_kernel<EXP_MAXC_BIG><<<std::min(nq, exp_persistent_ctas()), SEL_THREADS, select_smem_bytes(EXP_MAXC_BIG, p.kp2), st>>>(p);
  count_launch();
  return B200R_OK;
}

// exchange mode runs in four phases so that a host driving several ranks from one thread can issue phase i on every rank
// before phase i+1 on any (a kernel never waits for a kernel that has not been launched):
//   1  BFS of the home queries, candidates stored into their owners' windows
//   2  owners: plan + gather / re-score, cosines stored into the home ranks' windows
//   3  home: fusion + top-k, results stored into every rank's window
//   4  every rank: results -> the caller's arrays
template <typename T>
static int launch_expand_phase_t(const ExpandParams& p, int phase, int qdtype, cudaStream_t st) {
  typedef typename QueryType<T>::type QT;
  int rc = B200R_OK;
  if (phase == 1) {
    if ((rc = launch_bfs<QT>(p, st)) != B200R_OK) return rc;
  } else if (phase == 2) {
    if (p.metric == B200R_COS && (rc = launch_query_inv_norm(p.Q, p.B, p.D, qdtype, p.ws.qinv, st)) != B200R_OK) return rc;
    { ProfScope _ps(st, "graph_plan"); gather_plan_kernel<<<1, 1024, 0, st>>>(p); }
    count_launch();
    if ((rc = launch_gather<T>(p, st)) != B200R_OK) return rc;
  } else if (phase == 3) {
    if ((rc = launch_select(p, st)) != B200R_OK) return rc;
  } else {
    { ProfScope _ps(st, "graph_collect"); collect_outputs_kernel<<<2 * num_sms(), 256, 0, st>>>(p); }
    count_launch();
  }
  B200R_CUDA_CHECK(cudaGetLastError());
  return B200R_OK;
}

template <typename T>
static int launch_expand_t(const ExpandParams& p, cudaStream_t st) {
  typedef typename QueryType<T>::type QT;
  int rc = B200R_OK;
  if (p.mode != 2) {
    B200R_CUDA_CHECK(cudaMemsetAsync(p.ws.counter, 0, sizeof(u64), st));
    if ((rc = launch_bfs<QT>(p, st)) != B200R_OK) return rc;
    if ((rc = launch_gather<T>(p, st)) != B200R_OK) return rc;
    if (p.mode == 1) {
      minmax_decode_kernel<<<(p.B + 255) / 256, 256, 0, st>>>(p.ws.mm, p.minmax, p.out_status, p.B);
      count_launch();
    }
  }
  if (p.mode != 1 && (rc = launch_select(p, st)) != B200R_OK) return rc;
  B200R_CUDA_CHECK(cudaGetLastError());
  return B200R_OK;
}

size_t expand_xchg_req_bytes(int per) { return align_up(xreq_record_bytes(per), 16); }
size_t expand_xchg_cos_bytes(int per) { return (size_t)per * EXP_MAXC_BIG * 4; }
size_t expand_xchg_out_bytes(int B, int k) { return xout_bytes(B, k); }

static void fill_params(ExpandParams& p, const Index* ix, const Graph* g, const void* Q, const int64_t* seeds, int B, int S, int window,
                        float alpha_dense, float alpha_graph, int k, float* out_scores, int64_t* out_ids, int32_t* out_hops,
                        float* out_gn, float* out_dn, float* out_cos, int32_t* out_ncand, int32_t* out_status, void* ws) {
  memset(&p, 0, sizeof(p));
  p.X = ix->X; p.Q = Q; p.row_mult = ix->row_mult;
  p.row_ptr = reinterpret_cast<const long long*>(g->row_ptr); p.col_idx = g->col_idx;
  p.node_rows = reinterpret_cast<const long long*>(g->node_rows);
  p.seeds = reinterpret_cast<const long long*>(seeds);
  p.N = ix->N; p.V = g->V; p.id_base = ix->id_base;
  p.D = ix->D; p.B = B; p.S = S; p.window = window; p.k = k; p.metric = ix->metric;
  p.kp2 = 2; while (p.kp2 < k) p.kp2 <<= 1;
  p.alpha_dense = alpha_dense; p.alpha_graph = alpha_graph;
  p.out_scores = out_scores; p.out_ids = reinterpret_cast<long long*>(out_ids); p.out_hops = out_hops;
  p.out_gn = out_gn; p.out_dn = out_dn; p.out_cos = out_cos; p.out_ncand = out_ncand; p.out_status = out_status;
  p.ws = carve_ws(ws, B);
}

int launch_expand_rescore_exchange(const Index* ix, const Graph* g, const Comm* comm, const void* Q, const int64_t* seeds, int B, int S,
                                   int window, float alpha_dense, float alpha_graph, int k, int64_t n_total_rows, float* out_scores,
                                   int64_t* out_ids, int32_t* out_hops, float* out_gn, float* out_dn, float* out_cos, int32_t* out_ncand,
                                   int32_t* out_status, void* ws, int phase, cudaStream_t st) {
  if (B > (1 << 20)) { set_error("expand_rescore: at most %d queries per call", 1 << 20); return B200R_EINVAL; }
  if (phase < 1 || phase > 4) { set_error("phase must be 1..4"); return B200R_EINVAL; }
  if (n_total_rows <= 0) { set_error("n_total_rows must be > 0"); return B200R_EINVAL; }
  ExpandParams p;
  fill_params(p, ix, g, Q, seeds, B, S, window, alpha_dense, alpha_graph, k, out_scores, out_ids, out_hops, out_gn, out_dn, out_cos,
              out_ncand, out_status, ws);
  p.mode = 3;
  p.own_lo = INT64_MIN; p.own_hi = INT64_MAX;          // the home rank ranks every candidate of its queries
  const int G = comm->world;
  p.per = (B + G - 1) / G;
  p.q_lo = std::min(B, comm->rank * p.per);
  p.q_hi = std::min(B, p.q_lo + p.per);
  p.rows_per = (n_total_rows + G - 1) / G;
  int rc;
  if ((rc = comm_dev(comm, kChanGraphReq, expand_xchg_req_bytes(p.per), &p.c_req))) return rc;
  if ((rc = comm_dev(comm, kChanGraphCos, 0, &p.c_cos))) return rc;
  if ((rc = comm_dev(comm, kChanGraphOut, 0, &p.c_out))) return rc;
  if (expand_xchg_cos_bytes(p.per) > p.c_cos.slot_bytes || expand_xchg_out_bytes(B, k) > p.c_out.slot_bytes) {
    set_error("graph exchange of (B=%d, k=%d) exceeds the comm's windows (created for max_B=%d, max_graph_k=%d)", B, k, comm->max_B, comm->max_graph_k);
    return B200R_EINVAL;
  }
  if (phase == 1) B200R_CUDA_CHECK(cudaMemsetAsync(p.ws.counter, 0, sizeof(u64), st));
  const int qdtype = ix->dtype == B200R_I8 ? B200R_F32 : ix->dtype;
  switch (ix->dtype) {
    case B200R_BF16: return launch_expand_phase_t<__nv_bfloat16>(p, phase, qdtype, st);
    case B200R_F16: return launch_expand_phase_t<__half>(p, phase, qdtype, st);
    case B200R_F32: return launch_expand_phase_t<float>(p, phase, qdtype, st);
    case B200R_I8: return launch_expand_phase_t<int8_t>(p, phase, qdtype, st);
  }
  set_error("bad dtype");
  return B200R_EINVAL;
}


// ------------------------------------------------------------------------------------------------ unbounded frontier
// The reference's BFS has no candidate limit (graph_utils.py:109-127).  A query that overflows the 8192-candidate
// shared-memory instantiation is redone here with everything in global memory: a visited BITMAP over the graph's nodes,
// a candidate array that can hold every node, one grid-wide kernel per BFS level, the same gather kernel for the
// re-score, an element-wise fusion into 64-bit keys and the list-select kernel for the top-k.  One query per call.
struct BigWs {
  u32* visited;      // [V/32 + 1]
  u32* cand;         // [V] nodes in BFS order, then corpus rows
  uint8_t* dist;     // [V]
  float* cos;        // [V]
  u64* keys;         // [V]
  int* ncand;        // [1]
  int* level_end;    // [257]
  int* nown; float* qinv; u32* mm; u32* seg_start; u32* seg_q; u64* counter;   // what the gather kernel reads (B = 1)
};

size_t expand_big_workspace_bytes(long long V) {
  const size_t v = (size_t)V;
  return align_up((v / 32 + 1) * 4, 256) + align_up(v * 4, 256) * 2 + align_up(v, 256) + align_up(v * 8, 256) + 8 * 256 + align_up(257 * 4, 256);
}

static BigWs carve_big(void* ws, long long V) {
  const size_t v = (size_t)V;
  char* w = reinterpret_cast<char*>(ws);
  BigWs b;
  b.visited = reinterpret_cast<u32*>(w); w += align_up((v / 32 + 1) * 4, 256);
  b.cand = reinterpret_cast<u32*>(w); w += align_up(v * 4, 256);
  b.cos = reinterpret_cast<float*>(w); w += align_up(v * 4, 256);
  b.dist = reinterpret_cast<uint8_t*>(w); w += align_up(v, 256);
  b.keys = reinterpret_cast<u64*>(w); w += align_up(v * 8, 256);
  b.ncand = reinterpret_cast<int*>(w); w += 256;
  b.level_end = reinterpret_cast<int*>(w); w += align_up(257 * 4, 256);
  b.nown = reinterpret_cast<int*>(w); w += 256;
  b.qinv = reinterpret_cast<float*>(w); w += 256;
  b.mm = reinterpret_cast<u32*>(w); w += 256;
  b.seg_start = reinterpret_cast<u32*>(w); w += 256;
  b.seg_q = reinterpret_cast<u32*>(w); w += 256;
  b.counter = reinterpret_cast<u64*>(w); w += 256;
  return b;
}

__device__ __forceinline__ bool big_visit(u32* visited, u32 v) {
  const u32 bit = 1u << (v & 31u);
  return (atomicOr(&visited[v >> 5], bit) & bit) == 0u;
}

__global__ void bfs_big_seed_kernel(const long long* __restrict__ seeds, int S, long long V, BigWs b) {
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    const long long s = seeds[i];
    if (s >= 0 && s < V && big_visit(b.visited, (u32)s)) {
      const int pos = atomicAdd(b.ncand, 1);
      b.cand[pos] = (u32)s; b.dist[pos] = 0;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) b.level_end[0] = *b.ncand;
}

// one warp per frontier node, lanes over its adjacency
__global__ void bfs_big_level_kernel(const long long* __restrict__ row_ptr, const int* __restrict__ col_idx, long long V, int d, BigWs b) {
  const int lo = d >= 2 ? b.level_end[d - 2] : 0, hi = b.level_end[d - 1];
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int f = lo + warp; f < hi; f += nwarps) {
    const u32 u = b.cand[f];
    const long long e0 = __ldg(row_ptr + u), e1 = __ldg(row_ptr + u + 1);
    for (long long e = e0 + lane; e < e1; e += 32) {
      const int v = __ldg(col_idx + e);
      if (v >= 0 && (long long)v < V && big_visit(b.visited, (u32)v)) {
        const int pos = atomicAdd(b.ncand, 1);
        b.cand[pos] = (u32)v; b.dist[pos] = (uint8_t)min(d, 255);
      }
    }
  }
}
__global__ void bfs_big_close_level_kernel(int d, BigWs b) { b.level_end[d] = *b.ncand; }

// nodes -> corpus rows, and the gather kernel's bookkeeping for a single query
__global__ void bfs_big_finish_kernel(const long long* __restrict__ node_rows, BigWs b, int* out_ncand) {
  const int n = *b.ncand;
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x, step = gridDim.x * blockDim.x;
  if (node_rows)
    for (int i = i0; i < n; i += step) {
      const long long r = node_rows[b.cand[i]];
      b.cand[i] = (r >= 0 && r < 0xffffffffll) ? (u32)r : 0xfffffffeu;
    }
  if (i0 == 0) {
    b.nown[0] = n; b.seg_start[0] = 0; b.seg_q[0] = 0;
    b.counter[0] = (1ull << 32) | (u64)((n + GATHER_ROWS - 1) / GATHER_ROWS);
    b.mm[0] = 0xffffffffu; b.mm[1] = 0u;
    *out_ncand = n;
  }
}

__global__ void fuse_big_kernel(BigWs b, int window, float alpha_dense, float alpha_graph) {
  const int n = *b.ncand;
  int dmax = 0;
  for (int d = 1; d <= window; ++d) if (b.level_end[d] > b.level_end[d - 1]) dmax = d;
  const float gmax = hop_decay(0), gmin = hop_decay(dmax);
  const bool g_flat = !(gmax > gmin);
  const u32 lo = b.mm[0], hi = b.mm[1];
  const float cmin = lo <= hi ? ordered_to_f32(lo) : CUDART_INF_F, cmax = lo <= hi ? ordered_to_f32(hi) : -CUDART_INF_F;
  const bool c_flat = !(cmax > cmin);
  const float c_den = c_flat ? 1.0f : (cmax - cmin), g_den = g_flat ? 1.0f : (gmax - gmin);
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x, step = gridDim.x * blockDim.x;
  for (int i = i0; i < n; i += step) {
    const float dn = c_flat ? 0.0f : (b.cos[i] - cmin) / c_den;
    const float gn = g_flat ? 0.0f : (hop_decay(b.dist[i]) - gmin) / g_den;
    b.keys[i] = make_key(alpha_graph * gn + alpha_dense * dn, b.cand[i]);
  }
}

// winner j (one CTA): find its slot in the candidate array, emit the per-hit fields
__global__ void __launch_bounds__(256) lookup_big_kernel(BigWs b, int window, int k, const long long* __restrict__ ids, float* out_scores,
                                                          int* out_hops, float* out_gn, float* out_dn, float* out_cos) {
  __shared__ int s_slot;
  const int j = blockIdx.x, n = *b.ncand;
  if (threadIdx.x == 0) s_slot = -1;
  __syncthreads();
  const long long id = ids[j];
  if (id >= 0)
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      if ((long long)b.cand[i] == id) s_slot = i;          // rows are unique within a candidate list
  __syncthreads();
  if (threadIdx.x != 0) return;
  const int slot = s_slot;
  if (slot < 0) { out_scores[j] = -CUDART_INF_F; out_hops[j] = -1; out_gn[j] = 0; out_dn[j] = 0; out_cos[j] = 0; return; }
  int dmax = 0;
  for (int d = 1; d <= window; ++d) if (b.level_end[d] > b.level_end[d - 1]) dmax = d;
  const float gmax = hop_decay(0), gmin = hop_decay(dmax);
  const bool g_flat = !(gmax > gmin);
  const u32 lo = b.mm[0], hi = b.mm[1];
  const float cmin = lo <= hi ? ordered_to_f32(lo) : CUDART_INF_F, cmax = lo <= hi ? ordered_to_f32(hi) : -CUDART_INF_F;
  const bool c_flat = !(cmax > cmin);
  const int hd = b.dist[slot];
  const float cv = b.cos[slot];
  out_hops[j] = hd; out_cos[j] = cv;
  out_gn[j] = g_flat ? 0.0f : (hop_decay(hd) - gmin) / (gmax - gmin);
  out_dn[j] = c_flat ? 0.0f : (cv - cmin) / (cmax - cmin);
}

int launch_select_flat_keys(const void* keys, const int* count_dev, long long cap, int k, float* out_scores, int64_t* out_ids, cudaStream_t st);

template <typename T>
static int launch_gather_big(const ExpandParams& p, cudaStream_t st) { return launch_gather<T>(p, st); }

int launch_expand_rescore_big(const Index* ix, const Graph* g, const void* Q, const int64_t* seeds, int S, int window, float alpha_dense,
                              float alpha_graph, int k, float* out_scores, int64_t* out_ids, int32_t* out_hops, float* out_gn, float* out_dn,
                              float* out_cos, int32_t* out_ncand, void* ws, cudaStream_t st) {
  ProfScope _ps(st, "graph_big_frontier");
  BigWs b = carve_big(ws, g->V);
  const long long V = g->V;
  B200R_CUDA_CHECK(cudaMemsetAsync(b.visited, 0, ((size_t)V / 32 + 1) * 4, st));
  B200R_CUDA_CHECK(cudaMemsetAsync(b.ncand, 0, 256 + align_up(257 * 4, 256), st));
  bfs_big_seed_kernel<<<1, 256, 0, st>>>(reinterpret_cast<const long long*>(seeds), S, V, b);
  for (int d = 1; d <= window; ++d) {
    bfs_big_level_kernel<<<2 * num_sms(), 256, 0, st>>>(reinterpret_cast<const long long*>(g->row_ptr), g->col_idx, V, d, b);
    bfs_big_close_level_kernel<<<1, 1, 0, st>>>(d, b);
  }
  bfs_big_finish_kernel<<<2 * num_sms(), 256, 0, st>>>(reinterpret_cast<const long long*>(g->node_rows), b, out_ncand);
  count_launch(3 + 2 * window);
  int rc;
  const int qdtype = ix->dtype == B200R_I8 ? B200R_F32 : ix->dtype;
  if (ix->metric == B200R_COS) { if ((rc = launch_query_inv_norm(Q, 1, ix->D, qdtype, b.qinv, st))) return rc; }
  else { const float one = 1.0f; B200R_CUDA_CHECK(cudaMemcpyAsync(b.qinv, &one, 4, cudaMemcpyHostToDevice, st)); }
  // the re-score gather, unchanged: one reservation covering the whole list of the single query
  ExpandParams p;
  memset(&p, 0, sizeof(p));
  p.X = ix->X; p.Q = Q; p.row_mult = ix->row_mult; p.N = ix->N; p.V = V; p.id_base = ix->id_base; p.D = ix->D; p.B = 1; p.metric = ix->metric;
  p.own_lo = INT64_MIN; p.own_hi = INT64_MAX;
  p.ws.cand = b.cand; p.ws.cos = b.cos; p.ws.dist = b.dist; p.ws.ncand = b.ncand; p.ws.nown = b.nown; p.ws.qinv = b.qinv; p.ws.mm = b.mm;
  p.ws.seg_start = b.seg_start; p.ws.seg_q = b.seg_q; p.ws.counter = b.counter;
  switch (ix->dtype) {
    case B200R_BF16: rc = launch_gather_big<__nv_bfloat16>(p, st); break;
    case B200R_F16: rc = launch_gather_big<__half>(p, st); break;
    case B200R_F32: rc = launch_gather_big<float>(p, st); break;
    default: rc = launch_gather_big<int8_t>(p, st); break;
  }
  if (rc) return rc;
  fuse_big_kernel<<<2 * num_sms(), 256, 0, st>>>(b, window, alpha_dense, alpha_graph);
  count_launch();
  if ((rc = launch_select_flat_keys(b.keys, b.ncand, V, k, out_scores, out_ids, st))) return rc;
  lookup_big_kernel<<<k, 256, 0, st>>>(b, window, k, reinterpret_cast<const long long*>(out_ids), out_scores, out_hops, out_gn, out_dn, out_cos);
  count_launch();
  B200R_CUDA_CHECK(cudaGetLastError());
  return B200R_OK;
}

int launch_expand_rescore(const Index* ix, const Graph* g, const void* Q, const int64_t* seeds, int B, int S, int window,
                          float alpha_dense, float alpha_graph, int k, float* out_scores, int64_t* out_ids, int32_t* out_hops,
                          float* out_gn, float* out_dn, float* out_cos, int32_t* out_ncand, int32_t* out_status, int mode,
                          int64_t own_lo, int64_t own_hi, void* ws, float* minmax, cudaStream_t st) {
  if (B > (1 << 20)) { set_error("expand_rescore: at most %d queries per call", 1 << 20); return B200R_EINVAL; }
  ExpandParams p;
  memset(&p, 0, sizeof(p));
  p.X = ix->X; p.Q = Q; p.row_mult = ix->row_mult;
  p.row_ptr = reinterpret_cast<const long long*>(g->row_ptr); p.col_idx = g->col_idx;
  p.node_rows = reinterpret_cast<const long long*>(g->node_rows);
  p.seeds = reinterpret_cast<const long long*>(seeds);
  p.N = ix->N; p.V = g->V; p.id_base = ix->id_base;
  p.D = ix->D; p.B = B; p.S = S; p.window = window; p.k = k; p.metric = ix->metric;
  p.kp2 = 2; while (p.kp2 < k) p.kp2 <<= 1;
  p.alpha_dense = alpha_dense; p.alpha_graph = alpha_graph;
  p.out_scores = out_scores; p.out_ids = reinterpret_cast<long long*>(out_ids); p.out_hops = out_hops;
  p.out_gn = out_gn; p.out_dn = out_dn; p.out_cos = out_cos; p.out_ncand = out_ncand; p.out_status = out_status;
  p.mode = mode; p.own_lo = own_lo; p.own_hi = own_hi;
  p.minmax = minmax;
  p.ws = carve_ws(ws, B);
  switch (ix->dtype) {
    case B200R_BF16: return launch_expand_t<__nv_bfloat16>(p, st);
    case B200R_F16: return launch_expand_t<__half>(p, st);
    case B200R_F32: return launch_expand_t<float>(p, st);
    case B200R_I8: return launch_expand_t<int8_t>(p, st);
  }
  set_error("bad dtype");
  return B200R_EINVAL;
}

}  // namespace b200r
