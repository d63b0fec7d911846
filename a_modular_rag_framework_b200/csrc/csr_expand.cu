// csr_expand.cu -- K3: graph-enhanced second stage.  One CTA (512 threads) per query:
//   1. level-synchronous BFS from the seed passages over the CSR adjacency, depth <= window.  The
//      `visited` set is a shared-memory hash set (first arrival == minimum hop distance).  A frontier
//      level is expanded edge-parallel: adjacency ranges of the whole level are read at once, a block scan
//      turns degrees into edge offsets, and the flattened edge list is spread over all 512 threads (coalesced
//      col_idx reads, several in flight per thread; hubs need no special case), with a warp-level
//      __match_any_sync pre-dedup so each distinct neighbour costs one atomic per warp.
//   2. re-score every reached node against the query: one warp per 4 candidate rows, all 128-bit row
//      loads of the group in flight together (this gather is the HBM-bound part).
//   3. per-query min/max of both channels, fused = a_g*minmax(decay(hop)) + a_d*minmax(cos).
//   4. exact top-k by (fused desc, id asc): radix select of the k-th key in smem + bitonic sort of k.
//
// Two instantiations share the code: 4096 candidates (71 KB smem, 3 CTAs/SM) runs for every query; the
// 8192-candidate one re-runs only the queries that overflowed the first; beyond that B200R_EOVERFLOW.
//
// Replaces expand_qmatch_neighbors (reference app/modules/retrieval/graph_utils.py:58-129: BFS :109-127,
// decay :87-94), DenseReranker._cosine (retrieval_backend.py:192-197), _minmax_norm (:296-301) and the
// fusion / sort / slice of HybridRetrievalBackend.run (:353-372).
#include "common.cuh"
#include "internal.h"
#include <math_constants.h>
#include <cstdlib>

namespace b200r {

constexpr int EXP_THREADS = 512;
constexpr int EXP_WARPS = EXP_THREADS / 32;
constexpr int EXP_MAXC_BIG = 8192;
constexpr int EXP_FB = 1024;          // frontier nodes expanded per batch (2 per thread)
constexpr int EXP_EU = 4;             // col_idx loads in flight per thread
constexpr u32 EXP_EMPTY = 0xffffffffu;
constexpr int kStatusRetry = 1;       // internal: overflowed the small instantiation, big one will redo it

struct ExpandParams {
  const void* X;
  const void* Q;
  const float* row_mult;
  const long long* row_ptr;
  const int* col_idx;
  const long long* seeds;
  long long N, V, id_base;
  int D, B, S, window, k, kp2, metric;
  float alpha_dense, alpha_graph;
  float* out_scores; long long* out_ids; int* out_hops; float* out_gn; float* out_dn; float* out_cos;
  int* out_ncand; int* out_status;
  // row-sharded corpora (one rank per GPU): mode 0 = everything in one launch; 1 = BFS + re-score the candidates this
  // rank owns + local min/max, parked in the workspace; 2 = fuse with the all-reduced min/max and rank the owned ones
  int mode;
  long long own_lo, own_hi;       // global node ids this rank scores: [own_lo, own_hi)
  u32* ws_cand; float* ws_cos; uint8_t* ws_dist; int* ws_ncand; float* ws_minmax;   // [B][EXP_MAXC_BIG] / [B] / [B][2]
};

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

// EXP_ROWS: candidate rows gathered together by one warp (all their 128-bit loads in flight at once)
template <typename T, int MAXC, bool RETRY_PASS, int EXP_ROWS>
__global__ void __launch_bounds__(EXP_THREADS, 3) expand_rescore_kernel(const ExpandParams p) {
  typedef typename QueryType<T>::type QT;
  constexpr int VEC = ElemTraits<T>::kVec;
  constexpr int HASH = 2 * MAXC;
  const int q = blockIdx.x;
  if (RETRY_PASS && p.out_status[q] != kStatusRetry) return;
  if (p.mode == 2 && p.out_status[q] != 0) return;     // pass 1 already reported the overflow

  extern __shared__ uint8_t smem_raw[];
  u64* keys = reinterpret_cast<u64*>(smem_raw);                    // [MAXC]  (aliases the hash table, used after the BFS)
  u32* table = reinterpret_cast<u32*>(smem_raw);                   // [HASH] u32
  u32* cand = reinterpret_cast<u32*>(smem_raw + MAXC * 8);         // [MAXC] node ids
  float* cosv = reinterpret_cast<float*>(cand + MAXC);             // [MAXC]
  uint8_t* dist = reinterpret_cast<uint8_t*>(cosv + MAXC);         // [MAXC]
  float* qs = reinterpret_cast<float*>(dist + MAXC);               // [D]
  u64* sk = reinterpret_cast<u64*>(qs + ((p.D + 1) & ~1));         // [kp2]
  __shared__ int s_count, s_overflow, s_remaining, s_nout, s_done;
  __shared__ u32 s_wsum[EXP_WARPS];
  __shared__ unsigned long long s_wwide[EXP_WARPS];
  __shared__ float s_red[4][EXP_WARPS];
  __shared__ double s_qn[EXP_WARPS];
  __shared__ u32 hist[256];
  __shared__ u64 s_prefix;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  int ncand = 0;
  float qinv = 1.0f;
  if (p.mode != 2) {
  for (int i = tid; i < HASH; i += EXP_THREADS) table[i] = EXP_EMPTY;
  if (tid == 0) { s_count = 0; s_overflow = 0; }
  // query -> fp32 in smem, and its norm
  const QT* Qp = reinterpret_cast<const QT*>(p.Q) + (size_t)q * p.D;
  double qn = 0.0;
  for (int d = tid; d < p.D; d += EXP_THREADS) {
    const float v = q_to_float<QT>(Qp[d]);
    qs[d] = v;
    qn += (double)v * (double)v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) qn += __shfl_xor_sync(kFullMask, qn, o);
  if (lane == 0) s_qn[warp] = qn;
  __syncthreads();
  double qn_tot = 0.0;
  for (int w = 0; w < EXP_WARPS; ++w) qn_tot += s_qn[w];
  qinv = (p.metric == B200R_COS) ? (qn_tot > 0.0 ? (float)(1.0 / sqrt(qn_tot)) : 0.0f) : 1.0f;

  // ---- level 0: seeds ---------------------------------------------------------------------------
  for (int i = tid; i < p.S; i += EXP_THREADS) {
    const long long s = p.seeds[(size_t)q * p.S + i];   // node id == global corpus row id
    if (s >= 0 && s < p.V) {
      if (hash_insert<HASH>(table, (u32)s)) {
        const int pos = atomicAdd(&s_count, 1);
        if (pos < MAXC) { cand[pos] = (u32)s; dist[pos] = 0; } else s_overflow = 1;
      }
    }
  }
  __syncthreads();

  // ---- levels 1..window: edge-parallel expansion ----------------------------------------------------
  // A level is expanded in batches of EXP_FB frontier nodes: (1) every node's adjacency range (row_ptr) is read at
  // once, (2) a block scan turns the degrees into edge offsets, (3) the batch's edges are flattened over all 512
  // threads (binary search of the offsets), EXP_EU col_idx loads in flight per thread, consecutive threads on
  // consecutive edges.  A level therefore costs ~2 memory latencies whatever the degree distribution (hubs included).
  // The scratch lives in the cosine array, which is not written before the re-score.
  long long* f_e0 = reinterpret_cast<long long*>(cosv);               // [EXP_FB]
  u32* f_pre = reinterpret_cast<u32*>(cosv) + 2 * EXP_FB;             // [EXP_FB + 1] exclusive edge offsets
  int level_begin = 0;
  for (int d = 1; d <= p.window; ++d) {
    const int level_end = min(s_count, MAXC);
    __syncthreads();                      // everyone has read s_count before the level appends to it
    if (level_end == level_begin || s_overflow) break;  // block-uniform
    const uint8_t dd = (uint8_t)min(d, 255);
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
      for (int w = 0; w < EXP_WARPS; ++w) {
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
      // (3) the batch's edges, flattened
      for (u32 eb = 0; eb < total; eb += EXP_THREADS * EXP_EU) {
        u32 v[EXP_EU];
#pragma unroll
        for (int r = 0; r < EXP_EU; ++r) {
          const u32 e = eb + (u32)r * EXP_THREADS + (u32)tid;
          v[r] = EXP_EMPTY;
          if (e < total) {
            int lo = 0, hi = nb;              // largest node with f_pre[node] <= e
            while (hi - lo > 1) {
              const int mid = (lo + hi) >> 1;
              if (f_pre[mid] <= e) lo = mid; else hi = mid;
            }
            v[r] = (u32)__ldg(p.col_idx + f_e0[lo] + (long long)(e - f_pre[lo]));
          }
        }
#pragma unroll
        for (int r = 0; r < EXP_EU; ++r) {
          if (eb + (u32)r * EXP_THREADS >= total) break;        // block-uniform
          const u32 peers = __match_any_sync(kFullMask, v[r]);
          const bool leader = v[r] != EXP_EMPTY && (lane == __ffs(peers) - 1);
          if (leader && !*(volatile int*)&s_overflow && hash_insert<HASH>(table, v[r])) {
            const int pos = atomicAdd(&s_count, 1);
            if (pos < MAXC) { cand[pos] = v[r]; dist[pos] = dd; } else s_overflow = 1;
          }
        }
      }
      __syncthreads();                    // the scratch is reused by the next batch
    }
    level_begin = level_end;
  }
  __syncthreads();
  ncand = min(s_count, MAXC);
  if (s_overflow) {  // block-uniform
    if (tid == 0) { p.out_status[q] = (MAXC < EXP_MAXC_BIG) ? kStatusRetry : B200R_EOVERFLOW; p.out_ncand[q] = s_count; }
    for (int i = tid; i < p.k; i += EXP_THREADS) {
      const size_t o = (size_t)q * p.k + i;
      p.out_scores[o] = -CUDART_INF_F; p.out_ids[o] = -1; p.out_hops[o] = -1; p.out_gn[o] = 0; p.out_dn[o] = 0; p.out_cos[o] = 0;
    }
    return;
  }

  // ---- re-score: one warp per EXP_ROWS candidate rows, their loads in flight together ---------------
  const uint8_t* Xb = reinterpret_cast<const uint8_t*>(p.X);
  const int chunks = p.D / VEC;
  const size_t row_bytes = (size_t)p.D * sizeof(T);
  for (int c0 = warp * EXP_ROWS; c0 < ncand; c0 += EXP_WARPS * EXP_ROWS) {
    const uint8_t* rp[EXP_ROWS];
    bool have[EXP_ROWS];
    float acc[EXP_ROWS];
#pragma unroll
    for (int r = 0; r < EXP_ROWS; ++r) {
      const int c = min(c0 + r, ncand - 1);
      const long long node = (long long)cand[c];
      const long long row = node - p.id_base;
      have[r] = (c0 + r < ncand) && node >= p.own_lo && node < p.own_hi && row >= 0 && row < p.N;
      rp[r] = Xb + (size_t)(have[r] ? row : 0) * row_bytes;
      acc[r] = 0.0f;
    }
    // lane r fetches row r's multiplier together with the rows (not after the reduction, where its latency would be exposed)
    float mult = 0.0f;
    if (lane < EXP_ROWS && c0 + lane < ncand) {
      const long long node = (long long)cand[c0 + lane];
      const long long row = node - p.id_base;
      if (node >= p.own_lo && node < p.own_hi && row >= 0 && row < p.N) mult = p.row_mult ? __ldg(p.row_mult + row) : 1.0f;
    }
    for (int chunk = lane; chunk < chunks; chunk += 32) {
      uint4 xv[EXP_ROWS];
#pragma unroll
      for (int r = 0; r < EXP_ROWS; ++r) xv[r] = ld_stream_16B(rp[r] + (size_t)chunk * 16);
      const float* qc = qs + chunk * VEC;
#pragma unroll
      for (int r = 0; r < EXP_ROWS; ++r) {
        float xf[VEC];
        ElemTraits<T>::unpack(xv[r], xf);
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[r] = fmaf(xf[e], qc[e], acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < EXP_ROWS; ++r) {
      float a = acc[r];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(kFullMask, a, o);
      const float m = __shfl_sync(kFullMask, mult, r);
      if (lane == 0 && c0 + r < ncand) cosv[c0 + r] = (a * m) * qinv + 0.0f;
    }
  }
  __syncthreads();
  } else {
    // pass 2: candidates, hops and cosines were parked by pass 1
    ncand = p.ws_ncand[q];
    for (int c = tid; c < ncand; c += EXP_THREADS) {
      cand[c] = p.ws_cand[(size_t)q * EXP_MAXC_BIG + c];
      cosv[c] = p.ws_cos[(size_t)q * EXP_MAXC_BIG + c];
      dist[c] = p.ws_dist[(size_t)q * EXP_MAXC_BIG + c];
    }
    __syncthreads();
  }

  // ---- min/max of both channels (retrieval_backend.py:296-301) ----------------------------------
  float cmin = CUDART_INF_F, cmax = -CUDART_INF_F, gmin = CUDART_INF_F, gmax = -CUDART_INF_F;
  for (int c = tid; c < ncand; c += EXP_THREADS) {
    const float cv = cosv[c], gv = hop_decay(dist[c]);
    const long long node = (long long)cand[c];
    if (node >= p.own_lo && node < p.own_hi) { cmin = fminf(cmin, cv); cmax = fmaxf(cmax, cv); }
    gmin = fminf(gmin, gv); gmax = fmaxf(gmax, gv);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cmin = fminf(cmin, __shfl_xor_sync(kFullMask, cmin, o)); cmax = fmaxf(cmax, __shfl_xor_sync(kFullMask, cmax, o));
    gmin = fminf(gmin, __shfl_xor_sync(kFullMask, gmin, o)); gmax = fmaxf(gmax, __shfl_xor_sync(kFullMask, gmax, o));
  }
  if (lane == 0) { s_red[0][warp] = cmin; s_red[1][warp] = cmax; s_red[2][warp] = gmin; s_red[3][warp] = gmax; }
  __syncthreads();
  for (int w = 0; w < EXP_WARPS; ++w) {
    cmin = fminf(cmin, s_red[0][w]); cmax = fmaxf(cmax, s_red[1][w]);
    gmin = fminf(gmin, s_red[2][w]); gmax = fmaxf(gmax, s_red[3][w]);
  }
  if (p.mode == 1) {
    // park this rank's share; the host all-reduces ws_minmax (min, max) across ranks before pass 2
    for (int c = tid; c < ncand; c += EXP_THREADS) {
      p.ws_cand[(size_t)q * EXP_MAXC_BIG + c] = cand[c];
      p.ws_cos[(size_t)q * EXP_MAXC_BIG + c] = cosv[c];
      p.ws_dist[(size_t)q * EXP_MAXC_BIG + c] = dist[c];
    }
    if (tid == 0) { p.ws_ncand[q] = ncand; p.ws_minmax[2 * q] = cmin; p.ws_minmax[2 * q + 1] = cmax; p.out_ncand[q] = ncand; p.out_status[q] = 0; }
    return;
  }
  if (p.mode == 2) { cmin = p.ws_minmax[2 * q]; cmax = p.ws_minmax[2 * q + 1]; }
  const bool c_flat = !(cmax > cmin), g_flat = !(gmax > gmin);   // "vmax <= vmin -> all 0.0" (:300)
  const float c_den = c_flat ? 1.0f : (cmax - cmin), g_den = g_flat ? 1.0f : (gmax - gmin);

  // ---- fused keys (the hash table is dead now; its memory becomes the key array) ------------------
  for (int c = tid; c < ncand; c += EXP_THREADS) {
    const float dn = c_flat ? 0.0f : (cosv[c] - cmin) / c_den;
    const float gn = g_flat ? 0.0f : (hop_decay(dist[c]) - gmin) / g_den;
    const long long node = (long long)cand[c];
    const bool own = node >= p.own_lo && node < p.own_hi;
    keys[c] = own ? make_key(p.alpha_graph * gn + p.alpha_dense * dn, cand[c]) : 0ull;   // ties on the fused score -> lower node id
  }
  if (tid == 0) { s_prefix = 0; s_remaining = p.k; s_nout = 0; s_done = 0; }
  for (int i = tid; i < p.kp2; i += EXP_THREADS) sk[i] = 0;
  __syncthreads();

  // ---- exact k-th key by radix select, then sort the winners ------------------------------------------
  u64 kth = 0;
  if (ncand > p.k) {
    for (int pass = 7; pass >= 0; --pass) {
      const int shift = pass * 8;
      for (int i = tid; i < 256; i += EXP_THREADS) hist[i] = 0;
      __syncthreads();
      const u64 prefix = s_prefix;
      const u64 hi_mask = (pass == 7) ? 0ull : (~0ull << (shift + 8));
      for (int base = warp * 32; base < ncand; base += EXP_THREADS) {
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
  for (int i = tid; i < ncand; i += EXP_THREADS) {
    const u64 key = keys[i];
    if (key >= kth && key != 0) {
      const int pos = atomicAdd(&s_nout, 1);
      if (pos < p.kp2) sk[pos] = key;
    }
  }
  __syncthreads();
  block_bitonic_desc(sk, p.kp2);

  // ---- emit: winner i is looked up in the candidate list by one warp (k is small) ------------------
  const int nwin = min(min(ncand, p.k), s_nout);
  for (int i = warp; i < p.k; i += EXP_WARPS) {
    const size_t o = (size_t)q * p.k + i;
    if (i < nwin) {
      const u64 key = sk[i];
      const u32 node = key_id(key);
      int slot = -1;
      for (int base = 0; base < ncand && slot < 0; base += 32) {
        const int c = base + lane;
        const u32 hit = __ballot_sync(kFullMask, c < ncand && cand[c] == node);
        if (hit) slot = base + __ffs(hit) - 1;
      }
      if (lane == 0) {
        const float cv = cosv[slot];
        const int hd = dist[slot];
        p.out_scores[o] = key_score(key);
        p.out_ids[o] = (long long)node;
        p.out_hops[o] = hd;
        p.out_gn[o] = g_flat ? 0.0f : (hop_decay(hd) - gmin) / g_den;
        p.out_dn[o] = c_flat ? 0.0f : (cv - cmin) / c_den;
        p.out_cos[o] = cv;
      }
    } else if (lane == 0) {
      p.out_scores[o] = -CUDART_INF_F; p.out_ids[o] = -1; p.out_hops[o] = -1; p.out_gn[o] = 0; p.out_dn[o] = 0; p.out_cos[o] = 0;
    }
  }
  if (tid == 0) { p.out_ncand[q] = ncand; p.out_status[q] = 0; }
}

int32_t expand_max_candidates() { return EXP_MAXC_BIG; }

template <typename T, int ROWS>
static int launch_expand_r(const ExpandParams& p, cudaStream_t st) {
  constexpr int SMALL = 4096;
  auto smem_of = [&](int maxc) { return (size_t)maxc * 8 + (size_t)maxc * 4 * 2 + maxc + (size_t)((p.D + 1) & ~1) * 4 + (size_t)p.kp2 * 8; };
  if (smem_of(EXP_MAXC_BIG) > 220 * 1024) { set_error("expand_rescore: D=%d / k=%d too large for shared memory", p.D, p.k); return B200R_EINVAL; }
  if (p.mode != 2) {
    B200R_CUDA_CHECK(cudaFuncSetAttribute(expand_rescore_kernel<T, SMALL, false, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_of(SMALL)));
    expand_rescore_kernel<T, SMALL, false, ROWS><<<p.B, EXP_THREADS, smem_of(SMALL), st>>>(p);
    count_launch();
    B200R_CUDA_CHECK(cudaFuncSetAttribute(expand_rescore_kernel<T, EXP_MAXC_BIG, true, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_of(EXP_MAXC_BIG)));
    expand_rescore_kernel<T, EXP_MAXC_BIG, true, ROWS><<<p.B, EXP_THREADS, smem_of(EXP_MAXC_BIG), st>>>(p);
    count_launch();
  } else {
    B200R_CUDA_CHECK(cudaFuncSetAttribute(expand_rescore_kernel<T, EXP_MAXC_BIG, false, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_of(EXP_MAXC_BIG)));
    expand_rescore_kernel<T, EXP_MAXC_BIG, false, ROWS><<<p.B, EXP_THREADS, smem_of(EXP_MAXC_BIG), st>>>(p);
    count_launch();
  }
  B200R_CUDA_CHECK(cudaGetLastError());
  return B200R_OK;
}

// 4 rows per warp measured best (2: 0.95x, 8: 0.87x -- fewer resident CTAs)
template <typename T>
static int launch_expand_t(const ExpandParams& p, cudaStream_t st) { return launch_expand_r<T, 4>(p, st); }

size_t expand_workspace_bytes(int B) {
  return align_up((size_t)B * EXP_MAXC_BIG * 4, 256) * 2 + align_up((size_t)B * EXP_MAXC_BIG, 256) + align_up((size_t)B * 4, 256);
}

int launch_expand_rescore(const Index* ix, const Graph* g, const void* Q, const int64_t* seeds, int B, int S, int window,
                          float alpha_dense, float alpha_graph, int k, float* out_scores, int64_t* out_ids, int32_t* out_hops,
                          float* out_gn, float* out_dn, float* out_cos, int32_t* out_ncand, int32_t* out_status, int mode,
                          int64_t own_lo, int64_t own_hi, void* ws, float* minmax, cudaStream_t st) {
  ExpandParams p;
  p.X = ix->X; p.Q = Q; p.row_mult = ix->row_mult;
  p.row_ptr = reinterpret_cast<const long long*>(g->row_ptr); p.col_idx = g->col_idx;
  p.seeds = reinterpret_cast<const long long*>(seeds);
  p.N = ix->N; p.V = g->V; p.id_base = ix->id_base;
  p.D = ix->D; p.B = B; p.S = S; p.window = window; p.k = k; p.metric = ix->metric;
  p.kp2 = 2; while (p.kp2 < k) p.kp2 <<= 1;
  p.alpha_dense = alpha_dense; p.alpha_graph = alpha_graph;
  p.out_scores = out_scores; p.out_ids = reinterpret_cast<long long*>(out_ids); p.out_hops = out_hops;
  p.out_gn = out_gn; p.out_dn = out_dn; p.out_cos = out_cos; p.out_ncand = out_ncand; p.out_status = out_status;
  p.mode = mode; p.own_lo = own_lo; p.own_hi = own_hi;
  p.ws_cand = nullptr; p.ws_cos = nullptr; p.ws_dist = nullptr; p.ws_ncand = nullptr; p.ws_minmax = minmax;
  if (mode != 0) {
    char* w = reinterpret_cast<char*>(ws);
    const size_t a4 = align_up((size_t)B * EXP_MAXC_BIG * 4, 256);
    p.ws_cand = reinterpret_cast<u32*>(w);
    p.ws_cos = reinterpret_cast<float*>(w + a4);
    p.ws_dist = reinterpret_cast<uint8_t*>(w + 2 * a4);
    p.ws_ncand = reinterpret_cast<int*>(w + 2 * a4 + align_up((size_t)B * EXP_MAXC_BIG, 256));
  }
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
