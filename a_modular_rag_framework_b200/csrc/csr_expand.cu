// csr_expand.cu -- K3: graph-enhanced second stage.  One CTA per query:
//   1. level-synchronous BFS from the seed passages over the CSR adjacency, depth <= window, with a
//      shared-memory hash set as the `visited` set (first arrival == minimum hop distance) and a
//      warp-level __match_any_sync pre-dedup so each distinct neighbour issues one atomic per warp;
//   2. re-score every reached node against the query (one warp per candidate row, 128-bit gathers);
//   3. per-query min/max of both channels, fused = a_g*minmax(decay(hop)) + a_d*minmax(cos);
//   4. exact top-k by (fused desc, id asc) with an in-smem bitonic sort.
//
// Replaces expand_qmatch_neighbors (reference app/modules/retrieval/graph_utils.py:58-129: BFS :109-127,
// decay :87-94), DenseReranker._cosine (retrieval_backend.py:192-197), _minmax_norm (:296-301) and the
// fusion / sort / slice of HybridRetrievalBackend.run (:353-372).
#include "common.cuh"
#include "internal.h"
#include <math_constants.h>

namespace b200r {

constexpr int EXP_THREADS = 256;
constexpr int EXP_WARPS = EXP_THREADS / 32;
constexpr int EXP_MAXC = 8192;        // candidates per query
constexpr int EXP_HASH = 16384;       // hash slots (load factor <= 0.5); 64 KB, later reused for the sort keys
constexpr u32 EXP_EMPTY = 0xffffffffu;

struct ExpandParams {
  const void* X;
  const void* Q;
  const float* row_mult;
  const long long* row_ptr;
  const int* col_idx;
  const long long* seeds;
  long long N, V, id_base;
  int D, B, S, window, k, metric;
  float alpha_dense, alpha_graph;
  float* out_scores; long long* out_ids; int* out_hops; float* out_gn; float* out_dn; float* out_cos;
  int* out_ncand; int* out_status;
};

__device__ __forceinline__ float hop_decay(int d) {
  // graph_utils.py:87-94
  if (d == 0) return 1.0f;
  if (d == 1) return 0.7f;
  if (d == 2) return 0.5f;
  return fmaxf(0.5f - 0.1f * (float)(d - 2), 0.1f);
}

// insert `v` into the visited set; returns true if this call inserted it
__device__ __forceinline__ bool hash_insert(u32* table, u32 v) {
  u32 h = (v * 2654435761u) & (EXP_HASH - 1);
  while (true) {
    const u32 prev = atomicCAS(&table[h], EXP_EMPTY, v);
    if (prev == EXP_EMPTY) return true;
    if (prev == v) return false;
    h = (h + 1) & (EXP_HASH - 1);
  }
}

template <typename T>
__global__ void __launch_bounds__(EXP_THREADS) expand_rescore_kernel(const ExpandParams p) {
  typedef typename QueryType<T>::type QT;
  constexpr int VEC = ElemTraits<T>::kVec;
  extern __shared__ uint8_t smem_raw[];
  u64* keys = reinterpret_cast<u64*>(smem_raw);                    // [EXP_MAXC]  (aliases the hash table)
  u32* table = reinterpret_cast<u32*>(smem_raw);                   // [EXP_HASH] u32 == 64 KB
  u32* cand = reinterpret_cast<u32*>(smem_raw + EXP_MAXC * 8);     // [EXP_MAXC] node ids
  float* cosv = reinterpret_cast<float*>(cand + EXP_MAXC);         // [EXP_MAXC]
  uint8_t* dist = reinterpret_cast<uint8_t*>(cosv + EXP_MAXC);     // [EXP_MAXC]
  float* qs = reinterpret_cast<float*>(dist + EXP_MAXC);           // [D]
  __shared__ int s_count, s_overflow;
  __shared__ float s_red[4][EXP_WARPS];
  __shared__ double s_qn[EXP_WARPS];

  const int q = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int i = tid; i < EXP_HASH; i += EXP_THREADS) table[i] = EXP_EMPTY;
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
  const float qinv = (p.metric == B200R_COS) ? (qn_tot > 0.0 ? (float)(1.0 / sqrt(qn_tot)) : 0.0f) : 1.0f;

  // ---- level 0: seeds ---------------------------------------------------------------------------
  for (int i = tid; i < p.S; i += EXP_THREADS) {
    const long long s = p.seeds[(size_t)q * p.S + i] - p.id_base;
    if (s >= 0 && s < p.V) {
      if (hash_insert(table, (u32)s)) {
        const int pos = atomicAdd(&s_count, 1);
        if (pos < EXP_MAXC) { cand[pos] = (u32)s; dist[pos] = 0; } else s_overflow = 1;
      }
    }
  }
  __syncthreads();

  // ---- levels 1..window: one warp per frontier node, lanes stride over its neighbours -------------
  int level_begin = 0;
  for (int d = 1; d <= p.window; ++d) {
    const int level_end = min(s_count, EXP_MAXC);
    __syncthreads();                      // everyone has read s_count before the level appends to it
    if (level_end == level_begin) break;  // block-uniform
    for (int f = level_begin + warp; f < level_end; f += EXP_WARPS) {
      const u32 u = cand[f];
      const long long e0 = p.row_ptr[u], e1 = p.row_ptr[u + 1];
      for (long long base = e0; base < e1; base += 32) {
        const long long e = base + lane;
        const bool in = e < e1;
        const u32 v = in ? (u32)__ldg(p.col_idx + e) : EXP_EMPTY;
        // warp-level pre-dedup: one lane per distinct neighbour id goes to the hash set
        const u32 peers = __match_any_sync(kFullMask, v);
        const bool leader = in && (lane == __ffs(peers) - 1);
        if (leader && !*(volatile int*)&s_overflow && hash_insert(table, v)) {
          const int pos = atomicAdd(&s_count, 1);
          if (pos < EXP_MAXC) { cand[pos] = v; dist[pos] = (uint8_t)min(d, 255); } else s_overflow = 1;
        }
      }
    }
    __syncthreads();
    level_begin = level_end;
  }
  __syncthreads();
  const int ncand = min(s_count, EXP_MAXC);
  if (s_overflow) {  // block-uniform
    if (tid == 0) { p.out_status[q] = B200R_EOVERFLOW; p.out_ncand[q] = s_count; }
    for (int i = tid; i < p.k; i += EXP_THREADS) {
      const size_t o = (size_t)q * p.k + i;
      p.out_scores[o] = -CUDART_INF_F; p.out_ids[o] = -1; p.out_hops[o] = -1; p.out_gn[o] = 0; p.out_dn[o] = 0; p.out_cos[o] = 0;
    }
    return;
  }

  // ---- re-score: one warp per candidate row -------------------------------------------------------
  const uint8_t* Xb = reinterpret_cast<const uint8_t*>(p.X);
  const int chunks = p.D / VEC;
  for (int c = warp; c < ncand; c += EXP_WARPS) {
    const u32 row = cand[c];
    const bool have = (long long)row < p.N;
    float acc = 0.0f;
    if (have) {
      const uint8_t* rp = Xb + (size_t)row * p.D * sizeof(T);
      for (int chunk = lane; chunk < chunks; chunk += 32) {
        const uint4 xv = ld_stream_16B(rp + (size_t)chunk * 16);
        float xf[VEC];
        ElemTraits<T>::unpack(xv, xf);
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc = fmaf(xf[e], qs[chunk * VEC + e], acc);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFullMask, acc, o);
    if (lane == 0) {
      const float m = have ? (p.row_mult ? p.row_mult[row] : 1.0f) : 0.0f;
      cosv[c] = (acc * m) * qinv + 0.0f;
    }
  }
  __syncthreads();

  // ---- min/max of both channels (retrieval_backend.py:296-301) ----------------------------------
  float cmin = CUDART_INF_F, cmax = -CUDART_INF_F, gmin = CUDART_INF_F, gmax = -CUDART_INF_F;
  for (int c = tid; c < ncand; c += EXP_THREADS) {
    const float cv = cosv[c], gv = hop_decay(dist[c]);
    cmin = fminf(cmin, cv); cmax = fmaxf(cmax, cv); gmin = fminf(gmin, gv); gmax = fmaxf(gmax, gv);
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
  const bool c_flat = !(cmax > cmin), g_flat = !(gmax > gmin);   // "vmax <= vmin -> all 0.0" (:300)
  const float c_den = c_flat ? 1.0f : (cmax - cmin), g_den = g_flat ? 1.0f : (gmax - gmin);

  // ---- fused keys + sort (the hash table is dead now; its memory becomes the key array) -----------
  int np2 = 2;
  while (np2 < ncand) np2 <<= 1;
  __syncthreads();
  for (int c = tid; c < np2; c += EXP_THREADS) {
    u64 key = 0;
    if (c < ncand) {
      const float dn = c_flat ? 0.0f : (cosv[c] - cmin) / c_den;
      const float gn = g_flat ? 0.0f : (hop_decay(dist[c]) - gmin) / g_den;
      const float fused = p.alpha_graph * gn + p.alpha_dense * dn;
      key = make_key(fused, cand[c]);   // ties on the fused score go to the lower node id
    }
    keys[c] = key;
  }
  block_bitonic_desc(keys, np2);

  // ---- emit: winner i is looked up in the candidate list by one warp (k is small) ------------------
  for (int i = warp; i < p.k; i += EXP_WARPS) {
    const size_t o = (size_t)q * p.k + i;
    if (i < ncand) {
      const u64 key = keys[i];
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
        p.out_ids[o] = p.id_base + (long long)node;
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

int32_t expand_max_candidates() { return EXP_MAXC; }

int launch_expand_rescore(const Index* ix, const Graph* g, const void* Q, const int64_t* seeds, int B, int S, int window,
                          float alpha_dense, float alpha_graph, int k, float* out_scores, int64_t* out_ids, int32_t* out_hops,
                          float* out_gn, float* out_dn, float* out_cos, int32_t* out_ncand, int32_t* out_status, cudaStream_t st) {
  ExpandParams p;
  p.X = ix->X; p.Q = Q; p.row_mult = (ix->metric == B200R_COS || ix->dtype == B200R_I8) ? ix->row_mult : nullptr;
  p.row_ptr = reinterpret_cast<const long long*>(g->row_ptr); p.col_idx = g->col_idx;
  p.seeds = reinterpret_cast<const long long*>(seeds);
  p.N = ix->N; p.V = g->V; p.id_base = ix->id_base;
  p.D = ix->D; p.B = B; p.S = S; p.window = window; p.k = k; p.metric = ix->metric;
  p.alpha_dense = alpha_dense; p.alpha_graph = alpha_graph;
  p.out_scores = out_scores; p.out_ids = reinterpret_cast<long long*>(out_ids); p.out_hops = out_hops;
  p.out_gn = out_gn; p.out_dn = out_dn; p.out_cos = out_cos; p.out_ncand = out_ncand; p.out_status = out_status;
  const size_t smem = (size_t)EXP_MAXC * 8 + EXP_MAXC * 4 + EXP_MAXC * 4 + EXP_MAXC + (size_t)ix->D * 4;
  if (smem > 220 * 1024) { set_error("expand_rescore: D=%d too large", ix->D); return B200R_EINVAL; }
#define B200R_LAUNCH_EXP(T)                                                                                        \
  do {                                                                                                             \
    B200R_CUDA_CHECK(cudaFuncSetAttribute(expand_rescore_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    expand_rescore_kernel<T><<<B, EXP_THREADS, smem, st>>>(p);                                                     \
  } while (0)
  switch (ix->dtype) {
    case B200R_BF16: B200R_LAUNCH_EXP(__nv_bfloat16); break;
    case B200R_F16: B200R_LAUNCH_EXP(__half); break;
    case B200R_F32: B200R_LAUNCH_EXP(float); break;
    case B200R_I8: B200R_LAUNCH_EXP(int8_t); break;
    default: set_error("bad dtype"); return B200R_EINVAL;
  }
#undef B200R_LAUNCH_EXP
  count_launch();
  B200R_CUDA_CHECK(cudaGetLastError());
  return B200R_OK;
}

}  // namespace b200r
