// mmr.cu -- optional post-filter: MMR diversification of each query's candidate list on the device.
//
// Replaces mmr_diversify (reference app/utils/similarity.py:33-62): greedy selection of top_k items,
//   first pick:   val = score                                   (the reference does not apply lambda to the first pick)
//   later picks:  val = lambda * score - (1 - lambda) * max(0, max_{s in selected} cos(c, s))
// `val > best_val` is strict, so ties go to the candidate that comes FIRST in the input list.  The candidate vectors are
// the corpus rows behind the ids (a candidate whose row is not on this device has no vector: similarity 0, like a None
// vector in the reference).  One CTA per query; each pick costs one pass of P row dot products.
#include "common.cuh"
#include "internal.h"
#include <math_constants.h>

namespace b200r {

constexpr int MMR_THREADS = 256;
constexpr int MMR_WARPS = MMR_THREADS / 32;

struct MmrParams {
  const void* X;
  long long N, id_base;
  int D, B, P, top_k;
  float lambda;
  const long long* ids;      // [B][P] global ids, < 0 = padding
  const float* scores;       // [B][P]
  long long* out_ids;        // [B][top_k]
  float* out_scores;         // [B][top_k]
  int* out_pos;              // [B][top_k] position of every pick in the input list (-1 = padding)
};

template <typename T>
__global__ void __launch_bounds__(MMR_THREADS) mmr_kernel(const MmrParams p) {
  constexpr int VEC = ElemTraits<T>::kVec;
  extern __shared__ __align__(16) float smem_f[];
  float* cur = smem_f;                                  // [D] the vector picked last
  float* sim_max = cur + ((p.D + 3) & ~3);              // [P]
  float* inv_norm = sim_max + p.P;                      // [P] 1/|x_c| (0 = no vector)
  int* taken = reinterpret_cast<int*>(inv_norm + p.P);  // [P]
  __shared__ float s_val[MMR_WARPS];
  __shared__ int s_idx[MMR_WARPS];
  __shared__ int s_pick;
  const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long* ids = p.ids + (size_t)q * p.P;
  const float* sc = p.scores + (size_t)q * p.P;
  const uint8_t* Xb = reinterpret_cast<const uint8_t*>(p.X);
  const size_t row_bytes = (size_t)p.D * sizeof(T);
  const int chunks = p.D / VEC;

  // norms of the candidate vectors (one warp per candidate)
  for (int c = warp; c < p.P; c += MMR_WARPS) {
    const long long row = ids[c] - p.id_base;
    float acc = 0.0f;
    if (ids[c] >= 0 && row >= 0 && row < p.N) {
      for (int ch = lane; ch < chunks; ch += 32) {
        float f[VEC];
        ElemTraits<T>::unpack(ld_stream_16B(Xb + (size_t)row * row_bytes + (size_t)ch * 16), f);
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc = fmaf(f[e], f[e], acc);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFullMask, acc, o);
    if (lane == 0) { inv_norm[c] = acc > 0.0f ? rsqrtf(acc) : 0.0f; sim_max[c] = 0.0f; taken[c] = ids[c] < 0 ? 1 : 0; }
  }
  __syncthreads();

  for (int t = 0; t < p.top_k; ++t) {
    // ---- argmax of val over the remaining candidates, ties to the lowest position ----
    float best = -CUDART_INF_F; int bi = -1;
    for (int c = tid; c < p.P; c += MMR_THREADS) {
      if (taken[c]) continue;
      const float val = (t == 0) ? sc[c] : p.lambda * sc[c] - (1.0f - p.lambda) * sim_max[c];
      if (val > best || (val == best && (bi < 0 || c < bi))) { best = val; bi = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(kFullMask, best, o);
      const int oi = __shfl_xor_sync(kFullMask, bi, o);
      if (oi >= 0 && (bi < 0 || ov > best || (ov == best && oi < bi))) { best = ov; bi = oi; }
    }
    if (lane == 0) { s_val[warp] = best; s_idx[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      float bv = -CUDART_INF_F; int b = -1;
      for (int w = 0; w < MMR_WARPS; ++w)
        if (s_idx[w] >= 0 && (b < 0 || s_val[w] > bv || (s_val[w] == bv && s_idx[w] < b))) { bv = s_val[w]; b = s_idx[w]; }
      s_pick = b;
      const size_t o = (size_t)q * p.top_k + t;
      p.out_ids[o] = b >= 0 ? ids[b] : -1;
      p.out_scores[o] = b >= 0 ? sc[b] : -CUDART_INF_F;
      p.out_pos[o] = b;
      if (b >= 0) taken[b] = 1;
    }
    __syncthreads();
    const int pick = s_pick;
    if (pick < 0) continue;                               // block-uniform: the list is exhausted, the rest is padding
    if (t + 1 == p.top_k) break;
    // ---- the picked vector -> shared memory, then its cosine with every remaining candidate ----
    const long long prow = ids[pick] - p.id_base;
    const bool have = inv_norm[pick] > 0.0f;
    if (have)
      for (int ch = tid; ch < chunks; ch += MMR_THREADS) {
        float f[VEC];
        ElemTraits<T>::unpack(ld_stream_16B(Xb + (size_t)prow * row_bytes + (size_t)ch * 16), f);
#pragma unroll
        for (int e = 0; e < VEC; ++e) cur[ch * VEC + e] = f[e];
      }
    __syncthreads();
    if (have) {
      const float pin = inv_norm[pick];
      for (int c = warp; c < p.P; c += MMR_WARPS) {
        if (taken[c] || inv_norm[c] == 0.0f) continue;    // warp-uniform
        const long long row = ids[c] - p.id_base;
        float acc = 0.0f;
        for (int ch = lane; ch < chunks; ch += 32) {
          float f[VEC];
          ElemTraits<T>::unpack(ld_stream_16B(Xb + (size_t)row * row_bytes + (size_t)ch * 16), f);
#pragma unroll
          for (int e = 0; e < VEC; ++e) acc = fmaf(f[e], cur[ch * VEC + e], acc);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFullMask, acc, o);
        if (lane == 0) sim_max[c] = fmaxf(sim_max[c], acc * pin * inv_norm[c]);
      }
    }
    __syncthreads();
  }
}

int launch_mmr(const Index* ix, const int64_t* ids, const float* scores, int B, int P, int top_k, float lambda, int64_t* out_ids,
               float* out_scores, int32_t* out_pos, cudaStream_t st) {
  ProfScope _ps(st, "mmr");
  MmrParams p;
  p.X = ix->X; p.N = ix->N; p.id_base = ix->id_base; p.D = ix->D; p.B = B; p.P = P; p.top_k = top_k; p.lambda = lambda;
  p.ids = reinterpret_cast<const long long*>(ids); p.scores = scores;
  p.out_ids = reinterpret_cast<long long*>(out_ids); p.out_scores = out_scores; p.out_pos = out_pos;
  const size_t smem = (size_t)((ix->D + 3) & ~3) * 4 + (size_t)P * 12;
  if (smem > 200 * 1024) { set_error("mmr: D=%d / pool=%d too large for shared memory", ix->D, P); return B200R_EINVAL; }
#define B200R_MMR_LAUNCH(T)                                                                                              \
  do {                                                                                                                   \
    B200R_CUDA_CHECK(cudaFuncSetAttribute(mmr_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
    mmr_kernel<T><<<B, MMR_THREADS, smem, st>>>(p);                                                                      \
  } while (0)
  switch (ix->dtype) {
    case B200R_BF16: B200R_MMR_LAUNCH(__nv_bfloat16); break;
    case B200R_F16: B200R_MMR_LAUNCH(__half); break;
    case B200R_F32: B200R_MMR_LAUNCH(float); break;
    case B200R_I8: B200R_MMR_LAUNCH(int8_t); break;
    default: set_error("bad dtype"); return B200R_EINVAL;
  }
#undef B200R_MMR_LAUNCH
  count_launch();
  B200R_CUDA_CHECK(cudaGetLastError());
  return B200R_OK;
}

}  // namespace b200r
