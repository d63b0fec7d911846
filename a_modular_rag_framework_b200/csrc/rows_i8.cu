// rows_i8.cu -- K1c: int8 corpora on the tensor cores (tcgen05 kind::i8), 16 / 32 / 64 queries per corpus pass.
//
// Same contract as K1/K2 (exact cosine / IP top-k; reference retrieval_backend.py:192-197, 371-372).  The CUDA-core
// scans cannot keep up with HBM for int8 rows once there is more than one query (DP4A issues at ~16 lanes/clk/SM, the
// PRMT/FADD/FFMA unpack is slower still), so the dot products go to the int8 tensor pipe, with the operands laid out
// for streaming:
//   * the 128 corpus rows of a tile are the MMA's M (TMEM lanes); tiles of [128 rows x 128 bytes] arrive by TMA
//     (SWIZZLE_128B) in a ring of up to 12 stages that carries corpus data only;
//   * the queries are quantised ONCE per search to 23-bit fixed point relative to their largest component
//     (|q - q~| <= 2^-22 max|q|) and split into three balanced signed-byte limbs; the 3 x NQ limb rows of a group of
//     NQ queries are the MMA's N (48 / 96 / 192 TMEM columns) and stay RESIDENT in shared memory;
//   * accumulation is exact int32; the epilogue thread (= TMEM lane = one corpus row) recombines the limbs of each
//     query in fp32 (a2*65536 + a1*256 + a0), applies the query scale and the row multiplier and tests the scores, 16
//     queries at a time, against the per-query thresholds.  Survivors are appended per (warp, query) with a ballot; lists, thresholds,
//     phases and pruning exactly as in K1.
// Roofline: HBM, N*D bytes per group of queries.
#include "common.cuh"
#include "internal.h"
#include "tc_ptx.cuh"
#include <math_constants.h>
#include <cstring>

namespace b200r {

constexpr int RI_ROWS = 128;          // corpus rows per tile
constexpr int RI_BK = 128;            // int8 elements per K slice = one 128-byte swizzle row
constexpr int RI_STAGE_BYTES = RI_ROWS * RI_BK;   // 16 KB
constexpr int RI_THREADS = 192;
constexpr int RI_ACC_MAX = 4;         // accumulator stages in TMEM (2 when a group has 64 queries: 192 columns each)
// NQ = queries per corpus pass (16 / 32 / 64): 3*NQ limb rows = the MMA's N = TMEM columns of one accumulator
template <int NQ> struct RiGeo {
  static constexpr int NCOLS = 3 * NQ;
  static constexpr int ACC_STRIDE = NCOLS <= 64 ? 64 : (NCOLS <= 128 ? 128 : 256);   // TMEM columns between accumulators
  static constexpr int ACC = (512 / ACC_STRIDE) < RI_ACC_MAX ? (512 / ACC_STRIDE) : RI_ACC_MAX;
  static constexpr int Q_KB_BYTES = NCOLS * RI_BK;    // one K slice of the resident limb block (whole 1 KB swizzle atoms)
};

struct RowsI8Params {
  int B, k, D, num_kb, stages;
  int groups;              // query groups ganged in this launch: CTA b serves group b % groups (queries [g*NQ, g*NQ + NQ)) and is
                           // CTA b / groups of the gridDim.x / groups that share that group's units.  The CTAs of one gang index
                           // stream the same corpus tiles at the same time, so HBM sees each tile once and the rest hit in L2
  long long N;
  int tile_begin, tile_end, tiles_per_chunk, num_units;
  int cap, cap_soft, Bpad;
  const float* row_mult;
  const float* qscale;     // [groups * NQ] q_j ~= qscale[j] * (l2*65536 + l1*256 + l0)
  u64* lists;
  int* counts;
  u32* gthr;
  u32 idesc;
};

// ---- query quantisation: f32 [B, D] -> limb planes int8 [groups][3][nq][D] (one 2-D tensor map over all of them), then
// the scales f32 [groups][nq] -------------------------------------------------------------------------------------------
__global__ void quantize_queries_i8_kernel(const float* __restrict__ Q, int B, int D, int nq, int8_t* __restrict__ qws,
                                           size_t scale_off) {
  const int g = blockIdx.x / nq, j = blockIdx.x % nq;           // group of nq queries, slot inside it
  const int query = g * nq + j;
  int8_t* limbs = qws + (size_t)g * 3 * nq * D;
  float* scale = reinterpret_cast<float*>(qws + scale_off) + (size_t)g * nq;
  __shared__ float s_max[32];
  float m = 0.0f;
  if (query < B)
    for (int d = threadIdx.x; d < D; d += blockDim.x) m = fmaxf(m, fabsf(Q[(size_t)query * D + d]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(kFullMask, m, o));
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = m;
  __syncthreads();
  float qmax = 0.0f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) qmax = fmaxf(qmax, s_max[w]);
  const float inv = qmax > 0.0f ? 4194304.0f / qmax : 0.0f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const int qi = (query < B) ? __float2int_rn(Q[(size_t)query * D + d] * inv) : 0;      // |qi| <= 2^22
    const int l0 = ((qi + 128) & 255) - 128;
    const int q1 = (qi - l0) >> 8;
    const int l1 = ((q1 + 128) & 255) - 128;
    const int l2 = (q1 - l1) >> 8;                                                       // |l2| <= 64
    limbs[((size_t)0 * nq + j) * D + d] = (int8_t)l0;
    limbs[((size_t)1 * nq + j) * D + d] = (int8_t)l1;
    limbs[((size_t)2 * nq + j) * D + d] = (int8_t)l2;
  }
  if (threadIdx.x == 0) scale[j] = qmax * (1.0f / 4194304.0f);
}

template <int NQ>
__global__ void __launch_bounds__(RI_THREADS, 1)
rows_i8_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_x, const RowsI8Params p) {
  constexpr int RI_NQ = NQ;
  constexpr int RI_ACC = RiGeo<NQ>::ACC;
  constexpr int RI_ACC_STRIDE = RiGeo<NQ>::ACC_STRIDE;
  constexpr int RI_Q_KB_BYTES = RiGeo<NQ>::Q_KB_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const u32 raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  uint8_t* smem_q = smem;                                       // [num_kb][3*NQ x 128 B]
  uint8_t* smem_x = smem + (size_t)p.num_kb * RI_Q_KB_BYTES;    // [stages][128 x 128 B]
  uint8_t* tail = smem_x + (size_t)p.stages * RI_STAGE_BYTES;
  u32* smem_hist = reinterpret_cast<u32*>(tail);                // [4][256]
  float* thr_s = reinterpret_cast<float*>(smem_hist + 4 * 256); // [4][NQ]
  u64* kth_s = reinterpret_cast<u64*>(thr_s + 4 * RI_NQ);       // [4][NQ]
  int* cnt_s = reinterpret_cast<int*>(kth_s + 4 * RI_NQ);       // [4][NQ]
  float* qscale_s = reinterpret_cast<float*>(cnt_s + 4 * RI_NQ);   // [NQ]
  u64* bars = reinterpret_cast<u64*>(qscale_s + RI_NQ);
  u64* full_bar = bars;                      // [stages] (<= 12)
  u64* empty_bar = bars + 12;                // [stages]
  u64* tmem_full_bar = bars + 24;                // [RI_ACC]
  u64* tmem_empty_bar = bars + 24 + RI_ACC_MAX;  // [RI_ACC]
  u64* q_bar = bars + 24 + 2 * RI_ACC_MAX;
  u32* tmem_ptr_smem = reinterpret_cast<u32*>(bars + 25 + 2 * RI_ACC_MAX);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stages = p.stages;
  const int group = blockIdx.x % p.groups;          // which NQ queries
  const int u_first = blockIdx.x / p.groups;        // first unit of this CTA, then every u_step-th
  const int u_step = gridDim.x / p.groups;
  const int q0 = group * RI_NQ;                     // list / threshold / count slots q0 + j
  constexpr u32 TMEM_COLS = RI_ACC * RI_ACC_STRIDE;   // 256 (NQ = 16) or 512

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_q);
    prefetch_tmap(&tmap_x);
    for (int i = 0; i < stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < RI_ACC; ++i) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], 4); }
    mbar_init(q_bar, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < RI_NQ) qscale_s[threadIdx.x] = p.qscale[q0 + threadIdx.x];   // RI_NQ <= 64 < RI_THREADS
  if (warp_idx == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const u32 tmem_base = *tmem_ptr_smem;

  if (warp_idx == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      // the limb block, once: num_kb slices of [3*NQ limb rows x 128 bytes] (columns >= D are zero-filled by TMA)
      mbar_expect_tx(q_bar, (u32)(p.num_kb * RI_Q_KB_BYTES));
      for (int kb = 0; kb < p.num_kb; ++kb) tma_load_2d(smem_q + (size_t)kb * RI_Q_KB_BYTES, &tmap_q, q_bar, kb * RI_BK, group * RiGeo<NQ>::NCOLS, kEvictLast);
      int stage = 0; u32 phase = 0;
      for (int u = u_first; u < p.num_units; u += u_step) {
        const int t0 = p.tile_begin + u * p.tiles_per_chunk;
        const int t1 = min(p.tile_end, t0 + p.tiles_per_chunk);
        for (int t = t0; t < t1; ++t) {
          for (int kb = 0; kb < p.num_kb; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_expect_tx(&full_bar[stage], RI_STAGE_BYTES);
            tma_load_2d(smem_x + (size_t)stage * RI_STAGE_BYTES, &tmap_x, &full_bar[stage], kb * RI_BK, t * RI_ROWS, kEvictFirst);
            if (++stage == stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================================== MMA issuer =======================================
    if (lane == 0) {
      mbar_wait(q_bar, 0);
      int stage = 0; u32 phase = 0;
      u32 tile_iter = 0;
      for (int u = u_first; u < p.num_units; u += u_step) {
        const int t0 = p.tile_begin + u * p.tiles_per_chunk;
        const int t1 = min(p.tile_end, t0 + p.tiles_per_chunk);
        for (int t = t0; t < t1; ++t, ++tile_iter) {
          const u32 acc = tile_iter % RI_ACC;
          mbar_wait(&tmem_empty_bar[acc], ((tile_iter / RI_ACC) & 1) ^ 1);
          tc_fence_after();
          const u32 tmem_d = tmem_base + acc * RI_ACC_STRIDE;
          for (int kb = 0; kb < p.num_kb; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const u64 desc_a = make_smem_desc(smem_u32(smem_x + (size_t)stage * RI_STAGE_BYTES));   // corpus rows = M
            const u64 desc_b = make_smem_desc(smem_u32(smem_q + (size_t)kb * RI_Q_KB_BYTES));       // limb rows = N
#pragma unroll
            for (int kk = 0; kk < RI_BK / 32; ++kk)      // K = 32 int8 = 32 bytes per instruction
              umma_i8(tmem_d, desc_a + (u64)(kk * 2), desc_b + (u64)(kk * 2), p.idesc, (kb | kk) != 0 ? 1u : 0u);
            umma_commit(&empty_bar[stage]);
            if (kb == p.num_kb - 1) umma_commit(&tmem_full_bar[acc]);
            if (++stage == stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else {
    // ===================================== epilogue =========================================
    const int quarter = warp_idx & 3;          // TMEM lane quarter this warp may access == rows 32*quarter.. of the tile
    const int ew = warp_idx - 2;               // 0..3: index of this warp's lists / thresholds
    u32* hist = smem_hist + ew * 256;
    float* thr = thr_s + ew * RI_NQ;
    u64* kth = kth_s + ew * RI_NQ;
    int* cnt = cnt_s + ew * RI_NQ;
    u32 tile_iter = 0;
    for (int u = u_first; u < p.num_units; u += u_step) {
      const int t0 = p.tile_begin + u * p.tiles_per_chunk;
      const int t1 = min(p.tile_end, t0 + p.tiles_per_chunk);
      if (t0 >= t1) continue;
      const size_t seg = (size_t)u * 4 + ew;               // this warp's list segment for the unit's row chunk
      for (int j = lane; j < RI_NQ; j += 32) {
        const bool valid = q0 + j < p.B;
        thr[j] = valid ? ordered_to_f32(__ldcg(p.gthr + q0 + j)) : __int_as_float(0x7f800000);
        kth[j] = 0;
        cnt[j] = valid ? __ldcg(p.counts + seg * p.Bpad + q0 + j) : 0;
      }
      __syncwarp();
      for (int t = t0; t < t1; ++t, ++tile_iter) {
        const u32 acc = tile_iter % RI_ACC;
        const long long row = (long long)t * RI_ROWS + quarter * 32 + lane;
        const float m = (row < p.N) ? (p.row_mult ? __ldg(p.row_mult + row) : 1.0f) : __int_as_float(0x7fc00000);
        mbar_wait(&tmem_full_bar[acc], (tile_iter / RI_ACC) & 1);
        tc_fence_after();
        const u32 taddr = tmem_base + ((u32)(quarter * 32) << 16) + acc * RI_ACC_STRIDE;
        // 16 queries at a time: their three limb accumulators -> fp32 score -> threshold test
#pragma unroll 1
        for (int c = 0; c < RI_NQ / 16; ++c) {
          u32 a0[16], a1[16], a2[16];
          tmem_ld16(taddr + c * 16, a0);
          tmem_ld16(taddr + RI_NQ + c * 16, a1);
          tmem_ld16(taddr + 2 * RI_NQ + c * 16, a2);
          tmem_ld_wait();
          if (c == RI_NQ / 16 - 1) {          // everything of this accumulator has been read
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
          }
          float s[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            // exact integer limb sums -> fp32 (the same recombination as the DP4A scan); NaN m for rows beyond N never passes
            const float v = fmaf((float)(int)a2[j], 65536.0f, fmaf((float)(int)a1[j], 256.0f, (float)(int)a0[j]));
            s[j] = (v * qscale_s[c * 16 + j]) * m;
          }
          bool any = false;
#pragma unroll
          for (int j = 0; j < 16; ++j) any |= (s[j] >= thr[c * 16 + j]);
          if (__any_sync(kFullMask, any)) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int jj = c * 16 + j;
              const float tj = thr[jj];
              const u32 b = __ballot_sync(kFullMask, s[j] >= tj);
              if (b) {                                          // warp-uniform
                const u64 key = make_key(s[j], (u32)row);
                const u32 b2 = __ballot_sync(kFullMask, (s[j] >= tj) && key > kth[jj]);
                if (b2) {
                  const int base = cnt[jj];
                  u64* lp = p.lists + (seg * p.Bpad + q0 + jj) * (size_t)p.cap;
                  if ((b2 >> lane) & 1u) __stcg(lp + base + __popc(b2 & ((1u << lane) - 1u)), key);
                  const int n = base + __popc(b2);
                  __syncwarp();
                  if (n >= p.cap_soft) {
                    const u64 kk = warp_compact(lp, n, p.k, hist);
                    if (lane == 0) {
                      kth[jj] = kk; thr[jj] = fmaxf(tj, key_score(kk)); cnt[jj] = p.k;
                      atomicMax(p.gthr + q0 + jj, (u32)(kk >> 32));
                    }
                  } else if (lane == 0) {
                    cnt[jj] = n;
                  }
                  __syncwarp();
                }
              }
            }
          }
        }
        // thresholds published by other CTAs / warps
        for (int j = lane; j < RI_NQ; j += 32)
          if (q0 + j < p.B) thr[j] = fmaxf(thr[j], ordered_to_f32(__ldcg(p.gthr + q0 + j)));
        __syncwarp();
      }
      for (int j = lane; j < RI_NQ; j += 32)
        if (q0 + j < p.B) p.counts[seg * p.Bpad + q0 + j] = cnt[j];
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---- host side -----------------------------------------------------------------------------------------
static size_t rows_i8_smem_bytes(int nq, int num_kb, int stages) {
  return 1024 + (size_t)num_kb * 3 * nq * RI_BK + (size_t)stages * RI_STAGE_BYTES + 4 * 256 * 4 + (size_t)4 * nq * (4 + 8 + 4) +
         (size_t)nq * 4 + (26 + 2 * RI_ACC_MAX) * 8 + 16;
}
static int rows_i8_stages_for(int nq, int D) {
  const int num_kb = (D + RI_BK - 1) / RI_BK;
  int stages = 12;
  while (stages >= 4 && rows_i8_smem_bytes(nq, num_kb, stages) > 227 * 1024) --stages;
  return stages >= 4 ? stages : 0;
}

int rows_i8_tile_rows() { return RI_ROWS; }

// queries per corpus pass for a batch of B: the smallest of 16 / 32 / 64 that holds B and whose resident limb block
// leaves room for a >= 4-stage ring; 0 = the kernel does not apply (B200R_ROWS_I8=0 switches it off)
int rows_i8_nq(int B, int D) {
  static const bool off = [] { const char* e = getenv("B200R_ROWS_I8"); return e && e[0] == '0'; }();
  static const int max_nq = [] { const char* e = getenv("B200R_ROWS_I8_NQ"); const int v = e ? atoi(e) : 64; return (v == 16 || v == 32) ? v : 64; }();
  if (off) return 0;
  int best = 0;
  for (int nq = 16; nq <= max_nq; nq *= 2) {
    if (!rows_i8_stages_for(nq, D)) break;
    best = nq;
    if (nq >= B) break;
  }
  return best;
}

// workspace for the quantised queries: limb planes [groups][3][nq][D] int8 (contiguous: ONE tensor map covers every group),
// then groups * nq scales
static size_t rows_i8_limb_bytes(int groups, int nq, int D) { return align_up((size_t)groups * 3 * nq * D, 256); }
size_t rows_i8_query_bytes(int B, int nq, int D) {
  const int groups = (B + nq - 1) / nq;
  return rows_i8_limb_bytes(groups, nq, D) + align_up((size_t)groups * nq * 4, 256);
}

int launch_quantize_queries_i8(const void* Q, int B, int nq, int D, void* qws, cudaStream_t st) {
  ProfScope _ps(st, "quantize_queries_i8");
  const int groups = (B + nq - 1) / nq;
  quantize_queries_i8_kernel<<<groups * nq, 256, 0, st>>>(reinterpret_cast<const float*>(Q), B, D, nq, reinterpret_cast<int8_t*>(qws),
                                                        rows_i8_limb_bytes(groups, nq, D));
  count_launch();
  B200R_CUDA_CHECK(cudaGetLastError());
  return B200R_OK;
}

// one map over the limb rows of every group: box = the 3*nq limb rows of one group x one 128-byte K slice
int encode_query_tmap_i8(const void* qws, int groups, int nq, int D, CUtensorMap* tmap_q) {
  return encode_tmap_2d(tmap_q, B200R_I8, qws, (uint64_t)D, (uint64_t)groups * 3 * nq, RI_BK, (uint32_t)(3 * nq));
}

template <int NQ>
static int launch_rows_i8_nq(const Index* ix, const CUtensorMap* tmap_q, const void* qws, const SearchPlan& pl,
                             const GemmPhase& ph, void* ws, cudaStream_t st) {
  RowsI8Params p;
  p.B = pl.B; p.k = pl.k; p.D = ix->D; p.N = ix->N;
  p.groups = pl.rows_i8;
  p.num_kb = (ix->D + RI_BK - 1) / RI_BK;
  p.stages = rows_i8_stages_for(NQ, ix->D);
  if (!p.stages) { set_error("int8 tensor path: D=%d not supported with %d queries per pass", ix->D, NQ); return B200R_EINVAL; }
  p.tile_begin = ph.tile_begin; p.tile_end = ph.tile_end; p.tiles_per_chunk = ph.tiles_per_chunk;
  p.num_units = ph.num_chunks;
  p.cap = pl.cap; p.cap_soft = pl.cap_soft; p.Bpad = pl.Bpad;
  p.row_mult = ix->row_mult;
  p.qscale = reinterpret_cast<const float*>(reinterpret_cast<const char*>(qws) + rows_i8_limb_bytes(pl.rows_i8, NQ, ix->D));
  p.lists = reinterpret_cast<u64*>((char*)ws + pl.off_lists);
  p.counts = reinterpret_cast<int*>((char*)ws + pl.off_counts);
  p.gthr = reinterpret_cast<u32*>((char*)ws + pl.off_gthr);
  // instruction descriptor: D = s32 (2 << 4), A = B = signed 8 bit (1 << 7, 1 << 10), K-major both, N >> 3, M >> 4
  p.idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((u32)(RiGeo<NQ>::NCOLS >> 3) << 17) | ((u32)(RI_ROWS >> 4) << 24);
  const size_t smem = rows_i8_smem_bytes(NQ, p.num_kb, p.stages);
  B200R_CUDA_CHECK(cudaFuncSetAttribute(rows_i8_kernel<NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // a gang of `groups` CTAs per unit slot; the plan sized the chunks for num_sms / groups CTAs per group
  int per_group = rows_i8_ctas_per_group(pl.rows_i8);
  if (per_group > p.num_units) per_group = p.num_units;
  rows_i8_kernel<NQ><<<per_group * p.groups, RI_THREADS, smem, st>>>(*tmap_q, ix->tmap_x_i8, p);
  count_launch();
  B200R_CUDA_CHECK(cudaGetLastError());
  return B200R_OK;
}

// CTAs each query group gets when `groups` groups share one launch (one CTA per SM in total)
int rows_i8_ctas_per_group(int groups) {
  const int n = num_sms() / (groups > 0 ? groups : 1);
  return n > 0 ? n : 1;
}

// one launch = one phase of every group of <= pl.rows_i8_nq queries of the (<= 256-query) slice
int launch_rows_i8(const Index* ix, const CUtensorMap* tmap_q, const void* qws, const SearchPlan& pl, const GemmPhase& ph,
                   void* ws, cudaStream_t st) {
  ProfScope _ps(st, "rows_i8");
  switch (pl.rows_i8_nq) {
    case 16: return launch_rows_i8_nq<16>(ix, tmap_q, qws, pl, ph, ws, st);
    case 32: return launch_rows_i8_nq<32>(ix, tmap_q, qws, pl, ph, ws, st);
    case 64: return launch_rows_i8_nq<64>(ix, tmap_q, qws, pl, ph, ws, st);
  }
  set_error("int8 tensor path: bad group size %d", pl.rows_i8_nq);
  return B200R_EINVAL;
}

}  // namespace b200r
