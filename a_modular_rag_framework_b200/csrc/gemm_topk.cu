// gemm_topk.cu -- K1: Q x X^T on tcgen05/TMEM tiles fed by TMA, with the per-tile score epilogue
// fused into a running exact top-k, so the [B, N] score matrix never exists in HBM.
//
// Replaces DenseReranker._cosine over the candidate pool + sort/slice
// (reference app/modules/retrieval/retrieval_backend.py:192-197, 245, 371-372) for a whole batch
// against the whole corpus.
//
// Tile: 128 queries (TMEM lanes) x 256 corpus rows (TMEM columns), K streamed in 64-element
// (128-byte, SWIZZLE_128B) slices through a 4-stage TMA->smem ring.  Two 256-column accumulators
// alternate so that the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Warp roles (192 threads, 1 CTA/SM): warp 0 = TMA producer (1 thread), warp 1 = TMEM allocator and
// tcgen05.mma issuer (1 thread), warps 2..5 = epilogue; epilogue thread t owns TMEM lane
// (warp%4)*32+lane == one query of the tile and therefore sees that query's scores for consecutive
// corpus rows.  It keeps the query's current k-th best score in a register; a score passes only if it
// beats it (>=), which after warm-up almost never happens.  Survivors are appended to that query's
// list (global memory, L2 resident); when a list reaches cap_soft the warp cooperatively reduces it
// to its exact top-k (radix select, common.cuh) and tightens the threshold.  A global per-query
// threshold (gthr) lets later units start from the best threshold any earlier unit has published.
//
// Work unit = (query tile, chunk of consecutive row tiles); units are numbered chunk-major so the
// CTAs that share a chunk stream the same X tiles at the same time and hit them in L2.
#include "common.cuh"
#include "internal.h"
#include "tc_ptx.cuh"
#include <math_constants.h>
#include <cstring>
#include <type_traits>
#include <cstdlib>

namespace b200r {

constexpr int BM = 128;     // queries per CTA (TMEM lanes)
constexpr int BN = 256;     // corpus rows per tile (TMEM columns)
constexpr int BK = 64;
constexpr int A_BYTES = BM * BK * 2;   // 16 KB
constexpr int GEMM_THREADS = 192;
constexpr int TMEM_COLS = 512;

// CG = 1: one CTA per tile (M=128), it loads all 256 rows of B: 48 KB/stage, 4 stages.
// CG = 2: a CTA pair per tile (cta_group::2, M=256 = 2 x 128 queries), each CTA loads its own A and HALF of
//         B (128 rows): 32 KB/stage, 6 stages, half the shared-memory operand traffic per flop.
template <int CG> struct Geo {
  static constexpr int kBRows = BN / CG;
  static constexpr int kBBytes = kBRows * BK * 2;
  static constexpr int kStageBytes = A_BYTES + kBBytes;
  static constexpr int kStages = CG == 1 ? 4 : 6;
  static constexpr int kSmemBytes = 1024 /*align slack*/ + kStages * kStageBytes + 2 * BN * 4 /*row mult*/ + 2 * 8 * 4 /*chunk max mult*/ +
                                    4 * 256 * 4 /*hist*/ + 24 * 8 /*barriers*/ + 16;
};

struct GemmTopkParams {
  int B, k, D;
  long long N;
  int num_q_tiles, num_row_tiles;
  int tile_begin, tile_end, tiles_per_chunk, num_units;   // this phase: row tiles [tile_begin, tile_end) cut into chunks
  int cap, cap_soft, Bpad;
  const float* row_mult;
  u64* lists;
  int* counts;
  u32* gthr;
  u32 idesc;
  unsigned long long x_hint;   // L2 eviction policy of the corpus tiles
  int q_replicas;              // > 1: the (single) query tile exists in this many copies, CTA b reads copy b % q_replicas
  // range mode (b200r_range_count / b200r_range_fill): every row with FINAL score >= tau (and id > id_gt[q]) instead of a
  // top-k.  1 = count per (chunk, query) into `counts`; 2 = write the keys at range_off[slot] + i into range_out.
  int range_mode;
  float tau;
  const float* qinv;           // nullable (inner product)
  const long long* id_gt;      // nullable: only rows whose GLOBAL id exceeds id_gt[q]
  long long id_base;
  const long long* range_off;  // [num_chunks][Bpad]
  u64* range_out;
};

// s[j] for a runtime j without spilling s[] to local memory: a balanced tree of 31 selects
template <int LO, int N>
__device__ __forceinline__ float mux_tree(const float (&s)[32], int j) {
  if constexpr (N == 1) {
    return s[LO];
  } else {
    const float a = mux_tree<LO, N / 2>(s, j);
    const float b = mux_tree<LO + N / 2, N / 2>(s, j);
    return (j & (N / 2)) ? b : a;
  }
}
__device__ __forceinline__ float mux32(const float (&s)[32], int j) { return mux_tree<0, 32>(s, j); }

// ------------------------------------------------------------------------------------------ kernel
template <int CG>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_topk_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_x,
                 const GemmTopkParams p) {
  constexpr int STAGES = Geo<CG>::kStages;
  constexpr int B_BYTES = Geo<CG>::kBBytes;
  constexpr int STAGE_BYTES = Geo<CG>::kStageBytes;
  const u32 cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int unit0 = (CG == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;       // first unit of this CTA (pair)
  const int unit_stride = (CG == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  extern __shared__ uint8_t smem_raw[];
  const u32 raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  uint8_t* smem_a = smem;                               // STAGES x 16 KB
  uint8_t* smem_b = smem + STAGES * A_BYTES;            // STAGES x (32 | 16) KB
  float* smem_mult = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);  // [2][BN]
  float* smem_mmax = smem_mult + 2 * BN;                                     // [2][8] per-32-row-chunk max multiplier
  u32* smem_hist = reinterpret_cast<u32*>(smem_mmax + 2 * 8);                // [4][256]
  u64* bars = reinterpret_cast<u64*>(smem_hist + 4 * 256);
  u64* full_bar = bars;                  // [STAGES]
  u64* empty_bar = bars + STAGES;        // [STAGES]
  u64* tmem_full_bar = bars + 2 * STAGES;       // [2]
  u64* tmem_empty_bar = bars + 2 * STAGES + 2;  // [2]
  u32* tmem_ptr_smem = reinterpret_cast<u32*>(bars + 2 * STAGES + 4);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_q);
    prefetch_tmap(&tmap_x);
    // full: the leader's own arrive(+expect_tx) and, in a pair, the peer producer's remote arrive
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], CG); mbar_init(&empty_bar[i], 1); }
    // tmem_empty: one arrive per epilogue warp of every CTA of the group (the peer's arrive remotely)
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], 4 * CG); }
    fence_barrier_init();
  }
  if constexpr (CG == 2) cluster_sync_all();      // both CTAs' barriers exist before anything can reach them
  if (warp_idx == 1) {
    if constexpr (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const u32 tmem_base = *tmem_ptr_smem;

  const int num_kb = (p.D + BK - 1) / BK;

  if (warp_idx == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int stage = 0; u32 phase = 0;
      for (int u = unit0; u < p.num_units; u += unit_stride) {
        const int q_tile = (u % p.num_q_tiles) * CG + (int)cta_rank;    // this CTA's 128 queries
        const int chunk = u / p.num_q_tiles;
        const int t0 = p.tile_begin + chunk * p.tiles_per_chunk;
        const int t1 = min(p.tile_end, t0 + p.tiles_per_chunk);
        for (int t = t0; t < t1; ++t) {
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if constexpr (CG == 1) {
              mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
              tma_load_2d(smem_a + stage * A_BYTES, &tmap_q, &full_bar[stage], kb * BK,
                          (q_tile + (p.q_replicas > 1 ? (int)(blockIdx.x % p.q_replicas) : 0)) * BM, kEvictLast);
              tma_load_2d(smem_b + stage * B_BYTES, &tmap_x, &full_bar[stage], kb * BK, t * BN, p.x_hint);
            } else {
              // both CTAs' bytes are credited to the leader's barrier; this CTA loads its A and its half of B
              if (leader) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
              tma_load_2d_cg2(smem_a + stage * A_BYTES, &tmap_q, &full_bar[stage], kb * BK, q_tile * BM, kEvictLast);
              tma_load_2d_cg2(smem_b + stage * B_BYTES, &tmap_x, &full_bar[stage], kb * BK,
                              t * BN + (int)cta_rank * (BN / 2), p.x_hint);
              if (!leader) mbar_arrive_remote(&full_bar[stage], 0);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================================== MMA issuer =======================================
    if (lane == 0 && leader) {
      int stage = 0; u32 phase = 0;
      u32 tile_iter = 0;
      for (int u = unit0; u < p.num_units; u += unit_stride) {
        const int chunk = u / p.num_q_tiles;
        const int t0 = p.tile_begin + chunk * p.tiles_per_chunk;
        const int t1 = min(p.tile_end, t0 + p.tiles_per_chunk);
        for (int t = t0; t < t1; ++t, ++tile_iter) {
          const u32 acc = tile_iter & 1;
          mbar_wait(&tmem_empty_bar[acc], ((tile_iter >> 1) & 1) ^ 1);
          tc_fence_after();
          const u32 tmem_d = tmem_base + acc * BN;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const u64 desc_a = make_smem_desc(smem_u32(smem_a + stage * A_BYTES));
            const u64 desc_b = make_smem_desc(smem_u32(smem_b + stage * B_BYTES));
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk) {
              // advance 16 elements = 32 bytes inside the 128-byte swizzle span: +2 in the >>4 address field
              if constexpr (CG == 1) umma_f16(tmem_d, desc_a + (u64)(kk * 2), desc_b + (u64)(kk * 2), p.idesc, (kb | kk) != 0 ? 1u : 0u);
              else umma_f16_cg2(tmem_d, desc_a + (u64)(kk * 2), desc_b + (u64)(kk * 2), p.idesc, (kb | kk) != 0 ? 1u : 0u);
            }
            if constexpr (CG == 1) {
              umma_commit(&empty_bar[stage]);                 // frees the smem slot when these MMAs retire
              if (kb == num_kb - 1) umma_commit(&tmem_full_bar[acc]);  // accumulator complete
            } else {
              umma_commit_cg2(&empty_bar[stage]);             // ... in both CTAs of the pair
              if (kb == num_kb - 1) umma_commit_cg2(&tmem_full_bar[acc]);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else {
    // ===================================== epilogue =========================================
    const int quarter = warp_idx & 3;             // TMEM lane quarter this warp may access
    const int q_lane = quarter * 32 + lane;       // query within the tile
    const int epi_tid = (warp_idx - 2) * 32 + lane;  // 0..127
    u32* hist = smem_hist + (warp_idx - 2) * 256;
    u32 tile_iter = 0;
    for (int u = unit0; u < p.num_units; u += unit_stride) {
      const int q_tile = (u % p.num_q_tiles) * CG + (int)cta_rank;
      const int chunk = u / p.num_q_tiles;
      const int t0 = p.tile_begin + chunk * p.tiles_per_chunk;
      const int t1 = min(p.tile_end, t0 + p.tiles_per_chunk);
      if (t0 >= t1) continue;                 // empty chunk in a short phase (uniform across roles)
      const int q = q_tile * BM + q_lane;
      const bool valid = q < p.B;
      const size_t slot = (size_t)chunk * p.Bpad + (size_t)q_tile * BM + q_lane;
      u64* my_list = p.lists + slot * (size_t)p.cap;
      int cnt = valid ? __ldcg(p.counts + slot) : 0;   // lists persist across phases (ids keep increasing)
      u64 kth_key = 0;
      float thr = valid ? ordered_to_f32(__ldcg(p.gthr + q)) : __int_as_float(0x7f800000);
      // range mode: the test is on the final score acc*mult*qi >= tau, identical in the count and the fill pass
      const bool range = p.range_mode != 0;
      float qi = 1.0f;
      long long row_gt = -1;                    // local rows <= row_gt are excluded
      u64* range_dst = nullptr;
      if (range) {
        cnt = 0;
        thr = valid ? p.tau : __int_as_float(0x7f800000);
        if (valid) {
          qi = p.qinv ? p.qinv[q] : 1.0f;
          if (p.id_gt) row_gt = p.id_gt[q] - p.id_base;
          if (p.range_mode == 2) range_dst = p.range_out + p.range_off[slot];
        }
      }

      for (int t = t0; t < t1; ++t, ++tile_iter) {
        const u32 acc = tile_iter & 1;
        float* mult = smem_mult + (tile_iter & 1) * BN;
        float* mmax = smem_mmax + (tile_iter & 1) * 8;
        // per-row multipliers of this tile (1/|x| for cosine); NaN marks rows beyond N so they never pass.
        // Also the max multiplier of each 32-row chunk (slightly inflated): acc*mult <= max(0, max_acc * mmax), so a
        // chunk whose raw accumulator max cannot reach the threshold is rejected without touching the multipliers.
        {
          const long long r = (long long)t * BN + epi_tid * 2;
          float m0 = __int_as_float(0x7fc00000), m1 = __int_as_float(0x7fc00000);
          if (r < p.N) m0 = p.row_mult ? __ldg(p.row_mult + r) : 1.0f;
          if (r + 1 < p.N) m1 = p.row_mult ? __ldg(p.row_mult + r + 1) : 1.0f;
          mult[epi_tid * 2] = m0;
          mult[epi_tid * 2 + 1] = m1;
          float mm = fmaxf(fmaxf(m0, m1), 0.0f);      // fmaxf drops the NaNs
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) mm = fmaxf(mm, __shfl_xor_sync(kFullMask, mm, o));   // 16-lane halves = one chunk
          if ((lane & 15) == 0) mmax[epi_tid >> 4] = mm * 1.000001f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        mbar_wait(&tmem_full_bar[acc], (tile_iter >> 1) & 1);
        tc_fence_after();
        const u32 row0 = (u32)t * BN;
        const u32 taddr = tmem_base + ((u32)(quarter * 32) << 16) + acc * BN;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          u32 v[32];
          tmem_ld32(taddr + c * 32, v);
          tmem_ld_wait();
          if (c == BN / 32 - 1) {
            // the whole accumulator is in registers: hand the TMEM stage back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (CG == 1 || leader) mbar_arrive(&tmem_empty_bar[acc]);
              else mbar_arrive_remote(&tmem_empty_bar[acc], 0);
            }
          }
          float rmax = fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1]));
#pragma unroll
          for (int j = 2; j < 32; j += 2) rmax = fmaxf(rmax, fmaxf(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
          // conservative test on the raw accumulators (see mmax above); a non-positive threshold always enters
          const bool maybe = !((rmax * mmax[c]) * qi < thr) || !(thr > 0.0f);     // qi == 1 outside range mode
          if (range) {
            if (__any_sync(kFullMask, maybe)) {
              float s[32];
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 m4 = *reinterpret_cast<const float4*>(mult + c * 32 + j);
                s[j] = (__uint_as_float(v[j]) * m4.x) * qi;
                s[j + 1] = (__uint_as_float(v[j + 1]) * m4.y) * qi;
                s[j + 2] = (__uint_as_float(v[j + 2]) * m4.z) * qi;
                s[j + 3] = (__uint_as_float(v[j + 3]) * m4.w) * qi;
              }
              u32 mask = 0;
#pragma unroll
              for (int j = 0; j < 32; ++j) mask |= (s[j] >= thr) ? (1u << j) : 0u;     // NaN multipliers (rows beyond N) never pass
              // rows at or below the query's own id bound
              const long long first = (long long)row0 + c * 32;
              if (row_gt >= first) mask &= (row_gt - first >= 31) ? 0u : (0xffffffffu << (int)(row_gt - first + 1));
              if (p.range_mode == 1) {
                cnt += __popc(mask);
              } else {
                while (mask) {
                  const int j = __ffs(mask) - 1;
                  mask &= mask - 1;
                  range_dst[cnt++] = make_key(mux32(s, j), row0 + c * 32 + j);
                }
              }
            }
          } else if (__any_sync(kFullMask, maybe)) {
            // rare path: exact scores, survivor bitmask per lane (straight-line), then each lane walks its own bits
            // and picks the value with a 5-level select tree -- no dependent vote/branch chain
            float s[32];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 m4 = *reinterpret_cast<const float4*>(mult + c * 32 + j);
              s[j] = __uint_as_float(v[j]) * m4.x;
              s[j + 1] = __uint_as_float(v[j + 1]) * m4.y;
              s[j + 2] = __uint_as_float(v[j + 2]) * m4.z;
              s[j + 3] = __uint_as_float(v[j + 3]) * m4.w;
            }
            u32 mask = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) mask |= (s[j] >= thr) ? (1u << j) : 0u;
            while (mask) {
              const int j = __ffs(mask) - 1;
              mask &= mask - 1;
              const u64 key = make_key(mux32(s, j), row0 + c * 32 + j);
              if (key > kth_key) { __stcg(my_list + cnt, key); ++cnt; }
            }
            u32 full = __ballot_sync(kFullMask, cnt >= p.cap_soft);
            while (full) {
              const int l = __ffs(full) - 1;
              full &= full - 1;
              u64* lp = reinterpret_cast<u64*>(__shfl_sync(kFullMask, (u64)my_list, l));
              const int n = __shfl_sync(kFullMask, cnt, l);
              const u64 kth = warp_compact(lp, n, p.k, hist);
              if (lane == l) {
                kth_key = kth;
                thr = fmaxf(thr, key_score(kth));
                cnt = p.k;
                atomicMax(p.gthr + q, (u32)(kth >> 32));
              }
            }
          }
        }
        // pick up thresholds published by other CTAs working on the same queries
        if (valid && !range) thr = fmaxf(thr, ordered_to_f32(__ldcg(p.gthr + q)));
      }
      if (p.range_mode != 2) p.counts[slot] = valid ? cnt : 0;
    }
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();   // nobody leaves while the peer can still reach its smem
  if (warp_idx == 1) {
    __syncwarp();
    tc_fence_after();
    if constexpr (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ host
// One query tile (17 <= B <= 128): every CTA re-reads the same few query lines from L2 with every stage, which hot-spots
// a handful of L2 slices (measured: 79% of the copy bandwidth at B=16-class tiles vs 106% at B=128).  Spreading the
// tile over kQReplicas copies at different addresses removes the hot spot.
constexpr int kQReplicas = 16;

template <typename T>
__global__ void replicate_queries_kernel(const T* __restrict__ Q, T* __restrict__ rep, int B, int D) {
  const size_t per = (size_t)BM * D;
  const size_t total = per * kQReplicas;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t j = i % per;
    const int row = (int)(j / D);
    rep[i] = row < B ? Q[j] : T(0);
  }
}

size_t query_replica_bytes(int D) { return (size_t)kQReplicas * BM * D * 2; }

// Everything a K1 search needs before its first tile kernel, in ONE launch (three launches cost ~20 us of a 0.4 ms batch-1
// search on a 1.25 M-row shard): thresholds to -inf and list counts to zero, 1/|q| per query (cosine), and -- one query
// tile only -- the 16 replicas of the zero-padded query tile, written as 16-byte chunks.
template <typename QT>
__global__ void __launch_bounds__(256) search_prologue_kernel(u32* __restrict__ gthr, int n_gthr, int* __restrict__ counts, long long n_counts,
                                                              const QT* __restrict__ Q, int B, int D, float* __restrict__ qinv,
                                                              uint4* __restrict__ rep) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  for (long long j = tid; j < n_gthr; j += nthreads) gthr[j] = kOrderedNegInf;
  for (long long j = tid; j < n_counts; j += nthreads) counts[j] = 0;
  if (qinv) {                                      // one warp per query, same arithmetic as query_inv_norm_kernel
    const int lane = threadIdx.x & 31;
    for (long long b = tid >> 5; b < B; b += nthreads >> 5) {
      double acc = 0.0;
      for (int d = lane; d < D; d += 32) {
        const double v = (double)q_to_float<QT>(Q[(size_t)b * D + d]);
        acc += v * v;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFullMask, acc, o);
      if (lane == 0) qinv[b] = acc > 0.0 ? (float)(1.0 / sqrt(acc)) : 0.0f;
    }
  }
  if (rep) {                                       // D % 8 == 0 and Q 16-byte aligned on this path
    const int cpr = D / 8;                         // 16-byte chunks per row
    const long long per = (long long)BM * cpr;
    const uint4* Qv = reinterpret_cast<const uint4*>(Q);
    for (long long j = tid; j < per; j += nthreads) {
      const int row = (int)(j / cpr);
      const uint4 v = row < B ? Qv[j] : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int r = 0; r < kQReplicas; ++r) rep[(long long)r * per + j] = v;
    }
  }
}

int launch_search_prologue(const SearchPlan& pl, void* ws, const void* Q, int D, int qdtype, bool cosine, cudaStream_t st) {
  ProfScope _ps(st, "search_prologue");
  u32* gthr = reinterpret_cast<u32*>((char*)ws + pl.off_gthr);
  int* counts = reinterpret_cast<int*>((char*)ws + pl.off_counts);
  const long long n_counts = (long long)(pl.num_segments + 1) * pl.Bpad;   // + the carry lists' counts
  float* qinv = cosine ? reinterpret_cast<float*>((char*)ws + pl.off_qinv) : nullptr;
  uint4* rep = pl.q_replicas ? reinterpret_cast<uint4*>((char*)ws + pl.off_qrep) : nullptr;
  const int grid = num_sms() * 2;
  if (qdtype == B200R_BF16)
    search_prologue_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(gthr, pl.Bpad, counts, n_counts, reinterpret_cast<const __nv_bfloat16*>(Q), pl.B, D, qinv, rep);
  else if (qdtype == B200R_F16)
    search_prologue_kernel<__half><<<grid, 256, 0, st>>>(gthr, pl.Bpad, counts, n_counts, reinterpret_cast<const __half*>(Q), pl.B, D, qinv, rep);
  else { set_error("search_prologue: query dtype %d", qdtype); return B200R_EINVAL; }
  count_launch();
  B200R_CUDA_CHECK(cudaGetLastError());
  return B200R_OK;
}

// f32 host queries -> the index's 2-byte query dtype, round to nearest even (b200r_search_host_f32)
template <typename T>
__global__ void convert_queries_kernel(const float* __restrict__ src, T* __restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if constexpr (sizeof(T) == sizeof(__nv_bfloat16) && std::is_same<T, __nv_bfloat16>::value) dst[i] = __float2bfloat16_rn(src[i]);
    else dst[i] = __float2half_rn(src[i]);
  }
}

int launch_convert_queries(const float* src, void* dst, size_t n, int dtype, cudaStream_t st) {
  ProfScope _ps(st, "convert_queries");
  if (n == 0) return B200R_OK;
  const int threads = 256;
  size_t blocks = (n + threads - 1) / threads;
  if (blocks > (size_t)num_sms() * 8) blocks = (size_t)num_sms() * 8;
  if (dtype == B200R_BF16) convert_queries_kernel<__nv_bfloat16><<<(unsigned)blocks, threads, 0, st>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), n);
  else if (dtype == B200R_F16) convert_queries_kernel<__half><<<(unsigned)blocks, threads, 0, st>>>(src, reinterpret_cast<__half*>(dst), n);
  else { set_error("convert_queries: dtype %d is not a 2-byte float type", dtype); return B200R_EINVAL; }
  count_launch();
  B200R_CUDA_CHECK(cudaGetLastError());
  return B200R_OK;
}

int launch_replicate_queries(const void* Q, void* rep, int B, int D, cudaStream_t st) {
  ProfScope _ps(st, "replicate_queries");
  replicate_queries_kernel<unsigned short><<<num_sms(), 256, 0, st>>>(reinterpret_cast<const unsigned short*>(Q),
                                                                    reinterpret_cast<unsigned short*>(rep), B, D);
  count_launch();
  B200R_CUDA_CHECK(cudaGetLastError());
  return B200R_OK;
}

int encode_query_tmap_replicated(const Index* ix, const void* rep, CUtensorMap* tmap_q) {
  return encode_tmap_2d(tmap_q, ix->dtype, rep, (uint64_t)ix->D, (uint64_t)kQReplicas * BM, BK, BM);
}

int encode_query_tmap(const Index* ix, const void* Q, int B, CUtensorMap* tmap_q) {
  if (!ix->has_tmap) { set_error("gemm path needs a bf16/f16 corpus with D %% 8 == 0"); return B200R_EINVAL; }
  return encode_tmap_2d(tmap_q, ix->dtype, Q, (uint64_t)ix->D, (uint64_t)B, BK, BM);
}

template <int CG>
static int launch_gemm_cg(const Index* ix, const CUtensorMap* tmap_q, const SearchPlan& pl, const GemmPhase& ph, void* ws, cudaStream_t st,
                          const RangeArgs* rg) {
  GemmTopkParams p;
  p.B = pl.B; p.k = pl.k; p.D = ix->D; p.N = ix->N;
  p.num_q_tiles = pl.num_q_tiles; p.num_row_tiles = pl.num_row_tiles;     // num_q_tiles counts tiles of CG*128 queries
  p.tile_begin = ph.tile_begin; p.tile_end = ph.tile_end; p.tiles_per_chunk = ph.tiles_per_chunk;
  p.num_units = ph.num_chunks * pl.num_q_tiles;
  p.cap = pl.cap; p.cap_soft = pl.cap_soft; p.Bpad = pl.Bpad;
  p.row_mult = ix->row_mult;
  p.lists = reinterpret_cast<u64*>((char*)ws + pl.off_lists);
  p.counts = reinterpret_cast<int*>((char*)ws + pl.off_counts);
  p.gthr = reinterpret_cast<u32*>((char*)ws + pl.off_gthr);
  // instruction descriptor: D=f32 (1<<4), A/B format (bf16=1, f16=0) at [7,10)/[10,13), K-major A and B,
  // N>>3 at [17,23), M>>4 at [24,29) with M = 128 per CTA of the group
  const u32 fmt = (ix->dtype == B200R_BF16) ? 1u : 0u;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((u32)(BN >> 3) << 17) | ((u32)((BM * CG) >> 4) << 24);
  // one query tile: every corpus tile is read exactly once -> do not let it displace Q / the lists in L2
  p.x_hint = (pl.num_q_tiles == 1) ? kEvictFirst : kEvictNormal;
  p.q_replicas = pl.q_replicas ? kQReplicas : 1;
  p.range_mode = 0; p.tau = 0.0f; p.qinv = nullptr; p.id_gt = nullptr; p.id_base = ix->id_base; p.range_off = nullptr; p.range_out = nullptr;
  if (rg) {
    p.range_mode = rg->mode; p.tau = rg->tau; p.qinv = rg->qinv; p.id_gt = reinterpret_cast<const long long*>(rg->id_gt);
    p.range_off = reinterpret_cast<const long long*>(rg->offsets); p.range_out = reinterpret_cast<u64*>(rg->out);
  }
  B200R_CUDA_CHECK(cudaFuncSetAttribute(gemm_topk_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<CG>::kSmemBytes));
  const int groups = p.num_units < num_sms() / CG ? p.num_units : num_sms() / CG;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(groups * CG);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Geo<CG>::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  B200R_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_topk_kernel<CG>, *tmap_q, CG == 2 ? ix->tmap_x_half : ix->tmap_x, p));
  count_launch();
  B200R_CUDA_CHECK(cudaGetLastError());
  return B200R_OK;
}

int launch_gemm_topk(const Index* ix, const CUtensorMap* tmap_q, const SearchPlan& pl, const GemmPhase& ph, void* ws, cudaStream_t st,
                     const RangeArgs* rg) {
  ProfScope _ps(st, rg ? "gemm_range" : "gemm_topk");
  return pl.cta_group == 2 ? launch_gemm_cg<2>(ix, tmap_q, pl, ph, ws, st, rg) : launch_gemm_cg<1>(ix, tmap_q, pl, ph, ws, st, rg);
}

int gemm_tile_rows() { return BN; }
int gemm_tile_queries() { return BM; }

}  // namespace b200r
