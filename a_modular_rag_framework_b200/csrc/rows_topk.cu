// rows_topk.cu -- K1b: the tcgen05 path for SMALL query batches (B <= 16), laid out for HBM streaming.
//
// Same contract as K1 (exact cosine / IP top-k; reference retrieval_backend.py:192-197, 371-372) with the
// operands swapped: the 128 corpus rows of a tile are the MMA's M (TMEM lanes), the (padded) queries its N
// (TMEM columns).  The query block stays RESIDENT in shared memory (NQ x D, loaded once), so the whole ring
// -- up to 12 stages of 16 KB -- carries corpus data only and consecutive stages are consecutive 64-column
// slices of the SAME 128 rows: at any moment the TMA has a contiguous 128 x D block of the corpus in flight,
// which is what lets the stream approach the DRAM page-hit rate of a plain copy (K1 re-fetches Q with every
// stage and interleaves 128-byte pieces of 256 rows).
//
// Epilogue: thread = TMEM lane = one corpus row; it holds that row's NQ scores, scales them by the row's
// multiplier and tests them against the per-query thresholds (warp-uniform, in smem).  Survivors are
// appended per (warp, query) with a ballot (rows stay in id order inside a list), lists are reduced by the
// same warp radix select, thresholds shared through gthr, phases and pruning exactly as in K1.
#include "common.cuh"
#include "internal.h"
#include "tc_ptx.cuh"
#include <math_constants.h>
#include <cstring>

namespace b200r {

constexpr int RT_ROWS = 128;          // corpus rows per tile
constexpr int RT_BK = 64;
constexpr int RT_STAGE_BYTES = RT_ROWS * RT_BK * 2;   // 16 KB
constexpr int RT_THREADS = 192;
constexpr int RT_ACC = 4;             // accumulator stages in TMEM

struct RowsTopkParams {
  int B, k, D, num_kb, stages;
  long long N;
  int tile_begin, tile_end, tiles_per_chunk, num_units;
  int cap, cap_soft, Bpad;
  const float* row_mult;
  u64* lists;
  int* counts;
  u32* gthr;
  u32 idesc;
};

__device__ __forceinline__ void tmem_ld16(u32 taddr, u32* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

template <int NQ>
__global__ void __launch_bounds__(RT_THREADS, 1)
rows_topk_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_x,
                 const RowsTopkParams p) {
  extern __shared__ uint8_t smem_raw[];
  const u32 raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  constexpr int Q_KB_BYTES = NQ * RT_BK * 2;                 // one 64-column slice of the query block
  uint8_t* smem_q = smem;                                    // [num_kb][NQ x 128 B]
  uint8_t* smem_x = smem + (size_t)p.num_kb * Q_KB_BYTES;    // [stages][128 x 128 B]   (Q_KB_BYTES is a multiple of 1 KB)
  uint8_t* tail = smem_x + (size_t)p.stages * RT_STAGE_BYTES;
  u32* smem_hist = reinterpret_cast<u32*>(tail);             // [4][256]
  float* thr_s = reinterpret_cast<float*>(smem_hist + 4 * 256);   // [4][NQ]
  u64* kth_s = reinterpret_cast<u64*>(thr_s + 4 * NQ);       // [4][NQ]
  int* cnt_s = reinterpret_cast<int*>(kth_s + 4 * NQ);       // [4][NQ]
  u64* bars = reinterpret_cast<u64*>(cnt_s + 4 * NQ);
  u64* full_bar = bars;                      // [stages] (<= 12)
  u64* empty_bar = bars + 12;                // [stages]
  u64* tmem_full_bar = bars + 24;            // [RT_ACC]
  u64* tmem_empty_bar = bars + 24 + RT_ACC;  // [RT_ACC]
  u64* q_bar = bars + 24 + 2 * RT_ACC;
  u32* tmem_ptr_smem = reinterpret_cast<u32*>(bars + 25 + 2 * RT_ACC);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stages = p.stages;
  constexpr u32 TMEM_COLS = (RT_ACC * NQ) < 32 ? 32 : (RT_ACC * NQ);   // 64 / 128 / 256: powers of two

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_q);
    prefetch_tmap(&tmap_x);
    for (int i = 0; i < stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < RT_ACC; ++i) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], 4); }
    mbar_init(q_bar, 1);
    fence_barrier_init();
  }
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
      // the query block, once: num_kb slices of [NQ rows x 64 columns] (rows >= B are zero-filled by TMA)
      mbar_expect_tx(q_bar, (u32)(p.num_kb * Q_KB_BYTES));
      for (int kb = 0; kb < p.num_kb; ++kb) tma_load_2d(smem_q + (size_t)kb * Q_KB_BYTES, &tmap_q, q_bar, kb * RT_BK, 0, kEvictLast);
      int stage = 0; u32 phase = 0;
      for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
        const int t0 = p.tile_begin + u * p.tiles_per_chunk;
        const int t1 = min(p.tile_end, t0 + p.tiles_per_chunk);
        for (int t = t0; t < t1; ++t) {
          for (int kb = 0; kb < p.num_kb; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_expect_tx(&full_bar[stage], RT_STAGE_BYTES);
            tma_load_2d(smem_x + (size_t)stage * RT_STAGE_BYTES, &tmap_x, &full_bar[stage], kb * RT_BK, t * RT_ROWS, kEvictFirst);
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
      for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
        const int t0 = p.tile_begin + u * p.tiles_per_chunk;
        const int t1 = min(p.tile_end, t0 + p.tiles_per_chunk);
        for (int t = t0; t < t1; ++t, ++tile_iter) {
          const u32 acc = tile_iter % RT_ACC;
          mbar_wait(&tmem_empty_bar[acc], ((tile_iter / RT_ACC) & 1) ^ 1);
          tc_fence_after();
          const u32 tmem_d = tmem_base + acc * NQ;
          for (int kb = 0; kb < p.num_kb; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const u64 desc_a = make_smem_desc(smem_u32(smem_x + (size_t)stage * RT_STAGE_BYTES));   // corpus rows = M
            const u64 desc_b = make_smem_desc(smem_u32(smem_q + (size_t)kb * Q_KB_BYTES));          // queries = N
#pragma unroll
            for (int kk = 0; kk < RT_BK / 16; ++kk)
              umma_f16(tmem_d, desc_a + (u64)(kk * 2), desc_b + (u64)(kk * 2), p.idesc, (kb | kk) != 0 ? 1u : 0u);
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
    float* thr = thr_s + ew * NQ;
    u64* kth = kth_s + ew * NQ;
    int* cnt = cnt_s + ew * NQ;
    u32 tile_iter = 0;
    for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
      const int t0 = p.tile_begin + u * p.tiles_per_chunk;
      const int t1 = min(p.tile_end, t0 + p.tiles_per_chunk);
      if (t0 >= t1) continue;
      const size_t seg = (size_t)u * 4 + ew;               // this warp's list segment for the unit's row chunk
      for (int j = lane; j < NQ; j += 32) {
        const bool valid = j < p.B;
        thr[j] = valid ? ordered_to_f32(__ldcg(p.gthr + j)) : __int_as_float(0x7f800000);
        kth[j] = 0;
        cnt[j] = valid ? __ldcg(p.counts + seg * p.Bpad + j) : 0;
      }
      __syncwarp();
      for (int t = t0; t < t1; ++t, ++tile_iter) {
        const u32 acc = tile_iter % RT_ACC;
        const long long row = (long long)t * RT_ROWS + quarter * 32 + lane;
        const float m = (row < p.N) ? (p.row_mult ? __ldg(p.row_mult + row) : 1.0f) : __int_as_float(0x7fc00000);
        mbar_wait(&tmem_full_bar[acc], (tile_iter / RT_ACC) & 1);
        tc_fence_after();
        const u32 taddr = tmem_base + ((u32)(quarter * 32) << 16) + acc * NQ;
        float s[NQ];
#pragma unroll
        for (int c = 0; c < NQ / 16; ++c) {
          u32 v[16];
          tmem_ld16(taddr + c * 16, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) s[c * 16 + j] = __uint_as_float(v[j]) * m;   // NaN for rows beyond N: never passes
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);

        bool any = false;
#pragma unroll
        for (int j = 0; j < NQ; ++j) any |= (s[j] >= thr[j]);
        if (__any_sync(kFullMask, any)) {
#pragma unroll
          for (int j = 0; j < NQ; ++j) {
            const float tj = thr[j];
            const u32 b = __ballot_sync(kFullMask, s[j] >= tj);
            if (b) {                                          // warp-uniform
              const u64 key = make_key(s[j], (u32)row);
              const u32 b2 = __ballot_sync(kFullMask, (s[j] >= tj) && key > kth[j]);
              if (b2) {
                const int base = cnt[j];
                u64* lp = p.lists + (seg * p.Bpad + j) * (size_t)p.cap;
                if ((b2 >> lane) & 1u) __stcg(lp + base + __popc(b2 & ((1u << lane) - 1u)), key);
                const int n = base + __popc(b2);
                __syncwarp();
                if (n >= p.cap_soft) {
                  const u64 kk = warp_compact(lp, n, p.k, hist);
                  if (lane == 0) {
                    kth[j] = kk; thr[j] = fmaxf(tj, key_score(kk)); cnt[j] = p.k;
                    atomicMax(p.gthr + j, (u32)(kk >> 32));
                  }
                } else if (lane == 0) {
                  cnt[j] = n;
                }
                __syncwarp();
              }
            }
          }
        }
        // thresholds published by other CTAs / warps
        for (int j = lane; j < NQ; j += 32)
          if (j < p.B) thr[j] = fmaxf(thr[j], ordered_to_f32(__ldcg(p.gthr + j)));
        __syncwarp();
      }
      for (int j = lane; j < NQ; j += 32)
        if (j < p.B) p.counts[seg * p.Bpad + j] = cnt[j];
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

// shared-memory budget: resident queries + ring + per-warp state
static size_t rows_smem_bytes(int nq, int num_kb, int stages) {
  return 1024 + (size_t)num_kb * nq * RT_BK * 2 + (size_t)stages * RT_STAGE_BYTES + 4 * 256 * 4 + (size_t)4 * nq * (4 + 8 + 4) +
         (26 + 2 * RT_ACC) * 8 + 16;
}

// Measured (10M x 768 bf16): NQ=16 streams at 94% of the copy bandwidth (K1 at B=16: 79%, limited by 148 SMs
// re-reading the same few query lines from L2 every stage), but at NQ=64 the per-query ballot epilogue makes it
// slower than K1 (61% vs 89%), so only batches of up to 16 queries take this kernel.
int rows_mode_nq(int B) { return B <= 16 ? 16 : 0; }

// number of ring stages that fit next to the resident query block, or 0 if the rows-mode kernel cannot be used
int rows_mode_stages(int B, int D) {
  const int nq = rows_mode_nq(B);
  if (!nq) return 0;
  const int num_kb = (D + RT_BK - 1) / RT_BK;
  int stages = 12;
  while (stages >= 4 && rows_smem_bytes(nq, num_kb, stages) > 227 * 1024) --stages;
  return stages >= 4 ? stages : 0;
}

int rows_tile_rows() { return RT_ROWS; }

template <int NQ>
static int launch_rows_nq(const Index* ix, const CUtensorMap* tmap_q, const SearchPlan& pl, const GemmPhase& ph, void* ws, cudaStream_t st) {
  RowsTopkParams p;
  p.B = pl.B; p.k = pl.k; p.D = ix->D; p.N = ix->N;
  p.num_kb = (ix->D + RT_BK - 1) / RT_BK;
  p.stages = rows_mode_stages(pl.B, ix->D);
  p.tile_begin = ph.tile_begin; p.tile_end = ph.tile_end; p.tiles_per_chunk = ph.tiles_per_chunk;
  p.num_units = ph.num_chunks;
  p.cap = pl.cap; p.cap_soft = pl.cap_soft; p.Bpad = pl.Bpad;
  p.row_mult = ix->row_mult;
  p.lists = reinterpret_cast<u64*>((char*)ws + pl.off_lists);
  p.counts = reinterpret_cast<int*>((char*)ws + pl.off_counts);
  p.gthr = reinterpret_cast<u32*>((char*)ws + pl.off_gthr);
  const u32 fmt = (ix->dtype == B200R_BF16) ? 1u : 0u;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((u32)(NQ >> 3) << 17) | ((u32)(RT_ROWS >> 4) << 24);
  const size_t smem = rows_smem_bytes(NQ, p.num_kb, p.stages);
  B200R_CUDA_CHECK(cudaFuncSetAttribute(rows_topk_kernel<NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = p.num_units < kNumSMs ? p.num_units : kNumSMs;
  rows_topk_kernel<NQ><<<grid, RT_THREADS, smem, st>>>(*tmap_q, ix->tmap_x_half, p);
  count_launch();
  B200R_CUDA_CHECK(cudaGetLastError());
  return B200R_OK;
}

int encode_query_tmap_rows(const Index* ix, const void* Q, int B, CUtensorMap* tmap_q) {
  return encode_tmap_2d(tmap_q, ix->dtype, Q, (uint64_t)ix->D, (uint64_t)B, RT_BK, (uint32_t)rows_mode_nq(B));
}

int launch_rows_topk(const Index* ix, const CUtensorMap* tmap_q, const SearchPlan& pl, const GemmPhase& ph, void* ws, cudaStream_t st) {
  switch (rows_mode_nq(pl.B)) {
    case 16: return launch_rows_nq<16>(ix, tmap_q, pl, ph, ws, st);
  }
  set_error("rows-mode kernel supports B <= 16");
  return B200R_EINVAL;
}

}  // namespace b200r
