// rows_f32.cu -- K1d: fp32 corpora on the tensor cores (tcgen05 kind::tf32, split precision), 16 queries per corpus pass.
//
// Same contract as K1/K2 (exact cosine / IP top-k; reference retrieval_backend.py:192-197, 371-372) at the fp32 bar
// (1e-5 relative).  The CUDA-core scans are FFMA-issue-bound beyond one fp32 query per pass (scan_bulk.cu: 0.43 of the
// HBM roofline at B = 8), so the dot products go to the tensor pipe, which only takes 11 significant bits per operand:
//   * every fp32 value v is split EXACTLY into hi = v with the low 13 mantissa bits cleared and mid = (v - hi) with ITS
//     low 13 bits cleared; both are exactly representable in tf32 (so the result does not depend on how the MMA rounds its
//     inputs) and v - hi - mid is below 2^-22 |v|;
//   * q . x = q_hi.x_hi + q_mid.x_hi + q_hi.x_mid  (+ terms below 2^-22 |q||x|): three kind::tf32 MMAs per K step; every
//     product of two 11-bit significands is exact; the fp32 accumulation is spread over 8 partial accumulators per tile
//     (see RF_PARTS) that the epilogue adds with round-to-nearest;
//   * operand layout is the streaming one of rows_i8.cu: the 128 corpus rows of a tile are the MMA's M (TMEM lanes), tiles of
//     [128 rows x 128 bytes = 32 floats] arrive by TMA (SWIZZLE_128B) straight from the fp32 corpus; a ring STAGE carries up
//     to three such K blocks (48 KB of corpus) plus the matching query slices under ONE barrier, because what bounds this
//     kernel is the number of barrier hand-offs per byte (see rows_f32_kps); four CONVERTER warps split each arrived slice in shared memory (hi in place, mid into a
//     double-buffered side tile; 15 instructions per 4 elements, a sixth of the SM's issue slots at HBM speed) and hand it to
//     the MMA warp; the hi and mid planes of the pass's 16 queries (split once per search, 96 KB at D = 768, L2-resident) are
//     streamed slice by slice with the corpus, which leaves shared memory for an 8-deep ring at any D;
//   * epilogue thread = TMEM lane = one corpus row: 16 fp32 scores, row multiplier, per-query thresholds, ballot appends;
//     lists, thresholds, phases and pruning exactly as in K1.
// Roofline: HBM, N*D*4 bytes per group of 16 queries.
#include "common.cuh"
#include "internal.h"
#include "tc_ptx.cuh"
#include <math_constants.h>
#include <cstring>
#include <cstdlib>

namespace b200r {

constexpr int RF_ROWS = 128;                        // corpus rows per tile (MMA M)
constexpr int RF_BK = 32;                           // fp32 elements per K slice = one 128-byte swizzle row
constexpr int RF_STAGE_BYTES = RF_ROWS * RF_BK * 4; // 16 KB
constexpr int RF_NQ = 16;                           // queries per corpus pass (MMA N)
constexpr int RF_Q_KB_BYTES = 2 * RF_NQ * RF_BK * 4;   // hi rows then mid rows of one K slice: 4 KB
constexpr int RF_RING_BYTES = RF_STAGE_BYTES + RF_Q_KB_BYTES;   // a ring stage: the corpus slice and the matching query slice
constexpr int RF_THREADS = 320;                     // TMA, MMA, 4 epilogue warps, 4 converter warps
// The tensor core adds each instruction's partial sum into the fp32 accumulator with truncation, a bias that grows with the
// length of the accumulation chain (measured 1.5e-5 relative on 768 same-sign terms, above the fp32 bar).  The K slices of a
// tile are therefore dealt round-robin onto RF_PARTS accumulators (chains 8x shorter) that the epilogue adds in fp32 RN.
constexpr int RF_PARTS_SMEM = 8, RF_PARTS_TMEM = 4;
constexpr int RF_PART_COLS = 2 * 16;                // a partial accumulator: columns 0..15 = x.q_hi terms, 16..31 = x_hi.q_mid
constexpr int RF_ACC = 2;                           // tiles in flight in TMEM
// MID_TMEM variant: the mid piece never touches shared memory.  The kernel is bound by shared-memory bandwidth when the
// converters write a mid tile that the MMA then reads back (90 KB of traffic per 16 KB slice, 0.68 of the HBM roofline
// measured); here each converter thread owns one corpus ROW (conflict-free reads thanks to the 128-byte swizzle), and
// stores its 32 mid values into TMEM with tcgen05.st, where the MMA takes them as its A operand (58 KB per slice).
// TMEM columns: 2 tiles x 4 partial accumulators x 32 + 3 mid buffers x 32 = 352 (512 allocated).
constexpr int RF_MID_TMEM_BUFS = 3;
constexpr int RF_MID_TMEM_COL0 = RF_ACC * RF_PARTS_TMEM * RF_PART_COLS;   // 256
constexpr int RF_MAX_STAGES = 10;
constexpr u32 RF_TF32_MASK = 0xffffe000u;

struct RowsF32Params {
  int B, k, D, num_kb, stages, mids;          // raw ring stages, mid side buffers
  int backoff;             // 1: waiting epilogue warps sleep between polls
  int kps;                 // K blocks (32-float slices) per ring stage / barrier hand-off, 1..4 (see rows_f32_kps; selects the instantiation)
  int store_hi;            // 1: the converters also write the truncated hi piece back in place (see rows_f32_store_hi)
  int q0;                  // this launch serves queries [q0, min(B, q0 + 16))
  long long N;
  int tile_begin, tile_end, tiles_per_chunk, num_units;
  int cap, cap_soft, Bpad;
  const float* row_mult;
  u64* lists;
  int* counts;
  u32* gthr;
  u32 idesc, idesc32;      // N = 16 / N = 32 instruction descriptors
};

// ---- query split: f32 [B, D] -> per group of 16 queries: planes f32 [2][16][D] (hi, mid) -------------------------------
__global__ void split_queries_tf32_kernel(const float* __restrict__ Q, int B, int D, float* __restrict__ qws, size_t group_floats) {
  const int g = blockIdx.x / RF_NQ, j = blockIdx.x % RF_NQ;
  const int query = g * RF_NQ + j;
  float* hi = qws + (size_t)g * group_floats + (size_t)j * D;
  float* mid = hi + (size_t)RF_NQ * D;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float v = query < B ? Q[(size_t)query * D + d] : 0.0f;
    const float h = __uint_as_float(__float_as_uint(v) & RF_TF32_MASK);
    hi[d] = h;
    mid[d] = __uint_as_float(__float_as_uint(v - h) & RF_TF32_MASK);
  }
}

__device__ __forceinline__ void umma_tf32(u32 tmem_d, u64 desc_a, u64 desc_b, u32 idesc, u32 accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// A operand from tensor memory (128 lanes = rows, one 32-bit column per K element), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(u32 tmem_d, u32 tmem_a, u64 desc_b, u32 idesc, u32 accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(u32 taddr, const u32* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
        "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <bool MID_TMEM, int KPS>
__global__ void __launch_bounds__(RF_THREADS, 1)
rows_f32_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_x, const RowsF32Params p) {
  constexpr int RF_PARTS = MID_TMEM ? RF_PARTS_TMEM : RF_PARTS_SMEM;
  constexpr int RF_ACC_STRIDE = RF_PARTS * RF_PART_COLS;
  extern __shared__ uint8_t smem_raw[];
  const u32 raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  // ring stage = KPS x [128 x 128 B corpus slice] followed by KPS x [32 x 128 B query slice: hi rows 0..15, mid rows 16..31].
  // The query planes (2 x 16 x D floats, L2-resident) are streamed with the corpus instead of being kept in shared memory:
  // that leaves 144-200 KB of corpus reads in flight per SM at any D.
  uint8_t* smem_x = smem;                                         // [stages][20 KB]
  uint8_t* smem_m = smem_x + (size_t)p.stages * KPS * RF_RING_BYTES;    // [mids][128 x 128 B] mid
  uint8_t* tail = smem_m + (MID_TMEM ? (size_t)0 : (size_t)p.mids * RF_STAGE_BYTES);
  u32* smem_hist = reinterpret_cast<u32*>(tail);                // [4][256]
  float* thr_s = reinterpret_cast<float*>(smem_hist + 4 * 256); // [4][16]
  u64* kth_s = reinterpret_cast<u64*>(thr_s + 4 * RF_NQ);       // [4][16]
  int* cnt_s = reinterpret_cast<int*>(kth_s + 4 * RF_NQ);       // [4][16]
  u64* bars = reinterpret_cast<u64*>(cnt_s + 4 * RF_NQ);
  u64* full_bar = bars;                               // [stages]  TMA landed
  u64* conv_bar = bars + RF_MAX_STAGES;               // [stages]  split done (4 converter warps)
  u64* empty_bar = bars + 2 * RF_MAX_STAGES;          // [stages]  MMAs retired
  u64* mid_empty_bar = bars + 3 * RF_MAX_STAGES;      // [mids] (<= 4)
  u64* tmem_full_bar = mid_empty_bar + 4;             // [RF_ACC]
  u64* tmem_empty_bar = tmem_full_bar + RF_ACC;       // [RF_ACC]
  u32* tmem_ptr_smem = reinterpret_cast<u32*>(tmem_empty_bar + RF_ACC);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stages = p.stages;
  constexpr int kps = KPS;                                     // K blocks (32-float slices) per ring stage = per barrier hand-off
  const int num_kp = (p.num_kb + kps - 1) / kps;               // stages (hand-offs) per tile
  constexpr size_t stage_bytes = (size_t)kps * RF_RING_BYTES;  // [kps corpus slices | kps query slices]
  constexpr size_t q_off = (size_t)kps * RF_STAGE_BYTES;       // query slices start behind the corpus slices of the stage
  constexpr u32 TMEM_COLS = 512;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_q);
    prefetch_tmap(&tmap_x);
    for (int i = 0; i < stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&conv_bar[i], 4); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < p.mids; ++i) mbar_init(&mid_empty_bar[i], 1);
    for (int i = 0; i < RF_ACC; ++i) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], 4); }
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
      int stage = 0; u32 phase = 0;
      for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
        const int t0 = p.tile_begin + u * p.tiles_per_chunk;
        const int t1 = min(p.tile_end, t0 + p.tiles_per_chunk);
        for (int t = t0; t < t1; ++t) {
          for (int kp = 0; kp < num_kp; ++kp) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_expect_tx(&full_bar[stage], (u32)stage_bytes);
            uint8_t* sb = smem_x + (size_t)stage * stage_bytes;
#pragma unroll
            for (int h = 0; h < kps; ++h) {          // a K block past D is wholly out of bounds: TMA zero-fills it
              const int kb = kp * kps + h;
              tma_load_2d(sb + (size_t)h * RF_STAGE_BYTES, &tmap_x, &full_bar[stage], kb * RF_BK, t * RF_ROWS, kEvictFirst);
              tma_load_2d(sb + q_off + (size_t)h * RF_Q_KB_BYTES, &tmap_q, &full_bar[stage], kb * RF_BK, 0, kEvictLast);
            }
            if (++stage == stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================================== MMA issuer =======================================
    if (lane == 0) {
      int stage = 0; u32 phase = 0;
      u32 tile_iter = 0;
      int m = 0;
      for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
        const int t0 = p.tile_begin + u * p.tiles_per_chunk;
        const int t1 = min(p.tile_end, t0 + p.tiles_per_chunk);
        for (int t = t0; t < t1; ++t, ++tile_iter) {
          const u32 acc = tile_iter % RF_ACC;
          mbar_wait(&tmem_empty_bar[acc], ((tile_iter / RF_ACC) & 1) ^ 1);
          tc_fence_after();
          for (int kp = 0; kp < num_kp; ++kp) {
            mbar_wait(&conv_bar[stage], phase);          // the converters have split this stage (and fenced the async proxy)
            tc_fence_after();
            const uint8_t* sb = smem_x + (size_t)stage * stage_bytes;
#pragma unroll
            for (int h = 0; h < kps; ++h) {
              const int kb = kp * kps + h;
              if (kb >= p.num_kb) break;                   // zero-filled tail block of an odd D/32: nothing to add
              const u32 tmem_d = tmem_base + acc * RF_ACC_STRIDE + (u32)(kb % RF_PARTS) * RF_PART_COLS;    // this slice's partial accumulator
              const u64 x_hi = make_smem_desc(smem_u32(sb + (size_t)h * RF_STAGE_BYTES));
              const u64 x_mid = make_smem_desc(smem_u32(smem_m + (size_t)m * RF_STAGE_BYTES));
              const u64 q_all = make_smem_desc(smem_u32(sb + q_off + (size_t)h * RF_Q_KB_BYTES));   // 32 rows: q_hi then q_mid
#pragma unroll
              for (int kk = 0; kk < RF_BK / 8; ++kk) {      // K = 8 tf32 = 32 bytes per instruction
                // x_hi . [q_hi ; q_mid] -> 32 columns, then x_mid . q_hi onto the first 16
                umma_tf32(tmem_d, x_hi + (u64)(kk * 2), q_all + (u64)(kk * 2), p.idesc32, (kb >= RF_PARTS || kk != 0) ? 1u : 0u);
                if constexpr (MID_TMEM) umma_tf32_ts(tmem_d, tmem_base + RF_MID_TMEM_COL0 + (u32)(m * kps + h) * 32u + (u32)kk * 8u, q_all + (u64)(kk * 2), p.idesc, 1u);
                else umma_tf32(tmem_d, x_mid + (u64)(kk * 2), q_all + (u64)(kk * 2), p.idesc, 1u);
              }
            }
            umma_commit(&empty_bar[stage]);
            umma_commit(&mid_empty_bar[m]);
            if (kp == num_kp - 1) umma_commit(&tmem_full_bar[acc]);
            if (++m == p.mids) m = 0;
            if (++stage == stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp_idx >= 6) {
    // ===================================== converters ========================================
    // 128 threads split one 16 KB slice: thread t handles the 16-byte chunks t, t + 128, ... (conflict-free, layout-agnostic:
    // the split is elementwise, so the mid tile keeps the raw tile's swizzled layout)
    const int ct = threadIdx.x - 6 * 32;
    int stage = 0; u32 phase = 0;
    int m = 0; u32 mphase = 0;
    for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
      const int t0 = p.tile_begin + u * p.tiles_per_chunk;
      const int t1 = min(p.tile_end, t0 + p.tiles_per_chunk);
      for (int t = t0; t < t1; ++t) {
        for (int kp = 0; kp < num_kp; ++kp) {
          mbar_wait(&mid_empty_bar[m], mphase ^ 1u);                 // the MMAs that read this mid buffer `mids` stages ago retired
          mbar_wait(&full_bar[stage], phase);
          if constexpr (MID_TMEM) {
            // thread = corpus row (the TMEM lane its warp may address): its 128 bytes per K block, un-swizzled chunk by chunk
            const int row = (warp_idx & 3) * 32 + lane;
#pragma unroll
            for (int h = 0; h < kps; ++h) {
              const uint8_t* rp = smem_x + (size_t)stage * stage_bytes + (size_t)h * RF_STAGE_BYTES + (size_t)row * 128;
              u32 r[32];
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const uint4 v = *reinterpret_cast<const uint4*>(rp + ((c ^ (row & 7)) << 4));
                const u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  r[c * 4 + e] = __float_as_uint(__uint_as_float(w[e]) - __uint_as_float(w[e] & RF_TF32_MASK)) & RF_TF32_MASK;
              }
              tmem_st32(tmem_base + ((u32)((warp_idx & 3) * 32) << 16) + RF_MID_TMEM_COL0 + (u32)(m * kps + h) * 32u, r);
            }
            tmem_st_wait();
            tc_fence_before();
          } else {
          uint4* raw = reinterpret_cast<uint4*>(smem_x + (size_t)stage * RF_RING_BYTES);
          uint4* mid = reinterpret_cast<uint4*>(smem_m + (size_t)m * RF_STAGE_BYTES);
#pragma unroll
          for (int c = 0; c < RF_STAGE_BYTES / 16 / 128; ++c) {
            const uint4 v = raw[c * 128 + ct];
            uint4 h, r;
            h.x = v.x & RF_TF32_MASK; h.y = v.y & RF_TF32_MASK; h.z = v.z & RF_TF32_MASK; h.w = v.w & RF_TF32_MASK;
            r.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(h.x)) & RF_TF32_MASK;
            r.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(h.y)) & RF_TF32_MASK;
            r.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(h.z)) & RF_TF32_MASK;
            r.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(h.w)) & RF_TF32_MASK;
            if (p.store_hi) raw[c * 128 + ct] = h;
            mid[c * 128 + ct] = r;
          }
          fence_proxy_async_smem();            // generic-proxy stores -> visible to the tensor core's (async proxy) reads
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&conv_bar[stage]);
          if (++m == p.mids) { m = 0; mphase ^= 1u; }
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================================== epilogue =========================================
    const int quarter = warp_idx & 3;          // TMEM lane quarter this warp may access == rows 32*quarter.. of the tile
    const int ew = warp_idx - 2;               // 0..3: index of this warp's lists / thresholds
    u32* hist = smem_hist + ew * 256;
    float* thr = thr_s + ew * RF_NQ;
    u64* kth = kth_s + ew * RF_NQ;
    int* cnt = cnt_s + ew * RF_NQ;
    u32 tile_iter = 0;
    for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
      const int t0 = p.tile_begin + u * p.tiles_per_chunk;
      const int t1 = min(p.tile_end, t0 + p.tiles_per_chunk);
      if (t0 >= t1) continue;
      const size_t seg = (size_t)u * 4 + ew;               // this warp's list segment for the unit's row chunk
      if (lane < RF_NQ) {
        const bool valid = p.q0 + lane < p.B;
        thr[lane] = valid ? ordered_to_f32(__ldcg(p.gthr + p.q0 + lane)) : __int_as_float(0x7f800000);
        kth[lane] = 0;
        cnt[lane] = valid ? __ldcg(p.counts + seg * p.Bpad + p.q0 + lane) : 0;
      }
      __syncwarp();
      for (int t = t0; t < t1; ++t, ++tile_iter) {
        const u32 acc = tile_iter % RF_ACC;
        const long long row = (long long)t * RF_ROWS + quarter * 32 + lane;
        const float m = (row < p.N) ? (p.row_mult ? __ldg(p.row_mult + row) : 1.0f) : __int_as_float(0x7fc00000);
        // a tile takes ~10 us to arrive: sleep between polls so the spin does not take issue slots from the converters
        if (p.backoff) mbar_wait_backoff(&tmem_full_bar[acc], (tile_iter / RF_ACC) & 1, 200);
        else mbar_wait(&tmem_full_bar[acc], (tile_iter / RF_ACC) & 1);
        tc_fence_after();
        const u32 taddr = tmem_base + ((u32)(quarter * 32) << 16) + acc * RF_ACC_STRIDE;
        const int nparts = p.num_kb < RF_PARTS ? p.num_kb : RF_PARTS;      // partial accumulators this tile wrote
        float s[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) s[j] = 0.0f;
        for (int part = 0; part < nparts; ++part) {
          u32 a[32];
          tmem_ld32(taddr + part * RF_PART_COLS, a);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) s[j] += __uint_as_float(a[j]) + __uint_as_float(a[16 + j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        bool any = false;
#pragma unroll
        for (int j = 0; j < 16; ++j) { s[j] *= m; any |= (s[j] >= thr[j]); }   // NaN m (rows beyond N) never passes
        if (__any_sync(kFullMask, any)) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float tj = thr[j];
            const u32 b = __ballot_sync(kFullMask, s[j] >= tj);
            if (b) {                                          // warp-uniform
              const u64 key = make_key(s[j], (u32)row);
              const u32 b2 = __ballot_sync(kFullMask, (s[j] >= tj) && key > kth[j]);
              if (b2) {
                const int base = cnt[j];
                u64* lp = p.lists + (seg * p.Bpad + p.q0 + j) * (size_t)p.cap;
                if ((b2 >> lane) & 1u) __stcg(lp + base + __popc(b2 & ((1u << lane) - 1u)), key);
                const int n = base + __popc(b2);
                __syncwarp();
                if (n >= p.cap_soft) {
                  const u64 kk = warp_compact(lp, n, p.k, hist);
                  if (lane == 0) {
                    kth[j] = kk; thr[j] = fmaxf(tj, key_score(kk)); cnt[j] = p.k;
                    atomicMax(p.gthr + p.q0 + j, (u32)(kk >> 32));
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
        if (lane < RF_NQ && p.q0 + lane < p.B) thr[lane] = fmaxf(thr[lane], ordered_to_f32(__ldcg(p.gthr + p.q0 + lane)));
        __syncwarp();
      }
      if (lane < RF_NQ && p.q0 + lane < p.B) p.counts[seg * p.Bpad + p.q0 + lane] = cnt[lane];
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
// kind::tf32 reads the top 19 bits of each fp32 operand and IGNORES the low 13 mantissa bits (truncation), so the raw corpus
// slice already is the hi piece and the in-place store (a sixth of the kernel's shared-memory traffic, which is what bounds
// it) can be skipped; tests/test_gpu_search.py::test_fp32_split_is_tight_on_nearly_parallel_vectors would fail by three
// orders of magnitude if the hardware rounded instead.  B200R_F32_STORE_HI=1 restores the store.
static int rows_f32_store_hi() {
  static const int v = [] { const char* e = getenv("B200R_F32_STORE_HI"); return (e && e[0] == '1') ? 1 : 0; }();
  return v;
}
// B200R_F32_MID_TMEM=0 selects the variant that keeps the mid piece in shared memory (A/B measurements)
static bool rows_f32_mid_tmem() {
  static const bool v = [] { const char* e = getenv("B200R_F32_MID_TMEM"); return !(e && e[0] == '0'); }();
  return v;
}
// K blocks (32-float slices) per ring stage = per barrier hand-off (TMEM variant).  The kernel is bound by the serial chain
// of hand-offs per slice (TMA -> converters -> MMA thread -> commit -> release), not by HBM, shared memory or the tensor
// pipe: with 16 KB per hand-off it streams 5 M x 768 fp32 in 2.98 ms (0.79 of the measured copy bandwidth), with 32 KB in
// 2.49 ms (0.94), with 48 KB in 2.43 ms (0.97), with 64 KB (only two stages fit) in 2.58 ms (0.91) -- same box, back to
// back (profiles/r2_f32_kblocks_per_stage.jsonl).  Default: 3, or 2 when D/32 is even but not a multiple of 3 (a stage whose
// tail K blocks lie past D still moves zero-filled boxes into shared memory).  B200R_F32_KPS = 1..4 overrides.
static int rows_f32_kps(int D) {
  static const int env = [] { const char* e = getenv("B200R_F32_KPS"); const int k = e ? atoi(e) : 0; return (k >= 1 && k <= 4) ? k : 0; }();
  if (!rows_f32_mid_tmem()) return 1;
  if (env) return env;
  const int num_kb = (D + RF_BK - 1) / RF_BK;
  if (num_kb <= 2) return num_kb;
  if (num_kb % 3 == 0) return 3;
  return (num_kb % 2 == 0) ? 2 : 3;
}
static int rows_f32_mids(int D) {
  // TMEM variant: buffers of 32 x kps columns behind the 256 accumulator columns (<= 256 more)
  if (rows_f32_mid_tmem()) { const int fit = 8 / rows_f32_kps(D); return fit < RF_MID_TMEM_BUFS ? fit : RF_MID_TMEM_BUFS; }
  static const int m = [] { const char* e = getenv("B200R_F32_MIDS"); const int v = e ? atoi(e) : 3; return (v >= 2 && v <= 4) ? v : 3; }();
  return m;
}
static size_t rows_f32_smem_bytes(int stages, int D) {
  return 1024 + (size_t)stages * rows_f32_kps(D) * RF_RING_BYTES + (rows_f32_mid_tmem() ? 0 : (size_t)rows_f32_mids(D) * RF_STAGE_BYTES) + 4 * 256 * 4 +
         (size_t)4 * RF_NQ * (4 + 8 + 4) + (3 * RF_MAX_STAGES + 4 + 2 * RF_ACC + 2) * 8 + 16;
}
static int rows_f32_stages_for(int D) {
  int stages = RF_MAX_STAGES;
  while (stages >= 2 && rows_f32_smem_bytes(stages, D) > 227 * 1024) --stages;
  return stages >= 2 ? stages : 0;
}

int rows_f32_tile_rows() { return RF_ROWS; }
// queries per corpus pass (16); 0 = switched off (B200R_ROWS_F32=0)
int rows_f32_nq(int D) {
  static const bool off = [] { const char* e = getenv("B200R_ROWS_F32"); return e && e[0] == '0'; }();
  return (!off && rows_f32_stages_for(D)) ? RF_NQ : 0;
}

static size_t rows_f32_group_floats(int D) { return (size_t)2 * RF_NQ * D; }
size_t rows_f32_query_bytes(int B, int D) { return align_up((size_t)((B + RF_NQ - 1) / RF_NQ) * rows_f32_group_floats(D) * 4, 256); }

int launch_split_queries_tf32(const void* Q, int B, int D, void* qws, cudaStream_t st) {
  ProfScope _ps(st, "split_queries_tf32");
  const int groups = (B + RF_NQ - 1) / RF_NQ;
  split_queries_tf32_kernel<<<groups * RF_NQ, 256, 0, st>>>(reinterpret_cast<const float*>(Q), B, D, reinterpret_cast<float*>(qws),
                                                          rows_f32_group_floats(D));
  count_launch();
  B200R_CUDA_CHECK(cudaGetLastError());
  return B200R_OK;
}

int encode_query_tmap_f32(const void* qws, int group, int D, CUtensorMap* tmap_q) {
  const float* base = reinterpret_cast<const float*>(qws) + (size_t)group * rows_f32_group_floats(D);
  return encode_tmap_2d(tmap_q, B200R_F32, base, (uint64_t)D, (uint64_t)(2 * RF_NQ), RF_BK, (uint32_t)(2 * RF_NQ));
}

// one launch = one phase of one group of <= 16 queries
int launch_rows_f32(const Index* ix, const CUtensorMap* tmap_q, int group, const SearchPlan& pl, const GemmPhase& ph, void* ws,
                    cudaStream_t st) {
  ProfScope _ps(st, "rows_f32");
  RowsF32Params p;
  p.B = pl.B; p.k = pl.k; p.D = ix->D; p.N = ix->N;
  p.q0 = group * RF_NQ;
  p.num_kb = (ix->D + RF_BK - 1) / RF_BK;
  p.stages = rows_f32_stages_for(ix->D);
  p.mids = rows_f32_mids(ix->D);
  p.kps = rows_f32_kps(ix->D);
  p.store_hi = rows_f32_store_hi();
  { static const int b = [] { const char* e = getenv("B200R_F32_BACKOFF"); return (e && e[0] == '0') ? 0 : 1; }(); p.backoff = b; }
  if (!p.stages) { set_error("fp32 tensor path: D=%d does not fit", ix->D); return B200R_EINVAL; }
  p.tile_begin = ph.tile_begin; p.tile_end = ph.tile_end; p.tiles_per_chunk = ph.tiles_per_chunk;
  p.num_units = ph.num_chunks;
  p.cap = pl.cap; p.cap_soft = pl.cap_soft; p.Bpad = pl.Bpad;
  p.row_mult = ix->row_mult;
  p.lists = reinterpret_cast<u64*>((char*)ws + pl.off_lists);
  p.counts = reinterpret_cast<int*>((char*)ws + pl.off_counts);
  p.gthr = reinterpret_cast<u32*>((char*)ws + pl.off_gthr);
  // instruction descriptor: D = f32 (1 << 4), A = B = tf32 (2 << 7, 2 << 10), K-major both, N >> 3, M >> 4
  p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((u32)(RF_NQ >> 3) << 17) | ((u32)(RF_ROWS >> 4) << 24);
  p.idesc32 = (1u << 4) | (2u << 7) | (2u << 10) | ((u32)((2 * RF_NQ) >> 3) << 17) | ((u32)(RF_ROWS >> 4) << 24);
  const size_t smem = rows_f32_smem_bytes(p.stages, ix->D);
  const int grid = p.num_units < num_sms() ? p.num_units : num_sms();
#define B200R_F32_LAUNCH(TMEM, KPS)                                                                                              \
  do {                                                                                                                         \
    B200R_CUDA_CHECK(cudaFuncSetAttribute(rows_f32_kernel<TMEM, KPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    rows_f32_kernel<TMEM, KPS><<<grid, RF_THREADS, smem, st>>>(*tmap_q, ix->tmap_x_f32, p);                                     \
  } while (0)
  if (!rows_f32_mid_tmem()) B200R_F32_LAUNCH(false, 1);
  else if (p.kps == 4) B200R_F32_LAUNCH(true, 4);
  else if (p.kps == 3) B200R_F32_LAUNCH(true, 3);
  else if (p.kps == 2) B200R_F32_LAUNCH(true, 2);
  else B200R_F32_LAUNCH(true, 1);
#undef B200R_F32_LAUNCH
  count_launch();
  B200R_CUDA_CHECK(cudaGetLastError());
  return B200R_OK;
}

}  // namespace b200r
