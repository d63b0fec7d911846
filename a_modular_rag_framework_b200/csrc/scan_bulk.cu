// scan_bulk.cu -- K2b: the streaming scan fed by bulk copies (fp32 and int8 corpora, up to 8 queries per pass).
//
// K2 (scan_topk.cu) holds its in-flight loads in registers, which caps both the bytes a warp can have outstanding
// (int8: a warp-wide 128-bit load covers only 512 B of a row) and the number of accumulators left for extra queries.
// Here the corpus is moved by the copy engine instead: every CTA owns a contiguous block of rows, ONE elected thread
// issues one cp.async.bulk (global -> shared, mbarrier complete_tx) per tile of W*R consecutive rows (a contiguous byte
// range, so no tensor map is needed) into a ring of stages, and W consumer warps read their R rows of the tile from
// shared memory with conflict-free LDS.128, multiply against the queries staged as float4 "parts" (read once per R
// rows; int8 corpora instead run on DP4A against the query split into three signed-byte limbs), butterfly-reduce, and keep the same exact running top-k per (warp, query) as K2: threshold + list + radix
// select (common.cuh), finished by select_lists_kernel.  Registers are free for R*BQ accumulators, so one pass over the
// corpus serves up to 8 fp32 queries.
//
// Same contract as K1/K2: exact cosine / inner-product top-k (reference app/modules/retrieval/retrieval_backend.py
// :192-197 + :371-372).  Roofline: HBM, N*D*sizeof(elem) bytes per pass.
#include "common.cuh"
#include "internal.h"
#include "tc_ptx.cuh"

namespace b200r {

struct ScanBulkParams {
  const void* X;
  const void* Q;       // [B, D] in the query dtype
  long long N;
  int D, B, k;
  int q0, nq;          // queries [q0, q0+nq) handled by this launch
  int cap, cap_soft, Bpad;
  long long rows_per_cta;   // multiple of the tile height
  int stages;
  unsigned tile_bytes;      // W*R*row_bytes
  const float* row_mult;
  u64* lists;
  int* counts;
  u32* gthr;
};

__device__ __forceinline__ void bulk_load(u32 smem_dst, const void* gsrc, u32 bytes, u64* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// the ring is addressed through the 32-bit shared window (a generic pointer would compile to LD.E + 64-bit address math)
__device__ __forceinline__ uint4 lds128(u32 addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}

// T: corpus element type, BQ: queries per pass, R: rows per warp per tile, W: consumer warps (warp W is the producer)
template <typename T, int BQ, int R, int W>
__global__ void __launch_bounds__((W + 1) * 32, 1) scan_bulk_kernel(const ScanBulkParams p) {
  typedef typename QueryType<T>::type QT;
  constexpr int VEC = ElemTraits<T>::kVec;
  constexpr int PARTS = VEC / 4;
  constexpr int TR = W * R;
  constexpr bool kInt8 = (sizeof(T) == 1);
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int chunks = p.D / VEC;
  float4* q4 = reinterpret_cast<float4*>(smem_raw);                                   // [PARTS][BQ][chunks]
  u32* hist_all = reinterpret_cast<u32*>(q4 + (size_t)PARTS * BQ * chunks);          // [W][256]
  u64* full = reinterpret_cast<u64*>(hist_all + W * 256);                             // [stages]
  u64* empty = full + p.stages;                                                       // [stages]
  const u32 ring = (smem_u32(empty + p.stages) + 127u) & ~127u;                       // shared-window address, 128-aligned

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const QT* Qp = reinterpret_cast<const QT*>(p.Q);
  __shared__ u32 s_qmax[BQ];      // int8 path: bit image of max|q_b| (non-negative floats order like their bits)
  __shared__ float s_scale[BQ];   // int8 path: q_b ~= scale_b * (l2*65536 + l1*256 + l0)
  if constexpr (kInt8) {
    // int8 corpus: the dot products run on DP4A.  Each query is quantised ONCE to 23-bit fixed point relative to its
    // largest component (|q - q~| <= 2^-22 max|q|, about an fp32 rounding of that component) and split into three
    // balanced signed-byte limbs, staged as 16-byte planes qL[limb][b][chunk]; the integer sums are exact.
    if (threadIdx.x < BQ) s_qmax[threadIdx.x] = 0u;
    __syncthreads();
    for (int b = 0; b < BQ; ++b) {
      float m = 0.0f;
      if (b < p.nq)
        for (int d = threadIdx.x; d < p.D; d += (W + 1) * 32) m = fmaxf(m, fabsf(q_to_float<QT>(Qp[(size_t)(p.q0 + b) * p.D + d])));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(kFullMask, m, o));
      if (lane == 0) atomicMax(&s_qmax[b], __float_as_uint(m));
    }
    __syncthreads();
    int8_t* qL = reinterpret_cast<int8_t*>(smem_raw);
    for (int i = threadIdx.x; i < BQ * p.D; i += (W + 1) * 32) {
      const int b = i / p.D, d = i % p.D;
      const float qmax = __uint_as_float(s_qmax[b]);
      const float v = (b < p.nq && qmax > 0.0f) ? q_to_float<QT>(Qp[(size_t)(p.q0 + b) * p.D + d]) * (4194304.0f / qmax) : 0.0f;
      const int qi = __float2int_rn(v);                       // |qi| <= 2^22
      const int l0 = ((qi + 128) & 255) - 128;
      const int q1 = (qi - l0) >> 8;
      const int l1 = ((q1 + 128) & 255) - 128;
      const int l2 = (q1 - l1) >> 8;                          // |l2| <= 64
      const int chunk = d / VEC, e = d % VEC;
      qL[(((size_t)0 * BQ + b) * chunks + chunk) * 16 + e] = (int8_t)l0;
      qL[(((size_t)1 * BQ + b) * chunks + chunk) * 16 + e] = (int8_t)l1;
      qL[(((size_t)2 * BQ + b) * chunks + chunk) * 16 + e] = (int8_t)l2;
    }
    if (threadIdx.x < BQ) s_scale[threadIdx.x] = __uint_as_float(s_qmax[threadIdx.x]) * (1.0f / 4194304.0f);
  } else {
    for (int i = threadIdx.x; i < BQ * p.D; i += (W + 1) * 32) {
      const int b = i / p.D, d = i % p.D;
      const float v = (b < p.nq) ? q_to_float<QT>(Qp[(size_t)(p.q0 + b) * p.D + d]) : 0.0f;
      const int chunk = d / VEC, e = d % VEC;
      reinterpret_cast<float*>(&q4[((size_t)(e / 4) * BQ + b) * chunks + chunk])[e % 4] = v;
    }
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, W); }
    fence_barrier_init();
  }
  __syncthreads();

  const long long cta_begin = (long long)blockIdx.x * p.rows_per_cta;
  const long long cta_end = min(p.N, cta_begin + p.rows_per_cta);
  const long long cta_rows = cta_end > cta_begin ? cta_end - cta_begin : 0;
  const int ntiles = (int)((cta_rows + TR - 1) / TR);
  const u32 row_bytes = (u32)p.D * (u32)sizeof(T);
  const uint8_t* Xb = reinterpret_cast<const uint8_t*>(p.X);

  if (warp == W) {
    // ---- producer: one bulk copy per tile ----------------------------------------------------------
    if (lane == 0) {
      for (int t = 0; t < ntiles; ++t) {
        const int s = t % p.stages;
        if (t >= p.stages) mbar_wait(empty + s, (u32)((t / p.stages - 1) & 1));
        const long long r0 = cta_begin + (long long)t * TR;
        const u32 bytes = (u32)(min((long long)TR, cta_end - r0) * (long long)row_bytes);
        mbar_expect_tx(full + s, bytes);
        bulk_load(ring + (u32)s * p.tile_bytes, Xb + (size_t)r0 * row_bytes, bytes, full + s);
      }
    }
    return;
  }

  // ---- consumers ---------------------------------------------------------------------------------------
  u32* hist = hist_all + warp * 256;
  const long long gw = (long long)blockIdx.x * W + warp;
  float thr[BQ];
  u64 kth[BQ];
  int cnt[BQ];
  u64* list[BQ];
#pragma unroll
  for (int b = 0; b < BQ; ++b) {
    const bool valid = b < p.nq;
    thr[b] = valid ? ordered_to_f32(__ldcg(p.gthr + p.q0 + b)) : __int_as_float(0x7f800000);
    kth[b] = 0;
    cnt[b] = 0;
    list[b] = p.lists + ((size_t)gw * p.Bpad + (size_t)(p.q0 + (valid ? b : 0))) * (size_t)p.cap;
  }

  // lane r holds row r's multiplier (NaN = past the end: never passes), fetched ONE TILE AHEAD: the ring usually has the
  // next tile resident already, so a load issued at the top of the tile would be waited for in full
  auto load_mult = [&](int t) {
    const long long row = cta_begin + (long long)t * TR + (long long)warp * R + lane;
    float m = __int_as_float(0x7fc00000);
    if (lane < R && t < ntiles && row < cta_end) m = p.row_mult ? __ldg(p.row_mult + row) : 1.0f;
    return m;
  };
  float next_mult = load_mult(0);
  for (int t = 0; t < ntiles; ++t) {
    const int s = t % p.stages;
    const long long r0 = cta_begin + (long long)t * TR + (long long)warp * R;     // this warp's first row of the tile
    const float my_mult = next_mult;
    next_mult = load_mult(t + 1);
    mbar_wait(full + s, (u32)((t / p.stages) & 1));
    const u32 rows = ring + (u32)s * p.tile_bytes + (u32)(warp * R) * row_bytes + (u32)lane * 16u;   // this lane's first chunk

    float acc[R][BQ];
    if constexpr (kInt8) {
      int ai[R][BQ][3];
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int b = 0; b < BQ; ++b) { ai[r][b][0] = 0; ai[r][b][1] = 0; ai[r][b][2] = 0; }
      const uint4* qL = reinterpret_cast<const uint4*>(smem_raw);
      u32 xaddr = rows;
      for (int chunk = lane; chunk < chunks; chunk += 32, xaddr += 512u) {
        uint4 xv[R];
#pragma unroll
        for (int r = 0; r < R; ++r) xv[r] = lds128(xaddr + (u32)r * row_bytes);
#pragma unroll
        for (int limb = 0; limb < 3; ++limb) {
#pragma unroll
          for (int b = 0; b < BQ; ++b) {
            const uint4 qv = qL[((size_t)limb * BQ + b) * chunks + chunk];
#pragma unroll
            for (int r = 0; r < R; ++r) {
              int a = ai[r][b][limb];
              a = __dp4a((int)xv[r].x, (int)qv.x, a);
              a = __dp4a((int)xv[r].y, (int)qv.y, a);
              a = __dp4a((int)xv[r].z, (int)qv.z, a);
              a = __dp4a((int)xv[r].w, (int)qv.w, a);
              ai[r][b][limb] = a;
            }
          }
        }
      }
      // per-lane partial sums are exact integers (< 2^24 each); combining the limbs rounds like any fp32 sum
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int b = 0; b < BQ; ++b)
          acc[r][b] = fmaf((float)ai[r][b][2], 65536.0f, fmaf((float)ai[r][b][1], 256.0f, (float)ai[r][b][0]));
    } else {
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int b = 0; b < BQ; ++b) acc[r][b] = 0.0f;
      u32 xaddr = rows;
      for (int chunk = lane; chunk < chunks; chunk += 32, xaddr += 512u) {
        uint4 xv[R];
#pragma unroll
        for (int r = 0; r < R; ++r) xv[r] = lds128(xaddr + (u32)r * row_bytes);
#pragma unroll
        for (int part = 0; part < PARTS; ++part) {
          float4 qv[BQ];
#pragma unroll
          for (int b = 0; b < BQ; ++b) qv[b] = q4[((size_t)part * BQ + b) * chunks + chunk];
#pragma unroll
          for (int r = 0; r < R; ++r) {
            float xf[4];
            ElemTraits<T>::unpack4(xv[r], part, xf);
#pragma unroll
            for (int b = 0; b < BQ; ++b) {
              acc[r][b] = fmaf(xf[0], qv[b].x, acc[r][b]);
              acc[r][b] = fmaf(xf[1], qv[b].y, acc[r][b]);
              acc[r][b] = fmaf(xf[2], qv[b].z, acc[r][b]);
              acc[r][b] = fmaf(xf[3], qv[b].w, acc[r][b]);
            }
          }
        }
      }
    }
    // the tile's bytes are consumed: hand the stage back to the producer before the (shuffle-heavy) tail
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + s);

    // butterfly: every lane ends with the full sums (same order for every row => identical rows tie exactly)
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int b = 0; b < BQ; ++b) {
        float v = acc[r][b];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
        acc[r][b] = v;
      }
    bool pass = false;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float m = __shfl_sync(kFullMask, my_mult, r);
#pragma unroll
      for (int b = 0; b < BQ; ++b) {
        if constexpr (kInt8) acc[r][b] *= s_scale[b];
        acc[r][b] *= m;
        pass |= acc[r][b] >= thr[b];
      }
    }
    if (pass) {  // warp-uniform: every lane holds the same sums and thresholds
#pragma unroll
      for (int r = 0; r < R; ++r) {
#pragma unroll
        for (int b = 0; b < BQ; ++b) {
          if (acc[r][b] >= thr[b]) {
            const u64 key = make_key(acc[r][b], (u32)(r0 + r));
            if (key > kth[b]) {
              if (lane == 0) __stcg(list[b] + cnt[b], key);
              ++cnt[b];
              if (cnt[b] >= p.cap_soft) {
                kth[b] = warp_compact(list[b], cnt[b], p.k, hist);
                thr[b] = fmaxf(thr[b], key_score(kth[b]));
                cnt[b] = p.k;
                if (lane == 0) atomicMax(p.gthr + p.q0 + b, (u32)(kth[b] >> 32));
              }
            }
          }
        }
      }
    }
    if ((t & 15) == 15) {
#pragma unroll
      for (int b = 0; b < BQ; ++b)
        if (b < p.nq) thr[b] = fmaxf(thr[b], ordered_to_f32(__ldcg(p.gthr + p.q0 + b)));
    }
  }
#pragma unroll
  for (int b = 0; b < BQ; ++b) {
    if (b < p.nq) {
      if (cnt[b] > p.k) { warp_compact(list[b], cnt[b], p.k, hist); cnt[b] = p.k; }
      if (lane == 0) p.counts[(size_t)gw * p.Bpad + p.q0 + b] = cnt[b];
    }
  }
}

// ---- host side -----------------------------------------------------------------------------------------
// shapes: int8 rows are short -> 16 consumer warps x 4 rows (64-row tiles); fp32 rows are long -> 8 warps x 2 rows
template <typename T> struct BulkShape;
template <> struct BulkShape<int8_t> { static constexpr int W = 16, R = 4, BQ_MAX = 4; };
template <> struct BulkShape<float> { static constexpr int W = 8, R = 2, BQ_MAX = 8; };

constexpr size_t kBulkSmemBudget = 200 * 1024;

template <typename T>
static size_t bulk_fixed_smem(int D, int bq) {
  return (size_t)bq * D * 4 + (size_t)BulkShape<T>::W * 256 * 4 + 2 * 8 * 8 /* barriers, <= 8 stages */ + 128;
}

template <typename T>
static int bulk_stages(int D, int bq) {
  const size_t tile = (size_t)BulkShape<T>::W * BulkShape<T>::R * D * sizeof(T);
  const size_t fixed = bulk_fixed_smem<T>(D, bq);
  if (fixed + 2 * tile > kBulkSmemBudget) return 0;
  const size_t s = (kBulkSmemBudget - fixed) / tile;
  return (int)(s > 6 ? 6 : s);
}

// 0 = not available for this index (other dtype, rows too long for two stages, or switched off for an A/B run)
int scan_bulk_warps(const Index* ix) {
  static const bool off = [] { const char* e = getenv("B200R_SCAN_BULK"); return e && e[0] == '0'; }();
  if (off) return 0;
  if (ix->dtype == B200R_I8) return bulk_stages<int8_t>(ix->D, BulkShape<int8_t>::BQ_MAX) ? BulkShape<int8_t>::W : 0;
  if (ix->dtype == B200R_F32) return bulk_stages<float>(ix->D, BulkShape<float>::BQ_MAX) ? BulkShape<float>::W : 0;
  return 0;
}
int scan_bulk_tile_rows(const Index* ix) {
  return ix->dtype == B200R_I8 ? BulkShape<int8_t>::W * BulkShape<int8_t>::R : BulkShape<float>::W * BulkShape<float>::R;
}

template <typename T, int BQ>
static int launch_bulk_bq(const ScanBulkParams& p0, int blocks, cudaStream_t st) {
  constexpr int W = BulkShape<T>::W, R = BulkShape<T>::R;
  ScanBulkParams p = p0;
  p.stages = bulk_stages<T>(p.D, BQ);
  p.tile_bytes = (unsigned)((size_t)W * R * p.D * sizeof(T));
  const size_t smem = bulk_fixed_smem<T>(p.D, BQ) + (size_t)p.stages * p.tile_bytes;
  B200R_CUDA_CHECK(cudaFuncSetAttribute(scan_bulk_kernel<T, BQ, R, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  scan_bulk_kernel<T, BQ, R, W><<<blocks, (W + 1) * 32, smem, st>>>(p);
  count_launch();
  B200R_CUDA_CHECK(cudaGetLastError());
  return B200R_OK;
}

template <typename T>
static int launch_bulk_t(const Index* ix, const void* Q, const SearchPlan& pl, void* ws, cudaStream_t st) {
  constexpr int TR = BulkShape<T>::W * BulkShape<T>::R, BQ_MAX = BulkShape<T>::BQ_MAX;
  ScanBulkParams p;
  p.X = ix->X; p.Q = Q; p.N = ix->N; p.D = ix->D; p.B = pl.B; p.k = pl.k;
  p.cap = pl.cap; p.cap_soft = pl.cap_soft; p.Bpad = pl.Bpad;
  p.row_mult = ix->row_mult;
  p.lists = reinterpret_cast<u64*>((char*)ws + pl.off_lists);
  p.counts = reinterpret_cast<int*>((char*)ws + pl.off_counts);
  p.gthr = reinterpret_cast<u32*>((char*)ws + pl.off_gthr);
  long long rpc = (ix->N + pl.scan_blocks - 1) / pl.scan_blocks;
  p.rows_per_cta = (rpc + TR - 1) / TR * TR;
  for (int q0 = 0; q0 < pl.B; q0 += BQ_MAX) {
    p.q0 = q0;
    p.nq = pl.B - q0 < BQ_MAX ? pl.B - q0 : BQ_MAX;
    int rc;
    if (p.nq == 1) rc = launch_bulk_bq<T, 1>(p, pl.scan_blocks, st);
    else if (p.nq == 2) rc = launch_bulk_bq<T, 2>(p, pl.scan_blocks, st);
    else if (p.nq <= 4 || BQ_MAX == 4) rc = launch_bulk_bq<T, 4>(p, pl.scan_blocks, st);
    else rc = launch_bulk_bq<T, (BQ_MAX >= 8 ? 8 : 4)>(p, pl.scan_blocks, st);
    if (rc) return rc;
  }
  return B200R_OK;
}

int launch_scan_bulk(const Index* ix, const void* Q, const SearchPlan& pl, void* ws, cudaStream_t st) {
  ProfScope _ps(st, "scan_bulk");
  if (ix->dtype == B200R_I8) return launch_bulk_t<int8_t>(ix, Q, pl, ws, st);
  if (ix->dtype == B200R_F32) return launch_bulk_t<float>(ix, Q, pl, ws, st);
  set_error("bulk scan: unsupported dtype %d", ix->dtype);
  return B200R_EINVAL;
}

}  // namespace b200r
