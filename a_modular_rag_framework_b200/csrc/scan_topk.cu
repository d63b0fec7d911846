// scan_topk.cu -- K2: memory-bound streaming scan for small query batches (B <= 4 per pass).
//
// Same contract as K1 (exact cosine / inner-product top-k of reference
// app/modules/retrieval/retrieval_backend.py:192-197 + 371-372) but on CUDA cores: every warp streams
// a contiguous range of corpus rows with 128-bit L1-bypassing loads, R rows in flight, fp32
// accumulation against the queries held in shared memory, and keeps a running exact top-k per query
// with the same threshold + list + radix-select machinery as K1 (common.cuh).  Supports bf16, fp16,
// fp32 and int8 (per-row scale) corpora.  Also hosts the small per-index / per-search helper kernels.
#include "common.cuh"
#include "internal.h"

namespace b200r {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_WARPS = SCAN_THREADS / 32;

struct ScanParams {
  const void* X;
  const void* Q;       // [B, D] in the query dtype
  long long N;
  int D, B, k;
  int q0, nq;          // queries [q0, q0+nq) handled by this launch
  int cap, cap_soft, Bpad;
  long long rows_per_warp;
  const float* row_mult;
  u64* lists;
  int* counts;
  u32* gthr;
};

// T: corpus element type, BQ: queries per pass, R: rows in flight per warp
template <typename T, int BQ, int R>
__global__ void __launch_bounds__(SCAN_THREADS, 3) scan_topk_kernel(const ScanParams p) {
  typedef typename QueryType<T>::type QT;
  constexpr int VEC = ElemTraits<T>::kVec;
  constexpr int PARTS = VEC / 4;
  extern __shared__ uint8_t smem_raw[];
  // q4[part][b][chunk]: the 4*part.. 4*part+3 -th elements of 16-byte chunk `chunk` of query b (conflict-free float4 reads)
  float4* q4 = reinterpret_cast<float4*>(smem_raw);
  const int chunks = p.D / VEC;
  u32* hist_all = reinterpret_cast<u32*>(q4 + (size_t)PARTS * BQ * chunks);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  u32* hist = hist_all + warp * 256;

  const QT* Qp = reinterpret_cast<const QT*>(p.Q);
  for (int i = threadIdx.x; i < BQ * p.D; i += SCAN_THREADS) {
    const int b = i / p.D, d = i % p.D;
    const float v = (b < p.nq) ? q_to_float<QT>(Qp[(size_t)(p.q0 + b) * p.D + d]) : 0.0f;
    const int chunk = d / VEC, e = d % VEC;
    reinterpret_cast<float*>(&q4[((size_t)(e / 4) * BQ + b) * chunks + chunk])[e % 4] = v;
  }
  __syncthreads();

  const long long gw = (long long)blockIdx.x * SCAN_WARPS + warp;
  const long long row_begin = gw * p.rows_per_warp;
  const long long row_end = min(p.N, row_begin + p.rows_per_warp);

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

  const uint8_t* Xb = reinterpret_cast<const uint8_t*>(p.X);
  const size_t row_bytes = (size_t)p.D * sizeof(T);
  int iter = 0;
  for (long long r0 = row_begin; r0 < row_end; r0 += R, ++iter) {
    float acc[R][BQ];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int b = 0; b < BQ; ++b) acc[r][b] = 0.0f;

    const uint8_t* rowp[R];
#pragma unroll
    for (int r = 0; r < R; ++r) rowp[r] = Xb + (size_t)min(r0 + r, row_end - 1) * row_bytes;
    // lane r fetches row r's multiplier now, so its latency hides behind the row loads (NaN = past the end: never passes)
    float my_mult = __int_as_float(0x7fc00000);
    if (lane < R && r0 + lane < row_end) my_mult = p.row_mult ? __ldg(p.row_mult + r0 + lane) : 1.0f;

    if constexpr (PARTS == 4) {
      // int8 rows are short (D bytes): a warp-wide 128-bit load covers only 512 B of a row, so TWO chunk iterations of
      // every row are kept in flight together, and the values are consumed one 4-element part at a time so that only 4
      // unpacked values per row are live
      for (int cb = 0; cb < chunks; cb += 64) {
        uint4 xv[2][R];
        int ch[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = cb + h * 32 + lane;
          const bool ok = c < chunks;
          ch[h] = ok ? c : 0;
#pragma unroll
          for (int r = 0; r < R; ++r) xv[h][r] = ok ? ld_stream_16B(rowp[r] + (size_t)c * 16) : make_uint4(0, 0, 0, 0);   // int8 zeros add 0
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int part = 0; part < PARTS; ++part) {
            float4 qv[BQ];
#pragma unroll
            for (int b = 0; b < BQ; ++b) qv[b] = q4[((size_t)part * BQ + b) * chunks + ch[h]];
#pragma unroll
            for (int r = 0; r < R; ++r) {
              float xf[4];
              ElemTraits<T>::unpack4(xv[h][r], part, xf);
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
    } else {
      for (int chunk = lane; chunk < chunks; chunk += 32) {
        uint4 xv[R];
#pragma unroll
        for (int r = 0; r < R; ++r) xv[r] = ld_stream_16B(rowp[r] + (size_t)chunk * 16);
        float xf[R][VEC];
#pragma unroll
        for (int r = 0; r < R; ++r) ElemTraits<T>::unpack(xv[r], xf[r]);
#pragma unroll
        for (int part = 0; part < PARTS; ++part) {
#pragma unroll
          for (int b = 0; b < BQ; ++b) {
            const float4 qv = q4[((size_t)part * BQ + b) * chunks + chunk];
#pragma unroll
            for (int r = 0; r < R; ++r) {
              acc[r][b] = fmaf(xf[r][part * 4 + 0], qv.x, acc[r][b]);
              acc[r][b] = fmaf(xf[r][part * 4 + 1], qv.y, acc[r][b]);
              acc[r][b] = fmaf(xf[r][part * 4 + 2], qv.z, acc[r][b]);
              acc[r][b] = fmaf(xf[r][part * 4 + 3], qv.w, acc[r][b]);
            }
          }
        }
      }
    }
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
    if ((iter & 15) == 15) {
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

template <typename T, int R1, int R4>
static int launch_scan_t(const Index* ix, const void* Q, const SearchPlan& pl, void* ws, cudaStream_t st) {
  constexpr int VEC = ElemTraits<T>::kVec;
  ScanParams p;
  p.X = ix->X; p.Q = Q; p.N = ix->N; p.D = ix->D; p.B = pl.B; p.k = pl.k;
  p.cap = pl.cap; p.cap_soft = pl.cap_soft; p.Bpad = pl.Bpad;
  p.row_mult = ix->row_mult;
  p.lists = reinterpret_cast<u64*>((char*)ws + pl.off_lists);
  p.counts = reinterpret_cast<int*>((char*)ws + pl.off_counts);
  p.gthr = reinterpret_cast<u32*>((char*)ws + pl.off_gthr);
  const long long total_warps = (long long)pl.scan_blocks * SCAN_WARPS;
  long long rpw = (ix->N + total_warps - 1) / total_warps;
  rpw = (rpw + R1 - 1) / R1 * R1;   // R1 is a multiple of R4
  p.rows_per_warp = rpw;
  const int chunks = ix->D / VEC;
  for (int q0 = 0; q0 < pl.B; q0 += 4) {
    p.q0 = q0;
    p.nq = pl.B - q0 < 4 ? pl.B - q0 : 4;
    const int bq = p.nq >= 3 ? 4 : p.nq;
    const size_t smem = (size_t)(VEC / 4) * bq * chunks * sizeof(float4) + SCAN_WARPS * 256 * sizeof(u32);
    if (smem > 200 * 1024) { set_error("scan path: D=%d too large for shared-memory query staging", ix->D); return B200R_EINVAL; }
#define B200R_LAUNCH_SCAN(BQ, R)                                                                           \
    do {                                                                                                   \
      B200R_CUDA_CHECK(cudaFuncSetAttribute(scan_topk_kernel<T, BQ, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      scan_topk_kernel<T, BQ, R><<<pl.scan_blocks, SCAN_THREADS, smem, st>>>(p);                          \
    } while (0)
    if (bq == 1) B200R_LAUNCH_SCAN(1, R1);
    else if (bq == 2) B200R_LAUNCH_SCAN(2, R4);
    else B200R_LAUNCH_SCAN(4, R4);
#undef B200R_LAUNCH_SCAN
    count_launch();
    B200R_CUDA_CHECK(cudaGetLastError());
  }
  return B200R_OK;
}

int launch_scan_topk(const Index* ix, const void* Q, const SearchPlan& pl, void* ws, cudaStream_t st) {
  ProfScope _ps(st, "scan_topk");
  switch (ix->dtype) {
    case B200R_BF16: return launch_scan_t<__nv_bfloat16, 8, 4>(ix, Q, pl, ws, st);
    case B200R_F16: return launch_scan_t<__half, 8, 4>(ix, Q, pl, ws, st);
    case B200R_F32: return launch_scan_t<float, 8, 4>(ix, Q, pl, ws, st);
    case B200R_I8: return launch_scan_t<int8_t, 4, 2>(ix, Q, pl, ws, st);   // x2 chunk iterations in flight per row
  }
  set_error("bad dtype %d", ix->dtype);
  return B200R_EINVAL;
}

// ------------------------------------------------------------------------------------------------
// helper kernels
// ------------------------------------------------------------------------------------------------

// one warp per row: mult[n] = 1/|x_n| (0 when the norm is 0, like _cosine's zero-norm guard) or the int8 row scale
template <typename T>
__global__ void row_mult_kernel(const void* X, long long N, int D, int metric, const float* row_scale, float* mult) {
  constexpr int VEC = ElemTraits<T>::kVec;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  if (metric == B200R_IP) {
    if (lane == 0) mult[row] = row_scale ? row_scale[row] : 1.0f;
    return;
  }
  const uint8_t* rp = reinterpret_cast<const uint8_t*>(X) + (size_t)row * D * sizeof(T);
  double acc = 0.0;
  for (int chunk = lane; chunk < D / VEC; chunk += 32) {
    const uint4 v = ld_stream_16B(rp + (size_t)chunk * 16);
    float f[VEC];
    ElemTraits<T>::unpack(v, f);
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc += (double)f[e] * (double)f[e];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFullMask, acc, o);
  if (lane == 0) mult[row] = acc > 0.0 ? (float)(1.0 / sqrt(acc)) : 0.0f;
}

int launch_row_mult(const Index* ix, const float* row_scale, cudaStream_t st) {
  if (!ix->row_mult) return B200R_OK;
  const int wpb = 8;
  const unsigned blocks = (unsigned)((ix->N + wpb - 1) / wpb);
  switch (ix->dtype) {
    case B200R_BF16: row_mult_kernel<__nv_bfloat16><<<blocks, wpb * 32, 0, st>>>(ix->X, ix->N, ix->D, ix->metric, row_scale, ix->row_mult); break;
    case B200R_F16: row_mult_kernel<__half><<<blocks, wpb * 32, 0, st>>>(ix->X, ix->N, ix->D, ix->metric, row_scale, ix->row_mult); break;
    case B200R_F32: row_mult_kernel<float><<<blocks, wpb * 32, 0, st>>>(ix->X, ix->N, ix->D, ix->metric, row_scale, ix->row_mult); break;
    case B200R_I8: row_mult_kernel<int8_t><<<blocks, wpb * 32, 0, st>>>(ix->X, ix->N, ix->D, ix->metric, row_scale, ix->row_mult); break;
    default: set_error("bad dtype"); return B200R_EINVAL;
  }
  count_launch();
  B200R_CUDA_CHECK(cudaGetLastError());
  return B200R_OK;
}

// one warp per query: qinv[b] = 1/|q_b| (0 for a zero query)
template <typename QT>
__global__ void query_inv_norm_kernel(const QT* Q, int B, int D, float* qinv) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  double acc = 0.0;
  for (int d = lane; d < D; d += 32) {
    const double v = (double)q_to_float<QT>(Q[(size_t)b * D + d]);
    acc += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFullMask, acc, o);
  if (lane == 0) qinv[b] = acc > 0.0 ? (float)(1.0 / sqrt(acc)) : 0.0f;
}

int launch_query_inv_norm(const void* Q, int B, int D, int qdtype, float* qinv, cudaStream_t st) {
  ProfScope _ps(st, "query_inv_norm");
  const int wpb = 4;
  const unsigned blocks = (unsigned)((B + wpb - 1) / wpb);
  switch (qdtype) {
    case B200R_BF16: query_inv_norm_kernel<__nv_bfloat16><<<blocks, wpb * 32, 0, st>>>((const __nv_bfloat16*)Q, B, D, qinv); break;
    case B200R_F16: query_inv_norm_kernel<__half><<<blocks, wpb * 32, 0, st>>>((const __half*)Q, B, D, qinv); break;
    default: query_inv_norm_kernel<float><<<blocks, wpb * 32, 0, st>>>((const float*)Q, B, D, qinv); break;
  }
  count_launch();
  B200R_CUDA_CHECK(cudaGetLastError());
  return B200R_OK;
}

__global__ void init_search_kernel(u32* gthr, int n_gthr, int* counts, long long n_counts) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long j = i; j < n_gthr; j += stride) gthr[j] = kOrderedNegInf;
  for (long long j = i; j < n_counts; j += stride) counts[j] = 0;
}

// similarity floor (b200r_search_min_score): start every query's filter at min_score expressed in the key domain
// (keys carry acc*row_mult; the final score multiplies by 1/|q|), a hair lower so rounding cannot lose a boundary hit
__global__ void min_score_gthr_kernel(u32* gthr, const float* qinv, int B, float min_score) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= B) return;
  float t = min_score;
  if (qinv) t = qinv[q] > 0.0f ? min_score / qinv[q] : (min_score > 0.0f ? __int_as_float(0x7f800000) : -__int_as_float(0x7f800000));
  t = t > 0.0f ? t * (1.0f - 2e-6f) : t * (1.0f + 2e-6f);
  atomicMax(gthr + q, f32_to_ordered(t));
}

int launch_min_score(const SearchPlan& pl, void* ws, bool cosine, float min_score, cudaStream_t st) {
  ProfScope _ps(st, "min_score");
  u32* gthr = reinterpret_cast<u32*>((char*)ws + pl.off_gthr);
  const float* qinv = cosine ? reinterpret_cast<const float*>((char*)ws + pl.off_qinv) : nullptr;
  min_score_gthr_kernel<<<(pl.B + 255) / 256, 256, 0, st>>>(gthr, qinv, pl.B, min_score);
  count_launch();
  B200R_CUDA_CHECK(cudaGetLastError());
  return B200R_OK;
}

int launch_init_search(const SearchPlan& pl, void* ws, cudaStream_t st) {
  ProfScope _ps(st, "init_search");
  u32* gthr = reinterpret_cast<u32*>((char*)ws + pl.off_gthr);
  int* counts = reinterpret_cast<int*>((char*)ws + pl.off_counts);
  const long long n_counts = (long long)(pl.num_segments + 1) * pl.Bpad;   // + the carry lists' counts
  init_search_kernel<<<num_sms(), 256, 0, st>>>(gthr, pl.Bpad, counts, n_counts);
  count_launch();
  B200R_CUDA_CHECK(cudaGetLastError());
  return B200R_OK;
}

}  // namespace b200r
