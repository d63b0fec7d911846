// gather_bw.cu -- micro-benchmark: what HBM bandwidth can a RANDOM ROW GATHER reach on this GPU?
// (the access pattern of K3's re-score: ~2.2 k random 1536-byte corpus rows per query).  It bounds K3's roofline:
// random 1.5 KB reads do not stream like a copy does.  Two ways of moving a row are timed:
//   ldg  : a warp keeps ROWS rows in flight with 128-bit ld.global.nc loads into registers (K3 v2's scheme)
//   bulk : one lane issues ONE cp.async.bulk (global -> shared, mbarrier complete_tx) per row into a per-warp ring
// Both dot the row with a query held in shared memory, like the kernel does.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bw scripts/gather_bw.cu && ./gather_bw
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint4 ld_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float dot8(const uint4 v, const float* q) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  float a = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    a = fmaf(__uint_as_float(w[i] << 16), q[2 * i], a);
    a = fmaf(__uint_as_float(w[i] & 0xffff0000u), q[2 * i + 1], a);
  }
  return a;
}

__global__ void make_idx(uint32_t* idx, size_t n, uint32_t nrows, uint32_t seed) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint64_t x = (i + 1) * 0x9E3779B97F4A7C15ull + seed;
    x ^= x >> 31; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
    idx[i] = (uint32_t)(x % nrows);
  }
}

template <int ROWS>
__global__ void gather_ldg(const uint8_t* __restrict__ X, const uint32_t* __restrict__ idx, size_t n, int D, float* out) {
  extern __shared__ float qs[];
  for (int d = threadIdx.x; d < D; d += blockDim.x) qs[d] = 1.0f / (float)(d + 1);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const int chunks = D / 8;
  const size_t row_bytes = (size_t)D * 2;
  for (size_t g = warp * ROWS; g < n; g += nwarps * ROWS) {
    const uint8_t* rp[ROWS];
    float acc[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) { rp[r] = X + (size_t)idx[min(g + r, n - 1)] * row_bytes; acc[r] = 0.f; }
    for (int c = lane; c < chunks; c += 32) {
      uint4 v[ROWS];
#pragma unroll
      for (int r = 0; r < ROWS; ++r) v[r] = ld_stream(rp[r] + (size_t)c * 16);
#pragma unroll
      for (int r = 0; r < ROWS; ++r) acc[r] += dot8(v[r], qs + c * 8);
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      float a = acc[r];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0 && g + r < n) out[g + r] = a;
    }
  }
}

// per-warp ring of STAGES row slots in shared memory, one cp.async.bulk per row
template <int STAGES>
__global__ void gather_bulk(const uint8_t* __restrict__ X, const uint32_t* __restrict__ idx, size_t n, int D, float* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int nw = blockDim.x >> 5, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t row_bytes = (uint32_t)D * 2;
  float* qs = reinterpret_cast<float*>(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (((size_t)D * 4 + 127) & ~(size_t)127));
  uint8_t* ring = reinterpret_cast<uint8_t*>(bars) + (((size_t)nw * STAGES * 8 + 127) & ~(size_t)127);
  for (int d = threadIdx.x; d < D; d += blockDim.x) qs[d] = 1.0f / (float)(d + 1);
  uint64_t* mybar = bars + w * STAGES;
  uint8_t* myring = ring + (size_t)w * STAGES * row_bytes;
  if (lane == 0) {
    for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mybar + s)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const size_t warp = blockIdx.x * (size_t)nw + w, nwarps = (size_t)gridDim.x * nw;
  auto issue = [&](size_t i, int s) {
    const uint8_t* src = X + (size_t)idx[i] * row_bytes;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mybar + s)), "r"(row_bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(myring + (size_t)s * row_bytes)), "l"(src), "r"(row_bytes), "r"(smem_u32(mybar + s)) : "memory");
  };
  if (lane == 0) {
    for (int s = 0; s < STAGES; ++s) { const size_t i = warp + (size_t)s * nwarps; if (i < n) issue(i, s); }
  }
  const int chunks = D / 8;
  int s = 0; uint32_t parity = 0;
  for (size_t i = warp; i < n; i += nwarps) {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(mybar + s)), "r"(parity) : "memory");
    }
    const uint8_t* row = myring + (size_t)s * row_bytes;
    float a = 0.f;
    for (int c = lane; c < chunks; c += 32) a += dot8(*reinterpret_cast<const uint4*>(row + (size_t)c * 16), qs + c * 8);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    __syncwarp();
    if (lane == 0) {
      out[i] = a;
      const size_t nx = i + (size_t)STAGES * nwarps;
      if (nx < n) issue(nx, s);
    }
    if (++s == STAGES) { s = 0; parity ^= 1; }
  }
}

// streaming read of the same number of bytes (sequential rows) for reference
__global__ void stream_rows(const uint8_t* __restrict__ X, size_t nbytes, float* out) {
  float a = 0.f;
  const size_t n16 = nbytes / 16;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = ld_stream(X + i * 16);
    a += __uint_as_float(v.x << 16) + __uint_as_float(v.y << 16) + __uint_as_float(v.z << 16) + __uint_as_float(v.w << 16);
  }
  if (a == 123.456f) out[0] = a;
}

template <typename F>
static float time_ms(F f, int reps = 5) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main(int argc, char** argv) {
  const uint32_t nrows_bytes_gb = 8;   // table size ~7.7 GB regardless of D
  int sms = 148; cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0)); sms = prop.multiProcessorCount;
  const size_t n = (size_t)(argc > 1 ? atol(argv[1]) : 3000000);
  float* out; CK(cudaMalloc(&out, n * 4));
  uint32_t* idx; CK(cudaMalloc(&idx, n * 4));
  const size_t table_bytes = (size_t)5000000 * 1536;
  uint8_t* X; CK(cudaMalloc(&X, table_bytes)); CK(cudaMemset(X, 0x3c, table_bytes));
  (void)nrows_bytes_gb;
  {
    float ms = time_ms([&] { stream_rows<<<sms * 8, 512>>>(X, table_bytes, out); });
    printf("{\"variant\": \"stream\", \"GBps\": %.1f}\n", table_bytes / ms * 1e-6);
  }
  for (int D : {384, 768, 1536}) {
    const uint32_t nrows = (uint32_t)(table_bytes / ((size_t)D * 2));
    make_idx<<<sms * 4, 256>>>(idx, n, nrows, 17u * D); CK(cudaDeviceSynchronize());
    const double bytes = (double)n * D * 2;
#define RUN_LDG(ROWS, THREADS, BPS) { \
      float ms = time_ms([&] { gather_ldg<ROWS><<<sms * BPS, THREADS, D * 4>>>(X, idx, n, D, out); }); \
      printf("{\"variant\": \"ldg\", \"D\": %d, \"rows_in_flight_per_warp\": %d, \"threads\": %d, \"ctas_per_sm\": %d, \"inflight_KB_per_sm\": %.0f, \"GBps\": %.1f}\n", \
             D, ROWS, THREADS, BPS, (double)ROWS * (THREADS / 32) * BPS * D * 2 / 1024.0, bytes / ms * 1e-6); fflush(stdout); }
    RUN_LDG(1, 512, 1) RUN_LDG(1, 512, 2) RUN_LDG(1, 512, 4)
    RUN_LDG(2, 512, 1) RUN_LDG(2, 512, 2) RUN_LDG(2, 512, 4)
    RUN_LDG(4, 512, 1) RUN_LDG(4, 512, 2) RUN_LDG(4, 512, 3) RUN_LDG(4, 512, 4)
    RUN_LDG(8, 512, 1) RUN_LDG(8, 512, 2)
#define RUN_BULK(STAGES, THREADS, BPS) { \
      const size_t sm = (((size_t)D * 4 + 127) & ~(size_t)127) + ((((size_t)(THREADS / 32) * STAGES * 8) + 127) & ~(size_t)127) + (size_t)(THREADS / 32) * STAGES * D * 2; \
      if (sm * BPS <= 220 * 1024) { \
      CK(cudaFuncSetAttribute(gather_bulk<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
      float ms = time_ms([&] { gather_bulk<STAGES><<<sms * BPS, THREADS, sm>>>(X, idx, n, D, out); }); \
      printf("{\"variant\": \"bulk\", \"D\": %d, \"stages_per_warp\": %d, \"threads\": %d, \"ctas_per_sm\": %d, \"inflight_KB_per_sm\": %.0f, \"GBps\": %.1f}\n", \
             D, STAGES, THREADS, BPS, (double)STAGES * (THREADS / 32) * BPS * D * 2 / 1024.0, bytes / ms * 1e-6); fflush(stdout); } }
    RUN_BULK(2, 256, 1) RUN_BULK(2, 512, 1) RUN_BULK(4, 256, 1) RUN_BULK(4, 512, 1) RUN_BULK(8, 256, 1) RUN_BULK(8, 512, 1)
    RUN_BULK(2, 512, 2) RUN_BULK(4, 512, 2) RUN_BULK(4, 256, 2) RUN_BULK(8, 256, 2) RUN_BULK(2, 512, 3) RUN_BULK(2, 128, 1) RUN_BULK(4, 128, 1)
    RUN_BULK(8, 128, 1) RUN_BULK(8, 128, 2)
  }
  return 0;
}
