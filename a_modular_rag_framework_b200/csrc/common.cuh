// common.cuh -- shared device helpers: ordered 64-bit candidate keys and the warp-cooperative
// exact k-th selection / compaction that every search kernel uses for its running top-k lists.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace b200r {

typedef unsigned long long u64;
typedef unsigned int u32;

constexpr u32 kFullMask = 0xffffffffu;

// ---------------------------------------------------------------------------------------------
// Candidate key: (order-preserving image of the f32 score) << 32 | (0xffffffff - local row id).
// A larger key is a better candidate under the reference's ranking rule: score descending, then
// lower id first (stable sort over rows enumerated in id order, retrieval_backend.py:371-372).
// Key 0 is never produced by a real candidate and is used as padding.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ u32 f32_to_ordered(float f) {
#ifdef __CUDA_ARCH__
  u32 u = __float_as_uint(f);
#else
  union { float f; u32 u; } c; c.f = f; u32 u = c.u;
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_f32(u32 o) {
  u32 u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; u32 u; } c; c.u = u; return c.f;
#endif
}
__device__ __forceinline__ u64 make_key(float s, u32 local_id) {
  s = s + 0.0f;  // -0.0 -> +0.0 so that both zeros tie (python: -0.0 == 0.0)
  return ((u64)f32_to_ordered(s) << 32) | (u64)(0xffffffffu - local_id);
}
__device__ __forceinline__ float key_score(u64 key) { return ordered_to_f32((u32)(key >> 32)); }
__device__ __forceinline__ u32 key_id(u64 key) { return 0xffffffffu - (u32)(key & 0xffffffffu); }

constexpr u32 kOrderedNegInf = 0x007fffffu;  // f32_to_ordered(-inf)

// ---------------------------------------------------------------------------------------------
// warp_select_kth: exact k-th largest of keys[0..n) (n >= k >= 1), all 32 lanes cooperate.
// MSB-first radix select, 8 bits per pass, histogram in 256 words of shared memory owned by the warp.
// Keys are unique (ids are), so exactly k keys are >= the returned value.
// ---------------------------------------------------------------------------------------------
// Lists are read with ld.global.cg / written with st.global.cg (L2 only): they are produced and
// consumed inside one kernel by different lanes, so the non-coherent L1 path must not be used.
__device__ __forceinline__ u64 warp_select_kth(const u64* keys, int n, int k, u32* hist) {
  const int lane = threadIdx.x & 31;
  u64 prefix = 0;
  int remaining = k;
#pragma unroll 1
  for (int pass = 7; pass >= 0; --pass) {
    const int shift = pass * 8;
    for (int i = lane; i < 256; i += 32) hist[i] = 0;
    __syncwarp();
    const u64 hi_mask = (pass == 7) ? 0ull : (~0ull << (shift + 8));
    for (int i = lane; i < n; i += 32) {
      const u64 key = __ldcg(keys + i);
      if ((key & hi_mask) == prefix) atomicAdd(&hist[(u32)(key >> shift) & 255u], 1u);
    }
    __syncwarp();
    // lane l owns bins [8l, 8l+8); higher bins = larger keys
    u32 c[8];
    u32 lane_sum = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b) { c[b] = hist[lane * 8 + b]; lane_sum += c[b]; }
    u32 suffix = lane_sum;  // inclusive suffix sum over lanes >= this lane
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 t = __shfl_down_sync(kFullMask, suffix, o);
      if (lane + o < 32) suffix += t;
    }
    const u32 above = suffix - lane_sum;
    const bool mine = (above < (u32)remaining) && ((u32)remaining <= suffix);
    int digit = 0, new_remaining = 0;
    if (mine) {
      u32 run = above;
#pragma unroll
      for (int b = 7; b >= 0; --b) {
        if (run + c[b] >= (u32)remaining) { digit = lane * 8 + b; new_remaining = remaining - (int)run; break; }
        run += c[b];
      }
    }
    const int src = __ffs(__ballot_sync(kFullMask, mine)) - 1;
    digit = __shfl_sync(kFullMask, digit, src);
    remaining = __shfl_sync(kFullMask, new_remaining, src);
    prefix |= (u64)(u32)digit << shift;
    __syncwarp();
  }
  return prefix;
}

// warp_compact: keep the k largest keys of keys[0..n) in keys[0..k) (unordered); returns the k-th largest.
__device__ __forceinline__ u64 warp_compact(u64* keys, int n, int k, u32* hist) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
  const u64 kth = warp_select_kth(keys, n, k, hist);
  int out = 0;
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    const u64 key = (i < n) ? __ldcg(keys + i) : 0ull;
    const bool keep = (i < n) && (key >= kth);
    const u32 m = __ballot_sync(kFullMask, keep);
    __syncwarp();  // all lanes have read this round before anyone overwrites (out + rank <= i)
    if (keep) __stcg(keys + out + __popc(m & ((1u << lane) - 1u)), key);
    out += __popc(m);
  }
  __syncwarp();
  return kth;
}

// ---------------------------------------------------------------------------------------------
// element loads: 16 bytes -> floats
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ld_stream_16B(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

template <typename T> struct ElemTraits;
template <> struct ElemTraits<__nv_bfloat16> {
  static constexpr int kVec = 8;
  __device__ static __forceinline__ void unpack(const uint4& v, float* f) {
    const u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  }
  // elements [4*part, 4*part+4) only (part is a compile-time constant at every call site)
  __device__ static __forceinline__ void unpack4(const uint4& v, int part, float* f) {
    const u32 w0 = part == 0 ? v.x : v.z, w1 = part == 0 ? v.y : v.w;
    f[0] = __uint_as_float(w0 << 16); f[1] = __uint_as_float(w0 & 0xffff0000u);
    f[2] = __uint_as_float(w1 << 16); f[3] = __uint_as_float(w1 & 0xffff0000u);
  }
};
template <> struct ElemTraits<__half> {
  static constexpr int kVec = 8;
  __device__ static __forceinline__ void unpack(const uint4& v, float* f) {
    const u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
      const float2 t = __half22float2(h);
      f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
  }
  __device__ static __forceinline__ void unpack4(const uint4& v, int part, float* f) {
    const u32 w0 = part == 0 ? v.x : v.z, w1 = part == 0 ? v.y : v.w;
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&w0));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&w1));
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
  }
};
template <> struct ElemTraits<float> {
  static constexpr int kVec = 4;
  __device__ static __forceinline__ void unpack(const uint4& v, float* f) {
    f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y); f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
  }
  __device__ static __forceinline__ void unpack4(const uint4& v, int, float* f) { unpack(v, f); }
};
template <> struct ElemTraits<int8_t> {
  static constexpr int kVec = 16;
  __device__ static __forceinline__ void unpack(const uint4& v, float* f) {
    // int8 -> float without I2F: put (b ^ 0x80) into the mantissa of 2^23 and subtract 2^23 + 128
    const u32 w[4] = {v.x ^ 0x80808080u, v.y ^ 0x80808080u, v.z ^ 0x80808080u, v.w ^ 0x80808080u};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[4 * i + 0] = __uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7650)) - 8388736.0f;
      f[4 * i + 1] = __uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7651)) - 8388736.0f;
      f[4 * i + 2] = __uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7652)) - 8388736.0f;
      f[4 * i + 3] = __uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7653)) - 8388736.0f;
    }
  }
  __device__ static __forceinline__ void unpack4(const uint4& v, int part, float* f) {
    const u32 w = (part == 0 ? v.x : part == 1 ? v.y : part == 2 ? v.z : v.w) ^ 0x80808080u;
    f[0] = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650)) - 8388736.0f;
    f[1] = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7651)) - 8388736.0f;
    f[2] = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7652)) - 8388736.0f;
    f[3] = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7653)) - 8388736.0f;
  }
};

// query element -> float, and the query type that goes with a corpus type (int8 corpora take f32 queries)
template <typename QT> __device__ __forceinline__ float q_to_float(QT v);
template <> __device__ __forceinline__ float q_to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float q_to_float<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float q_to_float<float>(float v) { return v; }
template <typename T> struct QueryType { typedef T type; };
template <> struct QueryType<int8_t> { typedef float type; };

// bitonic sort of n (power of two) u64 keys in shared memory, descending, whole block
__device__ __forceinline__ void block_bitonic_desc(u64* a, int n) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < n / 2; t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const u64 x = a[lo], y = a[hi];
        if ((x < y) == desc) { a[lo] = y; a[hi] = x; }
      }
    }
  }
  __syncthreads();
}

}  // namespace b200r
