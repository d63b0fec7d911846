// select_merge.cu -- K4: exact top-k of candidate lists, ordered by (score desc, id asc).
//
//  * select_lists_kernel: final reduction of the per-unit candidate lists K1/K2 leave in the
//    workspace -> out_scores/out_ids [B, k].
//  * merge_topk_kernel: public merge of G per-shard top-k lists (b200r_merge_topk*; the shard exchange of comm.cu has its
//    own merge kernel that first waits for the peers' records).
// Both implement the reference's ranking rule -- python stable sort, descending, slice
// (app/modules/retrieval/retrieval_backend.py:371-372, retrieval_adapter.py:129-131) -- over
// candidates enumerated in id order, i.e. ties go to the lower id.
#include "common.cuh"
#include "comm.cuh"
#include "internal.h"
#include <math_constants.h>
#include <cstring>
#include <cstdlib>

namespace b200r {

constexpr int SEL_MAX_THREADS = 1024;  // block size is chosen at launch: the kernel is latency-bound, so many small CTAs
                                       // when there are thousands of queries, one wide CTA per query when there are few
constexpr int SEL_STAGE_MAX = 6144;   // keys staged in shared memory for the group-wise first level (48 KB)
constexpr int SEL_STAGE_LISTS = 2560; // ... and for the per-phase / final select over pruned lists (20 KB -> 11 CTAs/SM)
constexpr int SEL_MAX_SEGS = 1024;     // segments whose counts fit the in-smem prefix (flattened gather)
constexpr int SEL_GROUP = 64;         // segments per CTA in the first level of a two-level select
constexpr int SEL_BINS = 1024;        // linear score bins of the histogram path (see select_lists_kernel)

enum SelectMode { kSelOutput = 0, kSelThreshold = 1, kSelPartial = 2 };

struct SelectParams {
  const u64* lists;
  const int* counts;
  int S, Bpad, cap;            // input segments: lists[(s*Bpad + q)*cap + i], i < counts[s*Bpad + q]
  int seg_per_cta;             // kSelPartial: CTA (q, g) reduces segments [g*seg_per_cta, ...)
  int B, k, kp2, stage_cap, mode;
  const float* qinv;           // nullable
  u32* gthr;
  long long id_base;
  float* out_scores;
  long long* out_ids;
  u64* out_lists;              // kSelPartial: out_lists[(g*Bpad + q)*k + i]
  int* out_counts;
  // phased searches: the survivors of every per-phase prune are moved to ONE carry list per query
  // (carry[q*carry_cap + i], i < counts[S*Bpad + q]) and the per-unit lists restart empty, so a threshold select is two
  // dependent global loads + a few shared-memory passes instead of a serial walk over every unit list
  u64* carry;
  int carry_cap;
  // fused exchange (row-sharded corpora): instead of a local result, the final select writes this rank's packed
  // record [ids i64 [B,k] | scores f32 [B,k]] straight into its place in EVERY rank's exchange window (peer
  // memory over NVLink, comm.cuh), so the all-gather needs no separate collective kernel or staging copy
  CommDev comm;
  int use_comm;
  unsigned long long rec_stride;   // bytes between the records of consecutive ranks in a slot
  int final_keys;              // 1: the input lists already hold keys on the FINAL score (second level of a two-level select)
  float min_score;             // results below it are reported as padding (b200r_search_min_score)
};

// Output and partial modes rank on the FINAL score (raw key score x 1/|q|), not on the score the scan kernels filtered
// with: two different raw scores can round to the same final fp32, and the reported order must be (final score desc,
// id asc) -- the order b200r_merge_topk and the shard exchange rely on (lists strictly sorted under that rule).
__device__ __forceinline__ u64 final_key(u64 key, float qi) {
  const float s = key_score(key) * qi + 0.0f;
  return ((u64)f32_to_ordered(s) << 32) | (key & 0xffffffffull);
}

// Visit every valid key of the CTA's segments with a warp-uniform trip count: f(key, active).
template <typename F>
__device__ __forceinline__ void for_each_key(const SelectParams& p, int q, int s0, int s1, bool staged, const u64* stage,
                                             int total, bool xform, float qi, bool has_carry, F f) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int SEL_THREADS = blockDim.x, SEL_WARPS = blockDim.x >> 5;
  if (staged) {
    for (int base = warp * 32; base < total; base += SEL_THREADS) {
      const int i = base + lane;
      const bool active = i < total;
      f(active ? stage[i] : 0ull, active);
    }
  } else {
    const int s_end = s1 + (has_carry ? 1 : 0);    // pseudo segment s1 = the carry list (its count lives at index S)
    for (int s = s0 + warp; s < s_end; s += SEL_WARPS) {
      const bool is_carry = has_carry && s == s1;
      const size_t slot = (size_t)(is_carry ? p.S : s) * p.Bpad + q;
      const int c = p.counts[slot];
      const u64* lp = is_carry ? p.carry + (size_t)q * p.carry_cap : p.lists + slot * (size_t)p.cap;
      for (int base = 0; base < c; base += 32) {
        const int i = base + lane;
        const bool active = i < c;
        f(active ? (xform ? final_key(lp[i], qi) : lp[i]) : 0ull, active);
      }
    }
  }
}

// histogram increment with one shared-memory atomic per distinct digit per warp (candidate keys share
// their leading bytes, so un-aggregated atomics would serialise on one bin)
__device__ __forceinline__ void hist_add(u32* hist, bool active, u32 digit) {
  const u32 act = __ballot_sync(kFullMask, active);
  if (!act) return;
  const u32 d0 = __shfl_sync(kFullMask, digit, __ffs(act) - 1);
  if (__all_sync(kFullMask, !active || digit == d0)) {        // the leading bytes of candidate keys: everybody agrees
    if ((threadIdx.x & 31) == __ffs(act) - 1) atomicAdd(&hist[d0], (u32)__popc(act));
  } else if (active) {
    atomicAdd(&hist[digit], 1u);
  }
}

__global__ void __launch_bounds__(SEL_MAX_THREADS) select_lists_kernel(const SelectParams p) {
  const int SEL_THREADS = blockDim.x, SEL_WARPS = blockDim.x >> 5;
  extern __shared__ uint8_t smem_raw[];
  u64* stage = reinterpret_cast<u64*>(smem_raw);            // [stage_cap]
  u64* sk = stage + p.stage_cap;                            // [kp2] (output mode)
  __shared__ __align__(16) u32 hist[SEL_BINS];              // radix passes use the first 256 words
  __shared__ u64 s_prefix;
  __shared__ int s_remaining, s_total, s_nout;
  __shared__ u32 s_lo32, s_hi32, s_t32;                     // histogram path: score range of the staged keys, threshold
  __shared__ int s_bstar, s_reach;                          // ... bin of the k-th best, keys in bins >= it

  const int q = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  u32 epoch = 0;
  if (p.use_comm) epoch = comm_send_begin(p.comm);      // the peers have consumed what this slot held two exchanges ago
  // final-score keys (see final_key): everywhere except the per-phase threshold selects, which stay in filter space
  const float qi = p.qinv ? p.qinv[q] : 1.0f;
  const bool xform = p.mode != kSelThreshold && p.qinv != nullptr && !p.final_keys;
  int s0 = 0, s1 = p.S;
  if (p.mode == kSelPartial) { s0 = blockIdx.y * p.seg_per_cta; s1 = min(p.S, s0 + p.seg_per_cta); }

  // segment counts -> exclusive prefix in smem (all count loads in flight at once), then every thread fetches keys
  // of the flattened candidate array by binary search over the prefix: two dependent global loads in total
  __shared__ int seg_off[SEL_MAX_SEGS + 1];
  __shared__ int warp_tot[32];
  // the carry list rides along as one more (pseudo) segment after the real ones; a two-level select hands it to group 0
  const bool has_carry = p.carry != nullptr && (p.mode != kSelPartial || blockIdx.y == 0);
  const int nreal = s1 - s0;
  const int nseg = nreal + (has_carry ? 1 : 0);
  const bool flat = nseg <= SEL_MAX_SEGS;
  if (tid == 0) { s_total = 0; s_nout = 0; s_prefix = 0; s_lo32 = 0xffffffffu; s_hi32 = 0u; s_t32 = 0xffffffffu; }
  __syncthreads();
  if (flat) {
    // blocked scan: thread t owns segments [t*per, t*per+per)
    const int per = (nseg + SEL_THREADS - 1) / SEL_THREADS;
    int mine = 0;
    for (int j = 0; j < per; ++j) {
      const int sgi = tid * per + j;
      const int c = sgi < nseg ? p.counts[(size_t)(sgi < nreal ? s0 + sgi : p.S) * p.Bpad + q] : 0;
      if (sgi < nseg) seg_off[sgi] = c;
      mine += c;
    }
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(kFullMask, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < warp; ++w) wbase += warp_tot[w];
    int run = wbase + incl - mine;
    for (int j = 0; j < per; ++j) {
      const int sgi = tid * per + j;
      if (sgi < nseg) { const int c = seg_off[sgi]; seg_off[sgi] = run; run += c; }
    }
    if (tid == SEL_THREADS - 1) { seg_off[nseg] = run; s_total = run; }
    __syncthreads();
  } else {
    int local = (has_carry && tid == 0) ? p.counts[(size_t)p.S * p.Bpad + q] : 0;
    for (int s = s0 + tid; s < s1; s += SEL_THREADS) local += p.counts[(size_t)s * p.Bpad + q];
    if (local) atomicAdd(&s_total, local);
    __syncthreads();
  }
  const int total = s_total;
  if (p.mode == kSelThreshold && total < p.k) return;   // not enough candidates yet for a threshold
  const bool staged = total <= p.stage_cap;
  if (staged && flat) {
    // one warp per segment (the prefix gives every segment its place in the stage: deterministic, coalesced, ~6
    // instructions per 32 keys -- a per-key binary search over the prefix cost 60), four segments' first loads in flight;
    // the score range of what is staged feeds the histogram path below
    u32 lo32 = 0xffffffffu, hi32 = 0u;
    if (nseg < SEL_WARPS) {
      // fewer segments than warps (one long flat list, the second level of a two-level select): spread the KEYS over the
      // threads instead, each locating its segment by binary search over the prefix
      for (int i = tid; i < total; i += SEL_THREADS) {
        int lo = 0, hi = nseg;                       // last segment with seg_off <= i
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (seg_off[mid] <= i) lo = mid; else hi = mid; }
        const u64 raw = (lo >= nreal) ? p.carry[(size_t)q * p.carry_cap + (i - seg_off[lo])]
                                      : p.lists[((size_t)(s0 + lo) * p.Bpad + q) * (size_t)p.cap + (i - seg_off[lo])];
        const u64 key = xform ? final_key(raw, qi) : raw;
        stage[i] = key;
        lo32 = min(lo32, (u32)(key >> 32)); hi32 = max(hi32, (u32)(key >> 32));
      }
    } else
    for (int sg0 = warp; sg0 < nseg; sg0 += 4 * SEL_WARPS) {
      u64 v[4]; int o[4], c[4]; const u64* lp[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int sg = sg0 + r * SEL_WARPS;
        o[r] = sg < nseg ? seg_off[sg] : 0;
        c[r] = sg < nseg ? seg_off[sg + 1] - o[r] : 0;
        lp[r] = (sg >= nreal) ? p.carry + (size_t)q * p.carry_cap : p.lists + ((size_t)(s0 + sg) * p.Bpad + q) * (size_t)p.cap;
        v[r] = lane < c[r] ? lp[r][lane] : 0ull;
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (lane < c[r]) {
          const u64 key = xform ? final_key(v[r], qi) : v[r];
          stage[o[r] + lane] = key;
          lo32 = min(lo32, (u32)(key >> 32)); hi32 = max(hi32, (u32)(key >> 32));
        }
        for (int i = lane + 32; i < c[r]; i += 32) {
          const u64 raw = lp[r][i];
          const u64 key = xform ? final_key(raw, qi) : raw;
          stage[o[r] + i] = key;
          lo32 = min(lo32, (u32)(key >> 32)); hi32 = max(hi32, (u32)(key >> 32));
        }
      }
    }
    lo32 = __reduce_min_sync(kFullMask, lo32);
    hi32 = __reduce_max_sync(kFullMask, hi32);
    if (lane == 0 && lo32 <= hi32) { atomicMin(&s_lo32, lo32); atomicMax(&s_hi32, hi32); }
    __syncthreads();
  } else if (staged) {
    // gather the valid keys: one warp per segment, coalesced
    for (int s = s0 + warp; s < s1 + (has_carry ? 1 : 0); s += SEL_WARPS) {
      const bool is_carry = has_carry && s == s1;
      const size_t slot = (size_t)(is_carry ? p.S : s) * p.Bpad + q;
      const int c = p.counts[slot];
      if (c == 0) continue;
      int base = 0;
      if (lane == 0) base = atomicAdd(&s_nout, c);
      base = __shfl_sync(kFullMask, base, 0);
      const u64* lp = is_carry ? p.carry + (size_t)q * p.carry_cap : p.lists + slot * (size_t)p.cap;
      for (int i = lane; i < c; i += 32) stage[base + i] = xform ? final_key(lp[i], qi) : lp[i];
    }
    __syncthreads();
    if (tid == 0) s_nout = 0;
  }

  u64 kth = 0;
  // Histogram path (staged + flat; not the partial level, which must hand on exactly k keys): bin the staged keys ONCE
  // into SEL_BINS bins linear in the score between the smallest and the largest of them, find the bin b* of the k-th best
  // with one suffix scan, and take T = the smallest score found in bins >= b*.  Every key with score >= T is kept:
  // `reach` >= k of them, usually k plus a handful.  A threshold select needs nothing more (T is a valid, tight lower bound
  // of the k-th best); the final select sorts those `reach` keys and emits the best k -- unless they do not fit the sort
  // buffer (mass ties), in which case the exact radix select below runs.  One pass + two short ones instead of 3 / 8 radix
  // passes with warp-aggregated atomics: the per-phase select of a batch-1 search went from 179 k to ~35 k warp
  // instructions on its one SM.
  bool have_kth = false;
  if (staged && flat && p.mode != kSelPartial && (total > p.k || p.mode == kSelThreshold)) {
    for (int i = tid; i < SEL_BINS; i += SEL_THREADS) hist[i] = 0;
    __syncthreads();
    const float flo = ordered_to_f32(s_lo32), fhi = ordered_to_f32(s_hi32);
    const float range = fhi - flo;
    const float scale = (range > 0.0f && range < CUDART_INF_F) ? (float)(SEL_BINS - 1) / range : 0.0f;    // 0: every key in bin 0
    auto bin_of = [&](u64 key) -> int {
      const int b = (int)((key_score(key) - flo) * scale);      // monotone in the score
      return b < SEL_BINS - 1 ? b : SEL_BINS - 1;
    };
    for (int i = tid; i < total; i += SEL_THREADS) atomicAdd(&hist[bin_of(stage[i])], 1u);
    __syncthreads();
    if (tid < 32) {                                    // lane L owns bins [32 L, 32 L + 32)
      u32 w[32]; u32 lane_sum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint4 t = reinterpret_cast<const uint4*>(hist + lane * 32)[j];
        w[4 * j] = t.x; w[4 * j + 1] = t.y; w[4 * j + 2] = t.z; w[4 * j + 3] = t.w;
        lane_sum += t.x + t.y + t.z + t.w;
      }
      u32 suffix = lane_sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_down_sync(kFullMask, suffix, o); if (lane + o < 32) suffix += t; }
      const u32 above = suffix - lane_sum;
      if (above < (u32)p.k && (u32)p.k <= suffix) {    // exactly one lane (total >= k)
        u32 acc = above;
#pragma unroll
        for (int j = 31; j >= 0; --j) {
          acc += w[j];
          if (acc >= (u32)p.k) { s_bstar = lane * 32 + j; s_reach = (int)acc; break; }
        }
      }
    }
    __syncthreads();
    const int bstar = s_bstar;
    u32 t32 = 0xffffffffu;
    for (int i = tid; i < total; i += SEL_THREADS) {
      const u64 key = stage[i];
      if (bin_of(key) >= bstar) t32 = min(t32, (u32)(key >> 32));
    }
    t32 = __reduce_min_sync(kFullMask, t32);
    if (lane == 0 && t32 != 0xffffffffu) atomicMin(&s_t32, t32);
    __syncthreads();
    // reach = keys in bins >= b* = keys with score >= T (the bin is a function of the score)
    if (p.mode == kSelThreshold || s_reach <= p.kp2) { kth = (u64)s_t32 << 32; have_kth = true; }
  }
  if (!have_kth && (total > p.k || p.mode == kSelThreshold)) {
    // threshold mode stops after the top 24 bits: the truncated key is a valid (slightly low) threshold
    const int last_pass = (p.mode == kSelThreshold) ? 5 : 0;
    if (tid == 0) s_remaining = p.k;
    for (int pass = 7; pass >= last_pass; --pass) {
      const int shift = pass * 8;
      for (int i = tid; i < 256; i += SEL_THREADS) hist[i] = 0;
      __syncthreads();
      const u64 prefix = s_prefix;
      const u64 hi_mask = (pass == 7) ? 0ull : (~0ull << (shift + 8));
      for_each_key(p, q, s0, s1, staged, stage, total, xform, qi, has_carry, [&](u64 key, bool active) {
        hist_add(hist, active && (key & hi_mask) == prefix, (u32)(key >> shift) & 255u);
      });
      __syncthreads();
      if (tid < 32) {
        const int remaining = s_remaining;
        u32 c[8]; u32 lane_sum = 0;
#pragma unroll
        for (int b = 0; b < 8; ++b) { c[b] = hist[lane * 8 + b]; lane_sum += c[b]; }
        u32 suffix = lane_sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const u32 t = __shfl_down_sync(kFullMask, suffix, o);
          if (lane + o < 32) suffix += t;
        }
        const u32 above = suffix - lane_sum;
        if (above < (u32)remaining && (u32)remaining <= suffix) {
          u32 run = above;
#pragma unroll
          for (int b = 7; b >= 0; --b) {
            if (run + c[b] >= (u32)remaining) {
              s_prefix = prefix | ((u64)(u32)(lane * 8 + b) << shift);
              s_remaining = remaining - (int)run;
              break;
            }
            run += c[b];
          }
        }
      }
      __syncthreads();
    }
    kth = s_prefix;
  }

  if (p.mode == kSelThreshold) {
    if (tid == 0) atomicMax(p.gthr + q, (u32)(kth >> 32));
    // prune: keys below the (truncated) k-th key can never reach the final top-k, so drop them.
    // This keeps every later select (and the final one) at ~k + one phase's survivors per query.
    u64* wl = const_cast<u64*>(p.lists);
    int* wc = const_cast<int*>(p.counts);
    if (has_carry && staged && flat) {
      // survivors (already in shared memory) -> the carry list, compacted; the unit lists restart empty
      u64* cl = p.carry + (size_t)q * p.carry_cap;
      for (int base = warp * 32; base < total; base += SEL_THREADS) {
        const int i = base + lane;
        const u64 key = i < total ? stage[i] : 0ull;
        const bool keep = i < total && key >= kth;
        const u32 m = __ballot_sync(kFullMask, keep);
        int pos = 0;
        if (lane == 0 && m) pos = atomicAdd(&s_nout, __popc(m));
        pos = __shfl_sync(kFullMask, pos, 0);
        if (keep) cl[pos + __popc(m & ((1u << lane) - 1u))] = key;
      }
      for (int sgi = tid; sgi < nreal; sgi += SEL_THREADS) wc[(size_t)(s0 + sgi) * p.Bpad + q] = 0;
      __syncthreads();
      if (tid == 0) wc[(size_t)p.S * p.Bpad + q] = s_nout;
      return;
    }
    const int s_end = s1 + (has_carry ? 1 : 0);
    for (int s = s0 + warp; s < s_end; s += SEL_WARPS) {
      const bool is_carry = has_carry && s == s1;
      const size_t slot = (size_t)(is_carry ? p.S : s) * p.Bpad + q;
      const int c = wc[slot];
      if (c == 0) continue;
      u64* lp = is_carry ? p.carry + (size_t)q * p.carry_cap : wl + slot * (size_t)p.cap;
      int out = 0;
      for (int base = 0; base < c; base += 32) {
        const int i = base + lane;
        const u64 key = (i < c) ? lp[i] : 0ull;
        const bool keep = (i < c) && key >= kth;
        const u32 m = __ballot_sync(kFullMask, keep);
        __syncwarp();
        if (keep) lp[out + __popc(m & ((1u << lane) - 1u))] = key;
        out += __popc(m);
      }
      if (lane == 0) wc[slot] = out;
    }
    return;
  }

  const int want = total < p.k ? total : p.k;
  if (p.mode == kSelPartial) {
    const size_t oslot = (size_t)blockIdx.y * p.Bpad + q;
    u64* ol = p.out_lists + oslot * (size_t)p.k;
    for_each_key(p, q, s0, s1, staged, stage, total, xform, qi, has_carry, [&](u64 key, bool active) {
      if (active && key >= kth) {
        const int pos = atomicAdd(&s_nout, 1);
        if (pos < p.k) ol[pos] = key;
      }
    });
    if (tid == 0) p.out_counts[oslot] = want;
    return;
  }

  for (int i = tid; i < p.kp2; i += SEL_THREADS) sk[i] = 0;
  __syncthreads();
  for_each_key(p, q, s0, s1, staged, stage, total, xform, qi, has_carry, [&](u64 key, bool active) {
    if (active && key >= kth) {
      const int pos = atomicAdd(&s_nout, 1);
      if (pos < p.kp2) sk[pos] = key;
    }
  });
  __syncthreads();
  block_bitonic_desc(sk, p.kp2);
  const size_t n = (size_t)p.B * p.k;
  for (int i = tid; i < p.k; i += SEL_THREADS) {
    const u64 key = sk[i];
    const bool ok = (i < want) && key != 0;
    float sc = ok ? key_score(key) : -CUDART_INF_F;          // keys already carry the final score
    long long id = ok ? (p.id_base + (long long)key_id(key)) : -1ll;
    if (!(sc >= p.min_score)) { sc = -CUDART_INF_F; id = -1ll; }
    if (!p.use_comm) {
      p.out_scores[(size_t)q * p.k + i] = sc;
      p.out_ids[(size_t)q * p.k + i] = id;
    } else {
      for (int g = 0; g < p.comm.world; ++g) {
        char* base = comm_slot(p.comm, g, epoch) + (size_t)p.comm.rank * p.rec_stride;
        reinterpret_cast<long long*>(base)[(size_t)q * p.k + i] = id;
        reinterpret_cast<float*>(base + 8 * n)[(size_t)q * p.k + i] = sc;
      }
    }
  }
  if (p.use_comm) comm_send_end(p.comm, epoch, gridDim.x * gridDim.y);
}

// ------------------------------------------------------------------------------------------------
// threshold_warp_kernel: the per-phase select (threshold mode) with ONE WARP per query -- no block barriers.
// Used when a carry list exists and the per-query candidate count (~k + one phase's survivors) fits the warp's stage;
// a query that does not fit is handled from global memory with the radix select.
//
// The threshold only has to be a VALID lower bound of the final k-th score that is as tight as the candidates allow, so
// the staged path does not run an exact radix select: the candidates are gathered segment by segment (coalesced, four
// segments in flight), binned ONCE into 1024 bins that are linear in the score between the smallest and the largest
// candidate, the bin b* that holds the k-th best is found with one suffix scan, and the threshold is the smallest
// candidate score whose bin is >= b* -- an actual candidate's score, so exactly suffix(b*) >= k candidates reach it.
// About a third of the instructions of three 8-bit radix passes with warp-aggregated atomics.
// ------------------------------------------------------------------------------------------------
constexpr int TW_WARPS = 1;            // queries per CTA (the kernel is latency-bound: small CTAs pack the SMs to the smem limit)
constexpr int TW_STAGE = 1536;         // keys staged per warp (12 KB)
constexpr int TW_BINS = 1024;          // score bins, two 16-bit counters per word (a bin holds <= TW_STAGE keys)

__device__ __forceinline__ int tw_find_segment(const int* seg_off, int nseg, int i) {
  int lo = 0, hi = nseg;                       // last segment with seg_off <= i
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (seg_off[mid] <= i) lo = mid; else hi = mid; }
  return lo;
}

__device__ __forceinline__ int tw_bin(u64 key, float flo, float scale) {
  const int b = (int)((key_score(key) - flo) * scale);      // monotone in the score; flo is the smallest candidate score
  return b < TW_BINS - 1 ? b : TW_BINS - 1;
}

// the radix path for a query whose candidates do not fit the stage: keys are re-read from global memory in every pass
struct TwArgs { const u64* lists; int* counts; u64* carry; u32* gthr; int S, Bpad, cap, k, carry_cap; };
__device__ __noinline__ void tw_select_unstaged(const TwArgs p, int q, const int* seg_off, u32* hist, int total) {
  const int lane = threadIdx.x & 31;
  const int nreal = p.S, nseg = p.S + 1;
  const u64* carry = p.carry + (size_t)q * p.carry_cap;
  auto load_key = [&](int i) -> u64 {
    const int sg = tw_find_segment(seg_off, nseg, i);
    const int j = i - seg_off[sg];
    return sg >= nreal ? carry[j] : p.lists[((size_t)sg * p.Bpad + q) * (size_t)p.cap + j];
  };
  // radix select of the k-th key, top 24 bits (a truncated key is a valid, slightly low threshold)
  u64 prefix = 0;
  int remaining = p.k;
#pragma unroll 1
  for (int pass = 7; pass >= 5; --pass) {
    const int shift = pass * 8;
    for (int i = lane; i < 256; i += 32) hist[i] = 0;
    __syncwarp();
    const u64 hi_mask = (pass == 7) ? 0ull : (~0ull << (shift + 8));
    for (int base = 0; base < total; base += 32) {
      const int i = base + lane;
      const bool active = i < total;
      const u64 key = active ? load_key(i) : 0ull;
      const bool m = active && (key & hi_mask) == prefix;
      hist_add(hist, m, (u32)(key >> shift) & 255u);
    }
    __syncwarp();
    u32 c[8]; u32 lane_sum = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b) { c[b] = hist[lane * 8 + b]; lane_sum += c[b]; }
    u32 suffix = lane_sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_down_sync(kFullMask, suffix, o); if (lane + o < 32) suffix += t; }
    const u32 above = suffix - lane_sum;
    const bool mine = (above < (u32)remaining) && ((u32)remaining <= suffix);
    int digit = 0, new_remaining = 0;
    if (mine) {
      u32 r2 = above;
#pragma unroll
      for (int b = 7; b >= 0; --b) {
        if (r2 + c[b] >= (u32)remaining) { digit = lane * 8 + b; new_remaining = remaining - (int)r2; break; }
        r2 += c[b];
      }
    }
    const int src = __ffs(__ballot_sync(kFullMask, mine)) - 1;
    digit = __shfl_sync(kFullMask, digit, src);
    remaining = __shfl_sync(kFullMask, new_remaining, src);
    prefix |= (u64)(u32)digit << shift;
    __syncwarp();
  }
  const u64 kth = prefix;
  if (lane == 0) atomicMax(p.gthr + q, (u32)(kth >> 32));
  // prune in place, segment by segment (the carry may not hold everything)
  int* wc = p.counts;
  u64* wl = const_cast<u64*>(p.lists);
  for (int sg = 0; sg < nseg; ++sg) {
    const size_t slot = (size_t)sg * p.Bpad + q;
    const int c = seg_off[sg + 1] - seg_off[sg];
    if (c == 0) continue;
    u64* lp = sg >= nreal ? p.carry + (size_t)q * p.carry_cap : wl + slot * (size_t)p.cap;
    int out = 0;
    for (int base = 0; base < c; base += 32) {
      const int i = base + lane;
      const u64 key = (i < c) ? lp[i] : 0ull;
      const bool keep = (i < c) && key >= kth;
      const u32 m = __ballot_sync(kFullMask, keep);
      __syncwarp();
      if (keep) lp[out + __popc(m & ((1u << lane) - 1u))] = key;
      out += __popc(m);
    }
    if (lane == 0) wc[slot] = out;
  }
}

__global__ void __launch_bounds__(TW_WARPS * 32) threshold_warp_kernel(const SelectParams p, const int seg_stride) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * TW_WARPS + warp;
  if (q >= p.B) return;                                               // whole warp
  u64* stage = reinterpret_cast<u64*>(smem_raw) + (size_t)warp * TW_STAGE;
  u32* hist = reinterpret_cast<u32*>(reinterpret_cast<u64*>(smem_raw) + (size_t)TW_WARPS * TW_STAGE) + warp * (TW_BINS / 2);
  int* seg_off = reinterpret_cast<int*>(reinterpret_cast<u32*>(reinterpret_cast<u64*>(smem_raw) + (size_t)TW_WARPS * TW_STAGE) + TW_WARPS * (TW_BINS / 2)) +
                 (size_t)warp * seg_stride;
  const int nreal = p.S, nseg = p.S + 1;                               // + the carry list (count at index S)
  // ---- counts -> exclusive offsets ----
  int run = 0;
  for (int base = 0; base < nseg; base += 32) {
    const int sgi = base + lane;
    const int c = sgi < nseg ? p.counts[(size_t)sgi * p.Bpad + q] : 0;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(kFullMask, incl, o); if (lane >= o) incl += t; }
    if (sgi < nseg) seg_off[sgi] = run + incl - c;
    run += __shfl_sync(kFullMask, incl, 31);
  }
  if (lane == 0) seg_off[nseg] = run;
  __syncwarp();
  const int total = run;
  if (total < p.k) return;                                             // not enough candidates yet for a threshold
  if (total > TW_STAGE) {
    const TwArgs a = {p.lists, const_cast<int*>(p.counts), p.carry, p.gthr, p.S, p.Bpad, p.cap, p.k, p.carry_cap};
    tw_select_unstaged(a, q, seg_off, hist, total);
    return;
  }

  // ---- gather: segment by segment, four segments' loads in flight; track the score range ----
  u64* carry = p.carry + (size_t)q * p.carry_cap;
  u32 lo32 = 0xffffffffu, hi32 = 0u;
  for (int sg0 = 0; sg0 < nseg; sg0 += 4) {
    u64 v[4]; int o[4], c[4]; const u64* lp[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int sg = sg0 + r;
      o[r] = sg < nseg ? seg_off[sg] : 0;
      c[r] = sg < nseg ? seg_off[sg + 1] - o[r] : 0;
      lp[r] = sg >= nreal ? carry : p.lists + ((size_t)sg * p.Bpad + q) * (size_t)p.cap;
      v[r] = lane < c[r] ? lp[r][lane] : 0ull;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (lane < c[r]) { stage[o[r] + lane] = v[r]; lo32 = min(lo32, (u32)(v[r] >> 32)); hi32 = max(hi32, (u32)(v[r] >> 32)); }
      for (int i = lane + 32; i < c[r]; i += 32) {
        const u64 key = lp[r][i];
        stage[o[r] + i] = key;
        lo32 = min(lo32, (u32)(key >> 32)); hi32 = max(hi32, (u32)(key >> 32));
      }
    }
  }
  lo32 = __reduce_min_sync(kFullMask, lo32);
  hi32 = __reduce_max_sync(kFullMask, hi32);
  for (int i = lane; i < TW_BINS / 2; i += 32) hist[i] = 0;
  __syncwarp();

  // ---- one histogram over bins linear in the score, then the bin of the k-th best ----
  const float flo = ordered_to_f32(lo32), fhi = ordered_to_f32(hi32);
  const float range = fhi - flo;
  const float scale = (range > 0.0f && range < CUDART_INF_F) ? (float)(TW_BINS - 1) / range : 0.0f;    // 0: every key in bin 0
  for (int i = lane; i < total; i += 32) {
    const int b = tw_bin(stage[i], flo, scale);
    atomicAdd(&hist[b >> 1], 1u << ((b & 1) * 16));
  }
  __syncwarp();
  // lane L owns bins [32 L, 32 L + 32) = 16 words
  u32 w[16]; u32 lane_sum = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 t = reinterpret_cast<const uint4*>(hist + lane * 16)[j];
    w[4 * j] = t.x; w[4 * j + 1] = t.y; w[4 * j + 2] = t.z; w[4 * j + 3] = t.w;
    lane_sum += (t.x & 0xffffu) + (t.x >> 16) + (t.y & 0xffffu) + (t.y >> 16) + (t.z & 0xffffu) + (t.z >> 16) + (t.w & 0xffffu) + (t.w >> 16);
  }
  u32 suffix = lane_sum;                                             // keys in this lane's bins and all higher ones
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_down_sync(kFullMask, suffix, o); if (lane + o < 32) suffix += t; }
  const u32 above = suffix - lane_sum;
  const bool mine = above < (u32)p.k && (u32)p.k <= suffix;          // exactly one lane (total >= k)
  int bstar = 0;
  if (mine) {
    u32 acc = above;
#pragma unroll
    for (int j = 15; j >= 0; --j) {
      acc += w[j] >> 16;
      if (acc >= (u32)p.k) { bstar = lane * 32 + 2 * j + 1; break; }
      acc += w[j] & 0xffffu;
      if (acc >= (u32)p.k) { bstar = lane * 32 + 2 * j; break; }
    }
  }
  bstar = __shfl_sync(kFullMask, bstar, __ffs(__ballot_sync(kFullMask, mine)) - 1);
  // ---- threshold = the smallest candidate score in bins >= b*: at least k candidates reach it ----
  u32 t32 = 0xffffffffu;
  for (int i = lane; i < total; i += 32) {
    const u64 key = stage[i];
    if (tw_bin(key, flo, scale) >= bstar) t32 = min(t32, (u32)(key >> 32));
  }
  t32 = __reduce_min_sync(kFullMask, t32);
  const u64 kth = (u64)t32 << 32;
  if (lane == 0) atomicMax(p.gthr + q, t32);
  // ---- prune: survivors -> the carry list (every read of the old carry is in the stage by now); unit lists restart empty ----
  int* wc = const_cast<int*>(p.counts);
  int out = 0;
  for (int base = 0; base < total; base += 32) {
    const int i = base + lane;
    const u64 key = i < total ? stage[i] : 0ull;
    const bool keep = i < total && key >= kth;
    const u32 m = __ballot_sync(kFullMask, keep);
    if (keep) carry[out + __popc(m & ((1u << lane) - 1u))] = key;
    out += __popc(m);
  }
  for (int sgi = lane; sgi < nreal; sgi += 32) wc[(size_t)sgi * p.Bpad + q] = 0;
  if (lane == 0) wc[(size_t)p.S * p.Bpad + q] = out;
}

static int run_threshold_warp(const SelectParams& p, cudaStream_t st) {
  ProfScope _ps(st, "select_threshold_warp");
  const int seg_stride = (p.S + 2 + 3) & ~3;
  const size_t smem = (size_t)TW_WARPS * TW_STAGE * 8 + (size_t)TW_WARPS * (TW_BINS / 2) * 4 + (size_t)TW_WARPS * seg_stride * 4;
  B200R_CUDA_CHECK(cudaFuncSetAttribute(threshold_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  threshold_warp_kernel<<<(p.B + TW_WARPS - 1) / TW_WARPS, TW_WARPS * 32, smem, st>>>(p, seg_stride);
  count_launch();
  B200R_CUDA_CHECK(cudaGetLastError());
  return B200R_OK;
}

static int next_pow2(int x) { int p = 1; while (p < x) p <<= 1; return p; }

static int run_select(SelectParams& p, int grid_y, cudaStream_t st) {
  ProfScope _ps(st, p.mode == kSelThreshold ? "select_threshold" : (p.mode == kSelPartial ? "select_partial" : (p.use_comm ? "select_final_scatter" : "select_final")));
  const long long all = (long long)(p.mode == kSelPartial ? p.seg_per_cta : p.S) * p.cap + (p.carry ? p.carry_cap : 0);
  int limit = SEL_STAGE_MAX;
  const long long ctas = (long long)p.B * grid_y;
  if (p.mode != kSelPartial) limit = select_stage_limit(p.k, p.B);   // pruned lists hold ~k + one phase's survivors
  else if (ctas <= 148) limit = 24576;      // a handful of queries: one CTA per SM can afford a 192 KB stage
  p.stage_cap = (int)(all < limit ? all : limit);
  int threads = ctas >= 2048 ? 128 : (ctas >= 512 ? 256 : (ctas >= 128 ? 512 : 1024));
  // few keys per CTA (e.g. the second level of a two-level select: groups x k keys): the ~70 block barriers of the radix
  // passes and the sort dominate, and they are several times cheaper with 4 warps than with 32
  if (all <= 1024) threads = 128;
  else if (all <= 4096 && threads > 256) threads = 256;
  const size_t smem = (size_t)p.stage_cap * 8 + (p.mode == kSelOutput ? (size_t)p.kp2 * 8 : 0);
  B200R_CUDA_CHECK(cudaFuncSetAttribute(select_lists_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  select_lists_kernel<<<dim3(p.B, grid_y), threads, smem, st>>>(p);
  count_launch();
  B200R_CUDA_CHECK(cudaGetLastError());
  return B200R_OK;
}

int launch_select_lists(const Index* ix, const SearchPlan& pl, void* ws, float* out_scores, int64_t* out_ids, cudaStream_t st,
                        const CommDev* comm, size_t rec_stride, float min_score) {
  SelectParams p;
  memset(&p, 0, sizeof(p));
  p.min_score = min_score;
  if (comm) { p.comm = *comm; p.use_comm = 1; p.rec_stride = rec_stride; }
  p.lists = reinterpret_cast<const u64*>((char*)ws + pl.off_lists);
  p.counts = reinterpret_cast<const int*>((char*)ws + pl.off_counts);
  p.S = pl.num_segments; p.Bpad = pl.Bpad; p.cap = pl.cap;
  p.B = pl.B; p.k = pl.k; p.kp2 = next_pow2(pl.k);
  p.qinv = (ix->metric == B200R_COS) ? reinterpret_cast<const float*>((char*)ws + pl.off_qinv) : nullptr;
  p.gthr = reinterpret_cast<u32*>((char*)ws + pl.off_gthr);
  if (pl.carry_cap > 0) { p.carry = reinterpret_cast<u64*>((char*)ws + pl.off_carry); p.carry_cap = pl.carry_cap; }
  p.id_base = ix->id_base;
  p.out_scores = out_scores;
  p.out_ids = reinterpret_cast<long long*>(out_ids);
  if (!out_scores && !comm) {
    p.mode = kSelThreshold;
    // one warp per query when there are enough queries to fill the GPU that way and a phase's survivors fit its stage
    static const int env_warp = [] { const char* e = getenv("B200R_SELECT_WARP"); return e ? atoi(e) : 1; }();
    if (env_warp && p.carry && p.B >= 512 && 9 * p.k + 256 <= TW_STAGE && p.S + 1 <= SEL_MAX_SEGS) return run_threshold_warp(p, st);
    return run_select(p, 1, st);
  }
  if (pl.num_groups > 0) {
    // two levels: many short per-warp lists (scan path) are first reduced group-wise to k keys each
    SelectParams p1 = p;
    p1.use_comm = 0;           // the partial level's group 0 also reduces the carry list
    p.carry = nullptr;
    p1.mode = kSelPartial; p1.seg_per_cta = SEL_GROUP;
    p1.out_lists = reinterpret_cast<u64*>((char*)ws + pl.off_lists2);
    p1.out_counts = reinterpret_cast<int*>((char*)ws + pl.off_counts2);
    int rc = run_select(p1, pl.num_groups, st);
    if (rc) return rc;
    p.lists = p1.out_lists; p.counts = p1.out_counts; p.S = pl.num_groups; p.cap = pl.k;
    p.final_keys = 1;
  }
  p.mode = kSelOutput;
  return run_select(p, 1, st);
}

// ---- range search bookkeeping (b200r_range_count / b200r_range_fill) ---------------------------------------------------
// counts[c][q] (hits of query q in row chunk c) -> offsets[q] (exclusive over queries, offsets[B] = total) and
// slot_off[c][q] = offsets[q] + hits of q in the chunks before c: per query the hits land in ascending row order.  One CTA.
__global__ void __launch_bounds__(1024) range_offsets_kernel(const int* __restrict__ counts, int num_chunks, int Bpad, int B,
                                                             long long* __restrict__ offsets, long long* __restrict__ slot_off) {
  __shared__ long long s_warp[32];
  __shared__ long long s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int q0 = 0; q0 < B; q0 += 1024) {
    const int q = q0 + tid;
    long long tot = 0;
    if (q < B)
      for (int c = 0; c < num_chunks; ++c) { slot_off[(size_t)c * Bpad + q] = tot; tot += counts[(size_t)c * Bpad + q]; }
    long long incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const long long t = __shfl_up_sync(kFullMask, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    long long wbase = 0, all = 0;
    for (int w = 0; w < 32; ++w) { if (w < warp) wbase += s_warp[w]; all += s_warp[w]; }
    const long long mine = s_base + wbase + incl - tot;
    if (q < B) {
      offsets[q] = mine;
      for (int c = 0; c < num_chunks; ++c) slot_off[(size_t)c * Bpad + q] += mine;
    }
    __syncthreads();
    if (tid == 0) s_base += all;
    __syncthreads();
  }
  if (tid == 0) offsets[B] = s_base;
}

int launch_range_offsets(const SearchPlan& pl, void* ws, int64_t* offsets, int64_t* slot_off, cudaStream_t st) {
  ProfScope _ps(st, "range_offsets");
  range_offsets_kernel<<<1, 1024, 0, st>>>(reinterpret_cast<const int*>((char*)ws + pl.off_counts), pl.num_chunks, pl.Bpad, pl.B,
                                           reinterpret_cast<long long*>(offsets), reinterpret_cast<long long*>(slot_off));
  count_launch();
  B200R_CUDA_CHECK(cudaGetLastError());
  return B200R_OK;
}

// keys (left in the id array by the fill pass) -> global ids + scores, in place
__global__ void range_unpack_kernel(long long* __restrict__ ids, float* __restrict__ scores, long long total, long long id_base) {
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, step = (long long)gridDim.x * blockDim.x;
  for (long long i = i0; i < total; i += step) {
    const u64 key = (u64)ids[i];
    scores[i] = key_score(key);
    ids[i] = id_base + (long long)key_id(key);
  }
}

int launch_range_unpack(const void*, int64_t total, const int64_t*, int, float, int64_t id_base, int64_t* out_ids, float* out_scores,
                        cudaStream_t st) {
  if (total <= 0) return B200R_OK;
  ProfScope _ps(st, "range_unpack");
  long long blocks = (total + 255) / 256;
  if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
  range_unpack_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<long long*>(out_ids), out_scores, total, id_base);
  count_launch();
  B200R_CUDA_CHECK(cudaGetLastError());
  return B200R_OK;
}

// exact top-k of ONE flat list of keys in global memory (the unbounded-frontier path of the graph stage): *count_dev keys
int launch_select_flat_keys(const void* keys, const int* count_dev, long long cap, int k, float* out_scores, int64_t* out_ids, cudaStream_t st) {
  SelectParams p;
  memset(&p, 0, sizeof(p));
  p.min_score = -__builtin_huge_valf();
  p.lists = reinterpret_cast<const u64*>(keys); p.counts = count_dev;
  p.S = 1; p.Bpad = 1; p.cap = (int)(cap > 0x7fffffffll ? 0x7fffffffll : cap);
  p.B = 1; p.k = k; p.kp2 = next_pow2(k);
  p.out_scores = out_scores; p.out_ids = reinterpret_cast<long long*>(out_ids);
  p.mode = kSelOutput;
  return run_select(p, 1, st);
}

int select_group_size() { return SEL_GROUP; }
int select_max_flat_segments() { return SEL_MAX_SEGS; }

// ------------------------------------------------------------------------------------------------
// public merge: candidates [G, B, kin] (f32 score, i64 id) -> [B, k]
// ------------------------------------------------------------------------------------------------
constexpr int MERGE_THREADS = 256;

struct MergeParams {
  const float* cs;
  const long long* ci;
  size_t stride_s, stride_i;   // elements between consecutive shards
  int G, B, kin, k, np2;
  float* os;
  long long* oi;
  int* osrc;                   // nullable: [B, k] position g*kin + j of every winner in the gathered lists (-1 = padding)
};

__device__ __forceinline__ bool cand_better(u32 sa, long long ia, u32 sb, long long ib) {
  return sa > sb || (sa == sb && ia < ib);
}

__global__ void __launch_bounds__(MERGE_THREADS) merge_topk_kernel(const MergeParams p) {
  extern __shared__ uint8_t smem_raw[];
  long long* ids = reinterpret_cast<long long*>(smem_raw);  // [np2]
  u32* sc = reinterpret_cast<u32*>(ids + p.np2);            // [np2] ordered scores
  int* src = reinterpret_cast<int*>(sc + p.np2);            // [np2] where each candidate came from
  const int q = blockIdx.x;
  const int n = p.G * p.kin;
  for (int i = threadIdx.x; i < p.np2; i += MERGE_THREADS) {
    u32 s = 0; long long id = 0x7fffffffffffffffll;
    if (i < n) {
      const int g = i / p.kin, j = i % p.kin;
      const size_t off = (size_t)q * p.kin + j;
      const long long cid = p.ci[(size_t)g * p.stride_i + off];
      if (cid >= 0) { id = cid; s = f32_to_ordered(p.cs[(size_t)g * p.stride_s + off] + 0.0f); }
    }
    sc[i] = s; ids[i] = id; src[i] = i;
  }
  for (int size = 2; size <= p.np2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < p.np2 / 2; t += MERGE_THREADS) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const u32 sa = sc[lo], sb = sc[hi];
        const long long ia = ids[lo], ib = ids[hi];
        const bool lo_better = cand_better(sa, ia, sb, ib);
        if (lo_better != desc) {
          sc[lo] = sb; sc[hi] = sa; ids[lo] = ib; ids[hi] = ia;
          const int t2 = src[lo]; src[lo] = src[hi]; src[hi] = t2;
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < p.k; i += MERGE_THREADS) {
    const bool ok = i < p.np2 && ids[i] != 0x7fffffffffffffffll;
    p.os[(size_t)q * p.k + i] = ok ? ordered_to_f32(sc[i]) : -CUDART_INF_F;
    p.oi[(size_t)q * p.k + i] = ok ? ids[i] : -1ll;
    if (p.osrc) p.osrc[(size_t)q * p.k + i] = ok ? src[i] : -1;
  }
}

int launch_merge_topk(const float* cs, const int64_t* ci, size_t shard_stride_s, size_t shard_stride_i, int G, int B, int kin, int k,
                      float* os, int64_t* oi, cudaStream_t st, int32_t* osrc) {
  ProfScope _ps(st, "merge_topk");
  const long long n = (long long)G * kin;
  if (n > 16384) { set_error("merge_topk: G*kin = %lld exceeds 16384", n); return B200R_EINVAL; }
  MergeParams p;
  p.cs = cs; p.ci = reinterpret_cast<const long long*>(ci); p.stride_s = shard_stride_s; p.stride_i = shard_stride_i; p.G = G; p.B = B; p.kin = kin; p.k = k;
  p.np2 = next_pow2((int)n < 2 ? 2 : (int)n);
  p.os = os; p.oi = reinterpret_cast<long long*>(oi); p.osrc = osrc;
  const size_t smem = (size_t)p.np2 * 16;
  B200R_CUDA_CHECK(cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  merge_topk_kernel<<<B, MERGE_THREADS, smem, st>>>(p);
  count_launch();
  B200R_CUDA_CHECK(cudaGetLastError());
  return B200R_OK;
}

}  // namespace b200r
