// comm.cuh -- device side of the shard exchange (b200r_comm_*): every rank owns one WINDOW of device memory that all
// ranks of the box map (CUDA IPC between processes, peer access inside one process), and kernels exchange data by
// storing straight into the peers' windows over NVLink, ordered by epoch flags.  No NCCL, no host round trip.
//
// Window layout: [CommCtrl | channel 0 data | channel 1 data | ...]; a channel's data region is 2 slots (even / odd
// epoch) of `world` records each: rank r's record of epoch E lives at  data_off + (E & 1) * slot_bytes + r * rec_stride
// in EVERY rank's window (all-gather by peer stores).
//
// Protocol of one exchange on a channel (epochs count from 1, kept on the device so a captured CUDA graph replays):
//   sender kernel:   E = send_epoch + 1; wait until every peer has CONSUMED epoch E - 2 (its reader of this slot is done);
//                    store the payload into every peer's slot E & 1; the last CTA to finish publishes arrive[slot][rank] = E
//                    in every peer's window (release, system scope) and sets send_epoch = E.
//   receiver kernel: E = recv_epoch + 1; wait until arrive[slot][g] >= E for every rank g (acquire, system scope); read the
//                    local window; the last CTA to finish publishes consumed[rank] = E to every peer and sets recv_epoch = E.
// Waits are bounded (kCommTimeoutNs): on expiry the kernel sets the sticky status word and carries on, so a rank that
// never shows up turns into an error code (b200r_comm_status) instead of a hung GPU.
#pragma once
#include "common.cuh"
#include "../../include/b200r.h"

namespace b200r {

constexpr int kCommChannels = 4;
constexpr int kCommCtrlBytes = 8192;
constexpr unsigned long long kCommTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;

struct CommCtrl {
  // written only by kernels of the owning rank
  u32 send_epoch[kCommChannels];
  u32 recv_epoch[kCommChannels];
  u32 send_done[kCommChannels];
  u32 recv_done[kCommChannels];
  u32 status;
  u32 pad0[64 - 4 * kCommChannels - 1];
  // written by the peers (one 32-bit word per writer: no read-modify-write across NVLink)
  u32 arrive[kCommChannels][2][B200R_MAX_PEERS];
  u32 consumed[kCommChannels][B200R_MAX_PEERS];
};
static_assert(sizeof(CommCtrl) <= kCommCtrlBytes, "control block too large");

struct CommDev {                              // passed by value to kernels
  unsigned long long win[B200R_MAX_PEERS];    // base address of every rank's window, mapped in this process
  int world, rank, channel;
  unsigned long long data_off, slot_bytes;    // this channel's data region
};

#ifdef __CUDACC__
__device__ __forceinline__ CommCtrl* comm_ctrl(const CommDev& c, int g) { return reinterpret_cast<CommCtrl*>(c.win[g]); }
__device__ __forceinline__ char* comm_slot(const CommDev& c, int g, u32 epoch) {
  return reinterpret_cast<char*>(c.win[g]) + c.data_off + (size_t)(epoch & 1u) * c.slot_bytes;
}
__device__ __forceinline__ u32 ld_acquire_sys(const u32* p) {
  u32 v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(u32* p, u32 v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// spin until *flag has reached `want` (wrap-safe); false on timeout
__device__ __forceinline__ bool comm_wait_flag(const u32* flag, u32 want) {
  if ((int)(ld_acquire_sys(flag) - want) >= 0) return true;
  const unsigned long long t0 = global_timer_ns();
  while ((int)(ld_acquire_sys(flag) - want) < 0) {
    __nanosleep(64);
    if (global_timer_ns() - t0 > kCommTimeoutNs) return false;
  }
  return true;
}

// All threads of the CTA call these (they contain __syncthreads).  `nctas` = CTAs of the grid that take part.
__device__ __forceinline__ u32 comm_send_begin(const CommDev& c) {
  CommCtrl* me = comm_ctrl(c, c.rank);
  const u32 E = *reinterpret_cast<volatile u32*>(&me->send_epoch[c.channel]) + 1u;
  if ((int)threadIdx.x < c.world && E > 2u) {
    if (!comm_wait_flag(&me->consumed[c.channel][threadIdx.x], E - 2u)) atomicExch(&me->status, 1u);
  }
  __syncthreads();
  return E;
}
__device__ __forceinline__ void comm_send_end(const CommDev& c, u32 E, u32 nctas) {
  __threadfence_system();           // this thread's payload stores are performed system-wide before the CTA checks in
  __syncthreads();
  __shared__ u32 s_last;
  CommCtrl* me = comm_ctrl(c, c.rank);
  if (threadIdx.x == 0) s_last = (atomicAdd(&me->send_done[c.channel], 1u) == nctas - 1u) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();           // order the other CTAs' (fenced) stores before the flags below
  if ((int)threadIdx.x < c.world) st_release_sys(&comm_ctrl(c, threadIdx.x)->arrive[c.channel][E & 1u][c.rank], E);
  if (threadIdx.x == 0) { me->send_done[c.channel] = 0u; me->send_epoch[c.channel] = E; }
}
__device__ __forceinline__ u32 comm_recv_begin(const CommDev& c) {
  CommCtrl* me = comm_ctrl(c, c.rank);
  const u32 E = *reinterpret_cast<volatile u32*>(&me->recv_epoch[c.channel]) + 1u;
  if ((int)threadIdx.x < c.world) {
    if (!comm_wait_flag(&me->arrive[c.channel][E & 1u][threadIdx.x], E)) atomicExch(&me->status, 1u);
  }
  __syncthreads();
  return E;
}
// epoch of the exchange a receiver is about to read, without waiting (a kernel that runs after one that already waited)
__device__ __forceinline__ u32 comm_recv_peek(const CommDev& c) {
  return *reinterpret_cast<volatile u32*>(&comm_ctrl(c, c.rank)->recv_epoch[c.channel]) + 1u;
}
__device__ __forceinline__ void comm_recv_end(const CommDev& c, u32 E, u32 nctas) {
  __syncthreads();                  // every thread's reads of the window have returned
  __shared__ u32 r_last;
  CommCtrl* me = comm_ctrl(c, c.rank);
  if (threadIdx.x == 0) { __threadfence(); r_last = (atomicAdd(&me->recv_done[c.channel], 1u) == nctas - 1u) ? 1u : 0u; }
  __syncthreads();
  if (!r_last) return;
  if ((int)threadIdx.x < c.world) st_release_sys(&comm_ctrl(c, threadIdx.x)->consumed[c.channel][c.rank], E);
  if (threadIdx.x == 0) { me->recv_done[c.channel] = 0u; me->recv_epoch[c.channel] = E; }
}
#endif

}  // namespace b200r
