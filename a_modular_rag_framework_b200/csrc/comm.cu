// comm.cu -- host side of the shard exchange (b200r_comm_*) and the receiving merge kernel.
//
// SURVEY.md section 8(b) asks for the exchange under the C ABI "so Python never links NCCL".  The exchange here does
// not use NCCL at all: each rank allocates one window, the ranks swap 128-byte descriptors through ANY host channel
// (a file, a socket, torch.distributed's object gather -- control plane only), map each other's windows (CUDA IPC
// between processes, peer access inside one process) and from then on the search kernels themselves move the data:
// the final select kernel stores every rank's packed top-k into all windows over NVLink and the merge kernel below waits
// for the arrival flags (comm.cuh).  One process per GPU and one process driving several GPUs use the same code.
#include "comm.cuh"
#include "internal.h"

#include <math_constants.h>
#include <cstring>
#include <unistd.h>

namespace b200r {

struct CommBlob {                 // what b200r_comm_export writes (B200R_COMM_BLOB_BYTES = 128)
  uint32_t magic;
  int32_t pid, device, rank;
  unsigned long long ptr, bytes;
  int32_t has_ipc, pad;
  cudaIpcMemHandle_t handle;      // 64 bytes
};
static_assert(sizeof(CommBlob) <= B200R_COMM_BLOB_BYTES, "blob too large");

Comm* get_comm(uint64_t h) {
  Comm* c = reinterpret_cast<Comm*>(h);
  if (!c || c->magic != kCommMagic) { set_error("bad comm handle"); return nullptr; }
  return c;
}

int comm_dev(const Comm* c, int channel, size_t rec_stride, CommDev* out) {
  if (!c->connected) { set_error("comm is not connected (call b200r_comm_connect / b200r_comm_init_all first)"); return B200R_ESTATE; }
  if (channel < 0 || channel >= kCommChannels || c->rec_cap[channel] == 0) { set_error("exchange channel %d was not provisioned", channel); return B200R_EINVAL; }
  if (rec_stride > c->rec_cap[channel]) {
    set_error("exchange record of %zu bytes exceeds the channel's capacity %zu (comm created for max_B=%d max_k=%d max_graph_k=%d)",
              rec_stride, c->rec_cap[channel], c->max_B, c->max_k, c->max_graph_k);
    return B200R_EINVAL;
  }
  memset(out, 0, sizeof(*out));
  for (int g = 0; g < c->world; ++g) out->win[g] = c->peer[g];
  out->world = c->world; out->rank = c->rank; out->channel = channel;
  out->data_off = c->data_off[channel];
  out->slot_bytes = (unsigned long long)c->world * c->rec_cap[channel];
  return B200R_OK;
}

// ------------------------------------------------------------------------------------------------
// receiving side of the dense exchange: wait for every rank's record, merge the G sorted lists of each query.
// Every list is strictly sorted by (score desc, id asc) (select_lists_kernel ranks on the final score), ids are unique
// across shards, so the rank of an element in the union is its own position plus, for every other list, the number of
// elements that beat it -- a binary search per list, no sort and no barrier: out[rank] = element.
// ------------------------------------------------------------------------------------------------
constexpr int XM_THREADS = 128;

struct ExchMergeParams {
  CommDev comm;
  int B, kin, k;
  unsigned long long rec_stride;
  float* os;
  long long* oi;
};

__device__ __forceinline__ bool beats(u32 sa, long long ia, u32 sb, long long ib) { return sa > sb || (sa == sb && ia < ib); }

__global__ void __launch_bounds__(XM_THREADS) exchange_merge_kernel(const ExchMergeParams p) {
  extern __shared__ uint8_t smem_raw[];
  long long* ids = reinterpret_cast<long long*>(smem_raw);            // [G * kin]
  u32* sc = reinterpret_cast<u32*>(ids + (size_t)p.comm.world * p.kin);   // [G * kin] ordered scores
  __shared__ int s_real;
  const u32 E = comm_recv_begin(p.comm);
  const int G = p.comm.world, q = blockIdx.x;
  const char* slot = comm_slot(p.comm, p.comm.rank, E);
  const size_t n = (size_t)p.B * p.kin;
  if (threadIdx.x == 0) s_real = 0;
  __syncthreads();
  int real = 0;
  for (int i = threadIdx.x; i < G * p.kin; i += XM_THREADS) {
    const int g = i / p.kin, j = i - g * p.kin;
    const char* rec = slot + (size_t)g * p.rec_stride;
    const long long id = reinterpret_cast<const long long*>(rec)[(size_t)q * p.kin + j];
    const float s = reinterpret_cast<const float*>(rec + 8 * n)[(size_t)q * p.kin + j];
    const bool ok = id >= 0;
    ids[i] = ok ? id : 0x7fffffffffffffffll;          // padding sorts last and is never emitted
    sc[i] = ok ? f32_to_ordered(s + 0.0f) : 0u;
    real += ok ? 1 : 0;
  }
  for (int o = 16; o > 0; o >>= 1) real += __shfl_xor_sync(kFullMask, real, o);
  if ((threadIdx.x & 31) == 0 && real) atomicAdd(&s_real, real);
  __syncthreads();
  for (int i = threadIdx.x; i < G * p.kin; i += XM_THREADS) {
    const long long id = ids[i];
    if (id == 0x7fffffffffffffffll) continue;
    const u32 s = sc[i];
    const int g = i / p.kin;
    int rank = i - g * p.kin;
    for (int h = 0; h < G && rank < p.k; ++h) {
      if (h == g) continue;
      const int base = h * p.kin;
      int lo = 0, hi = p.kin;                             // first position of list h that does NOT beat (s, id)
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (beats(sc[base + mid], ids[base + mid], s, id)) lo = mid + 1; else hi = mid;
      }
      rank += lo;
    }
    if (rank < p.k) {
      p.os[(size_t)q * p.k + rank] = ordered_to_f32(s);
      p.oi[(size_t)q * p.k + rank] = id;
    }
  }
  for (int i = s_real + (int)threadIdx.x; i < p.k; i += XM_THREADS) {
    p.os[(size_t)q * p.k + i] = -CUDART_INF_F;
    p.oi[(size_t)q * p.k + i] = -1ll;
  }
  comm_recv_end(p.comm, E, gridDim.x);
}

int launch_exchange_merge(const Comm* c, int B, int kin, int k, float* out_scores, int64_t* out_ids, cudaStream_t st) {
  ProfScope _ps(st, "exchange_merge");
  ExchMergeParams p;
  int rc = comm_dev(c, kChanDense, dense_record_stride(B, kin), &p.comm);
  if (rc) return rc;
  p.B = B; p.kin = kin; p.k = k; p.rec_stride = dense_record_stride(B, kin);
  p.os = out_scores; p.oi = reinterpret_cast<long long*>(out_ids);
  const size_t smem = (size_t)c->world * kin * 12;
  if (smem > 200 * 1024) { set_error("exchange merge: world*k = %d exceeds the shared-memory stage", c->world * kin); return B200R_EINVAL; }
  B200R_CUDA_CHECK(cudaFuncSetAttribute(exchange_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  exchange_merge_kernel<<<B, XM_THREADS, smem, st>>>(p);
  count_launch();
  B200R_CUDA_CHECK(cudaGetLastError());
  return B200R_OK;
}

// consume an exchange without reading it (a rank that does not need the merged result still advances its epoch)
__global__ void exchange_skip_kernel(const CommDev comm) {
  const u32 E = comm_recv_begin(comm);
  comm_recv_end(comm, E, 1);
}

int launch_exchange_skip(const Comm* c, cudaStream_t st) {
  ProfScope _ps(st, "exchange_skip");
  CommDev cd;
  int rc = comm_dev(c, kChanDense, 0, &cd);
  if (rc) return rc;
  exchange_skip_kernel<<<1, 32, 0, st>>>(cd);
  count_launch();
  B200R_CUDA_CHECK(cudaGetLastError());
  return B200R_OK;
}

// all-gather of a finished local record by peer stores: record bytes (a multiple of 4) -> this rank's place in the
// current slot of every window
__global__ void __launch_bounds__(256) scatter_record_kernel(const CommDev comm, const u32* __restrict__ src, size_t words, size_t rec_stride) {
  const u32 E = comm_send_begin(comm);
  const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (rec_stride & 15) == 0) {
    const size_t vec = words / 4;
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    for (size_t i = i0; i < vec; i += step) {
      const uint4 v = s4[i];
      for (int g = 0; g < comm.world; ++g)
        reinterpret_cast<uint4*>(comm_slot(comm, g, E) + (size_t)comm.rank * rec_stride)[i] = v;
    }
    for (size_t i = vec * 4 + i0; i < words; i += step)
      for (int g = 0; g < comm.world; ++g)
        reinterpret_cast<u32*>(comm_slot(comm, g, E) + (size_t)comm.rank * rec_stride)[i] = src[i];
  } else {
    for (size_t i = i0; i < words; i += step)
      for (int g = 0; g < comm.world; ++g)
        reinterpret_cast<u32*>(comm_slot(comm, g, E) + (size_t)comm.rank * rec_stride)[i] = src[i];
  }
  comm_send_end(comm, E, gridDim.x);
}

int launch_scatter_record(const CommDev& cd, const void* src, size_t bytes, size_t rec_stride, cudaStream_t st) {
  ProfScope _ps(st, "scatter_record");
  if (bytes % 4 || bytes > rec_stride) { set_error("scatter_record: bad record size %zu (stride %zu)", bytes, rec_stride); return B200R_EINVAL; }
  const size_t words = bytes / 4;
  int blocks = (int)((words / 4 + 255) / 256);
  if (blocks < 1) blocks = 1;
  if (blocks > 2 * num_sms()) blocks = 2 * num_sms();
  scatter_record_kernel<<<blocks, 256, 0, st>>>(cd, reinterpret_cast<const u32*>(src), words, rec_stride);
  count_launch();
  B200R_CUDA_CHECK(cudaGetLastError());
  return B200R_OK;
}

static int comm_alloc(int device, int rank, int world, int max_B, int max_k, int max_graph_k, Comm** out) {
  if (world < 1 || world > B200R_MAX_PEERS || rank < 0 || rank >= world) { set_error("bad rank/world (%d/%d), world <= %d", rank, world, B200R_MAX_PEERS); return B200R_EINVAL; }
  if (max_B < 1 || max_k < 1 || max_k > B200R_MAX_K || max_graph_k < 0 || max_graph_k > B200R_MAX_K) { set_error("bad max_B/max_k/max_graph_k"); return B200R_EINVAL; }
  B200R_CUDA_CHECK(cudaSetDevice(device));
  Comm* c = new (std::nothrow) Comm();
  if (!c) { set_error("out of host memory"); return B200R_ENOMEM; }
  memset(c, 0, sizeof(*c));
  c->magic = kCommMagic; c->device = device; c->rank = rank; c->world = world;
  c->max_B = max_B; c->max_k = max_k; c->max_graph_k = max_graph_k;
  c->rec_cap[kChanDense] = dense_record_stride(max_B, max_k);
  if (max_graph_k > 0) {
    // graph stage (csr_expand.cu, mode 3): requests home -> owner (one record per sending rank), cosines owner -> home
    // and results home -> everybody (one shared region each, expressed as world equal parts)
    const int per = (max_B + world - 1) / world;
    c->rec_cap[kChanGraphReq] = expand_xchg_req_bytes(per);
    c->rec_cap[kChanGraphCos] = align_up((expand_xchg_cos_bytes(per) + world - 1) / world, 16);
    c->rec_cap[kChanGraphOut] = align_up((expand_xchg_out_bytes(max_B, max_graph_k) + world - 1) / world, 16);
  }
  size_t off = kCommCtrlBytes;
  for (int ch = 0; ch < kCommChannels; ++ch) {
    c->data_off[ch] = off;
    off += align_up(2 * (size_t)world * c->rec_cap[ch], 256);
  }
  c->window_bytes = off;
  cudaError_t e = cudaMalloc(&c->window, c->window_bytes);
  if (e != cudaSuccess) { set_error("cudaMalloc(exchange window, %zu bytes) failed: %s", c->window_bytes, cudaGetErrorString(e)); delete c; return B200R_ENOMEM; }
  e = cudaMemset(c->window, 0, c->window_bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { set_error("exchange window init failed: %s", cudaGetErrorString(e)); cudaFree(c->window); delete c; return B200R_ECUDA; }
  c->peer[rank] = reinterpret_cast<unsigned long long>(c->window);
  *out = c;
  return B200R_OK;
}

static void fill_blob(const Comm* c, CommBlob* b) {
  memset(b, 0, sizeof(*b));
  b->magic = kCommMagic; b->pid = (int32_t)getpid(); b->device = c->device; b->rank = c->rank;
  b->ptr = reinterpret_cast<unsigned long long>(c->window); b->bytes = c->window_bytes;
  cudaSetDevice(c->device);
  b->has_ipc = cudaIpcGetMemHandle(&b->handle, c->window) == cudaSuccess ? 1 : 0;
  if (!b->has_ipc) cudaGetLastError();
}

static int connect_blobs(Comm* c, const CommBlob* blobs) {
  B200R_CUDA_CHECK(cudaSetDevice(c->device));
  for (int g = 0; g < c->world; ++g) {
    const CommBlob& b = blobs[g];
    if (b.magic != kCommMagic || b.rank != g) { set_error("descriptor %d is not rank %d's b200r_comm_export output", g, g); return B200R_EINVAL; }
    if (b.bytes != c->window_bytes) { set_error("rank %d created its comm with different sizes (window %llu vs %zu bytes)", g, b.bytes, c->window_bytes); return B200R_EINVAL; }
    if (g == c->rank) continue;
    if (b.pid == (int32_t)getpid()) {
      // same process (one host driving several GPUs): plain peer access to the other device's allocation
      if (b.device != c->device) {
        int can = 0;
        B200R_CUDA_CHECK(cudaDeviceCanAccessPeer(&can, c->device, b.device));
        if (!can) { set_error("device %d cannot access device %d's memory (no NVLink/PCIe peer path)", c->device, b.device); return B200R_ECUDA; }
        cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { set_error("cudaDeviceEnablePeerAccess(%d) failed: %s", b.device, cudaGetErrorString(e)); return B200R_ECUDA; }
        cudaGetLastError();
      }
      c->peer[g] = b.ptr;
    } else {
      if (!b.has_ipc) { set_error("rank %d could not export a CUDA IPC handle for its window", g); return B200R_ECUDA; }
      void* p = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&p, b.handle, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) { set_error("cudaIpcOpenMemHandle(rank %d) failed: %s", g, cudaGetErrorString(e)); cudaGetLastError(); return B200R_ECUDA; }
      c->peer[g] = reinterpret_cast<unsigned long long>(p);
      c->ipc_opened[g] = true;
    }
  }
  c->connected = true;
  return B200R_OK;
}

}  // namespace b200r

using namespace b200r;

extern "C" {

int b200r_comm_create(int device, int32_t rank, int32_t world, int32_t max_B, int32_t max_k, int32_t max_graph_k, uint64_t* out_comm) {
  if (!out_comm) { set_error("out_comm is null"); return B200R_EINVAL; }
  *out_comm = 0;
  Comm* c = nullptr;
  int rc = comm_alloc(device, rank, world, max_B, max_k, max_graph_k, &c);
  if (rc) return rc;
  if (world == 1) c->connected = true;
  *out_comm = reinterpret_cast<uint64_t>(c);
  return B200R_OK;
}

int b200r_comm_export(uint64_t comm, void* blob_host) {
  Comm* c = get_comm(comm);
  if (!c) return B200R_ESTATE;
  if (!blob_host) { set_error("blob_host is null"); return B200R_EINVAL; }
  CommBlob b;
  fill_blob(c, &b);
  memset(blob_host, 0, B200R_COMM_BLOB_BYTES);
  memcpy(blob_host, &b, sizeof(b));
  return B200R_OK;
}

int b200r_comm_connect(uint64_t comm, const void* blobs_host) {
  Comm* c = get_comm(comm);
  if (!c) return B200R_ESTATE;
  if (!blobs_host) { set_error("blobs_host is null"); return B200R_EINVAL; }
  if (c->connected && c->world > 1) { set_error("comm is already connected"); return B200R_ESTATE; }
  CommBlob blobs[B200R_MAX_PEERS];
  for (int g = 0; g < c->world; ++g) memcpy(&blobs[g], (const char*)blobs_host + (size_t)g * B200R_COMM_BLOB_BYTES, sizeof(CommBlob));
  return connect_blobs(c, blobs);
}

int b200r_comm_init_all(const int32_t* devices, int32_t G, int32_t max_B, int32_t max_k, int32_t max_graph_k, uint64_t* out_comms) {
  if (!devices || !out_comms || G < 1 || G > B200R_MAX_PEERS) { set_error("bad devices / G (G <= %d)", B200R_MAX_PEERS); return B200R_EINVAL; }
  Comm* cs[B200R_MAX_PEERS] = {nullptr};
  CommBlob blobs[B200R_MAX_PEERS];
  int rc = B200R_OK;
  for (int g = 0; g < G && !rc; ++g) {
    rc = comm_alloc(devices[g], g, G, max_B, max_k, max_graph_k, &cs[g]);
    if (!rc) { fill_blob(cs[g], &blobs[g]); }
  }
  for (int g = 0; g < G && !rc; ++g) rc = connect_blobs(cs[g], blobs);
  if (rc) {
    for (int g = 0; g < G; ++g) if (cs[g]) { cudaSetDevice(cs[g]->device); cudaFree(cs[g]->window); delete cs[g]; }
    return rc;
  }
  for (int g = 0; g < G; ++g) out_comms[g] = reinterpret_cast<uint64_t>(cs[g]);
  return B200R_OK;
}

int b200r_comm_destroy(uint64_t comm) {
  Comm* c = get_comm(comm);
  if (!c) return B200R_ESTATE;
  cudaSetDevice(c->device);
  for (int g = 0; g < c->world; ++g)
    if (c->ipc_opened[g]) cudaIpcCloseMemHandle(reinterpret_cast<void*>(c->peer[g]));
  if (c->window) cudaFree(c->window);
  c->magic = 0;
  delete c;
  return B200R_OK;
}

int b200r_comm_status(uint64_t comm, void* stream) {
  Comm* c = get_comm(comm);
  if (!c) return B200R_ESTATE;
  B200R_CUDA_CHECK(cudaSetDevice(c->device));
  u32 status = 0;
  B200R_CUDA_CHECK(cudaMemcpyAsync(&status, &reinterpret_cast<CommCtrl*>(c->window)->status, sizeof(u32), cudaMemcpyDeviceToHost,
                                   reinterpret_cast<cudaStream_t>(stream)));
  B200R_CUDA_CHECK(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream)));
  if (status) { set_error("a device-side exchange wait timed out (a rank never arrived); results of that exchange are invalid"); return B200R_ETIMEOUT; }
  return B200R_OK;
}

int b200r_comm_info(uint64_t comm, int32_t* rank, int32_t* world, int32_t* device, size_t* window_bytes) {
  Comm* c = get_comm(comm);
  if (!c) return B200R_ESTATE;
  if (rank) *rank = c->rank;
  if (world) *world = c->world;
  if (device) *device = c->device;
  if (window_bytes) *window_bytes = c->window_bytes;
  return B200R_OK;
}

}  // extern "C"
