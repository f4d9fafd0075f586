// Gradient all-reduce over NVLink peer memory, one kernel per call (sm_100a, one process per GPU).
//
// Replaces the NCCL all-reduce of the trainer's flat gradient buffer — the collective of
// `MAPPOAgent.update` (reference gym_pybullet_drones/mappo/agent.py:702-772 runs one process; the
// data-parallel trainer all-reduces actor + critic gradients and the KL pair of the gate once per
// minibatch).  The buffers are small (1.8 MB) and the collective is latency-bound: NCCL takes
// ~31 us for the gradients plus ~25 us for the two doubles of the KL gate, 128 times per epoch.
//
// Every rank allocates the same block with cudaMalloc, exports it with cudaIpcGetMemHandle and maps
// its peers' blocks (cudaIpcOpenMemHandle, peer access through NVSwitch).  The gradient kernels write
// straight into the block's data region (the torch gradient tensors are views of it), so the
// collective has no staging copy.  One launch of `peer_allreduce_kernel` per rank then does, IN PLACE:
//   entry   block 0 copies the rank's `extra` doubles (the KL pair) into the block, fences at system
//           scope and stores the call's sequence number into every peer's flag_in[rank]; every CTA
//           waits until all peers' numbers have arrived in its own flag_in[].
//   reduce  rank r owns slice r of the vector: it loads that slice from every peer's data region
//           (volatile 128-bit loads over NVLink), adds the W values in rank order and stores the sum
//           back into slice r of EVERY rank's data region.  Slice r is read and written by rank r
//           alone, so in place is safe, and every rank ends up with bit-identical sums.
//           The extra doubles are summed by every rank itself (rank order, all peers' copies).
//   exit    the last CTA of the rank (ticket counter) fences, stores the sequence number into every
//           peer's flag_out[rank], and waits for all peers' flag_out: when the kernel ends, every slice
//           has arrived and no peer still reads this rank's data — the next kernel may overwrite it.
// The sequence number lives in device memory and is advanced by the kernel, so the launch can be
// captured in a CUDA graph and replayed.  Spins are bounded (trap after ~20 s): a lost peer fails the
// launch instead of hanging the GPU.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <new>

#include "../../include/batch_drones.h"

namespace {

constexpr int kMaxWorld = 16;
constexpr int kMaxExtra = 16;
constexpr int kThreads = 512;

thread_local char g_peer_err[256] = "";
int pfail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_peer_err, sizeof(g_peer_err), fmt, ap);
  va_end(ap);
  return code;
}

// control words of a rank's block; every flag array on its own 128-byte lines
struct PeerCtl {
  unsigned int flag_in[32];    // [peer]: sequence number of the peer's last "my data is ready"
  unsigned int flag_out[32];   // [peer]: sequence number of the peer's last "my slice has been written everywhere"
  double extra[kMaxExtra];     // this rank's extra doubles of the current call
  unsigned int seq;            // calls completed by this rank (local use)
  unsigned int done;           // CTAs of the current call that finished their part (local use)
  unsigned int pad[30];
};

struct PeerArgs {
  float* data[kMaxWorld];      // data region of every rank (own pointer for the own rank)
  PeerCtl* ctl[kMaxWorld];
  int rank, world;
  long long n4;                // vector length in float4
  double* extra;               // local, in place; may be null
  int n_extra;
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_volatile_f4(const float4* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
// wait until *flag has reached seq (sequence numbers wrap: compare as a signed difference)
__device__ __forceinline__ void wait_flag(const unsigned int* flag, unsigned int seq) {
  const long long t0 = clock64();
  while ((int)(ld_acquire_sys(flag) - seq) < 0) {
    __nanosleep(64);
    if (clock64() - t0 > 40000000000LL) __trap();      // ~20 s of SM clocks
  }
}

__global__ void __launch_bounds__(kThreads) peer_allreduce_kernel(PeerArgs A) {
  PeerCtl* const me = A.ctl[A.rank];
  const int W = A.world, r = A.rank, tid = threadIdx.x;
  const unsigned int seq = *reinterpret_cast<volatile unsigned int*>(&me->seq) + 1u;
  // ---- entry: my data (written by earlier kernels of this stream) and my extras are ready
  if (blockIdx.x == 0) {
    if (tid < A.n_extra) me->extra[tid] = A.extra[tid];
    __syncthreads();
    if (tid < W && tid != r) {
      __threadfence_system();
      st_release_sys(&A.ctl[tid]->flag_in[r], seq);
    }
  }
  if (tid < W && tid != r) wait_flag(&me->flag_in[tid], seq);
  __syncthreads();
  // ---- extras: every rank sums all copies itself, in rank order
  if (blockIdx.x == 0 && tid < A.n_extra) {
    double s = 0.0;
    for (int p = 0; p < W; ++p) s += ld_volatile_f64(&A.ctl[p]->extra[tid]);
    A.extra[tid] = s;
  }
  // ---- reduce my slice, write it to everybody
  const long long per = (A.n4 + W - 1) / W;
  const long long lo = per * r, hi = (lo + per < A.n4) ? lo + per : A.n4;
  for (long long i = lo + (long long)blockIdx.x * kThreads + tid; i < hi; i += (long long)gridDim.x * kThreads) {
    float4 acc = ld_volatile_f4(reinterpret_cast<const float4*>(A.data[0]) + i);
    for (int p = 1; p < W; ++p) {
      const float4 v = ld_volatile_f4(reinterpret_cast<const float4*>(A.data[p]) + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    for (int p = 0; p < W; ++p) reinterpret_cast<float4*>(A.data[p])[i] = acc;
  }
  // ---- exit: the rank's last CTA tells everybody, and waits for everybody
  __syncthreads();
  __shared__ unsigned int ticket;
  if (tid == 0) {
    __threadfence_system();
    ticket = atomicAdd(&me->done, 1u);
  }
  __syncthreads();
  if (ticket != gridDim.x - 1) return;
  if (tid < W && tid != r) {
    __threadfence_system();
    st_release_sys(&A.ctl[tid]->flag_out[r], seq);
    wait_flag(&me->flag_out[tid], seq);
  }
  __syncthreads();
  if (tid == 0) {
    me->done = 0u;
    *reinterpret_cast<volatile unsigned int*>(&me->seq) = seq;
  }
}

}  // namespace

struct bd_peer {
  int device = 0, rank = 0, world = 1;
  size_t floats = 0;           // capacity of the data region (multiple of 4)
  char* block = nullptr;       // own allocation: [data | PeerCtl]
  void* mapped[kMaxWorld] = {};   // peers' blocks (IPC mappings), null for the own rank
  bool opened = false;
  int sm_count = 0;
  int64_t launches = 0;
};

extern "C" {

const char* bd_peer_last_error(void) { return g_peer_err; }

int bd_peer_create(int device, int rank, int world, int64_t floats, bd_peer** out) {
  if (!out) return pfail(BD_EINVAL, "bd_peer_create: null out");
  *out = nullptr;
  if (world < 2 || world > kMaxWorld || rank < 0 || rank >= world) return pfail(BD_EINVAL, "bd_peer_create: world in [2,%d], 0 <= rank < world", kMaxWorld);
  if (floats < 4) return pfail(BD_EINVAL, "bd_peer_create: at least 4 floats");
  bd_peer* p = new (std::nothrow) bd_peer();
  if (!p) return pfail(BD_ECUDA, "bd_peer_create: out of host memory");
  p->device = device; p->rank = rank; p->world = world;
  p->floats = ((size_t)floats + 3) & ~(size_t)3;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaError_t e = cudaSetDevice(device);
  cudaDeviceProp prop;
  if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->block, p->floats * 4 + sizeof(PeerCtl));
  if (e == cudaSuccess) e = cudaMemset(p->block, 0, p->floats * 4 + sizeof(PeerCtl));
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (prev >= 0 && prev != device) cudaSetDevice(prev);
  if (e != cudaSuccess) {
    if (p->block) cudaFree(p->block);
    delete p;
    return pfail(BD_ECUDA, "bd_peer_create: %s", cudaGetErrorString(e));
  }
  p->sm_count = prop.multiProcessorCount;
  *out = p;
  return BD_OK;
}

int bd_peer_handle_size(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int bd_peer_get_handle(bd_peer* p, void* handle_out) {
  if (!p || !handle_out) return pfail(BD_EINVAL, "bd_peer_get_handle: null argument");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p->block);
  if (e != cudaSuccess) return pfail(BD_ECUDA, "bd_peer_get_handle: %s", cudaGetErrorString(e));
  memcpy(handle_out, &h, sizeof(h));
  return BD_OK;
}

int bd_peer_open(bd_peer* p, const void* handles, int count) {
  if (!p || !handles || count != p->world) return pfail(BD_EINVAL, "bd_peer_open: one handle per rank expected");
  if (p->opened) return pfail(BD_EINVAL, "bd_peer_open: already opened");
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(p->device);
  cudaError_t e = cudaSuccess;
  for (int r = 0; r < p->world && e == cudaSuccess; ++r) {
    if (r == p->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)r * sizeof(h), sizeof(h));
    e = cudaIpcOpenMemHandle(&p->mapped[r], h, cudaIpcMemLazyEnablePeerAccess);
  }
  if (e != cudaSuccess) {
    for (int r = 0; r < p->world; ++r)
      if (p->mapped[r]) { cudaIpcCloseMemHandle(p->mapped[r]); p->mapped[r] = nullptr; }
    cudaGetLastError();
  } else {
    p->opened = true;
  }
  if (prev >= 0 && prev != p->device) cudaSetDevice(prev);
  if (e != cudaSuccess) return pfail(BD_ECUDA, "bd_peer_open: %s", cudaGetErrorString(e));
  return BD_OK;
}

float* bd_peer_data(bd_peer* p) { return p ? reinterpret_cast<float*>(p->block) : nullptr; }

int bd_peer_allreduce(bd_peer* p, int64_t floats, double* extra_dev, int n_extra, void* stream) {
  if (!p || !p->opened) return pfail(BD_EINVAL, "bd_peer_allreduce: handle not opened");
  if (floats < 0 || (size_t)floats > p->floats) return pfail(BD_EINVAL, "bd_peer_allreduce: %lld floats > capacity %zu", (long long)floats, p->floats);
  if (n_extra < 0 || n_extra > kMaxExtra || (n_extra > 0 && !extra_dev)) return pfail(BD_EINVAL, "bd_peer_allreduce: 0..%d extra doubles", kMaxExtra);
  PeerArgs A;
  memset(&A, 0, sizeof(A));
  for (int r = 0; r < p->world; ++r) {
    char* base = (r == p->rank) ? p->block : (char*)p->mapped[r];
    A.data[r] = reinterpret_cast<float*>(base);
    A.ctl[r] = reinterpret_cast<PeerCtl*>(base + p->floats * 4);
  }
  A.rank = p->rank; A.world = p->world;
  A.n4 = (floats + 3) / 4;                 // the data region is zero-padded to a multiple of 4
  A.extra = extra_dev; A.n_extra = n_extra;
  const long long per = (A.n4 + p->world - 1) / p->world;
  long long grid = (per + kThreads - 1) / kThreads;
  if (grid < 1) grid = 1;
  if (grid > p->sm_count) grid = p->sm_count;      // all CTAs of the call must be resident: they wait for each other's peers
  peer_allreduce_kernel<<<(int)grid, kThreads, 0, (cudaStream_t)stream>>>(A);
  p->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return pfail(BD_ECUDA, "bd_peer_allreduce: %s", cudaGetErrorString(e));
  return BD_OK;
}

int64_t bd_peer_launch_count(const bd_peer* p) { return p ? p->launches : 0; }

/* Close the mappings of the peers' blocks (the own block stays).  Shutdown order: every rank unmaps, the ranks meet at a
 * barrier of the caller's transport, then every rank destroys — an exporter must not free a block a peer still maps. */
int bd_peer_unmap(bd_peer* p) {
  if (!p) return pfail(BD_EINVAL, "bd_peer_unmap: null handle");
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(p->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < p->world; ++r)
    if (p->mapped[r]) { cudaIpcCloseMemHandle(p->mapped[r]); p->mapped[r] = nullptr; }
  p->opened = false;
  if (prev >= 0 && prev != p->device) cudaSetDevice(prev);
  return BD_OK;
}

void bd_peer_destroy(bd_peer* p) {
  if (!p) return;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(p->device);
  for (int r = 0; r < p->world; ++r)
    if (p->mapped[r]) cudaIpcCloseMemHandle(p->mapped[r]);
  if (p->block) cudaFree(p->block);
  if (prev >= 0 && prev != p->device) cudaSetDevice(prev);
  delete p;
}

}  // extern "C"
