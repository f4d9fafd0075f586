// Issue rate of tcgen05.mma.cta_group::2 (M = 256, K = 16, bf16) on operands resident in the two CTAs' shared memory:
// the canonical no-swizzle K-major layout our kernels use against the 128-byte-swizzled layout CUTLASS uses.
// Only the descriptors differ (the data is irrelevant for a rate measurement).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_rate_2cta umma_rate_2cta.cu && ./umma_rate_2cta
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../marl_gym_pybullet_drones_b200/csrc/bd_umma.cuh"
using namespace bdu;

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint64_t desc_any(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout_type & 7) << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_m(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// swz: 0 = no swizzle (core matrices 128 B apart along K, 8-row groups K/8*128 B apart), 1 = SWIZZLE_128B (K = 64 element
// atoms of 8 rows x 128 B: SBO 1024, a K = 16 step advances the start address by 32 B)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) rate2(int n, int layers, int swz, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = tid; i < (65536 + 65536) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) { mbar_init(smem_u32(&done_bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s, aA = smem_u32(smem), aB = aA + 65536;
  if (rank == 0 && warp == 1 && lane == 0) {
    const uint64_t dA0 = swz ? desc_any(aA, 16, 1024, 2) : desc_any(aA, 128, 4096, 0);
    const uint64_t dB0 = swz ? desc_any(aB, 16, 1024, 2) : desc_any(aB, 128, 4096, 0);
    const uint32_t id = idesc_m(256, n);
    const long long t0 = clock64();
    for (int l = 0; l < layers; ++l) {
#pragma unroll
      for (int s = 0; s < 16; ++s) {
        // no swizzle: K-step = 256 B further; 128B swizzle: 32 B inside the 128-byte atom, next atom column every 4 steps
        const uint64_t off = swz ? (uint64_t)(((s & 3) * 32 + (s >> 2) * 16384) >> 4) : (uint64_t)(16 * s);
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem + (uint32_t)(l & 1) * 256),
            "l"(dA0 + off), "l"(dB0 + off), "r"(id), "r"((uint32_t)(s != 0))
            : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(&done_bar)),
                 "h"((uint16_t)3)
                 : "memory");
    mbar_wait(smem_u32(&done_bar), 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = (long long)layers * 16; }
  } else if (warp == 1 && lane == 0) {
    mbar_wait(smem_u32(&done_bar), 0);   // the peer keeps its shared memory until the pair's MMAs are complete
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

int main() {
  long long* out;
  cudaMalloc(&out, 16);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int smem = 131072;
  cudaFuncSetAttribute(rate2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int swz = 0; swz < 2; ++swz)
    for (int n : {256, 128, 64, 16}) {
      for (int grid : {2, (sms / 2) * 2}) {
        rate2<<<grid, 128, smem>>>(n, 64, swz, out);
        cudaDeviceSynchronize();
        rate2<<<grid, 128, smem>>>(n, 256, swz, out);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2];
        cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        printf("cta_group::2 M=256 N=%3d %-12s grid %3d: %7.1f cycles per MMA%s\n", n, swz ? "SWIZZLE_128B" : "no swizzle", grid,
               (double)h[0] / (double)h[1], e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
    }
  return 0;
}
