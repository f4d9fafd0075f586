// How fast does ONE CTA issue tcgen05.mma (cta_group::1, kind::f16, M = 128, K = 16) on operands resident in shared
// memory, in the canonical no-swizzle K-major layout the PPO / actor kernels use?  And how much of that survives when the
// B operand is streamed through a TMA slab ring while other warps write activations to shared memory?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu && ./umma_rate
//
// Variants (cycles per MMA instruction, one CTA per SM, all SMs):
//   resident        A [128 x 256], B [256 x 256] in smem, 16 K-steps per "layer", back-to-back
//   resident+stores the same while 512 other threads keep writing 16-byte vectors into a third buffer
//   streamed        B slabs (8 KB per K-step) arrive through a 6-stage cp.async.bulk ring from L2
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../marl_gym_pybullet_drones_b200/csrc/bd_umma.cuh"
using namespace bdu;

constexpr int kThreads = 576;
constexpr int kStages = 6;

// mode 0: resident, 1: resident + st.shared traffic, 2: streamed B
__global__ void __launch_bounds__(kThreads, 1) rate_kernel(int mode, int n, int layers, const __nv_bfloat16* slabs, long long* out,
                                                            uint32_t sbo_a, uint32_t lbo_a, uint32_t sbo_b, uint32_t kstep_b, int chains) {
  extern __shared__ __align__(128) unsigned char smem[];
  __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(smem);                 // 64 KB
  unsigned char* B = smem + 65536;                                           // resident: 128 KB; streamed: ring
  unsigned char* scratch = smem + 65536 + (mode == 2 ? kStages * 8192 : 131072);
  __shared__ __align__(8) uint64_t bars[2 * kStages + 2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (65536 + (mode == 2 ? kStages * 8192 : 131072)) / 4; i += kThreads) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) {
    for (int i = 0; i < 2 * kStages + 2; ++i) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s, aA = smem_u32(A), aB = smem_u32(B), bar0 = smem_u32(&bars[0]);
  auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  const int DONE = 2 * kStages;
  if (warp == 16 && lane == 0) {
    // tight issue loop: descriptors are precomputed, a K-step only adds a constant to their address field
    const uint64_t dA0 = umma_desc(aA, lbo_a, sbo_a), dB0 = umma_desc(aB, 128, sbo_b);
    const uint64_t incA = (uint64_t)((2 * lbo_a) >> 4), incB = (uint64_t)(kstep_b >> 4), incRing = (uint64_t)(8192 >> 4);
    const uint32_t idesc = umma_idesc(n);
    const uint32_t cstride = 512u / (uint32_t)chains;
    const long long t0 = clock64();
    int st = 0;
    uint32_t par = 0;
    for (int l = 0; l < layers; ++l) {
      const uint32_t acc_l = tmem + (uint32_t)(l & 1) * 256;
      if (mode != 2) {
#pragma unroll
        for (int s = 0; s < 16; ++s) {
          const uint32_t acc = chains > 1 ? tmem + (uint32_t)(s & (chains - 1)) * cstride : acc_l;
          umma_bf16(acc, dA0 + incA * s, dB0 + incB * s, idesc, chains > 1 ? (s >= chains) : (s != 0));
        }
      } else {
#pragma unroll
        for (int s = 0; s < 16; ++s) {
          mbar_wait(bar(st), par);
          umma_bf16(acc_l, dA0 + incA * s, dB0 + incRing * st, idesc, s != 0);
          umma_commit(bar(kStages + st));
          if (++st == kStages) { st = 0; par ^= 1; }
        }
      }
    }
    umma_commit(bar(DONE));
    mbar_wait(bar(DONE), 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = (long long)layers * 16; }
  } else if (warp == 17 && lane == 0 && mode == 2) {
    int st = 0;
    uint32_t epar = 1;
    const long long total = (long long)layers * 16;
    for (long long c = 0; c < total; ++c) {
      if (c >= kStages) mbar_wait(bar(kStages + st), epar);
      mbar_expect_tx(bar(st), 8192);
      bulk_g2s(aB + st * 8192, slabs + (c & 15) * 4096, 8192, bar(st));
      if (++st == kStages) { st = 0; epar ^= 1; }
    }
  } else if (warp < 16 && mode == 1) {
    // activation-like store traffic: 64 KB per ~2000 cycles is what the real epilogue writes; here as fast as it goes
    uint4* dst = reinterpret_cast<uint4*>(scratch);
    for (int it = 0; it < layers * 8; ++it) dst[(tid + it * 512) & 2047] = make_uint4(it, it, it, it);
  } else if (warp < 16 && mode >= 3) {
    // what an epilogue does while the tensor core works on the OTHER accumulator (columns 256..511 here, the MMAs write
    // 0..255): mode 3 = TMEM loads only, 4 = tanh (MUFU) only, 5 = canonical-layout 16-byte shared-memory stores only,
    // 6 = all three, i.e. the real epilogue loop
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const int grp = warp >> 2, row = tid & 127;
    float acc = 0.f;
    __nv_bfloat16* sH = reinterpret_cast<__nv_bfloat16*>(scratch);   // 32 KB: a quarter of a real activation tile (wraps)
    for (int it = 0; it < layers * 2; ++it) {
      uint32_t v[16];
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        if (mode == 3 || mode == 6) {
          tmem_ld16_nowait(tmem + 256u + lane_off + (uint32_t)(grp * 64 + j * 16), v);
          tmem_wait_ld();
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = __float_as_uint(acc + (float)e);
        }
        float z[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) z[e] = (mode == 4 || mode == 6) ? tanh_fast(__uint_as_float(v[e]) + 0.1f) : __uint_as_float(v[e]);
        if (mode == 5 || mode == 6) {
#pragma unroll
          for (int q = 0; q < 2; ++q)
            *reinterpret_cast<uint4*>(sH + (canon_off(row, (grp * 64 + j * 16 + q * 8) & 127, 128))) =
                make_uint4(pack_bf16(z[q * 8], z[q * 8 + 1]), pack_bf16(z[q * 8 + 2], z[q * 8 + 3]), pack_bf16(z[q * 8 + 4], z[q * 8 + 5]),
                           pack_bf16(z[q * 8 + 6], z[q * 8 + 7]));
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) acc += z[e];
        }
      }
    }
    if (acc == 123.456f) out[2] = 1;   // keep the work alive
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

int main() {
  long long* out;
  cudaMalloc(&out, 16);
  __nv_bfloat16* slabs;
  cudaMalloc(&slabs, 16 * 8192);
  cudaMemset(slabs, 0, 16 * 8192);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int smem = 65536 + 131072 + 32768;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[7] = {"resident", "resident+stores", "streamed (6 x 8 KB ring)", "resident + TMEM loads", "resident + tanh",
                          "resident + canonical stores", "resident + full epilogue loop"};
  // operand layouts: A K-major [128 x 256]: lbo = distance between the two core matrices of a K-step, sbo = between 8-row
  // groups.  "4096" = dense canonical tile; "4224" = row groups padded by 128 B; "kblock" = K-step-major blocks
  // [16 K-steps][16 row groups][2][128 B] (lbo 128, sbo 256, K-step stride 4096).  B resident: slab per K-step (sbo 256,
  // K-step stride 8192) or dense [256 x 256] tile (sbo 4096, K-step stride 256).
  struct Cfg { const char* name; uint32_t sbo_a, lbo_a, sbo_b, kstep_b; };
  const Cfg cfgs[] = {{"A dense sbo 4096, B slabs", 4096, 128, 256, 8192}, {"A padded sbo 4224, B slabs", 4224, 128, 256, 8192},
                      {"A dense, B dense sbo 4096", 4096, 128, 4096, 256}, {"A padded, B padded 4224", 4224, 128, 4224, 256}};
  // `chains`: consecutive MMAs go round-robin to this many independent accumulators (1: every K-step of a layer
  // accumulates into the same TMEM tile, as a GEMM main loop does)
  for (int chains : {1, 2, 4})
  for (const Cfg& c : cfgs)
    for (int n : {256, 128, 64}) {
      if (c.sbo_a != 4096 || c.sbo_b != 256) { if (chains > 1) continue; }
      if (n * chains > 512) continue;
      for (int mode = 0; mode < 7; ++mode) {
        if (mode == 2 && c.sbo_b != 256) continue;
        if ((mode == 1 || mode >= 3) && (n != 256 || chains != 1 || c.sbo_a != 4096 || c.sbo_b != 256)) continue;
        if (c.sbo_b == 4224 && n == 256) continue;   // 256 rows x 4224 B does not fit next to A
        const int grid = sms;
        rate_kernel<<<grid, kThreads, smem>>>(mode, n, 64, slabs, out, c.sbo_a, c.lbo_a, c.sbo_b, c.kstep_b, chains);
        cudaDeviceSynchronize();
        rate_kernel<<<grid, kThreads, smem>>>(mode, n, 256, slabs, out, c.sbo_a, c.lbo_a, c.sbo_b, c.kstep_b, chains);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2];
        cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        printf("chains %d %-28s N=%3d %-26s: %7.1f cycles per MMA%s\n", chains, c.name, n, names[mode], (double)h[0] / (double)h[1],
               e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
    }
  return 0;
}
