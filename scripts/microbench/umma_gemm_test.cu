// Bring-up test for the tcgen05 path used by the fused actor kernel:
//   D[128 x N] (fp32, TMEM) = A[128 x K] * B[N x K]^T, bf16 operands in shared memory,
//   K-major, no swizzle (8x8 core matrices, LBO = 128 B along K, SBO = K/8 * 128 B along M/N).
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o umma_gemm_test umma_gemm_test.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // version = 1 (Blackwell)
  return d;                 // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
      "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  }
}

template <int N, int K>
__global__ void __launch_bounds__(128) gemm_kernel(const __nv_bfloat16* __restrict__ Apk, const __nv_bfloat16* __restrict__ Bpk,
                                                   float* __restrict__ D) {
  extern __shared__ __align__(128) unsigned char smem[];
  __nv_bfloat16* sA = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sB = sA + 128 * K;
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * K / 8; i += 128) reinterpret_cast<uint4*>(sA)[i] = reinterpret_cast<const uint4*>(Apk)[i];
  for (int i = tid; i < N * K / 8; i += 128) reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(Bpk)[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic smem writes -> async proxy (UMMA)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    for (int s = 0; s < K / 16; ++s) {
      const uint64_t ad = make_desc(a0 + s * 256, 128, (K / 8) * 128);
      const uint64_t bd = make_desc(b0 + s * 256, 128, (K / 8) * 128);
      umma_f16(tmem_base, ad, bd, idesc, s > 0 ? 1u : 0u);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
  }
  mbar_wait(smem_u32(&mbar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // epilogue: thread = row (TMEM lane), 32 columns per tcgen05.ld
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
        "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32 && c0 + j < N; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u));
}

static void pack(const std::vector<float>& src, int rows, int K, std::vector<__nv_bfloat16>& dst) {
  dst.resize((size_t)rows * K);
  for (int r = 0; r < rows; ++r)
    for (int k = 0; k < K; ++k) {
      const size_t off = (size_t)(r / 8) * (K / 8) * 64 + (size_t)(k / 8) * 64 + (r % 8) * 8 + (k % 8);
      dst[off] = __float2bfloat16(src[(size_t)r * K + k]);
    }
}

template <int N, int K>
int run() {
  std::vector<float> A(128 * K), B((size_t)N * K);
  srand(1);
  for (auto& x : A) x = (rand() % 2001 - 1000) / 1000.0f;
  for (auto& x : B) x = (rand() % 2001 - 1000) / 1000.0f;
  std::vector<__nv_bfloat16> Ap, Bp;
  pack(A, 128, K, Ap); pack(B, N, K, Bp);
  __nv_bfloat16 *dA, *dB; float* dD;
  CK(cudaMalloc(&dA, Ap.size() * 2)); CK(cudaMalloc(&dB, Bp.size() * 2)); CK(cudaMalloc(&dD, 128 * N * 4));
  CK(cudaMemcpy(dA, Ap.data(), Ap.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bp.data(), Bp.size() * 2, cudaMemcpyHostToDevice));
  const size_t smem = (size_t)(128 + N) * K * 2;
  CK(cudaFuncSetAttribute(gemm_kernel<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gemm_kernel<N, K><<<1, 128, smem>>>(dA, dB, dD);
  CK(cudaDeviceSynchronize());
  std::vector<float> D(128 * N);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k)
        ref += (double)__bfloat162float(__float2bfloat16(A[m * K + k])) * (double)__bfloat162float(__float2bfloat16(B[(size_t)n * K + k]));
      maxerr = fmax(maxerr, fabs(ref - D[m * N + n]));
    }
  printf("N=%d K=%d max abs err %.3e  (D[0][0]=%f D[127][%d]=%f)\n", N, K, maxerr, D[0], N - 1, D[127 * N + N - 1]);
  return maxerr < 1e-3 ? 0 : 1;
}

int main() {
  int bad = 0;
  bad += run<256, 80>();
  bad += run<256, 256>();
  bad += run<16, 256>();
  printf(bad ? "FAILED\n" : "OK\n");
  return bad;
}
