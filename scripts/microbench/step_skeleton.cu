// Memory skeleton of one control step (no arithmetic): which structure moves the step's
// bytes fastest?  read 4 state planes + action + 14 history planes; write 4 state planes,
// one ring slot and complete 288-B observation rows.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o step_skeleton step_skeleton.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
constexpr int B = 15, NS = 14, D = 72;

struct Bufs { float4 *s0, *s1, *s2, *s3, *hist; const float4* act; float* obs; size_t n; int head; };

__device__ __forceinline__ void bulk_store(void* g, const void* s, unsigned bytes) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(s);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(g), "r"(sa), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
}

// ROWS per CTA = blockDim; persistent loop over tiles; MODE 0: cooperative STG, 1: TMA bulk store
template <int ROWS, int MODE>
__global__ void __launch_bounds__(ROWS) k_skel(Bufs b, int ntiles) {
  extern __shared__ __align__(128) float tile[];   // [ROWS][D]
  const int tid = threadIdx.x;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const size_t g = (size_t)t * ROWS + tid;
    float4 a0 = b.s0[g], a1 = b.s1[g], a2 = b.s2[g], a3 = b.s3[g], ac = b.act[g];
    float4 v[NS];
    int slot = b.head;
#pragma unroll
    for (int k = 0; k < NS; ++k) { slot = (slot + 1 == B) ? 0 : slot + 1; v[k] = b.hist[(size_t)slot * b.n + g]; }
    if (MODE == 1) { if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
    __syncthreads();
    float4* row = reinterpret_cast<float4*>(tile + (size_t)tid * D);
    row[0] = a0; row[1] = a1; row[2] = a2;
#pragma unroll
    for (int k = 0; k < NS; ++k) row[3 + k] = v[k];
    row[17] = ac;
    b.hist[(size_t)b.head * b.n + g] = ac;
    a0.x += 1.f; a1.x += 1.f; a2.x += 1.f; a3.x += 1.f;
    b.s0[g] = a0; b.s1[g] = a1; b.s2[g] = a2; b.s3[g] = a3;
    float* gob = b.obs + (size_t)t * ROWS * D;
    if (MODE == 0) {
      __syncthreads();
      const float4* s4 = reinterpret_cast<const float4*>(tile);
      float4* g4 = reinterpret_cast<float4*>(gob);
      for (int i = tid; i < ROWS * D / 4; i += ROWS) g4[i] = s4[i];
    } else {
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      __syncthreads();
      if (tid == 0) bulk_store(gob, tile, ROWS * D * 4);
    }
  }
  if (MODE == 1 && tid == 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
}

int main(int argc, char** argv) {
  const size_t n = argc > 1 ? atoll(argv[1]) : 262144;
  const int slots = 8, iters = 100;
  Bufs b; b.n = n;
  float4* act;
  CK(cudaMalloc(&b.s0, n * 16)); CK(cudaMalloc(&b.s1, n * 16)); CK(cudaMalloc(&b.s2, n * 16)); CK(cudaMalloc(&b.s3, n * 16));
  CK(cudaMalloc(&b.hist, (size_t)B * n * 16)); CK(cudaMalloc(&act, (size_t)slots * n * 16));
  CK(cudaMalloc(&b.obs, (size_t)slots * n * D * 4));
  CK(cudaMemset(b.s0, 0, n * 16)); CK(cudaMemset(b.s1, 0, n * 16)); CK(cudaMemset(b.s2, 0, n * 16)); CK(cudaMemset(b.s3, 0, n * 16));
  CK(cudaMemset(b.hist, 0, (size_t)B * n * 16)); CK(cudaMemset(act, 0, (size_t)slots * n * 16));
  float* obs0 = b.obs;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const double alg = (double)n * 647.5, act_bytes = (double)n * (304 + 368);
  auto run = [&](const char* name, auto&& launch) {
    for (int i = 0; i < 10; ++i) launch(i);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) launch(i);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize()); CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double t = ms * 1e-3 / iters;
    printf("%-46s %8.2f us  algorithmic %7.1f GB/s (frac of 6553: %.3f)  actual %7.1f GB/s\n", name, t * 1e6, alg / t / 1e9, alg / t / 6553e9, act_bytes / t / 1e9);
  };
  auto prep = [&](int i) { b.head = i % B; b.act = act + (size_t)(i % slots) * n; b.obs = obs0 + (size_t)(i % slots) * n * D; };
#define RUN(ROWS, MODE, GRID, label) { CK(cudaFuncSetAttribute(k_skel<ROWS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, ROWS * D * 4)); \
    run(label, [&](int i) { prep(i); k_skel<ROWS, MODE><<<GRID, ROWS, ROWS * D * 4>>>(b, (int)(n / ROWS)); }); }
  RUN(128, 0, (int)(n / 128), "128 rows/CTA, STG, one CTA per tile");
  RUN(128, 0, 148 * 6, "128 rows/CTA, STG, persistent 148x6");
  RUN(128, 1, (int)(n / 128), "128 rows/CTA, TMA store, one CTA per tile");
  RUN(128, 1, 148 * 6, "128 rows/CTA, TMA store, persistent 148x6");
  RUN(64, 1, 148 * 12, "64 rows/CTA, TMA store, persistent 148x12");
  RUN(32, 1, 148 * 24, "32 rows/CTA, TMA store, persistent 148x24");
  RUN(32, 1, (int)(n / 32), "32 rows/CTA, TMA store, one CTA per tile");
  RUN(256, 1, 148 * 3, "256 rows/CTA, TMA store, persistent 148x3");
  RUN(256, 0, 148 * 3, "256 rows/CTA, STG, persistent 148x3");
  return 0;
}
