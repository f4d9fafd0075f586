// Running mean / variance of observation batches on the device.
//
// Replaces `RunningMeanStd.update` + `MeanStdNormalizer.__call__` of the reference trainer
// (gym_pybullet_drones/safe_control_gym/math_and_models/normalization.py:13-96), which MAPPO
// applies to every observation batch of the rollout (`mappo/mappo.py:132,165,804`): batch moments
// over the env axis, parallel-variance merge into the running statistics, then
// clip((x - mean) / sqrt(var + eps), +-clip).
//
// One launch per update: `moments_flat_kernel` streams the (rows x cols) batch once with 128-bit loads
// (HBM-bound; per-CTA row slabs, register accumulators, one fp64 atomic pair per column and CTA); the
// last CTA to finish turns the sums into the batch's [mean | var | count] and either writes them out
// (the unit ranks exchange: one small all-gather when the envs are sharded over GPUs, followed by
// `merge_kernel`, which folds several such parts in rank order) or folds them straight into
// mean / var / count in fp64 and emits float mean and 1/sqrt(var + eps) vectors for fused consumers:
// the actor kernel normalises its input tile while it converts it to bf16 (bd_actor_set_input_norm), so
// the normalised observations are never written to HBM during the rollout.  `normalize_flat_kernel` is
// the standalone form (evaluation, critic input).
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <new>

#include "../../include/batch_drones.h"

namespace {

thread_local char g_rms_err[256] = "";
int rfail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_rms_err, sizeof(g_rms_err), fmt, ap);
  va_end(ap);
  return code;
}

constexpr int kMomThreads = 256;

// acc[c] += sum_r (x[r][c] - x[0][c]),  acc[cols + c] += sum_r (x[r][c] - x[0][c])^2
// (shifting by a sample of the batch keeps the one-pass variance well conditioned whatever the
// column's offset: the differences are of the order of the spread and exact for constant columns)
//
// Flat streaming form (the one that runs for every practical shape): G = 4 / gcd(4, cols) consecutive rows
// form a group of U = G cols / 4 units of 4 floats (G = 1 when cols % 4 == 0; odd widths such as Spiral's
// 5 x 119 = 595 take G = 4) and the CTA has R * U threads, so thread t reads units t, t + R U, t + 2 R U, ...
// of its slab — consecutive threads read consecutive 16-byte units (rows are contiguous), and the unit's four
// columns ((4u + k) mod cols) never change: 4 running sums and 4 sums of squares live in registers, 8 loads in
// flight per thread.  (W = 1: the scalar form for unaligned base pointers.)  The R row-groups are combined through shared memory, then
// one fp64 atomic pair per column and CTA.
// What the last CTA of a moments launch does with the finished sums (saves one or two tiny launches per update):
// mode 1: [mean | var | count] of the batch -> out (what ranks exchange); mode 2: fold the batch straight into the
// running statistics (RunningMeanStd.update_from_moments, normalization.py:44-58) and emit the float vectors.
struct Tail {
  int mode;
  unsigned* ticket;
  double* out;
  double *mean, *var, *count;
  double eps;
  float *mean_f, *rstd_f;
};

__device__ __forceinline__ void fold_column(int c, double bm, double bv, double batch_count, double cnt, double tot,
                                            double* mean, double* var, double eps, float* mean_f, float* rstd_f) {
  const double delta = bm - mean[c];
  const double new_mean = mean[c] + delta * batch_count / tot;
  const double m2 = var[c] * cnt + bv * batch_count + delta * delta * cnt * batch_count / tot;
  const double new_var = m2 / tot;
  mean[c] = new_mean;
  var[c] = new_var;
  mean_f[c] = (float)new_mean;
  rstd_f[c] = (float)(1.0 / sqrt(new_var + eps));
}

template <int W>
__global__ void __launch_bounds__(1024)
moments_flat_kernel(const float* __restrict__ x, long long rows, int cols, int G, int R, double* __restrict__ acc,
                    Tail tail) {
  extern __shared__ float sm[];   // [2][R][G * cols]
  const int GC = G * cols;        // floats per group of G rows (a multiple of W)
  const int U = GC / W;           // units per group
  const int t = threadIdx.x;
  const int u = t % U, rg = t / U;
  const long long groups = rows / G;
  const long long per = (groups + gridDim.x - 1) / gridDim.x;
  const long long g0 = (long long)blockIdx.x * per;
  const long long g1 = g0 + per < groups ? g0 + per : groups;
  float sh[W], s[W], q[W];
#pragma unroll
  for (int k = 0; k < W; ++k) { sh[k] = __ldg(x + (u * W + k) % cols); s[k] = 0.f; q[k] = 0.f; }
  if (g0 < g1) {
    const size_t n_units = (size_t)(g1 - g0) * U;
    const size_t stride = (size_t)R * U;
    const float* base = x + (size_t)g0 * GC;
    size_t i = t;
    constexpr int UNR = 8;
    for (; i + (UNR - 1) * stride < n_units; i += UNR * stride) {
      float v[UNR][W];
#pragma unroll
      for (int j = 0; j < UNR; ++j) {
        if constexpr (W == 4) {
          const float4 f = __ldcs(reinterpret_cast<const float4*>(base) + i + j * stride);
          v[j][0] = f.x; v[j][1] = f.y; v[j][2] = f.z; v[j][3] = f.w;
        } else {
          v[j][0] = __ldcs(base + i + j * stride);
        }
      }
#pragma unroll
      for (int j = 0; j < UNR; ++j)
#pragma unroll
        for (int k = 0; k < W; ++k) { const float d = v[j][k] - sh[k]; s[k] += d; q[k] = fmaf(d, d, q[k]); }
    }
    for (; i < n_units; i += stride) {
      if constexpr (W == 4) {
        const float4 f = __ldcs(reinterpret_cast<const float4*>(base) + i);
        const float v[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) { const float d = v[k] - sh[k]; s[k] += d; q[k] = fmaf(d, d, q[k]); }
      } else {
        const float d = __ldcs(base + i) - sh[0];
        s[0] += d; q[0] = fmaf(d, d, q[0]);
      }
    }
  }
  float* ss = sm;
  float* sq = sm + (size_t)R * GC;
#pragma unroll
  for (int k = 0; k < W; ++k) { ss[(size_t)rg * GC + u * W + k] = s[k]; sq[(size_t)rg * GC + u * W + k] = q[k]; }
  __syncthreads();
  if (g0 < g1 || blockIdx.x == 0) {
    for (int c = t; c < cols; c += blockDim.x) {
      double a = 0.0, b = 0.0;
      for (int e = c; e < R * GC; e += cols) { a += (double)ss[e]; b += (double)sq[e]; }   // row-groups x rows of a group
      if (blockIdx.x == 0) {          // the rows % G rows that do not fill a group
        const float sh0 = __ldg(x + c);
        for (long long r = groups * G; r < rows; ++r) {
          const double d = (double)(x[(size_t)r * cols + c] - sh0);
          a += d; b += d * d;
        }
      }
      atomicAdd(acc + c, a);
      atomicAdd(acc + cols + c, b);
    }
  }
  if (tail.mode == 0) return;
  // last CTA out finishes the update
  __shared__ int is_last;
  __threadfence();
  __syncthreads();
  if (t == 0) is_last = atomicAdd(tail.ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const double n = (double)rows;
  const double cnt = tail.mode == 2 ? tail.count[0] : 0.0;
  const double tot = cnt + n;
  for (int c = t; c < cols; c += blockDim.x) {
    const double sa = __ldcg(acc + c) / n, qa = __ldcg(acc + cols + c) / n;
    acc[c] = 0.0;
    acc[cols + c] = 0.0;
    const double bm = (double)x[c] + sa;      // the shift was the batch's first row
    double bv = qa - sa * sa;                 // np.var: population variance
    bv = bv > 0.0 ? bv : 0.0;
    if (tail.mode == 1) { tail.out[c] = bm; tail.out[cols + c] = bv; }
    else fold_column(c, bm, bv, n, cnt, tot, tail.mean, tail.var, tail.eps, tail.mean_f, tail.rstd_f);
  }
  __syncthreads();
  if (t == 0) {
    *tail.ticket = 0u;
    if (tail.mode == 1) tail.out[2 * cols] = n; else tail.count[0] = tot;
  }
}

// Fallback for rows wider than one CTA (cols / W > 1024): one thread per column, strided rows.
__global__ void __launch_bounds__(kMomThreads)
moments_kernel(const float* __restrict__ x, long long rows, int cols, double* __restrict__ acc) {
  const long long per = (rows + gridDim.x - 1) / gridDim.x;
  const long long r0 = (long long)blockIdx.x * per;
  const long long r1 = r0 + per < rows ? r0 + per : rows;
  for (int c = threadIdx.x; c < cols; c += kMomThreads) {
    const float sh = __ldg(x + c);
    float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};   // four independent chains per thread
    long long r = r0;
    for (; r + 3 < r1; r += 4) {
      float a[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] = x[(size_t)(r + u) * cols + c];
#pragma unroll
      for (int u = 0; u < 4; ++u) { const float d = a[u] - sh; s[u] += d; q[u] = fmaf(d, d, q[u]); }
    }
    for (; r < r1; ++r) { const float d = x[(size_t)r * cols + c] - sh; s[0] += d; q[0] = fmaf(d, d, q[0]); }
    const float s0 = s[0] + s[2], s1 = s[1] + s[3], q0 = q[0] + q[2], q1 = q[1] + q[3];
    if (r1 > r0) {
      atomicAdd(acc + c, (double)s0 + (double)s1);
      atomicAdd(acc + cols + c, (double)q0 + (double)q1);
    }
  }
}

// batch mean / population variance / row count from the shifted sums; clears the accumulators.
// out: [mean(cols) | var(cols) | count]
__global__ void finalize_kernel(const float* __restrict__ x, double* __restrict__ acc, double batch_count, int cols,
                                double* __restrict__ out) {
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    const double s = acc[c] / batch_count, q = acc[cols + c] / batch_count;
    const double bv = q - s * s;              // np.var: population variance
    out[c] = (double)x[c] + s;                // the shift was the batch's first row
    out[cols + c] = bv > 0.0 ? bv : 0.0;
    acc[c] = 0.0;
    acc[cols + c] = 0.0;
  }
  if (threadIdx.x == 0) out[2 * cols] = batch_count;
}

// `parts` batch moments ([mean | var | count] each, e.g. one per rank) are first combined, in order,
// into the moments of their union (Chan's parallel-variance formula = what np.mean / np.var of the
// concatenated batch give), then folded into the running statistics exactly like
// RunningMeanStd.update_from_moments (normalization.py:44-58).  All fp64.
__global__ void merge_kernel(const double* __restrict__ parts_buf, int parts, double* __restrict__ mean,
                             double* __restrict__ var, double* __restrict__ count, int cols, double eps,
                             float* __restrict__ mean_f, float* __restrict__ rstd_f) {
  const int stride = 2 * cols + 1;
  const double cnt = count[0];
  double batch_count = 0.0;
  for (int p = 0; p < parts; ++p) batch_count += parts_buf[(size_t)p * stride + 2 * cols];
  const double tot = cnt + batch_count;
  if (batch_count > 0.0) {
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
      double bn = 0.0, bm = 0.0, bm2 = 0.0;
      for (int p = 0; p < parts; ++p) {
        const double* P = parts_buf + (size_t)p * stride;
        const double n = P[2 * cols];
        if (n <= 0.0) continue;
        const double d = P[c] - bm, t = bn + n;
        bm2 += P[cols + c] * n + d * d * bn * n / t;
        bm += d * n / t;
        bn = t;
      }
      fold_column(c, bm, bm2 / bn, batch_count, cnt, tot, mean, var, eps, mean_f, rstd_f);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) count[0] = tot;
}

__global__ void export_kernel(const double* __restrict__ mean, const double* __restrict__ var, int cols, double eps,
                              float* __restrict__ mean_f, float* __restrict__ rstd_f) {
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    mean_f[c] = (float)mean[c];
    rstd_f[c] = (float)(1.0 / sqrt(var[c] + eps));
  }
}

// y = clip((x - mean) * rstd, +-clip).  Same flat form as the moments: thread t of a grid of G * R * U threads
// touches units t, t + G R U, ... whose columns never change, so mean / rstd sit in registers.
template <int W>
__global__ void __launch_bounds__(1024)
normalize_flat_kernel(const float* __restrict__ x, float* __restrict__ y, long long rows, int cols, int G,
                      const float* __restrict__ mean_f, const float* __restrict__ rstd_f, float clip) {
  const int U = G * cols / W;
  const int u = threadIdx.x % U;
  float m[W], r[W];
#pragma unroll
  for (int k = 0; k < W; ++k) { const int c = (u * W + k) % cols; m[k] = mean_f[c]; r[k] = rstd_f[c]; }
  const size_t n = (size_t)rows * cols;
  const size_t n_units = n / W;
  const size_t stride = (size_t)gridDim.x * blockDim.x;      // a multiple of U: the unit's columns never change
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_units; i += stride) {
    if constexpr (W == 4) {
      float4 f = __ldcs(reinterpret_cast<const float4*>(x) + i);
      f.x = fminf(fmaxf((f.x - m[0]) * r[0], -clip), clip);
      f.y = fminf(fmaxf((f.y - m[1]) * r[1], -clip), clip);
      f.z = fminf(fmaxf((f.z - m[2]) * r[2], -clip), clip);
      f.w = fminf(fmaxf((f.w - m[3]) * r[3], -clip), clip);
      reinterpret_cast<float4*>(y)[i] = f;
    } else {
      y[i] = fminf(fmaxf((__ldcs(x + i) - m[0]) * r[0], -clip), clip);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n - n_units * W)) {   // the n % 4 floats after the last full unit
    const size_t i = n_units * W + threadIdx.x;
    const int c = (int)(i % cols);
    y[i] = fminf(fmaxf((x[i] - mean_f[c]) * rstd_f[c], -clip), clip);
  }
}

__global__ void normalize_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, int cols,
                                 const float* __restrict__ mean_f, const float* __restrict__ rstd_f, float clip) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long k = i; k < n; k += stride) {
    const int c = (int)(k % cols);
    float v = (x[k] - mean_f[c]) * rstd_f[c];
    v = v < -clip ? -clip : (v > clip ? clip : v);
    y[k] = v;
  }
}

// launch geometry of the flat kernels: W floats per unit, R row-groups per CTA; returns false when a row does
// not fit one CTA (fallback kernels)
bool flat_geometry(const void* p0, const void* p1, int cols, int* W, int* G, int* R) {
  const bool al = (((uintptr_t)p0 | (uintptr_t)p1) & 15) == 0;
  *W = al ? 4 : 1;
  *G = al ? (cols % 4 == 0 ? 1 : (cols % 2 == 0 ? 2 : 4)) : 1;
  int U = *G * cols / *W;
  if (U > 1024 && al) { *W = 1; *G = 1; U = cols; }
  if (U > 1024) return false;
  const int r = 1024 / U;
  *R = r < 1 ? 1 : r;
  return true;
}

}  // namespace

struct bd_rms {
  int device = 0, cols = 0, sm_count = 0;
  double eps = 1e-8;
  double *mean = nullptr, *var = nullptr, *count = nullptr, *acc = nullptr, *part = nullptr;
  unsigned* ticket = nullptr;
  float *mean_f = nullptr, *rstd_f = nullptr;
  int64_t launches = 0;
};

extern "C" {

const char* bd_rms_last_error(void) { return g_rms_err; }

int bd_rms_create(int cols, int device, double count0, double eps, bd_rms** out) {
  if (!out) return rfail(BD_EINVAL, "bd_rms_create: null out");
  *out = nullptr;
  if (cols < 1) return rfail(BD_EINVAL, "bd_rms_create: cols must be positive");
  int prev = -1;
  if (cudaGetDevice(&prev) != cudaSuccess || cudaSetDevice(device) != cudaSuccess)
    return rfail(BD_ECUDA, "bd_rms_create: cannot select device %d", device);
  bd_rms* r = new (std::nothrow) bd_rms();
  if (!r) return rfail(BD_ENOMEM, "bd_rms_create: out of host memory");
  r->device = device; r->cols = cols; r->eps = eps;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  r->sm_count = prop.multiProcessorCount;
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); if (e == cudaSuccess) e = cudaMemset(*p, 0, bytes); };
  alloc((void**)&r->mean, cols * 8); alloc((void**)&r->var, cols * 8); alloc((void**)&r->count, 8);
  alloc((void**)&r->acc, 2 * cols * 8); alloc((void**)&r->part, (2 * cols + 1) * 8); alloc((void**)&r->ticket, 4); alloc((void**)&r->mean_f, cols * 4); alloc((void**)&r->rstd_f, cols * 4);
  if (e == cudaSuccess) {   // RunningMeanStd.__init__ (:24-32): mean 0, var 1, count = epsilon
    double* ones = new (std::nothrow) double[cols];
    if (ones) {
      for (int i = 0; i < cols; ++i) ones[i] = 1.0;
      e = cudaMemcpy(r->var, ones, cols * 8, cudaMemcpyHostToDevice);
      delete[] ones;
    }
    if (e == cudaSuccess) e = cudaMemcpy(r->count, &count0, 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) { export_kernel<<<1, 256>>>(r->mean, r->var, cols, eps, r->mean_f, r->rstd_f); e = cudaDeviceSynchronize(); }
  }
  if (prev >= 0 && prev != device) cudaSetDevice(prev);
  if (e != cudaSuccess) {
    cudaFree(r->mean); cudaFree(r->var); cudaFree(r->count); cudaFree(r->acc); cudaFree(r->part); cudaFree(r->ticket); cudaFree(r->mean_f); cudaFree(r->rstd_f);
    delete r;
    return rfail(BD_ECUDA, "bd_rms_create: %s", cudaGetErrorString(e));
  }
  *out = r;
  return BD_OK;
}

void bd_rms_destroy(bd_rms* r) {
  if (!r) return;
  cudaFree(r->mean); cudaFree(r->var); cudaFree(r->count); cudaFree(r->acc); cudaFree(r->part); cudaFree(r->ticket); cudaFree(r->mean_f); cudaFree(r->rstd_f);
  delete r;
}

// One pass over the batch; `mode` selects what the last CTA does with the sums (see Tail).  Returns false when the
// rows are too wide for the flat kernel (the caller then runs the fallback + separate finalize / merge launches).
static bool launch_moments_flat(bd_rms* r, const float* x_dev, int64_t rows, int mode, double* out, cudaStream_t st) {
  int W = 1, G = 1, R = 1;
  if (!flat_geometry(x_dev, x_dev, r->cols, &W, &G, &R)) return false;
  const int threads = R * (G * r->cols / W);
  // slabs of >= 8 unrolled iterations per thread; one CTA of ~1000 threads (or two of <= 512) per SM
  long long grid = (rows / G + (long long)R * 8 - 1) / ((long long)R * 8);
  const long long cap = (long long)r->sm_count * (threads > 512 ? 1 : 2);
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  const size_t smem = (size_t)2 * R * G * r->cols * sizeof(float);
  Tail tail{mode, r->ticket, out, r->mean, r->var, r->count, r->eps, r->mean_f, r->rstd_f};
  if (W == 4) moments_flat_kernel<4><<<(int)grid, threads, smem, st>>>(x_dev, rows, r->cols, G, R, r->acc, tail);
  else moments_flat_kernel<1><<<(int)grid, threads, smem, st>>>(x_dev, rows, r->cols, G, R, r->acc, tail);
  r->launches++;
  return true;
}

int bd_rms_batch_moments(bd_rms* r, const float* x_dev, int64_t rows, double* moments_dev, void* stream) {
  if (!r || !x_dev || !moments_dev) return rfail(BD_EINVAL, "bd_rms_batch_moments: null argument");
  if (rows <= 0) return rfail(BD_EINVAL, "bd_rms_batch_moments: rows must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  if (!launch_moments_flat(r, x_dev, rows, 1, moments_dev, st)) {
    long long grid = (long long)r->sm_count * 8;
    if (grid > rows) grid = rows;
    moments_kernel<<<(int)grid, kMomThreads, 0, st>>>(x_dev, rows, r->cols, r->acc);
    finalize_kernel<<<1, 256, 0, st>>>(x_dev, r->acc, (double)rows, r->cols, moments_dev);
    r->launches += 2;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return rfail(BD_ECUDA, "bd_rms_batch_moments: %s", cudaGetErrorString(e));
  return BD_OK;
}

int bd_rms_merge_moments(bd_rms* r, const double* moments_dev, int parts, void* stream) {
  if (!r || !moments_dev || parts < 1) return rfail(BD_EINVAL, "bd_rms_merge_moments: moments and parts >= 1 are required");
  merge_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(moments_dev, parts, r->mean, r->var, r->count, r->cols, r->eps,
                                                    r->mean_f, r->rstd_f);
  r->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return rfail(BD_ECUDA, "bd_rms_merge_moments: %s", cudaGetErrorString(e));
  return BD_OK;
}

int bd_rms_update(bd_rms* r, const float* x_dev, int64_t rows, void* stream) {
  if (!r || !x_dev) return rfail(BD_EINVAL, "bd_rms_update: null argument");
  if (rows <= 0) return BD_OK;
  if (launch_moments_flat(r, x_dev, rows, 2, nullptr, (cudaStream_t)stream)) {   // ONE launch
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return rfail(BD_ECUDA, "bd_rms_update: %s", cudaGetErrorString(e));
    return BD_OK;
  }
  int rc = bd_rms_batch_moments(r, x_dev, rows, r->part, stream);
  if (rc != BD_OK) return rc;
  return bd_rms_merge_moments(r, r->part, 1, stream);
}

int bd_rms_normalize(bd_rms* r, const float* x_dev, float* y_dev, int64_t rows, float clip, void* stream) {
  if (!r || !x_dev || !y_dev) return rfail(BD_EINVAL, "bd_rms_normalize: null argument");
  if (rows <= 0) return BD_OK;
  const long long n = (long long)rows * r->cols;
  int W = 1, G = 1, R = 1;
  if (flat_geometry(x_dev, y_dev, r->cols, &W, &G, &R)) {
    const int threads = R * (G * r->cols / W);
    long long grid = (n / W + threads - 1) / threads;
    const long long cap = (long long)r->sm_count * (threads > 512 ? 2 : 4) * 4;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    if (W == 4) normalize_flat_kernel<4><<<(int)grid, threads, 0, (cudaStream_t)stream>>>(x_dev, y_dev, rows, r->cols, G, r->mean_f, r->rstd_f, clip);
    else normalize_flat_kernel<1><<<(int)grid, threads, 0, (cudaStream_t)stream>>>(x_dev, y_dev, rows, r->cols, G, r->mean_f, r->rstd_f, clip);
  } else {
    long long grid = (n + 255) / 256;
    if (grid > (long long)r->sm_count * 32) grid = (long long)r->sm_count * 32;
    normalize_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(x_dev, y_dev, n, r->cols, r->mean_f, r->rstd_f, clip);
  }
  r->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return rfail(BD_ECUDA, "bd_rms_normalize: %s", cudaGetErrorString(e));
  return BD_OK;
}

int bd_rms_get(bd_rms* r, double* mean_dev, double* var_dev, double* count_dev, float* mean_f_dev, float* rstd_f_dev, void* stream) {
  if (!r) return rfail(BD_EINVAL, "bd_rms_get: null handle");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaSuccess;
  auto cp = [&](void* d, const void* s, size_t b) { if (d && e == cudaSuccess) e = cudaMemcpyAsync(d, s, b, cudaMemcpyDeviceToDevice, st); };
  cp(mean_dev, r->mean, r->cols * 8); cp(var_dev, r->var, r->cols * 8); cp(count_dev, r->count, 8);
  cp(mean_f_dev, r->mean_f, r->cols * 4); cp(rstd_f_dev, r->rstd_f, r->cols * 4);
  if (e != cudaSuccess) return rfail(BD_ECUDA, "bd_rms_get: %s", cudaGetErrorString(e));
  return BD_OK;
}

int bd_rms_set(bd_rms* r, const double* mean_dev, const double* var_dev, const double* count_dev, void* stream) {
  if (!r) return rfail(BD_EINVAL, "bd_rms_set: null handle");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaSuccess;
  auto cp = [&](void* d, const void* s, size_t b) { if (s && e == cudaSuccess) e = cudaMemcpyAsync(d, s, b, cudaMemcpyDeviceToDevice, st); };
  cp(r->mean, mean_dev, r->cols * 8); cp(r->var, var_dev, r->cols * 8); cp(r->count, count_dev, 8);
  if (e == cudaSuccess) { export_kernel<<<1, 256, 0, st>>>(r->mean, r->var, r->cols, r->eps, r->mean_f, r->rstd_f); e = cudaGetLastError(); r->launches++; }
  if (e != cudaSuccess) return rfail(BD_ECUDA, "bd_rms_set: %s", cudaGetErrorString(e));
  return BD_OK;
}

int64_t bd_rms_launch_count(const bd_rms* r) { return r ? r->launches : 0; }

}  // extern "C"
