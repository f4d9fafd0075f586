// Running mean / variance of observation batches on the device.
//
// Replaces `RunningMeanStd.update` + `MeanStdNormalizer.__call__` of the reference trainer
// (gym_pybullet_drones/safe_control_gym/math_and_models/normalization.py:13-96), which MAPPO
// applies to every observation batch of the rollout (`mappo/mappo.py:132,165,804`): batch moments
// over the env axis, parallel-variance merge into the running statistics, then
// clip((x - mean) / sqrt(var + eps), +-clip).
//
// Three kernels per update: `moments_kernel` reads the (rows x cols) batch once (HBM-bound; per-CTA
// row slabs, coalesced along the columns, one fp64 atomic per column and CTA), `finalize_kernel`
// turns the sums into [mean | var | count] (the unit ranks exchange: one small all-gather when the
// envs are sharded over GPUs), and `merge_kernel` (one CTA) folds one or several such parts into
// mean / var / count in fp64 and emits float mean and 1/sqrt(var + eps) vectors for fused consumers: the actor kernel normalises its input tile while it
// converts it to bf16 (bd_actor_set_input_norm), so the normalised observations are never written
// to HBM during the rollout.  `normalize_kernel` is the standalone form (evaluation, critic input).
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
      const double bv = bm2 / bn;
      const double delta = bm - mean[c];
      const double new_mean = mean[c] + delta * batch_count / tot;
      const double m2 = var[c] * cnt + bv * batch_count + delta * delta * cnt * batch_count / tot;
      const double new_var = m2 / tot;
      mean[c] = new_mean;
      var[c] = new_var;
      mean_f[c] = (float)new_mean;
      rstd_f[c] = (float)(1.0 / sqrt(new_var + eps));
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

// y = clip((x - mean) * rstd, +-clip); rows x cols, 4 columns per thread when cols % 4 == 0
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

}  // namespace

struct bd_rms {
  int device = 0, cols = 0, sm_count = 0;
  double eps = 1e-8;
  double *mean = nullptr, *var = nullptr, *count = nullptr, *acc = nullptr, *part = nullptr;
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
  alloc((void**)&r->acc, 2 * cols * 8); alloc((void**)&r->part, (2 * cols + 1) * 8); alloc((void**)&r->mean_f, cols * 4); alloc((void**)&r->rstd_f, cols * 4);
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
    cudaFree(r->mean); cudaFree(r->var); cudaFree(r->count); cudaFree(r->acc); cudaFree(r->part); cudaFree(r->mean_f); cudaFree(r->rstd_f);
    delete r;
    return rfail(BD_ECUDA, "bd_rms_create: %s", cudaGetErrorString(e));
  }
  *out = r;
  return BD_OK;
}

void bd_rms_destroy(bd_rms* r) {
  if (!r) return;
  cudaFree(r->mean); cudaFree(r->var); cudaFree(r->count); cudaFree(r->acc); cudaFree(r->part); cudaFree(r->mean_f); cudaFree(r->rstd_f);
  delete r;
}

int bd_rms_batch_moments(bd_rms* r, const float* x_dev, int64_t rows, double* moments_dev, void* stream) {
  if (!r || !x_dev || !moments_dev) return rfail(BD_EINVAL, "bd_rms_batch_moments: null argument");
  if (rows <= 0) return rfail(BD_EINVAL, "bd_rms_batch_moments: rows must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  long long grid = (long long)r->sm_count * 8;
  if (grid > rows) grid = rows;
  moments_kernel<<<(int)grid, kMomThreads, 0, st>>>(x_dev, rows, r->cols, r->acc);
  finalize_kernel<<<1, 256, 0, st>>>(x_dev, r->acc, (double)rows, r->cols, moments_dev);
  r->launches += 2;
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
  int rc = bd_rms_batch_moments(r, x_dev, rows, r->part, stream);
  if (rc != BD_OK) return rc;
  return bd_rms_merge_moments(r, r->part, 1, stream);
}

int bd_rms_normalize(bd_rms* r, const float* x_dev, float* y_dev, int64_t rows, float clip, void* stream) {
  if (!r || !x_dev || !y_dev) return rfail(BD_EINVAL, "bd_rms_normalize: null argument");
  if (rows <= 0) return BD_OK;
  const long long n = (long long)rows * r->cols;
  long long grid = (n + 255) / 256;
  if (grid > (long long)r->sm_count * 32) grid = (long long)r->sm_count * 32;
  normalize_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(x_dev, y_dev, n, r->cols, r->mean_f, r->rstd_f, clip);
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
