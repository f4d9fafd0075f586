// Batched drone-step kernels for sm_100a.
//
// One thread per drone, one CTA = E = 128/M whole environments, so the O(M^2)
// downwash term and the per-env reward / termination reductions never leave the
// CTA.  A control step is ONE launch of step_kernel:
//
//   action -> rpm (BaseRLAviary.py:191-192,224-225)
//   S substeps of explicit dynamics with the state in registers
//       (BaseAviary.py:343-374, _dynamics :815-877, _integrateQ :879-892,
//        Bullet quaternion round trip of _updateAndStoreKinematicInformation :509-519)
//   KIN observation (BaseRLAviary.py:307-319, SpiralAviary.py:120-146)
//   reward / terminated / truncated of the task
//       (HoverAviary.py:77-117, MultiHoverAviary.py:128-268, SpiralAviary.py:150-196)
//   step-counter advance (:382) and SubprocVecEnv's reset-on-done
//       (subproc_vec_env.py:195-206; MultiHoverAviary.py:75-110 jitter)
//
// HBM traffic per drone-step is the algorithmic minimum + one ring slot:
// 4 x 128-bit state loads/stores (SoA planes), the action, B-1 history slots
// (cp.async straight into the shared-memory observation tile), and the
// observation rows, which leave the CTA as one contiguous, fully coalesced block.
//
// Two arithmetic flavours are compiled from the same source:
//   exact  (double, or float with GENERIC): the reference's formulas verbatim
//          (s = 2/|q|^2 rotation, Bullet matrix->quaternion branches, sqrt/sin/cos
//          quaternion step with the 1e-8 early-out) — the fp64 parity mode.
//   fast   (float, !GENERIC): same maths re-associated for throughput: rotation
//          column only, per-step thrust/torque hoisting, polynomial sinc/cos for
//          the quaternion step, rsqrt normalisation with Bullet's sign rule.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "bd_params.h"

namespace bd {

enum { TASK_HOVER = 0, TASK_MULTIHOVER = 1, TASK_SPIRAL = 2 };
enum { MODEL_CF2X = 0, MODEL_CF2P = 1, MODEL_RACE = 2 };
enum { AERO_GND = 1, AERO_DRAG = 2, AERO_DW = 4 };
enum { RESET_FIXED = 0, RESET_PHILOX = 1, RESET_BUFFER = 2 };

// ---------------------------------------------------------------- small maths
__device__ __forceinline__ float  sqrt_(float x)  { return sqrtf(x); }
__device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
__device__ __forceinline__ float  exp_(float x)  { return expf(x); }
__device__ __forceinline__ double exp_(double x) { return exp(x); }
__device__ __forceinline__ float  atan2_(float y, float x)  { return atan2f(y, x); }
__device__ __forceinline__ double atan2_(double y, double x) { return atan2(y, x); }
__device__ __forceinline__ float  asin_(float x)  { return asinf(x); }
__device__ __forceinline__ double asin_(double x) { return asin(x); }
__device__ __forceinline__ void sincos_(float x, float* s, float* c)  { sincosf(x, s, c); }
__device__ __forceinline__ void sincos_(double x, double* s, double* c) { sincos(x, s, c); }
__device__ __forceinline__ float  abs_(float x)  { return fabsf(x); }
__device__ __forceinline__ double abs_(double x) { return fabs(x); }

__device__ __forceinline__ float4  make4(float a, float b, float c, float d)   { return make_float4(a, b, c, d); }
__device__ __forceinline__ double4 make4(double a, double b, double c, double d) { return make_double4(a, b, c, d); }

template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem_dst, const void* gmem_src) {
  unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  if constexpr (BYTES == 16) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
  } else {
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;\n" ::"r"(d), "l"(gmem_src), "n"(BYTES) : "memory");
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// ------------------------------------------------------------------- Philox
__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
__device__ __forceinline__ float  u01(uint32_t a, uint32_t, float)  { return (a >> 8) * (1.0f / 16777216.0f); }
__device__ __forceinline__ double u01(uint32_t a, uint32_t b, double) {
  return ((a >> 5) * 67108864.0 + (b >> 6)) * (1.0 / 9007199254740992.0);
}

// ------------------------------------------------- Bullet closed-form helpers
// btMatrix3x3::setRotation == pybullet getMatrixFromQuaternion (BaseAviary.py:836)
template <typename R>
__device__ __forceinline__ void quat_to_mat(R x, R y, R z, R w, R m[9]) {
  const R d = x * x + y * y + z * z + w * w;
  const R s = R(2) / d;
  const R xs = x * s, ys = y * s, zs = z * s;
  const R wx = w * xs, wy = w * ys, wz = w * zs;
  const R xx = x * xs, xy = x * ys, xz = x * zs;
  const R yy = y * ys, yz = y * zs, zz = z * zs;
  m[0] = R(1) - (yy + zz); m[1] = xy - wz;          m[2] = xz + wy;
  m[3] = xy + wz;          m[4] = R(1) - (xx + zz); m[5] = yz - wx;
  m[6] = xz - wy;          m[7] = yz + wx;          m[8] = R(1) - (xx + yy);
}

// btMatrix3x3::getRotation (what getBasePositionAndOrientation returns, :517)
template <typename R>
__device__ __forceinline__ void mat_to_quat(const R m[9], R& x, R& y, R& z, R& w) {
  const R trace = m[0] + m[4] + m[8];
  if (trace > R(0)) {
    R s = sqrt_(trace + R(1));
    w = s * R(0.5);
    s = R(0.5) / s;
    x = (m[7] - m[5]) * s;
    y = (m[2] - m[6]) * s;
    z = (m[3] - m[1]) * s;
  } else {
    const int i = m[0] < m[4] ? (m[4] < m[8] ? 2 : 1) : (m[0] < m[8] ? 2 : 0);
    if (i == 0) {
      R s = sqrt_(m[0] - m[4] - m[8] + R(1));
      x = s * R(0.5); s = R(0.5) / s;
      w = (m[7] - m[5]) * s; y = (m[3] + m[1]) * s; z = (m[6] + m[2]) * s;
    } else if (i == 1) {
      R s = sqrt_(m[4] - m[8] - m[0] + R(1));
      y = s * R(0.5); s = R(0.5) / s;
      w = (m[2] - m[6]) * s; z = (m[7] + m[5]) * s; x = (m[1] + m[3]) * s;
    } else {
      R s = sqrt_(m[8] - m[0] - m[4] + R(1));
      z = s * R(0.5); s = R(0.5) / s;
      w = (m[3] - m[1]) * s; x = (m[2] + m[6]) * s; y = (m[5] + m[7]) * s;
    }
  }
}

template <typename R>
__device__ __forceinline__ void bullet_roundtrip(R& x, R& y, R& z, R& w) {
  R m[9];
  quat_to_mat(x, y, z, w, m);
  mat_to_quat(m, x, y, z, w);
}

// Same map as bullet_roundtrip for float throughput: q/|q| with Bullet's sign rule
// (trace > 0 <=> 4w^2 > 1 -> w >= 0; else the largest diagonal's component >= 0).
__device__ __forceinline__ void fast_canonical(float& x, float& y, float& z, float& w) {
  const float rn = rsqrtf(fmaf(x, x, fmaf(y, y, fmaf(z, z, w * w))));
  x *= rn; y *= rn; z *= rn; w *= rn;
  bool neg;
  if (w * w > 0.25f) {
    neg = w < 0.0f;
  } else {
    const float xx = x * x, yy = y * y, zz = z * z;   // m00 < m11 <=> xx < yy
    const int i = xx < yy ? (yy < zz ? 2 : 1) : (xx < zz ? 2 : 0);
    neg = (i == 0 ? x : (i == 1 ? y : z)) < 0.0f;
  }
  if (neg) { x = -x; y = -y; z = -z; w = -w; }
}

// pybullet getEulerFromQuaternion (:518) incl. its gimbal branches
template <typename R>
__device__ __forceinline__ void quat_to_euler(R x, R y, R z, R w, R& roll, R& pitch, R& yaw) {
  const R sqx = x * x, sqy = y * y, sqz = z * z, squ = w * w;
  const R sarg = R(-2) * (x * z - w * y);
  const R half_pi = R(0.5 * 3.14159265358979323846);
  if (sarg <= R(-0.99999)) {
    roll = R(0); pitch = -half_pi; yaw = R(2) * atan2_(x, -y);
  } else if (sarg >= R(0.99999)) {
    roll = R(0); pitch = half_pi; yaw = R(2) * atan2_(-x, y);
  } else {
    roll = atan2_(R(2) * (y * z + w * x), squ - sqx - sqy + sqz);
    pitch = asin_(sarg);
    yaw = atan2_(R(2) * (x * y + w * z), squ + sqx - sqy - sqz);
  }
}

// pybullet getQuaternionFromEuler (:488), normalised
template <typename R>
__device__ __forceinline__ void euler_to_quat(R roll, R pitch, R yaw, R& x, R& y, R& z, R& w) {
  R sp, cp, st, ct, ss, cs;
  sincos_(roll / R(2), &sp, &cp);
  sincos_(pitch / R(2), &st, &ct);
  sincos_(yaw / R(2), &ss, &cs);
  x = sp * ct * cs - cp * st * ss;
  y = cp * st * cs + sp * ct * ss;
  z = cp * ct * ss - sp * st * cs;
  w = cp * ct * cs + sp * st * ss;
  const R n = sqrt_(x * x + y * y + z * z + w * w);
  x /= n; y /= n; z /= n; w /= n;
}

// BaseAviary._integrateQ (:879-892): q <- cos(th) q + sin(th)/|w| Lambda(w) q
template <typename R>
__device__ __forceinline__ void integrate_q_exact(R& x, R& y, R& z, R& w, R p, R q, R r, R dt) {
  const R n = sqrt_(p * p + q * q + r * r);
  if (n <= R(1e-8)) return;                       // np.isclose(norm, 0): atol 1e-8
  R s, c;
  sincos_(n * dt / R(2), &s, &c);
  const R k = R(2) / n * R(0.5) * s;
  const R nx = c * x + k * (r * y - q * z + p * w);
  const R ny = c * y + k * (-r * x + p * z + q * w);
  const R nz = c * z + k * (q * x - p * y + r * w);
  const R nw = c * w + k * (-p * x - q * y - r * z);
  x = nx; y = ny; z = nz; w = nw;
}

// float fast path: cos(th) and sin(th)/|w| = (dt/2) sinc(th) as polynomials in th^2
__device__ __forceinline__ void integrate_q_fast(float& x, float& y, float& z, float& w,
                                                 float p, float q, float r, float half_dt) {
  const float n2 = fmaf(p, p, fmaf(q, q, r * r));
  const float u = n2 * half_dt * half_dt;         // th^2
  float c, k;
  if (u <= 0.25f) {
    c = fmaf(u, fmaf(u, fmaf(u, fmaf(u, 2.4801587e-5f, -1.3888889e-3f), 4.1666668e-2f), -0.5f), 1.0f);
    k = half_dt * fmaf(u, fmaf(u, fmaf(u, fmaf(u, 2.7557319e-6f, -1.9841270e-4f), 8.3333338e-3f),
                                    -1.6666667e-1f), 1.0f);
  } else {
    const float n = sqrtf(n2);
    float s;
    sincosf(n * half_dt, &s, &c);
    k = s / n;
  }
  const float nx = fmaf(k, fmaf(r, y, fmaf(-q, z, p * w)), c * x);
  const float ny = fmaf(k, fmaf(-r, x, fmaf(p, z, q * w)), c * y);
  const float nz = fmaf(k, fmaf(q, x, fmaf(-p, y, r * w)), c * z);
  const float nw = fmaf(k, -fmaf(p, x, fmaf(q, y, r * z)), c * w);
  x = nx; y = ny; z = nz; w = nw;
}

// --------------------------------------------------------------- reset logic
template <typename R>
struct Drone {
  R px, py, pz, qx, qy, qz, qw, vx, vy, vz, wx, wy, wz, tx, ty, tz;
};

// INIT pose of (env, drone) -> freshly reset drone (BaseAviary._housekeeping :451-505
// + kinematic refresh :509-519) and its 12 kinematic observation entries.
template <typename R, int TASK>
__device__ __forceinline__ void reset_drone(const Params<R>& P, int env, int drone, const R* cand,
                                            Drone<R>& d, float kin[12]) {
  const long long ib = (long long)env * P.init_env_stride + drone * 3;
  R ix = P.init_xyz[ib], iy = P.init_xyz[ib + 1], iz = P.init_xyz[ib + 2];
  if (TASK == TASK_MULTIHOVER && cand != nullptr) { ix = cand[0]; iy = cand[1]; iz = cand[2]; }
  d.px = ix; d.py = iy; d.pz = iz;
  euler_to_quat(P.init_rpy[ib], P.init_rpy[ib + 1], P.init_rpy[ib + 2], d.qx, d.qy, d.qz, d.qw);
  bullet_roundtrip(d.qx, d.qy, d.qz, d.qw);
  d.vx = d.vy = d.vz = R(0);
  d.wx = d.wy = d.wz = R(0);
  if (TASK == TASK_HOVER) { d.tx = R(0); d.ty = R(0); d.tz = R(1); }                 // HoverAviary.py:51
  else if (TASK == TASK_MULTIHOVER) { d.tx = ix; d.ty = iy; d.tz = iz + R(1) / R(drone + 1); }  // :72,106
  else { d.tx = d.ty = d.tz = R(0); }
  R roll, pitch, yaw;
  quat_to_euler(d.qx, d.qy, d.qz, d.qw, roll, pitch, yaw);
  kin[0] = (float)d.px; kin[1] = (float)d.py; kin[2] = (float)d.pz;
  kin[3] = (float)roll; kin[4] = (float)pitch; kin[5] = (float)yaw;
  kin[6] = kin[7] = kin[8] = kin[9] = kin[10] = kin[11] = 0.0f;
}

// MultiHoverAviary.reset (:83-102) for one env, run by its leader thread:
// ORIGINAL_INIT_XYZS + U(-0.25,0.25)^3, z clipped to [0.1,1], redraw until every
// pair is >= 0.5 m apart.  `cand` = shared-memory rows of this env's M drones.
// The reference loops forever when no draw can succeed; here kMaxJitterTries
// draws, then the un-jittered layout (documented deviation).
template <typename R>
__device__ void sample_jitter(const Params<R>& P, int env, int total_steps, int epoch, R* cand) {
  const int M = P.M;
  const long long ib = (long long)env * P.init_env_stride;
  for (int attempt = 0; attempt <= kMaxJitterTries; ++attempt) {
    const bool last = attempt == kMaxJitterTries;
    for (int i = 0; i < M; ++i) {
      R j[3] = {R(0), R(0), R(0)};
      if (!last) {
        if (P.reset_mode == RESET_BUFFER && P.jitter != nullptr) {
          const long long jb = ((long long)env * M + i) * 3;
          j[0] = P.jitter[jb]; j[1] = P.jitter[jb + 1]; j[2] = P.jitter[jb + 2];
        } else {
          uint32_t c[4] = {(uint32_t)env, (uint32_t)total_steps, (uint32_t)(attempt * M + i), (uint32_t)epoch << 1};
          uint32_t c2[4] = {c[0], c[1], c[2], c[3] | 1u};
          philox4x32_10(c, (uint32_t)P.seed, (uint32_t)(P.seed >> 32));
          philox4x32_10(c2, (uint32_t)P.seed, (uint32_t)(P.seed >> 32));
          j[0] = R(-0.25) + R(0.5) * u01(c[0], c[1], R(0));
          j[1] = R(-0.25) + R(0.5) * u01(c[2], c[3], R(0));
          j[2] = R(-0.25) + R(0.5) * u01(c2[0], c2[1], R(0));
        }
      }
      R z = P.init_xyz[ib + i * 3 + 2] + j[2];
      z = z < R(0.1) ? R(0.1) : (z > R(1.0) ? R(1.0) : z);
      cand[i * 3 + 0] = P.init_xyz[ib + i * 3 + 0] + j[0];
      cand[i * 3 + 1] = P.init_xyz[ib + i * 3 + 1] + j[1];
      cand[i * 3 + 2] = z;
    }
    if (last || P.reset_mode == RESET_BUFFER) return;
    bool ok = true;
    for (int a = 0; a < M && ok; ++a)
      for (int b = a + 1; b < M; ++b) {
        const R dx = cand[a * 3] - cand[b * 3], dy = cand[a * 3 + 1] - cand[b * 3 + 1],
                dz = cand[a * 3 + 2] - cand[b * 3 + 2];
        if (sqrt_(dx * dx + dy * dy + dz * dz) < R(0.5)) { ok = false; break; }
      }
    if (ok) return;
  }
}

// --------------------------------------------------------------- task maths
template <typename R>
__device__ __forceinline__ void spiral_reference(const Params<R>& P, int step_counter, int drone,
                                                 R ref_p[3], R ref_v[3], R& sphi, R& cphi) {
  const R t = (R)((double)step_counter / P.pyb_freq);                            // SpiralAviary.py:84
  const R phase = P.sp_omega * t + R(2) * R(3.14159265358979323846) * R(drone) / R(P.M);
  sincos_(phase, &sphi, &cphi);
  ref_p[0] = P.sp_cx + P.sp_R * cphi;
  ref_p[1] = P.sp_cy + P.sp_R * sphi;
  ref_p[2] = R(0.3) + P.sp_vz * t;
  ref_v[0] = -P.sp_R * P.sp_omega * sphi;
  ref_v[1] = P.sp_R * P.sp_omega * cphi;
  ref_v[2] = P.sp_vz;
}

template <typename R, bool GENERIC> struct Bounds { static constexpr int kMinBlocks = 1; };
template <> struct Bounds<float, false> { static constexpr int kMinBlocks = 6; };
template <> struct Bounds<float, true> { static constexpr int kMinBlocks = 3; };

// =========================================================================
//                                step kernel
// =========================================================================
template <typename R, int TASK, int A, bool GENERIC>
__global__ void __launch_bounds__(kBlock, Bounds<R, GENERIC>::kMinBlocks)
step_kernel(const __grid_constant__ Params<R> P) {
  using R4 = typename V4<R>::type;
  constexpr bool FAST = (sizeof(R) == 4) && !GENERIC;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sobs = reinterpret_cast<float*>(smem_raw);                       // [kBlock][Ds]
  R* sred = reinterpret_cast<R*>(smem_raw + (size_t)kBlock * P.Ds * 4);   // [kBlock]
  R* spos = sred + kBlock;                                                // [kBlock][3]
  int* sflag = reinterpret_cast<int*>(spos + 3 * kBlock);                 // [kBlock]
  int* sdone = sflag + kBlock;                                            // [kBlock]

  const int tid = threadIdx.x;
  const int M = P.M, E = P.E, B = P.B, Ds = P.Ds;
  const int env_l = tid / M;
  const int drone = tid - env_l * M;
  const int env = blockIdx.x * E + env_l;
  const bool active = (env_l < E) && (env < P.N);
  const long long g = (long long)env * M + drone;
  float* myrow = sobs + (size_t)tid * Ds;

  int2 ec = make_int2(0, 0);
  if (active) ec = P.envc[env];
  const int head = ec.y % B;   // ring slot overwritten by this step's action

  // ---- 1. action history -> observation tile (async, no registers) ----------
  if (active) {
    int slot = head;
#pragma unroll 1
    for (int j = 0; j < B - 1; ++j) {
      slot = (slot + 1 == B) ? 0 : slot + 1;
      cp_async<A * 4>(myrow + 12 + j * A, P.hist + ((size_t)slot * P.n_total + g) * A);
    }
  }
  cp_async_commit();

  // ---- 2. state + action -----------------------------------------------------
  Drone<R> d;
  R rpm[4];
  float onep[4] = {1.f, 1.f, 1.f, 1.f};   // fl32(1 + 0.05 a) per motor (float actions)
  R last_rpm[4] = {R(0), R(0), R(0), R(0)};
  if (active) {
    const R4 a0 = P.s0[g], a1 = P.s1[g], a2 = P.s2[g], a3 = P.s3[g];
    d.px = a0.x; d.py = a0.y; d.pz = a0.z; d.qx = a0.w;
    d.qy = a1.x; d.qz = a1.y; d.qw = a1.z; d.vx = a1.w;
    d.vy = a2.x; d.vz = a2.y; d.wx = a2.z; d.wy = a2.w;
    d.wz = a3.x; d.tx = a3.y; d.ty = a3.z; d.tz = a3.w;
    float af[4] = {0.f, 0.f, 0.f, 0.f};
    bool from_double = false;
    if constexpr (sizeof(R) == 8) from_double = !P.action_is_f32;
    if (from_double) {
      const double* ap = reinterpret_cast<const double*>(P.actions) + g * A;
#pragma unroll
      for (int k = 0; k < A; ++k) {
        const double a = ap[k];
        af[k] = (float)a;
        const R r = P.hover_rpm * (R(1) + R(0.05) * (R)a);                 // BaseRLAviary.py:192
        if constexpr (A == 4) rpm[k] = r; else rpm[0] = rpm[1] = rpm[2] = rpm[3] = r;
      }
    } else {
      const float* ap = reinterpret_cast<const float*>(P.actions) + g * A;
      if constexpr (A == 4) {
        const float4 v = *reinterpret_cast<const float4*>(ap);
        af[0] = v.x; af[1] = v.y; af[2] = v.z; af[3] = v.w;
      } else {
        af[0] = ap[0];
      }
#pragma unroll
      for (int k = 0; k < A; ++k) {
        // numpy evaluates 1 + 0.05*a in float32 for float32 actions (two roundings)
        const float s = __fadd_rn(1.0f, __fmul_rn(0.05f, af[k]));
        const R r = P.hover_rpm * (R)s;
        if constexpr (A == 4) { rpm[k] = r; onep[k] = s; }
        else { rpm[0] = rpm[1] = rpm[2] = rpm[3] = r; onep[0] = onep[1] = onep[2] = onep[3] = s; }
      }
    }
    // newest history entry: ring slot `head` and the tail of the observation row
    float* hp = P.hist + ((size_t)head * P.n_total + g) * A;
    if constexpr (A == 4) {
      const float4 v = make_float4(af[0], af[1], af[2], af[3]);
      *reinterpret_cast<float4*>(hp) = v;
      *reinterpret_cast<float4*>(myrow + 12 + (B - 1) * A) = v;
    } else {
      hp[0] = af[0];
      myrow[12 + (B - 1) * A] = af[0];
    }
    if (GENERIC && (P.aero & AERO_DRAG) && ec.x > 0) {
      // last_clipped_action (BaseAviary.py:372,468): previous step's rpm, zero after a reset
      const int prev = head == 0 ? B - 1 : head - 1;
      const float* lp = P.hist + ((size_t)prev * P.n_total + g) * A;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float a = lp[A == 4 ? k : 0];
        float s = __fadd_rn(1.0f, __fmul_rn(0.05f, a));
        last_rpm[k] = P.hover_rpm * (R)s;
      }
    }
  } else {
    d = Drone<R>{};
    d.qw = R(1);
    rpm[0] = rpm[1] = rpm[2] = rpm[3] = R(0);
  }

  // ---- 3. S substeps of explicit dynamics ------------------------------------
  const R dt = P.dt;
  R avx = R(0), avy = R(0), avz = R(0);   // world angular velocity R_old * w_new (:873)

  if constexpr (FAST) {
    // rpm_i = H (1 + s_i)  =>  rpm_i^2 = H^2 (1 + u_i),  u_i = s_i (2 + s_i).  With
    // 4 KF H^2 = m g (BaseAviary.py:118) the hover terms cancel analytically, which
    // removes the float32 cancellation in thrust - gravity and in the torque mixes:
    //   thrust/m = g (1 + e), e = sum(u)/4;  f_i = (m g / 4)(1 + u_i);  KM rpm_i^2 = KM H^2 (1 + u_i)
    float u[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float sk = onep[k] - 1.0f;                // exact (Sterbenz), = 0.05 a as numpy rounds it
      u[k] = sk * (2.0f + sk);
    }
    const float qf = 0.25f * P.gravity;                 // KF H^2
    const float qm = P.km * (P.hover_rpm * P.hover_rpm);  // KM H^2
    float tz = qm * ((-u[0] + u[1]) + (-u[2] + u[3]));
    float tx, ty;
    if (P.model == MODEL_CF2X) { tx = -qf * ((u[0] + u[1]) - (u[2] + u[3])) * P.arm; ty = qf * ((-u[0] + u[1]) + (u[2] - u[3])) * P.arm; }
    else if (P.model == MODEL_CF2P) { tx = qf * (u[1] - u[3]) * P.arm; ty = qf * (-u[0] + u[2]) * P.arm; }
    else { tx = qf * ((u[0] + u[1]) - (u[2] + u[3])) * P.arm; ty = qf * ((-u[0] + u[1]) + (u[2] - u[3])) * P.arm; tz = -tz; }
    const float kx = dt * P.ijx * tx, ky = dt * P.ijy * ty, kz = dt * P.ijz * tz;
    const float gx = dt * P.ijx * (P.jz - P.jy), gy = dt * P.ijy * (P.jx - P.jz), gz = dt * P.ijz * (P.jy - P.jx);
    const float e4 = 0.25f * ((u[0] + u[1]) + (u[2] + u[3]));
    const float cg = dt * P.gravity * P.inv_m;          // dt g
    const float c1 = cg * (1.0f + e4), c2 = cg * e4, c3 = -2.0f * cg;
    const float half_dt = 0.5f * dt;
#pragma unroll 1
    for (int s = 0; s < P.S; ++s) {
      const float x = d.qx, y = d.qy, z = d.qz, w = d.qw;   // unit, canonical
      const float xxyy = fmaf(x, x, y * y);
      const float r02 = 2.0f * fmaf(x, z, w * y), r12 = 2.0f * fmaf(y, z, -w * x),
                  r22 = fmaf(-2.0f, xxyy, 1.0f);
      // a = g [(1+e) R[:,2] - e_z];  (1+e) r22 - 1 = e r22 - 2 (x^2 + y^2)
      d.vx = fmaf(c1, r02, d.vx);
      d.vy = fmaf(c1, r12, d.vy);
      d.vz = fmaf(c2, r22, fmaf(c3, xxyy, d.vz));
      const float owx = d.wx, owy = d.wy, owz = d.wz;
      d.wx = fmaf(-gx, owy * owz, owx + kx);
      d.wy = fmaf(-gy, owz * owx, owy + ky);
      d.wz = fmaf(-gz, owx * owy, owz + kz);
      d.px = fmaf(dt, d.vx, d.px);
      d.py = fmaf(dt, d.vy, d.py);
      d.pz = fmaf(dt, d.vz, d.pz);
      if (s == P.S - 1) {
        const float r00 = fmaf(-2.0f, fmaf(y, y, z * z), 1.0f), r01 = 2.0f * fmaf(x, y, -w * z);
        const float r10 = 2.0f * fmaf(x, y, w * z), r11 = fmaf(-2.0f, fmaf(x, x, z * z), 1.0f);
        const float r20 = 2.0f * fmaf(x, z, -w * y), r21 = 2.0f * fmaf(y, z, w * x);
        avx = fmaf(r00, d.wx, fmaf(r01, d.wy, r02 * d.wz));
        avy = fmaf(r10, d.wx, fmaf(r11, d.wy, r12 * d.wz));
        avz = fmaf(r20, d.wx, fmaf(r21, d.wy, r22 * d.wz));
      }
      integrate_q_fast(d.qx, d.qy, d.qz, d.qw, d.wx, d.wy, d.wz, half_dt);
      fast_canonical(d.qx, d.qy, d.qz, d.qw);
    }
  } else {
#pragma unroll 1
    for (int s = 0; s < P.S; ++s) {
      if (GENERIC && (P.aero & AERO_DW)) {   // Jacobi snapshot of the env's positions (:346-347,799-800)
        __syncthreads();
        spos[tid * 3 + 0] = d.px; spos[tid * 3 + 1] = d.py; spos[tid * 3 + 2] = d.pz;
        __syncthreads();
      }
      R m[9];
      quat_to_mat(d.qx, d.qy, d.qz, d.qw, m);
      R f[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) f[k] = rpm[k] * rpm[k] * P.kf;                      // :838
      R fx_extra = R(0), fy_extra = R(0), fz_extra = R(0);
      if (GENERIC && P.aero) {
        if (P.aero & AERO_GND) {             // :739-742, added to the propeller thrusts
          R roll, pitch, yaw;
          quat_to_euler(d.qx, d.qy, d.qz, d.qw, roll, pitch, yaw);
          const R half_pi = R(0.5 * 3.14159265358979323846);
          if (abs_(roll) < half_pi && abs_(pitch) < half_pi) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              R h = d.pz + (m[6] * P.prop_x[k] + m[7] * P.prop_y[k]);
              h = h < P.gnd_h_clip ? P.gnd_h_clip : h;
              const R ratio = P.prop_radius / (R(4) * h);
              f[k] += rpm[k] * rpm[k] * P.kf * P.gnd_coeff * (ratio * ratio);
            }
          }
        }
        if (P.aero & AERO_DRAG) {            // :773-774 -> world force k (.) v
          const R* lr = (s == 0) ? last_rpm : rpm;
          const R two_pi = R(2) * R(3.14159265358979323846);
          const R sum = two_pi * lr[0] / R(60) + two_pi * lr[1] / R(60) + two_pi * lr[2] / R(60) +
                        two_pi * lr[3] / R(60);
          fx_extra += R(-1) * P.drag_xy * sum * d.vx;
          fy_extra += R(-1) * P.drag_xy * sum * d.vy;
          fz_extra += R(-1) * P.drag_z * sum * d.vz;
        }
        if (P.aero & AERO_DW) {              // :798-804, body-z force -> R [0,0,F]
          R dw = R(0);
          const R* ep = spos + (size_t)env_l * M * 3;
          for (int i = 0; i < M; ++i) {
            const R dz = ep[i * 3 + 2] - d.pz;
            const R ddx = ep[i * 3] - d.px, ddy = ep[i * 3 + 1] - d.py;
            const R dxy = sqrt_(ddx * ddx + ddy * ddy);
            if (dz > R(0) && dxy < R(10)) {
              const R ratio = P.prop_radius / (R(4) * dz);
              const R alpha = P.dw1 * (ratio * ratio);
              const R beta = P.dw2 * dz + P.dw3;
              const R q_ = dxy / beta;
              dw += -alpha * exp_(R(-0.5) * (q_ * q_));
            }
          }
          fx_extra += m[2] * dw; fy_extra += m[5] * dw; fz_extra += m[8] * dw;
        }
      }
      const R thrust = ((f[0] + f[1]) + f[2]) + f[3];
      const R Fx = m[2] * thrust + fx_extra;
      const R Fy = m[5] * thrust + fy_extra;
      const R Fz = (m[8] * thrust - P.gravity) + fz_extra;                        // :839-841
      R zt[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) zt[k] = rpm[k] * rpm[k] * P.km;
      R tx, ty, tz;
      if (P.model == MODEL_RACE) {
        tz = ((zt[0] - zt[1]) + zt[2]) - zt[3];
        tx = (f[0] + f[1] - f[2] - f[3]) * P.arm;
        ty = (-f[0] + f[1] + f[2] - f[3]) * P.arm;
      } else if (P.model == MODEL_CF2X) {
        tz = ((-zt[0] + zt[1]) - zt[2]) + zt[3];
        tx = -(f[0] + f[1] - f[2] - f[3]) * P.arm;
        ty = (-f[0] + f[1] + f[2] - f[3]) * P.arm;
      } else {
        tz = ((-zt[0] + zt[1]) - zt[2]) + zt[3];
        tx = (f[1] - f[3]) * P.arm;
        ty = (-f[0] + f[2]) * P.arm;
      }
      // torques - w x (J w), J diagonal (:856-857)
      const R jwx = P.jx * d.wx, jwy = P.jy * d.wy, jwz = P.jz * d.wz;
      tx -= d.wy * jwz - d.wz * jwy;
      ty -= d.wz * jwx - d.wx * jwz;
      tz -= d.wx * jwy - d.wy * jwx;
      d.vx += dt * (Fx * P.inv_m);                                                // :860
      d.vy += dt * (Fy * P.inv_m);
      d.vz += dt * (Fz * P.inv_m);
      d.wx += dt * (P.ijx * tx);                                                  // :861
      d.wy += dt * (P.ijy * ty);
      d.wz += dt * (P.ijz * tz);
      d.px += dt * d.vx;                                                          // :862
      d.py += dt * d.vy;
      d.pz += dt * d.vz;
      if (!GENERIC || P.integrator == 0) {
        avx = m[0] * d.wx + m[1] * d.wy + m[2] * d.wz;                            // :873
        avy = m[3] * d.wx + m[4] * d.wy + m[5] * d.wz;
        avz = m[6] * d.wx + m[7] * d.wy + m[8] * d.wz;
        integrate_q_exact(d.qx, d.qy, d.qz, d.qw, d.wx, d.wy, d.wz, dt);          // :863
      } else {  // Euler-angle integrator (scg base_aviary.py:499-508)
        R roll, pitch, yaw;
        quat_to_euler(d.qx, d.qy, d.qz, d.qw, roll, pitch, yaw);
        euler_to_quat(roll + dt * d.wx, pitch + dt * d.wy, yaw + dt * d.wz, d.qx, d.qy, d.qz, d.qw);
        avx = d.wx; avy = d.wy; avz = d.wz;
      }
      bullet_roundtrip(d.qx, d.qy, d.qz, d.qw);                                   // :347 / :374
    }
  }

  // ---- 4. observation, reward terms, flags ----------------------------------
  R roll, pitch, yaw;
  quat_to_euler(d.qx, d.qy, d.qz, d.qw, roll, pitch, yaw);
  R contrib = R(0);
  int flags = 0;   // bit0: terminated condition, bit1: truncated condition (hover bounds)
  if (active) {
    {
      // kinematic part [pos rpy vel ang_v] (BaseRLAviary.py:314-315)
      float4* r4 = reinterpret_cast<float4*>(myrow);
      r4[0] = make_float4((float)d.px, (float)d.py, (float)d.pz, (float)roll);
      r4[1] = make_float4((float)pitch, (float)yaw, (float)d.vx, (float)d.vy);
      r4[2] = make_float4((float)d.vz, (float)avx, (float)avy, (float)avz);
    }
    if (TASK == TASK_HOVER) {
      const R ex = d.tx - d.px, ey = d.ty - d.py, ez = d.tz - d.pz;
      const R dist = sqrt_(ex * ex + ey * ey + ez * ez);
      const R d2 = dist * dist;
      const R r = R(2) - d2 * d2;                                                  // HoverAviary.py:78
      contrib = r > R(0) ? r : R(0);
      if (dist < R(.0001)) flags |= 1;                                             // :93
      if (abs_(d.px) > R(1.5) || abs_(d.py) > R(1.5) || d.pz > R(2.0) ||
          abs_(roll) > R(.4) || abs_(pitch) > R(.4)) flags |= 2;                   // :109-111
    } else if (TASK == TASK_MULTIHOVER) {
      const R ex = d.px - d.tx, ey = d.py - d.ty;
      const R err_xy = sqrt_(ex * ex + ey * ey);
      const R err_z = d.pz - d.tz;
      const R vel_z = d.vz;
      const R r_xy = R(1) / (R(1) + err_xy);
      const R r_z = exp_(R(-7.5) * abs_(err_z));
      const R r_vel = abs_(err_z) < R(0.2) ? R(-1.5) * (vel_z * vel_z) : R(0);
      const R bonus = (err_xy < R(0.03) && abs_(err_z) < R(0.03) && abs_(vel_z) < R(0.03)) ? R(0.5) : R(0);
      contrib = ((r_xy + r_z) + r_vel) + bonus;                                    // MultiHoverAviary.py:173-179
      if (d.pz < R(0.03)) flags |= 1;                                              // :226
      if (abs_(roll) > R(1.2) || abs_(pitch) > R(1.2)) flags |= 1;                 // :231
      if (abs_(d.px) > R(3.0) || abs_(d.py) > R(3.0)) flags |= 1;                  // :236
    } else {
      R rp[3], rv[3], sphi, cphi;
      spiral_reference(P, ec.x, drone, rp, rv, sphi, cphi);
      // SpiralAviary.py:130,156: "vel" = state[3:6] = quaternion x,y,z (reference quirk)
      const R qv[3] = {d.qx, d.qy, d.qz};
      float* ext = myrow + 12 + B * A;
      ext[0] = (float)(rp[0] - d.px); ext[1] = (float)(rp[1] - d.py); ext[2] = (float)(rp[2] - d.pz);
      ext[3] = (float)(rv[0] - qv[0]); ext[4] = (float)(rv[1] - qv[1]); ext[5] = (float)(rv[2] - qv[2]);
      ext[6] = (float)sphi; ext[7] = (float)cphi;
      ext[8] = (float)rv[0]; ext[9] = (float)rv[1]; ext[10] = (float)rv[2];
      const R dpx = d.px - rp[0], dpy = d.py - rp[1], dpz = d.pz - rp[2];
      const R npos = sqrt_(dpx * dpx + dpy * dpy + dpz * dpz);
      const R dvx = qv[0] - rv[0], dvy = qv[1] - rv[1], dvz = qv[2] - rv[2];
      const R nvel = sqrt_(dvx * dvx + dvy * dvy + dvz * dvz);
      const R r_pos = exp_(R(-4.0) * (npos * npos));
      const R r_vel = exp_(R(-2.0) * (nvel * nvel));
      R r_tan = R(0);
      const R rx = d.px - P.sp_cx, ry = d.py - P.sp_cy;
      const R nr = sqrt_(rx * rx + ry * ry);
      if (nr > R(1e-3)) {
        const R tgx = -(ry / nr), tgy = rx / nr;
        const R nv = sqrt_(qv[0] * qv[0] + qv[1] * qv[1]);
        if (nv > R(1e-3)) {
          const R dot = (qv[0] / nv) * tgx + (qv[1] / nv) * tgy;
          r_tan = dot > R(0) ? dot : R(0);
        }
      }
      contrib = (R(1.0) * r_pos + R(2.0) * r_vel) + R(1.0) * r_tan;               // :179
      if (d.pz < R(0.05) || d.pz > R(3.0)) flags |= 1;                             // :188-190
    }
  }
  sred[tid] = contrib;
  sflag[tid] = flags;
  cp_async_wait_all();
  __syncthreads();

  // ---- 5. per-env reduction by the env's first drone --------------------------
  int done_reset = 0;
  if (active && drone == 0) {
    R sum = R(0);
    int fl = 0;
    for (int i = 0; i < M; ++i) { sum += sred[tid + i]; fl |= sflag[tid + i]; }
    const R reward = (TASK == TASK_HOVER) ? sum : sum / R(M);
    const bool time_up = ((double)ec.x / P.pyb_freq) > P.episode_len;               // pre-increment (:379,:382)
    const bool terminated = (fl & 1) != 0;
    const bool truncated = ((fl & 2) != 0) || time_up;
    P.reward[env] = reward;
    P.terminated[env] = terminated ? 1 : 0;
    P.truncated[env] = truncated ? 1 : 0;
    done_reset = ((terminated || truncated) && P.auto_reset) ? 1 : 0;
    P.envc[env] = make_int2(done_reset ? 0 : ec.x + P.S, ec.y + 1);
    sdone[env_l] = done_reset;
  }
  const int any_reset = __syncthreads_or(done_reset);

  // ---- 6. reset-on-done (subproc_vec_env.py:195-206) ---------------------------
  if (any_reset) {
    const bool my_reset = active && sdone[env_l];
    const bool jit = (TASK == TASK_MULTIHOVER) && (P.reset_mode != RESET_FIXED);
    if (jit) {
      if (my_reset && drone == 0) sample_jitter(P, env, ec.y, 0, spos + (size_t)tid * 3);
      __syncthreads();
    }
    if (my_reset) {
      if (P.terminal_obs != nullptr) {
        float* to = P.terminal_obs + (size_t)g * P.D;
        for (int k = 0; k < P.D; ++k) to[k] = myrow[k];
      }
      float kin[12];
      reset_drone<R, TASK>(P, env, drone, jit ? spos + (size_t)tid * 3 : nullptr, d, kin);
#pragma unroll
      for (int k = 0; k < 12; ++k) myrow[k] = kin[k];
      avx = avy = avz = R(0);
      if (TASK == TASK_SPIRAL) {   // reset obs is evaluated at step_counter = 0
        R rp[3], rv[3], sphi, cphi;
        spiral_reference(P, 0, drone, rp, rv, sphi, cphi);
        float* ext = myrow + 12 + B * A;
        ext[0] = (float)(rp[0] - d.px); ext[1] = (float)(rp[1] - d.py); ext[2] = (float)(rp[2] - d.pz);
        ext[3] = (float)(rv[0] - d.qx); ext[4] = (float)(rv[1] - d.qy); ext[5] = (float)(rv[2] - d.qz);
        ext[6] = (float)sphi; ext[7] = (float)cphi;
        ext[8] = (float)rv[0]; ext[9] = (float)rv[1]; ext[10] = (float)rv[2];
      }
    }
    __syncthreads();
  }

  // ---- 7. state store + coalesced observation tile store -----------------------
  if (active) {
    P.s0[g] = make4(d.px, d.py, d.pz, d.qx);
    P.s1[g] = make4(d.qy, d.qz, d.qw, d.vx);
    P.s2[g] = make4(d.vy, d.vz, d.wx, d.wy);
    P.s3[g] = make4(d.wz, d.tx, d.ty, d.tz);
    if (GENERIC && P.keep_angv) P.s4[g] = make4(avx, avy, avz, R(0));
  }
  {
    const int envs_here = min(E, P.N - blockIdx.x * E);
    const int rows = envs_here * M;
    float* gobs = P.obs + (size_t)blockIdx.x * E * M * P.D;
    if (P.D == Ds) {
      const int n4 = rows * (Ds >> 2);
      const float4* s4p = reinterpret_cast<const float4*>(sobs);
      float4* g4p = reinterpret_cast<float4*>(gobs);
      for (int i = tid; i < n4; i += kBlock) g4p[i] = s4p[i];
    } else {
      const int n = rows * P.D;
      for (int i = tid; i < n; i += kBlock) {
        const int r = i / P.D, c = i - r * P.D;
        gobs[i] = sobs[(size_t)r * Ds + c];
      }
    }
  }
}

// =========================================================================
//                    reset / state access kernels (cold paths)
// =========================================================================
template <typename R, int TASK, int A>
__global__ void __launch_bounds__(kBlock)
reset_kernel(const __grid_constant__ Params<R> P) {
  using R4 = typename V4<R>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  R* spos = reinterpret_cast<R*>(smem_raw);   // [kBlock][3]
  const int tid = threadIdx.x;
  const int M = P.M, E = P.E, B = P.B;
  const int env_l = tid / M, drone = tid - env_l * M;
  const int env = blockIdx.x * E + env_l;
  const bool active = (env_l < E) && (env < P.N) && (P.reset_mask == nullptr || P.reset_mask[env] != 0);
  const long long g = (long long)env * M + drone;
  int2 ec = make_int2(0, 0);
  if (active) ec = P.envc[env];
  const bool jit = (TASK == TASK_MULTIHOVER) && (P.reset_mode != RESET_FIXED);
  if (jit) {
    if (active && drone == 0) sample_jitter(P, env, ec.y, P.reset_epoch, spos + (size_t)tid * 3);
    __syncthreads();
  }
  if (!active) return;
  Drone<R> d;
  float kin[12];
  reset_drone<R, TASK>(P, env, drone, jit ? spos + (size_t)tid * 3 : nullptr, d, kin);
  P.s0[g] = make4(d.px, d.py, d.pz, d.qx);
  P.s1[g] = make4(d.qy, d.qz, d.qw, d.vx);
  P.s2[g] = make4(d.vy, d.vz, d.wx, d.wy);
  P.s3[g] = make4(d.wz, d.tx, d.ty, d.tz);
  if (P.keep_angv) P.s4[g] = make4(R(0), R(0), R(0), R(0));
  if (drone == 0) P.envc[env] = make_int2(0, ec.y);   // ring head (total steps) is untouched by a reset
  if (P.obs != nullptr) {
    float* row = P.obs + (size_t)g * P.D;
    for (int k = 0; k < 12; ++k) row[k] = kin[k];
    const int head = ec.y % B;   // oldest entry lives in slot `head`
    int slot = head;
    for (int j = 0; j < B; ++j) {
      const float* hp = P.hist + ((size_t)slot * P.n_total + g) * A;
      for (int k = 0; k < A; ++k) row[12 + j * A + k] = hp[k];
      slot = (slot + 1 == B) ? 0 : slot + 1;
    }
    if (TASK == TASK_SPIRAL) {
      R rp[3], rv[3], sphi, cphi;
      spiral_reference(P, 0, drone, rp, rv, sphi, cphi);
      float* ext = row + 12 + B * A;
      ext[0] = (float)(rp[0] - d.px); ext[1] = (float)(rp[1] - d.py); ext[2] = (float)(rp[2] - d.pz);
      ext[3] = (float)(rv[0] - d.qx); ext[4] = (float)(rv[1] - d.qy); ext[5] = (float)(rv[2] - d.qz);
      ext[6] = (float)sphi; ext[7] = (float)cphi;
      ext[8] = (float)rv[0]; ext[9] = (float)rv[1]; ext[10] = (float)rv[2];
    }
  }
}

template <typename R>
__global__ void get_state_kernel(const __grid_constant__ Params<R> P, R* state20, R* rates,
                                 int32_t* step_counter) {
  using R4 = typename V4<R>::type;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= P.n_total) return;
  const int env = (int)(g / P.M);
  const R4 a0 = P.s0[g], a1 = P.s1[g], a2 = P.s2[g], a3 = P.s3[g];
  const int2 ec = P.envc[env];
  if (state20 != nullptr) {
    R* o = state20 + g * 20;
    o[0] = a0.x; o[1] = a0.y; o[2] = a0.z;
    o[3] = a0.w; o[4] = a1.x; o[5] = a1.y; o[6] = a1.z;
    R roll, pitch, yaw;
    quat_to_euler(a0.w, a1.x, a1.y, a1.z, roll, pitch, yaw);
    o[7] = roll; o[8] = pitch; o[9] = yaw;
    o[10] = a1.w; o[11] = a2.x; o[12] = a2.y;
    if (P.keep_angv) { const R4 av = P.s4[g]; o[13] = av.x; o[14] = av.y; o[15] = av.z; }
    else { const R nanv = R(nan("")); o[13] = o[14] = o[15] = nanv; }
    // last_clipped_action: rpm of the newest ring entry, zero right after a reset (:468)
    const int newest = (ec.y % P.B == 0) ? P.B - 1 : (ec.y % P.B) - 1;
    const float* hp = P.hist + ((size_t)newest * P.n_total + g) * P.A;
    for (int k = 0; k < 4; ++k) {
      const float a = hp[P.A == 4 ? k : 0];
      const float s = __fadd_rn(1.0f, __fmul_rn(0.05f, a));
      o[16 + k] = ec.x > 0 ? P.hover_rpm * (R)s : R(0);
    }
  }
  if (rates != nullptr) { rates[g * 3] = a2.z; rates[g * 3 + 1] = a2.w; rates[g * 3 + 2] = a3.x; }
  if (step_counter != nullptr && g % P.M == 0) step_counter[env] = ec.x;
}

template <typename R>
__global__ void set_state_kernel(const __grid_constant__ Params<R> P, const R* kin13,
                                 const R* targets, const int32_t* step_counter) {
  using R4 = typename V4<R>::type;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= P.n_total) return;
  const int env = (int)(g / P.M);
  if (kin13 != nullptr) {
    const R* k = kin13 + g * 13;
    const R4 old3 = P.s3[g];
    P.s0[g] = make4(k[0], k[1], k[2], k[3]);
    P.s1[g] = make4(k[4], k[5], k[6], k[7]);
    P.s2[g] = make4(k[8], k[9], k[10], k[11]);
    P.s3[g] = make4(k[12], old3.y, old3.z, old3.w);
    if (P.keep_angv) P.s4[g] = make4(R(0), R(0), R(0), R(0));
  }
  if (targets != nullptr) {
    const R4 old3 = P.s3[g];
    P.s3[g] = make4(old3.x, targets[g * 3], targets[g * 3 + 1], targets[g * 3 + 2]);
  }
  if (step_counter != nullptr && g % P.M == 0) {
    int2 ec = P.envc[env];
    ec.x = step_counter[env];
    P.envc[env] = ec;
  }
}

template <typename R>
__global__ void get_targets_kernel(const __grid_constant__ Params<R> P, R* targets) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= P.n_total) return;
  const auto a3 = P.s3[g];
  targets[g * 3] = a3.y; targets[g * 3 + 1] = a3.z; targets[g * 3 + 2] = a3.w;
}

// =========================================================================
//                              host-side dispatch
// =========================================================================
size_t step_smem_bytes(int precision, int Ds) {
  const size_t real = precision ? 8 : 4;
  return (size_t)kBlock * Ds * 4 + (size_t)kBlock * real * 4 + (size_t)kBlock * 8;
}

template <typename R, int TASK, int A, bool GENERIC>
static cudaError_t launch_step_t(const Params<R>& P, int device, cudaStream_t st) {
  const size_t smem = step_smem_bytes(sizeof(R) == 8, P.Ds);
  auto kern = step_kernel<R, TASK, A, GENERIC>;
  static size_t configured[64] = {0};   // per device: the attribute is per context
  if (smem > configured[device & 63]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured[device & 63] = smem;
  }
  const int grid = (P.N + P.E - 1) / P.E;
  kern<<<grid, kBlock, smem, st>>>(P);
  return cudaGetLastError();
}

template <typename R, int TASK, int A>
static cudaError_t launch_reset_t(const Params<R>& P, cudaStream_t st) {
  const int grid = (P.N + P.E - 1) / P.E;
  reset_kernel<R, TASK, A><<<grid, kBlock, kBlock * 3 * sizeof(R), st>>>(P);
  return cudaGetLastError();
}

template <typename R, int TASK, int A>
static cudaError_t step_g(const LaunchSpec& ls, const void* p, cudaStream_t st) {
  const Params<R>& P = *static_cast<const Params<R>*>(p);
  return ls.generic ? launch_step_t<R, TASK, A, true>(P, ls.device, st)
                    : launch_step_t<R, TASK, A, false>(P, ls.device, st);
}
template <typename R, int TASK>
static cudaError_t step_a(const LaunchSpec& ls, const void* p, cudaStream_t st) {
  return ls.act_a == 4 ? step_g<R, TASK, 4>(ls, p, st) : step_g<R, TASK, 1>(ls, p, st);
}
template <typename R>
static cudaError_t step_t(const LaunchSpec& ls, const void* p, cudaStream_t st) {
  switch (ls.task) {
    case TASK_HOVER: return step_a<R, TASK_HOVER>(ls, p, st);
    case TASK_MULTIHOVER: return step_a<R, TASK_MULTIHOVER>(ls, p, st);
    default: return step_a<R, TASK_SPIRAL>(ls, p, st);
  }
}
cudaError_t launch_step(const LaunchSpec& ls, const void* params, cudaStream_t st) {
  return ls.precision ? step_t<double>(ls, params, st) : step_t<float>(ls, params, st);
}

template <typename R, int TASK>
static cudaError_t reset_a(const LaunchSpec& ls, const void* p, cudaStream_t st) {
  const Params<R>& P = *static_cast<const Params<R>*>(p);
  return ls.act_a == 4 ? launch_reset_t<R, TASK, 4>(P, st) : launch_reset_t<R, TASK, 1>(P, st);
}
template <typename R>
static cudaError_t reset_t(const LaunchSpec& ls, const void* p, cudaStream_t st) {
  switch (ls.task) {
    case TASK_HOVER: return reset_a<R, TASK_HOVER>(ls, p, st);
    case TASK_MULTIHOVER: return reset_a<R, TASK_MULTIHOVER>(ls, p, st);
    default: return reset_a<R, TASK_SPIRAL>(ls, p, st);
  }
}
cudaError_t launch_reset(const LaunchSpec& ls, const void* params, cudaStream_t st) {
  return ls.precision ? reset_t<double>(ls, params, st) : reset_t<float>(ls, params, st);
}

template <typename R>
static int flat_grid(const Params<R>& P) { return (int)((P.n_total + 255) / 256); }

cudaError_t launch_get_state(int precision, const void* p, void* state20, void* rates,
                             int32_t* step_counter, cudaStream_t st) {
  if (precision) {
    const auto& P = *static_cast<const Params<double>*>(p);
    get_state_kernel<double><<<flat_grid(P), 256, 0, st>>>(P, (double*)state20, (double*)rates, step_counter);
  } else {
    const auto& P = *static_cast<const Params<float>*>(p);
    get_state_kernel<float><<<flat_grid(P), 256, 0, st>>>(P, (float*)state20, (float*)rates, step_counter);
  }
  return cudaGetLastError();
}

cudaError_t launch_set_state(int precision, const void* p, const void* kin13, const void* targets,
                             const int32_t* step_counter, cudaStream_t st) {
  if (precision) {
    const auto& P = *static_cast<const Params<double>*>(p);
    set_state_kernel<double><<<flat_grid(P), 256, 0, st>>>(P, (const double*)kin13, (const double*)targets, step_counter);
  } else {
    const auto& P = *static_cast<const Params<float>*>(p);
    set_state_kernel<float><<<flat_grid(P), 256, 0, st>>>(P, (const float*)kin13, (const float*)targets, step_counter);
  }
  return cudaGetLastError();
}

cudaError_t launch_get_targets(int precision, const void* p, void* targets, cudaStream_t st) {
  if (precision) {
    const auto& P = *static_cast<const Params<double>*>(p);
    get_targets_kernel<double><<<flat_grid(P), 256, 0, st>>>(P, (double*)targets);
  } else {
    const auto& P = *static_cast<const Params<float>*>(p);
    get_targets_kernel<float><<<flat_grid(P), 256, 0, st>>>(P, (float*)targets);
  }
  return cudaGetLastError();
}

}  // namespace bd
