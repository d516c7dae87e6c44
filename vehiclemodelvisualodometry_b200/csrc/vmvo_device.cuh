// Device-side building blocks shared by the VMVO kernels (sm_100a).
//
// Arithmetic follows vmvo/bicycle_model.py:66-75 of the reference:
//   delta = radians(s) / ratio;  th' = th + ((v / L) * tan(delta)) * dt
//   x' = x + (v * cos(th')) * dt;  y' = y + (v * sin(th')) * dt
// float64 paths use explicit round-to-nearest intrinsics so nvcc never contracts a
// product into an FMA: grid controls V_k, S_k are then bit-identical to the NumPy oracle
// and only the libm-level functions (tan, sincos, atan) differ, by <= 2 ulp.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vmvo_b200.h"

namespace vmvo {

constexpr unsigned FULL = 0xffffffffu;
constexpr double kPi = 3.14159265358979323846;
constexpr double kTwoPi = 6.28318530717958647692;
constexpr double kDegToRad = kPi / 180.0;   // np.radians multiplies by this constant
constexpr double kRadToDeg = 180.0 / kPi;

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

// ---- typed wrappers so the rollout code is written once for float and double ----------
template <typename T> struct Num;
template <> struct Num<double> {
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
  static __device__ __forceinline__ double tan_(double a) { return tan(a); }
  static __device__ __forceinline__ void sincos_(double a, double* s, double* c) { sincos(a, s, c); }
  static __device__ __forceinline__ double abs_(double a) { return fabs(a); }
  static __device__ __forceinline__ double deg2rad() { return kDegToRad; }
};
template <> struct Num<float> {
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
  static __device__ __forceinline__ float tan_(float a) { return tanf(a); }
  static __device__ __forceinline__ void sincos_(float a, float* s, float* c) { sincosf(a, s, c); }
  static __device__ __forceinline__ float abs_(float a) { return fabsf(a); }
  static __device__ __forceinline__ float deg2rad() { return (float)kDegToRad; }
};

// ---- warp primitives -------------------------------------------------------------------
// Inclusive scan over aligned groups of 2^STEPS lanes (the whole warp by default), Sklansky form:
// in step s the upper half of every aligned block of 2^(s+1) lanes adds the total of its lower
// half.  Same five shuffle + add steps as Hillis-Steele, but the association of the additions is
// anchored at lane 0 of the group instead of at the receiving lane, which gives two properties the
// search relies on:
//   * every op is sign-symmetric, so negated inputs give exactly negated outputs (mirror hypotheses
//     tie exactly);
//   * if the inputs are zero from lane m on, every lane >= m - 1 holds the SAME value, bit for bit
//     (x + 0 = x), and the lanes below 2^s never see steps >= s -- so a hypothesis that stops after
//     m <= G steps gets identical sums from a G-lane group as from the whole warp (the packed
//     float64 re-score of vmvo_deferred_rescore_kernel).
//
// HS = true: the Hillis-Steele form over the whole warp (additions associated from the receiving
// lane; sign-symmetric as well).  The dense-grid kernel, whose windows are never parked for the
// second kernel, re-scores with it: the same instruction count, but the kernel measures 3 % faster
// (profiles/README.md).  A launch uses ONE form throughout (SearchParams::scan_hs).
template <typename T, int STEPS = 5, bool HS = false>
__device__ __forceinline__ T warp_scan_add(T v, int lane) {
  if (HS) {
    static_assert(!HS || STEPS == 5, "the Hillis-Steele form is for whole-warp scans");
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      T t = __shfl_up_sync(FULL, v, o);
      if (lane >= o) v = Num<T>::add(v, t);
    }
    return v;
  }
#pragma unroll
  for (int s = 0; s < STEPS; ++s) {
    const int src = (((lane >> s) << s) - 1) & 31;   // last lane of the block to the left
    T t = __shfl_sync(FULL, v, src);
    if ((lane >> s) & 1) v = Num<T>::add(v, t);
  }
  return v;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = Num<T>::add(v, __shfl_xor_sync(FULL, v, o));
  return v;
}

__device__ __forceinline__ float warp_min_f32_nonneg(float v) {
  // non-negative floats (and +inf) order like their bit patterns
  return __uint_as_float(__reduce_min_sync(FULL, __float_as_uint(v)));
}

// ---- hypothesis grid -----------------------------------------------------
// rate_i = limit * (2i - (g-1)) / (g-1): exactly antisymmetric about the grid centre.
__device__ __forceinline__ double grid_rate(double limit, int idx, int g) {
  if (g <= 1) return 0.0;
  return ddiv(dmul(limit, (double)(2 * idx - (g - 1))), (double)(g - 1));
}

struct GridCtl {
  double v_seed, s_seed, dt;
  double accel, srate;       // a_i, r_j of this hypothesis
  double max_steer;
  __device__ __forceinline__ void at(int k, double* v, double* s) const {
    double t = dmul((double)k, dt);
    double vv = dadd(v_seed, dmul(accel, t));
    *v = vv > 0.0 ? vv : (vv == vv ? 0.0 : vv);          // np.maximum(0, v): NaN propagates
    double ss = dadd(s_seed, dmul(srate, t));
    ss = ss < -max_steer ? -max_steer : ss;
    ss = ss > max_steer ? max_steer : ss;
    *s = ss;
  }
};

// ---- one 32-step round of the model, one step per lane ------------------------------------
// carry = pose before the round's first step; returns this lane's pose after its step and
// advances the carry to the pose after the round's last step.
template <typename T>
struct Pose { T x, y, th; };

// (STEPS < 5: independent rounds in aligned groups of 2^STEPS lanes, one carry per group)
template <typename T, int STEPS = 5, bool HS = false>
__device__ __forceinline__ Pose<T> warp_model_round(T v, T s_deg, bool active, T dt, T L, T ratio,
                                                    Pose<T>& carry, int lane) {
  using N = Num<T>;
  T inc_th = (T)0;
  if (active && v != (T)0) {
    T delta = N::div(N::mul(s_deg, N::deg2rad()), ratio);
    inc_th = N::mul(N::mul(N::div(v, L), N::tan_(delta)), dt);
  }
  T th = N::add(carry.th, warp_scan_add<T, STEPS, HS>(inc_th, lane));
  T ix = (T)0, iy = (T)0;
  if (active && v != (T)0) {
    T sn, cs;
    N::sincos_(th, &sn, &cs);
    ix = N::mul(N::mul(v, cs), dt);
    iy = N::mul(N::mul(v, sn), dt);
  }
  Pose<T> p;
  p.th = th;
  p.x = N::add(carry.x, warp_scan_add<T, STEPS, HS>(ix, lane));
  p.y = N::add(carry.y, warp_scan_add<T, STEPS, HS>(iy, lane));
  const int last = lane | ((1 << STEPS) - 1);
  carry.th = __shfl_sync(FULL, p.th, last);
  carry.x = __shfl_sync(FULL, p.x, last);
  carry.y = __shfl_sync(FULL, p.y, last);
  return p;
}

// tan of a road-wheel angle for the search's TL table.  Steering is mechanically limited
// (MAX_STEER / ratio = 0.605 rad with the reference's constants), so the argument is small: the
// Taylor polynomial through x^19, Horner in x^2 with FMAs, is within 0.74 ulp of tan(x) for every
// float |x| <= 0.62 (all 1.1e8 of them checked against float64; tests/test_gpu_operators.py
// re-checks the compiled code), at a third of tanf's instructions.  Larger angles (a caller's own
// steering limits) go through tanf (<= 4 ulp).
constexpr float kTanPolyMax = 0.62f;
__device__ __forceinline__ float tan_poly(float x) {   // |x| <= kTanPolyMax
  const float x2 = x * x;
  float p = 0.0002391291200183332f;           // 443861162/1856156927625  x^19
  p = fmaf(p, x2, 0.0005900274263694882f);    // 6404582/10854718875       x^17
  p = fmaf(p, x2, 0.0014558343682438135f);    // 929569/638512875          x^15
  p = fmaf(p, x2, 0.0035921279340982437f);    // 21844/6081075             x^13
  p = fmaf(p, x2, 0.0088632358238101f);       // 1382/155925               x^11
  p = fmaf(p, x2, 0.021869488060474396f);     // 62/2835                   x^9
  p = fmaf(p, x2, 0.05396825447678566f);      // 17/315                    x^7
  p = fmaf(p, x2, 0.13333334028720856f);      // 2/15                      x^5
  p = fmaf(p, x2, 0.3333333432674408f);       // 1/3                       x^3
  return fmaf(x, x2 * p, x);
}
__device__ __forceinline__ float tan_steer(float x) {
  if (!(fabsf(x) <= kTanPolyMax)) return tanf(x);
  return tan_poly(x);
}

// Python / NumPy float modulo by a positive divisor (floor-mod): result in [0, b).
__device__ __forceinline__ double pymod_pos(double a, double b) {
  double r = fmod(a, b);
  if (r != 0.0) {
    if (r < 0.0) r = dadd(r, b);
  } else {
    r = 0.0;
  }
  return r;
}

// ---- 1-D bulk async copy (TMA unit, SASS UBLKCP) + mbarrier --------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// bytes must be a multiple of 16; both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

}  // namespace vmvo
