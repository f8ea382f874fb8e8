// Device functions shared by the update kernels: the restated exp of the reference's softmax.
// (expAndNormalize, pydensecrf/densecrf/src/densecrf.cpp:98-106 in /root/reference, evaluates `b.array().exp()` with
// Eigen 3.3.9's SSE packet exp for whole 4-packets of a column and with libm's expf for the L % 4 tail.)
#pragma once
#include <cuda_runtime.h>

namespace dcrf {

// One lane of Eigen 3.3.9's pexp<Packet4f> (Eigen/src/Core/arch/SSE/MathFunctions.h:107-170), which is what
// `b.array().exp()` (densecrf.cpp:103) evaluates for the whole packets of a column under SSE2: Cephes range
// reduction and a degree-5 polynomial, no FMA, floor by truncate-and-fix.  Arguments here are x - column max,
// i.e. <= 0, where the upper clamp and the final max(result, x) of the reference code are no-ops and are dropped.
__device__ __forceinline__ float eigen_pexp_nonpos(float x0) {
  float x = fmaxf(x0, -88.3762626647949f);
  float fx = __fadd_rn(__fmul_rn(x, 1.44269504088896341f), 0.5f);
  float t = __int2float_rn(__float2int_rz(fx));
  if (t > fx) t = __fsub_rn(t, 1.f);
  const int n = __float2int_rz(t);
  const float tmp = __fmul_rn(t, 0.693359375f);
  float z = __fmul_rn(t, -2.12194440e-4f);
  x = __fsub_rn(x, tmp);
  x = __fsub_rn(x, z);
  z = __fmul_rn(x, x);
  float y = 1.9875691500E-4f;
  y = __fadd_rn(__fmul_rn(y, x), 1.3981999507E-3f);
  y = __fadd_rn(__fmul_rn(y, x), 8.3334519073E-3f);
  y = __fadd_rn(__fmul_rn(y, x), 4.1665795894E-2f);
  y = __fadd_rn(__fmul_rn(y, x), 1.6666665459E-1f);
  y = __fadd_rn(__fmul_rn(y, x), 5.0000001201E-1f);
  y = __fadd_rn(__fmul_rn(y, z), x);
  y = __fadd_rn(y, 1.f);
  return __fmul_rn(y, __int_as_float((n + 0x7f) << 23));
}

// exp(x) for x <= 0 in double, rounded once to float: range reduction by ln2, degree-11 Taylor in
// FMA form (|r| <= 0.347: truncation 2^-45), scaling through the exponent field.  The L%4 tail of a column
// goes through libm's expf in the reference; this agrees with it except in rare last-ulp cases.
__device__ __forceinline__ float tail_exp_nonpos(float xf) {
  const double x = (double)fmaxf(xf, -200.f);
  const double kd0 = fma(x, 1.4426950408889634074, 6755399441055744.0);
  const int k = __double2loint(kd0);
  const double kd = kd0 - 6755399441055744.0;
  double r = fma(kd, -6.93147180369123816490e-01, x);
  r = fma(kd, -1.90821492927058770002e-10, r);
  double p = 2.50521083854417187751e-08;           // 1/11!
  p = fma(p, r, 2.75573192239858906526e-07);       // 1/10!
  p = fma(p, r, 2.75573192239858906526e-06);       // 1/9!
  p = fma(p, r, 2.48015873015873015873e-05);       // 1/8!
  p = fma(p, r, 1.98412698412698412698e-04);       // 1/7!
  p = fma(p, r, 1.38888888888888888889e-03);       // 1/6!
  p = fma(p, r, 8.33333333333333333333e-03);       // 1/5!
  p = fma(p, r, 4.16666666666666666667e-02);       // 1/4!
  p = fma(p, r, 1.66666666666666666667e-01);       // 1/3!
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  // p * 2^k, k in [-289, 0]: add k to the exponent field (p in [0.7, 1.42], no carry problems)
  const int hi = __double2hiint(p) + (k << 20);
  return (float)__hiloint2double(hi, __double2loint(p));
}

// b / s for the elements of a column that share the divisor: the instruction sequence of the IEEE division
// fast path (reciprocal refined once, quotient refined once from the exact remainder) with the reciprocal
// computed once.  Valid while no intermediate can leave the normal range; callers guarantee 1 <= s <= 2^8 and
// take __fdiv_rn for any numerator below 2^-63 (checked against __fdiv_rn by dcrf_debug_selftest).
__device__ __forceinline__ float refined_rcp(float s) {
  float r0;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(s));
  const float e = __fmaf_rn(-s, r0, 1.f);
  return __fmaf_rn(r0, e, r0);
}
__device__ __forceinline__ float div_shared(float b, float s, float r) {
  const float q0 = __fmaf_rn(b, r, 0.f);
  const float rem = __fmaf_rn(-s, q0, b);
  return __fmaf_rn(r, rem, q0);
}

}  // namespace dcrf
