// Packed two-float arithmetic (sm_100a FFMA2) shared by the update / filter kernels.
//
// Blackwell issues one fma.rn.f32x2 per two floats.  Bit-exactness against the reference's SSE2 code needs
// separately rounded multiplies and adds, but ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 (even
// under -fmad=false) and folds fma(a,b,-0) / fma(a,1,c) back into mul / add when it can see the constants.  So
// the identities are spelled with operands it cannot see through: a*b = fma(a, b, -0), a+b = fma(a, 1, b),
// a-b = fma(b, -1, a) with -0, 1, -1 read from the kernel arguments.  Each is exact: the product by +-1 is exact,
// and x + (-0) == x for every x.
#pragma once
#include <cuda_runtime.h>

namespace dcrf {

typedef unsigned long long u64;
struct Pk {
  u64 one, nzero, mone;
};
__device__ __forceinline__ u64 pk2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ u64 pk1(float v) { return pk2(v, v); }
__device__ __forceinline__ float lo2(u64 v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi2(u64 v) { return __uint_as_float((unsigned)(v >> 32)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b, const Pk& c) { return fma2(a, b, c.nzero); }
__device__ __forceinline__ u64 add2(u64 a, u64 b, const Pk& c) { return fma2(a, c.one, b); }
__device__ __forceinline__ u64 sub2(u64 a, u64 b, const Pk& c) { return fma2(b, c.mone, a); }

// eigen_pexp_nonpos (softmax_math.cuh) for two elements.  Same operations in the same order; two steps are
// restated in forms that give identical bits (checked exhaustively over every float in [-100, -0] by
// dcrf_debug_selftest): floor(fx) instead of truncate-and-fix, and 2^n built from the mantissa of t + 1.5*2^23
// (exact for the integer-valued t in [-127, 0]) instead of a float->int conversion.
__device__ __forceinline__ u64 eigen_pexp_nonpos2(u64 x0, const Pk& c) {
  u64 x = pk2(fmaxf(lo2(x0), -88.3762626647949f), fmaxf(hi2(x0), -88.3762626647949f));
  const u64 fx = add2(mul2(x, pk1(1.44269504088896341f), c), pk1(0.5f), c);
  const u64 t = pk2(floorf(lo2(fx)), floorf(hi2(fx)));
  const u64 tmp = mul2(t, pk1(0.693359375f), c);
  u64 z = mul2(t, pk1(-2.12194440e-4f), c);
  x = sub2(x, tmp, c);
  x = sub2(x, z, c);
  z = mul2(x, x, c);
  u64 y = pk1(1.9875691500E-4f);
  y = add2(mul2(y, x, c), pk1(1.3981999507E-3f), c);
  y = add2(mul2(y, x, c), pk1(8.3334519073E-3f), c);
  y = add2(mul2(y, x, c), pk1(4.1665795894E-2f), c);
  y = add2(mul2(y, x, c), pk1(1.6666665459E-1f), c);
  y = add2(mul2(y, x, c), pk1(5.0000001201E-1f), c);
  y = add2(mul2(y, z, c), x, c);
  y = add2(y, pk1(1.f), c);
  const u64 nb = add2(t, pk1(12582912.f), c);  // low mantissa bits = n (two's complement)
  const float s0 = __int_as_float((__float_as_int(lo2(nb)) << 23) + 0x3f800000);
  const float s1 = __int_as_float((__float_as_int(hi2(nb)) << 23) + 0x3f800000);
  return mul2(y, pk2(s0, s1), c);
}

// ---- 32-byte pieces of a row: eight floats as four register pairs, moved with one LDG.256 / STG.256 ---------
struct F8 {
  u64 p[4];
};
__device__ __forceinline__ F8 ld8_nc(const float* __restrict__ ptr) {  // read-only data (never written by this launch)
  F8 r;
  asm("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(r.p[0]), "=l"(r.p[1]), "=l"(r.p[2]), "=l"(r.p[3]) : "l"(ptr));
  return r;
}
__device__ __forceinline__ void st8(float* ptr, const F8& v) {
  asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(ptr), "l"(v.p[0]), "l"(v.p[1]), "l"(v.p[2]), "l"(v.p[3])
               : "memory");
}
__device__ __forceinline__ F8 zero8() {
  F8 r;
  r.p[0] = r.p[1] = r.p[2] = r.p[3] = 0ull;
  return r;
}
__device__ __forceinline__ float get8(const F8& v, int k) { return (k & 1) ? hi2(v.p[k >> 1]) : lo2(v.p[k >> 1]); }

}  // namespace dcrf
