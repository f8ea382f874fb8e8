// "Wide-lane" variants of the three kernel families of one mean-field iteration (sm_100a): every lane owns a
// 32-byte piece (eight floats) of a row and moves it with ONE 256-bit access (LDG.E.256 / STG.E.256, new on
// sm_100), so a 21-label row (24 floats, 96 bytes) is three lanes instead of six.  Same arithmetic, same order,
// same bits as the float4 kernels in filter.cu (packed FFMA2 with opaque -0 / 1 / -1 operands, packed_math.cuh);
// what changes is the instruction count per byte: index loads, address arithmetic, weight products and the
// shuffles of the softmax are per lane, and there are half as many lanes.
//
// Replaces (paths relative to /root/reference/pydensecrf/densecrf/), exactly like filter.cu:
//   Permutohedral::sseCompute   src/permutohedral.cpp:545-553 (splat), :556-569 (blur), :571-584 (slice)
//   DenseKernel::filter         src/pairwise.cpp:63-80
//   Potts/DiagonalCompatibility src/labelcompatibility.cpp:46-48,66-69
//   expAndNormalize             src/densecrf.cpp:98-106
// Rows must be a whole number of 32-byte pieces (RowLayout::Lp % 8 == 0: 5-8, 13-16, 21-24, ... labels).
#include <cfloat>
#include <cstdlib>

#include "common.h"
#include "packed_math.cuh"
#include "softmax_math.cuh"

#ifndef UPD8_MINBLOCKS
#define UPD8_MINBLOCKS 6
#endif
#ifndef UPD8_GB
#define UPD8_GB 2  // vertex rows in flight per gather batch
#endif
#ifndef UPD8_STRIP_ROWS
#define UPD8_STRIP_ROWS 4
#endif

namespace dcrf {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ PixEntry ld_pix(const PixEntry* __restrict__ p) {
  const int2 v = __ldg(reinterpret_cast<const int2*>(p));
  PixEntry e;
  e.row = v.x;
  e.w = __int_as_float(v.y);
  return e;
}

// ---- fused update (k_fused_update4 in filter.cu) -------------------------------------------------------------
// C = Lp/8 lanes own one pixel (3 for 21 labels, 10 pixels per warp).  Lane h holds elements 8h .. 8h+7, i.e. the
// Eigen packets 2h and 2h+1 of the column, which is exactly how Eigen's unrolled redux pairs them (accumulator p0
// takes the even packets, p1 the odd ones): lane 0 adds the other lanes' eight floats element by element, then
// p0 + p1, the odd last packet, (a0+a2)+(a1+a3) and the scalar tail -- the same sum in the same order.
template <int DP1, int GB>
__device__ __forceinline__ void slice8(F8& m, const PixEntry* __restrict__ pix, float alpha, const float* __restrict__ val,
                                       int Lp, int coff, const Pk& c) {
  m = zero8();
#pragma unroll
  for (int j0 = 0; j0 < DP1; j0 += GB) {
    PixEntry pe[GB];
    F8 x[GB];
#pragma unroll
    for (int u = 0; u < GB; u++)
      if (j0 + u < DP1) pe[u] = ld_pix(pix + j0 + u);
#pragma unroll
    for (int u = 0; u < GB; u++)
      if (j0 + u < DP1) x[u] = ld8_nc(val + (size_t)pe[u].row * Lp + coff);
#pragma unroll
    for (int u = 0; u < GB; u++)
      if (j0 + u < DP1) {
        const u64 wa = pk1(__fmul_rn(pe[u].w, alpha));
#pragma unroll
        for (int k = 0; k < 4; k++) m.p[k] = add2(m.p[k], mul2(wa, x[u].p[k], c), c);
      }
  }
}
__device__ __forceinline__ void finish8(F8& acc, F8& m, const FusedTerm& ft, int i, int coff, const Pk& c) {
  if (ft.post_norm) {  // out = out * diag(norm) (pairwise.cpp:78-79)
    const u64 n = pk1(__ldg(ft.post_norm + i));
#pragma unroll
    for (int k = 0; k < 4; k++) m.p[k] = mul2(m.p[k], n, c);
  }
  if (ft.ckind == DCRF_COMPAT_POTTS) {
    const u64 w = pk1(ft.neg_w);
#pragma unroll
    for (int k = 0; k < 4; k++) m.p[k] = mul2(w, m.p[k], c);
  } else if (ft.ckind == DCRF_COMPAT_DIAGONAL) {
    const F8 w = ld8_nc(ft.diag + coff);
#pragma unroll
    for (int k = 0; k < 4; k++) m.p[k] = mul2(w.p[k], m.p[k], c);
  }
#pragma unroll
  for (int k = 0; k < 4; k++) acc.p[k] = sub2(acc.p[k], m.p[k], c);  // tmp1 -= tmp2
}

template <int DPA, int DPB, int LT>
__global__ void __launch_bounds__(kThreads, UPD8_MINBLOCKS) k_fused_update8(FusedArgs a) {
  constexpr int L = LT, Lp = ((LT + 7) / 8) * 8, C = Lp / 8, P = 32 / C;
  constexpr int NF = L / 4, NT = L % 4, NPAIR = NF / 2;
  constexpr int kTailLane = (4 * NF) / 8, kTailOff = (4 * NF) % 8;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = lane / C, h = lane - grp * C, base = grp * C;
  const int coff = h * 8;
  Pk c;
  c.one = pk1(a.c_one), c.nzero = pk1(a.c_nzero), c.mone = pk1(a.c_mone);
  const int strips_x = a.W > 0 ? (a.W + P - 1) / P : 1;
  const int sx = a.W > 0 ? (int)(blockIdx.x % (unsigned)strips_x) : 0;
  const int sy = a.W > 0 ? (int)(blockIdx.x / (unsigned)strips_x) : 0;
  const int passes = a.W > 0 ? UPD8_STRIP_ROWS : 1;
  for (int pass = 0; pass < passes; pass++) {
    bool act;
    unsigned gi;
    if (a.W > 0) {
      const int x = sx * P + grp, y = (sy * UPD8_STRIP_ROWS + pass) * (kThreads / 32) + warp;
      act = (grp < P) && x < a.W && y < a.H;
      gi = (unsigned)y * (unsigned)a.W + (unsigned)x;
    } else {
      const unsigned warp_g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
      gi = warp_g * (unsigned)P + (unsigned)grp;
      act = (grp < P) && (gi < (unsigned)a.N);
    }
    const unsigned i = act ? gi : 0u;
    const unsigned off = i * (unsigned)Lp + (unsigned)coff;  // host guarantees N*Lp < 2^31

    F8 acc = zero8();
    if (act) {
      const u64 z = pk1(a.neg_unary ? -0.f : 0.f);  // (-0) - u == -u bit for bit
      if (a.U) {
        const F8 u = ld8_nc(a.U + off);
#pragma unroll
        for (int k = 0; k < 4; k++) acc.p[k] = sub2(z, u.p[k], c);
      } else {
#pragma unroll
        for (int k = 0; k < 4; k++) acc.p[k] = z;
      }
      if constexpr (DPA > 0) {
        F8 m;
        slice8<DPA, UPD8_GB>(m, a.t[0].pix + i * (unsigned)DPA, a.t[0].alpha, a.t[0].val, Lp, coff, c);
        finish8(acc, m, a.t[0], (int)i, coff, c);
        if constexpr (DPB > 0) {
          slice8<DPB, UPD8_GB>(m, a.t[1].pix + i * (unsigned)DPB, a.t[1].alpha, a.t[1].val, Lp, coff, c);
          finish8(acc, m, a.t[1], (int)i, coff, c);
        }
      }
    }

    bool real[8];
#pragma unroll
    for (int k = 0; k < 8; k++) real[k] = act && coff + k < L;
    // column max (exact, order-free)
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 8; k++)
      if (real[k]) mx = fmaxf(mx, get8(acc, k));
    {
      const float mine = mx;
#pragma unroll
      for (int k = 0; k < C; k++) mx = fmaxf(mx, __shfl_sync(full, mine, base + k));
    }
    const u64 mx2 = pk1(mx);
#pragma unroll
    for (int k = 0; k < 4; k++) acc.p[k] = sub2(acc.p[k], mx2, c);
    float xs[8];
#pragma unroll
    for (int k = 0; k < 8; k++) xs[k] = get8(acc, k);
    float lo = xs[0];
#pragma unroll
    for (int k = 1; k < 8; k++) lo = fminf(lo, xs[k]);
    const bool small = lo < -44.f;  // keeps the shared-reciprocal division in range (pads may trip it: harmless)

    float b[8];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const u64 e = eigen_pexp_nonpos2(acc.p[k], c);
      b[2 * k] = lo2(e), b[2 * k + 1] = hi2(e);
    }
#pragma unroll
    for (int k = 0; k < 8; k++)
      if (!real[k]) b[k] = 0.f;
    if (NT && act && h == kTailLane) {
#pragma unroll
      for (int k = 0; k < NT; k++) b[kTailOff + k] = tail_exp_nonpos(xs[kTailOff + k]);
    }
    // Eigen's redux order (Eigen/src/Core/Redux.h:213-262), see the comment above
    float s;
    if constexpr (NF == 0) {
      s = b[0];
      if (L > 1) s = __fadd_rn(s, b[1]);
      if (L > 2) s = __fadd_rn(s, b[2]);
    } else {
      u64 p8[4] = {pk2(b[0], b[1]), pk2(b[2], b[3]), pk2(b[4], b[5]), pk2(b[6], b[7])};
#pragma unroll
      for (int m = 1; m < NPAIR; m++) {
        const int src = base + m;
#pragma unroll
        for (int k = 0; k < 4; k++)
          p8[k] = add2(p8[k], pk2(__shfl_sync(full, b[2 * k], src), __shfl_sync(full, b[2 * k + 1], src)), c);
      }
      u64 p01 = p8[0], p23 = p8[1];
      if constexpr (NF > 1) {
        p01 = add2(p8[0], p8[2], c), p23 = add2(p8[1], p8[3], c);  // p0 + p1
        if constexpr (NF & 1) {
          const int src = base + NPAIR;  // the odd last packet is the first half of lane NPAIR
          p01 = add2(p01, pk2(__shfl_sync(full, b[0], src), __shfl_sync(full, b[1], src)), c);
          p23 = add2(p23, pk2(__shfl_sync(full, b[2], src), __shfl_sync(full, b[3], src)), c);
        }
      }
      const u64 hs = add2(p01, p23, c);  // (a0 + a2, a1 + a3)
      s = __fadd_rn(lo2(hs), hi2(hs));
#pragma unroll
      for (int k = 0; k < NT; k++) s = __fadd_rn(s, __shfl_sync(full, b[kTailOff + k], base + kTailLane));
      s = __shfl_sync(full, s, base);
    }

    if (act) {
      float qv[8];
      if (!small) {  // s in [1, L]: numerators are 0 (pads) or >= e^-44
        const float r = refined_rcp(s);
#pragma unroll
        for (int k = 0; k < 8; k++) qv[k] = (coff + k < L) ? div_shared(b[k], s, r) : 0.f;
      } else {
#pragma unroll
        for (int k = 0; k < 8; k++) qv[k] = (coff + k < L) ? __fdiv_rn(b[k], s) : 0.f;  // out.col = b / sum
      }
      F8 o;
#pragma unroll
      for (int k = 0; k < 4; k++) o.p[k] = pk2(qv[2 * k], qv[2 * k + 1]);
      if (a.write_q) st8(a.Q + off, o);
      for (int j = 0; j < a.nqn; j++) {  // Q * diag(norm_j) for term j's next splat
        const u64 n = pk1(__ldg(a.qn_norm[j] + i));
        F8 qs;
#pragma unroll
        for (int k = 0; k < 4; k++) qs.p[k] = mul2(o.p[k], n, c);
        st8(a.qn[j] + off, qs);
      }
    }
  }
}

}  // namespace

bool wide_rows_supported(RowLayout rl) { return rl.vec == 4 && rl.Lp % 8 == 0 && rl.Lp / 8 <= 32; }

// Returns false when this model shape has no wide instantiation (the caller then takes launch_fused_update).
bool launch_fused_update8(Ctx cx, RowLayout rl, const FusedArgs& a_in) {
  if (rl.L != 21 || rl.Lp != 24) return false;
  const int d0 = a_in.nterms > 0 ? a_in.t[0].dp1 : 0, d1 = a_in.nterms > 1 ? a_in.t[1].dp1 : 0;
  int shape = -1;
  if (a_in.nterms == 0) shape = 0;
  else if (a_in.nterms == 2 && d0 == 3 && d1 == 6) shape = 1;
  else if (a_in.nterms == 1 && d0 == 3) shape = 2;
  else if (a_in.nterms == 1 && d0 == 6) shape = 3;
  if (shape < 0) return false;
  LaunchScope ls(cx.prof, cx.stream, "fused_update");
  FusedArgs a = a_in;
  a.c_one = 1.f, a.c_nzero = -0.f, a.c_mone = -1.f;
  constexpr int P = 10;
  if (a.W <= 0 || (size_t)a.W * a.H != (size_t)a.N) a.W = a.H = 0;
  const int strip_h = (kThreads / 32) * UPD8_STRIP_ROWS;
  const int ppb = (kThreads / 32) * P;
  const unsigned grid = a.W > 0 ? (unsigned)(((a.W + P - 1) / P) * ((a.H + strip_h - 1) / strip_h))
                                : (unsigned)(((size_t)a.N + ppb - 1) / ppb);
  switch (shape) {
    case 0: k_fused_update8<0, 0, 21><<<grid, kThreads, 0, cx.stream>>>(a); break;
    case 1: k_fused_update8<3, 6, 21><<<grid, kThreads, 0, cx.stream>>>(a); break;
    case 2: k_fused_update8<3, 0, 21><<<grid, kThreads, 0, cx.stream>>>(a); break;
    default: k_fused_update8<6, 0, 21><<<grid, kThreads, 0, cx.stream>>>(a); break;
  }
  return true;
}

}  // namespace dcrf
