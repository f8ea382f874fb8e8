// Learning-path kernels on the lattice: Permutohedral::gradient, the derivative of a^T K b with respect to
// the features (pydensecrf/densecrf/src/permutohedral.cpp:612-695 in /root/reference).  DenseKernel::kernelGradient /
// featureGradient (src/pairwise.cpp:82-114) are sums and products around it and around Permutohedral::compute(reverse).
//
// The reference does, for dir in {0, 1}: splat (dir ? b : a) -> blur the axes in ascending (dir 0) or descending (dir 1)
// order with seqCompute's arithmetic (0.5 is a double literal, :657) -> per pixel, "rotate" the blurred values of its
// simplex corners with the stored ranks, multiply by the elevation matrix and dot with (dir ? a : b), added to df.
// Here: the splat and the blurs are the filter's own kernels (k_splat without pre-normalisation walks the same (pixel,
// corner) order; launch_blur with seq_math), the per-pixel part is k_lattice_grad, one thread per pixel.  Every float
// operation is separately rounded and in the reference's order: the k loop is hoisted outside the j loop, which does not
// change any sum because every (j, k) entry of r_a and every sm[k] is touched by one k only and every gradient entry
// accumulates over k in ascending order either way.
#include <cmath>

#include "common.h"

namespace dcrf {
namespace {

constexpr int kThreads = 128;

struct GradParams {
  int N, d, L, Lp;
  float alpha;                    // 1 / (1 + 2^-d) / (d + 1), float arithmetic (:629)
  float scale[kMaxFeatureDims];   // scale_factor (:624-626)
};

__global__ void __launch_bounds__(kThreads)
    k_lattice_grad(GradParams p, const PixEntry* __restrict__ pix, const uint8_t* __restrict__ rank,
                   const float* __restrict__ val, const float* __restrict__ other, float* __restrict__ df) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.N) return;
  const int d = p.d, dp1 = d + 1;
  // rows of the two corners entry j of r_a combines: r0 = d - rank[j], r1 = r0 + 1 (wrapping to 0)  (:662-666)
  int o0[kMaxFeatureDims + 1], o1[kMaxFeatureDims + 1];
  {
    int last = d * dp1 / 2;
    for (int j = 0; j <= d; j++) {
      int rk;
      if (j < d) {
        rk = rank[(size_t)i * d + j];
        last -= rk;
      } else {
        rk = last;  // the ranks are a permutation of 0..d
      }
      const int r0 = d - rk, r1 = r0 + 1 > d ? 0 : r0 + 1;
      o0[j] = pix[(size_t)i * dp1 + r0].row;
      o1[j] = pix[(size_t)i * dp1 + r1].row;
    }
  }
  float g[kMaxFeatureDims];
  for (int j = 0; j < d; j++) g[j] = 0.f;
  for (int k = 0; k < p.L; k++) {
    const float ok = other[(size_t)i * p.Lp + k];
    // r_a[0][k], then sm = r_a[0][k]; for j >= 1: v = scale[j-1] * (sm - j * r_a[j][k]); grad_j += other * v; sm += r_a[j][k]
    float sm = __fsub_rn(__fadd_rn(0.f, __fmul_rn(p.alpha, val[(size_t)o0[0] * p.Lp + k])),
                         __fmul_rn(p.alpha, val[(size_t)o1[0] * p.Lp + k]));
    for (int j = 1; j <= d; j++) {
      const float ra = __fsub_rn(__fadd_rn(0.f, __fmul_rn(p.alpha, val[(size_t)o0[j] * p.Lp + k])),
                                 __fmul_rn(p.alpha, val[(size_t)o1[j] * p.Lp + k]));
      const float v = __fmul_rn(p.scale[j - 1], __fsub_rn(sm, __fmul_rn((float)j, ra)));
      g[j - 1] = __fadd_rn(g[j - 1], __fmul_rn(ok, v));
      sm = __fadd_rn(sm, ra);
    }
  }
  for (int j = 0; j < d; j++) df[(size_t)i * d + j] = __fadd_rn(df[(size_t)i * d + j], g[j]);  // df[i*d_+j-1] += grad
}

}  // namespace

void lattice_gradient(Ctx cx, const Lattice& lt, RowLayout rl, const float* a, const float* b, float* A, float* B, float* df) {
  DCRF_REQUIRE(lt.rank != nullptr, "this lattice was built without its rank table");
  const int d = lt.d;
  GradParams p;
  p.N = lt.N, p.d = d, p.L = rl.L, p.Lp = rl.Lp;
  p.alpha = 1.0f / (1 + powf(2, -d)) / (d + 1);
  const float inv_std_dev = (float)(std::sqrt(2.0 / 3.0) * (double)(d + 1));
  for (int i = 0; i < kMaxFeatureDims; i++)
    p.scale[i] = (i < d) ? (float)(1.0 / std::sqrt((double)((i + 2) * (i + 1))) * (double)inv_std_dev) : 0.f;
  DCRF_CUDA(cudaMemsetAsync(df, 0, (size_t)lt.N * d * sizeof(float), cx.stream));
  const size_t vbytes = ((size_t)lt.V + 1) * rl.Lp * sizeof(float);
  for (int dir = 0; dir < 2; dir++) {
    // values = new_values = 0, then splat a (dir 0) or b (dir 1)
    DCRF_CUDA(cudaMemsetAsync(A, 0, vbytes, cx.stream));
    DCRF_CUDA(cudaMemsetAsync(B, 0, vbytes, cx.stream));
    launch_splat(cx, lt, rl, dir ? b : a, nullptr, A);
    float *src = A, *dst = B;
    for (int t = 0; t <= d; t++) {
      launch_blur(cx, lt, rl, dir ? d - t : t, src, dst, /*seq_math=*/true);
      std::swap(src, dst);
    }
    LaunchScope ls(cx.prof, cx.stream, "lattice_grad");
    k_lattice_grad<<<(lt.N + kThreads - 1) / kThreads, kThreads, 0, cx.stream>>>(p, lt.pix, lt.rank, src, dir ? a : b, df);
  }
}

}  // namespace dcrf
