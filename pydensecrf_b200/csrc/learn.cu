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

template <bool VEC4>
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
  if (VEC4) {
    // four labels per row access (rows are whole float4 chunks for L >= 3); every label still runs the same operations
    // in the same order, and a gradient entry still accumulates over the labels in ascending order
    for (int k4 = 0; k4 < p.Lp; k4 += 4) {
      const float4 ok4 = *reinterpret_cast<const float4*>(other + (size_t)i * p.Lp + k4);
      const float okc[4] = {ok4.x, ok4.y, ok4.z, ok4.w};
      float sm[4];
      {
        const float4 a = *reinterpret_cast<const float4*>(val + (size_t)o0[0] * p.Lp + k4);
        const float4 b = *reinterpret_cast<const float4*>(val + (size_t)o1[0] * p.Lp + k4);
        const float ac[4] = {a.x, a.y, a.z, a.w}, bc[4] = {b.x, b.y, b.z, b.w};
        for (int c = 0; c < 4; c++) sm[c] = __fsub_rn(__fadd_rn(0.f, __fmul_rn(p.alpha, ac[c])), __fmul_rn(p.alpha, bc[c]));
      }
      for (int j = 1; j <= d; j++) {
        const float4 a = *reinterpret_cast<const float4*>(val + (size_t)o0[j] * p.Lp + k4);
        const float4 b = *reinterpret_cast<const float4*>(val + (size_t)o1[j] * p.Lp + k4);
        const float ac[4] = {a.x, a.y, a.z, a.w}, bc[4] = {b.x, b.y, b.z, b.w};
        float gj = g[j - 1];
        for (int c = 0; c < 4; c++) {
          if (k4 + c >= p.L) break;
          const float ra = __fsub_rn(__fadd_rn(0.f, __fmul_rn(p.alpha, ac[c])), __fmul_rn(p.alpha, bc[c]));
          const float v = __fmul_rn(p.scale[j - 1], __fsub_rn(sm[c], __fmul_rn((float)j, ra)));
          gj = __fadd_rn(gj, __fmul_rn(okc[c], v));
          sm[c] = __fadd_rn(sm[c], ra);
        }
        g[j - 1] = gj;
      }
    }
  } else
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

// The same per-variable computation with the C = Lp/4 label chunks of a variable on C neighbouring lanes (32/C variables
// per warp): a row read is one contiguous piece per variable instead of one scattered float4 per thread (the thread-per-
// variable form above kept the L1 pipe 75 % busy).  Every lane forms the products of its four labels; the sum over the
// labels stays in ascending label order by handing the running sums from lane to lane.  D = d (<= 8: the products of all
// axes stay in registers).
template <int D>
__global__ void __launch_bounds__(kThreads)
    k_lattice_grad_rows(GradParams p, int C, const PixEntry* __restrict__ pix, const uint8_t* __restrict__ rank,
                        const float* __restrict__ val, const float* __restrict__ other, float* __restrict__ df) {
  constexpr int dp1 = D + 1;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, P = 32 / C;
  const int grp = lane / C, h = lane - grp * C, base = grp * C;
  const long long warp_g = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long gi = warp_g * P + grp;
  const bool act = grp < P && gi < p.N;
  const int i = act ? (int)gi : 0;
  int o0[dp1], o1[dp1];
  {
    int last = D * dp1 / 2;
#pragma unroll
    for (int j = 0; j <= D; j++) {
      int rk;
      if (j < D) {
        rk = rank[(size_t)i * D + j];
        last -= rk;
      } else {
        rk = last;
      }
      const int r0 = D - rk, r1 = r0 + 1 > D ? 0 : r0 + 1;
      o0[j] = pix[(size_t)i * dp1 + r0].row;
      o1[j] = pix[(size_t)i * dp1 + r1].row;
    }
  }
  const int k4 = 4 * h;
  float term[D][4];
  {
    const float4 ok4 = *reinterpret_cast<const float4*>(other + (size_t)i * p.Lp + k4);
    const float okc[4] = {ok4.x, ok4.y, ok4.z, ok4.w};
    float sm[4];
    {
      const float4 a = *reinterpret_cast<const float4*>(val + (size_t)o0[0] * p.Lp + k4);
      const float4 b = *reinterpret_cast<const float4*>(val + (size_t)o1[0] * p.Lp + k4);
      const float ac[4] = {a.x, a.y, a.z, a.w}, bc[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int c = 0; c < 4; c++) sm[c] = __fsub_rn(__fadd_rn(0.f, __fmul_rn(p.alpha, ac[c])), __fmul_rn(p.alpha, bc[c]));
    }
#pragma unroll
    for (int j = 1; j <= D; j++) {
      const float4 a = *reinterpret_cast<const float4*>(val + (size_t)o0[j] * p.Lp + k4);
      const float4 b = *reinterpret_cast<const float4*>(val + (size_t)o1[j] * p.Lp + k4);
      const float ac[4] = {a.x, a.y, a.z, a.w}, bc[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const float ra = __fsub_rn(__fadd_rn(0.f, __fmul_rn(p.alpha, ac[c])), __fmul_rn(p.alpha, bc[c]));
        const float v = __fmul_rn(p.scale[j - 1], __fsub_rn(sm[c], __fmul_rn((float)j, ra)));
        term[j - 1][c] = __fmul_rn(okc[c], v);
        sm[c] = __fadd_rn(sm[c], ra);
      }
    }
  }
  float g[D];
#pragma unroll
  for (int j = 0; j < D; j++) g[j] = 0.f;
  for (int st = 0; st < C; st++) {  // lane st of every variable adds its labels, then hands the sums on
    if (h == st) {
#pragma unroll
      for (int c = 0; c < 4; c++)
        if (k4 + c < p.L) {
#pragma unroll
          for (int j = 0; j < D; j++) g[j] = __fadd_rn(g[j], term[j][c]);
        }
    }
#pragma unroll
    for (int j = 0; j < D; j++) g[j] = __shfl_sync(full, g[j], base + st);
  }
  if (act && h == 0) {
#pragma unroll
    for (int j = 0; j < D; j++) df[(size_t)i * D + j] = __fadd_rn(df[(size_t)i * D + j], g[j]);
  }
}

template <int D>
void launch_grad_rows(cudaStream_t s, const GradParams& p, int C, const Lattice& lt, const float* val, const float* other, float* df) {
  const int P = 32 / C;
  const long long warps = ((long long)lt.N + P - 1) / P;
  const unsigned blocks = (unsigned)((warps * 32 + kThreads - 1) / kThreads);
  k_lattice_grad_rows<D><<<blocks, kThreads, 0, s>>>(p, C, lt.pix, lt.rank, val, other, df);
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
      const int j = dir ? d - t : t;
      if (t + 1 <= d) {  // two axes per launch, bit-identical to two single-axis passes
        launch_blur2(cx, lt, rl, j, dir ? j - 1 : j + 1, src, dst, /*seq_math=*/true);
        t++;
      } else {
        launch_blur(cx, lt, rl, j, src, dst, /*seq_math=*/true);
      }
      std::swap(src, dst);
    }
    LaunchScope ls(cx.prof, cx.stream, "lattice_grad");
    const float* other = dir ? a : b;
    const int C = rl.C;
    if (rl.vec == 4 && C <= 32 && d <= 8) {
      switch (d) {
        case 1: launch_grad_rows<1>(cx.stream, p, C, lt, src, other, df); break;
        case 2: launch_grad_rows<2>(cx.stream, p, C, lt, src, other, df); break;
        case 3: launch_grad_rows<3>(cx.stream, p, C, lt, src, other, df); break;
        case 4: launch_grad_rows<4>(cx.stream, p, C, lt, src, other, df); break;
        case 5: launch_grad_rows<5>(cx.stream, p, C, lt, src, other, df); break;
        case 6: launch_grad_rows<6>(cx.stream, p, C, lt, src, other, df); break;
        case 7: launch_grad_rows<7>(cx.stream, p, C, lt, src, other, df); break;
        default: launch_grad_rows<8>(cx.stream, p, C, lt, src, other, df); break;
      }
    } else if (rl.vec == 4)
      k_lattice_grad<true><<<(lt.N + kThreads - 1) / kThreads, kThreads, 0, cx.stream>>>(p, lt.pix, lt.rank, src, dir ? a : b, df);
    else
      k_lattice_grad<false><<<(lt.N + kThreads - 1) / kThreads, kThreads, 0, cx.stream>>>(p, lt.pix, lt.rank, src, dir ? a : b, df);
  }
}

}  // namespace dcrf

// ================================================================================================================
// DenseCRF::gradient (src/densecrf.cpp:237-297): the per-pixel and reduction kernels around the filters.  Everything
// works on the device row layout (N rows of Lp floats, one row per variable); sums over the variables are accumulated
// in double through a fixed two-level order (a block per chunk of variables, then the chunks in ascending order), so
// they are deterministic and at least as accurate as the reference's float Eigen reductions they stand for.
// ================================================================================================================
namespace dcrf {
namespace {

constexpr int kRedThreads = 256;
constexpr int kOutPerThread = 4;                       // outputs a thread accumulates per pass
constexpr int kOutPerPass = kRedThreads * kOutPerThread;
constexpr int kTileBytes = 40 * 1024;                  // shared-memory tile of A and B rows

// part[chunk][o] = sum over the chunk's variables i of A[i*sai + r*sar] * B[i*sbi + c*sbr], o = r*rb + c (diag: r = c = o)
__global__ void __launch_bounds__(kRedThreads)
    k_outer_partial(int N, int per_chunk, const float* __restrict__ A, size_t sai, size_t sar, int ra,
                    const float* __restrict__ B, size_t sbi, size_t sbr, int rb, int diag, int nout, int TP,
                    double* __restrict__ part) {
  extern __shared__ float sh[];
  float* As = sh;                      // TP x ra
  float* Bs = sh + (size_t)TP * ra;    // TP x rb
  const int chunk = blockIdx.x, pass = blockIdx.y;
  const int i0 = chunk * per_chunk, i1 = min(N, i0 + per_chunk);
  int r[kOutPerThread], c[kOutPerThread];
  double acc[kOutPerThread];
  for (int j = 0; j < kOutPerThread; j++) {
    const int o = pass * kOutPerPass + j * kRedThreads + threadIdx.x;
    acc[j] = 0.0;
    if (o < nout) {
      r[j] = diag ? o : o / rb;
      c[j] = diag ? o : o % rb;
    } else {
      r[j] = -1, c[j] = 0;
    }
  }
  for (int t0 = i0; t0 < i1; t0 += TP) {
    const int tp = min(TP, i1 - t0);
    __syncthreads();
    for (int e = threadIdx.x; e < tp * ra; e += kRedThreads) As[e] = A[(size_t)(t0 + e / ra) * sai + (size_t)(e % ra) * sar];
    for (int e = threadIdx.x; e < tp * rb; e += kRedThreads) Bs[e] = B[(size_t)(t0 + e / rb) * sbi + (size_t)(e % rb) * sbr];
    __syncthreads();
    for (int j = 0; j < kOutPerThread; j++) {
      if (r[j] < 0) continue;
      double a = acc[j];
      for (int p = 0; p < tp; p++) a += (double)As[p * ra + r[j]] * (double)Bs[p * rb + c[j]];
      acc[j] = a;
    }
  }
  for (int j = 0; j < kOutPerThread; j++) {
    const int o = pass * kOutPerPass + j * kRedThreads + threadIdx.x;
    if (o < nout) part[(size_t)chunk * nout + o] = acc[j];
  }
}
__global__ void k_outer_final(int nchunks, int nout, const double* __restrict__ part, double* __restrict__ G) {
  // a warp per output: lane l adds chunks l, l + 32, ... in ascending order, then a fixed shuffle tree (deterministic)
  const int o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (o >= nout) return;
  double s = 0.0;
  for (int ch = lane; ch < nchunks; ch += 32) s += part[(size_t)ch * nout + o];
  for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
  if (lane == 0) G[o] = s;
}

// one thread per variable ------------------------------------------------------------------------------------------
// LogLikelihood::evaluate (objective.cpp:37-50) and Hamming::evaluate (:67-80): b = d_mul_Q, term[i] = the variable's
// addend of the objective value (double)
__global__ void k_objective_point(int N, int L, int Lp, int kind, const int16_t* __restrict__ gt, float robust,
                                  const float* __restrict__ cw, const float* __restrict__ Q, float* __restrict__ b,
                                  double* __restrict__ term) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float* bi = b + (size_t)i * Lp;
  for (int l = 0; l < Lp; l++) bi[l] = 0.f;
  const int g = gt[i];
  double t = 0.0;
  if (g >= 0 && g < L) {
    const float q = Q[(size_t)i * Lp + g];
    if (kind == DCRF_OBJECTIVE_LOGLIKELIHOOD) {
      const float qq = fmaxf(__fadd_rn(q, robust), 1e-20f);
      t = log((double)qq) / N;
      bi[g] = __fdiv_rn(__fdiv_rn(q, qq), (float)N);
    } else {
      const float qq = __fmul_rn(cw[g], q);
      t = (double)qq;
      bi[g] = qq;
    }
  }
  term[i] = t;
}
// IntersectionOverUnion::evaluate (:83-109), first loop: rows whose column sums are in[] and un[] - 1e-20
__global__ void k_iou_rows(int N, int L, int Lp, const int16_t* __restrict__ gt, const float* __restrict__ Q,
                           float* __restrict__ in_rows, float* __restrict__ un_rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int g = gt[i];
  const bool ok = g >= 0 && g < L;
  for (int l = 0; l < Lp; l++) {
    const float q = l < L ? Q[(size_t)i * Lp + l] : 0.f;
    in_rows[(size_t)i * Lp + l] = (ok && l == g) ? q : 0.f;
    un_rows[(size_t)i * Lp + l] = !ok || l >= L ? 0.f : (l == g ? 1.f : q);
  }
}
// second loop: d_mul_Q from the class sums (double arithmetic, as the reference's VectorXd in / un make it)
__global__ void k_iou_b(int N, int L, int Lp, const int16_t* __restrict__ gt, const float* __restrict__ Q,
                        const double* __restrict__ in, const double* __restrict__ un, float* __restrict__ b) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int g = gt[i];
  const bool ok = g >= 0 && g < L;
  for (int l = 0; l < Lp; l++) {
    float v = 0.f;
    if (ok && l < L) {
      const double q = (double)Q[(size_t)i * Lp + l];
      const double u = un[l] + 1e-20;
      v = l == g ? (float)(q / (u * L)) : (float)(-q * in[l] / (u * u * L));
    }
    b[(size_t)i * Lp + l] = v;
  }
}
// sumAndNormalize (densecrf.cpp:107-114): out = sum(x) * q - x per variable, x = in (mul == 0) or in .* Q (:290)
__global__ void k_sum_and_normalize(int N, int L, int Lp, const float* __restrict__ in, const float* __restrict__ Q, int mul,
                                    float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float* x = in + (size_t)i * Lp;
  const float* q = Q + (size_t)i * Lp;
  float s = 0.f;
  for (int l = 0; l < L; l++) s = __fadd_rn(s, mul ? __fmul_rn(x[l], q[l]) : x[l]);
  float* o = out + (size_t)i * Lp;
  for (int l = 0; l < Lp; l++) {
    const float xl = l < L ? (mul ? __fmul_rn(x[l], q[l]) : x[l]) : 0.f;
    o[l] = l < L ? __fsub_rn(__fmul_rn(s, q[l]), xl) : 0.f;
  }
}

// one thread per element -------------------------------------------------------------------------------------------
enum { EW_SCALE_ROWS = 0, EW_POTTS = 1, EW_DIAG = 2, EW_ADD = 3, EW_SYM = 4, EW_PROD_NORM2 = 5 };
// EW_SCALE_ROWS: out = x * norm[i]            (a * norm_.asDiagonal(), pairwise.cpp:92)
// EW_POTTS:      out = s * x                  (-w_ * Q, labelcompatibility.cpp:47)
// EW_DIAG:       out = vec[l] * x             (w_.asDiagonal() * Q, :67)
// EW_ADD:        out = out + x                (tmp1 += tmp2, densecrf.cpp:288)
// EW_SYM:        out = 0.5 * (x*y + z*w) * norm^3   (pairwise.cpp:96-97)
// EW_PROD_NORM2: out = (x*y) * norm^2               (:105-106, :113-114)
__global__ void k_elementwise(int op, size_t n, int Lp, const float* __restrict__ x, const float* __restrict__ y,
                              const float* __restrict__ z, const float* __restrict__ w, const float* __restrict__ vec,
                              float s, float* __restrict__ out) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const size_t i = e / Lp;
  const int l = (int)(e % Lp);
  float r;
  switch (op) {
    case EW_SCALE_ROWS: r = __fmul_rn(x[e], vec[i]); break;
    case EW_POTTS: r = __fmul_rn(s, x[e]); break;
    case EW_DIAG: r = __fmul_rn(vec[l], x[e]); break;
    case EW_ADD: r = __fadd_rn(out[e], x[e]); break;
    case EW_SYM: {
      const float nn = vec[i], n3 = __fmul_rn(__fmul_rn(nn, nn), nn);
      r = __fmul_rn(__fmul_rn(0.5f, __fadd_rn(__fmul_rn(x[e], y[e]), __fmul_rn(z[e], w[e]))), n3);
      break;
    }
    default: {
      const float nn = vec[i];
      r = __fmul_rn(__fmul_rn(x[e], y[e]), __fmul_rn(nn, nn));
    }
  }
  out[e] = r;
}
// featureGradient's last line: fg = -r + kg (pairwise.cpp:98, :107, :115), in place on kg
__global__ void k_neg_add(size_t n, const float* __restrict__ r, float* __restrict__ kg) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) kg[e] = __fadd_rn(-r[e], kg[e]);
}
__global__ void k_sum_doubles(int n, const double* __restrict__ x, int per, int nparts, double* __restrict__ part) {
  // fixed order: thread t adds its contiguous slice, the slices are added in ascending order by k_outer_final
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nparts) return;
  const int i0 = t * per;
  double s = 0.0;
  for (int i = i0; i < min(n, i0 + per); i++) s += x[i];
  part[t] = s;
}

inline int blocks_for(size_t n, int threads) { return (int)((n + threads - 1) / threads); }

}  // namespace

void launch_outer_sums(Ctx cx, int N, const float* A, size_t sai, size_t sar, int ra, const float* B, size_t sbi, size_t sbr,
                       int rb, bool diag, double* G) {
  DCRF_REQUIRE(ra >= 1 && rb >= 1 && (size_t)(ra + rb) * sizeof(float) <= kTileBytes, "too many rows for the reduction kernel");
  const int nout = diag ? ra : ra * rb;
  const int TP = std::max(1, std::min(64, (int)(kTileBytes / ((size_t)(ra + rb) * sizeof(float)))));
  int nchunks = std::max(1, std::min(592, (N + 4 * TP - 1) / (4 * TP)));
  nchunks = (int)std::max<size_t>(1, std::min<size_t>((size_t)nchunks, ((size_t)256 << 20) / ((size_t)nout * sizeof(double))));  // partials <= 256 MB
  const int per_chunk = (N + nchunks - 1) / nchunks;
  nchunks = (N + per_chunk - 1) / per_chunk;
  double* part = dev_alloc<double>((size_t)nchunks * nout, cx.stream);
  {
    LaunchScope ls(cx.prof, cx.stream, "grad_outer_sums");
    dim3 grid(nchunks, (nout + kOutPerPass - 1) / kOutPerPass);
    k_outer_partial<<<grid, kRedThreads, (size_t)TP * (ra + rb) * sizeof(float), cx.stream>>>(N, per_chunk, A, sai, sar, ra, B, sbi,
                                                                                              sbr, rb, diag ? 1 : 0, nout, TP, part);
  }
  {
    LaunchScope ls(cx.prof, cx.stream, "grad_outer_final");
    k_outer_final<<<blocks_for((size_t)nout * 32, 128), 128, 0, cx.stream>>>(nchunks, nout, part, G);
  }
  dev_free(part, cx.stream);
}

void launch_objective_point(Ctx cx, RowLayout rl, int N, int kind, const int16_t* gt, float robust, const float* cw, const float* Q,
                            float* b, double* value) {
  double* term = dev_alloc<double>(N, cx.stream);
  {
    LaunchScope ls(cx.prof, cx.stream, "grad_objective");
    k_objective_point<<<blocks_for(N, 128), 128, 0, cx.stream>>>(N, rl.L, rl.Lp, kind, gt, robust, cw, Q, b, term);
  }
  const int per = 256, nt = (N + per - 1) / per;
  double* part = dev_alloc<double>(nt, cx.stream);
  {
    LaunchScope ls(cx.prof, cx.stream, "grad_objective_sum");
    k_sum_doubles<<<blocks_for(nt, 128), 128, 0, cx.stream>>>(N, term, per, nt, part);
    k_outer_final<<<1, 32, 0, cx.stream>>>(nt, 1, part, value);
  }
  dev_free(term, cx.stream);
  dev_free(part, cx.stream);
}
void launch_iou_rows(Ctx cx, RowLayout rl, int N, const int16_t* gt, const float* Q, float* in_rows, float* un_rows) {
  LaunchScope ls(cx.prof, cx.stream, "grad_objective");
  k_iou_rows<<<blocks_for(N, 128), 128, 0, cx.stream>>>(N, rl.L, rl.Lp, gt, Q, in_rows, un_rows);
}
void launch_iou_b(Ctx cx, RowLayout rl, int N, const int16_t* gt, const float* Q, const double* in, const double* un, float* b) {
  LaunchScope ls(cx.prof, cx.stream, "grad_objective");
  k_iou_b<<<blocks_for(N, 128), 128, 0, cx.stream>>>(N, rl.L, rl.Lp, gt, Q, in, un, b);
}
void launch_sum_and_normalize(Ctx cx, RowLayout rl, int N, const float* in, const float* Q, bool mul, float* out) {
  LaunchScope ls(cx.prof, cx.stream, "grad_sum_normalize");
  k_sum_and_normalize<<<blocks_for(N, 128), 128, 0, cx.stream>>>(N, rl.L, rl.Lp, in, Q, mul ? 1 : 0, out);
}
static void elementwise(Ctx cx, int op, size_t n, int Lp, const float* x, const float* y, const float* z, const float* w,
                        const float* vec, float s, float* out) {
  LaunchScope ls(cx.prof, cx.stream, "grad_elementwise");
  k_elementwise<<<blocks_for(n, 256), 256, 0, cx.stream>>>(op, n, Lp, x, y, z, w, vec, s, out);
}
void launch_scale_rows(Ctx cx, RowLayout rl, int N, const float* x, const float* norm, float* out) {
  elementwise(cx, EW_SCALE_ROWS, (size_t)N * rl.Lp, rl.Lp, x, nullptr, nullptr, nullptr, norm, 0.f, out);
}
void launch_compat_pointwise(Ctx cx, RowLayout rl, int N, const CompatDev& c, const float* x, float* out) {
  if (c.kind == DCRF_COMPAT_POTTS)
    elementwise(cx, EW_POTTS, (size_t)N * rl.Lp, rl.Lp, x, nullptr, nullptr, nullptr, nullptr, c.neg_w, out);
  else
    elementwise(cx, EW_DIAG, (size_t)N * rl.Lp, rl.Lp, x, nullptr, nullptr, nullptr, c.diag, 0.f, out);
}
void launch_add_rows(Ctx cx, RowLayout rl, int N, const float* x, float* acc) {
  elementwise(cx, EW_ADD, (size_t)N * rl.Lp, rl.Lp, x, nullptr, nullptr, nullptr, nullptr, 0.f, acc);
}
void launch_feature_gradient_input(Ctx cx, RowLayout rl, int N, int ntype, const float* a, const float* fb, const float* fa,
                                   const float* b, const float* norm, float* out) {
  const size_t n = (size_t)N * rl.Lp;
  if (ntype == DCRF_NORMALIZE_SYMMETRIC)
    elementwise(cx, EW_SYM, n, rl.Lp, a, fb, fa, b, norm, 0.f, out);
  else if (ntype == DCRF_NORMALIZE_AFTER)
    elementwise(cx, EW_PROD_NORM2, n, rl.Lp, a, fb, nullptr, nullptr, norm, 0.f, out);
  else
    elementwise(cx, EW_PROD_NORM2, n, rl.Lp, fa, b, nullptr, nullptr, norm, 0.f, out);
}
void launch_neg_add(Ctx cx, size_t n, const float* r, float* kg) {
  LaunchScope ls(cx.prof, cx.stream, "grad_elementwise");
  k_neg_add<<<blocks_for(n, 256), 256, 0, cx.stream>>>(n, r, kg);
}

}  // namespace dcrf
