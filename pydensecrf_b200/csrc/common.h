// Internal declarations shared by the .cu files of libdcrf_b200.so.  Not part of the C-ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/dcrf_b200.h"

namespace dcrf {

// ---------------------------------------------------------------------------------------------
// errors: internal code throws, the extern "C" layer converts to dcrf_status + dcrf_last_error()
// ---------------------------------------------------------------------------------------------
struct Error : public std::runtime_error {
  dcrf_status code;
  Error(dcrf_status c, const std::string& what) : std::runtime_error(what), code(c) {}
};
[[noreturn]] void fail_cuda(cudaError_t e, const char* expr, const char* file, int line);
#define DCRF_CUDA(expr)                                                  \
  do {                                                                   \
    cudaError_t _e = (expr);                                             \
    if (_e != cudaSuccess) ::dcrf::fail_cuda(_e, #expr, __FILE__, __LINE__); \
  } while (0)
#define DCRF_REQUIRE(cond, msg)                                          \
  do {                                                                   \
    if (!(cond)) throw ::dcrf::Error(DCRF_EINVAL, (msg));                \
  } while (0)

// ---------------------------------------------------------------------------------------------
// row layout of every per-pixel / per-vertex value buffer: the L labels of a row are contiguous and
// padded to a whole number of vector chunks, so one row is C aligned chunks of VEC floats.
//   L == 1 -> VEC 1;  L == 2 -> VEC 2;  L >= 3 -> VEC 4 (float4 rows, Lp = ceil4(L))
// ---------------------------------------------------------------------------------------------
struct RowLayout {
  int L, Lp, vec, C;
  explicit RowLayout(int L_ = 1) : L(L_) {
    vec = (L_ <= 1) ? 1 : (L_ == 2 ? 2 : 4);
    C = (L_ + vec - 1) / vec;
    Lp = C * vec;
  }
  // L<=2 takes the reference's seqCompute arithmetic, otherwise sseCompute (permutohedral.cpp:596-604)
  bool seq() const { return L <= 2; }
};

// A slab of device memory a batch worker hands out by bumping a pointer (api.cu: BatchChunk).  While a thread has an
// arena installed, dev_alloc serves from it and dev_free of its pointers is a no-op; the slab is recycled as a whole when
// the chunk is done.  A chunk of a batch makes ~40 allocations of up to a gigabyte with sizes that depend on the image
// content; through the stream-ordered pool that meant fresh driver allocations for many calls (and seconds of them when
// several processes grow their pools at once), through the arena it costs nothing.
struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0;
  void* take(size_t bytes) {
    const size_t o = (off + 255) & ~(size_t)255;
    if (o + bytes > cap) return nullptr;
    off = o + bytes;
    return base + o;
  }
  bool owns(const void* p) const { return p >= (const void*)base && p < (const void*)(base + cap); }
};
extern thread_local Arena* t_arena;  // api.cu
// the library's own stream-ordered pool of the calling thread's current device (created on first use; freed blocks stay
// cached, so steady-state allocations cost no driver call).  The device's default pool, which other libraries in the
// process may use, is left alone.
cudaMemPool_t device_pool();
void trim_device_caches();  // what dcrf_empty_cache does for the current device (called once before an allocation gives up)

// stream-ordered device memory from device_pool()
template <typename T>
T* dev_alloc(size_t n, cudaStream_t s) {
  void* p = nullptr;
  if (n == 0) n = 1;
  if (t_arena && (p = t_arena->take(n * sizeof(T))) != nullptr) return static_cast<T*>(p);
  cudaError_t e = cudaMallocFromPoolAsync(&p, n * sizeof(T), device_pool(), s);
  if (e == cudaErrorMemoryAllocation) {  // return every cached block to the driver and try once more
    cudaGetLastError();
    trim_device_caches();
    e = cudaMallocFromPoolAsync(&p, n * sizeof(T), device_pool(), s);
  }
  if (e == cudaErrorMemoryAllocation) {
    cudaGetLastError();
    throw Error(DCRF_ENOMEM, "out of device memory allocating " + std::to_string(n * sizeof(T)) + " bytes");
  }
  if (e != cudaSuccess) fail_cuda(e, "cudaMallocFromPoolAsync", __FILE__, __LINE__);
  return static_cast<T*>(p);
}
inline void dev_free(void* p, cudaStream_t s) {
  if (!p || (t_arena && t_arena->owns(p))) return;
  cudaFreeAsync(p, s);
}

// per-kernel-class profiling / launch accounting, owned by the model
struct Profiler {
  struct Rec {
    std::string name;
    int launches = 0;
    float ms = 0.f;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
  };
  bool enabled = false;
  int64_t launches = 0;
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> free_events;
  cudaEvent_t get_event();
  int index_of(const char* name);
  void collect();  // requires the stream to be idle
  void destroy();
};

// RAII bracket around one kernel launch
struct LaunchScope {
  Profiler* p;
  cudaStream_t s;
  int idx = -1;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  LaunchScope(Profiler* prof, cudaStream_t st, const char* name) : p(prof), s(st) {
    p->launches++;
    if (p->enabled) {
      idx = p->index_of(name);
      e0 = p->get_event();
      e1 = p->get_event();
      cudaEventRecord(e0, s);
    }
  }
  ~LaunchScope() {
    if (idx >= 0) {
      cudaEventRecord(e1, s);
      p->recs[idx].pending.emplace_back(e0, e1);
      p->recs[idx].launches++;
    }
  }
};

struct Ctx {  // what every launcher needs
  cudaStream_t stream;
  Profiler* prof;
};

// ---------------------------------------------------------------------------------------------
// lattice (lattice.cu)
// ---------------------------------------------------------------------------------------------
struct PixEntry {  // one corner of a pixel's enclosing simplex
  int row;         // vertex id + 1 (row 0 of a value buffer is the all-zero "absent" row)
  float w;         // barycentric weight
};
struct CsrEntry {  // one contribution to a vertex, rows sorted by (pixel, corner) ascending
  int pix;
  float w;
};

enum FeatureKind { FEAT_GENERIC = 0, FEAT_GAUSS2D = 1, FEAT_BILATERAL2D = 2, FEAT_ND = 3 };
static const int kMaxNdAxes = 6;
struct FeatureSource {
  int kind;
  const float* feat;  // device, d x N row-major (FEAT_GENERIC)
  const uint8_t* rgb; // device, N x 3 (FEAT_BILATERAL2D)
  int W;
  int H;              // image height when the variables are a batch of images stacked along y (0: one image)
  float s[5];         // sx, sy, sr, sg, sb
  // FEAT_ND: what utils.create_pairwise_gaussian / create_pairwise_bilateral build on the host (utils.py:89-162):
  // nsp spatial coordinates (C order over dims[]) / sdims, then nch channels of a channel-last image / schan
  int nsp, nch, img_is_float;
  int dims[kMaxNdAxes];
  int stride[kMaxNdAxes];
  float sdims[kMaxNdAxes];
  float schan[kMaxNdAxes];
  const void* img;    // device, N x nch, uint8 or float32
};

struct Lattice {
  int N = 0, d = 0, V = 0, Npad = 0;
  int64_t E = 0;               // N*(d+1) contributions
  PixEntry* pix = nullptr;     // [N*(d+1)]
  int* csr_start = nullptr;    // [V+1]
  CsrEntry* csr = nullptr;     // [E]
  int2* nbr = nullptr;         // [(d+1)*V] rows (id+1) of the -/+ neighbours, 0 = absent
  int* order = nullptr;        // [V] splat work order: windows of 4096 vertex ids, inside a window by CSR row length
  int16_t* keys = nullptr;     // [V*d]
  int* vimg = nullptr;         // [V] image index of every vertex (batched lattices only)
  uint8_t* rank = nullptr;     // [Npad*d] rank of the first d residuals of every pixel (rank_ of the reference; the last one
                               // is what is missing from the permutation 0..d); read by the gradient only
  float alpha = 1.f;           // 1/(1+2^-d)
  void release(cudaStream_t s);
};

static const int kMaxFeatureDims = 32;
// npi > 0: the N variables are N / npi images of npi pixels; the lattices of the images are built side by side (the image
// index is an extra key coordinate) and come out concatenated with id offsets, identical to separate builds
void lattice_build(Ctx cx, const FeatureSource& src, int d, int N, Lattice* out, int npi = 0);

// the features of `src` (optionally transformed by kernel parameters, see lattice.cu) as a d x N row-major device array
void materialize_features(Ctx cx, const FeatureSource& src, int d, int N, int ktype, const float* params, float* out_dxN);

// ---------------------------------------------------------------------------------------------
// filter / update kernels (filter.cu)
// ---------------------------------------------------------------------------------------------
// out rows (V+1) x Lp.  `in` is N x Lp; pre_norm (may be null) scales row i by norm[i] first.
void launch_splat(Ctx cx, const Lattice& lt, RowLayout rl, const float* in, const float* pre_norm, float* val);
// one Jacobi blur pass along axis j: dst = src + 0.5*(src[n1] + src[n2])
// seq_math: seqCompute's arithmetic (0.5 as a double literal, permutohedral.cpp:505 / :657) whatever the channel count
void launch_blur(Ctx cx, const Lattice& lt, RowLayout rl, int axis, const float* src, float* dst, bool seq_math = false);
// axis `axis` then axis `axis2` in one launch (bit-identical to two launch_blur calls)
void launch_blur2(Ctx cx, const Lattice& lt, RowLayout rl, int axis, int axis2, const float* src, float* dst, bool seq_math = false);

enum SliceMode { SLICE_WRITE = 0, SLICE_ACC_FIRST = 1, SLICE_ACC_NEXT = 2 };
struct CompatDev {
  int kind;          // dcrf_compat_kind; MATRIX is handled outside the slice/update kernels
  float neg_w;       // Potts: -w
  const float* diag; // Diagonal: Lp floats on the device (pads 0)
};
// standalone slice: m = alpha-weighted gather of val rows, optional post-norm, optional Potts/Diagonal;
//   SLICE_WRITE      out  = m
//   SLICE_ACC_FIRST  out  = (neg_unary ? -U : 0-U) - m       (U may be null -> treated as 0)
//   SLICE_ACC_NEXT   out  = out - m
void launch_slice(Ctx cx, const Lattice& lt, RowLayout rl, const float* val, const float* post_norm,
                  const CompatDev* compat, int mode, const float* U, bool neg_unary, float* out);

static const int kMaxFusedTerms = 8;
struct FusedTerm {
  const float* val;
  const PixEntry* pix;
  const float* post_norm;
  const float* diag;
  float neg_w;
  float alpha;
  int dp1;
  int ckind;
};
struct FusedArgs {
  int N, nterms;
  int W = 0, H = 0;  // image size of a 2-D model (pixel i = y*W + x), 0 = unknown: lets the update kernel walk 2-D tiles
  const float* U;  // may be null (no unary)
  int neg_unary;   // 1: start from -U (inference), 0: start from 0-U (stepInference)
  float* Q;
  float c_one, c_nzero, c_mone;  // 1, -0, -1: filled in by launch_fused_update (opaque operands of the packed kernel)
  // Pre-scaled copies of the new Q for the next iteration's splats: qn[j][i] = Q[i] * qn_norm[j][i], the product
  // DenseKernel::filter forms first (pairwise.cpp:65-66), computed once per pixel here instead of once per
  // contribution in the splat.  write_q = 0 skips storing Q itself (an inner iteration nobody reads Q of).
  int write_q = 1, nqn = 0;
  float* qn[kMaxFusedTerms];
  const float* qn_norm[kMaxFusedTerms];
  FusedTerm t[kMaxFusedTerms];
};
// arithmetic self-tests (see k_selftest); *bad (device) receives the number of mismatching elements
void launch_selftest(Ctx cx, int which, unsigned long long n, unsigned long long seed, unsigned long long* bad);
bool fused_update_supported(RowLayout rl);
// Q = softmax_col( -U - sum_k compat_k(norm_k * slice_k) ), everything in one pass
void launch_fused_update(Ctx cx, RowLayout rl, const FusedArgs& a, bool fast);
// wide-lane build of the same kernel; false = no instantiation for this shape (take launch_fused_update)
bool launch_fused_update8(Ctx cx, RowLayout rl, const FusedArgs& a);
// Q = softmax_col(acc) for any L (staged path)
void launch_softmax(Ctx cx, RowLayout rl, int N, const float* acc, float* Q);
// acc = -U / 0-U / -0 (no terms at all)
void launch_init_acc(Ctx cx, RowLayout rl, int N, const float* U, bool neg_unary, float* acc);
// out(mode) <- W * msg per pixel, W row-major LxL on the device; Eigen's GEBP summation order
void launch_matrix_compat(Ctx cx, RowLayout rl, int N, const float* W, const float* msg, int mode, const float* U,
                          bool neg_unary, float* out);

// compatibility GEMMs of all terms + unary + softmax in one tensor-core kernel when every term has a
// MatrixCompatibility (matrix_tc.cu)
static const int kMaxTCUTerms = 4;
struct TCUTerm {
  const float* msg;  // N x Lp: the term's normalised slice (PairwisePotential::apply before the compatibility)
  const float* Bhi;  // Np x Kp: tf32(W)
  const float* Blo;  // Np x Kp: W - tf32(W)
};
struct TCUArgs {
  int N, L, Lp, Np, Kp, neg_unary, nterms;
  const float* U;  // may be null
  float* Q;
  TCUTerm t[kMaxTCUTerms];
};
bool update_matrix_tc_supported(RowLayout rl, int nterms);
void launch_update_matrix_tc(Ctx cx, RowLayout rl, TCUArgs a);
// the same update fed by TMA with a thread-per-row epilogue out of TMEM (matrix_tc_tma.cu); launch_update_matrix_tc
// dispatches to it when it is supported
bool update_matrix_tc_tma_supported(RowLayout rl, int nterms);
void launch_update_matrix_tc_tma(Ctx cx, RowLayout rl, TCUArgs a);

// tensor-core (tcgen05, 3xTF32) MatrixCompatibility (matrix_tc.cu)
bool matrix_tc_supported(RowLayout rl);
void matrix_tc_dims(RowLayout rl, int* Np, int* Kp);
void launch_split_w(Ctx cx, RowLayout rl, const float* W, float* hi, float* lo);
void launch_matrix_compat_tc(Ctx cx, RowLayout rl, int N, const float* Bhi, const float* Blo, const float* msg, int mode,
                             const float* U, bool neg_unary, float* out);

// Permutohedral::gradient (permutohedral.cpp:612-695), learn.cu: df (N x d, pixel-major, device) = d/df a^T K b for the
// pixel-major N x Lp row arrays a and b; A and B are two zeroed (V+1) x Lp scratch buffers
void lattice_gradient(Ctx cx, const Lattice& lt, RowLayout rl, const float* a, const float* b, float* A, float* B, float* df);

// DenseCRF::gradient building blocks (learn.cu).  G[r*rb + c] (diag: G[r], ra == rb) = sum over the N variables of
// A[i*sai + r*sar] * B[i*sbi + c*sbr], accumulated in double in a fixed order (G: device doubles)
void launch_outer_sums(Ctx cx, int N, const float* A, size_t sai, size_t sar, int ra, const float* B, size_t sbi, size_t sbr,
                       int rb, bool diag, double* G);
// LogLikelihood / Hamming (objective.cpp:37-50, :67-80): b = d_mul_Q, *value (device) = the objective
void launch_objective_point(Ctx cx, RowLayout rl, int N, int kind, const int16_t* gt, float robust, const float* cw, const float* Q,
                            float* b, double* value);
// IntersectionOverUnion (objective.cpp:83-109): rows whose column sums are in[] and un[] - 1e-20; then d_mul_Q from them
void launch_iou_rows(Ctx cx, RowLayout rl, int N, const int16_t* gt, const float* Q, float* in_rows, float* un_rows);
void launch_iou_b(Ctx cx, RowLayout rl, int N, const int16_t* gt, const float* Q, const double* in, const double* un, float* b);
// sumAndNormalize (densecrf.cpp:107-114) on x = in (mul false) or in .* Q
void launch_sum_and_normalize(Ctx cx, RowLayout rl, int N, const float* in, const float* Q, bool mul, float* out);
void launch_scale_rows(Ctx cx, RowLayout rl, int N, const float* x, const float* norm, float* out);   // x * norm.asDiagonal()
void launch_compat_pointwise(Ctx cx, RowLayout rl, int N, const CompatDev& c, const float* x, float* out);  // Potts / Diagonal apply
void launch_add_rows(Ctx cx, RowLayout rl, int N, const float* x, float* acc);                          // acc += x
// the first argument of featureGradient's inner kernelGradient for the three normalised cases (pairwise.cpp:96-97, :105-106, :113-114)
void launch_feature_gradient_input(Ctx cx, RowLayout rl, int N, int ntype, const float* a, const float* fb, const float* fa,
                                   const float* b, const float* norm, float* out);
void launch_neg_add(Ctx cx, size_t n, const float* r, float* kg);  // kg = -r + kg

// layout conversion at the boundary
// batch > 1: `batch` images, each with its own (L, N) planes and N consecutive rows
void launch_planes_to_rows(Ctx cx, RowLayout rl, int N, const float* planes_LxN, float* rows_NxLp, int batch = 1);
// same transposition with utils.unary_from_softmax (utils.py:55-79) applied on the way: -log(clip(scale*p + (1-scale)/L))
void launch_softmax_planes_to_unary(Ctx cx, RowLayout rl, int N, const float* planes_LxN, double scale, double clip,
                                    float* rows_NxLp);
void launch_unpack_rows(Ctx cx, RowLayout rl, int N, const float* packed_NxL, float* rows_NxLp);
void launch_pack_rows(Ctx cx, RowLayout rl, int N, const float* rows_NxLp, float* packed_NxL);
void launch_fill(Ctx cx, float* p, size_t n, float v);
// out[e] = {csr[e].pix, csr[e].w * norm[csr[e].pix]}
void launch_bake_csr(Ctx cx, size_t E, const CsrEntry* csr, const float* norm, CsrEntry* out);
// norm_raw -> 1/sqrt(x+1e-20) (symmetric) or 1/(x+1e-20), evaluated in double (pairwise.cpp:52-61)
void launch_norm_transform(Ctx cx, int N, float* norm, bool symmetric);
// U = Lmat * f for the logistic unary (unary.cpp:50-52); rows_NxLp out
void launch_logistic_unary(Ctx cx, RowLayout rl, int N, int F, const float* Lmat_LxF, const float* f_FxN, float* U);
// sums[0] += sum Q log max(Q,1e-20), sums[1] += sum U*Q (T == null) or sums[2] += sum Q*T: double accumulators on the
// device, fixed reduction tree through `part` (kl_partials() doubles): deterministic
void launch_kl_terms(Ctx cx, RowLayout rl, int N, const float* Q, const float* U, const float* T, double* sums, double* part);
size_t kl_partials();
// *sum += (Q.array() * T.array()).sum() as the reference's float Eigen reduction evaluates it (bit-exact, sequential)
void launch_kl_pair_exact(Ctx cx, RowLayout rl, int N, const float* Q, const float* T, double* sum);
void launch_argmax(Ctx cx, RowLayout rl, int N, const float* Q, int16_t* labels);
void launch_argmax64(Ctx cx, RowLayout rl, int N, const float* Q, long long* labels);

}  // namespace dcrf
