// Host driver and C-ABI of libdcrf_b200.so (include/dcrf_b200.h).
//
// Replaces the model layer of the reference (paths relative to /root/reference/pydensecrf/densecrf/):
//   DenseCRF / DenseCRF2D            src/densecrf.cpp:39-131,178-235
//   DenseKernel::initLattice/filter  src/pairwise.cpp:40-80
//   PairwisePotential::apply         src/pairwise.cpp:173-178
//   Const/LogisticUnaryEnergy        src/unary.cpp:42-52
// and the Eigen-backed matrix objects of pydensecrf/eigen.pyx.  Everything lattice- or label-sized
// lives in HBM for the life of the model; one CUDA stream per model; no CPU compute path.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>

#include "common.h"

namespace dcrf {

static thread_local std::string g_last_error;
thread_local Arena* t_arena = nullptr;

void fail_cuda(cudaError_t e, const char* expr, const char* file, int line) {
  cudaGetLastError();  // clear sticky-less errors
  const char* base = std::strrchr(file, '/');
  throw Error(e == cudaErrorMemoryAllocation ? DCRF_ENOMEM : DCRF_ECUDA,
              std::string("CUDA error: ") + cudaGetErrorString(e) + " [" + expr + "] at " + (base ? base + 1 : file) +
                  ":" + std::to_string(line));
}

// ---- private stream-ordered pools, one per device ------------------------------------------------------
static std::mutex& g_pool_mu = *new std::mutex();
static cudaMemPool_t g_pools[64] = {};
cudaMemPool_t device_pool() {
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (g_pools[dev]) return g_pools[dev];
  std::lock_guard<std::mutex> g(g_pool_mu);
  if (!g_pools[dev]) {
    cudaMemPoolProps props{};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    cudaMemPool_t pool = nullptr;
    DCRF_CUDA(cudaMemPoolCreate(&pool, &props));
    uint64_t keep = UINT64_MAX;  // keep freed blocks cached: no driver calls in steady state (dcrf_empty_cache trims)
    DCRF_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    g_pools[dev] = pool;
  }
  return g_pools[dev];
}

// ---- profiler ----------------------------------------------------------------------------------
cudaEvent_t Profiler::get_event() {
  if (!free_events.empty()) {
    cudaEvent_t e = free_events.back();
    free_events.pop_back();
    return e;
  }
  cudaEvent_t e;
  DCRF_CUDA(cudaEventCreate(&e));
  return e;
}
int Profiler::index_of(const char* name) {
  for (size_t i = 0; i < recs.size(); i++)
    if (recs[i].name == name) return (int)i;
  recs.emplace_back();
  recs.back().name = name;
  return (int)recs.size() - 1;
}
void Profiler::collect() {
  for (auto& r : recs) {
    for (auto& pr : r.pending) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) r.ms += ms;
      free_events.push_back(pr.first);
      free_events.push_back(pr.second);
    }
    r.pending.clear();
  }
}
void Profiler::destroy() {
  collect();
  for (cudaEvent_t e : free_events) cudaEventDestroy(e);
  free_events.clear();
}

// ---- pinned host memory pool: result matrices are re-used across inference calls ----------------
// Pooled blocks come in size classes (eighths of a power of two, at least 4 KB; <= 12.5 % slack), so that a loop over images
// of DIFFERENT sizes -- a dataset -- recycles its result buffers like a loop over one size does (measured on 60 different
// sizes around 500 x 375: 4.8 ms per image with exact-size pools, every image a pool miss, against 2.4 ms for one size).
static size_t size_class(size_t n) {
  if (n <= 4096) return 4096;
  int k = 63 - __builtin_clzll((unsigned long long)n);
  const size_t step = (size_t)1 << (k - 3);
  return (n + step - 1) / step * step;
}

struct PinnedPool {
  std::mutex mu;
  std::multimap<size_t, void*> free_blocks;
  std::map<size_t, int> live;  // blocks of a size that are handed out right now
  size_t cached = 0;
  static constexpr size_t kMaxCached = 8ull << 30;
  // Result buffers (the host mirror of a result matrix, a label map): page-locked while at most `max_live` blocks of this
  // size are out (eight: several models with two alternating results each stay page-locked), else nullptr and the caller takes ordinary memory.  Pinning costs ~0.5 ms per MB (82 ms for a 1080p Q),
  // which a loop that drops or alternates its results earns back through the pool -- a caller who KEEPS every result
  // (a list of label maps) would pay it on every call (measured: np.argmax 8.5 instead of 0.6 ms at 1080p).
  void* get_result(size_t bytes, int max_live = 8) {
    bytes = size_class(bytes);
    {
      std::lock_guard<std::mutex> g(mu);
      auto it = free_blocks.find(bytes);
      if (it != free_blocks.end()) {
        void* p = it->second;
        free_blocks.erase(it);
        cached -= bytes;
        live[bytes]++;
        return p;
      }
      if (live[bytes] >= max_live) return nullptr;
    }
    return get(bytes);
  }
  void* get(size_t bytes) {
    bytes = size_class(bytes);
    {
      std::lock_guard<std::mutex> g(mu);
      live[bytes]++;
      auto it = free_blocks.find(bytes);
      if (it != free_blocks.end()) {
        void* p = it->second;
        free_blocks.erase(it);
        cached -= bytes;
        return p;
      }
    }
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) {
      cudaGetLastError();
      std::lock_guard<std::mutex> g(mu);
      live[bytes]--;
      throw Error(DCRF_ENOMEM, "cudaHostAlloc failed for " + std::to_string(bytes) + " bytes");
    }
    return p;
  }
  void put(void* p, size_t bytes) {
    bytes = size_class(bytes);
    std::lock_guard<std::mutex> g(mu);
    live[bytes]--;
    if (cached + bytes > kMaxCached) {
      cudaFreeHost(p);
      return;
    }
    free_blocks.emplace(bytes, p);
    cached += bytes;
  }
  void trim() {
    std::lock_guard<std::mutex> g(mu);
    for (auto& b : free_blocks) cudaFreeHost(b.second);
    free_blocks.clear();
    cached = 0;
  }
};
static PinnedPool g_pinned;

// device buffers of result matrices, recycled per (device, size): inference() returns a fresh matrix on
// every call, and cudaMalloc/cudaFree of ~200 MB would otherwise sit on the end-to-end path
struct DevicePool {
  std::mutex mu;
  std::multimap<std::pair<int, size_t>, void*> free_blocks;
  size_t cached = 0;
  static constexpr size_t kMaxCached = 48ull << 30;
  void* get(int device, size_t bytes) {
    bytes = size_class(bytes);
    {
      std::lock_guard<std::mutex> g(mu);
      auto it = free_blocks.find({device, bytes});
      if (it != free_blocks.end()) {
        void* p = it->second;
        free_blocks.erase(it);
        cached -= bytes;
        return p;
      }
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
    if (e == cudaErrorMemoryAllocation) {  // give the cached blocks back to the driver and try once more
      cudaGetLastError();
      trim(device);
      e = cudaMalloc(&p, bytes ? bytes : 1);
    }
    if (e == cudaErrorMemoryAllocation) {
      cudaGetLastError();
      throw Error(DCRF_ENOMEM, "out of device memory allocating " + std::to_string(bytes) + " bytes");
    }
    if (e != cudaSuccess) fail_cuda(e, "cudaMalloc", __FILE__, __LINE__);
    return p;
  }
  void put(int device, void* p, size_t bytes) {
    bytes = size_class(bytes);
    std::lock_guard<std::mutex> g(mu);
    if (cached + bytes > kMaxCached) {
      cudaFree(p);
      return;
    }
    free_blocks.emplace(std::make_pair(device, bytes), p);
    cached += bytes;
  }
  // free every cached block of `device` (-1: all devices) and trim that device's stream-ordered pool
  void trim(int device) {
    std::vector<std::pair<int, void*>> drop;
    {
      std::lock_guard<std::mutex> g(mu);
      for (auto it = free_blocks.begin(); it != free_blocks.end();) {
        if (device < 0 || it->first.first == device) {
          drop.emplace_back(it->first.first, it->second);
          cached -= it->first.second;
          it = free_blocks.erase(it);
        } else {
          ++it;
        }
      }
    }
    int prev = -1;
    cudaGetDevice(&prev);
    for (auto& d : drop) {
      cudaSetDevice(d.first);
      cudaFree(d.second);
    }
    for (int dv = 0; dv < 64; dv++)
      if ((device < 0 || dv == device) && g_pools[dv]) cudaMemPoolTrimTo(g_pools[dv], 0);
    cudaGetLastError();
    if (prev >= 0) cudaSetDevice(prev);
  }
};
static DevicePool g_devpool;

// arenas of the batch workers, recycled per (device, size) like the result buffers above
struct ArenaScope {  // installs an arena for the calling thread (nullptr: suspends the current one) for a scope
  Arena* prev;
  explicit ArenaScope(Arena* a) : prev(t_arena) { t_arena = a; }
  ~ArenaScope() { t_arena = prev; }
};

void trim_device_caches() {
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceSynchronize();  // frees queued on any stream have to have happened before the pool can give memory back
  g_devpool.trim(dev);
}

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) DCRF_CUDA(cudaSetDevice(dev));
  }
  ~DeviceGuard() {
    int cur = -1;
    cudaGetDevice(&cur);
    if (prev >= 0 && cur != prev) cudaSetDevice(prev);
  }
};

}  // namespace dcrf

using namespace dcrf;

// ---- matrix object -----------------------------------------------------------------------------
struct dcrf_mat {
  int rows = 0, cols = 0;
  float* host = nullptr;  // rows*cols floats, col-major (pixel-major); malloc'd or pinned
  bool host_pinned = false;
  float* dev = nullptr;   // cols rows of RowLayout(rows).Lp floats
  size_t dev_bytes = 0;
  int device = -1;
  cudaEvent_t ready = nullptr;  // recorded after the last device use (read or write)
  bool host_valid = false, dev_valid = false;
  int exports = 0;              // live buffer views
  size_t bytes() const { return (size_t)rows * cols * sizeof(float); }
};

static void mat_free_host(dcrf_mat* a) {
  if (!a->host) return;
  if (a->host_pinned)
    g_pinned.put(a->host, a->bytes());
  else
    std::free(a->host);
  a->host = nullptr;
}
static void mat_free_dev(dcrf_mat* a) {
  if (a->dev) {
    DeviceGuard g(a->device);
    if (a->ready) cudaEventSynchronize(a->ready);  // nothing queued may still touch the buffer
    g_devpool.put(a->device, a->dev, a->dev_bytes);
    a->dev = nullptr;
  }
  if (a->ready) {
    cudaEventDestroy(a->ready);
    a->ready = nullptr;
  }
}

// ---- model -------------------------------------------------------------------------------------
// A lattice that depends on the geometry only (DenseCRF2D::addPairwiseGaussian: image size and sxy) is the same for
// every image of a video or batch, so it is built once per process and shared, read-only, by all terms that ask for it,
// together with norm_ = compute(1) before the NormalizationType transform.  The last user frees it.
struct SharedLattice {
  int device = 0;
  Lattice lat;
  float* norm_raw = nullptr;  // [N]
  ~SharedLattice() {
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(device);
    lat.release(nullptr);  // every user has synchronised its stream before dropping its reference
    dev_free(norm_raw, nullptr);
    if (prev >= 0) cudaSetDevice(prev);
  }
};

struct Term {
  Lattice lat;                          // shallow copy of shared->lat when shared is set
  std::shared_ptr<SharedLattice> shared;
  float* norm = nullptr;  // [N]
  int ntype = DCRF_NORMALIZE_SYMMETRIC;
  int ckind = DCRF_COMPAT_POTTS;
  float potts_w = 0.f;
  float* diag = nullptr;  // Lp floats
  float* Wm = nullptr;    // L*L floats, symmetrised
  float *Whi = nullptr, *Wlo = nullptr;  // tf32 split of Wm, zero padded (tensor-core path)
  float *valA = nullptr, *valB = nullptr;  // (V+1) x Lp ping-pong value buffers (one allocation)
  float* msg = nullptr;                    // N x Lp message of this term (tensor-core update path only)
  float* qn = nullptr;                     // N x Lp: Q * diag(norm), written by the update kernel for this term's next splat
  CsrEntry* csr_baked = nullptr;           // the lattice's contribution list with w * norm[pix] as weights (VAR_BAKED_NORM)
  size_t val_bytes = 0;
  // learning path (DenseKernel's f_, ktype_, parameters_; pairwise.cpp:35-39): where the features came from, kept for
  // DIAG_KERNEL / FULL_KERNEL terms so that kernelParameters / kernelGradient can form f_ when they are first asked for
  int ktype = DCRF_CONST_KERNEL, d = 0;
  FeatureSource src{};
  void* src_buf = nullptr;       // the device buffer src points into (features / image), owned by the term
  float* feat0 = nullptr;        // f_: d x N row-major, materialised on first use
  std::vector<float> compat_host;  // Potts: [w]; Diagonal: L; Matrix: L x L symmetrised, row-major (parameters())
  std::vector<float> kparams;    // parameters_ (DIAG: d, FULL: d*d column-major); empty = never set (ones / identity)
  bool pre() const { return ntype == DCRF_NORMALIZE_SYMMETRIC || ntype == DCRF_NORMALIZE_BEFORE; }
  bool post() const { return ntype == DCRF_NORMALIZE_SYMMETRIC || ntype == DCRF_NORMALIZE_AFTER; }
};

struct dcrf_model {
  int N = 0, L = 0, W = 0, H = 0, device = 0;
  int npi = 0;  // > 0: the variables are N / npi images of npi pixels stacked along y (internal batch models)
  RowLayout rl;
  cudaStream_t stream = nullptr;
  Profiler prof;
  float* U = nullptr;  // N x Lp, null = no unary set
  // LogisticUnaryEnergy (unary.cpp:47-71) keeps its two factors for parameters() / gradient(): L_ on the host (L x F
  // row-major), f_ on the device (F x N row-major).  A constant unary has neither.
  std::vector<float> ulog_L;
  float* ulog_f = nullptr;
  int ulog_F = 0;
  std::vector<Term*> terms;
  float* acc = nullptr;  // staged path: running tmp1
  float* msg = nullptr;  // staged path / KL: one message
  double* kl_sums = nullptr;   // 3 sums + the per-block partials of the ordered reduction
  dcrf_mat* map_q = nullptr;   // result matrix of dcrf_map (kept: same device buffer -> same CUDA graph on every call)
  int kl_mode = 0;             // 0: ordered double sums; 1: pairwise terms in the reference's float reduction order
  int iter_mode = 0;
  bool fast_softmax = false;
  int variant = 0;      // kernel variants, see default_variant()
  int matrix_mode = 0;  // 0: tensor cores from 64 labels up, 1: always the exact-order kernel, 2: tensor cores whenever possible
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  char* flush_buf = nullptr;
  size_t flush_bytes = 0;
  // side streams: the filters of terms 1.. run beside term 0 (they only meet in the update kernel)
  std::vector<cudaStream_t> aux;
  std::vector<cudaEvent_t> aux_done;
  cudaEvent_t fork = nullptr;
  bool overlap_terms = true;
  // CUDA graphs of whole inference(n) calls, keyed by everything a captured launch sequence depends on
  struct GraphEntry {
    const float* q;
    int n_iter;
    cudaGraphExec_t exec;
    int64_t launches;
    int direct_runs = 1;  // calls with this key that ran without a graph
  };
  std::vector<GraphEntry> graphs;
  bool use_graphs = true;
  void drop_graphs() {
    for (auto& g : graphs)
      if (g.exec) cudaGraphExecDestroy(g.exec);
    graphs.clear();
  }
  Ctx cx() { return Ctx{stream, &prof}; }
};

namespace {

// Kernel variants (bit mask; every combination gives identical bits).  1: wide-lane update kernel (filter_wide.cu);
// 2: the update kernel also writes Q * diag(norm_k) for every pre-normalised term and the next splat reads that copy
// instead of Q and norm.  $DCRF_B200_VARIANT / dcrf_set_kernel_variant choose for A/B timing and tests.
// 4: the pre-normalisation of a term is folded into its contribution weights, (w * norm_i) * Q_i instead of the
// reference's w * (norm_i * Q_i) (pairwise.cpp:65-66 then permutohedral.cpp:553): one extra rounding per contribution
// (~1e-7 on Q), for which the update kernel writes ONE Q that every splat reads instead of a pre-scaled copy per term.
// Overrides 2.  NOT bit-identical to the other variants.
enum { VAR_WIDE_UPDATE = 1, VAR_PRESCALED_Q = 2, VAR_BAKED_NORM = 4, VAR_ALL = 7, VAR_DEFAULT = 3 };
int default_variant() {
  static const int v = [] {
    const char* e = std::getenv("DCRF_B200_VARIANT");
    return e ? (std::atoi(e) & VAR_ALL) : VAR_DEFAULT;
  }();
  return v;
}

template <typename F>
dcrf_status guarded(F&& f) {
  try {
    f();
    return DCRF_OK;
  } catch (const Error& e) {
    g_last_error = e.what();
    return e.code;
  } catch (const std::bad_alloc&) {
    g_last_error = "out of host memory";
    return DCRF_ENOMEM;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return DCRF_ECUDA;
  }
}

int pick_device(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    throw Error(DCRF_ECUDA, "no CUDA device available: libdcrf_b200 has no CPU path");
  }
  if (device < 0) {
    const char* env = std::getenv("DCRF_B200_DEVICE");
    device = env ? std::atoi(env) : 0;
  }
  DCRF_REQUIRE(device >= 0 && device < n, "device index " + std::to_string(device) + " out of range (" + std::to_string(n) + " visible)");
  return device;
}

dcrf_model* model_new(int N, int L, int W, int H, int device) {
  DCRF_REQUIRE(N >= 1, "need at least one variable");
  DCRF_REQUIRE(L >= 1 && L <= 4096, "label count must be in [1, 4096]");
  device = pick_device(device);
  DCRF_CUDA(cudaSetDevice(device));
  dcrf_model* m = new dcrf_model();
  m->N = N, m->L = L, m->W = W, m->H = H, m->device = device;
  m->rl = RowLayout(L);
  DCRF_CUDA(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
  device_pool();  // the library's own stream-ordered pool of this device
  DCRF_CUDA(cudaEventCreate(&m->t0));
  DCRF_CUDA(cudaEventCreate(&m->t1));
  m->kl_sums = dev_alloc<double>(3 + kl_partials(), m->stream);
  m->use_graphs = std::getenv("DCRF_B200_NO_GRAPHS") == nullptr;
  m->variant = default_variant();
  return m;
}

void term_free(dcrf_model* m, Term* t) {
  cudaStream_t s = m->stream;
  if (!t->shared) t->lat.release(s);
  for (void* p : {(void*)t->norm, (void*)t->diag, (void*)t->Wm, (void*)t->Whi, (void*)t->Wlo, (void*)t->valA, (void*)t->msg, (void*)t->qn,
                  t->src_buf, (void*)t->feat0, (void*)t->csr_baked})
    dev_free(p, s);
  delete t;
}

void ensure_val(dcrf_model* m, Term* t) {
  if (t->valA) return;
  const size_t n = ((size_t)t->lat.V + 1) * m->rl.Lp;
  t->valA = dev_alloc<float>(2 * n, m->stream);  // one allocation: A | B
  t->valB = t->valA + n;
  // row 0 (absent neighbour) must read as zero forever; the kernels never write it
  DCRF_CUDA(cudaMemsetAsync(t->valA, 0, 2 * n * sizeof(float), m->stream));
  t->val_bytes = 2 * n * sizeof(float);
}
void ensure_msg(dcrf_model* m, Term* t) {
  if (!t->msg) t->msg = dev_alloc<float>((size_t)m->N * m->rl.Lp, m->stream);
}
void ensure_scratch(dcrf_model* m) {
  const size_t n = (size_t)m->N * m->rl.Lp;
  if (!m->acc) m->acc = dev_alloc<float>(n, m->stream);
  if (!m->msg) m->msg = dev_alloc<float>(n, m->stream);
}

// Permutohedral::compute: splat, d+1 blurs; returns the buffer holding the blurred vertex values.  Inference
// always blurs the axes in forward order; reverse = true is Permutohedral::compute(..., reverse) as used by
// applyTranspose (pairwise.cpp:71-72), reachable through dcrf_lattice_filter for parity checks.
const float* run_filter(Ctx cx, const Lattice& lt, RowLayout rl, const float* in, const float* pre_norm, float* A,
                        float* B, bool reverse = false) {
  // two axes per launch by default (measured 7.89 -> 7.50 ms per 1080p inference); DCRF_B200_BLUR2=0 turns it off
  static const int pair_blur = [] {
    const char* e = std::getenv("DCRF_B200_BLUR2");
    return e ? std::atoi(e) : 1;
  }();
  launch_splat(cx, lt, rl, in, pre_norm, A);
  float *src = A, *dst = B;
  const int n = lt.d + 1;
  for (int i = 0; i < n; i++) {
    const int j = reverse ? n - 1 - i : i;
    const int j2 = reverse ? j - 1 : j + 1;  // the axis after j in blur order
    if (pair_blur && i + 1 < n && lt.V >= pair_blur) {
      launch_blur2(cx, lt, rl, j, j2, src, dst);  // axes j and j2 together
      i++;
    } else {
      launch_blur(cx, lt, rl, j, src, dst);
    }
    std::swap(src, dst);
  }
  return src;
}

// Tensor cores for a MatrixCompatibility term?  Mode 2: always; mode 0: from 64 labels up and only for normalised
// kernels (un-normalised messages are large enough for the 3xTF32 products to cost ~7e-5 on Q at 250 labels).
bool tc_wanted(const dcrf_model* m, const Term* t) {
  if (m->matrix_mode == 2) return true;
  return m->matrix_mode == 0 && m->L >= 64 && t->ntype != DCRF_NO_NORMALIZATION;
}

bool can_fuse(dcrf_model* m) {
  if (!fused_update_supported(m->rl) || (int)m->terms.size() > kMaxFusedTerms) return false;
  if ((uint64_t)m->N * m->rl.Lp >= (1ull << 31)) return false;  // the fused kernel indexes rows with 32 bits
  for (Term* t : m->terms)
    if (t->ckind == DCRF_COMPAT_MATRIX) return false;
  return true;
}

CompatDev compat_of(const Term* t) {
  CompatDev c;
  c.kind = t->ckind;
  c.neg_w = -t->potts_w;
  c.diag = t->diag;
  return c;
}

// message of term k for the current Q into `out` (mode as in launch_slice): PairwisePotential::apply, or with
// transpose = true PairwisePotential::applyTranspose (pairwise.cpp:179-183): the normalisations trade places, the blur
// runs in reverse axis order (DenseKernel::filter, :63-80) and the compatibility is applied transposed -- which for
// Potts, Diagonal and the symmetrised MatrixCompatibility is the same operator
void term_apply(dcrf_model* m, Term* t, const float* Q, int mode, bool neg_unary, float* out, bool transpose = false) {
  Ctx cx = m->cx();
  ensure_val(m, t);
  const bool pre = t->ntype == DCRF_NORMALIZE_SYMMETRIC || t->ntype == (transpose ? DCRF_NORMALIZE_AFTER : DCRF_NORMALIZE_BEFORE);
  const bool post = t->ntype == DCRF_NORMALIZE_SYMMETRIC || t->ntype == (transpose ? DCRF_NORMALIZE_BEFORE : DCRF_NORMALIZE_AFTER);
  const float* fin = run_filter(cx, t->lat, m->rl, Q, pre ? t->norm : nullptr, t->valA, t->valB, transpose);
  if (t->ckind != DCRF_COMPAT_MATRIX) {
    CompatDev c = compat_of(t);
    launch_slice(cx, t->lat, m->rl, fin, post ? t->norm : nullptr, &c, mode, m->U, neg_unary, out);
  } else {
    ensure_scratch(m);
    DCRF_REQUIRE(out != m->msg, "internal: matrix term needs msg as scratch");
    launch_slice(cx, t->lat, m->rl, fin, post ? t->norm : nullptr, nullptr, SLICE_WRITE, nullptr, false, m->msg);
    const bool tc = t->Whi && tc_wanted(m, t);
    if (tc)
      launch_matrix_compat_tc(cx, m->rl, m->N, t->Whi, t->Wlo, m->msg, mode, m->U, neg_unary, out);
    else
      launch_matrix_compat(cx, m->rl, m->N, t->Wm, m->msg, mode, m->U, neg_unary, out);
  }
}

void ensure_scratch_if_staged(dcrf_model* m);

void ensure_aux(dcrf_model* m, int nterms) {
  while ((int)m->aux.size() < nterms - 1) {
    cudaStream_t s2;
    cudaEvent_t e2;
    // the first side stream runs the heaviest term's filter (filter_all_terms) at the highest priority: the shorter
    // filters then only fill the gaps it leaves instead of sharing the SMs with it (1080p: 6.78 -> 6.61 ms per
    // inference(10); DCRF_B200_AUX_PRIO=0 for the A/B).  Graphs keep it through cudaGraphInstantiateFlagUseNodePriority.
    static const bool prio = [] {
      const char* e = std::getenv("DCRF_B200_AUX_PRIO");
      return !e || std::atoi(e) != 0;
    }();
    if (prio && m->aux.empty()) {
      int least = 0, greatest = 0;
      DCRF_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
      DCRF_CUDA(cudaStreamCreateWithPriority(&s2, cudaStreamNonBlocking, greatest));
    } else {
      DCRF_CUDA(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    }
    DCRF_CUDA(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
    m->aux.push_back(s2);
    m->aux_done.push_back(e2);
  }
  if (!m->fork) DCRF_CUDA(cudaEventCreateWithFlags(&m->fork, cudaEventDisableTiming));
}

// every term's filter (splat + blurs) for the current Q, side by side: the heaviest term on the first (high-priority)
// side stream, one term on the model's stream, the others on further side streams; all join the model's stream again
// before returning.  vals[k] = the buffer holding term k's blurred vertex values.
// slice_msgs: also slice + post-normalise each term into its own message array (Term::msg) on the same stream.
void filter_all_terms(dcrf_model* m, const float* Q, const float** vals, bool slice_msgs = false, bool prescaled = false) {
  Ctx cx = m->cx();
  const int nterms = (int)m->terms.size();
  for (int k = 0; k < nterms; k++) ensure_val(m, m->terms[k]);
  // per-launch profiling wants clean, serialised durations: no overlap while it is on
  static const int overlap_env = [] {
    const char* e = std::getenv("DCRF_B200_OVERLAP");
    return e ? std::atoi(e) : -1;
  }();
  const bool want = overlap_env >= 0 ? overlap_env != 0 : m->overlap_terms;
  const bool overlap = want && nterms > 1 && !m->prof.enabled;
  if (overlap) {
    ensure_aux(m, nterms);
    DCRF_CUDA(cudaEventRecord(m->fork, m->stream));
  }
  int heavy = 0;  // by rows moved per launch sequence: (d + 1) x (contributions + blur reads)
  auto cost = [&](const Term* t) { return (double)(t->lat.d + 1) * ((double)t->lat.N + 3.0 * (double)t->lat.V); };
  for (int k = 1; k < nterms; k++)
    if (cost(m->terms[k]) > cost(m->terms[heavy])) heavy = k;
  int next_aux = 1;
  bool main_taken = false;
  for (int k = 0; k < nterms; k++) {
    Term* t = m->terms[k];
    Ctx ck = cx;
    int side = -1;  // index into m->aux, -1 = the model's stream
    if (overlap) {
      if (k == heavy)
        side = 0;
      else if (!main_taken)
        main_taken = true;
      else
        side = next_aux++;
    }
    if (side >= 0) {
      ck.stream = m->aux[side];
      DCRF_CUDA(cudaStreamWaitEvent(ck.stream, m->fork, 0));
    }
    // prescaled: the previous update kernel left Q * diag(norm) in t->qn (bit-identical to scaling in the splat)
    const bool pre_done = prescaled && t->pre() && t->qn;
    if ((m->variant & VAR_BAKED_NORM) && t->pre() && t->csr_baked) {
      Lattice lb = t->lat;  // shallow: the same lattice with the norm-carrying contribution list
      lb.csr = t->csr_baked;
      vals[k] = run_filter(ck, lb, m->rl, Q, nullptr, t->valA, t->valB);
    } else
    vals[k] = run_filter(ck, t->lat, m->rl, pre_done ? t->qn : Q, (t->pre() && !pre_done) ? t->norm : nullptr, t->valA, t->valB);
    if (slice_msgs) {
      ensure_msg(m, t);
      launch_slice(ck, t->lat, m->rl, vals[k], t->post() ? t->norm : nullptr, nullptr, SLICE_WRITE, nullptr, false, t->msg);
    }
    if (side >= 0) DCRF_CUDA(cudaEventRecord(m->aux_done[side], ck.stream));
  }
  // join after everything has been queued: a wait placed earlier would hold back the term on the model's stream
  if (overlap)
    for (int a = 0; a < nterms - 1; a++) DCRF_CUDA(cudaStreamWaitEvent(m->stream, m->aux_done[a], 0));
}

// all terms MatrixCompatibility on the tensor-core path: one kernel does every term's GEMM, the unary and the softmax
// (matrix_tc.cu: k_update_matrix_tc)
bool can_tc_update(dcrf_model* m) {
  if (m->iter_mode == 1 || m->terms.empty()) return false;
  for (Term* t : m->terms)
    if (t->ckind != DCRF_COMPAT_MATRIX || !t->Whi || !tc_wanted(m, t)) return false;
  return update_matrix_tc_supported(m->rl, (int)m->terms.size());
}

bool fuse_wanted(dcrf_model* m) { return m->iter_mode == 2 || (m->iter_mode == 0 && can_fuse(m)); }
// inference(n) with the fused kernel and at least one pre-normalised term: the iterations hand Q * diag(norm) on
bool prescale_wanted(dcrf_model* m) {
  if (!(m->variant & VAR_PRESCALED_Q) || (m->variant & VAR_BAKED_NORM) || !fuse_wanted(m) || !can_fuse(m)) return false;
  for (Term* t : m->terms)
    if (t->pre()) return true;
  return false;
}
void ensure_baked(dcrf_model* m) {
  if (!(m->variant & VAR_BAKED_NORM)) return;
  for (Term* t : m->terms)
    if (t->pre() && !t->csr_baked && t->lat.E > 0) {
      t->csr_baked = dev_alloc<CsrEntry>((size_t)t->lat.E, m->stream);
      launch_bake_csr(m->cx(), (size_t)t->lat.E, t->lat.csr, t->norm, t->csr_baked);
    }
}
void ensure_qn(dcrf_model* m) {
  for (Term* t : m->terms)
    if (t->pre() && !t->qn) t->qn = dev_alloc<float>((size_t)m->N * m->rl.Lp, m->stream);
}
// which copies of the new Q does an update kernel have to store?  qn_out: the next iteration's splats read the pre-scaled
// copies; Q itself is then only needed by terms without a pre-normalisation (or by the caller: keep_q)
void set_outputs(dcrf_model* m, FusedArgs& a, bool qn_out, bool keep_q) {
  a.nqn = 0;
  bool q_needed = keep_q || !qn_out;
  if (qn_out)
    for (Term* t : m->terms) {
      if (t->pre() && t->qn) {
        a.qn[a.nqn] = t->qn, a.qn_norm[a.nqn] = t->norm;
        a.nqn++;
      } else {
        q_needed = true;
      }
    }
  a.write_q = q_needed ? 1 : 0;
}
void run_fused(dcrf_model* m, const FusedArgs& a) {
  if (m->fast_softmax || !(m->variant & VAR_WIDE_UPDATE) || !launch_fused_update8(m->cx(), m->rl, a))
    launch_fused_update(m->cx(), m->rl, a, m->fast_softmax);
}

// one mean-field update of Q in place (the loop body of densecrf.cpp:122-129 / :187-200).  qn_in: the terms' pre-scaled
// copies of Q are current (written by the previous update of the same inference call); qn_out / keep_q: see set_outputs.
void mf_step(dcrf_model* m, float* Q, bool neg_unary, bool qn_in = false, bool qn_out = false, bool keep_q = true) {
  Ctx cx = m->cx();
  const bool fuse = fuse_wanted(m);
  const float* vals[std::max(kMaxFusedTerms, kMaxTCUTerms)];
  if (fuse) {
    DCRF_REQUIRE(can_fuse(m), "fused iteration not available for this model (matrix compatibility, >128 labels or >8 terms)");
    FusedArgs a;
    a.N = m->N;
    a.W = m->W, a.H = m->H;
    a.nterms = (int)m->terms.size();
    a.U = m->U;
    a.neg_unary = neg_unary ? 1 : 0;
    a.Q = Q;
    filter_all_terms(m, Q, vals, false, qn_in);
    for (int k = 0; k < a.nterms; k++) {
      Term* t = m->terms[k];
      FusedTerm& ft = a.t[k];
      ft.val = vals[k];
      ft.pix = t->lat.pix;
      ft.post_norm = t->post() ? t->norm : nullptr;
      ft.diag = t->diag;
      ft.neg_w = -t->potts_w;
      ft.alpha = t->lat.alpha;
      ft.dp1 = t->lat.d + 1;
      ft.ckind = t->ckind;
    }
    set_outputs(m, a, qn_out, keep_q);
    run_fused(m, a);
  } else if (can_tc_update(m)) {
    TCUArgs a;
    a.N = m->N;
    a.nterms = (int)m->terms.size();
    a.U = m->U;
    a.neg_unary = neg_unary ? 1 : 0;
    a.Q = Q;
    filter_all_terms(m, Q, vals, true);
    for (int k = 0; k < a.nterms; k++) {
      Term* t = m->terms[k];
      a.t[k].msg = t->msg;
      a.t[k].Bhi = t->Whi, a.t[k].Blo = t->Wlo;
    }
    launch_update_matrix_tc(cx, m->rl, a);
  } else {
    ensure_scratch(m);
    if (m->terms.empty()) launch_init_acc(cx, m->rl, m->N, m->U, neg_unary, m->acc);
    for (size_t k = 0; k < m->terms.size(); k++)
      term_apply(m, m->terms[k], Q, k == 0 ? SLICE_ACC_FIRST : SLICE_ACC_NEXT, neg_unary, m->acc);
    launch_softmax(cx, m->rl, m->N, m->acc, Q);
  }
}

void ensure_scratch_if_staged(dcrf_model* m) {
  const bool init_fused = m->iter_mode != 1 && fused_update_supported(m->rl) && (uint64_t)m->N * m->rl.Lp < (1ull << 31);
  if (can_tc_update(m)) {  // per-term messages; the accumulator pair only for a staged initial softmax
    for (Term* t : m->terms) ensure_msg(m, t);
    if (!init_fused) ensure_scratch(m);
    return;
  }
  const bool fuse = m->iter_mode == 2 || (m->iter_mode == 0 && can_fuse(m));
  if (!fuse || !(m->iter_mode != 1 && fused_update_supported(m->rl) && (uint64_t)m->N * m->rl.Lp < (1ull << 31))) ensure_scratch(m);
  for (Term* t : m->terms)
    if (t->ckind == DCRF_COMPAT_MATRIX) ensure_scratch(m);
}

// Q0 = expAndNormalize(-unary) (densecrf.cpp:120)
void mf_init(dcrf_model* m, float* Q, bool qn_out = false, bool keep_q = true) {
  Ctx cx = m->cx();
  if (m->iter_mode != 1 && fused_update_supported(m->rl) && (uint64_t)m->N * m->rl.Lp < (1ull << 31)) {
    FusedArgs a;
    a.N = m->N, a.W = m->W, a.H = m->H, a.nterms = 0, a.U = m->U, a.neg_unary = 1, a.Q = Q;
    set_outputs(m, a, qn_out, keep_q);
    run_fused(m, a);
  } else if (m->iter_mode != 1 && m->U && update_matrix_tc_tma_supported(m->rl, 0)) {
    // more than 128 labels: the thread-per-row softmax of the tensor-core update kernel without its main loop (one pass
    // over U and Q instead of the accumulator round trip of the staged form)
    TCUArgs a;
    a.N = m->N, a.nterms = 0, a.U = m->U, a.neg_unary = 1, a.Q = Q;
    launch_update_matrix_tc_tma(cx, m->rl, a);
  } else {
    ensure_scratch(m);
    launch_init_acc(cx, m->rl, m->N, m->U, true, m->acc);
    launch_softmax(cx, m->rl, m->N, m->acc, Q);
  }
}

// inference(n): Q0, then n updates.  With pre-scaled copies the inner iterations never store Q itself.
void mf_run(dcrf_model* m, float* q, int n_iterations) {
  const bool ps = n_iterations > 0 && prescale_wanted(m);
  mf_init(m, q, ps, !ps);
  for (int it = 0; it < n_iterations; it++) {
    const bool last = it + 1 == n_iterations;
    mf_step(m, q, true, ps, ps && !last, last);
  }
}

void mat_reset(dcrf_mat* a, int rows, int cols) {
  if (a->rows == rows && a->cols == cols) return;
  DCRF_REQUIRE(a->exports == 0, "cannot resize a matrix while a buffer view of it is alive");
  mat_free_host(a);
  mat_free_dev(a);
  a->rows = rows, a->cols = cols;
  a->host_valid = a->dev_valid = false;
}

// make sure `a` has a device buffer of this model's layout holding its current contents
float* mat_device(dcrf_model* m, dcrf_mat* a, bool need_contents) {
  DCRF_REQUIRE(a->rows == m->L && a->cols == m->N,
               "matrix has shape (" + std::to_string(a->rows) + ", " + std::to_string(a->cols) + "), model needs (" +
                   std::to_string(m->L) + ", " + std::to_string(m->N) + ")");
  if (a->dev && a->device != m->device) {
    if (!a->host_valid && a->dev_valid) throw Error(DCRF_EINVAL, "matrix lives on another device");
    mat_free_dev(a);
    a->dev_valid = false;
  }
  const size_t n = (size_t)m->N * m->rl.Lp;
  if (!a->dev) {
    a->dev_bytes = n * sizeof(float);
    a->dev = (float*)g_devpool.get(m->device, a->dev_bytes);
    a->device = m->device;
    a->dev_valid = false;
  }
  if (!a->ready)
    DCRF_CUDA(cudaEventCreateWithFlags(&a->ready, cudaEventDisableTiming));
  else if (a->dev)  // another model's stream may have produced or may still be reading the buffer: order after it
    DCRF_CUDA(cudaStreamWaitEvent(m->stream, a->ready, 0));
  if (need_contents && !a->dev_valid) {
    DCRF_REQUIRE(a->host && a->host_valid, "matrix has no contents");
    float* stage = dev_alloc<float>((size_t)m->N * m->L, m->stream);
    DCRF_CUDA(cudaMemcpyAsync(stage, a->host, a->bytes(), cudaMemcpyHostToDevice, m->stream));
    launch_unpack_rows(m->cx(), m->rl, m->N, stage, a->dev);
    dev_free(stage, m->stream);
    a->dev_valid = true;
  }
  return a->dev;
}

// the device copy was just rewritten on the model's stream
void mat_written(dcrf_model* m, dcrf_mat* a);

void download_to(void* host, const char* d, size_t bytes, cudaStream_t s);

void mat_download(dcrf_mat* a, cudaStream_t s, Profiler* prof) {
  // pack rows (cols x Lp) -> col-major rows*cols on the device, then one contiguous D2H copy
  DeviceGuard g(a->device);
  RowLayout rl(a->rows);
  if (!a->host) {
    a->host = (float*)g_pinned.get_result(a->bytes());
    a->host_pinned = a->host != nullptr;
    if (!a->host) {  // many results of this size are being kept alive: ordinary memory, download_to stages the copy
      a->host = (float*)std::malloc(std::max<size_t>(a->bytes(), 1));
      if (!a->host) throw Error(DCRF_ENOMEM, "out of host memory for a result matrix");
    }
  }
  float* stage = dev_alloc<float>((size_t)a->rows * a->cols, s);
  launch_pack_rows(Ctx{s, prof}, rl, a->cols, a->dev, stage);
  if (a->host_pinned) {
    DCRF_CUDA(cudaMemcpyAsync(a->host, stage, a->bytes(), cudaMemcpyDeviceToHost, s));
    DCRF_CUDA(cudaStreamSynchronize(s));
  } else {
    download_to(a->host, (const char*)stage, a->bytes(), s);
  }
  dev_free(stage, s);
  a->host_valid = true;
}

void mat_written(dcrf_model* m, dcrf_mat* a) {
  a->dev_valid = true;
  a->host_valid = false;
  DCRF_CUDA(cudaEventRecord(a->ready, m->stream));
  if (a->exports > 0) {  // a live numpy view must observe the update, like the reference's shared Eigen memory
    mat_download(a, m->stream, &m->prof);
    a->dev_valid = false;  // ...and may be written through before the next call
  }
}

// Host array -> temporary device buffer on the model's stream.  Page-locked memory (utils.pinned_empty,
// dcrf_host_alloc) goes by DMA at PCIe rate (~53 GB/s).  Ordinary pageable memory -- what a numpy array is -- would
// reach the device at ~10 GB/s through the driver's own staging copy, so large pageable uploads are staged here: a few
// threads memcpy the next chunk into one of two page-locked buffers while the previous chunk is on the bus.
// Helper threads for those memcpys: started on first use and kept (a thread start per chunk would cost more than a small
// upload takes).  One upload at a time drives them; a second concurrent upload just copies on its own thread.
class CopyPool {
 public:
  void copy(char* dst, const char* src, size_t n, int nthreads) {
    std::unique_lock<std::mutex> own(owner_, std::try_to_lock);
    if (!own.owns_lock() || nthreads <= 1 || n < (1u << 20)) {
      std::memcpy(dst, src, n);
      return;
    }
    std::unique_lock<std::mutex> lk(mu_);
    while (nhelpers_ < nthreads - 1) {
      std::thread([this] { worker(); }).detach();
      nhelpers_++;
    }
    dst_ = dst, src_ = src, n_ = n;
    parts_ = nthreads;
    per_ = ((n + parts_ - 1) / parts_ + 63) & ~(size_t)63;
    next_ = 0, done_ = 0;
    gen_++;
    cv_.notify_all();
    take_parts(lk);
    cv_done_.wait(lk, [&] { return done_ == parts_; });
  }

 private:
  void take_parts(std::unique_lock<std::mutex>& lk) {  // called with mu_ held
    while (next_ < parts_) {
      const int p = next_++;
      const size_t o = (size_t)p * per_;
      char* d = dst_;
      const char* s = src_;
      const size_t len = o < n_ ? std::min(per_, n_ - o) : 0;
      lk.unlock();
      if (len) std::memcpy(d + o, s + o, len);
      lk.lock();
      if (++done_ == parts_) cv_done_.notify_all();
    }
  }
  void worker() {
    uint64_t seen = 0;
    std::unique_lock<std::mutex> lk(mu_);
    for (;;) {
      cv_.wait(lk, [&] { return gen_ != seen; });
      seen = gen_;
      take_parts(lk);
    }
  }
  std::mutex owner_, mu_;
  std::condition_variable cv_, cv_done_;
  char* dst_ = nullptr;
  const char* src_ = nullptr;
  size_t n_ = 0, per_ = 0;
  int parts_ = 0, next_ = 0, done_ = 0, nhelpers_ = 0;
  uint64_t gen_ = 0;
};
static CopyPool& g_copy_pool = *new CopyPool();  // never destroyed: its detached helpers may outlive static destructors

void upload_to(dcrf_model* m, char* d, const void* host, size_t bytes) {
  static const int nthreads = [] {
    const char* e = std::getenv("DCRF_B200_UPLOAD_THREADS");
    return e ? std::max(0, std::atoi(e)) : 4;
  }();
  cudaPointerAttributes at{};
  const bool pageable = cudaPointerGetAttributes(&at, host) != cudaSuccess || at.type == cudaMemoryTypeUnregistered;
  cudaGetLastError();
  if (!pageable || nthreads == 0 || bytes < (4u << 20)) {
    DCRF_CUDA(cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, m->stream));
    return;
  }
  const size_t mb = 1u << 20;
  const size_t chunk = std::min((size_t)32 * mb, std::max((size_t)2 * mb, ((bytes / 4 + mb - 1) / mb) * mb));
  char* stage[2] = {(char*)g_pinned.get(chunk), (char*)g_pinned.get(chunk)};
  cudaEvent_t done[2] = {nullptr, nullptr};
  try {
    for (int k = 0; k < 2; k++) DCRF_CUDA(cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming));
    int i = 0;
    for (size_t off = 0; off < bytes; off += chunk, i++) {
      const size_t n = std::min(chunk, bytes - off);
      if (i >= 2) DCRF_CUDA(cudaEventSynchronize(done[i & 1]));  // the DMA that last read this staging buffer
      g_copy_pool.copy(stage[i & 1], (const char*)host + off, n, nthreads);
      DCRF_CUDA(cudaMemcpyAsync(d + off, stage[i & 1], n, cudaMemcpyHostToDevice, m->stream));
      DCRF_CUDA(cudaEventRecord(done[i & 1], m->stream));
    }
    for (int k = 0; k < 2; k++) DCRF_CUDA(cudaEventSynchronize(done[k]));
  } catch (...) {
    cudaStreamSynchronize(m->stream);
    for (int k = 0; k < 2; k++) {
      if (done[k]) cudaEventDestroy(done[k]);
      g_pinned.put(stage[k], chunk);
    }
    throw;
  }
  for (int k = 0; k < 2; k++) {
    cudaEventDestroy(done[k]);
    g_pinned.put(stage[k], chunk);
  }
}

// The other direction: device -> host array that may be ordinary (pageable) memory -- a result the caller keeps, see
// PinnedPool::get_result.  The driver's own staging of such a copy runs at ~3.5 GB/s (measured: 5 ms for a 16.6 MB label
// map); here the DMA lands in one of two page-locked staging blocks while the helper threads copy the previous chunk out.
// Returns with the copy COMPLETE.  Page-locked destinations take the plain asynchronous copy + one synchronisation.
void download_to(void* host, const char* d, size_t bytes, cudaStream_t s) {
  static const int nthreads = [] {
    const char* e = std::getenv("DCRF_B200_UPLOAD_THREADS");
    return e ? std::max(0, std::atoi(e)) : 4;
  }();
  cudaPointerAttributes at{};
  const bool pageable = cudaPointerGetAttributes(&at, host) != cudaSuccess || at.type == cudaMemoryTypeUnregistered;
  cudaGetLastError();
  if (!pageable || nthreads == 0 || bytes < (1u << 20)) {
    DCRF_CUDA(cudaMemcpyAsync(host, d, bytes, cudaMemcpyDeviceToHost, s));
    DCRF_CUDA(cudaStreamSynchronize(s));
    return;
  }
  const size_t mb = 1u << 20;
  const size_t chunk = std::min((size_t)32 * mb, std::max((size_t)1 * mb, ((bytes / 4 + mb - 1) / mb) * mb));
  char* stage[2] = {(char*)g_pinned.get(chunk), (char*)g_pinned.get(chunk)};
  cudaEvent_t landed[2] = {nullptr, nullptr};
  try {
    for (int k = 0; k < 2; k++) DCRF_CUDA(cudaEventCreateWithFlags(&landed[k], cudaEventDisableTiming));
    const size_t nchunks = (bytes + chunk - 1) / chunk;
    auto issue = [&](size_t c) {
      const size_t off = c * chunk, n = std::min(chunk, bytes - off);
      DCRF_CUDA(cudaMemcpyAsync(stage[c & 1], d + off, n, cudaMemcpyDeviceToHost, s));
      DCRF_CUDA(cudaEventRecord(landed[c & 1], s));
    };
    issue(0);
    for (size_t c = 0; c < nchunks; c++) {
      DCRF_CUDA(cudaEventSynchronize(landed[c & 1]));
      const size_t off = c * chunk, n = std::min(chunk, bytes - off);
      // the other staging block is free (its chunk was copied out in the previous round): start the next DMA first
      if (c + 1 < nchunks) issue(c + 1);
      g_copy_pool.copy((char*)host + off, stage[c & 1], n, nthreads);
    }
  } catch (...) {
    cudaStreamSynchronize(s);
    for (int k = 0; k < 2; k++) {
      if (landed[k]) cudaEventDestroy(landed[k]);
      g_pinned.put(stage[k], chunk);
    }
    throw;
  }
  for (int k = 0; k < 2; k++) {
    cudaEventDestroy(landed[k]);
    g_pinned.put(stage[k], chunk);
  }
}
float* upload_tmp(dcrf_model* m, const void* host, size_t bytes) {
  char* d = dev_alloc<char>(bytes, m->stream);
  try {
    upload_to(m, d, host, bytes);
  } catch (...) {
    dev_free(d, m->stream);
    throw;
  }
  return (float*)d;
}

// norm_ = lattice_.compute(ones)  (pairwise.cpp:44): a 1-label pass, which takes the seqCompute arithmetic
float* raw_norm(dcrf_model* m, const Lattice& lat) {
  Ctx cx = m->cx();
  cudaStream_t s = m->stream;
  const int N = m->N;
  RowLayout r1(1);
  float* ones = dev_alloc<float>(N, s);
  launch_fill(cx, ones, N, 1.0f);
  float* A = dev_alloc<float>((size_t)lat.V + 1, s);
  float* B = dev_alloc<float>((size_t)lat.V + 1, s);
  DCRF_CUDA(cudaMemsetAsync(A, 0, ((size_t)lat.V + 1) * sizeof(float), s));
  DCRF_CUDA(cudaMemsetAsync(B, 0, ((size_t)lat.V + 1) * sizeof(float), s));
  const float* fin = run_filter(cx, lat, r1, ones, nullptr, A, B);
  float* norm = dev_alloc<float>(N, s);
  launch_slice(cx, lat, r1, fin, nullptr, nullptr, SLICE_WRITE, nullptr, false, norm);
  dev_free(ones, s);
  dev_free(A, s);
  dev_free(B, s);
  return norm;
}

struct GeoKey {
  int device, W, H, npi;
  uint32_t sx, sy;
  bool operator==(const GeoKey& o) const {
    return device == o.device && W == o.W && H == o.H && npi == o.npi && sx == o.sx && sy == o.sy;
  }
};
// most recent last, at most 4 entries; heap objects that are never destroyed, so that no CUDA call runs from a static
// destructor at process exit (entries still referenced by live models stay alive through their shared_ptr anyway)
static std::mutex& g_geo_mu = *new std::mutex();
static std::vector<std::pair<GeoKey, std::shared_ptr<SharedLattice>>>& g_geo_cache =
    *new std::vector<std::pair<GeoKey, std::shared_ptr<SharedLattice>>>();

std::shared_ptr<SharedLattice> shared_gaussian_lattice(dcrf_model* m, const FeatureSource& src) {
  static const bool off = [] {
    const char* e = std::getenv("DCRF_B200_NO_LATTICE_CACHE");
    return e && std::atoi(e) != 0;
  }();
  GeoKey key{m->device, m->W, m->H, m->npi, 0, 0};
  std::memcpy(&key.sx, &src.s[0], 4);
  std::memcpy(&key.sy, &src.s[1], 4);
  if (!off) {
    std::lock_guard<std::mutex> g(g_geo_mu);
    for (auto& e : g_geo_cache)
      if (e.first == key) return e.second;
  }
  ArenaScope outside(nullptr);  // the shared lattice outlives the caller's chunk: never from a worker's arena
  auto sp = std::make_shared<SharedLattice>();
  sp->device = m->device;
  lattice_build(m->cx(), src, 2, m->N, &sp->lat, m->npi);
  sp->norm_raw = raw_norm(m, sp->lat);
  DCRF_CUDA(cudaGetLastError());
  DCRF_CUDA(cudaStreamSynchronize(m->stream));  // complete before any other stream may read it
  if (!off) {
    std::lock_guard<std::mutex> g(g_geo_mu);
    for (auto& e : g_geo_cache)
      if (e.first == key) return e.second;  // another thread was faster: use its copy, ours dies with sp
    if (g_geo_cache.size() >= 4) g_geo_cache.erase(g_geo_cache.begin());
    g_geo_cache.emplace_back(key, sp);
  }
  return sp;
}

// DenseKernel::initLattice (pairwise.cpp:40-62): the lattice of `src` and the normalisation vector of t->ntype
void term_init_lattice(dcrf_model* m, Term* t, const FeatureSource& src, int d) {
  Ctx cx = m->cx();
  cudaStream_t s = m->stream;
  if (src.kind == FEAT_GAUSS2D) {
    t->shared = shared_gaussian_lattice(m, src);
    t->lat = t->shared->lat;  // pointers only; owned by the shared object
  } else {
    lattice_build(cx, src, d, m->N, &t->lat, m->npi);
  }
  const int N = m->N, ntype = t->ntype;
  if (t->shared) {
    t->norm = dev_alloc<float>(N, s);
    DCRF_CUDA(cudaMemcpyAsync(t->norm, t->shared->norm_raw, (size_t)N * sizeof(float), cudaMemcpyDeviceToDevice, s));
  } else {
    t->norm = raw_norm(m, t->lat);
  }
  if (ntype == DCRF_NO_NORMALIZATION) {
    // mean_norm = N / sum(norm): the reference adds N floats one after the other in float
    // (pairwise.cpp:47-51); that rounding sequence is inherently serial, so it runs on the host once
    std::vector<float> h(N);
    DCRF_CUDA(cudaMemcpyAsync(h.data(), t->norm, (size_t)N * sizeof(float), cudaMemcpyDeviceToHost, s));
    DCRF_CUDA(cudaStreamSynchronize(s));
    const int per = m->npi > 0 ? m->npi : N;  // a batch: every image has its own mean, as separate models would
    for (int i0 = 0; i0 < N; i0 += per) {
      float mean = 0.f;
      for (int i = i0; i < i0 + per; i++) mean += h[i];
      mean = (float)per / mean;
      launch_fill(cx, t->norm + i0, per, mean);
    }
  } else {
    launch_norm_transform(cx, N, t->norm, ntype == DCRF_NORMALIZE_SYMMETRIC);
  }
  DCRF_CUDA(cudaGetLastError());
  DCRF_CUDA(cudaStreamSynchronize(s));
}

// LabelCompatibility of a term from its host description (labelcompatibility.cpp:43, :63, :76); replaces what was there
void term_set_compat(dcrf_model* m, Term* t, int compat_kind, const float* compat) {
  Ctx cx = m->cx();
  cudaStream_t s = m->stream;
  const int L = m->L, Lp = m->rl.Lp;
  t->ckind = compat_kind;
  if (compat_kind == DCRF_COMPAT_POTTS) {
    t->potts_w = compat[0];
    t->compat_host.assign(1, compat[0]);
  } else if (compat_kind == DCRF_COMPAT_DIAGONAL) {
    std::vector<float> w(Lp, 0.f);
    std::copy(compat, compat + L, w.begin());
    t->compat_host.assign(compat, compat + L);
    if (!t->diag) t->diag = dev_alloc<float>(Lp, s);
    DCRF_CUDA(cudaMemcpyAsync(t->diag, w.data(), Lp * sizeof(float), cudaMemcpyHostToDevice, s));
    DCRF_CUDA(cudaStreamSynchronize(s));
  } else {
    std::vector<float> w((size_t)L * L);
    for (int a = 0; a < L; a++)
      for (int b = 0; b < L; b++) w[(size_t)a * L + b] = 0.5f * (compat[(size_t)a * L + b] + compat[(size_t)b * L + a]);
    t->compat_host = w;
    if (!t->Wm) t->Wm = dev_alloc<float>((size_t)L * L, s);
    DCRF_CUDA(cudaMemcpyAsync(t->Wm, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice, s));
    if (matrix_tc_supported(m->rl)) {
      int Np, Kp;
      matrix_tc_dims(m->rl, &Np, &Kp);
      if (!t->Whi) t->Whi = dev_alloc<float>((size_t)Np * Kp, s);
      if (!t->Wlo) t->Wlo = dev_alloc<float>((size_t)Np * Kp, s);
      launch_split_w(cx, m->rl, t->Wm, t->Whi, t->Wlo);
    }
    DCRF_CUDA(cudaStreamSynchronize(s));
  }
}

// returns true when the term took ownership of `src_buf` (the device buffer `src` points into): DIAG_KERNEL / FULL_KERNEL
// terms keep their feature source for the learning path, CONST_KERNEL terms have no kernel parameters
bool add_term(dcrf_model* m, const FeatureSource& src, int d, int compat_kind, const float* compat, int ktype,
              int ntype, void* src_buf = nullptr) {
  DCRF_REQUIRE(ntype >= 0 && ntype <= 3, "bad NormalizationType");
  DCRF_REQUIRE(ktype >= 0 && ktype <= 2, "bad KernelType");  // no effect on inference (pairwise.cpp:116-121)
  DCRF_REQUIRE(compat_kind >= 0 && compat_kind <= 2 && compat, "bad label compatibility");
  Term* t = new Term();
  try {
    t->ntype = ntype;
    t->ktype = ktype;
    t->d = d;
    term_init_lattice(m, t, src, d);
    term_set_compat(m, t, compat_kind, compat);
  } catch (...) {
    term_free(m, t);
    throw;
  }
  const bool adopt = ktype != DCRF_CONST_KERNEL;
  if (adopt) {
    t->src = src;
    t->src_buf = src_buf;
  }
  m->terms.push_back(t);
  m->drop_graphs();
  return adopt;
}

dcrf_mat* new_result_mat(dcrf_model* m) {
  dcrf_mat* a = new dcrf_mat();
  a->rows = m->L, a->cols = m->N;
  return a;
}

}  // namespace

// =================================================================================================
// C-ABI
// =================================================================================================
extern "C" {

const char* dcrf_last_error(void) { return g_last_error.c_str(); }
int dcrf_abi_version(void) { return DCRF_ABI_VERSION; }
int dcrf_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

dcrf_status dcrf_create(int nvar, int nlabels, int device, dcrf_t** out) {
  return guarded([&] {
    DCRF_REQUIRE(out, "null output pointer");
    *out = model_new(nvar, nlabels, 0, 0, device);
  });
}
dcrf_status dcrf_create_2d(int width, int height, int nlabels, int device, dcrf_t** out) {
  return guarded([&] {
    DCRF_REQUIRE(out, "null output pointer");
    DCRF_REQUIRE(width >= 1 && height >= 1, "width and height must be positive");
    DCRF_REQUIRE((int64_t)width * height < (1ll << 31), "image too large");
    *out = model_new(width * height, nlabels, width, height, device);
  });
}
void dcrf_destroy(dcrf_t* m) {
  if (!m) return;
  if (m->map_q) dcrf_mat_destroy(m->map_q);
  try {
    DeviceGuard g(m->device);
    cudaStreamSynchronize(m->stream);
    for (Term* t : m->terms) term_free(m, t);
    for (void* p : {(void*)m->U, (void*)m->acc, (void*)m->msg, (void*)m->kl_sums, (void*)m->flush_buf, (void*)m->ulog_f}) dev_free(p, m->stream);
    cudaStreamSynchronize(m->stream);
    m->drop_graphs();
    for (cudaStream_t s2 : m->aux) cudaStreamDestroy(s2);
    for (cudaEvent_t e2 : m->aux_done) cudaEventDestroy(e2);
    if (m->fork) cudaEventDestroy(m->fork);
    m->prof.destroy();
    cudaEventDestroy(m->t0);
    cudaEventDestroy(m->t1);
    cudaStreamDestroy(m->stream);
  } catch (...) {
  }
  delete m;
}

// a new unary replaces the old one (densecrf.cpp:88-96): a logistic unary's factors go with it
static void forget_logistic(dcrf_model* m) {
  if (m->ulog_f) dev_free(m->ulog_f, m->stream);
  m->ulog_f = nullptr;
  m->ulog_F = 0;
  m->ulog_L.clear();
}
dcrf_status dcrf_set_unary_const(dcrf_t* m, const float* u_LxN) {
  return guarded([&] {
    DCRF_REQUIRE(m && u_LxN, "null argument");
    DeviceGuard g(m->device);
    forget_logistic(m);
    const size_t n = (size_t)m->N * m->L;
    float* planes = upload_tmp(m, u_LxN, n * sizeof(float));
    if (!m->U) {
      m->U = dev_alloc<float>((size_t)m->N * m->rl.Lp, m->stream);
      m->drop_graphs();
    }
    launch_planes_to_rows(m->cx(), m->rl, m->N, planes, m->U);
    dev_free(planes, m->stream);
    DCRF_CUDA(cudaGetLastError());
    DCRF_CUDA(cudaStreamSynchronize(m->stream));  // the caller may free/overwrite u now
  });
}
dcrf_status dcrf_set_unary_device(dcrf_t* m, const float* dev_u_LxN) {
  return guarded([&] {
    DCRF_REQUIRE(m && dev_u_LxN, "null argument");
    DeviceGuard g(m->device);
    forget_logistic(m);
    cudaPointerAttributes at{};
    DCRF_CUDA(cudaPointerGetAttributes(&at, dev_u_LxN));
    DCRF_REQUIRE(at.type == cudaMemoryTypeDevice && at.device == m->device,
                 "setUnaryEnergyDevice needs a device pointer on the model's GPU (device " + std::to_string(m->device) + ")");
    // the producer (another library's stream) is unknown to this engine: order after everything on the device
    DCRF_CUDA(cudaDeviceSynchronize());
    if (!m->U) {
      m->U = dev_alloc<float>((size_t)m->N * m->rl.Lp, m->stream);
      m->drop_graphs();
    }
    launch_planes_to_rows(m->cx(), m->rl, m->N, dev_u_LxN, m->U);
    DCRF_CUDA(cudaGetLastError());
    DCRF_CUDA(cudaStreamSynchronize(m->stream));  // the caller may reuse its buffer now
  });
}
dcrf_status dcrf_set_unary_from_softmax(dcrf_t* m, const float* sm_LxN, double scale, double clip) {
  return guarded([&] {
    DCRF_REQUIRE(m && sm_LxN, "null argument");
    DCRF_REQUIRE(!(scale > 1.0), "`scale` needs to be in (0,1]");
    DeviceGuard g(m->device);
    forget_logistic(m);
    const size_t n = (size_t)m->N * m->L;
    float* planes = upload_tmp(m, sm_LxN, n * sizeof(float));
    if (!m->U) {
      m->U = dev_alloc<float>((size_t)m->N * m->rl.Lp, m->stream);
      m->drop_graphs();
    }
    launch_softmax_planes_to_unary(m->cx(), m->rl, m->N, planes, scale, clip, m->U);
    dev_free(planes, m->stream);
    DCRF_CUDA(cudaGetLastError());
    DCRF_CUDA(cudaStreamSynchronize(m->stream));
  });
}
// utils.unary_from_labels (pydensecrf/utils.py:6-46) evaluated on the device: one int32 label per variable crosses the bus
// instead of nlabels floats.  row / unsure follow numpy's indexing of `U[labels - 1 if zero_unsure else labels, arange]`
// (negative indices wrap once) and `labels == 0`.
namespace {
__global__ void k_unary_from_labels(int N, int L, int Lp, const int32_t* __restrict__ labels, int zero_unsure, float e_label,
                                    float e_other, float e_unsure, float* __restrict__ U) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (size_t)N * Lp) return;
  const size_t i = q / Lp;
  const int l = (int)(q - i * Lp);
  const int lab = labels[i];
  int row = zero_unsure ? lab - 1 : lab;
  if (row < 0) row += L;
  float v = 0.f;
  if (l < L) v = (zero_unsure && lab == 0) ? e_unsure : (l == row ? e_label : e_other);
  U[q] = v;
}
}  // namespace
dcrf_status dcrf_set_unary_from_labels(dcrf_t* m, const int32_t* labels_N, double gt_prob, int zero_unsure) {
  return guarded([&] {
    DCRF_REQUIRE(m && labels_N, "null argument");
    DCRF_REQUIRE(gt_prob > 0.0 && gt_prob < 1.0, "`gt_prob must be in (0,1).");
    const int N = m->N, L = m->L;
    for (int i = 0; i < N; i++) {  // numpy raises IndexError for a row outside [-L, L)
      const long long row = (long long)labels_N[i] - (zero_unsure ? 1 : 0);
      DCRF_REQUIRE(row >= -(long long)L && row < (long long)L, "label " + std::to_string(labels_N[i]) + " of variable " +
                                                                   std::to_string(i) + " is out of bounds for " + std::to_string(L) + " labels");
    }
    DeviceGuard g(m->device);
    forget_logistic(m);
    int32_t* dl = (int32_t*)upload_tmp(m, labels_N, (size_t)N * sizeof(int32_t));
    if (!m->U) {
      m->U = dev_alloc<float>((size_t)N * m->rl.Lp, m->stream);
      m->drop_graphs();
    }
    const float e_other = (float)-std::log((1.0 - gt_prob) / (L - 1)), e_label = (float)-std::log(gt_prob),
                e_unsure = (float)-std::log(1.0 / L);
    {
      LaunchScope ls(&m->prof, m->stream, "unary_from_labels");
      const size_t n = (size_t)N * m->rl.Lp;
      k_unary_from_labels<<<(unsigned)((n + 255) / 256), 256, 0, m->stream>>>(N, L, m->rl.Lp, dl, zero_unsure ? 1 : 0, e_label, e_other,
                                                                              e_unsure, m->U);
    }
    dev_free(dl, m->stream);
    DCRF_CUDA(cudaGetLastError());
    DCRF_CUDA(cudaStreamSynchronize(m->stream));
  });
}
dcrf_status dcrf_set_unary_logistic(dcrf_t* m, const float* Lmat_LxF, int F, const float* f_FxN) {
  return guarded([&] {
    DCRF_REQUIRE(m && Lmat_LxF && f_FxN && F >= 1, "bad argument");
    DeviceGuard g(m->device);
    float* dL = upload_tmp(m, Lmat_LxF, (size_t)m->L * F * sizeof(float));
    float* df = upload_tmp(m, f_FxN, (size_t)F * m->N * sizeof(float));
    if (!m->U) {
      m->U = dev_alloc<float>((size_t)m->N * m->rl.Lp, m->stream);
      m->drop_graphs();
    }
    launch_logistic_unary(m->cx(), m->rl, m->N, F, dL, df, m->U);
    dev_free(dL, m->stream);
    forget_logistic(m);
    m->ulog_f = df;  // kept: LogisticUnaryEnergy::gradient multiplies by f_ (unary.cpp:67-71)
    m->ulog_F = F;
    m->ulog_L.assign(Lmat_LxF, Lmat_LxF + (size_t)m->L * F);
    DCRF_CUDA(cudaGetLastError());
    DCRF_CUDA(cudaStreamSynchronize(m->stream));
  });
}

dcrf_status dcrf_add_pairwise_energy(dcrf_t* m, const float* feat_dxN, int d, int compat_kind, const float* compat_data,
                                     int kernel_type, int norm_type) {
  return guarded([&] {
    DCRF_REQUIRE(m && feat_dxN, "null argument");
    DCRF_REQUIRE(d >= 1 && d <= kMaxFeatureDims, "feature dimension must be in [1, 32]");
    DeviceGuard g(m->device);
    float* f = upload_tmp(m, feat_dxN, (size_t)d * m->N * sizeof(float));
    FeatureSource src{};
    src.kind = FEAT_GENERIC;
    src.feat = f;
    bool kept = false;
    try {
      kept = add_term(m, src, d, compat_kind, compat_data, kernel_type, norm_type, f);
    } catch (...) {
      dev_free(f, m->stream);
      throw;
    }
    if (!kept) dev_free(f, m->stream);
  });
}
dcrf_status dcrf_add_pairwise_gaussian(dcrf_t* m, float sx, float sy, int compat_kind, const float* compat_data,
                                       int kernel_type, int norm_type) {
  return guarded([&] {
    DCRF_REQUIRE(m, "null argument");
    DCRF_REQUIRE(m->W > 0, "addPairwiseGaussian needs a 2D model (dcrf_create_2d)");
    DeviceGuard g(m->device);
    FeatureSource src{};
    src.kind = FEAT_GAUSS2D;
    src.W = m->W;
    src.H = m->npi > 0 ? m->npi / m->W : 0;
    src.s[0] = sx, src.s[1] = sy;
    add_term(m, src, 2, compat_kind, compat_data, kernel_type, norm_type);
  });
}
dcrf_status dcrf_add_pairwise_bilateral(dcrf_t* m, float sx, float sy, float sr, float sg, float sb,
                                        const uint8_t* rgb_HxWx3, int compat_kind, const float* compat_data,
                                        int kernel_type, int norm_type) {
  return guarded([&] {
    DCRF_REQUIRE(m && rgb_HxWx3, "null argument");
    DCRF_REQUIRE(m->W > 0, "addPairwiseBilateral needs a 2D model (dcrf_create_2d)");
    DeviceGuard g(m->device);
    uint8_t* im = (uint8_t*)upload_tmp(m, rgb_HxWx3, (size_t)m->N * 3);
    FeatureSource src{};
    src.kind = FEAT_BILATERAL2D;
    src.W = m->W;
    src.H = m->npi > 0 ? m->npi / m->W : 0;
    src.rgb = im;
    src.s[0] = sx, src.s[1] = sy, src.s[2] = sr, src.s[3] = sg, src.s[4] = sb;
    bool kept = false;
    try {
      kept = add_term(m, src, 5, compat_kind, compat_data, kernel_type, norm_type, im);
    } catch (...) {
      dev_free(im, m->stream);
      throw;
    }
    if (!kept) dev_free(im, m->stream);
  });
}

dcrf_status dcrf_add_pairwise_nd(dcrf_t* m, int nsp, const int* dims, const float* sdims, int nch, const float* schan,
                                 const void* img, int img_is_float, int compat_kind, const float* compat_data,
                                 int kernel_type, int norm_type) {
  return guarded([&] {
    DCRF_REQUIRE(m && dims && sdims, "null argument");
    DCRF_REQUIRE(nsp >= 1 && nsp <= kMaxNdAxes && nch >= 0 && nch <= kMaxNdAxes, "at most 6 spatial axes and 6 channels");
    DCRF_REQUIRE(nch == 0 || (schan && img), "channels given without an image");
    int64_t total = 1;
    for (int k = 0; k < nsp; k++) {
      DCRF_REQUIRE(dims[k] >= 1, "bad shape");
      total *= dims[k];
    }
    DCRF_REQUIRE(total == m->N, "shape has " + std::to_string(total) + " elements, model has " + std::to_string(m->N) + " variables");
    DeviceGuard g(m->device);
    FeatureSource src{};
    src.kind = FEAT_ND;
    src.nsp = nsp, src.nch = nch, src.img_is_float = img_is_float;
    int st = 1;
    for (int k = nsp - 1; k >= 0; k--) {
      src.dims[k] = dims[k];
      src.stride[k] = st;
      st *= dims[k];
      src.sdims[k] = sdims[k];
    }
    for (int c = 0; c < nch; c++) src.schan[c] = schan[c];
    void* dimg = nullptr;
    if (nch > 0) {
      dimg = upload_tmp(m, img, (size_t)m->N * nch * (img_is_float ? 4 : 1));
      src.img = dimg;
    }
    bool kept = false;
    try {
      kept = add_term(m, src, nsp + nch, compat_kind, compat_data, kernel_type, norm_type, dimg);
    } catch (...) {
      dev_free(dimg, m->stream);
      throw;
    }
    if (!kept) dev_free(dimg, m->stream);
  });
}

dcrf_status dcrf_inference_into(dcrf_t* m, int n_iterations, dcrf_mat_t* Q) {
  return guarded([&] {
    DCRF_REQUIRE(m && Q, "null argument");
    DCRF_REQUIRE(n_iterations >= 0, "negative iteration count");
    DeviceGuard g(m->device);
    mat_reset(Q, m->L, m->N);
    float* q = mat_device(m, Q, false);
    // everything the launches need is allocated up front, so the sequence below is capturable
    for (Term* t : m->terms) ensure_val(m, t);
    ensure_scratch_if_staged(m);
    if (n_iterations > 0 && prescale_wanted(m)) ensure_qn(m);
    ensure_baked(m);
    ensure_aux(m, (int)m->terms.size());
    if (m->use_graphs && !m->prof.enabled && n_iterations > 0) {
      dcrf_model::GraphEntry* hit = nullptr;
      for (auto& g : m->graphs)
        if (g.q == q && g.n_iter == n_iterations) hit = &g;
      if (!hit) {
        // first call with this (Q, n): run directly and only remember the key -- capturing and instantiating a
        // graph costs more than it saves for one-shot models (a fresh model per image, cfg 4)
        if (m->graphs.size() >= 8) m->drop_graphs();
        m->graphs.push_back({q, n_iterations, nullptr, 0});
        mf_run(m, q, n_iterations);
      } else if (!hit->exec && hit->direct_runs < 2) {
        // second call: still direct -- capturing and instantiating costs several milliseconds, which only a caller
        // who keeps repeating the same inference earns back
        hit->direct_runs++;
        mf_run(m, q, n_iterations);
      } else {
        if (!hit->exec) {  // third call: capture the launch sequence (side streams included) once
          const int64_t l0 = m->prof.launches;
          cudaGraph_t graph = nullptr;
          DCRF_CUDA(cudaStreamBeginCapture(m->stream, cudaStreamCaptureModeThreadLocal));
          try {
            mf_run(m, q, n_iterations);
          } catch (...) {
            cudaStreamEndCapture(m->stream, &graph);
            if (graph) cudaGraphDestroy(graph);
            throw;
          }
          DCRF_CUDA(cudaStreamEndCapture(m->stream, &graph));
          cudaGraphExec_t exec = nullptr;
          cudaError_t ge = cudaGraphInstantiate(&exec, graph, cudaGraphInstantiateFlagUseNodePriority);  // side-stream priorities survive the capture
          cudaGraphDestroy(graph);
          if (ge != cudaSuccess) fail_cuda(ge, "cudaGraphInstantiate", __FILE__, __LINE__);
          hit->exec = exec;
          hit->launches = m->prof.launches - l0;
          m->prof.launches = l0;  // the capture itself launched nothing; replays are counted below
        }
        DCRF_CUDA(cudaGraphLaunch(hit->exec, m->stream));
        m->prof.launches += hit->launches;
      }
    } else {
      mf_run(m, q, n_iterations);
    }
    DCRF_CUDA(cudaGetLastError());
    mat_written(m, Q);
  });
}
dcrf_status dcrf_inference(dcrf_t* m, int n_iterations, dcrf_mat_t** Q) {
  if (!m || !Q) {
    g_last_error = "null argument";
    return DCRF_EINVAL;
  }
  dcrf_mat* a = new dcrf_mat();
  dcrf_status st = dcrf_inference_into(m, n_iterations, a);
  if (st != DCRF_OK) {
    dcrf_mat_destroy(a);
    return st;
  }
  *Q = a;
  return DCRF_OK;
}
dcrf_status dcrf_start_inference(dcrf_t* m, dcrf_mat_t** Q) {
  if (!m || !Q) {
    g_last_error = "null argument";
    return DCRF_EINVAL;
  }
  dcrf_mat* a = nullptr;
  dcrf_status st = guarded([&] {
    DeviceGuard g(m->device);
    a = new_result_mat(m);
    float* q = mat_device(m, a, false);
    if (m->U)
      mf_init(m, q);
    else  // startInference without a unary leaves Q all-zero (densecrf.cpp:178-186)
      DCRF_CUDA(cudaMemsetAsync(q, 0, (size_t)m->N * m->rl.Lp * sizeof(float), m->stream));
    DCRF_CUDA(cudaGetLastError());
    mat_written(m, a);
  });
  if (st != DCRF_OK) {
    dcrf_mat_destroy(a);
    return st;
  }
  *Q = a;
  return DCRF_OK;
}
dcrf_status dcrf_step_inference(dcrf_t* m, dcrf_mat_t* Q, dcrf_mat_t* tmp1, dcrf_mat_t* tmp2) {
  (void)tmp1;
  (void)tmp2;
  return guarded([&] {
    DCRF_REQUIRE(m && Q, "null argument");
    DeviceGuard g(m->device);
    float* q = mat_device(m, Q, true);
    ensure_baked(m);
    mf_step(m, q, false);
    DCRF_CUDA(cudaGetLastError());
    mat_written(m, Q);
  });
}
dcrf_status dcrf_kl_divergence(dcrf_t* m, dcrf_mat_t* Q, double* out) {
  return guarded([&] {
    DCRF_REQUIRE(m && Q && out, "null argument");
    DeviceGuard g(m->device);
    Ctx cx = m->cx();
    float* q = mat_device(m, Q, true);
    ensure_scratch(m);
    DCRF_CUDA(cudaMemsetAsync(m->kl_sums, 0, 3 * sizeof(double), m->stream));
    double* part = m->kl_sums + 3;
    launch_kl_terms(cx, m->rl, m->N, q, m->U, nullptr, m->kl_sums, part);
    for (Term* t : m->terms) {
      term_apply(m, t, q, SLICE_WRITE, false, m->acc);
      if (m->kl_mode == 1)
        launch_kl_pair_exact(cx, m->rl, m->N, q, m->acc, m->kl_sums + 2);
      else
        launch_kl_terms(cx, m->rl, m->N, q, nullptr, m->acc, m->kl_sums, part);
    }
    DCRF_CUDA(cudaEventRecord(Q->ready, m->stream));
    double h[3];
    DCRF_CUDA(cudaMemcpyAsync(h, m->kl_sums, sizeof(h), cudaMemcpyDeviceToHost, m->stream));
    DCRF_CUDA(cudaStreamSynchronize(m->stream));
    *out = h[0] + h[1] + h[2];
  });
}
dcrf_status dcrf_current_map(dcrf_t* m, dcrf_mat_t* Q, int16_t* labels_N) {
  return guarded([&] {
    DCRF_REQUIRE(m && Q && labels_N, "null argument");
    DeviceGuard g(m->device);
    float* q = mat_device(m, Q, true);
    int16_t* d = dev_alloc<int16_t>(m->N, m->stream);
    launch_argmax(m->cx(), m->rl, m->N, q, d);
    DCRF_CUDA(cudaEventRecord(Q->ready, m->stream));  // a later writer on another stream has to wait for this read
    download_to(labels_N, (const char*)d, (size_t)m->N * sizeof(int16_t), m->stream);
    dev_free(d, m->stream);
  });
}

dcrf_status dcrf_map(dcrf_t* m, int n_iterations, int16_t* labels_N) {
  if (!m || !labels_N) {
    g_last_error = "null argument";
    return DCRF_EINVAL;
  }
  if (!m->map_q) m->map_q = new dcrf_mat();
  dcrf_status st = dcrf_inference_into(m, n_iterations, m->map_q);
  if (st == DCRF_OK) st = dcrf_current_map(m, m->map_q, labels_N);
  return st;
}

// ---- matrices ------------------------------------------------------------------------------------
dcrf_status dcrf_mat_create_empty(dcrf_mat_t** out) {
  return guarded([&] {
    DCRF_REQUIRE(out, "null output pointer");
    *out = new dcrf_mat();
  });
}
dcrf_status dcrf_mat_from_host(const float* rowmajor, int rows, int cols, dcrf_mat_t** out) {
  return guarded([&] {
    DCRF_REQUIRE(out && rows >= 0 && cols >= 0 && (rowmajor || rows * cols == 0), "bad argument");
    dcrf_mat* a = new dcrf_mat();
    a->rows = rows, a->cols = cols;
    a->host = (float*)std::malloc(std::max<size_t>(a->bytes(), 4));
    if (!a->host) {
      delete a;
      throw std::bad_alloc();
    }
    for (int r = 0; r < rows; r++)
      for (int c = 0; c < cols; c++) a->host[(size_t)c * rows + r] = rowmajor[(size_t)r * cols + c];
    a->host_valid = true;
    *out = a;
  });
}
dcrf_status dcrf_vec_from_host(const float* data, int n, dcrf_mat_t** out) { return dcrf_mat_from_host(data, n, 1, out); }
void dcrf_mat_destroy(dcrf_mat_t* a) {
  if (!a) return;
  try {
    mat_free_dev(a);
  } catch (...) {
  }
  mat_free_host(a);
  delete a;
}
int dcrf_mat_rows(const dcrf_mat_t* a) { return a ? a->rows : 0; }
int dcrf_mat_cols(const dcrf_mat_t* a) { return a ? a->cols : 0; }
int dcrf_mat_on_device(const dcrf_mat_t* a) { return (a && a->dev && a->dev_valid) ? 1 : 0; }
dcrf_status dcrf_mat_host_ptr(dcrf_mat_t* a, float** ptr) {
  return guarded([&] {
    DCRF_REQUIRE(a && ptr, "null argument");
    if (!a->host_valid && a->dev_valid) {
      DeviceGuard g(a->device);
      // order after the producer without needing its stream: wait on the matrix' own event
      cudaStream_t s;
      DCRF_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
      try {
        DCRF_CUDA(cudaStreamWaitEvent(s, a->ready, 0));
        Profiler none;
        mat_download(a, s, &none);
      } catch (...) {
        cudaStreamDestroy(s);
        throw;
      }
      cudaStreamDestroy(s);
    }
    if (!a->host) {  // empty 0x0 matrix or never written
      a->host = (float*)std::calloc(std::max<size_t>((size_t)a->rows * a->cols, 1), sizeof(float));
      a->host_valid = true;
    }
    a->exports++;
    // the view is writable (readonly=0, eigen.pyx:40,92) and shares memory with the matrix in the reference: while it
    // is alive the host copy is the authoritative one, so the next device consumer uploads it again
    a->dev_valid = false;
    *ptr = a->host;
  });
}
dcrf_status dcrf_mat_argmax(dcrf_mat_t* a, void* labels_cols, int label_bytes) {
  return guarded([&] {
    DCRF_REQUIRE(a && labels_cols, "null argument");
    DCRF_REQUIRE(label_bytes == 2 || label_bytes == 8, "labels are int16 (label_bytes = 2) or int64 (8)");
    DCRF_REQUIRE(a->dev && a->dev_valid, "matrix has no current device copy (it was created or modified on the host)");
    DCRF_REQUIRE(a->rows >= 1 && a->rows <= 32767, "bad label count for a label map");
    DeviceGuard g(a->device);
    cudaStream_t s;
    DCRF_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    char* d = nullptr;
    try {
      DCRF_CUDA(cudaStreamWaitEvent(s, a->ready, 0));  // order after the producer without needing its stream
      Profiler none;
      d = dev_alloc<char>((size_t)a->cols * label_bytes, s);
      if (label_bytes == 2)
        launch_argmax(Ctx{s, &none}, RowLayout(a->rows), a->cols, a->dev, (int16_t*)d);
      else
        launch_argmax64(Ctx{s, &none}, RowLayout(a->rows), a->cols, a->dev, (long long*)d);
      DCRF_CUDA(cudaGetLastError());
      DCRF_CUDA(cudaEventRecord(a->ready, s));  // a later writer / free has to wait for this read
      download_to(labels_cols, d, (size_t)a->cols * label_bytes, s);
      dev_free(d, s);
    } catch (...) {
      cudaStreamSynchronize(s);
      cudaStreamDestroy(s);
      throw;
    }
    cudaStreamDestroy(s);
  });
}
dcrf_status dcrf_mat_device_ptr(dcrf_mat_t* a, float** ptr, int* pitch_floats, int* device) {
  return guarded([&] {
    DCRF_REQUIRE(a && ptr && pitch_floats && device, "null argument");
    DCRF_REQUIRE(a->dev && a->dev_valid, "matrix has no current device copy (it was created or modified on the host)");
    DeviceGuard g(a->device);
    if (a->ready) DCRF_CUDA(cudaEventSynchronize(a->ready));  // producer finished: safe for any stream to read
    *ptr = a->dev;
    *pitch_floats = RowLayout(a->rows).Lp;
    *device = a->device;
  });
}
void dcrf_mat_release_host(dcrf_mat_t* a) {
  if (!a || a->exports <= 0) return;
  a->exports--;
  // the view was writable (readonly=0 in the reference's buffer, eigen.pyx:40,92): assume it was written
  a->dev_valid = false;
  a->host_valid = true;
}

// ---- pinned host buffers for callers ---------------------------------------------------------------
dcrf_status dcrf_host_alloc(size_t bytes, void** out) {
  return guarded([&] {
    DCRF_REQUIRE(out, "null argument");
    pick_device(-1);  // pinned memory needs a CUDA context
    *out = g_pinned.get(bytes);
  });
}
void dcrf_host_free(void* p, size_t bytes) {
  if (p) g_pinned.put(p, bytes);
}
dcrf_status dcrf_host_alloc_result(size_t bytes, void** out) {
  return guarded([&] {
    DCRF_REQUIRE(out, "null argument");
    pick_device(-1);
    *out = g_pinned.get_result(bytes);
  });
}

// ---- measurement -----------------------------------------------------------------------------------
void dcrf_empty_cache(void) {
  try {
    g_devpool.trim(-1);
    g_pinned.trim();
  } catch (...) {
  }
}
dcrf_status dcrf_synchronize(dcrf_t* m) {
  return guarded([&] {
    DCRF_REQUIRE(m, "null argument");
    DeviceGuard g(m->device);
    DCRF_CUDA(cudaStreamSynchronize(m->stream));
  });
}
dcrf_status dcrf_timer_start(dcrf_t* m) {
  return guarded([&] {
    DCRF_REQUIRE(m, "null argument");
    DeviceGuard g(m->device);
    DCRF_CUDA(cudaEventRecord(m->t0, m->stream));
  });
}
dcrf_status dcrf_timer_stop(dcrf_t* m, float* ms) {
  return guarded([&] {
    DCRF_REQUIRE(m && ms, "null argument");
    DeviceGuard g(m->device);
    DCRF_CUDA(cudaEventRecord(m->t1, m->stream));
    DCRF_CUDA(cudaEventSynchronize(m->t1));
    DCRF_CUDA(cudaEventElapsedTime(ms, m->t0, m->t1));
  });
}
dcrf_status dcrf_profile_enable(dcrf_t* m, int on) {
  return guarded([&] {
    DCRF_REQUIRE(m, "null argument");
    DeviceGuard g(m->device);
    DCRF_CUDA(cudaStreamSynchronize(m->stream));
    m->prof.collect();
    m->prof.enabled = on != 0;
  });
}
dcrf_status dcrf_profile_read(dcrf_t* m, dcrf_prof_entry* out, int cap, int* n, int reset) {
  return guarded([&] {
    DCRF_REQUIRE(m && n, "null argument");
    DeviceGuard g(m->device);
    DCRF_CUDA(cudaStreamSynchronize(m->stream));
    m->prof.collect();
    int k = 0;
    for (auto& r : m->prof.recs) {
      if (out && k < cap) {
        std::memset(&out[k], 0, sizeof(out[k]));
        std::strncpy(out[k].name, r.name.c_str(), sizeof(out[k].name) - 1);
        out[k].launches = r.launches;
        out[k].ms_total = r.ms;
      }
      k++;
      if (reset) r.launches = 0, r.ms = 0.f;
    }
    *n = k;
  });
}
int64_t dcrf_launch_count(const dcrf_t* m) { return m ? m->prof.launches : 0; }
dcrf_status dcrf_flush_l2(dcrf_t* m) {
  return guarded([&] {
    DCRF_REQUIRE(m, "null argument");
    DeviceGuard g(m->device);
    if (!m->flush_buf) {
      m->flush_bytes = 256ull << 20;  // 2x the 126 MB L2
      m->flush_buf = dev_alloc<char>(m->flush_bytes, m->stream);
    }
    m->prof.launches++;
    DCRF_CUDA(cudaMemsetAsync(m->flush_buf, 0x5a, m->flush_bytes, m->stream));
  });
}

}  // extern "C"

namespace {

// end of the iterations of the chunk that last held a device's lock (guarded by that lock)
cudaEvent_t& last_compute_done(int dev) {
  static cudaEvent_t* ev = new cudaEvent_t[64]();
  return ev[((dev < 0 ? 0 : dev)) & 63];
}

struct BatchParams {
  int W, H, L, n_iterations;
  float g_sx, g_sy, g_w, b_sx, b_sy, b_sr, b_sg, b_sb, b_w;
};

// Images [b0, b0 + Bc) as ONE model of Bc * W * H variables ("super-lattice"): the images are stacked along y,
// both lattices are built once for the whole chunk with the image index as an extra key coordinate (lattice.cu), and
// every kernel of the iteration runs once over the chunk -- at the efficiency of a large image instead of Bc small
// ones.  Tables, ids and arithmetic per image are exactly those of a separate model, so the results are bit-identical.
// Two phases so that a worker can upload its next chunk while another worker's chunk owns the SMs: upload() moves the
// unaries and images (copy engine + one transposition launch), compute() does everything else under the device's lock.
struct BatchChunk {
  const BatchParams& p;
  int b0, Bc;
  size_t Ni, N;
  dcrf_model* m = nullptr;
  uint8_t* im = nullptr;
  BatchChunk(const BatchParams& p_, int dev, int b0_, int Bc_) : p(p_), b0(b0_), Bc(Bc_) {
    Ni = (size_t)p.W * p.H, N = Ni * Bc;
    m = model_new((int)N, p.L, p.W, p.H * Bc, dev);
    m->npi = (int)Ni;
    m->use_graphs = false;  // one inference per model
  }
  ~BatchChunk() {
    if (m) {
      dev_free(im, m->stream);
      dcrf_destroy(m);
    }
  }
  void upload(const float* const* unary, const uint8_t* const* rgb) {
    DeviceGuard g(m->device);
    cudaStream_t s = m->stream;
    const int L = p.L;
    // unaries: every image's (L, Ni) planes into one staging buffer, one transposition launch for the chunk
    float* planes = dev_alloc<float>(N * L, s);
    for (int b = 0; b < Bc; b++) upload_to(m, (char*)(planes + (size_t)b * L * Ni), unary[b0 + b], Ni * L * sizeof(float));
    m->U = dev_alloc<float>(N * m->rl.Lp, s);
    launch_planes_to_rows(m->cx(), m->rl, (int)Ni, planes, m->U, Bc);
    dev_free(planes, s);
    im = dev_alloc<uint8_t>(N * 3, s);
    for (int b = 0; b < Bc; b++) upload_to(m, (char*)(im + (size_t)b * Ni * 3), rgb[b0 + b], Ni * 3);
    DCRF_CUDA(cudaGetLastError());
    DCRF_CUDA(cudaStreamSynchronize(s));  // the caller's buffers are free again; compute() starts from resident data
  }
  // both lattices and their norm vectors (hash insert, sort, neighbour tables; one host round trip for the vertex count)
  void build() {
    DeviceGuard g(m->device);
    cudaStream_t s = m->stream;
    FeatureSource src{};
    src.kind = FEAT_GAUSS2D;
    src.W = p.W, src.H = p.H;
    src.s[0] = p.g_sx, src.s[1] = p.g_sy;
    add_term(m, src, 2, DCRF_COMPAT_POTTS, &p.g_w, DCRF_CONST_KERNEL, DCRF_NORMALIZE_SYMMETRIC);  // no learning path on batch chunks
    src.kind = FEAT_BILATERAL2D;
    src.rgb = im;
    src.s[0] = p.b_sx, src.s[1] = p.b_sy, src.s[2] = p.b_sr, src.s[3] = p.b_sg, src.s[4] = p.b_sb;
    add_term(m, src, 5, DCRF_COMPAT_POTTS, &p.b_w, DCRF_CONST_KERNEL, DCRF_NORMALIZE_SYMMETRIC);
    dev_free(im, s);
    im = nullptr;
  }
  // the mean-field iterations and the label map (+ Q) back into page-locked memory
  // Enqueue only (call under the device lock): the stream first waits for the event that marks the end of the previous
  // chunk's iterations on this device, so chunks own the SMs one after the other WITHOUT a host round trip between
  // them -- the next worker queues its launches while this chunk is still running.  finish() synchronises, outside the lock.
  void compute(int16_t* const* labels, float* const* q_out) {
    DeviceGuard g(m->device);
    cudaStream_t s = m->stream;
    Ctx cx = m->cx();
    const int L = p.L, Lp = m->rl.Lp;
    cudaEvent_t& prev = last_compute_done(m->device);
    if (prev) DCRF_CUDA(cudaStreamWaitEvent(s, prev, 0));
    float* q = dev_alloc<float>(N * Lp, s);
    for (Term* t : m->terms) ensure_val(m, t);
    ensure_scratch_if_staged(m);
    if (p.n_iterations > 0 && prescale_wanted(m)) ensure_qn(m);
    ensure_baked(m);
    ensure_aux(m, (int)m->terms.size());
    mf_run(m, q, p.n_iterations);
    int16_t* dl = dev_alloc<int16_t>(N, s);
    launch_argmax(cx, m->rl, (int)N, q, dl);
    {  // the next chunk may start its kernels here (its copies back to the host overlap)
      cudaEvent_t done;
      DCRF_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
      DCRF_CUDA(cudaEventRecord(done, s));
      if (prev) cudaEventDestroy(prev);  // legal while a stream still waits on it: released when it has completed
      prev = done;
    }
    int16_t* hl = (int16_t*)g_pinned.get(N * sizeof(int16_t));
    float* hq = nullptr;
    try {
      DCRF_CUDA(cudaMemcpyAsync(hl, dl, N * sizeof(int16_t), cudaMemcpyDeviceToHost, s));
      if (q_out) {
        float* packed = dev_alloc<float>(N * L, s);
        launch_pack_rows(cx, m->rl, (int)N, q, packed);
        hq = (float*)g_pinned.get(N * L * sizeof(float));
        DCRF_CUDA(cudaMemcpyAsync(hq, packed, N * L * sizeof(float), cudaMemcpyDeviceToHost, s));
        dev_free(packed, s);
      }
      dev_free(dl, s);
      dev_free(q, s);
      DCRF_CUDA(cudaGetLastError());
    } catch (...) {
      cudaStreamSynchronize(s);
      g_pinned.put(hl, N * sizeof(int16_t));
      if (hq) g_pinned.put(hq, N * L * sizeof(float));
      throw;
    }
    hl_ = hl, hq_ = hq;
    (void)labels;
    (void)q_out;
  }
  // wait for the chunk's kernels and copies (outside the device lock)
  void finish() {
    DeviceGuard g(m->device);
    const cudaError_t e = cudaStreamSynchronize(m->stream);
    if (e != cudaSuccess) {
      const int L = p.L;
      if (hl_) g_pinned.put(hl_, N * sizeof(int16_t));
      if (hq_) g_pinned.put(hq_, N * L * sizeof(float));
      hl_ = nullptr, hq_ = nullptr;
      DCRF_CUDA(e);
    }
  }
  // host-side scatter of the results into the caller's per-image arrays (outside the device lock)
  void deliver(int16_t* const* labels, float* const* q_out) {
    const int L = p.L;
    for (int b = 0; b < Bc; b++) {
      std::memcpy(labels[b0 + b], hl_ + (size_t)b * Ni, Ni * sizeof(int16_t));
      if (hq_ && q_out && q_out[b0 + b]) {  // Eigen layout (Ni columns of L floats) -> (L, Ni) row-major for the caller
        const float* src_q = hq_ + (size_t)b * Ni * L;
        float* dst = q_out[b0 + b];
        for (size_t i = 0; i < Ni; i++)
          for (int l = 0; l < L; l++) dst[(size_t)l * Ni + i] = src_q[i * L + l];
      }
    }
    g_pinned.put(hl_, N * sizeof(int16_t));
    if (hq_) g_pinned.put(hq_, N * L * sizeof(float));
    hl_ = nullptr, hq_ = nullptr;
  }
  int16_t* hl_ = nullptr;
  float* hq_ = nullptr;
};

// bytes one chunk may allocate: the N x Lp arrays (U, Q, two pre-scaled copies, value buffers of a lattice with up to
// ~1.5 vertices per pixel), the lattice tables and the build scratch (hash table, sort buffers) -- a bound, not a need:
// an allocation that does not fit falls through to the stream-ordered pool
size_t chunk_arena_bytes(size_t N, int Lp) { return N * (600 + (size_t)Lp * 4 * 9) + (64u << 20); }

// one chunk at a time owns a device's SMs (the kernels of two chunks would only evict each other's L2 lines); uploads
// and the host-side delivery of other chunks run beside it
std::mutex& device_lock(int dev) {
  static std::mutex* locks = new std::mutex[64];
  return locks[((dev < 0 ? 0 : dev)) & 63];
}

std::mutex& build_lock(int dev) {
  static std::mutex* locks = new std::mutex[64];
  return locks[((dev < 0 ? 0 : dev)) & 63];
}

// images per chunk of the super-lattice path: large enough for full-size kernels, small enough that the upload of one
// chunk hides under the kernels of another (about 0.25 GB of device memory per 512x512 image and chunk in flight)
int batch_chunk_images(int W, int H, int L) {
  const char* e = std::getenv("DCRF_B200_BATCH_CHUNK");  // read per call: tests vary it
  const int env = e ? std::atoi(e) : 0;
  const int64_t Ni = (int64_t)W * H, Lp = RowLayout(L).Lp;
  int64_t c = env > 0 ? env : std::max<int64_t>(1, (4ll << 20) / Ni);  // ~4 MPix per chunk
  c = std::min<int64_t>(c, ((1ll << 31) - 64) / (Ni * 6));              // contribution indices stay below 2^31 (d = 5)
  c = std::min<int64_t>(c, ((1ll << 31) - 64) / (Ni * Lp));             // 32-bit row offsets of the fused kernel
  return (int)std::max<int64_t>(c, 1);
}

}  // namespace

extern "C" {

dcrf_status dcrf_batch_map_2d(int W, int H, int L, int device, int B, const float* const* unary_LxN,
                              const uint8_t* const* rgb_HxWx3, float g_sx, float g_sy, float g_w, float b_sx, float b_sy,
                              float b_sr, float b_sg, float b_sb, float b_w, int n_iterations, int n_workers,
                              int16_t* const* labels, float* const* q_out, const int* devices, int n_devices) {
  return guarded([&] {
    DCRF_REQUIRE(W > 0 && H > 0 && L > 0 && B >= 0 && n_iterations >= 0, "bad batch shape");
    DCRF_REQUIRE(B == 0 || (unary_LxN && rgb_HxWx3 && labels), "null argument");
    DCRF_REQUIRE(n_devices >= 0 && (n_devices == 0 || devices), "bad device list");
    const int ndev = devices ? std::max(n_devices, 1) : 1;
    // $DCRF_B200_BATCH_MODE: 1 (default) = chunks of images as one super-lattice model; 0 = one model per image
    const char* mode_env = std::getenv("DCRF_B200_BATCH_MODE");
    const int mode = mode_env ? std::atoi(mode_env) : 1;
    const int chunk = batch_chunk_images(W, H, L);
    const bool super = mode != 0 && ((int64_t)W * H) % 4 == 0 && chunk >= 2 && B >= 2;
    const int nchunks = super ? (B + chunk - 1) / chunk : B;
    const int nw = std::max(1, std::min(n_workers > 0 ? n_workers : (super ? 3 : 4) * ndev, nchunks));
    std::atomic<int> next_worker{0};
    std::atomic<int> next{0};
    std::atomic<bool> failed{false};
    std::mutex mu;
    std::string first_error;
    dcrf_status first_status = DCRF_OK;
    const BatchParams bp{W, H, L, n_iterations, g_sx, g_sy, g_w, b_sx, b_sy, b_sr, b_sg, b_sb, b_w};
    auto fail = [&](dcrf_status st, const std::string& what) {
      std::lock_guard<std::mutex> g(mu);
      if (!failed.exchange(true)) {
        first_status = st;
        first_error = what;
      }
    };
    auto worker = [&] {
      const size_t N = (size_t)W * H;
      const int wk = next_worker.fetch_add(1);
      const int dev = (devices && n_devices > 0) ? devices[wk % n_devices] : device;
      // super-lattice chunks allocate from a slab owned by this worker for the duration of the call (see Arena)
      Arena arena;
      int arena_dev = -1;
      if (super) {
        try {
          arena_dev = pick_device(dev);
          DeviceGuard g(arena_dev);
          arena.cap = chunk_arena_bytes(N * (size_t)std::min(chunk, B), RowLayout(L).Lp);
          arena.base = (char*)g_devpool.get(arena_dev, arena.cap);
        } catch (...) {
          arena = Arena();  // no slab: the chunks take the stream-ordered pool
        }
      }
      struct ReturnSlab {
        Arena& a;
        int dev;
        ~ReturnSlab() {
          if (a.base) g_devpool.put(dev, a.base, a.cap);
        }
      } return_slab{arena, arena_dev};
      for (;;) {
        const int b = next.fetch_add(1);
        if (b >= nchunks || failed.load()) return;
        if (super) {
          const int b0 = b * chunk, Bc = std::min(chunk, B - b0);
          try {
            arena.off = 0;  // the previous chunk of this worker is complete (its model synchronised and destroyed)
            ArenaScope scope(arena.base ? &arena : nullptr);
            BatchChunk ch(bp, dev, b0, Bc);
            ch.upload(unary_LxN, rgb_HxWx3);
            // two locks per device: the lattice build of one chunk (latency- and atomics-bound kernels with a host round
            // trip in the middle) runs beside the iterations of another and fills the gaps its host-side work leaves
            static const bool split = [] {
              const char* e = std::getenv("DCRF_B200_BATCH_SPLIT_LOCK");
              return !e || std::atoi(e) != 0;
            }();
            if (split) {
              {
                std::lock_guard<std::mutex> own(build_lock(ch.m->device));
                ch.build();
              }
              {
                std::lock_guard<std::mutex> own(device_lock(ch.m->device));
                ch.compute(labels, q_out);
              }
              ch.finish();
            } else {
              std::lock_guard<std::mutex> own(device_lock(ch.m->device));
              ch.build();
              ch.compute(labels, q_out);
              ch.finish();
            }
            ch.deliver(labels, q_out);
          } catch (const Error& e) {
            fail(e.code, "images " + std::to_string(b0) + ".." + std::to_string(b0 + Bc - 1) + ": " + e.what());
          } catch (const std::exception& e) {
            fail(DCRF_ECUDA, "images " + std::to_string(b0) + ".." + std::to_string(b0 + Bc - 1) + ": " + e.what());
          }
          continue;
        }
        dcrf_t* m = nullptr;
        dcrf_mat_t* Q = nullptr;
        dcrf_status st = dcrf_create_2d(W, H, L, dev, &m);
        if (st == DCRF_OK) st = dcrf_set_unary_const(m, unary_LxN[b]);
        if (st == DCRF_OK) st = dcrf_add_pairwise_gaussian(m, g_sx, g_sy, DCRF_COMPAT_POTTS, &g_w, 1, DCRF_NORMALIZE_SYMMETRIC);
        if (st == DCRF_OK)
          st = dcrf_add_pairwise_bilateral(m, b_sx, b_sy, b_sr, b_sg, b_sb, rgb_HxWx3[b], DCRF_COMPAT_POTTS, &b_w, 1,
                                           DCRF_NORMALIZE_SYMMETRIC);
        if (st == DCRF_OK) st = dcrf_inference(m, n_iterations, &Q);
        if (st == DCRF_OK) st = dcrf_current_map(m, Q, labels[b]);
        if (st == DCRF_OK && q_out && q_out[b]) {
          float* host = nullptr;  // Eigen layout: N columns of L floats -> (L, N) row-major for the caller
          st = dcrf_mat_host_ptr(Q, &host);
          if (st == DCRF_OK) {
            float* dst = q_out[b];
            for (size_t i = 0; i < N; i++)
              for (int l = 0; l < L; l++) dst[(size_t)l * N + i] = host[i * L + l];
            dcrf_mat_release_host(Q);
          }
        }
        if (st != DCRF_OK) fail(st, "image " + std::to_string(b) + ": " + dcrf_last_error());
        if (Q) dcrf_mat_destroy(Q);
        if (m) dcrf_destroy(m);
      }
    };
    std::vector<std::thread> pool;
    for (int k = 1; k < nw; k++) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
    if (failed.load()) throw Error(first_status, first_error);
  });
}

dcrf_status dcrf_debug_selftest(dcrf_t* m, int which, uint64_t n, uint64_t seed, uint64_t* mismatches) {
  return guarded([&] {
    DCRF_REQUIRE(m && mismatches && (which == 0 || which == 1), "bad argument");
    DeviceGuard g(m->device);
    Ctx cx = m->cx();
    unsigned long long* bad = dev_alloc<unsigned long long>(1, m->stream);
    DCRF_CUDA(cudaMemsetAsync(bad, 0, sizeof(unsigned long long), m->stream));
    launch_selftest(cx, which, n, seed, bad);
    DCRF_CUDA(cudaGetLastError());
    unsigned long long h = 0;
    DCRF_CUDA(cudaMemcpyAsync(&h, bad, sizeof(h), cudaMemcpyDeviceToHost, m->stream));
    DCRF_CUDA(cudaStreamSynchronize(m->stream));
    dev_free(bad, m->stream);
    *mismatches = h;
  });
}

// ---- introspection -----------------------------------------------------------------------------------
int dcrf_num_pairwise(const dcrf_t* m) { return m ? (int)m->terms.size() : 0; }
dcrf_status dcrf_lattice_info(const dcrf_t* m, int k, int* d, int* N, int* V) {
  return guarded([&] {
    DCRF_REQUIRE(m && k >= 0 && k < (int)m->terms.size(), "no such pairwise term");
    const Lattice& lt = m->terms[k]->lat;
    if (d) *d = lt.d;
    if (N) *N = lt.N;
    if (V) *V = lt.V;
  });
}
dcrf_status dcrf_lattice_dump(dcrf_t* m, int k, int32_t* offset, float* bary, int16_t* keys, int32_t* nbr, float* norm) {
  return guarded([&] {
    DCRF_REQUIRE(m && k >= 0 && k < (int)m->terms.size(), "no such pairwise term");
    DeviceGuard g(m->device);
    const Term* t = m->terms[k];
    const Lattice& lt = t->lat;
    cudaStream_t s = m->stream;
    const size_t E = (size_t)lt.E;
    if (offset || bary) {
      std::vector<PixEntry> h(E);
      DCRF_CUDA(cudaMemcpyAsync(h.data(), lt.pix, E * sizeof(PixEntry), cudaMemcpyDeviceToHost, s));
      DCRF_CUDA(cudaStreamSynchronize(s));
      for (size_t e = 0; e < E; e++) {
        if (offset) offset[e] = h[e].row - 1;
        if (bary) bary[e] = h[e].w;
      }
    }
    if (keys) DCRF_CUDA(cudaMemcpyAsync(keys, lt.keys, (size_t)lt.V * lt.d * sizeof(int16_t), cudaMemcpyDeviceToHost, s));
    if (nbr) {
      const size_t n = (size_t)lt.V * (lt.d + 1);
      std::vector<int2> h(n);
      DCRF_CUDA(cudaMemcpyAsync(h.data(), lt.nbr, n * sizeof(int2), cudaMemcpyDeviceToHost, s));
      DCRF_CUDA(cudaStreamSynchronize(s));
      for (size_t e = 0; e < n; e++) nbr[2 * e] = h[e].x - 1, nbr[2 * e + 1] = h[e].y - 1;
    }
    if (norm) DCRF_CUDA(cudaMemcpyAsync(norm, t->norm, (size_t)lt.N * sizeof(float), cudaMemcpyDeviceToHost, s));
    DCRF_CUDA(cudaStreamSynchronize(s));
  });
}
dcrf_status dcrf_lattice_filter(dcrf_t* m, int k, const float* in_NxL, int L, int reverse, float* out_NxL) {
  return guarded([&] {
    DCRF_REQUIRE(m && k >= 0 && k < (int)m->terms.size(), "no such pairwise term");
    DCRF_REQUIRE(in_NxL && out_NxL && L >= 1 && L <= 4096, "bad argument");
    DeviceGuard g(m->device);
    Ctx cx = m->cx();
    cudaStream_t s = m->stream;
    const Lattice& lt = m->terms[k]->lat;
    RowLayout rl(L);
    const int N = lt.N;
    float* packed = upload_tmp(m, in_NxL, (size_t)N * L * sizeof(float));
    float* rows = dev_alloc<float>((size_t)N * rl.Lp, s);
    launch_unpack_rows(cx, rl, N, packed, rows);
    const size_t vn = ((size_t)lt.V + 1) * rl.Lp;
    float* A = dev_alloc<float>(vn, s);
    float* B = dev_alloc<float>(vn, s);
    DCRF_CUDA(cudaMemsetAsync(A, 0, vn * sizeof(float), s));
    DCRF_CUDA(cudaMemsetAsync(B, 0, vn * sizeof(float), s));
    const float* fin = run_filter(cx, lt, rl, rows, nullptr, A, B, reverse != 0);
    launch_slice(cx, lt, rl, fin, nullptr, nullptr, SLICE_WRITE, nullptr, false, rows);
    launch_pack_rows(cx, rl, N, rows, packed);
    DCRF_CUDA(cudaMemcpyAsync(out_NxL, packed, (size_t)N * L * sizeof(float), cudaMemcpyDeviceToHost, s));
    for (void* p : {(void*)packed, (void*)rows, (void*)A, (void*)B}) dev_free(p, s);
    DCRF_CUDA(cudaGetLastError());
    DCRF_CUDA(cudaStreamSynchronize(s));
  });
}
static dcrf_status pairwise_apply_impl(dcrf_t* m, int k, dcrf_mat_t* Q, dcrf_mat_t** out, bool transpose) {
  if (!out) {
    g_last_error = "null argument";
    return DCRF_EINVAL;
  }
  dcrf_mat* a = nullptr;
  dcrf_status st = guarded([&] {
    DCRF_REQUIRE(m && Q && k >= 0 && k < (int)m->terms.size(), "no such pairwise term");
    DeviceGuard g(m->device);
    float* q = mat_device(m, Q, true);
    a = new_result_mat(m);
    float* o = mat_device(m, a, false);
    term_apply(m, m->terms[k], q, SLICE_WRITE, false, o, transpose);
    DCRF_CUDA(cudaGetLastError());
    DCRF_CUDA(cudaEventRecord(Q->ready, m->stream));
    mat_written(m, a);
  });
  if (st != DCRF_OK) {
    dcrf_mat_destroy(a);
    return st;
  }
  *out = a;
  return DCRF_OK;
}
dcrf_status dcrf_pairwise_apply(dcrf_t* m, int k, dcrf_mat_t* Q, dcrf_mat_t** out) { return pairwise_apply_impl(m, k, Q, out, false); }
dcrf_status dcrf_pairwise_apply_transpose(dcrf_t* m, int k, dcrf_mat_t* Q, dcrf_mat_t** out) {
  return pairwise_apply_impl(m, k, Q, out, true);
}
dcrf_status dcrf_lattice_gradient(dcrf_t* m, int k, const float* a_NxL, const float* b_NxL, int L, float* df_Nxd) {
  return guarded([&] {
    DCRF_REQUIRE(m && k >= 0 && k < (int)m->terms.size(), "no such pairwise term");
    DCRF_REQUIRE(a_NxL && b_NxL && df_Nxd && L >= 1 && L <= 4096, "bad argument");
    DeviceGuard g(m->device);
    Ctx cx = m->cx();
    cudaStream_t s = m->stream;
    const Lattice& lt = m->terms[k]->lat;
    RowLayout rl(L);
    const int N = lt.N;
    float* rows[2];
    const float* host[2] = {a_NxL, b_NxL};
    for (int i = 0; i < 2; i++) {
      float* packed = upload_tmp(m, host[i], (size_t)N * L * sizeof(float));
      rows[i] = dev_alloc<float>((size_t)N * rl.Lp, s);
      launch_unpack_rows(cx, rl, N, packed, rows[i]);
      dev_free(packed, s);
    }
    const size_t vn = ((size_t)lt.V + 1) * rl.Lp;
    float* A = dev_alloc<float>(vn, s);
    float* B = dev_alloc<float>(vn, s);
    float* df = dev_alloc<float>((size_t)N * lt.d, s);
    lattice_gradient(cx, lt, rl, rows[0], rows[1], A, B, df);
    DCRF_CUDA(cudaMemcpyAsync(df_Nxd, df, (size_t)N * lt.d * sizeof(float), cudaMemcpyDeviceToHost, s));
    for (void* p : {(void*)rows[0], (void*)rows[1], (void*)A, (void*)B, (void*)df}) dev_free(p, s);
    DCRF_CUDA(cudaGetLastError());
    DCRF_CUDA(cudaStreamSynchronize(s));
  });
}
// ---- energies of a labelling: DenseCRF::unaryEnergy / pairwiseEnergy (densecrf.cpp:141-177, the "Debug" block) -------------
namespace {
struct DevScratch {  // device buffers of one call, released in stream order however the call ends
  cudaStream_t s;
  std::vector<void*> v;
  void add(void* p) { v.push_back(p); }
  ~DevScratch() {
    for (void* p : v) dev_free(p, s);
  }
};
__global__ void k_one_hot(int N, int L, int Lp, const int16_t* __restrict__ lab, float* __restrict__ Q) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (size_t)N * Lp) return;
  const size_t i = q / Lp;
  const int l = (int)(q - i * Lp);
  Q[q] = (l < L && lab[i] == l) ? 1.f : 0.f;  // Q(j, i) = (l[i] == j)  (:169)
}
// out[i] (+)= scale * src(l[i], i) for labels inside [0, L), else 0  (:149-150, :172-175)
__global__ void k_pick_label(int N, int L, int Lp, const int16_t* __restrict__ lab, const float* __restrict__ src, float scale,
                             int accumulate, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int l = lab[i];
  const float v = (l >= 0 && l < L) ? __fmul_rn(scale, src[(size_t)i * Lp + l]) : 0.f;
  out[i] = accumulate ? __fadd_rn(out[i], v) : v;
}
dcrf_status labelling_energy(dcrf_t* m, const int16_t* labels_N, bool pairwise, int term, float* out_N) {
  return guarded([&] {
    DCRF_REQUIRE(m && labels_N && out_N, "null argument");
    const int nt = (int)m->terms.size();
    DCRF_REQUIRE(!pairwise || (term >= -1 && term < nt), "no such pairwise term");
    DeviceGuard g(m->device);
    Ctx cx = m->cx();
    cudaStream_t s = m->stream;
    const int N = m->N, L = m->L, Lp = m->rl.Lp;
    const size_t n = (size_t)N * Lp;
    DevScratch tmp{s, {}};
    int16_t* lab = (int16_t*)upload_tmp(m, labels_N, (size_t)N * sizeof(int16_t));
    tmp.add(lab);
    float* r = dev_alloc<float>(N, s);
    tmp.add(r);
    DCRF_CUDA(cudaMemsetAsync(r, 0, (size_t)N * sizeof(float), s));
    float *Q = nullptr, *msg = nullptr;
    const unsigned nb = (unsigned)((N + 255) / 256);
    if (!pairwise) {
      if (m->U) {
        LaunchScope ls(cx.prof, s, "labelling_energy");
        k_pick_label<<<nb, 256, 0, s>>>(N, L, Lp, lab, m->U, 1.f, 0, r);
      }
    } else if (nt > 0) {
      Q = dev_alloc<float>(n, s);
      tmp.add(Q);
      msg = dev_alloc<float>(n, s);
      tmp.add(msg);
      {
        LaunchScope ls(cx.prof, s, "labelling_energy");
        k_one_hot<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(N, L, Lp, lab, Q);
      }
      for (int k = (term < 0 ? 0 : term); k < (term < 0 ? nt : term + 1); k++) {
        term_apply(m, m->terms[k], Q, SLICE_WRITE, false, msg);  // pairwise_[term]->apply(Q, Q)  (:170)
        LaunchScope ls(cx.prof, s, "labelling_energy");
        k_pick_label<<<nb, 256, 0, s>>>(N, L, Lp, lab, msg, -0.5f, term < 0 ? 1 : 0, r);  // r += ... over the terms (:160-161)
      }
    }
    DCRF_CUDA(cudaMemcpyAsync(out_N, r, (size_t)N * sizeof(float), cudaMemcpyDeviceToHost, s));
    DCRF_CUDA(cudaGetLastError());
    DCRF_CUDA(cudaStreamSynchronize(s));
  });
}
}  // namespace
dcrf_status dcrf_unary_energy(dcrf_t* m, const int16_t* labels_N, float* out_N) { return labelling_energy(m, labels_N, false, 0, out_N); }
dcrf_status dcrf_pairwise_energy(dcrf_t* m, const int16_t* labels_N, int term, float* out_N) {
  return labelling_energy(m, labels_N, true, term, out_N);
}

// ---- learning path: parameter vectors and DenseCRF::gradient ------------------------------------------------------
namespace {
int n_compat_params(const dcrf_model* m, const Term* t) {
  return t->ckind == DCRF_COMPAT_POTTS ? 1 : t->ckind == DCRF_COMPAT_DIAGONAL ? m->L : m->L * (m->L + 1) / 2;
}
int n_kernel_params(const Term* t) {
  return t->ktype == DCRF_CONST_KERNEL ? 0 : t->ktype == DCRF_DIAG_KERNEL ? t->d : t->d * t->d;
}
int n_params(const dcrf_model* m, int group) {
  int n = 0;
  if (group == DCRF_PARAMS_UNARY) return (int)m->ulog_L.size();
  for (const Term* t : m->terms) n += group == DCRF_PARAMS_LABEL_COMPATIBILITY ? n_compat_params(m, t) : n_kernel_params(t);
  return n;
}
// f_ of the term's DenseKernel, d x N row-major
const float* term_features(dcrf_model* m, Term* t) {
  if (!t->feat0) {
    DCRF_REQUIRE(t->ktype != DCRF_CONST_KERNEL, "a CONST_KERNEL term keeps no features");
    t->feat0 = dev_alloc<float>((size_t)t->d * m->N, m->stream);
    materialize_features(m->cx(), t->src, t->d, m->N, 0, nullptr, t->feat0);
  }
  return t->feat0;
}
// DenseKernel::setParameters (pairwise.cpp:139-151): initLattice(P * f_)
void term_set_kernel_params(dcrf_model* m, Term* t, const float* v) {
  const int d = t->d, N = m->N, np = n_kernel_params(t);
  if (np == 0) return;
  cudaStream_t s = m->stream;
  const float* f0 = term_features(m, t);
  float* dp = upload_tmp(m, v, (size_t)np * sizeof(float));
  float* tf = dev_alloc<float>((size_t)d * N, s);
  FeatureSource g{};
  g.kind = FEAT_GENERIC;
  g.feat = f0;
  materialize_features(m->cx(), g, d, N, t->ktype, dp, tf);
  // build the new lattice first, swap it in only when that succeeded (an allocation failure leaves the term as it was)
  Term fresh;
  fresh.ntype = t->ntype;
  g.feat = tf;
  try {
    term_init_lattice(m, &fresh, g, d);
  } catch (...) {
    if (!fresh.shared) fresh.lat.release(s);
    dev_free(fresh.norm, s);
    dev_free(dp, s);
    dev_free(tf, s);
    throw;
  }
  dev_free(dp, s);
  dev_free(tf, s);
  DCRF_CUDA(cudaStreamSynchronize(s));  // nothing queued may still read the old lattice
  for (cudaStream_t a : m->aux) DCRF_CUDA(cudaStreamSynchronize(a));
  if (!t->shared) t->lat.release(s);
  dev_free(t->norm, s);
  dev_free(t->valA, s);
  dev_free(t->csr_baked, s);
  t->shared = fresh.shared;
  t->lat = fresh.lat;
  t->norm = fresh.norm;
  t->valA = t->valB = nullptr, t->val_bytes = 0, t->csr_baked = nullptr;
  fresh.norm = nullptr;  // ownership moved
  t->kparams.assign(v, v + np);
  m->drop_graphs();
}
// Permutohedral::compute(x, reverse) of the term's lattice on device rows (no normalisation, no compatibility)
void term_compute(dcrf_model* m, Term* t, const float* x, bool reverse, float* out) {
  Ctx cx = m->cx();
  const float* fin = run_filter(cx, t->lat, m->rl, x, nullptr, t->valA, t->valB, reverse);
  launch_slice(cx, t->lat, m->rl, fin, nullptr, nullptr, SLICE_WRITE, nullptr, false, out);
}
struct GradScratch {  // N x Lp buffers of the backward pass
  float *an, *bn, *fa, *fb, *rin, *ones, *df, *dr;
};
// DenseKernel::featureGradient (pairwise.cpp:87-116) into w.df (N x d)
void term_feature_gradient(dcrf_model* m, Term* t, const float* a, const float* b, GradScratch& w) {
  Ctx cx = m->cx();
  const RowLayout rl = m->rl;
  const int N = m->N;
  if (t->ntype == DCRF_NO_NORMALIZATION) {
    lattice_gradient(cx, t->lat, rl, a, b, t->valA, t->valB, w.df);
    return;
  }
  const float *ka = a, *kb = b;  // arguments of the second kernelGradient
  if (t->ntype == DCRF_NORMALIZE_SYMMETRIC) {
    launch_scale_rows(cx, rl, N, a, t->norm, w.an);
    launch_scale_rows(cx, rl, N, b, t->norm, w.bn);
    term_compute(m, t, w.an, true, w.fa);
    term_compute(m, t, w.bn, false, w.fb);
    ka = w.an, kb = w.bn;
  } else if (t->ntype == DCRF_NORMALIZE_AFTER) {
    term_compute(m, t, b, false, w.fb);
    launch_scale_rows(cx, rl, N, a, t->norm, w.an);
    ka = w.an;
  } else {
    term_compute(m, t, a, true, w.fa);
    launch_scale_rows(cx, rl, N, b, t->norm, w.bn);
    kb = w.bn;
  }
  launch_feature_gradient_input(cx, rl, N, t->ntype, a, w.fb, w.fa, b, t->norm, w.rin);
  lattice_gradient(cx, t->lat, rl, w.rin, w.ones, t->valA, t->valB, w.dr);
  lattice_gradient(cx, t->lat, rl, ka, kb, t->valA, t->valB, w.df);
  launch_neg_add(cx, (size_t)N * t->lat.d, w.dr, w.df);
}
}  // namespace

dcrf_status dcrf_num_parameters(dcrf_t* m, int group, int* n) {
  return guarded([&] {
    DCRF_REQUIRE(m && n && group >= 0 && group <= 2, "bad argument");
    *n = n_params(m, group);
  });
}
dcrf_status dcrf_get_parameters(dcrf_t* m, int group, float* out) {
  return guarded([&] {
    DCRF_REQUIRE(m && group >= 0 && group <= 2, "bad argument");
    if (n_params(m, group) == 0) return;
    DCRF_REQUIRE(out, "null argument");
    const int L = m->L;
    if (group == DCRF_PARAMS_UNARY) {
      const int F = m->ulog_F;
      for (int l = 0; l < L; l++)
        for (int f = 0; f < F; f++) out[l + (size_t)L * f] = m->ulog_L[(size_t)l * F + f];
      return;
    }
    for (const Term* t : m->terms) {
      if (group == DCRF_PARAMS_LABEL_COMPATIBILITY) {
        if (t->ckind != DCRF_COMPAT_MATRIX) {
          out = std::copy(t->compat_host.begin(), t->compat_host.end(), out);
        } else {
          for (int i = 0; i < L; i++)
            for (int j = i; j < L; j++) *out++ = t->compat_host[(size_t)i * L + j];
        }
      } else {
        const int np = n_kernel_params(t), d = t->d;
        if (!t->kparams.empty()) {
          out = std::copy(t->kparams.begin(), t->kparams.end(), out);
        } else {
          for (int e = 0; e < np; e++) *out++ = (t->ktype == DCRF_DIAG_KERNEL || e % d == e / d) ? 1.f : 0.f;
        }
      }
    }
  });
}
dcrf_status dcrf_set_parameters(dcrf_t* m, int group, const float* v, int n) {
  return guarded([&] {
    DCRF_REQUIRE(m && group >= 0 && group <= 2, "bad argument");
    DCRF_REQUIRE(n == n_params(m, group), "this model has " + std::to_string(n_params(m, group)) + " parameters of that group, got " +
                                              std::to_string(n));
    if (n == 0) return;
    DCRF_REQUIRE(v, "null argument");
    DeviceGuard g(m->device);
    const int L = m->L;
    DCRF_CUDA(cudaStreamSynchronize(m->stream));
    if (group == DCRF_PARAMS_UNARY) {
      const int F = m->ulog_F;
      for (int l = 0; l < L; l++)
        for (int f = 0; f < F; f++) m->ulog_L[(size_t)l * F + f] = v[l + (size_t)L * f];
      float* dL = upload_tmp(m, m->ulog_L.data(), m->ulog_L.size() * sizeof(float));
      launch_logistic_unary(m->cx(), m->rl, m->N, F, dL, m->ulog_f, m->U);
      dev_free(dL, m->stream);
      DCRF_CUDA(cudaGetLastError());
      DCRF_CUDA(cudaStreamSynchronize(m->stream));
      return;
    }
    for (Term* t : m->terms) {
      if (group == DCRF_PARAMS_LABEL_COMPATIBILITY) {
        const int np = n_compat_params(m, t);
        if (t->ckind != DCRF_COMPAT_MATRIX) {
          term_set_compat(m, t, t->ckind, v);
        } else {
          std::vector<float> w((size_t)L * L);
          int k = 0;
          for (int i = 0; i < L; i++)
            for (int j = i; j < L; j++, k++) w[(size_t)i * L + j] = w[(size_t)j * L + i] = v[k];
          term_set_compat(m, t, t->ckind, w.data());
        }
        v += np;
      } else {
        term_set_kernel_params(m, t, v);
        v += n_kernel_params(t);
      }
    }
    m->drop_graphs();  // Potts weights are launch arguments of the captured kernels
    DCRF_CUDA(cudaGetLastError());
    DCRF_CUDA(cudaStreamSynchronize(m->stream));
  });
}

dcrf_status dcrf_gradient(dcrf_t* m, int n_iterations, const dcrf_objective* obj, double* value, float* unary_grad,
                          float* lbl_cmp_grad, float* kernel_grad) {
  return guarded([&] {
    DCRF_REQUIRE(m && obj && obj->gt && value, "null argument");
    DCRF_REQUIRE(n_iterations >= 0, "negative iteration count");
    DCRF_REQUIRE(obj->kind >= 0 && obj->kind <= 2, "unknown objective");
    DeviceGuard g(m->device);
    Ctx cx = m->cx();
    cudaStream_t s = m->stream;
    const RowLayout rl = m->rl;
    const int N = m->N, L = m->L, Lp = rl.Lp, nt = (int)m->terms.size();
    const size_t n = (size_t)N * Lp;
    if (n_params(m, DCRF_PARAMS_UNARY) == 0) unary_grad = nullptr;
    if (n_params(m, DCRF_PARAMS_LABEL_COMPATIBILITY) == 0) lbl_cmp_grad = nullptr;
    if (n_params(m, DCRF_PARAMS_KERNEL) == 0) kernel_grad = nullptr;
    if (obj->kind == DCRF_OBJECTIVE_HAMMING) {
      DCRF_REQUIRE(obj->class_weight && obj->n_class_weight >= 1, "the Hamming objective needs class weights");
      for (int i = 0; i < N; i++)
        DCRF_REQUIRE(obj->gt[i] < 0 || obj->gt[i] >= L || obj->gt[i] < obj->n_class_weight, "class weights do not cover every label of gt");
    }
    std::vector<void*> owned;  // every device buffer of this call
    auto rows = [&]() {
      float* p = dev_alloc<float>(n, s);
      owned.push_back(p);
      return p;
    };
    struct Cleanup {
      std::vector<void*>& v;
      cudaStream_t s;
      ~Cleanup() {
        for (void* p : v) dev_free(p, s);
      }
    } cleanup{owned, s};

    // mean field with every Q kept (densecrf.cpp:239-252)
    for (Term* t : m->terms) ensure_val(m, t);
    ensure_scratch_if_staged(m);
    ensure_aux(m, nt);
    std::vector<float*> Q(n_iterations + 1);
    ensure_baked(m);
    Q[0] = rows();
    mf_init(m, Q[0]);
    for (int it = 0; it < n_iterations; it++) {
      Q[it + 1] = rows();
      DCRF_CUDA(cudaMemcpyAsync(Q[it + 1], Q[it], n * sizeof(float), cudaMemcpyDeviceToDevice, s));
      mf_step(m, Q[it + 1], true);
    }
    const float* Qn = Q[n_iterations];

    // objective value and b = d_mul_Q (densecrf.cpp:255-257)
    float *b = rows(), *b2 = rows();
    int16_t* gt = (int16_t*)upload_tmp(m, obj->gt, (size_t)N * sizeof(int16_t));
    owned.push_back(gt);
    double* dsc = dev_alloc<double>(2 * (size_t)L + 1, s);
    owned.push_back(dsc);
    if (obj->kind != DCRF_OBJECTIVE_IOU) {
      float* cw = nullptr;
      if (obj->kind == DCRF_OBJECTIVE_HAMMING) {
        std::vector<float> w(L, 0.f);
        std::copy(obj->class_weight, obj->class_weight + std::min(L, obj->n_class_weight), w.begin());
        cw = upload_tmp(m, w.data(), (size_t)L * sizeof(float));
        owned.push_back(cw);
      }
      launch_objective_point(cx, rl, N, obj->kind, gt, obj->robust, cw, Qn, b, dsc);
      DCRF_CUDA(cudaMemcpyAsync(value, dsc, sizeof(double), cudaMemcpyDeviceToHost, s));
      DCRF_CUDA(cudaStreamSynchronize(s));
    } else {
      float *in_rows = rows(), *un_rows = rows();
      const float one = 1.f;
      float* d_one = upload_tmp(m, &one, sizeof(float));
      owned.push_back(d_one);
      launch_iou_rows(cx, rl, N, gt, Qn, in_rows, un_rows);
      launch_outer_sums(cx, N, in_rows, Lp, 1, L, d_one, 0, 0, 1, false, dsc);
      launch_outer_sums(cx, N, un_rows, Lp, 1, L, d_one, 0, 0, 1, false, dsc + L);
      launch_iou_b(cx, rl, N, gt, Qn, dsc, dsc + L, b);
      std::vector<double> h(2 * (size_t)L);
      DCRF_CUDA(cudaMemcpyAsync(h.data(), dsc, h.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
      DCRF_CUDA(cudaStreamSynchronize(s));
      double r = 0;
      for (int l = 0; l < L; l++) r += h[l] / (h[L + l] + 1e-20);
      *value = r / L;
    }
    launch_sum_and_normalize(cx, rl, N, b, Qn, false, b2);  // sumAndNormalize(b, b, Q[n])  (:258)
    std::swap(b, b2);

    // one slot of doubles per sum over the variables that the backward pass needs, reduced on the host at the end
    const int F = m->ulog_F;
    std::vector<size_t> lbl_slot, ker_slot;
    size_t total = 0;
    const size_t un_sz = unary_grad ? (size_t)L * F : 0;
    const size_t un_slot0 = total;
    total += un_sz * (n_iterations + 1);
    for (int it = 0; it < n_iterations; it++)
      for (Term* t : m->terms) {
        lbl_slot.push_back(total);
        if (lbl_cmp_grad) total += t->ckind == DCRF_COMPAT_MATRIX ? (size_t)L * L : (size_t)L;
        ker_slot.push_back(total);
        if (kernel_grad) total += n_kernel_params(t);
      }
    double* res = dev_alloc<double>(std::max<size_t>(total, 1), s);
    owned.push_back(res);
    int un_count = 0;
    auto unary_gradient = [&]() {  // LogisticUnaryEnergy::gradient: b * f_^T (unary.cpp:67-71)
      if (unary_grad) launch_outer_sums(cx, N, b, Lp, 1, L, m->ulog_f, 1, (size_t)N, F, false, res + un_slot0 + un_sz * un_count++);
    };
    unary_gradient();  // (:261-262)

    float *tmp1 = nullptr, *tmp2 = nullptr, *fQ = nullptr, *lblQ = nullptr;
    GradScratch w{};
    if (n_iterations > 0 && nt > 0) {
      tmp1 = rows(), tmp2 = rows();
      if (lbl_cmp_grad) fQ = rows();
      if (kernel_grad) {
        lblQ = rows();
        DCRF_CUDA(cudaMemsetAsync(lblQ, 0, n * sizeof(float), s));  // MatrixCompatibility leaves the pad columns alone
        w.an = rows(), w.bn = rows(), w.fa = rows(), w.fb = rows(), w.rin = rows(), w.ones = rows();
        launch_fill(cx, w.ones, n, 1.0f);
        int dmax = 1;
        for (Term* t : m->terms) dmax = std::max(dmax, t->lat.d);
        w.df = dev_alloc<float>((size_t)N * dmax, s);
        w.dr = dev_alloc<float>((size_t)N * dmax, s);
        owned.push_back(w.df);
        owned.push_back(w.dr);
      }
    }
    // inverse message passing (:268-295)
    for (int it = n_iterations - 1; it >= 0; it--) {
      if (nt == 0) {  // tmp1 = 0: b = sumAndNormalize(0, Q) = 0
        DCRF_CUDA(cudaMemsetAsync(b, 0, n * sizeof(float), s));
        unary_gradient();
        continue;
      }
      DCRF_CUDA(cudaMemsetAsync(tmp1, 0, n * sizeof(float), s));
      for (int k = 0; k < nt; k++) {
        Term* t = m->terms[k];
        const size_t slot = (size_t)(n_iterations - 1 - it) * nt + k;
        if (lbl_cmp_grad) {  // compatibility_->gradient(b, kernel_->apply(Q))  (pairwise.cpp:190-195)
          const bool pre = t->pre(), post = t->post();
          const float* fin = run_filter(cx, t->lat, rl, Q[it], pre ? t->norm : nullptr, t->valA, t->valB, false);
          launch_slice(cx, t->lat, rl, fin, post ? t->norm : nullptr, nullptr, SLICE_WRITE, nullptr, false, fQ);
          launch_outer_sums(cx, N, b, Lp, 1, L, fQ, Lp, 1, L, t->ckind != DCRF_COMPAT_MATRIX, res + lbl_slot[slot]);
        }
        if (kernel_grad && t->ktype != DCRF_CONST_KERNEL) {  // kernel_->gradient(b, compatibility_->apply(Q))  (:202-207)
          if (t->ckind == DCRF_COMPAT_MATRIX) {
            launch_matrix_compat(cx, rl, N, t->Wm, Q[it], SLICE_WRITE, nullptr, false, lblQ);
          } else {
            CompatDev c = compat_of(t);
            launch_compat_pointwise(cx, rl, N, c, Q[it], lblQ);
          }
          term_feature_gradient(m, t, b, lblQ, w);
          const float* f0 = term_features(m, t);
          const int d = t->d;
          if (t->ktype == DCRF_DIAG_KERNEL)  // (f_ .* fg).rowwise().sum()  (pairwise.cpp:157)
            launch_outer_sums(cx, N, f0, 1, (size_t)N, d, w.df, d, 1, d, true, res + ker_slot[slot]);
          else  // fg * f_^T  (:159)
            launch_outer_sums(cx, N, w.df, d, 1, d, f0, 1, (size_t)N, d, false, res + ker_slot[slot]);
        }
        term_apply(m, t, b, SLICE_WRITE, false, tmp2, /*transpose=*/true);
        launch_add_rows(cx, rl, N, tmp2, tmp1);
      }
      launch_sum_and_normalize(cx, rl, N, tmp1, Q[it], true, b2);  // (:290)
      std::swap(b, b2);
      unary_gradient();
    }
    std::vector<double> h(std::max<size_t>(total, 1));
    DCRF_CUDA(cudaGetLastError());
    if (total > 0) DCRF_CUDA(cudaMemcpyAsync(h.data(), res, total * sizeof(double), cudaMemcpyDeviceToHost, s));
    DCRF_CUDA(cudaStreamSynchronize(s));

    // the reference's float accumulation over iterations and terms, in its order
    if (unary_grad) {
      for (size_t e = 0; e < un_sz; e++) unary_grad[e] = 0.f;
      for (int c = 0; c < un_count; c++)
        for (int l = 0; l < L; l++)
          for (int f = 0; f < F; f++) {
            const float gv = (float)h[un_slot0 + un_sz * c + (size_t)l * F + f];
            float& o = unary_grad[l + (size_t)L * f];
            o = c == 0 ? gv : o + gv;
          }
    }
    if (lbl_cmp_grad) std::fill(lbl_cmp_grad, lbl_cmp_grad + n_params(m, DCRF_PARAMS_LABEL_COMPATIBILITY), 0.f);
    if (kernel_grad) std::fill(kernel_grad, kernel_grad + n_params(m, DCRF_PARAMS_KERNEL), 0.f);
    for (int it = n_iterations - 1; it >= 0 && nt > 0; it--) {
      int ip = 0, ik = 0;
      for (int k = 0; k < nt; k++) {
        Term* t = m->terms[k];
        const size_t slot = (size_t)(n_iterations - 1 - it) * nt + k;
        if (lbl_cmp_grad) {
          const double* G = h.data() + lbl_slot[slot];
          float* o = lbl_cmp_grad + ip;
          if (t->ckind == DCRF_COMPAT_POTTS) {  // -(b .* Q).sum()  (labelcompatibility.cpp:59)
            double sum = 0;
            for (int l = 0; l < L; l++) sum += G[l];
            o[0] += -(float)sum;
          } else if (t->ckind == DCRF_COMPAT_DIAGONAL) {  // (b .* Q).rowwise().sum()  (:73-75)
            for (int l = 0; l < L; l++) o[l] += (float)G[l];
          } else {  // g = b * Q^T; r[k] = g(i,j) + (i != j ? g(j,i) : 0)  (:101-108)
            int e = 0;
            for (int i = 0; i < L; i++)
              for (int j = i; j < L; j++, e++)
                o[e] += (float)G[(size_t)i * L + j] + (i != j ? (float)G[(size_t)j * L + i] : 0.f);
          }
          ip += n_compat_params(m, t);
        }
        if (kernel_grad) {
          const double* G = h.data() + ker_slot[slot];
          float* o = kernel_grad + ik;
          const int d = t->d;
          if (t->ktype == DCRF_DIAG_KERNEL) {
            for (int j = 0; j < d; j++) o[j] += (float)G[j];
          } else if (t->ktype == DCRF_FULL_KERNEL) {  // p(r, c) at r + c*d (column-major resize, pairwise.cpp:160)
            for (int r = 0; r < d; r++)
              for (int c = 0; c < d; c++) o[r + (size_t)c * d] += (float)G[(size_t)r * d + c];
          }
          ik += n_kernel_params(t);
        }
      }
    }
  });
}

dcrf_status dcrf_set_iteration_mode(dcrf_t* m, int mode) {
  return guarded([&] {
    DCRF_REQUIRE(m && mode >= 0 && mode <= 2, "bad iteration mode");
    m->iter_mode = mode;
    m->drop_graphs();
  });
}
dcrf_status dcrf_set_matrix_mode(dcrf_t* m, int mode) {
  return guarded([&] {
    DCRF_REQUIRE(m && mode >= 0 && mode <= 2, "bad matrix compatibility mode");
    m->matrix_mode = mode;
    m->drop_graphs();
  });
}
dcrf_status dcrf_set_kl_mode(dcrf_t* m, int mode) {
  return guarded([&] {
    DCRF_REQUIRE(m && (mode == 0 || mode == 1), "bad KL mode");
    m->kl_mode = mode;
  });
}
dcrf_status dcrf_set_kernel_variant(dcrf_t* m, int mask) {
  return guarded([&] {
    DCRF_REQUIRE(m && mask >= 0 && mask <= VAR_ALL, "bad kernel variant mask");
    m->variant = mask;
    m->drop_graphs();
  });
}
dcrf_status dcrf_set_softmax_mode(dcrf_t* m, int fast) {
  return guarded([&] {
    DCRF_REQUIRE(m, "null argument");
    m->fast_softmax = fast != 0;
    m->drop_graphs();
  });
}

}  // extern "C"
