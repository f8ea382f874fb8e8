// Permutohedral lattice construction on the GPU.
//
// Replaces Permutohedral::init + HashTable (pydensecrf/densecrf/src/permutohedral.cpp:54-131,
// 140-321 in /root/reference).  Same result, different construction:
//
//   k_elevate     one thread per pixel: elevate, round, rank, barycentric weights.  Every float
//                 operation is an explicitly rounded, unfused intrinsic in the order the reference's
//                 SSE2 code performs it, because the enclosing simplex (an integer result) is
//                 decided by those roundings.
//   k_insert      every (pixel, corner) offers its vertex key to an open-addressing table whose
//                 slots hold a *contribution index* t = pixel*(d+1)+corner instead of a key: the
//                 slot's key is re-derived from t's pixel descriptor, is immutable once the slot is
//                 claimed, and atomicMin leaves the smallest t per vertex = the reference's
//                 first-touch order.  Lanes of a warp that hold the same key are aggregated with
//                 __match_any_sync so only one of them probes.
//   scan + k_assign   rank the first-touch indices -> vertex ids identical to the reference's
//                 insertion-order ids (so `offset_` matches element for element).
//   radix sort    contributions grouped by vertex, ascending (pixel, corner) inside a vertex: the
//                 CSR that lets splat run as a deterministic gather in the reference's summation
//                 order instead of as float atomics.
//   k_neighbours  2(d+1) lookups per vertex.
//
// Batches (SURVEY 8e "super-lattice"): with npi > 0 the N variables are B images of npi pixels each and the image
// index is one more (exact) key coordinate: vertices of different images never meet, the first-touch numbering walks
// image after image, so tables and ids are the per-image lattices concatenated with id offsets -- bit for bit what B
// separate builds give -- while every kernel here and in the filter runs once over the whole batch.
//
// CUB (bundled with the CUDA toolkit) provides the prefix sums and the radix sort; everything
// lattice-specific is written here.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "common.h"

namespace dcrf {
namespace {

constexpr uint32_t kEmpty = 0xFFFFFFFFu;
constexpr int kThreads = 256;

struct ElevateParams {
  int N, Npad, d;
  int npi;  // pixels per image of a batch (0: one image)
  FeatureSource src;
  float scale[kMaxFeatureDims];
  float inv_dp1, dp1;
};

__device__ __forceinline__ float feature_of(const FeatureSource& s, int k, int i, int N) {
  if (s.kind == FEAT_GENERIC) return __ldg(s.feat + (size_t)k * N + i);
  if (s.kind == FEAT_ND) {
    if (k < s.nsp) return __fdiv_rn((float)((i / s.stride[k]) % s.dims[k]), s.sdims[k]);
    const size_t e = (size_t)i * s.nch + (k - s.nsp);
    const float v = s.img_is_float ? __ldg(static_cast<const float*>(s.img) + e) : (float)__ldg(static_cast<const uint8_t*>(s.img) + e);
    return __fdiv_rn(v, s.schan[k - s.nsp]);
  }
  // DenseCRF2D::addPairwiseGaussian / Bilateral (densecrf.cpp:61-81): int coordinate / byte divided
  // by the float sigma, pixel index = y*W + x
  if (k == 0) return __fdiv_rn((float)(i % s.W), s.s[0]);
  if (k == 1) return __fdiv_rn((float)(s.H > 0 ? (i / s.W) % s.H : i / s.W), s.s[1]);  // s.H: image height of a batch
  return __fdiv_rn((float)__ldg(s.rgb + (size_t)i * 3 + (k - 2)), s.s[k]);
}

// D > 0: compile-time feature dimension (everything stays in registers); D == 0: runtime d
template <int D>
__global__ void __launch_bounds__(kThreads) k_elevate(ElevateParams p, float* __restrict__ bary,
                                                     int16_t* __restrict__ rem, uint8_t* __restrict__ rnk) {
  constexpr int MD = D ? D : kMaxFeatureDims;
  const int d = D ? D : p.d;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.Npad) return;

  float el[MD + 1], r0[MD + 1], rk[MD + 1], b[MD + 2];

  // elevate: cf = f*scale; el[j] = sm - j*cf; sm += cf   (permutohedral.cpp:201-207)
  // pixels past N are the zero-feature lanes of the reference's last 4-wide block (:196)
  float sm = 0.f;
#pragma unroll
  for (int j = MD; j > 0; j--) {
    if (j <= d) {
      const float f = (i < p.N) ? feature_of(p.src, j - 1, i, p.N) : 0.f;
      const float cf = __fmul_rn(f, p.scale[j - 1]);
      el[j] = __fsub_rn(sm, __fmul_rn((float)j, cf));
      sm = __fadd_rn(sm, cf);
    }
  }
  el[0] = sm;

  // nearest multiple of d+1, ties to even (:210-220)
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k <= MD; k++) {
    if (k <= d) {
      float v = __fmul_rn(p.inv_dp1, el[k]);
      v = __int2float_rn(__float2int_rn(v));
      r0[k] = __fmul_rn(v, p.dp1);
      sum = __fadd_rn(sum, v);
      rk[k] = 0.f;
    }
  }
  // rank residuals, strict '<' (:223-233)
#pragma unroll
  for (int a = 0; a < MD; a++) {
    if (a < d) {
      const float da = __fsub_rn(el[a], r0[a]);
#pragma unroll
      for (int c = a + 1; c <= MD; c++) {
        if (c <= d) {
          const float dc = __fsub_rn(el[c], r0[c]);
          const float lt = (da < dc) ? 1.f : 0.f;
          rk[a] = __fadd_rn(rk[a], lt);
          rk[c] = __fadd_rn(rk[c], __fsub_rn(1.f, lt));
        }
      }
    }
  }
  // back onto the plane (:236-242)
#pragma unroll
  for (int k = 0; k <= MD; k++) {
    if (k <= d) {
      rk[k] = __fadd_rn(rk[k], sum);
      const float add = (rk[k] < 0.f) ? p.dp1 : 0.f;
      const float sub = (rk[k] >= p.dp1) ? p.dp1 : 0.f;
      const float adj = __fsub_rn(add, sub);
      rk[k] = __fadd_rn(rk[k], adj);
      r0[k] = __fadd_rn(r0[k], adj);
    }
  }
  // barycentric weights (:245-263); b[q] is touched through selects so it stays in registers
#pragma unroll
  for (int q = 0; q <= MD + 1; q++) b[q] = 0.f;
#pragma unroll
  for (int k = 0; k <= MD; k++) {
    if (k <= d) {
      const float v = __fmul_rn(__fsub_rn(el[k], r0[k]), p.inv_dp1);
      const int pos = (int)__fsub_rn((float)d, rk[k]);
#pragma unroll
      for (int q = 0; q <= MD + 1; q++) {
        if (q == pos) b[q] = __fadd_rn(b[q], v);
        if (q == pos + 1) b[q] = __fsub_rn(b[q], v);
      }
    }
  }
  float wrap = 0.f;
#pragma unroll
  for (int q = 0; q <= MD + 1; q++)
    if (q == d + 1) wrap = b[q];
  b[0] = __fadd_rn(b[0], __fadd_rn(1.f, wrap));

  const size_t t0 = (size_t)i * (d + 1);
#pragma unroll
  for (int k = 0; k <= MD; k++)
    if (k <= d) bary[t0 + k] = b[k];
  // pixel descriptor: enough to re-derive any of the pixel's d+1 vertex keys later
#pragma unroll
  for (int k = 0; k < MD; k++)
    if (k < d) {
      rem[(size_t)i * d + k] = (int16_t)(int)r0[k];
      rnk[(size_t)i * d + k] = (uint8_t)(int)rk[k];
    }
}

// coordinate k of the vertex with remainder r of a pixel: rem0[k] + canonical[r][rank[k]] truncated to
// int16 (permutohedral.cpp:171-176, :270)
__device__ __forceinline__ int16_t key_coord(int16_t rem0, int rank, int r, int d) {
  return (int16_t)(rem0 + ((rank <= d - r) ? r : r - (d + 1)));
}

__device__ __forceinline__ uint64_t mix64(uint64_t h) {
  h ^= h >> 33;
  h *= 0xff51afd7ed558ccdull;
  h ^= h >> 33;
  h *= 0xc4ceb9fe1a85ec53ull;
  h ^= h >> 33;
  return h;
}
template <int MD>
__device__ __forceinline__ uint64_t hash_key(const int16_t (&k)[MD], int d) {
  uint64_t h = 0x9E3779B97F4A7C15ull;
#pragma unroll
  for (int a = 0; a < MD; a++)
    if (a < d) h = (h ^ (uint16_t)k[a]) * 0x100000001b3ull;
  return mix64(h);
}

__device__ __forceinline__ uint64_t hash_img(uint64_t h, int img) { return mix64(h ^ ((uint64_t)(uint32_t)img * 0x9E3779B97F4A7C15ull)); }

// does the contribution `cur` (some pixel corner) name the same vertex as key[] (of image `img` when npi > 0)?
template <int MD>
__device__ __forceinline__ bool same_vertex(uint32_t cur, const int16_t (&key)[MD], int d,
                                            const int16_t* __restrict__ rem, const uint8_t* __restrict__ rnk, int npi, int img) {
  const uint32_t dp1 = d + 1;
  const uint32_t pc = cur / dp1;
  const int rc = (int)(cur - pc * dp1);
  if (npi && (int)(pc / (uint32_t)npi) != img) return false;
  bool eq = true;
#pragma unroll
  for (int a = 0; a < MD; a++)
    if (a < d) eq &= (key_coord(__ldg(rem + (size_t)pc * d + a), __ldg(rnk + (size_t)pc * d + a), rc, d) == key[a]);
  return eq;
}

template <int D>
__global__ void __launch_bounds__(kThreads)
    k_insert(int Npad, int d_rt, int npi, const int16_t* __restrict__ rem, const uint8_t* __restrict__ rnk,
             uint32_t* __restrict__ table, uint32_t mask, uint32_t* __restrict__ slot_of) {
  constexpr int MD = D ? D : kMaxFeatureDims;
  const int d = D ? D : d_rt;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < Npad;
  const unsigned vmask = __ballot_sync(0xffffffffu, valid);
  if (!valid) return;
  const int lane = threadIdx.x & 31;
  const int img = npi ? i / npi : 0;

  int16_t r0[MD];
  int rk[MD];
#pragma unroll
  for (int a = 0; a < MD; a++)
    if (a < d) {
      r0[a] = rem[(size_t)i * d + a];
      rk[a] = rnk[(size_t)i * d + a];
    }

  for (int r = 0; r <= d; r++) {
    int16_t key[MD];
#pragma unroll
    for (int a = 0; a < MD; a++)
      if (a < d) key[a] = key_coord(r0[a], rk[a], r, d);
    uint64_t h = hash_key<MD>(key, d);
    if (npi) h = hash_img(h, img);
    const uint32_t t = (uint32_t)i * (uint32_t)(d + 1) + (uint32_t)r;

    // warp aggregation: neighbouring pixels mostly share vertices; one lane per distinct key probes
    const unsigned peers = __match_any_sync(vmask, h);
    const int leader = __ffs(peers) - 1;
    bool same = __shfl_sync(peers, img, leader) == img;
#pragma unroll
    for (int a = 0; a < MD; a++)
      if (a < d) same &= (__shfl_sync(peers, (int)key[a], leader) == (int)key[a]);

    uint32_t slot = 0;
    if (lane == leader || !same) {
      slot = (uint32_t)h & mask;
      for (;;) {
        uint32_t cur = *((volatile uint32_t*)(table + slot));
        if (cur == kEmpty) {
          cur = atomicCAS(table + slot, kEmpty, t);
          if (cur == kEmpty) break;  // claimed an empty slot: we are (for now) its representative
        }
        if (same_vertex<MD>(cur, key, d, rem, rnk, npi, img)) {
          if (t < cur) atomicMin(table + slot, t);
          break;
        }
        slot = (slot + 1) & mask;
      }
    }
    const uint32_t lslot = __shfl_sync(peers, slot, leader);
    if (same) slot = lslot;
    slot_of[t] = slot;
  }
}

__global__ void k_mark_first(uint32_t T, const uint32_t* __restrict__ table, const uint32_t* __restrict__ slot_of,
                             int* __restrict__ flag) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  flag[t] = (table[slot_of[t]] == t) ? 1 : 0;
}

template <int D>
__global__ void __launch_bounds__(kThreads)
    k_assign(uint32_t T, uint32_t E, int d_rt, const uint32_t* __restrict__ table, const uint32_t* __restrict__ slot_of,
             const int* __restrict__ vidx, const float* __restrict__ bary, const int16_t* __restrict__ rem,
             const uint8_t* __restrict__ rnk, PixEntry* __restrict__ pix, uint32_t* __restrict__ sort_key,
             uint32_t* __restrict__ sort_val, int* __restrict__ count, int16_t* __restrict__ keys, int npi,
             int* __restrict__ vimg) {
  constexpr int MD = D ? D : kMaxFeatureDims;
  const int d = D ? D : d_rt;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const uint32_t rep = table[slot_of[t]];
  const int vid = vidx[rep];
  if (t < E) {
    PixEntry e;
    e.row = vid + 1;
    e.w = bary[t];
    pix[t] = e;
    sort_key[t] = (uint32_t)vid;
    sort_val[t] = t;
    atomicAdd(count + vid, 1);
  }
  if (rep == t) {
    const uint32_t pc = t / (uint32_t)(d + 1);
    const int rc = (int)(t - pc * (uint32_t)(d + 1));
#pragma unroll
    for (int a = 0; a < MD; a++)
      if (a < d) keys[(size_t)vid * d + a] = key_coord(rem[(size_t)pc * d + a], rnk[(size_t)pc * d + a], rc, d);
    if (npi) vimg[vid] = (int)(pc / (uint32_t)npi);
  }
}

// after this the table maps key -> vertex id
__global__ void k_table_to_ids(uint32_t T, const int* __restrict__ flag, const int* __restrict__ vidx,
                               const uint32_t* __restrict__ slot_of, uint32_t* __restrict__ table) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  if (flag[t]) table[slot_of[t]] = (uint32_t)vidx[t];
}

__global__ void k_fill_csr(uint32_t E, int dp1, const uint32_t* __restrict__ sorted_t, const float* __restrict__ bary,
                           CsrEntry* __restrict__ csr) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const uint32_t t = sorted_t[e];
  CsrEntry c;
  c.pix = (int)(t / (uint32_t)dp1);
  c.w = bary[t];
  csr[e] = c;
}

// splat work order: keep the coarse first-touch (scan-line) order for cache locality but, inside a window of
// 4096 vertices, line vertices up by CSR row length so the lanes of a warp loop about equally often
__global__ void k_order_keys(int V, const int* __restrict__ count, uint32_t* __restrict__ key, uint32_t* __restrict__ val) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  key[v] = ((uint32_t)(v >> 12) << 8) | (uint32_t)min(count[v], 255);
  val[v] = (uint32_t)v;
}

template <int MD>
__device__ __forceinline__ int lookup(const int16_t (&key)[MD], int d, const uint32_t* __restrict__ table,
                                      uint32_t mask, const int16_t* __restrict__ keys, const int* __restrict__ vimg, int img) {
  uint64_t h = hash_key<MD>(key, d);
  if (vimg) h = hash_img(h, img);
  uint32_t slot = (uint32_t)h & mask;
  for (;;) {
    const uint32_t cur = __ldg(table + slot);
    if (cur == kEmpty) return -1;
    bool eq = !vimg || __ldg(vimg + cur) == img;
#pragma unroll
    for (int a = 0; a < MD; a++)
      if (a < d) eq &= (__ldg(keys + (size_t)cur * d + a) == key[a]);
    if (eq) return (int)cur;
    slot = (slot + 1) & mask;
  }
}

// neighbours along lattice axis j: key -/+ (1,...,1) with coordinate j moved by +/- d instead
// (permutohedral.cpp:305-318); the implied (d+1)-th coordinate is not part of the key
template <int D>
__global__ void __launch_bounds__(kThreads)
    k_neighbours(int V, int d_rt, const int16_t* __restrict__ keys, const uint32_t* __restrict__ table, uint32_t mask,
                 const int* __restrict__ vimg, int2* __restrict__ nbr) {
  constexpr int MD = D ? D : kMaxFeatureDims;
  const int d = D ? D : d_rt;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (v >= V) return;
  const int img = vimg ? vimg[v] : 0;
  int16_t n1[MD], n2[MD];
#pragma unroll
  for (int a = 0; a < MD; a++)
    if (a < d) {
      const int16_t k = keys[(size_t)v * d + a];
      n1[a] = (int16_t)(k - 1);
      n2[a] = (int16_t)(k + 1);
      if (a == j) {
        n1[a] = (int16_t)(k + d);
        n2[a] = (int16_t)(k - d);
      }
    }
  int2 o;
  o.x = lookup<MD>(n1, d, table, mask, keys, vimg, img) + 1;
  o.y = lookup<MD>(n2, d, table, mask, keys, vimg, img) + 1;
  nbr[(size_t)j * V + v] = o;
}

inline unsigned blocks_for(size_t n, int threads = kThreads) { return (unsigned)((n + threads - 1) / threads); }

template <int D>
void build_impl(Ctx cx, const ElevateParams& ep, Lattice* L) {
  cudaStream_t s = cx.stream;
  const int d = ep.d, N = ep.N, Npad = ep.Npad, dp1 = d + 1;
  const uint32_t T = (uint32_t)((size_t)Npad * dp1), E = (uint32_t)((size_t)N * dp1);

  float* bary = dev_alloc<float>(T, s);
  int16_t* rem = dev_alloc<int16_t>((size_t)Npad * d, s);
  uint8_t* rnk = dev_alloc<uint8_t>((size_t)Npad * d, s);
  {
    LaunchScope ls(cx.prof, s, "lattice_elevate");
    k_elevate<D><<<blocks_for(Npad), kThreads, 0, s>>>(ep, bary, rem, rnk);
  }

  uint64_t cap = 1024;
  while (cap < 2ull * T) cap <<= 1;
  uint32_t* table = dev_alloc<uint32_t>(cap, s);
  DCRF_CUDA(cudaMemsetAsync(table, 0xff, cap * sizeof(uint32_t), s));
  uint32_t* slot_of = dev_alloc<uint32_t>(T, s);
  {
    LaunchScope ls(cx.prof, s, "lattice_insert");
    k_insert<D><<<blocks_for(Npad), kThreads, 0, s>>>(Npad, d, ep.npi, rem, rnk, table, (uint32_t)(cap - 1), slot_of);
  }

  int* flag = dev_alloc<int>(T, s);
  int* vidx = dev_alloc<int>(T, s);
  {
    LaunchScope ls(cx.prof, s, "lattice_rank");
    k_mark_first<<<blocks_for(T), kThreads, 0, s>>>(T, table, slot_of, flag);
    size_t tmp_bytes = 0;
    DCRF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, flag, vidx, (int)T, s));
    void* tmp = dev_alloc<char>(tmp_bytes, s);
    DCRF_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, flag, vidx, (int)T, s));
    dev_free(tmp, s);
  }
  int last[2];
  DCRF_CUDA(cudaMemcpyAsync(&last[0], vidx + (T - 1), sizeof(int), cudaMemcpyDeviceToHost, s));
  DCRF_CUDA(cudaMemcpyAsync(&last[1], flag + (T - 1), sizeof(int), cudaMemcpyDeviceToHost, s));
  DCRF_CUDA(cudaStreamSynchronize(s));
  const int V = last[0] + last[1];

  L->N = N;
  L->d = d;
  L->V = V;
  L->Npad = Npad;
  L->E = E;
  L->alpha = 1.0f / (1.0f + powf(2.0f, (float)-d));  // permutohedral.cpp:571
  L->pix = dev_alloc<PixEntry>(E, s);
  L->keys = dev_alloc<int16_t>((size_t)V * d, s);
  L->csr = dev_alloc<CsrEntry>(E, s);
  L->csr_start = dev_alloc<int>((size_t)V + 1, s);
  if (ep.npi) L->vimg = dev_alloc<int>((size_t)V, s);
  L->nbr = dev_alloc<int2>((size_t)V * dp1, s);

  int* count = dev_alloc<int>((size_t)V + 2, s);
  DCRF_CUDA(cudaMemsetAsync(count, 0, ((size_t)V + 2) * sizeof(int), s));
  uint32_t* skey = dev_alloc<uint32_t>(E, s);
  uint32_t* sval = dev_alloc<uint32_t>(E, s);
  uint32_t* skey2 = dev_alloc<uint32_t>(E, s);
  uint32_t* sval2 = dev_alloc<uint32_t>(E, s);
  {
    LaunchScope ls(cx.prof, s, "lattice_assign");
    k_assign<D><<<blocks_for(T), kThreads, 0, s>>>(T, E, d, table, slot_of, vidx, bary, rem, rnk, L->pix, skey, sval,
                                                   count, L->keys, ep.npi, L->vimg);
    k_table_to_ids<<<blocks_for(T), kThreads, 0, s>>>(T, flag, vidx, slot_of, table);
  }
  {
    LaunchScope ls(cx.prof, s, "lattice_csr");
    size_t tmp_bytes = 0;
    DCRF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, count, L->csr_start, V + 1, s));
    void* tmp = dev_alloc<char>(tmp_bytes, s);
    DCRF_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, count, L->csr_start, V + 1, s));
    dev_free(tmp, s);
    if (E > 0) {
      int bits = 1;
      while ((1ull << bits) < (uint64_t)V) bits++;
      // stable LSD radix sort by vertex id: inside a vertex the contributions stay in ascending t
      // order = the order the reference's sequential splat adds them (permutohedral.cpp:545-553)
      tmp_bytes = 0;
      DCRF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, skey, skey2, sval, sval2, (int)E, 0, bits, s));
      tmp = dev_alloc<char>(tmp_bytes, s);
      DCRF_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, skey, skey2, sval, sval2, (int)E, 0, bits, s));
      dev_free(tmp, s);
      k_fill_csr<<<blocks_for(E), kThreads, 0, s>>>(E, dp1, sval2, bary, L->csr);
    }
  }
  L->order = dev_alloc<int>((size_t)V, s);
  if (V > 0) {
    LaunchScope ls(cx.prof, s, "lattice_order");
    uint32_t* okey = dev_alloc<uint32_t>(V, s);
    uint32_t* okey2 = dev_alloc<uint32_t>(V, s);
    uint32_t* oval = dev_alloc<uint32_t>(V, s);
    k_order_keys<<<blocks_for(V), kThreads, 0, s>>>(V, count, okey, oval);
    int bits = 9;
    while ((1ull << (bits - 8)) <= ((uint64_t)V >> 12)) bits++;
    size_t tmp_bytes = 0;
    DCRF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, okey, okey2, oval, (uint32_t*)L->order, V, 0, bits, s));
    void* tmp = dev_alloc<char>(tmp_bytes, s);
    DCRF_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, okey, okey2, oval, (uint32_t*)L->order, V, 0, bits, s));
    for (void* q : {(void*)tmp, (void*)okey, (void*)okey2, (void*)oval}) dev_free(q, s);
  }
  if (V > 0) {
    LaunchScope ls(cx.prof, s, "lattice_neighbours");
    dim3 grid(blocks_for(V), dp1);
    k_neighbours<D><<<grid, kThreads, 0, s>>>(V, d, L->keys, table, (uint32_t)(cap - 1), L->vimg, L->nbr);
  }
  DCRF_CUDA(cudaGetLastError());
  // no synchronisation here: the temporaries are freed in stream order and every caller synchronises once its term is
  // complete (add_term)

  L->rank = rnk;  // kept: Permutohedral::gradient reads rank_ (permutohedral.cpp:662)
  for (void* p : {(void*)bary, (void*)rem, (void*)table, (void*)slot_of, (void*)flag, (void*)vidx,
                  (void*)count, (void*)skey, (void*)sval, (void*)skey2, (void*)sval2})
    dev_free(p, s);
}

}  // namespace

// f_ of DenseKernel (pairwise.cpp:117) as a d x N row-major device array, whatever the features came from; with
// `params` (device) the transformed features DenseKernel::setParameters builds the lattice from (:139-151):
// ktype 1 (DIAG_KERNEL): p[k] * f[k][i]; ktype 2 (FULL_KERNEL): sum_j P(k, j) f[j][i], P column-major, the sum taken in
// ascending j from zero with every product and sum rounded (the order of Eigen's product kernel over the depth)
namespace {
__global__ void k_materialize_features(FeatureSource src, int d, int N, int ktype, const float* __restrict__ params,
                                       float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float f[kMaxFeatureDims];
  for (int k = 0; k < d; k++) f[k] = feature_of(src, k, i, N);
  for (int k = 0; k < d; k++) {
    float v = f[k];
    if (ktype == 1) {
      v = __fmul_rn(params[k], f[k]);
    } else if (ktype == 2) {
      v = 0.f;
      for (int j = 0; j < d; j++) v = __fadd_rn(v, __fmul_rn(params[k + j * d], f[j]));
    }
    out[(size_t)k * N + i] = v;
  }
}
}  // namespace
void materialize_features(Ctx cx, const FeatureSource& src, int d, int N, int ktype, const float* params, float* out_dxN) {
  LaunchScope ls(cx.prof, cx.stream, "materialize_features");
  k_materialize_features<<<(N + kThreads - 1) / kThreads, kThreads, 0, cx.stream>>>(src, d, N, ktype, params, out_dxN);
}

// contribution weights with a term's pre-normalisation folded in: w * norm[pix] (VAR_BAKED_NORM in api.cu)
namespace {
__global__ void k_bake_csr(size_t E, const CsrEntry* __restrict__ csr, const float* __restrict__ norm, CsrEntry* __restrict__ out) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  CsrEntry c = csr[e];
  c.w = __fmul_rn(c.w, __ldg(norm + c.pix));
  out[e] = c;
}
}  // namespace
void launch_bake_csr(Ctx cx, size_t E, const CsrEntry* csr, const float* norm, CsrEntry* out) {
  LaunchScope ls(cx.prof, cx.stream, "bake_csr");
  k_bake_csr<<<blocks_for(E), kThreads, 0, cx.stream>>>(E, csr, norm, out);
}

void Lattice::release(cudaStream_t s) {
  dev_free(pix, s);
  dev_free(csr_start, s);
  dev_free(csr, s);
  dev_free(nbr, s);
  dev_free(keys, s);
  dev_free(order, s);
  dev_free(vimg, s);
  dev_free(rank, s);
  rank = nullptr;
  order = nullptr;
  vimg = nullptr;
  pix = nullptr;
  csr_start = nullptr;
  csr = nullptr;
  nbr = nullptr;
  keys = nullptr;
}

void lattice_build(Ctx cx, const FeatureSource& src, int d, int N, Lattice* out, int npi) {
  DCRF_REQUIRE(d >= 1 && d <= kMaxFeatureDims,
               "feature dimension must be in [1, " + std::to_string(kMaxFeatureDims) + "], got " + std::to_string(d));
  DCRF_REQUIRE(N >= 1, "need at least one variable");
  ElevateParams ep;
  ep.N = N;
  ep.Npad = (N + 3) & ~3;  // the reference walks pixels in blocks of 4 and inserts the padding lanes too
  ep.d = d;
  ep.npi = npi;
  ep.src = src;
  DCRF_REQUIRE(npi == 0 || (npi % 4 == 0 && N % npi == 0), "a batch needs images of a multiple of four pixels");
  DCRF_REQUIRE((uint64_t)ep.Npad * (uint64_t)(d + 1) < (1ull << 31), "nvar*(d+1) must stay below 2^31");
  // permutohedral.cpp:179-182: float inv_std_dev, scale evaluated in double then narrowed
  const float inv_std_dev = (float)(std::sqrt(2.0 / 3.0) * (double)(d + 1));
  for (int i = 0; i < kMaxFeatureDims; i++)
    ep.scale[i] = (i < d) ? (float)(1.0 / std::sqrt((double)((i + 2) * (i + 1))) * (double)inv_std_dev) : 0.f;
  ep.inv_dp1 = 1.0f / (float)(d + 1);
  ep.dp1 = (float)(d + 1);
  switch (d) {
    case 1: build_impl<1>(cx, ep, out); break;
    case 2: build_impl<2>(cx, ep, out); break;
    case 3: build_impl<3>(cx, ep, out); break;
    case 4: build_impl<4>(cx, ep, out); break;
    case 5: build_impl<5>(cx, ep, out); break;
    case 6: build_impl<6>(cx, ep, out); break;
    case 7: build_impl<7>(cx, ep, out); break;
    default: build_impl<0>(cx, ep, out); break;
  }
}

}  // namespace dcrf
