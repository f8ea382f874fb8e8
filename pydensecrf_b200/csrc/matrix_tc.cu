// MatrixCompatibility for large label counts on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Replaces MatrixCompatibility::apply, out = W * Q (pydensecrf/densecrf/src/labelcompatibility.cpp:82-84 in
// /root/reference), i.e. per pixel out[l] = sum_k W[l][k] * msg[k]: a (N x L) * (L x L) product, the only GEMM-shaped
// operation on the mean-field path (373 GFLOP per kernel-iteration at 4K / 150 labels).
//
// Shape on the hardware: D[128 pixels x Np labels] (fp32, in TMEM) += A[128 x 8] * B[8 x Np] per tcgen05.mma,
// kind::tf32, cta_group::1.  A = message rows (K-major as stored), B = W rows (W is symmetric, so its rows are the
// K-major B operand).  Precision: TF32 alone (10-bit mantissa) would cost ~1e-3 relative, far outside the 1e-4
// Q tolerance, so every operand is split x = hi + lo with hi = tf32(x) and the product is taken as
// hi*hi + lo*hi + hi*lo (three MMAs, fp32 accumulation): ~2^-20 relative per product.  The exact-order CUDA-core
// kernel (filter.cu: k_matrix_compat_tiled) stays the default; this path is chosen for L >= 64 unless the caller
// asks for bit-faithful arithmetic.
//
// Up to two CTAs (512 threads each) per SM walk 128-pixel tiles.  K is consumed in slabs of 16 (two k-steps), double
// buffered: all threads load + split the slab into the canonical no-swizzle K-major core-matrix layout in shared
// memory, one thread issues the 6 MMAs of the slab and commits them to an mbarrier that frees the buffer.  The
// epilogue reads the accumulator with tcgen05.ld (a warp reads TMEM lanes 32*(warp%4).. = 32 pixels) into a
// shared-memory tile that aliases the operand stages, then all threads apply tmp1 -= out (or write out) with
// coalesced row accesses.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "common.h"
#include "softmax_math.cuh"

namespace dcrf {
namespace {

#ifndef TCU_WS_ROWS
#define TCU_WS_ROWS 2  // rows per epilogue warp in flight in the warp-specialised update kernel
#endif
constexpr int kTM = 128;  // pixels per tile = UMMA M
constexpr int kKS = 16;   // K elements per slab (two k-steps): small stages so that two CTAs fit on an SM
constexpr int kCH = kKS / 4;  // 16-byte K chunks per slab
constexpr int kTCThreads = 512;  // 16 warps fill the slabs; warps w, w+4, w+8, .. share TMEM lane quarter w%4 in the epilogue

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, K-major, no swizzle: 8-row x 16-byte core matrices; consecutive 8-row groups
// are `sbo` bytes apart, consecutive 16-byte K chunks `lbo` bytes apart (version 1 = Blackwell)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a), "l"(b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, 0x989680;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}\n"
        : "=r"(done)
        : "r"(mbar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  uint32_t h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  lo = x - hi;  // exact
}

struct TCArgs {
  int N, L, Lp, Np, Kp, mode, neg_unary;
  const float* msg;   // N x Lp
  const float* Bhi;   // Np x Kp: tf32(W), zero padded
  const float* Blo;   // Np x Kp: W - tf32(W)
  const float* U;
  float* out;
};

__global__ void __launch_bounds__(kTCThreads, 2) k_matrix_compat_tc(TCArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  // +32: consecutive K chunks land two 16-byte bank groups apart, so the 2 rows x 4 chunks a quarter-warp stores hit
  // eight different bank groups (with +16 rows and chunks collide pairwise: 230 M conflicts in profiles/r01j)
  const uint32_t lboA = kTM * 16 + 32, lboB = (uint32_t)a.Np * 16 + 32;
  const uint32_t szA = kCH * lboA, szB = kCH * lboB;
  const uint32_t stage_bytes = 2 * szA + 2 * szB;
  uint8_t* stage0 = smem;
  const uint32_t tile_bytes = (uint32_t)(kTM * (a.Lp + 4) * sizeof(float));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + max(2 * stage_bytes, tile_bytes));  // [0,1]: stage free, [2]: accumulator done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  if (tid == 0) {
    mbar_init(smem_u32(bars + 0), 1);
    mbar_init(smem_u32(bars + 1), 1);
    mbar_init(smem_u32(bars + 2), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {  // one warp allocates 256 TMEM columns (power of two >= Np) and publishes the base address
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  // instruction descriptor: D fp32, A/B tf32, both K-major, N = Np, M = 128
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.Np >> 3) << 17) | ((uint32_t)(kTM >> 4) << 24);
  const int nslabs = a.Kp / kKS;
  const int ntiles = (a.N + kTM - 1) / kTM;
  uint32_t it = 0;  // slabs issued so far by this CTA (stage = it & 1)
  uint32_t tl = 0;  // tiles finished by this CTA

  // Register prefetch: the global loads of slab n+1 are issued right after slab n has been stored to shared
  // memory, so their latency runs under the fence / barrier / MMA issue of slab n (and under the epilogue at a
  // tile boundary).  kTM * kCH == kTCThreads: one A chunk per thread; B needs up to kNB chunks per thread.
  static_assert(kTM * kCH == kTCThreads, "one A chunk per thread");
  constexpr int kNB = (256 * kCH + kTCThreads - 1) / kTCThreads;
  float4 pa, pbh[kNB], pbl[kNB];
  const int arow = tid / kCH, ac = tid % kCH;
  auto prefetch = [&](int tile_p0, int slab) {
    const int k = slab * kKS + ac * 4;
    pa = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tile_p0 + arow < a.N && k < a.Lp) pa = __ldg(reinterpret_cast<const float4*>(a.msg + (size_t)(tile_p0 + arow) * a.Lp + k));
#pragma unroll
    for (int u = 0; u < kNB; u++) {
      const int e = u * kTCThreads + tid;
      if (e < a.Np * kCH) {
        const size_t g = (size_t)(e / kCH) * a.Kp + slab * kKS + (e % kCH) * 4;
        pbh[u] = __ldg(reinterpret_cast<const float4*>(a.Bhi + g));
        pbl[u] = __ldg(reinterpret_cast<const float4*>(a.Blo + g));
      }
    }
  };
  if ((int)blockIdx.x < ntiles) prefetch(blockIdx.x * kTM, 0);

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, tl++) {
    const int p0 = tile * kTM;
    for (int s = 0; s < nslabs; s++, it++) {
      const uint32_t st = it & 1;
      uint8_t* sb = stage0 + st * stage_bytes;
      if (it >= 2) mbar_wait(smem_u32(bars + st), ((it >> 1) - 1) & 1);  // MMAs that read this buffer are done
      {  // registers -> shared memory (A split into hi/lo on the way)
        float4 h, l;
        split_tf32(pa.x, h.x, l.x);
        split_tf32(pa.y, h.y, l.y);
        split_tf32(pa.z, h.z, l.z);
        split_tf32(pa.w, h.w, l.w);
        const uint32_t o = (uint32_t)ac * lboA + (uint32_t)arow * 16;
        *reinterpret_cast<float4*>(sb + o) = h;
        *reinterpret_cast<float4*>(sb + szA + o) = l;
#pragma unroll
        for (int u = 0; u < kNB; u++) {
          const int e = u * kTCThreads + tid;
          if (e < a.Np * kCH) {
            const uint32_t ob = (uint32_t)(e % kCH) * lboB + (uint32_t)(e / kCH) * 16;
            *reinterpret_cast<float4*>(sb + 2 * szA + ob) = pbh[u];
            *reinterpret_cast<float4*>(sb + 2 * szA + szB + ob) = pbl[u];
          }
        }
      }
      if (s + 1 < nslabs)
        prefetch(p0, s + 1);
      else if (tile + (int)gridDim.x < ntiles)
        prefetch((tile + (int)gridDim.x) * kTM, 0);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the MMA's async proxy
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sA = smem_u32(sb), sB = smem_u32(sb + 2 * szA);
#pragma unroll
        for (int ks = 0; ks < kKS / 8; ks++) {  // one k-step = 8 tf32 = two 16-byte chunks
          const uint64_t ahi = umma_desc(sA + ks * 2 * lboA, lboA, 128), alo = umma_desc(sA + szA + ks * 2 * lboA, lboA, 128);
          const uint64_t bhi = umma_desc(sB + ks * 2 * lboB, lboB, 128), blo = umma_desc(sB + szB + ks * 2 * lboB, lboB, 128);
          umma_tf32(tmem, ahi, bhi, idesc, (s | ks) ? 1u : 0u);
          umma_tf32(tmem, alo, bhi, idesc, 1u);
          umma_tf32(tmem, ahi, blo, idesc, 1u);
        }
        umma_commit(smem_u32(bars + st));                      // frees this stage when the MMAs above have read it
        if (s == nslabs - 1) umma_commit(smem_u32(bars + 2));  // accumulator complete
      }
    }
    // ---- epilogue: accumulator -> shared-memory tile -> coalesced read-modify-write of the output rows ----------
    mbar_wait(smem_u32(bars + 2), tl & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // all MMAs of this tile are complete, so both operand stages are free: the output tile aliases them
    float* T = reinterpret_cast<float*>(stage0);
    const int TLD = a.Lp + 4;  // row pitch in floats: 16-byte aligned, 28 mod 32 -> conflict-free float4 stores
    {
      const int lane_q = warp & 3;                  // TMEM lane quarter this warp may read
      const int trow = lane_q * 32 + (tid & 31);     // tile row = TMEM lane
      const uint32_t taddr = tmem + ((uint32_t)(lane_q * 32) << 16);
      const int nchunk = (a.Lp + 15) / 16, per = (nchunk + kTCThreads / 128 - 1) / (kTCThreads / 128);
      const int cbeg = (warp >> 2) * per * 16, cend = min(a.Lp, ((warp >> 2) + 1) * per * 16);
      for (int c0 = cbeg; c0 < cend; c0 += 16) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          if (c0 + j < a.Lp)
            *reinterpret_cast<float4*>(T + (size_t)trow * TLD + c0 + j) =
                make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();  // tile complete in shared memory; TMEM may be overwritten by the next tile's MMAs
    {
      const int cpr = a.Lp >> 2;  // float4 chunks per row
      for (int e = tid; e < kTM * cpr; e += kTCThreads) {
        const int row = e / cpr, c4 = e - row * cpr;
        const int i = p0 + row;
        if (i >= a.N) continue;
        const int col = c4 * 4;
        const float4 t4 = *reinterpret_cast<const float4*>(T + (size_t)row * TLD + col);
        float o[4] = {t4.x, t4.y, t4.z, t4.w};
        const size_t off = (size_t)i * a.Lp + col;
        if (a.mode != SLICE_WRITE) {
          float b[4];
          if (a.mode == SLICE_ACC_FIRST) {
            const float z = a.neg_unary ? -0.f : 0.f;
            if (a.U) {
              const float4 u = __ldg(reinterpret_cast<const float4*>(a.U + off));
              b[0] = z - u.x, b[1] = z - u.y, b[2] = z - u.z, b[3] = z - u.w;
            } else {
              b[0] = b[1] = b[2] = b[3] = z;
            }
          } else {
            const float4 v = *reinterpret_cast<const float4*>(a.out + off);
            b[0] = v.x, b[1] = v.y, b[2] = v.z, b[3] = v.w;
          }
#pragma unroll
          for (int q = 0; q < 4; q++) o[q] = b[q] - o[q];
        }
#pragma unroll
        for (int q = 0; q < 4; q++)
          if (col + q >= a.L) o[q] = 0.f;
        *reinterpret_cast<float4*>(a.out + off) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
    __syncthreads();  // the output tile has been consumed: the next tile may refill the operand stages
  }

  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

// =================================================================================================
// The update for models whose terms all carry a MatrixCompatibility, in one kernel after the slices:
//   Q = softmax( -U - sum_k W_k * msg_k ),   msg_k = norm_k . slice_k (written by k_slice)
// Same tile / slab / TMEM structure as k_matrix_compat_tc, but
//   * all terms accumulate into the same TMEM accumulator (the K dimension is the concatenation of the terms), so the
//     N x L accumulator array in HBM and its read-modify-write per term disappear;
//   * the epilogue turns the accumulator tile into x = (-U) - acc in shared memory and runs the bit-faithful column
//     softmax on it (a warp per pixel, the code of k_softmax_rows), writing Q: no init / softmax passes.
// Per iteration at 4K / 150 labels / 2 terms: 1.5 (init) + 2 x 4.8 (GEMM) + 4.9 (softmax) ms -> one launch.
// (A variant that also gathered the vertex rows itself, so that the messages never reached HBM, was measured at
// 21.8 ms per launch: one slab of prefetch distance cannot cover gather latency under block-wide barriers.)
// =================================================================================================
__global__ void __launch_bounds__(kTCThreads, 2) k_update_matrix_tc(TCUArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t lboA = kTM * 16 + 32, lboB = (uint32_t)a.Np * 16 + 32;
  const uint32_t szA = kCH * lboA, szB = kCH * lboB;
  const uint32_t stage_bytes = 2 * szA + 2 * szB;
  uint8_t* stage0 = smem;
  const uint32_t tile_bytes = (uint32_t)(kTM * (a.Lp + 4) * sizeof(float));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + max(2 * stage_bytes, tile_bytes));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  if (tid == 0) {
    mbar_init(smem_u32(bars + 0), 1);
    mbar_init(smem_u32(bars + 1), 1);
    mbar_init(smem_u32(bars + 2), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.Np >> 3) << 17) | ((uint32_t)(kTM >> 4) << 24);
  const int nslabs = a.Kp / kKS;          // per term
  const int gslabs = nslabs * a.nterms;   // per tile
  const int ntiles = (a.N + kTM - 1) / kTM;
  uint32_t it = 0, tl = 0;
#ifdef TCU_TIMING  // -DTCU_TIMING: block 0 prints the clock64 cycles per tile spent in each phase (profiles/build_variant.py)
  long long t_main = 0, t_stage = 0, t_soft = 0, t_mark = clock64();
#endif

  static_assert(kTM * kCH == kTCThreads, "one A chunk per thread");
  constexpr int kNB = (256 * kCH + kTCThreads - 1) / kTCThreads;
  const int arow = tid / kCH, ac = tid % kCH;
  float4 pa, pbh[kNB], pbl[kNB];
  auto prefetch = [&](int tile_p0, int g) {
    const int term = g / nslabs, slab = g - term * nslabs;
    const TCUTerm& t = a.t[term];
    const int k = slab * kKS + ac * 4;
    pa = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tile_p0 + arow < a.N && k < a.Lp) pa = __ldg(reinterpret_cast<const float4*>(t.msg + (size_t)(tile_p0 + arow) * a.Lp + k));
#pragma unroll
    for (int u = 0; u < kNB; u++) {
      const int e = u * kTCThreads + tid;
      if (e < a.Np * kCH) {
        const size_t gi = (size_t)(e / kCH) * a.Kp + slab * kKS + (e % kCH) * 4;
        pbh[u] = __ldg(reinterpret_cast<const float4*>(t.Bhi + gi));
        pbl[u] = __ldg(reinterpret_cast<const float4*>(t.Blo + gi));
      }
    }
  };
  if ((int)blockIdx.x < ntiles) prefetch(blockIdx.x * kTM, 0);

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, tl++) {
    const int p0 = tile * kTM;
    for (int g = 0; g < gslabs; g++, it++) {
      const uint32_t st = it & 1;
      uint8_t* sb = stage0 + st * stage_bytes;
      if (it >= 2) mbar_wait(smem_u32(bars + st), ((it >> 1) - 1) & 1);
      {
        float4 h, l;
        split_tf32(pa.x, h.x, l.x);
        split_tf32(pa.y, h.y, l.y);
        split_tf32(pa.z, h.z, l.z);
        split_tf32(pa.w, h.w, l.w);
        const uint32_t o = (uint32_t)ac * lboA + (uint32_t)arow * 16;
        *reinterpret_cast<float4*>(sb + o) = h;
        *reinterpret_cast<float4*>(sb + szA + o) = l;
#pragma unroll
        for (int u = 0; u < kNB; u++) {
          const int e = u * kTCThreads + tid;
          if (e < a.Np * kCH) {
            const uint32_t ob = (uint32_t)(e % kCH) * lboB + (uint32_t)(e / kCH) * 16;
            *reinterpret_cast<float4*>(sb + 2 * szA + ob) = pbh[u];
            *reinterpret_cast<float4*>(sb + 2 * szA + szB + ob) = pbl[u];
          }
        }
      }
      if (g + 1 < gslabs)
        prefetch(p0, g + 1);
      else if (tile + (int)gridDim.x < ntiles)
        prefetch((tile + (int)gridDim.x) * kTM, 0);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sA = smem_u32(sb), sB = smem_u32(sb + 2 * szA);
#pragma unroll
        for (int ks = 0; ks < kKS / 8; ks++) {
          const uint64_t ahi = umma_desc(sA + ks * 2 * lboA, lboA, 128), alo = umma_desc(sA + szA + ks * 2 * lboA, lboA, 128);
          const uint64_t bhi = umma_desc(sB + ks * 2 * lboB, lboB, 128), blo = umma_desc(sB + szB + ks * 2 * lboB, lboB, 128);
          umma_tf32(tmem, ahi, bhi, idesc, (g | ks) ? 1u : 0u);
          umma_tf32(tmem, alo, bhi, idesc, 1u);
          umma_tf32(tmem, ahi, blo, idesc, 1u);
        }
        umma_commit(smem_u32(bars + st));
        if (g == gslabs - 1) umma_commit(smem_u32(bars + 2));
      }
    }
#ifdef TCU_TIMING
    { const long long n = clock64(); t_main += n - t_mark; t_mark = n; }
#endif
    // ---- epilogue: accumulator -> shared-memory tile -> x = (-U) - acc -> column softmax -> Q ------------------
    mbar_wait(smem_u32(bars + 2), tl & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* T = reinterpret_cast<float*>(stage0);
    const int TLD = a.Lp + 4;
    {
      const int lane_q = warp & 3;
      const int trow = lane_q * 32 + lane;
      const uint32_t taddr = tmem + ((uint32_t)(lane_q * 32) << 16);
      const int nchunk = (a.Lp + 15) / 16, per = (nchunk + kTCThreads / 128 - 1) / (kTCThreads / 128);
      const int cbeg = (warp >> 2) * per * 16, cend = min(a.Lp, ((warp >> 2) + 1) * per * 16);
      for (int c0 = cbeg; c0 < cend; c0 += 16) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          if (c0 + j < a.Lp)
            *reinterpret_cast<float4*>(T + (size_t)trow * TLD + c0 + j) =
                make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();  // tile complete in shared memory; TMEM may be overwritten by the next tile's MMAs
#ifdef TCU_TIMING
    { const long long n = clock64(); t_stage += n - t_mark; t_mark = n; }
#endif
    {
      const unsigned full = 0xffffffffu;
      const int L = a.L, nfull = L / 4;
      const float z = a.neg_unary ? -0.f : 0.f;
      for (int row = warp; row < kTM; row += kTCThreads / 32) {
        const int i = p0 + row;
        if (i >= a.N) break;  // warp-uniform
        float* b = T + (size_t)row * TLD;
        const float* u = a.U ? a.U + (size_t)i * a.Lp : nullptr;
        float mx = -INFINITY;
        for (int l = lane; l < L; l += 32) {
          const float base = u ? __fsub_rn(z, __ldg(u + l)) : z;
          const float v = __fsub_rn(base, b[l]);  // tmp1 = -unary; tmp1 -= sum of the terms' outputs
          b[l] = v;
          mx = fmaxf(mx, v);
        }
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(full, mx, o));
        __syncwarp();
        for (int l = lane; l < L; l += 32) {
          const float xs = __fsub_rn(b[l], mx);
          b[l] = (l < nfull * 4) ? eigen_pexp_nonpos(xs) : tail_exp_nonpos(xs);
        }
        __syncwarp();
        // Eigen's redux order (see k_softmax_rows in filter.cu)
        float s = 0.f;
        if (nfull == 0) {
          if (lane == 0) {
            s = b[0];
            for (int l = 1; l < L; l++) s = __fadd_rn(s, b[l]);
          }
        } else {
          const int par = (lane >> 2) & 1, e = lane & 3;
          const int npair = nfull / 2;
          float ch = 0.f;
          if (lane < 8 && (par == 0 || nfull > 1)) {
            ch = b[4 * par + e];
            for (int m2 = 1; m2 < npair; m2++) ch = __fadd_rn(ch, b[8 * m2 + 4 * par + e]);
          }
          const float other = __shfl_down_sync(full, ch, 4);
          if (nfull > 1) {
            ch = __fadd_rn(ch, other);
            if ((nfull & 1) && lane < 4) ch = __fadd_rn(ch, b[4 * (nfull - 1) + e]);
          }
          const float c0 = __shfl_sync(full, ch, 0), c1 = __shfl_sync(full, ch, 1), c2 = __shfl_sync(full, ch, 2),
                      c3 = __shfl_sync(full, ch, 3);
          if (lane == 0) {
            s = __fadd_rn(__fadd_rn(c0, c2), __fadd_rn(c1, c3));
            for (int l = nfull * 4; l < L; l++) s = __fadd_rn(s, b[l]);
          }
        }
        s = __shfl_sync(full, s, 0);
        float* qo = a.Q + (size_t)i * a.Lp;
        for (int l = lane; l < a.Lp; l += 32) qo[l] = (l < L) ? __fdiv_rn(b[l], s) : 0.f;
      }
    }
    __syncthreads();  // the tile has been consumed: the next tile may refill the operand stages
#ifdef TCU_TIMING
    { const long long n = clock64(); t_soft += n - t_mark; t_mark = n; }
#endif
  }

  __syncthreads();
#ifdef TCU_TIMING
  if (tid == 0 && blockIdx.x == 0) printf("[tcu] tiles %u: mainloop %lld, tmem->smem %lld, softmax %lld cycles per tile\n", tl, t_main / max(tl, 1u), t_stage / max(tl, 1u), t_soft / max(tl, 1u));
#endif
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

// =================================================================================================
// Warp-specialised form of k_update_matrix_tc (the default when its shared memory fits, i.e. up to ~200 labels): one
// CTA of 1024 threads per SM, 16 producer warps (operand slabs + MMA issue, the main loop above unchanged) and 16
// epilogue warps; the accumulator is double-buffered in TMEM (2 x 256 columns) and the softmax tile has its own shared
// memory, so the epilogue of tile t (TMEM -> x = (-U) - acc -> column softmax -> Q) runs while the producers are already
// feeding tile t+1.  The epilogue requests its unary values before it waits for the accumulator and keeps two rows per
// warp in flight.  4K / 150 labels / 2 terms: 11.9 ms per launch against 13.7 for the two-CTA form (clock64 per tile:
// the epilogue never waits for the accumulator; it is the longer half at ~45 k cycles).
// =================================================================================================
constexpr int kProdThreads = kTCThreads;  // 512: warps 0..15
#ifndef TCU_WS_EPI
#define TCU_WS_EPI 16
#endif
constexpr int kEpiWarps = TCU_WS_EPI;     // warps 16..
constexpr int kWSThreads = kProdThreads + 32 * kEpiWarps;

__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}

template <int KU>
__global__ void __launch_bounds__(kWSThreads, 1) k_update_matrix_tc_ws(TCUArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t lboA = kTM * 16 + 32, lboB = (uint32_t)a.Np * 16 + 32;
  const uint32_t szA = kCH * lboA, szB = kCH * lboB;
  const uint32_t stage_bytes = 2 * szA + 2 * szB;
  uint8_t* stage0 = smem;
  const int TLD = a.Lp + 4;
  float* T = reinterpret_cast<float*>(smem + 2 * stage_bytes);
  const uint32_t tile_bytes = (uint32_t)(kTM * TLD * sizeof(float));
  // bars[0,1]: operand stage free; bars[2,3]: accumulator b complete; bars[4,5]: accumulator b drained by the epilogue
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * stage_bytes + tile_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

  if (tid == 0) {
    for (int i = 0; i < 4; i++) mbar_init(smem_u32(bars + i), 1);
    mbar_init(smem_u32(bars + 4), kEpiWarps);
    mbar_init(smem_u32(bars + 5), kEpiWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const int ntiles = (a.N + kTM - 1) / kTM;

  if (tid < kProdThreads) {
    // ---------------- producers + MMA issue -----------------------------------------------------------------------
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.Np >> 3) << 17) | ((uint32_t)(kTM >> 4) << 24);
    const int nslabs = a.Kp / kKS;          // per term
    const int gslabs = nslabs * a.nterms;   // per tile
    uint32_t it = 0, tl = 0;
    static_assert(kTM * kCH == kProdThreads, "one A chunk per producer thread");
    constexpr int kNB = (256 * kCH + kProdThreads - 1) / kProdThreads;
    const int arow = tid / kCH, ac = tid % kCH;
    float4 pa, pbh[kNB], pbl[kNB];
    auto prefetch = [&](int tile_p0, int g) {
      const int term = g / nslabs, slab = g - term * nslabs;
      const TCUTerm& t = a.t[term];
      const int k = slab * kKS + ac * 4;
      pa = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tile_p0 + arow < a.N && k < a.Lp) pa = __ldg(reinterpret_cast<const float4*>(t.msg + (size_t)(tile_p0 + arow) * a.Lp + k));
#pragma unroll
      for (int u = 0; u < kNB; u++) {
        const int e = u * kProdThreads + tid;
        if (e < a.Np * kCH) {
          const size_t gi = (size_t)(e / kCH) * a.Kp + slab * kKS + (e % kCH) * 4;
          pbh[u] = __ldg(reinterpret_cast<const float4*>(t.Bhi + gi));
          pbl[u] = __ldg(reinterpret_cast<const float4*>(t.Blo + gi));
        }
      }
    };
    if ((int)blockIdx.x < ntiles) prefetch(blockIdx.x * kTM, 0);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, tl++) {
      const int p0 = tile * kTM;
      const uint32_t acc = tmem + (tl & 1) * 256;
      for (int g = 0; g < gslabs; g++, it++) {
        const uint32_t st = it & 1;
        uint8_t* sb = stage0 + st * stage_bytes;
        if (it >= 2) mbar_wait(smem_u32(bars + st), ((it >> 1) - 1) & 1);
        {
          float4 h, l;
          split_tf32(pa.x, h.x, l.x);
          split_tf32(pa.y, h.y, l.y);
          split_tf32(pa.z, h.z, l.z);
          split_tf32(pa.w, h.w, l.w);
          const uint32_t o = (uint32_t)ac * lboA + (uint32_t)arow * 16;
          *reinterpret_cast<float4*>(sb + o) = h;
          *reinterpret_cast<float4*>(sb + szA + o) = l;
#pragma unroll
          for (int u = 0; u < kNB; u++) {
            const int e = u * kProdThreads + tid;
            if (e < a.Np * kCH) {
              const uint32_t ob = (uint32_t)(e % kCH) * lboB + (uint32_t)(e / kCH) * 16;
              *reinterpret_cast<float4*>(sb + 2 * szA + ob) = pbh[u];
              *reinterpret_cast<float4*>(sb + 2 * szA + szB + ob) = pbl[u];
            }
          }
        }
        if (g + 1 < gslabs)
          prefetch(p0, g + 1);
        else if (tile + (int)gridDim.x < ntiles)
          prefetch((tile + (int)gridDim.x) * kTM, 0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, %0;" ::"n"(kProdThreads) : "memory");  // producers only
        if (tid == 0) {
          // the accumulator buffer must have been drained by the epilogue of the tile that used it two tiles ago
          if (g == 0 && tl >= 2) mbar_wait(smem_u32(bars + 4 + (tl & 1)), ((tl >> 1) - 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sA = smem_u32(sb), sB = smem_u32(sb + 2 * szA);
#pragma unroll
          for (int ks = 0; ks < kKS / 8; ks++) {
            const uint64_t ahi = umma_desc(sA + ks * 2 * lboA, lboA, 128), alo = umma_desc(sA + szA + ks * 2 * lboA, lboA, 128);
            const uint64_t bhi = umma_desc(sB + ks * 2 * lboB, lboB, 128), blo = umma_desc(sB + szB + ks * 2 * lboB, lboB, 128);
            umma_tf32(acc, ahi, bhi, idesc, (g | ks) ? 1u : 0u);
            umma_tf32(acc, alo, bhi, idesc, 1u);
            umma_tf32(acc, ahi, blo, idesc, 1u);
          }
          umma_commit(smem_u32(bars + st));
          if (g == gslabs - 1) umma_commit(smem_u32(bars + 2 + (tl & 1)));
        }
      }
    }
  } else {
    // ---------------- epilogue warps ------------------------------------------------------------------------------
    const int ew = warp - kProdThreads / 32;  // 0..7
    const unsigned full = 0xffffffffu;
    const int L = a.L, nfull = L / 4, ntail = L & 3;
    uint32_t tl = 0;
#ifdef TCU_TIMING
    long long t_wait = 0, t_stage = 0, t_soft = 0, t_mark = clock64();
#endif
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, tl++) {
      const int p0 = tile * kTM;
      const uint32_t acc = tmem + (tl & 1) * 256;
      {
        // x = (-U) - accumulator -> T; the unary chunk of the first columns is requested before the accumulator wait
        const int lane_q = ew & 3;
        const int trow = lane_q * 32 + lane;
        const uint32_t taddr = acc + ((uint32_t)(lane_q * 32) << 16);
        const int nchunk = (a.Lp + 15) / 16, per = (nchunk + kEpiWarps / 4 - 1) / (kEpiWarps / 4);
        const int cbeg = (ew >> 2) * per * 16, cend = min(a.Lp, ((ew >> 2) + 1) * per * 16);
        const float z = a.neg_unary ? -0.f : 0.f;
        const bool have_u = a.U && p0 + trow < a.N;
        const float* urow = a.U + (size_t)(have_u ? p0 + trow : 0) * a.Lp;
        float4 un[4], uc[4];
        auto load_u = [&](int c0) {
#pragma unroll
          for (int j = 0; j < 4; j++)
            un[j] = (have_u && c0 + 4 * j < a.Lp) ? __ldg(reinterpret_cast<const float4*>(urow + c0 + 4 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        if (cbeg < cend) load_u(cbeg);
        mbar_wait(smem_u32(bars + 2 + (tl & 1)), (tl >> 1) & 1);
#ifdef TCU_TIMING
        { const long long n = clock64(); t_wait += n - t_mark; t_mark = n; }
#endif
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int c0 = cbeg; c0 < cend; c0 += 16) {
#pragma unroll
          for (int j = 0; j < 4; j++) uc[j] = un[j];
          uint32_t r[16];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
              : "r"(taddr + (uint32_t)c0));
          if (c0 + 16 < cend) load_u(c0 + 16);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 4; j++) {
            if (c0 + 4 * j < a.Lp) {
              float4 v;  // tmp1 = -unary (or 0 - unary); tmp1 -= sum of the terms' outputs
              v.x = __fsub_rn(__fsub_rn(z, uc[j].x), __uint_as_float(r[4 * j + 0]));
              v.y = __fsub_rn(__fsub_rn(z, uc[j].y), __uint_as_float(r[4 * j + 1]));
              v.z = __fsub_rn(__fsub_rn(z, uc[j].z), __uint_as_float(r[4 * j + 2]));
              v.w = __fsub_rn(__fsub_rn(z, uc[j].w), __uint_as_float(r[4 * j + 3]));
              *reinterpret_cast<float4*>(T + (size_t)trow * TLD + c0 + 4 * j) = v;
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("bar.sync 2, %0;" ::"n"(32 * kEpiWarps) : "memory");  // tile complete in T, TMEM buffer drained
      if (lane == 0) mbar_arrive(smem_u32(bars + 4 + (tl & 1)));
#ifdef TCU_TIMING
      { const long long n = clock64(); t_stage += n - t_mark; t_mark = n; }
#endif
      {
        // column softmax, bit-faithful, RG rows per warp in flight (see k_update_matrix_tc)
        constexpr int RG = TCU_WS_ROWS;
        for (int kb = 0; kb < kTM / kEpiWarps; kb += RG) {
          const int row0 = ew + kb * kEpiWarps;
          if (p0 + row0 >= a.N) break;  // warp-uniform: rows ascend with g
          float* bg[RG];
          bool valid[RG];
          float x[RG][KU], mx[RG];
#pragma unroll
          for (int g = 0; g < RG; g++) {
            const int row = row0 + g * kEpiWarps;
            valid[g] = p0 + row < a.N;
            bg[g] = T + (size_t)row * TLD;
            mx[g] = -INFINITY;
#pragma unroll
            for (int j = 0; j < KU; j++) {
              const int l = lane + 32 * j;
              x[g][j] = (l < L) ? bg[g][l] : -INFINITY;
              mx[g] = fmaxf(mx[g], x[g][j]);
            }
          }
#pragma unroll
          for (int o = 16; o; o >>= 1) {
#pragma unroll
            for (int g = 0; g < RG; g++) mx[g] = fmaxf(mx[g], __shfl_xor_sync(full, mx[g], o));
          }
          bool small[RG];
#pragma unroll
          for (int g = 0; g < RG; g++) {
            small[g] = false;
#pragma unroll
            for (int j = 0; j < KU; j++) {
              const int l = lane + 32 * j;
              const float xs = __fsub_rn(x[g][j], mx[g]);
              if (l < L) small[g] |= xs < -44.f;
              x[g][j] = (l < nfull * 4) ? eigen_pexp_nonpos(xs) : 0.f;
              if (l < nfull * 4) bg[g][l] = x[g][j];
            }
          }
          if (ntail) {  // lane t takes tail element t % ntail of row t / ntail
            const int tg = lane / ntail, tk = lane - tg * ntail;
            if (tg < RG) {
              const int l = nfull * 4 + tk;
              float m = mx[0];
#pragma unroll
              for (int g = 1; g < RG; g++)
                if (tg == g) m = mx[g];
              float* bt = T + (size_t)(row0 + tg * kEpiWarps) * TLD;
              bt[l] = tail_exp_nonpos(__fsub_rn(bt[l], m));
            }
          }
          __syncwarp();
          // Eigen's redux order, row g in the lanes [32/RG * g, 32/RG * (g+1)): the first four lanes of the group run
          // accumulator p0 (packets 0, 2, ..), the next four accumulator p1 (packets 1, 3, ..), element e = lane & 3
          constexpr int GL = 32 / RG;  // lanes per row group (8 or 16)
          float s = 0.f;
          {
            const int gsel = lane / GL, gl = lane - gsel * GL;
            const float* b = T + (size_t)(row0 + gsel * kEpiWarps) * TLD;
            if (nfull == 0) {
              s = b[0];
              for (int l = 1; l < L; l++) s = __fadd_rn(s, b[l]);
            } else {
              const int par = (gl >> 2) & 1, e = gl & 3;
              const int npair = nfull / 2;
              float ch = 0.f;
              if (gl < 8 && (par == 0 || nfull > 1)) {
                const float* bp = b + 4 * par + e;
                ch = bp[0];
                int m2 = 1;
                for (; m2 + 4 <= npair; m2 += 4) {
                  const float v0 = bp[8 * m2], v1 = bp[8 * m2 + 8], v2 = bp[8 * m2 + 16], v3 = bp[8 * m2 + 24];
                  ch = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(ch, v0), v1), v2), v3);
                }
                for (; m2 < npair; m2++) ch = __fadd_rn(ch, bp[8 * m2]);
              }
              const float other = __shfl_down_sync(full, ch, 4);
              if (nfull > 1) {
                ch = __fadd_rn(ch, other);
                if ((nfull & 1) && gl < 4) ch = __fadd_rn(ch, b[4 * (nfull - 1) + e]);
              }
              const int g0 = gsel * GL;
              const float c0 = __shfl_sync(full, ch, g0), c1 = __shfl_sync(full, ch, g0 + 1), c2 = __shfl_sync(full, ch, g0 + 2),
                          c3 = __shfl_sync(full, ch, g0 + 3);
              s = __fadd_rn(__fadd_rn(c0, c2), __fadd_rn(c1, c3));
              for (int l = nfull * 4; l < L; l++) s = __fadd_rn(s, b[l]);
            }
          }
          // out.col = b / sum: one refined reciprocal per row, IEEE division outside its proven range
#pragma unroll
          for (int g = 0; g < RG; g++) {
            const float sg = __shfl_sync(full, s, GL * g);
            const bool sm = __any_sync(full, small[g]);
            if (!valid[g]) continue;  // warp-uniform
            const float rg = refined_rcp(sg);
            float* qo = a.Q + (size_t)(p0 + row0 + g * kEpiWarps) * a.Lp;
#pragma unroll
            for (int j = 0; j < KU; j++) {
              const int l = lane + 32 * j;
              if (l < a.Lp) {
                float e = x[g][j];
                if (l >= nfull * 4 && l < L) e = bg[g][l];  // tail elements were evaluated by other lanes
                qo[l] = (l < L) ? (sm ? __fdiv_rn(e, sg) : div_shared(e, sg, rg)) : 0.f;
              }
            }
          }
          __syncwarp();
        }
      }
      asm volatile("bar.sync 2, %0;" ::"n"(32 * kEpiWarps) : "memory");  // T may be overwritten by the next tile
#ifdef TCU_TIMING
      { const long long n = clock64(); t_soft += n - t_mark; t_mark = n; }
#endif
    }
#ifdef TCU_TIMING
    if (ew == 0 && lane == 0 && blockIdx.x == 0) printf("[tcu-ws] epilogue, %u tiles: wait for accumulator %lld, tmem->smem %lld, softmax %lld cycles per tile\n", tl, t_wait / max(tl, 1u), t_stage / max(tl, 1u), t_soft / max(tl, 1u));
#endif
  }

  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// W (L x L, symmetrised) -> tf32 hi / lo parts in zero-padded Np x Kp arrays
__global__ void k_split_w(int L, int Np, int Kp, const float* __restrict__ W, float* __restrict__ hi, float* __restrict__ lo) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= Np * Kp) return;
  const int n = e / Kp, k = e - n * Kp;
  float h = 0.f, l = 0.f;
  if (n < L && k < L) split_tf32(W[(size_t)n * L + k], h, l);
  hi[e] = h;
  lo[e] = l;
}

}  // namespace

bool matrix_tc_supported(RowLayout rl) { return rl.vec == 4 && rl.L >= 16 && rl.Lp <= 256; }

void matrix_tc_dims(RowLayout rl, int* Np, int* Kp) {
  *Np = (rl.Lp + 15) & ~15;             // UMMA N: multiple of 16 for M = 128
  *Kp = (rl.Lp + 31) & ~31;               // whole slabs of 32 (the TMA kernel's; two of this file's 16-wide ones)
}

void launch_split_w(Ctx cx, RowLayout rl, const float* W, float* hi, float* lo) {
  LaunchScope ls(cx.prof, cx.stream, "split_w");
  int Np, Kp;
  matrix_tc_dims(rl, &Np, &Kp);
  k_split_w<<<(Np * Kp + 255) / 256, 256, 0, cx.stream>>>(rl.L, Np, Kp, W, hi, lo);
}

void launch_matrix_compat_tc(Ctx cx, RowLayout rl, int N, const float* Bhi, const float* Blo, const float* msg, int mode,
                             const float* U, bool neg_unary, float* out) {
  LaunchScope ls(cx.prof, cx.stream, "matrix_compat_tc");
  TCArgs a;
  a.N = N, a.L = rl.L, a.Lp = rl.Lp, a.mode = mode, a.neg_unary = neg_unary ? 1 : 0;
  matrix_tc_dims(rl, &a.Np, &a.Kp);
  a.msg = msg, a.Bhi = Bhi, a.Blo = Blo, a.U = U, a.out = out;
  const size_t stage = 2 * (size_t)(kCH * (kTM * 16 + 32)) + 2 * (size_t)(kCH * (a.Np * 16 + 32));
  // the epilogue's output tile (128 x (Lp+4) floats) aliases the operand stages; the barriers sit behind both
  const size_t body = std::max(2 * stage, (size_t)kTM * (a.Lp + 4) * sizeof(float));
  const size_t smem = body + 64;
  DCRF_REQUIRE(smem <= 227 * 1024, "too many labels for the tensor-core matrix compatibility kernel");
  DCRF_CUDA(cudaFuncSetAttribute(k_matrix_compat_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int ntiles = (N + kTM - 1) / kTM;
  const int per_sm = (2 * smem + 4096 <= 227 * 1024) ? 2 : 1;  // two CTAs per SM when shared memory (and 2 x 256 TMEM columns) allow
  k_matrix_compat_tc<<<std::min(ntiles, sms * per_sm), kTCThreads, smem, cx.stream>>>(a);
}

bool update_matrix_tc_supported(RowLayout rl, int nterms) {
  return matrix_tc_supported(rl) && nterms >= 1 && nterms <= kMaxTCUTerms;
}

void launch_update_matrix_tc(Ctx cx, RowLayout rl, TCUArgs a) {
  if (update_matrix_tc_tma_supported(rl, a.nterms)) {  // TMA-fed, thread-per-row epilogue (matrix_tc_tma.cu)
    launch_update_matrix_tc_tma(cx, rl, a);
    return;
  }
  LaunchScope ls(cx.prof, cx.stream, "update_matrix_tc");
  a.L = rl.L, a.Lp = rl.Lp;
  matrix_tc_dims(rl, &a.Np, &a.Kp);
  const size_t stage = 2 * (size_t)(kCH * (kTM * 16 + 32)) + 2 * (size_t)(kCH * (a.Np * 16 + 32));
  const size_t body = std::max(2 * stage, (size_t)kTM * (a.Lp + 4) * sizeof(float));
  const size_t smem = body + 64;
  DCRF_REQUIRE(smem <= 227 * 1024, "too many labels for the tensor-core update kernel");
  DCRF_CUDA(cudaFuncSetAttribute(k_update_matrix_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int ntiles = (a.N + kTM - 1) / kTM;
  const int per_sm = (2 * smem + 4096 <= 227 * 1024) ? 2 : 1;
  static const bool ws = [] {  // DCRF_B200_TCU_WS=0: the two-CTA form for every label count (A/B timing)
    const char* e = std::getenv("DCRF_B200_TCU_WS");
    return !e || std::atoi(e) != 0;
  }();
  const size_t smem_ws = 2 * stage + (size_t)kTM * (a.Lp + 4) * sizeof(float) + 128;
  if (ws && smem_ws <= 227 * 1024) {
    const int grid_ws = std::min(ntiles, sms);
    auto go = [&](auto kern) {
      DCRF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ws));
      kern<<<grid_ws, kWSThreads, smem_ws, cx.stream>>>(a);
    };
    const int ku = (rl.L + 31) / 32;
    if (ku <= 2)
      go(k_update_matrix_tc_ws<2>);
    else if (ku <= 3)
      go(k_update_matrix_tc_ws<3>);
    else if (ku <= 4)
      go(k_update_matrix_tc_ws<4>);
    else if (ku <= 5)
      go(k_update_matrix_tc_ws<5>);  // 129..160 labels (ADE20K's 150)
    else if (ku <= 6)
      go(k_update_matrix_tc_ws<6>);
    else
      go(k_update_matrix_tc_ws<8>);
    return;
  }
  k_update_matrix_tc<<<std::min(ntiles, sms * per_sm), kTCThreads, smem, cx.stream>>>(a);
}

}  // namespace dcrf
