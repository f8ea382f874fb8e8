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
  return matrix_tc_supported(rl) && nterms >= 1 && nterms <= kMaxTCUTerms && update_matrix_tc_tma_supported(rl, nterms);
}

void launch_update_matrix_tc(Ctx cx, RowLayout rl, TCUArgs a) { launch_update_matrix_tc_tma(cx, rl, a); }

}  // namespace dcrf
