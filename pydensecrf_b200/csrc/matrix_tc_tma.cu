// The mean-field update of models whose terms all carry a MatrixCompatibility, TMA-fed (round 2):
//   Q = softmax( -U - sum_k W_k * msg_k )        (labelcompatibility.cpp:82-84, densecrf.cpp:98-106,123-127 in
//                                                  /root/reference/pydensecrf/densecrf/src)
// Same mathematics as k_update_matrix_tc_ws (matrix_tc.cu): per 128-pixel tile all terms' GEMMs accumulate into one TMEM
// accumulator (3xTF32: hi*hi + lo*hi + hi*lo, fp32 accumulation), the epilogue forms x = (-U) - acc and runs the
// bit-faithful column softmax.  What changed is who moves the data and who owns a row:
//
//   warp 0   (one lane)  TMA producer: per K slab of 32 floats one cp.async.bulk.tensor for the message tile (128 rows x
//                        128 bytes) and one each for tf32(W) and W - tf32(W) (Np rows x 128 bytes), SWIZZLE_128B, landing on
//                        an mbarrier with expect-tx; rows past N and columns past Lp are zero-filled by the TMA unit.
//   warps 4-7            split the message slab in place into hi = tf32(x) and lo = x - hi (the only operand work left on
//                        the SM: 8 float4 per thread and slab).
//   warp 1   (one lane)  issues the 12 tcgen05.mma of a slab (4 k-steps x 3 products, descriptors advance 32 bytes inside
//                        the swizzle atom) and commits them to the stage's "empty" barrier / the accumulator's "full" one.
//   warps 8-15           epilogue, ONE THREAD PER PIXEL ROW straight out of TMEM (tcgen05.ld 32x32b: lane = row), two sets
//                        of four warps taking alternate tiles so a set has two tile times for its row.  Row maximum, the
//                        eight Eigen reduction chains and the shared reciprocal are thread-local: no shuffles, no
//                        shared-memory tile.  The row lives in TMEM between the passes (x, then exp(x - max), written back
//                        with tcgen05.st), so the registers hold 16 columns at a time; arithmetic is packed (FFMA2) with
//                        the opaque-constant identities of packed_math.cuh.
// The accumulator is double-buffered in TMEM (2 x 256 columns): tile t+1's main loop runs under tile t's epilogue.
#include <cuda.h>  // CUtensorMap and the encoder's signature only; the entry point comes from cudaGetDriverEntryPoint

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "common.h"
#include "packed_math.cuh"
#include "softmax_math.cuh"

namespace dcrf {
namespace {

constexpr int kTM = 128;        // pixels per tile = UMMA M
#ifndef TCU_KB
#define TCU_KB 32
#endif
constexpr int kKB = TCU_KB;     // floats per K slab = one swizzle row: 32 (SWIZZLE_128B) or 16 (SWIZZLE_64B, twice the stages)
constexpr int kRB = kKB * 4;    // bytes per operand row in shared memory
static_assert(kKB == 32 || kKB == 16, "K slab = one 128- or 64-byte swizzle row");
// warps: 0 TMA, 1 MMA, 2 TMEM allocation, 4-7 split, 8.. epilogue (four per accumulator buffer)
constexpr int kMaxBufs = 3;
constexpr int kSplitWarp0 = 4, kSplitThreads = 128;
constexpr int kEpiWarp0 = 8;
constexpr int kMaxStages = 8;

struct TmaMaps {
  CUtensorMap a[kMaxTCUTerms], bhi[kMaxTCUTerms], blo[kMaxTCUTerms];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
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
// for the roles with slack (epilogue, TMA producer): poll, then sleep, so that their waiting does not take issue slots from
// the warps on the critical path (ncu: 36 % of all issued instructions were wait-loop iterations without this)
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t mbar, uint32_t parity, uint32_t ns) {
  for (;;) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}\n"
        : "=r"(done)
        : "r"(mbar), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(ns);
  }
}
__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
// one 2-D tile: coordinates (c0 = element along the row, c1 = row), completes `bytes of the box` on the mbarrier
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t mbar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(mbar)
      : "memory");
}
// shared-memory matrix descriptor, K-major, SWIZZLE_128B / SWIZZLE_64B: rows of kRB bytes, 8-row groups 8 * kRB bytes
// apart (SBO), LBO unused for swizzled K-major operands, version 1 (Blackwell), layout type 2 / 4 in bits [61,64)
__device__ __forceinline__ uint64_t umma_desc_sw(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)((8 * kRB) >> 4) << 32) | (1ull << 46) |
         ((kRB == 128 ? 2ull : 4ull) << 61);
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
// ---- CTA pairs (cta_group::2): the even CTA of a pair issues the MMAs for both, its barriers collect the arrivals of both ----
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared-window address: same offset in the even CTA
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// plain arrive on the even CTA's barrier (what CUTLASS' umma_arrive_2x1SM_sm0 issues).  The data it announces is in the
// arriving CTA's OWN shared memory / TMEM and has been fenced for the async proxy before; a release at cluster scope here
// was measured at ~2 k cycles per arrival (profiles/r03i_tma_pair.txt).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t mbar_cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(mbar_cluster_addr) : "memory");
}
// the pair's form of the tile load: executed by both CTAs, each into its own shared memory; the bytes complete on the
// EVEN CTA's barrier (`mbar` has the rank bit cleared)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t mbar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(mbar), "r"(c0), "r"(c1), "l"(0x1000000000000000ull /* evict normal */)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
  const uint32_t z = 0;  // no output lanes disabled
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a), "l"(b), "r"(idesc), "r"(accumulate), "r"(z)
      : "memory");
}
// arrives (once the MMAs issued so far have completed) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint32_t mbar) {
  const uint16_t both = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(mbar), "h"(both)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// hi = tf32(x), round to nearest with ties away from zero (what cvt.rna.tf32.f32 gives for finite x below 2^127: the
// messages are finite), as two integer operations instead of the conversion's special-case handling; lo = x - hi is exact
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
  lo = x - hi;
}

#ifdef TCU_TIMING  // -DTCU_TIMING: block 0 prints, per role, the clock64 cycles per tile it spent waiting / working
#define TT_DECL(...) long long __VA_ARGS__
#define TT_MARK(acc, mark) { const long long n_ = clock64(); acc += n_ - mark; mark = n_; }
#else
#define TT_DECL(...)
#define TT_MARK(acc, mark)
#endif

struct TmaArgs {
  TCUArgs a;
  int nstages;
  float c_one, c_nzero, c_mone;  // 1, -0, -1 behind the kernel-argument curtain (packed_math.cuh)
};

// NB: accumulator buffers in TMEM = sets of four epilogue warps (3 when 3 * Np <= 512 columns, else 2);
// W32: rows of U and Q are 32-byte aligned (Lp % 8 == 0) and move in 256-bit pieces;
// PAIR: clusters of two CTAs on one 256-pixel tile pair (tcgen05 cta_group::2, M = 256): each CTA holds its own 128 message
//       rows and HALF of every W slab (Np / 2 rows); the even CTA's MMA lane issues for both SMs, so an SM reads and stores
//       half the W operand bytes of the single-CTA form -- the shared-memory pipe is what bounds this kernel.
template <int NB, bool W32, bool PAIR>
__global__ void __launch_bounds__(256 + 128 * NB, 1) k_update_matrix_tc_tma(const __grid_constant__ TmaMaps maps, const TmaArgs ta) {
  extern __shared__ uint8_t smem_raw[];
  const TCUArgs& a = ta.a;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = ta.nstages;
  constexpr int TPG = PAIR ? 2 : 1;  // tiles per group (CTA or CTA pair) and step
  uint32_t rank = 0;
  if (PAIR) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int ngroups = (int)gridDim.x / TPG, group = (int)blockIdx.x / TPG;
  // SWIZZLE_128B wants the stage bases on 1024-byte boundaries of the shared window
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const uint32_t szA = kTM * kRB, szB = (uint32_t)(a.Np / TPG) * kRB;
  const uint32_t stage_bytes = 2 * szA + 2 * szB;  // [A -> A_hi][A_lo][B_hi][B_lo]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + (size_t)S * stage_bytes);
  // bars: afull[S] (message tile landed; local), bfull[S] (W slabs landed), split[S] (A_hi / A_lo ready), empty[S] (MMAs
  // done reading; local), accfull[NB] (local), accempty[NB].  In a pair bfull, split and accempty are the EVEN CTA's: its
  // MMA lane is the only waiter, both CTAs arrive there.
  const uint32_t b_afull = smem_u32(bars), b_bfull = b_afull + 8 * kMaxStages, b_split = b_bfull + 8 * kMaxStages,
                 b_empty = b_split + 8 * kMaxStages, b_accfull = b_empty + 8 * kMaxStages, b_accempty = b_accfull + 8 * kMaxBufs;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * kMaxStages + 2 * kMaxBufs);

  if (tid == 0) {
    for (int s = 0; s < S; s++) {
      mbar_init(b_afull + 8 * s, 1);
      mbar_init(b_bfull + 8 * s, 1);
      mbar_init(b_split + 8 * s, TPG);
      mbar_init(b_empty + 8 * s, 1);
    }
    for (int b = 0; b < NB; b++) {
      mbar_init(b_accfull + 8 * b, 1);
      mbar_init(b_accempty + 8 * b, 4 * TPG);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if (PAIR) {  // the same warp of both CTAs
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (PAIR) cluster_sync_all();  // the peer's barriers exist before anything arrives on them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const int ntiles = (a.N + kTM - 1) / kTM;
  const int nslabs = a.Kp / kKB;         // per term
  const int gslabs = nslabs * a.nterms;  // per tile
  const uint32_t acc_cols = NB == 2 ? 256u : (uint32_t)a.Np;  // TMEM columns between the accumulator buffers

  if (warp == 0) {
    // ---------------- TMA producer ---------------------------------------------------------------------------------
    if (lane == 0) {
      int s = 0;
      uint32_t round = 0;
      TT_DECL(t_wait = 0, t_issue = 0, t_mark = clock64(); int nt = 0);
      for (int pt = group; pt * TPG < ntiles; pt += ngroups) {
        const int tile = pt * TPG + (int)rank;  // past the last tile for the odd CTA of the last pair: zero rows
        int term = 0, slab = 0;
#ifdef TCU_TIMING
        nt++;
#endif
        for (int g = 0; g < gslabs; g++) {
          TT_MARK(t_issue, t_mark);
          if (round > 0) mbar_wait_relaxed(b_empty + 8 * s, (round - 1) & 1, 64);
          TT_MARK(t_wait, t_mark);
          const uint32_t st = base + (uint32_t)s * stage_bytes, fa = b_afull + 8 * s, fb = b_bfull + 8 * s;
          mbar_expect_tx(fa, szA);
          tma_load_2d(st, &maps.a[term], slab * kKB, tile * kTM, fa);
          if (PAIR) {  // this CTA's half of the W rows; all four half-slabs of the pair complete on the even CTA's barrier
            if (rank == 0) mbar_expect_tx(fb, 4 * szB);
            const int row0 = (int)rank * (a.Np / 2);
            tma_load_2d_pair(st + 2 * szA, &maps.bhi[term], slab * kKB, row0, fb & kPeerMask);
            tma_load_2d_pair(st + 2 * szA + szB, &maps.blo[term], slab * kKB, row0, fb & kPeerMask);
          } else {
            mbar_expect_tx(fb, 2 * szB);
            tma_load_2d(st + 2 * szA, &maps.bhi[term], slab * kKB, 0, fb);
            tma_load_2d(st + 2 * szA + szB, &maps.blo[term], slab * kKB, 0, fb);
          }
          if (++slab == nslabs) slab = 0, term++;
          if (++s == S) s = 0, round++;
        }
      }
#ifdef TCU_TIMING
      if (blockIdx.x == 0) printf("[tma] producer, %d tiles: wait empty %lld, issue %lld cycles per tile\n", nt, t_wait / nt, t_issue / nt);
#endif
    }
  } else if (warp == 1) {
    // ---------------- MMA issue ------------------------------------------------------------------------------------
    if (lane == 0 && gslabs > 0 && rank == 0) {
      // instruction descriptor: D fp32, A/B tf32, both K-major, N = Np, M = 128 (256 over the two SMs of a pair)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.Np >> 3) << 17) | ((uint32_t)((kTM * TPG) >> 4) << 24);
      int s = 0;
      uint32_t round = 0, tl = 0;
      TT_DECL(t_acc = 0, t_full = 0, t_split = 0, t_issue = 0, t_mark = clock64());
      for (int pt = group; pt * TPG < ntiles; pt += ngroups, tl++) {
        const uint32_t buf = tl % NB, use = tl / NB, acc = tmem + buf * acc_cols;
        TT_MARK(t_issue, t_mark);
        if (use > 0) {  // drained by the epilogue (of both CTAs) of its previous tile
          mbar_wait(b_accempty + 8 * buf, (use - 1) & 1);
        }
        TT_MARK(t_acc, t_mark);
        for (int g = 0; g < gslabs; g++) {
          TT_MARK(t_issue, t_mark);
          mbar_wait(b_bfull + 8 * s, round & 1);
          TT_MARK(t_full, t_mark);
          mbar_wait(b_split + 8 * s, round & 1);
          TT_MARK(t_split, t_mark);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t st = base + (uint32_t)s * stage_bytes;
#pragma unroll
          for (int ks = 0; ks < kKB / 8; ks++) {  // one k-step = 8 tf32 = 32 bytes inside the 128-byte swizzle row
            const uint64_t ahi = umma_desc_sw(st + ks * 32), alo = umma_desc_sw(st + szA + ks * 32);
            const uint64_t bhi = umma_desc_sw(st + 2 * szA + ks * 32), blo = umma_desc_sw(st + 2 * szA + szB + ks * 32);
            if (PAIR) {
              umma_tf32_pair(acc, ahi, bhi, idesc, (g | ks) ? 1u : 0u);
              umma_tf32_pair(acc, alo, bhi, idesc, 1u);
              umma_tf32_pair(acc, ahi, blo, idesc, 1u);
            } else {
              umma_tf32(acc, ahi, bhi, idesc, (g | ks) ? 1u : 0u);
              umma_tf32(acc, alo, bhi, idesc, 1u);
              umma_tf32(acc, ahi, blo, idesc, 1u);
            }
          }
          if (PAIR) {
            umma_commit_pair(b_empty + 8 * s);
            if (g == gslabs - 1) umma_commit_pair(b_accfull + 8 * buf);
          } else {
            umma_commit(b_empty + 8 * s);
            if (g == gslabs - 1) umma_commit(b_accfull + 8 * buf);
          }
          if (++s == S) s = 0, round++;
        }
      }
#ifdef TCU_TIMING
      if (blockIdx.x == 0 && tl) printf("[tma] mma, %u tiles: wait accumulator drained %lld, wait TMA %lld, wait split %lld, issue %lld cycles per tile\n", tl, t_acc / tl, t_full / tl, t_split / tl, t_issue / tl);
#endif
    }
  } else if (warp >= kSplitWarp0 && warp < kSplitWarp0 + kSplitThreads / 32) {
    // ---------------- message slab -> tf32 hi (in place) + lo ----------------------------------------------------------
    const int t = tid - kSplitWarp0 * 32;
    int s = 0;
    uint32_t round = 0;
    TT_DECL(t_wait = 0, t_work = 0, t_fence = 0, t_mark = clock64(); int nt = 0);
    for (int pt = group; pt * TPG < ntiles; pt += ngroups) {
#ifdef TCU_TIMING
      nt++;
#endif
      for (int g = 0; g < gslabs; g++) {
        TT_MARK(t_work, t_mark);
        mbar_wait(b_afull + 8 * s, round & 1);
        TT_MARK(t_wait, t_mark);
        float4* A = reinterpret_cast<float4*>(sm + (size_t)s * stage_bytes);
        float4* Alo = reinterpret_cast<float4*>(sm + (size_t)s * stage_bytes + szA);
        constexpr int kPer = (int)(kTM * kRB / 16) / kSplitThreads;
        float4 v[kPer];
#pragma unroll
        for (int i = 0; i < kPer; i++) v[i] = A[t + i * kSplitThreads];
#pragma unroll
        for (int i = 0; i < kPer; i++) {
          float4 h, l;
          split_tf32(v[i].x, h.x, l.x);
          split_tf32(v[i].y, h.y, l.y);
          split_tf32(v[i].z, h.z, l.z);
          split_tf32(v[i].w, h.w, l.w);
          A[t + i * kSplitThreads] = h;
          Alo[t + i * kSplitThreads] = l;
        }
        TT_MARK(t_work, t_mark);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the MMA's async proxy
        // the four split warps meet, ONE thread arrives for the CTA (128 release.cluster arrivals on the peer's barrier
        // cost 19 k cycles per tile)
        asm volatile("bar.sync 1, %0;" ::"n"(kSplitThreads) : "memory");
        if (t == 0) {
          if (PAIR) mbar_arrive_cluster((b_split + 8 * s) & kPeerMask); else mbar_arrive(b_split + 8 * s);
        }
        TT_MARK(t_fence, t_mark);
        if (++s == S) s = 0, round++;
      }
    }
#ifdef TCU_TIMING
    if (blockIdx.x == 0 && t == 0) printf("[tma] split, %d tiles: wait TMA %lld, work %lld, fence + arrive %lld cycles per tile\n", nt, t_wait / nt, t_work / nt, t_fence / nt);
#endif
  } else if (warp >= kEpiWarp0) {
    // ---------------- epilogue: one thread per pixel row -----------------------------------------------------------------
    const int set = (warp - kEpiWarp0) >> 2, q = warp & 3;  // a warp may touch the TMEM lanes 32 * (warp % 4) ..
    const int L = a.L, Lp = a.Lp, nfull = L >> 2, ntail = L & 3;
    const int lim4 = nfull * 4, lim8 = (nfull >> 1) * 8;
    Pk c;
    c.one = pk1(ta.c_one), c.nzero = pk1(ta.c_nzero), c.mone = pk1(ta.c_mone);
    const u64 Z = pk1(a.neg_unary ? -0.f : 0.f);
    uint32_t tl = 0;
    TT_DECL(t_wait = 0, t_p1 = 0, t_p2 = 0, t_p3 = 0, t_mark = clock64(); int nt = 0);
    for (int pt = group; pt * TPG < ntiles; pt += ngroups, tl++) {
      if ((int)(tl % NB) != set) continue;
      const int tile = pt * TPG + (int)rank;
#ifdef TCU_TIMING
      nt++;
#endif
      const uint32_t buf = set, use = tl / NB;
      const uint32_t tad = tmem + buf * acc_cols + ((uint32_t)(q * 32) << 16);
      const int i = tile * kTM + q * 32 + lane;
      const bool valid = i < a.N;
      const float* urow = (a.U && valid) ? a.U + (size_t)i * Lp : nullptr;
      // 16 unary values as pairs; two register sets that trade places chunk by chunk, so the loads of chunk c+1 are in
      // flight while chunk c is processed (a copy from a single "next" set would wait for them at once)
      u64 ua[8], ub[8];
      auto load_u = [&](int c0, u64 (&un)[8]) {
        if (W32) {
#pragma unroll
          for (int j = 0; j < 2; j++) {
            const F8 v = (urow && c0 + 8 * j < Lp) ? ld8_nc(urow + c0 + 8 * j) : zero8();
#pragma unroll
            for (int k = 0; k < 4; k++) un[4 * j + k] = v.p[k];
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const float4 v = (urow && c0 + 4 * j < Lp) ? __ldg(reinterpret_cast<const float4*>(urow + c0 + 4 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
            un[2 * j] = pk2(v.x, v.y), un[2 * j + 1] = pk2(v.z, v.w);
          }
        }
      };
      load_u(0, ua);
      // the rest of this tile's unary rows on their way into L2 while the set waits for the accumulator (about one tile
      // time: prefetching the set's NEXT tile instead, three tile times ahead, was evicted again before use -- ncu showed
      // the unary read twice from DRAM)
      if (urow) {
        const char* nu = reinterpret_cast<const char*>(urow);
        for (int b = 128; b < Lp * 4; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nu + b));
      }
      TT_MARK(t_p3, t_mark);
      if (gslabs > 0) mbar_wait_relaxed(b_accfull + 8 * buf, use & 1, 256);  // no terms: Q = softmax(-U), the accumulator reads as 0
      TT_MARK(t_wait, t_mark);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

      // Every pass walks the row in chunks of 16 columns; chunks that lie wholly inside the region a pass treats uniformly
      // take the branch-free body (edge = false), the last one or two the general one.
      // pass 1: x = (-U) - accumulator back into TMEM, row maximum and minimum
      float mx = -INFINITY, mn = INFINITY;
      auto pass1 = [&](int c0, u64 (&uc)[8], u64 (&un)[8], auto edge_t) {
        constexpr bool EDGE = decltype(edge_t)::value;
        uint32_t r[16];
        if (gslabs > 0) {
          tmem_ld16(tad + c0, r);
        } else {
#pragma unroll
          for (int j = 0; j < 16; j++) r[j] = 0u;
        }
        if (c0 + 16 < Lp) load_u(c0 + 16, un);
        tmem_wait_ld();
#pragma unroll
        for (int p = 0; p < 8; p++) {
          const int l = c0 + 2 * p;
          const u64 x = sub2(sub2(Z, uc[p], c), pk2(__uint_as_float(r[2 * p]), __uint_as_float(r[2 * p + 1])), c);
          if (!EDGE || l + 1 < L) {
            mx = fmaxf(mx, fmaxf(lo2(x), hi2(x)));
            mn = fminf(mn, fminf(lo2(x), hi2(x)));
          } else if (l < L) {
            mx = fmaxf(mx, lo2(x));
            mn = fminf(mn, lo2(x));
          }
          r[2 * p] = __float_as_uint(lo2(x)), r[2 * p + 1] = __float_as_uint(hi2(x));
        }
        tmem_st16(tad + c0, r);
      };
      auto pass1_at = [&](int c0, u64 (&uc)[8], u64 (&un)[8]) {
        if (c0 + 16 <= L)
          pass1(c0, uc, un, std::false_type());
        else
          pass1(c0, uc, un, std::true_type());
      };
      int c0 = 0;
      for (; c0 < Lp; c0 += 32) {
        pass1_at(c0, ua, ub);
        if (c0 + 16 < Lp) pass1_at(c0 + 16, ub, ua);
      }
      tmem_wait_st();
      TT_MARK(t_p1, t_mark);
      // some exp(x - max) of this row is below the range the shared-reciprocal division is proven for (x - max is
      // monotonic in x, so the row minimum decides)
      const bool small = __fsub_rn(mn, mx) < -44.f;

      // pass 2: e = exp(x - max) back into TMEM; Eigen's redux order: eight sequential chains over the packet pairs,
      // p0 + p1, the odd trailing packet, (c0 + c2) + (c1 + c3), then the L % 4 tail one by one
      const u64 MX = pk1(mx);
      u64 ch[4] = {0ull, 0ull, 0ull, 0ull}, tp01 = 0ull, tp23 = 0ull;
      float tx0 = 0.f, tx1 = 0.f, tx2 = 0.f;
      auto pass2 = [&](int c0, auto edge_t) {
        constexpr bool EDGE = decltype(edge_t)::value;
        uint32_t r[16];
        tmem_ld16(tad + c0, r);
        tmem_wait_ld();
#pragma unroll
        for (int p = 0; p < 8; p++) {
          const int l = c0 + 2 * p;
          const u64 xs = sub2(pk2(__uint_as_float(r[2 * p]), __uint_as_float(r[2 * p + 1])), MX, c);
          u64 e = 0ull;
          if (!EDGE) {
            e = eigen_pexp_nonpos2(xs, c);
            ch[p & 3] = add2(ch[p & 3], e, c);
          } else if (l < lim4) {
            e = eigen_pexp_nonpos2(xs, c);
            if (l < lim8) {
              ch[p & 3] = add2(ch[p & 3], e, c);
            } else if ((p & 3) == 0) {
              tp01 = e;
            } else if ((p & 3) == 1) {
              tp23 = e;
            }
          } else if (l == lim4) {
            tx0 = lo2(xs), tx1 = hi2(xs);
          } else if (l == lim4 + 2) {
            tx2 = lo2(xs);
          }
          r[2 * p] = (uint32_t)e, r[2 * p + 1] = (uint32_t)(e >> 32);
        }
        tmem_st16(tad + c0, r);
      };
      for (c0 = 0; c0 + 16 <= lim8; c0 += 16) pass2(c0, std::false_type());
      for (; c0 < Lp; c0 += 16) pass2(c0, std::true_type());
      const float te0 = ntail > 0 ? tail_exp_nonpos(tx0) : 0.f;
      const float te1 = ntail > 1 ? tail_exp_nonpos(tx1) : 0.f;
      const float te2 = ntail > 2 ? tail_exp_nonpos(tx2) : 0.f;
      u64 c01 = add2(ch[0], ch[2], c), c23 = add2(ch[1], ch[3], c);
      if (nfull & 1) c01 = add2(c01, tp01, c), c23 = add2(c23, tp23, c);
      float s = __fadd_rn(__fadd_rn(lo2(c01), lo2(c23)), __fadd_rn(hi2(c01), hi2(c23)));
      if (ntail > 0) s = __fadd_rn(s, te0);
      if (ntail > 1) s = __fadd_rn(s, te1);
      if (ntail > 2) s = __fadd_rn(s, te2);
      tmem_wait_st();
      TT_MARK(t_p2, t_mark);

      // pass 3: out.col = e / sum (one refined reciprocal per row, IEEE division outside its proven range)
      const float rg = refined_rcp(s);
      const u64 R = pk1(rg), NS = pk1(-s);
      float* qrow = a.Q + (size_t)(valid ? i : 0) * Lp;
      auto pass3 = [&](int c0, auto edge_t) {
        constexpr bool EDGE = decltype(edge_t)::value;
        uint32_t r[16];
        tmem_ld16(tad + c0, r);
        tmem_wait_ld();
        float o[16];
#pragma unroll
        for (int p = 0; p < 8; p++) {
          const int l = c0 + 2 * p;
          u64 e = pk2(__uint_as_float(r[2 * p]), __uint_as_float(r[2 * p + 1]));
          if (EDGE) {
            if (l == lim4) e = pk2(te0, te1);
            if (l == lim4 + 2) e = pk2(te2, 0.f);
          }
          u64 qv;
          if (small) {
            qv = pk2(__fdiv_rn(lo2(e), s), __fdiv_rn(hi2(e), s));
          } else {  // div_shared (softmax_math.cuh) for two
            const u64 q0 = fma2(e, R, 0ull);
            const u64 rem = fma2(NS, q0, e);
            qv = fma2(R, rem, q0);
          }
          o[2 * p] = lo2(qv), o[2 * p + 1] = hi2(qv);
        }
        if (valid) {
          if (W32) {
#pragma unroll
            for (int j = 0; j < 2; j++)
              if (!EDGE || c0 + 8 * j < Lp) {
                F8 v;
#pragma unroll
                for (int k = 0; k < 4; k++) v.p[k] = pk2(o[8 * j + 2 * k], o[8 * j + 2 * k + 1]);
                st8(qrow + c0 + 8 * j, v);
              }
          } else {
#pragma unroll
            for (int j = 0; j < 4; j++)
              if (!EDGE || c0 + 4 * j < Lp)
                *reinterpret_cast<float4*>(qrow + c0 + 4 * j) = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
          }
        }
      };
      for (c0 = 0; c0 + 16 <= lim4; c0 += 16) pass3(c0, std::false_type());
      for (; c0 < Lp; c0 += 16) pass3(c0, std::true_type());
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {  // the accumulator buffer may be overwritten
        if (PAIR) mbar_arrive_cluster((b_accempty + 8 * buf) & kPeerMask); else mbar_arrive(b_accempty + 8 * buf);
      }
      TT_MARK(t_p3, t_mark);
    }
#ifdef TCU_TIMING
    if (blockIdx.x == 0 && lane == 0 && q == 0 && nt) printf("[tma] epilogue set %d, %d tiles: wait accumulator %lld, pass 1 %lld, pass 2 %lld, pass 3 %lld cycles per tile\n", set, nt, t_wait / nt, t_p1 / nt, t_p2 / nt, t_p3 / nt);
#endif
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (PAIR) cluster_sync_all();  // neither CTA leaves (or frees TMEM) while the other may still touch it
  if (warp == 2) {
    if (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encoder() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// fp32 matrix of `rows` rows of `cols` floats (row pitch `pitch` floats) -> boxes of 32 floats x `box_rows` rows, SWIZZLE_128B
void make_map(CUtensorMap* m, const float* ptr, uint64_t cols, uint64_t rows, uint64_t pitch, uint32_t box_rows) {
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {pitch * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)kKB, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encoder()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, kRB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DCRF_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed");
}

}  // namespace

// shared-memory plan of the TMA kernel: stages of [A][A_lo][B_hi][B_lo], 1024-byte aligned, barriers behind them
static int tma_stages(int Np, bool pair, size_t* smem) {
  const size_t stage = 2 * (size_t)kTM * kRB + 2 * (size_t)(pair ? Np / 2 : Np) * kRB;
  const size_t fixed = 1024 /* alignment slack */ + 8 * (4 * kMaxStages + 2 * kMaxBufs) + 16;
  int S = (int)((227 * 1024 - fixed) / stage);
  if (S > kMaxStages) S = kMaxStages;
  *smem = fixed + (size_t)S * stage;
  return S;
}

bool update_matrix_tc_tma_supported(RowLayout rl, int nterms) {
  static const bool on = [] {  // DCRF_B200_TCU_TMA=0: no fused tensor-core update; the staged slice -> GEMM -> softmax path (A/B timing)
    const char* e = std::getenv("DCRF_B200_TCU_TMA");
    return !e || std::atoi(e) != 0;
  }();
  if (!on || !matrix_tc_supported(rl) || nterms < 0 || nterms > kMaxTCUTerms || !encoder()) return false;
  int Np, Kp;
  matrix_tc_dims(rl, &Np, &Kp);
  size_t smem;
  return Kp % kKB == 0 && tma_stages(Np, false, &smem) >= 2;
}

// a.nterms == 0: Q = softmax(-U) alone (the start of inference, densecrf.cpp:118-119), the epilogue without a main loop
void launch_update_matrix_tc_tma(Ctx cx, RowLayout rl, TCUArgs a) {
  LaunchScope ls(cx.prof, cx.stream, a.nterms ? "update_matrix_tc" : "init_softmax_rows");
  a.L = rl.L, a.Lp = rl.Lp;
  matrix_tc_dims(rl, &a.Np, &a.Kp);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int ntiles = (a.N + kTM - 1) / kTM;
  // CTA pairs (cta_group::2) when there is a main loop and more than one tile: DCRF_B200_TCU_PAIR=0 for the single-CTA form
  static const bool pair_on = [] {
    const char* e = std::getenv("DCRF_B200_TCU_PAIR");
    return !e || std::atoi(e) != 0;
  }();
  static std::atomic<bool> pair_ok{true};  // cleared when the device cannot run the two-CTA clusters
  bool pair = pair_on && pair_ok && a.nterms > 0 && ntiles >= 2 && sms >= 2;
  for (;;) {
    TmaArgs ta;
    size_t smem;
    ta.nstages = tma_stages(a.Np, pair, &smem);
    ta.c_one = 1.f, ta.c_nzero = -0.f, ta.c_mone = -1.f;
    TmaMaps maps;
    std::memset(&maps, 0, sizeof(maps));
    const uint32_t brows = (uint32_t)(pair ? a.Np / 2 : a.Np);  // W rows per CTA and slab
    for (int k = 0; k < a.nterms; k++) {
      make_map(&maps.a[k], a.t[k].msg, (uint64_t)a.Lp, (uint64_t)a.N, (uint64_t)a.Lp, kTM);
      make_map(&maps.bhi[k], a.t[k].Bhi, (uint64_t)a.Kp, (uint64_t)a.Np, (uint64_t)a.Kp, brows);
      make_map(&maps.blo[k], a.t[k].Blo, (uint64_t)a.Kp, (uint64_t)a.Np, (uint64_t)a.Kp, brows);
    }
    ta.a = a;
    auto go = [&](auto kern, int nb) {
      DCRF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      cudaLaunchConfig_t cfg{};
      cfg.blockDim = dim3(256 + 128 * nb);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = cx.stream;
      cudaLaunchAttribute attr[1];
      if (pair) {
        cfg.gridDim = dim3(2 * std::min((ntiles + 1) / 2, sms / 2));
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr, cfg.numAttrs = 1;
        int clusters = 0;  // a device (partition) that cannot co-schedule two such CTAs reports 0: the caller falls back
        if (cudaOccupancyMaxActiveClusters(&clusters, kern, &cfg) != cudaSuccess || clusters < 1) {
          cudaGetLastError();
          return false;
        }
      } else {
        cfg.gridDim = dim3(std::min(ntiles, sms));
      }
      DCRF_CUDA(cudaLaunchKernelEx(&cfg, kern, maps, ta));
      return true;
    };
    const bool w32 = a.Lp % 8 == 0;
    const int nb = 3 * a.Np <= 512 ? 3 : 2;
    if (pair) {
      bool ok;
      if (nb == 3)
        ok = w32 ? go(k_update_matrix_tc_tma<3, true, true>, 3) : go(k_update_matrix_tc_tma<3, false, true>, 3);
      else
        ok = w32 ? go(k_update_matrix_tc_tma<2, true, true>, 2) : go(k_update_matrix_tc_tma<2, false, true>, 2);
      if (ok) return;
      pair_ok = false, pair = false;  // once more in the single-CTA form
      continue;
    } else {
      if (nb == 3) {
        if (w32) go(k_update_matrix_tc_tma<3, true, false>, 3); else go(k_update_matrix_tc_tma<3, false, false>, 3);
      } else {
        if (w32) go(k_update_matrix_tc_tma<2, true, false>, 2); else go(k_update_matrix_tc_tma<2, false, false>, 2);
      }
    }
    return;
  }
}

}  // namespace dcrf
