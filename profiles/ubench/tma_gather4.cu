// Microbenchmark: row-gather rate of the Blackwell TMA gather4 path (cp.async.bulk.tensor.2d ... tile::gather4: four rows
// of a 2-D tensor, given by four row indices, land in shared memory per instruction) against the per-thread LSU gathers
// the round-1 kernels use (profiles/ubench/l2_gather.cu: ~120 G rows/s of 96-byte rows out of L2).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_gather4 tma_gather4.cu -lcuda && ./tma_gather4 [box_rows]
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int kRowFloats = 24;          // 96 bytes
constexpr int kRowsPerStage = 32;       // 8 gather4 instructions per warp and stage
constexpr int kStages = 2;
constexpr int kWarps = 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, 0x989680;\n\tselp.b32 %0, 1, 0, P1;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* map, int col, int r0, int r1, int r2, int r3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
      "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
      : "memory");
}

__global__ void __launch_bounds__(kWarps * 32) k_tma(const __grid_constant__ CUtensorMap map, const int* __restrict__ idx,
                                                    int stages_per_warp, float* __restrict__ out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[kWarps * kStages];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t stage_bytes = (size_t)kRowsPerStage * kRowFloats * 4;
  uint8_t* mine = smem + (size_t)warp * kStages * stage_bytes;
  if (lane == 0)
    for (int s = 0; s < kStages; s++) mbar_init(smem_u32(&bars[warp * kStages + s]), 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const size_t gw = (size_t)blockIdx.x * kWarps + warp;
  const int* my = idx + gw * (size_t)stages_per_warp * kRowsPerStage;
  auto issue = [&](int st) {
    const uint32_t bar = smem_u32(&bars[warp * kStages + (st % kStages)]);
    if (lane == 0) mbar_expect_tx(bar, (uint32_t)stage_bytes);
    __syncwarp();
    if (lane < kRowsPerStage / 4) {
      const int4 r = *reinterpret_cast<const int4*>(my + (size_t)st * kRowsPerStage + lane * 4);
      tma_gather4(smem_u32(mine + (st % kStages) * stage_bytes + (size_t)lane * 4 * kRowFloats * 4), &map, 0, r.x, r.y, r.z, r.w, bar);
    }
  };
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  issue(0);
  for (int st = 0; st < stages_per_warp; st++) {
    if (st + 1 < stages_per_warp) issue(st + 1);
    mbar_wait(smem_u32(&bars[warp * kStages + (st % kStages)]), (st / kStages) & 1);
    const float4* rows = reinterpret_cast<const float4*>(mine + (st % kStages) * stage_bytes);
    for (int e = lane; e < kRowsPerStage * kRowFloats / 4; e += 32) {
      const float4 v = rows[e];
      acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
    __syncwarp();
  }
  if (acc.x == 123.456f) out[0] = acc.x + acc.y + acc.z + acc.w;
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s\n", cudaGetErrorString(e_), #x); return 1; } } while (0)

int main(int argc, char** argv) {
  const int box_rows = argc > 1 ? atoi(argv[1]) : 1;
  cudaFree(0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  const int grid = 148 * 4, stages_per_warp = 64;
  const size_t total_rows = (size_t)grid * kWarps * stages_per_warp * kRowsPerStage;
  int* idx;
  float* out;
  CK(cudaMalloc(&idx, total_rows * sizeof(int)));
  CK(cudaMalloc(&out, 1024));
  std::vector<int> h(total_rows);
  const size_t smem = (size_t)kWarps * kStages * kRowsPerStage * kRowFloats * 4;
  CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (double mb : {24.0, 64.0, 96.0, 200.0, 1024.0}) {
    const size_t nrows = (size_t)(mb * 1e6 / (kRowFloats * 4));
    float* rows;
    CK(cudaMalloc(&rows, nrows * kRowFloats * sizeof(float)));
    CK(cudaMemset(rows, 0, nrows * kRowFloats * sizeof(float)));
    unsigned long long s = 88172645463325252ull;
    for (auto& v : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; v = (int)(s % nrows); }
    CK(cudaMemcpy(idx, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice));
    CUtensorMap map;
    cuuint64_t gdim[2] = {(cuuint64_t)kRowFloats, (cuuint64_t)nrows};
    cuuint64_t gstride[1] = {(cuuint64_t)kRowFloats * 4};
    cuuint32_t box[2] = {(cuuint32_t)kRowFloats, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, rows, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return 1; }
    k_tma<<<grid, kWarps * 32, smem>>>(map, idx, stages_per_warp, out);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int rep = 0; rep < 5; rep++) k_tma<<<grid, kWarps * 32, smem>>>(map, idx, stages_per_warp, out);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 5;
    printf("table %6.0f MB: TMA gather4 (box %d row) %.1f G rows/s = %.0f GB/s of 96-byte rows\n", mb, box_rows, total_rows / ms / 1e6,
           total_rows * 96.0 / ms / 1e6);
    cudaFree(rows);
  }
  return 0;
}
