// Microbenchmark behind DESIGN.md's "second roofline": how fast can B200 SMs pull 96-byte rows (the 21-label row of
// this engine) out of L2 / HBM when the rows are gathered at random, compared with streaming the same bytes?
// Every mean-field kernel here is such a gather (splat: pixel rows by vertex, blur: neighbour rows, update: simplex
// rows by pixel), so this number, not the HBM copy rate, is what bounds them.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_gather l2_gather.cu && ./l2_gather
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int kRowFloats = 24;  // 96 bytes
constexpr int kLanes = 6;       // lanes per row (float4 each), as in the engine

template <int INFLIGHT>
__global__ void __launch_bounds__(256) k_gather(const float* __restrict__ rows, const int* __restrict__ idx, size_t ngroups,
                                               int per_group, float* __restrict__ out, int pitch = kRowFloats) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t g = q / kLanes;
  const int c = (int)(q % kLanes);
  if (g >= ngroups) return;
  const int* my = idx + g * per_group;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = 0; t + INFLIGHT <= per_group; t += INFLIGHT) {
    int r[INFLIGHT];
    float4 x[INFLIGHT];
#pragma unroll
    for (int u = 0; u < INFLIGHT; u++) r[u] = __ldg(my + t + u);
#pragma unroll
    for (int u = 0; u < INFLIGHT; u++) x[u] = __ldg(reinterpret_cast<const float4*>(rows + (size_t)r[u] * pitch) + c);
#pragma unroll
    for (int u = 0; u < INFLIGHT; u++) acc.x += x[u].x, acc.y += x[u].y, acc.z += x[u].z, acc.w += x[u].w;
  }
  reinterpret_cast<float4*>(out + g * kRowFloats)[c] = acc;
}

// LANES lanes per row, the first `active` of them load 16 bytes each (row pitch in floats given): does it matter whether
// a quarter-warp (8 lanes = one 128-byte L1 request) spans one row or pieces of two?
template <int LANES>
__global__ void __launch_bounds__(256) k_gather_lanes(const float* __restrict__ rows, const int* __restrict__ idx, size_t ngroups,
                                                     int per_group, float* __restrict__ out, int pitch, int active) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t g = q / LANES;
  const int c = (int)(q % LANES);
  if (g >= ngroups) return;
  const int* my = idx + g * per_group;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = 0; t + 4 <= per_group; t += 4) {
    int r[4];
    float4 x[4];
#pragma unroll
    for (int u = 0; u < 4; u++) r[u] = __ldg(my + t + u);
#pragma unroll
    for (int u = 0; u < 4; u++)
      x[u] = c < active ? __ldg(reinterpret_cast<const float4*>(rows + (size_t)r[u] * pitch) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 4; u++) acc.x += x[u].x, acc.y += x[u].y, acc.z += x[u].z, acc.w += x[u].w;
  }
  if (c < active) reinterpret_cast<float4*>(out + g * 32)[c] = acc;
}

__global__ void __launch_bounds__(256) k_stream(const float4* __restrict__ p, size_t n, int reps, float* __restrict__ out) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int r = 0; r < reps; r++)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      const float4 v = __ldg(p + i);
      acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
  if (acc.x == 12345.678f) out[0] = acc.x + acc.y + acc.z + acc.w;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

int main() {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  const size_t ngroups = 2u << 20;  // 2M gathering groups (about one per 1080p pixel)
  const int per_group = 16;
  int* idx;
  float* out;
  cudaMalloc(&idx, ngroups * per_group * sizeof(int));
  cudaMalloc(&out, ngroups * kRowFloats * sizeof(float));
  std::vector<int> h(ngroups * per_group);
  for (double mb : {24.0, 64.0, 96.0, 200.0, 1024.0}) {
    const size_t nrows = (size_t)(mb * 1e6 / (kRowFloats * 4));
    float* rows;
    cudaMalloc(&rows, nrows * kRowFloats * sizeof(float));
    cudaMemset(rows, 0, nrows * kRowFloats * sizeof(float));
    unsigned long long s = 88172645463325252ull;
    for (auto& v : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; v = (int)(s % nrows); }
    cudaMemcpy(idx, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice);
    const unsigned grid = (unsigned)((ngroups * kLanes + 255) / 256);
    for (int w = 0; w < 2; w++) k_gather<4><<<grid, 256>>>(rows, idx, ngroups, per_group, out);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) k_gather<4><<<grid, 256>>>(rows, idx, ngroups, per_group, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    const double ms4 = time_ms(e0, e1) / 5;
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) k_gather<8><<<grid, 256>>>(rows, idx, ngroups, per_group, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    const double ms8 = time_ms(e0, e1) / 5;
    const double bytes = (double)ngroups * per_group * kRowFloats * 4;
    // streaming the same buffer (coalesced float4), enough repetitions to move the same number of bytes
    const size_t n4 = nrows * kRowFloats / 4;
    const int reps = (int)(bytes / (n4 * 16.0)) + 1;
    k_stream<<<148 * 16, 256>>>(reinterpret_cast<const float4*>(rows), n4, 1, out);
    cudaEventRecord(e0);
    k_stream<<<148 * 16, 256>>>(reinterpret_cast<const float4*>(rows), n4, reps, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    const double mss = time_ms(e0, e1);
    // same number of rows at a 128-byte pitch (each row inside one line): is the limit bytes or requests?
    float* rows128;
    cudaMalloc(&rows128, nrows * 32 * sizeof(float));
    cudaMemset(rows128, 0, nrows * 32 * sizeof(float));
    for (int w = 0; w < 2; w++) k_gather<4><<<grid, 256>>>(rows128, idx, ngroups, per_group, out, 32);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) k_gather<4><<<grid, 256>>>(rows128, idx, ngroups, per_group, out, 32);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    const double ms128 = time_ms(e0, e1) / 5;
    cudaFree(rows128);
    printf("   rows/s: 96-byte pitch %.1f G, 128-byte pitch %.1f G\n", ngroups * per_group / ms4 / 1e6, ngroups * per_group / ms128 / 1e6);
    {
      float* rows128b;
      cudaMalloc(&rows128b, nrows * 32 * sizeof(float));
      cudaMemset(rows128b, 0, nrows * 32 * sizeof(float));
      float* out2;
      cudaMalloc(&out2, ngroups * 32 * sizeof(float));
      auto run = [&](const char* name, int lanes, const float* tab, int pitch, int active) {
        const unsigned gr = (unsigned)((ngroups * lanes + 255) / 256);
        for (int w = 0; w < 6; w++) {
          if (w == 1) cudaEventRecord(e0);
          if (lanes == 8) k_gather_lanes<8><<<gr, 256>>>(tab, idx, ngroups, per_group, out2, pitch, active);
          else k_gather_lanes<4><<<gr, 256>>>(tab, idx, ngroups, per_group, out2, pitch, active);
        }
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        const double ms = time_ms(e0, e1) / 5;
        printf("   %-52s %.1f G rows/s\n", name, ngroups * per_group / ms / 1e6);
      };
      run("8 lanes per row, 6 active, 96-byte pitch", 8, rows, 24, 6);
      run("8 lanes per row, 6 active, 128-byte pitch", 8, rows128b, 32, 6);
      run("8 lanes per row, 8 active, 128-byte rows", 8, rows128b, 32, 8);
      run("4 lanes per row, 4 active, 64-byte rows (pitch 64)", 4, rows128b, 16, 4);
      cudaFree(rows128b);
      cudaFree(out2);
    }
    printf("buffer %6.0f MB: random 96-byte rows %.0f GB/s (4 in flight) %.0f GB/s (8 in flight); streamed %.0f GB/s   [%s]\n", mb,
           bytes / ms4 / 1e6, bytes / ms8 / 1e6, (double)n4 * 16.0 * reps / mss / 1e6, cudaGetErrorString(cudaGetLastError()));
    cudaFree(rows);
  }
  return 0;
}
