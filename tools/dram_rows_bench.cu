// Micro-benchmark: what does a B200 deliver for COLD random 256-byte rows (the backward gather's access pattern:
// a half-warp per row, 16 lanes x 16 bytes, 8 rows in flight per half-warp), every row read exactly once from a
// table far larger than L2?  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dram_rows_bench dram_rows_bench.cu
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include <cuda_runtime.h>

template <int WIN>
__global__ void gather_rows(const float4* __restrict__ table, const int* __restrict__ idx, int n_per_hw, float4* out) {
  const int hwid = (blockIdx.x * blockDim.x + threadIdx.x) >> 4, hl = threadIdx.x & 15;
  const int* my = idx + (size_t)hwid * n_per_hw;
  float4 acc = make_float4(0, 0, 0, 0);
  for (int i = 0; i < n_per_hw; i += WIN) {
    float4 a[WIN];
#pragma unroll
    for (int u = 0; u < WIN; ++u) a[u] = __ldg(table + (size_t)__ldg(my + i + u) * 16 + hl);
#pragma unroll
    for (int u = 0; u < WIN; ++u) { acc.x += a[u].x; acc.y += a[u].y; acc.z += a[u].z; acc.w += a[u].w; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void stream_rows(const float4* __restrict__ table, size_t n4, float4* out) {
  float4 acc = make_float4(0, 0, 0, 0);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(table + i);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main(int argc, char** argv) {
  const int rows = argc > 1 ? atoi(argv[1]) : 2000000;           // 2 M rows = 512 MB
  float4* table; int* idx; float4* out;
  cudaMalloc(&table, (size_t)rows * 256); cudaMemset(table, 0, (size_t)rows * 256);
  std::vector<int> h(rows);
  for (int i = 0; i < rows; ++i) h[i] = i;
  std::mt19937 rng(1);
  std::shuffle(h.begin(), h.end(), rng);
  cudaMalloc(&idx, (size_t)rows * 4); cudaMemcpy(idx, h.data(), (size_t)rows * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&out, (size_t)148 * 64 * 32 * 16);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int wps : {16, 24, 32, 48, 64}) {
    const int threads = 256, blocks = 148 * wps * 32 / threads, halfwarps = blocks * threads / 16;
    const int nph = rows / halfwarps / 8 * 8;
    for (int w = 0; w < 2; ++w) gather_rows<8><<<blocks, threads>>>(table, idx, nph, out);
    cudaEventRecord(e0);
    for (int w = 0; w < 5; ++w) gather_rows<8><<<blocks, threads>>>(table, idx, nph, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    printf("random 256-B rows, each once: %2d warps/SM, 8 rows in flight per half-warp: %7.1f us  %5.2f TB/s\n", wps, ms * 1e3,
           (double)halfwarps * nph * 256 / ms / 1e9);
  }
  cudaEventRecord(e0);
  for (int w = 0; w < 5; ++w) stream_rows<<<148 * 8, 256>>>(table, (size_t)rows * 16, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  printf("the same table streamed linearly: %7.1f us  %5.2f TB/s\n", ms * 1e3, (double)rows * 256 / ms / 1e9);
  return 0;
}
