// Micro-benchmark: how fast can sm_100a gather random 256-byte rows from an L2-resident table?
// (ceiling for the splat's feature gather).  nvcc -arch=sm_100a -O3 -o gather_bench gather_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

template <int WIN, bool NOALLOC>
__global__ void gather(const float4* __restrict__ table, const int* __restrict__ idx, int n_per_stream, float4* out) {
  const int stream = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, ql = threadIdx.x & 7;
  const int* my = idx + (size_t)stream * n_per_stream;
  float4 acc = make_float4(0, 0, 0, 0);
  for (int i = 0; i < n_per_stream; i += WIN) {
    int r[WIN];
#pragma unroll
    for (int u = 0; u < WIN; ++u) r[u] = __ldg(my + i + u);
    float4 a[WIN], b[WIN];
#pragma unroll
    for (int u = 0; u < WIN; ++u) {
      const float4* row = table + (size_t)r[u] * 16 + ql;
      if (NOALLOC) {
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a[u].x), "=f"(a[u].y), "=f"(a[u].z), "=f"(a[u].w) : "l"(row));
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b[u].x), "=f"(b[u].y), "=f"(b[u].z), "=f"(b[u].w) : "l"(row + 8));
      } else {
        a[u] = __ldg(row); b[u] = __ldg(row + 8);
      }
    }
#pragma unroll
    for (int u = 0; u < WIN; ++u) { acc.x += a[u].x + b[u].x; acc.y += a[u].y + b[u].y; acc.z += a[u].z + b[u].z; acc.w += a[u].w + b[u].w; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int WIN, bool NOALLOC>
void run(const float4* table, const int* idx, float4* out, int blocks, int threads, int total_recs, const char* tag,
         size_t smem = 0) {
  cudaFuncSetAttribute(gather<WIN, NOALLOC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int streams = blocks * threads / 8;
  const int nps = (total_recs / streams) / WIN * WIN;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 2; ++w) gather<WIN, NOALLOC><<<blocks, threads, smem>>>(table, idx, nps, out);
  cudaEventRecord(e0);
  for (int w = 0; w < 5; ++w) gather<WIN, NOALLOC><<<blocks, threads, smem>>>(table, idx, nps, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  const double bytes = (double)streams * nps * 256;
  printf("%-28s blocks %5d x %3d thr (%.1f warps/SM) win %d smem %3zu KB/CTA: %7.1f us  %6.2f TB/s\n", tag, blocks, threads,
         blocks * threads / 32.0 / 148, WIN, smem / 1024, ms * 1e3, bytes / ms / 1e9);
}

int main(int argc, char** argv) {
  // rows of the table: 65536 (16.8 MB, L2-resident: the splat's feature table) by default;
  // 400000 (102 MB) is the size of the backward's cell-major gradient
  const int rows = argc > 1 ? atoi(argv[1]) : 65536, total = 2500000 * 2;
  float4* table; int* idx; float4* out;
  cudaMalloc(&table, (size_t)rows * 256); cudaMemset(table, 0, (size_t)rows * 256);
  std::vector<int> h(total);
  srand(1);
  for (auto& v : h) v = rand() % rows;
  cudaMalloc(&idx, total * 4); cudaMemcpy(idx, h.data(), total * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&out, 148 * 32 * 1024 * 16);
  for (int wps : {8, 16, 24, 32, 48, 64}) {
    const int threads = 256, blocks = 148 * wps * 32 / threads;
    run<4, false>(table, idx, out, blocks, threads, total, "ldg win4");
    run<8, false>(table, idx, out, blocks, threads, total, "ldg win8");
    run<4, true>(table, idx, out, blocks, threads, total, "no_allocate win4");
  }
  // the splat's residency: 5 CTAs of 128 threads per SM, with and without its 35 KB of shared memory per CTA
  // (the shared-memory carve-out shrinks L1, which also holds the lines of the loads in flight)
  for (size_t kb : {0, 16, 35, 44}) {
    run<4, false>(table, idx, out, 148 * 5, 128, total, "ldg win4, 5 CTAs/SM", kb * 1024);
    run<4, true>(table, idx, out, 148 * 5, 128, total, "no_allocate win4, 5 CTAs/SM", kb * 1024);
  }
  return 0;
}
