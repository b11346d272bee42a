// Micro-benchmark: random 256-byte row gathers with cp.async.bulk (TMA 1-D bulk copy) into a
// shared-memory ring, one producer lane per warp, mbarrier transaction-count completion.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_bench bulk_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\tbra WAIT;\n\tDONE:\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(smem)), "l"(gmem), "r"(bytes),
               "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

// each warp: ring of STAGES windows, each window = WIN rows of 256 B
template <int WIN, int STAGES>
__global__ void gather_bulk(const float4* __restrict__ table, const int* __restrict__ idx, int n_per_warp, float4* out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  float4* ring = reinterpret_cast<float4*>(smem_raw) + (size_t)warp * STAGES * WIN * 16;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)nw * STAGES * WIN * 256) + warp * STAGES;
  if (lane == 0) for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const int gw = blockIdx.x * nw + warp;
  const int* my = idx + (size_t)gw * n_per_warp;
  const int nwin = n_per_warp / WIN;
  auto issue = [&](int w) {
    if (lane == 0) {
      const int s = w % STAGES;
      mbar_expect_tx(&bars[s], WIN * 256);
      for (int u = 0; u < WIN; ++u) bulk_g2s(ring + (size_t)(s * WIN + u) * 16, table + (size_t)__ldg(my + w * WIN + u) * 16, 256, &bars[s]);
    }
  };
  for (int w = 0; w < STAGES - 1 && w < nwin; ++w) issue(w);
  float4 acc = make_float4(0, 0, 0, 0);
  for (int w = 0; w < nwin; ++w) {
    if (w + STAGES - 1 < nwin) issue(w + STAGES - 1);
    mbar_wait(&bars[w % STAGES], (w / STAGES) & 1);
    const float4* buf = ring + (size_t)(w % STAGES) * WIN * 16;
    for (int u = lane >> 4; u < WIN; u += 2) { const float4 v = buf[u * 16 + (lane & 15)]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
    __syncwarp();
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int WIN, int STAGES>
void run(const float4* table, const int* idx, float4* out, int blocks, int threads, int total) {
  const int warps = blocks * threads / 32;
  const int npw = (total / warps) / WIN * WIN;
  const size_t smem = (size_t)(threads / 32) * STAGES * WIN * 256 + (threads / 32) * STAGES * 8;
  cudaFuncSetAttribute(gather_bulk<WIN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 2; ++w) gather_bulk<WIN, STAGES><<<blocks, threads, smem>>>(table, idx, npw, out);
  cudaEventRecord(e0);
  for (int w = 0; w < 5; ++w) gather_bulk<WIN, STAGES><<<blocks, threads, smem>>>(table, idx, npw, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  cudaError_t err = cudaGetLastError();
  printf("bulk win %2d stages %d: blocks %4d x %3d (%4.1f warps/SM, smem %5.1f KB/CTA): %7.1f us %6.2f TB/s %s\n", WIN, STAGES, blocks,
         threads, warps / 148.0, smem / 1024.0, ms * 1e3, (double)warps * npw * 256 / ms / 1e9, err ? cudaGetErrorString(err) : "");
}

int main() {
  const int rows = 65536, total = 2500000 * 2;
  float4* table; int* idx; float4* out;
  cudaMalloc(&table, (size_t)rows * 256); cudaMemset(table, 0, (size_t)rows * 256);
  std::vector<int> h(total); srand(1);
  for (auto& v : h) v = rand() % rows;
  cudaMalloc(&idx, total * 4); cudaMemcpy(idx, h.data(), total * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&out, 148 * 64 * 32 * 16);
  run<8, 3>(table, idx, out, 148 * 2, 128, total);
  run<8, 3>(table, idx, out, 148 * 4, 128, total);
  run<8, 3>(table, idx, out, 148 * 8, 128, total);
  run<16, 3>(table, idx, out, 148 * 2, 128, total);
  run<16, 3>(table, idx, out, 148 * 4, 128, total);
  run<16, 4>(table, idx, out, 148 * 3, 128, total);
  run<4, 4>(table, idx, out, 148 * 8, 128, total);
  run<4, 4>(table, idx, out, 148 * 8, 256, total);
  return 0;
}
