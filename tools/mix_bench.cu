// Micro-benchmark: are the LDG path and the TMA bulk-copy path additive for random 256-byte row
// gathers?  Per warp and window: NL rows through quarter-warp LDG.128 pairs (the splat's pattern)
// and NB rows through cp.async.bulk into a shared-memory ring (consumed with LDS one window later).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mix_bench mix_bench.cu
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

// NLQ = LDG rows per quarter-warp per window (0 or 4), NB = bulk rows per warp per window (multiple of 4 or 0)
template <int NLQ, int NB, int STAGES>
__global__ void mix(const float4* __restrict__ table, const int* __restrict__ idx, int nwin, float4* out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5, ql = lane & 7, qw = lane >> 3;
  constexpr int NBS = NB ? NB : 1;
  float4* ring = reinterpret_cast<float4*>(smem_raw) + (size_t)warp * STAGES * NBS * 16;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)nw * STAGES * NBS * 256) + warp * STAGES;
  if (NB) {
    if (lane == 0) for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
  }
  constexpr int PER = 4 * NLQ + NB;            // rows per warp per window
  const int gw = blockIdx.x * nw + warp;
  const int* my = idx + (size_t)gw * nwin * PER;
  auto issue = [&](int w) {
    if (NB && lane == 0) {
      const int s = w % STAGES;
      mbar_expect_tx(&bars[s], NB * 256);
      for (int u = 0; u < NB; ++u)
        bulk_g2s(ring + (size_t)(s * NB + u) * 16, table + (size_t)__ldg(my + w * PER + 4 * NLQ + u) * 16, 256, &bars[s]);
    }
  };
  for (int w = 0; w < STAGES - 1 && w < nwin; ++w) issue(w);
  float4 acc = make_float4(0, 0, 0, 0);
  for (int w = 0; w < nwin; ++w) {
    if (w + STAGES - 1 < nwin) issue(w + STAGES - 1);
    float4 a[NLQ ? NLQ : 1], b[NLQ ? NLQ : 1];
    if (NLQ) {
      int r[NLQ ? NLQ : 1];
#pragma unroll
      for (int u = 0; u < NLQ; ++u) r[u] = __ldg(my + w * PER + qw * NLQ + u);
#pragma unroll
      for (int u = 0; u < NLQ; ++u) {
        const float4* row = table + (size_t)r[u] * 16 + ql;
        a[u] = __ldg(row); b[u] = __ldg(row + 8);
      }
    }
    if (NB) {
      mbar_wait(&bars[w % STAGES], (w / STAGES) & 1);
      const float4* buf = ring + (size_t)(w % STAGES) * NB * 16;
#pragma unroll
      for (int u = 0; u < NB / 4; ++u) {
        const float4 v0 = buf[(qw * (NB / 4) + u) * 16 + ql], v1 = buf[(qw * (NB / 4) + u) * 16 + ql + 8];
        acc.x += v0.x + v1.x; acc.y += v0.y + v1.y; acc.z += v0.z + v1.z; acc.w += v0.w + v1.w;
      }
      __syncwarp();
    }
    if (NLQ) {
#pragma unroll
      for (int u = 0; u < NLQ; ++u) { acc.x += a[u].x + b[u].x; acc.y += a[u].y + b[u].y; acc.z += a[u].z + b[u].z; acc.w += a[u].w + b[u].w; }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int NLQ, int NB, int STAGES>
void run(const float4* table, const int* idx, float4* out, int ctas_per_sm, int threads, int total) {
  const int blocks = 148 * ctas_per_sm, warps = blocks * threads / 32;
  constexpr int PER = 4 * NLQ + NB;
  const int nwin = total / warps / PER;
  const size_t smem = (size_t)(threads / 32) * STAGES * (NB ? NB : 1) * 256 + (threads / 32) * STAGES * 8;
  cudaFuncSetAttribute(mix<NLQ, NB, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 2; ++w) mix<NLQ, NB, STAGES><<<blocks, threads, smem>>>(table, idx, nwin, out);
  cudaEventRecord(e0);
  for (int w = 0; w < 5; ++w) mix<NLQ, NB, STAGES><<<blocks, threads, smem>>>(table, idx, nwin, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  cudaError_t err = cudaGetLastError();
  printf("ldg %2d + bulk %2d rows/warp/window, stages %d, %4.1f warps/SM, smem %5.1f KB/CTA: %7.1f us  %6.2f TB/s %s\n", 4 * NLQ, NB,
         STAGES, warps / 148.0, smem / 1024.0, ms * 1e3, (double)warps * nwin * PER * 256 / ms / 1e9, err ? cudaGetErrorString(err) : "");
}

int main() {
  const int rows = 65536, total = 2500000 * 2;
  float4* table; int* idx; float4* out;
  cudaMalloc(&table, (size_t)rows * 256); cudaMemset(table, 0, (size_t)rows * 256);
  std::vector<int> h(total); srand(1);
  for (auto& v : h) v = rand() % rows;
  cudaMalloc(&idx, total * 4); cudaMemcpy(idx, h.data(), total * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&out, 148 * 64 * 32 * 16);
  for (int c : {3, 5, 8}) {                       // CTAs of 128 threads per SM: 12, 20, 32 warps/SM
    run<4, 0, 3>(table, idx, out, c, 128, total);
    run<4, 4, 3>(table, idx, out, c, 128, total);
    run<4, 8, 3>(table, idx, out, c, 128, total);
    run<4, 16, 3>(table, idx, out, c, 128, total);
    run<0, 16, 3>(table, idx, out, c, 128, total);
  }
  return 0;
}
