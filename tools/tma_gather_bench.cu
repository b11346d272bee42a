// Micro-benchmark: random 256-byte row gathers through the Blackwell TMA gather4 path
// (cp.async.bulk.tensor.2d.tile::gather4: four arbitrary rows of a 2-D tensor per instruction, landing in
// shared memory, completion on an mbarrier) against the LDG.128 gather the library uses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_gather_bench tools/tma_gather_bench.cu
//   tools/tma_gather_bench [rows]      (640000 rows = 164 MB: the gradient of a B=16 step; 65536 = L2-resident)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

#define ROW_FLOATS 64
#define ROW_BYTES 256
#define ROWS_PER_STAGE 16      // one depth window of the gather kernel: 4 gather4 instructions, 4 KB

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_gather4(unsigned dst, const CUtensorMap* tm, unsigned bar, int col, int4 r) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst), "l"(tm), "r"(bar), "r"(col), "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w)
      : "memory");
}

// A warp owns a ring of NS stages of 16 rows.  Lane 0 issues the four gather4 copies of a stage; all lanes
// wait on the stage's mbarrier, read the 16 rows (half-warp per row, 16 bytes per lane) and add them up.
template <int NS>
__global__ void __launch_bounds__(256) tma_gather(const __grid_constant__ CUtensorMap tm, const int* __restrict__ idx,
                                                  int stages_per_warp, float4* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  unsigned char* ring = smem + (size_t)warp * NS * ROWS_PER_STAGE * ROW_BYTES;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + (size_t)nwarps * NS * ROWS_PER_STAGE * ROW_BYTES) + warp * NS;
  if (lane == 0) {
    for (int s = 0; s < NS; ++s) mbar_init(smem_u32(bars + s), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  const int gw = blockIdx.x * nwarps + warp;
  const int4* my = reinterpret_cast<const int4*>(idx + (size_t)gw * stages_per_warp * ROWS_PER_STAGE);
  auto issue = [&](int stage_no) {       // lane 0 only
    const int s = stage_no % NS;
    const unsigned bar = smem_u32(bars + s), dst = smem_u32(ring + (size_t)s * ROWS_PER_STAGE * ROW_BYTES);
    mbar_expect_tx(bar, ROWS_PER_STAGE * ROW_BYTES);
#pragma unroll
    for (int q = 0; q < ROWS_PER_STAGE / 4; ++q) tma_gather4(dst + q * 4 * ROW_BYTES, &tm, bar, 0, __ldg(my + stage_no * 4 + q));
  };
  if (lane == 0)
    for (int s = 0; s < NS && s < stages_per_warp; ++s) issue(s);
  float4 acc = make_float4(0, 0, 0, 0);
  for (int st = 0; st < stages_per_warp; ++st) {
    const int s = st % NS;
    mbar_wait(smem_u32(bars + s), (st / NS) & 1);
    const float4* rows = reinterpret_cast<const float4*>(ring + (size_t)s * ROWS_PER_STAGE * ROW_BYTES);
#pragma unroll
    for (int u = 0; u < ROWS_PER_STAGE / 2; ++u) {
      const float4 v = rows[(2 * u + (lane >> 4)) * 16 + (lane & 15)];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    __syncwarp();
    if (lane == 0 && st + NS < stages_per_warp) issue(st + NS);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// the library's way: half-warp per row, 8 rows in flight per half-warp (LDG.128)
__global__ void __launch_bounds__(256) ldg_gather(const float4* __restrict__ table, const int* __restrict__ idx,
                                                  int stages_per_warp, float4* out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int gw = blockIdx.x * nwarps + warp;
  const int* my = idx + (size_t)gw * stages_per_warp * ROWS_PER_STAGE;
  float4 acc = make_float4(0, 0, 0, 0);
  for (int st = 0; st < stages_per_warp; ++st) {
    const int rec = __ldg(my + st * ROWS_PER_STAGE + (lane & 15));
    float4 g[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int r = __shfl_sync(0xffffffffu, rec, 2 * u + (lane >> 4));
      g[u] = __ldg(table + (size_t)r * 16 + (lane & 15));
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) { acc.x += g[u].x; acc.y += g[u].y; acc.z += g[u].z; acc.w += g[u].w; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static double checksum(const float4* d_out, int n) {
  std::vector<float4> h(n);
  cudaMemcpy(h.data(), d_out, (size_t)n * 16, cudaMemcpyDeviceToHost);
  double s = 0;
  for (auto& v : h) s += (double)v.x + v.y + v.z + v.w;
  return s;
}

int main(int argc, char** argv) {
  const int rows = argc > 1 ? atoi(argv[1]) : 640000;
  const int box_rows = argc > 2 ? atoi(argv[2]) : 1;
  const int total = 2490368;      // ~ the kept points of a B=16 step, a multiple of 16 * 148 * 8 * ...
  float* table;
  cudaMalloc(&table, (size_t)rows * ROW_BYTES);
  {
    std::vector<float> h((size_t)rows * ROW_FLOATS);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)((i * 2654435761u) >> 20 & 1023) * (1.0f / 1024);
    cudaMemcpy(table, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  }
  std::vector<int> hi(total + 4096);
  srand(1);
  for (auto& v : hi) v = (int)(((unsigned)rand() * 32768u + (unsigned)rand()) % (unsigned)rows);
  int* idx;
  cudaMalloc(&idx, hi.size() * 4);
  cudaMemcpy(idx, hi.data(), hi.size() * 4, cudaMemcpyHostToDevice);
  float4* out;
  cudaMalloc(&out, (size_t)148 * 64 * 256 * 16);

  EncodeTiled encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres) != cudaSuccess || !encode) {
    printf("cuTensorMapEncodeTiled not available\n");
    return 1;
  }
  CUtensorMap tm;
  const cuuint64_t gdim[2] = {ROW_FLOATS, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {ROW_BYTES};
  const cuuint32_t box[2] = {ROW_FLOATS, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult cr = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, table, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("rows %d (%.0f MB), box {%d, %d}, encode rc %d\n", rows, rows * 256.0 / 1e6, ROW_FLOATS, box_rows, (int)cr);
  if (cr != CUDA_SUCCESS) return 1;

  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  auto flush = [&]() { cudaMemset(out, 0, (size_t)148 * 64 * 256 * 16); };
  double ref_sum = 0;
  // LDG reference at 8/16/24/32 warps per SM
  for (int wps : {16, 24, 32, 48}) {
    const int threads = 256, blocks = 148 * wps / 8, warps = blocks * 8;
    const int spw = total / ROWS_PER_STAGE / warps;
    ldg_gather<<<blocks, threads>>>((const float4*)table, idx, spw, out);
    cudaDeviceSynchronize();
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
      flush();
      cudaEventRecord(e0);
      ldg_gather<<<blocks, threads>>>((const float4*)table, idx, spw, out);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      best = ms < best ? ms : best;
    }
    const double bytes = (double)warps * spw * ROWS_PER_STAGE * ROW_BYTES;
    printf("ldg   %2d warps/SM                          : %7.1f us  %6.2f TB/s  (%.1f G rows/s)  err %s\n", wps, best * 1e3,
           bytes / best / 1e9, bytes / 256 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    if (wps == 16) ref_sum = checksum(out, blocks * threads);
  }
  // TMA gather4: CTAs of 8 warps, NS stages of 4 KB per warp
#define RUN_TMA(NS, CTAS)                                                                                              \
  {                                                                                                                    \
    const int threads = 256, blocks = 148 * CTAS, warps = blocks * 8;                                                  \
    const size_t smem = (size_t)8 * NS * ROWS_PER_STAGE * ROW_BYTES + 8 * NS * 8;                                      \
    cudaFuncSetAttribute(tma_gather<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                      \
    const int spw = total / ROWS_PER_STAGE / warps;                                                                    \
    tma_gather<NS><<<blocks, threads, smem>>>(tm, idx, spw, out);                                                      \
    cudaError_t err = cudaDeviceSynchronize();                                                                         \
    float best = 1e9f;                                                                                                 \
    for (int rep = 0; rep < 5 && err == cudaSuccess; ++rep) {                                                          \
      flush();                                                                                                         \
      cudaEventRecord(e0);                                                                                             \
      tma_gather<NS><<<blocks, threads, smem>>>(tm, idx, spw, out);                                                    \
      cudaEventRecord(e1);                                                                                             \
      cudaEventSynchronize(e1);                                                                                        \
      float ms;                                                                                                        \
      cudaEventElapsedTime(&ms, e0, e1);                                                                               \
      best = ms < best ? ms : best;                                                                                    \
    }                                                                                                                  \
    const double bytes = (double)warps * spw * ROWS_PER_STAGE * ROW_BYTES;                                             \
    printf("tma4  %d CTAs/SM x 8 warps, %d stages (%3zu KB/CTA, %3zu KB in flight/SM): %7.1f us  %6.2f TB/s  (%.1f G rows/s)  err %s", \
           CTAS, NS, smem / 1024, (size_t)CTAS * 8 * NS * 4, best * 1e3, bytes / best / 1e9, bytes / 256 / best / 1e6,  \
           cudaGetErrorString(err));                                                                                   \
    if (CTAS == 2 && NS == 2 && err == cudaSuccess) printf("  checksum %s (%.6e vs ldg %.6e)",                          \
           fabs(checksum(out, blocks * threads) - ref_sum) <= 1e-6 * fabs(ref_sum) ? "OK" : "MISMATCH",                \
           checksum(out, blocks * threads), ref_sum);                                                                  \
    printf("\n");                                                                                                      \
    if (err != cudaSuccess) return 1;                                                                                  \
  }
  RUN_TMA(2, 2)
  RUN_TMA(2, 1)
  RUN_TMA(4, 1)
  RUN_TMA(6, 1)
  RUN_TMA(3, 2)
  RUN_TMA(2, 3)
  RUN_TMA(1, 6)
  RUN_TMA(4, 1)
  return 0;
}
