// Dense helpers around the splat: depth softmax (model/bev_model.py:64) forward/backward and
// the NCHW <-> padded-NHWC staging transposes.  All bandwidth-bound, coalesced both ways.
#include <stdlib.h>

#include "ls_internal.h"

#define LS_SM_MAXK 16   // depth bins per thread held in registers: D <= 8 * 16

// =====================================================================================
// softmax over depth.  CTA = (image, 32 pixels); thread (lane = pixel, dg = depth group 0..7)
// owns bins dg, dg+8, ... in registers; max and sum are combined through shared memory.
// =====================================================================================
template <typename T, int K>
__global__ void __launch_bounds__(256)
ls_softmax_kernel(const T* __restrict__ logits, int D, int HW, T* __restrict__ prob) {
  ls_pdl_trigger();
  ls_pdl_wait();
  __shared__ float red[8][33];
  const int img = blockIdx.y;
  const int lane = threadIdx.x & 31, dg = threadIdx.x >> 5;
  const int rc = blockIdx.x * 32 + lane;
  const bool on = rc < HW;
  const T* src = logits + (size_t)img * D * HW + rc;
  T* dst = prob + (size_t)img * D * HW + rc;
  float x[K];
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int d = dg + 8 * k;
    x[k] = (on && d < D) ? ls_to_float(src[(size_t)d * HW]) : -INFINITY;
    m = fmaxf(m, x[k]);
  }
  red[dg][lane] = m;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) m = fmaxf(m, red[j][lane]);
  __syncthreads();
  float s = 0.0f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int d = dg + 8 * k;
    x[k] = (on && d < D) ? expf(x[k] - m) : 0.0f;
    s += x[k];
  }
  red[dg][lane] = s;
  __syncthreads();
  s = 0.0f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += red[j][lane];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int d = dg + 8 * k;
    if (on && d < D) dst[(size_t)d * HW] = ls_from_float<T>(__fdiv_rn(x[k], s));
  }
}

// generic fallback for D > 128: one thread per pixel, three passes
template <typename T>
__global__ void __launch_bounds__(256)
ls_softmax_serial_kernel(const T* __restrict__ logits, int D, int HW, T* __restrict__ prob) {
  ls_pdl_trigger();
  ls_pdl_wait();
  const int img = blockIdx.y;
  const int rc = blockIdx.x * blockDim.x + threadIdx.x;
  if (rc >= HW) return;
  const T* src = logits + (size_t)img * D * HW + rc;
  T* dst = prob + (size_t)img * D * HW + rc;
  float m = -INFINITY;
  for (int d = 0; d < D; ++d) m = fmaxf(m, ls_to_float(src[(size_t)d * HW]));
  float s = 0.0f;
  for (int d = 0; d < D; ++d) s += expf(ls_to_float(src[(size_t)d * HW]) - m);
  for (int d = 0; d < D; ++d)
    dst[(size_t)d * HW] = ls_from_float<T>(__fdiv_rn(expf(ls_to_float(src[(size_t)d * HW]) - m), s));
}

template <typename T>
static int ls_softmax_dispatch(const T* logits, const LsDims& dm, T* prob, dim3 grid, cudaStream_t s) {
  const int k = (dm.D + 7) / 8;
#define LS_SM(KK) LS_LAUNCH((ls_softmax_kernel<T, KK>), grid, dim3(256), 0, s, logits, dm.D, dm.HW, prob)
  if (k <= 2) LS_SM(2);
  else if (k <= 4) LS_SM(4);
  else if (k <= 6) LS_SM(6);
  else if (k <= 8) LS_SM(8);
  else if (k <= 12) LS_SM(12);
  else LS_SM(16);
#undef LS_SM
  return LS_OK;
}

int ls_launch_softmax(const void* logits, int dtype, const LsDims& dm, void* prob, cudaStream_t s) {
  const int images = dm.B * dm.N;
  if (dm.D <= 8 * LS_SM_MAXK) {
    dim3 grid((dm.HW + 31) / 32, images);
    if (dtype == LS_F32) return ls_softmax_dispatch<float>((const float*)logits, dm, (float*)prob, grid, s);
    return ls_softmax_dispatch<__nv_bfloat16>((const __nv_bfloat16*)logits, dm, (__nv_bfloat16*)prob, grid, s);
  }
  dim3 grid((dm.HW + 255) / 256, images);
  if (dtype == LS_F32)
    LS_LAUNCH(ls_softmax_serial_kernel<float>, grid, dim3(256), 0, s, (const float*)logits, dm.D, dm.HW, (float*)prob);
  else
    LS_LAUNCH(ls_softmax_serial_kernel<__nv_bfloat16>, grid, dim3(256), 0, s, (const __nv_bfloat16*)logits, dm.D,
              dm.HW, (__nv_bfloat16*)prob);
  return LS_OK;
}

// =====================================================================================
// softmax backward: gl = prob * (g - sum_d prob*g),  g = grad_prob (+ ext).
// grad_prob arrives PIXEL-major [image][pixel][D] from the gather kernel; it is staged through
// shared memory so that both its reads and the depth-major writes are coalesced.
// =====================================================================================
template <typename T, int K>
__device__ __forceinline__ void ls_softmax_bwd_tile(const T* __restrict__ prob, const float* __restrict__ gprob_pm,
                                                    const T* __restrict__ gext, int D, int HW, T* __restrict__ glogits,
                                                    int img, int rc0, float* sm) {
  const int Dp = D | 1;
  float* stage = sm;                   // [32][Dp]
  float* red = sm + 32 * Dp;           // [8][33]
  const int lane = threadIdx.x & 31, dg = threadIdx.x >> 5;
  const int valid = min(32, HW - rc0);
  const int rc = rc0 + lane;
  const bool on = rc < HW;
  const size_t base = (size_t)img * D * HW + rc;
  // depth-major operands first (independent loads, all in flight), pixel-major rows meanwhile
  float p[K], g[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int d = dg + 8 * k;
    p[k] = 0.0f; g[k] = 0.0f;
    if (on && d < D) {
      p[k] = ls_to_float(prob[base + (size_t)d * HW]);
      if (gext) g[k] = ls_to_float(gext[base + (size_t)d * HW]);
    }
  }
  const float* gsrc = gprob_pm + ((size_t)img * HW + rc0) * D;
  for (int r = dg; r < valid; r += 8)
    for (int d = lane; d < D; d += 32) stage[r * Dp + d] = gsrc[(size_t)r * D + d];
  __syncthreads();
  float dot = 0.0f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int d = dg + 8 * k;
    if (on && d < D) { g[k] += stage[lane * Dp + d]; dot = fmaf(p[k], g[k], dot); }
  }
  red[dg * 33 + lane] = dot;
  __syncthreads();
  dot = 0.0f;
#pragma unroll
  for (int j = 0; j < 8; ++j) dot += red[j * 33 + lane];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int d = dg + 8 * k;
    if (on && d < D) glogits[base + (size_t)d * HW] = ls_from_float<T>(p[k] * (g[k] - dot));
  }
}

template <typename T, int K>
__global__ void __launch_bounds__(256)
ls_softmax_bwd_kernel(const T* __restrict__ prob, const float* __restrict__ gprob_pm, const T* __restrict__ gext,
                      int D, int HW, T* __restrict__ glogits) {
  ls_pdl_trigger();
  ls_pdl_wait();
  extern __shared__ float sm[];
  ls_softmax_bwd_tile<T, K>(prob, gprob_pm, gext, D, HW, glogits, blockIdx.y, blockIdx.x * 32, sm);
}

// generic fallback for D > 128
template <typename T>
__global__ void __launch_bounds__(256)
ls_softmax_bwd_serial_kernel(const T* __restrict__ prob, const float* __restrict__ gprob_pm, const T* __restrict__ gext,
                             int D, int HW, T* __restrict__ glogits) {
  ls_pdl_trigger();
  ls_pdl_wait();
  const int img = blockIdx.y;
  const int rc = blockIdx.x * blockDim.x + threadIdx.x;
  if (rc >= HW) return;
  const size_t base = (size_t)img * D * HW + rc;
  const float* gsrc = gprob_pm + ((size_t)img * HW + rc) * D;
  float dot = 0.0f;
  for (int d = 0; d < D; ++d) {
    float g = gsrc[d];
    if (gext) g += ls_to_float(gext[base + (size_t)d * HW]);
    dot = fmaf(ls_to_float(prob[base + (size_t)d * HW]), g, dot);
  }
  for (int d = 0; d < D; ++d) {
    float g = gsrc[d];
    if (gext) g += ls_to_float(gext[base + (size_t)d * HW]);
    glogits[base + (size_t)d * HW] = ls_from_float<T>(ls_to_float(prob[base + (size_t)d * HW]) * (g - dot));
  }
}

template <typename T>
static int ls_softmax_bwd_dispatch(const T* prob, const float* gprob_pm, const T* gext, const LsDims& dm, T* glogits,
                                    dim3 grid, size_t smem, cudaStream_t s) {
  const int k = (dm.D + 7) / 8;
#define LS_SB(KK) \
  LS_LAUNCH((ls_softmax_bwd_kernel<T, KK>), grid, dim3(256), smem, s, prob, gprob_pm, gext, dm.D, dm.HW, glogits)
  if (k <= 2) LS_SB(2);
  else if (k <= 4) LS_SB(4);
  else if (k <= 6) LS_SB(6);
  else if (k <= 8) LS_SB(8);
  else if (k <= 12) LS_SB(12);
  else LS_SB(16);
#undef LS_SB
  return LS_OK;
}

int ls_launch_bwd_epilogue(const void* prob, const float* gprob_pm, const void* gext, const void* gfeatT, int dtype,
                           const LsDims& dm, void* glogits, void* gfeat_nchw, int* ready, int target, cudaStream_t s);
bool ls_epilogue_supports(const LsDims& dm);

int ls_launch_softmax_bwd(const void* prob, const float* gprob_pm, const void* gext, int dtype, const LsDims& dm,
                          void* glogits, cudaStream_t s) {
  // the common depth ranges: thread-per-pixel kernel (no staging, same bits as the staged one below)
  if (ls_epilogue_supports(dm))
    return ls_launch_bwd_epilogue(prob, gprob_pm, gext, nullptr, dtype, dm, glogits, nullptr, nullptr, 0, s);
  const size_t smem = ((size_t)32 * (dm.D | 1) + 8 * 33) * sizeof(float);
  if (dm.D <= 8 * LS_SM_MAXK && smem <= 48 * 1024) {
    dim3 grid((dm.HW + 31) / 32, dm.B * dm.N);
    if (dtype == LS_F32)
      return ls_softmax_bwd_dispatch<float>((const float*)prob, gprob_pm, (const float*)gext, dm, (float*)glogits, grid,
                                            smem, s);
    return ls_softmax_bwd_dispatch<__nv_bfloat16>((const __nv_bfloat16*)prob, gprob_pm, (const __nv_bfloat16*)gext, dm,
                                                  (__nv_bfloat16*)glogits, grid, smem, s);
  }
  dim3 grid((dm.HW + 255) / 256, dm.B * dm.N);
  if (dtype == LS_F32)
    LS_LAUNCH(ls_softmax_bwd_serial_kernel<float>, grid, dim3(256), 0, s, (const float*)prob, gprob_pm,
              (const float*)gext, dm.D, dm.HW, (float*)glogits);
  else
    LS_LAUNCH(ls_softmax_bwd_serial_kernel<__nv_bfloat16>, grid, dim3(256), 0, s, (const __nv_bfloat16*)prob, gprob_pm,
              (const __nv_bfloat16*)gext, dm.D, dm.HW, (__nv_bfloat16*)glogits);
  return LS_OK;
}

// =====================================================================================
// [image][C][HW] -> [image][HW][Cp]  (Cp = C rounded up to 4, zero padded), 64x64 tiles.
// Reads are 128-byte rows over HW; writes are 16-byte channel quads (256 B per pixel at C=64).
// =====================================================================================
template <typename T>
__global__ void __launch_bounds__(256)
ls_to_nhwc_kernel(const T* __restrict__ src, int C, int Cp, int HW, T* __restrict__ dst) {
  ls_pdl_trigger();
  ls_pdl_wait();
  __shared__ float tile[64][65];
  const int img = blockIdx.z, hw0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  const T* in = src + (size_t)img * C * HW;
  T* out = dst + (size_t)img * HW * Cp;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int c = w; c < 64; c += 8) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int hw = hw0 + lane + 32 * h;
      float v = 0.0f;
      if (c0 + c < C && hw < HW) v = ls_to_float(in[(size_t)(c0 + c) * HW + hw]);
      tile[c][lane + 32 * h] = v;
    }
  }
  __syncthreads();
  const int q = threadIdx.x & 15;
  for (int h = threadIdx.x >> 4; h < 64; h += 16) {
    const int hw = hw0 + h, c = c0 + 4 * q;
    if (hw < HW && c < Cp)
      ls_store4<T>(out + (size_t)hw * Cp + c,
                   make_float4(tile[4 * q][h], tile[4 * q + 1][h], tile[4 * q + 2][h], tile[4 * q + 3][h]));
  }
}

template <typename T>
__device__ __forceinline__ void ls_from_nhwc_tile(const T* __restrict__ src, int C, int Cp, int HW, T* __restrict__ dst,
                                                  int img, int hw0, int c0, float (*tile)[65]) {
  const T* in = src + (size_t)img * HW * Cp;
  T* out = dst + (size_t)img * C * HW;
  const int q = threadIdx.x & 15;
  for (int h = threadIdx.x >> 4; h < 64; h += 16) {
    const int hw = hw0 + h, c = c0 + 4 * q;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (hw < HW && c < Cp) v = ls_load4<T>(in + (size_t)hw * Cp + c);
    tile[4 * q][h] = v.x; tile[4 * q + 1][h] = v.y; tile[4 * q + 2][h] = v.z; tile[4 * q + 3][h] = v.w;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int c = w; c < 64; c += 8) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int hw = hw0 + lane + 32 * h;
      if (c0 + c < C && hw < HW) out[(size_t)(c0 + c) * HW + hw] = ls_from_float<T>(tile[c][lane + 32 * h]);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
ls_from_nhwc_kernel(const T* __restrict__ src, int C, int Cp, int HW, T* __restrict__ dst) {
  ls_pdl_trigger();
  ls_pdl_wait();
  __shared__ float tile[64][65];
  ls_from_nhwc_tile<T>(src, C, Cp, HW, dst, blockIdx.z, blockIdx.x * 64, blockIdx.y * 64, tile);
}

int ls_launch_to_nhwc(const void* src, int dtype, int images, int C, int Cp, int HW, void* dst, cudaStream_t s) {
  dim3 grid((HW + 63) / 64, (Cp + 63) / 64, images);
  if (dtype == LS_F32)
    LS_LAUNCH(ls_to_nhwc_kernel<float>, grid, dim3(256), 0, s, (const float*)src, C, Cp, HW, (float*)dst);
  else
    LS_LAUNCH(ls_to_nhwc_kernel<__nv_bfloat16>, grid, dim3(256), 0, s, (const __nv_bfloat16*)src, C, Cp, HW,
              (__nv_bfloat16*)dst);
  return LS_OK;
}

int ls_launch_from_nhwc(const void* src, int dtype, int images, int C, int Cp, int HW, void* dst, cudaStream_t s) {
  dim3 grid((HW + 63) / 64, (Cp + 63) / 64, images);
  if (dtype == LS_F32)
    LS_LAUNCH(ls_from_nhwc_kernel<float>, grid, dim3(256), 0, s, (const float*)src, C, Cp, HW, (float*)dst);
  else
    LS_LAUNCH(ls_from_nhwc_kernel<__nv_bfloat16>, grid, dim3(256), 0, s, (const __nv_bfloat16*)src, C, Cp, HW,
              (__nv_bfloat16*)dst);
  return LS_OK;
}

// =====================================================================================
// Backward epilogue, pixel-stationary: ONE launch does the softmax backward and (gfeatT != NULL) the
// NHWC -> NCHW fix-up of grad_feat, with a THREAD per pixel and no shared memory:
//   part A  a thread reads its pixel's kD pixel-major gradients (16-byte loads of its own row), the kD
//           depth-major probabilities / upstream gradients (coalesced over the warp's 32 pixels), and
//           writes kD depth-major logit gradients; the sum over depth is thread-local, added up in the
//           order of ls_softmax_bwd_tile (8 strided chains, then 0..7), so both kernels give the same bits;
//   part B  a thread reads its pixel's row of grad_feat (16-byte loads) and writes one element of
//           every channel plane (coalesced over the warp).
// ~8 instructions per element instead of ~80 for the staged version (index math, shared-memory
// staging and two barriers per 32 pixels): the kernel is bound by its 110 MB of traffic, not by issue.
// kOverlap: launched as the programmatic dependent of ls_bwd_gather_occ_kernel (ready != NULL) WITHOUT
// a dependency wait: its CTAs take the SM slots the gather's last wave leaves free, in the gather's
// own image order, and each waits (acquire) until ready[image] has `target` arrivals - the image's
// rows are then complete.  Departures are counted per image; the last CTA of an image zeroes both
// counters for the next call.  No CTA blocks on the dependency at all: one that did would hold its
// slot until the whole gather has drained.  The grid still cannot complete before the gather's last
// warp has signalled, because every image has CTAs here.
// =====================================================================================
#ifndef LS_EPI_THREADS
#define LS_EPI_THREADS 128
#endif
template <typename T, int kD>
__device__ __forceinline__ void ls_softmax_bwd_pixel(const T* prob, const float* gprob_pm, const T* gext, int HW,
                                                     T* glogits, int img, int hw) {
  static_assert(kD % 8 == 0, "depth bins in groups of 8");
  float g[kD], p[kD];
  const float4* gs = reinterpret_cast<const float4*>(gprob_pm + ((size_t)img * HW + hw) * kD);
#pragma unroll
  for (int i = 0; i < kD / 4; ++i) {
    const float4 v = gs[i];
    g[4 * i] = v.x; g[4 * i + 1] = v.y; g[4 * i + 2] = v.z; g[4 * i + 3] = v.w;
  }
  const size_t base = (size_t)img * kD * HW + hw;
#pragma unroll
  for (int d = 0; d < kD; ++d) p[d] = ls_to_float(prob[base + (size_t)d * HW]);
  if (gext) {
#pragma unroll
    for (int d = 0; d < kD; ++d) g[d] = ls_to_float(gext[base + (size_t)d * HW]) + g[d];
  }
  // sum_d p g in the staged kernel's order: chain j takes d = j, j + 8, ...; chains added 0..7 from zero
  float dot = 0.0f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float c = 0.0f;
#pragma unroll
    for (int d = j; d < kD; d += 8) c = fmaf(p[d], g[d], c);
    dot += c;
  }
#pragma unroll
  for (int d = 0; d < kD; ++d) glogits[base + (size_t)d * HW] = ls_from_float<T>(p[d] * (g[d] - dot));
}

template <typename T> __device__ __forceinline__ float4 ls_load4_plain(const T* p);
template <> __device__ __forceinline__ float4 ls_load4_plain<float>(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
template <> __device__ __forceinline__ float4 ls_load4_plain<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 raw = *reinterpret_cast<const uint2*>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

template <typename T>
__device__ __forceinline__ void ls_from_nhwc_pixel(const T* src, int C, int Cp, int HW, T* dst, int img, int hw) {
  const T* in = src + ((size_t)img * HW + hw) * Cp;
  T* out = dst + (size_t)img * C * HW + hw;
  for (int c0 = 0; c0 < Cp; c0 += 32) {          // 8 quads in flight
    float4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (c0 + 4 * i < Cp) ? ls_load4_plain<T>(in + c0 + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = c0 + 4 * i;
      if (c + 0 < C) out[(size_t)(c + 0) * HW] = ls_from_float<T>(v[i].x);
      if (c + 1 < C) out[(size_t)(c + 1) * HW] = ls_from_float<T>(v[i].y);
      if (c + 2 < C) out[(size_t)(c + 2) * HW] = ls_from_float<T>(v[i].z);
      if (c + 3 < C) out[(size_t)(c + 3) * HW] = ls_from_float<T>(v[i].w);
    }
  }
}

// (no __restrict__ / read-only path on the inputs: with kOverlap they are written while this grid is resident)
template <typename T, int kD, bool kOverlap>
__global__ void __launch_bounds__(LS_EPI_THREADS)
ls_bwd_epilogue_kernel(const T* prob, const float* gprob_pm, const T* gext, const T* gfeatT, int HW, int C, int Cp,
                       T* glogits, T* gfeat, int* ready, int target, int images) {
  const int nA = (HW + LS_EPI_THREADS - 1) / LS_EPI_THREADS, nB = gfeatT ? nA : 0;
  const int per = nA + nB;
  const int order = blockIdx.x / per, j = blockIdx.x % per;
  const int img = LS_GATHER_REVERSE ? images - 1 - order : order;
  if (kOverlap) {
    if (threadIdx.x == 0) {
      int seen;
      for (;;) {
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(ready + img) : "memory");
        if (seen >= target) break;
        __nanosleep(200);
      }
    }
    __syncthreads();
  } else {
    ls_pdl_trigger();
    ls_pdl_wait();
  }
  // blocks alternate between the two parts, so both streams are in flight on every SM
  const bool partA = nB == 0 || (j & 1) == 0;
  const int hw = (nB == 0 ? j : j >> 1) * LS_EPI_THREADS + threadIdx.x;
  if (hw < HW) {
    if (partA) ls_softmax_bwd_pixel<T, kD>(prob, gprob_pm, gext, HW, glogits, img, hw);
    else ls_from_nhwc_pixel<T>(gfeatT, C, Cp, HW, gfeat, img, hw);
  }
  if (kOverlap) {
    if (threadIdx.x == 0 && atomicAdd(ready + images + img, 1) == per - 1) {
      ready[img] = 0;                  // every CTA of the image is past its wait: clean for the next call
      ready[images + img] = 0;
    }
  }
}

// depth-bin counts with a pixel-stationary instantiation (registers: 2 kD floats per thread)
// (LS_SOFTMAX_BWD_STAGED=1: never - the staged kernels everywhere, for A/B runs and the same-bits test)
bool ls_epilogue_supports(const LsDims& dm) {
  static const bool staged_only = getenv("LS_SOFTMAX_BWD_STAGED") != nullptr;
  return !staged_only && (dm.D == 48 || dm.D == 32 || dm.D == 64);
}

template <typename T>
static int ls_bwd_epilogue_dispatch(const T* prob, const float* gprob_pm, const T* gext, const T* gfeatT, const LsDims& dm,
                                    T* glogits, T* gfeat, int* ready, int target, cudaStream_t s) {
  const int images = dm.B * dm.N;
  const int nA = (dm.HW + LS_EPI_THREADS - 1) / LS_EPI_THREADS;
  const dim3 grid(images * (gfeatT ? 2 * nA : nA));
#define LS_EP(DD)                                                                                                       \
  do {                                                                                                                  \
    if (ready)                                                                                                          \
      LS_LAUNCH((ls_bwd_epilogue_kernel<T, DD, true>), grid, dim3(LS_EPI_THREADS), 0, s, prob, gprob_pm, gext, gfeatT, dm.HW, \
                dm.C, dm.Cp, glogits, gfeat, ready, target, images);                                                   \
    else                                                                                                                \
      LS_LAUNCH((ls_bwd_epilogue_kernel<T, DD, false>), grid, dim3(LS_EPI_THREADS), 0, s, prob, gprob_pm, gext, gfeatT, dm.HW, \
                dm.C, dm.Cp, glogits, gfeat, ready, target, images);                                                   \
  } while (0)
  if (dm.D == 48) LS_EP(48);
  else if (dm.D == 32) LS_EP(32);
  else if (dm.D == 64) LS_EP(64);
  else return LS_ERR_UNSUPPORTED;
#undef LS_EP
  return LS_OK;
}

// ready == NULL: plain stream-ordered launch (dependency wait at the top)
int ls_launch_bwd_epilogue(const void* prob, const float* gprob_pm, const void* gext, const void* gfeatT, int dtype,
                           const LsDims& dm, void* glogits, void* gfeat_nchw, int* ready, int target, cudaStream_t s) {
  if (dtype == LS_F32)
    return ls_bwd_epilogue_dispatch<float>((const float*)prob, gprob_pm, (const float*)gext, (const float*)gfeatT, dm,
                                           (float*)glogits, (float*)gfeat_nchw, ready, target, s);
  return ls_bwd_epilogue_dispatch<__nv_bfloat16>((const __nv_bfloat16*)prob, gprob_pm, (const __nv_bfloat16*)gext,
                                                 (const __nv_bfloat16*)gfeatT, dm, (__nv_bfloat16*)glogits,
                                                 (__nv_bfloat16*)gfeat_nchw, ready, target, s);
}

// =====================================================================================
// Target channel of ParkingModel.add_target_bev (model/parking_model.py:28-46): an 8x8 stamp
// of ones around the (noised) target pixel, zeros elsewhere, written straight into its channel
// of the BEV tensor (any strides: a [B,1,X,Y] tensor of its own, or channel C of a
// channels-last [B,X,Y,C+1] buffer whose first C channels the splat wrote) - no torch.zeros,
// no per-sample python loop, no torch.cat copy of the 10 MB/sample BEV.
// The slice bounds follow python's semantics exactly (negative starts wrap, then clamp).
// =====================================================================================
__device__ __forceinline__ int ls_py_slice_bound(int v, int n) {
  if (v < 0) v += n;
  return v < 0 ? 0 : (v > n ? n : v);
}

__global__ void __launch_bounds__(256)
ls_target_bev_kernel(const int* __restrict__ pix, int X, int Y, float* __restrict__ out, long long sb, long long sx,
                     long long sy) {
  ls_pdl_trigger();
  ls_pdl_wait();
  const int b = blockIdx.y;
  const int px = pix[2 * b + 0], py = pix[2 * b + 1];
  const int x0 = ls_py_slice_bound(px - 4, X), x1 = ls_py_slice_bound(px + 4, X);
  const int y0 = ls_py_slice_bound(py - 4, Y), y1 = ls_py_slice_bound(py + 4, Y);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < X * Y; i += gridDim.x * blockDim.x) {
    const int x = i / Y, y = i % Y;
    const bool in = x >= x0 && x < x1 && y >= y0 && y < y1;
    out[(size_t)b * sb + (size_t)x * sx + (size_t)y * sy] = in ? 1.0f : 0.0f;
  }
}

int ls_launch_target_bev(const int* pix, int B, int X, int Y, float* out, long long sb, long long sx, long long sy,
                         cudaStream_t s) {
  const int blocks = (X * Y + 255) / 256;
  LS_LAUNCH(ls_target_bev_kernel, dim3(blocks < 1 ? 1 : (blocks > 592 ? 592 : blocks), B), dim3(256), 0, s, pix, X, Y,
            out, sb, sx, sy);
  return LS_OK;
}
