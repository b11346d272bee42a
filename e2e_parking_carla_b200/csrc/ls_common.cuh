// Shared device/host helpers for the lift-splat kernels (sm_100a).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ls_b200.h"

#define LS_TX 8                      // BEV tile: 8 x-rows ...
#define LS_TY 32                     // ... by 32 y-columns (128 B of fp32 along Y)
#define LS_TILE (LS_TX * LS_TY)      // 256 cells per tile
#define LS_TILE_PAD (LS_TILE + 1)    // odd smem row stride -> conflict-free column walks
#define LS_CCHUNK 64                 // channels handled per pass by one warp (float2 per lane)
#define LS_WARPS 8
#define LS_THREADS (LS_WARPS * 32)

// Everything a kernel needs to know about the BEV grid, derived once on the host.
struct LsGrid {
  int X, Y, Z;
  int tiles_x, tiles_y, tiles;
  int Vc;          // padded cell count = tiles * LS_TILE
  float off[3];    // bev_start_pos - bev_res / 2   (float32 ops, model/bev_model.py:85)
  float res[3];
  float fdim[3];   // (float)dim, exact (dim < 2^24)
};

struct LsDims {
  int B, N, D, fh, fw, C;
  int HW;      // fh*fw
  int DHW;     // D*fh*fw   points per camera
  int Npts;    // N*D*fh*fw points per sample
};

static inline LsDims ls_dims(const LsShape* s) {
  LsDims d;
  d.B = s->B; d.N = s->N; d.D = s->D; d.fh = s->fh; d.fw = s->fw; d.C = s->C;
  d.HW = s->fh * s->fw;
  d.DHW = s->D * d.HW;
  d.Npts = s->N * d.DHW;
  return d;
}

static inline LsGrid ls_grid(const LsShape* s) {
  LsGrid g;
  g.X = s->X; g.Y = s->Y; g.Z = s->Z;
  g.tiles_x = (s->X + LS_TX - 1) / LS_TX;
  g.tiles_y = (s->Y + LS_TY - 1) / LS_TY;
  g.tiles = g.tiles_x * g.tiles_y;
  g.Vc = g.tiles * LS_TILE;
  for (int i = 0; i < 3; ++i) {
    // x86-64 float arithmetic is IEEE single (SSE), same bits as torch's float32 ops.
    volatile float half = s->res[i] / 2.0f;
    volatile float off = s->start[i] - half;
    g.off[i] = off;
    g.res[i] = s->res[i];
  }
  g.fdim[0] = (float)s->X; g.fdim[1] = (float)s->Y; g.fdim[2] = (float)s->Z;
  return g;
}

// rank (reference convention, Z == 1) -> tile-major cell id
__device__ __forceinline__ int ls_cell_of_xy(int gx, int gy, int tiles_y) {
  const int tile = (gx / LS_TX) * tiles_y + (gy / LS_TY);
  return tile * LS_TILE + (gx % LS_TX) * LS_TY + (gy % LS_TY);
}
__device__ __forceinline__ int ls_cell_of_rank(int r, int Y, int tiles_y) {
  const int gx = r / Y;
  const int gy = r - gx * Y;
  return ls_cell_of_xy(gx, gy, tiles_y);
}

// Voxel coordinate of one frustum point, operation by operation as torch-CPU does it
// (model/bev_model.py:50-55,85): p=(u*d, v*d, d); g_i=((0+m_i0*px)+m_i1*py)+m_i2*pz; g_i+=t_i;
// c_i=(g_i-off_i)/res_i.  No FMA contraction, IEEE divide.
__device__ __forceinline__ void ls_point_geom(const float* __restrict__ m, const float* __restrict__ t,
                                              float u, float v, float d, float g[3]) {
  const float px = __fmul_rn(u, d);
  const float py = __fmul_rn(v, d);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float a = __fadd_rn(0.0f, __fmul_rn(m[3 * i + 0], px));
    a = __fadd_rn(a, __fmul_rn(m[3 * i + 1], py));
    a = __fadd_rn(a, __fmul_rn(m[3 * i + 2], d));
    g[i] = __fadd_rn(a, t[i]);
  }
}

__device__ __forceinline__ bool ls_point_voxel(const float g[3], const LsGrid& grid, float c[3], int v[3]) {
  bool keep = true;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    c[i] = __fdiv_rn(__fsub_rn(g[i], grid.off[i]), grid.res[i]);
    // trunc(c) in [0, dim)  <=>  -1 < c < dim   (NaN fails both, like x86's INT64_MIN)
    keep = keep && (c[i] > -1.0f) && (c[i] < grid.fdim[i]);
    v[i] = __float2int_rz(c[i]);
  }
  return keep;
}

template <typename T> __device__ __forceinline__ float ls_to_float(T v);
template <> __device__ __forceinline__ float ls_to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float ls_to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T ls_from_float(float v);
template <> __device__ __forceinline__ float ls_from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 ls_from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// two consecutive channels
template <typename T> __device__ __forceinline__ float2 ls_load2(const T* p);
template <> __device__ __forceinline__ float2 ls_load2<float>(const float* p) {
  return *reinterpret_cast<const float2*>(p);
}
template <> __device__ __forceinline__ float2 ls_load2<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}
template <typename T> __device__ __forceinline__ void ls_store2(T* p, float2 v);
template <> __device__ __forceinline__ void ls_store2<float>(float* p, float2 v) {
  *reinterpret_cast<float2*>(p) = v;
}
template <> __device__ __forceinline__ void ls_store2<__nv_bfloat16>(__nv_bfloat16* p, float2 v) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __float22bfloat162_rn(v);
}

__device__ __forceinline__ float ls_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
