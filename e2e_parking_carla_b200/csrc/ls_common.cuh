// Shared device/host helpers for the lift-splat kernels (sm_100a).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "ls_b200.h"

// ---- BEV tiling ---------------------------------------------------------------------
// The grid is cut into tiles of LS_TILE = 128 cells, tx x ty (LsShape.tile_x; tx a power of two
// dividing 128); cells are numbered tile-major (cell = tile*128 + lx*ty + ly).  A tile is one CTA
// of the canonicaliser and of the splat.
//   1 x 128  long runs of one x-row: what the NCHW write-out / gradient staging need (a tile row is
//            a contiguous 512-byte run of every channel plane) - LS_TX / LS_TY below, compile time;
//   8 x 16   (or any other tx) for the channels-last direct splat, whose output unit is a cell's own
//            256-byte row, so the tile can be square: a ray then stays inside the tile for ~3.6
//            consecutive depth bins instead of 1.15 and its feature row is re-read from L1, not L2.
#ifndef LS_TX
#define LS_TX 1                      // x-rows per tile of the compile-time (NCHW) geometry
#endif
#ifndef LS_TY
#define LS_TY 128                    // y-columns per tile of the compile-time (NCHW) geometry
#endif
#define LS_TILE (LS_TX * LS_TY)      // cells per tile (cell-in-tile fits 8 bits)
static_assert(LS_TX == 1 && LS_TY == 128, "the tile kernels of the NCHW path assume 1 x 128 strips");
#define LS_CCHUNK 64                 // channels per pass
#define LS_THREADS LS_TILE           // tile kernels: one thread per cell of the tile
#ifndef LS_GATHER_THREADS
#define LS_GATHER_THREADS 256
#endif
#ifndef LS_GATHER_MINB
#define LS_GATHER_MINB 2
#endif
#define LS_HALFWARPS (LS_GATHER_THREADS / 16)
#ifndef LS_GATHER_REVERSE
#define LS_GATHER_REVERSE 1          // the gradient gather (and its overlapped epilogue) walk the images last to first
#endif
#define LS_SEG_PAD 4                 // seg_start row stride = Vc + LS_SEG_PAD (16 B aligned rows)

// Everything a kernel needs to know about the BEV grid, derived once on the host.
struct LsGrid {
  int X, Y, Z;
  int XY;          // X * Y: cells of one sample in the reference's row-major rank order
  int tx, ty;      // tile shape (tx * ty == LS_TILE), powers of two
  int tx_shift, ty_shift;
  int tiles_x, tiles_y, tiles;
  int Vc;          // padded cell count = tiles * LS_TILE
  int seg_stride;  // Vc + LS_SEG_PAD
  unsigned long long ty_magic;   // ceil(2^42 / tiles_y): tile / tiles_y == (tile * ty_magic) >> 42 for tile < 2^21
  float off[3];    // bev_start_pos - bev_res / 2   (float32 ops, model/bev_model.py:85)
  float res[3];
  float fdim[3];   // (float)dim, exact (dim < 2^24)
  int zfast;       // Z == 1 and res_z > 0: the z keep-test needs no division (see ls_point_voxel)
};

struct LsDims {
  int B, N, D, fh, fw, C;
  int Cp;      // channels padded to a multiple of 4 (internal NHWC staging rows)
  int HW;      // fh*fw
  int DHW;     // D*fh*fw   points per camera
  int Npts;    // N*D*fh*fw points per sample
  int dbits;   // ceil(log2(D)): sort key = cell_in_tile<<24 | (pix << dbits | d)
  int policy;  // LsGeomPolicy
  int bev_bf16; // LsShape.bev_dtype == LS_BF16: BEV tensor and its gradient are bf16 (opt-in)
};

static inline LsDims ls_dims(const LsShape* s) {
  LsDims d;
  d.B = s->B; d.N = s->N; d.D = s->D; d.fh = s->fh; d.fw = s->fw; d.C = s->C;
  d.Cp = (s->C + 3) & ~3;
  d.HW = s->fh * s->fw;
  d.DHW = s->D * d.HW;
  d.Npts = s->N * d.DHW;
  d.dbits = 0;
  while ((1 << d.dbits) < s->D) ++d.dbits;
  d.policy = s->geom_policy;
  d.bev_bf16 = s->bev_dtype == LS_BF16;
  return d;
}

static inline LsGrid ls_grid(const LsShape* s) {
  LsGrid g;
  g.X = s->X; g.Y = s->Y; g.Z = s->Z;
  g.tx = s->tile_x > 0 ? s->tile_x : 1;
  g.ty = LS_TILE / g.tx;
  g.tx_shift = 0;
  while ((1 << g.tx_shift) < g.tx) ++g.tx_shift;
  g.ty_shift = 0;
  while ((1 << g.ty_shift) < g.ty) ++g.ty_shift;
  g.tiles_x = (s->X + g.tx - 1) / g.tx;
  g.tiles_y = (s->Y + g.ty - 1) / g.ty;
  g.tiles = g.tiles_x * g.tiles_y;
  g.Vc = g.tiles * LS_TILE;
  g.seg_stride = g.Vc + LS_SEG_PAD;
  g.XY = s->X * s->Y;
  g.ty_magic = ((1ULL << 42) + (unsigned long long)g.tiles_y - 1) / (unsigned long long)g.tiles_y;
  for (int i = 0; i < 3; ++i) {
    // x86-64 float arithmetic is IEEE single (SSE), same bits as torch's float32 ops.
    volatile float half = s->res[i] / 2.0f;
    volatile float off = s->start[i] - half;
    g.off[i] = off;
    g.res[i] = s->res[i];
  }
  g.fdim[0] = (float)s->X; g.fdim[1] = (float)s->Y; g.fdim[2] = (float)s->Z;
  g.zfast = (s->Z == 1 && s->res[2] > 0.0f && s->res[2] < 3.0e38f) ? 1 : 0;
  return g;
}

// (gx, gy) -> tile-major cell id (Z == 1)
__host__ __device__ __forceinline__ int ls_cell_of_xy(int gx, int gy, const LsGrid& g) {
  const int tile = (gx >> g.tx_shift) * g.tiles_y + (gy >> g.ty_shift);
  return tile * LS_TILE + ((gx & (g.tx - 1)) << g.ty_shift) + (gy & (g.ty - 1));
}
// tile-major cell id -> reference rank gx*Y + gy (-1 for padding cells outside the grid)
__host__ __device__ __forceinline__ int ls_rank_of_cell(int cell, const LsGrid& g) {
  const int tile = cell / LS_TILE, local = cell % LS_TILE;
  const int gx = (tile / g.tiles_y) * g.tx + (local >> g.ty_shift);
  const int gy = (tile % g.tiles_y) * g.ty + (local & (g.ty - 1));
  return (gx < g.X && gy < g.Y) ? gx * g.Y + gy : -1;
}

// the same for a cell known to lie inside the grid, without a hardware division (tile < 2^21 and
// tiles_y < 2^21 are guaranteed by ls_check_splat_shape: (tile * magic) >> 42 is then exact)
__device__ __forceinline__ int ls_rank_of_cell_fast(int cell, const LsGrid& g) {
  const unsigned tile = (unsigned)cell / LS_TILE, local = (unsigned)cell % LS_TILE;
  const unsigned tgx = (unsigned)(((unsigned long long)tile * g.ty_magic) >> 42);
  const unsigned gx = (tgx << g.tx_shift) + (local >> g.ty_shift);
  const unsigned gy = ((tile - tgx * (unsigned)g.tiles_y) << g.ty_shift) + (local & (unsigned)(g.ty - 1));
  return (int)(gx * (unsigned)g.Y + gy);
}

// ---- BEV tensor layouts (include/ls_b200.h LsBevStrides) --------------------------------
enum LsBevOut {
  LS_OUT_NCHW_VEC4 = 0,    // y stride 1, 16-byte aligned runs: 4x4 register transposes + 16-byte stores
  LS_OUT_NCHW_SCALAR = 1,  // y stride 1, anything else
  LS_OUT_NHWC_BULK = 2,    // c stride 1, dense 64-channel rows: the tile leaves as ONE bulk (TMA) store
  LS_OUT_NHWC_ROWS = 3,    // c stride 1, any row pitch / channel count: coalesced 4-byte stores
  LS_OUT_NHWC_DIRECT_VEC = 4,     // c stride 1, 64 channels of a wider 16-byte aligned row: direct row stores
  LS_OUT_NHWC_DIRECT_SCALAR = 5,  // the same with unaligned rows (64 of 65 channels)
  LS_OUT_BAD = -1
};
enum LsGradIn {
  LS_GRAD_STAGED = 0,      // NCHW gradient -> cell rows staged by ls_bwd_transpose_kernel
  LS_GRAD_DIRECT_VEC = 1,  // channels-last gradient, 16-byte aligned rows: gathered in place
  LS_GRAD_DIRECT_SCALAR = 2,   // channels-last gradient, unaligned rows (e.g. a 64-channel slice of 65)
  LS_GRAD_BAD = -1
};

// Voxel coordinate of one frustum point, operation by operation as torch-CPU does it
// (model/bev_model.py:50-55,85): p=(u*d, v*d, d); g_i=((0+m_i0*px)+m_i1*py)+m_i2*pz; g_i+=t_i;
// c_i=(g_i-off_i)/res_i.  No FMA contraction, IEEE divide.  kPolicy = LS_GEOM_TORCH_CUDA instead
// follows torch's CUDA matmul (include/ls_b200.h LsGeomPolicy).
template <int kPolicy>
__device__ __forceinline__ void ls_point_geom(const float* __restrict__ m, const float* __restrict__ t,
                                              float u, float v, float d, float g[3]) {
  const float px = __fmul_rn(u, d);
  const float py = __fmul_rn(v, d);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float a;
    if (kPolicy == LS_GEOM_TORCH_CUDA) {
      // torch-CUDA (cuBLAS batched 3x3 . 3x1): the first two terms share one fused multiply-add
      a = __fmaf_rn(m[3 * i + 1], py, __fmul_rn(m[3 * i + 0], px));
      a = __fadd_rn(a, __fmul_rn(m[3 * i + 2], d));
    } else {
      a = __fadd_rn(0.0f, __fmul_rn(m[3 * i + 0], px));
      a = __fadd_rn(a, __fmul_rn(m[3 * i + 1], py));
      a = __fadd_rn(a, __fmul_rn(m[3 * i + 2], d));
    }
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

// The same keep decision and (x, y) voxel for the hot kernel when there is a single z cell
// (grid.zfast): with a = g_z - off_z and r = res_z > 0,  -1 < RN(a / r) < 1  <=>  -r < a < r.
// Proof: the floats next to +-r toward zero are at least r * 2^-24 away from it, so their quotient is
// at most 1 - 2^-24 in magnitude - a representable number, hence rounds to itself or below it, never
// to 1; |a| >= r gives |a / r| >= 1, which rounds to >= 1.  (NaN fails both forms.)  One IEEE
// division less per point; the x and y coordinates keep theirs (they need the truncated quotient).
__device__ __forceinline__ bool ls_point_voxel_xy(const float g[3], const LsGrid& grid, int v[3]) {
  bool keep = true;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float c = __fdiv_rn(__fsub_rn(g[i], grid.off[i]), grid.res[i]);
    keep = keep && (c > -1.0f) && (c < grid.fdim[i]);
    v[i] = __float2int_rz(c);
  }
  const float a = __fsub_rn(g[2], grid.off[2]);
  v[2] = 0;
  return keep && (a > -grid.res[2]) && (a < grid.res[2]);
}

template <typename T> __device__ __forceinline__ float ls_to_float(T v);
template <> __device__ __forceinline__ float ls_to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float ls_to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T ls_from_float(float v);
template <> __device__ __forceinline__ float ls_from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 ls_from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// four consecutive channels (16 B of fp32 / 8 B of bf16), read-only path
template <typename T> __device__ __forceinline__ float4 ls_load4(const T* p);
template <> __device__ __forceinline__ float4 ls_load4<float>(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
template <> __device__ __forceinline__ float4 ls_load4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 raw = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
template <typename T> __device__ __forceinline__ void ls_store4(T* p, float4 v);
template <> __device__ __forceinline__ void ls_store4<float>(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}
template <> __device__ __forceinline__ void ls_store4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __float22bfloat162_rn(make_float2(v.x, v.y));
  __nv_bfloat162 b = __float22bfloat162_rn(make_float2(v.z, v.w));
  uint2 raw;
  raw.x = *reinterpret_cast<unsigned*>(&a);
  raw.y = *reinterpret_cast<unsigned*>(&b);
  *reinterpret_cast<uint2*>(p) = raw;
}

// ---- packed float32 pairs (Blackwell FFMA2 / FMUL2: two fused multiply-adds per issue slot) ----
#ifndef LS_FFMA2
#define LS_FFMA2 0   // measured slower on a B200 (splat +3 us, gather +4 us: the pack/unpack moves and register-pair constraints cost more than the saved issue slots)
#endif
__device__ __forceinline__ unsigned long long ls_pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void ls_unpack2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long ls_fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ unsigned long long ls_mul2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// Shared-memory tile [cell][channel] with row stride (cc + 4) floats; the channel quad of a
// cell is XOR-swizzled with bits of the cell index so that BOTH the row-wise float4 accesses
// (one cell, 16 quads) and the column-wise scalar walks (fixed channel, consecutive cells)
// are bank-conflict free.  nq = number of quads per row (power of two <= 16).
__device__ __forceinline__ int ls_tile_quad(int cl, int q, int nq) {
  return q ^ ((cl >> 3) & (nq - 1));
}

// ---- launch bookkeeping (defined in ls_api.cu) ----------------------------------------
void ls_note_launch();
int ls_note_cuda_error(cudaError_t e, const char* file, int line);
#define LS_CUDA(call)                                                     \
  do {                                                                    \
    cudaError_t e__ = (call);                                             \
    if (e__ != cudaSuccess) return ls_note_cuda_error(e__, __FILE__, __LINE__); \
  } while (0)
#define LS_LAUNCHED() do { ls_note_launch(); LS_CUDA(cudaGetLastError()); } while (0)

// ---- programmatic dependent launch ------------------------------------------------------
// Every kernel of the pipeline is launched with the programmatic-stream-serialization
// attribute: it may be scheduled while the tail of its predecessor in the stream is still
// running, does its private prologue (index math, shared-memory zeroing) and blocks in
// ls_pdl_wait() until the predecessor has completed and its writes are visible.  EVERY kernel
// calls ls_pdl_wait() before touching global memory, which keeps the ordering transitive
// along the chain.  ls_pdl_trigger() lets the successor start being scheduled as soon as all
// CTAs of this grid are resident.  Set LS_NO_PDL=1 to launch with plain stream ordering.
__device__ __forceinline__ void ls_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#ifndef LS_PDL_TRIGGER
#define LS_PDL_TRIGGER 0
#endif
__device__ __forceinline__ void ls_pdl_trigger() {
  if (LS_PDL_TRIGGER) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
bool ls_pdl_enabled();   // ls_api.cu

template <typename... KArgs, typename... Args>
static inline cudaError_t ls_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                    Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = ls_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// kernel names with template commas go in parentheses: LS_LAUNCH((k<T, 4>), grid, block, smem, s, args...)
#define LS_LAUNCH(kernel, grid, block, smem, stream, ...)                                   \
  do {                                                                                      \
    cudaError_t e__ = ls_launch(kernel, grid, block, smem, stream, __VA_ARGS__);            \
    if (e__ != cudaSuccess) return ls_note_cuda_error(e__, __FILE__, __LINE__);             \
    ls_note_launch();                                                                       \
  } while (0)
