// Index side of the lift-splat path: camera transform, fused geometry -> voxel -> rank,
// counting sort (histogram with ticket, scan, placement).  Integer work is bit-exact with
// the reference (model/bev_model.py:45-57,85-97); see include/ls_b200.h.
#include <stdlib.h>

#include "ls_internal.h"

// =====================================================================================
// camera transform: E^-1, K^-1 (fp64 Gauss-Jordan, partial pivoting), M = R . K^-1
// reference: model/bev_model.py:46-47,53     oracle: camera_transform
// =====================================================================================
// Column-parallel Gauss-Jordan: lane `base + j` owns column j of the augmented matrix
// [A | I] (n rows in registers); pivot column and factors travel by warp shuffles.  Every
// element sees exactly the operations of oracle/_gauss_jordan_f64 (divide the pivot row, then
// row_i -= a[i][k] * row_k with separate multiply and subtract), so results are bit-identical
// to the serial form; the dependent chain is n pivots instead of n * 2n * n operations.
template <int n>
__device__ __forceinline__ void ls_gauss_jordan_cols(double (&col)[n], int base) {
#pragma unroll
  for (int k = 0; k < n; ++k) {
    double ck[n];
#pragma unroll
    for (int i = 0; i < n; ++i) ck[i] = __shfl_sync(0xffffffffu, col[i], base + k);
    int p = k;
    double best = fabs(ck[k]);
#pragma unroll
    for (int i = k + 1; i < n; ++i) {
      const double v = fabs(ck[i]);
      if (v > best) { best = v; p = i; }   // first maximum wins (numpy argmax)
    }
#pragma unroll
    for (int i = k + 1; i < n; ++i) {
      if (p == i) {
        double tmp = col[k]; col[k] = col[i]; col[i] = tmp;
        tmp = ck[k]; ck[k] = ck[i]; ck[i] = tmp;
      }
    }
    const double rowk = __ddiv_rn(col[k], ck[k]);
#pragma unroll
    for (int i = 0; i < n; ++i)
      if (i != k) col[i] = __dsub_rn(col[i], __dmul_rn(ck[i], rowk));
    col[k] = rowk;
  }
}

// One warp per camera: lanes 0-7 invert the 4x4 extrinsic, lanes 8-13 the 3x3 intrinsic.
__global__ void __launch_bounds__(32)
ls_camera_transform_kernel(const float* __restrict__ intr, const float* __restrict__ extr, int BN,
                           float* __restrict__ M, float* __restrict__ t) {
  ls_pdl_wait();
  // 64 warps of dependent fp64 arithmetic leave the GPU empty: once everything before this kernel is
  // complete (the wait above), the next launch may start - ls_forward's first kernel zeroes the
  // histogram without needing M, t and only joins this kernel's completion at its end
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int i = blockIdx.x, lane = threadIdx.x;
  double e[4], k[3];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int c = lane & 7;
    e[r] = (c < 4) ? (double)extr[i * 16 + r * 4 + c] : ((c - 4 == r) ? 1.0 : 0.0);
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int c = (lane >= 8 && lane < 14) ? lane - 8 : 0;
    k[r] = (c < 3) ? (double)intr[i * 9 + r * 3 + c] : ((c - 3 == r) ? 1.0 : 0.0);
  }
  ls_gauss_jordan_cols<4>(e, 0);
  ls_gauss_jordan_cols<3>(k, 8);
  // E^-1[r][c] sits in lane 4+c (e[r]); K^-1[q][c] in lane 11+c (k[q]).  Lane 11+c builds
  // column c of M = R . K^-1 (aten's small-matrix bmm: unfused, q ascending, from +0).
  float rot[3][3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int q = 0; q < 3; ++q) rot[r][q] = (float)__shfl_sync(0xffffffffu, e[r], 4 + q);
  if (lane >= 11 && lane < 14) {
    const int c = lane - 11;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float acc = 0.0f;
#pragma unroll
      for (int q = 0; q < 3; ++q) acc = __fadd_rn(acc, __fmul_rn(rot[r][q], (float)k[q]));
      M[i * 9 + r * 3 + c] = acc;
    }
  }
  if (lane == 7) {
#pragma unroll
    for (int r = 0; r < 3; ++r) t[i * 3 + r] = (float)e[r];
  }
}

int ls_launch_camera_transform(const float* intr, const float* extr, int BN, float* M, float* t, cudaStream_t s) {
  LS_LAUNCH(ls_camera_transform_kernel, dim3(BN), dim3(32), 0, s, intr, extr, BN, M, t);
  return LS_OK;
}

// =====================================================================================
// K1: fused lift geometry -> voxel index -> keep -> rank  (+ histogram ticket)
// reference: model/bev_model.py:49-55 (get_geometry), :85-95 (voxelise, mask, rank)
// One thread per frustum point; geom never leaves registers.  The per-cell histogram is one
// integer atomicAdd per kept point whose return value ("ticket") is the point's slot inside
// its cell, so placement later needs no second atomic pass.
// =====================================================================================
#ifndef LS_IDX_ILP
#define LS_IDX_ILP 2
#endif

// kGeomIn: the ego-frame coordinates are given (proj_bev_feature(geom, x) of the reference API,
// model/bev_model.py:74-107) instead of computed from the camera transform and the frustum.
template <bool kExport, bool kGeomIn, int kPolicy>
__global__ void __launch_bounds__(256)
ls_index_kernel(const float* __restrict__ M, const float* __restrict__ t, const float* __restrict__ frustum,
                LsDims dm, LsGrid grid, int* __restrict__ rank, int* __restrict__ cell, int* __restrict__ within,
                int* __restrict__ counts, float* __restrict__ geom_out, long long* __restrict__ vox_out,
                unsigned char* __restrict__ keep_out, long long* __restrict__ rank64_out) {
  ls_pdl_trigger();
  ls_pdl_wait();
  const int b = blockIdx.z, n = blockIdx.y;
  float cam[12];   // warp-uniform loads: every thread keeps the camera's 3x3 + translation in registers
  if (!kGeomIn) {
#pragma unroll
    for (int k = 0; k < 9; ++k) cam[k] = __ldg(M + (b * dm.N + n) * 9 + k);
#pragma unroll
    for (int k = 0; k < 3; ++k) cam[9 + k] = __ldg(t + (b * dm.N + n) * 3 + k);
  }
  // LS_IDX_ILP points per thread (strided by the CTA width): their loads, divisions and the
  // histogram atomics are independent, so the returning atomics overlap instead of serialising
  const int i0 = blockIdx.x * (blockDim.x * LS_IDX_ILP) + threadIdx.x;
  float u[LS_IDX_ILP], v[LS_IDX_ILP], d[LS_IDX_ILP];
#pragma unroll
  for (int k = 0; k < LS_IDX_ILP; ++k) {
    const int i = i0 + k * blockDim.x;
    const bool in = i < dm.DHW;
    // kGeomIn: `frustum` is geom[B,Npts,3]; (u,v,d) carry the point's coordinates
    const float* src = kGeomIn ? frustum + 3 * ((size_t)b * dm.Npts + (size_t)n * dm.DHW + i) : frustum + 3 * i;
    u[k] = in ? __ldg(src + 0) : 0.0f;
    v[k] = in ? __ldg(src + 1) : 0.0f;
    d[k] = in ? __ldg(src + 2) : 0.0f;
  }
  bool keep[LS_IDX_ILP];
  int vx[LS_IDX_ILP][3], cid[LS_IDX_ILP], tk[LS_IDX_ILP];
  float g[LS_IDX_ILP][3], c[LS_IDX_ILP][3];
#pragma unroll
  for (int k = 0; k < LS_IDX_ILP; ++k) {
    if (kGeomIn) { g[k][0] = u[k]; g[k][1] = v[k]; g[k][2] = d[k]; }
    else ls_point_geom<kPolicy>(cam, cam + 9, u[k], v[k], d[k], g[k]);
    if (!kExport && grid.zfast) keep[k] = ls_point_voxel_xy(g[k], grid, vx[k]);
    else keep[k] = ls_point_voxel(g[k], grid, c[k], vx[k]);
    keep[k] = keep[k] && (i0 + k * blockDim.x < dm.DHW);
    cid[k] = -1;
    tk[k] = 0;
  }
  if (!kExport && cell) {
#pragma unroll
    for (int k = 0; k < LS_IDX_ILP; ++k) {
      if (keep[k]) {
        cid[k] = ls_cell_of_xy(vx[k][0], vx[k][1], grid);
        tk[k] = atomicAdd(&counts[(size_t)b * grid.Vc + cid[k]], 1);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < LS_IDX_ILP; ++k) {
    const int i = i0 + k * blockDim.x;
    if (i >= dm.DHW) continue;
    const size_t p = (size_t)b * dm.Npts + (size_t)n * dm.DHW + i;
    const int r = keep[k] ? (vx[k][0] * (grid.Y * grid.Z) + vx[k][1] * grid.Z + vx[k][2]) : -1;
    if (kExport) {
      if (geom_out) { geom_out[3 * p + 0] = g[k][0]; geom_out[3 * p + 1] = g[k][1]; geom_out[3 * p + 2] = g[k][2]; }
      if (vox_out) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          // Tensor.long() on x86: cvttss2si -> INT64_MIN for NaN / out of range
          const bool ok = fabsf(c[k][a]) < 9.2e18f;
          vox_out[3 * p + a] = ok ? __float2ll_rz(c[k][a]) : (long long)0x8000000000000000ULL;
        }
      }
      if (keep_out) keep_out[p] = keep[k] ? 1 : 0;
      if (rank64_out) rank64_out[p] = (long long)r;
    } else {
      if (rank) rank[p] = r;
      if (cell) { cell[p] = cid[k]; within[p] = tk[k]; }
    }
  }
}


// The hot form of K1 (what ls_forward launches: sort outputs only, single z cell): LS_IDX_PIPE points per thread,
// one after the other, software-pipelined around the histogram atomic.  The kernel above issues a thread's atomics
// together and then WAITS for the tickets before it can store them and retire (the stores of `within` collect most
// of its stall samples; without the returning atomic it runs in 21 us instead of 27).  Here the ticket of point i
// is stored after the geometry of point i+1 has been computed, so the L2 round trip of the atomic hides behind
// ~150 arithmetic instructions of the same thread.
#ifndef LS_IDX_PIPE
#define LS_IDX_PIPE 4
#endif
template <int kPolicy>
__global__ void __launch_bounds__(256)
ls_index_pipe_kernel(const float* __restrict__ M, const float* __restrict__ t, const float* __restrict__ frustum,
                     LsDims dm, LsGrid grid, int* __restrict__ cell, int* __restrict__ within, int* __restrict__ counts) {
  ls_pdl_trigger();
  ls_pdl_wait();
  const int b = blockIdx.z, n = blockIdx.y;
  float cam[12];
#pragma unroll
  for (int k = 0; k < 9; ++k) cam[k] = __ldg(M + (b * dm.N + n) * 9 + k);
#pragma unroll
  for (int k = 0; k < 3; ++k) cam[9 + k] = __ldg(t + (b * dm.N + n) * 3 + k);
  const int i0 = blockIdx.x * (256 * LS_IDX_PIPE) + threadIdx.x;
  int* __restrict__ cnt = counts + (size_t)b * grid.Vc;
  const size_t base = (size_t)b * dm.Npts + (size_t)n * dm.DHW;
  int* __restrict__ cellp = cell + base;
  int* __restrict__ withp = within + base;
  // the frustum of every point of this thread up front (independent loads, L2-resident table)
  float u[LS_IDX_PIPE], v[LS_IDX_PIPE], d[LS_IDX_PIPE];
#pragma unroll
  for (int k = 0; k < LS_IDX_PIPE; ++k) {
    const int i = i0 + k * 256;
    const float* src = frustum + 3 * (i < dm.DHW ? i : 0);
    u[k] = __ldg(src + 0); v[k] = __ldg(src + 1); d[k] = __ldg(src + 2);
  }
  int prev_tk = 0, prev_i = -1;
#pragma unroll
  for (int k = 0; k < LS_IDX_PIPE; ++k) {
    const int i = i0 + k * 256;
    float g[3];
    int vx[3];
    ls_point_geom<kPolicy>(cam, cam + 9, u[k], v[k], d[k], g);
    const bool keep = ls_point_voxel_xy(g, grid, vx) && i < dm.DHW;
    const int cid = keep ? ls_cell_of_xy(vx[0], vx[1], grid) : -1;
    if (prev_i >= 0) withp[prev_i] = prev_tk;           // the previous point's ticket has had a whole point's arithmetic to arrive
    prev_tk = 0;
    prev_i = -1;
    if (i < dm.DHW) {
      if (keep) prev_tk = atomicAdd(cnt + cid, 1);
      cellp[i] = cid;
      prev_i = i;
    }
  }
  if (prev_i >= 0) withp[prev_i] = prev_tk;
}

// zero the per-cell histogram (B * Vc ints, a multiple of 128): one 16-byte store per thread and trip
__global__ void __launch_bounds__(256)
ls_zero_counts_kernel(int4* __restrict__ p, int n16, int* __restrict__ extra, int n_extra) {
  ls_pdl_trigger();
  // The histogram is private scratch of this call and nothing written here is read by the predecessor,
  // so the zeroing does not wait for it.  A predecessor lets this kernel start early only after ITS
  // dependency wait (ls_camera_transform_kernel) or at its completion, i.e. when all earlier users of
  // the scratch blob are done.  The dependency wait comes last: this grid completes after the
  // predecessor, which keeps the stream order transitive for the kernels that wait on this one.
  const int4 z = make_int4(0, 0, 0, 0);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) p[i] = z;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_extra; i += gridDim.x * blockDim.x) extra[i] = 0;
  ls_pdl_wait();
}

int ls_launch_zero_counts(int* counts, const LsDims& dm, const LsGrid& g, int* extra, int n_extra, cudaStream_t s) {
  const int n16 = (int)(((size_t)dm.B * g.Vc) / 4);      // Vc is a multiple of LS_TILE
  const int blocks = (n16 + 1023) / 1024;
  LS_LAUNCH(ls_zero_counts_kernel, dim3(blocks < 1 ? 1 : blocks), dim3(256), 0, s, reinterpret_cast<int4*>(counts), n16,
            extra, extra ? n_extra : 0);
  return LS_OK;
}

int ls_launch_index(const float* M, const float* t, const float* frustum, const LsDims& dm, const LsGrid& g,
                    int* rank, int* cell, int* within, int* counts, cudaStream_t s) {
  static const bool no_pipe = getenv("LS_INDEX_NO_PIPE") != nullptr;
  if (!rank && cell && within && counts && g.zfast && !no_pipe) {      // the hot form: ls_forward
    dim3 pgrid((dm.DHW + 256 * LS_IDX_PIPE - 1) / (256 * LS_IDX_PIPE), dm.N, dm.B);
    if (dm.policy == LS_GEOM_TORCH_CUDA)
      LS_LAUNCH(ls_index_pipe_kernel<LS_GEOM_TORCH_CUDA>, pgrid, dim3(256), 0, s, M, t, frustum, dm, g, cell, within, counts);
    else
      LS_LAUNCH(ls_index_pipe_kernel<LS_GEOM_TORCH_CPU>, pgrid, dim3(256), 0, s, M, t, frustum, dm, g, cell, within, counts);
    return LS_OK;
  }
  dim3 grid((dm.DHW + 256 * LS_IDX_ILP - 1) / (256 * LS_IDX_ILP), dm.N, dm.B);
  if (dm.policy == LS_GEOM_TORCH_CUDA)
    LS_LAUNCH((ls_index_kernel<false, false, LS_GEOM_TORCH_CUDA>), grid, dim3(256), 0, s, M, t, frustum, dm, g, rank,
              cell, within, counts, (float*)nullptr, (long long*)nullptr, (unsigned char*)nullptr, (long long*)nullptr);
  else
    LS_LAUNCH((ls_index_kernel<false, false, LS_GEOM_TORCH_CPU>), grid, dim3(256), 0, s, M, t, frustum, dm, g, rank,
              cell, within, counts, (float*)nullptr, (long long*)nullptr, (unsigned char*)nullptr, (long long*)nullptr);
  return LS_OK;
}

int ls_launch_index_geom(const float* geom, const LsDims& dm, const LsGrid& g, int* rank, int* cell, int* within,
                         int* counts, cudaStream_t s) {
  dim3 grid((dm.DHW + 256 * LS_IDX_ILP - 1) / (256 * LS_IDX_ILP), dm.N, dm.B);
  LS_LAUNCH((ls_index_kernel<false, true, LS_GEOM_TORCH_CPU>), grid, dim3(256), 0, s, (const float*)nullptr, (const float*)nullptr, geom,
            dm, g, rank, cell, within, counts, (float*)nullptr, (long long*)nullptr, (unsigned char*)nullptr,
            (long long*)nullptr);
  return LS_OK;
}

int ls_launch_export(const float* M, const float* t, const float* frustum, const LsDims& dm, const LsGrid& g,
                     float* geom, long long* vox, unsigned char* keep, long long* rank64, cudaStream_t s) {
  dim3 grid((dm.DHW + 256 * LS_IDX_ILP - 1) / (256 * LS_IDX_ILP), dm.N, dm.B);
  if (dm.policy == LS_GEOM_TORCH_CUDA)
    ls_index_kernel<true, false, LS_GEOM_TORCH_CUDA><<<grid, 256, 0, s>>>(M, t, frustum, dm, g, nullptr, nullptr, nullptr,
                                                                          nullptr, geom, vox, keep, rank64);
  else
    ls_index_kernel<true, false, LS_GEOM_TORCH_CPU><<<grid, 256, 0, s>>>(M, t, frustum, dm, g, nullptr, nullptr, nullptr,
                                                                         nullptr, geom, vox, keep, rank64);
  LS_LAUNCHED();
  return LS_OK;
}

// =====================================================================================
// K2a: exclusive scan of the per-cell histogram -> CSR offsets (one CTA per sample).
// Replaces the boundary mask of tool/geometry.py:295-296.  Three phases per chunk of 1024
// tiles: per-tile totals (one warp per tile, 8 cells per lane), block scan of the totals,
// then in-tile scans written as 16-byte stores.
// =====================================================================================
// A warp covers one tile: LS_TILE/32 consecutive cells per lane (8 -> two int4, 4 -> one).
__device__ __forceinline__ void ls_load_counts(const int* __restrict__ tile_counts, int lane, int4& a, int4& c) {
  static_assert(LS_TILE == 256 || LS_TILE == 128, "tile must have 128 or 256 cells");
  const int4* q = reinterpret_cast<const int4*>(tile_counts) + lane * (LS_TILE / 128);
  a = q[0];
  c = (LS_TILE == 256) ? q[LS_TILE == 256 ? 1 : 0] : make_int4(0, 0, 0, 0);
}
__device__ __forceinline__ void ls_store_offsets(int* __restrict__ tile_seg, int lane, int4 o0, int4 o1) {
  int4* dst = reinterpret_cast<int4*>(tile_seg) + lane * (LS_TILE / 128);
  dst[0] = o0;
  if (LS_TILE == 256) dst[1] = o1;
}

__device__ __forceinline__ int ls_warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += y;
  }
  return v;
}

__global__ void __launch_bounds__(1024)
ls_scan_kernel(const int* __restrict__ counts, LsGrid g, int* __restrict__ seg_start, int* __restrict__ tile_order) {
  ls_pdl_trigger();
  ls_pdl_wait();
  __shared__ int tile_base[1024];
  __shared__ int warp_tot[32];
  __shared__ int chunk_total;
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int* cnt = counts + (size_t)b * g.Vc;
  int* seg = seg_start + (size_t)b * g.seg_stride;
  int carry = 0;
  for (int t0 = 0; t0 < g.tiles; t0 += 1024) {
    const int tend = min(g.tiles, t0 + 1024);
    for (int t = t0 + warp; t < tend; t += 32) {
      int4 a, c;
      ls_load_counts(cnt + (size_t)t * LS_TILE, lane, a, c);
      int s = a.x + a.y + a.z + a.w + c.x + c.y + c.z + c.w;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) tile_base[t - t0] = s;
    }
    __syncthreads();
    const int v = (t0 + tid < tend) ? tile_base[tid] : 0;
    if (tile_order) {
      // heaviest tile first (longest-processing-time order for the splat's CTAs); exact for
      // up to 1024 tiles, identity beyond
      int* ord = tile_order + (size_t)b * g.tiles;
      if (g.tiles <= 1024) {
        if (tid < g.tiles) {
          int r = 0;
          for (int j = 0; j < g.tiles; ++j) {
            const int o = tile_base[j];
            r += (o > v || (o == v && j < tid)) ? 1 : 0;
          }
          ord[r] = tid;
        }
      } else if (t0 + tid < tend) {
        ord[t0 + tid] = t0 + tid;
      }
    }
    const int incl = ls_warp_incl_scan(v, lane);
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const int w = warp_tot[lane];
      const int wi = ls_warp_incl_scan(w, lane);
      warp_tot[lane] = wi - w;
      if (lane == 31) chunk_total = wi;
    }
    __syncthreads();
    tile_base[tid] = carry + warp_tot[warp] + incl - v;
    __syncthreads();
    for (int t = t0 + warp; t < tend; t += 32) {
      int4 a, c;
      ls_load_counts(cnt + (size_t)t * LS_TILE, lane, a, c);
      const int s = a.x + a.y + a.z + a.w + c.x + c.y + c.z + c.w;
      const int base = tile_base[t - t0] + ls_warp_incl_scan(s, lane) - s;
      int4 o0, o1;
      o0.x = base; o0.y = o0.x + a.x; o0.z = o0.y + a.y; o0.w = o0.z + a.z;
      o1.x = o0.w + a.w; o1.y = o1.x + c.x; o1.z = o1.y + c.y; o1.w = o1.z + c.z;
      ls_store_offsets(seg + (size_t)t * LS_TILE, lane, o0, o1);
    }
    carry += chunk_total;
    __syncthreads();
  }
  if (tid == 0) seg[g.Vc] = carry;
}

// Parallel form for grids of up to 2048 tiles: (1) one warp per tile sums its 256 counts;
// (2) one warp per tile rebuilds its own base (sum of the totals of the tiles before it), its
// place in the heaviest-first order, and the in-tile scan.  No cross-CTA dependency.
__global__ void __launch_bounds__(256)
ls_tile_totals_kernel(const int* __restrict__ counts, int ntiles_all, int* __restrict__ tile_tot) {
  ls_pdl_trigger();
  ls_pdl_wait();
  const int t = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (t >= ntiles_all) return;
  int4 a, c;
  ls_load_counts(counts + (size_t)t * LS_TILE, lane, a, c);
  int s = a.x + a.y + a.z + a.w + c.x + c.y + c.z + c.w;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) tile_tot[t] = s;
}

__global__ void __launch_bounds__(256)
ls_tile_scan_kernel(const int* __restrict__ counts, const int* __restrict__ tile_tot, LsGrid g,
                    int* __restrict__ seg_start, int* __restrict__ tile_order) {
  ls_pdl_trigger();
  ls_pdl_wait();
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (t >= g.tiles) return;
  const int* tot = tile_tot + (size_t)b * g.tiles;
  const int mine = tot[t];
  int base = 0, rank = 0;
  for (int j = lane; j < g.tiles; j += 32) {
    const int o = tot[j];
    base += (j < t) ? o : 0;
    rank += (o > mine || (o == mine && j < t)) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    base += __shfl_xor_sync(0xffffffffu, base, o);
    rank += __shfl_xor_sync(0xffffffffu, rank, o);
  }
  int4 a, c;
  ls_load_counts(counts + ((size_t)b * g.tiles + t) * LS_TILE, lane, a, c);
  const int s = a.x + a.y + a.z + a.w + c.x + c.y + c.z + c.w;
  const int e0 = base + ls_warp_incl_scan(s, lane) - s;
  int4 o0, o1;
  o0.x = e0; o0.y = o0.x + a.x; o0.z = o0.y + a.y; o0.w = o0.z + a.z;
  o1.x = o0.w + a.w; o1.y = o1.x + c.x; o1.z = o1.y + c.y; o1.w = o1.z + c.z;
  int* seg = seg_start + (size_t)b * g.seg_stride;
  ls_store_offsets(seg + (size_t)t * LS_TILE, lane, o0, o1);
  if (lane == 0) {
    if (tile_order) tile_order[(size_t)b * g.tiles + rank] = t;
    if (t == g.tiles - 1) seg[g.Vc] = base + mine;
  }
}

int ls_launch_scan(const int* counts, const LsDims& dm, const LsGrid& g, int* seg_start, int* tile_order,
                   int* tile_tot, cudaStream_t s) {
  if (g.tiles <= 2048 && tile_tot) {
    const int all = g.tiles * dm.B;
    LS_LAUNCH(ls_tile_totals_kernel, dim3((all + 7) / 8), dim3(256), 0, s, counts, all, tile_tot);
    LS_LAUNCH(ls_tile_scan_kernel, dim3((g.tiles + 7) / 8, dm.B), dim3(256), 0, s, counts, (const int*)tile_tot, g,
              seg_start, tile_order);
    return LS_OK;
  }
  LS_LAUNCH(ls_scan_kernel, dim3(dm.B), dim3(1024), 0, s, counts, g, seg_start, tile_order);
  return LS_OK;
}

// =====================================================================================
// K2b: placement.  Replaces argsort + gathers of model/bev_model.py:96-97: every kept point
// writes one 8-byte record (sort key, prob) into slot seg_start[cell] + ticket.  The ticket
// order inside a cell is arbitrary (atomics); ls_splat_fwd re-orders each cell by key, so the
// sums are deterministic.  key = cell_in_tile << 24 | (pixel << dbits | d).
// With pix_recs != NULL it also emits, pixel-major, the (rank, prob) pair of every depth bin
// of every pixel - the index the pixel-stationary backward walks (dropped points carry
// rank X*Y: one past the sample's last cell).
// CTA = (sample, camera, 32 consecutive pixels) x all depth bins; loads are coalesced over
// pixels, the pixel-major rows are transposed through shared memory.
// =====================================================================================
#ifndef LS_PLACE_GROUPS
#define LS_PLACE_GROUPS 16    // depth groups (warps) per CTA; a thread handles ceil(D / groups) bins of one pixel
#endif
template <typename T, int K>
__global__ void __launch_bounds__(32 * LS_PLACE_GROUPS)
ls_place_kernel(const int* __restrict__ cell, const int* __restrict__ within, const T* __restrict__ prob, LsDims dm,
                LsGrid grid, const int* __restrict__ seg_start, int2* __restrict__ recs, int2* __restrict__ pix_recs) {
  ls_pdl_trigger();
  ls_pdl_wait();
  extern __shared__ int2 stage[];                 // [32][Dp], Dp odd
  const int Dp = dm.D | 1;
  const int b = blockIdx.z, n = blockIdx.y, rc0 = blockIdx.x * 32;
  const int lane = threadIdx.x & 31, dg = threadIdx.x >> 5;
  const int rc = rc0 + lane;
  // block-uniform 64-bit bases; everything per thread is a 32-bit offset inside the camera / the sample
  // (B*Npts < 2^31 is checked by the API), loads are unconditional on clamped addresses: no branch per load
  const size_t cam0 = (size_t)b * dm.Npts + (size_t)n * dm.DHW;
  const int* __restrict__ cellp = cell + cam0;
  const int* __restrict__ withp = within + cam0;
  const T* __restrict__ probp = prob + cam0;
  const int* __restrict__ seg = seg_start + (size_t)b * grid.seg_stride;
  int2* __restrict__ rb = recs + (size_t)b * dm.Npts;
  if (rc < dm.HW) {
    const int pixd = (n * dm.HW + rc) << dm.dbits;
    // K = ceil(D/groups) depth bins per thread, unrolled: the three streams and the dependent
    // seg_start lookups of all bins are in flight together
    int c[K], tk[K], wb[K], sg[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int d = dg + LS_PLACE_GROUPS * k;
      const int off = (d < dm.D ? d : 0) * dm.HW + rc;
      c[k] = cellp[off];
      tk[k] = withp[off];
      wb[k] = __float_as_int(ls_to_float(probp[off]));
      if (d >= dm.D) c[k] = -1;
    }
#pragma unroll
    for (int k = 0; k < K; ++k) sg[k] = __ldg(seg + (c[k] >= 0 ? c[k] : 0));
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int d = dg + LS_PLACE_GROUPS * k;
      if (c[k] >= 0) rb[sg[k] + tk[k]] = make_int2(((c[k] & (LS_TILE - 1)) << 24) | pixd | d, wb[k]);
      // row of the cell in rank order (gx*Y + gy: a row of a channels-last gradient, or of the
      // staged cell-major copy of an NCHW one); X*Y for a dropped point
      if (pix_recs && d < dm.D)
        stage[lane * Dp + d] = make_int2(c[k] >= 0 ? ls_rank_of_cell_fast(c[k], grid) : grid.XY, wb[k]);
    }
  }
  if (!pix_recs) return;
  __syncthreads();
  const int valid = min(32, dm.HW - rc0);
  int2* __restrict__ dst = pix_recs + ((size_t)(b * dm.N + n) * dm.HW + rc0) * dm.D;
  for (int r = dg; r < valid; r += LS_PLACE_GROUPS)
    for (int d = lane; d < dm.D; d += 32) dst[r * dm.D + d] = stage[r * Dp + d];
}

template <typename T>
static int ls_place_dispatch(const int* cell, const int* within, const T* prob, const LsDims& dm, const LsGrid& g,
                             const int* seg_start, int2* recs, int2* pix_recs, dim3 grid, size_t smem, cudaStream_t s) {
  const int k = (dm.D + LS_PLACE_GROUPS - 1) / LS_PLACE_GROUPS;
#define LS_PL(KK) \
  LS_LAUNCH((ls_place_kernel<T, KK>), grid, dim3(32 * LS_PLACE_GROUPS), smem, s, cell, within, prob, dm, g, seg_start, recs, pix_recs)
  if (k <= 1) LS_PL(1);
  else if (k <= 2) LS_PL(2);
  else if (k <= 3) LS_PL(3);
  else if (k <= 4) LS_PL(4);
  else if (k <= 6) LS_PL(6);
  else if (k <= 8) LS_PL(8);
  else if (k <= 12) LS_PL(12);
  else if (k <= 16) LS_PL(16);
  else if (k <= 32) LS_PL(32);
  else return LS_ERR_UNSUPPORTED;
#undef LS_PL
  return LS_OK;
}

int ls_launch_place(const int* cell, const int* within, const void* prob, int dtype, const LsDims& dm,
                    const LsGrid& g, const int* seg_start, int2* recs, int2* pix_recs, cudaStream_t s) {
  dim3 grid((dm.HW + 31) / 32, dm.N, dm.B);
  const size_t smem = pix_recs ? (size_t)32 * (dm.D | 1) * sizeof(int2) : 0;
  if (smem > 48 * 1024) return LS_ERR_UNSUPPORTED;
  int rc;
  if (dtype == LS_F32)
    rc = ls_place_dispatch<float>(cell, within, (const float*)prob, dm, g, seg_start, recs, pix_recs, grid, smem, s);
  else
    rc = ls_place_dispatch<__nv_bfloat16>(cell, within, (const __nv_bfloat16*)prob, dm, g, seg_start, recs, pix_recs,
                                          grid, smem, s);
  return rc;
}

// (row-major rank, point count) of every cell of sample b, from the CSR (test export)
__global__ void ls_export_cell_counts_kernel(const int* __restrict__ seg_start, LsGrid grid, int b,
                                             long long* __restrict__ out) {
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= grid.Vc) return;
  const int* seg = seg_start + (size_t)b * grid.seg_stride;
  const int r = ls_rank_of_cell(cell, grid);
  const int k = seg[cell + 1] - seg[cell];
  out[2 * (size_t)cell + 0] = (r >= 0 && k > 0) ? r : 0;
  out[2 * (size_t)cell + 1] = (r >= 0) ? k : 0;
}

int ls_launch_export_cell_counts(const int* seg_start, const LsGrid& g, int B, int b, long long* out, int* kept,
                                 cudaStream_t s) {
  if (out) {
    ls_export_cell_counts_kernel<<<(g.Vc + 255) / 256, 256, 0, s>>>(seg_start, g, b, out);
    LS_LAUNCHED();
  }
  if (kept) {
    LS_CUDA(cudaMemcpy2DAsync(kept, sizeof(int), seg_start + g.Vc, (size_t)g.seg_stride * sizeof(int), sizeof(int), B,
                              cudaMemcpyDeviceToDevice, s));
    ls_note_launch();
  }
  return LS_OK;
}
