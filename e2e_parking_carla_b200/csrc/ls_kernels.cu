// Lift-splat kernels for B200 (sm_100a).  See include/ls_b200.h for the contract and
// DESIGN.md for the data layout.  Reference semantics: model/bev_model.py,
// tool/geometry.py:285-317 of qintonguav/e2e-parking-carla (cited per kernel).
#include "ls_common.cuh"

// =====================================================================================
// camera transform: E^-1, K^-1 (fp64 Gauss-Jordan, partial pivoting), M = R . K^-1
// reference: model/bev_model.py:46-47,53     oracle: camera_transform
// =====================================================================================
template <int n>
__device__ void ls_gauss_jordan(double (*a)[2 * n]) {
  for (int k = 0; k < n; ++k) {
    int p = k;
    double best = fabs(a[k][k]);
    for (int i = k + 1; i < n; ++i) {
      const double v = fabs(a[i][k]);
      if (v > best) { best = v; p = i; }   // first maximum wins (numpy argmax)
    }
    if (p != k) {
      for (int j = 0; j < 2 * n; ++j) { const double tmp = a[k][j]; a[k][j] = a[p][j]; a[p][j] = tmp; }
    }
    const double piv = a[k][k];
    for (int j = 0; j < 2 * n; ++j) a[k][j] = __ddiv_rn(a[k][j], piv);
    for (int i = 0; i < n; ++i) {
      if (i == k) continue;
      const double f = a[i][k];
      for (int j = 0; j < 2 * n; ++j) a[i][j] = __dsub_rn(a[i][j], __dmul_rn(f, a[k][j]));
    }
  }
}

__global__ void ls_camera_transform_kernel(const float* __restrict__ intr, const float* __restrict__ extr,
                                           int BN, float* __restrict__ M, float* __restrict__ t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BN) return;
  double e[4][8];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) { e[r][c] = (double)extr[i * 16 + r * 4 + c]; e[r][4 + c] = (r == c) ? 1.0 : 0.0; }
  ls_gauss_jordan<4>(e);
  double k[3][6];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) { k[r][c] = (double)intr[i * 9 + r * 3 + c]; k[r][3 + c] = (r == c) ? 1.0 : 0.0; }
  ls_gauss_jordan<3>(k);
  float rot[3][3], kin[3][3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) { rot[r][c] = (float)e[r][4 + c]; kin[r][c] = (float)k[r][3 + c]; }
  for (int r = 0; r < 3; ++r) {
    t[i * 3 + r] = (float)e[r][7];
    for (int c = 0; c < 3; ++c) {
      float acc = 0.0f;  // aten's small-matrix bmm: unfused, k ascending, from +0
      for (int q = 0; q < 3; ++q) acc = __fadd_rn(acc, __fmul_rn(rot[r][q], kin[q][c]));
      M[i * 9 + r * 3 + c] = acc;
    }
  }
}

// =====================================================================================
// K1: fused lift geometry -> voxel index -> keep -> rank (+ per-cell histogram)
// reference: model/bev_model.py:49-55 (get_geometry), :85-95 (voxelise, mask, rank)
// One thread per frustum point; geom never leaves registers.
// =====================================================================================
template <bool kExport>
__global__ void __launch_bounds__(256)
ls_index_kernel(const float* __restrict__ M, const float* __restrict__ t, const float* __restrict__ frustum,
                LsDims dm, LsGrid grid, int* __restrict__ rank, int* __restrict__ counts,
                float* __restrict__ geom_out, long long* __restrict__ vox_out,
                unsigned char* __restrict__ keep_out, long long* __restrict__ rank64_out) {
  __shared__ float cam[12];
  const int b = blockIdx.z, n = blockIdx.y;
  if (threadIdx.x < 9) cam[threadIdx.x] = M[(b * dm.N + n) * 9 + threadIdx.x];
  else if (threadIdx.x < 12) cam[threadIdx.x] = t[(b * dm.N + n) * 3 + threadIdx.x - 9];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= dm.DHW) return;
  const float u = frustum[3 * i + 0], v = frustum[3 * i + 1], d = frustum[3 * i + 2];
  float g[3], c[3];
  int vx[3];
  ls_point_geom(cam, cam + 9, u, v, d, g);
  const bool keep = ls_point_voxel(g, grid, c, vx);
  const size_t p = (size_t)b * dm.Npts + (size_t)n * dm.DHW + i;
  const int r = keep ? (vx[0] * (grid.Y * grid.Z) + vx[1] * grid.Z + vx[2]) : -1;
  if (kExport) {
    if (geom_out) { geom_out[3 * p + 0] = g[0]; geom_out[3 * p + 1] = g[1]; geom_out[3 * p + 2] = g[2]; }
    if (vox_out) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        // Tensor.long() on x86: cvttss2si -> INT64_MIN for NaN / out of range
        const bool ok = fabsf(c[a]) < 9.2e18f;
        vox_out[3 * p + a] = ok ? __float2ll_rz(c[a]) : (long long)0x8000000000000000ULL;
      }
    }
    if (keep_out) keep_out[p] = keep ? 1 : 0;
    if (rank64_out) rank64_out[p] = (long long)r;
  } else {
    if (rank) rank[p] = r;
    if (counts && keep) atomicAdd(&counts[(size_t)b * grid.Vc + ls_cell_of_xy(vx[0], vx[1], grid.tiles_y)], 1);
  }
}

// histogram from a rank array (ls_sort with have_hist == 0)
__global__ void __launch_bounds__(256)
ls_hist_kernel(const int* __restrict__ rank, LsDims dm, LsGrid grid, int* __restrict__ counts) {
  const int b = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= dm.Npts) return;
  const int r = rank[(size_t)b * dm.Npts + p];
  if (r >= 0) atomicAdd(&counts[(size_t)b * grid.Vc + ls_cell_of_rank(r, grid.Y, grid.tiles_y)], 1);
}

// =====================================================================================
// K2a: exclusive scan of the per-cell histogram -> CSR offsets (one CTA per sample)
// replaces the boundary mask of tool/geometry.py:295-296
// =====================================================================================
__global__ void __launch_bounds__(1024)
ls_scan_kernel(const int* __restrict__ counts, int Vc, int* __restrict__ seg_start) {
  __shared__ int warp_tot[32];
  __shared__ int block_tot;
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int* cnt = counts + (size_t)b * Vc;
  int* seg = seg_start + (size_t)b * (Vc + 1);
  int running = 0;
  for (int base = 0; base < Vc; base += 4096) {
    const int i = base + tid * 4;
    int4 v = make_int4(0, 0, 0, 0);
    if (i < Vc) v = *reinterpret_cast<const int4*>(cnt + i);   // Vc is a multiple of 256
    const int tsum = v.x + v.y + v.z + v.w;
    int incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = warp_tot[lane];
      int wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += y;
      }
      warp_tot[lane] = wi - w;           // exclusive prefix of warp totals
      if (lane == 31) block_tot = wi;
    }
    __syncthreads();
    if (i < Vc) {
      const int e0 = running + warp_tot[warp] + incl - tsum;
      seg[i + 0] = e0;
      seg[i + 1] = e0 + v.x;
      seg[i + 2] = e0 + v.x + v.y;
      seg[i + 3] = e0 + v.x + v.y + v.z;
    }
    running += block_tot;
    __syncthreads();
  }
  if (tid == 0) seg[Vc] = running;
}

// =====================================================================================
// K2b: placement - every kept point takes one slot of its cell's segment.
// replaces argsort + gathers of model/bev_model.py:96-97.  Slot order inside a cell is
// arbitrary here (atomics); ls_splat_fwd sums each cell in ascending point id, so the
// result is deterministic.  counts is decremented back to all-zero.
// =====================================================================================
__global__ void __launch_bounds__(256)
ls_place_kernel(const int* __restrict__ rank, LsDims dm, LsGrid grid, int* __restrict__ counts,
                const int* __restrict__ seg_start, int* __restrict__ order) {
  const int b = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= dm.Npts) return;
  const int r = rank[(size_t)b * dm.Npts + p];
  if (r < 0) return;
  const int cell = ls_cell_of_rank(r, grid.Y, grid.tiles_y);
  const int old = atomicSub(&counts[(size_t)b * grid.Vc + cell], 1);
  const int slot = seg_start[(size_t)b * (grid.Vc + 1) + cell] + old - 1;
  order[(size_t)b * dm.Npts + slot] = p;
}

// (row-major rank, point count) of every cell of sample b, from the CSR (test export)
__global__ void ls_export_cell_counts_kernel(const int* __restrict__ seg_start, LsGrid grid, int b,
                                              long long* __restrict__ out) {
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= grid.Vc) return;
  const int* seg = seg_start + (size_t)b * (grid.Vc + 1);
  const int tile = cell / LS_TILE, local = cell % LS_TILE;
  const int gx = (tile / grid.tiles_y) * LS_TX + local / LS_TY;
  const int gy = (tile % grid.tiles_y) * LS_TY + local % LS_TY;
  if (gx >= grid.X || gy >= grid.Y) return;
  const int k = seg[cell + 1] - seg[cell];
  if (k == 0) return;
  out[2 * (size_t)cell + 0] = (long long)gx * grid.Y + gy;
  out[2 * (size_t)cell + 1] = k;
}

// =====================================================================================
// softmax over depth (model/bev_model.py:64) - one thread per pixel, coalesced over fw
// =====================================================================================
template <typename T>
__global__ void __launch_bounds__(256)
ls_softmax_kernel(const T* __restrict__ logits, int D, int HW, T* __restrict__ prob) {
  const int img = blockIdx.y;
  const int rc = blockIdx.x * blockDim.x + threadIdx.x;
  if (rc >= HW) return;
  const T* src = logits + (size_t)img * D * HW + rc;
  T* dst = prob + (size_t)img * D * HW + rc;
  float m = -INFINITY;
  for (int d = 0; d < D; ++d) m = fmaxf(m, ls_to_float(src[(size_t)d * HW]));
  float s = 0.0f;
  for (int d = 0; d < D; ++d) s += expf(ls_to_float(src[(size_t)d * HW]) - m);
  for (int d = 0; d < D; ++d) dst[(size_t)d * HW] = ls_from_float<T>(__fdiv_rn(expf(ls_to_float(src[(size_t)d * HW]) - m), s));
}

// softmax backward: gl = prob * (g - sum_d prob*g),  g = grad_prob (+ ext)
template <typename T>
__global__ void __launch_bounds__(256)
ls_softmax_bwd_kernel(const T* __restrict__ prob, const float* __restrict__ gprob, const T* __restrict__ gext,
                      int D, int HW, T* __restrict__ glogits) {
  const int img = blockIdx.y;
  const int rc = blockIdx.x * blockDim.x + threadIdx.x;
  if (rc >= HW) return;
  const size_t base = (size_t)img * D * HW + rc;
  float dot = 0.0f;
  for (int d = 0; d < D; ++d) {
    const size_t o = base + (size_t)d * HW;
    float g = gprob[o];
    if (gext) g += ls_to_float(gext[o]);
    dot = fmaf(ls_to_float(prob[o]), g, dot);
  }
  for (int d = 0; d < D; ++d) {
    const size_t o = base + (size_t)d * HW;
    float g = gprob[o];
    if (gext) g += ls_to_float(gext[o]);
    glogits[o] = ls_from_float<T>(ls_to_float(prob[o]) * (g - dot));
  }
}

// =====================================================================================
// [img][R][S] -> [img][S][R] transpose through shared memory (both directions)
// =====================================================================================
template <typename T>
__global__ void __launch_bounds__(256)
ls_transpose_kernel(const T* __restrict__ src, int R, int S, T* __restrict__ dst) {
  __shared__ T tile[32][33];
  const int img = blockIdx.z;
  const int s0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const T* in = src + (size_t)img * R * S;
  T* out = dst + (size_t)img * R * S;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = r0 + j, s = s0 + threadIdx.x;
    if (r < R && s < S) tile[j][threadIdx.x] = in[(size_t)r * S + s];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int s = s0 + j, r = r0 + threadIdx.x;
    if (r < R && s < S) out[(size_t)s * R + r] = tile[threadIdx.x][j];
  }
}

// =====================================================================================
// K3: deterministic ranked segment-reduce splat
// reference: outer product model/bev_model.py:66, VoxelsSumming.forward
// tool/geometry.py:289-305, scatter + permute model/bev_model.py:101-105.
// One CTA per (sample, 8x32-voxel tile).  A warp owns a cell at a time: lanes = channel
// pairs, points of the cell are visited in ascending point id, each contributing
// prob[p] * feat[pix(p), :] from the NHWC feature copy (256 B coalesced per point).  The
// tile is assembled in shared memory [channel][cell] and written out as 128-byte rows of
// the [B,C,X,Y] tensor, zeros included (no separate memset of the BEV grid).
// =====================================================================================
template <typename T>
__global__ void __launch_bounds__(LS_THREADS)
ls_splat_fwd_kernel(const T* __restrict__ featT, const T* __restrict__ prob, const int* __restrict__ order,
                    const int* __restrict__ seg_start, int* __restrict__ order_tmp, LsDims dm, LsGrid grid,
                    float* __restrict__ bev, LsBevStrides st) {
  extern __shared__ float smem[];
  const int cc = dm.C < LS_CCHUNK ? dm.C : LS_CCHUNK;
  float* tile = smem;                                            // [cc][LS_TILE_PAD]
  int* seg = reinterpret_cast<int*>(smem + cc * LS_TILE_PAD);    // [LS_TILE + 1]
  int* scratch = seg + LS_TILE + 1;                              // [LS_WARPS][32]

  const int b = blockIdx.y, tile_id = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx0 = (tile_id / grid.tiles_y) * LS_TX, ty0 = (tile_id % grid.tiles_y) * LS_TY;
  const int* segg = seg_start + (size_t)b * (grid.Vc + 1) + (size_t)tile_id * LS_TILE;
  for (int i = tid; i <= LS_TILE; i += LS_THREADS) seg[i] = segg[i];
  __syncthreads();

  const int* ord = order + (size_t)b * dm.Npts;
  int* otmp = order_tmp + (size_t)b * dm.Npts;
  const T* pr = prob + (size_t)b * dm.Npts;
  const T* fb = featT + (size_t)b * dm.N * dm.HW * dm.C;
  int* wscr = scratch + warp * 32;
  const bool tile_empty = seg[LS_TILE] == seg[0];

  for (int cbase = 0; cbase < dm.C; cbase += LS_CCHUNK) {
    const bool lane_on = (cbase + 2 * lane) < dm.C;
    if (!tile_empty) {
      for (int cl = warp; cl < LS_TILE; cl += LS_WARPS) {
        const int s = seg[cl], k = seg[cl + 1] - s;
        float2 acc = make_float2(0.0f, 0.0f);
        if (k > 0) {
          const int* src = ord + s;
          if (k > 32) {
            // canonicalise a long segment: rank every id among the k ids of the cell
            for (int c0 = 0; c0 < k; c0 += 32) {
              const int mine = (c0 + lane < k) ? ord[s + c0 + lane] : 0x7fffffff;
              int r = 0;
              for (int c1 = 0; c1 < k; c1 += 32) {
                const int other = (c1 + lane < k) ? ord[s + c1 + lane] : 0x7fffffff;
                const int m = min(32, k - c1);
                for (int j = 0; j < m; ++j) r += (__shfl_sync(0xffffffffu, other, j) < mine) ? 1 : 0;
              }
              if (c0 + lane < k) otmp[s + r] = mine;
            }
            __syncwarp();
            src = otmp + s;
          }
          for (int c0 = 0; c0 < k; c0 += 32) {
            const int m = min(32, k - c0);
            int id = (lane < m) ? src[c0 + lane] : 0x7fffffff;
            if (k <= 32 && k > 1) {
              int r = 0;
              for (int j = 0; j < m; ++j) r += (__shfl_sync(0xffffffffu, id, j) < id) ? 1 : 0;
              if (lane < m) wscr[r] = id;
              __syncwarp();
              if (lane < m) id = wscr[lane];
              __syncwarp();
            }
            float w = 0.0f;
            int row = 0;                         // element offset of the point's pixel row in featT
            if (lane < m) {
              w = ls_to_float(pr[id]);
              const int n = id / dm.DHW;
              const int rc = id % dm.HW;
              row = (n * dm.HW + rc) * dm.C;
            }
#pragma unroll 4
            for (int j = 0; j < m; ++j) {
              const float wj = __shfl_sync(0xffffffffu, w, j);
              const int rj = __shfl_sync(0xffffffffu, row, j);
              if (lane_on) {
                const float2 f = ls_load2<T>(fb + rj + cbase + 2 * lane);
                acc.x = fmaf(wj, f.x, acc.x);
                acc.y = fmaf(wj, f.y, acc.y);
              }
            }
          }
        }
        if (2 * lane < cc) {
          tile[(2 * lane) * LS_TILE_PAD + cl] = acc.x;
          tile[(2 * lane + 1) * LS_TILE_PAD + cl] = acc.y;
        }
      }
    }
    __syncthreads();
    const int nch = min(cc, dm.C - cbase);
    for (int idx = tid; idx < nch * LS_TILE; idx += LS_THREADS) {
      const int c = idx / LS_TILE, cl = idx % LS_TILE;
      const int gx = tx0 + cl / LS_TY, gy = ty0 + cl % LS_TY;
      if (gx < grid.X && gy < grid.Y)
        bev[(size_t)b * st.b + (size_t)(cbase + c) * st.c + (size_t)gx * st.x + gy] =
            tile_empty ? 0.0f : tile[c * LS_TILE_PAD + cl];
    }
    __syncthreads();
  }
}

// =====================================================================================
// K4a: grad_bev [B,C,X,Y] -> cell-major gT [B, Vc, C] (tile transposes through smem)
// =====================================================================================
__global__ void __launch_bounds__(LS_THREADS)
ls_bwd_transpose_kernel(const float* __restrict__ gbev, LsBevStrides st, LsDims dm, LsGrid grid,
                        float* __restrict__ gT) {
  extern __shared__ float smem[];
  const int cc = dm.C < LS_CCHUNK ? dm.C : LS_CCHUNK;
  float* tile = smem;  // [cc][LS_TILE_PAD]
  const int b = blockIdx.y, tile_id = blockIdx.x, tid = threadIdx.x;
  const int tx0 = (tile_id / grid.tiles_y) * LS_TX, ty0 = (tile_id % grid.tiles_y) * LS_TY;
  float* dst = gT + ((size_t)b * grid.Vc + (size_t)tile_id * LS_TILE) * dm.C;
  for (int cbase = 0; cbase < dm.C; cbase += LS_CCHUNK) {
    const int nch = min(cc, dm.C - cbase);
    for (int idx = tid; idx < nch * LS_TILE; idx += LS_THREADS) {
      const int c = idx / LS_TILE, cl = idx % LS_TILE;
      const int gx = tx0 + cl / LS_TY, gy = ty0 + cl % LS_TY;
      float v = 0.0f;
      if (gx < grid.X && gy < grid.Y)
        v = gbev[(size_t)b * st.b + (size_t)(cbase + c) * st.c + (size_t)gx * st.x + gy];
      tile[c * LS_TILE_PAD + cl] = v;
    }
    __syncthreads();
    for (int idx = tid; idx < nch * LS_TILE; idx += LS_THREADS) {
      const int cl = idx / nch, c = idx % nch;
      dst[(size_t)cl * dm.C + cbase + c] = tile[c * LS_TILE_PAD + cl];
    }
    __syncthreads();
  }
}

// =====================================================================================
// K4b: gradient gather, pixel-stationary (deterministic, no atomics)
// reference: VoxelsSumming.backward tool/geometry.py:307-317 + autograd of
// model/bev_model.py:66,91-97.  One warp per pixel; lanes = channel pairs; depth bins in order.
// =====================================================================================
template <typename T, int NCH>
__global__ void __launch_bounds__(LS_THREADS)
ls_bwd_gather_kernel(const float* __restrict__ gT, const T* __restrict__ featT, const T* __restrict__ prob,
                     const int* __restrict__ rank, LsDims dm, LsGrid grid, float* __restrict__ gprob,
                     T* __restrict__ gfeatT) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pix = blockIdx.x * LS_WARPS + warp;          // over B*N*HW
  if (pix >= dm.B * dm.N * dm.HW) return;
  const int b = pix / (dm.N * dm.HW);
  const int rem = pix - b * (dm.N * dm.HW);
  const int n = rem / dm.HW, rc = rem - n * dm.HW;
  const T* frow = featT + (size_t)pix * dm.C;
  float2 f[NCH], gf[NCH];
  bool on[NCH];
#pragma unroll
  for (int q = 0; q < NCH; ++q) {
    on[q] = (q * LS_CCHUNK + 2 * lane) < dm.C;
    f[q] = on[q] ? ls_load2<T>(frow + q * LS_CCHUNK + 2 * lane) : make_float2(0.f, 0.f);
    gf[q] = make_float2(0.f, 0.f);
  }
  const size_t pbase = (size_t)b * dm.Npts + (size_t)n * dm.DHW + rc;
  const float* gTb = gT + (size_t)b * grid.Vc * dm.C;
  for (int d = 0; d < dm.D; ++d) {
    const size_t p = pbase + (size_t)d * dm.HW;
    const int r = rank[p];                       // warp-uniform
    float dot = 0.0f;
    if (r >= 0) {
      const float w = ls_to_float(prob[p]);
      const float* g = gTb + (size_t)ls_cell_of_rank(r, grid.Y, grid.tiles_y) * dm.C;
#pragma unroll
      for (int q = 0; q < NCH; ++q) {
        if (on[q]) {
          const float2 gv = *reinterpret_cast<const float2*>(g + q * LS_CCHUNK + 2 * lane);
          dot = fmaf(f[q].x, gv.x, dot);
          dot = fmaf(f[q].y, gv.y, dot);
          gf[q].x = fmaf(w, gv.x, gf[q].x);
          gf[q].y = fmaf(w, gv.y, gf[q].y);
        }
      }
      dot = ls_warp_sum(dot);
    }
    if (lane == 0) gprob[p] = dot;
  }
  T* grow = gfeatT + (size_t)pix * dm.C;
#pragma unroll
  for (int q = 0; q < NCH; ++q)
    if (on[q]) ls_store2<T>(grow + q * LS_CCHUNK + 2 * lane, gf[q]);
}

// =====================================================================================
// host side: C ABI
// =====================================================================================
#include <atomic>
#include <stdio.h>
#include <string.h>

static std::atomic<long long> g_launches{0};
static thread_local char g_cuda_err[256] = "";

#define LS_COUNT() g_launches.fetch_add(1, std::memory_order_relaxed)
#define LS_CUDA(call)                                                                  \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess) {                                                          \
      snprintf(g_cuda_err, sizeof(g_cuda_err), "%s at %s:%d", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return LS_ERR_CUDA;                                                              \
    }                                                                                  \
  } while (0)
#define LS_LAUNCHED() do { LS_COUNT(); LS_CUDA(cudaGetLastError()); } while (0)

static int ls_check_shape(const LsShape* s) {
  if (!s) return LS_ERR_BAD_ARG;
  if (s->B <= 0 || s->N <= 0 || s->D <= 0 || s->fh <= 0 || s->fw <= 0 || s->C <= 0) return LS_ERR_BAD_ARG;
  if (s->X <= 0 || s->Y <= 0 || s->Z <= 0) return LS_ERR_BAD_ARG;
  if ((long long)s->X * s->Y * s->Z >= (1LL << 30)) return LS_ERR_UNSUPPORTED;
  if ((long long)s->N * s->D * s->fh * s->fw >= (1LL << 30)) return LS_ERR_UNSUPPORTED;
  return LS_OK;
}
static int ls_check_splat_shape(const LsShape* s) {
  int rc = ls_check_shape(s);
  if (rc) return rc;
  if (s->Z != 1) return LS_ERR_UNSUPPORTED;   // reference: squeeze(0) needs Z == 1 (bev_model.py:104)
  if (s->C % 2) return LS_ERR_BAD_ARG;
  if (s->C > 4 * LS_CCHUNK) return LS_ERR_UNSUPPORTED;
  return LS_OK;
}

extern "C" {

const char* ls_version(void) { return "ls_b200 0.1 (sm_100a)"; }

const char* ls_strerror(int status) {
  switch (status) {
    case LS_OK: return "ok";
    case LS_ERR_BAD_ARG: return "bad argument";
    case LS_ERR_UNSUPPORTED: return "unsupported configuration";
    case LS_ERR_WORKSPACE: return "workspace too small";
    case LS_ERR_CUDA: return "CUDA error";
    default: return "unknown status";
  }
}
const char* ls_last_cuda_error(void) { return g_cuda_err; }
int64_t ls_launch_count(void) { return (int64_t)g_launches.load(); }

int ls_grid_cells(const LsShape* s, int32_t* tiles, int32_t* cells_padded) {
  int rc = ls_check_shape(s);
  if (rc) return rc;
  LsGrid g = ls_grid(s);
  if (tiles) *tiles = g.tiles;
  if (cells_padded) *cells_padded = g.Vc;
  return LS_OK;
}

int ls_camera_transform(const float* intrinsics, const float* extrinsics, int32_t BN, float* M, float* t,
                        ls_stream_t stream) {
  if (!intrinsics || !extrinsics || !M || !t || BN <= 0) return LS_ERR_BAD_ARG;
  ls_camera_transform_kernel<<<(BN + 63) / 64, 64, 0, (cudaStream_t)stream>>>(intrinsics, extrinsics, BN, M, t);
  LS_LAUNCHED();
  return LS_OK;
}

static int ls_launch_index(const float* M, const float* t, const float* frustum, const LsShape* s, int32_t* rank,
                           int32_t* counts, float* geom, int64_t* vox, uint8_t* keep, int64_t* rank64,
                           bool do_export, cudaStream_t stream) {
  int rc = ls_check_shape(s);
  if (rc) return rc;
  if (!M || !t || !frustum) return LS_ERR_BAD_ARG;
  LsDims dm = ls_dims(s);
  LsGrid g = ls_grid(s);
  if (counts && s->Z != 1) return LS_ERR_UNSUPPORTED;
  dim3 grid((dm.DHW + 255) / 256, dm.N, dm.B);
  if (do_export)
    ls_index_kernel<true><<<grid, 256, 0, stream>>>(M, t, frustum, dm, g, nullptr, nullptr, geom, (long long*)vox,
                                                    keep, (long long*)rank64);
  else
    ls_index_kernel<false><<<grid, 256, 0, stream>>>(M, t, frustum, dm, g, rank, counts, nullptr, nullptr, nullptr,
                                                     nullptr);
  LS_LAUNCHED();
  return LS_OK;
}

int ls_geometry(const float* M, const float* t, const float* frustum, const LsShape* s, float* geom,
                ls_stream_t stream) {
  if (!geom) return LS_ERR_BAD_ARG;
  return ls_launch_index(M, t, frustum, s, nullptr, nullptr, geom, nullptr, nullptr, nullptr, true,
                         (cudaStream_t)stream);
}

int ls_index(const float* M, const float* t, const float* frustum, const LsShape* s, int32_t* rank,
             int32_t* counts, ls_stream_t stream) {
  if (!rank && !counts) return LS_ERR_BAD_ARG;
  return ls_launch_index(M, t, frustum, s, rank, counts, nullptr, nullptr, nullptr, nullptr, false,
                         (cudaStream_t)stream);
}

int ls_export_indices(const float* M, const float* t, const float* frustum, const LsShape* s, int64_t* vox,
                      uint8_t* keep, int64_t* rank, ls_stream_t stream) {
  return ls_launch_index(M, t, frustum, s, nullptr, nullptr, nullptr, vox, keep, rank, true, (cudaStream_t)stream);
}

int ls_sort(const int32_t* rank, const LsShape* s, int32_t* counts, int have_hist, int32_t* seg_start,
            int32_t* order, ls_stream_t stream_) {
  int rc = ls_check_splat_shape(s);
  if (rc) return rc;
  if (!rank || !counts || !seg_start || !order) return LS_ERR_BAD_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  LsDims dm = ls_dims(s);
  LsGrid g = ls_grid(s);
  dim3 pgrid((dm.Npts + 255) / 256, dm.B);
  if (!have_hist) {
    LS_CUDA(cudaMemsetAsync(counts, 0, (size_t)dm.B * g.Vc * sizeof(int), stream));
    LS_COUNT();
    ls_hist_kernel<<<pgrid, 256, 0, stream>>>(rank, dm, g, counts);
    LS_LAUNCHED();
  }
  ls_scan_kernel<<<dm.B, 1024, 0, stream>>>(counts, g.Vc, seg_start);
  LS_LAUNCHED();
  ls_place_kernel<<<pgrid, 256, 0, stream>>>(rank, dm, g, counts, seg_start, order);
  LS_LAUNCHED();
  return LS_OK;
}

int ls_export_cell_counts(const int32_t* seg_start, const LsShape* s, int32_t b, int64_t* out, int32_t* kept,
                          ls_stream_t stream) {
  int rc = ls_check_splat_shape(s);
  if (rc) return rc;
  if (!seg_start || b < 0 || b >= s->B) return LS_ERR_BAD_ARG;
  LsGrid g = ls_grid(s);
  if (out) {
    LS_CUDA(cudaMemsetAsync(out, 0, (size_t)g.Vc * 2 * sizeof(int64_t), (cudaStream_t)stream));
    LS_COUNT();
    ls_export_cell_counts_kernel<<<(g.Vc + 255) / 256, 256, 0, (cudaStream_t)stream>>>(seg_start, g, b,
                                                                                        (long long*)out);
    LS_LAUNCHED();
  }
  if (kept) {
    // kept[b'] = seg_start[b'][Vc] for every sample (strided device->device copy)
    LS_CUDA(cudaMemcpy2DAsync(kept, sizeof(int), seg_start + g.Vc, (size_t)(g.Vc + 1) * sizeof(int), sizeof(int),
                              s->B, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    LS_COUNT();
  }
  return LS_OK;
}

int ls_softmax(const void* logits, int dtype, const LsShape* s, void* prob, ls_stream_t stream) {
  int rc = ls_check_shape(s);
  if (rc) return rc;
  if (!logits || !prob) return LS_ERR_BAD_ARG;
  LsDims dm = ls_dims(s);
  dim3 grid((dm.HW + 255) / 256, dm.B * dm.N);
  if (dtype == LS_F32)
    ls_softmax_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)logits, dm.D, dm.HW, (float*)prob);
  else if (dtype == LS_BF16)
    ls_softmax_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)logits, dm.D, dm.HW,
                                                                             (__nv_bfloat16*)prob);
  else
    return LS_ERR_BAD_ARG;
  LS_LAUNCHED();
  return LS_OK;
}

int ls_softmax_bwd(const void* prob, const float* grad_prob, const void* grad_prob_ext, int dtype, const LsShape* s,
                   void* grad_logits, ls_stream_t stream) {
  int rc = ls_check_shape(s);
  if (rc) return rc;
  if (!prob || !grad_prob || !grad_logits) return LS_ERR_BAD_ARG;
  LsDims dm = ls_dims(s);
  dim3 grid((dm.HW + 255) / 256, dm.B * dm.N);
  if (dtype == LS_F32)
    ls_softmax_bwd_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)prob, grad_prob,
                                                                         (const float*)grad_prob_ext, dm.D, dm.HW,
                                                                         (float*)grad_logits);
  else if (dtype == LS_BF16)
    ls_softmax_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)prob, grad_prob, (const __nv_bfloat16*)grad_prob_ext, dm.D, dm.HW,
        (__nv_bfloat16*)grad_logits);
  else
    return LS_ERR_BAD_ARG;
  LS_LAUNCHED();
  return LS_OK;
}

static int ls_transpose(const void* src, int dtype, int images, int R, int S, void* dst, cudaStream_t stream) {
  if (!src || !dst || images <= 0 || R <= 0 || S <= 0) return LS_ERR_BAD_ARG;
  dim3 grid((S + 31) / 32, (R + 31) / 32, images), block(32, 8);
  if (dtype == LS_F32)
    ls_transpose_kernel<float><<<grid, block, 0, stream>>>((const float*)src, R, S, (float*)dst);
  else if (dtype == LS_BF16)
    ls_transpose_kernel<__nv_bfloat16><<<grid, block, 0, stream>>>((const __nv_bfloat16*)src, R, S, (__nv_bfloat16*)dst);
  else
    return LS_ERR_BAD_ARG;
  LS_LAUNCHED();
  return LS_OK;
}

int ls_nchw_to_nhwc(const void* src, int dtype, int32_t images, int32_t C, int32_t HW, void* dst,
                    ls_stream_t stream) {
  return ls_transpose(src, dtype, images, C, HW, dst, (cudaStream_t)stream);
}
int ls_nhwc_to_nchw(const void* src, int dtype, int32_t images, int32_t C, int32_t HW, void* dst,
                    ls_stream_t stream) {
  return ls_transpose(src, dtype, images, HW, C, dst, (cudaStream_t)stream);
}

static size_t ls_tile_smem(const LsDims& dm) {
  const int cc = dm.C < LS_CCHUNK ? dm.C : LS_CCHUNK;
  return (size_t)cc * LS_TILE_PAD * sizeof(float) + (LS_TILE + 1 + LS_WARPS * 32) * sizeof(int);
}

int ls_splat_fwd(const void* feat_nhwc, const void* prob, int dtype, const int32_t* order, const int32_t* seg_start,
                 int32_t* order_tmp, const LsShape* s, float* bev, const LsBevStrides* st, ls_stream_t stream_) {
  int rc = ls_check_splat_shape(s);
  if (rc) return rc;
  if (!feat_nhwc || !prob || !order || !seg_start || !order_tmp || !bev || !st) return LS_ERR_BAD_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  LsDims dm = ls_dims(s);
  LsGrid g = ls_grid(s);
  const size_t smem = ls_tile_smem(dm);
  dim3 grid(g.tiles, dm.B);
  static bool attr_done = false;
  if (!attr_done) {
    LsDims big = dm; big.C = LS_CCHUNK;
    const int max_smem = (int)ls_tile_smem(big);
    LS_CUDA(cudaFuncSetAttribute(ls_splat_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    LS_CUDA(cudaFuncSetAttribute(ls_splat_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    LS_CUDA(cudaFuncSetAttribute(ls_bwd_transpose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    attr_done = true;
  }
  if (dtype == LS_F32) {
    ls_splat_fwd_kernel<float><<<grid, LS_THREADS, smem, stream>>>((const float*)feat_nhwc, (const float*)prob, order,
                                                                   seg_start, order_tmp, dm, g, bev, *st);
  } else if (dtype == LS_BF16) {
    ls_splat_fwd_kernel<__nv_bfloat16><<<grid, LS_THREADS, smem, stream>>>(
        (const __nv_bfloat16*)feat_nhwc, (const __nv_bfloat16*)prob, order, seg_start, order_tmp, dm, g, bev, *st);
  } else {
    return LS_ERR_BAD_ARG;
  }
  LS_LAUNCHED();
  return LS_OK;
}

}  // extern "C"

template <typename T>
static int ls_launch_gather(const float* gT, const void* featT, const void* prob, const int32_t* rank,
                            const LsDims& dm, const LsGrid& g, float* gprob, void* gfeatT, cudaStream_t stream) {
  const int nch = (dm.C + LS_CCHUNK - 1) / LS_CCHUNK;
  const int pixels = dm.B * dm.N * dm.HW;
  dim3 grid((pixels + LS_WARPS - 1) / LS_WARPS);
#define LS_GATHER(NCH)                                                                                      \
  ls_bwd_gather_kernel<T, NCH><<<grid, LS_THREADS, 0, stream>>>(gT, (const T*)featT, (const T*)prob, rank, dm, g, \
                                                                gprob, (T*)gfeatT)
  switch (nch) {
    case 1: LS_GATHER(1); break;
    case 2: LS_GATHER(2); break;
    case 3: LS_GATHER(3); break;
    case 4: LS_GATHER(4); break;
    default: return LS_ERR_UNSUPPORTED;
  }
#undef LS_GATHER
  LS_LAUNCHED();
  return LS_OK;
}

extern "C" {

int ls_splat_bwd(const float* grad_bev, const LsBevStrides* gst, const void* feat_nhwc, const void* prob, int dtype,
                 const int32_t* rank, const LsShape* s, float* gT_ws, float* grad_prob, void* grad_feat_nhwc,
                 ls_stream_t stream_) {
  int rc = ls_check_splat_shape(s);
  if (rc) return rc;
  if (!grad_bev || !gst || !feat_nhwc || !prob || !rank || !gT_ws || !grad_prob || !grad_feat_nhwc)
    return LS_ERR_BAD_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  LsDims dm = ls_dims(s);
  LsGrid g = ls_grid(s);
  const int cc = dm.C < LS_CCHUNK ? dm.C : LS_CCHUNK;
  const size_t smem = (size_t)cc * LS_TILE_PAD * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    LS_CUDA(cudaFuncSetAttribute(ls_bwd_transpose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(LS_CCHUNK * LS_TILE_PAD * sizeof(float))));
    attr_done = true;
  }
  ls_bwd_transpose_kernel<<<dim3(g.tiles, dm.B), LS_THREADS, smem, stream>>>(grad_bev, *gst, dm, g, gT_ws);
  LS_LAUNCHED();
  if (dtype == LS_F32) return ls_launch_gather<float>(gT_ws, feat_nhwc, prob, rank, dm, g, grad_prob, grad_feat_nhwc, stream);
  if (dtype == LS_BF16)
    return ls_launch_gather<__nv_bfloat16>(gT_ws, feat_nhwc, prob, rank, dm, g, grad_prob, grad_feat_nhwc, stream);
  return LS_ERR_BAD_ARG;
}

// ---- workspace carving -------------------------------------------------------------
struct LsWs {
  int32_t *rank, *counts, *seg_start, *order, *order_tmp;
  void* featT;
  float *gT, *gprob;
  void* gfeatT;
  size_t bytes;
};
static inline size_t ls_align(size_t v) { return (v + 255) & ~(size_t)255; }
static LsWs ls_carve(const LsShape* s, int dtype, int with_backward, void* base) {
  LsDims dm = ls_dims(s);
  LsGrid g = ls_grid(s);
  const size_t es = dtype == LS_BF16 ? 2 : 4;
  char* p = (char*)base;
  size_t off = 0;
  LsWs w;
  auto take = [&](size_t n) { void* r = p ? (void*)(p + off) : nullptr; off += ls_align(n); return r; };
  w.rank = (int32_t*)take((size_t)dm.B * dm.Npts * 4);
  w.featT = take((size_t)dm.B * dm.N * dm.HW * dm.C * es);
  w.counts = (int32_t*)take((size_t)dm.B * g.Vc * 4);
  w.seg_start = (int32_t*)take((size_t)dm.B * (g.Vc + 1) * 4);
  w.order = (int32_t*)take((size_t)dm.B * dm.Npts * 4);
  w.order_tmp = (int32_t*)take((size_t)dm.B * dm.Npts * 4);
  w.gT = nullptr; w.gprob = nullptr; w.gfeatT = nullptr;
  if (with_backward) {
    w.gT = (float*)take((size_t)dm.B * g.Vc * dm.C * 4);
    w.gprob = (float*)take((size_t)dm.B * dm.Npts * 4);
    w.gfeatT = take((size_t)dm.B * dm.N * dm.HW * dm.C * es);
  }
  w.bytes = off;
  return w;
}

size_t ls_workspace_bytes(const LsShape* s, int dtype, int with_backward) {
  if (ls_check_splat_shape(s)) return 0;
  return ls_carve(s, dtype, with_backward, nullptr).bytes;
}

int ls_forward(const void* feat, const void* logits, int dtype, const float* M, const float* t, const float* frustum,
               const LsShape* s, void* ws, size_t ws_bytes, float* bev, const LsBevStrides* bev_strides, void* prob,
               ls_stream_t stream) {
  int rc = ls_check_splat_shape(s);
  if (rc) return rc;
  if (!feat || !logits || !M || !t || !frustum || !ws || !bev || !bev_strides || !prob) return LS_ERR_BAD_ARG;
  if (dtype != LS_F32 && dtype != LS_BF16) return LS_ERR_BAD_ARG;
  if (ws_bytes < ls_carve(s, dtype, 0, nullptr).bytes) return LS_ERR_WORKSPACE;
  LsWs w = ls_carve(s, dtype, 0, ws);
  LsDims dm = ls_dims(s);
  LsGrid g = ls_grid(s);
  LS_CUDA(cudaMemsetAsync(w.counts, 0, (size_t)dm.B * g.Vc * sizeof(int), (cudaStream_t)stream));
  LS_COUNT();
  if ((rc = ls_index(M, t, frustum, s, w.rank, w.counts, stream))) return rc;
  if ((rc = ls_sort(w.rank, s, w.counts, 1, w.seg_start, w.order, stream))) return rc;
  if ((rc = ls_softmax(logits, dtype, s, prob, stream))) return rc;
  if ((rc = ls_nchw_to_nhwc(feat, dtype, dm.B * dm.N, dm.C, dm.HW, w.featT, stream))) return rc;
  return ls_splat_fwd(w.featT, prob, dtype, w.order, w.seg_start, w.order_tmp, s, bev, bev_strides, stream);
}

int ls_backward(const float* grad_bev, const LsBevStrides* grad_strides, const void* grad_prob_ext, const void* prob,
                int dtype, const LsShape* s, void* ws, size_t ws_bytes, void* grad_feat, void* grad_logits,
                ls_stream_t stream) {
  int rc = ls_check_splat_shape(s);
  if (rc) return rc;
  if (!grad_bev || !grad_strides || !prob || !ws || !grad_feat || !grad_logits) return LS_ERR_BAD_ARG;
  if (dtype != LS_F32 && dtype != LS_BF16) return LS_ERR_BAD_ARG;
  if (ws_bytes < ls_carve(s, dtype, 1, nullptr).bytes) return LS_ERR_WORKSPACE;
  LsWs w = ls_carve(s, dtype, 1, ws);
  LsDims dm = ls_dims(s);
  if ((rc = ls_splat_bwd(grad_bev, grad_strides, w.featT, prob, dtype, w.rank, s, w.gT, w.gprob, w.gfeatT, stream)))
    return rc;
  if ((rc = ls_nhwc_to_nchw(w.gfeatT, dtype, dm.B * dm.N, dm.C, dm.HW, grad_feat, stream))) return rc;
  return ls_softmax_bwd(prob, w.gprob, grad_prob_ext, dtype, s, grad_logits, stream);
}

}  // extern "C"
