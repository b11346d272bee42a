// The splat itself: deterministic ranked segment-reduce (forward) and its gradient
// (cell-major gradient staging + pixel-stationary gather).  Reference semantics:
// model/bev_model.py:66-72,99-105 and VoxelsSumming (tool/geometry.py:285-317).
#include "ls_internal.h"

#define LS_WIN 8   // points whose feature rows a half-warp keeps in flight

__device__ __forceinline__ unsigned ls_half_mask() { return 0xFFFFu << (threadIdx.x & 16); }

struct LsTileGeom {
  int cc;       // channels of this pass (<= 64, multiple of 4)
  int nqp;      // quads per smem row rounded to a power of two
  int stride;   // smem row stride in floats = 4*nqp + 4
};
__host__ __device__ __forceinline__ LsTileGeom ls_tile_geom(int Cp) {
  LsTileGeom t;
  t.cc = Cp < LS_CCHUNK ? Cp : LS_CCHUNK;
  int nq = t.cc / 4;
  t.nqp = 1;
  while (t.nqp < nq) t.nqp <<= 1;
  t.stride = 4 * t.nqp + 4;
  return t;
}

// =====================================================================================
// K3: forward splat.  One CTA per (sample, 16x16-voxel tile), three phases:
//  A  canonicalise: one thread per point record of the tile; its position inside its cell =
//     number of records of that cell with a smaller key (keys are unique), which makes the
//     summation order independent of the atomics that placed the records.  The re-ordered
//     records {pixel | cell_in_tile<<20 | last_of_cell<<28, prob} go to a scratch array.
//  B  reduce: a half-warp owns 16 consecutive cells (one x-row of the tile) = one contiguous
//     run of records; it streams them with LS_WIN feature rows (16 B per lane, 256 B per point)
//     in flight, accumulates prob*feat in registers and drops the sum into the shared-memory
//     tile [cell][channel] (swizzled, conflict-free) when a record carries the last-of-cell flag.
//  C  write-out: the tile is read column-wise and written as 16-byte pieces of the
//     [B,C,X,Y] tensor, zeros included - the BEV grid is never memset.
// =====================================================================================
#define LS_REC_PIX_MASK 0xFFFFF
#define LS_REC_LAST (1 << 28)

#define LS_ITEM 64   // target records per work item of phase B

__device__ __forceinline__ void ls_load_recs(const int2* __restrict__ rso, int i, int n, int2 (&r)[LS_WIN]) {
#pragma unroll
  for (int u = 0; u < LS_WIN; ++u) {
    r[u] = rso[i + min(u, n - 1)];                          // half-warp-uniform 8-byte loads
    if (u >= n) { r[u].x &= ~LS_REC_LAST; r[u].y = 0; }     // padding: weight 0, never flushes
  }
}

// kCC = 64: the common case (Cp == 64) with compile-time tile geometry; kCC = 0: any Cp.
template <typename T, bool VEC4, int kCC>
__global__ void __launch_bounds__(LS_THREADS, 3)
ls_splat_fwd_kernel(const T* __restrict__ featT, const int2* __restrict__ recs, const int* __restrict__ seg_start,
                    const int* __restrict__ tile_order, int2* __restrict__ recs_sorted, LsDims dm, LsGrid grid,
                    float* __restrict__ bev, LsBevStrides st) {
  extern __shared__ float smem[];
  const LsTileGeom tgr = ls_tile_geom(dm.Cp);
  const int Cp = kCC ? kCC : dm.Cp;
  const int stride = kCC ? kCC + 4 : tgr.stride;
  const int nqp = kCC ? kCC / 4 : tgr.nqp;
  const int ccmax = kCC ? kCC : tgr.cc;
  float* tile = smem;                                            // [LS_TILE][stride]
  int* seg = reinterpret_cast<int*>(smem + LS_TILE * stride);    // [LS_TILE + 1]
  int* heads = seg + LS_TILE + 1;                                // [LS_TILE + 1] first cell of each work item
  int* ctl = heads + LS_TILE + 1;                                // [0] item count, [1] next item, [2..9] warp sums

  // heaviest tiles first, all samples interleaved: blockIdx.x = order_index * B + b
  const int b = blockIdx.x % dm.B;
  const int tile_id = tile_order[(size_t)b * grid.tiles + blockIdx.x / dm.B];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx0 = (tile_id / grid.tiles_y) * LS_TX, ty0 = (tile_id % grid.tiles_y) * LS_TY;
  const int* segg = seg_start + (size_t)b * grid.seg_stride + (size_t)tile_id * LS_TILE;
  for (int i = tid; i <= LS_TILE; i += LS_THREADS) seg[i] = segg[i];
  __syncthreads();
  const int s0 = seg[0], s1 = seg[LS_TILE];
  const bool tile_empty = (s0 == s1);
  const int2* rin = recs + (size_t)b * dm.Npts;
  int2* rso = recs_sorted + (size_t)b * dm.Npts;

  if (!tile_empty) {
    // ---- phase A --------------------------------------------------------------------
    for (int i = s0 + tid; i < s1; i += LS_THREADS) {
      const int2 r = rin[i];
      const int cl = (unsigned)r.x >> 24;
      const int a = seg[cl], e = seg[cl + 1];
      int pos = a;
      for (int j = a; j < e; ++j) pos += (__ldg(&rin[j].x) < r.x) ? 1 : 0;
      const int pix = (r.x & 0xFFFFFF) >> dm.dbits;
      rso[pos] = make_int2(pix | (cl << 20) | (pos == e - 1 ? LS_REC_LAST : 0), r.y);
    }
    // ---- work items: runs of whole cells of about LS_ITEM records ----------------------
    // cell `tid` opens an item when its first record falls in a new LS_ITEM-sized bucket
    {
      const int id = (seg[tid] - s0) / LS_ITEM;
      const bool head = (tid == 0) || (id != (seg[tid - 1] - s0) / LS_ITEM);
      const unsigned bal = __ballot_sync(0xffffffffu, head);
      if (lane == 0) ctl[2 + warp] = __popc(bal);
      __syncthreads();
      int before = 0;
      for (int w = 0; w < warp; ++w) before += ctl[2 + w];
      if (head) heads[before + __popc(bal & ((1u << lane) - 1))] = tid;
      if (tid == LS_THREADS - 1) {
        const int nitems = before + __popc(bal);
        heads[nitems] = LS_TILE;
        ctl[0] = nitems;
      }
    }
  }

  const int hl = tid & 15;
  const unsigned hmask = ls_half_mask();
  const T* fbase = featT + (size_t)b * dm.N * dm.HW * Cp;
  const unsigned row_bytes = (unsigned)(Cp * sizeof(T));
  // phase C geometry of this thread (fixed): 4 consecutive y, one channel of a quad, one x-row
  const int y4 = tid & 3, cq = (tid >> 2) & 3, xr = (tid >> 4) & 15;
  const int gx = tx0 + xr, gy = ty0 + 4 * y4;
  const bool inb = gx < grid.X && gy < grid.Y;
  const int clc = xr * LS_TY + 4 * y4;

  for (int cbase = 0; cbase < Cp; cbase += LS_CCHUNK) {
    const int cc = kCC ? kCC : min(ccmax, Cp - cbase);
    const int nquads = cc >> 2;
    if (!tile_empty) {
      // zero the tile: cells nobody hits are never touched by phase B
      for (int i = tid; i < LS_TILE * stride / 4; i += LS_THREADS)
        reinterpret_cast<float4*>(tile)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tid == 0) ctl[1] = 0;
    }
    __syncthreads();   // also orders phase A's scratch records / work items before phase B
    // ---- phase B --------------------------------------------------------------------
    if (!tile_empty) {
      const bool lane_on = 4 * hl < cc;
      // lanes beyond the channel count read lane 0's (valid) bytes and never store
      const char* fbytes = reinterpret_cast<const char*>(fbase + cbase + (lane_on ? 4 * hl : 0));
      const int nitems = ctl[0];
      for (;;) {
        int it = 0;
        if (hl == 0) it = atomicAdd(&ctl[1], 1);
        it = __shfl_sync(hmask, it, 0, 16);
        if (it >= nitems) break;
        int i = seg[heads[it]];
        const int iend = seg[heads[it + 1]];
        if (i >= iend) continue;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int2 r[LS_WIN];
        ls_load_recs(rso, i, min(LS_WIN, iend - i), r);
        while (i < iend) {
          float4 f[LS_WIN];
#pragma unroll
          for (int u = 0; u < LS_WIN; ++u)
            f[u] = ls_load4<T>(reinterpret_cast<const T*>(
                fbytes + (unsigned long long)(unsigned)(r[u].x & LS_REC_PIX_MASK) * row_bytes));
          // records of the next window are fetched while this window's feature rows are in flight
          i += LS_WIN;
          int2 rn[LS_WIN];
          if (i < iend) ls_load_recs(rso, i, min(LS_WIN, iend - i), rn);
#pragma unroll
          for (int u = 0; u < LS_WIN; ++u) {
            const float w = __int_as_float(r[u].y);
            acc.x = fmaf(w, f[u].x, acc.x);
            acc.y = fmaf(w, f[u].y, acc.y);
            acc.z = fmaf(w, f[u].z, acc.z);
            acc.w = fmaf(w, f[u].w, acc.w);
            if (r[u].x & LS_REC_LAST) {
              const int cl = (r[u].x >> 20) & 255;
              if (lane_on) *reinterpret_cast<float4*>(tile + cl * stride + 4 * ls_tile_quad(cl, hl, nqp)) = acc;
              acc = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
#pragma unroll
          for (int u = 0; u < LS_WIN; ++u) r[u] = rn[u];
        }
      }
    }
    __syncthreads();
    // ---- phase C --------------------------------------------------------------------
    if (VEC4) {
      if (inb) {
        const int swz4 = 4 * ((clc >> 3) & (nqp - 1));
        const float* srow = tile + clc * stride + cq;
        float* gptr = bev + (size_t)b * st.b + (size_t)(cbase + cq) * st.c + (size_t)gx * st.x + gy;
        const size_t qstep = (size_t)4 * st.c;
#pragma unroll 8
        for (int q = 0; q < nquads; ++q) {
          if (kCC || cbase + 4 * q + cq < dm.C) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!tile_empty) {
              const float* src = srow + ((4 * q) ^ swz4);
              v.x = src[0]; v.y = src[stride]; v.z = src[2 * stride]; v.w = src[3 * stride];
            }
            *reinterpret_cast<float4*>(gptr) = v;
          }
          gptr += qstep;
        }
      }
    } else {
      for (int idx = tid; idx < 4 * nquads * LS_TILE; idx += LS_THREADS) {
        const int y = idx & 15, x = (idx >> 4) & 15, cr = idx >> 8;
        const int c = cbase + cr;
        const int ox = tx0 + x, oy = ty0 + y;
        if (c < dm.C && ox < grid.X && oy < grid.Y) {
          const int cl = x * LS_TY + y;
          const float v = tile_empty ? 0.0f : tile[cl * stride + 4 * ls_tile_quad(cl, cr >> 2, nqp) + (cr & 3)];
          bev[(size_t)b * st.b + (size_t)c * st.c + (size_t)ox * st.x + oy] = v;
        }
      }
    }
    __syncthreads();
  }
}

static size_t ls_tile_smem_bytes(const LsDims& dm) {
  const LsTileGeom tg = ls_tile_geom(dm.Cp);
  return (size_t)LS_TILE * tg.stride * sizeof(float) + (2 * (LS_TILE + 1) + 16) * sizeof(int);
}
static size_t ls_tile_smem_max() {
  LsDims d;
  d.Cp = LS_CCHUNK;
  return ls_tile_smem_bytes(d);
}

static bool ls_bev_vec4(const float* p, const LsBevStrides& st, const LsGrid& g) {
  return ((uintptr_t)p % 16 == 0) && (st.b % 4 == 0) && (st.c % 4 == 0) && (st.x % 4 == 0) && (g.Y % 4 == 0);
}

template <typename T>
static int ls_splat_dispatch(const void* featT, const int2* recs, const int* seg_start, const int* tile_order,
                             int2* recs_sorted, const LsDims& dm, const LsGrid& g, float* bev, const LsBevStrides& st, cudaStream_t s) {
  static bool attr_done = false;
  if (!attr_done) {
    const int m = (int)ls_tile_smem_max();
    LS_CUDA(cudaFuncSetAttribute(ls_splat_fwd_kernel<T, true, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, m));
    LS_CUDA(cudaFuncSetAttribute(ls_splat_fwd_kernel<T, true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, m));
    LS_CUDA(cudaFuncSetAttribute(ls_splat_fwd_kernel<T, false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, m));
    attr_done = true;
  }
  const size_t smem = ls_tile_smem_bytes(dm);
  dim3 grid(g.tiles * dm.B);
  const bool v4 = ls_bev_vec4(bev, st, g);
  if (v4 && dm.Cp == 64 && dm.C == 64)
    ls_splat_fwd_kernel<T, true, 64><<<grid, LS_THREADS, smem, s>>>((const T*)featT, recs, seg_start, tile_order, recs_sorted,
                                                                   dm, g, bev, st);
  else if (v4)
    ls_splat_fwd_kernel<T, true, 0><<<grid, LS_THREADS, smem, s>>>((const T*)featT, recs, seg_start, tile_order, recs_sorted,
                                                                  dm, g, bev, st);
  else
    ls_splat_fwd_kernel<T, false, 0><<<grid, LS_THREADS, smem, s>>>((const T*)featT, recs, seg_start, tile_order, recs_sorted,
                                                                   dm, g, bev, st);
  LS_LAUNCHED();
  return LS_OK;
}

int ls_launch_splat_fwd(const void* featT, int dtype, const int2* recs, const int* seg_start, const int* tile_order,
                        int2* recs_sorted, const LsDims& dm, const LsGrid& g, float* bev, const LsBevStrides& st,
                        cudaStream_t s) {
  if (dtype == LS_F32)
    return ls_splat_dispatch<float>(featT, recs, seg_start, tile_order, recs_sorted, dm, g, bev, st, s);
  return ls_splat_dispatch<__nv_bfloat16>(featT, recs, seg_start, tile_order, recs_sorted, dm, g, bev, st, s);
}

// =====================================================================================
// K4a: grad_bev [B,C,X,Y] -> cell-major gT [B, Vc, Cp]; rows of cells nobody hit are skipped
// (they are never read).  Same tile / swizzle as the forward write-out, run backwards.
// =====================================================================================
template <bool VEC4>
__global__ void __launch_bounds__(LS_THREADS, 3)
ls_bwd_transpose_kernel(const float* __restrict__ gbev, LsBevStrides st, const int* __restrict__ seg_start,
                        LsDims dm, LsGrid grid, float* __restrict__ gT) {
  extern __shared__ float smem[];
  const LsTileGeom tg = ls_tile_geom(dm.Cp);
  float* tile = smem;
  int* seg = reinterpret_cast<int*>(smem + LS_TILE * tg.stride);
  const int b = blockIdx.y, tile_id = blockIdx.x, tid = threadIdx.x;
  const int tx0 = (tile_id / grid.tiles_y) * LS_TX, ty0 = (tile_id % grid.tiles_y) * LS_TY;
  const int* segg = seg_start + (size_t)b * grid.seg_stride + (size_t)tile_id * LS_TILE;
  for (int i = tid; i <= LS_TILE; i += LS_THREADS) seg[i] = segg[i];
  if (tile_id == 0) {   // row Vc of every sample = zeros: where dropped points gather from
    float* zrow = gT + ((size_t)b * (grid.Vc + 1) + grid.Vc) * dm.Cp;
    for (int i = tid; i < dm.Cp; i += LS_THREADS) zrow[i] = 0.0f;
  }
  __syncthreads();
  if (seg[0] == seg[LS_TILE]) return;          // nobody reads this tile's gradient
  float* dst = gT + ((size_t)b * (grid.Vc + 1) + (size_t)tile_id * LS_TILE) * dm.Cp;
  const int y4 = tid & 3, cq = (tid >> 2) & 3, xr = (tid >> 4) & 15;
  const int gx = tx0 + xr, gy = ty0 + 4 * y4;
  const bool inb = gx < grid.X && gy < grid.Y;
  const int clc = xr * LS_TY + 4 * y4;
  for (int cbase = 0; cbase < dm.Cp; cbase += LS_CCHUNK) {
    const int cc = min(tg.cc, dm.Cp - cbase);
    const int nquads = cc >> 2;
    if (VEC4) {
      const int swz = (clc >> 3) & (tg.nqp - 1);
      float* drow = tile + clc * tg.stride + cq;
      const float* gptr = gbev + (size_t)b * st.b + (size_t)(cbase + cq) * st.c + (size_t)gx * st.x + gy;
      const size_t qstep = (size_t)4 * st.c;
#pragma unroll 4
      for (int q = 0; q < nquads; ++q) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (inb && cbase + 4 * q + cq < dm.C) v = __ldg(reinterpret_cast<const float4*>(gptr + q * qstep));
        float* d = drow + 4 * (q ^ swz);
        d[0] = v.x; d[tg.stride] = v.y; d[2 * tg.stride] = v.z; d[3 * tg.stride] = v.w;
      }
    } else {
      for (int idx = tid; idx < 4 * nquads * LS_TILE; idx += LS_THREADS) {
        const int y = idx & 15, x = (idx >> 4) & 15, cr = idx >> 8;
        const int c = cbase + cr;
        const int ox = tx0 + x, oy = ty0 + y;
        float v = 0.0f;
        if (c < dm.C && ox < grid.X && oy < grid.Y)
          v = gbev[(size_t)b * st.b + (size_t)c * st.c + (size_t)ox * st.x + oy];
        const int cl = x * LS_TY + y;
        tile[cl * tg.stride + 4 * ls_tile_quad(cl, cr >> 2, tg.nqp) + (cr & 3)] = v;
      }
    }
    __syncthreads();
    // rows of non-empty cells, 16 B per lane: thread = (quad, cell mod 16)
    if ((tid & 15) < nquads) {
      const int q = tid & 15;
      for (int cl = tid >> 4; cl < LS_TILE; cl += LS_THREADS / 16) {
        if (seg[cl + 1] != seg[cl])
          *reinterpret_cast<float4*>(dst + (size_t)cl * dm.Cp + cbase + 4 * q) =
              *reinterpret_cast<const float4*>(tile + cl * tg.stride + 4 * ls_tile_quad(cl, q, tg.nqp));
      }
    }
    __syncthreads();
  }
}

int ls_launch_bwd_transpose(const float* gbev, const LsBevStrides& st, const int* seg_start, const LsDims& dm,
                            const LsGrid& g, float* gT, cudaStream_t s) {
  static bool attr_done = false;
  if (!attr_done) {
    const int m = (int)ls_tile_smem_max();
    LS_CUDA(cudaFuncSetAttribute(ls_bwd_transpose_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, m));
    LS_CUDA(cudaFuncSetAttribute(ls_bwd_transpose_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, m));
    attr_done = true;
  }
  const size_t smem = ls_tile_smem_bytes(dm);
  dim3 grid(g.tiles, dm.B);
  if (ls_bev_vec4(gbev, st, g))
    ls_bwd_transpose_kernel<true><<<grid, LS_THREADS, smem, s>>>(gbev, st, seg_start, dm, g, gT);
  else
    ls_bwd_transpose_kernel<false><<<grid, LS_THREADS, smem, s>>>(gbev, st, seg_start, dm, g, gT);
  LS_LAUNCHED();
  return LS_OK;
}

// =====================================================================================
// K4b: gradient gather, pixel-stationary (deterministic, no atomics).
// reference: VoxelsSumming.backward tool/geometry.py:307-317 + autograd of
// model/bev_model.py:66,91-97.  CTA = one feature-map column of one camera (its rays sweep one
// radial line of the BEV, so the cell-major gradient rows it gathers are re-used from L1);
// a half-warp owns a pixel: 16 lanes x float4 channels, depth bins in windows of 16:
//   gf[c]  += prob[d] * g[cell(d), c]                     (registers, d ascending)
//   gp[d]   = sum_c feat[c] * g[cell(d), c]               (16 dots reduced together by a
//                                                          transposing butterfly: 15 shuffles)
// Outputs: grad_feat NHWC-padded [pix][Cp] and grad_prob PIXEL-major [pix][D].
// =====================================================================================
template <typename T, int NCH>
__global__ void __launch_bounds__(LS_THREADS, 2)
ls_bwd_gather_kernel(const float* __restrict__ gT, const T* __restrict__ featT, const int2* __restrict__ pix_recs,
                     LsDims dm, LsGrid grid, float* __restrict__ gprob_pm, T* __restrict__ gfeatT) {
  const int col = blockIdx.x, bn = blockIdx.y;
  const int b = bn / dm.N;
  const int hw = threadIdx.x >> 4, hl = threadIdx.x & 15;
  const unsigned hmask = ls_half_mask();
  // dropped points carry row index Vc: the all-zero row written by the transpose kernel
  const float* gTb = gT + (size_t)b * (grid.Vc + 1) * dm.Cp + 4 * hl;
  bool on[NCH];
#pragma unroll
  for (int q = 0; q < NCH; ++q) on[q] = (q * LS_CCHUNK + 4 * hl) < dm.Cp;

  for (int row = hw; row < dm.fh; row += LS_HALFWARPS) {
    const size_t pix = (size_t)bn * dm.HW + (size_t)row * dm.fw + col;
    const T* frow = featT + pix * dm.Cp + 4 * hl;
    float4 f[NCH], gf[NCH];
#pragma unroll
    for (int q = 0; q < NCH; ++q) {
      f[q] = on[q] ? ls_load4<T>(frow + q * LS_CCHUNK) : make_float4(0.f, 0.f, 0.f, 0.f);
      gf[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int2* pr = pix_recs + pix * dm.D;
    for (int d0 = 0; d0 < dm.D; d0 += 16) {
      const int n = min(16, dm.D - d0);
      int2 r[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) r[u] = (u < n) ? __ldg(pr + d0 + u) : make_int2(grid.Vc, 0);  // uniform
      float dot[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) dot[u] = 0.0f;
#pragma unroll
      for (int q = 0; q < NCH; ++q) {
        float4 g[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (on[q]) g[u] = __ldg(reinterpret_cast<const float4*>(gTb + (size_t)r[u].x * dm.Cp + q * LS_CCHUNK));
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const float w = __int_as_float(r[u].y);
          dot[u] = fmaf(f[q].x, g[u].x, dot[u]);
          dot[u] = fmaf(f[q].y, g[u].y, dot[u]);
          dot[u] = fmaf(f[q].z, g[u].z, dot[u]);
          dot[u] = fmaf(f[q].w, g[u].w, dot[u]);
          gf[q].x = fmaf(w, g[u].x, gf[q].x);
          gf[q].y = fmaf(w, g[u].y, gf[q].y);
          gf[q].z = fmaf(w, g[u].z, gf[q].z);
          gf[q].w = fmaf(w, g[u].w, gf[q].w);
        }
      }
      // transposing butterfly: lane hl ends up with sum over the 16 lanes of dot[hl]
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const bool hi = hl & 8;
        const float send = hi ? dot[u] : dot[u + 8];
        const float keep = hi ? dot[u + 8] : dot[u];
        dot[u] = keep + __shfl_xor_sync(hmask, send, 8, 16);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool hi = hl & 4;
        const float send = hi ? dot[u] : dot[u + 4];
        const float keep = hi ? dot[u + 4] : dot[u];
        dot[u] = keep + __shfl_xor_sync(hmask, send, 4, 16);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const bool hi = hl & 2;
        const float send = hi ? dot[u] : dot[u + 2];
        const float keep = hi ? dot[u + 2] : dot[u];
        dot[u] = keep + __shfl_xor_sync(hmask, send, 2, 16);
      }
      {
        const bool hi = hl & 1;
        const float send = hi ? dot[0] : dot[1];
        const float keep = hi ? dot[1] : dot[0];
        dot[0] = keep + __shfl_xor_sync(hmask, send, 1, 16);
      }
      if (hl < n) gprob_pm[pix * dm.D + d0 + hl] = dot[0];
    }
    T* grow = gfeatT + pix * dm.Cp + 4 * hl;
#pragma unroll
    for (int q = 0; q < NCH; ++q)
      if (on[q]) ls_store4<T>(grow + q * LS_CCHUNK, gf[q]);
  }
}

template <typename T>
static int ls_gather_dispatch(const float* gT, const void* featT, const int2* pix_recs, const LsDims& dm,
                              const LsGrid& g, float* gprob_pm, void* gfeatT, cudaStream_t s) {
  const int nch = (dm.Cp + LS_CCHUNK - 1) / LS_CCHUNK;
  dim3 grid(dm.fw, dm.B * dm.N);
#define LS_GATHER(NCH) \
  ls_bwd_gather_kernel<T, NCH><<<grid, LS_THREADS, 0, s>>>(gT, (const T*)featT, pix_recs, dm, g, gprob_pm, (T*)gfeatT)
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

int ls_launch_bwd_gather(const float* gT, const void* featT, int dtype, const int2* pix_recs, const LsDims& dm,
                         const LsGrid& g, float* gprob_pm, void* gfeatT, cudaStream_t s) {
  if (dtype == LS_F32) return ls_gather_dispatch<float>(gT, featT, pix_recs, dm, g, gprob_pm, gfeatT, s);
  return ls_gather_dispatch<__nv_bfloat16>(gT, featT, pix_recs, dm, g, gprob_pm, gfeatT, s);
}
