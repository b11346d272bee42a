// The splat itself: deterministic ranked segment-reduce (forward) and its gradient
// (cell-major gradient staging + pixel-stationary gather).  Reference semantics:
// model/bev_model.py:66-72,99-105 and VoxelsSumming (tool/geometry.py:285-317).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda.h>      // CUtensorMap (types only: the encoder is looked up through the runtime at first use)

#include "ls_internal.h"


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

// canonical record: x = pixel << 12 | last_of_cell << 11 | valid << 10 | cell_in_tile,  y = prob bits
#define LS_REC_LAST 0x800
#define LS_REC_VALID 0x400
#ifndef LS_QWIN
#define LS_QWIN 4    // records a quarter-warp keeps in flight
#endif
#ifndef LS_SPLAT_MINB
#define LS_SPLAT_MINB 5   // register budget of the splat: 5 CTAs/SM worth (measured best: 4 and 6 are ~6 us slower)
#endif
#define LS_QWARPS (LS_THREADS / 8)

#ifdef LS_PROFILE
__device__ unsigned long long ls_dbg_phase[8];
__device__ int ls_dbg_cta[8192][8];   // per CTA: records, then cycles of each phase
#define LS_TICK(k)                                                                  \
  do {                                                                              \
    if (threadIdx.x == 0) {                                                         \
      const long long now__ = clock64();                                            \
      atomicAdd(&ls_dbg_phase[k], (unsigned long long)(now__ - tick__));            \
      if (blockIdx.x < 8192) ls_dbg_cta[blockIdx.x][k + 1] = (int)(now__ - tick__); \
      tick__ = now__;                                                               \
    }                                                                               \
  } while (0)
#define LS_TICK_INIT() long long tick__ = clock64()
#else
#define LS_TICK(k) do {} while (0)
#define LS_TICK_INIT() do {} while (0)
#endif

__device__ __forceinline__ unsigned ls_quarter_mask() { return 0xFFu << (threadIdx.x & 24); }

// Canonical records of a sample live in the CSR layout of `recs` (slot seg_start[cell] + rank);
// the slack lets a stream prefetch one window past the end of the sample.
#define LS_SORTED_SLACK 8
__host__ __device__ __forceinline__ size_t ls_sorted_capacity(int Npts) { return (size_t)Npts + LS_SORTED_SLACK; }

// =====================================================================================
// K3a: canonical order.  The placement wrote every cell's records in ticket (atomic arrival)
// order; here each record finds its rank among the records of its cell (keys are unique:
// pixel, depth bin) and moves to slot seg_start[cell] + rank of a second buffer, so the
// splat's summation order - and with it every bit of the BEV tensor - is independent of
// the atomics.  Flat: one thread per record, one CTA per tile (heaviest first), the compare
// loop runs over L1-resident keys.  Replaces the (unstable) argsort of model/bev_model.py:96.
// =====================================================================================
#ifndef LS_CANON_THREADS
#define LS_CANON_THREADS 256
#endif
#ifndef LS_CANON_BIG
#define LS_CANON_BIG 512      // cells with more records than this take the bucketed path (its ~50 barriers cost more than 512^2/256 compares per thread)
#endif
#define LS_CANON_BUCKETS 1024 // by the top 10 of the 24 key bits (4 KB of shared memory: keeps the L1 share of the hot path)
#define LS_CANON_BSHIFT 14

// one tile of sample b by a whole CTA of LS_CANON_THREADS threads (CTA-uniform control flow)
__device__ __forceinline__ void ls_canon_tile(int2* __restrict__ recs, const int* __restrict__ seg_start, int b, int tile_id,
                                              const LsDims& dm, const LsGrid& grid, int2* __restrict__ recs_sorted,
                                              int* __restrict__ perm) {
  __shared__ int seg[LS_TILE + 1];
  __shared__ int any_big;
  const int* segg = seg_start + (size_t)b * grid.seg_stride + (size_t)tile_id * LS_TILE;
  __syncthreads();                       // (called in a loop: the previous tile's readers of seg / any_big are done)
  if (threadIdx.x == 0) any_big = 0;
  for (int i = threadIdx.x; i <= LS_TILE; i += LS_CANON_THREADS) seg[i] = segg[i];
  __syncthreads();
  const int s0 = seg[0], s1 = seg[LS_TILE];
  int2* rin = recs + (size_t)b * dm.Npts;
  int2* out = recs_sorted + (size_t)b * ls_sorted_capacity(dm.Npts);
  // perm (static-rig cache only): canonical slot -> point id (n*D + d)*HW + rc, the index of the point's
  // probability inside the sample; lets a later step refresh the weights without redoing the sort
  int* pm = perm ? perm + (size_t)b * dm.Npts : nullptr;
  const int dmask = (1 << dm.dbits) - 1;
  auto point_of = [&](int key, int pix) { const int n = pix / dm.HW; return (n * dm.D + (key & dmask)) * dm.HW + (pix - n * dm.HW); };
  // (read-only path: the light cells' records are never written by this kernel)
  const int2* __restrict__ rro = rin;
  // (a thread's next record is fetched before the compare loop of the current one: the kernel is bound by
  // the length of this per-thread chain of dependent loads, not by work)
  int2 rnext = make_int2(0, 0);
  if (s0 + (int)threadIdx.x < s1) rnext = __ldg(rro + s0 + threadIdx.x);
  for (int i = s0 + threadIdx.x; i < s1; i += LS_CANON_THREADS) {
    const int2 r = rnext;
    if (i + LS_CANON_THREADS < s1) rnext = __ldg(rro + i + LS_CANON_THREADS);
    const int cl = (unsigned)r.x >> 24;
    const int a = seg[cl], e = seg[cl + 1];
    if (e - a > LS_CANON_BIG) { any_big = 1; continue; }      // handled below by the whole CTA
    int pos = a;
#pragma unroll 4
    for (int j = a; j < e; ++j) pos += (__ldg(&rro[j].x) < r.x) ? 1 : 0;
    const int pix = (r.x & 0xFFFFFF) >> dm.dbits;
    out[pos] = make_int2((pix << 12) | cl | LS_REC_VALID | (pos == e - 1 ? LS_REC_LAST : 0), r.y);
    if (pm) pm[pos] = point_of(r.x, pix);
  }
  __syncthreads();
  if (!any_big) return;
  // ---- heavy cells (coarse grids, degenerate rigs: up to every point of the sample in one cell) ----
  // Rank-by-counting is quadratic in the cell's record count; here the cell is first grouped by the top
  // 10 key bits (shared-memory histogram + scan + scatter, integer atomics only), then every record is
  // ranked against its own bucket only: k * (k / 1024) compares instead of k^2.  The input run of the
  // cell (dead after the scatter) is the scratch for the final order.
  __shared__ int bucket_end[LS_CANON_BUCKETS];      // histogram -> scan -> scatter cursor -> end of each bucket
  __shared__ int scan_tmp[LS_CANON_THREADS];
  for (int cl = 0; cl < LS_TILE; ++cl) {
    const int a = seg[cl], e = seg[cl + 1];
    if (e - a <= LS_CANON_BIG) continue;             // uniform over the CTA
    for (int i = threadIdx.x; i < LS_CANON_BUCKETS; i += LS_CANON_THREADS) bucket_end[i] = 0;
    __syncthreads();
    for (int i = a + threadIdx.x; i < e; i += LS_CANON_THREADS)
      atomicAdd(&bucket_end[(__ldcg(&rin[i].x) & 0xFFFFFF) >> LS_CANON_BSHIFT], 1);
    __syncthreads();
    {
      // exclusive scan of 4096 counts: 16 per thread, then a block scan of the thread totals
      constexpr int kPer = LS_CANON_BUCKETS / LS_CANON_THREADS;
      int loc[kPer], tot = 0;
#pragma unroll
      for (int j = 0; j < kPer; ++j) { loc[j] = bucket_end[threadIdx.x * kPer + j]; tot += loc[j]; }
      scan_tmp[threadIdx.x] = tot;
      __syncthreads();
      for (int o = 1; o < LS_CANON_THREADS; o <<= 1) {
        const int v = (int)threadIdx.x >= o ? scan_tmp[threadIdx.x - o] : 0;
        __syncthreads();
        scan_tmp[threadIdx.x] += v;
        __syncthreads();
      }
      int run = scan_tmp[threadIdx.x] - tot;
#pragma unroll
      for (int j = 0; j < kPer; ++j) { bucket_end[threadIdx.x * kPer + j] = run; run += loc[j]; }
    }
    __syncthreads();
    // scatter into bucket order (arbitrary inside a bucket); the cursor ends at the bucket's end
    for (int i = a + threadIdx.x; i < e; i += LS_CANON_THREADS) {
      const int2 r = __ldcg(&rin[i]);
      const int pos = atomicAdd(&bucket_end[(r.x & 0xFFFFFF) >> LS_CANON_BSHIFT], 1);
      out[a + pos] = r;
    }
    __syncthreads();
    // rank inside the bucket -> final slot, written in the output format into the (dead) input run
    for (int i = a + threadIdx.x; i < e; i += LS_CANON_THREADS) {
      const int2 r = __ldcg(&out[i]);
      const int bk = (r.x & 0xFFFFFF) >> LS_CANON_BSHIFT;
      const int lo = a + (bk ? bucket_end[bk - 1] : 0), hi = a + bucket_end[bk];
      int pos = lo;
      for (int j = lo; j < hi; ++j) pos += (__ldcg(&out[j].x) < r.x) ? 1 : 0;
      const int pix = (r.x & 0xFFFFFF) >> dm.dbits;
      rin[pos] = make_int2((pix << 12) | cl | LS_REC_VALID | (pos == e - 1 ? LS_REC_LAST : 0), r.y);
      if (pm) pm[pos] = point_of(r.x, pix);
    }
    __syncthreads();
    for (int i = a + threadIdx.x; i < e; i += LS_CANON_THREADS) out[i] = __ldcg(&rin[i]);
    __syncthreads();
  }
}

__global__ void __launch_bounds__(LS_CANON_THREADS)
ls_canon_kernel(int2* __restrict__ recs, const int* __restrict__ seg_start, const int* __restrict__ tile_order,
                LsDims dm, LsGrid grid, int2* __restrict__ recs_sorted, int* __restrict__ perm) {
  ls_pdl_trigger();
  ls_pdl_wait();
  const int b = blockIdx.x % dm.B;
  ls_canon_tile(recs, seg_start, b, tile_order[(size_t)b * grid.tiles + blockIdx.x / dm.B], dm, grid, recs_sorted, perm);
}

// =====================================================================================
// K3b: forward splat.  One CTA per (sample, LS_TX x LS_TY voxel tile), heaviest tiles first:
//  B  reduce: the tile's run of canonical records is cut into LS_QWARPS pieces of equal record
//     count, one per quarter-warp.  A lane owns 8 channels (two 16-byte pieces of the 256-byte
//     feature row); LS_QWIN rows are in flight while the next records are prefetched; prob*feat
//     accumulates in registers and is dropped into the shared-memory tile [cell][channel]
//     (swizzled, conflict-free) on a last-of-cell record.  A cell cut by a piece boundary is
//     stored by the piece holding its last record; the open sums of the pieces before it are
//     added after the barrier in piece order (fixed association).
//  C  write-out: the tile is read column-wise and written as 16-byte pieces of the
//     [B,C,X,Y] tensor, zeros included - the BEV grid is never memset.
//     With a channels-last BEV tensor (OUT = LS_OUT_NHWC_BULK) the tile [cell][channel] IS the
//     memory image of 128 consecutive cells: rows are kept dense (no padding, no swizzle - the
//     column walks that needed them are gone) and the whole tile leaves as one bulk async
//     (TMA) store issued by a single thread.
// kCC = 64: the common case (Cp == 64) with compile-time tile geometry; kCC = 0: any Cp.
// =====================================================================================
template <typename T, int OUT, int kCC>
__global__ void __launch_bounds__(LS_THREADS, LS_SPLAT_MINB)
ls_splat_fwd_kernel(const T* __restrict__ featT, const int* __restrict__ seg_start,
                    const int* __restrict__ tile_order, const int2* __restrict__ recs_sorted, LsDims dm, LsGrid grid,
                    float* __restrict__ bev, LsBevStrides st) {
  extern __shared__ __align__(128) float smem[];
  static_assert(OUT != LS_OUT_NHWC_BULK || (kCC == 64 && LS_TX == 1), "bulk store: dense 64-channel rows of one x-row");
  constexpr bool kDense = OUT == LS_OUT_NHWC_BULK;
  const LsTileGeom tgr = ls_tile_geom(dm.Cp);
  const int Cp = kCC ? kCC : dm.Cp;
  const int stride = kDense ? kCC : (kCC ? kCC + 4 : tgr.stride);
  const int nqp = kDense ? 1 : (kCC ? kCC / 4 : tgr.nqp);     // nqp == 1: no swizzle
  const int ccmax = kCC ? kCC : tgr.cc;
  float* tile = smem;                                            // [LS_TILE][stride]
  int* part_cell = reinterpret_cast<int*>(smem + LS_TILE * stride);   // [LS_QWARPS] cell of each piece's open partial sum (-1: none)

  // heaviest tiles first, all samples interleaved: blockIdx.x = order_index * B + b
  const int b = blockIdx.x % dm.B;
  const int tid = threadIdx.x;
  LS_TICK_INIT();
  ls_pdl_trigger();
  // private prologue of a programmatic dependent launch: zero the accumulator tile, then wait
  // for the producers of seg_start / tile_order / the canonical records
  for (int i = tid; i < LS_TILE * stride / 4; i += LS_THREADS)
    reinterpret_cast<float4*>(tile)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  ls_pdl_wait();
  const int tile_id = tile_order[(size_t)b * grid.tiles + blockIdx.x / dm.B];
  const int tx0 = (tile_id / grid.tiles_y) * LS_TX, ty0 = (tile_id % grid.tiles_y) * LS_TY;
  const int* segg = seg_start + (size_t)b * grid.seg_stride + (size_t)tile_id * LS_TILE;
  // only the ends of the tile's record run are needed (uniform loads, one line each)
  const int s0 = __ldg(segg), s1 = __ldg(segg + LS_TILE);

  const int ql = tid & 7;
  const T* fbase = featT + (size_t)b * dm.N * dm.HW * Cp;
  const int2* rs = recs_sorted + (size_t)b * ls_sorted_capacity(dm.Npts);
  unsigned row_bytes = (unsigned)(Cp * sizeof(T));
  asm volatile("" : "+r"(row_bytes));   // opaque: pixel * row_bytes + base stays one wide multiply-add
  // phase C geometry of this thread (fixed): 4 consecutive y of one x-row; channel quads qg, qg+4, ...
  const int y4 = tid % (LS_TY / 4), xr = (tid / (LS_TY / 4)) % LS_TX, qg = tid / (LS_TILE / 4);
  const int gx = tx0 + xr, gy = ty0 + 4 * y4;
  const bool inb = gx < grid.X && gy < grid.Y;
  const int clc = xr * LS_TY + 4 * y4;

  for (int cbase = 0; cbase < Cp; cbase += LS_CCHUNK) {
    const int cc = kCC ? kCC : min(ccmax, Cp - cbase);
    const int nquads = cc >> 2;
    // zero the tile (cells nobody hits are never touched by phase B); the first pass did it
    // in the prologue
    if (cbase > 0) {
      for (int i = tid; i < LS_TILE * stride / 4; i += LS_THREADS)
        reinterpret_cast<float4*>(tile)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    LS_TICK(0);
    const bool tile_empty = (s0 == s1);
#ifdef LS_PROFILE
    if (threadIdx.x == 0 && blockIdx.x < 8192) ls_dbg_cta[blockIdx.x][0] = s1 - s0;
#endif
    // ---- phase B: lane = channels [4ql,4ql+4) and [32+4ql,..) of its quarter-warp's records ----
    if (!tile_empty) {
      // Equal pieces: quarter-warp q reduces records [s0 + q*n/16, s0 + (q+1)*n/16) of the tile's
      // canonical run, cut wherever that falls.  A cell that straddles a cut is finished by the
      // piece that holds its last record (plain store); the pieces before it keep their partial
      // sums in registers and add them after the barrier, in piece order - a fixed association,
      // so the result is still independent of the atomics and of the schedule.
      const int qw = tid >> 3, n = s1 - s0;
      int idx = s0 + (int)(((long long)qw * n) / LS_QWARPS);
      const int end = s0 + (int)(((long long)(qw + 1) * n) / LS_QWARPS);
      const int idx_first = idx;
      const bool on0 = 4 * ql < cc, on1 = 32 + 4 * ql < cc;
      // lanes beyond the channel count read valid bytes (lane 0's) and never store
      const char* f0 = reinterpret_cast<const char*>(fbase + cbase + (on0 ? 4 * ql : 0));
      const unsigned f1off = (unsigned)((on1 ? 32 : 0) * sizeof(T));       // second 16-byte piece of the row
      unsigned tile_s = (unsigned)__cvta_generic_to_shared(tile);          // 32-bit shared-window address
      asm volatile("" : "+r"(tile_s));                                     // computed once, not per flush
      const unsigned row_sbytes = (unsigned)stride * 4u, ql16 = (unsigned)ql << 4, swz_mask = (unsigned)(nqp - 1) << 4;
      float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = make_float4(0.f, 0.f, 0.f, 0.f);
      int pc = -1;                                           // cell of this piece's open partial sum
      if (idx < end) {
        const int xl = __ldg(&rs[end - 1].x);
        if (!(xl & LS_REC_LAST)) pc = xl & 255;
        const int2* p = rs + idx;
        int2 r[LS_QWIN], rn[LS_QWIN];
#pragma unroll
        for (int u = 0; u < LS_QWIN; ++u) {                     // quarter-warp-uniform 8-byte loads
          r[u] = p[u];
          if (idx + u >= end) r[u] = make_int2(0, 0);           // beyond the piece: pixel 0, not valid
        }
        // One window: gather the rows of `cur`, fetch the next window's records into `nxt`
        // while they are in flight, then accumulate.  Two windows per trip with the record
        // buffers swapped, so no registers are copied between trips.
#define LS_SPLAT_WINDOW(cur, nxt)                                                                    \
  {                                                                                                  \
    float4 fa[LS_QWIN], fb[LS_QWIN];                                                                 \
    _Pragma("unroll") for (int u = 0; u < LS_QWIN; ++u) {                                            \
      /* base + pixel * row_bytes as one 32x32+64 multiply-add */                                    \
      const char* row = f0 + (unsigned long long)((unsigned)cur[u].x >> 12) * row_bytes;             \
      fa[u] = ls_load4<T>(reinterpret_cast<const T*>(row));                                          \
      fb[u] = ls_load4<T>(reinterpret_cast<const T*>(row + f1off));                                  \
    }                                                                                                \
    idx += LS_QWIN;                                                                                  \
    p += LS_QWIN;                                                                                    \
    _Pragma("unroll") for (int u = 0; u < LS_QWIN; ++u) nxt[u] = (idx + u < end) ? p[u] : make_int2(0, 0); \
    _Pragma("unroll") for (int u = 0; u < LS_QWIN; ++u) {                                            \
      if (cur[u].x & LS_REC_VALID) {                                                                 \
        const float wt = __int_as_float(cur[u].y);                                                   \
        acc0.x = fmaf(wt, fa[u].x, acc0.x); acc0.y = fmaf(wt, fa[u].y, acc0.y);                      \
        acc0.z = fmaf(wt, fa[u].z, acc0.z); acc0.w = fmaf(wt, fa[u].w, acc0.w);                      \
        acc1.x = fmaf(wt, fb[u].x, acc1.x); acc1.y = fmaf(wt, fb[u].y, acc1.y);                      \
        acc1.z = fmaf(wt, fb[u].z, acc1.z); acc1.w = fmaf(wt, fb[u].w, acc1.w);                      \
      }                                                                                              \
      if (cur[u].x & LS_REC_LAST) {                                                                  \
        const unsigned cl = (unsigned)cur[u].x & 255u;                                               \
        /* row of the cell + this lane's swizzled quads (q and q+8 differ by one address bit) */     \
        const unsigned rowb = tile_s + cl * row_sbytes, x0 = ql16 ^ ((cl << 1) & swz_mask);          \
        const unsigned a0 = rowb + x0, a1 = rowb + (x0 ^ 0x80u);                                     \
        if (on0)                                                                                     \
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a0), "f"(acc0.x), "f"(acc0.y), \
                       "f"(acc0.z), "f"(acc0.w) : "memory");                                         \
        if (on1)                                                                                     \
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a1), "f"(acc1.x), "f"(acc1.y), \
                       "f"(acc1.z), "f"(acc1.w) : "memory");                                         \
        acc0 = make_float4(0.f, 0.f, 0.f, 0.f);                                                      \
        acc1 = make_float4(0.f, 0.f, 0.f, 0.f);                                                      \
      }                                                                                              \
    }                                                                                                \
  }
        for (;;) {
          LS_SPLAT_WINDOW(r, rn);
          if (idx >= end) break;
          LS_SPLAT_WINDOW(rn, r);
          if (idx >= end) break;
        }
#undef LS_SPLAT_WINDOW
      }
      (void)idx_first;
      if (ql == 0) part_cell[qw] = pc;
      __syncthreads();
      // open partial sums: round r adds the partial of every piece that has exactly r open
      // pieces of the same cell directly before it (a cell longer than two pieces), so no two
      // quarter-warps touch the same row in one round
      int depth = 0, maxd = -1;
      {
        int run = 0, prev = -1;
#pragma unroll
        for (int j = 0; j < LS_QWARPS; ++j) {
          const int c = part_cell[j];
          if (c >= 0) {                 // pieces without an open sum (empty, or ending on a cell's last record) are transparent
            run = (c == prev) ? run + 1 : 0;
            maxd = max(maxd, run);
            prev = c;
          }
          if (j == qw) depth = run;
        }
      }
      for (int rd = 0; rd <= maxd; ++rd) {
        if (pc >= 0 && depth == rd) {
          const unsigned cl = (unsigned)pc;
          const unsigned rowb = tile_s + cl * row_sbytes, x0 = ql16 ^ ((cl << 1) & swz_mask);
          const unsigned a0 = rowb + x0, a1 = rowb + (x0 ^ 0x80u);
          float4 t0, t1;
          if (on0) {
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t0.x), "=f"(t0.y), "=f"(t0.z), "=f"(t0.w) : "r"(a0) : "memory");
            t0.x += acc0.x; t0.y += acc0.y; t0.z += acc0.z; t0.w += acc0.w;
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a0), "f"(t0.x), "f"(t0.y), "f"(t0.z), "f"(t0.w) : "memory");
          }
          if (on1) {
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t1.x), "=f"(t1.y), "=f"(t1.z), "=f"(t1.w) : "r"(a1) : "memory");
            t1.x += acc1.x; t1.y += acc1.y; t1.z += acc1.z; t1.w += acc1.w;
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a1), "f"(t1.x), "f"(t1.y), "f"(t1.z), "f"(t1.w) : "memory");
          }
        }
        __syncthreads();
      }
    } else {
      __syncthreads();
    }
    LS_TICK(3);
    // ---- phase C --------------------------------------------------------------------
    if (OUT == LS_OUT_NHWC_BULK) {
      // generic-proxy writes of the tile -> visible to the async proxy, then one thread hands the
      // valid part of the tile (whole 256-byte rows, zeros included) to the TMA unit
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        const int valid = min(LS_TY, grid.Y - ty0);
        float* dstp = bev + (size_t)b * st.b + (size_t)tx0 * st.x + (size_t)ty0 * kCC;
        const unsigned src = (unsigned)__cvta_generic_to_shared(tile);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dstp), "r"(src),
                     "r"((unsigned)(valid * kCC * 4)) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the tile must outlive the read
      }
    } else if (OUT == LS_OUT_NHWC_ROWS) {
      // channels-last with any row pitch / channel count: a warp writes one cell's channels as
      // consecutive 4-byte stores (the chunk of a row is contiguous in memory)
      const int lane = tid & 31;
      for (int cl = tid >> 5; cl < LS_TILE; cl += LS_THREADS / 32) {
        const int ox = tx0 + cl / LS_TY, oy = ty0 + cl % LS_TY;
        if (ox >= grid.X || oy >= grid.Y) continue;
        float* g = bev + (size_t)b * st.b + (size_t)ox * st.x + (size_t)oy * st.y + cbase;
        for (int cr = lane; cr < cc; cr += 32) {
          if (cbase + cr < dm.C)
            g[cr] = tile_empty ? 0.0f : tile[cl * stride + 4 * ls_tile_quad(cl, cr >> 2, nqp) + (cr & 3)];
        }
      }
    } else if (OUT == LS_OUT_NCHW_VEC4) {
      if (inb) {
        // a thread reads one channel quad of 4 consecutive cells (4 x 16 B, conflict-free),
        // transposes the 4x4 block in registers and writes 4 channels x 4 y as 16-byte
        // streaming stores (the 164 MB output stream must not evict the feature rows from L2)
        const int swz = (clc >> 3) & (nqp - 1);
        const float* srow = tile + clc * stride;
        float* gbase = bev + (size_t)b * st.b + (size_t)cbase * st.c + (size_t)gx * st.x + gy;
#pragma unroll 4
        for (int q = qg; q < nquads; q += 4) {
          float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0, r2 = r0, r3 = r0;
          if (!tile_empty) {
            const float* src = srow + 4 * (q ^ swz);
            r0 = *reinterpret_cast<const float4*>(src);
            r1 = *reinterpret_cast<const float4*>(src + stride);
            r2 = *reinterpret_cast<const float4*>(src + 2 * stride);
            r3 = *reinterpret_cast<const float4*>(src + 3 * stride);
          }
          float* g = gbase + (size_t)(4 * q) * st.c;
          const int c = cbase + 4 * q;
          if (kCC || c + 0 < dm.C) __stcs(reinterpret_cast<float4*>(g), make_float4(r0.x, r1.x, r2.x, r3.x));
          if (kCC || c + 1 < dm.C) __stcs(reinterpret_cast<float4*>(g + st.c), make_float4(r0.y, r1.y, r2.y, r3.y));
          if (kCC || c + 2 < dm.C) __stcs(reinterpret_cast<float4*>(g + 2 * st.c), make_float4(r0.z, r1.z, r2.z, r3.z));
          if (kCC || c + 3 < dm.C) __stcs(reinterpret_cast<float4*>(g + 3 * st.c), make_float4(r0.w, r1.w, r2.w, r3.w));
        }
      }
    } else {
      for (int idx = tid; idx < 4 * nquads * LS_TILE; idx += LS_THREADS) {
        const int y = idx % LS_TY, x = (idx / LS_TY) % LS_TX, cr = idx / LS_TILE;
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
    LS_TICK(4);
  }
}

// =====================================================================================
// K3c: forward splat for a dense channels-last BEV tensor, WITHOUT the shared-memory tile.
// With rows as the unit of both accumulation and output there is nothing to transpose: a
// quarter-warp that finishes a cell writes its 256-byte row straight to global memory, cells
// nobody hits get their zero rows from the warp that owns them.  What this buys: the 32 KB per CTA that
// the tile took from the unified L1/shared array stay L1 - the feature rows a tile re-reads
// (the same ray crosses it in consecutive depth bins) are served on-chip instead of going back
// to L2, which the tile version saturates (9.6 of ~12.4 TB/s of L2 throughput).
// =====================================================================================
#ifndef LS_ABLATE
#define LS_ABLATE 0       // developer builds only (tools/ab_build.sh): 1 = memory traffic without the arithmetic, 2 = arithmetic
                          // without the row gathers, in the direct splat and the register-lean gather.  Results are WRONG.
#endif
#ifndef LS_SPLATD_MINB
#define LS_SPLATD_MINB 6
#endif
#ifndef LS_SPLAT_FASTLOOP
#define LS_SPLAT_FASTLOOP 0   // unmasked record prefetch for full windows: 72 -> 80 registers, splat +3 us (measured); off
#endif
// 16 bytes of a row: one vector access when rows are 16-byte aligned (row pitch a multiple of 4
// floats), four scalar ones otherwise (a 64-channel slice of a 65-channel tensor: 260-byte pitch)
template <bool kVec> __device__ __forceinline__ void ls_row_store4(float* g, float4 v) {
  if (kVec) { __stcs(reinterpret_cast<float4*>(g), v); }
  else { __stcs(g, v.x); __stcs(g + 1, v.y); __stcs(g + 2, v.z); __stcs(g + 3, v.w); }
}
// bf16 BEV rows (opt-in, always 8-byte aligned): four channels = one 8-byte store
template <bool kVec> __device__ __forceinline__ void ls_row_store4(__nv_bfloat16* g, float4 v) {
  __nv_bfloat162 a = __float22bfloat162_rn(make_float2(v.x, v.y)), b = __float22bfloat162_rn(make_float2(v.z, v.w));
  uint2 raw;
  raw.x = *reinterpret_cast<unsigned*>(&a);
  raw.y = *reinterpret_cast<unsigned*>(&b);
  __stcs(reinterpret_cast<uint2*>(g), raw);
}

// eight consecutive bf16 channels (16 B) -> two float4
__device__ __forceinline__ void ls_load8_bf16(const char* p, float4& a, float4& b) {
  const uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
  const float2 v0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
  const float2 v1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
  const float2 v2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.z));
  const float2 v3 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.w));
  a = make_float4(v0.x, v0.y, v1.x, v1.y);
  b = make_float4(v2.x, v2.y, v3.x, v3.y);
}

// kStrip: 1 x 128 tiles (the default): a cell's row offset is cl * y-stride, no shift/mask/second multiply
template <typename T, bool kVec, typename TO, bool kStrip>
__global__ void __launch_bounds__(LS_THREADS, LS_SPLATD_MINB)
ls_splat_fwd_direct_kernel(const T* __restrict__ featT, const int* __restrict__ seg_start,
                           const int* __restrict__ tile_order, const int2* __restrict__ recs_sorted, LsDims dm,
                           LsGrid grid, TO* __restrict__ bev, LsBevStrides st) {
  constexpr int kC = 64;
  __shared__ int seg[LS_TILE + 1];
  __shared__ int cuts[LS_QWARPS + 1];
  const int b = blockIdx.x % dm.B;
  const int tid = threadIdx.x;
  ls_pdl_trigger();
  ls_pdl_wait();
  const int tile_id = tile_order[(size_t)b * grid.tiles + blockIdx.x / dm.B];
  const int tx0 = (tile_id / grid.tiles_y) << grid.tx_shift, ty0 = (tile_id % grid.tiles_y) << grid.ty_shift;
  const int* segg = seg_start + (size_t)b * grid.seg_stride + (size_t)tile_id * LS_TILE;
  TO* tile0 = bev + (size_t)b * st.b + (size_t)tx0 * st.x + (size_t)ty0 * st.y;
  // row of cell-in-tile cl (elements from tile0): tile row cl / ty, column cl % ty
  const unsigned ty_shift = (unsigned)grid.ty_shift, ty_mask = (unsigned)grid.ty - 1u;
  const unsigned sxu = (unsigned)st.x, syu = (unsigned)st.y;      // < 2^32 floats (checked by the classifier)
  auto row_off = [&](unsigned cl) -> size_t {
    if (kStrip) return (size_t)(cl * syu);
    return (size_t)((cl >> ty_shift) * sxu) + (size_t)((cl & ty_mask) * syu);
  };
  {
    // this thread's cell: empty cells inside the grid get a zero row (the BEV tensor is never memset)
    const int a = __ldg(segg + tid), e = __ldg(segg + tid + 1);
    seg[tid] = a;
    if (tid == LS_TILE - 1) seg[LS_TILE] = e;
    const bool in_grid = kStrip ? (ty0 + tid < grid.Y)
                                : (tx0 + (int)((unsigned)tid >> ty_shift)) < grid.X && (ty0 + (int)((unsigned)tid & ty_mask)) < grid.Y;
    unsigned m = __ballot_sync(0xffffffffu, in_grid && a == e);
    const int lane = tid & 31, wbase = tid & ~31;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    while (m) {                                   // two rows per trip: 16 lanes x 16 B each
      const int c0 = __ffs(m) - 1;
      m &= m - 1;
      int c1 = -1;
      if (m) { c1 = __ffs(m) - 1; m &= m - 1; }
      const int c = lane < 16 ? c0 : c1;
      if (c >= 0) ls_row_store4<kVec>(tile0 + row_off(wbase + c) + 4 * (lane & 15), z);
    }
  }
  __syncthreads();
  const int s0 = seg[0], s1 = seg[LS_TILE];
  if (s0 == s1) return;
  // Pieces: quarter-warp q takes the records from the start of the cell that holds record
  // s0 + q*n/16 to the start of the cell that holds record s0 + (q+1)*n/16 - equal shares rounded
  // down to cell boundaries, so every cell is summed by ONE quarter-warp, front to back, in
  // canonical order: the bits of a cell do not depend on the tiling or on where the cuts fall.
  const int ql = tid & 7, qw = tid >> 3, n = s1 - s0;
  if (tid <= LS_QWARPS) {                        // 17 threads find the 17 cuts, everyone else just reads them
    int c = s1;
    if (tid < LS_QWARPS) {
      const int target = s0 + (int)(((long long)tid * n) / LS_QWARPS);
      int lo = 0, hi = LS_TILE;                  // largest cell with seg[cell] <= target
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int mid = (lo + hi + 1) >> 1;
        if (seg[mid] <= target) lo = mid; else hi = mid - 1;
      }
      c = seg[lo];
    }
    cuts[tid] = c;
  }
  __syncthreads();
  int idx = cuts[qw];
  const int end = cuts[qw + 1];
  if (idx >= end) return;
  const T* fbase = featT + (size_t)b * dm.N * dm.HW * kC;
  const int2* rs = recs_sorted + (size_t)b * ls_sorted_capacity(dm.Npts);
  unsigned row_bytes = (unsigned)(kC * sizeof(T));
  asm volatile("" : "+r"(row_bytes));
  // channels of this lane: fp32 rows are 256 B = two 16-byte pieces per lane, 128 B apart (quads ql and
  // ql + 8); bf16 rows are 128 B = ONE 16-byte piece per lane (channels 8*ql .. 8*ql+7), which halves
  // the load instructions and L1 wavefronts of the gather - the bf16 fast path
  constexpr bool kHalf = sizeof(T) == 2;
  constexpr int kSecond = kHalf ? 4 : 32;        // channel distance between the lane's two float4 accumulators
  const char* f0 = reinterpret_cast<const char*>(fbase + (kHalf ? 8 : 4) * ql);
  const unsigned f1off = (unsigned)(32 * sizeof(T));
  TO* lane0 = tile0 + (kHalf ? 8 : 4) * ql;       // this lane's first quad of row 0
  // strips: a finished cell's row address is ONE 32x32+64 multiply-add (cell * row pitch in bytes + base)
  unsigned pitch_bytes = syu * (unsigned)sizeof(TO);
  asm volatile("" : "+r"(pitch_bytes));
  const char* lane0b = reinterpret_cast<const char*>(lane0);
  float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = make_float4(0.f, 0.f, 0.f, 0.f);
#if LS_FFMA2
  // the 8 channel sums of this lane live in four packed float32 pairs: 4 FFMA2 per record instead of 8 FFMA
  unsigned long long a01 = ls_pack2(0.f, 0.f), a23 = a01, a45 = a01, a67 = a01;
#define LS_ACC8(WT__, A, B)                                                      \
  {                                                                              \
    const unsigned long long w2__ = ls_pack2(WT__, WT__);                        \
    a01 = ls_fma2(w2__, ls_pack2(A.x, A.y), a01);                                \
    a23 = ls_fma2(w2__, ls_pack2(A.z, A.w), a23);                                \
    a45 = ls_fma2(w2__, ls_pack2(B.x, B.y), a45);                                \
    a67 = ls_fma2(w2__, ls_pack2(B.z, B.w), a67);                                \
  }
#define LS_ACC_GET() { ls_unpack2(a01, acc0.x, acc0.y); ls_unpack2(a23, acc0.z, acc0.w); ls_unpack2(a45, acc1.x, acc1.y); ls_unpack2(a67, acc1.z, acc1.w); }
#define LS_ACC_ZERO() { a01 = ls_pack2(0.f, 0.f); a23 = a01; a45 = a01; a67 = a01; }
#elif LS_ABLATE == 1      /* developer ablation: every load and store, two adds instead of eight FMAs per record */
#define LS_X3(a, b, c) __int_as_float(__float_as_int(a) ^ __float_as_int(b) ^ __float_as_int(c))
#define LS_ACC8(WT__, A, B) { acc0.x = LS_X3(acc0.x, A.x, A.y); acc0.y = LS_X3(acc0.y, A.z, A.w); acc1.x = LS_X3(acc1.x, B.x, B.y); acc1.y = LS_X3(acc1.y, B.z, B.w); acc0.z += WT__; }
#define LS_ACC_GET() {}
#define LS_ACC_ZERO() { acc0 = make_float4(0.f, 0.f, 0.f, 0.f); acc1 = make_float4(0.f, 0.f, 0.f, 0.f); }
#else
#define LS_ACC8(WT__, A, B)                                                                          \
  {                                                                                                  \
    acc0.x = fmaf(WT__, A.x, acc0.x); acc0.y = fmaf(WT__, A.y, acc0.y);                              \
    acc0.z = fmaf(WT__, A.z, acc0.z); acc0.w = fmaf(WT__, A.w, acc0.w);                              \
    acc1.x = fmaf(WT__, B.x, acc1.x); acc1.y = fmaf(WT__, B.y, acc1.y);                              \
    acc1.z = fmaf(WT__, B.z, acc1.z); acc1.w = fmaf(WT__, B.w, acc1.w);                              \
  }
#define LS_ACC_GET() {}
#define LS_ACC_ZERO() { acc0 = make_float4(0.f, 0.f, 0.f, 0.f); acc1 = make_float4(0.f, 0.f, 0.f, 0.f); }
#endif
  const int2* p = rs + idx;
  int2 r[LS_QWIN], rn[LS_QWIN];
#pragma unroll
  for (int u = 0; u < LS_QWIN; ++u) {
    r[u] = p[u];
    if (idx + u >= end) r[u] = make_int2(0, 0);
  }
#define LS_SPLATD_WINDOW(cur, nxt, MASKED)                                                           \
  {                                                                                                  \
    float4 fa[LS_QWIN], fb[LS_QWIN];                                                                 \
    _Pragma("unroll") for (int u = 0; u < LS_QWIN; ++u) {                                            \
      const char* row = f0 + (unsigned long long)((unsigned)cur[u].x >> 12) * row_bytes;             \
      if (LS_ABLATE == 2) {          /* developer ablation: no feature-row traffic, same arithmetic */ \
        fa[u] = make_float4(__int_as_float(cur[u].x), 1.f, 2.f, 3.f);                                \
        fb[u] = make_float4(__int_as_float(cur[u].y), 1.f, 2.f, 3.f);                                \
      } else if (kHalf) ls_load8_bf16(row, fa[u], fb[u]);                                            \
      else {                                                                                         \
        fa[u] = ls_load4<T>(reinterpret_cast<const T*>(row));                                        \
        fb[u] = ls_load4<T>(reinterpret_cast<const T*>(row + f1off));                                \
      }                                                                                              \
    }                                                                                                \
    idx += LS_QWIN;                                                                                  \
    p += LS_QWIN;                                                                                    \
    _Pragma("unroll") for (int u = 0; u < LS_QWIN; ++u)                                              \
        nxt[u] = (!(MASKED) || idx + u < end) ? p[u] : make_int2(0, 0);                              \
    _Pragma("unroll") for (int u = 0; u < LS_QWIN; ++u) {                                            \
      if (cur[u].x & LS_REC_VALID) {                                                                 \
        const float wt = __int_as_float(cur[u].y);                                                   \
        LS_ACC8(wt, fa[u], fb[u]);                                                                   \
      }                                                                                              \
      if (cur[u].x & LS_REC_LAST) {                                                                  \
        const unsigned cl__ = (unsigned)cur[u].x & 255u;                                             \
        TO* g = kStrip ? reinterpret_cast<TO*>(const_cast<char*>(lane0b) + (unsigned long long)cl__ * pitch_bytes) \
                       : lane0 + row_off(cl__);                                                      \
        LS_ACC_GET();                                                                                \
        ls_row_store4<kVec>(g, acc0);                                                                \
        ls_row_store4<kVec>(g + kSecond, acc1);                                                      \
        LS_ACC_ZERO();                                                                               \
      }                                                                                              \
    }                                                                                                \
  }
  // full windows first: while the window after the next still lies inside the piece, the record prefetch
  // needs no bounds test (4 of ~27 instructions per record); the masked form finishes the piece
  while (LS_SPLAT_FASTLOOP && idx + 3 * LS_QWIN <= end) {
    LS_SPLATD_WINDOW(r, rn, 0);
    LS_SPLATD_WINDOW(rn, r, 0);
  }
  for (;;) {
    LS_SPLATD_WINDOW(r, rn, 1);
    if (idx >= end) break;
    LS_SPLATD_WINDOW(rn, r, 1);
    if (idx >= end) break;
  }
#undef LS_SPLATD_WINDOW
#undef LS_ACC8
#undef LS_ACC_GET
#undef LS_ACC_ZERO
}

int ls_debug_fetch_phase_cycles(unsigned long long* out8) {
#ifdef LS_PROFILE
  unsigned long long zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  LS_CUDA(cudaMemcpyFromSymbol(out8, ls_dbg_phase, sizeof(zero)));
  LS_CUDA(cudaMemcpyToSymbol(ls_dbg_phase, zero, sizeof(zero)));
  if (FILE* f = fopen("gpurun_out/cta_phases.bin", "wb")) {
    static int host[8192][8];
    if (cudaMemcpyFromSymbol(host, ls_dbg_cta, sizeof(host)) == cudaSuccess) fwrite(host, 1, sizeof(host), f);
    fclose(f);
  }
  return LS_OK;
#else
  for (int i = 0; i < 8; ++i) out8[i] = 0;
  return LS_ERR_UNSUPPORTED;
#endif
}

size_t ls_sorted_records_capacity(const LsDims& dm, const LsGrid& g) { (void)g; return ls_sorted_capacity(dm.Npts); }

static size_t ls_tile_smem_bytes(const LsDims& dm) {
  const LsTileGeom tg = ls_tile_geom(dm.Cp);
  return (size_t)LS_TILE * tg.stride * sizeof(float) + (LS_QWARPS + 4) * sizeof(int);
}
static size_t ls_tile_smem_max() {
  LsDims d;
  d.Cp = LS_CCHUNK;
  return ls_tile_smem_bytes(d);
}

static bool ls_bev_vec4(const float* p, const LsBevStrides& st, const LsGrid& g) {
  return ((uintptr_t)p % 16 == 0) && (st.b % 4 == 0) && (st.c % 4 == 0) && (st.x % 4 == 0) && (g.Y % 4 == 0);
}

// How the splat writes a BEV tensor with these strides (include/ls_b200.h LsBevStrides).
int ls_classify_bev_out(const float* p, const LsBevStrides& st, const LsDims& dm, const LsGrid& g) {
  if (dm.bev_bf16) {          // strides in bf16 elements: rows must be 16-byte aligned
    const bool ok = st.c == 1 && dm.C == 64 && st.y >= 64 && st.y % 8 == 0 && st.x % 8 == 0 && st.b % 8 == 0 &&
                    (uintptr_t)p % 16 == 0 && st.x < (1LL << 32) && st.y < (1LL << 32);
    return ok ? LS_OUT_NHWC_DIRECT_VEC : LS_OUT_BAD;
  }
  if (st.y == 1 && (st.c != 1 || dm.C == 1)) return ls_bev_vec4(p, st, g) ? LS_OUT_NCHW_VEC4 : LS_OUT_NCHW_SCALAR;
  if (st.c == 1) {
    const bool vec = (uintptr_t)p % 16 == 0 && st.b % 4 == 0 && st.x % 4 == 0 && st.y % 4 == 0;
    if (dm.C == 64 && st.y == 64 && vec) return LS_OUT_NHWC_BULK;            // dense rows: bulk or direct
    if (dm.C == 64 && st.y >= 64 && st.x < (1LL << 32) && st.y < (1LL << 32))
      return vec ? LS_OUT_NHWC_DIRECT_VEC : LS_OUT_NHWC_DIRECT_SCALAR;      // 64 channels of a wider row
    return LS_OUT_NHWC_ROWS;
  }
  return LS_OUT_BAD;
}

// cudaFuncSetAttribute is per device: remember which devices have seen it
static bool ls_attr_needed(unsigned long long* done_mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
  const unsigned long long bit = 1ULL << dev;
  const unsigned long long prev = __atomic_fetch_or(done_mask, bit, __ATOMIC_RELAXED);
  return !(prev & bit);
}

// recs == nullptr: recs_sorted already holds the canonical records (static-rig cache), no re-ordering pass
template <typename T>
static int ls_splat_dispatch(const void* featT, const int2* recs, const int* seg_start, const int* tile_order,
                             int2* recs_sorted, int* perm, const LsDims& dm, const LsGrid& g, float* bev,
                             const LsBevStrides& st, cudaStream_t s) {
  static unsigned long long attr_done = 0;
  if (ls_attr_needed(&attr_done)) {
    const int m = (int)ls_tile_smem_max();
    LS_CUDA(cudaFuncSetAttribute(ls_splat_fwd_kernel<T, LS_OUT_NCHW_VEC4, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, m));
    LS_CUDA(cudaFuncSetAttribute(ls_splat_fwd_kernel<T, LS_OUT_NCHW_VEC4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, m));
    LS_CUDA(cudaFuncSetAttribute(ls_splat_fwd_kernel<T, LS_OUT_NCHW_SCALAR, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, m));
    LS_CUDA(cudaFuncSetAttribute(ls_splat_fwd_kernel<T, LS_OUT_NHWC_BULK, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, m));
    LS_CUDA(cudaFuncSetAttribute(ls_splat_fwd_kernel<T, LS_OUT_NHWC_ROWS, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, m));
  }
  const int out = ls_classify_bev_out(bev, st, dm, g);
  if (out == LS_OUT_BAD) return LS_ERR_UNSUPPORTED;
  size_t smem = ls_tile_smem_bytes(dm);
  dim3 grid(g.tiles * dm.B);
  // dense channels-last rows: direct row stores (no shared-memory tile, L1 kept for the feature
  // rows) unless LS_SPLAT_OUT=bulk asks for the one-bulk-store-per-tile (TMA) variant
  static const bool want_bulk = getenv("LS_SPLAT_OUT") && !strcmp(getenv("LS_SPLAT_OUT"), "bulk");
  const bool direct = out == LS_OUT_NHWC_DIRECT_VEC || out == LS_OUT_NHWC_DIRECT_SCALAR ||
                      (out == LS_OUT_NHWC_BULK && !(want_bulk && g.tx == 1));
  if (recs)
    LS_LAUNCH(ls_canon_kernel, grid, dim3(LS_CANON_THREADS), 0, s, const_cast<int2*>(recs), seg_start, tile_order, dm, g,
              recs_sorted, perm);
  const dim3 block(LS_THREADS);
  const int2* rs = recs_sorted;
#define LS_SPLAT(OUT, CC)                                                                                          \
  LS_LAUNCH((ls_splat_fwd_kernel<T, OUT, CC>), grid, block, smem, s, (const T*)featT, seg_start, tile_order, rs, dm, g, \
            bev, st)
  if (!direct && g.tx != 1) return LS_ERR_UNSUPPORTED;      // the tile kernels know 1 x 128 strips only
#define LS_DIRECT(VEC, TOUT, PTR)                                                                                      \
  do {                                                                                                                 \
    if (g.tx == 1)                                                                                                     \
      LS_LAUNCH((ls_splat_fwd_direct_kernel<T, VEC, TOUT, true>), grid, block, 0, s, (const T*)featT, seg_start, tile_order, \
                rs, dm, g, PTR, st);                                                                                   \
    else                                                                                                               \
      LS_LAUNCH((ls_splat_fwd_direct_kernel<T, VEC, TOUT, false>), grid, block, 0, s, (const T*)featT, seg_start,       \
                tile_order, rs, dm, g, PTR, st);                                                                       \
  } while (0)
  if (dm.bev_bf16) {          // opt-in bf16 BEV: the classifier only lets dense, aligned 64-channel rows through
    LS_DIRECT(true, __nv_bfloat16, reinterpret_cast<__nv_bfloat16*>(bev));
  } else if (direct && out != LS_OUT_NHWC_DIRECT_SCALAR) {
    LS_DIRECT(true, float, bev);
  } else if (direct) {
    LS_DIRECT(false, float, bev);
  } else if (out == LS_OUT_NHWC_BULK) {
    smem = (size_t)LS_TILE * 64 * sizeof(float) + (LS_QWARPS + 4) * sizeof(int);
    LS_SPLAT(LS_OUT_NHWC_BULK, 64);
  } else if (out == LS_OUT_NHWC_ROWS) {
    LS_SPLAT(LS_OUT_NHWC_ROWS, 0);
  } else if (out == LS_OUT_NCHW_VEC4 && dm.Cp == 64 && dm.C == 64) {
    LS_SPLAT(LS_OUT_NCHW_VEC4, 64);
  } else if (out == LS_OUT_NCHW_VEC4) {
    LS_SPLAT(LS_OUT_NCHW_VEC4, 0);
  } else {
    LS_SPLAT(LS_OUT_NCHW_SCALAR, 0);
  }
#undef LS_SPLAT
#undef LS_DIRECT
  return LS_OK;
}

int ls_launch_splat_fwd(const void* featT, int dtype, const int2* recs, const int* seg_start, const int* tile_order,
                        int2* recs_sorted, int* perm, const LsDims& dm, const LsGrid& g, float* bev,
                        const LsBevStrides& st, cudaStream_t s) {
  if (dtype == LS_F32)
    return ls_splat_dispatch<float>(featT, recs, seg_start, tile_order, recs_sorted, perm, dm, g, bev, st, s);
  return ls_splat_dispatch<__nv_bfloat16>(featT, recs, seg_start, tile_order, recs_sorted, perm, dm, g, bev, st, s);
}

// =====================================================================================
// Static-rig cache (opt-in): the camera rig of the reference's data never moves
// (dataset/carla_dataset.py:392-393), so voxel index, counting sort and canonical order of a
// batch are the same every step; only the depth probabilities carried by the records change.
// Refresh = two streaming passes:  recs_sorted[slot].y = prob[perm[slot]]  (canonical records)
// and  pix_recs[pix][d].y = prob[d][pix]  (pixel-major index of the backward).
// =====================================================================================
template <typename T>
__global__ void __launch_bounds__(256)
ls_refresh_records_kernel(const T* __restrict__ prob, const int* __restrict__ perm, const int* __restrict__ seg_start,
                          LsDims dm, LsGrid grid, int2* __restrict__ recs_sorted, int* __restrict__ zero_ints, int n_zero) {
  ls_pdl_trigger();
  ls_pdl_wait();
  const int b = blockIdx.y;
  if (b == 0)      // the backward's per-image arrival counters start from zero
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_zero; i += gridDim.x * blockDim.x) zero_ints[i] = 0;
  const int kept = __ldg(seg_start + (size_t)b * grid.seg_stride + grid.Vc);
  const int* pm = perm + (size_t)b * dm.Npts;
  const T* pb = prob + (size_t)b * dm.Npts;
  int2* rs = recs_sorted + (size_t)b * ls_sorted_capacity(dm.Npts);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kept; i += gridDim.x * blockDim.x)
    rs[i].y = __float_as_int(ls_to_float(pb[__ldg(pm + i)]));
}

template <typename T>
__global__ void __launch_bounds__(256)
ls_refresh_pixel_index_kernel(const T* __restrict__ prob, LsDims dm, int2* __restrict__ pix_recs) {
  ls_pdl_trigger();
  ls_pdl_wait();
  extern __shared__ int wstage[];                 // [32][Dp]
  const int Dp = dm.D | 1;
  const int bn = blockIdx.y, rc0 = blockIdx.x * 32;
  const int lane = threadIdx.x & 31, dg = threadIdx.x >> 5;
  const int valid = min(32, dm.HW - rc0);
  if (lane < valid)
    for (int d = dg; d < dm.D; d += 8)
      wstage[lane * Dp + d] = __float_as_int(ls_to_float(prob[((size_t)bn * dm.D + d) * dm.HW + rc0 + lane]));
  __syncthreads();
  int2* dst = pix_recs + ((size_t)bn * dm.HW + rc0) * dm.D;
  for (int r = dg; r < valid; r += 8)
    for (int d = lane; d < dm.D; d += 32) dst[(size_t)r * dm.D + d].y = wstage[r * Dp + d];
}

int ls_launch_refresh(const void* prob, int dtype, const int* perm, const int* seg_start, const LsDims& dm,
                      const LsGrid& g, int2* recs_sorted, int2* pix_recs, int* zero_ints, int n_zero, cudaStream_t s) {
  if (!zero_ints) n_zero = 0;
  dim3 ga((dm.Npts + 256 * 8 - 1) / (256 * 8), dm.B);
  dim3 gb((dm.HW + 31) / 32, dm.B * dm.N);
  const size_t smem = (size_t)32 * (dm.D | 1) * sizeof(int);
  if (dtype == LS_F32) {
    LS_LAUNCH(ls_refresh_records_kernel<float>, ga, dim3(256), 0, s, (const float*)prob, perm, seg_start, dm, g, recs_sorted,
              zero_ints, n_zero);
    if (pix_recs) LS_LAUNCH(ls_refresh_pixel_index_kernel<float>, gb, dim3(256), smem, s, (const float*)prob, dm, pix_recs);
  } else {
    LS_LAUNCH(ls_refresh_records_kernel<__nv_bfloat16>, ga, dim3(256), 0, s, (const __nv_bfloat16*)prob, perm, seg_start,
              dm, g, recs_sorted, zero_ints, n_zero);
    if (pix_recs)
      LS_LAUNCH(ls_refresh_pixel_index_kernel<__nv_bfloat16>, gb, dim3(256), smem, s, (const __nv_bfloat16*)prob, dm, pix_recs);
  }
  return LS_OK;
}

// =====================================================================================
// K4a (NCHW gradients only): grad_bev [B,C,X,Y] -> cell-major gT [B, X*Y + 1, Cp] in rank order
// (row X*Y = zeros: where dropped points gather from); rows of cells nobody hit are skipped
// (they are never read).  Same tile / swizzle as the forward write-out, run backwards.
// A channels-last gradient needs none of this: its rows are gathered in place.
// =====================================================================================
#ifndef LS_GOCC_ROWS
#define LS_GOCC_ROWS 8    // rows gathered at a time by a half-warp of the register-lean gather
#endif
#ifndef LS_GOCC_MINB
#define LS_GOCC_MINB 4    // its CTAs per SM (32 warps: 64 registers; 3 CTAs at 80 registers measured 2 us slower)
#endif
#ifndef LS_GATHER_PF1
#define LS_GATHER_PF1 1   // prefetch.global.L1 (SASS CCTL.E.PF1) of the NEXT batch's gradient rows: gather 71.7 -> 69.7 us; two batches ahead (2): 73 us
#endif
#ifndef LS_GOCC_BFLY8
#define LS_GOCC_BFLY8 1   // reduce each batch of 8 dot products right away (needs LS_GOCC_ROWS == 8, LS_GATHER_SKIP_DEAD != 2): what lets 8 rows in flight fit 64 registers
#endif
#ifndef LS_GATHER_OCC
#define LS_GATHER_OCC 1   // register-lean gradient gather (32 warps/SM) for Cp <= 64, D % 16 == 0
#endif
#ifndef LS_GATHER_REVERSE
#define LS_GATHER_REVERSE 1
#endif
#ifndef LS_GATHER_SKIP_DEAD
#define LS_GATHER_SKIP_DEAD 1
#endif
#ifndef LS_GATHER_L2PF
#define LS_GATHER_L2PF 0    // prefetch.global.L2 of the next window's gradient rows, one window ahead: gather 74 -> 82 us (measured), off
#endif
#ifndef LS_TCHUNK
#define LS_TCHUNK 32   // channels per CTA of the gradient transposer (16, 32 or 64): 8 x 16 B in flight per thread
#endif

template <bool VEC4>
__global__ void __launch_bounds__(LS_THREADS)
ls_bwd_transpose_kernel(const float* __restrict__ gbev, LsBevStrides st, const int* __restrict__ seg_start,
                        LsDims dm, LsGrid grid, float* __restrict__ gT) {
  ls_pdl_trigger();
  ls_pdl_wait();
  extern __shared__ float smem[];
  const LsTileGeom tg = ls_tile_geom(min(dm.Cp, LS_TCHUNK));
  float* tile = smem;
  int* hit = reinterpret_cast<int*>(smem + LS_TILE * tg.stride);      // [LS_TILE] does anybody read this cell's row?
  // this kernel's own strips: 128 consecutive y of one x-row, whatever tiling the forward used
  const int strips_y = (grid.Y + LS_TY - 1) / LS_TY;
  const int b = blockIdx.y, strip = blockIdx.x, tid = threadIdx.x;
  const int cbase = blockIdx.z * LS_TCHUNK;
  const int tx0 = strip / strips_y, ty0 = (strip % strips_y) * LS_TY;
  const int nquads = min(tg.cc, dm.Cp - cbase) >> 2;
  // thread = (4 consecutive y, x-row, channel quad): four 16-byte loads (4 channels) per pass.
  // The gradient loads do not depend on the cell offsets, so they are issued first and
  // both round trips overlap.
  const int y4 = tid % (LS_TY / 4), q0 = tid / (LS_TILE / 4);
  const int gx = tx0, gy = ty0 + 4 * y4;
  constexpr int kPasses = LS_TCHUNK / 16;
  float4 c[kPasses][4];
  if (VEC4) {
#pragma unroll
    for (int ps = 0; ps < kPasses; ++ps) {
      const int q = q0 + 4 * ps;
#pragma unroll
      for (int k = 0; k < 4; ++k) c[ps][k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q < nquads && gx < grid.X && gy < grid.Y) {
        const int ch = cbase + 4 * q;
        const float* g = gbev + (size_t)b * st.b + (size_t)ch * st.c + (size_t)gx * st.x + gy;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (ch + k < dm.C) c[ps][k] = __ldg(reinterpret_cast<const float4*>(g + (size_t)k * st.c));
      }
    }
  }
  {
    // cell tid of the strip in the forward's tile-major numbering: hit by any point?
    const int cy = ty0 + tid;
    int h = 0;
    if (cy < grid.Y) {
      const int* seg = seg_start + (size_t)b * grid.seg_stride + ls_cell_of_xy(gx, cy, grid);
      h = __ldg(seg + 1) != __ldg(seg);
    }
    hit[tid] = h;
    if (strip == 0 && blockIdx.z == 0) {   // row X*Y of every sample = zeros: where dropped points gather from
      float* zrow = gT + ((size_t)b * (grid.XY + 1) + grid.XY) * dm.Cp;
      for (int i = tid; i < dm.Cp; i += LS_THREADS) zrow[i] = 0.0f;
    }
    if (!__syncthreads_or(h)) return;          // nobody reads this strip's gradient
  }
  float* dst = gT + ((size_t)b * (grid.XY + 1) + (size_t)gx * grid.Y + ty0) * dm.Cp + cbase;     // + cl * Cp
  if (VEC4) {
    // 4x4 register transpose, four 16-byte conflict-free shared stores (4 cells, one quad each)
    const int clc = 4 * y4;
    const int swz = (clc >> 3) & (tg.nqp - 1);
#pragma unroll
    for (int ps = 0; ps < kPasses; ++ps) {
      const int q = q0 + 4 * ps;
      if (q < nquads) {
        float* d = tile + clc * tg.stride + 4 * (q ^ swz);
        *reinterpret_cast<float4*>(d) = make_float4(c[ps][0].x, c[ps][1].x, c[ps][2].x, c[ps][3].x);
        *reinterpret_cast<float4*>(d + tg.stride) = make_float4(c[ps][0].y, c[ps][1].y, c[ps][2].y, c[ps][3].y);
        *reinterpret_cast<float4*>(d + 2 * tg.stride) = make_float4(c[ps][0].z, c[ps][1].z, c[ps][2].z, c[ps][3].z);
        *reinterpret_cast<float4*>(d + 3 * tg.stride) = make_float4(c[ps][0].w, c[ps][1].w, c[ps][2].w, c[ps][3].w);
      }
    }
  } else {
    for (int idx = tid; idx < 4 * nquads * LS_TILE; idx += LS_THREADS) {
      const int y = idx % LS_TY, cr = idx / LS_TILE;
      const int ch = cbase + cr;
      const int oy = ty0 + y;
      float v = 0.0f;
      if (ch < dm.C && gx < grid.X && oy < grid.Y)
        v = gbev[(size_t)b * st.b + (size_t)ch * st.c + (size_t)gx * st.x + oy];
      tile[y * tg.stride + 4 * ls_tile_quad(y, cr >> 2, tg.nqp) + (cr & 3)] = v;
    }
  }
  __syncthreads();
  // rows of hit cells: thread = (quad, cell), 16 B per lane, LS_TCHUNK*4 B per cell
  {
    constexpr int kQ = LS_TCHUNK / 4;
    const int q = tid % kQ;
    if (q < nquads) {
#pragma unroll 4
      for (int cl = tid / kQ; cl < LS_TILE; cl += LS_THREADS / kQ) {
        if (hit[cl])
          *reinterpret_cast<float4*>(dst + (size_t)cl * dm.Cp + 4 * q) =
              *reinterpret_cast<const float4*>(tile + cl * tg.stride + 4 * ls_tile_quad(cl, q, tg.nqp));
      }
    }
  }
}

int ls_launch_bwd_transpose(const float* gbev, const LsBevStrides& st, const int* seg_start, const LsDims& dm,
                            const LsGrid& g, float* gT, cudaStream_t s) {
  const LsTileGeom tg = ls_tile_geom(dm.Cp < LS_TCHUNK ? dm.Cp : LS_TCHUNK);
  const size_t smem = (size_t)LS_TILE * tg.stride * sizeof(float) + (LS_TILE + 1) * sizeof(int);
  dim3 grid(g.X * ((g.Y + LS_TY - 1) / LS_TY), dm.B, (dm.Cp + LS_TCHUNK - 1) / LS_TCHUNK);
  if (ls_bev_vec4(gbev, st, g))
    LS_LAUNCH(ls_bwd_transpose_kernel<true>, grid, dim3(LS_THREADS), smem, s, gbev, st, seg_start, dm, g, gT);
  else
    LS_LAUNCH(ls_bwd_transpose_kernel<false>, grid, dim3(LS_THREADS), smem, s, gbev, st, seg_start, dm, g, gT);
  return LS_OK;
}

// =====================================================================================
// K4b: gradient gather, pixel-stationary (deterministic, no atomics).
// reference: VoxelsSumming.backward tool/geometry.py:307-317 + autograd of
// model/bev_model.py:66,91-97.  CTA = one feature-map column of one camera (its rays sweep one
// radial line of the BEV, so the gradient rows it gathers are re-used from L1);
// a half-warp owns a pixel: 16 lanes x float4 channels, depth bins in windows of 16:
//   gf[c]  += prob[d] * g[cell(d), c]                     (registers, d ascending)
//   gp[d]   = sum_c feat[c] * g[cell(d), c]               (16 dots reduced together by a
//                                                          transposing butterfly: 15 shuffles)
// Outputs: grad_feat NHWC-padded [pix][Cp] and grad_prob PIXEL-major [pix][D].
// Where the rows g[cell] come from (MODE, enum LsGradIn):
//   STAGED         the cell-major copy of an NCHW gradient (row X*Y = zeros for dropped points)
//   DIRECT_VEC     the channels-last gradient tensor itself, 16-byte aligned rows
//   DIRECT_SCALAR  the same with unaligned rows (a 64-channel slice of a 65-channel tensor)
// pix_recs.x is the rank (row number) of the point's cell, X*Y for a dropped point.
// =====================================================================================
// An all-zero row (never written): where dropped points "gather" from when the rows are read in
// place.  Selecting the ADDRESS keeps every row load unconditional, so the compiler keeps all of
// a window's loads in flight (a predicated load needs its destination zeroed first, and ptxas
// then serialised the window into pairs: 2x slower).
__device__ float4 ls_zero_row[4 * LS_CCHUNK / 4];

// TG: element type of the gradient rows (float; __nv_bfloat16 for the opt-in bf16 BEV, direct-vector mode only)
template <int MODE, typename TG>
__device__ __forceinline__ float4 ls_grad_row4(const char* __restrict__ base, const char* __restrict__ zero,
                                               unsigned rank, unsigned row_bytes, unsigned nrows) {
  if (MODE == LS_GRAD_STAGED) return __ldg(reinterpret_cast<const float4*>(base + (size_t)rank * row_bytes));
  const char* p = rank < nrows ? base + (size_t)rank * row_bytes : zero;
  if (sizeof(TG) == 2) return ls_load4<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(p));
  if (MODE == LS_GRAD_DIRECT_VEC) return __ldg(reinterpret_cast<const float4*>(p));
  const float* f = reinterpret_cast<const float*>(p);
  return make_float4(__ldg(f), __ldg(f + 1), __ldg(f + 2), __ldg(f + 3));
}

// transposing butterfly over a half-warp: lane hl ends up with the sum over the 16 lanes of dot[hl]
__device__ __forceinline__ float ls_half_butterfly(float (&dot)[16], int hl, unsigned hmask) {
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
  const bool hi = hl & 1;
  const float send = hi ? dot[0] : dot[1];
  const float keep = hi ? dot[1] : dot[0];
  return keep + __shfl_xor_sync(hmask, send, 1, 16);
}

// the same tree for EIGHT values (one batch of rows of the register-lean gather): the xor-8, -4, -2 levels leave row
// hl >> 1 in each lane pair, the xor-1 level adds the pair's two partial sums (a + b == b + a: same bits as above)
__device__ __forceinline__ float ls_half_butterfly8(float (&dot)[8], int hl, unsigned hmask) {
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const bool hi = hl & 8;
    const float send = hi ? dot[u] : dot[u + 4];
    const float keep = hi ? dot[u + 4] : dot[u];
    dot[u] = keep + __shfl_xor_sync(hmask, send, 8, 16);
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const bool hi = hl & 4;
    const float send = hi ? dot[u] : dot[u + 2];
    const float keep = hi ? dot[u + 2] : dot[u];
    dot[u] = keep + __shfl_xor_sync(hmask, send, 4, 16);
  }
  const bool hi = hl & 2;
  const float send = hi ? dot[0] : dot[1];
  const float keep = hi ? dot[1] : dot[0];
  const float v = keep + __shfl_xor_sync(hmask, send, 2, 16);
  return v + __shfl_xor_sync(hmask, v, 1, 16);
}

// Where a sample's gradient rows live: base + b * sample_stride floats, rows row_bytes apart.
struct LsRows {
  const void* base;
  long long sample_stride;     // elements
  unsigned row_bytes;
  unsigned nrows;       // X*Y
};

// General shape: any D, up to 4 chunks of 64 channels.
template <typename T, int NCH, int MODE>
__global__ void __launch_bounds__(LS_GATHER_THREADS, LS_GATHER_MINB)
ls_bwd_gather_kernel(LsRows rows, const T* __restrict__ featT, const int2* __restrict__ pix_recs,
                     LsDims dm, float* __restrict__ gprob_pm, T* __restrict__ gfeatT) {
  ls_pdl_trigger();
  ls_pdl_wait();
  // images in reverse order: the rows written last (staged copy) are the ones still in L2
  const int col = blockIdx.x, bn = LS_GATHER_REVERSE ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
  const int b = bn / dm.N;
  const int hw = threadIdx.x >> 4, hl = threadIdx.x & 15;
  const unsigned hmask = ls_half_mask();
  bool on[NCH];
#pragma unroll
  for (int q = 0; q < NCH; ++q) on[q] = (q * LS_CCHUNK + 4 * hl) < dm.Cp;
  // lanes beyond the channel count read valid bytes (lane 0's) and never store
  const char* gb = reinterpret_cast<const char*>(reinterpret_cast<const float*>(rows.base) + (size_t)b * rows.sample_stride);
  const char* zrow = reinterpret_cast<const char*>(ls_zero_row);

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
      for (int u = 0; u < 16; ++u) r[u] = (u < n) ? __ldg(pr + d0 + u) : make_int2((int)rows.nrows, 0);   // half-warp-uniform
      float dot[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) dot[u] = 0.0f;
#pragma unroll
      for (int q = 0; q < NCH; ++q) {
        float4 g[16];
        const char* gq = gb + (on[q] ? (q * LS_CCHUNK + 4 * hl) * 4 : 0);
#pragma unroll
        for (int u = 0; u < 16; ++u) g[u] = ls_grad_row4<MODE, float>(gq, zrow, (unsigned)r[u].x, rows.row_bytes, rows.nrows);
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
      const float mine = ls_half_butterfly(dot, hl, hmask);
      if (hl < n) gprob_pm[pix * dm.D + d0 + hl] = mine;
    }
    T* grow = gfeatT + pix * dm.Cp + 4 * hl;
#pragma unroll
    for (int q = 0; q < NCH; ++q)
      if (on[q]) ls_store4<T>(grow + q * LS_CCHUNK, gf[q]);
  }
}

// Higher-occupancy variant for the common shape (Cp <= 64, D a multiple of 16).  The random
// 256-byte row gather from a gradient larger than L2 scales with the number of resident
// warps, not with the rows in flight per warp (tools/gather_bench.cu 400000: 7.8 / 11.0 / 13.7
// TB/s at 16 / 24 / 32 warps per SM), so this version trades registers for warps: the sixteen
// records of a depth window are loaded one per lane (a single coalesced 128-byte load per
// half-warp, prefetched a window ahead, broadcast with 16-wide shuffles) and the rows are
// gathered eight at a time, each batch's eight dot products reduced right away (ls_half_butterfly8) - 64 registers,
// four CTAs (32 warps) per SM; with the sixteen dot products of a window kept until its end it needed 80 registers
// (24 warps) and ran 2 us longer.
// ready != NULL: the backward's epilogue (ls_bwd_epilogue_kernel: softmax backward + layout fix-up of
// grad_feat) is the next launch in the stream and runs OVERLAPPED with this kernel's tail.  Every CTA
// lets the dependent launch be scheduled as soon as it has passed its own dependency wait; the
// epilogue's CTAs then take the SM slots this grid's last wave leaves free and wait, per image, on
// ready[image]: each warp adds 1 (release) when its rows are written, fw * warps-per-CTA arrivals = image done.
template <typename T, int MODE, typename TG>
__global__ void __launch_bounds__(LS_GATHER_THREADS, LS_GOCC_MINB)
ls_bwd_gather_occ_kernel(LsRows rows, const T* __restrict__ featT, const int2* __restrict__ pix_recs,
                         LsDims dm, float* __restrict__ gprob_pm, T* __restrict__ gfeatT, int* __restrict__ ready) {
  ls_pdl_trigger();
  ls_pdl_wait();
  if (ready) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int col = blockIdx.x, bn = LS_GATHER_REVERSE ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
  const int b = bn / dm.N;
  const int hw = threadIdx.x >> 4, hl = threadIdx.x & 15;
  const unsigned hmask = ls_half_mask();
  const bool on = 4 * hl < dm.Cp;
  // lanes beyond the channel count read valid bytes (lane 0's) and never store
  const char* gb = reinterpret_cast<const char*>(reinterpret_cast<const TG*>(rows.base) + (size_t)b * rows.sample_stride +
                                                 (on ? 4 * hl : 0));
  const char* zrow = reinterpret_cast<const char*>(ls_zero_row);
  const int wpp = dm.D >> 4;                                                  // windows per pixel
#if LS_GATHER_L2PF
  // Every lane holds one record of the NEXT window a whole window ahead (the next pixel's first window
  // included): it asks L2 for that record's gradient row now (two 128-byte lines of a 256-byte row), one
  // window of arithmetic before the half-warp gathers it - the row loads then find most of their
  // 170 MB of cold DRAM rows in L2.  Two instructions per lane and window.
  const char* grow0 = reinterpret_cast<const char*>(reinterpret_cast<const TG*>(rows.base) + (size_t)b * rows.sample_stride);
  const unsigned row_len = (unsigned)(dm.Cp * sizeof(TG));
  auto l2_prefetch = [&](int2 r) {
    if ((unsigned)r.x < rows.nrows) {
      const char* p = grow0 + (size_t)(unsigned)r.x * rows.row_bytes;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
      if (row_len > 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 128));
    }
  };
  int2 rec_first = make_int2((int)rows.nrows, 0);
  if (hw < dm.fh) {
    rec_first = __ldg(pix_recs + ((size_t)bn * dm.HW + (size_t)hw * dm.fw + col) * dm.D + hl);
    l2_prefetch(rec_first);
  }
#endif
  for (int row = hw; row < dm.fh; row += LS_HALFWARPS) {
    const size_t pix = (size_t)bn * dm.HW + (size_t)row * dm.fw + col;
    const float4 f = on ? ls_load4<T>(featT + pix * dm.Cp + 4 * hl) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 gf = make_float4(0.f, 0.f, 0.f, 0.f);
#if LS_FFMA2
    const unsigned long long fxy = ls_pack2(f.x, f.y), fzw = ls_pack2(f.z, f.w);
    unsigned long long gfxy = ls_pack2(0.f, 0.f), gfzw = gfxy;
#endif
    const int2* pr = pix_recs + pix * dm.D + hl;
#if LS_GATHER_L2PF
    int2 rec = rec_first;
#else
    int2 rec = __ldg(pr);
#endif
    for (int w = 0; w < wpp; ++w) {
      int2 recn = rec;
      if (w + 1 < wpp) recn = __ldg(pr + 16 * (w + 1));
#if LS_GATHER_L2PF
      else if (row + LS_HALFWARPS < dm.fh)      // last window: the first record of this half-warp's next pixel
        recn = rec_first = __ldg(pr + (size_t)LS_HALFWARPS * dm.fw * dm.D);
      if (w + 1 < wpp || row + LS_HALFWARPS < dm.fh) l2_prefetch(recn);
#endif
#if LS_GATHER_SKIP_DEAD
      // a ray leaves the grid at some depth and stays out: whole windows of dropped points (about one in
      // seven at the default rig) need no rows, no FMAs, no butterfly - their probability gradient is 0
      // (LS_GATHER_SKIP_DEAD == 2: the same test per batch of LS_GOCC_ROWS rows - the window a ray leaves the grid in)
      const unsigned kept16 = (__ballot_sync(hmask, (unsigned)rec.x < rows.nrows) >> (threadIdx.x & 16)) & 0xFFFFu;
      if (!kept16) {
        gprob_pm[pix * dm.D + 16 * w + hl] = 0.0f;
        rec = recn;
        continue;
      }
#endif
#if LS_GOCC_BFLY8
      // the eight dot products of a batch are reduced right after the batch (same butterfly tree, same bits):
      // eight live registers instead of sixteen across the second batch's loads
#define LS_DOT(i) dot8[(i) - LS_GOCC_ROWS * h]
#else
      float dot[16];
#define LS_DOT(i) dot[i]
#endif
#pragma unroll
      for (int h = 0; h < 16 / LS_GOCC_ROWS; ++h) {
#if LS_GOCC_BFLY8
        float dot8[8];
#endif
#if LS_GATHER_SKIP_DEAD == 2
        if (!((kept16 >> (LS_GOCC_ROWS * h)) & ((1u << LS_GOCC_ROWS) - 1u))) {      // half-warp uniform
#pragma unroll
          for (int u = 0; u < LS_GOCC_ROWS; ++u) dot[LS_GOCC_ROWS * h + u] = 0.0f;
          continue;
        }
#endif
        float4 g[LS_GOCC_ROWS];
#pragma unroll
        for (int u = 0; u < LS_GOCC_ROWS; ++u) {
          const unsigned rank = (unsigned)__shfl_sync(hmask, rec.x, LS_GOCC_ROWS * h + u, 16);
#if LS_ABLATE == 2        // developer ablation: no gradient-row traffic, same arithmetic
          g[u] = make_float4(__uint_as_float(rank), 1.f, 2.f, 3.f);
#else
          g[u] = ls_grad_row4<MODE, TG>(gb, zrow, rank, rows.row_bytes, rows.nrows);
#endif
        }
#if LS_GATHER_PF1
        // The rows of the NEXT batch of eight are asked into L1 now, one row (two 128-byte lines) per lane that holds
        // its record: lanes 8-15 of this window for its second batch, lanes 0-7 of the next window for its first.
        // No registers: the batch in flight lives in registers, the one behind it in L1.
        {
#if LS_GATHER_PF1 == 2      // two batches ahead: the same batch of the next window
          const unsigned nx = (unsigned)recn.x;
          const bool mine = ((h == 0) ? (hl < 8) : (hl >= 8)) && w + 1 < wpp;
#else
          const unsigned nx = (h == 0) ? (unsigned)rec.x : (unsigned)recn.x;
          const bool mine = (h == 0) ? (hl >= 8) : (hl < 8 && w + 1 < wpp);
#endif
          if (mine && nx < rows.nrows) {
            const char* pfp = gb - (on ? 4 * hl * (int)sizeof(TG) : 0) + (size_t)nx * rows.row_bytes;
            asm volatile("prefetch.global.L1 [%0];" ::"l"(pfp));
            if (sizeof(TG) == 4) asm volatile("prefetch.global.L1 [%0];" ::"l"(pfp + 128));
          }
        }
#endif
#pragma unroll
        for (int u = 0; u < LS_GOCC_ROWS; ++u) {
          const float wgt = __int_as_float(__shfl_sync(hmask, rec.y, LS_GOCC_ROWS * h + u, 16));
#if LS_ABLATE == 1        // developer ablation: every load and store, one add instead of nine multiply-adds per row
          LS_DOT(LS_GOCC_ROWS * h + u) = wgt;
          gf.x = __int_as_float(__float_as_int(gf.x) ^ __float_as_int(g[u].x) ^ __float_as_int(g[u].y));
          gf.y = __int_as_float(__float_as_int(gf.y) ^ __float_as_int(g[u].z) ^ __float_as_int(g[u].w));
#elif LS_FFMA2
          // packed pairs: (x,y) and (z,w) of the row take one FFMA2 each for the feature gradient and
          // one FMUL2 + one FFMA2 for the dot product (6 issue slots per row instead of 8)
          const unsigned long long gxy = ls_pack2(g[u].x, g[u].y), gzw = ls_pack2(g[u].z, g[u].w);
          const unsigned long long w2 = ls_pack2(wgt, wgt);
          float d0, d1;
          ls_unpack2(ls_fma2(fzw, gzw, ls_mul2(fxy, gxy)), d0, d1);
          LS_DOT(LS_GOCC_ROWS * h + u) = d0 + d1;
          gfxy = ls_fma2(w2, gxy, gfxy);
          gfzw = ls_fma2(w2, gzw, gfzw);
#else
          float dv = f.x * g[u].x;
          dv = fmaf(f.y, g[u].y, dv);
          dv = fmaf(f.z, g[u].z, dv);
          dv = fmaf(f.w, g[u].w, dv);
          LS_DOT(LS_GOCC_ROWS * h + u) = dv;
          gf.x = fmaf(wgt, g[u].x, gf.x); gf.y = fmaf(wgt, g[u].y, gf.y);
          gf.z = fmaf(wgt, g[u].z, gf.z); gf.w = fmaf(wgt, g[u].w, gf.w);
#endif
        }
#if LS_GOCC_BFLY8
        const float v8 = ls_half_butterfly8(dot8, hl, hmask);
        if (!(hl & 1)) gprob_pm[pix * dm.D + 16 * w + 8 * h + (hl >> 1)] = v8;
#endif
      }
#undef LS_DOT
#if !LS_GOCC_BFLY8
      gprob_pm[pix * dm.D + 16 * w + hl] = ls_half_butterfly(dot, hl, hmask);
#endif
      rec = recn;
    }
#if LS_FFMA2
    ls_unpack2(gfxy, gf.x, gf.y);
    ls_unpack2(gfzw, gf.z, gf.w);
#endif
    if (on) ls_store4<T>(gfeatT + pix * dm.Cp + 4 * hl, gf);
  }
  if (ready) {
    // this warp's rows of image bn are written: one release-add per warp (the warp barrier orders the
    // other lanes' stores before lane 0's release)
    __syncwarp();
    if ((threadIdx.x & 31) == 0)
      asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(ready + bn) : "memory");
  }
}


// =====================================================================================
// Gradient gather through the Blackwell TMA gather4 path (opt-in, LS_GATHER_TMA=1).  Same work split, same
// arithmetic and the same bits as ls_bwd_gather_occ_kernel; what changes is how the 256-byte gradient rows
// reach the SM.  The LDG gather keeps 8 rows per half-warp in REGISTERS (128 KB in flight per SM at 32 warps),
// and `tools/tma_gather_bench.cu` shows the chip's random-row throughput growing with the bytes in flight
// per SM.  Here lane 0 of a warp hands the copy engine the sixteen row indices of a half depth window as four
// `cp.async.bulk.tensor.2d.tile::gather4` instructions (SASS UTMALDG.2D.GATHER4: four arbitrary rows of a 2-D
// tensor map per instruction) that land in a 4 KB shared-memory stage and complete on the stage's mbarrier;
// two stages per warp = 8 KB per warp, 192 KB per SM in flight with three CTAs, none of it in registers.
// Dropped points carry a row index beyond the tensor: the copy engine fills their rows with zeros.
// =====================================================================================
#define LS_TMA_STAGE_BYTES 4096      // 16 rows of 64 floats: rows 0-7 of half-warp 0, rows 8-15 of half-warp 1
__device__ __forceinline__ unsigned ls_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ls_mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nLS_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LS_DONE_%=;\nbra LS_WAIT_%=;\nLS_DONE_%=:\n}" ::"r"(bar), "r"(parity) : "memory");
}

template <typename T>
__global__ void __launch_bounds__(LS_GATHER_THREADS, 3)
ls_bwd_gather_tma_kernel(const __grid_constant__ CUtensorMap tm, unsigned rows_per_sample, unsigned nrows, unsigned oob_row,
                         const T* __restrict__ featT, const int2* __restrict__ pix_recs, LsDims dm,
                         float* __restrict__ gprob_pm, T* __restrict__ gfeatT) {
  extern __shared__ __align__(1024) unsigned char ls_tma_smem[];
  constexpr int kWarps = LS_GATHER_THREADS / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned ring = ls_smem_u32(ls_tma_smem) + (unsigned)warp * 2u * LS_TMA_STAGE_BYTES;
  const unsigned bar0 = ls_smem_u32(ls_tma_smem) + (unsigned)kWarps * 2u * LS_TMA_STAGE_BYTES + (unsigned)warp * 16u;
  if (lane == 0) {      // private prologue: the two stage barriers of this warp (one arrival each: the issuing lane)
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  ls_pdl_trigger();
  ls_pdl_wait();
  const int col = blockIdx.x, bn = LS_GATHER_REVERSE ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
  const int b = bn / dm.N;
  const int hw = threadIdx.x >> 4, hl = threadIdx.x & 15, h = lane >> 4;
  const unsigned hmask = ls_half_mask();
  const int wpp = dm.D >> 4;                                                  // windows per pixel
  const int npass = (dm.fh + LS_HALFWARPS - 1) / LS_HALFWARPS;
  const int nwin = npass * wpp;                                               // windows of this half-warp, all passes
  const unsigned row_base = (unsigned)b * rows_per_sample;
  const int2 dropped = make_int2((int)nrows, 0);
  // records of window k of this half-warp (lane hl: depth bin 16 * (k % wpp) + hl of the pixel of pass k / wpp)
  auto load_rec = [&](int k) -> int2 {
    const int row = hw + (k / wpp) * LS_HALFWARPS;
    if (k >= nwin || row >= dm.fh) return dropped;
    return __ldg(pix_recs + ((size_t)bn * dm.HW + (size_t)row * dm.fw + col) * dm.D + 16 * (k % wpp) + hl);
  };
  // Stage `half` <- rows [8 * half, 8 * half + 8) of both half-warps' windows.  All lanes take part in the
  // shuffles that bring the sixteen row indices to lane 0, which arms the barrier and issues four gather4 copies.
  auto issue = [&](const int2& rec, int half) {
    unsigned r[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const unsigned rk = (unsigned)__shfl_sync(0xffffffffu, rec.x, (j >> 3) * 16 + half * 8 + (j & 7));
      r[j] = rk < nrows ? row_base + rk : oob_row;
    }
    if (lane == 0) {
      const unsigned bar = bar0 + 8u * half, dst = ring + (unsigned)half * LS_TMA_STAGE_BYTES;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(LS_TMA_STAGE_BYTES) : "memory");
#pragma unroll
      for (int q = 0; q < 4; ++q)
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst + q * 1024u), "l"(&tm), "r"(bar), "r"(0),
            "r"(r[4 * q]), "r"(r[4 * q + 1]), "r"(r[4 * q + 2]), "r"(r[4 * q + 3])
            : "memory");
    }
  };
  int2 rec = load_rec(0);
  bool live = __ballot_sync(0xffffffffu, (unsigned)rec.x < nrows) != 0u;     // warp-uniform: any kept point in the window
  if (live) { issue(rec, 0); issue(rec, 1); }
  unsigned ncons = 0;                                                          // windows consumed = phase of both barriers
  float4 f = make_float4(0.f, 0.f, 0.f, 0.f), gf = f;
  for (int k = 0; k < nwin; ++k) {
    const int w = k % wpp, row = hw + (k / wpp) * LS_HALFWARPS;
    const bool have = row < dm.fh;
    const size_t pix = (size_t)bn * dm.HW + (size_t)(have ? row : 0) * dm.fw + col;
    const bool on = 4 * hl < dm.Cp;
    if (w == 0) {
      f = (have && on) ? ls_load4<T>(featT + pix * dm.Cp + 4 * hl) : make_float4(0.f, 0.f, 0.f, 0.f);
      gf = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int2 recn = load_rec(k + 1);
    const bool live_n = __ballot_sync(0xffffffffu, (unsigned)recn.x < nrows) != 0u;
    if (live) {
      float dot[16];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        ls_mbar_wait(bar0 + 8u * half, ncons & 1u);
        const float4* rows = reinterpret_cast<const float4*>(ls_tma_smem + (size_t)warp * 2 * LS_TMA_STAGE_BYTES +
                                                             (size_t)half * LS_TMA_STAGE_BYTES + (size_t)h * 2048) + hl;
        float4 g[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) g[u] = rows[u * 16];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float wgt = __int_as_float(__shfl_sync(hmask, rec.y, half * 8 + u, 16));
          float dv = f.x * g[u].x;
          dv = fmaf(f.y, g[u].y, dv);
          dv = fmaf(f.z, g[u].z, dv);
          dv = fmaf(f.w, g[u].w, dv);
          dot[half * 8 + u] = dv;
          gf.x = fmaf(wgt, g[u].x, gf.x); gf.y = fmaf(wgt, g[u].y, gf.y);
          gf.z = fmaf(wgt, g[u].z, gf.z); gf.w = fmaf(wgt, g[u].w, gf.w);
        }
        __syncwarp();                               // every lane has its rows in registers: the stage is free
        if (live_n) issue(recn, half);
      }
      ++ncons;
      const float mine = ls_half_butterfly(dot, hl, hmask);
      if (have) gprob_pm[pix * dm.D + 16 * w + hl] = mine;
    } else {
      if (have) gprob_pm[pix * dm.D + 16 * w + hl] = 0.0f;
      if (live_n) { issue(recn, 0); issue(recn, 1); }
    }
    if (w == wpp - 1 && have && on) ls_store4<T>(gfeatT + pix * dm.Cp + 4 * hl, gf);
    rec = recn;
    live = live_n;
  }
}

typedef CUresult (*LsEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// LS_OK and *launched = true when the TMA variant took the launch; *launched = false: the caller uses the LDG gather
template <typename T>
static int ls_gather_tma_try(const LsRows& rows, const void* featT, const int2* pix_recs, const LsDims& dm, float* gprob_pm,
                             void* gfeatT, cudaStream_t s, bool* launched) {
  *launched = false;
  static LsEncodeTiled encode = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess) fn = nullptr;
    (void)cudaGetLastError();
    return (LsEncodeTiled)fn;
  }();
  const long long row_elems = rows.row_bytes / 4;
  if (!encode || dm.Cp != 64 || dm.D % 16 != 0 || rows.row_bytes % 16 != 0 || (uintptr_t)rows.base % 16 != 0 ||
      row_elems <= 0 || rows.sample_stride % row_elems != 0)
    return LS_OK;
  const unsigned long long rps = (unsigned long long)(rows.sample_stride / row_elems);
  const unsigned long long total = rps * (unsigned long long)(dm.B - 1) + rows.nrows;
  if (rps < rows.nrows || total >= (1ULL << 31)) return LS_OK;
  CUtensorMap tm;
  const cuuint64_t gdim[2] = {64, (cuuint64_t)total};
  const cuuint64_t gstride[1] = {rows.row_bytes};
  const cuuint32_t box[2] = {64, 1};
  const cuuint32_t estr[2] = {1, 1};
  if (encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(rows.base), gdim, gstride, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return LS_OK;
  constexpr int kSmem = (LS_GATHER_THREADS / 32) * (2 * LS_TMA_STAGE_BYTES + 16);
  static unsigned long long attr_done = 0;
  if (ls_attr_needed(&attr_done))
    LS_CUDA(cudaFuncSetAttribute(ls_bwd_gather_tma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  LS_LAUNCH((ls_bwd_gather_tma_kernel<T>), dim3(dm.fw, dm.B * dm.N), dim3(LS_GATHER_THREADS), (size_t)kSmem, s, tm,
            (unsigned)rps, rows.nrows, (unsigned)total, (const T*)featT, pix_recs, dm, gprob_pm, (T*)gfeatT);
  *launched = true;
  return LS_OK;
}

// arrivals that complete an image in ready[] (ls_bwd_gather_occ_kernel)
int ls_gather_ready_target(const LsDims& dm) { return dm.fw * (LS_GATHER_THREADS / 32); }
// Opt-in (LS_OVERLAP_BWD=1): measured on a B200 at the default sizes, the overlap hides ~10 us of the
// epilogue but the gather's tail slows down by the same amount (its last wave is not idle capacity) -
// 239.4 vs 239.9 us per step; off by default, kept with a test.
bool ls_gather_can_overlap(const LsDims& dm) {
  const int nch = (dm.Cp + LS_CCHUNK - 1) / LS_CCHUNK;
  static const bool on = getenv("LS_OVERLAP_BWD") != nullptr;
  return on && LS_GATHER_OCC && dm.D % 16 == 0 && ls_epilogue_supports(dm) && nch == 1 && ls_pdl_enabled();
}

// How the backward reads a gradient with these strides (include/ls_b200.h LsBevStrides).
int ls_classify_grad_in(const float* p, const LsBevStrides& st, const LsDims& dm, const LsGrid& g) {
  if (dm.bev_bf16) {      // bf16 gradient rows: in place only, through the register-lean gather
    const bool ok = st.c == 1 && dm.C == 64 && dm.D % 16 == 0 && st.x == (long long)g.Y * st.y && st.y >= 64 &&
                    st.y % 4 == 0 && st.b % 4 == 0 && (uintptr_t)p % 8 == 0 &&
                    (unsigned long long)g.XY * (unsigned long long)st.y * 2ULL < (1ULL << 32);
    return ok ? LS_GRAD_DIRECT_VEC : LS_GRAD_BAD;
  }
  if (st.y == 1 && (st.c != 1 || dm.C == 1)) return LS_GRAD_STAGED;
  if (st.c == 1 && dm.C == dm.Cp && st.x == (long long)g.Y * st.y && st.y >= dm.C &&
      (unsigned long long)g.XY * (unsigned long long)st.y * 4ULL < (1ULL << 32)) {
    const bool vec = (uintptr_t)p % 16 == 0 && st.y % 4 == 0 && st.b % 4 == 0;
    return vec ? LS_GRAD_DIRECT_VEC : LS_GRAD_DIRECT_SCALAR;
  }
  return LS_GRAD_BAD;
}

template <typename T, int MODE>
static int ls_gather_dispatch(const LsRows& rows, const void* featT, const int2* pix_recs, const LsDims& dm,
                              float* gprob_pm, void* gfeatT, int* ready, cudaStream_t s) {
  const int nch = (dm.Cp + LS_CCHUNK - 1) / LS_CCHUNK;
  dim3 grid(dm.fw, dm.B * dm.N);
  if (ready && !ls_gather_can_overlap(dm)) return LS_ERR_UNSUPPORTED;
  if (dm.bev_bf16) {
    if (MODE != LS_GRAD_DIRECT_VEC) return LS_ERR_UNSUPPORTED;
    LS_LAUNCH((ls_bwd_gather_occ_kernel<T, LS_GRAD_DIRECT_VEC, __nv_bfloat16>), grid, dim3(LS_GATHER_THREADS), 0, s, rows,
              (const T*)featT, pix_recs, dm, gprob_pm, (T*)gfeatT, ready);
    return LS_OK;
  }
  static const bool want_tma = getenv("LS_GATHER_TMA") && atoi(getenv("LS_GATHER_TMA")) != 0;
  if (want_tma && MODE == LS_GRAD_DIRECT_VEC && !ready && nch == 1) {      // rows through the TMA gather4 path
    bool launched = false;
    const int rc = ls_gather_tma_try<T>(rows, featT, pix_recs, dm, gprob_pm, gfeatT, s, &launched);
    if (rc != LS_OK || launched) return rc;
  }
  if (LS_GATHER_OCC && dm.D % 16 == 0 && nch == 1) {
    LS_LAUNCH((ls_bwd_gather_occ_kernel<T, MODE, float>), grid, dim3(LS_GATHER_THREADS), 0, s, rows, (const T*)featT,
              pix_recs, dm, gprob_pm, (T*)gfeatT, ready);
    return LS_OK;
  }
#define LS_GATHER(NCH)                                                                                        \
  LS_LAUNCH((ls_bwd_gather_kernel<T, NCH, MODE>), grid, dim3(LS_GATHER_THREADS), 0, s, rows, (const T*)featT, \
            pix_recs, dm, gprob_pm, (T*)gfeatT)
  switch (nch) {
    case 1: LS_GATHER(1); break;
    case 2: LS_GATHER(2); break;
    case 3: LS_GATHER(3); break;
    case 4: LS_GATHER(4); break;
    default: return LS_ERR_UNSUPPORTED;
  }
#undef LS_GATHER
  return LS_OK;
}

// rows: staged (gT, mode LS_GRAD_STAGED) or the channels-last gradient itself (direct modes)
int ls_launch_bwd_gather(const void* rows_base, long long sample_stride, long long row_stride, int mode,
                         const void* featT, int dtype, const int2* pix_recs, const LsDims& dm, const LsGrid& g,
                         float* gprob_pm, void* gfeatT, int* ready, cudaStream_t s) {
  LsRows rows;
  rows.base = rows_base;
  rows.sample_stride = sample_stride;
  rows.row_bytes = (unsigned)(row_stride * (dm.bev_bf16 && mode != LS_GRAD_STAGED ? 2 : 4));
  rows.nrows = (unsigned)g.XY;
#define LS_GD(TT, MODE) return ls_gather_dispatch<TT, MODE>(rows, featT, pix_recs, dm, gprob_pm, gfeatT, ready, s)
  if (dtype == LS_F32) {
    if (mode == LS_GRAD_STAGED) LS_GD(float, LS_GRAD_STAGED);
    if (mode == LS_GRAD_DIRECT_VEC) LS_GD(float, LS_GRAD_DIRECT_VEC);
    if (mode == LS_GRAD_DIRECT_SCALAR) LS_GD(float, LS_GRAD_DIRECT_SCALAR);
  } else {
    if (mode == LS_GRAD_STAGED) LS_GD(__nv_bfloat16, LS_GRAD_STAGED);
    if (mode == LS_GRAD_DIRECT_VEC) LS_GD(__nv_bfloat16, LS_GRAD_DIRECT_VEC);
    if (mode == LS_GRAD_DIRECT_SCALAR) LS_GD(__nv_bfloat16, LS_GRAD_DIRECT_SCALAR);
  }
#undef LS_GD
  return LS_ERR_UNSUPPORTED;
}
