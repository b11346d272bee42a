// C ABI of libls_b200.so (include/ls_b200.h): argument checking, workspace carving and the
// forward / backward pipelines.  No torch types, no exceptions, no allocation, no host sync.
#include <atomic>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ls_internal.h"

static std::atomic<long long> g_launches{0};
static thread_local char g_cuda_err[256] = "";

void ls_note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
int ls_note_cuda_error(cudaError_t e, const char* file, int line) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s at %s:%d", cudaGetErrorString(e), file, line);
  return LS_ERR_CUDA;
}

bool ls_pdl_enabled() {
  static const bool on = getenv("LS_NO_PDL") == nullptr;
  return on;
}

// ---- side streams ---------------------------------------------------------------------
// ls_forward / ls_backward fork independent stages (softmax, layout staging) onto two
// library-owned non-blocking streams and join them back into the caller's stream with events:
// still one asynchronous unit of work on `stream` for the caller (and capturable in a CUDA
// graph as a fork/join), but the small kernels overlap instead of queueing behind each other.
// The streams and events are per device and shared by every caller, so the whole enqueue
// sequence of a pipeline call runs under the device's mutex (a wait captures the state of
// its event at the time of the call, so later calls re-recording it do no harm), and LsFork's
// destructor joins whatever was forked on EVERY exit path - an error return never leaves a
// side stream running on buffers the caller is about to free, nor a capture unjoined.
#include <mutex>
struct LsAsync {
  cudaStream_t s1 = nullptr, s2 = nullptr;
  cudaEvent_t fork = nullptr, j1 = nullptr, j2 = nullptr;
  bool ready = false;
  std::mutex mu;     // held while a pipeline call enqueues
};
static LsAsync g_async[64];
static std::mutex g_async_mu;
static LsAsync* ls_async() {
  static const bool disabled = getenv("LS_NO_SIDE_STREAMS") != nullptr;
  if (disabled) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(g_async_mu);
  LsAsync& a = g_async[dev];
  if (!a.ready) {
    if (cudaStreamCreateWithFlags(&a.s1, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaStreamCreateWithFlags(&a.s2, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&a.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&a.j1, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&a.j2, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    a.ready = true;
  }
  return &a;
}

// Scope of one pipeline call's use of the side streams.
struct LsFork {
  LsAsync* a;
  cudaStream_t main;
  bool open1 = false, open2 = false;
  explicit LsFork(cudaStream_t stream) : a(ls_async()), main(stream) {
    if (a) a->mu.lock();
  }
  // side stream k (1 or 2) ordered after everything enqueued on the caller's stream so far;
  // the caller's stream itself when side streams are off or the fork fails
  cudaStream_t side(int k) {
    if (!a) return main;
    bool& open = (k == 1) ? open1 : open2;
    cudaStream_t st = (k == 1) ? a->s1 : a->s2;
    if (cudaEventRecord(a->fork, main) != cudaSuccess || cudaStreamWaitEvent(st, a->fork, 0) != cudaSuccess) return main;
    open = true;
    return st;
  }
  // the caller's stream waits for side stream k
  int join(int k) {
    if (!a) return LS_OK;
    bool& open = (k == 1) ? open1 : open2;
    if (!open) return LS_OK;
    open = false;
    cudaEvent_t ev = (k == 1) ? a->j1 : a->j2;
    LS_CUDA(cudaEventRecord(ev, (k == 1) ? a->s1 : a->s2));
    LS_CUDA(cudaStreamWaitEvent(main, ev, 0));
    return LS_OK;
  }
  ~LsFork() {
    if (!a) return;
    (void)join(1);
    (void)join(2);
    a->mu.unlock();
  }
};

// Product of positive 32-bit factors, saturating at `cap` (<= 2^31): the 64-bit product of five
// int32 sizes can wrap, and a wrapped product must not pass a limit test.
static long long ls_prod_capped(const int* f, int n, long long cap) {
  long long p = 1;
  for (int i = 0; i < n; ++i) {
    p *= f[i];                       // p < cap <= 2^31 and f[i] < 2^31: no overflow
    if (p >= cap) return cap;
  }
  return p;
}

static int ls_check_shape(const LsShape* s) {
  if (!s) return LS_ERR_BAD_ARG;
  if (s->B <= 0 || s->N <= 0 || s->D <= 0 || s->fh <= 0 || s->fw <= 0 || s->C <= 0) return LS_ERR_BAD_ARG;
  if (s->X <= 0 || s->Y <= 0 || s->Z <= 0) return LS_ERR_BAD_ARG;
  if (s->geom_policy != LS_GEOM_TORCH_CPU && s->geom_policy != LS_GEOM_TORCH_CUDA) return LS_ERR_BAD_ARG;
  if (s->tile_x < 0 || s->tile_x > LS_TILE || (s->tile_x & (s->tile_x - 1))) return LS_ERR_BAD_ARG;
  if (s->bev_dtype != LS_F32 && s->bev_dtype != LS_BF16) return LS_ERR_BAD_ARG;
  const int cells[3] = {s->X, s->Y, s->Z};
  const int points[5] = {s->B, s->N, s->D, s->fh, s->fw};
  if (ls_prod_capped(cells, 3, 1LL << 28) >= (1LL << 28)) return LS_ERR_UNSUPPORTED;
  if (ls_prod_capped(points, 5, 1LL << 31) >= (1LL << 31)) return LS_ERR_UNSUPPORTED;
  return LS_OK;
}
static int ls_check_splat_shape(const LsShape* s) {
  int rc = ls_check_shape(s);
  if (rc) return rc;
  if (s->Z != 1) return LS_ERR_UNSUPPORTED;   // reference: squeeze(0) needs Z == 1 (bev_model.py:104)
  if (s->C > 4 * LS_CCHUNK) return LS_ERR_UNSUPPORTED;
  // placement: at most 32 depth bins per thread of 16 depth groups, and its pixel-major staging
  // tile (32 pixels x (D|1) records) has to fit the default 48 KB of shared memory
  if (s->D > ls_max_depth_bins()) return LS_ERR_UNSUPPORTED;
  LsDims dm = ls_dims(s);
  if (ls_grid(s).tiles >= (1 << 21)) return LS_ERR_UNSUPPORTED;        // exact magic division tile / tiles_y
  // sort key: 8 bits cell-in-tile | (pixel << dbits | d) must fit 24 bits
  if (((long long)s->N * dm.HW) << dm.dbits > (1LL << 24)) return LS_ERR_UNSUPPORTED;
  if ((long long)s->N * dm.HW > (1LL << 20)) return LS_ERR_UNSUPPORTED;   // pixel id field of a sorted record
  if ((long long)s->N * dm.HW * dm.Cp >= (1LL << 31)) return LS_ERR_UNSUPPORTED;
  if (((long long)ls_grid(s).Vc + 1) * dm.Cp * 4 >= (1LL << 31)) return LS_ERR_UNSUPPORTED;   // 32-bit row byte offsets
  return LS_OK;
}
static bool ls_layout_ok(int layout, const LsShape* s) {
  return layout == LS_FEAT_NCHW || (layout == LS_FEAT_NHWC && s->C % 4 == 0);
}
static bool ls_dtype_ok(int dtype) { return dtype == LS_F32 || dtype == LS_BF16; }

// ---- workspace carving -------------------------------------------------------------
// scratch: transient buffers of one call.  saved: what ls_backward needs from ls_forward.
struct LsWs {
  // scratch (forward)
  int *cell, *within, *counts, *tile_order, *tile_tot;
  int2 *recs, *recs_sorted;
  // scratch (backward)
  float *gT, *gprob_pm;
  void* gfeatT;
  // saved
  void* featT;
  int* seg_start;
  int2* pix_recs;
  int* bwd_flags;      // per-image arrival / departure counters of the overlapped backward epilogue (zeroed by the forward, self-resetting)
  size_t scratch_bytes, saved_bytes;
};
static inline size_t ls_align(size_t v) { return (v + 255) & ~(size_t)255; }
// phase: 0 forward scratch, 1 backward scratch.  saved may be NULL (forward without backward
// state): seg_start then lives in the scratch blob.
static LsWs ls_carve(const LsShape* s, int dtype, int feat_layout, int phase, void* scratch, void* saved, bool with_saved) {
  LsDims dm = ls_dims(s);
  LsGrid g = ls_grid(s);
  const size_t es = dtype == LS_BF16 ? 2 : 4;
  const size_t pts = (size_t)dm.B * dm.Npts;
  const size_t feat = (size_t)dm.B * dm.N * dm.HW * dm.Cp * es;
  LsWs w;
  memset(&w, 0, sizeof(w));
  char* p = (char*)scratch;
  size_t off = 0;
  auto take = [&](size_t n) { void* r = p ? (void*)(p + off) : nullptr; off += ls_align(n); return r; };
  if (phase == 0) {
    w.tile_order = (int*)take((size_t)dm.B * g.tiles * 4);
    w.tile_tot = (int*)take((size_t)dm.B * g.tiles * 4);
    w.counts = (int*)take((size_t)dm.B * g.Vc * 4);
    w.cell = (int*)take(pts * 4);
    w.within = (int*)take(pts * 4);
    w.recs = (int2*)take(pts * 8);
    w.recs_sorted = (int2*)take((size_t)dm.B * ls_sorted_records_capacity(dm, g) * 8);
    if (!with_saved) {
      w.seg_start = (int*)take((size_t)dm.B * g.seg_stride * 4);
      if (feat_layout == LS_FEAT_NCHW) w.featT = take(feat);
    }
  } else {
    w.gprob_pm = (float*)take(pts * 4);
    if (feat_layout == LS_FEAT_NCHW) w.gfeatT = take(feat);
    w.gT = (float*)take((size_t)dm.B * (g.XY + 1) * dm.Cp * 4);     // only read for NCHW gradients
  }
  w.scratch_bytes = off;
  p = (char*)saved;
  off = 0;
  if (with_saved) {
    w.seg_start = (int*)take((size_t)dm.B * g.seg_stride * 4);
    w.pix_recs = (int2*)take(pts * 8);
    if (feat_layout == LS_FEAT_NCHW) w.featT = take(feat);
    w.bwd_flags = (int*)take((size_t)ls_bwd_flag_ints(dm) * 4);
  }
  w.saved_bytes = off;
  return w;
}

// ---- static-rig cache (opt-in) ------------------------------------------------------------
struct LsCache {
  int *seg_start, *tile_order, *perm, *bwd_flags;
  int2 *recs_sorted, *pix_recs;
  size_t bytes;
};
static LsCache ls_carve_cache(const LsShape* s, void* base) {
  LsDims dm = ls_dims(s);
  LsGrid g = ls_grid(s);
  const size_t pts = (size_t)dm.B * dm.Npts;
  char* p = (char*)base;
  size_t off = 0;
  auto take = [&](size_t n) { void* r = p ? (void*)(p + off) : nullptr; off += ls_align(n); return r; };
  LsCache c;
  c.seg_start = (int*)take((size_t)dm.B * g.seg_stride * 4);
  c.tile_order = (int*)take((size_t)dm.B * g.tiles * 4);
  c.perm = (int*)take(pts * 4);
  c.recs_sorted = (int2*)take((size_t)dm.B * ls_sorted_records_capacity(dm, g) * 8);
  c.pix_recs = (int2*)take(pts * 8);
  c.bwd_flags = (int*)take((size_t)ls_bwd_flag_ints(dm) * 4);
  c.bytes = off;
  return c;
}

extern "C" {

const char* ls_version(void) { return "ls_b200 0.2 (sm_100a)"; }

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
int ls_debug_phase_cycles(uint64_t* out8) { return out8 ? ls_debug_fetch_phase_cycles((unsigned long long*)out8) : LS_ERR_BAD_ARG; }

int ls_grid_cells(const LsShape* s, int32_t* tiles, int32_t* cells_padded, int32_t* seg_stride) {
  int rc = ls_check_shape(s);
  if (rc) return rc;
  LsGrid g = ls_grid(s);
  if (tiles) *tiles = g.tiles;
  if (cells_padded) *cells_padded = g.Vc;
  if (seg_stride) *seg_stride = g.seg_stride;
  return LS_OK;
}

int32_t ls_padded_channels(int32_t C) { return (C + 3) & ~3; }

size_t ls_sorted_records(const LsShape* s) {
  if (ls_check_splat_shape(s)) return 0;
  return ls_sorted_records_capacity(ls_dims(s), ls_grid(s));
}

int ls_camera_transform(const float* intrinsics, const float* extrinsics, int32_t BN, float* M, float* t,
                        ls_stream_t stream) {
  if (!intrinsics || !extrinsics || !M || !t || BN <= 0) return LS_ERR_BAD_ARG;
  return ls_launch_camera_transform(intrinsics, extrinsics, BN, M, t, (cudaStream_t)stream);
}

int ls_geometry(const float* M, const float* t, const float* frustum, const LsShape* s, float* geom,
                ls_stream_t stream) {
  int rc = ls_check_shape(s);
  if (rc) return rc;
  if (!M || !t || !frustum || !geom) return LS_ERR_BAD_ARG;
  return ls_launch_export(M, t, frustum, ls_dims(s), ls_grid(s), geom, nullptr, nullptr, nullptr,
                          (cudaStream_t)stream);
}

int ls_export_indices(const float* M, const float* t, const float* frustum, const LsShape* s, int64_t* vox,
                      uint8_t* keep, int64_t* rank, ls_stream_t stream) {
  int rc = ls_check_shape(s);
  if (rc) return rc;
  if (!M || !t || !frustum) return LS_ERR_BAD_ARG;
  return ls_launch_export(M, t, frustum, ls_dims(s), ls_grid(s), nullptr, (long long*)vox, keep, (long long*)rank,
                          (cudaStream_t)stream);
}

int ls_index(const float* M, const float* t, const float* frustum, const LsShape* s, int32_t* rank, int32_t* cell,
             int32_t* within, int32_t* counts, ls_stream_t stream) {
  int rc = ls_check_shape(s);
  if (rc) return rc;
  if (!M || !t || !frustum) return LS_ERR_BAD_ARG;
  const bool sorting = cell || within || counts;
  if (sorting && !(cell && within && counts)) return LS_ERR_BAD_ARG;
  if (!sorting && !rank) return LS_ERR_BAD_ARG;
  if (sorting && (rc = ls_check_splat_shape(s))) return rc;
  return ls_launch_index(M, t, frustum, ls_dims(s), ls_grid(s), rank, cell, within, counts, (cudaStream_t)stream);
}

int ls_index_geom(const float* geom, const LsShape* s, int32_t* rank, int32_t* cell, int32_t* within,
                  int32_t* counts, ls_stream_t stream) {
  int rc = ls_check_shape(s);
  if (rc) return rc;
  if (!geom) return LS_ERR_BAD_ARG;
  const bool sorting = cell || within || counts;
  if (sorting && !(cell && within && counts)) return LS_ERR_BAD_ARG;
  if (!sorting && !rank) return LS_ERR_BAD_ARG;
  if (sorting && (rc = ls_check_splat_shape(s))) return rc;
  return ls_launch_index_geom(geom, ls_dims(s), ls_grid(s), rank, cell, within, counts, (cudaStream_t)stream);
}

int ls_sort(const int32_t* cell, const int32_t* within, const int32_t* counts, const void* prob, int dtype,
            const LsShape* s, int32_t* seg_start, int32_t* tile_order, int32_t* tile_scratch, void* recs,
            void* pix_recs, ls_stream_t stream) {
  int rc = ls_check_splat_shape(s);
  if (rc) return rc;
  if (!cell || !within || !counts || !prob || !seg_start || !tile_order || !recs || !ls_dtype_ok(dtype))
    return LS_ERR_BAD_ARG;
  LsDims dm = ls_dims(s);
  LsGrid g = ls_grid(s);
  if ((rc = ls_launch_scan(counts, dm, g, seg_start, tile_order, tile_scratch, (cudaStream_t)stream))) return rc;
  return ls_launch_place(cell, within, prob, dtype, dm, g, seg_start, (int2*)recs, (int2*)pix_recs,
                         (cudaStream_t)stream);
}

int ls_export_cell_counts(const int32_t* seg_start, const LsShape* s, int32_t b, int64_t* out, int32_t* kept,
                          ls_stream_t stream) {
  int rc = ls_check_splat_shape(s);
  if (rc) return rc;
  if (!seg_start || b < 0 || b >= s->B) return LS_ERR_BAD_ARG;
  return ls_launch_export_cell_counts(seg_start, ls_grid(s), s->B, b, (long long*)out, kept, (cudaStream_t)stream);
}

int ls_softmax(const void* logits, int dtype, const LsShape* s, void* prob, ls_stream_t stream) {
  int rc = ls_check_shape(s);
  if (rc) return rc;
  if (!logits || !prob || !ls_dtype_ok(dtype)) return LS_ERR_BAD_ARG;
  return ls_launch_softmax(logits, dtype, ls_dims(s), prob, (cudaStream_t)stream);
}

int ls_softmax_bwd(const void* prob, const float* grad_prob_pm, const void* grad_prob_ext, int dtype,
                   const LsShape* s, void* grad_logits, ls_stream_t stream) {
  int rc = ls_check_shape(s);
  if (rc) return rc;
  if (!prob || !grad_prob_pm || !grad_logits || !ls_dtype_ok(dtype)) return LS_ERR_BAD_ARG;
  return ls_launch_softmax_bwd(prob, grad_prob_pm, grad_prob_ext, dtype, ls_dims(s), grad_logits,
                               (cudaStream_t)stream);
}

int ls_nchw_to_nhwc(const void* src, int dtype, int32_t images, int32_t C, int32_t HW, void* dst,
                    ls_stream_t stream) {
  if (!src || !dst || images <= 0 || C <= 0 || HW <= 0 || !ls_dtype_ok(dtype)) return LS_ERR_BAD_ARG;
  return ls_launch_to_nhwc(src, dtype, images, C, ls_padded_channels(C), HW, dst, (cudaStream_t)stream);
}
int ls_nhwc_to_nchw(const void* src, int dtype, int32_t images, int32_t C, int32_t HW, void* dst,
                    ls_stream_t stream) {
  if (!src || !dst || images <= 0 || C <= 0 || HW <= 0 || !ls_dtype_ok(dtype)) return LS_ERR_BAD_ARG;
  return ls_launch_from_nhwc(src, dtype, images, C, ls_padded_channels(C), HW, dst, (cudaStream_t)stream);
}

int ls_splat_fwd(const void* feat_nhwc, int dtype, const void* recs, const int32_t* seg_start,
                 const int32_t* tile_order, void* recs_scratch, const LsShape* s, float* bev, const LsBevStrides* st,
                 ls_stream_t stream) {
  int rc = ls_check_splat_shape(s);
  if (rc) return rc;
  if (!feat_nhwc || !recs || !seg_start || !tile_order || !recs_scratch || !bev || !st || !ls_dtype_ok(dtype))
    return LS_ERR_BAD_ARG;
  return ls_launch_splat_fwd(feat_nhwc, dtype, (const int2*)recs, seg_start, tile_order, (int2*)recs_scratch, nullptr,
                             ls_dims(s), ls_grid(s), bev, *st, (cudaStream_t)stream);
}

int ls_splat_bwd(const float* grad_bev, const LsBevStrides* gst, const void* feat_nhwc, int dtype,
                 const void* pix_recs, const int32_t* seg_start, const LsShape* s, float* gT_ws, float* grad_prob_pm,
                 void* grad_feat_nhwc, ls_stream_t stream) {
  int rc = ls_check_splat_shape(s);
  if (rc) return rc;
  if (!grad_bev || !gst || !feat_nhwc || !pix_recs || !grad_prob_pm || !grad_feat_nhwc || !ls_dtype_ok(dtype))
    return LS_ERR_BAD_ARG;
  LsDims dm = ls_dims(s);
  LsGrid g = ls_grid(s);
  const int mode = ls_classify_grad_in(grad_bev, *gst, dm, g);
  if (mode == LS_GRAD_BAD) return LS_ERR_UNSUPPORTED;
  if (mode == LS_GRAD_STAGED) {
    if (!gT_ws || !seg_start) return LS_ERR_BAD_ARG;
    if ((rc = ls_launch_bwd_transpose(grad_bev, *gst, seg_start, dm, g, gT_ws, (cudaStream_t)stream))) return rc;
    return ls_launch_bwd_gather(gT_ws, (long long)(g.XY + 1) * dm.Cp, dm.Cp, mode, feat_nhwc, dtype,
                                (const int2*)pix_recs, dm, g, grad_prob_pm, grad_feat_nhwc, nullptr, (cudaStream_t)stream);
  }
  return ls_launch_bwd_gather(grad_bev, gst->b, gst->y, mode, feat_nhwc, dtype, (const int2*)pix_recs, dm, g,
                              grad_prob_pm, grad_feat_nhwc, nullptr, (cudaStream_t)stream);
}

int ls_target_bev(const int32_t* target_pix, int32_t B, int32_t X, int32_t Y, float* out, int64_t stride_b,
                  int64_t stride_x, int64_t stride_y, ls_stream_t stream) {
  if (!target_pix || !out || B <= 0 || X <= 0 || Y <= 0) return LS_ERR_BAD_ARG;
  if ((long long)X * Y >= (1LL << 31)) return LS_ERR_UNSUPPORTED;
  return ls_launch_target_bev(target_pix, B, X, Y, out, stride_b, stride_x, stride_y, (cudaStream_t)stream);
}

size_t ls_scratch_bytes(const LsShape* s, int dtype, int with_backward) {
  if (ls_check_splat_shape(s) || !ls_dtype_ok(dtype)) return 0;
  // sized for the layout that needs most (NCHW features and gradients), so one blob serves any call
  size_t n = ls_carve(s, dtype, LS_FEAT_NCHW, 0, nullptr, nullptr, false).scratch_bytes;
  if (with_backward) {
    const size_t b = ls_carve(s, dtype, LS_FEAT_NCHW, 1, nullptr, nullptr, true).scratch_bytes;
    n = b > n ? b : n;
  }
  return n;
}

size_t ls_saved_bytes(const LsShape* s, int dtype, int feat_layout) {
  if (ls_check_splat_shape(s) || !ls_dtype_ok(dtype) || !ls_layout_ok(feat_layout, s)) return 0;
  return ls_carve(s, dtype, feat_layout, 0, nullptr, nullptr, true).saved_bytes;
}

int ls_forward(const void* feat, int feat_layout, const void* logits, int dtype, const float* M, const float* t,
               const float* frustum, const LsShape* s, void* scratch, size_t scratch_bytes, void* saved,
               size_t saved_bytes, float* bev, const LsBevStrides* bev_strides, void* prob, ls_stream_t stream_) {
  int rc = ls_check_splat_shape(s);
  if (rc) return rc;
  if (!feat || !logits || !M || !t || !frustum || !scratch || !bev || !bev_strides || !prob || !ls_dtype_ok(dtype))
    return LS_ERR_BAD_ARG;
  if (!ls_layout_ok(feat_layout, s)) return LS_ERR_UNSUPPORTED;
  const bool with_saved = saved != nullptr;
  LsWs w = ls_carve(s, dtype, feat_layout, 0, scratch, saved, with_saved);
  if (scratch_bytes < w.scratch_bytes || saved_bytes < w.saved_bytes) return LS_ERR_WORKSPACE;
  cudaStream_t stream = (cudaStream_t)stream_;
  LsDims dm = ls_dims(s);
  LsGrid g = ls_grid(s);
  if (ls_classify_bev_out(bev, *bev_strides, dm, g) == LS_OUT_BAD) return LS_ERR_UNSUPPORTED;
  const void* featT = feat_layout == LS_FEAT_NHWC ? feat : w.featT;
  LsFork fk(stream);
  // side stream 1: depth softmax; side stream 2: NHWC staging of NCHW features
  if ((rc = ls_launch_softmax(logits, dtype, dm, prob, fk.side(1)))) return rc;
  if (feat_layout == LS_FEAT_NCHW &&
      (rc = ls_launch_to_nhwc(feat, dtype, dm.B * dm.N, dm.C, dm.Cp, dm.HW, w.featT, fk.side(2))))
    return rc;
  // caller's stream: index -> scan, then (after softmax) placement, then (after staging) the splat
  if ((rc = ls_launch_zero_counts(w.counts, dm, g, w.bwd_flags, ls_bwd_flag_ints(dm), stream))) return rc;
  if ((rc = ls_launch_index(M, t, frustum, dm, g, nullptr, w.cell, w.within, w.counts, stream))) return rc;
  if ((rc = ls_launch_scan(w.counts, dm, g, w.seg_start, w.tile_order, w.tile_tot, stream))) return rc;
  if ((rc = fk.join(1))) return rc;
  if ((rc = ls_launch_place(w.cell, w.within, prob, dtype, dm, g, w.seg_start, w.recs, w.pix_recs, stream))) return rc;
  if ((rc = fk.join(2))) return rc;
  return ls_launch_splat_fwd(featT, dtype, w.recs, w.seg_start, w.tile_order, w.recs_sorted, nullptr, dm, g, bev,
                             *bev_strides, stream);
}

int ls_backward(const float* grad_bev, const LsBevStrides* grad_strides, const void* grad_prob_ext, const void* prob,
                const void* feat, int feat_layout, int dtype, const LsShape* s, void* scratch, size_t scratch_bytes,
                const void* saved, size_t saved_bytes, void* grad_feat, void* grad_logits, ls_stream_t stream_) {
  int rc = ls_check_splat_shape(s);
  if (rc) return rc;
  if (!grad_bev || !grad_strides || !prob || !scratch || !saved || !grad_feat || !grad_logits || !ls_dtype_ok(dtype))
    return LS_ERR_BAD_ARG;
  if (!ls_layout_ok(feat_layout, s)) return LS_ERR_UNSUPPORTED;
  if (feat_layout == LS_FEAT_NHWC && !feat) return LS_ERR_BAD_ARG;
  LsWs wf = ls_carve(s, dtype, feat_layout, 0, nullptr, const_cast<void*>(saved), true);
  LsWs w = ls_carve(s, dtype, feat_layout, 1, scratch, nullptr, false);
  if (scratch_bytes < w.scratch_bytes || saved_bytes < wf.saved_bytes) return LS_ERR_WORKSPACE;
  cudaStream_t stream = (cudaStream_t)stream_;
  LsDims dm = ls_dims(s);
  LsGrid g = ls_grid(s);
  const int mode = ls_classify_grad_in(grad_bev, *grad_strides, dm, g);
  if (mode == LS_GRAD_BAD) return LS_ERR_UNSUPPORTED;
  const void* featT = feat_layout == LS_FEAT_NHWC ? feat : wf.featT;
  void* gfeatT = feat_layout == LS_FEAT_NHWC ? grad_feat : w.gfeatT;
  // common shapes: the epilogue (softmax backward + grad_feat layout) is ONE pixel-stationary launch; with
  // LS_OVERLAP_BWD=1 it overlaps the gather's tail, synchronised per image through the counters in the saved blob
  int* ready = ls_gather_can_overlap(dm) ? wf.bwd_flags : nullptr;
  if (mode == LS_GRAD_STAGED) {
    if ((rc = ls_launch_bwd_transpose(grad_bev, *grad_strides, wf.seg_start, dm, g, w.gT, stream))) return rc;
    rc = ls_launch_bwd_gather(w.gT, (long long)(g.XY + 1) * dm.Cp, dm.Cp, mode, featT, dtype, wf.pix_recs, dm, g,
                              w.gprob_pm, gfeatT, ready, stream);
  } else {
    rc = ls_launch_bwd_gather(grad_bev, grad_strides->b, grad_strides->y, mode, featT, dtype, wf.pix_recs, dm, g,
                              w.gprob_pm, gfeatT, ready, stream);
  }
  if (rc) return rc;
  if (ready || ls_epilogue_supports(dm))
    return ls_launch_bwd_epilogue(prob, w.gprob_pm, grad_prob_ext, feat_layout == LS_FEAT_NCHW ? w.gfeatT : nullptr, dtype,
                                  dm, grad_logits, grad_feat, ready, ls_gather_ready_target(dm), stream);
  // the two layout fix-ups are independent: grad_feat on a side stream, grad_logits on the caller's
  LsFork fk(stream);
  if (feat_layout == LS_FEAT_NCHW &&
      (rc = ls_launch_from_nhwc(w.gfeatT, dtype, dm.B * dm.N, dm.C, dm.Cp, dm.HW, grad_feat, fk.side(1))))
    return rc;
  if ((rc = ls_launch_softmax_bwd(prob, w.gprob_pm, grad_prob_ext, dtype, dm, grad_logits, stream))) return rc;
  return fk.join(1);
}

size_t ls_cache_bytes(const LsShape* s) {
  if (ls_check_splat_shape(s)) return 0;
  return ls_carve_cache(s, nullptr).bytes;
}

int ls_forward_cached(const void* feat, int feat_layout, const void* logits, int dtype, const float* M, const float* t,
                      const float* frustum, const LsShape* s, void* scratch, size_t scratch_bytes, void* saved,
                      size_t saved_bytes, void* cache, size_t cache_bytes, int rebuild, float* bev,
                      const LsBevStrides* bev_strides, void* prob, ls_stream_t stream_) {
  int rc = ls_check_splat_shape(s);
  if (rc) return rc;
  if (!feat || !logits || !frustum || !scratch || !cache || !bev || !bev_strides || !prob || !ls_dtype_ok(dtype))
    return LS_ERR_BAD_ARG;
  if (rebuild && (!M || !t)) return LS_ERR_BAD_ARG;
  if (!ls_layout_ok(feat_layout, s)) return LS_ERR_UNSUPPORTED;
  const bool with_saved = saved != nullptr;
  LsWs w = ls_carve(s, dtype, feat_layout, 0, scratch, saved, with_saved);
  LsCache c = ls_carve_cache(s, cache);
  if (scratch_bytes < w.scratch_bytes || saved_bytes < w.saved_bytes || cache_bytes < c.bytes) return LS_ERR_WORKSPACE;
  cudaStream_t stream = (cudaStream_t)stream_;
  LsDims dm = ls_dims(s);
  LsGrid g = ls_grid(s);
  if (ls_classify_bev_out(bev, *bev_strides, dm, g) == LS_OUT_BAD) return LS_ERR_UNSUPPORTED;
  const void* featT = feat_layout == LS_FEAT_NHWC ? feat : w.featT;
  LsFork fk(stream);
  if ((rc = ls_launch_softmax(logits, dtype, dm, prob, fk.side(1)))) return rc;
  if (feat_layout == LS_FEAT_NCHW &&
      (rc = ls_launch_to_nhwc(feat, dtype, dm.B * dm.N, dm.C, dm.Cp, dm.HW, w.featT, fk.side(2))))
    return rc;
  if (rebuild) {
    // the full index / sort / canonical-order pipeline, written into the cache instead of the scratch
    if ((rc = ls_launch_zero_counts(w.counts, dm, g, c.bwd_flags, ls_bwd_flag_ints(dm), stream))) return rc;
    if ((rc = ls_launch_index(M, t, frustum, dm, g, nullptr, w.cell, w.within, w.counts, stream))) return rc;
    if ((rc = ls_launch_scan(w.counts, dm, g, c.seg_start, c.tile_order, w.tile_tot, stream))) return rc;
    if ((rc = fk.join(1))) return rc;
    if ((rc = ls_launch_place(w.cell, w.within, prob, dtype, dm, g, c.seg_start, w.recs, c.pix_recs, stream))) return rc;
    if ((rc = fk.join(2))) return rc;
    return ls_launch_splat_fwd(featT, dtype, w.recs, c.seg_start, c.tile_order, c.recs_sorted, c.perm, dm, g, bev,
                               *bev_strides, stream);
  }
  // cached: only the weights the records carry are new
  if ((rc = fk.join(1))) return rc;
  if ((rc = ls_launch_refresh(prob, dtype, c.perm, c.seg_start, dm, g, c.recs_sorted, c.pix_recs, c.bwd_flags,
                              ls_bwd_flag_ints(dm), stream)))
    return rc;
  if ((rc = fk.join(2))) return rc;
  return ls_launch_splat_fwd(featT, dtype, nullptr, c.seg_start, c.tile_order, c.recs_sorted, nullptr, dm, g, bev,
                             *bev_strides, stream);
}

int ls_backward_cached(const float* grad_bev, const LsBevStrides* grad_strides, const void* grad_prob_ext,
                       const void* prob, const void* feat, int feat_layout, int dtype, const LsShape* s, void* scratch,
                       size_t scratch_bytes, const void* saved, size_t saved_bytes, const void* cache,
                       size_t cache_bytes, void* grad_feat, void* grad_logits, ls_stream_t stream_) {
  int rc = ls_check_splat_shape(s);
  if (rc) return rc;
  if (!grad_bev || !grad_strides || !prob || !scratch || !cache || !grad_feat || !grad_logits || !ls_dtype_ok(dtype))
    return LS_ERR_BAD_ARG;
  if (!ls_layout_ok(feat_layout, s)) return LS_ERR_UNSUPPORTED;
  if (feat_layout == LS_FEAT_NHWC ? !feat : !saved) return LS_ERR_BAD_ARG;
  LsWs wf = ls_carve(s, dtype, feat_layout, 0, nullptr, const_cast<void*>(saved), true);
  LsWs w = ls_carve(s, dtype, feat_layout, 1, scratch, nullptr, false);
  LsCache c = ls_carve_cache(s, const_cast<void*>(cache));
  if (scratch_bytes < w.scratch_bytes || cache_bytes < c.bytes) return LS_ERR_WORKSPACE;
  if (feat_layout == LS_FEAT_NCHW && saved_bytes < wf.saved_bytes) return LS_ERR_WORKSPACE;
  cudaStream_t stream = (cudaStream_t)stream_;
  LsDims dm = ls_dims(s);
  LsGrid g = ls_grid(s);
  const int mode = ls_classify_grad_in(grad_bev, *grad_strides, dm, g);
  if (mode == LS_GRAD_BAD) return LS_ERR_UNSUPPORTED;
  const void* featT = feat_layout == LS_FEAT_NHWC ? feat : wf.featT;
  void* gfeatT = feat_layout == LS_FEAT_NHWC ? grad_feat : w.gfeatT;
  int* ready = ls_gather_can_overlap(dm) ? c.bwd_flags : nullptr;
  if (mode == LS_GRAD_STAGED) {
    if ((rc = ls_launch_bwd_transpose(grad_bev, *grad_strides, c.seg_start, dm, g, w.gT, stream))) return rc;
    rc = ls_launch_bwd_gather(w.gT, (long long)(g.XY + 1) * dm.Cp, dm.Cp, mode, featT, dtype, c.pix_recs, dm, g,
                              w.gprob_pm, gfeatT, ready, stream);
  } else {
    rc = ls_launch_bwd_gather(grad_bev, grad_strides->b, grad_strides->y, mode, featT, dtype, c.pix_recs, dm, g,
                              w.gprob_pm, gfeatT, ready, stream);
  }
  if (rc) return rc;
  if (ready || ls_epilogue_supports(dm))
    return ls_launch_bwd_epilogue(prob, w.gprob_pm, grad_prob_ext, feat_layout == LS_FEAT_NCHW ? w.gfeatT : nullptr, dtype,
                                  dm, grad_logits, grad_feat, ready, ls_gather_ready_target(dm), stream);
  LsFork fk(stream);
  if (feat_layout == LS_FEAT_NCHW &&
      (rc = ls_launch_from_nhwc(w.gfeatT, dtype, dm.B * dm.N, dm.C, dm.Cp, dm.HW, grad_feat, fk.side(1))))
    return rc;
  if ((rc = ls_launch_softmax_bwd(prob, w.gprob_pm, grad_prob_ext, dtype, dm, grad_logits, stream))) return rc;
  return fk.join(1);
}

}  // extern "C"
