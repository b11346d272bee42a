/*
 * ls_b200.h - C ABI of the B200 (sm_100a) lift-splat library  (libls_b200.so)
 *
 * Drop-in boundary for the camera->BEV projection hot path of E2E Parking
 * (reference: model/bev_model.py, tool/geometry.py:40-59,285-317).  The reference
 * is pure Python/PyTorch and has no FFI of its own; these entry points are what a
 * ctypes (or TORCH_LIBRARY) binding inside the reference's BevModel would bind -
 * INTEGRATION.md shows that stub.  Plain pointers and sizes only: no torch types,
 * no exceptions across the boundary; every function returns an LsStatus.
 *
 * All pointers are DEVICE pointers unless the name ends in _host.  Every call is
 * asynchronous on `stream` (pass the caller's current stream), performs no host
 * synchronisation and no allocation, so a whole forward/backward is CUDA-graph
 * capturable.
 *
 * Point order inside a sample is the reference's: p = ((cam*D + d)*fh + row)*fw + col
 * (model/bev_model.py:83).  rank = gx*(Y*Z) + gy*Z + gz (model/bev_model.py:94-95),
 * -1 for a dropped point.
 */
#ifndef LS_B200_H_
#define LS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* ls_stream_t; /* cudaStream_t */

typedef enum LsStatus {
  LS_OK = 0,
  LS_ERR_BAD_ARG = -1,      /* null pointer, non-positive size, odd channel count ... */
  LS_ERR_UNSUPPORTED = -2,  /* e.g. Z != 1 for the splat (reference squeezes Z, bev_model.py:104) */
  LS_ERR_WORKSPACE = -3,    /* workspace too small */
  LS_ERR_CUDA = -4          /* a CUDA runtime call / launch failed; see ls_last_cuda_error() */
} LsStatus;

typedef enum LsDtype { LS_F32 = 0, LS_BF16 = 1 } LsDtype;

/* Sizes of one lift-splat problem.  start/res/dim are BevModel.bev_start_pos /
 * bev_res / bev_dim (tool/geometry.py:40-59, model/bev_model.py:15-20). */
typedef struct LsShape {
  int32_t B, N;        /* batch, cameras                                       */
  int32_t D, fh, fw;   /* depth bins, feature rows, feature cols (frustum dims) */
  int32_t C;           /* feature channels (even)                               */
  int32_t X, Y, Z;     /* bev_dim                                               */
  float start[3];      /* bev_start_pos                                         */
  float res[3];        /* bev_res                                               */
  int32_t geom_policy; /* LsGeomPolicy: float32 evaluation order of M.(u*d, v*d, d) + t    */
  int32_t tile_x;      /* internal BEV tiling: tiles of 128 cells, tile_x x (128 / tile_x); 0 or 1 =
                        * 1 x 128 strips (the default; required for NCHW BEV tensors), 8 = 8 x 16
                        * (accepted by the channels-last 64-channel splat; same bits, measured slower
                        * on a B200); a power of two <= 128.  Use the same value for every call of
                        * one forward/backward pair.                                          */
  int32_t bev_dtype;   /* LsDtype of the BEV tensor and of the gradient arriving on it.  LS_F32 (0) is the
                        * reference's contract (model/bev_model.py:76: always float32).  LS_BF16 is an opt-in
                        * for autocast training (the consumer convolution runs in bf16 anyway): 128-byte
                        * rows, half the output and gradient-row traffic; needs a dense channels-last
                        * tensor with C == 64 (and D % 16 == 0 for the backward); the `float*` BEV
                        * arguments then point at bf16 data.  Sums are still taken in float32.      */
} LsShape;

/* model/bev_model.py:54 is a broadcast batched 3x3 matmul; its float32 rounding depends on the
 * device the reference runs on, and 1 ulp decides voxel indices of points on cell boundaries:
 *   LS_GEOM_TORCH_CPU   ((m0*px + m1*py) + m2*pz) + t, every product and sum rounded (aten's CPU
 *                       kernel; what the golden fixtures pin)
 *   LS_GEOM_TORCH_CUDA  (fma(m1, py, m0*px) + m2*pz) + t: what torch's CUDA matmul (cuBLAS batched
 *                       SGEMM, m=3 n=1 k=3) computes on a B200 - probed bit for bit with
 *                       tools/geom_policy_probe.py, pinned by tests/test_gpu_parity.py */
typedef enum LsGeomPolicy { LS_GEOM_TORCH_CPU = 0, LS_GEOM_TORCH_CUDA = 1 } LsGeomPolicy;

/* Element strides of a BEV tensor of logical shape [B, C, X, Y] (the reference's bev_feature,
 * model/bev_model.py:76,105, and the gradient arriving on it).  Two families are accepted:
 *   y == 1  "NCHW"        - the reference's own layout (and channel-slice views of it, e.g. the
 *                           gradient torch.cat's backward hands back, model/parking_model.py:45);
 *   c == 1  "channels-last" - torch.channels_last storage [B, X, Y, C] (and channel-slice views of
 *                           it: y >= C).  The splat accumulates cell rows, so this is the layout
 *                           it writes with one bulk (TMA) store per tile, and the layout whose
 *                           gradient the backward gathers from directly (no staging pass).
 * Anything else is LS_ERR_UNSUPPORTED (make the tensor contiguous in one of the two first). */
typedef struct LsBevStrides {
  int64_t b, c, x, y;
} LsBevStrides;

/* Memory layout of the per-camera feature maps handed to ls_forward / returned by ls_backward. */
typedef enum LsFeatLayout {
  LS_FEAT_NCHW = 0,   /* [B*N, C, fh, fw] contiguous, as CamEncoder returns them (model/cam_encoder.py:102-111) */
  LS_FEAT_NHWC = 1    /* torch.channels_last storage [B*N, fh, fw, C]; needs C % 4 == 0 (no staging copies) */
} LsFeatLayout;

const char* ls_version(void);
const char* ls_strerror(int status);
const char* ls_last_cuda_error(void);

/* Internal BEV tiling: the grid is cut into tiles of 128 cells (LsShape.tile_x rows each);
 * cells are numbered tile-major.  cells_padded = tiles * 128 (>= X*Y); seg_stride = row
 * stride of seg_start. */
int ls_grid_cells(const LsShape* s, int32_t* tiles, int32_t* cells_padded, int32_t* seg_stride);

/* Channel count of the internal NHWC staging rows: C rounded up to a multiple of 4. */
int32_t ls_padded_channels(int32_t C);

/* Per-sample capacity (in 8-byte records) of ls_splat_fwd's recs_scratch: the records of
 * every cell re-ordered by key, in the CSR slots of `recs`, plus a few records of read slack. */
size_t ls_sorted_records(const LsShape* s);

/* a4 first half - model/bev_model.py:46-47,53:  E^-1 = inverse(extrinsics),
 * M = E^-1[:3,:3] . inverse(intrinsics),  t = E^-1[:3,3].
 * intrinsics f32[BN,3,3], extrinsics f32[BN,4,4] -> M f32[BN,3,3], t f32[BN,3].
 * Inverse = float64 Gauss-Jordan with partial pivoting rounded once to float32
 * (oracle/lift_splat_oracle.py:camera_transform is the bit-exact statement). */
int ls_camera_transform(const float* intrinsics, const float* extrinsics, int32_t BN,
                        float* M, float* t, ls_stream_t stream);

/* a4 second half - model/bev_model.py:49-55 (BevModel.get_geometry):
 * geom f32[B,N,D,fh,fw,3] = M.(u*d, v*d, d) + t, unfused float32, bit-exact with
 * torch-CPU.  Only for API compatibility / tests: the hot path never stores geom. */
int ls_geometry(const float* M, const float* t, const float* frustum, const LsShape* s,
                float* geom, ls_stream_t stream);

/* a4+a6.1-a6.3 fused - model/bev_model.py:49-55,85-95: per-point voxel index in
 * registers, keep test, rank.
 *   rank   i32[B,Npts]  reference rank, -1 = dropped (may be NULL)
 *   cell   i32[B,Npts]  tile-major cell id, -1 = dropped       } all three or none;
 *   within i32[B,Npts]  ticket of the point inside its cell    } counts must be zero
 *   counts i32[B,cells_padded] per-cell histogram (int atomics)} on entry            */
int ls_index(const float* M, const float* t, const float* frustum, const LsShape* s,
             int32_t* rank, int32_t* cell, int32_t* within, int32_t* counts, ls_stream_t stream);

/* a6.1-a6.3 from GIVEN coordinates - the index half of proj_bev_feature(geom, x)
 * (model/bev_model.py:74-107, the reference's public method that takes a materialised geom):
 * geom f32[B,Npts,3] in the reference's point order; outputs as ls_index. */
int ls_index_geom(const float* geom, const LsShape* s, int32_t* rank, int32_t* cell, int32_t* within,
                  int32_t* counts, ls_stream_t stream);

/* Debug/test export in the reference's own convention (model/bev_model.py:85-97):
 * vox i64[B,Npts,3] = .long() of the voxel coordinate (INT64_MIN where x86 gives
 * 'integer indefinite'), keep u8[B,Npts], rank i64[B,Npts] (-1 dropped). Any may be NULL. */
int ls_export_indices(const float* M, const float* t, const float* frustum, const LsShape* s,
                      int64_t* vox, uint8_t* keep, int64_t* rank, ls_stream_t stream);

/* a6.3 - model/bev_model.py:96-97 (argsort + gathers) and the segment boundaries of
 * tool/geometry.py:295-296, as a counting sort by cell: exclusive scan of counts ->
 * seg_start i32[B,seg_stride] (CSR offsets, entry [cells_padded] = kept count) and
 * tile_order i32[B,tiles] (tile ids by descending point count: the launch order of the
 * splat's CTAs), then every
 * kept point writes an 8-byte record {key, prob bits} to recs[B,Npts] at
 * seg_start[cell] + within;  key = cell_in_tile << 24 | (pixel << ceil(log2 D) | d).
* pix_recs (may be NULL): i32x2[B*N*fh*fw, D] = {rank gx*Y+gy of the point's cell (X*Y for a dropped
 * point), prob bits} per depth bin of every pixel, pixel-major - the index ls_splat_bwd walks.
 * prob: [B*N,D,fh,fw] of `dtype`. */
int ls_sort(const int32_t* cell, const int32_t* within, const int32_t* counts, const void* prob,
            int dtype, const LsShape* s, int32_t* seg_start, int32_t* tile_order,
            int32_t* tile_scratch /* i32[B,tiles] or NULL: enables the parallel scan */, void* recs,
            void* pix_recs, ls_stream_t stream);

/* Test export: for sample b, out i64[cells_padded,2] = (row-major rank, number of kept
 * points) of every cell, zeros for empty / padding cells; from it the reference's
 * ``ranks[ranks.argsort()]`` (model/bev_model.py:97) is rebuilt by repeating each rank
 * count times in ascending rank order.  kept i32[B] receives every sample's kept count.
 * Either output may be NULL. */
int ls_export_cell_counts(const int32_t* seg_start, const LsShape* s, int32_t b, int64_t* out,
                          int32_t* kept, ls_stream_t stream);

/* a5 - model/bev_model.py:64: prob = softmax(depth_logits, dim=1).
 * logits/prob: [B*N, D, fh, fw] of `dtype`. */
int ls_softmax(const void* logits, int dtype, const LsShape* s, void* prob, ls_stream_t stream);

/* Layout staging: [images, C, HW] -> [images, HW, Cp] (Cp = ls_padded_channels(C), zero
 * padded) and back (padding dropped); coalesced shared-memory transposes. */
int ls_nchw_to_nhwc(const void* src, int dtype, int32_t images, int32_t C, int32_t HW,
                    void* dst, ls_stream_t stream);
int ls_nhwc_to_nchw(const void* src, int dtype, int32_t images, int32_t C, int32_t HW,
                    void* dst, ls_stream_t stream);

/* a5(outer product)+a6.4-a6.5 fused - model/bev_model.py:66-72,99-105 and
 * VoxelsSumming.forward (tool/geometry.py:289-305): deterministic segment sum of
 * prob[p]*feat[pix(p),:] per cell, written (zeros included) to bev f32[B,C,X,Y].
 * feat_nhwc: [B*N, fh, fw, Cp] of `dtype`; recs/seg_start/tile_order from ls_sort;
 * recs_scratch: 8 B x [B, ls_sorted_records()], receives the records re-ordered by key inside each cell
 * (this is what makes the sums independent of the atomics' arrival order). */
int ls_splat_fwd(const void* feat_nhwc, int dtype, const void* recs, const int32_t* seg_start,
                 const int32_t* tile_order, void* recs_scratch, const LsShape* s, float* bev,
                 const LsBevStrides* bev_strides, ls_stream_t stream);

/* a7 + autograd of a5/a6 - VoxelsSumming.backward (tool/geometry.py:307-317) and the
 * backward of the outer product: every kept point receives its cell's gradient;
 *   grad_prob[p]      = sum_c feat[pix,c] * g[c, cell(p)]
 *   grad_feat[pix, c] = sum_d prob[d,pix] * g[c, cell(d,pix)]      (pixel-stationary,
 * no atomics, fixed order).  grad_bev f32 [B,C,X,Y] with strides: a channels-last gradient
 * (c stride 1, x stride == Y * y stride, C % 4 == 0) is gathered in place; an NCHW one (y stride 1)
 * is first staged as cell rows into gT_ws f32[B, X*Y+1, Cp] (may be NULL for channels-last).
 * Outputs grad_prob_pm f32[B*N*fh*fw, D] (PIXEL-major, consumed by ls_softmax_bwd) and
 * grad_feat_nhwc [B*N,fh,fw,Cp] of `dtype`. */
int ls_splat_bwd(const float* grad_bev, const LsBevStrides* grad_strides, const void* feat_nhwc,
                 int dtype, const void* pix_recs, const int32_t* seg_start, const LsShape* s,
                 float* gT_ws, float* grad_prob_pm, void* grad_feat_nhwc, ls_stream_t stream);

/* backward of a5's softmax (model/bev_model.py:64): grad_logits = prob * (g - sum_d prob*g),
 * g = grad_prob_pm (pixel-major, from ls_splat_bwd) + grad_prob_ext (the gradient arriving
 * on the returned pred_depth, [B*N,D,fh,fw] of `dtype`, may be NULL). */
int ls_softmax_bwd(const void* prob, const float* grad_prob_pm, const void* grad_prob_ext, int dtype,
                   const LsShape* s, void* grad_logits, ls_stream_t stream);

/* ---- one-call pipelines (what BevModel.calc_bev_feature uses) -------------------
 * Two device blobs:
 *   scratch  transient, >= ls_scratch_bytes(); free for other use as soon as the call's kernels
 *            have run (the Python side keeps one per stream and reuses it for every call);
 *   saved    >= ls_saved_bytes(); what ls_backward needs from ls_forward (NHWC feature copy for
 *            NCHW inputs, CSR offsets, pixel-major index): keep it untouched until ls_backward
 *            of the same step.  Pass NULL/0 to ls_forward for inference (no backward state). */
size_t ls_scratch_bytes(const LsShape* s, int dtype, int with_backward);
size_t ls_saved_bytes(const LsShape* s, int dtype, int feat_layout);

/* feat [B*N,C,fh,fw] in `feat_layout` (LsFeatLayout), logits [B*N,D,fh,fw] contiguous, both of
 * `dtype`, as CamEncoder returns them (model/cam_encoder.py:102-111); M/t from
 * ls_camera_transform or from the caller's own torch.inverse; frustum f32[D,fh,fw,3]
 * (BevModel.frustum).  Outputs: bev f32[B,C,X,Y] (strided, see LsBevStrides), prob
 * [B*N,D,fh,fw] of `dtype` (= pred_depth). */
int ls_forward(const void* feat, int feat_layout, const void* logits, int dtype, const float* M, const float* t,
               const float* frustum, const LsShape* s, void* scratch, size_t scratch_bytes, void* saved,
               size_t saved_bytes, float* bev, const LsBevStrides* bev_strides, void* prob, ls_stream_t stream);

/* grad_bev f32 (strided), grad_prob_ext (`dtype`, may be NULL), prob = forward's output, feat =
 * forward's input (read only when feat_layout == LS_FEAT_NHWC: no copy of it was saved).
 * Outputs grad_feat [B*N,C,fh,fw] in `feat_layout`, grad_logits [B*N,D,fh,fw] of `dtype`.
 * = [ls_bwd_transpose for NCHW gradients ->] the gradient gather of ls_splat_bwd -> ONE epilogue kernel
 * (the softmax backward of ls_softmax_bwd + the NHWC -> NCHW fix-up of grad_feat, a thread per pixel;
 * D = 32, 48, 64 - other depth counts run ls_softmax_bwd and ls_nhwc_to_nchw side by side); the same
 * bits either way.  `saved` also carries a few per-image counters (zeroed by ls_forward, left zero by
 * ls_backward) that the opt-in overlapped epilogue (environment LS_OVERLAP_BWD=1) synchronises on. */
int ls_backward(const float* grad_bev, const LsBevStrides* grad_strides, const void* grad_prob_ext, const void* prob,
                const void* feat, int feat_layout, int dtype, const LsShape* s, void* scratch, size_t scratch_bytes,
                const void* saved, size_t saved_bytes, void* grad_feat, void* grad_logits, ls_stream_t stream);

/* ---- static-rig cache (opt-in) ---------------------------------------------------------------
 * The reference's camera rig never moves (intrinsics / extrinsics are constants of the dataset,
 * dataset/carla_dataset.py:392-393), so voxel indices, the counting sort and the canonical record order
 * of a batch are identical every step; only the depth probabilities the records carry change.
 * `cache` (>= ls_cache_bytes(), device, owned by the caller, kept across steps) holds the CSR offsets,
 * the tile order, the canonical records, their slot -> point permutation and the pixel-major index.
 *   rebuild != 0  same work as ls_forward, results kept in the cache (first step, or the rig changed);
 *   rebuild == 0  softmax + a streaming refresh of the record weights + the splat: no index kernel, no
 *                 histogram, no scan, no placement, no re-ordering.  M / t are not read.
 * The caller vouches that M, t, frustum and the shape are those of the rebuild step.  When a cache is
 * in use the backward state in `saved` only carries the NHWC feature copy (NULL for channels-last
 * features).  Results are bit-identical to ls_forward / ls_backward. */
size_t ls_cache_bytes(const LsShape* s);
int ls_forward_cached(const void* feat, int feat_layout, const void* logits, int dtype, const float* M,
                      const float* t, const float* frustum, const LsShape* s, void* scratch,
                      size_t scratch_bytes, void* saved, size_t saved_bytes, void* cache, size_t cache_bytes,
                      int rebuild, float* bev, const LsBevStrides* bev_strides, void* prob,
                      ls_stream_t stream);
int ls_backward_cached(const float* grad_bev, const LsBevStrides* grad_strides, const void* grad_prob_ext,
                       const void* prob, const void* feat, int feat_layout, int dtype, const LsShape* s,
                       void* scratch, size_t scratch_bytes, const void* saved, size_t saved_bytes,
                       const void* cache, size_t cache_bytes, void* grad_feat, void* grad_logits,
                       ls_stream_t stream);

/* ---- consumer of bev: ParkingModel.add_target_bev (model/parking_model.py:28-46) ----------
 * Writes the target channel: ones on [px-4, px+4) x [py-4, py+4) (python slice semantics) around
 * target_pix i32[B,2] = the (noised) target pixel, zeros elsewhere, into `out` with element
 * strides (b, x, y) - a [B,1,X,Y] tensor, or channel C of a channels-last [B,X,Y,C+1] buffer
 * whose first C channels ls_forward wrote (no torch.cat copy of the BEV). */
int ls_target_bev(const int32_t* target_pix, int32_t B, int32_t X, int32_t Y, float* out, int64_t stride_b,
                  int64_t stride_x, int64_t stride_y, ls_stream_t stream);

/* ---- consumer of pred_depth: DepthLoss (loss/depth_loss.py:18-48) ------------------------
 * prob [BN,D,fh,fw] of `dtype` (ls_forward's pred_depth), gt_depth f32[BN, fh*down, fw*down]
 * (metric depth per pixel, 0 = no return; data['depth'] of dataset/carla_dataset.py viewed as
 * [B*N,H,W]).  d_off = float32(d_bound[0] - d_bound[2]), d_step = float32(d_bound[2]).
 * Outputs: labels i32[BN*fh*fw] (0 background, k >= 1: depth bin k-1 is the positive class;
 * kept for the backward), out2 f32[2] = {loss, 1 / max(1, #foreground pixels)}.
 * ws: >= ls_depth_loss_ws_bytes() of scratch (per-CTA partial sums; sums have a fixed order).
 * Limits: BN <= 65535, fh*fw < 2^31, BN * ceil(fh*fw / 32) < 2^31 (LS_ERR_UNSUPPORTED; ws_bytes() = 0). */
size_t ls_depth_loss_ws_bytes(int32_t BN, int32_t fh, int32_t fw);
int ls_depth_loss_fwd(const void* prob, int dtype, const float* gt_depth, int32_t BN, int32_t D, int32_t fh,
                      int32_t fw, int32_t down, float d_off, float d_step, int32_t* labels, void* ws,
                      size_t ws_bytes, float* out2, ls_stream_t stream);
/* grad_prob [BN,D,fh,fw] of `dtype` = grad_out * d loss / d prob (aten's binary_cross_entropy
 * backward, zero on background pixels); grad_out: device scalar (NULL = 1). */
int ls_depth_loss_bwd(const void* prob, int dtype, const int32_t* labels, const float* fwd_out2,
                      const float* grad_out, int32_t BN, int32_t D, int32_t fh, int32_t fw, void* grad_prob,
                      ls_stream_t stream);

/* Developer aid: per-phase clock64 totals of ls_splat_fwd (summed over CTAs) since the last
 * call; LS_ERR_UNSUPPORTED unless the library was built with -DLS_PROFILE. */
int ls_debug_phase_cycles(uint64_t* out8);

/* Number of kernel launches (and copies) issued by this library since load; bench.py
 * reads it around the timed region to report gpu_launches. */
int64_t ls_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* LS_B200_H_ */
