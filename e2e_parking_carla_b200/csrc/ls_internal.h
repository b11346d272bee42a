// Host-side launchers shared between translation units (not part of the C ABI).
#pragma once
#include "ls_common.cuh"

// ls_index.cu
// placement handles at most 32 depth bins per thread of 16 depth groups, and its pixel-major
// staging tile (32 pixels x (D|1) 8-byte records) must fit 48 KB of shared memory
static inline int ls_max_depth_bins() { return 191; }
int ls_launch_camera_transform(const float* intr, const float* extr, int BN, float* M, float* t, cudaStream_t s);
// also zeroes the n_extra ints at `extra` (the backward's group tickets; may be NULL)
int ls_launch_zero_counts(int* counts, const LsDims& dm, const LsGrid& g, int* extra, int n_extra, cudaStream_t s);
int ls_launch_index(const float* M, const float* t, const float* frustum, const LsDims& dm, const LsGrid& g,
                    int* rank, int* cell, int* within, int* counts, cudaStream_t s);
int ls_launch_index_geom(const float* geom, const LsDims& dm, const LsGrid& g, int* rank, int* cell, int* within,
                         int* counts, cudaStream_t s);
int ls_launch_export(const float* M, const float* t, const float* frustum, const LsDims& dm, const LsGrid& g,
                     float* geom, long long* vox, unsigned char* keep, long long* rank64, cudaStream_t s);
int ls_launch_scan(const int* counts, const LsDims& dm, const LsGrid& g, int* seg_start, int* tile_order,
                   int* tile_tot /* scratch i32[B*tiles] or NULL */, cudaStream_t s);
int ls_launch_place(const int* cell, const int* within, const void* prob, int dtype, const LsDims& dm,
                    const LsGrid& g, const int* seg_start, int2* recs, int2* pix_recs, cudaStream_t s);
int ls_launch_export_cell_counts(const int* seg_start, const LsGrid& g, int B, int b, long long* out, int* kept,
                                 cudaStream_t s);

// ls_dense.cu
int ls_launch_softmax(const void* logits, int dtype, const LsDims& dm, void* prob, cudaStream_t s);
int ls_launch_softmax_bwd(const void* prob, const float* gprob_pm, const void* gext, int dtype, const LsDims& dm,
                          void* glogits, cudaStream_t s);
// [images][C][HW] -> [images][HW][Cp] (zero padded) and back (drops the padding)
int ls_launch_to_nhwc(const void* src, int dtype, int images, int C, int Cp, int HW, void* dst, cudaStream_t s);
int ls_launch_from_nhwc(const void* src, int dtype, int images, int C, int Cp, int HW, void* dst, cudaStream_t s);
// softmax backward + (gfeatT != NULL) NHWC -> NCHW of grad_feat in ONE launch that overlaps the tail of the
// gather: no dependency wait at its top, every CTA waits for its image's arrivals in ready[] instead
bool ls_epilogue_supports(const LsDims& dm);     // depth-bin counts with a pixel-stationary instantiation
int ls_launch_bwd_epilogue(const void* prob, const float* gprob_pm, const void* gext, const void* gfeatT, int dtype,
                           const LsDims& dm, void* glogits, void* gfeat_nchw, int* ready, int target, cudaStream_t s);

int ls_launch_target_bev(const int* pix, int B, int X, int Y, float* out, long long sb, long long sx, long long sy,
                         cudaStream_t s);

// ls_splat.cu
int ls_debug_fetch_phase_cycles(unsigned long long* out8);   // only with -DLS_PROFILE
size_t ls_sorted_records_capacity(const LsDims& dm, const LsGrid& g);   // per sample, in 8-byte records
// recs == NULL: recs_sorted is already canonical (static-rig cache); perm (may be NULL): slot -> point id
int ls_launch_splat_fwd(const void* featT, int dtype, const int2* recs, const int* seg_start, const int* tile_order,
                        int2* recs_sorted, int* perm, const LsDims& dm, const LsGrid& g, float* bev,
                        const LsBevStrides& st, cudaStream_t s);
int ls_launch_refresh(const void* prob, int dtype, const int* perm, const int* seg_start, const LsDims& dm,
                      const LsGrid& g, int2* recs_sorted, int2* pix_recs, int* zero_ints, int n_zero, cudaStream_t s);
int ls_launch_bwd_transpose(const float* gbev, const LsBevStrides& st, const int* seg_start, const LsDims& dm,
                            const LsGrid& g, float* gT, cudaStream_t s);
// ready (may be NULL): i32 per-image arrival counters, zero on entry; with it the gather triggers its
// dependent launch early and the NEXT launch in the stream must be ls_launch_bwd_epilogue with the same
// counters (it waits on them per image, and leaves them zero).
bool ls_gather_can_overlap(const LsDims& dm);
int ls_gather_ready_target(const LsDims& dm);      // arrivals that complete an image
static inline int ls_bwd_flag_ints(const LsDims& dm) { return 2 * dm.B * dm.N; }   // arrivals + departures
int ls_launch_bwd_gather(const void* rows_base, long long sample_stride, long long row_stride, int mode /* LsGradIn */,
                         const void* featT, int dtype, const int2* pix_recs, const LsDims& dm, const LsGrid& g,
                         float* gprob_pm, void* gfeatT, int* ready, cudaStream_t s);
int ls_classify_bev_out(const float* p, const LsBevStrides& st, const LsDims& dm, const LsGrid& g);   // LsBevOut
int ls_classify_grad_in(const float* p, const LsBevStrides& st, const LsDims& dm, const LsGrid& g);   // LsGradIn
