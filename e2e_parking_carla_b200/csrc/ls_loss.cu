// Depth supervision of the lift-splat's second output:
// DepthLoss (loss/depth_loss.py:18-48) as two kernels forward and one backward.
//   labels: 8x8 (down x down) min-pool of the metric ground-truth depth ignoring zeros
//           (:32-40), bin index (gt - (d_lo - d_step)) / d_step truncated, valid in [0, D+1)
//           else 0 (:42-44); the one-hot row drops class 0 (:45-46), so label k >= 1 marks bin
//           k-1 and label 0 marks a background pixel (no positive bin, excluded by fg_mask :21);
//   loss:   sum over foreground pixels and all D bins of binary cross entropy on the depth
//           PROBABILITIES (:24-28; aten clamps both logs at -100), divided by max(1, #foreground).
// Everything the reference does with ~10 elementwise kernels, two boolean-index host syncs
// and a [B*N*h*w, D+1] one-hot tensor happens in registers here; sums are taken in a fixed
// order (per-CTA tree, then one CTA over the partials in double), so the loss is deterministic.
#include "ls_internal.h"

#define LS_DL_GROUPS 8

template <typename T>
__global__ void __launch_bounds__(32 * LS_DL_GROUPS)
ls_depth_loss_fwd_kernel(const T* __restrict__ prob, const float* __restrict__ gt, int D, int fh, int fw, int down,
                         float off, float step, int* __restrict__ labels, float* __restrict__ partial) {
  ls_pdl_trigger();
  ls_pdl_wait();
  __shared__ float red[LS_DL_GROUPS][33];
  __shared__ int lab[32];
  const int HW = fh * fw, W = fw * down;
  const int img = blockIdx.y, rc0 = blockIdx.x * 32;
  const int lane = threadIdx.x & 31, dg = threadIdx.x >> 5;
  const int rc = rc0 + lane;
  const bool on = rc < HW;
  // ---- min-pool of the pixel's down x down block, zeros ignored (depth_loss.py:34-39) ----
  float m = 1e5f;
  if (on) {
    const int row = rc / fw, col = rc % fw;
    const float* blk = gt + ((size_t)img * fh * down + (size_t)row * down) * W + (size_t)col * down;
    for (int r = dg; r < down; r += LS_DL_GROUPS) {
      const float* p = blk + (size_t)r * W;
      if (down == 8 && (W & 3) == 0 && ((uintptr_t)gt & 15) == 0) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) m = fminf(m, v[k] == 0.0f ? 1e5f : v[k]);
      } else {
        for (int k = 0; k < down; ++k) {
          const float v = __ldg(p + k);
          m = fminf(m, v == 0.0f ? 1e5f : v);
        }
      }
    }
  }
  red[dg][lane] = m;
  __syncthreads();
  if (dg == 0) {
#pragma unroll
    for (int j = 1; j < LS_DL_GROUPS; ++j) m = fminf(m, red[j][lane]);
    // (gt - (d_lo - d_step)) / d_step, kept where 0 <= v < D + 1, truncated (depth_loss.py:42-46)
    const float v = __fdiv_rn(__fsub_rn(m, off), step);
    const int l = (v < (float)(D + 1) && v >= 0.0f) ? __float2int_rz(v) : 0;
    lab[lane] = on ? l : 0;
    if (on) labels[(size_t)img * HW + rc] = l;
  }
  __syncthreads();
  // ---- binary cross entropy over this thread's depth bins of a foreground pixel ----
  const int l = lab[lane];
  float s = 0.0f;
  if (l >= 1) {
    const T* src = prob + (size_t)img * D * HW + rc;
    for (int d = dg; d < D; d += LS_DL_GROUPS) {
      const float p = ls_to_float(src[(size_t)d * HW]);
      // aten binary_cross_entropy: (y - 1) * max(log(1 - p), -100) - y * max(log(p), -100)
      s += (d == l - 1) ? -fmaxf(logf(p), -100.0f) : -fmaxf(logf(1.0f - p), -100.0f);
    }
  }
  __syncthreads();
  red[dg][lane] = s;
  __syncthreads();
  if (dg == 0) {
#pragma unroll
    for (int j = 1; j < LS_DL_GROUPS; ++j) s += red[j][lane];
    float cnt = l >= 1 ? 1.0f : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if (lane == 0) {
      const size_t blk = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
      partial[2 * blk + 0] = s;
      partial[2 * blk + 1] = cnt;
    }
  }
}

// one CTA: partial sums in a fixed order (double) -> out[0] = loss, out[1] = 1 / max(1, #foreground)
__global__ void __launch_bounds__(256)
ls_depth_loss_finish_kernel(const float* __restrict__ partial, int nblocks, float* __restrict__ out) {
  ls_pdl_trigger();
  ls_pdl_wait();
  __shared__ double ss[256], cc[256];
  double s = 0.0, c = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += 256) {
    s += (double)partial[2 * i];
    c += (double)partial[2 * i + 1];
  }
  ss[threadIdx.x] = s;
  cc[threadIdx.x] = c;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      ss[threadIdx.x] += ss[threadIdx.x + o];
      cc[threadIdx.x] += cc[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double n = cc[0] > 1.0 ? cc[0] : 1.0;      // max(1.0, fg_mask.sum()) (depth_loss.py:28)
    out[0] = (float)(ss[0] / n);
    out[1] = (float)(1.0 / n);
  }
}

// d loss / d prob: grad_out / max(1, #fg) * (p - y) / max((1 - p) * p, 1e-12) on foreground pixels
// (aten binary_cross_entropy_backward), zero elsewhere.
template <typename T>
__global__ void __launch_bounds__(256)
ls_depth_loss_bwd_kernel(const T* __restrict__ prob, const int* __restrict__ labels, const float* __restrict__ fwd_out,
                         const float* __restrict__ grad_out, int D, int HW, T* __restrict__ gprob) {
  ls_pdl_trigger();
  ls_pdl_wait();
  const int img = blockIdx.y;
  const int lane = threadIdx.x & 31, dg = threadIdx.x >> 5;
  const int rc = blockIdx.x * 32 + lane;
  if (rc >= HW) return;
  const int l = labels[(size_t)img * HW + rc];
  const float scale = (grad_out ? __ldg(grad_out) : 1.0f) * __ldg(fwd_out + 1);
  const size_t base = (size_t)img * D * HW + rc;
  for (int d = dg; d < D; d += 8) {
    float g = 0.0f;
    if (l >= 1) {
      const float p = ls_to_float(prob[base + (size_t)d * HW]);
      const float y = (d == l - 1) ? 1.0f : 0.0f;
      g = scale * (p - y) / fmaxf((1.0f - p) * p, 1e-12f);
    }
    gprob[base + (size_t)d * HW] = ls_from_float<T>(g);
  }
}

static inline int ls_dl_blocks(int BN, int HW) { return ((HW + 31) / 32) * BN; }

extern "C" {

// Launch limits: grid.y = BN, fh * fw and the block count are 32-bit.
static inline bool ls_dl_shape_ok(int BN, int fh, int fw) {
  if (BN <= 0 || fh <= 0 || fw <= 0 || BN > 65535) return false;
  const long long hw = (long long)fh * fw;
  return hw < (1LL << 31) && ((hw + 31) / 32) * BN < (1LL << 31);
}

size_t ls_depth_loss_ws_bytes(int32_t BN, int32_t fh, int32_t fw) {
  if (!ls_dl_shape_ok(BN, fh, fw)) return 0;
  return (size_t)ls_dl_blocks(BN, fh * fw) * 2 * sizeof(float);
}

int ls_depth_loss_fwd(const void* prob, int dtype, const float* gt_depth, int32_t BN, int32_t D, int32_t fh, int32_t fw,
                      int32_t down, float d_off, float d_step, int32_t* labels, void* ws, size_t ws_bytes,
                      float* out2, ls_stream_t stream) {
  if (!prob || !gt_depth || !labels || !ws || !out2) return LS_ERR_BAD_ARG;
  if (BN <= 0 || D <= 0 || fh <= 0 || fw <= 0 || down <= 0 || !(d_step > 0.0f)) return LS_ERR_BAD_ARG;
  if (dtype != LS_F32 && dtype != LS_BF16) return LS_ERR_BAD_ARG;
  if (!ls_dl_shape_ok(BN, fh, fw)) return LS_ERR_UNSUPPORTED;
  if (ws_bytes < ls_depth_loss_ws_bytes(BN, fh, fw)) return LS_ERR_WORKSPACE;
  const int HW = fh * fw;
  dim3 grid((HW + 31) / 32, BN);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == LS_F32)
    LS_LAUNCH(ls_depth_loss_fwd_kernel<float>, grid, dim3(32 * LS_DL_GROUPS), 0, s, (const float*)prob, gt_depth, D, fh,
              fw, down, d_off, d_step, labels, (float*)ws);
  else
    LS_LAUNCH(ls_depth_loss_fwd_kernel<__nv_bfloat16>, grid, dim3(32 * LS_DL_GROUPS), 0, s, (const __nv_bfloat16*)prob,
              gt_depth, D, fh, fw, down, d_off, d_step, labels, (float*)ws);
  LS_LAUNCH(ls_depth_loss_finish_kernel, dim3(1), dim3(256), 0, s, (const float*)ws, ls_dl_blocks(BN, HW), out2);
  return LS_OK;
}

int ls_depth_loss_bwd(const void* prob, int dtype, const int32_t* labels, const float* fwd_out2, const float* grad_out,
                      int32_t BN, int32_t D, int32_t fh, int32_t fw, void* grad_prob, ls_stream_t stream) {
  if (!prob || !labels || !fwd_out2 || !grad_prob) return LS_ERR_BAD_ARG;
  if (BN <= 0 || D <= 0 || fh <= 0 || fw <= 0) return LS_ERR_BAD_ARG;
  if (dtype != LS_F32 && dtype != LS_BF16) return LS_ERR_BAD_ARG;
  if (!ls_dl_shape_ok(BN, fh, fw)) return LS_ERR_UNSUPPORTED;
  const int HW = fh * fw;
  dim3 grid((HW + 31) / 32, BN);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == LS_F32)
    LS_LAUNCH(ls_depth_loss_bwd_kernel<float>, grid, dim3(256), 0, s, (const float*)prob, labels, fwd_out2, grad_out, D,
              HW, (float*)grad_prob);
  else
    LS_LAUNCH(ls_depth_loss_bwd_kernel<__nv_bfloat16>, grid, dim3(256), 0, s, (const __nv_bfloat16*)prob, labels,
              fwd_out2, grad_out, D, HW, (__nv_bfloat16*)grad_prob);
  return LS_OK;
}

}  // extern "C"
