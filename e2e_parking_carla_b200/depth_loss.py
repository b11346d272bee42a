"""Drop-in ``DepthLoss`` (reference: loss/depth_loss.py:10-48) on the sm_100a kernels.

Same constructor argument (``cfg`` with ``d_bound`` and ``bev_down_sample``), same call:
``loss = DepthLoss(cfg)(depth_preds, depth_labels)`` with ``depth_preds`` the lift-splat's
``pred_depth`` ``[B*N, D, h, w]`` (probabilities) and ``depth_labels`` ``[B, N, H, W]`` metric
depth.  One kernel builds the min-pooled bin labels and the per-CTA cross-entropy sums, a
second adds them in a fixed order; the backward is one elementwise kernel.  No one-hot
tensor, no boolean-index host syncs.  CUDA tensors only.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
from torch import nn

from . import _lib
from .lift_splat import _dtype_code, _need_cuda, _ptr, _stream


class _DepthLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prob, gt, down, d_off, d_step):
        _need_cuda(prob, gt)
        lib = _lib.load()
        prob_c = prob.contiguous()
        bn, d, fh, fw = prob_c.shape
        gt_c = gt.detach().to(torch.float32).contiguous()
        if gt_c.numel() != bn * fh * down * fw * down:
            raise ValueError("depth labels %s do not match predictions %s at down-sample %d"
                             % (tuple(gt.shape), tuple(prob.shape), down))
        labels = torch.empty(bn * fh * fw, dtype=torch.int32, device=prob.device)
        ws = torch.empty(lib.ls_depth_loss_ws_bytes(bn, fh, fw), dtype=torch.uint8, device=prob.device)
        out2 = torch.empty(2, dtype=torch.float32, device=prob.device)
        code = _dtype_code(prob_c)
        _lib.check(lib.ls_depth_loss_fwd(_ptr(prob_c), code, _ptr(gt_c), bn, d, fh, fw, down, d_off, d_step,
                                         _ptr(labels), _ptr(ws), ws.numel(), _ptr(out2), _stream(prob_c)),
                   "ls_depth_loss_fwd")
        ctx.save_for_backward(prob_c, labels, out2)
        ctx.code = code
        return out2[0].clone()

    @staticmethod
    def backward(ctx, grad_out):
        prob, labels, out2 = ctx.saved_tensors
        bn, d, fh, fw = prob.shape
        g = grad_out.detach().to(torch.float32).reshape(1).contiguous()
        gprob = torch.empty_like(prob)
        _lib.check(_lib.load().ls_depth_loss_bwd(_ptr(prob), ctx.code, _ptr(labels), _ptr(out2), _ptr(g), bn, d, fh,
                                                 fw, _ptr(gprob), _stream(prob)), "ls_depth_loss_bwd")
        return gprob, None, None, None, None


class DepthLoss(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.d_bound = self.cfg.d_bound
        self.down_sample_factor = self.cfg.bev_down_sample
        self.depth_channels = int((self.cfg.d_bound[1] - self.cfg.d_bound[0]) / self.cfg.d_bound[2])
        # (gt - (d_lo - d_step)) / d_step: python-float scalars that torch casts to float32
        self._off = float(np.float32(self.d_bound[0] - self.d_bound[2]))
        self._step = float(np.float32(self.d_bound[2]))

    def forward(self, depth_preds, depth_labels):
        if depth_preds.shape[1] != self.depth_channels:
            raise ValueError("expected %d depth bins, got %d" % (self.depth_channels, depth_preds.shape[1]))
        return _DepthLossFn.apply(depth_preds, depth_labels, int(self.down_sample_factor), self._off, self._step)
