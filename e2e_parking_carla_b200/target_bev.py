"""``add_target_bev`` of the reference's ``ParkingModel`` (model/parking_model.py:28-46), the
immediate consumer of the lift-splat output, kept in the splat's native channels-last layout.

The reference allocates a zero map, stamps an 8x8 square of ones around the noised target pixel
in a python loop over the batch and ``torch.cat``s it behind the 64 BEV channels (a 10 MB/sample
copy that also turns a channels_last BEV back into NCHW).  Here the target pixel is computed with
the reference's own tensor ops (same float ops, same ``torch.rand_like`` call, so the same
generator state gives the same noise), and ONE kernel writes the target channel:

* when ``bev_feature`` came from a ``BevModel(..., spare_channels=1)`` it is a view of a
  channels-last ``[B, X, Y, C+1]`` buffer: the stamp goes into channel C in place, nothing is copied;
* otherwise the stamp is written into a channels-last ``[B, 1, X, Y]`` map and concatenated, which
  keeps the result (and the gradient coming back) channels_last.

Returns ``(bev_feature[B, C+1, X, Y], bev_target[B, 1, X, Y])`` like the reference.
"""
from __future__ import annotations

import torch

from . import _lib
from .lift_splat import _need_cuda, _ptr, _stream


def target_pixels(bev_shape, target_point: torch.Tensor, cfg, noise: bool = True) -> torch.Tensor:
    """i32[B,2]: the target pixel, op for op as model/parking_model.py:32-37."""
    _, _, h, w = bev_shape
    x_pixel = (h / 2 + target_point[:, 0] / cfg.bev_x_bound[2]).unsqueeze(0).T.int()
    y_pixel = (w / 2 + target_point[:, 1] / cfg.bev_y_bound[2]).unsqueeze(0).T.int()
    pix = torch.cat([x_pixel, y_pixel], dim=1)
    if noise:
        pix = pix + (torch.rand_like(pix, dtype=torch.float) * 10 - 5).int()
    return pix.contiguous()


def _stamp(pix: torch.Tensor, out: torch.Tensor) -> None:
    """out: [B,1,X,Y] view with arbitrary strides."""
    b, _, x, y = out.shape
    _lib.check(_lib.load().ls_target_bev(_ptr(pix), b, x, y, _ptr(out), out.stride(0), out.stride(2), out.stride(3),
                                         _stream(out)), "ls_target_bev")


class _StampInPlace(torch.autograd.Function):
    """bev = first C channels of a channels-last [B,C+1,X,Y] buffer -> the whole buffer with the
    target channel filled in.  Backward hands the first C channels of the gradient back (a view)."""

    @staticmethod
    def forward(ctx, bev, pix):
        full = bev._base
        _stamp(pix, full[:, bev.shape[1]:])
        ctx.channels = bev.shape[1]
        return torch.as_strided(full, full.shape, full.stride())

    @staticmethod
    def backward(ctx, grad_full):
        return grad_full[:, :ctx.channels], None


def _has_spare_channel(bev: torch.Tensor) -> bool:
    full = bev._base
    if full is None or full.dim() != 4 or bev.storage_offset() != full.storage_offset():
        return False
    return (full.shape[1] == bev.shape[1] + 1 and full.shape[0] == bev.shape[0] and full.shape[2:] == bev.shape[2:]
            and full.stride() == bev.stride() and full.is_contiguous(memory_format=torch.channels_last))


def add_target_bev(bev_feature: torch.Tensor, target_point: torch.Tensor, cfg, noise: bool = True):
    _need_cuda(bev_feature, target_point)
    pix = target_pixels(bev_feature.shape, target_point, cfg, noise)
    if _has_spare_channel(bev_feature):
        full = _StampInPlace.apply(bev_feature, pix)
        return full, full[:, bev_feature.shape[1]:].detach()
    b, _, h, w = bev_feature.shape
    target = torch.empty((b, h, w, 1), dtype=torch.float32, device=bev_feature.device).permute(0, 3, 1, 2)
    _stamp(pix, target)
    if not bev_feature.is_contiguous(memory_format=torch.channels_last):
        target = target.contiguous()          # an NCHW BEV stays NCHW, as in the reference
    return torch.cat([bev_feature, target], dim=1), target
