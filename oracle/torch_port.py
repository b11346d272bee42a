"""TEST / BASELINE INFRASTRUCTURE ONLY - the reference's CPU algorithm, restated.

A port of the reference lift-splat *as the reference computes it* (same aten op chain:
materialised outer product, per-sample loop, boolean masks, argsort, float32 cumsum
trick, index_put scatter) written against plain torch CPU ops so that it can travel to
the GPU box, where ``/root/reference`` does not exist.  It is what ``bench.py`` times as
``cpu_baseline`` (kind "port") and under ``--impl reference``, with all host threads.
It is never imported by the product package.

Pinned by ``tests/test_oracle_golden.py``: bit-identical ranks and float32 BEV output to
the unmodified reference on the golden rigs (and to the live reference when present).

Reference lines restated: model/bev_model.py:45-57 (geometry), :59-72 (lift),
:74-107 (projection loop), tool/geometry.py:285-317 (VoxelsSumming).
"""
from __future__ import annotations

import torch


class SegmentCumsum(torch.autograd.Function):
    """Sum of consecutive rows that share a rank, via prefix sums (the 'cumsum trick',
    tool/geometry.py:289-305); backward hands each row the gradient of its segment
    (tool/geometry.py:307-317)."""

    @staticmethod
    def forward(ctx, rows, ranks):
        prefix = torch.cumsum(rows, dim=0)
        last = torch.ones(rows.shape[0], dtype=torch.bool, device=rows.device)
        last[:-1] = ranks[:-1] != ranks[1:]
        ends = prefix[last]
        sums = ends.clone()
        sums[1:] = ends[1:] - ends[:-1]
        ctx.save_for_backward(last)
        return sums, last

    @staticmethod
    def backward(ctx, grad_sums, _):
        (last,) = ctx.saved_tensors
        seg_id = torch.cumsum(last, 0)
        seg_id[last] -= 1
        return grad_sums[seg_id], None


def camera_geometry(frustum, intrinsics, extrinsics):
    """geom[B,N,D,h,w,3] with the reference's op chain (two inverses, small matmuls)."""
    cam_to_ego = torch.inverse(extrinsics)
    rot, trans = cam_to_ego[..., :3, :3], cam_to_ego[..., :3, 3]
    b, n = trans.shape[:2]
    uvd = frustum[None, None, ..., None]                                    # [1,1,D,h,w,3,1]
    rays = torch.cat((uvd[..., :2, :] * uvd[..., 2:3, :], uvd[..., 2:3, :]), dim=5)
    lift = rot.matmul(torch.inverse(intrinsics)).view(b, n, 1, 1, 1, 3, 3)
    pts = lift.matmul(rays).squeeze(-1)
    pts = pts + trans.view(b, n, 1, 1, 1, 3)
    return pts


def lift(feat, depth_logits, batch, cams):
    """(x[B,N,D,h,w,C] view of the materialised outer product, prob)."""
    prob = depth_logits.softmax(dim=1)
    vol = prob.unsqueeze(1) * feat.unsqueeze(2)                              # [BN,C,D,h,w]
    vol = vol.view(batch, cams, *vol.shape[1:]).permute(0, 1, 3, 4, 5, 2)
    return vol, prob


def project(geom, vol, bev_start, bev_res, bev_dim):
    """Per-sample voxelise / mask / rank / argsort / cumsum-trick / scatter."""
    batch, n, d, h, w, c = vol.shape
    X, Y, Z = (int(v) for v in bev_dim)
    out = torch.zeros((batch, c, X, Y), dtype=torch.float, device=vol.device)
    npts = n * d * h * w
    origin = bev_start - bev_res / 2.0
    for b in range(batch):
        rows = vol[b].reshape(npts, c)
        vox = ((geom[b] - origin) / bev_res).view(npts, 3).long()
        inside = ((vox[:, 0] >= 0) & (vox[:, 0] < X) & (vox[:, 1] >= 0) & (vox[:, 1] < Y)
                  & (vox[:, 2] >= 0) & (vox[:, 2] < Z))
        rows, vox = rows[inside], vox[inside]
        ranks = vox[:, 0] * (Y * Z) + vox[:, 1] * Z + vox[:, 2]
        perm = ranks.argsort()
        rows, vox, ranks = rows[perm], vox[perm], ranks[perm]
        sums, last = SegmentCumsum.apply(rows, ranks)
        vox = vox[last]
        cells = torch.zeros((Z, X, Y, c), dtype=sums.dtype, device=vol.device)
        cells[vox[:, 2], vox[:, 0], vox[:, 1]] = sums
        out[b] = cells.permute(0, 3, 1, 2).squeeze(0)
    return out


def lift_splat_cpu(feat, depth_logits, intrinsics, extrinsics, frustum, bev_start, bev_res, bev_dim):
    """Whole path: (bev f32[B,C,X,Y], prob).  Differentiable w.r.t. feat / depth_logits."""
    b, n = intrinsics.shape[:2]
    geom = camera_geometry(frustum, intrinsics, extrinsics)
    vol, prob = lift(feat, depth_logits, b, n)
    return project(geom, vol, bev_start, bev_res, bev_dim), prob


def ranks_cpu(intrinsics, extrinsics, frustum, bev_start, bev_res, bev_dim):
    """int64 rank per point, -1 where masked out (test helper)."""
    geom = camera_geometry(frustum, intrinsics, extrinsics)
    b = geom.shape[0]
    X, Y, Z = (int(v) for v in bev_dim)
    vox = ((geom - (bev_start - bev_res / 2.0)) / bev_res).view(b, -1, 3).long()
    inside = ((vox[..., 0] >= 0) & (vox[..., 0] < X) & (vox[..., 1] >= 0) & (vox[..., 1] < Y)
              & (vox[..., 2] >= 0) & (vox[..., 2] < Z))
    ranks = vox[..., 0] * (Y * Z) + vox[..., 1] * Z + vox[..., 2]
    return torch.where(inside, ranks, torch.full_like(ranks, -1))


def fwd_bwd_step(feat, depth_logits, intrinsics, extrinsics, frustum, bev_start, bev_res, bev_dim,
                 grad_bev, grad_prob):
    """One forward+backward pass; returns (bev, prob, grad_feat, grad_logits)."""
    f = feat.detach().requires_grad_(True)
    z = depth_logits.detach().requires_grad_(True)
    bev, prob = lift_splat_cpu(f, z, intrinsics, extrinsics, frustum, bev_start, bev_res, bev_dim)
    torch.autograd.backward([bev, prob], [grad_bev, grad_prob])
    return bev.detach(), prob.detach(), f.grad, z.grad


def add_target_bev_ref(bev_feature, target_point, x_res, y_res, noise=True):
    """model/parking_model.py:28-46 restated (test oracle): zero map, 8x8 stamp of ones around
    the noised target pixel via python slices (negative bounds wrap like python's), channel cat.
    Draws its noise with the same ``torch.rand_like`` call, so a seeded generator reproduces it."""
    b, c, h, w = bev_feature.shape
    target_map = torch.zeros((b, 1, h, w), dtype=torch.float, device=bev_feature.device)
    px = (h / 2 + target_point[:, 0] / x_res).unsqueeze(0).T.int()
    py = (w / 2 + target_point[:, 1] / y_res).unsqueeze(0).T.int()
    pix = torch.cat([px, py], dim=1)
    if noise:
        pix = pix + (torch.rand_like(pix, dtype=torch.float) * 10 - 5).int()
    for i in range(b):
        cx, cy = int(pix[i, 0]), int(pix[i, 1])
        target_map[i, 0][cx - 4:cx + 4, cy - 4:cy + 4] = 1.0
    return torch.cat([bev_feature, target_map], dim=1), target_map
