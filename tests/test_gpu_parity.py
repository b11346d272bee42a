"""Parity of the sm_100a kernels (through the C ABI) against the reference.

Three sources of truth, all CPU:
  * golden fixtures frozen from the unmodified reference (tests/golden/*.npz);
  * the numpy oracle (oracle/lift_splat_oracle.py), itself pinned to those fixtures;
  * size-independent properties at BASELINE.json's full sizes.

Bars (BASELINE.json north_star): voxel indices / keep mask / sorted ranks BIT-EXACT;
BEV features and gradients <= 1e-5 relative (fp32) or <= 2e-2 (bf16) against the
reference run in float64.
"""
import os

import numpy as np
import pytest
import torch

from _util import GOLDEN_SHAPES, Golden, frustum_of, grid_of, relerr, sha
from e2e_parking_carla_b200.synthetic import (LiftSplatShape, make_cfg, make_encoder_outputs, make_rig,
                                              make_upstream_grads)

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-5     # north_star: "<=1e-5 relative error in fp32"
BF16_TOL = 2e-2     # north_star: "<=2e-2 in bf16"
DEV = "cuda:0"


def _ls():
    from e2e_parking_carla_b200 import lift_splat as ls
    return ls


def _grid_spec(shape: LiftSplatShape):
    ls = _ls()
    res, start, dim = grid_of(shape)
    return ls.GridSpec(tuple(float(v) for v in start), tuple(float(v) for v in res), tuple(int(v) for v in dim))


def _ls_shape(shape: LiftSplatShape, channels=None):
    ls = _ls()
    return ls.make_shape(shape.batch, shape.cams, shape.depth_bins, shape.fh, shape.fw,
                         channels or shape.channels, _grid_spec(shape))


def _dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else a
    t = t.to(DEV)
    return t.to(dtype) if dtype is not None else t


# ------------------------------------------------------------------------------------
# indices: bit-exact
# ------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(GOLDEN_SHAPES))
def test_rank_bit_exact_vs_reference(lib, name):
    """ls_index on the reference's own M,t reproduces the reference's rank / keep mask
    for every point (model/bev_model.py:85-95)."""
    ls = _ls()
    g = Golden(name)
    s = _ls_shape(g.shape)
    fr = _dev(frustum_of(g.shape))
    assert sha(frustum_of(g.shape)) == str(g["frustum_sha"])
    rank = ls.index(_dev(g["M_ref"]), _dev(g["t_ref"]), fr, s)
    ref = g["rank_ref"]
    got = rank.cpu().numpy()
    assert got.dtype == np.int32 and got.shape == ref.shape
    assert np.array_equal(got, ref), "%d of %d ranks differ" % ((got != ref).sum(), ref.size)
    # keep mask and kept counts per camera
    kept = (got >= 0).reshape(g.shape.batch, g.shape.cams, -1).sum(-1)
    assert np.array_equal(kept, g["kept_per_cam"])


@pytest.mark.parametrize("name", list(GOLDEN_SHAPES))
def test_export_indices_and_geometry_vs_oracle(lib, name):
    """The debug export (vox int64, keep, rank int64) and ls_geometry equal the oracle's
    per-op float32 restatement bit for bit."""
    from oracle import lift_splat_oracle as lo
    ls = _ls()
    g = Golden(name)
    s = _ls_shape(g.shape)
    fr_np = frustum_of(g.shape)
    res, start, dim = grid_of(g.shape)
    geom_o = lo.geometry(g["M_ref"], g["t_ref"], fr_np)
    vox_o, keep_o, rank_o = lo.voxel_index(geom_o, start, res, dim)
    M, t, fr = _dev(g["M_ref"]), _dev(g["t_ref"]), _dev(fr_np)
    geom = ls.geometry(M, t, fr, s).cpu().numpy()
    assert sha(geom) == str(g["geom_sha"]), "geometry differs from the reference's get_geometry"
    assert np.array_equal(geom, geom_o)
    vox, keep, rank = ls.export_indices(M, t, fr, s)
    assert np.array_equal(vox.cpu().numpy(), vox_o)
    assert np.array_equal(keep.cpu().numpy(), keep_o)
    assert np.array_equal(rank.cpu().numpy(), rank_o)


@pytest.mark.parametrize("name", list(GOLDEN_SHAPES))
def test_sorted_ranks_bit_exact(lib, name):
    """ranks[ranks.argsort()] (model/bev_model.py:96-97) and the segment count of
    VoxelsSumming (tool/geometry.py:295-296), from the CSR the counting sort builds."""
    ls = _ls()
    g = Golden(name)
    s = _ls_shape(g.shape)
    sh = g.shape
    rank, cell, within, counts = ls.index(_dev(g["M_ref"]), _dev(g["t_ref"]), _dev(frustum_of(sh)), s, for_sort=True)
    assert np.array_equal(rank.cpu().numpy(), g["rank_ref"])
    assert torch.equal(cell >= 0, rank >= 0)
    _, logits, _, _ = g.inputs()
    prob = ls.softmax(logits.to(DEV), s)
    seg, order, recs, pix = ls.sort(cell, within, counts, prob, s, with_pixel_index=True)
    # launch order of the splat: every tile exactly once, heaviest first
    tiles, cells_all, _ = ls.grid_cells(s)
    tc = cells_all // tiles
    assert torch.equal(order.sort(dim=1).values, torch.arange(tiles, device=DEV, dtype=torch.int32).expand(sh.batch, -1))
    tot = (seg[:, tc:tiles * tc + 1:tc] - seg[:, 0:tiles * tc:tc]).long()
    assert bool((torch.gather(tot, 1, order.long()).diff(dim=1) <= 0).all())
    # the single-CTA scan (no scratch) builds the same CSR and order
    counts2 = ls.index(_dev(g["M_ref"]), _dev(g["t_ref"]), _dev(frustum_of(sh)), s, for_sort=True)[3]
    seg1, order1, _, _ = ls.sort(cell, within, counts2, prob, s, parallel_scan=False)
    assert torch.equal(seg1[:, :-3], seg[:, :-3])
    if ls.grid_cells(s)[0] <= 1024:      # beyond that the single-CTA scan keeps the identity order
        assert torch.equal(order1, order)
    kept = ls.kept_counts(seg, s).cpu().numpy()
    assert np.array_equal(kept, g["kept_per_cam"].sum(-1))
    hw, D = sh.fh * sh.fw, sh.depth_bins
    dbits = int(np.ceil(np.log2(D)))
    for b in range(sh.batch):
        sr = ls.export_sorted_ranks(seg, s, b).cpu().numpy()
        assert sr.dtype == np.int64
        assert sha(sr) == str(g["sorted_rank_sha"][b])
        assert np.unique(sr).size == int(g["segments"][b])
        # the records are a permutation of exactly the kept points, each carrying its prob
        key = recs[b, :kept[b], 0].cpu().numpy().astype(np.int64) & 0xFFFFFF
        pixel, d = key >> dbits, key & ((1 << dbits) - 1)
        n, rc = pixel // hw, pixel % hw
        p = (n * D + d) * hw + rc
        r = g["rank_ref"][b]
        assert np.array_equal(np.sort(p), np.nonzero(r >= 0)[0])
        w = recs[b, :kept[b], 1].cpu().numpy().view(np.float32)
        assert np.array_equal(w, prob.view(sh.batch, -1)[b].cpu().numpy()[p])
    # pixel-major index: {cell, prob} of every depth bin of every pixel
    row_bytes = ls.padded_channels(sh.channels) * 4
    assert bool((pix[..., 0] % row_bytes == 0).all())
    pc = (pix[..., 0] // row_bytes).view(sh.batch, sh.cams, hw, D).permute(0, 1, 3, 2).reshape(sh.batch, -1)
    cells_padded = ls.grid_cells(s)[1]
    assert torch.equal(pc, torch.where(cell >= 0, cell, torch.full_like(cell, cells_padded)))   # dropped -> zero row
    pw = pix[..., 1].contiguous().view(torch.float32).view(sh.batch, sh.cams, hw, D).permute(0, 1, 3, 2)
    assert torch.equal(pw.reshape(-1), prob.view(sh.batch, sh.cams, D, hw).reshape(-1))


def test_camera_transform_bit_exact_vs_oracle(lib):
    """ls_camera_transform == oracle.camera_transform (fp64 Gauss-Jordan rounded once,
    unfused float32 R.K^-1), and differs from torch's LAPACK inverse by rounding only."""
    from oracle import lift_splat_oracle as lo
    ls = _ls()
    for jitter, seed in ((False, 0), (True, 4), (True, 5)):
        intr, extr = make_rig(8, 6, jitter=jitter, seed=seed)
        M, t = ls.camera_transform(intr.to(DEV), extr.to(DEV))
        Mo, to = lo.camera_transform(intr.numpy(), extr.numpy())
        assert np.array_equal(M.cpu().numpy(), Mo)
        assert np.array_equal(t.cpu().numpy(), to)
        inv = torch.inverse(extr)
        Mt = inv[..., :3, :3].matmul(torch.inverse(intr))
        assert (M.cpu() - Mt).abs().max() < 5e-7 * max(1.0, Mt.abs().max())
        assert (t.cpu() - inv[..., :3, 3]).abs().max() < 2e-6


def test_rig_a_end_to_end_indices_match_reference(lib):
    """From raw intrinsics/extrinsics (native inverse) the CARLA rig gives exactly the
    reference's ranks, including the 16 384 points that sit on voxel boundaries."""
    ls = _ls()
    g = Golden("rigA_b1_c4")
    s = _ls_shape(g.shape)
    M, t = ls.camera_transform(_dev(g["intrinsics"]), _dev(g["extrinsics"]))
    rank = ls.index(M, t, _dev(frustum_of(g.shape)), s).cpu().numpy()
    assert np.array_equal(rank, g["rank_ref"])
    assert (rank >= 0).sum() == 155296          # SURVEY.md 8c known answer (floor would give 150016)


# ------------------------------------------------------------------------------------
# forward / backward values
# ------------------------------------------------------------------------------------
def _run(shape, feat, logits, M, t, gb=None, gp=None, dtype=torch.float32):
    ls = _ls()
    feat = feat.to(DEV, dtype).requires_grad_(gb is not None)
    logits = logits.to(DEV, dtype).requires_grad_(gb is not None)
    bev, prob = ls.lift_splat(feat, logits, _dev(M), _dev(t), _dev(frustum_of(shape)), _grid_spec(shape))
    out = {"bev": bev.detach(), "prob": prob.detach()}
    if gb is not None:
        loss = (bev * gb.to(DEV)).sum() + (prob.float() * gp.to(DEV).float()).sum()
        loss.backward()
        out["grad_feat"], out["grad_logits"] = feat.grad, logits.grad
    return out


@pytest.mark.parametrize("name", list(GOLDEN_SHAPES))
def test_forward_backward_vs_reference_fp64(lib, name):
    g = Golden(name)
    feat, logits, gb, gp = g.inputs()
    out = _run(g.shape, feat, logits, g["M_ref"], g["t_ref"], gb, gp)
    ds = g.dstride
    assert out["bev"].dtype == torch.float32 and tuple(out["bev"].shape) == g["bev_ref64"].shape
    e = relerr(out["bev"], g["bev_ref64"])
    assert e <= FP32_TOL, e
    assert e <= float(g["ref32_vs_ref64_bev"]), "must be at least as close to fp64 as the reference's own fp32 run"
    # voxels nobody hits are exactly zero, and only those
    assert np.array_equal(out["bev"].cpu().numpy() == 0, g["bev_ref64"] == 0)
    assert relerr(out["prob"][:, ::ds], g["prob_ref"]) <= 1e-6
    assert relerr(out["grad_feat"], g["grad_feat_ref64"]) <= FP32_TOL
    assert relerr(out["grad_logits"][:, ::ds], g["grad_logits_ref64"]) <= FP32_TOL


def test_full_channels_vs_oracle(lib):
    """C=64 (config/training.yaml:22), jittered rig, B=2, against the numpy oracle
    (exact fp64 segment sums) - forward and backward."""
    from oracle import lift_splat_oracle as lo
    shape = LiftSplatShape(batch=2, channels=64)
    intr, extr = make_rig(2, 4, jitter=True, seed=21)
    feat, logits = make_encoder_outputs(shape, seed=5)
    gb, gp = make_upstream_grads(shape, seed=5)
    M, t = lo.camera_transform(intr.numpy(), extr.numpy())
    res, start, dim = grid_of(shape)
    _, _, rank = lo.voxel_index(lo.geometry(M, t, frustum_of(shape)), start, res, dim)
    bev_o, prob_o = lo.splat_forward(feat.numpy(), logits.numpy(), rank, dim, shape.cams)
    gf_o, gl_o = lo.splat_backward(feat.numpy(), logits.numpy(), rank, dim, shape.cams, gb.numpy(), gp.numpy())
    out = _run(shape, feat, logits, M, t, gb, gp)
    assert relerr(out["bev"], bev_o) <= FP32_TOL
    assert relerr(out["prob"], prob_o) <= 1e-6
    assert relerr(out["grad_feat"], gf_o) <= FP32_TOL
    assert relerr(out["grad_logits"], gl_o) <= FP32_TOL


def test_zero_mean_inputs_tight(lib):
    """Zero-mean variant: cancellation-heavy sums, still <= 1e-5."""
    from oracle import lift_splat_oracle as lo
    shape = LiftSplatShape(batch=1, channels=8)
    intr, extr = make_rig(1, 4, jitter=True, seed=33)
    feat, logits = make_encoder_outputs(shape, seed=6, relu=False)
    M, t = lo.camera_transform(intr.numpy(), extr.numpy())
    res, start, dim = grid_of(shape)
    _, _, rank = lo.voxel_index(lo.geometry(M, t, frustum_of(shape)), start, res, dim)
    bev_o, _ = lo.splat_forward(feat.numpy(), logits.numpy(), rank, dim, shape.cams)
    out = _run(shape, feat, logits, M, t)
    assert relerr(out["bev"], bev_o) <= FP32_TOL


def test_bf16_vs_reference_fp64(lib):
    g = Golden("rigB_b2_c4")
    feat, logits, gb, gp = g.inputs()
    out = _run(g.shape, feat, logits, g["M_ref"], g["t_ref"], gb, gp, dtype=torch.bfloat16)
    ds = g.dstride
    assert out["bev"].dtype == torch.float32          # reference output is always fp32 (bev_model.py:76)
    assert out["prob"].dtype == torch.bfloat16
    assert relerr(out["bev"], g["bev_ref64"]) <= BF16_TOL
    assert relerr(out["prob"][:, ::ds], g["prob_ref"]) <= BF16_TOL
    assert relerr(out["grad_feat"], g["grad_feat_ref64"]) <= BF16_TOL
    assert relerr(out["grad_logits"][:, ::ds], g["grad_logits_ref64"]) <= BF16_TOL


def test_deterministic_bitwise(lib):
    """Two runs give bit-identical BEV features and gradients although the counting sort
    places points with atomics (the splat sums each cell in ascending point id)."""
    shape = LiftSplatShape(batch=3, channels=16)
    intr, extr = make_rig(3, 4, jitter=True, seed=9)
    feat, logits = make_encoder_outputs(shape, seed=7)
    gb, gp = make_upstream_grads(shape, seed=7)
    ls = _ls()
    M, t = ls.camera_transform(intr.to(DEV), extr.to(DEV))
    a = _run(shape, feat, logits, M, t, gb, gp)
    for _ in range(3):
        b = _run(shape, feat, logits, M, t, gb, gp)
        for k in a:
            assert torch.equal(a[k], b[k]), k


# ------------------------------------------------------------------------------------
# edge cases
# ------------------------------------------------------------------------------------
def test_coarse_grid_long_segments(lib):
    """1 m voxels: hundreds to thousands of points per cell - exercises the >32-point
    canonicalisation path of the splat."""
    from oracle import lift_splat_oracle as lo
    shape = LiftSplatShape(batch=1, channels=4, bev_x_bound=[-10.0, 10.0, 1.0], bev_y_bound=[-10.0, 10.0, 1.0])
    intr, extr = make_rig(1, 4, jitter=True, seed=2)
    feat, logits = make_encoder_outputs(shape, seed=8)
    gb, gp = make_upstream_grads(shape, seed=8)
    M, t = lo.camera_transform(intr.numpy(), extr.numpy())
    res, start, dim = grid_of(shape)
    _, _, rank = lo.voxel_index(lo.geometry(M, t, frustum_of(shape)), start, res, dim)
    counts = np.bincount(rank[0][rank[0] >= 0])
    assert counts.max() > 1000
    bev_o, _ = lo.splat_forward(feat.numpy(), logits.numpy(), rank, dim, shape.cams)
    gf_o, gl_o = lo.splat_backward(feat.numpy(), logits.numpy(), rank, dim, shape.cams, gb.numpy(), gp.numpy())
    a = _run(shape, feat, logits, M, t, gb, gp)
    b = _run(shape, feat, logits, M, t, gb, gp)
    assert relerr(a["bev"], bev_o) <= FP32_TOL
    assert relerr(a["grad_feat"], gf_o) <= FP32_TOL
    assert relerr(a["grad_logits"], gl_o) <= FP32_TOL
    assert torch.equal(a["bev"], b["bev"])


def test_all_points_dropped_and_one_camera_dropped(lib):
    """A rig looking away from the grid: every point masked out -> BEV exactly zero, zero
    feature gradient; with one valid camera only that camera contributes."""
    from oracle import lift_splat_oracle as lo
    shape = LiftSplatShape(batch=2, channels=4)
    intr, extr = make_rig(2, 4, jitter=False)
    extr = extr.clone()
    extr[0, :, :3, 3] += 1000.0           # sample 0: all cameras far away
    extr[1, 1:, :3, 3] += 1000.0          # sample 1: only the front camera stays
    feat, logits = make_encoder_outputs(shape, seed=9)
    gb, gp = make_upstream_grads(shape, seed=9)
    M, t = lo.camera_transform(intr.numpy(), extr.numpy())
    res, start, dim = grid_of(shape)
    _, keep, rank = lo.voxel_index(lo.geometry(M, t, frustum_of(shape)), start, res, dim)
    assert keep[0].sum() == 0 and keep[1].reshape(4, -1)[1:].sum() == 0 and keep[1].sum() > 0
    bev_o, _ = lo.splat_forward(feat.numpy(), logits.numpy(), rank, dim, shape.cams)
    gf_o, gl_o = lo.splat_backward(feat.numpy(), logits.numpy(), rank, dim, shape.cams, gb.numpy(), gp.numpy())
    out = _run(shape, feat, logits, M, t, gb, gp)
    assert float(out["bev"][0].abs().max()) == 0.0
    assert float(out["grad_feat"][:4].abs().max()) == 0.0
    assert relerr(out["bev"], bev_o) <= FP32_TOL
    assert relerr(out["grad_feat"], gf_o) <= FP32_TOL
    assert relerr(out["grad_logits"], gl_o) <= FP32_TOL


def test_nan_extrinsics_drop_points(lib):
    """NaN coordinates convert to INT64_MIN on x86 and are masked out (bev_model.py:86-90);
    the kernel must drop them too instead of keeping voxel 0."""
    ls = _ls()
    shape = LiftSplatShape(batch=1, channels=2)
    s = _ls_shape(shape)
    M = torch.full((1, 4, 3, 3), float("nan"), device=DEV)
    t = torch.zeros(1, 4, 3, device=DEV)
    rank = ls.index(M, t, _dev(frustum_of(shape)), s)
    assert int((rank >= 0).sum()) == 0
    vox, keep, _ = ls.export_indices(M, t, _dev(frustum_of(shape)), s)
    assert int(keep.sum()) == 0 and int(vox.max()) == np.iinfo(np.int64).min


def test_odd_sizes(lib):
    """Grid that is not a multiple of the tile (and whose rows are not 16-byte multiples), non-square feature map, D not a
    multiple of anything, C=6."""
    from oracle import lift_splat_oracle as lo
    shape = LiftSplatShape(batch=2, cams=3, channels=6, bev_x_bound=[-7.0, 6.0, 0.2], bev_y_bound=[-5.0, 9.0, 0.2],
                           d_bound=[1.0, 9.0, 0.7], final_dim=[160, 224], bev_down_sample=8)
    intr, extr = make_rig(2, 3, jitter=True, seed=17)
    feat, logits = make_encoder_outputs(shape, seed=10)
    gb, gp = make_upstream_grads(shape, seed=10)
    gb = gb[:, :, :int(grid_of(shape)[2][0]), :int(grid_of(shape)[2][1])].contiguous()
    M, t = lo.camera_transform(intr.numpy(), extr.numpy())
    res, start, dim = grid_of(shape)
    assert gb.shape[2:] == (int(dim[0]), int(dim[1]))
    _, _, rank = lo.voxel_index(lo.geometry(M, t, frustum_of(shape)), start, res, dim)
    bev_o, prob_o = lo.splat_forward(feat.numpy(), logits.numpy(), rank, dim, shape.cams)
    gf_o, gl_o = lo.splat_backward(feat.numpy(), logits.numpy(), rank, dim, shape.cams, gb.numpy(), gp.numpy())
    ls = _ls()
    r = ls.index(_dev(M), _dev(t), _dev(frustum_of(shape)), _ls_shape(shape)).cpu().numpy()
    assert np.array_equal(r, rank.astype(np.int32))
    out = _run(shape, feat, logits, M, t, gb, gp)
    assert relerr(out["bev"], bev_o) <= FP32_TOL
    assert relerr(out["prob"], prob_o) <= 1e-6
    assert relerr(out["grad_feat"], gf_o) <= FP32_TOL
    assert relerr(out["grad_logits"], gl_o) <= FP32_TOL


def test_strided_grad_bev(lib):
    """torch.cat's backward hands a channel-slice view of a [B,65,X,Y] tensor
    (model/parking_model.py:45); the ABI takes strides instead of forcing a copy."""
    shape = LiftSplatShape(batch=2, channels=4)
    intr, extr = make_rig(2, 4, jitter=True, seed=3)
    feat, logits = make_encoder_outputs(shape, seed=11)
    gb, gp = make_upstream_grads(shape, seed=11)
    ls = _ls()
    M, t = ls.camera_transform(intr.to(DEV), extr.to(DEV))
    a = _run(shape, feat, logits, M, t, gb, gp)
    f = feat.to(DEV).requires_grad_(True)
    lg = logits.to(DEV).requires_grad_(True)
    bev, prob = ls.lift_splat(f, lg, M, t, _dev(frustum_of(shape)), _grid_spec(shape))
    wide = torch.cat([bev, torch.zeros(2, 1, 200, 200, device=DEV)], dim=1)
    gwide = torch.cat([gb.to(DEV), torch.ones(2, 1, 200, 200, device=DEV)], dim=1)
    ((wide * gwide).sum() + (prob * gp.to(DEV)).sum()).backward()
    assert torch.equal(f.grad, a["grad_feat"]) and torch.equal(lg.grad, a["grad_logits"])


# ------------------------------------------------------------------------------------
# properties at full size (BASELINE.json configs[1]: B=16, 4 cams, D=48, C=64, 200x200)
# ------------------------------------------------------------------------------------
def test_full_size_properties(lib):
    shape = LiftSplatShape(batch=16, channels=64)
    intr, extr = make_rig(16, 4, jitter=True, seed=1)
    feat, logits = make_encoder_outputs(shape, seed=12)
    ls = _ls()
    M, t = ls.camera_transform(intr.to(DEV), extr.to(DEV))
    fr, grid = _dev(frustum_of(shape)), _grid_spec(shape)
    f, lg = feat.to(DEV), logits.to(DEV)
    bev, prob = ls.lift_splat(f, lg, M, t, fr, grid)
    assert tuple(bev.shape) == (16, 64, 200, 200) and tuple(prob.shape) == (64, 48, 32, 32)
    # (1) probabilities sum to one
    assert float((prob.sum(1) - 1).abs().max()) < 1e-5
    # (2) mass conservation (checksum of checksums): sum over voxels of every channel equals
    #     sum over kept points of prob * feat, computed independently with torch ops
    rank = ls.index(M, t, fr, _ls_shape(shape))
    keep = (rank >= 0).view(16, 4, 48, 32 * 32).double()
    w = (prob.view(16, 4, 48, 32 * 32).double() * keep).sum(2)                   # [B,N,HW]
    expect = torch.einsum("bnp,bncp->bc", w, f.view(16, 4, 64, 32 * 32).double())
    got = bev.double().sum((2, 3))
    assert float(((got - expect).abs() / expect.abs().clamp_min(1e-6)).max()) < 1e-6
    # (3) linearity in the features
    bev2, _ = ls.lift_splat(2.0 * f, lg, M, t, fr, grid)
    assert torch.equal(bev2, 2.0 * bev)
    # (4) per-sample independence: sample 5 alone gives the same BEV slice
    sl = slice(5 * 4, 6 * 4)
    bev5, _ = ls.lift_splat(f[sl], lg[sl], M[5:6], t[5:6], fr, grid)
    assert torch.equal(bev5[0], bev[5])


def test_bev_model_drop_in(lib):
    """BevModel keeps the reference's call contract: forward(images, intrinsics,
    extrinsics) -> (bev[B,C,X,Y] fp32, pred_depth[B*N,D,h,w]); grads reach the encoder."""
    from e2e_parking_carla_b200 import BevModel

    class TinyEncoder(torch.nn.Module):
        def __init__(self, c, d):
            super().__init__()
            self.f = torch.nn.Conv2d(3, c, 8, stride=8)
            self.d = torch.nn.Conv2d(3, d, 8, stride=8)

        def forward(self, x):
            return self.f(x).relu(), self.d(x).relu()

    shape = LiftSplatShape(batch=2, channels=8)
    cfg = make_cfg(shape)
    model = BevModel(cfg, cam_encoder=TinyEncoder(8, 48)).to(DEV)
    assert model.bev_dim.dtype == torch.int64 and model.frustum.shape == (48, 32, 32, 3)
    intr, extr = make_rig(2, 4, jitter=True, seed=6)
    images = torch.randn(2, 4, 3, 256, 256, device=DEV)
    bev, depth = model(images, intr.to(DEV), extr.to(DEV))
    assert tuple(bev.shape) == (2, 8, 200, 200) and bev.dtype == torch.float32
    assert tuple(depth.shape) == (8, 48, 32, 32)
    (bev.sum() + depth.square().sum()).backward()
    assert model.cam_encoder.f.weight.grad is not None and model.cam_encoder.d.weight.grad.abs().sum() > 0
    # geometry="torch" (the reference's own torch.inverse calls) agrees up to index flips
    model_t = BevModel(cfg, cam_encoder=model.cam_encoder, geometry="torch").to(DEV)
    bev_t, _ = model_t(images, intr.to(DEV), extr.to(DEV))
    assert relerr(bev_t, bev) < 1e-2
    geom = model.get_geometry(intr.to(DEV), extr.to(DEV))
    assert tuple(geom.shape) == (2, 4, 48, 32, 32, 3)


def test_bench_paths_agree(lib):
    """The three ways bench.py drives the library - kernel-by-kernel launches, the captured
    CUDA graph, and the host-buffer pipeline (per-group workspaces, forward of every group
    before the backwards) - produce bit-identical outputs."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench

    st = bench.Stepper(LiftSplatShape(batch=5, channels=16), torch.float32, DEV)
    outs = ("bev", "prob", "gfeat", "glogits")
    st.step()
    torch.cuda.synchronize()
    ref = {k: getattr(st, k).clone() for k in outs}
    assert ref["bev"].abs().sum() > 0 and ref["gfeat"].abs().sum() > 0
    graph = st.capture()
    for k in outs:
        getattr(st, k).zero_()
    graph.replay()
    torch.cuda.synchronize()
    for k in outs:
        assert torch.equal(getattr(st, k), ref[k]), k
    assert st.graph_launches >= 10
    for k in outs:
        getattr(st, k).zero_()
    st.step_e2e(chunks=3)       # ragged groups: 2 + 2 + 1 samples
    for k in outs:
        assert torch.equal(st.out_host[k], ref[k].cpu()), k
    st.capture_e2e(chunks=3)    # the same pipeline as one graph, copies included
    for k in outs:
        st.out_host[k].zero_()
    st.step_e2e(chunks=3)
    for k in outs:
        assert torch.equal(st.out_host[k], ref[k].cpu()), k


def test_plain_stream_order_matches(lib):
    """LS_NO_PDL=1 (no programmatic dependent launches) and LS_NO_SIDE_STREAMS=1 (everything on
    the caller's stream) are read once per process, so they are exercised in a child process:
    the step must give the same bits as the default configuration."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, torch, hashlib; sys.path.insert(0, %r); import bench\n"
        "from e2e_parking_carla_b200.synthetic import LiftSplatShape\n"
        "st = bench.Stepper(LiftSplatShape(batch=3, channels=16), torch.float32, torch.device('cuda:0'))\n"
        "st.step(); torch.cuda.synchronize()\n"
        "h = hashlib.sha256()\n"
        "for k in ('bev', 'prob', 'gfeat', 'glogits'): h.update(getattr(st, k).cpu().numpy().tobytes())\n"
        "print('DIGEST', h.hexdigest())\n" % root)
    digests = []
    for extra in ({}, {"LS_NO_PDL": "1", "LS_NO_SIDE_STREAMS": "1"}):
        env = dict(os.environ, **extra)
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        digests.append([l for l in out.stdout.splitlines() if l.startswith("DIGEST")][0])
    assert digests[0] == digests[1]


def test_many_tiles_single_cta_scan(lib):
    """A long, narrow grid with more than 2048 tiles takes the single-CTA scan (identity tile
    order) instead of the two-kernel parallel scan; everything downstream must not care."""
    from oracle import lift_splat_oracle as lo
    shape = LiftSplatShape(batch=1, cams=2, channels=8, bev_x_bound=[-105.0, 105.0, 0.1], bev_y_bound=[-4.0, 4.0, 0.1],
                           d_bound=[0.5, 60.5, 2.5], final_dim=[128, 128], bev_down_sample=8)
    ls = _ls()
    tiles = ls.grid_cells(_ls_shape(shape))[0]
    assert tiles > 2048
    intr, extr = make_rig(1, 2, jitter=True, seed=23)
    feat, logits = make_encoder_outputs(shape, seed=12)
    gb, gp = make_upstream_grads(shape, seed=12)
    res, start, dim = grid_of(shape)
    gb = gb[:, :, :int(dim[0]), :int(dim[1])].contiguous()
    M, t = lo.camera_transform(intr.numpy(), extr.numpy())
    _, _, rank = lo.voxel_index(lo.geometry(M, t, frustum_of(shape)), start, res, dim)
    assert (rank >= 0).sum() > 1000
    bev_o, prob_o = lo.splat_forward(feat.numpy(), logits.numpy(), rank, dim, shape.cams)
    gf_o, gl_o = lo.splat_backward(feat.numpy(), logits.numpy(), rank, dim, shape.cams, gb.numpy(), gp.numpy())
    r = ls.index(_dev(M), _dev(t), _dev(frustum_of(shape)), _ls_shape(shape)).cpu().numpy()
    assert np.array_equal(r, rank.astype(np.int32))
    out = _run(shape, feat, logits, M, t, gb, gp)
    assert relerr(out["bev"], bev_o) <= FP32_TOL
    assert relerr(out["grad_feat"], gf_o) <= FP32_TOL
    assert relerr(out["grad_logits"], gl_o) <= FP32_TOL
