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

from _util import GOLDEN_SHAPES, Golden, assert_close, frustum_of, grid_of, maxerr, relerr, sha
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


def _ls_shape(shape: LiftSplatShape, channels=None, tile_x=1):
    ls = _ls()
    return ls.make_shape(shape.batch, shape.cams, shape.depth_bins, shape.fh, shape.fw,
                         channels or shape.channels, _grid_spec(shape), 0, tile_x)


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


@pytest.mark.parametrize("tile_x", [1, 8, 32])
@pytest.mark.parametrize("name", list(GOLDEN_SHAPES))
def test_sorted_ranks_bit_exact(lib, name, tile_x):
    """ranks[ranks.argsort()] (model/bev_model.py:96-97) and the segment count of
    VoxelsSumming (tool/geometry.py:295-296), from the CSR the counting sort builds - for the
    1 x 128 strips of the NCHW path and the square tiles of the channels-last splat."""
    ls = _ls()
    g = Golden(name)
    s = _ls_shape(g.shape, tile_x=tile_x)
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
    # pixel-major index: {rank of the cell (X*Y when dropped), prob} of every depth bin of every pixel
    xy = int(s.X) * int(s.Y)
    pc = pix[..., 0].view(sh.batch, sh.cams, hw, D).permute(0, 1, 3, 2).reshape(sh.batch, -1)
    assert torch.equal(pc, torch.where(rank >= 0, rank, torch.full_like(rank, xy)))   # dropped -> one past the grid
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
def _run(shape, feat, logits, M, t, gb=None, gp=None, dtype=torch.float32, bev_format=torch.contiguous_format,
         feat_format=torch.contiguous_format):
    """One forward (+ backward) through the autograd function.  ``bev_format`` is the memory
    format of the BEV output AND of the upstream gradient handed to backward (torch.autograd
    delivers exactly the tensor given to torch.autograd.backward)."""
    ls = _ls()
    feat = feat.to(DEV, dtype).contiguous(memory_format=feat_format).requires_grad_(gb is not None)
    logits = logits.to(DEV, dtype).requires_grad_(gb is not None)
    bev, prob = ls.lift_splat(feat, logits, _dev(M), _dev(t), _dev(frustum_of(shape)), _grid_spec(shape), bev_format)
    assert bev.is_contiguous(memory_format=bev_format)
    out = {"bev": bev.detach(), "prob": prob.detach()}
    if gb is not None:
        g_bev = gb.to(DEV).contiguous(memory_format=bev_format)
        torch.autograd.backward([bev, prob], [g_bev, gp.to(DEV, prob.dtype)])
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
    # ... and LS_SPLAT_OUT=bulk: the shared-memory-tile splat that leaves as one bulk (TMA) store per tile
    # instead of the default direct row stores
    # ... LS_OVERLAP_BWD=1: the backward's epilogue launched as the gather's programmatic dependent, waiting per
    # image on the gather's arrival counters instead of on the whole grid; LS_SOFTMAX_BWD_STAGED=1: the
    # shared-memory-staged softmax backward instead of the thread-per-pixel one (same summation order)
    for extra in ({}, {"LS_NO_PDL": "1", "LS_NO_SIDE_STREAMS": "1"}, {"LS_SPLAT_OUT": "bulk"}, {"LS_OVERLAP_BWD": "1"},
                  {"LS_SOFTMAX_BWD_STAGED": "1"}):
        env = dict(os.environ, **extra)
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        digests.append([l for l in out.stdout.splitlines() if l.startswith("DIGEST")][0])
    assert digests[0] == digests[1]
    assert digests[0] == digests[3], "overlapped epilogue"
    assert digests[0] == digests[4], "staged softmax backward / layout kernels"
    # the bulk (shared-memory tile) variant cuts its pieces inside cells: same gradients, BEV equal up to
    # the last bit - checked on values in test_bulk_store_variant_matches, here it only has to run
    assert len(digests[2]) > 10


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


# ------------------------------------------------------------------------------------
# round 2: layouts (channels_last BEV / gradient / features), real channel counts in bf16,
# C > 64, the stress grid at C = 64, full-size values, max element-wise bounds
# ------------------------------------------------------------------------------------
def _oracle_case(shape, rig_seed, in_seed, relu=True):
    """Inputs + numpy-oracle outputs (exact float64 segment sums) for one configuration."""
    from oracle import lift_splat_oracle as lo
    intr, extr = make_rig(shape.batch, shape.cams, jitter=True, seed=rig_seed)
    feat, logits = make_encoder_outputs(shape, seed=in_seed, relu=relu)
    gb, gp = make_upstream_grads(shape, seed=in_seed)
    res, start, dim = grid_of(shape)
    gb = gb[:, :, :int(dim[0]), :int(dim[1])].contiguous()
    M, t = lo.camera_transform(intr.numpy(), extr.numpy())
    _, _, rank = lo.voxel_index(lo.geometry(M, t, frustum_of(shape)), start, res, dim)
    return {"feat": feat, "logits": logits, "gb": gb, "gp": gp, "M": M, "t": t, "rank": rank, "dim": dim}


def _oracle_outputs(shape, c, dtype=torch.float32):
    """The oracle on the values the kernels actually see (inputs rounded to ``dtype`` first)."""
    from oracle import lift_splat_oracle as lo
    f = c["feat"].to(dtype).float().numpy()
    z = c["logits"].to(dtype).float().numpy()
    gp = c["gp"].to(dtype).float().numpy()
    bev_o, prob_o = lo.splat_forward(f, z, c["rank"], c["dim"], shape.cams)
    gf_o, gl_o = lo.splat_backward(f, z, c["rank"], c["dim"], shape.cams, c["gb"].numpy(), gp)
    return {"bev": bev_o, "prob": prob_o, "grad_feat": gf_o, "grad_logits": gl_o}


def _check_all(out, ref, tol, prob_tol=None):
    assert_close(out["bev"], ref["bev"], tol, "bev")
    assert_close(out["prob"], ref["prob"], prob_tol or tol, "prob")
    assert_close(out["grad_feat"], ref["grad_feat"], tol, "grad_feat")
    assert_close(out["grad_logits"], ref["grad_logits"], tol, "grad_logits")
    assert np.array_equal(out["bev"].cpu().numpy() == 0, ref["bev"] == 0) or tol > 1e-4, "zero pattern"


def _same_across_layouts(a, b):
    """Gradients and probabilities: the same bits in every layout (the backward's arithmetic does not
    depend on it).  BEV features: every cell is summed in the same canonical record order, but the
    NCHW tile kernel cuts its quarter-warp pieces inside cells (partial sums merged afterwards)
    while the channels-last kernel sums each cell front to back - last-bit differences allowed."""
    for k in ("prob", "grad_feat", "grad_logits"):
        if k in a:
            assert torch.equal(a[k], b[k]), k
    assert maxerr(a["bev"], b["bev"]) <= 1e-6
    assert torch.equal(a["bev"] == 0, b["bev"] == 0)


@pytest.mark.parametrize("channels", [64, 16, 6])
def test_channels_last_bev_bit_identical_to_nchw(lib, channels):
    """channels_last BEV output + channels_last gradient (the splat's native layout: bulk tile
    stores forward, gradient rows gathered in place backward) give the SAME BITS as the
    reference's NCHW layout.  C=64: bulk-store kernel + 16-byte row gathers; C=16: row-store
    kernel; C=6: channels_last output, gradient falls back to the staged path."""
    shape = LiftSplatShape(batch=2, channels=channels)
    c = _oracle_case(shape, rig_seed=41, in_seed=13)
    a = _run(shape, c["feat"], c["logits"], c["M"], c["t"], c["gb"], c["gp"])
    b = _run(shape, c["feat"], c["logits"], c["M"], c["t"], c["gb"], c["gp"], bev_format=torch.channels_last)
    assert b["bev"].stride(1) == 1 and a["bev"].stride(3) == 1
    _same_across_layouts(a, b)
    _check_all(b, _oracle_outputs(shape, c), FP32_TOL, 1e-6)


def test_channels_last_gradient_slice_of_65_channels(lib):
    """Downstream of a channels_last torch.cat with the target channel (model/parking_model.py:45)
    the gradient is a 64-channel slice of a [B,X,Y,65] tensor: rows 260 bytes apart, not 16-byte
    aligned -> the scalar in-place gather.  Also the forward writing INTO such a slice."""
    ls = _ls()
    shape = LiftSplatShape(batch=2, channels=64)
    c = _oracle_case(shape, rig_seed=42, in_seed=14)
    a = _run(shape, c["feat"], c["logits"], c["M"], c["t"], c["gb"], c["gp"])
    f = c["feat"].to(DEV).requires_grad_(True)
    z = c["logits"].to(DEV).requires_grad_(True)
    bev, prob = ls.lift_splat(f, z, _dev(c["M"]), _dev(c["t"]), _dev(frustum_of(shape)), _grid_spec(shape),
                              torch.channels_last)
    wide = torch.cat([bev, torch.zeros(2, 200, 200, 1, device=DEV).permute(0, 3, 1, 2)], dim=1)
    assert wide.is_contiguous(memory_format=torch.channels_last)
    gwide = torch.cat([c["gb"].to(DEV), torch.ones(2, 1, 200, 200, device=DEV)], dim=1).contiguous(
        memory_format=torch.channels_last)
    torch.autograd.backward([wide, prob], [gwide, c["gp"].to(DEV)])
    assert torch.equal(f.grad, a["grad_feat"]) and torch.equal(z.grad, a["grad_logits"])
    # forward straight into channels 0..63 of a 65-channel channels_last buffer (row-store kernel)
    buf = torch.full((2, 65, 200, 200), -7.0, device=DEV).contiguous(memory_format=torch.channels_last)
    s = _ls_shape(shape)
    prob2 = torch.empty_like(prob)
    scratch = ls.scratch_for(ls.scratch_bytes(s, ls.LS_F32, False), torch.device(DEV))
    view = buf[:, :64]
    st = ls._bev_strides(view)
    assert (st.c, st.y) == (1, 65)
    import ctypes as C
    from e2e_parking_carla_b200 import _lib
    P = lambda x: C.c_void_p(x.data_ptr())
    fr = _dev(frustum_of(shape))
    Md, td = _dev(c["M"]), _dev(c["t"])
    ls.check(_lib.load().ls_forward(P(f.detach()), ls.LS_FEAT_NCHW, P(z.detach()), ls.LS_F32, P(Md), P(td), P(fr),
                                    C.byref(s), P(scratch), scratch.numel(), None, 0, P(view), C.byref(st), P(prob2),
                                    C.c_void_p(torch.cuda.current_stream().cuda_stream)), "ls_forward")
    assert maxerr(buf[:, :64], a["bev"]) <= 1e-6 and bool((buf[:, 64] == -7.0).all())


def test_channels_last_features_in_place(lib):
    """A channels_last feature map (what a channels_last CamEncoder emits) is consumed without the
    NHWC staging copy and its gradient comes back channels_last - same bits as the NCHW route."""
    shape = LiftSplatShape(batch=2, channels=64)
    c = _oracle_case(shape, rig_seed=43, in_seed=15)
    for dtype in (torch.float32, torch.bfloat16):
        a = _run(shape, c["feat"], c["logits"], c["M"], c["t"], c["gb"], c["gp"], dtype=dtype)
        b = _run(shape, c["feat"], c["logits"], c["M"], c["t"], c["gb"], c["gp"], dtype=dtype,
                 bev_format=torch.channels_last, feat_format=torch.channels_last)
        assert b["grad_feat"].is_contiguous(memory_format=torch.channels_last)
        _same_across_layouts(a, b)


@pytest.mark.parametrize("bev_format", [torch.contiguous_format, torch.channels_last])
def test_bf16_real_channel_count_vs_oracle(lib, bev_format):
    """bf16 at C=64 on the jittered rig, B=2, forward + backward <= 2e-2 (north_star) against the
    float64 oracle - the kernels bench.py --dtype bf16 times (ls_splat_fwd_kernel<bf16,...,64>,
    every lane of ls_bwd_gather_occ_kernel<bf16>)."""
    shape = LiftSplatShape(batch=2, channels=64)
    c = _oracle_case(shape, rig_seed=44, in_seed=16)
    out = _run(shape, c["feat"], c["logits"], c["M"], c["t"], c["gb"], c["gp"], dtype=torch.bfloat16,
               bev_format=bev_format)
    assert out["bev"].dtype == torch.float32 and out["grad_feat"].dtype == torch.bfloat16
    # against the oracle on the original float32 inputs: the north_star bound
    _check_all(out, _oracle_outputs(shape, c), BF16_TOL)
    # against the oracle on the bf16-rounded inputs the error is the kernels' own: prob is
    # rounded to bf16 once (2^-9), everything is accumulated in float32
    ref = _oracle_outputs(shape, c, torch.bfloat16)
    assert relerr(out["bev"], ref["bev"]) <= 4e-3


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_stress_grid_real_channel_count_vs_oracle(lib, dtype):
    """BASELINE.json configs[3] geometry (6 cameras, 96 depth bins, 400x400 at 0.05 m) at C=64,
    B=2, against the numpy oracle: placement with 6 bins per thread, 1 600 tiles, 590 k points."""
    from oracle import lift_splat_oracle as lo
    shape = LiftSplatShape.stress(batch=2)
    assert shape.channels == 64 and shape.cams == 6 and shape.depth_bins == 96
    c = _oracle_case(shape, rig_seed=45, in_seed=17)
    ls = _ls()
    r = ls.index(_dev(c["M"]), _dev(c["t"]), _dev(frustum_of(shape)), _ls_shape(shape)).cpu().numpy()
    assert np.array_equal(r, c["rank"].astype(np.int32))
    out = _run(shape, c["feat"], c["logits"], c["M"], c["t"], c["gb"], c["gp"], dtype=dtype,
               bev_format=torch.channels_last)
    tol = FP32_TOL if dtype == torch.float32 else BF16_TOL
    _check_all(out, _oracle_outputs(shape, c), tol, 1e-6 if dtype == torch.float32 else None)
    if dtype == torch.float32:
        nchw = _run(shape, c["feat"], c["logits"], c["M"], c["t"], c["gb"], c["gp"])
        _same_across_layouts(nchw, out)


@pytest.mark.parametrize("channels,dtype", [(96, torch.float32), (128, torch.float32), (256, torch.float32),
                                            (130, torch.float32), (128, torch.bfloat16)])
def test_more_than_64_channels(lib, channels, dtype):
    """C > 64: the splat's channel-chunk loop (tile re-zeroed per 64-channel pass) and the
    2-/3-/4-chunk gradient gathers, in both BEV layouts."""
    shape = LiftSplatShape(batch=1, channels=channels)
    c = _oracle_case(shape, rig_seed=46, in_seed=18)
    ref = _oracle_outputs(shape, c)
    tol = FP32_TOL if dtype == torch.float32 else BF16_TOL
    a = _run(shape, c["feat"], c["logits"], c["M"], c["t"], c["gb"], c["gp"], dtype=dtype)
    _check_all(a, ref, tol, 1e-6 if dtype == torch.float32 else None)
    b = _run(shape, c["feat"], c["logits"], c["M"], c["t"], c["gb"], c["gp"], dtype=dtype,
             bev_format=torch.channels_last)
    _same_across_layouts(a, b)


def test_more_than_256_channels_rejected(lib):
    ls = _ls()
    with pytest.raises(ValueError):
        ls.scratch_bytes(_ls_shape(LiftSplatShape(batch=1, channels=260)), ls.LS_F32, True)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_many_depth_bins_serial_softmax(lib, dtype):
    """D = 130 > 128: the one-thread-per-pixel softmax forward/backward kernels, placement with
    9 bins per thread, D not a multiple of 16 in the gather."""
    shape = LiftSplatShape(batch=1, cams=2, channels=8, d_bound=[0.5, 13.5, 0.1], final_dim=[64, 96])
    assert shape.depth_bins == 130
    c = _oracle_case(shape, rig_seed=47, in_seed=19)
    tol = FP32_TOL if dtype == torch.float32 else BF16_TOL
    out = _run(shape, c["feat"], c["logits"], c["M"], c["t"], c["gb"], c["gp"], dtype=dtype)
    _check_all(out, _oracle_outputs(shape, c), tol, 1e-6 if dtype == torch.float32 else None)


@pytest.mark.parametrize("d_bound, bins", [([0.5, 8.5, 0.25], 32), ([0.5, 16.5, 0.25], 64)])
@pytest.mark.parametrize("feat_format", [torch.contiguous_format, torch.channels_last])
def test_pixel_stationary_epilogue_depth_counts(lib, d_bound, bins, feat_format):
    """The thread-per-pixel backward epilogue (softmax backward + NHWC->NCHW of grad_feat in one launch) has
    instantiations for 32, 48 and 64 depth bins; 48 is what every default-shape test runs, this covers the other
    two, for NCHW features (both parts) and channels-last features (softmax part only), with and without an
    upstream gradient on prob."""
    shape = LiftSplatShape(batch=2, cams=3, channels=64, d_bound=d_bound)
    assert shape.depth_bins == bins
    c = _oracle_case(shape, rig_seed=53, in_seed=23)
    ref = _oracle_outputs(shape, c)
    out = _run(shape, c["feat"], c["logits"], c["M"], c["t"], c["gb"], c["gp"], bev_format=torch.channels_last,
               feat_format=feat_format)
    _check_all(out, ref, FP32_TOL, 1e-6)
    # no upstream gradient on prob (grad_prob_ext == NULL): the epilogue's gext branch is off
    ls = _ls()
    feat = c["feat"].to(DEV).contiguous(memory_format=feat_format).requires_grad_(True)
    logits = c["logits"].to(DEV).requires_grad_(True)
    bev, _ = ls.lift_splat(feat, logits, _dev(c["M"]), _dev(c["t"]), _dev(frustum_of(shape)), _grid_spec(shape),
                           torch.channels_last)
    bev.backward(c["gb"].to(DEV).contiguous(memory_format=torch.channels_last))
    ref0 = _oracle_outputs(shape, dict(c, gp=torch.zeros_like(c["gp"])))
    assert_close(logits.grad, ref0["grad_logits"], FP32_TOL, "grad_logits without upstream prob gradient")
    assert_close(feat.grad, ref0["grad_feat"], FP32_TOL, "grad_feat without upstream prob gradient")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_pixel_stationary_epilogue_ragged_feature_map(lib, dtype):
    """D = 48 (the thread-per-pixel epilogue) on a 25 x 33 feature map - 825 pixels per image, neither a multiple of the
    epilogue's 128-thread CTAs nor of a warp - with 6 channels padded to 8 (the layout part drops the padding)."""
    shape = LiftSplatShape(batch=2, cams=3, channels=6, final_dim=[200, 264])
    assert (shape.fh, shape.fw, shape.depth_bins) == (25, 33, 48)
    c = _oracle_case(shape, rig_seed=59, in_seed=29)
    tol = FP32_TOL if dtype == torch.float32 else BF16_TOL
    out = _run(shape, c["feat"], c["logits"], c["M"], c["t"], c["gb"], c["gp"], dtype=dtype,
               bev_format=torch.channels_last)
    _check_all(out, _oracle_outputs(shape, c, dtype), tol, 1e-6 if dtype == torch.float32 else None)


def test_two_host_threads_share_one_device(lib):
    """Two host threads drive the same device at the same time, each on its own stream (ctypes releases the GIL, so
    the enqueue sequences really interleave): the library-owned side streams and events are shared per device and
    guarded by a mutex for the duration of a call, the scratch blob is per stream - every iteration of both threads
    must give the bits of a single-threaded run."""
    import threading
    shape = LiftSplatShape(batch=2, channels=64)
    c = _oracle_case(shape, rig_seed=61, in_seed=31)
    args = (shape, c["feat"], c["logits"], c["M"], c["t"], c["gb"], c["gp"])
    ref = _run(*args, bev_format=torch.channels_last)
    torch.cuda.synchronize()
    results, errors = {}, []

    def worker(i):
        try:
            stream = torch.cuda.Stream()
            fmt = torch.channels_last if i == 0 else torch.contiguous_format     # different pipelines side by side
            with torch.cuda.stream(stream):
                for _ in range(6):
                    out = _run(*args, bev_format=fmt)
                    stream.synchronize()
                    for k in ("prob", "grad_feat", "grad_logits"):
                        assert torch.equal(out[k], ref[k]), (i, k)
                    assert maxerr(out["bev"], ref["bev"]) <= 1e-6
            results[i] = True
        except BaseException as exc:      # surfaced in the main thread below
            errors.append(repr(exc))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors
    assert results == {0: True, 1: True}


def test_full_size_values_on_sampled_batch_indices(lib):
    """BASELINE.json configs[1] at its full size (B=16, C=64): the BEV features and gradients of
    three batch indices against the oracle's values (the rest of the batch is covered by the
    properties of test_full_size_properties)."""
    shape = LiftSplatShape(batch=16, channels=64)
    c = _oracle_case(shape, rig_seed=1, in_seed=20)
    out = _run(shape, c["feat"], c["logits"], c["M"], c["t"], c["gb"], c["gp"], bev_format=torch.channels_last)
    n = shape.cams
    one = LiftSplatShape(batch=1, channels=64)
    for b in (0, 7, 15):
        sub = {"feat": c["feat"][b * n:(b + 1) * n], "logits": c["logits"][b * n:(b + 1) * n],
               "gb": c["gb"][b:b + 1], "gp": c["gp"][b * n:(b + 1) * n], "rank": c["rank"][b:b + 1], "dim": c["dim"]}
        ref = _oracle_outputs(one, sub)
        got = {"bev": out["bev"][b:b + 1], "prob": out["prob"][b * n:(b + 1) * n],
               "grad_feat": out["grad_feat"][b * n:(b + 1) * n], "grad_logits": out["grad_logits"][b * n:(b + 1) * n]}
        _check_all(got, ref, FP32_TOL, 1e-6)


# ------------------------------------------------------------------------------------
# DepthLoss kernels (loss/depth_loss.py:18-48)
# ------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_depth_loss_vs_golden_and_oracle(lib, dtype):
    """ls_depth_loss_fwd/bwd against the fixture frozen from the unmodified reference and
    against the float64 oracle: bin labels bit-exact, loss and gradient <= 1e-5 (fp32)."""
    from oracle import depth_loss_oracle as dlo
    from e2e_parking_carla_b200 import DepthLoss
    from e2e_parking_carla_b200.synthetic import make_depth_labels
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "depth_loss_b1.npz"))
    shape = LiftSplatShape(batch=1, channels=4)
    _, logits = make_encoder_outputs(shape, seed=31)
    gt = make_depth_labels(shape, seed=31)
    prob = logits.softmax(dim=1).to(DEV, dtype).requires_grad_(True)
    crit = DepthLoss(make_cfg(shape))
    assert crit.depth_channels == 48
    loss = crit(prob, gt.to(DEV))
    (3.0 * loss).backward()
    lo_loss, lo_grad, labels = dlo.depth_loss(prob.detach().float().cpu().numpy(), gt.numpy(), shape.d_bound, 8)
    assert np.array_equal(labels, z["labels"].astype(np.int64))
    tol = FP32_TOL if dtype == torch.float32 else BF16_TOL
    assert abs(loss.item() - lo_loss) <= tol * abs(lo_loss)
    assert_close(prob.grad.float() / 3.0, lo_grad, tol, "grad_prob")
    if dtype == torch.float32:
        assert abs(loss.item() - float(z["loss32"])) <= 1e-6 * abs(lo_loss)
        assert_close(prob.grad / 3.0, z["grad32"], FP32_TOL, "grad_prob vs reference")
        # labels as the kernel saw them (saved for backward): zero gradient exactly on background pixels
        bg = torch.from_numpy(labels == 0).view(4, 32, 32).to(DEV)
        assert float(prob.grad.permute(0, 2, 3, 1)[bg].abs().max()) == 0.0
    # deterministic
    prob2 = prob.detach().clone().requires_grad_(True)
    loss2 = crit(prob2, gt.to(DEV))
    assert torch.equal(loss2, loss)


def test_depth_loss_edge_cases(lib):
    """No foreground at all -> loss 0 (division by max(1, 0)); probabilities of exactly 0 and 1
    hit aten's log clamp (-100) and the backward's 1e-12 floor; odd down-sample / sizes."""
    from oracle import depth_loss_oracle as dlo
    from e2e_parking_carla_b200 import DepthLoss
    from types import SimpleNamespace
    cfg = SimpleNamespace(d_bound=[1.0, 9.0, 0.5], bev_down_sample=4)
    crit = DepthLoss(cfg)
    D = crit.depth_channels
    assert D == 16
    gen = torch.Generator().manual_seed(5)
    prob = torch.rand(6, D, 5, 7, generator=gen).softmax(dim=1)
    prob[0, 3] = 0.0
    prob[1, 2] = 1.0
    gt = torch.rand(2, 3, 20, 28, generator=gen) * 10.0
    gt[0, 0, :8] = 0.0
    p = prob.to(DEV).requires_grad_(True)
    loss = crit(p, gt.to(DEV))
    loss.backward()
    lo_loss, lo_grad, _ = dlo.depth_loss(prob.numpy(), gt.numpy(), cfg.d_bound, 4)
    assert abs(loss.item() - lo_loss) <= 1e-5 * abs(lo_loss)
    assert_close(p.grad, lo_grad, 1e-5, "grad")
    p0 = prob.to(DEV).requires_grad_(True)
    loss0 = crit(p0, torch.zeros(2, 3, 20, 28, device=DEV))
    loss0.backward()
    assert loss0.item() == 0.0 and float(p0.grad.abs().max()) == 0.0


# ------------------------------------------------------------------------------------
# add_target_bev (model/parking_model.py:28-46): the BEV's immediate consumer
# ------------------------------------------------------------------------------------
def test_add_target_bev_matches_reference_semantics(lib):
    """Same values as the reference's zero-map + python-slice stamp + torch.cat (including its
    torch.rand_like noise under the same seed and python's negative-slice wrap at the border),
    for an NCHW BEV, a channels_last BEV (cat stays channels_last) and the in-place variant
    (BevModel(spare_channels=1): no copy of the BEV); the gradient reaches the encoder outputs
    through the 65-channel tensor and is bit-identical in all three."""
    from oracle import torch_port as tp
    from e2e_parking_carla_b200 import BevModel, add_target_bev

    class Preset(torch.nn.Module):
        def forward(self, images):
            return self.feat, self.logits

    shape = LiftSplatShape(batch=4, channels=64)
    cfg = make_cfg(shape)
    intr, extr = make_rig(4, 4, jitter=True, seed=51)
    feat, logits = make_encoder_outputs(shape, seed=21)
    # targets: inside, near the low border (negative slice start -> python wraps: empty stamp), far corner
    target = torch.tensor([[1.3, -2.7, 0.0], [-9.9, 0.2, 0.0], [7.9, 7.9, 0.0], [-3.3, 9.7, 0.0]], device=DEV)
    images = torch.zeros(4, 4, 3, 8, 8, device=DEV)
    conv = torch.nn.Conv2d(65, 4, 3, padding=1).to(DEV)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False       # the consumer conv is not what is under test
    results = []
    for fmt, spare in ((torch.contiguous_format, 0), (torch.channels_last, 0), (torch.channels_last, 1)):
        enc = Preset()
        enc.feat = feat.to(DEV).requires_grad_(True)
        enc.logits = logits.to(DEV).requires_grad_(True)
        model = BevModel(cfg, cam_encoder=enc, bev_memory_format=fmt, spare_channels=spare).to(DEV)
        bev, depth = model(images, intr.to(DEV), extr.to(DEV))
        torch.manual_seed(77)
        wide, tmap = add_target_bev(bev, target, cfg)
        torch.manual_seed(77)
        wide_ref, tmap_ref = tp.add_target_bev_ref(bev.detach(), target, cfg.bev_x_bound[2], cfg.bev_y_bound[2])
        assert tuple(wide.shape) == (4, 65, 200, 200) and tuple(tmap.shape) == (4, 1, 200, 200)
        assert torch.equal(wide, wide_ref) and torch.equal(tmap, tmap_ref)
        assert float(tmap.sum()) > 0
        if fmt == torch.channels_last:
            assert wide.is_contiguous(memory_format=torch.channels_last)
        if spare:
            assert wide.data_ptr() == bev.data_ptr()          # nothing was copied
        conv.zero_grad()
        (conv(wide).square().sum() + depth.sum()).backward()
        results.append((bev.detach().clone(), enc.feat.grad.clone(), enc.logits.grad.clone()))
    torch.backends.cudnn.allow_tf32 = tf32
    for r in results[1:]:
        assert maxerr(r[0], results[0][0]) <= 1e-6
        # the gradient handed to the backward differs only in layout; conv's own backward picks
        # another algorithm per layout, so compare to tolerance, not bits
        assert relerr(r[1], results[0][1]) < 1e-4 and relerr(r[2], results[0][2]) < 1e-4


@pytest.mark.parametrize("name", ["inner_b16", "border_b12", "stress_b8", "ragged_b6"])
def test_target_bev_kernel_vs_reference_golden(lib, name):
    """ls_target_bev against maps frozen from the UNMODIFIED reference method (tests/golden/
    target_bev.npz): the noised pixels are drawn as the reference draws them (torch's CPU generator
    under the fixture's seed, the product's own target_pixels on CPU tensors), the stamp runs on the GPU
    into an NCHW map, a channels-last map and channel C of a channels-last [B,C+1,X,Y] buffer whose
    other channels must stay untouched."""
    import types
    from _util import target_bev_golden
    from e2e_parking_carla_b200.target_bev import _stamp, target_pixels
    pts, xr, yr, h, w, seed, tmap = target_bev_golden(name)
    b = pts.shape[0]
    cfg = types.SimpleNamespace(bev_x_bound=[0.0, 0.0, xr], bev_y_bound=[0.0, 0.0, yr])
    torch.manual_seed(seed)
    pix = target_pixels((b, 2, h, w), pts.clone(), cfg).to(DEV)
    nchw = torch.full((b, 1, h, w), 7.0, device=DEV)
    _stamp(pix, nchw)
    assert torch.equal(nchw.cpu(), tmap)
    cl = torch.full((b, h, w, 1), 7.0, device=DEV).permute(0, 3, 1, 2)
    _stamp(pix, cl)
    assert torch.equal(cl.cpu(), tmap)
    wide = torch.full((b, h, w, 5), 3.0, device=DEV).permute(0, 3, 1, 2)     # channels-last [B,5,H,W]
    _stamp(pix, wide[:, 4:])
    assert torch.equal(wide[:, 4:].cpu(), tmap) and bool((wide[:, :4] == 3.0).all())


def test_proj_bev_feature_compat_api(lib):
    """The reference's three-call form get_geometry -> encoder_forward -> proj_bev_feature
    (model/bev_model.py:109-113) on the materialised tensors gives the fused path's result
    (forward <= 1e-5; identical zero pattern) and gradients reach the encoder."""
    from e2e_parking_carla_b200 import BevModel

    class Preset(torch.nn.Module):
        def forward(self, images):
            return self.feat, self.logits

    shape = LiftSplatShape(batch=2, channels=8)
    c = _oracle_case(shape, rig_seed=52, in_seed=22)
    ref = _oracle_outputs(shape, c)
    intr, extr = make_rig(2, 4, jitter=True, seed=52)
    images = torch.zeros(2, 4, 3, 8, 8, device=DEV)
    for fmt in (torch.contiguous_format, torch.channels_last):
        enc = Preset()
        enc.feat = c["feat"].to(DEV).requires_grad_(True)
        enc.logits = c["logits"].to(DEV).requires_grad_(True)
        model = BevModel(make_cfg(shape), cam_encoder=enc, bev_memory_format=fmt).to(DEV)
        geom = model.get_geometry(intr.to(DEV), extr.to(DEV))
        x, prob = model.encoder_forward(images)
        assert tuple(x.shape) == (2, 4, 48, 32, 32, 8)
        bev = model.proj_bev_feature(geom, x)
        assert_close(bev, ref["bev"], FP32_TOL, "bev")
        assert np.array_equal(bev.detach().cpu().numpy() == 0, ref["bev"] == 0)
        torch.autograd.backward([bev, prob], [c["gb"].to(DEV), c["gp"].to(DEV)])
        assert_close(enc.feat.grad, ref["grad_feat"], FP32_TOL, "grad_feat")
        assert_close(enc.logits.grad, ref["grad_logits"], FP32_TOL, "grad_logits")


# ------------------------------------------------------------------------------------
# a4 on the reference's own device: geometry="torch" == the reference's torch ops on this GPU
# ------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["rigA_b1", "rigB_b16", "stress_b2"])
def test_torch_geometry_same_device_rank_parity(lib, case):
    """The reference hard-codes .cuda() (model/bev_model.py:46,53): its real path inverts with
    cuSOLVER and multiplies with cuBLAS.  BevModel(geometry="torch") takes M, t from the same
    torch calls and evaluates the per-point transform in cuBLAS's float32 order
    (LS_GEOM_TORCH_CUDA, probed with tools/geom_policy_probe.py): every voxel rank equals the one
    the reference's own op chain (oracle/torch_port.py on CUDA tensors) produces on this GPU."""
    from oracle import torch_port as tp
    from e2e_parking_carla_b200 import BevModel
    ls = _ls()
    shape, jitter, seed = {"rigA_b1": (LiftSplatShape(batch=1, channels=4), False, 0),
                           "rigB_b16": (LiftSplatShape(batch=16, channels=4), True, 1),
                           "stress_b2": (LiftSplatShape.stress(batch=2), True, 61)}[case]
    intr, extr = make_rig(shape.batch, shape.cams, jitter=jitter, seed=seed)
    res, start, dim = grid_of(shape)
    fr = _dev(frustum_of(shape))
    want = tp.ranks_cpu(intr.to(DEV), extr.to(DEV), fr, _dev(start), _dev(res), [int(v) for v in dim])
    model = BevModel(make_cfg(shape), cam_encoder=torch.nn.Identity(), geometry="torch").to(DEV)
    M, t = model.camera_transform(intr.to(DEV), extr.to(DEV))
    s = model._shape(shape.batch, shape.cams, 4)
    assert s.geom_policy == 1
    got = ls.index(M, t, fr, s)
    flips = int((got.long() != want).sum())
    assert flips == 0, "%d of %d ranks differ from the reference's torch-CUDA run" % (flips, want.numel())
    # the CPU-order policy on the same M, t is NOT the same function (that is the point of the policy)
    s_cpu = ls.make_shape(shape.batch, shape.cams, shape.depth_bins, shape.fh, shape.fw, 4, _grid_spec(shape))
    geom_cuda = ls.geometry(M, t, fr, s)
    if jitter:      # (the axis-aligned CARLA rig has too many exact zeros in M for the orders to differ)
        assert not torch.equal(geom_cuda, ls.geometry(M, t, fr, s_cpu))
    # coordinates: bit-equal to torch-CUDA's everywhere except the tail of its batched matmul (the last
    # ~30 points of the last camera of the batch go through another cuBLAS code path and differ by
    # <= 4 ulp; tools/geom_policy_probe.py --library: 76 of 9 437 184 coordinates at B=16)
    bad = geom_cuda != tp.camera_geometry(fr, intr.to(DEV), extr.to(DEV))
    assert int(bad.sum()) <= 128
    assert int(bad.view(shape.batch * shape.cams, -1)[:-1].sum()) == 0          # only the last camera
    assert int(bad[-1, -1].reshape(-1, 3)[:-64].sum()) == 0                     # only its last points


def test_bulk_store_variant_matches(lib):
    """LS_SPLAT_OUT=bulk (shared-memory tile leaving as ONE bulk/TMA store per tile, the first
    channels-last design) against the default direct-row-store splat: a child process per variant
    dumps the BEV tensor, values must agree to the last bit or two."""
    import subprocess
    import sys
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    with tempfile.TemporaryDirectory() as tmp:
        for mode in ("direct", "bulk"):
            path = os.path.join(tmp, mode + ".pt")
            code = (
                "import sys, torch; sys.path.insert(0, %r); import bench\n"
                "from e2e_parking_carla_b200.synthetic import LiftSplatShape\n"
                "st = bench.Stepper(LiftSplatShape(batch=3, channels=64), torch.float32, torch.device('cuda:0'))\n"
                "st.step(); torch.cuda.synchronize()\n"
                "torch.save({k: getattr(st, k).cpu() for k in ('bev', 'prob', 'gfeat', 'glogits')}, %r)\n" % (root, path))
            env = dict(os.environ, LS_SPLAT_OUT=mode)
            out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
            assert out.returncode == 0, out.stderr[-2000:]
            outs.append(torch.load(path))
    a, b = outs
    assert a["bev"].abs().sum() > 0
    assert maxerr(a["bev"], b["bev"]) <= 1e-6 and torch.equal(a["bev"] == 0, b["bev"] == 0)
    for k in ("prob", "gfeat", "glogits"):
        assert torch.equal(a[k], b[k]), k


def test_tma_gather_variant_matches(lib):
    """LS_GATHER_TMA=1 (gradient rows fetched by the copy engine: cp.async.bulk.tensor gather4 into
    shared-memory stages, tools/tma_gather_bench.cu) against the default LDG gather: a child process
    per variant, fp32 and bf16 features, every gradient must have the same bits."""
    import subprocess
    import sys
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.TemporaryDirectory() as tmp:
        for dt in ("float32", "bfloat16"):
            outs = []
            for mode in ("0", "1"):
                path = os.path.join(tmp, dt + mode + ".pt")
                code = (
                    "import sys, torch; sys.path.insert(0, %r); import bench\n"
                    "from e2e_parking_carla_b200.synthetic import LiftSplatShape\n"
                    "st = bench.Stepper(LiftSplatShape(batch=3, channels=64), torch.%s, torch.device('cuda:0'))\n"
                    "st.step(); torch.cuda.synchronize()\n"
                    "torch.save({k: getattr(st, k).float().cpu() for k in ('bev', 'prob', 'gfeat', 'glogits')}, %r)\n"
                    % (root, dt, path))
                env = dict(os.environ, LS_GATHER_TMA=mode)
                out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
                assert out.returncode == 0, out.stderr[-2000:]
                outs.append(torch.load(path))
            a, b = outs
            assert a["gfeat"].abs().sum() > 0 and a["glogits"].abs().sum() > 0
            for k in ("bev", "prob", "gfeat", "glogits"):
                assert torch.equal(a[k], b[k]), (dt, k)


@pytest.mark.parametrize("tile_x", [8, 32])
def test_square_tiles_same_bits_as_strips(lib, tile_x):
    """The channels-last splat with tile_x x (128 / tile_x) tiles (LsShape.tile_x) against 1 x 128
    strips: every cell is summed front to back in canonical order by one quarter-warp, so the BEV
    tensor has the same bits whatever the tiling; gradients do not depend on it at all."""
    import ctypes as C
    from e2e_parking_carla_b200 import _lib
    ls = _ls()
    shape = LiftSplatShape(batch=3, channels=64)
    c = _oracle_case(shape, rig_seed=48, in_seed=23)
    lib_ = _lib.load()
    P = lambda x: C.c_void_p(x.data_ptr())
    fr, Md, td = _dev(frustum_of(shape)), _dev(c["M"]), _dev(c["t"])
    feat, logits = c["feat"].to(DEV), c["logits"].to(DEV)
    gb = c["gb"].to(DEV).contiguous(memory_format=torch.channels_last)
    outs = []
    for tx in (1, tile_x):
        s = _ls_shape(shape, tile_x=tx)
        scratch = torch.empty(ls.scratch_bytes(s, ls.LS_F32, True), dtype=torch.uint8, device=DEV)
        saved = torch.empty(ls.saved_bytes(s, ls.LS_F32), dtype=torch.uint8, device=DEV)
        bev = torch.empty((3, 64, 200, 200), device=DEV).contiguous(memory_format=torch.channels_last)
        prob, gfeat, glogits = torch.empty_like(logits), torch.empty_like(feat), torch.empty_like(logits)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        st, gst = ls._bev_strides(bev), ls._bev_strides(gb)
        ls.check(lib_.ls_forward(P(feat), 0, P(logits), 0, P(Md), P(td), P(fr), C.byref(s), P(scratch), scratch.numel(),
                                 P(saved), saved.numel(), P(bev), C.byref(st), P(prob), stream), "ls_forward")
        ls.check(lib_.ls_backward(P(gb), C.byref(gst), None, P(prob), P(feat), 0, 0, C.byref(s), P(scratch),
                                  scratch.numel(), P(saved), saved.numel(), P(gfeat), P(glogits), stream), "ls_backward")
        outs.append((bev, prob, gfeat, glogits))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    assert_close(outs[1][0], _oracle_outputs(shape, c)["bev"], FP32_TOL, "bev")


@pytest.mark.parametrize("zres", [20.0, 0.3, 7.1, 8.0])
def test_single_z_cell_keep_test_without_division(lib, zres):
    """With one z cell the hot kernel decides -1 < RN((z - off) / res) < 1 as -res < z - off < res
    (no division).  Adversarial z: exactly on both boundaries, one ulp inside / outside, NaN, inf -
    against the oracle, which divides like the reference (model/bev_model.py:85-90)."""
    import ctypes as C
    from oracle import lift_splat_oracle as lo
    from e2e_parking_carla_b200 import _lib
    ls = _ls()
    res = np.array([0.1, 0.1, zres], np.float32)
    start = np.array([-9.95, -9.95, 0.0], np.float32)
    dim = np.array([200, 200, 1], np.int64)
    off = lo.grid_offset(start, res)
    r = np.float32(zres)
    edge = []
    for sgn in (-1.0, 1.0):
        b = np.float32(sgn) * r
        for v in (b, np.nextafter(b, np.float32(0)), np.nextafter(b, np.float32(sgn * np.inf)),
                  np.nextafter(np.nextafter(b, np.float32(0)), np.float32(0))):
            edge.append(v)
    edge += [np.float32(0), np.float32(np.nan), np.float32(np.inf), np.float32(-np.inf), np.float32(1e-30), r / 2]
    rng = np.random.RandomState(3)
    n = 32 * 16
    geom = np.empty((1, n, 3), np.float32)
    geom[0, :, :2] = rng.uniform(-9.9, 9.9, size=(n, 2)).astype(np.float32)
    a = np.array([edge[i % len(edge)] for i in range(n)], np.float32)
    a[len(edge) * 8:] = rng.uniform(-1.5, 1.5, size=n - len(edge) * 8).astype(np.float32) * r
    geom[0, :, 2] = a + off[2]           # z - off reproduces a exactly for these magnitudes? checked below
    back = (geom[0, :, 2] - off[2]).astype(np.float32)
    _, keep_o, rank_o = lo.voxel_index(geom.reshape(1, 1, 1, 16, 32, 3), start, res, dim)
    grid = ls.GridSpec(tuple(float(v) for v in start), tuple(float(v) for v in res), (200, 200, 1))
    s = ls.make_shape(1, 1, 1, 16, 32, 4, grid)
    rank = torch.empty(1, n, dtype=torch.int32, device=DEV)
    g = _dev(geom)
    ls.check(_lib.load().ls_index_geom(C.c_void_p(g.data_ptr()), C.byref(s), C.c_void_p(rank.data_ptr()), None, None,
                                       None, C.c_void_p(torch.cuda.current_stream().cuda_stream)), "ls_index_geom")
    assert np.array_equal(rank.cpu().numpy()[0], rank_o[0].astype(np.int32))
    # the boundary values really are in the input (z - off is exact for them), on both sides of the test
    assert (np.abs(back) == r).sum() >= 8 and 0 < keep_o.sum() < n


# ------------------------------------------------------------------------------------
# static-rig cache (opt-in) and degenerate cells
# ------------------------------------------------------------------------------------
@pytest.mark.parametrize("fmt,dtype", [(torch.channels_last, torch.float32), (torch.contiguous_format, torch.float32),
                                       (torch.channels_last, torch.bfloat16)])
def test_static_rig_cache_bit_identical(lib, fmt, dtype):
    """BevModel(static_rig=True): the first call builds the index structures, later calls only refresh
    the record weights.  Every call must give the bits of the uncached model, also with NEW encoder
    outputs, and a changed rig must be picked up after invalidate_rig_cache()."""
    from e2e_parking_carla_b200 import BevModel

    class Preset(torch.nn.Module):
        def forward(self, images):
            return self.feat, self.logits

    shape = LiftSplatShape(batch=2, channels=64)
    cfg = make_cfg(shape)
    images = torch.zeros(2, 4, 3, 8, 8, device=DEV)
    rigs = [make_rig(2, 4, jitter=True, seed=s) for s in (71, 72)]
    gb, gp = make_upstream_grads(shape, seed=30)
    gb = gb.to(DEV).contiguous(memory_format=fmt)

    def run(model, rig, seed):
        feat, logits = make_encoder_outputs(shape, seed=seed)
        model.cam_encoder.feat = feat.to(DEV, dtype).requires_grad_(True)
        model.cam_encoder.logits = logits.to(DEV, dtype).requires_grad_(True)
        bev, prob = model(images, rig[0].to(DEV), rig[1].to(DEV))
        torch.autograd.backward([bev, prob], [gb, gp.to(DEV, dtype)])
        return bev.detach(), prob.detach(), model.cam_encoder.feat.grad, model.cam_encoder.logits.grad

    plain = BevModel(cfg, cam_encoder=Preset(), bev_memory_format=fmt).to(DEV)
    cached = BevModel(cfg, cam_encoder=Preset(), bev_memory_format=fmt, static_rig=True).to(DEV)
    for step, (rig, seed) in enumerate([(rigs[0], 31), (rigs[0], 32), (rigs[0], 33), (rigs[1], 34), (rigs[1], 35)]):
        if step == 3:
            cached.invalidate_rig_cache()
        want, got = run(plain, rig, seed), run(cached, rig, seed)
        for a, b in zip(want, got):
            assert torch.equal(a, b), step
    assert cached._rig_cache.valid and plain._rig_cache is None


def test_every_point_in_one_cell(lib):
    """Degenerate grid: 2 x 2 cells of 10 m - tens of thousands of points per cell.  The canonical
    ordering must not be quadratic (bucketed path of ls_canon_kernel) and the sums stay deterministic."""
    from oracle import lift_splat_oracle as lo
    import time
    shape = LiftSplatShape(batch=1, channels=4, bev_x_bound=[-10.0, 10.0, 10.0], bev_y_bound=[-10.0, 10.0, 10.0])
    c = _oracle_case(shape, rig_seed=49, in_seed=24)
    counts = np.bincount(c["rank"][0][c["rank"][0] >= 0])
    assert counts.max() > 30000
    ref = _oracle_outputs(shape, c)
    a = _run(shape, c["feat"], c["logits"], c["M"], c["t"], c["gb"], c["gp"])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    b = _run(shape, c["feat"], c["logits"], c["M"], c["t"], c["gb"], c["gp"], bev_format=torch.channels_last)
    torch.cuda.synchronize()
    assert time.perf_counter() - t0 < 2.0
    _check_all(a, ref, 2e-5)          # 40 000-term float32 sums: a little above the 1e-5 of realistic cells
    for k in a:
        assert relerr(a[k], b[k]) <= 1e-5, k
    assert torch.equal(a["bev"], _run(shape, c["feat"], c["logits"], c["M"], c["t"])["bev"])


def test_autocast_and_mixed_dtypes(lib):
    """Under torch.autocast the encoder emits bf16 feature maps / logits: they go to the bf16 kernels
    unchanged, the BEV stays float32 (model/bev_model.py:76), gradients come back in bf16 and reach
    float32 parameters.  Mixed head dtypes are computed in float32."""
    from e2e_parking_carla_b200 import BevModel

    class Enc(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.f = torch.nn.Conv2d(3, 64, 8, stride=8)
            self.d = torch.nn.Conv2d(3, 48, 8, stride=8)

        def forward(self, x):
            return self.f(x).relu(), self.d(x).relu()

    shape = LiftSplatShape(batch=1, channels=64)
    model = BevModel(make_cfg(shape), cam_encoder=Enc()).to(DEV)
    intr, extr = make_rig(1, 4, jitter=True, seed=8)
    images = torch.randn(1, 4, 3, 256, 256, device=DEV)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        bev, depth = model(images, intr.to(DEV), extr.to(DEV))
    assert bev.dtype == torch.float32 and depth.dtype == torch.bfloat16
    (bev.sum() + depth.float().square().sum()).backward()
    assert model.cam_encoder.f.weight.grad.dtype == torch.float32 and model.cam_encoder.f.weight.grad.abs().sum() > 0
    ref, _ = model(images, intr.to(DEV), extr.to(DEV))            # float32 run of the same weights
    assert relerr(bev, ref) <= BF16_TOL
    ls = _ls()
    f, z = model.cam_encoder(images.view(4, 3, 256, 256))
    M, t = model.camera_transform(intr.to(DEV), extr.to(DEV))
    mixed, prob = ls.lift_splat(f.bfloat16(), z, M, t, model.frustum, model._grid)
    assert prob.dtype == torch.float32 and relerr(mixed, ref) <= BF16_TOL


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_bf16_bev_opt_in(lib, dtype):
    """BevModel(bev_dtype=torch.bfloat16): bf16 BEV tensor and bf16 gradient rows (128 bytes per cell),
    sums in float32.  Forward = the float32-BEV result rounded once to bf16; gradients within the bf16
    tolerance of the float64 oracle evaluated on the bf16-rounded upstream gradient; a gradient that
    cannot be read in place (NCHW float32) falls back to the float32 path with the same result."""
    ls = _ls()
    shape = LiftSplatShape(batch=2, channels=64)
    c = _oracle_case(shape, rig_seed=53, in_seed=25)
    fr, grid = _dev(frustum_of(shape)), _grid_spec(shape)
    Md, td = _dev(c["M"]), _dev(c["t"])

    def run(bev_dtype, gb):
        f = c["feat"].to(DEV, dtype).requires_grad_(True)
        z = c["logits"].to(DEV, dtype).requires_grad_(True)
        bev, prob = ls.lift_splat(f, z, Md, td, fr, grid, torch.channels_last, bev_dtype=bev_dtype)
        torch.autograd.backward([bev, prob], [gb, c["gp"].to(DEV, dtype)])
        return bev.detach(), f.grad, z.grad

    gb16 = c["gb"].to(DEV).contiguous(memory_format=torch.channels_last).bfloat16()
    b16, gf16, gl16 = run(torch.bfloat16, gb16)
    assert b16.dtype == torch.bfloat16 and b16.is_contiguous(memory_format=torch.channels_last)
    b32, gf32, gl32 = run(torch.float32, gb16.float())           # same upstream values, float32 tensors
    assert torch.equal(b16, b32.bfloat16())                      # one rounding on the store
    assert torch.equal(gf16, gf32) and torch.equal(gl16, gl32)   # the rows carry the same values either way
    ref = _oracle_outputs(shape, c, dtype)
    assert_close(b16.float(), ref["bev"], BF16_TOL, "bev")
    assert_close(gf16.float(), ref["grad_feat"], BF16_TOL, "grad_feat")
    # fallback: an NCHW float32 gradient on a bf16-BEV forward
    f = c["feat"].to(DEV, dtype).requires_grad_(True)
    z = c["logits"].to(DEV, dtype).requires_grad_(True)
    bev, prob = ls.lift_splat(f, z, Md, td, fr, grid, torch.channels_last, bev_dtype=torch.bfloat16)
    (bev.float() * gb16.float().contiguous()).sum().backward()
    assert relerr(f.grad, gf32) < (1e-5 if dtype == torch.float32 else 2e-2)
