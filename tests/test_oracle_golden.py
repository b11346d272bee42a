"""Pin the CPU oracle (oracle/lift_splat_oracle.py, oracle/torch_port.py) to the reference.

The reference ships no tests or golden vectors for the lift-splat path, so
the pin is (a) fixtures frozen from the UNMODIFIED reference run in the build container
(tests/golden/make_golden.py) and (b), when /root/reference is present, the live
reference itself.  No GPU needed.
"""
import os

import numpy as np
import pytest
import torch

from _util import GOLDEN_SHAPES, Golden, frustum_of, grid_of, relerr, sha
from e2e_parking_carla_b200.synthetic import (LiftSplatShape, make_cfg, make_encoder_outputs, make_rig,
                                              make_upstream_grads)
from oracle import lift_splat_oracle as lo
from oracle import ref_harness as rh
from oracle import torch_port as tp


def test_grid_params_known_answers():
    """tool/geometry.py:40-59 at config/training.yaml:26-28 and the stress grid."""
    res, start, dim = lo.bev_grid_params([-10.0, 10.0, 0.1], [-10.0, 10.0, 0.1], [-10.0, 10.0, 20.0])
    assert res.dtype == np.float32 and start.dtype == np.float32 and dim.dtype == np.int64
    assert dim.tolist() == [200, 200, 1]
    assert np.array_equal(res, np.array([0.1, 0.1, 20.0], np.float32))
    assert np.array_equal(start, np.array([-9.95, -9.95, 0.0], np.float32))
    assert np.array_equal(lo.grid_offset(start, res), np.array([-10.0, -10.0, -10.0], np.float32))
    _, _, dim = lo.bev_grid_params([-10.0, 10.0, 0.05], [-10.0, 10.0, 0.05], [-10.0, 10.0, 20.0])
    assert dim.tolist() == [400, 400, 1]


def test_frustum_matches_torch():
    """create_frustum restates torch.arange / torch.linspace (model/bev_model.py:28-43)."""
    for shape in GOLDEN_SHAPES.values():
        fr = frustum_of(shape)
        d = torch.arange(*shape.d_bound, dtype=torch.float)
        u = torch.linspace(0, shape.final_dim[1] - 1, shape.fw, dtype=torch.float)
        v = torch.linspace(0, shape.final_dim[0] - 1, shape.fh, dtype=torch.float)
        assert np.array_equal(fr[:, 0, 0, 2], d.numpy())
        assert np.array_equal(fr[0, 0, :, 0], u.numpy())
        assert np.array_equal(fr[0, :, 0, 1], v.numpy())
    fr = frustum_of(GOLDEN_SHAPES["rigA_b1_c4"])
    assert np.allclose(fr[0, 0, :4, 0], [0, 8.2258062, 16.4516125, 24.6774178], rtol=0, atol=1e-6)


@pytest.mark.parametrize("name", list(GOLDEN_SHAPES))
def test_indices_bit_exact_vs_golden(name):
    """geometry / voxel index / keep mask / rank / sorted ranks: bit-exact with the reference
    given the reference's own M, t."""
    g = Golden(name)
    fr = frustum_of(g.shape)
    assert sha(fr) == str(g["frustum_sha"])
    res, start, dim = grid_of(g.shape)
    geom = lo.geometry(g["M_ref"], g["t_ref"], fr)
    assert sha(geom) == str(g["geom_sha"])
    vox, keep, rank = lo.voxel_index(geom, start, res, dim)
    assert np.array_equal(rank.astype(np.int32), g["rank_ref"])
    assert np.array_equal(keep.reshape(g.shape.batch, g.shape.cams, -1).sum(-1), g["kept_per_cam"])
    for b in range(g.shape.batch):
        sr = lo.sorted_ranks(rank[b])
        assert sha(sr.astype(np.int64)) == str(g["sorted_rank_sha"][b])
        assert np.unique(sr).size == int(g["segments"][b])


def test_rig_a_known_answers():
    """196 608 points, 155 296 kept (32 768 / 42 848 / 42 624 / 37 056 per
    camera), 28 909 voxels hit, at most 128 points per voxel; floor instead of trunc would
    keep only 150 016."""
    g = Golden("rigA_b1_c4")
    rank = g["rank_ref"][0]
    assert rank.size == 196608 and (rank >= 0).sum() == 155296
    assert g["kept_per_cam"][0].tolist() == [32768, 42848, 42624, 37056]
    counts = np.bincount(rank[rank >= 0])
    assert (counts > 0).sum() == 28909 and counts.max() == 128
    res, start, dim = grid_of(g.shape)
    geom = lo.geometry(g["M_ref"], g["t_ref"], frustum_of(g.shape))
    c = ((geom - lo.grid_offset(start, res)) / res).reshape(-1, 3)
    fl = np.floor(c)
    kept_floor = ((fl[:, 0] >= 0) & (fl[:, 0] < 200) & (fl[:, 1] >= 0) & (fl[:, 1] < 200) & (fl[:, 2] >= 0) & (fl[:, 2] < 1))
    assert kept_floor.sum() == 150016
    # camera centres of the CARLA rig (dataset/carla_dataset.py:209-230)
    centres = np.linalg.inv(g["extrinsics"][0].astype(np.float64))[:, :3, 3]
    assert np.allclose(centres, [[1.5, 0, 1.5], [0, -0.8, 1.5], [0, 0.8, 1.5], [-2.2, 0, 1.5]], atol=1e-6)
    assert np.allclose(g["intrinsics"][0, 0], [[167.8199, 0, 128], [0, 167.8199, 128], [0, 0, 1]], atol=1e-4)


@pytest.mark.parametrize("name", list(GOLDEN_SHAPES))
def test_values_vs_golden_fp64(name):
    """Oracle forward / backward against the reference executed in float64."""
    g = Golden(name)
    feat, logits, gb, gp = g.inputs()
    res, start, dim = grid_of(g.shape)
    rank = g["rank_ref"].astype(np.int64)
    bev, prob = lo.splat_forward(feat.numpy(), logits.numpy(), rank, dim, g.shape.cams)
    gf, gl = lo.splat_backward(feat.numpy(), logits.numpy(), rank, dim, g.shape.cams, gb.numpy(), gp.numpy())
    ds = g.dstride
    assert relerr(bev, g["bev_ref64"]) < 5e-7
    assert np.array_equal(bev == 0, g["bev_ref64"] == 0)
    assert relerr(prob[:, ::ds], g["prob_ref"]) < 5e-7
    assert relerr(gf, g["grad_feat_ref64"]) < 1e-6
    assert relerr(gl[:, ::ds], g["grad_logits_ref64"]) < 1e-6
    # the reference's own fp32 cumsum trick is far noisier than that
    assert float(g["ref32_vs_ref64_bev"]) > 1e-4


def test_own_camera_transform_close_to_lapack():
    """camera_transform (fp64 Gauss-Jordan rounded once) vs torch.inverse (MKL LAPACK):
    rounding-level differences only; identical voxel ranks on the CARLA rig."""
    for jitter, seed in ((False, 0), (True, 3)):
        intr, extr = make_rig(4, 6, jitter=jitter, seed=seed)
        M, t = lo.camera_transform(intr.numpy(), extr.numpy())
        inv = torch.inverse(extr)
        Mt = inv[..., :3, :3].matmul(torch.inverse(intr)).numpy()
        assert np.abs(M - Mt).max() < 5e-7 and np.abs(t - inv[..., :3, 3].numpy()).max() < 2e-6
    g = Golden("rigA_b1_c4")
    M, t = lo.camera_transform(g["intrinsics"], g["extrinsics"])
    res, start, dim = grid_of(g.shape)
    _, _, rank = lo.voxel_index(lo.geometry(M, t, frustum_of(g.shape)), start, res, dim)
    assert np.array_equal(rank.astype(np.int32), g["rank_ref"])
    # jittered rigs: a handful of boundary points may flip (SURVEY.md 7 "hard parts")
    g = Golden("rigB_b2_c4")
    M, t = lo.camera_transform(g["intrinsics"], g["extrinsics"])
    _, _, rank = lo.voxel_index(lo.geometry(M, t, frustum_of(g.shape)), start, res, dim)
    assert (rank.astype(np.int32) != g["rank_ref"]).sum() <= 8


def test_to_long_semantics():
    """Tensor.long() on x86: truncation toward zero, INT64_MIN for NaN / out of range."""
    x = np.array([-0.5, -1.0, -1.5, 0.999, 199.99998, np.nan, np.inf, -np.inf, 1e30], np.float32)
    got = lo._to_long(x)
    imin = np.iinfo(np.int64).min
    assert got.tolist() == [0, -1, -1, 0, 199, imin, imin, imin, imin]


@pytest.mark.parametrize("name", ["rigA_b1_c4", "rigB_b2_c4"])
def test_torch_port_matches_golden(name):
    """oracle/torch_port.py (what bench.py times as the CPU baseline) reproduces the
    reference's ranks exactly and its fp32 BEV to within the reference's own fp32 noise."""
    g = Golden(name)
    feat, logits, gb, gp = g.inputs()
    res, start, dim = grid_of(g.shape)
    fr = torch.from_numpy(frustum_of(g.shape))
    intr, extr = torch.from_numpy(g["intrinsics"]), torch.from_numpy(g["extrinsics"])
    a = (fr, torch.from_numpy(start), torch.from_numpy(res), torch.from_numpy(dim))
    rank = tp.ranks_cpu(intr, extr, *a)
    assert np.array_equal(rank.numpy().astype(np.int32), g["rank_ref"])
    bev, prob, gf, gl = tp.fwd_bwd_step(feat, logits, intr, extr, *a, gb, gp)
    noise = float(g["ref32_vs_ref64_bev"])
    assert abs(relerr(bev, g["bev_ref64"]) - noise) < 0.05 * noise      # same algorithm, same error
    ds = g.dstride
    assert relerr(gf, g["grad_feat_ref64"]) < 1e-6 and relerr(gl[:, ::ds], g["grad_logits_ref64"]) < 1e-6


@pytest.mark.skipif(not rh.available(), reason="reference tree not present")
@pytest.mark.parametrize("kw", [dict(channels=6), dict(channels=2, bev_x_bound=[-10.0, 10.0, 1.0], bev_y_bound=[-10.0, 10.0, 1.0]),
                                dict(channels=2, cams=6, d_bound=[0.5, 12.5, 0.125]),
                                dict(channels=3, final_dim=[128, 192], bev_down_sample=16,
                                     bev_x_bound=[-7.0, 9.8, 0.3], bev_y_bound=[-12.0, 4.0, 0.125])],
                         ids=["default_c6", "1m_voxels_long_segments", "6_cams_96_bins", "ragged_everything"])
def test_live_reference_equals_port_and_oracle(kw):
    """With /root/reference mounted: the unmodified reference, the torch port and the numpy
    oracle agree (bit-exact ranks and fp32 outputs for the port; fp64-level for the oracle)."""
    shape = LiftSplatShape(batch=2, **kw)
    cfg = make_cfg(shape)
    intr, extr = make_rig(2, shape.cams, jitter=True, seed=77)
    feat, logits = make_encoder_outputs(shape, seed=70)
    gb, gp = make_upstream_grads(shape, seed=70)
    ref32 = rh.run_reference(cfg, feat, logits, intr, extr, double=False, backward_with=(gb, gp))
    ref64 = rh.run_reference(cfg, feat, logits, intr, extr, double=True, backward_with=(gb, gp))
    model = rh.reference_bev_model(cfg)
    a = (model.frustum.data, model.bev_start_pos.data, model.bev_res.data, model.bev_dim.data)
    bev, prob, gf, gl = tp.fwd_bwd_step(feat, logits, intr, extr, *a, gb, gp)
    assert torch.equal(bev, ref32["bev"]) and torch.equal(prob, ref32["prob"])
    assert torch.equal(gf, ref32["grad_feat"]) and torch.equal(gl, ref32["grad_logits"])
    _, vox, keep, ranks = rh.reference_indices(cfg, intr, extr)
    inv = torch.inverse(extr)
    M = inv[..., :3, :3].matmul(torch.inverse(intr)).numpy()
    res, start, dim = grid_of(shape)
    vox_o, keep_o, rank_o = lo.voxel_index(lo.geometry(M, inv[..., :3, 3].numpy(), frustum_of(shape)), start, res, dim)
    assert np.array_equal(vox_o, vox.numpy()) and np.array_equal(keep_o, keep.numpy())
    bev_o, _ = lo.splat_forward(feat.numpy(), logits.numpy(), rank_o, dim, shape.cams)
    assert relerr(bev_o, ref64["bev"]) < 5e-7


_SWEEP = {
    # name: (LiftSplatShape kwargs, cameras, jitter, rig seed)
    "default_seed1": (dict(), 4, True, 1),
    "default_seed2": (dict(), 4, True, 2),
    "default_seed3": (dict(), 4, True, 3),
    "carla_rig_exact": (dict(), 4, False, 0),                      # 16 384 points exactly on voxel boundaries
    "stress_6cam": (dict(cams=6, bev_x_bound=[-10.0, 10.0, 0.05], bev_y_bound=[-10.0, 10.0, 0.05],
                         d_bound=[0.5, 12.5, 0.125]), 6, True, 4),
    "coarse_1m": (dict(bev_x_bound=[-10.0, 10.0, 1.0], bev_y_bound=[-10.0, 10.0, 1.0]), 4, True, 5),
    "ragged_grid": (dict(bev_x_bound=[-7.0, 9.8, 0.3], bev_y_bound=[-12.0, 4.0, 0.125]), 4, True, 6),
    "far_depths": (dict(d_bound=[2.0, 58.0, 1.0]), 4, True, 7),
    "thin_z_slab": (dict(bev_z_bound=[-1.0, 1.0, 2.0]), 4, True, 8),  # most points fail the z test
    "small_maps": (dict(final_dim=[128, 192], bev_down_sample=16), 4, True, 9),
}


@pytest.mark.skipif(not rh.available(), reason="reference tree not present")
@pytest.mark.parametrize("name", list(_SWEEP))
def test_live_reference_index_sweep(name):
    """Voxel indices, keep mask and sorted ranks of the numpy oracle against the unmodified reference's
    own tensor ops on rigs, grids and depth ranges beyond the three frozen fixtures (indices only: the
    bit-exact half of the bar, and what every GPU parity test at other sizes leans on)."""
    kw, cams, jitter, seed = _SWEEP[name]
    shape = LiftSplatShape(batch=2, channels=4, **kw)
    assert shape.cams == cams
    cfg = make_cfg(shape)
    intr, extr = make_rig(2, cams, jitter=jitter, seed=seed)
    geom, vox, keep, ranks = rh.reference_indices(cfg, intr, extr)
    res, start, dim = grid_of(shape)
    # grid parameters themselves: tool/geometry.py:40-59
    model = rh.reference_bev_model(cfg)
    assert np.array_equal(model.bev_res.numpy(), res) and np.array_equal(model.bev_start_pos.numpy(), start)
    assert np.array_equal(model.bev_dim.numpy(), dim) and np.array_equal(model.frustum.numpy(), frustum_of(shape))
    inv = torch.inverse(extr)
    M = inv[..., :3, :3].matmul(torch.inverse(intr)).numpy()
    geom_o = lo.geometry(M, inv[..., :3, 3].numpy(), frustum_of(shape))
    assert np.array_equal(geom_o, geom.numpy())                       # torch-CPU's float32 evaluation order
    vox_o, keep_o, rank_o = lo.voxel_index(geom_o, start, res, dim)
    assert np.array_equal(vox_o, vox.numpy()) and np.array_equal(keep_o, keep.numpy())
    for b in range(2):
        kept = np.sort(rank_o[b][rank_o[b] >= 0])
        assert np.array_equal(kept, ranks[b].numpy())
    assert 0 < keep_o.sum() < keep_o.size or name == "thin_z_slab"


# ------------------------------------------------------------------------------------
# DepthLoss (loss/depth_loss.py:18-48): the consumer of pred_depth
# ------------------------------------------------------------------------------------
def _depth_loss_case():
    from e2e_parking_carla_b200.synthetic import make_depth_labels
    shape = LiftSplatShape(batch=1, channels=4)
    _, logits = make_encoder_outputs(shape, seed=31)
    return shape, logits, make_depth_labels(shape, seed=31)


def test_depth_loss_oracle_vs_golden():
    """The numpy restatement reproduces the unmodified reference DepthLoss: identical bin labels
    (min-pool ignoring zeros, truncation, range test), loss and gradient of its float32 run."""
    from oracle import depth_loss_oracle as dlo
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "depth_loss_b1.npz"))
    shape, logits, gt = _depth_loss_case()
    prob = logits.softmax(dim=1).numpy()
    loss, grad, labels = dlo.depth_loss(prob, gt.numpy(), shape.d_bound, shape.bev_down_sample)
    assert np.array_equal(labels, z["labels"].astype(np.int64))
    assert int((labels > 0).sum()) == int(z["fg"]) and set(np.unique(labels)) == set(range(49))
    assert abs(loss - float(z["loss32"])) <= 1e-6 * abs(loss)
    assert np.abs(grad - z["grad32"]).max() <= 1e-6 * np.abs(z["grad32"]).max()


@pytest.mark.parametrize("kw", [dict(), dict(d_bound=[0.5, 12.5, 0.125]), dict(d_bound=[2.0, 58.0, 1.0]),
                                dict(final_dim=[128, 192], bev_down_sample=16), dict(cams=6)],
                         ids=["default", "96_bins", "1m_bins", "ds16_ragged", "6_cams"])
def test_depth_loss_live_reference(kw):
    """Against the live reference module when the tree is mounted (labels with exact-boundary
    depths, all-zero blocks and out-of-range blocks), for several bin widths, down-sampling factors
    and camera counts."""
    import sys
    if not os.path.isfile("/root/reference/loss/depth_loss.py"):
        pytest.skip("reference tree not present")
    sys.path.insert(0, "/root/reference")
    from loss.depth_loss import DepthLoss as RefLoss
    from oracle import depth_loss_oracle as dlo
    from e2e_parking_carla_b200.synthetic import make_cfg, make_depth_labels
    shape = LiftSplatShape(batch=2, channels=4, **kw)
    _, logits = make_encoder_outputs(shape, seed=32)
    gt = make_depth_labels(shape, seed=32)
    prob = logits.softmax(dim=1).requires_grad_(True)
    ref = RefLoss(make_cfg(shape))
    loss_ref = ref(prob, gt)
    loss_ref.backward()
    loss, grad, labels = dlo.depth_loss(prob.detach().numpy(), gt.numpy(), shape.d_bound, shape.bev_down_sample)
    onehot = ref.get_down_sampled_gt_depth(gt)
    lab_ref = torch.where(onehot.sum(1) > 0, onehot.argmax(1) + 1, torch.zeros(onehot.shape[0], dtype=torch.long))
    assert np.array_equal(labels, lab_ref.numpy())
    assert abs(loss - loss_ref.item()) <= 1e-6 * abs(loss)
    assert np.abs(grad - prob.grad.numpy()).max() <= 1e-6 * np.abs(grad).max()


# ------------------------------------------------------------------------------------
# add_target_bev (model/parking_model.py:28-46): fixtures frozen from the unmodified method
# ------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["inner_b16", "border_b12", "stress_b8", "ragged_b6"])
def test_target_bev_port_vs_golden(name):
    """The port (what the GPU tests check ``add_target_bev`` against) reproduces the unmodified
    reference method under the same generator seed: noise draw, truncation, python-slice borders."""
    from _util import target_bev_golden
    tp_pts, xr, yr, h, w, seed, tmap = target_bev_golden(name)
    bev = torch.zeros(tp_pts.shape[0], 2, h, w)
    torch.manual_seed(seed)
    wide, got = tp.add_target_bev_ref(bev, tp_pts.clone(), xr, yr)
    assert torch.equal(got, tmap) and torch.equal(wide[:, 2:], tmap) and float(wide[:, :2].abs().sum()) == 0.0
    # the host half of the product (the pixel computation is torch ops on whatever device the points
    # live on) lands on the same pixels: the stamp is rows cx-4..cx+3 / cols cy-4..cy+3 where in range
    from e2e_parking_carla_b200.target_bev import target_pixels
    import types
    cfg = types.SimpleNamespace(bev_x_bound=[0.0, 0.0, xr], bev_y_bound=[0.0, 0.0, yr])
    torch.manual_seed(seed)
    pix = target_pixels((tp_pts.shape[0], 2, h, w), tp_pts.clone(), cfg)
    assert pix.dtype == torch.int32 and tuple(pix.shape) == (tp_pts.shape[0], 2)
    for i in range(pix.shape[0]):
        ref = torch.zeros(h, w)
        cx, cy = int(pix[i, 0]), int(pix[i, 1])
        ref[cx - 4:cx + 4, cy - 4:cy + 4] = 1.0
        assert torch.equal(ref, tmap[i, 0])


def test_target_bev_known_answers():
    """What the border fixture pins: a stamp fully inside has 64 ones; a negative slice start next to
    a positive stop selects nothing; the far border clips; two negative bounds wrap around."""
    from _util import target_bev_golden
    sums = {n: target_bev_golden(n)[6].sum((1, 2, 3)).int().tolist() for n in ("inner_b16", "border_b12")}
    assert sums["inner_b16"] == [64] * 16
    assert sums["border_b12"] == [0, 25, 64, 64, 0, 0, 64, 0, 32, 0, 0, 64]
    tmap = target_bev_golden("border_b12")[6]
    assert tmap[6, 0, 150:, 150:].sum() == 64        # target (-11, -11) m: both bounds negative -> wraps to the far corner


def test_target_bev_live_reference():
    """The fixtures regenerate from the reference tree when it is mounted."""
    if not os.path.isfile("/root/reference/model/parking_model.py"):
        pytest.skip("reference tree not present")
    import importlib.util
    from _util import TARGET_BEV_CASES, target_bev_golden
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden_target_bev",
                                                  os.path.join(here, "golden", "make_golden_target_bev.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    fn, lines = gen.reference_add_target_bev()
    assert tuple(lines) == (28, 46)                  # the lines every citation in this repo names
    cases = gen.cases()
    assert set(cases) == set(TARGET_BEV_CASES)
    for name, (pts, xr, yr, h, w, seed) in cases.items():
        g_pts, g_xr, g_yr, g_h, g_w, g_seed, tmap = target_bev_golden(name)
        assert torch.equal(pts, g_pts) and (xr, yr, h, w, seed) == (g_xr, g_yr, g_h, g_w, g_seed)
        assert torch.equal(gen.run_reference(fn, pts, xr, yr, h, w, seed), tmap)
