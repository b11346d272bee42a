"""Shared helpers for the test-suite: golden fixtures, regenerated inputs, error norms."""
from __future__ import annotations

import hashlib
import os

import numpy as np
import torch

from e2e_parking_carla_b200.synthetic import LiftSplatShape, make_encoder_outputs, make_upstream_grads

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

GOLDEN_SHAPES = {
    "rigA_b1_c4": LiftSplatShape(batch=1, channels=4),
    "rigB_b2_c4": LiftSplatShape(batch=2, channels=4),
    "stress_b1_c2": LiftSplatShape(batch=1, cams=6, channels=2, bev_x_bound=[-10.0, 10.0, 0.05],
                                   bev_y_bound=[-10.0, 10.0, 0.05], d_bound=[0.5, 12.5, 0.125]),
}


def sha(a) -> str:
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().contiguous().numpy()
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def relerr(a, b) -> float:
    """Frobenius-relative error |a-b| / |b| in float64."""
    a = np.asarray(a.detach().cpu().float() if isinstance(a, torch.Tensor) else a, np.float64)
    b = np.asarray(b.detach().cpu().float() if isinstance(b, torch.Tensor) else b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def maxerr(a, b) -> float:
    """Largest element-wise error relative to the largest reference magnitude:
    max|a-b| / max|b| (float64).  One wrong voxel out of millions passes a Frobenius bound;
    it does not pass this one."""
    a = np.asarray(a.detach().cpu().float() if isinstance(a, torch.Tensor) else a, np.float64)
    b = np.asarray(b.detach().cpu().float() if isinstance(b, torch.Tensor) else b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def assert_close(a, b, tol, what=""):
    """Frobenius-relative AND max-element-wise (relative to max|b|) error within tol."""
    e, m = relerr(a, b), maxerr(a, b)
    assert e <= tol, "%s: Frobenius-relative error %.3g > %.3g" % (what, e, tol)
    assert m <= tol, "%s: max element-wise error %.3g (of max|ref|) > %.3g" % (what, m, tol)


class Golden:
    """One fixture frozen from the unmodified reference (tests/golden/make_golden.py)."""

    def __init__(self, name: str):
        self.name = name
        self.shape = GOLDEN_SHAPES[name]
        self.z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.dstride = int(self.z["dstride"])

    def __getitem__(self, k):
        return self.z[k]

    def inputs(self):
        """(feat, logits, grad_bev, grad_prob) regenerated from the seed and verified
        against the checksums stored with the fixture."""
        seed = int(self.z["in_seed"])
        feat, logits = make_encoder_outputs(self.shape, seed=seed)
        gb, gp = make_upstream_grads(self.shape, seed=seed)
        assert sha(feat) == str(self.z["feat_sha"]), "input generator drifted (feat)"
        assert sha(logits) == str(self.z["logits_sha"]), "input generator drifted (logits)"
        assert sha(gb) == str(self.z["grad_bev_sha"]), "input generator drifted (grad_bev)"
        assert sha(gp) == str(self.z["grad_prob_sha"]), "input generator drifted (grad_prob)"
        return feat, logits, gb, gp


def grid_of(shape: LiftSplatShape):
    from oracle import lift_splat_oracle as lo
    return lo.bev_grid_params(shape.bev_x_bound, shape.bev_y_bound, shape.bev_z_bound)


def frustum_of(shape: LiftSplatShape) -> np.ndarray:
    from oracle import lift_splat_oracle as lo
    return lo.create_frustum(shape.d_bound, shape.final_dim, shape.bev_down_sample)


# ------------------------------------------------------------------------------------
# add_target_bev fixtures (tests/golden/make_golden_target_bev.py)
# ------------------------------------------------------------------------------------
TARGET_BEV_CASES = ("inner_b16", "border_b12", "stress_b8", "ragged_b6")


def target_bev_golden(name):
    """(target points f32[B,3], x_res, y_res, H, W, seed, target map f32[B,1,H,W]) of one frozen case.
    The reference stamps one python-slice box per sample, so the stored row / column occupancy is
    the map: map[b] = rows[b] (outer) cols[b]; the stored per-sample sums cross-check it."""
    import torch
    z = np.load(os.path.join(GOLDEN_DIR, "target_bev.npz"))
    xr, yr, h, w, seed = z[name + "_meta"]
    h, w = int(h), int(w)
    rows = np.unpackbits(z[name + "_rows"], axis=1)[:, :h].astype(np.float32)
    cols = np.unpackbits(z[name + "_cols"], axis=1)[:, :w].astype(np.float32)
    tmap = rows[:, :, None] * cols[:, None, :]
    assert np.array_equal(tmap.sum((1, 2)).astype(np.int32), z[name + "_sum"])
    return torch.from_numpy(z[name + "_target"]), float(xr), float(yr), h, w, int(seed), torch.from_numpy(tmap)[:, None]
