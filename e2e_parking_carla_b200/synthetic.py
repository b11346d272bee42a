"""Synthetic camera rigs and lift-splat inputs (no CARLA, no dataset).

The reference builds its rig in ``dataset/carla_dataset.py:206-270`` from
``carla.Transform`` objects.  ``carla`` is not installable here, so the UE4
left-handed transform is restated from its published definition
(``carla.Transform.get_matrix``: yaw about Z, pitch about Y, roll about X, all
in degrees) and checked against the camera centres the survey probed.

Everything here is input generation for tests / bench / smoke.  It is shared by
the product tests and by the oracle tests so both sides see identical tensors.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import numpy as np
import torch

# cam axes (x right, y down, z forward)  <-  UE4 axes (x forward, y right, z up)
# reference: dataset/carla_dataset.py:254-259
_CAM2PIXEL = np.array([[0, 1, 0, 0],
                       [0, 0, -1, 0],
                       [1, 0, 0, 0],
                       [0, 0, 0, 1]], dtype=np.float64)

# reference rig: dataset/carla_dataset.py:209-230  (x, y, z, roll, pitch, yaw)
CARLA_RIG: Tuple[Tuple[float, ...], ...] = (
    (1.5, 0.0, 1.5, 0.0, 0.0, 0.0),       # rgb_front
    (0.0, -0.8, 1.5, 0.0, -40.0, -90.0),  # rgb_left
    (0.0, 0.8, 1.5, 0.0, -40.0, 90.0),    # rgb_right
    (-2.2, 0.0, 1.5, 0.0, -30.0, 180.0),  # rgb_rear
)
# two extra corner cameras for the 6-camera stress rig
CORNER_CAMS: Tuple[Tuple[float, ...], ...] = (
    (0.8, -0.8, 1.5, 0.0, -20.0, -45.0),
    (0.8, 0.8, 1.5, 0.0, -20.0, 45.0),
)


def carla_transform_matrix(x, y, z, roll, pitch, yaw) -> np.ndarray:
    """4x4 camera->vehicle matrix of a ``carla.Transform`` (float64)."""
    cy, sy = math.cos(math.radians(yaw)), math.sin(math.radians(yaw))
    cr, sr = math.cos(math.radians(roll)), math.sin(math.radians(roll))
    cp, sp = math.cos(math.radians(pitch)), math.sin(math.radians(pitch))
    return np.array([
        [cp * cy, cy * sp * sr - sy * cr, -cy * sp * cr - sy * sr, x],
        [cp * sy, sy * sp * sr + cy * cr, -sy * sp * cr + cy * sr, y],
        [sp, -cp * sr, cp * cr, z],
        [0.0, 0.0, 0.0, 1.0]], dtype=np.float64)


def vehicle_to_camera(spec: Sequence[float]) -> np.ndarray:
    """veh2cam extrinsic as the dataset builds it (carla_dataset.py:260-264)."""
    return _CAM2PIXEL @ np.linalg.inv(carla_transform_matrix(*spec))


def cropped_intrinsics(width=400, height=300, fov=100.0, crop=256,
                       focal_scale=1.0, pp_shift=(0.0, 0.0)) -> np.ndarray:
    """Pinhole K after the centre crop (carla_dataset.py:233-251, tool/geometry.py:16-37)."""
    f = width / (2.0 * math.tan(fov * math.pi / 360.0))
    k = np.array([[f, 0.0, width / 2.0],
                  [0.0, f, height / 2.0],
                  [0.0, 0.0, 1.0]], dtype=np.float64)
    k = k.astype(np.float32)          # the dataset casts before cropping
    k[0, 2] -= np.float32((width - crop) / 2.0)
    k[1, 2] -= np.float32((height - crop) / 2.0)
    k[0, 0] *= np.float32(focal_scale)
    k[1, 1] *= np.float32(focal_scale)
    k[0, 2] += np.float32(pp_shift[0])
    k[1, 2] += np.float32(pp_shift[1])
    return k


def make_rig(batch: int, cams: int = 4, jitter: bool = False, seed: int = 0
             ) -> Tuple[torch.Tensor, torch.Tensor]:
    """(intrinsics f32[B,N,3,3], extrinsics f32[B,N,4,4]).

    ``jitter=False`` is rig A (every sample the exact CARLA rig);
    ``jitter=True`` is rig B: per sample and camera, position sigma 0.2/0.2/0.1 m,
    roll 2 deg, pitch 5 deg, yaw 10 deg, focal x(1+N(0,.05)), principal point
    +N(0,3 px)  (SURVEY.md §8d).
    """
    specs = list(CARLA_RIG) + list(CORNER_CAMS)
    if cams > len(specs):
        raise ValueError("at most %d cameras" % len(specs))
    rng = np.random.RandomState(seed)
    intr = np.zeros((batch, cams, 3, 3), np.float32)
    extr = np.zeros((batch, cams, 4, 4), np.float32)
    for b in range(batch):
        for n in range(cams):
            s = list(specs[n])
            fs, pp = 1.0, (0.0, 0.0)
            if jitter:
                s[0] += rng.normal(0, 0.2)
                s[1] += rng.normal(0, 0.2)
                s[2] += rng.normal(0, 0.1)
                s[3] += rng.normal(0, 2.0)
                s[4] += rng.normal(0, 5.0)
                s[5] += rng.normal(0, 10.0)
                fs = 1.0 + rng.normal(0, 0.05)
                pp = (rng.normal(0, 3.0), rng.normal(0, 3.0))
            intr[b, n] = cropped_intrinsics(focal_scale=fs, pp_shift=pp)
            extr[b, n] = vehicle_to_camera(s).astype(np.float32)
    return torch.from_numpy(intr), torch.from_numpy(extr)


@dataclass
class LiftSplatShape:
    """Sizes of one lift-splat problem (defaults = config/training.yaml:22-33)."""
    batch: int = 1
    cams: int = 4
    channels: int = 64
    bev_x_bound: List[float] = field(default_factory=lambda: [-10.0, 10.0, 0.1])
    bev_y_bound: List[float] = field(default_factory=lambda: [-10.0, 10.0, 0.1])
    bev_z_bound: List[float] = field(default_factory=lambda: [-10.0, 10.0, 20.0])
    d_bound: List[float] = field(default_factory=lambda: [0.5, 12.5, 0.25])
    final_dim: List[int] = field(default_factory=lambda: [256, 256])
    bev_down_sample: int = 8

    @property
    def fh(self) -> int:
        return self.final_dim[0] // self.bev_down_sample

    @property
    def fw(self) -> int:
        return self.final_dim[1] // self.bev_down_sample

    @property
    def depth_bins(self) -> int:
        return int(torch.arange(*self.d_bound, dtype=torch.float).numel())

    @classmethod
    def stress(cls, batch: int = 32) -> "LiftSplatShape":
        """BASELINE.json configs[3]: 400x400 @0.05 m, 96 bins, 6 cameras."""
        return cls(batch=batch, cams=6,
                   bev_x_bound=[-10.0, 10.0, 0.05], bev_y_bound=[-10.0, 10.0, 0.05],
                   d_bound=[0.5, 12.5, 0.125])


def make_cfg(shape: LiftSplatShape):
    """Duck-typed stand-in for the reference ``tool.config.Configuration``
    (only the fields ``BevModel`` reads: tool/config.py:29-37)."""
    from types import SimpleNamespace
    return SimpleNamespace(
        bev_x_bound=list(shape.bev_x_bound), bev_y_bound=list(shape.bev_y_bound),
        bev_z_bound=list(shape.bev_z_bound), d_bound=list(shape.d_bound),
        final_dim=list(shape.final_dim), bev_down_sample=shape.bev_down_sample,
        use_depth_distribution=1, backbone="efficientnet-b4",
        bev_encoder_in_channel=shape.channels, device=torch.device("cpu"))


def _irwin_hall_normal(rng: np.random.RandomState, shape) -> np.ndarray:
    """Unit-variance, zero-mean, bell-shaped float32 noise built ONLY from integer draws
    (sum of four uniform 16-bit integers, one IEEE float32 divide), so that every host
    regenerates bit-identical inputs - no libm, no vectorised normal sampler involved."""
    acc = np.zeros(shape, dtype=np.int64)
    for _ in range(4):
        acc += rng.randint(0, 65536, size=shape).astype(np.int64)
    centred = (acc - 131070).astype(np.float32)           # mean of the sum is 4*32767.5
    return centred / np.float32(37837.2265625)            # std of the sum is 65536/sqrt(3)


def make_encoder_outputs(shape: LiftSplatShape, seed: int = 0, relu: bool = True,
                         dtype=torch.float32) -> Tuple[torch.Tensor, torch.Tensor]:
    """Stand-ins for ``CamEncoder`` outputs: feat[B*N,C,h,w], depth_logits[B*N,D,h,w].

    Both reference heads end in BN+ReLU (model/convolutions.py:189-196), hence
    ``relu=True`` by default; ``relu=False`` gives the zero-mean variant used for
    the tight-tolerance check.
    """
    rng = np.random.RandomState(seed)
    bn = shape.batch * shape.cams
    feat = torch.from_numpy(_irwin_hall_normal(rng, (bn, shape.channels, shape.fh, shape.fw)))
    logit = torch.from_numpy(_irwin_hall_normal(rng, (bn, shape.depth_bins, shape.fh, shape.fw)))
    if relu:
        feat, logit = feat.relu(), logit.relu()
    return feat.to(dtype), logit.to(dtype)


def make_upstream_grads(shape: LiftSplatShape, seed: int = 0, dtype=torch.float32
                        ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Synthetic gradients arriving on the two outputs: grad_bev f32[B,C,X,Y] and
    grad_pred_depth[B*N,D,h,w]."""
    rng = np.random.RandomState(1000 + seed)
    x = int(round((shape.bev_x_bound[1] - shape.bev_x_bound[0]) / shape.bev_x_bound[2]))
    y = int(round((shape.bev_y_bound[1] - shape.bev_y_bound[0]) / shape.bev_y_bound[2]))
    gb = torch.from_numpy(_irwin_hall_normal(rng, (shape.batch, shape.channels, x, y)))
    gp = torch.from_numpy(_irwin_hall_normal(rng, (shape.batch * shape.cams, shape.depth_bins, shape.fh, shape.fw)))
    return gb, gp.to(dtype)


def make_depth_labels(shape: LiftSplatShape, seed: int = 0) -> torch.Tensor:
    """Synthetic metric ground-truth depth f32[B, N, H, W] for DepthLoss (data['depth'] of
    dataset/carla_dataset.py:379-423): 0.5 .. 15 m on a 1.25 cm lattice (so that many values sit
    EXACTLY on depth-bin boundaries, every 20th lattice point at 0.25 m bins), ~12 % zeros
    ("no return", ignored by the min-pool), and whole 8x8 blocks of zeros (background pixels).
    Built from integer draws and two float32 ops: bit-identical on every host."""
    rng = np.random.RandomState(2000 + seed)
    h, w = shape.final_dim
    n = (shape.batch, shape.cams, h, w)
    ds = shape.bev_down_sample
    # a base depth per ds x ds block (so the min-pool covers every bin) plus 0 .. 0.5 m of per-pixel relief
    base = rng.randint(0, 1121, size=(shape.batch, shape.cams, h // ds, w // ds))
    k = (np.repeat(np.repeat(base, ds, axis=2), ds, axis=3) + rng.randint(0, 41, size=n)).astype(np.float32)
    depth = (k * np.float32(0.0125) + np.float32(0.5)).astype(np.float32)
    depth[rng.randint(0, 100, size=n) < 12] = 0.0
    blocks = rng.randint(0, 100, size=(shape.batch, shape.cams, h // ds, w // ds)) < 7
    depth[np.repeat(np.repeat(blocks, ds, axis=2), ds, axis=3)] = 0.0
    # a band of far depths whose min-pool lands beyond the last bin (label 0 through the range test)
    depth[:, :, : 2 * ds, :] = np.where(depth[:, :, : 2 * ds, :] > 0, depth[:, :, : 2 * ds, :] + np.float32(12.5), 0.0)
    return torch.from_numpy(depth)
