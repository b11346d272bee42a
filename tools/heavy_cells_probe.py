#!/usr/bin/env python
"""Forward time of the lift-splat on grids with very heavy cells (the canonical ordering must not be
quadratic): default 0.1 m grid, 1 m grid (> 1 000 points per cell), 10 m grid (> 30 000 per cell)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from e2e_parking_carla_b200 import lift_splat as ls  # noqa: E402
from e2e_parking_carla_b200.bev_model import BevModel  # noqa: E402
from e2e_parking_carla_b200.synthetic import LiftSplatShape, make_cfg, make_encoder_outputs, make_rig  # noqa: E402

dev = torch.device("cuda:0")
for res in (0.1, 1.0, 10.0):
    shape = LiftSplatShape(batch=4, channels=64, bev_x_bound=[-10.0, 10.0, res], bev_y_bound=[-10.0, 10.0, res])
    model = BevModel(make_cfg(shape), cam_encoder=torch.nn.Identity()).to(dev)
    intr, extr = make_rig(4, 4, jitter=True, seed=2)
    feat, logits = make_encoder_outputs(shape, seed=8)
    f, z = feat.to(dev), logits.to(dev)
    M, t = model.camera_transform(intr.to(dev), extr.to(dev))
    for _ in range(3):
        ls.lift_splat(f, z, M, t, model.frustum, model._grid, torch.channels_last)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        ls.lift_splat(f, z, M, t, model.frustum, model._grid, torch.channels_last)
    e1.record()
    torch.cuda.synchronize()
    rank = ls.index(M, t, model.frustum, model._shape(4, 4, 64))
    counts = torch.bincount(rank[rank >= 0].long())
    print("voxel %.1f m: %d x %d cells, max %d points per cell, forward %.3f ms (B=4)"
          % (res, model._grid.dim[0], model._grid.dim[1], int(counts.max()), e0.elapsed_time(e1) / 10))
