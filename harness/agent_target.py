"""The agent's next-target estimate (agent/parking_agent.py:290-318: ``save_prev_target`` +
``get_target_point_ego_coord``) as a handful of device-side reductions.

The reference takes the arg-max class map of the predicted segmentation to the host, flips it
vertically, walks all 200 x 200 pixels in a python loop collecting the coordinates of the target-slot
class, averages them (``int(np.average(...))``: truncation) and converts the pixel to ego metres.
Here the same numbers come from two integer row / column histograms of the class map: no device->host
copy of the map, no python loop, nothing that breaks CUDA-graph capture, so the whole agent tick
(``ParkingModel.predict`` + this) replays as one graph (bench.py --workload agent).

Exactness: pixel counts and coordinate sums are int64 (exact); ``int(np.average(v))`` of non-negative
integers equals ``sum(v) // len(v)`` (the float64 quotient of two integers below 2**53 cannot round
across an integer: a non-integer quotient is at least 1 / len(v) >= 2.5e-5 away from one); the metre
conversion is evaluated in float64 like the reference's python floats and rounded to float32 once, as
``torch.tensor(target_point, dtype=torch.float)`` does (agent/parking_agent.py:476).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

TARGET_SLOT_CLASS = 2        # recoloured to 255 at agent/parking_agent.py:295 and tested at :303


def prev_target_point(pred_segmentation: torch.Tensor, x_res: float, y_res: float,
                      prev: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """pred_segmentation f32[B, classes, H, W] (only sample 0 is read, :291) ->
    (target f32[2] = ego (x, y) in metres, found bool[]).

    ``found`` is False when no pixel has the target-slot class; the reference then leaves
    ``self.pre_target_point`` untouched (:308), so ``target`` is ``prev`` (or zeros when there is none).
    Everything stays on the tensor's device; no synchronisation."""
    seg = torch.argmax(pred_segmentation[0], dim=0)                    # :291-292, [H, W] int64
    slot = seg == TARGET_SLOT_CLASS
    h, w = slot.shape
    dev = slot.device
    # the reference scans the vertically flipped image (:296): its row r is row h-1-r of the map
    flipped_rows = torch.arange(h - 1, -1, -1, device=dev, dtype=torch.int64)
    cols = torch.arange(w, device=dev, dtype=torch.int64)
    per_row = slot.sum(dim=1, dtype=torch.int64)
    per_col = slot.sum(dim=0, dtype=torch.int64)
    n = per_row.sum()
    found = n > 0
    den = n.clamp_min(1)
    px = torch.div((per_row * flipped_rows).sum(), den, rounding_mode="floor")     # int(np.average(rows)), :309
    py = torch.div((per_col * cols).sum(), den, rounding_mode="floor")             # :310
    half = h / 2                                                                   # bev_shape / 2 for BOTH axes (:314-316)
    x = -(px.to(torch.float64) - half) * x_res                                     # :315, :317
    y = (py.to(torch.float64) - half) * y_res                                      # :316, :317
    new = torch.stack([x, y]).to(torch.float32)
    if prev is None:
        prev = torch.zeros(2, dtype=torch.float32, device=dev)
    return torch.where(found, new, prev.to(torch.float32)), found
