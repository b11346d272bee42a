#!/usr/bin/env python
"""Row-reuse statistics of the lift-splat's two gathers, from the CPU oracle (no GPU).

    python tools/reuse_stats.py [--batch 4] [--workload cfg2|stress] [out.md]

DESIGN.md section 6 argues that the forward splat and the backward gather are bound by one 256-byte row
per kept point moving from L2 to an SM, and that lowering this traffic needs another blocking.  This
script measures, on the bench's own synthetic rig (rig B: CARLA rig with per-sample jitter, seed 1),
what each blocking could reuse:

* forward: the splat sums `prob * feat_row(pixel)` per cell; a CTA owns a BEV tile.  A ray (pixel)
  whose consecutive depth bins fall into the same tile could fetch its feature row once.  Reported:
  records per distinct (tile, pixel) pair for 1 x 128 strips (the shipped tiling) and square-ish tiles.
* backward: a CTA owns one feature-map column (camera, image column: 32 pixels x D bins); records
  whose points fall into the same cell read the same gradient row.  Reported: records per distinct
  gradient row inside a column, inside one pixel (ray) and inside a warp-sized window of 16 bins.

Test/analysis infrastructure: imports the oracle, never the product library.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from e2e_parking_carla_b200.synthetic import LiftSplatShape, make_rig  # noqa: E402
from oracle import lift_splat_oracle as lo  # noqa: E402


def distinct_pairs(a: np.ndarray, b: np.ndarray) -> int:
    """Number of distinct (a, b) pairs of two int64 arrays."""
    return int(np.unique(a.astype(np.int64) * (int(b.max()) + 1) + b.astype(np.int64)).size)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "stress"])
    ap.add_argument("out", nargs="?")
    args = ap.parse_args()
    shape = LiftSplatShape(batch=args.batch, channels=64) if args.workload == "cfg2" else LiftSplatShape.stress(args.batch)
    intr, extr = make_rig(shape.batch, shape.cams, jitter=True, seed=1)
    res, start, dim = lo.bev_grid_params(shape.bev_x_bound, shape.bev_y_bound, shape.bev_z_bound)
    fr = lo.create_frustum(shape.d_bound, shape.final_dim, shape.bev_down_sample)
    M, t = lo.camera_transform(intr.numpy(), extr.numpy())
    vox, keep, rank = lo.voxel_index(lo.geometry(M, t, fr), start, res, dim)
    B, N, D, fh, fw = shape.batch, shape.cams, shape.depth_bins, shape.fh, shape.fw
    X, Y = int(dim[0]), int(dim[1])
    npts = N * D * fh * fw
    # point order inside a sample: (camera, depth bin, row, col) - model/bev_model.py:83
    idx = np.arange(npts)
    cam, dbin, row, col = idx // (D * fh * fw), (idx // (fh * fw)) % D, (idx // fw) % fh, idx % fw
    pixel = (cam * fh + row) * fw + col
    column = cam * fw + col
    lines = ["# Row reuse available to the two gathers (CPU oracle, `python tools/reuse_stats.py --workload %s --batch %d`)"
             % (args.workload, B), "",
             "%d samples of rig B (seed 1), %d cameras, D=%d, %dx%d feature maps, %dx%d cells: %d points per sample."
             % (B, N, D, fh, fw, X, Y, npts), ""]
    kept_total = int(keep.sum())
    bench_batch = 16 if args.workload == "cfg2" else 32
    lines += ["Kept points: %d of %d (%.1f %%) = %.0f per sample -> %.1f MB of 256-byte rows per sample and gather "
              "(x%d samples, the bench batch = %.0f MB per step and gather)."
              % (kept_total, B * npts, 100.0 * kept_total / (B * npts), kept_total / B, kept_total / B * 256 / 1e6,
                 bench_batch, kept_total / B * 256 * bench_batch / 1e6), ""]

    # ---------------- forward: records per distinct (tile, pixel) ----------------
    lines += ["## Forward splat: records per distinct (tile, pixel) pair", "",
              "| tile (x by y cells) | tiles per sample | records per distinct feature row in a tile | "
              "feature-row traffic with perfect per-tile staging (MB / sample) | distinct pixels in the heaviest tile |",
              "|---|---|---|---|---|"]
    for tx in (1, 2, 4, 8, 16, 32):
        ty = 128 // tx
        pairs, heaviest = 0, 0
        for b in range(B):
            k = keep[b]
            gx, gy = vox[b][k, 0], vox[b][k, 1]
            tile = (gx // tx) * ((Y + ty - 1) // ty) + gy // ty
            pairs += distinct_pairs(tile, pixel[k])
            key = np.unique(tile.astype(np.int64) * (N * fh * fw) + pixel[k])
            heaviest = max(heaviest, int(np.bincount((key // (N * fh * fw)).astype(np.int64)).max()))
        tiles = ((X + tx - 1) // tx) * ((Y + ty - 1) // ty)
        lines.append("| %d x %d | %d | %.2f | %.1f | %d |" % (tx, ty, tiles, kept_total / pairs, pairs / B * 256 / 1e6, heaviest))
    lines += ["", "(A distinct pixel of a tile costs 256 bytes of shared memory if its row is staged; the heaviest tile bounds the "
              "stage size: 1 000 pixels = 256 KB would not fit, which is why staging needs a per-tile pixel directory and "
              "a split of heavy tiles.)", ""]

    # ---------------- backward: records per distinct gradient row ----------------
    lines += ["## Backward gather: records per distinct gradient row", "",
              "| scope that could share a row | kept records per distinct row | gradient-row traffic if shared perfectly (MB / sample) |",
              "|---|---|---|"]
    scopes = {
        "one feature-map column (camera, image column): the CTA of `ls_bwd_gather_occ_kernel`": column,
        "one pixel (ray), all its depth bins: the half-warp": pixel,
        "one pixel, a window of 16 consecutive depth bins": pixel * ((D + 15) // 16) + dbin // 16,
        "one pixel, a batch of 8 consecutive depth bins (rows in flight together)": pixel * ((D + 7) // 8) + dbin // 8,
        "one image row of a camera (32 pixels x D bins)": cam * fh + row,
        "a whole camera": cam,
    }
    for name, scope in scopes.items():
        pairs = 0
        for b in range(B):
            k = keep[b]
            pairs += distinct_pairs(scope[k], rank[b][k])
        lines.append("| %s | %.2f | %.1f |" % (name, kept_total / pairs, pairs / B * 256 / 1e6))
    distinct_cells = sum(int(np.unique(rank[b][keep[b]]).size) for b in range(B))
    lines += ["", "Distinct cells hit per sample: %.0f of %d (%.1f %%): the compulsory gradient-row traffic is %.1f MB per sample "
              "(the algorithmic figure counts the whole %.1f MB gradient tensor)."
              % (distinct_cells / B, X * Y, 100.0 * distinct_cells / B / (X * Y), distinct_cells / B * 256 / 1e6,
                 X * Y * 256 / 1e6), ""]
    text = "\n".join(lines) + "\n"
    if args.out:
        with open(args.out, "w") as f:
            f.write(text)
    sys.stdout.write(text)


if __name__ == "__main__":
    main()
