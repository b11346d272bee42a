"""Freeze the UNMODIFIED reference DepthLoss (loss/depth_loss.py) on seeded inputs.

Run in the build container (needs /root/reference):  python tests/golden/make_golden_depth_loss.py
Stores the labels the reference derives from the ground-truth depth, the loss and the
gradient w.r.t. the depth probabilities (the reference's own float32 run).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from e2e_parking_carla_b200.synthetic import LiftSplatShape, make_cfg, make_depth_labels, make_encoder_outputs  # noqa: E402

REF = "/root/reference"


def main():
    sys.path.insert(0, REF)
    from loss.depth_loss import DepthLoss  # the reference's own module
    shape = LiftSplatShape(batch=1, channels=4)
    cfg = make_cfg(shape)
    _, logits = make_encoder_outputs(shape, seed=31)
    gt = make_depth_labels(shape, seed=31)
    ref = DepthLoss(cfg)
    out = {}
    for name, dt in (("32", torch.float32),):     # the reference casts its labels to float32 (:48): no float64 run
        prob = logits.to(dt).softmax(dim=1).requires_grad_(True)
        loss = ref(prob, gt.to(dt))
        loss.backward()
        out["loss" + name] = np.float64(loss.item())
        out["grad" + name] = prob.grad.to(torch.float32).numpy()
    labels = ref.get_down_sampled_gt_depth(gt)          # one-hot without class 0
    lab = torch.where(labels.sum(1) > 0, labels.argmax(1) + 1, torch.zeros(labels.shape[0], dtype=torch.long))
    np.savez_compressed(os.path.join(HERE, "depth_loss_b1.npz"), labels=lab.numpy().astype(np.int16),
                        in_seed=31, fg=int((lab > 0).sum()), **out)
    print("loss32 %.9g fg %d of %d" % (out["loss32"], int((lab > 0).sum()), lab.numel()))


if __name__ == "__main__":
    main()
