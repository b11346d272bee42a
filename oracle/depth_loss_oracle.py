"""TEST INFRASTRUCTURE ONLY - numpy restatement of the reference DepthLoss
(loss/depth_loss.py:18-48), the consumer of the lift-splat's ``pred_depth``.

Pinned by tests/golden/depth_loss_b1.npz, frozen from the UNMODIFIED reference
(tests/golden/make_golden_depth_loss.py imports loss/depth_loss.py where it lies);
tests/test_oracle_golden.py checks this restatement against that fixture.  Integer
decisions (min-pool, bin index) are float32 op by op; sums are float64.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def down_sampled_labels(gt, d_bound, down, depth_channels):
    """labels i64[B*N*h*w]: 0 background, k >= 1 -> bin k-1 positive  (depth_loss.py:32-46)."""
    gt = np.asarray(gt, F32)
    B, N, H, W = gt.shape
    g = gt.reshape(B * N, H // down, down, W // down, down).transpose(0, 1, 3, 2, 4).reshape(-1, down * down)
    g = np.where(g == 0.0, F32(1e5), g).min(axis=-1)
    off = F32(d_bound[0] - d_bound[2])
    v = ((g - off).astype(F32) / F32(d_bound[2])).astype(F32)
    v = np.where((v < F32(depth_channels + 1)) & (v >= F32(0.0)), v, F32(0.0))
    return np.trunc(v).astype(np.int64)


def depth_loss(prob, gt, d_bound, down):
    """(loss f64, grad wrt prob f64[BN,D,h,w], labels).  BCE with aten's log clamp at -100 and
    backward (p - y) / max((1 - p) p, 1e-12); mean over foreground pixels (depth_loss.py:21-28)."""
    prob = np.asarray(prob, np.float64)
    bn, D, fh, fw = prob.shape
    labels = down_sampled_labels(gt, d_bound, down, D)
    p = prob.transpose(0, 2, 3, 1).reshape(-1, D)
    fg = labels >= 1
    y = np.zeros_like(p)
    y[np.nonzero(fg)[0], labels[fg] - 1] = 1.0
    with np.errstate(divide="ignore"):
        lp = np.maximum(np.log(p), -100.0)
        l1p = np.maximum(np.log1p(-p), -100.0)
    elem = -(y * lp + (1.0 - y) * l1p)
    n = max(1.0, float(fg.sum()))
    loss = elem[fg].sum() / n
    g = (p - y) / np.maximum((1.0 - p) * p, 1e-12) / n
    g[~fg] = 0.0
    grad = g.reshape(bn, fh, fw, D).transpose(0, 3, 1, 2)
    return loss, grad, labels
