"""Freeze outputs of the UNMODIFIED reference lift-splat as test fixtures.

Run in the build container (needs /root/reference):
    python tests/golden/make_golden.py
It imports the reference through ``oracle/ref_harness.py`` (stub modules for the three
absent third-party imports, preset encoder outputs), runs its own ``get_geometry``,
``encoder_forward``, ``proj_bev_feature`` (+ autograd backward) on CPU and stores:

  * the rig (intrinsics, extrinsics) and the reference's own ``M = R.K^-1`` and ``t``
    (torch.inverse on CPU = MKL LAPACK);
  * voxel rank per point (-1 dropped), int32 - bit-exact target for ls_index;
  * BEV features from the reference run in float64, prob, and the
    gradients w.r.t. feat / depth logits for a fixed upstream gradient;
  * the reference's own float32 result error vs its float64 run (noise floor);
  * sha256 of the regenerated inputs so drift of the input generator is detected.

The reference ships no tests or golden vectors for this path; these files
are the pin for both the oracle (tests/test_oracle_golden.py) and the CUDA kernels.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from e2e_parking_carla_b200.synthetic import (LiftSplatShape, make_cfg, make_encoder_outputs,  # noqa: E402
                                              make_rig, make_upstream_grads)
from oracle import ref_harness as rh  # noqa: E402

CASES = {
    # name: (shape, jitter, rig seed, input seed)
    "rigA_b1_c4": (LiftSplatShape(batch=1, channels=4), False, 0, 1),
    "rigB_b2_c4": (LiftSplatShape(batch=2, channels=4), True, 11, 2),
    "stress_b1_c2": (LiftSplatShape(batch=1, cams=6, channels=2, bev_x_bound=[-10.0, 10.0, 0.05],
                                    bev_y_bound=[-10.0, 10.0, 0.05], d_bound=[0.5, 12.5, 0.125]), True, 12, 3),
}


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


def relerr(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm())


def main():
    torch.manual_seed(0)
    for name, (shape, jitter, rig_seed, in_seed) in CASES.items():
        cfg = make_cfg(shape)
        intr, extr = make_rig(shape.batch, shape.cams, jitter=jitter, seed=rig_seed)
        feat, logits = make_encoder_outputs(shape, seed=in_seed)
        gb, gp = make_upstream_grads(shape, seed=in_seed)
        model = rh.reference_bev_model(cfg)
        inv = torch.inverse(extr)
        M = inv[..., :3, :3].matmul(torch.inverse(intr)).contiguous()
        t = inv[..., :3, 3].contiguous()
        geom, vox, keep, ranks = rh.reference_indices(cfg, intr, extr)
        dim = model.bev_dim
        rank = (vox[..., 0] * (dim[1] * dim[2]) + vox[..., 1] * dim[2] + vox[..., 2])
        rank = torch.where(keep, rank, torch.full_like(rank, -1)).to(torch.int32)
        r64 = rh.run_reference(cfg, feat, logits, intr, extr, double=True, backward_with=(gb, gp))
        r32 = rh.run_reference(cfg, feat, logits, intr, extr, double=False, backward_with=(gb, gp))
        dstride = 1 if name.startswith("rigA") else 4
        out = {
            "intrinsics": intr.numpy(), "extrinsics": extr.numpy(),
            "M_ref": M.numpy(), "t_ref": t.numpy(),
            "frustum_sha": np.array(sha(model.frustum.detach())),
            "rank_ref": rank.numpy(),
            "kept_per_cam": keep.view(shape.batch, shape.cams, -1).sum(-1).numpy().astype(np.int64),
            "segments": np.array([int(r.unique().numel()) for r in ranks], np.int64),
            "sorted_rank_sha": np.array([hashlib.sha256(r.numpy().tobytes()).hexdigest() for r in ranks]),
            "geom_sha": np.array(sha(geom)),
            # depth-major tensors are stored every `dstride`-th bin to keep fixtures small
            "dstride": np.array(dstride),
            "bev_ref64": r64["bev"].numpy(), "prob_ref": r64["prob"][:, ::dstride].numpy(),
            "grad_feat_ref64": r64["grad_feat"].numpy(),
            "grad_logits_ref64": r64["grad_logits"][:, ::dstride].numpy(),
            "ref32_vs_ref64_bev": np.array(relerr(r32["bev"], r64["bev"])),
            "feat_sha": np.array(sha(feat)), "logits_sha": np.array(sha(logits)),
            "grad_bev_sha": np.array(sha(gb)), "grad_prob_sha": np.array(sha(gp)),
            "rig_seed": np.array(rig_seed), "in_seed": np.array(in_seed), "jitter": np.array(jitter),
            "torch_version": np.array(torch.__version__),
        }
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print("%s: kept %s segments %s  ref32-vs-ref64 %.2e  -> %s (%.1f KB)" % (
            name, out["kept_per_cam"].sum(-1).tolist(), out["segments"].tolist(),
            float(out["ref32_vs_ref64_bev"]), path, os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
