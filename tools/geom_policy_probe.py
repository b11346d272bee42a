#!/usr/bin/env python
"""Which float32 evaluation order does torch-CUDA's get_geometry use?  (SURVEY.md 7 'hard parts')

The reference computes geom = (R.K^-1) . (u*d, v*d, d) + t with a broadcast batched matmul
(model/bev_model.py:54).  On CPU that is unfused mul/add, k ascending (pinned by the golden
fixtures).  On CUDA it goes through cuBLAS; this probe runs the reference's op chain
(oracle/torch_port.py:camera_geometry) on the GPU and compares it bit for bit with candidate
orders emulated in float64 (an fp32 FMA = exact product + add, rounded once), then reports
how many voxel ranks each candidate would flip.  Output: one JSON line.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from e2e_parking_carla_b200.synthetic import LiftSplatShape, make_rig  # noqa: E402
from oracle import lift_splat_oracle as lo  # noqa: E402
from oracle import torch_port as tp  # noqa: E402


def f32(x):
    return x.astype(np.float32)


def fma(a, b, c):
    return f32(a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64))


def main():
    shape = LiftSplatShape(batch=16, channels=4)
    intr, extr = make_rig(16, 4, jitter=True, seed=1)
    fr = torch.from_numpy(lo.create_frustum(shape.d_bound, shape.final_dim, shape.bev_down_sample))
    dev = torch.device("cuda")
    geom_gpu = tp.camera_geometry(fr.to(dev), intr.to(dev), extr.to(dev)).cpu().numpy()
    geom_cpu = tp.camera_geometry(fr, intr, extr).numpy()
    # the per-camera transform exactly as torch-CUDA computed it
    inv = torch.inverse(extr.to(dev))
    M = inv[..., :3, :3].matmul(torch.inverse(intr.to(dev))).cpu().numpy()
    t = inv[..., :3, 3].cpu().numpy()
    u, v, d = (fr[..., k].numpy()[None, None] for k in range(3))
    px, py, pz = f32(u * d), f32(v * d), d
    m = [[M[:, :, i, k][:, :, None, None, None] for k in range(3)] for i in range(3)]
    tt = [t[:, :, i][:, :, None, None, None] for i in range(3)]
    zero = np.zeros((), np.float32)
    cands = {}

    def build(fn):
        return np.stack([fn(m[i][0], m[i][1], m[i][2], tt[i]) for i in range(3)], axis=-1)

    cands["unfused_asc"] = build(lambda a, b, c, s: f32(f32(f32(f32(a * px) + f32(b * py)) + f32(c * pz)) + s))
    cands["fma_asc"] = build(lambda a, b, c, s: f32(fma(c, pz, fma(b, py, f32(a * px))) + s))
    cands["fma_desc"] = build(lambda a, b, c, s: f32(fma(a, px, fma(b, py, f32(c * pz))) + s))
    cands["fma_asc_t_fused"] = build(lambda a, b, c, s: fma(c, pz, fma(b, py, fma(a, px, s + zero))))
    cands["fma_mixed"] = build(lambda a, b, c, s: f32(f32(fma(a, px, f32(b * py)) + f32(c * pz)) + s))
    res, start, dim = lo.bev_grid_params(shape.bev_x_bound, shape.bev_y_bound, shape.bev_z_bound)
    _, _, rank_gpu = lo.voxel_index(geom_gpu, start, res, dim)
    out = {"points": int(geom_gpu.size // 3),
           "gpu_vs_cpu_coords_differ": int((geom_gpu != geom_cpu).sum()),
           "gpu_vs_cpu_rank_flips": int((rank_gpu != lo.voxel_index(geom_cpu, start, res, dim)[2]).sum())}
    for name, g in cands.items():
        _, _, r = lo.voxel_index(g, start, res, dim)
        out[name] = {"coords_differ": int((g != geom_gpu).sum()), "rank_flips": int((r != rank_gpu).sum())}
    # a slice for offline search of the evaluation order: camera (0,0), every 13th point
    sel = np.arange(0, geom_gpu[0, 0].size // 3, 13)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "geom_probe_sample.npz"), M=M[0, 0], t=t[0, 0],
                        px=np.broadcast_to(px, geom_gpu.shape[:-1])[0, 0].reshape(-1)[sel],
                        py=np.broadcast_to(py, geom_gpu.shape[:-1])[0, 0].reshape(-1)[sel],
                        pz=np.broadcast_to(pz, geom_gpu.shape[:-1])[0, 0].reshape(-1)[sel],
                        geom=geom_gpu[0, 0].reshape(-1, 3)[sel])
    print(json.dumps(out))


if __name__ == "__main__" and "--library" not in sys.argv:
    main()


def library_check():
    """ls_geometry with LS_GEOM_TORCH_CUDA against torch-CUDA's own geometry: mismatch census."""
    from e2e_parking_carla_b200 import lift_splat as ls
    from e2e_parking_carla_b200.bev_model import BevModel
    from e2e_parking_carla_b200.synthetic import make_cfg
    shape = LiftSplatShape(batch=16, channels=4)
    intr, extr = make_rig(16, 4, jitter=True, seed=1)
    dev = torch.device("cuda")
    model = BevModel(make_cfg(shape), cam_encoder=torch.nn.Identity(), geometry="torch").to(dev)
    M, t = model.camera_transform(intr.to(dev), extr.to(dev))
    fr = model.frustum.data
    ours = ls.geometry(M, t, fr, model._shape(16, 4, 4))
    ref = tp.camera_geometry(fr, intr.to(dev), extr.to(dev))
    bad = (ours != ref)
    out = {"coords": int(bad.numel()), "mismatch": int(bad.sum()),
           "per_axis": [int(bad[..., i].sum()) for i in range(3)],
           "per_cam": bad.sum(dim=(2, 3, 4, 5)).flatten().tolist()}
    idx = bad.nonzero()[:6]
    ex = []
    u, v, d = fr[..., 0], fr[..., 1], fr[..., 2]
    for b, n, dd, r, c, a in idx.tolist():
        ex.append({"b": b, "n": n, "axis": a, "m": M[b, n, a].tolist(), "t": float(t[b, n, a]),
                   "u": float(u[dd, r, c]), "v": float(v[dd, r, c]), "d": float(d[dd, r, c]),
                   "ours": float(ours[b, n, dd, r, c, a]), "ref": float(ref[b, n, dd, r, c, a]),
                   "ours_hex": ours[b, n, dd, r, c, a].cpu().numpy().tobytes().hex(),
                   "ref_hex": ref[b, n, dd, r, c, a].cpu().numpy().tobytes().hex()})
    out["examples"] = ex
    print(json.dumps(out))


if __name__ == "__main__" and "--library" in sys.argv:
    library_check()
