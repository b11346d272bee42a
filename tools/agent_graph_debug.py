#!/usr/bin/env python
"""Which stage of the agent tick refuses CUDA-graph capture?  (developer aid)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from harness.parking_stack import ParkingStack, default_cfg, synthetic_batch  # noqa: E402

dev = torch.device("cuda")
cfg = default_cfg(dev)
model = ParkingStack(cfg).to(dev).eval()
data = synthetic_batch(cfg, 1, dev)
data["gt_control"] = data["gt_control"][:, :1]


def stages():
    bev, depth = model.bev_model(data["image"], data["intrinsics"], data["extrinsics"])
    yield "bev_model", bev
    bev, tmap = model.add_target(bev, data["target_point"])
    yield "add_target", bev
    tok = model.bev_encoder(bev)
    yield "bev_encoder", tok
    fused = model.fusion(tok, data["ego_motion"])
    yield "fusion", fused
    seg = model.seg_head(fused)
    yield "seg_head", seg
    t = model.control.predict(fused, data["gt_control"])
    yield "control.predict", t


with torch.no_grad():
    for _ in range(3):
        for _n, _v in stages():
            pass
    torch.cuda.synchronize()
    names = [n for n, _ in stages()]
    for upto in range(1, len(names) + 1):
        for mode in ("thread_local",):
            g = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(g, capture_error_mode=mode):
                    for i, (n, v) in enumerate(stages()):
                        if i + 1 == upto:
                            break
                g.replay()
                torch.cuda.synchronize()
                print("capture up to %-16s [%s]: ok" % (names[upto - 1], mode))
            except Exception as exc:
                print("capture up to %-16s [%s]: FAILED %s" % (names[upto - 1], mode, repr(exc)[:300]))
                torch.cuda.synchronize()
                raise SystemExit(0)
