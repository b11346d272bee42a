"""Freeze the UNMODIFIED reference ``ParkingModel.add_target_bev`` (model/parking_model.py:28-46).

Run in the build container (needs /root/reference):  python tests/golden/make_golden_target_bev.py

``model/parking_model.py`` cannot be imported here (its encoder imports need packages this image
does not have), and the method needs nothing of the class but ``self.cfg``: the generator parses the
file where it lies, compiles the ``add_target_bev`` FunctionDef node as is and calls it with a stand-in
``self``.  No reference text is copied into the repo.  For every case it stores the target points, the
seed of torch's CPU generator (the method draws its +-5 pixel noise with ``torch.rand_like``) and the
rows/columns of the stamped target map (the map is a union of axis-aligned boxes, so its row and
column occupancy per sample plus its sum pin it; the full map of the small cases is stored too).
"""
from __future__ import annotations

import ast
import os
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_FILE = "/root/reference/model/parking_model.py"


def reference_add_target_bev():
    """The reference's method as a plain function ``f(self, bev_feature, target_point)``."""
    tree = ast.parse(open(REF_FILE).read(), REF_FILE)
    for cls in tree.body:
        if isinstance(cls, ast.ClassDef) and cls.name == "ParkingModel":
            for fn in cls.body:
                if isinstance(fn, ast.FunctionDef) and fn.name == "add_target_bev":
                    mod = ast.Module(body=[fn], type_ignores=[])
                    ns = {"torch": torch}
                    exec(compile(mod, REF_FILE, "exec"), ns)
                    return ns["add_target_bev"], (fn.lineno, fn.end_lineno)
    raise RuntimeError("add_target_bev not found in %s" % REF_FILE)


def cases():
    """name -> (target points [B,3], x_res, y_res, H, W, seed)."""
    g = torch.Generator().manual_seed(5)
    out = {
        # |x|, |y| < 8 m keeps the +-4 px stamp (and its +-5 px noise) inside the grid
        "inner_b16": ((torch.rand(16, 3, generator=g) * 16 - 8), 0.1, 0.1, 200, 200, 11),
        # stamps that reach or cross a border: python slice semantics (clipping at the far border, a
        # negative start next to a positive stop selects nothing, both negative wraps around)
        "border_b12": (torch.tensor([[-9.9, -9.9, 0.], [9.9, 9.9, 0.], [-9.5, 0.0, 0.], [0.0, 9.6, 0.],
                                     [-10.4, 3.0, 0.], [3.0, -10.4, 0.], [-11.0, -11.0, 0.], [10.5, -2.0, 0.],
                                     [-9.7, 9.7, 0.], [9.95, -9.95, 0.], [-10.0, 10.0, 0.], [0.04, -0.04, 0.]]),
                       0.1, 0.1, 200, 200, 12),
        # the stress grid (0.05 m, 400 x 400) and a non-square map with different resolutions
        "stress_b8": ((torch.rand(8, 3, generator=g) * 18 - 9), 0.05, 0.05, 400, 400, 13),
        "ragged_b6": ((torch.rand(6, 3, generator=g) * 10 - 5), 0.2, 0.1, 56, 104, 14),
    }
    return out


def run_reference(fn, tp, x_res, y_res, h, w, seed, channels=2):
    cfg = types.SimpleNamespace(device=torch.device("cpu"), bev_x_bound=[0.0, 0.0, x_res], bev_y_bound=[0.0, 0.0, y_res])
    bev = torch.zeros(tp.shape[0], channels, h, w)
    torch.manual_seed(seed)
    wide, tmap = fn(types.SimpleNamespace(cfg=cfg), bev, tp.clone())
    assert wide.shape == (tp.shape[0], channels + 1, h, w) and torch.equal(wide[:, channels:], tmap)
    return tmap


def main():
    fn, lines = reference_add_target_bev()
    out = {"ref_lines": np.array(lines)}
    for name, (tp, xr, yr, h, w, seed) in cases().items():
        tmap = run_reference(fn, tp, xr, yr, h, w, seed)[:, 0]
        out[name + "_target"] = tp.numpy()
        out[name + "_meta"] = np.array([xr, yr, h, w, seed], dtype=np.float64)
        out[name + "_rows"] = np.packbits(tmap.amax(2).numpy().astype(np.uint8), axis=1)
        out[name + "_cols"] = np.packbits(tmap.amax(1).numpy().astype(np.uint8), axis=1)
        out[name + "_sum"] = tmap.sum((1, 2)).numpy().astype(np.int32)
        print(name, "ones per sample:", out[name + "_sum"].tolist())
    np.savez_compressed(os.path.join(HERE, "target_bev.npz"), **out)


if __name__ == "__main__":
    main()
