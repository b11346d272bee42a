"""The device-side next-target estimate of the agent tick (harness/agent_target.py) against the
UNMODIFIED reference methods ``save_prev_target`` / ``get_target_point_ego_coord``
(agent/parking_agent.py:290-318).  ``agent/parking_agent.py`` imports carla and pygame, so the
two FunctionDef nodes are compiled from the file where it lies (nothing is copied) and called with
a stand-in ``self``; without the reference tree the same cases are checked against a plain
restatement of the python loop."""
import ast
import os
import types

import numpy as np
import pytest
import torch

from harness.agent_target import prev_target_point

REF_FILE = "/root/reference/agent/parking_agent.py"


def _reference_methods():
    tree = ast.parse(open(REF_FILE).read(), REF_FILE)
    want = ("save_prev_target", "get_target_point_ego_coord")
    for cls in tree.body:
        if isinstance(cls, ast.ClassDef):
            fns = [f for f in cls.body if isinstance(f, ast.FunctionDef) and f.name in want]
            if len(fns) == 2:
                ns = {"torch": torch, "np": np}
                exec(compile(ast.Module(body=fns, type_ignores=[]), REF_FILE, "exec"), ns)
                return ns, {f.name: (f.lineno, f.end_lineno) for f in fns}
    raise RuntimeError("agent methods not found")


def _loop_restatement(seg_logits, x_res, y_res, prev):
    """agent/parking_agent.py:290-318 as the python loop it is (used when the tree is absent)."""
    img = seg_logits[0].argmax(dim=0).numpy()[::-1]
    xs, ys = [], []
    for r in range(img.shape[0]):
        for c in range(img.shape[1]):
            if img[r, c] == 2:
                xs.append(r)
                ys.append(c)
    if not xs:
        return prev
    px, py = int(np.average(xs)), int(np.average(ys))
    half = img.shape[0] / 2
    return [-(px - half) * x_res, (py - half) * y_res]


def _expected(seg_logits, x_res, y_res, prev):
    if not os.path.isfile(REF_FILE):
        return _loop_restatement(seg_logits, x_res, y_res, prev)
    ns, _ = _reference_methods()
    me = types.SimpleNamespace(cfg=types.SimpleNamespace(bev_x_bound=[-10.0, 10.0, x_res], bev_y_bound=[-10.0, 10.0, y_res]),
                               pre_target_point=prev)
    me.get_target_point_ego_coord = types.MethodType(ns["get_target_point_ego_coord"], me)
    ns["save_prev_target"](me, seg_logits)
    return me.pre_target_point


def _cases():
    g = torch.Generator().manual_seed(9)
    out = {}
    out["random_200"] = torch.randn(1, 3, 200, 200, generator=g)                      # ~1/3 of the pixels per class
    blob = torch.zeros(2, 3, 200, 200)
    blob[:, 0] = 1.0
    blob[0, 2, 37:52, 120:151] = 5.0                                                  # one slot rectangle
    blob[1, 2] = 9.0                                                                  # sample 1 is never read
    out["slot_rectangle"] = blob
    one = torch.zeros(1, 3, 200, 200)
    one[0, 1] = 1.0
    one[0, 2, 199, 0] = 2.0
    out["single_pixel_corner"] = one
    none = torch.zeros(1, 3, 200, 200)
    none[0, 1] = 1.0
    out["no_slot"] = none
    ties = torch.zeros(1, 3, 200, 200)                                                # all-equal logits: argmax takes class 0
    out["ties"] = ties
    full = torch.zeros(1, 3, 200, 200)
    full[0, 2] = 1.0
    out["everything_is_slot"] = full
    two = torch.zeros(1, 3, 200, 200)
    two[0, 0] = 1.0
    two[0, 2, 3:5, 7:9] = 2.0
    two[0, 2, 180:190, 100:111] = 2.0                                                 # odd sums: truncation matters
    out["two_blobs"] = two
    out["stress_400"] = torch.randn(1, 3, 400, 400, generator=g)
    return out


@pytest.mark.parametrize("name", list(_cases()))
@pytest.mark.parametrize("res", [(0.1, 0.1), (0.05, 0.05), (0.1, 0.2)])
def test_prev_target_point_matches_the_reference_methods(name, res):
    seg = _cases()[name]
    for prev in (None, [1.25, -3.5]):
        want = _expected(seg.clone(), res[0], res[1], prev)
        prev_t = None if prev is None else torch.tensor(prev, dtype=torch.float32)
        got, found = prev_target_point(seg, res[0], res[1], prev_t)
        assert got.dtype == torch.float32 and tuple(got.shape) == (2,) and found.dtype == torch.bool
        if want is None:                               # nothing found and nothing remembered
            assert not bool(found) and torch.equal(got, torch.zeros(2))
            continue
        assert bool(found) == (want is not prev)
        # what the agent feeds the model next tick: torch.tensor(target_point, dtype=torch.float) (:476)
        assert torch.equal(got, torch.tensor(want, dtype=torch.float32)), (got, want)


def test_reference_lines_are_the_cited_ones():
    if not os.path.isfile(REF_FILE):
        pytest.skip("reference tree not present")
    _, lines = _reference_methods()
    assert lines == {"save_prev_target": (290, 311), "get_target_point_ego_coord": (313, 318)}


@pytest.mark.gpu
def test_prev_target_point_on_the_device_and_inside_a_graph():
    """Same bits on the GPU, no synchronisation: capturable into a CUDA graph with the rest of the tick."""
    cases = _cases()
    dev = torch.device("cuda:0")
    for name in ("random_200", "slot_rectangle", "no_slot", "two_blobs"):
        want, wf = prev_target_point(cases[name], 0.1, 0.1)
        got, gf = prev_target_point(cases[name].to(dev), 0.1, 0.1)
        assert torch.equal(got.cpu(), want) and bool(gf) == bool(wf)
    seg = cases["two_blobs"].to(dev)
    prev_target_point(seg, 0.1, 0.1)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out, found = prev_target_point(seg, 0.1, 0.1)
    seg.copy_(cases["slot_rectangle"][:1].to(dev))
    g.replay()
    torch.cuda.synchronize()
    want, _ = prev_target_point(cases["slot_rectangle"], 0.1, 0.1)
    assert torch.equal(out.cpu(), want) and bool(found)
