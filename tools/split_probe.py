#!/usr/bin/env python
"""Does running the step as two half-batches on two streams overlap complementary kernels
(ALU-bound index / L1-bound splat / DRAM-bound gather)?  Developer probe (GPU box)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from e2e_parking_carla_b200.synthetic import LiftSplatShape  # noqa: E402

dev = torch.device("cuda:0")
st = bench.Stepper(LiftSplatShape(batch=16, channels=64), torch.float32, dev)
ls, lib = st.ls, st.lib
P = st._p


def make(groups):
    sizes = [16 // groups] * groups
    los = [sum(sizes[:i]) for i in range(groups)]
    shapes = [ls.make_shape(sz, 4, 48, 32, 32, 64, st.grid, 0, st.tile_x) for sz in sizes]
    scr = [torch.empty(ls.scratch_bytes(s, st.code, True), dtype=torch.uint8, device=dev) for s in shapes]
    sav = [torch.empty(ls.saved_bytes(s, st.code, st.layout), dtype=torch.uint8, device=dev) for s in shapes]
    streams = [torch.cuda.Stream() for _ in range(groups)]
    d, n = st.dev, 4

    def step(fwd_then_bwd=True):
        main = torch.cuda.current_stream()
        for k, (lo, sz) in enumerate(zip(los, sizes)):
            hi = lo + sz
            s = streams[k]
            s.wait_stream(main)
            with torch.cuda.stream(s):
                cs = C.c_void_p(s.cuda_stream)
                lib.ls_camera_transform(P(d["intr"][lo:hi]), P(d["extr"][lo:hi]), sz * n, P(st.M[lo:hi]), P(st.t[lo:hi]), cs)
                lib.ls_forward(P(d["feat"][lo * n:hi * n]), st.layout, P(d["logits"][lo * n:hi * n]), st.code, P(st.M[lo:hi]),
                               P(st.t[lo:hi]), P(st.frustum), C.byref(shapes[k]), P(scr[k]), scr[k].numel(), P(sav[k]),
                               sav[k].numel(), P(st.bev[lo:hi]), C.byref(st.st), P(st.prob[lo * n:hi * n]), cs)
                lib.ls_backward(P(d["gbev"][lo:hi]), C.byref(st.gst), P(d["gprob"][lo * n:hi * n]), P(st.prob[lo * n:hi * n]),
                                P(d["feat"][lo * n:hi * n]), st.layout, st.code, C.byref(shapes[k]), P(scr[k]), scr[k].numel(),
                                P(sav[k]), sav[k].numel(), P(st.gfeat[lo * n:hi * n]), P(st.glogits[lo * n:hi * n]), cs)
        for s in streams:
            main.wait_stream(s)
    return step


for groups in (1, 2, 4):
    step = make(groups)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        step()
    for _ in range(5):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(30):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print("groups=%d  %.1f us/step" % (groups, e0.elapsed_time(e1) / 30 * 1e3))
