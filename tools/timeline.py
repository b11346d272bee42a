"""Kernel timeline of the device-resident step from CUPTI (torch.profiler): per-kernel in-situ
durations and the idle gaps between kernels.  python tools/timeline.py [graph]"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from e2e_parking_carla_b200.synthetic import LiftSplatShape

use_graph = len(sys.argv) > 1 and sys.argv[1] == "graph"
st = bench.Stepper(LiftSplatShape(batch=16, channels=64), torch.float32, torch.device("cuda:0"))
for _ in range(5):
    st.step()
g = st.capture() if use_graph else None
torch.cuda.synchronize()
STEPS = 10
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(STEPS):
        g.replay() if use_graph else st.step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0, t1 = ev[0].time_range.start, max(e.time_range.end for e in ev)
print("events", len(ev), "span per step %.1f us" % ((t1 - t0) / STEPS))
dur, cnt = {}, {}
for e in ev:
    n = e.name.split("(")[0].replace("void ", "")[:40]
    dur[n] = dur.get(n, 0.0) + (e.time_range.end - e.time_range.start)
    cnt[n] = cnt.get(n, 0) + 1
for n in dur:
    print("%-42s n=%3d  mean %.1f us" % (n, cnt[n], dur[n] / cnt[n]))
# union of busy intervals -> idle time
busy, cur_s, cur_e = 0.0, None, None
for e in ev:
    s, t = e.time_range.start, e.time_range.end
    if cur_e is None or s > cur_e:
        if cur_e is not None:
            busy += cur_e - cur_s
        cur_s, cur_e = s, t
    else:
        cur_e = max(cur_e, t)
busy += cur_e - cur_s
print("busy per step %.1f us, idle per step %.1f us, sum of kernel durations per step %.1f us" %
      (busy / STEPS, (t1 - t0 - busy) / STEPS, sum(dur.values()) / STEPS))
