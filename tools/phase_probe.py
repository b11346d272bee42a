import ctypes as C, json, sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from e2e_parking_carla_b200.synthetic import LiftSplatShape
shape = LiftSplatShape(batch=16, channels=64)
st = bench.Stepper(shape, torch.float32, torch.device("cuda:0"))
for _ in range(3): st.step()
torch.cuda.synchronize()
out = (C.c_uint64 * 8)()
st.lib.ls_debug_phase_cycles(out)
for _ in range(5): st.step()
torch.cuda.synchronize()
rc = st.lib.ls_debug_phase_cycles(out)
v = [x / 5 for x in out]
ctas = 350 * 16
names = ["seg+zero", "", "", "phaseB", "phaseC", "", "", ""]
print("rc", rc)
for n, x in zip(names, v):
    if n: print("%-12s avg cycles per CTA %9.0f" % (n, x / ctas))
