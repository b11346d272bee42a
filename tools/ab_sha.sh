#!/bin/bash
# developer A/B (GPU box): sha256 of the gradients of one small step under every variant built by tools/ab_build.sh
cd "$(dirname "$0")/.."
for so in e2e_parking_carla_b200/build/variants/*.so; do
  cp "$so" e2e_parking_carla_b200/libls_b200.so
  python - "$(basename $so .so)" <<'PY' 2>/dev/null
import sys, hashlib, torch
sys.path.insert(0, '.')
import bench
from e2e_parking_carla_b200.synthetic import LiftSplatShape
st = bench.Stepper(LiftSplatShape(batch=3, channels=64), torch.float32, torch.device('cuda:0'))
st.step(); torch.cuda.synchronize()
h = hashlib.sha256()
for k in ('bev', 'gfeat', 'glogits'):
    h.update(getattr(st, k).float().cpu().contiguous().numpy().tobytes())
print('%-16s sha256 %s' % (sys.argv[1], h.hexdigest()[:16]))
PY
done
