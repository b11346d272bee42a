#!/bin/bash
run() {
  python -m e2e_parking_carla_b200.build --force > /dev/null 2>&1
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-train > gpurun_out/ab.json 2>/dev/null
  python -c "
import json; d=json.loads(open('gpurun_out/ab.json').read().strip().splitlines()[-1]); s=d['stage_ms']; print('$1', round(d['ms_per_step']*1e3,1), 'us | splat', round(s['splat_fwd']*1e3,1), 'gather', round(s['splat_bwd(transpose+gather)']*1e3,1))"
}
LS_FFMA2=0 run "ffma2=0"
LS_FFMA2=1 run "ffma2=1"
LS_FFMA2=0 run "ffma2=0"
LS_FFMA2=1 run "ffma2=1"
python -m e2e_parking_carla_b200.build --force > /dev/null 2>&1
