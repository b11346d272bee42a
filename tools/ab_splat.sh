#!/bin/bash
# A/B builds of the direct splat: records in flight per quarter-warp x register budget (run on the GPU box)
for cfg in "4 6" "4 7" "4 8" "8 4" "8 5" "6 5" "2 8"; do
  set -- $cfg
  LS_QWIN=$1 LS_SPLATD_MINB=$2 python -m e2e_parking_carla_b200.build --force > /dev/null 2>&1
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-train > gpurun_out/ab.json 2>/dev/null
  python -c "
import json; d=json.loads(open('gpurun_out/ab.json').read().strip().splitlines()[-1]); print('QWIN=$1 MINB=$2', round(d['ms_per_step']*1e3,1), 'us; splat_fwd stage', round(d['stage_ms']['splat_fwd']*1e3,1))"
done
python -m e2e_parking_carla_b200.build --force > /dev/null 2>&1
