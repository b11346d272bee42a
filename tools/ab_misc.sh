#!/bin/bash
# A/B builds (run on the GPU box): dead-window skip in the gather, index ILP, placement groups
run() {
  python -m e2e_parking_carla_b200.build --force > /dev/null 2>&1
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-train > gpurun_out/ab.json 2>/dev/null
  python -c "
import json; d=json.loads(open('gpurun_out/ab.json').read().strip().splitlines()[-1]); s=d['stage_ms']; print('$1', round(d['ms_per_step']*1e3,1), 'us | index', round(s['index+hist']*1e3,1), 'place', round(s['sort(scan+place)']*1e3,1), 'splat', round(s['splat_fwd']*1e3,1), 'gather', round(s['splat_bwd(transpose+gather)']*1e3,1))"
}
LS_GATHER_SKIP_DEAD=0 run "skip_dead=0"
LS_GATHER_SKIP_DEAD=1 run "skip_dead=1"
LS_IDX_ILP=1 run "idx_ilp=1"
LS_IDX_ILP=4 run "idx_ilp=4"
LS_PLACE_GROUPS=8 run "place_groups=8"
LS_PLACE_GROUPS=24 run "place_groups=24"
LS_GOCC_MINB=4 LS_GOCC_ROWS=4 run "gocc 4x4"
python -m e2e_parking_carla_b200.build --force > /dev/null 2>&1
