#!/bin/bash
# one bench line of the current tree, compact (GPU box)
for i in 1 2; do
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-train > gpurun_out/ab.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/ab.json').read().strip().splitlines()[-1]); s=d['stage_ms']; print(round(d['ms_per_step']*1e3,1), 'us | index', round(s['index+hist']*1e3,1), 'place', round(s['sort(scan+place)']*1e3,1), 'splat', round(s['splat_fwd']*1e3,1), 'bwd', round(s['backward(gather+epilogue)']*1e3,1), '(unfused gather', round(s['splat_bwd(transpose+gather)']*1e3,1), 'smbwd', round(s['softmax_bwd']*1e3,1), ')', 'cached', round(d['static_rig_cache']['ms_per_step']*1e3,1))"
done
