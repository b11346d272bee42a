#!/bin/bash
# developer A/B (GPU box): time every library variant built by tools/ab_build.sh with the same bench command.
cd "$(dirname "$0")/.."
for so in e2e_parking_carla_b200/build/variants/*.so; do
  cp "$so" e2e_parking_carla_b200/libls_b200.so
  for rep in 1 2; do
    python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-train --no-compat "$@" 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['stage_ms']
print('%-14s step %.1f us | idx %.1f sort %.1f fwd %.1f bwd %.1f (unfused: gather %.1f smbwd %.1f)' % ('$(basename $so .so)', 1e3*d['ms_per_step'], 1e3*s['index+hist'], 1e3*s['sort(scan+place)'], 1e3*s['splat_fwd'], 1e3*s['backward(gather+epilogue)'], 1e3*s['splat_bwd(transpose+gather)'], 1e3*s['softmax_bwd']))"
  done
done
