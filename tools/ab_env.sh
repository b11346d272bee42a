#!/bin/bash
# developer A/B (GPU box): time the library under several values of one runtime switch.
#   usage: tools/ab_env.sh VAR v1 v2 ... [-- bench args]
cd "$(dirname "$0")/.."
var=$1; shift
vals=()
while [ $# -gt 0 ] && [ "$1" != "--" ]; do vals+=("$1"); shift; done
[ "$1" == "--" ] && shift
for v in "${vals[@]}"; do
  for rep in 1 2; do
    env $var=$v python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-train --no-compat "$@" 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['stage_ms']
print('%-14s step %.1f us | idx %.1f sort %.1f fwd %.1f bwd %.1f (unfused: gather %.1f smbwd %.1f)' % ('$var=$v', 1e3*d['ms_per_step'], 1e3*s['index+hist'], 1e3*s['sort(scan+place)'], 1e3*s['splat_fwd'], 1e3*s['backward(gather+epilogue)'], 1e3*s['splat_bwd(transpose+gather)'], 1e3*s['softmax_bwd']))"
  done
done
