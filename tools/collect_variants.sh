#!/bin/bash
# The variant bench lines of tools/collect_profiles.sh only (GPU box), most important first.
O=gpurun_out
last() { python -c "import sys,json; print(json.dumps(json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]), indent=1))" "$1" > "$2"; }
Q="--no-train --no-gpu-reference --no-cpu-baseline"
python bench.py --steps 20 --warmup 5 --dtype bf16 $Q > $O/bf16.out 2>> $O/bench.err && last $O/bf16.out $O/r02_bench_bf16.json
python bench.py --steps 10 --warmup 3 --workload stress $Q > $O/stress.out 2>> $O/bench.err && last $O/stress.out $O/r02_bench_stress.json
python bench.py --steps 20 --warmup 5 --bev-format nchw $Q > $O/nchw.out 2>> $O/bench.err && last $O/nchw.out $O/r02_bench_nchw.json
python bench.py --steps 20 --warmup 5 --feat-format channels_last $Q > $O/featcl.out 2>> $O/bench.err && last $O/featcl.out $O/r02_bench_featcl.json
python bench.py --steps 20 --warmup 5 --dtype bf16 --bev-dtype bf16 $Q > $O/bevbf16.out 2>> $O/bench.err && last $O/bevbf16.out $O/r02_bench_bf16_bev.json
python bench.py --steps 10 --warmup 3 --workload stress --dtype bf16 $Q > $O/stressb.out 2>> $O/bench.err && last $O/stressb.out $O/r02_bench_stress_bf16.json
LS_OVERLAP_BWD=1 python bench.py --steps 20 --warmup 5 $Q --no-compat > $O/overlap.out 2>> $O/bench.err && last $O/overlap.out $O/r02_bench_overlap_bwd.json
LS_SOFTMAX_BWD_STAGED=1 python bench.py --steps 20 --warmup 5 $Q --no-compat > $O/staged.out 2>> $O/bench.err && last $O/staged.out $O/r02_bench_staged_epilogue.json
LS_SPLAT_OUT=bulk python bench.py --steps 20 --warmup 5 $Q > $O/bulk.out 2>> $O/bench.err && last $O/bulk.out $O/r02_bench_bulk_tma.json
LS_GATHER_TMA=1 python bench.py --steps 20 --warmup 5 $Q --no-compat > $O/gtma.out 2>> $O/bench.err && last $O/gtma.out $O/r02_bench_gather_tma.json
ls -la $O/r02_bench_*.json | wc -l
