#!/bin/bash
# Round-2 evidence, run on the GPU box:  bash tools/collect_profiles.sh   (outputs under gpurun_out/)
O=gpurun_out
last() { python -c "import sys,json; print(json.dumps(json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]), indent=1))" "$1" > "$2"; }
python bench.py --steps 20 --warmup 5 > $O/bench.out 2> $O/bench.err && last $O/bench.out $O/r02_bench.json
python bench.py --steps 20 --warmup 5 --dtype bf16 --no-train --no-gpu-reference --no-cpu-baseline > $O/bf16.out 2>> $O/bench.err && last $O/bf16.out $O/r02_bench_bf16.json
python bench.py --steps 10 --warmup 3 --workload stress --no-train --no-gpu-reference --no-cpu-baseline > $O/stress.out 2>> $O/bench.err && last $O/stress.out $O/r02_bench_stress.json
python bench.py --steps 10 --warmup 3 --workload stress --dtype bf16 --no-train --no-gpu-reference --no-cpu-baseline > $O/stressb.out 2>> $O/bench.err && last $O/stressb.out $O/r02_bench_stress_bf16.json
python bench.py --steps 20 --warmup 5 --bev-format nchw --no-train --no-gpu-reference --no-cpu-baseline > $O/nchw.out 2>> $O/bench.err && last $O/nchw.out $O/r02_bench_nchw.json
python bench.py --steps 20 --warmup 5 --dtype bf16 --bev-dtype bf16 --no-train --no-gpu-reference --no-cpu-baseline > $O/bevbf16.out 2>> $O/bench.err && last $O/bevbf16.out $O/r02_bench_bf16_bev.json
python bench.py --steps 20 --warmup 5 --feat-format channels_last --no-train --no-gpu-reference --no-cpu-baseline > $O/featcl.out 2>> $O/bench.err && last $O/featcl.out $O/r02_bench_featcl.json
LS_SPLAT_OUT=bulk python bench.py --steps 20 --warmup 5 --no-train --no-gpu-reference --no-cpu-baseline > $O/bulk.out 2>> $O/bench.err && last $O/bulk.out $O/r02_bench_bulk_tma.json
LS_OVERLAP_BWD=1 python bench.py --steps 20 --warmup 5 --no-train --no-gpu-reference --no-cpu-baseline --no-compat > $O/overlap.out 2>> $O/bench.err && last $O/overlap.out $O/r02_bench_overlap_bwd.json
LS_SOFTMAX_BWD_STAGED=1 python bench.py --steps 20 --warmup 5 --no-train --no-gpu-reference --no-cpu-baseline --no-compat > $O/staged.out 2>> $O/bench.err && last $O/staged.out $O/r02_bench_staged_epilogue.json
python bench.py --impl reference --steps 3 --warmup 1 > $O/ref.out 2>> $O/bench.err && last $O/ref.out $O/r02_bench_reference.json
python bench.py --workload train --steps 20 --warmup 5 > $O/train.out 2>> $O/bench.err && last $O/train.out $O/r02_bench_train.json
python bench.py --workload train --steps 5 --warmup 3 --impl reference > $O/trainref.out 2>> $O/bench.err && last $O/trainref.out $O/r02_bench_train_reference.json
python bench.py --workload agent --steps 1000 > $O/agent.out 2>> $O/bench.err && last $O/agent.out $O/r02_bench_agent.json
python bench.py --workload agent --steps 100 --impl reference > $O/agentref.out 2>> $O/bench.err && last $O/agentref.out $O/r02_bench_agent_reference.json
python tools/timeline.py graph 2>> $O/bench.err | grep -v Warn > $O/r02_timeline.txt
PROF="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-train --no-graph"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv $PROF > /dev/null 2> $O/ncu1.err
ncu --set full --clock-control none --import-source on -k regex:^ls_ -c 14 -o $O/r02_full $PROF > /dev/null 2> $O/ncu2.err
LS_SPLAT_OUT=bulk ncu --set full --clock-control none -k regex:splat_fwd -c 1 -o $O/r02_bulk_tma $PROF > /dev/null 2> $O/ncu3.err
ncu --set full --clock-control none -k regex:"canon|transpose|gather|splat_fwd|epilogue" -c 5 -o $O/r02_nchw $PROF --bev-format nchw > /dev/null 2> $O/ncu4.err
tail -2 $O/bench.err $O/ncu2.err
ls -la $O/*.ncu-rep
