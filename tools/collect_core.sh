#!/bin/bash
# The core of tools/collect_profiles.sh (GPU box): default bench line, CUPTI timeline, ncu launch list and the
# --set full capture of every ls_ kernel.  Outputs under gpurun_out/.
O=gpurun_out
last() { python -c "import sys,json; print(json.dumps(json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]), indent=1))" "$1" > "$2"; }
python bench.py --steps 20 --warmup 5 > $O/bench.out 2> $O/bench.err && last $O/bench.out $O/r02_bench.json
python tools/timeline.py graph 2>> $O/bench.err | grep -v Warn > $O/r02_timeline.txt
PROF="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-train --no-graph"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv $PROF > /dev/null 2> $O/ncu1.err
ncu --set full --clock-control none --import-source on -k regex:^ls_ -c 14 -o $O/r02_full -f $PROF > /dev/null 2> $O/ncu2.err
tail -2 $O/bench.err $O/ncu2.err; cat $O/r02_timeline.txt
