#!/bin/bash
out=gpurun_out/gap.txt
: > $out
run() { label=$1; d=$2; shift; shift; env "$@" python scripts/pipe_time.py $d 300 "$label" >> $out 2>&1; }
run plain 5 X=0
run sweep 5 SWEEP=1
run sweep_fill 5 SWEEP=1 FILL=1
run fill 5 FILL=1
cat $out
python bench.py --steps 20 --warmup 5 --no-e2e --cpu-sample 0 > gpurun_out/bench20.json 2>gpurun_out/bench20.err; cat gpurun_out/bench20.json
