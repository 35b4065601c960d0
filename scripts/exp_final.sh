#!/bin/bash
out=gpurun_out/final_cfg.txt
: > $out
run() { label=$1; d=$2; shift; shift; env "$@" python scripts/pipe_time.py $d 300 "$label" >> $out 2>&1; }
run nt512_d5 5 X=0
run nt256_d5 5 NMS_NT=256
run nt512_d4 4 X=0
run nt512_d6 6 X=0
run nt256_d6 6 NMS_NT=256
run nt512_d10 10 X=0
cat $out
