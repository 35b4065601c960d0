#!/bin/bash
out=gpurun_out/l2g.txt
: > $out
run() { label=$1; d=$2; shift; shift; env "$@" python scripts/pipe_time.py $d 300 "$label" >> $out 2>&1; }
run base 5 X=0
run l2g32 5 L2G=32
run l2g128 5 L2G=128
run l2g32_c8x8 5 L2G=32 BTPOST_C_MINB=8 BTPOST_C_CTAS=8
cat $out
