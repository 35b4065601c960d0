#!/bin/bash
out=gpurun_out/gs.txt
: > $out
run() { label=$1; d=$2; shift; shift; env "$@" python scripts/pipe_time.py $d 300 "$label" >> $out 2>&1; }
run gs8_mb8 5 X=0
run gs16_mb8 5 BTPOST_C_GS=16
run gs32_mb8 5 BTPOST_C_GS=32
run gs8_mb7 5 BTPOST_C_MINB=7 BTPOST_C_CTAS=7
run gs16_mb7 5 BTPOST_C_GS=16 BTPOST_C_MINB=7 BTPOST_C_CTAS=7
run gs16_mb8_d6 6 BTPOST_C_GS=16
cat $out
