#!/bin/bash
out=gpurun_out/radix.txt
: > $out
run() { label=$1; d=$2; shift; shift; env "$@" python scripts/pipe_time.py $d 300 "$label" >> $out 2>&1; }
run radix 5 X=0
run radix_c8x8 5 BTPOST_C_MINB=8 BTPOST_C_CTAS=8
run radix_c8x8_d6 6 BTPOST_C_MINB=8 BTPOST_C_CTAS=8
cat $out
