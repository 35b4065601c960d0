#!/bin/bash
out=gpurun_out/ilv.txt
: > $out
run() { label=$1; d=$2; shift; shift; env "$@" python scripts/pipe_time.py $d 300 "$label" >> $out 2>&1; }
run base 5 X=0
run ilv 5 BTPOST_A_ILV=1
run ilv_c8x8 5 BTPOST_A_ILV=1 BTPOST_C_MINB=8 BTPOST_C_CTAS=8
run ilv_pref1 5 BTPOST_A_ILV=1 BTPOST_A_PREF=1
run ilv_a3 5 BTPOST_A_ILV=1 BTPOST_A_CTAS=3
cat $out
