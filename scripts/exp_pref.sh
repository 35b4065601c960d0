#!/bin/bash
out=gpurun_out/pref.txt
: > $out
run() { label=$1; d=$2; shift; shift; env "$@" python scripts/pipe_time.py $d 300 "$label" >> $out 2>&1; }
run pref0 5 BTPOST_A_PREF=0
run pref1 5 BTPOST_A_PREF=1
run pref2 5 BTPOST_A_PREF=2
run pref4 5 BTPOST_A_PREF=4
run pref1_c8x8 5 BTPOST_A_PREF=1 BTPOST_C_MINB=8 BTPOST_C_CTAS=8
run pref2_c8x8 5 BTPOST_A_PREF=2 BTPOST_C_MINB=8 BTPOST_C_CTAS=8
run pref2_c8x8_d6 6 BTPOST_A_PREF=2 BTPOST_C_MINB=8 BTPOST_C_CTAS=8
cat $out
