#!/bin/bash
out=gpurun_out/nms256.txt
: > $out
run() { label=$1; d=$2; shift; shift; env "$@" python scripts/pipe_time.py $d 300 "$label" >> $out 2>&1; }
run base_d5 5 X=0
run nt256_d5 5 NMS_NT=256
run nt256_d6 6 NMS_NT=256
run nt256_d4 4 NMS_NT=256
run nt256_c8x8_d5 5 NMS_NT=256 BTPOST_C_MINB=8 BTPOST_C_CTAS=8
run nt256_c8x8_d6 6 NMS_NT=256 BTPOST_C_MINB=8 BTPOST_C_CTAS=8
run nt1024_d5 5 NMS_NT=1024
cat $out
