#!/bin/bash
# Developer experiment: can contract_kernel (HBM-bound) and cells_kernel (issue-bound) share an SM?  Needs `make sw`.
out=gpurun_out/coresident.txt
: > $out
run() { label=$1; shift; env "$@" python scripts/pipe_time.py ${DEPTH:-5} 300 "$label" >> $out 2>&1; }
run base            X=0
run c8x8            BTPOST_C_MINB=8 BTPOST_C_CTAS=8
run a2buf2_c8x4     BTPOST_A_NBUF=2 BTPOST_A_MINB=4 BTPOST_A_CTAS=2 BTPOST_C_MINB=8 BTPOST_C_CTAS=4
run a2buf2_c8x8     BTPOST_A_NBUF=2 BTPOST_A_MINB=4 BTPOST_A_CTAS=2 BTPOST_C_MINB=8 BTPOST_C_CTAS=8
run a2buf2_c7       BTPOST_A_NBUF=2 BTPOST_A_MINB=4 BTPOST_A_CTAS=2
run a2buf1_c8x4     BTPOST_A_CTAS=2 BTPOST_C_MINB=8 BTPOST_C_CTAS=4
run a3buf1_c8x8     BTPOST_A_CTAS=3 BTPOST_C_MINB=8 BTPOST_C_CTAS=8
run a3buf2_c8x8     BTPOST_A_NBUF=2 BTPOST_A_MINB=4 BTPOST_A_CTAS=3 BTPOST_C_MINB=8 BTPOST_C_CTAS=8
run a3buf2_c8x2     BTPOST_A_NBUF=2 BTPOST_A_MINB=4 BTPOST_A_CTAS=3 BTPOST_C_MINB=8 BTPOST_C_CTAS=2
cat $out
