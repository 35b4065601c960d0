#!/bin/bash
# Round-2 (third part): state check of HEAD on one GPU.  Outputs under gpurun_out/r02c_*.
o=gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -3 > $o/r02c_gputests.txt
python bench.py --steps 20 --warmup 5 > $o/r02c_bench_20steps.json 2> $o/r02c_bench_20steps.err
for i in 1 2 3; do python bench.py --steps 20 --warmup 5 --no-e2e --cpu-sample 0 >> $o/r02c_bench_20steps_rep.json 2>/dev/null; done
python scripts/pipe_time.py > $o/r02c_pipe_time.txt 2>&1
python scripts/phase_timing.py > $o/r02c_phase.txt 2>&1
cat $o/r02c_gputests.txt $o/r02c_bench_20steps.json $o/r02c_bench_20steps_rep.json $o/r02c_pipe_time.txt; tail -20 $o/r02c_phase.txt
