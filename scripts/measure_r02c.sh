#!/bin/bash
# Round-2 (third part) measurement set, one GPU.  Outputs under gpurun_out/r02c_*.
o=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -2 > $o/r02c_gputests.txt
python bench.py --steps 20 --warmup 5 > $o/r02c_bench_20steps.json 2> $o/r02c_bench_20steps.err
python bench.py > $o/r02c_bench_200steps.json 2> $o/r02c_bench_200steps.err
rm -f $o/r02c_depths.txt
for d in 4 6 10; do echo "depth $d" >> $o/r02c_depths.txt; python bench.py --steps 20 --warmup 5 --pipeline $d --no-e2e --cpu-sample 0 >> $o/r02c_depths.txt 2>/dev/null; done
python scripts/phase_timing.py 64 640 1024 > $o/r02c_phase_1024.txt 2>&1
python scripts/phase_timing.py 64 640 512 > $o/r02c_phase_512.txt 2>&1
rm -f $o/r02c_configs.txt
for c in hires dense; do python scripts/config_timing.py $c pipe >> $o/r02c_configs.txt 2>&1; done
python scripts/run_sweep.py --images 16384 > $o/r02c_sweep16k.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/r02c_launches.csv python bench.py --steps 6 --warmup 3 --no-e2e --cpu-sample 0 > $o/r02c_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nms_kernel -c 2 -o $o/r02c_nms -f python scripts/one_step.py 64 640 3 512 > $o/r02c_ncu_nms.log 2>&1
cat $o/r02c_gputests.txt $o/r02c_depths.txt $o/r02c_phase_1024.txt $o/r02c_configs.txt
