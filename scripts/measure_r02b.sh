#!/bin/bash
# Round-2 (second half) measurement set, one GPU.  Outputs under gpurun_out/r02b_*.
o=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -2 > $o/r02b_gputests.txt
python bench.py --steps 20 --warmup 5 > $o/r02b_bench_20steps.json 2> $o/r02b_bench_20steps.err
python bench.py > $o/r02b_bench_200steps.json 2> $o/r02b_bench_200steps.err
python bench.py --workload c2 --masks-out dense --steps 20 --warmup 5 --cpu-sample 0 > $o/r02b_c2_dense.json 2>/dev/null
python bench.py --workload c2 --masks-out bits --steps 20 --warmup 5 --cpu-sample 0 > $o/r02b_c2_bits.json 2>/dev/null
for c in hires dense bf16 l1; do python scripts/config_timing.py $c pipe >> $o/r02b_configs.txt 2>&1; done
python scripts/ablate.py 5 > $o/r02b_ablate.txt 2>&1
python scripts/run_sweep.py --images 16384 > $o/r02b_sweep16k.json 2>/dev/null
python scripts/run_sweep.py --images 256 > $o/r02b_sweep256.json 2>/dev/null
cat $o/r02b_gputests.txt $o/r02b_bench_20steps.json $o/r02b_configs.txt $o/r02b_ablate.txt $o/r02b_sweep16k.json
