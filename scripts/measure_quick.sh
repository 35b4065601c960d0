#!/bin/bash
# Quick check of a kernel change on one GPU: NMS edge + parity tests, phase counters, pipelined step, bench at the driver's settings.
o=gpurun_out; t=${1:-q}
python -m pytest tests/test_gpu_nms_edge.py tests/test_gpu_parity.py tests/test_gpu_pipeline.py -m gpu -q -x 2>&1 | tail -15 > $o/${t}_tests.txt
python scripts/phase_timing.py > $o/${t}_phase.txt 2>&1
python scripts/phase_timing.py 64 640 512 > $o/${t}_phase512.txt 2>&1
python scripts/pipe_time.py > $o/${t}_pipe.txt 2>&1
for i in 1 2 3; do python bench.py --steps 20 --warmup 5 --no-e2e --cpu-sample 0 >> $o/${t}_bench20.json 2>/dev/null; done
cat $o/${t}_tests.txt $o/${t}_phase.txt $o/${t}_pipe.txt; python - <<PY
import json
for l in open("$o/${t}_bench20.json"):
    if l.startswith("{"):
        d=json.loads(l); print(round(d["value"]), d["ms_per_step"], d["pipeline"]["frac_of_peak"], d["pipeline"]["ms_per_step_one_batch_in_flight"], d["pipeline"]["stage_ms"])
PY
