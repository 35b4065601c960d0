"""Developer tool: replay a few pipelined steps (btpost.Pipeline, distinct inputs per slot) between cudaProfilerStart/Stop;
used under `ncu --replay-mode range` / `--graph-profiling graph` to read the DRAM bytes of a pipelined step.
usage: python scripts/pipe_step.py [depth] [steps]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "multitask-bonetumor-yolo_b200")]
import torch
from btpost import Pipeline, PostConfig, synth

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 5
steps = int(sys.argv[2]) if len(sys.argv) > 2 else depth
B, S = 64, 640
dev = torch.device("cuda:0")
first = synth.make_batch_device(synth.SynthConfig(batch=B, img_size=S, seed=20262), dev)
pipe = Pipeline(PostConfig(batch=B, img_size=S), dev, depth=depth, proj_weight=first["proj_weight"], proj_bias=first["proj_bias"])
for i in range(depth):
    d = first if i == 0 else synth.make_batch_device(synth.SynthConfig(batch=B, img_size=S, seed=20262, image_offset=i * B), dev)
    pipe.load(i, d["head"], d["protos"], d["det_boxes_gt"], d["masks_gt"])
torch.cuda.synchronize()
pipe.fork()
for _ in range(2 * depth):
    pipe.replay()
pipe.join()
torch.cuda.synchronize()
torch.cuda.profiler.start()
pipe.fork()
for _ in range(steps):
    pipe.replay()
pipe.join()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("det_count", pipe.procs[0].out["det_count"][:4].tolist())
