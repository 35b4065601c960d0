"""Developer tool (debug build, `make dbg`): what does each kernel cost with several batches in flight?  Captures the
step with one kernel left out (btpost_debug_skip; the buffers keep the valid data of a complete eager step) and times
the replays over distinct input sets.  usage: python scripts/ablate.py [depth]"""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
os.environ["BTPOST_LIB"] = str(ROOT / "multitask-bonetumor-yolo_b200" / "btpost" / "libbtpost_dbg.so")
sys.path[:0] = [str(ROOT), str(ROOT / "multitask-bonetumor-yolo_b200")]
import torch
from btpost import Pipeline, PostConfig, synth, _lib

B, S, depth = 64, 640, int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device("cuda:0")
L = _lib.load()
first = synth.make_batch_device(synth.SynthConfig(batch=B, img_size=S, seed=20262), dev)
pipe = Pipeline(PostConfig(batch=B, img_size=S), dev, depth=depth, proj_weight=first["proj_weight"], proj_bias=first["proj_bias"])
for i in range(depth):
    d = first if i == 0 else synth.make_batch_device(synth.SynthConfig(batch=B, img_size=S, seed=20262, image_offset=i * B), dev)
    pipe.load(i, d["head"], d["protos"], d["det_boxes_gt"], d["masks_gt"])
torch.cuda.synchronize()
# the plan cannot be left out without starving cells_kernel of its work items
names = {0: "nothing", 1: "gt_pack", 2: "decode_filter", 4: "nms", 16: "gather", 32: "match", 64: "contract", 128: "cells+finalize"}
base = None
for mask, name in names.items():
    L.btpost_debug_skip(0)
    for p, inp in zip(pipe.procs, pipe.inputs):
        p.run(inp["head"], inp["protos"], inp["det_boxes_gt"], inp["masks_gt"], pipe.proj_weight, pipe.proj_bias)   # valid data everywhere
    torch.cuda.synchronize()
    L.btpost_debug_skip(mask)
    pipe.graphs = []
    for p, inp in zip(pipe.procs, pipe.inputs):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            p.run(inp["head"], inp["protos"], inp["det_boxes_gt"], inp["masks_gt"], pipe.proj_weight, pipe.proj_bias)
        pipe.graphs.append(g)
    pipe._bind_graphs()
    L.btpost_debug_skip(0)
    n = 300
    pipe.fork()
    for _ in range(30): pipe.replay()
    pipe.join(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pipe.fork()
    for _ in range(n): pipe.replay()
    pipe.join(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    base = us if base is None else base
    print(f"without {name:16s} {us:7.1f} us/step   (marginal cost {base - us:6.1f} us)")
