"""Diagnostic: where does the step time go?  Times graph replays of the whole step and of each stage."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "multitask-bonetumor-yolo_b200")]
import numpy as np, torch
from btpost import PostConfig, PostProcessor, synth

B, S = int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else 640
max_det = int(sys.argv[3]) if len(sys.argv) > 3 else 300
cfg = synth.SynthConfig(batch=B, img_size=S, seed=20262)
b = synth.make_batch(cfg)
dev = torch.device("cuda:0")
d = {k: torch.from_numpy(np.ascontiguousarray(b[k])).to(dev) for k in ("head", "protos", "det_boxes_gt", "masks_gt", "proj_weight")}
pp = PostProcessor(PostConfig(batch=B, img_size=S, max_det=max_det), dev)
args = (d["head"], d["protos"], d["det_boxes_gt"], d["masks_gt"], d["proj_weight"], float(b["proj_bias"]))

def timeit(g, n=50):
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

print("n_cand", pp.run(*args)["n_cand"][:8].tolist(), "det_count", pp.out["det_count"][:8].tolist())
print("whole step  us:", timeit(pp.capture(*args)))
for st in ("decode_filter", "nms_match", "masks"):
    print(f"{st:14s} us:", timeit(pp.capture(*args, stage=st)))
