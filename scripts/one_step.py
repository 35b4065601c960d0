"""Developer tool: run a few plain (non-graph) steps of the hot path; used under ncu for launch lists."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "multitask-bonetumor-yolo_b200")]
import numpy as np, torch
from btpost import PostConfig, PostProcessor, synth

B, S = int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else 640
n = int(sys.argv[3]) if len(sys.argv) > 3 else 5
cfg = synth.SynthConfig(batch=B, img_size=S, seed=20262)
b = synth.make_batch(cfg)
dev = torch.device("cuda:0")
d = {k: torch.from_numpy(np.ascontiguousarray(b[k])).to(dev) for k in ("head", "protos", "det_boxes_gt", "masks_gt", "proj_weight")}
pp = PostProcessor(PostConfig(batch=B, img_size=S, nms_threads=int(sys.argv[4]) if len(sys.argv) > 4 else 0), dev)
for _ in range(n):
    pp.run(d["head"], d["protos"], d["det_boxes_gt"], d["masks_gt"], d["proj_weight"], float(b["proj_bias"]))
torch.cuda.synchronize()
print("det_count", pp.out["det_count"][:8].tolist())
