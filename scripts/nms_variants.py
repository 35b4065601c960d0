"""Developer tool: strictly serial step and NMS stage with the 512- and the 1024-thread NMS kernel on the same inputs."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "multitask-bonetumor-yolo_b200")]
import torch
from btpost import PostConfig, PostProcessor, synth

B, S = 64, 640
dev = torch.device("cuda:0")
for off in (0, 64, 128):
    d = synth.make_batch_device(synth.SynthConfig(batch=B, img_size=S, seed=20262, image_offset=off), dev)
    args = (d["head"], d["protos"], d["det_boxes_gt"], d["masks_gt"], d["proj_weight"], d["proj_bias"])
    for nt in (1024, 512):
        pp = PostProcessor(PostConfig(batch=B, img_size=S, nms_threads=nt), dev)
        out = pp.run(*args)
        torch.cuda.synchronize()
        res = []
        for stage in ("run", "nms_match"):
            g = pp.capture(*args, stage=stage)
            for _ in range(5): g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50): g.replay()
            e1.record(); torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / 50 * 1e3)
        print(f"images {off}..{off+63}  nms_threads {nt}: step {res[0]:.1f} us, nms stage {res[1]:.1f} us, mean candidates {float(out['n_cand'].float().mean()):.0f}")
