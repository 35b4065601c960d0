"""Developer tool: step time of the BASELINE configs that are not the bench headline (graph replays, CUDA events).
usage: python scripts/config_timing.py dense|dense1024|hires|l1|bf16 [pipe]   (BTPOST_LIB picks the library build)"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "multitask-bonetumor-yolo_b200")]
import numpy as np, torch
from btpost import Pipeline, PostConfig, PostProcessor, synth, _lib

which = sys.argv[1] if len(sys.argv) > 1 else "dense"
dev = torch.device("cuda:0")
if which == "dense":      # config 4: conf 0.001, every anchor a candidate, max_det 300, batch 128
    B, S, kw, l1 = 128, 640, dict(conf_thres=0.001), False
elif which == "dense1024":  # config 4 at its stated size: 21 504 candidates / image at 1024^2, batch 128
    B, S, kw, l1 = 128, 1024, dict(conf_thres=0.001), False
elif which == "hires":    # config 3: 1024^2, 21504 anchors, 256^2 protos
    B, S, kw, l1 = 64, 1024, dict(), False
elif which == "bf16":     # headline shapes with bfloat16 prototypes (what the reference's bf16-mixed forward produces)
    B, S, kw, l1 = 64, 640, dict(proto_bf16=True), False
else:                     # the reference's own layout: three raw maps, DFL decode in the kernel
    B, S, kw, l1 = 64, 640, dict(layout=_lib.LAYOUT_L1), True
small = synth.make_batch(synth.SynthConfig(batch=8, img_size=S, seed=20264), l1=l1)   # 8 distinct images, tiled to B
rep = lambda a: np.ascontiguousarray(np.concatenate([a] * (B // 8), 0))
gt = np.concatenate([small["det_boxes_gt"] + np.array([8 * i, 0, 0, 0, 0, 0], np.float32) for i in range(B // 8)], 0)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
pp = PostProcessor(PostConfig(batch=B, img_size=S, **kw), dev)
extra = dict(maps=[t(rep(m)) for m in small["maps"]], coeffs=t(rep(small["coeffs"]))) if l1 else {}
protos = t(rep(small["protos"]))
if which == "bf16":
    protos = protos.bfloat16()
args = (None if l1 else t(rep(small["head"])), protos, t(gt), t(rep(small["masks_gt"])), t(small["proj_weight"]),
        float(small["proj_bias"]))

def timeit(g, n=30):
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

out = pp.run(*args, **extra)
print(which, "B", B, "S", S, "n_cand", out["n_cand"][:4].tolist(), "dets", out["det_count"][:4].tolist())
us = timeit(pp.capture(*args, **extra))
print(f"whole step us: {us:.1f}  -> {B / us * 1e6:.0f} images/s")
for st in ("decode_filter", "nms_match", "masks", "masks_contract"):
    print(f"  {st:14s} us: {timeit(pp.capture(*args, stage=st, **extra)):.1f}")

# the same step with four batches in flight (own inputs, workspace and outputs each)
if len(sys.argv) > 2 and sys.argv[2] == "pipe" and not l1:
    pipe = Pipeline(PostConfig(batch=B, img_size=S, **kw), dev, depth=4, proj_weight=args[4], proj_bias=args[5], gt_rows_cap=len(gt))
    for i in range(4):
        pipe.load(i, args[0], args[1], args[2], args[3])
    pipe.fork()
    for _ in range(12): pipe.replay()
    pipe.join(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 80
    e0.record(); pipe.fork()
    for _ in range(n): pipe.replay()
    pipe.join(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    print(f"four batches in flight: {us:.1f} us/step -> {B / us * 1e6:.0f} images/s")
