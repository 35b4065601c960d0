"""Developer tool: per-phase cycle breakdown of the NMS kernel (debug build, `make dbg`)."""
import ctypes as C, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
os.environ.setdefault("BTPOST_LIB", str(ROOT / "multitask-bonetumor-yolo_b200" / "btpost" / "libbtpost_dbg.so"))
sys.path[:0] = [str(ROOT), str(ROOT / "multitask-bonetumor-yolo_b200")]
import numpy as np, torch
from btpost import PostConfig, PostProcessor, synth, _lib

B, S = int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else 640
cfg = synth.SynthConfig(batch=B, img_size=S, seed=20262)
b = synth.make_batch(cfg)
dev = torch.device("cuda:0")
d = {k: torch.from_numpy(np.ascontiguousarray(b[k])).to(dev) for k in ("head", "protos", "det_boxes_gt", "masks_gt", "proj_weight")}
pp = PostProcessor(PostConfig(batch=B, img_size=S, nms_threads=int(sys.argv[3]) if len(sys.argv) > 3 else 0), dev)
args = (d["head"], d["protos"], d["det_boxes_gt"], d["masks_gt"], d["proj_weight"], float(b["proj_bias"]))
L = _lib.load()
buf = (C.c_ulonglong * 48)()
for _ in range(3): pp.run(*args)
torch.cuda.synchronize(); L.btpost_debug_phase_cycles(buf, 1)
n = 10
for _ in range(n): pp.run(*args)
torch.cuda.synchronize(); L.btpost_debug_phase_cycles(buf, 1)
names = {1: {0: "sort", 1: "stage window", 3: "chunks (tail mark)", 8: "chunk A", 11: "chunk B", 9: "chunk C", 10: "chunk insert", 6: "package", 7: "COCO match (other kernel)"}}
for k, nb in ((1, B),):   # the mask stage is four plain kernels now: time them with bench.py / ncu
    tot = sum(buf[k * 16 + i] for i in range(16))
    print(f"kernel {k}: total cycles/launch {tot / n:.0f}  ({tot / n / nb:.0f} per image, thread 0 of every CTA)")
    for i in range(16):
        v = buf[k * 16 + i]
        if v: print(f"   {names[k].get(i, i)!s:28s} {v / n:14.0f} cycles/launch  {100 * v / tot:5.1f}%")
