"""Generates the committed golden fixtures by running the UNMODIFIED reference
(`/root/reference/src/running_main_v{2,3}.py::MultiTaskLitModel.validation_step`, loaded through
oracle/ref_shim.py) on seeded synthetic inputs.  Only usable in the build container (needs
/root/reference); the fixtures it writes travel with the repo:

    python tests/golden/make_golden.py        ->  tests/golden/ref_*.npz

Inputs are NOT stored: tests regenerate them bit-identically from btpost.synth (counter-based
generator) with the seeds recorded in each fixture.
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path[:0] = [str(ROOT), str(ROOT / "multitask-bonetumor-yolo_b200")]

from btpost import synth  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = Path(__file__).resolve().parent

CASES = [
    # name, version, batch, img, seed, conf, iou, top_k
    ("ref_v3_s160", "v3", 3, 160, 4242, None, None, None),     # v3 defaults: CONF .05, IOU .6, TOP_K 100
    ("ref_v2_s160", "v2", 3, 160, 4243, None, None, None),     # v2 defaults: TOP_K 300
    ("ref_v3_s160_loose", "v3", 2, 160, 4244, 0.01, 0.45, 20),
    ("ref_v3_s640", "v3", 1, 640, 4245, None, None, None),
    # peaked DFL logits: decoded boxes sit on the objects -> NMS clusters and many anchor<->GT confusion-matrix pairs
    ("ref_v3_s160_peaked", "v3", 6, 160, 4246, None, None, None),
    ("ref_v2_s160_peaked", "v2", 6, 160, 4247, None, None, None),
    ("ref_v3_s640_peaked", "v3", 2, 640, 4248, None, None, None),
]


def run_case(name, version, B, S, seed, conf, iou, top_k):
    cfg = synth.SynthConfig(batch=B, img_size=S, seed=seed)
    peaked = name.endswith("_peaked")
    batch = synth.make_batch(cfg, l1=True, l1_peaked=peaked)
    t = torch.from_numpy
    maps = [t(m) for m in batch["maps"]]
    res = ref_shim.run_validation_step(
        version, maps, t(batch["protos"]), t(batch["det_boxes_gt"]), t(batch["masks_gt"].astype(np.float32)),
        img_size=S, nc_det=cfg.nc, proj_weight=t(batch["proj_weight"]), proj_bias=t(np.float32(batch["proj_bias"]).reshape(1)),
        conf_th=conf, nms_iou=iou, top_k=top_k)
    mod = res["module"]
    preds, targets = res["map_update"]
    out = {"version": version, "batch": B, "img_size": S, "seed": seed, "l1_peaked": int(peaked),
           "conf_th": conf if conf is not None else mod.CONF_TH, "nms_iou": iou if iou is not None else mod.NMS_IOU,
           "top_k": top_k if top_k is not None else mod.TOP_K, "n_images": len(preds)}
    for i, (p, g) in enumerate(zip(preds, targets)):
        out[f"pred_boxes_{i}"] = p["boxes"].numpy().astype(np.float32)
        out[f"pred_scores_{i}"] = p["scores"].numpy().astype(np.float32)
        out[f"pred_labels_{i}"] = p["labels"].numpy().astype(np.int64)
        out[f"gt_boxes_{i}"] = g["boxes"].numpy().astype(np.float32)
        out[f"gt_labels_{i}"] = g["labels"].numpy().astype(np.int64)
    seg_pred, seg_gt = res["seg_update"]
    seg_pred = seg_pred.detach()
    if seg_pred.dtype.is_floating_point:      # v3 feeds probabilities; torchmetrics thresholds them at 0.5
        seg_bin = (seg_pred > 0.5)
    else:                                     # v2 feeds (probs > 0.5).int()
        seg_bin = seg_pred != 0
    out["seg_pred_bits"] = np.packbits(seg_bin.numpy().astype(np.uint8).reshape(B, -1), axis=1)
    logits = res["seg_logits"].detach().numpy().astype(np.float32).reshape(B, S, S)
    if S <= 160:
        out["seg_logits"] = logits
    else:                                     # keep the fixture small: every 7th pixel + a checksum of all bits
        out["seg_logits_sub"] = logits.reshape(B, -1)[:, ::7].copy()
        out["seg_logits_xor"] = np.bitwise_xor.reduce(logits.view(np.uint32).reshape(B, -1), axis=1)
    out["cm_pairs"] = np.asarray(res["cm_pairs"], np.int64).reshape(-1, 2)   # (pred_cls, gt_cls)
    np.savez_compressed(OUT / f"{name}.npz", **out)
    print(name, "dets/img", [len(p["scores"]) for p in preds], "cm pairs", len(res["cm_pairs"]),
          "seg px", int(seg_bin.sum()))


def pin_constants():
    """Known-answer facts about torch the oracle's constants rely on (also asserted in tests)."""
    # sigmoid(x) > 0.5 in fp32  <=>  x > 1.5 * 2^-24 : bisection over bit patterns
    lo, hi = np.float32(0.0).view(np.uint32), np.float32(1e-6).view(np.uint32)
    lo, hi = int(lo), int(hi)
    while hi - lo > 1:
        mid = (lo + hi) // 2
        x = torch.tensor(np.uint32(mid).view(np.float32))
        if bool(torch.sigmoid(x) > 0.5):
            hi = mid
        else:
            lo = mid
    print("largest x with sigmoid(x) <= 0.5:", hex(lo), np.uint32(lo).view(np.float32))
    return lo


if __name__ == "__main__":
    # oneDNN picks its 1x1-conv kernel (and with it the summation order) by thread count: with 8 threads the
    # projector conv is the sequential fma chain from the bias that the oracle pins; with 1 thread it is not
    # (1266 of 1600 logits differ by up to 1833 ulp).  The fixtures are generated with 8 threads.
    torch.set_num_threads(8)
    pin_constants()
    for c in CASES:
        run_case(*c)
