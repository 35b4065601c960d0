"""Generates tests/golden/ref_wb_payloads.json by running the UNMODIFIED reference loggers
(`/root/reference/src/multitask_logging.py::log_det_examples`, `log_seg_examples`) with `wandb.Image` replaced by a
recorder, on seeded inputs.  Only usable in the build container (needs /root/reference); the fixture travels.

    python tests/golden/make_payload_golden.py
"""
import importlib
import json
import sys
import types
from pathlib import Path

import numpy as np
import torch

OUT = Path(__file__).resolve().parent / "ref_wb_payloads.json"
sys.path.insert(0, "/root/reference/src")


class _Image:                      # stands in for wandb.Image: keeps what the reference hands over
    def __init__(self, img, boxes=None, masks=None):
        self.boxes, self.masks = boxes, masks


class _Run:
    def __init__(self):
        self.logged = {}

    def log(self, d, step=None, commit=False):
        self.logged.update(d)


class _Logger:
    def __init__(self):
        self.experiment = _Run()


def inputs(seed=7, B=3, K=12, G=3, S=64):
    g = torch.Generator().manual_seed(seed)
    counts = [K, 5, 0]
    gcounts = [2, 0, 3]
    dets = torch.zeros(B, K, 6)
    for b in range(B):
        n = counts[b]
        xy = torch.rand(n, 2, generator=g) * 40
        wh = torch.rand(n, 2, generator=g) * 20 + 1
        sc = torch.sort(torch.rand(n, generator=g), descending=True).values
        dets[b, :n] = torch.cat([xy, xy + wh, sc[:, None], torch.randint(0, 3, (n, 1), generator=g).float()], 1)
    gtb = torch.rand(B, G, 4, generator=g) * 60
    gtl = torch.randint(0, 3, (B, G), generator=g)
    seg_mask = (torch.rand(B, S, S, generator=g) > 0.5).to(torch.uint8)
    masks_gt = (torch.rand(B, 1, S, S, generator=g) > 0.5).to(torch.uint8)
    return dict(dets=dets, det_count=torch.tensor(counts, dtype=torch.int32), gt_boxes=gtb, gt_labels=gtl.int(),
                gt_count=torch.tensor(gcounts, dtype=torch.int32), seg_mask=seg_mask), masks_gt


def main():
    wandb = importlib.import_module("wandb")
    wandb.Image = _Image
    ml = importlib.import_module("multitask_logging")
    out, masks_gt = inputs()
    B = out["dets"].shape[0]
    names = {0: "benign", 1: "malignant", 2: "other"}
    preds = [out["dets"][b, :int(out["det_count"][b])] for b in range(B)]
    gts = [torch.cat([out["gt_boxes"][b, :int(out["gt_count"][b])], out["gt_labels"][b, :int(out["gt_count"][b])].float()[:, None]], 1)
           for b in range(B)]
    imgs = torch.zeros(B, 3, 64, 64)
    lg = _Logger()
    ml.log_det_examples(lg, imgs, preds, gts=gts, class_id_to_name=names, stage="val", conf_th=0.25, max_samples=8, max_boxes=4)
    det = [im.boxes for im in lg.experiment.logged["det_examples_val"]]
    # log_seg_examples thresholds logits itself: feed logits whose sign encodes the mask
    logits = (out["seg_mask"].float() * 2 - 1)[:, None]
    lg2 = _Logger()
    ml.log_seg_examples(lg2, imgs, logits, masks_gt=masks_gt, stage="val", max_samples=8)
    seg = [{k: {"mask_sum": int(v["mask_data"].sum()), "mask_sha": int(np.frombuffer(v["mask_data"].tobytes(), np.uint8).astype(np.int64).dot(
        np.arange(v["mask_data"].size) % 251)), "class_labels": {str(a): b for a, b in v["class_labels"].items()}} for k, v in im.masks.items()}
           for im in lg2.experiment.logged["seg_examples_val"]]
    def clean(o):
        if isinstance(o, dict):
            return {str(k): clean(v) for k, v in o.items()}
        if isinstance(o, (list, tuple)):
            return [clean(v) for v in o]
        return o
    OUT.write_text(json.dumps({"det": clean(det), "seg": seg}, indent=1))
    print("wrote", OUT, len(det), "det payloads")


if __name__ == "__main__":
    main()
