"""W&B payload builders: the host-side formatting `log_det_examples` / `log_seg_examples` do on the hot
path's outputs (`/root/reference/src/multitask_logging.py:80-256`), from the padded batch tensors of `PostProcessor.run`
instead of ragged per-image lists: ONE device->host copy for the selected images, no per-box `.cpu()`.

The functions return exactly the dicts the reference hands to `wandb.Image(img, boxes=...)` / `wandb.Image(img, masks=...)`
(`multitask_logging.py:140-170`, `:111-127`); building the `wandb.Image` and calling `run.log` stays with the caller, so
this module does not import wandb.
"""
from __future__ import annotations

from typing import Mapping, Sequence

import torch


def boxes_to_wb(boxes_xyxy, scores, labels, class_id_to_name: Mapping[int, str], caption_prefix: str = "pred"):
    """`_tensor_boxes_to_wb` (`multitask_logging.py:140-170`) on host lists."""
    out = []
    for (x1, y1, x2, y2), s, l in zip(boxes_xyxy, scores, labels):
        l = int(l)
        out.append({
            "position": {"minX": x1, "minY": y1, "maxX": x2, "maxY": y2},
            "class_id": l,
            "domain": "pixel",
            "scores": {"conf": float(s)},
            "box_caption": f"{caption_prefix} {class_id_to_name.get(l, str(l))} {s:.2f}",
        })
    return out


@torch.no_grad()
def det_box_payloads(out: dict, indices: Sequence[int], class_id_to_name: Mapping[int, str], conf_th: float = 0.25,
                     max_boxes: int = 100, with_gt: bool = True):
    """Per selected image the `boxes=` payload of `log_det_examples` (`multitask_logging.py:208-250`): predictions with
    score > conf_th, first `max_boxes` (they are already in descending score order), and the GT boxes with score 1.
    `out` = outputs of `PostProcessor.run` (dets [B,K,6], det_count, gt_boxes [B,G,4], gt_labels, gt_count)."""
    idx = torch.as_tensor(list(indices), dtype=torch.long, device=out["dets"].device)
    dets = out["dets"].index_select(0, idx).cpu()
    cnt = out["det_count"].index_select(0, idx).cpu().tolist()
    if with_gt:
        gtb = out["gt_boxes"].index_select(0, idx).cpu()
        gtl = out["gt_labels"].index_select(0, idx).cpu()
        gcnt = out["gt_count"].index_select(0, idx).cpu().tolist()
    payloads = []
    for j in range(len(cnt)):
        d = dets[j, :cnt[j]]
        d = d[d[:, 4] > conf_th][:max_boxes]
        p = {"pred": {"box_data": boxes_to_wb(d[:, :4].tolist(), d[:, 4].tolist(), d[:, 5].tolist(), class_id_to_name),
                      "class_labels": class_id_to_name}}
        if with_gt:
            g = min(gcnt[j], max_boxes)
            p["gt"] = {"box_data": boxes_to_wb(gtb[j, :g].tolist(), [1.0] * g, gtl[j, :g].tolist(), class_id_to_name, "gt"),
                       "class_labels": class_id_to_name}
        payloads.append(p)
    return payloads


@torch.no_grad()
def seg_mask_payloads(out: dict, indices: Sequence[int], masks_gt=None):
    """Per selected image the `masks=` payload of `log_seg_examples` (`multitask_logging.py:111-127`): the thresholded
    projector mask (`seg_mask` output: sigmoid(logit) > 0.5, needs `PostConfig.with_seg_mask`) and, optionally, the GT."""
    if "seg_mask" not in out:
        raise ValueError("seg_mask_payloads needs PostConfig(with_seg_mask=True)")
    idx = torch.as_tensor(list(indices), dtype=torch.long, device=out["seg_mask"].device)
    pred = out["seg_mask"].index_select(0, idx).cpu().numpy()
    gt = masks_gt.index_select(0, idx.to(masks_gt.device)).cpu().byte().numpy() if masks_gt is not None else None
    payloads = []
    for j in range(len(indices)):
        p = {"prediction": {"mask_data": pred[j], "class_labels": {1: "pred"}}}
        if gt is not None:
            p["ground_truth"] = {"mask_data": gt[j, 0] if gt.ndim == 4 else gt[j], "class_labels": {1: "gt"}}
        payloads.append(p)
    return payloads
