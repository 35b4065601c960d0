"""TEST INFRASTRUCTURE / CPU BASELINE ONLY -- torch port of the reference's post-processing + evaluation ops.

`/root/reference` is not on the GPU box, so `bench.py --impl reference` and the `cpu_baseline` leg time THIS port: the
statements of `MultiTaskLitModel.validation_step` (`/root/reference/src/running_main_v3.py:447-575`,
`running_main_v2.py:672-892`) and of `test_model.py:80-89`, written with the same library calls the reference makes
(torch elementwise ops, `torchvision.ops.nms`, `F.conv2d`, `F.interpolate`, `torch.einsum`), on the whole batch, with
every host thread torch can use.  The input is the L2 head the benchmark workload names (`segment_preds_cat`,
`main_modelv2.py:367`), so decode is Ultralytics' xywh -> xyxy; everything after it follows the reference line by line:

  filter / NMS / packaging   running_main_v2.py:780-839   (max over classes, > CONF_TH, clamp_, torchvision nms[:TOP_K], .cpu())
  GT prep                    running_main_v2.py:842-882
  CM pairs                   running_main_v2.py:402-449 + batch_bbox_iou :68-94 (all N raw boxes vs GT)
  projector mask + counters  running_main_v2.py:689-713   (conv 1x1 -> bilinear -> sigmoid -> > 0.5 -> tp/fp/fn/tn, Dice)
  instance masks             test_model.py:80-89          (einsum -> bilinear -> sigmoid > 0.5 -> IoU / Dice), with the
                             Ultralytics crop at prototype resolution
  COCO matching              pycocotools evaluateImg is C code: the oracle's C restatement (`bto_coco_match`) is called

Numerically it agrees with the oracle up to summation order (einsum / conv): tests/test_ref_port.py.  Not used by the
product.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F
import torchvision

from . import oracle


def batch_bbox_iou(b1, b2, eps=1e-7):
    """running_main_v2.py:68-94."""
    x1 = torch.max(b1[:, None, 0], b2[None, :, 0]); y1 = torch.max(b1[:, None, 1], b2[None, :, 1])
    x2 = torch.min(b1[:, None, 2], b2[None, :, 2]); y2 = torch.min(b1[:, None, 3], b2[None, :, 3])
    inter = (x2 - x1).clamp(min=0) * (y2 - y1).clamp(min=0)
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1]); a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    return inter / (a1[:, None] + a2[None, :] - inter + eps)


@torch.no_grad()
def run_batch(batch, conf_thres=0.05, iou_thres=0.6, max_det=300, img_size=640, nc=3, iou_match_thresh=0.5, crop=True,
              with_coco=True, with_instances=True):
    """One batch of the workload through the reference's ops.  Returns the metrics a validation step would update."""
    head = torch.from_numpy(batch["head"]) if isinstance(batch["head"], np.ndarray) else batch["head"]
    protos = torch.from_numpy(batch["protos"]) if isinstance(batch["protos"], np.ndarray) else batch["protos"]
    gt_rows = torch.from_numpy(np.asarray(batch["det_boxes_gt"], np.float32))
    masks_gt = torch.from_numpy(batch["masks_gt"]) if isinstance(batch["masks_gt"], np.ndarray) else batch["masks_gt"]
    w = torch.from_numpy(np.asarray(batch["proj_weight"], np.float32)).view(1, -1, 1, 1)
    bias = torch.tensor([float(batch["proj_bias"])])
    B, _, N = head.shape
    S = img_size
    thrs = oracle.iou_thresholds()
    # ---- decode (Ultralytics xywh2xyxy on segment_preds_cat) + class max, whole batch
    xy, wh = head[:, 0:2], head[:, 2:4]
    boxes = torch.cat([xy - wh * 0.5, xy + wh * 0.5], 1).permute(0, 2, 1).contiguous()          # [B, N, 4]
    scores_all = head[:, 4:4 + nc].permute(0, 2, 1)                                              # [B, N, nc]
    coeffs = head[:, 4 + nc:].permute(0, 2, 1)                                                   # [B, N, 32]
    # ---- projector mask (running_main_v2.py:689-713)
    seg_logits = F.interpolate(F.conv2d(protos, w, bias), size=(S, S), mode="bilinear", align_corners=False)
    seg_pred = (seg_logits.sigmoid() > 0.5).int()
    gt_i = masks_gt.int()
    tp = int((seg_pred & gt_i).sum()); fp = int((seg_pred & (1 - gt_i)).sum()); fn = int(((1 - seg_pred) & gt_i).sum())
    inter = (seg_pred & gt_i).flatten(1).sum(1).float(); psum = seg_pred.flatten(1).sum(1).float(); gsum = gt_i.flatten(1).sum(1).float()
    seg_dice = (2 * inter + 1e-7) / (psum + gsum + 1e-7)
    out = {"det_count": [], "cm": torch.zeros(nc, nc, dtype=torch.int64), "seg_cnt": (tp, fp, fn, B * S * S - tp - fp - fn),
           "seg_dice": seg_dice, "uni_dice": [], "uni_iou": [], "matched": 0, "preds": []}
    rx = (S // 4) / S
    for b in range(B):                                                   # the reference's Python loop over images (:778)
        top_scores, top_labels = scores_all[b].max(dim=1)
        keep = top_scores > conf_thres
        item_boxes, item_scores, item_labels = boxes[b][keep], top_scores[keep], top_labels[keep]
        item_boxes.clamp_(0, S)
        if item_boxes.shape[0]:
            k = torchvision.ops.nms(item_boxes, item_scores, iou_thres)[:max_det]
        else:
            k = torch.zeros(0, dtype=torch.long)
        kb, ks, kl = item_boxes[k].cpu(), item_scores[k].cpu(), item_labels[k].cpu()
        out["preds"].append({"boxes": kb, "scores": ks, "labels": kl})
        out["det_count"].append(len(k))
        # GT prep (:842-882) + CM pairs on all raw boxes (:402-449)
        rows = gt_rows[gt_rows[:, 0] == b]
        if rows.shape[0]:
            cx, cy, gw, gh = rows[:, 2], rows[:, 3], rows[:, 4], rows[:, 5]
            gtb = torch.stack([(cx - gw / 2) * S, (cy - gh / 2) * S, (cx + gw / 2) * S, (cy + gh / 2) * S], 1)
            gtl = rows[:, 1].long()
            iou = batch_bbox_iou(boxes[b], gtb)
            best, arg = iou.max(dim=1)
            pos = best > iou_match_thresh
            pc, gc = scores_all[b].argmax(1)[pos], gtl[arg[pos]]
            out["cm"].index_put_((gc, pc), torch.ones_like(gc), accumulate=True)
            gtb_c = gtb.clamp(0, S)
        else:
            gtb_c, gtl = torch.zeros(0, 4), torch.zeros(0, dtype=torch.long)
        # instance masks (test_model.py:80-89) with the Ultralytics crop at prototype resolution
        if with_instances and len(k):
            cf = coeffs[b][keep][k]                                       # [K, 32]
            m = torch.einsum("qc,chw->qhw", cf, protos[b])                # [K, ph, pw]
            if crop:
                x1, y1, x2, y2 = (kb[:, i, None, None] * rx for i in range(4))
                r = torch.arange(m.shape[2], dtype=torch.float32)[None, None, :]
                c = torch.arange(m.shape[1], dtype=torch.float32)[None, :, None]
                m = m * ((r >= x1) & (r < x2) & (c >= y1) & (c < y2))
            up = F.interpolate(m[None], size=(S, S), mode="bilinear", align_corners=False)[0]
            inst = up.sigmoid() > 0.5                                     # [K, S, S] bool
            uni = inst.any(0)
        else:
            uni = torch.zeros(S, S, dtype=torch.bool)
        g = masks_gt[b, 0].bool()
        i_, u_ = float((uni & g).sum()), float((uni | g).sum())
        out["uni_iou"].append((i_ + 1e-7) / (u_ + 1e-7))
        out["uni_dice"].append((2 * i_ + 1e-7) / (float(uni.sum()) + float(g.sum()) + 1e-7))
        # COCO evaluateImg (C code in the reference's stack too)
        if with_coco:
            for cidx in range(nc):
                di, gi = torch.nonzero(kl == cidx)[:, 0], torch.nonzero(gtl == cidx)[:, 0]
                if len(di) == 0 and len(gi) == 0:
                    continue
                dm, _, _ = oracle.coco_match(kb[di].numpy(), gtb_c[gi].numpy(), thrs)
                out["matched"] += int((dm > 0).sum())
    return out
