"""v3 segmentation-mAP prep (`/root/reference/src/running_main_v3.py:478-498`): the reference hands torchmetrics
ONE predicted mask (sigmoid > 0.5 of the projector logits), its score (mean foreground probability) and ONE target
mask per image, class 0.  On the device all of that reduces to numbers the mask kernel already leaves behind --
`seg_img3` (inter, |P|, |G|) and `seg_prob_sum` -- so the prep is a handful of tensor ops on [B] vectors; the result
has the layout of the detection outputs with K = G = 1 and goes through the same `SweepState` (nc = 1)."""
from __future__ import annotations

import torch


@torch.no_grad()
def seg_map_outputs(out: dict, iou_thrs) -> dict:
    img3 = out["seg_img3"].to(torch.float64)
    inter, P, G = img3[:, 0], img3[:, 1], img3[:, 2]
    B, dev = img3.shape[0], img3.device
    # score = (probs * mask).sum() / (mask.sum() + 1e-6) in fp32 (running_main_v3.py:486)
    score = out["seg_prob_sum"].to(torch.float32) / (P.to(torch.float32) + 1e-6)
    union = P + G - inter
    iou = torch.where(union > 0, inter / union.clamp(min=1.0), torch.zeros_like(union))
    lo = torch.tensor([0.0, 0.0, 32.0 ** 2, 96.0 ** 2], dtype=torch.float64, device=dev)
    hi = torch.tensor([1e10, 32.0 ** 2, 96.0 ** 2, 1e10], dtype=torch.float64, device=dev)
    gt_ig = (G[:, None] < lo) | (G[:, None] > hi)                                    # [B, A]
    dt_out = (P[:, None] < lo) | (P[:, None] > hi)
    thr = torch.tensor([min(float(t), 1 - 1e-10) for t in iou_thrs], dtype=torch.float64, device=dev)
    hit = iou[:, None] >= thr                                                        # [B, T]
    T = thr.numel()
    dt_match = hit[:, None, :].expand(B, 4, T).to(torch.int32)[..., None].contiguous()
    dt_ignore = torch.where(hit[:, None, :], gt_ig[:, :, None], dt_out[:, :, None]).to(torch.uint8)[..., None].contiguous()
    dets = torch.zeros(B, 1, 6, dtype=torch.float32, device=dev)
    dets[:, 0, 4] = score
    ones = torch.ones(B, dtype=torch.int32, device=dev)
    keep = ("seg_dice", "seg_iou", "uni_dice", "uni_iou")
    res = {k: out[k] for k in keep if k in out}
    res.update(dets=dets, det_count=ones, dt_match=dt_match, dt_ignore=dt_ignore, gt_ignore=gt_ig.to(torch.uint8)[..., None].contiguous(),
               gt_labels=torch.zeros(B, 1, dtype=torch.int32, device=dev), gt_count=ones.clone(), seg_map_score=score)
    return res
