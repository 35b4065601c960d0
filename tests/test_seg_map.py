"""v3 segmentation-mAP prep (`/root/reference/src/running_main_v3.py:478-498`): the oracle's restatement
against the reference's own expressions written in torch, the record builder against a plain-Python COCOeval of
one mask pair per image, the device-side prep (`btpost.segmap`) against both, and -- on the GPU -- the mask kernel's
`seg_prob_sum` against the oracle (1e-5 relative: a float metric)."""
import numpy as np
import pytest
import torch

import helpers
from btpost.segmap import seg_map_outputs
from btpost.sweep import SweepState
from oracle import oracle


def reference_expressions(logits, masks_gt):
    """running_main_v3.py:463-486, verbatim in torch, for one batch of upsampled logits."""
    seg_probs = torch.from_numpy(logits)[:, None].sigmoid()
    scores, preds = [], []
    for i in range(seg_probs.shape[0]):
        pred_mask_bool = (seg_probs[i] > 0.5)
        score_tensor = (seg_probs[i] * pred_mask_bool.float()).sum() / (pred_mask_bool.float().sum() + 1e-6)
        scores.append(float(score_tensor)); preds.append(pred_mask_bool[0].numpy())
    return np.array(scores, np.float32), np.stack(preds), (torch.from_numpy(masks_gt)[:, 0].int() > 0.5).numpy()


def coco_single(iou, area_d, area_g, thr, lo, hi):
    """COCOeval.evaluateImg for one detection and one ground truth of the same class."""
    g_ig = area_g < lo or area_g > hi
    matched = iou >= min(thr, 1 - 1e-10)
    d_ig = g_ig if matched else (area_d < lo or area_d > hi)
    return matched, d_ig, g_ig


def test_oracle_score_and_records_follow_the_reference():
    batch = helpers.make(batch=3, img_size=160, seed=41)
    ref = oracle.run_pipeline(batch, img_size=160, max_det=10)
    scores, preds, gts = reference_expressions(ref["seg_logits"], batch["masks_gt"])
    np.testing.assert_array_equal(preds, ref["seg_mask"].astype(bool))           # sigmoid > 0.5 == the oracle's threshold
    np.testing.assert_allclose(ref["seg_map_score"], scores, rtol=1e-5)
    thrs = oracle.iou_thresholds()
    rec = oracle.seg_map_records(ref["seg_img3"], ref["seg_map_score"], thrs)
    lo = [0.0, 0.0, 32.0 ** 2, 96.0 ** 2]; hi = [1e10, 32.0 ** 2, 96.0 ** 2, 1e10]
    for b in range(3):
        inter = int((preds[b] & gts[b]).sum()); P = int(preds[b].sum()); G = int(gts[b].sum())
        assert [inter, P, G] == ref["seg_img3"][b].tolist()
        iou = inter / (P + G - inter) if P + G - inter else 0.0
        for a in range(4):
            for t, thr in enumerate(thrs):
                m, dig, gig = coco_single(iou, P, G, thr, lo[a], hi[a])
                assert rec["dt_match"][b, a, t, 0] == int(m) and rec["dt_ignore"][b, a, t, 0] == int(dig)
                assert rec["gt_ignore"][b, a, 0] == int(gig)


def test_device_prep_equals_oracle_records_and_sweeps():
    batch = helpers.make(batch=6, img_size=160, seed=42)
    ref = oracle.run_pipeline(batch, img_size=160, max_det=10)
    thrs = oracle.iou_thresholds()
    rec = oracle.seg_map_records(ref["seg_img3"], ref["seg_map_score"], thrs)
    out = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in ref.items() if isinstance(v, np.ndarray)}
    got = seg_map_outputs(out, thrs)
    for k in ("dt_match", "dt_ignore", "gt_ignore", "det_count", "gt_count", "gt_labels"):
        np.testing.assert_array_equal(got[k].numpy(), rec[k], err_msg=k)
    np.testing.assert_allclose(got["dets"].numpy(), rec["dets"], rtol=1e-6)
    st = SweepState(1, len(thrs), thrs, (1, 10, 100))
    st.add(got, 0)
    res = st.compute()
    # the same through the oracle's numpy COCO accumulate
    recs = [dict(labels=np.zeros(1, np.int64), scores=rec["dets"][b, :1, 4], matched=rec["dt_match"][b][:, :, :1] > 0,
                 ignored=rec["dt_ignore"][b][:, :, :1] > 0) for b in range(6)]
    npig = np.array([[int((1 - rec["gt_ignore"][:, a, 0]).sum())] for a in range(4)], np.int64)
    want = oracle.accumulate_ap(recs, npig, thrs, (1, 10, 100), 1)
    for k in ("map", "map_50", "map_75", "mar_1", "mar_100"):
        assert float(res[k]) == pytest.approx(float(want[k]), rel=1e-12, abs=1e-12), k


@pytest.mark.gpu
def test_mask_kernel_score_numerator_matches_oracle():
    batch = helpers.make(batch=4, img_size=640, seed=43)
    ref = oracle.run_pipeline(batch, max_det=20)
    got, pp = helpers.run_cuda(batch, max_det=20)
    helpers.assert_same(got, ref, 4, 20)
    np.testing.assert_allclose(got["seg_prob_sum"], ref["seg_prob_sum"], rtol=1e-5)
    thrs = oracle.iou_thresholds()
    dev = seg_map_outputs(pp.out, thrs)
    rec = oracle.seg_map_records(ref["seg_img3"], ref["seg_map_score"], thrs)
    np.testing.assert_allclose(dev["seg_map_score"].cpu().numpy(), ref["seg_map_score"], rtol=1e-5)
    for k in ("dt_match", "dt_ignore", "gt_ignore"):
        np.testing.assert_array_equal(dev[k].cpu().numpy(), rec[k], err_msg=k)
