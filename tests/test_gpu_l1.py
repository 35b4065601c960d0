"""-m gpu: the reference's own input layout (three raw maps, DFL decode in the kernel) through the
C ABI, against the oracle (bit-exact) and against the fixtures produced by the unmodified reference."""
import numpy as np
import pytest
import torch

import btpost
import helpers
from oracle import oracle
from test_oracle_golden import CASES, check_against_reference, load

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", CASES)
def test_l1_matches_oracle_and_reference_fixture(name):
    f, cfg, batch, kw = load(name)
    batch["cfg"] = cfg
    kw2 = dict(conf_thres=kw["conf_thres"], iou_thres=kw["iou_thres"], max_det=kw["max_det"])
    got, _ = helpers.run_cuda(batch, l1=True, **kw2)
    ref = oracle.run_pipeline(batch, layout="l1", img_size=cfg.img_size, **kw2)
    helpers.assert_same(got, ref, cfg.batch, kw["max_det"])          # CUDA == oracle, bit for bit
    check_against_reference(got, f, cfg.batch, cfg.img_size)         # CUDA vs the reference's own outputs


def test_drop_in_signature_returns_reference_formats():
    """`_prepare_det_outputs_for_metrics_and_logging(det_outputs, det_boxes_gt, device, batch_size)`
    (evaluate_model.py:174-178) with the reference's list-of-maps input."""
    f, cfg, batch, kw = load("ref_v3_s160")
    dev = torch.device("cuda:0")
    btpost.api.TOP_K = int(f["top_k"])
    try:
        maps = [torch.from_numpy(m).to(dev) for m in batch["maps"]]
        preds, targets, log_preds, log_gts = btpost.prepare_det_outputs_for_metrics_and_logging(
            maps, torch.from_numpy(batch["det_boxes_gt"]).to(dev), dev, cfg.batch, img_size=cfg.img_size, nc=cfg.nc)
    finally:
        btpost.api.TOP_K = 300
    assert len(preds) == len(targets) == len(log_preds) == len(log_gts) == cfg.batch
    for i in range(cfg.batch):
        assert preds[i]["boxes"].device.type == "cpu" and preds[i]["labels"].dtype == torch.int64
        assert log_preds[i].device.type == "cuda" and log_preds[i].shape[1] == 6 and log_gts[i].shape[1] == 5
        np.testing.assert_array_equal(preds[i]["labels"].numpy(), f[f"pred_labels_{i}"])
        np.testing.assert_allclose(preds[i]["scores"].numpy(), f[f"pred_scores_{i}"], rtol=2e-6)
        np.testing.assert_allclose(preds[i]["boxes"].numpy(), f[f"pred_boxes_{i}"], atol=2e-4)
        assert targets[i]["boxes"].numpy().tobytes() == f[f"gt_boxes_{i}"].tobytes()
        np.testing.assert_array_equal(targets[i]["labels"].numpy(), f[f"gt_labels_{i}"])
        np.testing.assert_array_equal(log_preds[i][:, :4].cpu().numpy(), preds[i]["boxes"].numpy())
