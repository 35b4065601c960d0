"""CPU: the oracle (oracle/) against fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden.py ran `MultiTaskLitModel.validation_step` of running_main_v2/v3 verbatim).
Inputs are regenerated from the seeded counter-based generator; nothing here reads /root/reference."""
from pathlib import Path

import numpy as np
import pytest

from btpost import synth
from oracle import oracle

GOLD = Path(__file__).resolve().parent / "golden"
CASES = ["ref_v3_s160", "ref_v2_s160", "ref_v3_s160_loose", "ref_v3_s640", "ref_v3_s160_peaked", "ref_v2_s160_peaked",
         "ref_v3_s640_peaked"]


def load(name):
    f = np.load(GOLD / f"{name}.npz")
    cfg = synth.SynthConfig(batch=int(f["batch"]), img_size=int(f["img_size"]), seed=int(f["seed"]))
    batch = synth.make_batch(cfg, l1=True, l1_peaked=bool(int(f["l1_peaked"])) if "l1_peaked" in f else False)
    kw = dict(layout="l1", img_size=cfg.img_size, conf_thres=float(f["conf_th"]), iou_thres=float(f["nms_iou"]),
              max_det=int(f["top_k"]), gt_mode=0, with_instances=False)
    return f, cfg, batch, kw


def check_against_reference(out, f, B, S):
    """`out` in the library/oracle layout vs what the reference handed to torchmetrics."""
    cm_ref = np.zeros((3, 3), np.int64)
    for pc, gc in f["cm_pairs"]:
        cm_ref[gc, pc] += 1
    np.testing.assert_array_equal(out["cm"], cm_ref)
    for i in range(B):
        k = int(out["det_count"][i])
        rb, rs, rl = f[f"pred_boxes_{i}"], f[f"pred_scores_{i}"], f[f"pred_labels_{i}"]
        assert k == len(rs), f"image {i}: {k} detections vs reference {len(rs)}"
        np.testing.assert_array_equal(out["dets"][i, :k, 5].astype(np.int64), rl)
        # the DFL decode goes through torch's softmax + einsum (summation order unspecified):
        # boxes / scores agree to a few ulp, not bit-for-bit (SURVEY.md §7 "Decode parity")
        np.testing.assert_allclose(out["dets"][i, :k, 4], rs, rtol=2e-6, atol=0)
        np.testing.assert_allclose(out["dets"][i, :k, :4], rb, rtol=0, atol=2e-4 * S / 160)
        g = len(f[f"gt_labels_{i}"])
        if f["version"] == "v2" and int(out["n_cand"][i]) == 0:
            g = 0        # v2 drops the target of an image without candidates (running_main_v2.py:797-814)
        assert int(out["gt_count"][i]) >= g
        np.testing.assert_array_equal(out["gt_boxes"][i, :g].tobytes(), f[f"gt_boxes_{i}"].tobytes())
        np.testing.assert_array_equal(out["gt_labels"][i, :g], f[f"gt_labels_{i}"])
    bits = np.packbits(out["seg_mask"].reshape(B, -1), axis=1)
    np.testing.assert_array_equal(bits, f["seg_pred_bits"])
    if "seg_logits" in f:
        assert out["seg_logits"].astype(np.float32).tobytes() == f["seg_logits"].tobytes()
    else:
        np.testing.assert_array_equal(out["seg_logits"].reshape(B, -1)[:, ::7], f["seg_logits_sub"])
        np.testing.assert_array_equal(np.bitwise_xor.reduce(out["seg_logits"].view(np.uint32).reshape(B, -1), axis=1),
                                      f["seg_logits_xor"])


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_validation_step(name):
    f, cfg, batch, kw = load(name)
    out = oracle.run_pipeline(batch, **kw)
    check_against_reference(out, f, cfg.batch, cfg.img_size)
