"""a1: unpacking what the model hands over, with the reference's own checks.

CPU: `unpack_seg_outputs` reproduces `/root/reference/src/running_main_v2.py:286-316` (strict: the ValueErrors of
`_multitask_loss`) and `:672-688` / `evaluate_model.py:142-159` (lenient: the seg block is skipped); `unpack_infer_dict` takes
the dict of `ConvNeXtBiFPNYOLO.forward(x, mode="infer")` (`main_modelv2.py:362-378`).  GPU: `postprocess_infer` end to end."""
import numpy as np
import pytest
import torch

import btpost


def _parts(B=2, N=84, S=64):
    return [torch.zeros(B, 67, S // s, S // s) for s in (8, 16, 32)], torch.zeros(B, 32, N), torch.zeros(B, 32, S // 4, S // 4)


def test_unpack_seg_outputs_both_structures():
    maps, mc, protos = _parts()
    d, c, p = btpost.unpack_seg_outputs((maps, mc, protos))                       # train mode: (x, mc, p)
    assert d is maps and c is mc and p is protos
    d, c, p = btpost.unpack_seg_outputs((torch.zeros(2, 39, 84), (mc, protos)))   # eval mode: (cat, (mc, p))
    assert c is mc and p is protos


def test_unpack_seg_outputs_errors_match_the_reference():
    maps, mc, protos = _parts()
    with pytest.raises(ValueError, match="unhandled structure"):
        btpost.unpack_seg_outputs((maps, mc, protos, None))
    with pytest.raises(ValueError, match="unhandled structure"):
        btpost.unpack_seg_outputs((maps, (mc, protos, protos)))
    with pytest.raises(ValueError, match="must be 4D"):
        btpost.unpack_seg_outputs((maps, mc, protos[0]))
    with pytest.raises(ValueError, match="must be 4D"):
        btpost.unpack_seg_outputs((maps, mc, [protos]))
    with pytest.raises(ValueError, match="channel mismatch. Expected 32, got 16"):
        btpost.unpack_seg_outputs((maps, mc, protos[:, :16]))
    # the validation / eval variant never raises: the seg block is skipped (running_main_v2.py:680-688)
    assert btpost.unpack_seg_outputs((maps, mc, protos, None), strict=False) == (None, None, None)
    assert btpost.unpack_seg_outputs((maps, mc, protos[:, :16]), strict=False)[2] is None
    assert btpost.unpack_seg_outputs("nonsense", strict=False) == (None, None, None)


def test_unpack_infer_dict():
    _, mc, protos = _parts()
    head = torch.zeros(2, 39, 84)
    for sp in (protos, (mc, protos), (None, mc, protos)):                          # tensor, (mc, p), (feats, mc, p)
        h, p = btpost.unpack_infer_dict({"segment_preds_cat": head, "segment_protos": sp, "img_cls_logits": None})
        assert h is head and p is protos
    with pytest.raises(ValueError, match="segment_preds_cat"):
        btpost.unpack_infer_dict({"segment_protos": protos})
    with pytest.raises(ValueError, match="must be 4D"):
        btpost.unpack_infer_dict({"segment_preds_cat": head, "segment_protos": mc})
    with pytest.raises(ValueError, match=r"\[B, 4\+nc\+nm = 39, N\]"):
        btpost.unpack_infer_dict({"segment_preds_cat": head[:, :7], "segment_protos": protos})


def test_drop_in_argument_validation_cpu():
    with pytest.raises(ValueError, match="variant"):
        btpost.prepare_det_outputs_for_metrics_and_logging(torch.zeros(1, 39, 84), None, "cpu", 1, variant="v4")
    with pytest.raises(RuntimeError, match="CUDA devices only"):
        btpost.PostProcessor(btpost.PostConfig(batch=1), "cpu")


@pytest.mark.gpu
def test_postprocess_infer_end_to_end():
    import helpers
    from oracle import oracle
    batch = helpers.make(batch=2, img_size=640, seed=20276)
    d = helpers.to_dev(batch, "cuda:0")
    coeffs = d["head"][:, 7:].contiguous()
    out_dict = {"detect_preds_cat": d["head"][:, :7], "segment_protos": (coeffs, d["protos"]), "segment_preds_cat": d["head"],
                "img_cls_logits": torch.zeros(2, 3), "img_cls_probs": torch.zeros(2, 3)}
    out, pp = btpost.postprocess_infer(out_dict, d["det_boxes_gt"], d["masks_gt"], d["proj_weight"].view(1, 32, 1, 1), d["proj_bias"],
                                       max_det=40)
    torch.cuda.synchronize()
    ref = oracle.run_pipeline(batch, max_det=40, with_masks_out=False)
    got = {k: v.cpu().numpy() for k, v in out.items()}
    helpers.assert_same(got, ref, 2, 40, check_masks=False)
    preds, targets, log_preds, log_gts = pp.to_reference_lists()
    assert len(preds) == 2 and preds[0]["boxes"].shape[1] == 4 and preds[0]["labels"].dtype == torch.int64
    assert log_preds[0].shape[1] == 6 and log_gts[0].shape[1] == 5 and log_preds[0].is_cuda
    with pytest.raises(ValueError, match="with_inst_masks"):
        btpost.PostProcessor(btpost.PostConfig(batch=1, with_inst_masks="png"), "cuda:0")
    # float32 GT masks (the dataset's dtype, dataset_btxrdv2.py:164-166): another processor, same counters
    out2, _ = btpost.postprocess_infer(out_dict, d["det_boxes_gt"], d["masks_gt"].float(), d["proj_weight"], d["proj_bias"], max_det=40)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out2["seg_img3"].cpu().numpy(), ref["seg_img3"])
