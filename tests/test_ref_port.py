"""CPU: the torch port of the reference's ops (oracle/ref_torch.py -- what `bench.py --impl reference` and the cpu_baseline
leg time on the GPU box, where /root/reference does not exist) agrees with the oracle: identical detections, confusion
matrix and projector counters; instance-mask Dice within the summation-order noise of einsum / conv."""
import numpy as np

from btpost import synth
from oracle import oracle, ref_torch


def test_torch_port_agrees_with_oracle():
    cfg = synth.SynthConfig(batch=2, img_size=640, seed=20262)
    batch = synth.make_batch(cfg)
    kw = dict(max_det=40)
    got = ref_torch.run_batch(batch, **kw)
    ref = oracle.run_pipeline(batch, gt_mode=1, with_masks_out=False, **kw)
    assert got["det_count"] == ref["det_count"].tolist()
    for b in range(2):
        k = got["det_count"][b]
        assert got["preds"][b]["boxes"].numpy().tobytes() == ref["dets"][b, :k, :4].tobytes()
        assert got["preds"][b]["scores"].numpy().tobytes() == ref["dets"][b, :k, 4].tobytes()
        np.testing.assert_array_equal(got["preds"][b]["labels"].numpy(), ref["dets"][b, :k, 5].astype(np.int64))
    np.testing.assert_array_equal(got["cm"].numpy(), ref["cm"])
    np.testing.assert_array_equal(np.asarray(got["seg_cnt"]), ref["seg_cnt4"])
    np.testing.assert_allclose(got["seg_dice"].numpy(), ref["seg_dice"], rtol=1e-5)
    np.testing.assert_allclose(got["uni_dice"], ref["uni_dice"], rtol=2e-3)     # a few pixels at |logit| ~ 1e-7 may flip
    np.testing.assert_allclose(got["uni_iou"], ref["uni_iou"], rtol=2e-3)
    assert got["matched"] == int((ref["dt_match"] > 0).sum())
