"""-m gpu: the scheduling layers above the kernels must not change a single bit of the results:
`btpost_masks_parts` (the mask stage kernel group by kernel group), `btpost_run`'s forked helper stream, and
`Pipeline` (several batches in flight on several streams, each with its own workspace)."""
import numpy as np
import pytest
import torch

import helpers
from btpost import Pipeline, PostConfig, PostProcessor

pytestmark = pytest.mark.gpu

KEYS_ACC = ("cm", "seg_cnt4", "uni_cnt4")   # accumulated across calls


def _cfg(batch):
    return PostConfig(batch=batch, img_size=640, conf_thres=0.05, iou_thres=0.6, max_det=300, with_coco=True)


def _args(batch, dev):
    d = helpers.to_dev(batch, dev)
    return (d["head"], d["protos"], d["det_boxes_gt"], d["masks_gt"], d["proj_weight"], d["proj_bias"])


def test_stage_by_stage_equals_run():
    dev = torch.device("cuda:0")
    batch = helpers.make(batch=4, img_size=640, seed=31)
    args = _args(batch, dev)
    a, b = PostProcessor(_cfg(4), dev), PostProcessor(_cfg(4), dev)
    ra = {k: v.clone() for k, v in a.run(*args).items()}
    for stage in ("decode_filter", "nms_match", "masks_pack", "masks_contract", "masks_cells"):
        rb = b.run(*args, stage=stage)
    torch.cuda.synchronize()
    for k in ra:
        assert ra[k].cpu().numpy().tobytes() == rb[k].cpu().numpy().tobytes(), k


def test_pipeline_three_batches_in_flight_bit_identical():
    dev = torch.device("cuda:0")
    batch = helpers.make(batch=8, img_size=640, seed=32)
    args = _args(batch, dev)
    ref_pp = PostProcessor(_cfg(8), dev)
    ref = {k: v.clone() for k, v in ref_pp.run(*args).items()}
    pipe = Pipeline(_cfg(8), dev, depth=3).capture(*args)
    pipe.reset_metrics()
    n = 12
    pipe.fork()
    for _ in range(n):
        pipe.replay()
    pipe.join()
    torch.cuda.synchronize()
    for p in pipe.procs:
        for k, v in p.out.items():
            if k in KEYS_ACC:
                continue
            assert v.cpu().numpy().tobytes() == ref[k].cpu().numpy().tobytes(), k
    for k in KEYS_ACC:   # every step added its counters exactly once
        np.testing.assert_array_equal(pipe.counters(k).cpu().numpy(), n * ref[k].cpu().numpy())
