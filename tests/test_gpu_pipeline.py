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


def test_pipeline_six_different_batches_in_flight():
    """Six DIFFERENT batches in flight, each in its own slot (own static inputs, workspace, outputs, stream), two rounds:
    every slot's outputs must equal a serial PostProcessor's on the same batch bit for bit, and the shared counters must
    be the sum over all twelve steps.  The second round feeds the slots in a different order through `submit`."""
    dev = torch.device("cuda:0")
    B, depth = 4, 6
    batches = [helpers.make(batch=B, img_size=640, seed=600, image_offset=100 * i) for i in range(depth)]   # one seed: same projector weights
    serial = PostProcessor(_cfg(B), dev)
    refs = []
    for bt in batches:
        serial.reset_metrics()
        refs.append({k: v.clone() for k, v in serial.run(*_args(bt, dev)).items()})
    torch.cuda.synchronize()
    w, bias = _args(batches[0], dev)[4:6]
    pipe = Pipeline(_cfg(B), dev, depth=depth, proj_weight=w, proj_bias=bias)
    # round 1: zero-copy style -- inputs are written into the slots' buffers, then the steps are replayed round robin
    for i, bt in enumerate(batches):
        a = _args(bt, dev)
        pipe.load(i, a[0], a[1], a[2], a[3])
    pipe.fork()
    for i in range(depth):
        pipe.replay()
    pipe.join()
    torch.cuda.synchronize()
    for i in range(depth):
        for k, v in pipe.procs[i].out.items():
            if k in KEYS_ACC:
                continue
            assert v.cpu().numpy().tobytes() == refs[i][k].cpu().numpy().tobytes(), (i, k)
    # round 2: submit() from pinned host tensors, shifted by one slot; wait(slot) hands the outputs back
    order = [(i + 1) % depth for i in range(depth)]
    tickets = []
    for j in order:
        bt = batches[j]
        host = [torch.from_numpy(np.ascontiguousarray(bt[k])).pin_memory() for k in ("head", "protos", "det_boxes_gt", "masks_gt")]
        tickets.append((pipe.submit(*host), j))
    for slot, j in tickets:
        out = pipe.wait(slot)
        for k, v in out.items():
            if k in KEYS_ACC:
                continue
            assert v.cpu().numpy().tobytes() == refs[j][k].cpu().numpy().tobytes(), (slot, j, k)
    torch.cuda.synchronize()
    for k in KEYS_ACC:   # every step added its counters exactly once into the shared accumulators
        want = 2 * sum(r[k].cpu().numpy() for r in refs)
        np.testing.assert_array_equal(pipe.counters(k).cpu().numpy(), want)
