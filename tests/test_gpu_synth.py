"""-m gpu: the device-side generator of the synthetic workload (csrc/synth.cu, include/btpost_synth.h) is
bit-identical to the numpy generator the oracle-checked tests use, for any image offset (sharding)."""
import numpy as np
import pytest
import torch

from btpost import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("img_size,batch,offset", [(640, 3, 0), (160, 5, 1234), (1024, 1, 7)])
def test_device_generator_matches_numpy(img_size, batch, offset):
    cfg = synth.SynthConfig(batch=batch, img_size=img_size, seed=20271, image_offset=offset)
    ref = synth.make_batch(cfg)
    got = synth.make_batch_device(cfg, "cuda:0")
    torch.cuda.synchronize()
    for k in ("head", "protos", "masks_gt", "det_boxes_gt", "proj_weight"):
        assert got[k].cpu().numpy().tobytes() == np.ascontiguousarray(ref[k]).tobytes(), k
    assert got["proj_bias"] == float(ref["proj_bias"])
