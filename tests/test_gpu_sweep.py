"""-m gpu: sweep aggregation on the device (the production placement) from the CUDA library's own outputs,
against the same aggregation on the CPU from the oracle's outputs."""
import numpy as np
import pytest
import torch

import helpers
from btpost.sweep import SweepState
from oracle import oracle

pytestmark = pytest.mark.gpu


def test_device_sweep_matches_cpu_sweep_of_oracle_outputs():
    kw = dict(max_det=100, gt_mode=1)
    st_dev = SweepState(3, 10, oracle.iou_thresholds(), (1, 10, 100), device="cuda:0")
    st_cpu = SweepState(3, 10, oracle.iou_thresholds(), (1, 10, 100))
    pp = None
    for i in range(2):                                   # two batches of 4 images, 160^2
        batch = helpers.make(batch=4, img_size=160, seed=31, image_offset=4 * i)
        ref = oracle.run_pipeline(batch, img_size=160, **kw)
        got, pp_i = helpers.run_cuda(batch, **kw)
        helpers.assert_same(got, ref, 4, 100)
        st_dev.add(pp_i.out, 4 * i, accumulate_counters=True)
        st_cpu.add({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in ref.items() if isinstance(v, np.ndarray)}, 4 * i,
                   accumulate_counters=True)
    a, b = st_dev.compute(), st_cpu.compute()
    assert a["n_images"] == b["n_images"] == 8
    for k in ("map", "map_50", "map_75", "mar_1", "mar_10", "mar_100", "seg_f1", "seg_dice", "seg_iou", "uni_dice", "uni_iou"):
        assert float(a[k]) == pytest.approx(float(b[k]), rel=1e-6, abs=1e-9), k
    np.testing.assert_array_equal(a["cm"].cpu().numpy(), b["cm"].numpy())
    np.testing.assert_allclose(a["precision"].cpu().numpy(), b["precision"].numpy(), rtol=1e-12)
