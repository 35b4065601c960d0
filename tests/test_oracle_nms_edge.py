"""CPU: the NMS edge-case suite of tests/test_gpu_nms_edge.py (known answers k1-k6, quantised fuzz, class-aware modes,
Ultralytics max_nms) with the ORACLE in place of the CUDA kernels: pins the oracle's filter + NMS against
`torchvision.ops.nms` / torch `.max` (the reference's own calls, `/root/reference/src/running_main_v2.py:788-817`) on exactly
the inputs the GPU run is later held to."""
import numpy as np
import pytest

import test_gpu_nms_edge as T
from oracle import oracle
from test_gpu_nms_edge import (test_fuzz_against_torchvision, test_fuzz_class_aware_against_oracle_and_torchvision,  # noqa: F401
                               test_k1_ties_keep_the_lower_index, test_k2_iou_exactly_at_the_threshold_is_kept,
                               test_k3_identical_zero_area_boxes_are_both_kept, test_k4_nan_and_negative_zero_scores,
                               test_k5_k6_empty_input_and_output_format, test_max_cand_is_top_k_by_score_like_ultralytics)


def oracle_run_det(heads, conf=0.05, iou=0.6, max_det=300, S=64, class_mode=0, max_cand=0, nms_threads=0):
    heads = np.ascontiguousarray(heads, np.float32)
    B, _, N = heads.shape
    out = dict(det_count=np.zeros(B, np.int32), det_keep=np.full((B, max_det), -1, np.int64), dets=np.zeros((B, max_det, 6), np.float32),
               n_cand=np.zeros(B, np.int32), det_anchor=np.full((B, max_det), -1, np.int32))
    for b in range(B):
        boxes, score, label = oracle.decode_l2(heads[b], T.NC)
        cb = np.empty((N, 4), np.float32); cs = np.empty(N, np.float32); cl = np.empty(N, np.int32); ca = np.empty(N, np.int32)
        m = oracle.lib().bto_filter(boxes, score, label, N, np.float32(conf), 1, np.float32(S), np.float32(S), cb, cs, cl, ca)
        cb, cs, cl, ca = cb[:m].copy(), cs[:m].copy(), cl[:m].copy(), ca[:m].copy()
        keep = oracle.nms(cb, cs, iou, cl, class_mode, 7680.0, max_det, max_cand) if m else np.zeros(0, np.int64)
        k = len(keep)
        out["n_cand"][b], out["det_count"][b] = m, k
        out["det_keep"][b, :k] = keep
        out["det_anchor"][b, :k] = ca[keep]
        out["dets"][b, :k, :4] = cb[keep]; out["dets"][b, :k, 4] = cs[keep]; out["dets"][b, :k, 5] = cl[keep]
    return out


@pytest.fixture(autouse=True)
def _oracle_in_place_of_the_kernels(monkeypatch):
    monkeypatch.setattr(T, "run_det", oracle_run_det)
