"""-m gpu: the CUDA library (through the C ABI) against the CPU oracle on identical seeded inputs.
Integer / index / byte outputs are compared bit-exactly; fp32 boxes, scores, coefficients and logits
are compared as raw bytes (the kernels reproduce the oracle's rounding sequence); Dice / IoU floats
within 1e-5 relative (north-star tolerance)."""
import numpy as np
import pytest

import helpers
from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kw", [
    dict(),                                         # reference defaults (CONF_TH .05, NMS_IOU .6, TOP_K 300)
    dict(max_det=100),                              # v3 TOP_K
    dict(class_mode=1), dict(class_mode=2),         # class-aware variants
    dict(gt_mode=1), dict(crop=0, max_det=20), dict(clamp=0),
    dict(conf_thres=0.999),                         # nothing passes the filter
    dict(conf_thres=0.5, iou_thres=0.3),
    dict(max_cand=500),
    dict(crop=0, max_det=100),                      # every detection on every strip: several rounds / scratch batches
    dict(max_det=400),                              # more detections than fit one round of the contract kernel's tables
    dict(nms_threads=1024), dict(nms_threads=512, class_mode=1),   # both NMS kernel variants whatever the heuristic picks
])
def test_pipeline_matches_oracle_640(kw):
    batch = helpers.make(batch=2, img_size=640)
    ref = oracle.run_pipeline(batch, **{k: v for k, v in kw.items() if k != "nms_threads"})   # a scheduling knob, not semantics
    got, _ = helpers.run_cuda(batch, **kw)
    helpers.assert_same(got, ref, 2, kw.get("max_det", 300))


@pytest.mark.parametrize("nms_threads", [0, 512, 1024])
def test_pipeline_dense_candidates(nms_threads):
    """conf 0.001: every anchor is a candidate (8400 per image) -> the 16-keys-per-thread register sort of the
    1024-thread NMS kernel (auto / 1024) and the global-memory sort of the 512-thread one."""
    batch = helpers.make(batch=2, img_size=640, seed=20264)
    kw = dict(conf_thres=0.001, max_det=300)
    ref = oracle.run_pipeline(batch, with_instances=False, **kw)
    got, _ = helpers.run_cuda(batch, nms_threads=nms_threads, **kw)
    assert int(ref["n_cand"].min()) > 8000
    helpers.assert_same(got, ref, 2, 300)


def test_pipeline_1024():
    batch = helpers.make(batch=1, img_size=1024, seed=20263)
    kw = dict(max_det=50)
    ref = oracle.run_pipeline(batch, img_size=1024, **kw)
    got, _ = helpers.run_cuda(batch, **kw)
    helpers.assert_same(got, ref, 1, 50)


@pytest.mark.parametrize("kw", [dict(max_det=300), dict(crop=0, max_det=40)])
def test_bf16_prototypes(kw):
    """Prototypes as bfloat16 (the reference validates under bf16-mixed and upcasts with .float()): the kernel widens
    them exactly, so every output equals the oracle's on the bf16-rounded prototypes, bit for bit (pool path and the
    on-the-fly corner contraction of detections without pool room)."""
    import torch
    batch = helpers.make(batch=2, img_size=640, seed=41)
    batch["protos"] = torch.from_numpy(batch["protos"]).bfloat16().float().numpy()
    ref = oracle.run_pipeline(batch, **kw)
    got, _ = helpers.run_cuda(batch, proto_bf16=True, **kw)
    helpers.assert_same(got, ref, 2, kw["max_det"])


def test_bf16_head_and_prototypes():
    """Both big inputs as bfloat16 (what the reference's bf16-mixed forward hands over): bit-exact against the oracle
    on the rounded tensors, detections and masks alike."""
    import torch
    batch = helpers.make(batch=2, img_size=640, seed=42)
    for k in ("head", "protos"):
        batch[k] = torch.from_numpy(batch[k]).bfloat16().float().numpy()
    ref = oracle.run_pipeline(batch)
    got, _ = helpers.run_cuda(batch, proto_bf16=True, head_bf16=True)
    helpers.assert_same(got, ref, 2, 300)


def test_gt_mask_f32():
    batch = helpers.make(batch=2, img_size=640, seed=7)
    kw = dict(max_det=30)
    ref = oracle.run_pipeline(batch, **kw)
    got, _ = helpers.run_cuda(batch, gt_f32=True, **kw)
    helpers.assert_same(got, ref, 2, 30)


def test_metrics_accumulate_and_reset():
    batch = helpers.make(batch=2, img_size=640, seed=11)
    kw = dict(max_det=20)
    ref = oracle.run_pipeline(batch, **kw)
    got, pp = helpers.run_cuda(batch, **kw)
    d = helpers.to_dev(batch, "cuda:0")
    out = pp.run(d["head"], d["protos"], d["det_boxes_gt"], d["masks_gt"], d["proj_weight"], d["proj_bias"])
    np.testing.assert_array_equal(out["cm"].cpu().numpy(), 2 * ref["cm"])
    np.testing.assert_array_equal(out["seg_cnt4"].cpu().numpy(), 2 * ref["seg_cnt4"])
    np.testing.assert_array_equal(out["seg_img3"].cpu().numpy(), ref["seg_img3"])       # per-batch, not accumulated
    np.testing.assert_array_equal(out["inst_area"].cpu().numpy(), ref["inst_area"])
    pp.reset_metrics()
    assert int(out["cm"].sum()) == 0


def test_large_batch_and_determinism():
    """16 images (persistent mask CTAs cross image boundaries, the NMS kernel runs on 16 SMs at once) against
    the oracle, then the same step 10 more times: every output must be bit-identical each time (a race between
    warps or CTAs would show up as run-to-run differences)."""
    from concurrent.futures import ThreadPoolExecutor
    import os
    import torch
    batch = helpers.make(batch=16, img_size=640, seed=20269)
    with ThreadPoolExecutor(os.cpu_count() or 4) as pool:
        ref = oracle.run_pipeline(batch, pool=pool)
    got, pp = helpers.run_cuda(batch)
    helpers.assert_same(got, ref, 16, 300)
    d = helpers.to_dev(batch, "cuda:0")
    keys = [k for k in got if k not in ("cm", "seg_cnt4", "uni_cnt4")]          # accumulated across calls
    for _ in range(10):
        out = pp.run(d["head"], d["protos"], d["det_boxes_gt"], d["masks_gt"], d["proj_weight"], d["proj_bias"])
        torch.cuda.synchronize()
        for k in keys:
            assert out[k].cpu().numpy().tobytes() == got[k].tobytes(), k


@pytest.mark.parametrize("mode", ["no_gt_rows", "gt_in_one_image_only", "empty_gt_mask"])
def test_ragged_and_empty_ground_truth(mode):
    """Edge cases of the GT side (the reference's collate_fn yields [0, 6] when a batch has no boxes)."""
    batch = helpers.make(batch=2, img_size=640, seed=33)
    if mode == "no_gt_rows":
        batch["det_boxes_gt"] = np.zeros((0, 6), np.float32)
    elif mode == "gt_in_one_image_only":
        batch["det_boxes_gt"] = batch["det_boxes_gt"][batch["det_boxes_gt"][:, 0] == 1]
    else:
        batch["masks_gt"] = np.zeros_like(batch["masks_gt"])
    kw = dict(max_det=40)
    ref = oracle.run_pipeline(batch, **kw)
    got, _ = helpers.run_cuda(batch, **kw)
    helpers.assert_same(got, ref, 2, 40)


def test_full_size_batch_is_consistent_with_small_batches():
    """BASELINE config size (batch 64 x 640^2; the oracle would need minutes): 8 distinct images checked against the
    oracle, then tiled 8x into a batch of 64 -- every per-image output must not depend on the batch size or on the
    image's position in it (at 64 images the persistent mask CTAs own ~12 strips each and cross image boundaries,
    the NMS kernel runs 64 CTAs at once), and the accumulated counters must be exactly 8x."""
    from concurrent.futures import ThreadPoolExecutor
    import os
    small = helpers.make(batch=8, img_size=640, seed=20270)
    with ThreadPoolExecutor(os.cpu_count() or 4) as pool:
        ref = oracle.run_pipeline(small, pool=pool)
    got8, _ = helpers.run_cuda(small)
    helpers.assert_same(got8, ref, 8, 300)
    big = dict(small)
    rep = lambda a: np.ascontiguousarray(np.concatenate([a] * 8, 0))
    for k in ("head", "protos", "masks_gt"):
        big[k] = rep(small[k])
    big["det_boxes_gt"] = np.concatenate([small["det_boxes_gt"] + np.array([8 * i, 0, 0, 0, 0, 0], np.float32) for i in range(8)], 0)
    from btpost import synth
    big["cfg"] = synth.SynthConfig(batch=64, img_size=640, seed=20270)
    got64, _ = helpers.run_cuda(big)
    for k, v in got8.items():
        if k in ("cm", "seg_cnt4", "uni_cnt4"):
            np.testing.assert_array_equal(got64[k], 8 * v, err_msg=k)
        else:
            assert got64[k].tobytes() == rep(v).tobytes(), k
