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
    dict(nms_threads=1024), dict(nms_threads=512, class_mode=1), dict(nms_threads=256), dict(nms_threads=256, iou_thres=0.5),   # every NMS kernel variant whatever the heuristic picks
])
def test_pipeline_matches_oracle_640(kw):
    batch = helpers.make(batch=2, img_size=640)
    ref = oracle.run_pipeline(batch, **{k: v for k, v in kw.items() if k != "nms_threads"})   # a scheduling knob, not semantics
    got, _ = helpers.run_cuda(batch, **kw)
    helpers.assert_same(got, ref, 2, kw.get("max_det", 300))


@pytest.mark.parametrize("nms_threads", [0, 256, 512, 1024])
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


@pytest.mark.parametrize("kw", [dict(with_inst_masks="dense"), dict(with_inst_masks="bits", crop=0, max_det=24),
                                dict(with_inst_masks="dense", max_det=40, proto_bf16=True)])
def test_instance_masks_pixel_for_pixel(kw):
    """g1: every instance bitmap `[B,K,S,S]` (bit-packed, and expanded to bytes) against oracle.instance_mask
    (`/root/reference/src/test_model.py:80-85`: einsum -> bilinear -> sigmoid > 0.5, with / without the Ultralytics crop):
    a detection whose pixels are wrong inside another instance's footprint cannot hide behind equal popcounts."""
    import torch
    batch = helpers.make(batch=2, img_size=640, seed=20271)
    if kw.get("proto_bf16"):
        batch["protos"] = torch.from_numpy(batch["protos"]).bfloat16().float().numpy()
    okw = {k: v for k, v in kw.items() if k not in ("with_inst_masks", "proto_bf16")}
    ref = helpers.oracle_parallel(batch, **okw)
    got, _ = helpers.run_cuda(batch, **kw)
    assert "inst_bits" in got and sum(len(m) for m in ref["inst_masks"]) == int(ref["det_count"].sum()) > 0
    helpers.assert_same(got, ref, 2, kw.get("max_det", 300))


def test_pipeline_1024_batch8_full():
    """BASELINE config 3 shapes at batch 8: 1024^2 (21504 anchors, 256^2 prototypes), max_det 300, instance masks on and
    compared bitmap by bitmap."""
    batch = helpers.make(batch=8, img_size=1024, seed=20273)
    ref = helpers.oracle_parallel(batch, img_size=1024)
    got, _ = helpers.run_cuda(batch, with_inst_masks="bits")
    assert int(ref["det_count"].max()) > 200
    helpers.assert_same(got, ref, 8, 300)


def test_config4_dense_at_stated_size():
    """BASELINE config 4 as stated: batch 128 x 1024^2, conf 0.001 -> every one of the 21504 anchors is a candidate
    (above the 16384-key register sort: the global-memory sort network of nms_kernel), max_det 300.  Inputs come from the
    device generator (bit-identical to the numpy one, tests/test_gpu_synth.py); the oracle checks every detection-side
    and projector-side output of all 128 images (its instance masks at this size would take half an hour: those are
    compared at batch 8 above)."""
    import torch
    from btpost import PostConfig, PostProcessor, synth
    B, S = 128, 1024
    scfg = synth.SynthConfig(batch=B, img_size=S, seed=20274)
    d = synth.make_batch_device(scfg, "cuda:0")
    torch.cuda.synchronize()
    batch = {k: d[k].cpu().numpy() for k in ("head", "protos", "masks_gt", "det_boxes_gt", "proj_weight")}
    batch["proj_bias"] = d["proj_bias"]
    kw = dict(conf_thres=0.001, max_det=300)
    ref = helpers.oracle_parallel(batch, img_size=S, with_instances=False, with_masks_out=False, **kw)
    assert int(ref["n_cand"].min()) == 21504
    cfg = PostConfig(batch=B, img_size=S, conf_thres=0.001, iou_thres=0.6, max_det=300, with_coco=True, with_seg_map=True)
    pp = PostProcessor(cfg, "cuda:0")
    out = pp.run(d["head"], d["protos"], d["det_boxes_gt"], d["masks_gt"], d["proj_weight"], d["proj_bias"])
    torch.cuda.synchronize()
    got = {k: v.cpu().numpy() for k, v in out.items()}
    helpers.assert_same(got, ref, B, 300, check_masks=False)


def test_dense_30000_candidates_padded():
    """The `up to 30k boxes / image` end of config 4: the head is padded to N = 30000 anchors that all pass conf 0.001
    (32768-key global-memory sort)."""
    batch = helpers.make(batch=2, img_size=1024, seed=20275)
    rng = np.random.default_rng(5)
    B, C, N0 = batch["head"].shape
    extra = 30000 - N0
    pad = np.zeros((B, C, extra), np.float32)
    pad[:, 0:2] = rng.uniform(0, 1024, (B, 2, extra))
    pad[:, 2:4] = rng.uniform(4, 200, (B, 2, extra))
    pad[:, 4:7] = rng.uniform(0.002, 0.5, (B, 3, extra))
    pad[:, 7:] = rng.normal(0, 1, (B, C - 7, extra))
    batch["head"] = np.ascontiguousarray(np.concatenate([batch["head"], pad.astype(np.float32)], 2))
    kw = dict(conf_thres=0.001, max_det=300)
    ref = helpers.oracle_parallel(batch, img_size=1024, with_instances=False, **kw)
    assert int(ref["n_cand"].min()) == 30000
    got, _ = helpers.run_cuda(batch, num_anchors=30000, **kw)
    helpers.assert_same(got, ref, 2, 300)


def test_bf16_l1_maps():
    """f1: the reference's three raw maps as bfloat16 (bf16-mixed validation, `/root/reference/src/running_main_v2.py:1324`):
    widened exactly on load, so every output equals the oracle's on the rounded maps."""
    import torch
    batch = helpers.make(batch=2, img_size=640, seed=43, l1=True)
    rnd = lambda a: torch.from_numpy(a).bfloat16().float().numpy()
    batch["maps"] = [rnd(m) for m in batch["maps"]]
    batch["coeffs"] = rnd(batch["coeffs"])
    kw = dict(max_det=40)
    ref = oracle.run_pipeline(batch, layout="l1", **kw)
    got, _ = helpers.run_cuda(batch, l1=True, head_bf16=True, **kw)
    helpers.assert_same(got, ref, 2, 40)


def test_drop_gt_no_cand_and_gt_overflow():
    """(i) v2 drops the target of an image without any candidate above CONF_TH (`running_main_v2.py:797-814`),
    v3 keeps it (`running_main_v3.py:541-571`): an explicit flag, not tied to gt_mode; (ii) more GT rows than max_gt
    is reported, not silent."""
    import torch
    from btpost import PostConfig, PostProcessor
    batch = helpers.make(batch=2, img_size=640, seed=44)
    batch["head"][1, 4:7] = 0.0                                   # image 1: nothing passes the filter
    d = helpers.to_dev(batch, "cuda:0")
    args = (d["head"], d["protos"], d["det_boxes_gt"], d["masks_gt"], d["proj_weight"], d["proj_bias"])
    for drop in (False, True):
        pp = PostProcessor(PostConfig(batch=2, img_size=640, drop_gt_no_cand=drop, gt_mode=0), "cuda:0")
        pp.run(*args)
        preds, targets, _, _ = pp.to_reference_lists()
        assert len(preds[1]["boxes"]) == 0 and int(pp.out["gt_count"][1]) > 0
        assert len(targets[1]["boxes"]) == (0 if drop else int(pp.out["gt_count"][1]))
        assert len(targets[0]["boxes"]) == int(pp.out["gt_count"][0])
    rows = np.tile(batch["det_boxes_gt"][:1], (40, 1)).astype(np.float32)      # 40 rows for one image, max_gt = 32
    pp = PostProcessor(PostConfig(batch=2, img_size=640), "cuda:0")
    pp.run(d["head"], d["protos"], torch.from_numpy(rows).cuda(), d["masks_gt"], d["proj_weight"], d["proj_bias"])
    assert pp.out["gt_rows_total"].cpu().tolist()[int(rows[0, 0])] == 40
    with pytest.raises(ValueError, match="max_gt"):
        pp.to_reference_lists()
