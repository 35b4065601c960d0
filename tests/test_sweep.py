"""CPU: sweep-level aggregation (counter all-reduce + AP record gather + COCO accumulate) -- against the
oracle's numpy restatement of COCOeval.accumulate, and across a 2-rank gloo group (images sharded)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from btpost import synth
from btpost import _lib
from btpost.sweep import HDR, REC_DTYPE, SweepState, decode_records, merge_shards
from oracle import oracle

KW = dict(conf_thres=0.05, iou_thres=0.6, max_det=100, gt_mode=1, with_instances=False, with_masks_out=False)
MAX_DETS = (1, 10, 100)


def oracle_outputs(n_images, image_offset=0, seed=77):
    cfg = synth.SynthConfig(batch=n_images, img_size=160, seed=seed, image_offset=image_offset)
    batch = synth.make_batch(cfg)
    out = oracle.run_pipeline(batch, img_size=160, **KW)
    keys = ["dets", "det_count", "dt_match", "dt_ignore", "gt_ignore", "gt_labels", "gt_count", "seg_dice", "seg_iou",
            "uni_dice", "uni_iou", "cm", "seg_cnt4", "uni_cnt4"]
    return out, {k: torch.from_numpy(np.ascontiguousarray(out[k])) for k in keys}


def oracle_ap(out, n_images):
    thrs = oracle.iou_thresholds()
    recs, npig = [], np.zeros((4, 3), np.int64)
    for b in range(n_images):
        k, g = int(out["det_count"][b]), int(out["gt_count"][b])
        recs.append(dict(labels=out["dets"][b, :k, 5].astype(np.int64), scores=out["dets"][b, :k, 4],
                         matched=out["dt_match"][b][:, :, :k] > 0, ignored=out["dt_ignore"][b][:, :, :k] > 0))
        for a in range(4):
            for gi in range(g):
                if not out["gt_ignore"][b, a, gi]:
                    npig[a, out["gt_labels"][b, gi]] += 1
    return oracle.accumulate_ap(recs, npig, thrs, MAX_DETS, 3)


def assert_ap_equal(res, ref):
    for k in ("map", "map_50", "map_75", "map_small", "map_medium", "map_large", "mar_1", "mar_10", "mar_100",
              "mar_small", "mar_medium", "mar_large"):
        assert res[k] == pytest.approx(ref[k], rel=1e-12, abs=1e-12), k
    np.testing.assert_allclose(res["precision"].numpy(), ref["precision"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(res["recall"].numpy(), ref["recall"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(res["map_per_class"], ref["map_per_class"], rtol=1e-12)


def test_compute_matches_oracle_accumulate():
    n = 12
    out, tens = oracle_outputs(n)
    st = SweepState(3, 10, oracle.iou_thresholds(), MAX_DETS)
    st.add(tens, 0, accumulate_counters=True)
    res = st.compute()
    ref = oracle_ap(out, n)
    assert ref["map_50"] > 0.05           # the synthetic detections do hit the GT boxes
    assert_ap_equal(res, ref)
    assert res["n_images"] == n
    np.testing.assert_array_equal(res["cm"].numpy(), out["cm"])
    tp, fp, fn, tn = out["seg_cnt4"]
    assert res["seg_f1"] == pytest.approx(2 * tp / (2 * tp + fp + fn))
    assert res["seg_dice"] == pytest.approx(float(out["seg_dice"].astype(np.float64).mean()))


def test_batches_and_order_do_not_matter():
    out, tens = oracle_outputs(8)
    ref = oracle_ap(out, 8)
    st = SweepState(3, 10, oracle.iou_thresholds(), MAX_DETS)
    for lo, hi in ((4, 8), (0, 4)):      # added out of order, in two batches
        part = {k: (v[lo:hi] if v.dim() and v.shape[0] == 8 and k not in ("cm", "seg_cnt4", "uni_cnt4") else v) for k, v in tens.items()}
        st.add(part, lo)
    assert_ap_equal(st.compute(), ref)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_per_rank, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _, tens = oracle_outputs(n_per_rank, image_offset=rank * n_per_rank)
        st = SweepState(3, 10, oracle.iou_thresholds(), MAX_DETS)
        st.add(tens, rank * n_per_rank, accumulate_counters=True)
        st.all_reduce()
        st.gather()
        res = st.compute()
        if rank == 0:
            q.put({k: (v.numpy() if torch.is_tensor(v) else v) for k, v in res.items()})
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharded_sweep_equals_single_process():
    n_per_rank, world = 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, PORT, n_per_rank, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single process over the same 10 images (global indices 0..9)
    outs = [oracle_outputs(n_per_rank, image_offset=r * n_per_rank) for r in range(world)]
    st = SweepState(3, 10, oracle.iou_thresholds(), MAX_DETS)
    for r, (_, tens) in enumerate(outs):
        st.add(tens, r * n_per_rank, accumulate_counters=True)
    ref = st.compute()
    assert res["n_images"] == ref["n_images"] == 10
    np.testing.assert_array_equal(res["cm"], ref["cm"].numpy())
    np.testing.assert_array_equal(res["precision"], ref["precision"].numpy())
    np.testing.assert_array_equal(res["recall"], ref["recall"].numpy())
    for k in ("map", "map_50", "mar_100", "seg_f1", "seg_dice", "uni_iou"):
        assert res[k] == pytest.approx(ref[k], rel=1e-12)


def encode_records(out, image_offset, T=10):
    """What match_kernel appends to the device sweep (include/btpost.h BtSweepRecord), built on the host from oracle
    outputs: the CPU stand-in for the shards `merge_shards` moves."""
    rows = []
    bits = (np.uint64(1) << np.arange(4 * T, dtype=np.uint64))
    for b in range(len(out["det_count"])):
        k = int(out["det_count"][b])
        lab = out["dets"][b, :k, 5].astype(np.int64)
        r = np.zeros(k, REC_DTYPE)
        mt = np.moveaxis(out["dt_match"][b][:, :, :k] > 0, 2, 0).reshape(k, 4 * T)
        ig = np.moveaxis(out["dt_ignore"][b][:, :, :k] > 0, 2, 0).reshape(k, 4 * T)
        r["matched"] = (mt * bits).sum(1).astype(np.uint64)
        r["ignored"] = (ig * bits).sum(1).astype(np.uint64)
        u = out["dets"][b, :k, 4].astype(np.float32).view(np.uint32)
        asc = np.where(u & np.uint32(0x80000000), ~u, u | np.uint32(0x80000000)).astype(np.uint32)
        r["score_key"] = ~asc
        r["image"], r["rank"], r["label"] = image_offset + b, np.arange(k), lab
        r["class_rank"] = [int((lab[:j] == lab[j]).sum()) for j in range(k)]
        rows.append(r)
    return np.concatenate(rows) if rows else np.zeros(0, REC_DTYPE)


def _shard(rank, n_per_rank):
    out, _ = oracle_outputs(n_per_rank, image_offset=rank * n_per_rank)
    rec = encode_records(out, rank * n_per_rank)
    hdr = np.zeros(HDR, np.int64)
    hdr[_lib.SWEEP_N_RECORDS], hdr[_lib.SWEEP_N_IMAGES] = len(rec), n_per_rank
    hdr[_lib.SWEEP_USER: _lib.SWEEP_USER + 9] = out["cm"].ravel()
    ring = np.zeros((len(rec) + 7 * rank, 32), np.uint8)            # rings of different capacity per rank
    ring[:len(rec)] = rec.view(np.uint8).reshape(-1, 32)
    return out, hdr, ring


def _merge_worker(rank, world, port, n_per_rank, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _, hdr, ring = _shard(rank, n_per_rank)
        h = torch.from_numpy(hdr)
        ns, merged = merge_shards(h, torch.from_numpy(ring))
        if rank == 1:                                               # every rank holds the merged state; check a non-zero one
            q.put((ns, merged.numpy().copy(), h.numpy().copy()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_merge_of_device_sweep_shards():
    """The N > 1 path of DeviceSweep.finish on CPU tensors over gloo: ONE all-reduce of the header, one
    all_gather_into_tensor of the (ragged, padded) record rings; the merged records decode to exactly the two shards in
    rank order and feed the oracle's accumulate to the single-process result."""
    n_per_rank, world = 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_merge_worker, args=(r, world, port, n_per_rank, q)) for r in range(world)]
    for p in procs:
        p.start()
    ns, merged, hdr = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shards = [_shard(r, n_per_rank) for r in range(world)]
    assert ns == [int(s[1][0]) for s in shards]
    want = np.concatenate([s[2][:n] for s, n in zip(shards, ns)])
    np.testing.assert_array_equal(merged, want)
    assert hdr[_lib.SWEEP_N_RECORDS] == sum(ns) and hdr[_lib.SWEEP_N_IMAGES] == world * n_per_rank
    np.testing.assert_array_equal(hdr[_lib.SWEEP_USER: _lib.SWEEP_USER + 9], sum(s[0]["cm"] for s in shards).ravel())
    rec = decode_records(torch.from_numpy(merged), 10)
    np.testing.assert_array_equal(rec["image"], np.concatenate([np.repeat(np.arange(r * n_per_rank, (r + 1) * n_per_rank), s[0]["det_count"])
                                                                for r, s in enumerate(shards)]))
    assert rec["score"].tobytes() == np.concatenate([s[0]["dets"][b, :int(s[0]["det_count"][b]), 4] for s in shards
                                                     for b in range(n_per_rank)]).tobytes()


PORT = _free_port()
