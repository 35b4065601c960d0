#!/usr/bin/env python
"""Sharded evaluation sweep (BASELINE config 5, scaled by --images): every rank post-processes its own
images (synthetic head outputs generated per batch from the counter-based generator, so any sharding sees
the same images), the only collectives are the counter all-reduce and the AP record gather at the end.

    python scripts/run_sweep.py --images 256                                   # 1 GPU
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/run_sweep.py --images 512
Prints one JSON line with mAP / mAP50 / Dice / F1 and the device time of the sweep."""
import argparse, json, os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "multitask-bonetumor-yolo_b200")]
import numpy as np, torch, torch.distributed as dist
from btpost import PostConfig, PostProcessor, synth
from btpost.api import map_iou_thresholds
from btpost.sweep import SweepState
from btpost.segmap import seg_map_outputs

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=256)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--img", type=int, default=640)
args = ap.parse_args()
rank, world, lrank = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(lrank)
dev = torch.device("cuda", lrank)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    dist.all_reduce(torch.zeros(1, device=dev))   # connect the peers now: communicator set-up is not part of the sweep
    dist.all_gather([torch.zeros(1, device=dev) for _ in range(world)], torch.zeros(1, device=dev))
B = args.batch
nbatches = args.images // B
mine = [i for i in range(nbatches) if i % world == rank]          # whole batches b = r (mod G)
pp = PostProcessor(PostConfig(batch=B, img_size=args.img, gt_mode=1, max_det=100, with_seg_map=True), dev)
st = SweepState(3, 10, map_iou_thresholds(), (1, 10, 100), device=dev)
st_seg = SweepState(1, 10, map_iou_thresholds(), (1, 10, 100), device=dev)   # v3 segmentation mAP: one mask pair per image
pp.reset_metrics()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
gpu_ms, t_host = 0.0, time.perf_counter()
for i in mine:
    # this rank's images, generated on the device (bit-identical to the numpy generator: tests/test_gpu_synth.py)
    d = synth.make_batch_device(synth.SynthConfig(batch=B, img_size=args.img, seed=20265, image_offset=i * B), dev)
    ev0.record()
    out = pp.run(d["head"], d["protos"], d["det_boxes_gt"], d["masks_gt"], d["proj_weight"], d["proj_bias"])
    st.add(out, i * B)
    st_seg.add(seg_map_outputs(out, map_iou_thresholds()), i * B)
    ev1.record()
    torch.cuda.synchronize()
    gpu_ms += ev0.elapsed_time(ev1)
st.take_counters(pp.out)
ev0.record()
if world > 1:
    st.all_reduce(); st.gather()
    st_seg.all_reduce(); st_seg.gather()
res = st.compute()
res_seg = st_seg.compute()
ev1.record()
torch.cuda.synchronize()
if rank == 0:
    keys = ("n_images", "map", "map_50", "map_75", "mar_100", "seg_f1", "seg_dice", "seg_iou", "uni_dice", "uni_iou")
    line = {k: (float(res[k]) if not isinstance(res[k], int) else res[k]) for k in keys if k in res}
    line.update(seg_map=float(res_seg["map"]), seg_map_50=float(res_seg["map_50"]))
    line.update(world=world, images=nbatches * B, device_ms_per_rank_batches=gpu_ms, reduce_gather_compute_ms=ev0.elapsed_time(ev1),
                host_s=time.perf_counter() - t_host)
    print(json.dumps(line))
if world > 1:
    dist.destroy_process_group()
