#!/usr/bin/env python
"""bench.py -- images/sec of the post-processing + evaluation hot path (BASELINE.json metric).

A "step" is one pass of the whole hot path (decode+filter -> NMS+COCO matching -> mask assembly +
Dice/IoU counters) over one batch of synthetic head outputs.  At N GPUs every rank owns its own
batch of `--batch` images per step (images are independent units: weak scaling, no data-path
collective); the only NCCL traffic is one all-reduce of the metric counters at the end of the
timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference ...                           # the reference's CPU path (oracle port)

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definition of every key.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for _p in (ROOT, ROOT / "multitask-bonetumor-yolo_b200"):
    if str(_p) not in sys.path:
        sys.path.insert(0, str(_p))

import numpy as np  # noqa: E402

METRIC = "images/sec post-proc+eval"
UNIT = "images/s"


def algorithmic_bytes_per_image(S, nc=3, nm=32):
    """head + protos + u8 GT mask (full eval, masks fused away)."""
    N = sum((S // s) ** 2 for s in (8, 16, 32))
    head = 4 * (4 + nc + nm) * N
    protos = 4 * nm * (S // 4) ** 2
    gt_mask = S * S
    return dict(head=head, protos=protos, gt_mask=gt_mask, total=head + protos + gt_mask)


def workload_name(args):
    return (f"batch {args.batch} x {args.img}^2 synthetic head outputs (L2 [B,39,N] + 32ch protos), "
            f"decode+NMS+mask assembly+Dice/IoU+COCO matching, conf {args.conf}, iou {args.iou}, max_det {args.max_det}")


# --------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock + throttle reasons during the timed region with the recipe's own
    `nvidia-smi --query-gpu=... -lms` line, run as a SEPARATE process (NVML calls made from a thread
    of this process serialise against CUDA launches and were measured to double the step time)."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=100):
        import shutil
        import subprocess
        import tempfile
        self.proc, self.path = None, None
        exe = shutil.which("nvidia-smi")
        if exe is None or period_ms <= 0:
            return
        f = tempfile.NamedTemporaryFile(prefix="btpost_clocks_", suffix=".csv", delete=False)
        self.path = f.name
        self.proc = subprocess.Popen([exe, "-i", str(index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                      "-lms", str(period_ms)], stdout=f, stderr=subprocess.DEVNULL)

    def start(self):
        return self

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "note": "nvidia-smi not found"}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        mhz, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path).read().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 6 or not parts[0].isdigit():
                continue
            mhz.append(int(parts[0]))
            mx = int(parts[1]) if parts[1].isdigit() else mx
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        mhz.sort()
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(mhz)}


# --------------------------------------------------------------------------------------------
def make_inputs(args, rank):
    from btpost import synth
    cfg = synth.SynthConfig(batch=args.batch, img_size=args.img, seed=20262, image_offset=rank * args.batch)
    b = synth.make_batch(cfg)
    b["cfg"] = cfg
    return b


def cpu_oracle_rate(args, n_images, threads):
    """Times the CPU oracle (port of the reference path) on `n_images` images of the same workload."""
    from concurrent.futures import ThreadPoolExecutor

    from btpost import synth
    from oracle import oracle
    cfg = synth.SynthConfig(batch=n_images, img_size=args.img, seed=20262)
    batch = synth.make_batch(cfg)
    kw = dict(conf_thres=args.conf, iou_thres=args.iou, max_det=args.max_det, img_size=args.img, with_masks_out=False)
    pool = ThreadPoolExecutor(threads) if threads > 1 else None
    t0 = time.perf_counter()
    oracle.run_pipeline(batch, pool=pool, **kw)
    dt = time.perf_counter() - t0
    if pool:
        pool.shutdown()
    return n_images / dt, dt


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path.  The reference is pure
    Python/PyTorch and is not present on the GPU box, so this is the oracle port (oracle/), run on
    all host threads, each step a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ThreadPoolExecutor

    from btpost import synth
    from oracle import oracle
    threads = os.cpu_count() or 1
    pool = ThreadPoolExecutor(threads)
    kw = dict(conf_thres=args.conf, iou_thres=args.iou, max_det=args.max_det, img_size=args.img, with_masks_out=False)
    sample = 1  # images per step
    cfg = synth.SynthConfig(batch=sample, img_size=args.img, seed=20262)
    batch = synth.make_batch(cfg)
    for _ in range(max(args.warmup, 1)):
        oracle.run_pipeline(batch, pool=pool, **kw)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.run_pipeline(batch, pool=pool, **kw)
    dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "sample_images_per_step": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} image(s) of the workload per step, {args.steps} steps, instance masks spread over {threads} host threads"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--img", type=int, default=640)
    ap.add_argument("--conf", type=float, default=0.05)
    ap.add_argument("--iou", type=float, default=0.6)
    ap.add_argument("--max-det", dest="max_det", type=int, default=300)
    ap.add_argument("--cpu-sample", type=int, default=8, help="images the cpu_baseline leg times (0 = skip)")
    ap.add_argument("--pipeline", type=int, default=6,
                    help="batches in flight: steps are replayed round-robin on this many streams (1 = strictly serial steps)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--clock-period-ms", type=int, default=100, help="nvidia-smi sampling period during the timed region (0 = off)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    from btpost import Pipeline, PostConfig, PostProcessor, _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    _lib.load()  # fail loudly if libbtpost.so is missing
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B, S = args.batch, args.img
    batch = make_inputs(args, rank)
    host = {k: torch.from_numpy(np.ascontiguousarray(batch[k])).pin_memory()
            for k in ("head", "protos", "det_boxes_gt", "masks_gt", "proj_weight")}
    d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    bias = float(batch["proj_bias"])
    cfg = PostConfig(batch=B, img_size=S, conf_thres=args.conf, iou_thres=args.iou, max_det=args.max_det, with_coco=True)
    pipe = Pipeline(cfg, dev, depth=max(1, args.pipeline))
    pp = PostProcessor(cfg, dev)   # one batch in flight: strictly serial steps, per-stage timings, e2e
    counters = torch.zeros(3 * 3 + 4 + 4 + 4, dtype=torch.float64, device=dev)

    def step(inp=d, stage="run"):
        return pp.run(inp["head"], inp["protos"], inp["det_boxes_gt"], inp["masks_gt"], inp["proj_weight"], bias, stage=stage)

    def pack_counters(out):
        # cm (nc*nc), seg tp/fp/fn/tn, uni tp/fp/fn/tn, [sum seg dice, sum seg iou, sum uni dice, sum uni iou]
        torch.cat([pipe.counters("cm").flatten().double(), pipe.counters("seg_cnt4").double(), pipe.counters("uni_cnt4").double(),
                   torch.stack([out["seg_dice"].sum(), out["seg_iou"].sum(), out["uni_dice"].sum(),
                                out["uni_iou"].sum()]).double()], out=counters)
        return counters

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- warm-up; the step is captured once into a CUDA graph (9 kernels per replay)
    pipe.capture(d["head"], d["protos"], d["det_boxes_gt"], d["masks_gt"], d["proj_weight"], bias)
    graph = pp.capture(d["head"], d["protos"], d["det_boxes_gt"], d["masks_gt"], d["proj_weight"], bias)
    out = pipe.procs[0].out
    pipe.fork()
    for _ in range(args.warmup):
        pipe.replay()
    pipe.join()
    # first use of the counter-packing ops / the NCCL communicator loads modules and connects peers:
    # done once here so that the timed region only holds the steps and the one counter all-reduce
    c = pack_counters(out)
    if world > 1:
        dist.all_reduce(c)
    barrier()
    sampler = ClockSampler(local_rank, args.clock_period_ms)
    sampler.start()

    # ---------------- timed region: K steps, inputs resident in HBM (320 MB/step at 640^2 > L2)
    pipe.reset_metrics()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    pipe.fork()
    for _ in range(args.steps):
        pipe.replay()      # step i on stream i % depth: consecutive batches overlap, every step does all of its work
    pipe.join()
    c = pack_counters(out)
    if world > 1:
        dist.all_reduce(c)  # the only collective: metric counters (NCCL over NVLink)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = B * world * args.steps / (ms / 1e3)

    # ---------------- strictly serial steps (one stream, one batch in flight): the latency of a step
    torch.cuda.synchronize()
    es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    es0.record()
    for _ in range(args.steps):
        graph.replay()
    es1.record()
    torch.cuda.synchronize()
    serial_ms = es0.elapsed_time(es1) / args.steps

    # ---------------- per-stage device times (CUDA events on the launching stream)
    stage_ms = {}
    ORDER = ("decode_filter", "nms_match", "masks_pack", "masks_contract", "masks_cells")
    for stage in ORDER:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tot = 0.0
        for _ in range(args.steps):
            for s2 in ORDER:   # keep the real order so caches look like a step
                if s2 == stage:
                    e0.record()
                step(stage=s2)
                if s2 == stage:
                    e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        stage_ms[stage] = tot / args.steps

    # ---------------- end to end through the public API with HOST buffers
    e2e = None
    if not args.no_e2e:
        res_host = {k: torch.empty_like(out[k], device="cpu").pin_memory()
                    for k in ("det_count", "dets", "seg_dice", "seg_iou", "uni_dice", "uni_iou", "cm", "seg_cnt4")}
        dd = {k: torch.empty_like(v) for k, v in d.items()}
        graph2 = pp.capture(dd["head"], dd["protos"], dd["det_boxes_gt"], dd["masks_gt"], dd["proj_weight"], bias)

        def e2e_step():
            for k in dd:
                dd[k].copy_(host[k], non_blocking=True)
            graph2.replay()
            o = pp.out
            for k, hbuf in res_host.items():
                hbuf.copy_(o[k], non_blocking=True)
            torch.cuda.current_stream().synchronize()   # the caller reads the step's result

        for _ in range(3):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": B * world * args.steps / float(dt.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(sum(v.numel() * v.element_size() for v in host.values())),
               "d2h_bytes_per_step": int(sum(v.numel() * v.element_size() for v in res_host.values()))}
    clocks = sampler.stop()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    peak_gbs, peak_src = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
    ab = algorithmic_bytes_per_image(S)
    mask_bytes = ab["protos"] * B       # algorithmic bytes of one contract_kernel launch: the prototypes, once
    achieved = mask_bytes / (stage_ms["masks_contract"] / 1e3) / 1e9
    traffic = None
    tf = ROOT / "profiles" / "roofline_traffic.json"
    if tf.exists():
        try:
            traffic = json.loads(tf.read_text()).get(f"contract_kernel_B{B}_S{S}")
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "global_batch": B * world, "sharding": f"images sharded, {B}/GPU",
                   "batches_in_flight": max(1, args.pipeline),
                   "l2": "inputs (%.0f MB/step/GPU) larger than the 126 MB L2" % (ab["total"] * B / 1e6)},
        "roofline": {"bound": "hbm", "kernel": "contract_kernel", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                     "frac": achieved / peak_gbs, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": mask_bytes},
        "pipeline": {"algorithmic_bytes_per_image": ab["total"],
                     "achieved_gbs": ab["total"] * B / (ms / args.steps / 1e3) / 1e9,
                     "frac_of_peak": ab["total"] * B / (ms / args.steps / 1e3) / 1e9 / peak_gbs,
                     "stage_ms": stage_ms, "ms_per_step_one_batch_in_flight": serial_ms},
        "clocks": clocks, "gpu_launches": 9 * args.steps,
        "kernels_per_step": ["gt_pack_kernel", "decode_filter_l2_kernel", "nms_kernel", "plan_kernel", "coeff_gather_kernel", "match_kernel",
                             "contract_kernel", "cells_kernel", "finalize_kernel"],
    }
    if e2e:
        line["e2e"] = e2e
    if world == 1 and args.cpu_sample > 0:
        v, dt = cpu_oracle_rate(args, args.cpu_sample, 1)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": f"{args.cpu_sample} images of the same workload, scalar C oracle, {dt:.1f} s"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
