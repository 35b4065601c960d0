"""Host-side mirror of the reference's post-processing / evaluation interface.

The reference has one operator boundary for this path:

    map_preds, map_targets, det_log_preds, det_log_gts = \
        model._prepare_det_outputs_for_metrics_and_logging(det_outputs, det_boxes_gt, device, batch_size)

(`/root/reference/src/evaluate_model.py:174-178`; behaviour = `running_main_v2.py:720-882`), plus the
segmentation block `running_main_v2.py:672-713`.  This module keeps those names, argument meanings and
return formats, and adds a batched `PostProcessor.run` that returns padded device tensors so the
performance path never builds ragged Python lists.  All arithmetic happens in ``libbtpost.so``
(hand-written sm_100a CUDA) behind the C ABI of ``include/btpost.h``; PyTorch only owns the device
memory and the stream.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from dataclasses import dataclass, field

import torch

from . import _lib
from ._lib import BtIO, BtParams

# Module constants with the reference's names and defaults (running_main_v2.py:51-53); overridable.
CONF_TH = 0.05
NMS_IOU = 0.6
TOP_K = 300


def map_iou_thresholds():
    """fp32 linspace(0.5, 0.95, 10) widened to double (running_main_v2.py:247)."""
    return [float(v) for v in torch.linspace(0.5, 0.95, 10).tolist()]


def num_anchors(img_size: int) -> int:
    return sum((img_size // s) ** 2 for s in (8, 16, 32))


@dataclass
class PostConfig:
    batch: int
    img_size: int = 640
    nc: int = 3
    nm: int = 32
    reg_max: int = 16
    conf_thres: float | None = None      # None -> module CONF_TH at call time
    iou_thres: float | None = None       # None -> module NMS_IOU
    max_det: int | None = None           # None -> module TOP_K
    max_cand: int = 0
    class_mode: int = _lib.CLASS_AGNOSTIC
    max_wh: float = 7680.0
    clamp_boxes: bool = True
    gt_mode: int = _lib.GT_LITERAL
    max_gt: int = 32
    iou_match_thresh: float = 0.5
    crop: bool = True
    layout: int = _lib.LAYOUT_L2
    gt_mask_dtype: int = _lib.MASK_U8
    iou_thrs: list = field(default_factory=map_iou_thresholds)
    # optional dense outputs
    with_seg_mask: bool = False
    with_seg_logits: bool = False
    with_uni_mask: bool = False
    with_coco: bool = True
    with_seg_map: bool = False           # v3 segmentation-mAP prep (score numerator per image)
    num_anchors: int | None = None
    nms_threads: int = 0                 # 0 = auto; 512 / 1024 force a variant of the NMS kernel
    proto_bf16: bool = False             # prototypes arrive as torch.bfloat16 (widened exactly in the kernel)
    head_bf16: bool = False              # same for the L2 head tensor


class PostProcessor:
    """Owns params, workspace and output tensors for one (batch, shape) configuration on one GPU."""

    def __init__(self, cfg: PostConfig, device="cuda:0"):
        self.lib = _lib.load()
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("btpost runs on CUDA devices only (sm_100a); there is no CPU path")
        B, S = cfg.batch, cfg.img_size
        K = cfg.max_det if cfg.max_det is not None else TOP_K
        N = cfg.num_anchors if cfg.num_anchors is not None else num_anchors(S)
        self.B, self.S, self.K, self.N = B, S, K, N
        p = BtParams()
        p.batch, p.num_anchors, p.nc, p.nm, p.reg_max = B, N, cfg.nc, cfg.nm, cfg.reg_max
        p.img_h = p.img_w = S
        p.proto_h = p.proto_w = S // 4
        p.layout = cfg.layout
        p.conf_thres = cfg.conf_thres if cfg.conf_thres is not None else CONF_TH
        p.iou_thres = cfg.iou_thres if cfg.iou_thres is not None else NMS_IOU
        p.max_det, p.max_cand, p.class_mode, p.max_wh = K, cfg.max_cand, cfg.class_mode, cfg.max_wh
        p.clamp_boxes, p.gt_mode, p.max_gt = int(cfg.clamp_boxes), cfg.gt_mode, cfg.max_gt
        p.iou_match_thresh, p.crop, p.gt_mask_dtype = cfg.iou_match_thresh, int(cfg.crop), cfg.gt_mask_dtype
        p.num_iou_thrs = len(cfg.iou_thrs)
        p.nms_threads = cfg.nms_threads
        p.proto_dtype = _lib.PROTO_BF16 if cfg.proto_bf16 else _lib.PROTO_F32
        p.head_dtype = _lib.HEAD_BF16 if cfg.head_bf16 else _lib.HEAD_F32
        for i, v in enumerate(cfg.iou_thrs):
            p.iou_thrs[i] = v
        self.params = p
        nbytes = C.c_size_t()
        _lib.check(self.lib.btpost_workspace_bytes(C.byref(p), C.byref(nbytes)), "btpost_workspace_bytes")
        dev = self.device
        self.workspace = torch.empty(nbytes.value + 256, dtype=torch.uint8, device=dev)
        self._ws_ptr = (self.workspace.data_ptr() + 255) // 256 * 256
        self._ws_bytes = nbytes.value
        G, A, T, nc, nm = cfg.max_gt, _lib.BT_NUM_AREA, len(cfg.iou_thrs), cfg.nc, cfg.nm
        z = lambda *shape, dtype: torch.zeros(*shape, dtype=dtype, device=dev)
        o = {
            "det_count": z(B, dtype=torch.int32), "dets": z(B, K, 6, dtype=torch.float32),
            "det_keep": z(B, K, dtype=torch.int64), "det_anchor": z(B, K, dtype=torch.int32),
            "det_coeff": z(B, K, nm, dtype=torch.float32), "n_cand": z(B, dtype=torch.int32),
            "gt_count": z(B, dtype=torch.int32), "gt_boxes": z(B, G, 4, dtype=torch.float32),
            "gt_boxes_raw": z(B, G, 4, dtype=torch.float32), "gt_labels": z(B, G, dtype=torch.int32),
            "cm": z(nc, nc, dtype=torch.int64), "seg_cnt4": z(4, dtype=torch.int64), "uni_cnt4": z(4, dtype=torch.int64),
            "cm_pos": z(B, dtype=torch.int32), "seg_img3": z(B, 3, dtype=torch.int64),
            "seg_dice": z(B, dtype=torch.float32), "seg_iou": z(B, dtype=torch.float32),
            "uni_img3": z(B, 3, dtype=torch.int64), "uni_dice": z(B, dtype=torch.float32),
            "uni_iou": z(B, dtype=torch.float32), "inst_area": z(B, K, dtype=torch.int32),
            "inst_inter": z(B, K, dtype=torch.int32),
        }
        if cfg.with_seg_mask:
            o["seg_mask"] = z(B, S, S, dtype=torch.uint8)
        if cfg.with_seg_logits:
            o["seg_logits"] = z(B, S, S, dtype=torch.float32)
        if cfg.with_uni_mask:
            o["uni_mask"] = z(B, S, S, dtype=torch.uint8)
        if cfg.with_coco:
            o["dt_match"] = z(B, A, T, K, dtype=torch.int32)
            o["dt_ignore"] = z(B, A, T, K, dtype=torch.uint8)
            o["gt_ignore"] = z(B, A, G, dtype=torch.uint8)
        if cfg.with_seg_map:
            o["seg_prob_sum"] = z(B, dtype=torch.float64)
        self.out = o
        self._empty_gt = torch.zeros(1, 6, dtype=torch.float32, device=dev)

    # -- metric state ---------------------------------------------------------------------------
    def reset_metrics(self):
        """Zero the accumulated counters (confusion matrix, global pixel tp/fp/fn/tn)."""
        for k in ("cm", "seg_cnt4", "uni_cnt4"):
            self.out[k].zero_()

    # -- launch ---------------------------------------------------------------------------------
    def _check_in(self, t, shape, dtype, name):
        if t.device != self.device or t.dtype != dtype or tuple(t.shape) != tuple(shape) or not t.is_contiguous():
            raise ValueError(f"{name}: expected contiguous {dtype} {tuple(shape)} on {self.device}, "
                             f"got {t.dtype} {tuple(t.shape)} on {t.device}")

    def _io(self, head, protos, det_boxes_gt, masks_gt, proj_weight, maps=None, coeffs=None):
        io = BtIO()
        cfg, B, S, N = self.cfg, self.B, self.S, self.N
        if cfg.layout == _lib.LAYOUT_L2:
            self._check_in(head, (B, 4 + cfg.nc + cfg.nm, N), torch.bfloat16 if cfg.head_bf16 else torch.float32, "head")
            io.head = head.data_ptr()
        else:
            for l, (m, s) in enumerate(zip(maps, (8, 16, 32))):
                self._check_in(m, (B, 4 * cfg.reg_max + cfg.nc, S // s, S // s), torch.float32, f"maps[{l}]")
                setattr(io, f"maps{l}", m.data_ptr())
            if coeffs is not None:
                self._check_in(coeffs, (B, cfg.nm, N), torch.float32, "coeffs")
                io.coeffs = coeffs.data_ptr()
        self._check_in(protos, (B, cfg.nm, S // 4, S // 4), torch.bfloat16 if cfg.proto_bf16 else torch.float32, "protos")
        io.protos = protos.data_ptr()
        if det_boxes_gt is None or det_boxes_gt.numel() == 0:
            det_boxes_gt, rows = self._empty_gt, 0
        else:
            rows = det_boxes_gt.shape[0]
            self._check_in(det_boxes_gt, (rows, 6), torch.float32, "det_boxes_gt")
        io.det_boxes_gt = det_boxes_gt.data_ptr()
        self.params.num_gt_rows = rows
        mdt = torch.uint8 if cfg.gt_mask_dtype == _lib.MASK_U8 else torch.float32
        if masks_gt.dim() == 3:
            masks_gt = masks_gt.unsqueeze(1)
        self._check_in(masks_gt, (B, 1, S, S), mdt, "masks_gt")
        io.masks_gt = masks_gt.data_ptr()
        self._check_in(proj_weight, (cfg.nm,), torch.float32, "proj_weight")
        io.proj_weight = proj_weight.data_ptr()
        for k, v in self.out.items():
            setattr(io, k, v.data_ptr())
        self._keepalive = (head, protos, det_boxes_gt, masks_gt, proj_weight, maps, coeffs)
        return io

    def run(self, head, protos, det_boxes_gt, masks_gt, proj_weight, proj_bias=0.0, *, maps=None, coeffs=None,
            stage="run", stream=None):
        """Enqueue the hot path for one batch on the current (or given) CUDA stream; returns the
        dict of output tensors (device, padded to ``max_det`` / ``max_gt``; see include/btpost.h)."""
        # module constants are read at call time, as the reference reads its globals inside the loop
        if self.cfg.conf_thres is None:
            self.params.conf_thres = CONF_TH
        if self.cfg.iou_thres is None:
            self.params.iou_thres = NMS_IOU
        self.params.proj_bias = float(proj_bias)
        io = self._io(head, protos, det_boxes_gt, masks_gt, proj_weight, maps, coeffs)
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        parts = {"masks_pack": _lib.MASKS_PACK, "masks_contract": _lib.MASKS_CONTRACT, "masks_cells": _lib.MASKS_CELLS}.get(stage)
        with torch.cuda.device(self.device):
            if parts is not None:   # one kernel group of the mask stage (bench.py times the HBM-side kernel alone)
                rc = self.lib.btpost_masks_parts(C.byref(self.params), C.byref(io), C.c_void_p(self._ws_ptr),
                                                 C.c_size_t(self._ws_bytes), C.c_void_p(st.cuda_stream), C.c_int(parts))
            else:
                fn = getattr(self.lib, f"btpost_{stage}")
                rc = fn(C.byref(self.params), C.byref(io), C.c_void_p(self._ws_ptr), C.c_size_t(self._ws_bytes),
                        C.c_void_p(st.cuda_stream))
        _lib.check(rc, f"btpost_{stage}")
        return self.out

    def capture(self, head, protos, det_boxes_gt, masks_gt, proj_weight, proj_bias=0.0, **kw):
        """Capture one step on these (static) buffers into a CUDA graph; returns the graph, whose
        ``replay()`` re-runs the whole hot path with one host call (the library is capture-safe:
        no allocation, no synchronisation, caller's stream only)."""
        self.run(head, protos, det_boxes_gt, masks_gt, proj_weight, proj_bias, **kw)   # warm-up: attributes, module load
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.run(head, protos, det_boxes_gt, masks_gt, proj_weight, proj_bias, **kw)
        return g

    # -- reference-shaped views -----------------------------------------------------------------
    def to_reference_lists(self, out=None):
        """(map_preds, map_targets, det_log_preds, det_log_gts) exactly as
        `running_main_v2.py:720-882` builds them: CPU dicts for torchmetrics, device [K,6]/[G,5]
        tensors for `log_det_examples` (`multitask_logging.py:216,240`)."""
        o = out or self.out
        counts = o["det_count"].cpu().tolist()       # one D->H sync for the whole batch
        gcounts = o["gt_count"].cpu().tolist()
        ncand = o["n_cand"].cpu().tolist()
        dets_cpu, gtb_cpu, gtl_cpu = o["dets"].cpu(), o["gt_boxes"].cpu(), o["gt_labels"].cpu()
        map_preds, map_targets, log_preds, log_gts = [], [], [], []
        for b in range(self.B):
            k, g = counts[b], gcounts[b]
            if ncand[b] == 0 and self.cfg.gt_mode == _lib.GT_LITERAL:
                # v2 appends an empty target too when no box passes CONF_TH (running_main_v2.py:797-814)
                g = 0
            d = dets_cpu[b, :k]
            map_preds.append({"boxes": d[:, :4].clone(), "scores": d[:, 4].clone(), "labels": d[:, 5].long()})
            log_preds.append(o["dets"][b, :k])
            map_targets.append({"boxes": gtb_cpu[b, :g].clone(), "labels": gtl_cpu[b, :g].long()})
            log_gts.append(torch.cat([o["gt_boxes"][b, :g], o["gt_labels"][b, :g].float().unsqueeze(1)], dim=1))
        return map_preds, map_targets, log_preds, log_gts


class Pipeline:
    """``depth`` independent PostProcessors (own workspace and outputs each) whose captured steps are replayed
    round-robin on ``depth`` streams: consecutive batches overlap on the GPU, so the NMS of batch i+1 (one CTA per
    image, most SMs idle) runs under the mask kernels of batch i.  Every step still does all of its work; per-step
    outputs are read from ``procs[i % depth].out`` after ``join()``; the accumulated counters (cm, seg_cnt4,
    uni_cnt4) are the sums over the processors (``counters()``)."""

    def __init__(self, cfg: PostConfig, device="cuda:0", depth: int = 2):
        conf = cfg.conf_thres if cfg.conf_thres is not None else CONF_TH
        if depth > 1 and cfg.nms_threads == 0 and conf >= 0.01:
            # small-footprint NMS kernel: its CTAs share SMs with the mask kernels of the other batches in flight
            # (dense candidate lists keep the 1024-thread variant, which sorts them in registers)
            cfg = dataclasses.replace(cfg, nms_threads=512)
        self.procs = [PostProcessor(cfg, device) for _ in range(depth)]
        self.device = self.procs[0].device
        self.streams = [torch.cuda.Stream(self.device) for _ in range(depth)]
        self.graphs = []
        self._i = 0

    def capture(self, *args, **kw):
        self.graphs = [p.capture(*args, **kw) for p in self.procs]
        return self

    def fork(self):
        cur = torch.cuda.current_stream(self.device)
        for st in self.streams:
            st.wait_stream(cur)

    def replay(self):
        i = self._i % len(self.procs)
        self._i += 1
        with torch.cuda.stream(self.streams[i]):
            self.graphs[i].replay()
        return self.procs[i]

    def join(self):
        cur = torch.cuda.current_stream(self.device)
        for st in self.streams:
            cur.wait_stream(st)

    def reset_metrics(self):
        for p in self.procs:
            p.reset_metrics()

    def counters(self, key):
        return sum(p.out[key] for p in self.procs)


_cached: dict = {}


def prepare_det_outputs_for_metrics_and_logging(det_outputs, det_boxes_gt, device, batch_size, *, img_size=640,
                                                nc=3, protos=None, masks_gt=None, proj_weight=None, proj_bias=0.0):
    """Drop-in for ``MultiTaskLitModel._prepare_det_outputs_for_metrics_and_logging``
    (`/root/reference/src/evaluate_model.py:174-178`).

    ``det_outputs`` is either the reference's list of three raw maps ``[B, 4*16+nc, H_l, W_l]`` (L1) or the
    concatenated ``[B, 4+nc+32, N]`` tensor (L2, ``segment_preds_cat``).  Returns
    ``(map_preds, map_targets, det_log_preds, det_log_gts)`` in the reference's formats.  Only the
    detection stages run (decode/filter + NMS/matching); the mask stage needs prototypes and is
    reached through ``PostProcessor.run``.
    """
    device = torch.device(device)
    l1 = isinstance(det_outputs, (list, tuple))
    key = (batch_size, img_size, nc, l1, str(device), CONF_TH, NMS_IOU, TOP_K)
    pp = _cached.get(key)
    if pp is None:
        cfg = PostConfig(batch=batch_size, img_size=img_size, nc=nc,
                         layout=_lib.LAYOUT_L1 if l1 else _lib.LAYOUT_L2)
        pp = _cached[key] = PostProcessor(cfg, device)
    S = img_size
    dummy_protos = getattr(pp, "_dummy_protos", None)
    if dummy_protos is None:
        pp._dummy_protos = torch.zeros(batch_size, 32, S // 4, S // 4, device=device)
        pp._dummy_masks = torch.zeros(batch_size, 1, S, S, dtype=torch.uint8, device=device)
        pp._dummy_w = torch.zeros(32, device=device)
        pp._dummy_coeffs = torch.zeros(batch_size, 32, pp.N, device=device)
    kw = dict(maps=[m.contiguous().float() for m in det_outputs], coeffs=pp._dummy_coeffs) if l1 else {}
    head = None if l1 else det_outputs.contiguous().float()
    gt = det_boxes_gt.to(device).contiguous().float() if det_boxes_gt is not None else None
    pp.run(head, pp._dummy_protos, gt, pp._dummy_masks, pp._dummy_w, 0.0, stage="decode_filter", **kw)
    pp.run(head, pp._dummy_protos, gt, pp._dummy_masks, pp._dummy_w, 0.0, stage="nms_match", **kw)
    return pp.to_reference_lists()
