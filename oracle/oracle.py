"""TEST INFRASTRUCTURE ONLY — Python front-end of the CPU oracle (oracle/btpost_oracle.c).

Loads ``oracle/_build/libbtoracle.so`` (built by ``oracle/Makefile`` / ``__graft_entry__.build``)
and strings its functions into the same per-batch pipeline the CUDA library runs, returning plain
numpy arrays in the layout of the library's batched outputs so tests can compare them
element-for-element.  The AP accumulation (`accumulate_ap`) restates pycocotools
``COCOeval.accumulate``/``summarize`` as configured by the reference
(`/root/reference/src/running_main_v2.py:228-251`, `/root/reference/src/evaluate_model.py:81-94`;
SURVEY.md A.3) in float64 numpy.  pycocotools/torchmetrics are not installed: that part is
**parity unpinned** (restated from the published algorithm).

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg may import this.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libbtoracle.so"
_lib = None

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")

AREA_RANGES = np.array([[0.0, 1e10], [0.0, 32.0 ** 2], [32.0 ** 2, 96.0 ** 2], [96.0 ** 2, 1e10]], np.float64)


def build():
    subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            build()
        L = C.CDLL(str(_LIB_PATH))
        L.bto_decode_l2.argtypes = [f32p, C.c_int, C.c_int, f32p, f32p, i32p]
        L.bto_decode_l1_level.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, f32p, f32p]
        L.bto_filter.argtypes = [f32p, f32p, i32p, C.c_int, C.c_float, C.c_int, C.c_float, C.c_float,
                                 f32p, f32p, i32p, i32p]
        L.bto_filter.restype = C.c_int
        L.bto_nms.argtypes = [f32p, f32p, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_float, i64p, C.c_int, C.c_int]
        L.bto_nms.restype = C.c_int
        L.bto_gt_prep.argtypes = [f32p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, f32p, i32p, C.c_int]
        L.bto_gt_prep.restype = C.c_int
        L.bto_cm_match.argtypes = [f32p, i32p, C.c_int, f32p, i32p, C.c_int, C.c_float, C.c_int, i64p, C.c_void_p]
        L.bto_cm_match.restype = C.c_int
        L.bto_project.argtypes = [f32p, f32p, C.c_float, C.c_int, C.c_int, f32p]
        L.bto_upsample.argtypes = [f32p, C.c_int, C.c_int, f32p, C.c_int, C.c_int]
        L.bto_threshold.argtypes = [f32p, C.c_size_t, u8p]
        L.bto_instance_mask.argtypes = [f32p, f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, f32p, u8p]
        L.bto_mask_counts.argtypes = [u8p, u8p, C.c_size_t, i64p, i64p]
        L.bto_dice_iou.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.bto_coco_match.argtypes = [f32p, C.c_int, f32p, C.c_int, f64p, C.c_int, f64p, C.c_int, i32p, u8p, u8p]
        L.bto_expf_public.argtypes = [C.c_float]
        L.bto_expf_public.restype = C.c_float
        _lib = L
    return _lib


DEFAULTS = dict(conf_thres=0.05, iou_thres=0.6, max_det=300, nc=3, nm=32, img_size=640,
                class_mode=0, max_wh=7680.0, clamp=1, gt_mode=0, iou_match_thresh=0.5, max_gt=32,
                crop=1, max_cand=0, with_instances=True, with_masks_out=True, drop_gt_no_cand=0)


def iou_thresholds():
    """fp32 linspace(0.5, 0.95, 10) widened to double (src/running_main_v2.py:247)."""
    import torch
    return np.asarray(torch.linspace(0.5, 0.95, 10).tolist(), np.float64)


def nms(boxes, scores, iou_thr, labels=None, class_mode=0, max_wh=7680.0, max_keep=None, max_cand=0):
    boxes = np.ascontiguousarray(boxes, np.float32).reshape(-1, 4)
    scores = np.ascontiguousarray(scores, np.float32)
    n = len(scores)
    keep = np.zeros(max(n, 1), np.int64)
    lab = None
    if labels is not None:
        lab_arr = np.ascontiguousarray(labels, np.int32)
        lab = lab_arr.ctypes.data_as(C.c_void_p)
    k = lib().bto_nms(boxes, scores, lab, n, float(iou_thr), class_mode, max_wh, keep,
                      n if max_keep is None else max_keep, int(max_cand))
    return keep[:k].copy()


def decode_l2(head_img, nc):
    Ctot, N = head_img.shape
    boxes = np.empty((N, 4), np.float32)
    score = np.empty(N, np.float32)
    label = np.empty(N, np.int32)
    lib().bto_decode_l2(np.ascontiguousarray(head_img), N, nc, boxes, score, label)
    return boxes, score, label


def decode_l1(maps_img, img_size, nc, reg_max=16):
    """maps_img: list of [4R+nc, H, W] for one image -> (boxes [N,4] xyxy, scores [N,nc])."""
    bs, ss = [], []
    for m in maps_img:
        _, H, W = m.shape
        b = np.empty((H * W, 4), np.float32)
        s = np.empty((H * W, nc), np.float32)
        lib().bto_decode_l1_level(np.ascontiguousarray(m), H, W, reg_max, nc, np.float32(img_size / W), b, s)
        bs.append(b)
        ss.append(s)
    return np.concatenate(bs), np.concatenate(ss)


def l1_to_l2(maps, coeffs, img_size, nc, reg_max=16):
    """Ultralytics Detect._inference restatement: L1 maps (+ mask coeffs [B,nm,N])
    -> L2 [B, 4+nc+nm, N] with xywh boxes, using the oracle's own decode arithmetic."""
    B = maps[0].shape[0]
    out = []
    for b in range(B):
        boxes, scores = decode_l1([m[b] for m in maps], img_size, nc, reg_max)
        cxcy = (boxes[:, :2] + boxes[:, 2:]) / np.float32(2)
        wh = boxes[:, 2:] - boxes[:, :2]
        out.append(np.concatenate([cxcy.T, wh.T, scores.T, coeffs[b]], 0))
    return np.stack(out).astype(np.float32)


def gt_prep(det_boxes_gt, b, S, mode, clamp, max_gt=32):
    gt = np.ascontiguousarray(det_boxes_gt, np.float32).reshape(-1, 6)
    ob = np.zeros((max_gt, 4), np.float32)
    ol = np.zeros(max_gt, np.int32)
    g = lib().bto_gt_prep(gt, len(gt), b, np.float32(S), mode, clamp, ob, ol, max_gt)
    return ob[:g].copy(), ol[:g].copy()


def project_upsample(protos_img, w, bias, S):
    nm, ph, pw = protos_img.shape
    logit = np.empty(ph * pw, np.float32)
    lib().bto_project(np.ascontiguousarray(protos_img), np.ascontiguousarray(w, np.float32), np.float32(bias), nm, ph * pw, logit)
    up = np.empty((S, S), np.float32)
    lib().bto_upsample(logit, ph, pw, up, S, S)
    return logit.reshape(ph, pw), up


def threshold(logits):
    m = np.empty(logits.size, np.uint8)
    lib().bto_threshold(np.ascontiguousarray(logits, np.float32).ravel(), logits.size, m)
    return m.reshape(logits.shape)


def instance_mask(protos_img, coeff, box, S, crop=1):
    nm, ph, pw = protos_img.shape
    scratch = np.empty(ph * pw, np.float32)
    mask = np.empty((S, S), np.uint8)
    lib().bto_instance_mask(np.ascontiguousarray(protos_img), np.ascontiguousarray(coeff, np.float32),
                            np.ascontiguousarray(box, np.float32), nm, ph, pw, S, S, crop, scratch, mask)
    return mask


def mask_counts(pred, gt, cnt4):
    img3 = np.zeros(3, np.int64)
    lib().bto_mask_counts(np.ascontiguousarray(pred).ravel(), np.ascontiguousarray(gt).ravel(), pred.size, cnt4, img3)
    return img3


def dice_iou(inter, p, g):
    d, i = C.c_float(), C.c_float()
    lib().bto_dice_iou(int(inter), int(p), int(g), C.byref(d), C.byref(i))
    return d.value, i.value


def coco_match(det_xyxy, gt_xyxy, thrs, area_rng=AREA_RANGES):
    D, G, T, A = len(det_xyxy), len(gt_xyxy), len(thrs), len(area_rng)
    dm = np.zeros((A, T, max(D, 1)), np.int32)
    di = np.zeros((A, T, max(D, 1)), np.uint8)
    gi = np.zeros((A, max(G, 1)), np.uint8)
    lib().bto_coco_match(np.ascontiguousarray(det_xyxy, np.float32).reshape(-1, 4) if D else np.zeros((1, 4), np.float32), D,
                         np.ascontiguousarray(gt_xyxy, np.float32).reshape(-1, 4) if G else np.zeros((1, 4), np.float32), G,
                         np.ascontiguousarray(thrs, np.float64), T, np.ascontiguousarray(area_rng, np.float64), A,
                         dm.reshape(A, T, -1) if D else dm, di, gi)
    return dm[:, :, :D], di[:, :, :D], gi[:, :G]


def run_pipeline(batch, pool=None, **kw):
    """Whole hot path on one batch (L2 head).  Returns numpy outputs in the CUDA library's layout."""
    p = dict(DEFAULTS)
    p.update(kw)
    l1 = p.get("layout", "l2") == "l1"
    head, protos = (None if l1 else batch["head"]), batch["protos"]
    gt_rows, masks_gt = batch["det_boxes_gt"], batch["masks_gt"]
    if l1:   # reference layout: three raw maps + Segment `mc` coefficients [B, nm, N]
        maps, coeffs = batch["maps"], batch["coeffs"]
        B, N = maps[0].shape[0], sum(m.shape[2] * m.shape[3] for m in maps)
    else:
        B, _, N = head.shape
    nc, nm, S, max_det = p["nc"], p["nm"], p["img_size"], p["max_det"]
    out = {
        "n_cand": np.zeros(B, np.int32), "det_count": np.zeros(B, np.int32),
        "dets": np.zeros((B, max_det, 6), np.float32), "det_anchor": np.full((B, max_det), -1, np.int32),
        "det_keep": np.full((B, max_det), -1, np.int32), "det_coeff": np.zeros((B, max_det, nm), np.float32),
        "cand_box": [], "cand_score": [], "cand_label": [], "cand_anchor": [],
        "gt_count": np.zeros(B, np.int32), "gt_boxes": np.zeros((B, p["max_gt"], 4), np.float32),
        "gt_boxes_raw": np.zeros((B, p["max_gt"], 4), np.float32),
        "gt_labels": np.zeros((B, p["max_gt"]), np.int32),
        "cm": np.zeros((nc, nc), np.int64), "cm_pos": np.zeros(B, np.int32),
        "seg_cnt4": np.zeros(4, np.int64), "seg_img3": np.zeros((B, 3), np.int64),
        "seg_dice": np.zeros(B, np.float32), "seg_iou": np.zeros(B, np.float32),
        "seg_prob_sum": np.zeros(B, np.float64), "seg_map_score": np.zeros(B, np.float32),
        "seg_mask": np.zeros((B, S, S), np.uint8) if p["with_masks_out"] else None,
        "seg_logits": np.zeros((B, S, S), np.float32) if p["with_masks_out"] else None,
        "inst_area": np.zeros((B, max_det), np.int32), "inst_inter": np.zeros((B, max_det), np.int32),
        "inst_masks": [],
        "uni_cnt4": np.zeros(4, np.int64), "uni_img3": np.zeros((B, 3), np.int64),
        "uni_dice": np.zeros(B, np.float32), "uni_iou": np.zeros(B, np.float32),
        "uni_mask": np.zeros((B, S, S), np.uint8) if p["with_masks_out"] else None,
    }
    thrs = iou_thresholds()
    A, T = len(AREA_RANGES), len(thrs)
    out["dt_match"] = np.zeros((B, A, T, max_det), np.int32)
    out["dt_ignore"] = np.zeros((B, A, T, max_det), np.uint8)
    out["gt_ignore"] = np.zeros((B, A, p["max_gt"]), np.uint8)
    for b in range(B):
        if l1:
            boxes, sc_all = decode_l1([m[b] for m in maps], S, nc, p.get("reg_max", 16))
            label = np.argmax(sc_all, axis=1).astype(np.int32)          # first max on ties (torch .max(dim=1))
            score = np.ascontiguousarray(sc_all[np.arange(N), label])
        else:
            boxes, score, label = decode_l2(head[b], nc)
        cb = np.empty((N, 4), np.float32); cs = np.empty(N, np.float32)
        cl = np.empty(N, np.int32); ca = np.empty(N, np.int32)
        m = lib().bto_filter(boxes, score, label, N, np.float32(p["conf_thres"]), p["clamp"], np.float32(S), np.float32(S), cb, cs, cl, ca)
        cb, cs, cl, ca = cb[:m].copy(), cs[:m].copy(), cl[:m].copy(), ca[:m].copy()
        out["n_cand"][b] = m
        out["cand_box"].append(cb); out["cand_score"].append(cs); out["cand_label"].append(cl); out["cand_anchor"].append(ca)
        keep = nms(cb, cs, p["iou_thres"], cl, p["class_mode"], p["max_wh"], max_det, p["max_cand"]) if m else np.zeros(0, np.int64)
        k = len(keep)
        out["det_count"][b] = k
        out["dets"][b, :k, :4] = cb[keep]; out["dets"][b, :k, 4] = cs[keep]; out["dets"][b, :k, 5] = cl[keep].astype(np.float32)
        out["det_anchor"][b, :k] = ca[keep]; out["det_keep"][b, :k] = keep
        csrc = coeffs[b] if l1 else head[b, 4 + nc:]
        out["det_coeff"][b, :k] = csrc[:, ca[keep]].T.reshape(k, nm) if k else 0
        # GT (mAP copy clamped, loss copy unclamped) + CM matching on raw decoded boxes
        gb, gl = gt_prep(gt_rows, b, S, p["gt_mode"], 1, p["max_gt"])
        gbr, _ = gt_prep(gt_rows, b, S, p["gt_mode"], 0, p["max_gt"])
        g = len(gl)
        out["gt_count"][b] = g
        out["gt_boxes"][b, :g] = gb; out["gt_boxes_raw"][b, :g] = gbr; out["gt_labels"][b, :g] = gl
        if g:
            out["cm_pos"][b] = lib().bto_cm_match(boxes, label, N, gbr, gl, g, np.float32(p["iou_match_thresh"]), nc, out["cm"], None)
        # M1 projector mask + counters
        gtm = np.ascontiguousarray(masks_gt[b, 0])
        _, up = project_upsample(protos[b], batch["proj_weight"], batch["proj_bias"], S)
        pm = threshold(up)
        img3 = mask_counts(pm, gtm, out["seg_cnt4"])
        out["seg_img3"][b] = img3
        out["seg_dice"][b], out["seg_iou"][b] = dice_iou(*img3)
        out["seg_prob_sum"][b], out["seg_map_score"][b] = seg_map_score(up, pm)
        if p["with_masks_out"]:
            out["seg_mask"][b] = pm; out["seg_logits"][b] = up
        # M2 instance masks
        if p["with_instances"]:
            im = np.zeros((k, S, S), np.uint8)

            def _one(i, b=b, im=im, gtm=gtm):
                im[i] = instance_mask(protos[b], out["det_coeff"][b, i], out["dets"][b, i, :4], S, p["crop"])
                out["inst_area"][b, i] = int(im[i].sum()); out["inst_inter"][b, i] = int((im[i] & gtm).sum())

            if pool is not None:          # ctypes releases the GIL: instances run on all host threads
                list(pool.map(_one, range(k)))
            else:
                for i in range(k):
                    _one(i)
            out["inst_masks"].append(im)
            # union of the instance masks = the image-level mask the per-image Dice / IoU is taken on
            um = im.any(0).astype(np.uint8) if k else np.zeros((S, S), np.uint8)
            u3 = mask_counts(um, gtm, out["uni_cnt4"])
            out["uni_img3"][b] = u3
            out["uni_dice"][b], out["uni_iou"][b] = dice_iou(*u3)
            if p["with_masks_out"]:
                out["uni_mask"][b] = um
        # COCO matching per class
        labels = out["dets"][b, :k, 5].astype(np.int32)
        for c in range(nc):
            di_ = np.nonzero(labels == c)[0]
            gi_ = np.nonzero(gl == c)[0]
            if len(di_) == 0 and len(gi_) == 0:
                continue
            dm, dig, gig = coco_match(out["dets"][b, di_, :4], gb[gi_], thrs)
            # matched index refers to the per-class GT list; map back to the image's GT index + 1
            mapped = np.where(dm > 0, gi_[np.maximum(dm - 1, 0)] + 1 if len(gi_) else 0, 0)
            out["dt_match"][b][:, :, di_] = mapped
            out["dt_ignore"][b][:, :, di_] = dig
            out["gt_ignore"][b][:, gi_] = gig
    return out


# ------------------------------------------------------------------ AP accumulation (a9)
def seg_map_score(logits, pred_mask):
    """v3 segmentation-mAP prep (running_main_v3.py:478-486): score = (probs * mask).sum() / (mask.sum() + 1e-6)
    with probs = sigmoid(logits) in fp32.  Returns (sum of the foreground probabilities as float64, fp32 score)."""
    x = logits.astype(np.float32)
    probs = (np.float32(1.0) / (np.float32(1.0) + np.exp(-x, dtype=np.float32))).astype(np.float32)
    m = pred_mask.astype(bool)
    psum = float(probs[m].astype(np.float64).sum())
    score = np.float32(np.float32(psum) / (np.float32(m.sum()) + np.float32(1e-6)))
    return psum, score


def seg_map_records(seg_img3, scores, thrs):
    """What torchmetrics MeanAveragePrecision(iou_type='segm') -> COCOeval.evaluateImg does with v3's ONE predicted
    mask and ONE target mask per image, class 0 (running_main_v3.py:478-498; SURVEY.md A.3 with mask IoU =
    inter / union, areas = mask pixel counts).  Returns the arrays of the detection pipeline's layout with K = G = 1."""
    seg_img3 = np.asarray(seg_img3, np.int64)
    B, T = seg_img3.shape[0], len(thrs)
    inter, P, G = (seg_img3[:, i].astype(np.float64) for i in range(3))
    union = P + G - inter
    iou = np.where(union > 0, inter / np.where(union > 0, union, 1.0), 0.0)
    lo = np.array([0.0, 0.0, 32.0 ** 2, 96.0 ** 2]); hi = np.array([1e10, 32.0 ** 2, 96.0 ** 2, 1e10])
    gt_ig = (G[:, None] < lo[None]) | (G[:, None] > hi[None])                       # [B, A]
    dt_out = (P[:, None] < lo[None]) | (P[:, None] > hi[None])
    thr = np.minimum(np.asarray(thrs, np.float64), 1 - 1e-10)
    hit = iou[:, None] >= thr[None]                                                  # [B, T]
    dt_match = np.broadcast_to(hit[:, None, :], (B, 4, T)).astype(np.int32)[..., None]
    dt_ignore = np.where(hit[:, None, :], gt_ig[:, :, None], dt_out[:, :, None]).astype(np.uint8)[..., None]
    dets = np.zeros((B, 1, 6), np.float32); dets[:, 0, 4] = scores
    return dict(dets=dets, det_count=np.ones(B, np.int32), dt_match=np.ascontiguousarray(dt_match),
                dt_ignore=np.ascontiguousarray(dt_ignore), gt_ignore=gt_ig.astype(np.uint8)[..., None],
                gt_labels=np.zeros((B, 1), np.int32), gt_count=np.ones(B, np.int32))


def accumulate_ap(records, n_gt_valid, thrs, max_dets=(1, 10, 100), nc=3):
    """COCOeval.accumulate + summarize from per-detection match records.

    records: list over images of dict(labels[K], scores[K], matched[A,T,K] bool, ignored[A,T,K] bool)
             in per-image descending-score order;  n_gt_valid[a][c] = non-ignored GT count.
    Returns dict with map, map_50, map_75, map_small/medium/large, mar_<k>, per-class arrays.
    """
    T, A, M = len(thrs), 4, len(max_dets)
    rec_thrs = np.linspace(0.0, 1.0, 101)
    precision = -np.ones((T, 101, nc, A, M))
    recall = -np.ones((T, nc, A, M))
    for c in range(nc):
        for a in range(A):
            npig = int(n_gt_valid[a][c])
            for mi, md in enumerate(max_dets):
                sc, tpm, igm = [], [], []
                for r in records:
                    sel = np.nonzero(r["labels"] == c)[0][:md]
                    sc.append(r["scores"][sel]); tpm.append(r["matched"][a][:, sel]); igm.append(r["ignored"][a][:, sel])
                if not sc or npig == 0:
                    continue
                sc = np.concatenate(sc)
                if True:
                    order = np.argsort(-sc, kind="mergesort")
                    tpm_ = np.concatenate(tpm, 1)[:, order]; igm_ = np.concatenate(igm, 1)[:, order]
                    tps = np.logical_and(tpm_, ~igm_); fps = np.logical_and(~tpm_, ~igm_)
                    tp_sum = np.cumsum(tps, 1).astype(np.float64); fp_sum = np.cumsum(fps, 1).astype(np.float64)
                    for t in range(T):
                        tp, fp = tp_sum[t], fp_sum[t]
                        nd = len(tp)
                        rc = tp / npig
                        pr = tp / (fp + tp + np.spacing(1))
                        q = np.zeros(101)
                        recall[t, c, a, mi] = rc[-1] if nd else 0
                        pr = pr.tolist(); q = q.tolist()
                        for i in range(nd - 1, 0, -1):
                            if pr[i] > pr[i - 1]:
                                pr[i - 1] = pr[i]
                        inds = np.searchsorted(rc, rec_thrs, side="left")
                        try:
                            for ri, pi in enumerate(inds):
                                q[ri] = pr[pi]
                        except IndexError:
                            pass
                        precision[t, :, c, a, mi] = np.array(q)

    def _mean(x):
        x = x[x > -1]
        return float(x.mean()) if x.size else -1.0

    def _thr_idx(v):
        hit = np.nonzero(np.isclose(thrs, v))[0]
        return int(hit[0]) if len(hit) else None

    res = {"map": _mean(precision[:, :, :, 0, -1])}
    for name, v in (("map_50", 0.5), ("map_75", 0.75)):
        ti = _thr_idx(v)
        res[name] = _mean(precision[ti, :, :, 0, -1]) if ti is not None else -1.0
    for ai, name in ((1, "small"), (2, "medium"), (3, "large")):
        res[f"map_{name}"] = _mean(precision[:, :, :, ai, -1])
        res[f"mar_{name}"] = _mean(recall[:, :, ai, -1])
    for mi, md in enumerate(max_dets):
        res[f"mar_{md}"] = _mean(recall[:, :, 0, mi])
    res["map_per_class"] = np.array([_mean(precision[:, :, c, 0, -1]) for c in range(nc)])
    res[f"mar_{max_dets[-1]}_per_class"] = np.array([_mean(recall[:, c, 0, -1]) for c in range(nc)])
    res["precision"] = precision
    res["recall"] = recall
    return res
