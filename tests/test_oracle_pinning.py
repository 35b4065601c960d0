"""CPU: pins the oracle's third-party arithmetic against what is installed here (torch 2.11 /
torchvision 0.26 CPU) and against the reference's own helper statements restated in torch.
SURVEY.md §8c known-answer tests k1..k12."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F
import torchvision

from oracle import oracle


def tv_nms(b, s, thr):
    return torchvision.ops.nms(torch.from_numpy(np.asarray(b, np.float32)).reshape(-1, 4),
                               torch.from_numpy(np.asarray(s, np.float32)), thr).numpy()


# ---------------------------------------------------------------- NMS (a4)
def test_nms_known_answers():
    A, B = [0, 0, 10, 10], [20, 20, 30, 30]
    np.testing.assert_array_equal(oracle.nms([A, B, A, B], [0.5] * 4, 0.5), [0, 1])                    # k1 stable ties
    np.testing.assert_array_equal(oracle.nms([[0, 0, 2, 2], [0, 0, 2, 1]], [0.9, 0.8], 0.5), [0, 1])    # k2 IoU == thr
    np.testing.assert_array_equal(oracle.nms([[0, 0, 2, 2], [0, 0, 2, 1]], [0.9, 0.8], 0.4999), [0])
    np.testing.assert_array_equal(oracle.nms([[5, 5, 5, 5], [5, 5, 5, 5]], [0.9, 0.8], 0.5), [0, 1])    # k3 0/0 = NaN
    np.testing.assert_array_equal(oracle.nms([A, B, [40, 40, 50, 50]], [0.1, np.nan, 0.3], 0.5), [1, 2, 0])  # k4 NaN first
    assert oracle.nms(np.zeros((0, 4)), np.zeros(0), 0.5).shape == (0,)                                # k5 empty
    assert oracle.nms([A], [1.0], 0.5).dtype == np.int64                                               # k6
    for args in (([A, B, A, B], [0.5] * 4, 0.5), ([[0, 0, 2, 2], [0, 0, 2, 1]], [0.9, 0.8], 0.5),
                 ([[5, 5, 5, 5], [5, 5, 5, 5]], [0.9, 0.8], 0.5), ([A, B, [40, 40, 50, 50]], [0.1, np.nan, 0.3], 0.5)):
        np.testing.assert_array_equal(oracle.nms(*args), tv_nms(*args))


@pytest.mark.parametrize("seed", range(6))
def test_nms_random_matches_torchvision(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 1500))
    xy = rng.uniform(0, 600, (n, 2)).astype(np.float32)
    wh = rng.uniform(1, 120, (n, 2)).astype(np.float32)
    boxes = np.concatenate([xy, xy + wh], 1)
    if seed % 2:   # clusters of near-duplicates and exact score ties
        boxes[n // 2:] = boxes[: n - n // 2] + rng.normal(0, 2, (n - n // 2, 4)).astype(np.float32)
    scores = np.round(rng.uniform(0, 1, n), 2 if seed % 3 == 0 else 6).astype(np.float32)
    for thr in (0.3, 0.6, 0.45):
        np.testing.assert_array_equal(oracle.nms(boxes, scores, thr), tv_nms(boxes, scores, thr))
        np.testing.assert_array_equal(oracle.nms(boxes, scores, thr, max_keep=50), tv_nms(boxes, scores, thr)[:50])


def test_class_aware_nms_equals_per_class_merge():
    rng = np.random.default_rng(3)
    n = 400
    xy = rng.uniform(0, 300, (n, 2)).astype(np.float32)
    boxes = np.concatenate([xy, xy + rng.uniform(5, 80, (n, 2)).astype(np.float32)], 1)
    scores = rng.uniform(0, 1, n).astype(np.float32)
    labels = rng.integers(0, 3, n).astype(np.int32)
    got = oracle.nms(boxes, scores, 0.5, labels, class_mode=1)
    keep = np.concatenate([np.nonzero(labels == c)[0][tv_nms(boxes[labels == c], scores[labels == c], 0.5)] for c in range(3)])
    keep = keep[np.argsort(-scores[keep], kind="stable")]
    np.testing.assert_array_equal(np.sort(got), np.sort(keep))
    np.testing.assert_array_equal(scores[got], scores[keep])
    # Ultralytics offset trick == agnostic NMS on boxes + label * max_wh
    off = boxes + (labels.astype(np.float32) * np.float32(7680))[:, None]
    np.testing.assert_array_equal(oracle.nms(boxes, scores, 0.5, labels, class_mode=2), tv_nms(off, scores, 0.5))


# ---------------------------------------------------------------- helpers of the reference (a2, a10)
def ref_batch_bbox_iou(b1, b2, eps=1e-7):
    """running_main_v2.py:68-94 restated."""
    ix1 = torch.max(b1[:, 0].unsqueeze(1), b2[:, 0].unsqueeze(0)); iy1 = torch.max(b1[:, 1].unsqueeze(1), b2[:, 1].unsqueeze(0))
    ix2 = torch.min(b1[:, 2].unsqueeze(1), b2[:, 2].unsqueeze(0)); iy2 = torch.min(b1[:, 3].unsqueeze(1), b2[:, 3].unsqueeze(0))
    inter = (ix2 - ix1).clamp(min=0) * (iy2 - iy1).clamp(min=0)
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1]); a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    return inter / (a1.unsqueeze(1) + a2.unsqueeze(0) - inter + eps)


def test_cm_match_follows_batch_bbox_iou():
    rng = np.random.default_rng(0)
    N, G, nc = 3000, 3, 3
    xy = rng.uniform(0, 500, (N, 2)).astype(np.float32)
    boxes = np.concatenate([xy, xy + rng.uniform(10, 200, (N, 2)).astype(np.float32)], 1)
    gxy = rng.uniform(50, 300, (G, 2)).astype(np.float32)
    gtb = np.concatenate([gxy, gxy + rng.uniform(50, 200, (G, 2)).astype(np.float32)], 1)
    boxes[:300] = gtb[rng.integers(0, G, 300)] + rng.normal(0, 6, (300, 4)).astype(np.float32)
    pred = rng.integers(0, nc, N).astype(np.int32); gtl = np.array([2, 0, 1], np.int32)
    cm = np.zeros((nc, nc), np.int64)
    npos = oracle.lib().bto_cm_match(boxes, pred, N, gtb, gtl, G, np.float32(0.5), nc, cm, None)
    iou = ref_batch_bbox_iou(torch.from_numpy(boxes), torch.from_numpy(gtb))
    v, a = iou.max(dim=1)
    pos = v > 0.5
    ref = np.zeros((nc, nc), np.int64)
    for p, g in zip(pred[pos.numpy()], gtl[a[pos].numpy()]):
        ref[g, p] += 1
    assert npos == int(pos.sum()) and npos > 50
    np.testing.assert_array_equal(cm, ref)
    k10 = ref_batch_bbox_iou(torch.tensor([[0., 0, 10, 10]]), torch.tensor([[5., 5, 15, 15], [0., 0, 10, 10]]))
    np.testing.assert_allclose(k10.numpy(), [[25 / 175, 1.0]], rtol=1e-6)


def test_l1_decode_within_ulps_of_torch():
    """running_main_v2.py:743-775 restated in torch; the oracle's exp is its own fma polynomial, so
    agreement is to a few ulp of the box coordinates, not bit-for-bit."""
    rng = np.random.default_rng(1)
    S, nc, R = 160, 3, 16
    maps = [rng.normal(0, 2, (4 * R + nc, S // s, S // s)).astype(np.float32) for s in (8, 16, 32)]
    boxes, scores = oracle.decode_l1(maps, S, nc, R)
    ref_b, ref_s = [], []
    proj = torch.arange(R, dtype=torch.float32)
    for m in maps:
        t = torch.from_numpy(m)[None]
        _, ch, h, w = t.shape
        stride = S / w
        flat = t.permute(0, 2, 3, 1).reshape(1, h * w, ch)
        ltrb = torch.einsum("ijkl,l->ijk", F.softmax(flat[..., :4 * R].view(1, h * w, 4, R), dim=-1), proj)
        gy, gx = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
        anc = torch.stack((gx + 0.5, gy + 0.5), dim=-1).view(1, h * w, 2)
        lt, rb = torch.split(ltrb * stride, 2, dim=-1)
        ref_b.append(torch.cat((anc * stride - lt, anc * stride + rb), dim=-1)[0]); ref_s.append(flat[0, :, 4 * R:].sigmoid())
    ref_b, ref_s = torch.cat(ref_b).numpy(), torch.cat(ref_s).numpy()
    np.testing.assert_allclose(boxes, ref_b, rtol=0, atol=1e-4)       # k11-style: (anchor -/+ ltrb) * stride
    np.testing.assert_allclose(scores, ref_s, rtol=3e-6, atol=0)
    assert sum((S // s) ** 2 for s in (8, 16, 32)) == 525 and sum((640 // s) ** 2 for s in (8, 16, 32)) == 8400   # k12
    assert sum((1024 // s) ** 2 for s in (8, 16, 32)) == 21504


# ---------------------------------------------------------------- masks (a7, a8)
def test_upsample_and_threshold_bit_exact_vs_torch():
    rng = np.random.default_rng(2)
    x = rng.normal(0, 1, (40, 40)).astype(np.float32)
    up = np.empty((160, 160), np.float32)
    oracle.lib().bto_upsample(x.ravel().copy(), 40, 40, up, 160, 160)
    ref = F.interpolate(torch.from_numpy(x)[None, None], size=(160, 160), mode="bilinear", align_corners=False)[0, 0]
    assert up.tobytes() == ref.numpy().tobytes()                                                     # k8
    np.testing.assert_array_equal(oracle.threshold(up), (ref.sigmoid() > 0.5).numpy().astype(np.uint8))  # k9
    edge = np.array([0.0, 5.9604645e-08, 8.940697e-08, np.nextafter(np.float32(8.940697e-08), np.float32(1)), 1e-7, -1e-9],
                    np.float32)
    np.testing.assert_array_equal(oracle.threshold(edge), (torch.from_numpy(edge).sigmoid() > 0.5).numpy().astype(np.uint8))


def test_instance_mask_follows_test_model_statement():
    """test_model.py:80-85: einsum(coeff, protos) -> bilinear -> sigmoid > 0.5 (no crop).  The matmul's
    summation order is the library's, so pixels may differ only where |logit| is at rounding level."""
    rng = np.random.default_rng(5)
    nm, ph, S = 32, 40, 160
    protos = rng.normal(0, 1, (nm, ph, ph)).astype(np.float32)
    coeff = rng.normal(0, 1, nm).astype(np.float32)
    m = oracle.instance_mask(protos, coeff, np.array([0, 0, S, S], np.float32), S, crop=0)
    logits = torch.einsum("bqc,bchw->bqhw", torch.from_numpy(coeff)[None, None], torch.from_numpy(protos)[None])
    up = F.interpolate(logits, size=(S, S), mode="bilinear", align_corners=False)
    ref = (up.sigmoid() > 0.5)[0, 0].numpy()
    diff = m.astype(bool) != ref
    assert diff.mean() < 1e-3
    assert np.all(np.abs(up[0, 0].numpy()[diff]) < 1e-5)
    # crop at prototype resolution (Ultralytics crop_mask: r >= x1 & r < x2 & c >= y1 & c < y2 on the scaled box)
    box = np.array([33.0, 50.0, 101.0, 120.0], np.float32)
    mc = oracle.instance_mask(protos, coeff, box, S, crop=1)
    r = torch.arange(ph, dtype=torch.float32)
    x1, y1, x2, y2 = (torch.tensor(box) * 0.25).tolist()
    keep = ((r[None, :] >= x1) & (r[None, :] < x2) & (r[:, None] >= y1) & (r[:, None] < y2)).float()
    upc = F.interpolate(logits * keep, size=(S, S), mode="bilinear", align_corners=False)
    refc = (upc.sigmoid() > 0.5)[0, 0].numpy()
    d2 = mc.astype(bool) != refc
    assert d2.mean() < 1e-3 and np.all(np.abs(upc[0, 0].numpy()[d2]) < 1e-5)


def test_dice_iou_and_counters():
    rng = np.random.default_rng(6)
    p = (rng.uniform(size=(64, 64)) > 0.6).astype(np.uint8); g = (rng.uniform(size=(64, 64)) > 0.5).astype(np.uint8)
    cnt4 = np.zeros(4, np.int64)
    inter, ps, gs = oracle.mask_counts(p, g, cnt4)
    assert (inter, ps, gs) == (int((p & g).sum()), int(p.sum()), int(g.sum()))
    np.testing.assert_array_equal(cnt4, [inter, ps - inter, gs - inter, p.size - ps - gs + inter])
    d, i = oracle.dice_iou(inter, ps, gs)
    pt, gt = torch.from_numpy(p.astype(bool)), torch.from_numpy(g.astype(bool))
    it = (pt & gt).float().sum(); un = (pt | gt).float().sum()                  # test_model.py:15-23
    assert np.float32(i) == ((it + 1e-7) / (un + 1e-7)).numpy()
    assert np.float32(d) == ((2 * it + 1e-7) / (pt.float().sum() + gt.float().sum() + 1e-7)).numpy()


# ---------------------------------------------------------------- COCO matching (a9) -- parity unpinned
def py_coco_match(dets, gts, thrs, rng_lo, rng_hi):
    """Independent pure-Python restatement of COCOeval.evaluateImg for one class."""
    d = [(float(b[0]), float(b[1]), float(np.float32(b[2]) - np.float32(b[0])), float(np.float32(b[3]) - np.float32(b[1]))) for b in dets]
    g = [(float(b[0]), float(b[1]), float(np.float32(b[2]) - np.float32(b[0])), float(np.float32(b[3]) - np.float32(b[1]))) for b in gts]

    def iou(a, b):
        w = min(a[0] + a[2], b[0] + b[2]) - max(a[0], b[0])
        h = min(a[1] + a[3], b[1] + b[3]) - max(a[1], b[1])
        if w <= 0 or h <= 0:
            return 0.0
        return w * h / (a[2] * a[3] + b[2] * b[3] - w * h)
    gi = [(x[2] * x[3] < rng_lo or x[2] * x[3] > rng_hi) for x in g]
    order = sorted(range(len(g)), key=lambda k: gi[k])
    dm = np.zeros((len(thrs), len(d)), np.int32); di = np.zeros((len(thrs), len(d)), np.uint8)
    for ti, t in enumerate(thrs):
        gm = set()
        for k, dd in enumerate(d):
            best, m = min(t, 1 - 1e-10), -1
            for gg in order:
                if gg in gm:
                    continue
                if m > -1 and not gi[m] and gi[gg]:
                    break
                v = iou(dd, g[gg])
                if v < best:
                    continue
                best, m = v, gg
            if m >= 0:
                gm.add(m); dm[ti, k] = m + 1; di[ti, k] = gi[m]
            else:
                di[ti, k] = dd[2] * dd[3] < rng_lo or dd[2] * dd[3] > rng_hi
    return dm, di, np.array(gi, np.uint8)


def test_coco_match_against_independent_restatement():
    rng = np.random.default_rng(7)
    thrs = oracle.iou_thresholds()
    for trial in range(20):
        G, D = int(rng.integers(0, 6)), int(rng.integers(0, 40))
        gxy = rng.uniform(0, 400, (G, 2)).astype(np.float32)
        gts = np.concatenate([gxy, gxy + rng.uniform(8, 200, (G, 2)).astype(np.float32)], 1)
        dxy = rng.uniform(0, 400, (D, 2)).astype(np.float32)
        dets = np.concatenate([dxy, dxy + rng.uniform(8, 200, (D, 2)).astype(np.float32)], 1)
        if G and D:
            dets[: D // 2] = gts[rng.integers(0, G, D // 2)] + rng.normal(0, 8, (D // 2, 4)).astype(np.float32)
        dm, di, gi = oracle.coco_match(dets, gts, thrs)
        for a, (lo, hi) in enumerate(oracle.AREA_RANGES):
            rdm, rdi, rgi = py_coco_match(dets, gts, thrs, lo, hi)
            np.testing.assert_array_equal(dm[a], rdm); np.testing.assert_array_equal(di[a], rdi)
            np.testing.assert_array_equal(gi[a], rgi)
