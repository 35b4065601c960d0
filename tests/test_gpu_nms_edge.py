"""-m gpu: NMS edge cases ON THE CUDA KERNELS (through the C ABI: btpost_decode_filter + btpost_nms_match on hand-built
L2 heads), against `torchvision.ops.nms` itself (the reference's call, `/root/reference/src/running_main_v2.py:817`) and
against the oracle.

* known answers k1-k6 of SURVEY.md §8(c): score ties -> lower index, IoU exactly at the threshold is kept (strict >),
  identical zero-area boxes both kept (0/0), empty input, int64 descending-score output; NaN and -0.0 scores;
* the three regimes of the kernel's pair tests: iou 0.5 (no centre cull), 0.55 and 0.6 (centre cull on);
* a fuzz over 208 seeds with scores quantised to 1/64 (masses of ties) and coordinates to 0.5 px (IoUs that hit the
  threshold exactly), compared index for index with torchvision;
* `max_cand` = Ultralytics `max_nms` (top-k by score BEFORE the NMS), against a torch restatement of that statement.
"""
import numpy as np
import pytest
import torch
import torchvision

from btpost import PostConfig, PostProcessor

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
NC, NM = 3, 32


def head_from_boxes(boxes, scores, labels=None, n_pad=None):
    """[n,4] xyxy + [n] scores (+ labels) -> one image of an L2 head [4+nc+nm, N] (xywh rows, class-score rows)."""
    boxes = np.asarray(boxes, np.float32).reshape(-1, 4)
    n = len(boxes)
    N = n_pad or max(4, (n + 3) // 4 * 4)
    h = np.zeros((4 + NC + NM, N), np.float32)
    h[0, :n] = (boxes[:, 0] + boxes[:, 2]) * np.float32(0.5)
    h[1, :n] = (boxes[:, 1] + boxes[:, 3]) * np.float32(0.5)
    h[2, :n] = boxes[:, 2] - boxes[:, 0]
    h[3, :n] = boxes[:, 3] - boxes[:, 1]
    h[4:4 + NC] = -1.0                                   # padding anchors never pass the filter
    lab = np.zeros(n, np.int64) if labels is None else np.asarray(labels)
    for i in range(n):
        h[4 + lab[i], i] = scores[i]
    return h


def decoded(head_img, conf, S):
    """numpy restatement of decode + filter (fp32, one rounding per op) -> candidates in anchor order."""
    cx, cy, w, hh = head_img[:4]
    hw, hh2 = w * np.float32(0.5), hh * np.float32(0.5)
    b = np.stack([cx - hw, cy - hh2, cx + hw, cy + hh2], 1).astype(np.float32)
    sc = head_img[4:4 + NC]
    lab = sc.argmax(0)
    best = sc.max(0)
    keep = best > np.float32(conf)
    return np.clip(b[keep], 0, S).astype(np.float32), best[keep].astype(np.float32), lab[keep]


_PP = {}


def run_det(heads, conf=0.05, iou=0.6, max_det=300, S=64, class_mode=0, max_cand=0, nms_threads=0):
    heads = np.ascontiguousarray(heads, np.float32)
    B, _, N = heads.shape
    key = (B, N, conf, iou, max_det, S, class_mode, max_cand, nms_threads)
    if key not in _PP:
        cfg = PostConfig(batch=B, img_size=S, num_anchors=N, conf_thres=conf, iou_thres=iou, max_det=max_det, class_mode=class_mode,
                         max_cand=max_cand, with_coco=False, nms_threads=nms_threads)
        pp = PostProcessor(cfg, DEV)
        pp._dummy = (torch.zeros(B, NM, S // 4, S // 4, device=DEV), torch.zeros(B, 1, S, S, dtype=torch.uint8, device=DEV),
                     torch.zeros(NM, device=DEV))
        _PP[key] = pp
    pp = _PP[key]
    pr, mk, w = pp._dummy
    hd = torch.from_numpy(heads).to(DEV)
    for stage in ("decode_filter", "nms_match"):
        out = pp.run(hd, pr, None, mk, w, 0.0, stage=stage)
    torch.cuda.synchronize()
    return {k: out[k].cpu().numpy() for k in ("det_count", "det_keep", "dets", "n_cand", "det_anchor")}


def tv_keep(head_img, conf, iou, max_det, S):
    b, s, _ = decoded(head_img, conf, S)
    if len(s) == 0:
        return np.zeros(0, np.int64)
    return torchvision.ops.nms(torch.from_numpy(b), torch.from_numpy(s), iou)[:max_det].numpy()


def check_vs_torchvision(heads, conf, iou, max_det, S, nms_threads=0):
    got = run_det(heads, conf, iou, max_det, S, nms_threads=nms_threads)
    for b in range(heads.shape[0]):
        want = tv_keep(heads[b], conf, iou, max_det, S)
        k = int(got["det_count"][b])
        assert got["det_keep"].dtype == np.int64
        np.testing.assert_array_equal(got["det_keep"][b, :k], want, err_msg=f"image {b}, iou {iou}")
        assert (got["det_keep"][b, k:] == -1).all()
    return got


A, Bx = [0, 0, 10, 10], [20, 20, 30, 30]


@pytest.mark.parametrize("iou", [0.5, 0.55, 0.6])
def test_k1_ties_keep_the_lower_index(iou):
    h = head_from_boxes([A, Bx, A, Bx], [0.5, 0.5, 0.5, 0.5])[None]
    got = check_vs_torchvision(h, 0.05, iou, 300, 64)
    assert got["det_keep"][0, :2].tolist() == [0, 1] and got["det_count"][0] == 2


def test_k2_iou_exactly_at_the_threshold_is_kept():
    h = head_from_boxes([[0, 0, 2, 2], [0, 0, 2, 1]], [0.9, 0.8])[None]
    assert check_vs_torchvision(h, 0.05, 0.5, 300, 64)["det_count"][0] == 2          # IoU == 0.5, strict >
    assert check_vs_torchvision(h, 0.05, 0.4999, 300, 64)["det_keep"][0, :1].tolist() == [0]
    assert check_vs_torchvision(h, 0.05, 0.4999, 300, 64)["det_count"][0] == 1
    # the same where the centre cull is on.  0.55 and 0.6 are not binary fractions: the fp32 quotient 55/100 is
    # 0.5500000119 > 0.55 (the comparison is made in double, as torchvision does), so the pair IS suppressed at the
    # nominal threshold and kept once the threshold is the fp32 value itself
    for frac, hgt in ((0.55, 5.5), (0.6, 6.0)):
        hh = head_from_boxes([[0, 0, 10, 10], [0, 0, 10, hgt]], [0.9, 0.8])[None]
        assert check_vs_torchvision(hh, 0.05, frac, 300, 64)["det_count"][0] == 1
        assert check_vs_torchvision(hh, 0.05, float(np.float32(frac)), 300, 64)["det_count"][0] == 2
        assert check_vs_torchvision(hh, 0.05, float(np.nextafter(np.float32(frac), np.float32(0))), 300, 64)["det_count"][0] == 1


@pytest.mark.parametrize("iou", [0.5, 0.6])
def test_k3_identical_zero_area_boxes_are_both_kept(iou):
    h = head_from_boxes([[5, 5, 5, 5], [5, 5, 5, 5], [7, 3, 7, 9], [7, 3, 7, 9]], [0.9, 0.8, 0.7, 0.6])[None]
    assert check_vs_torchvision(h, 0.05, iou, 300, 64)["det_count"][0] == 4           # 0/0 = NaN never suppresses


def test_k4_nan_and_negative_zero_scores():
    # a NaN score cannot pass the strict filter (NaN > CONF_TH is false in the reference too): the anchor is dropped,
    # and a NaN in ANY class row poisons the max exactly as torch's .max(dim) does
    h = head_from_boxes([A, Bx, [40, 40, 50, 50]], [0.1, 0.2, 0.3])
    h[4, 1] = np.nan
    h[5, 2] = np.nan                                          # class 0 holds 0.3, class 1 NaN -> max is NaN
    got = run_det(h[None], 0.05, 0.6, 300, 64)
    assert got["n_cand"][0] == 1 and got["det_count"][0] == 1 and got["det_keep"][0, 0] == 0
    sc = torch.from_numpy(h[4:4 + NC].T.copy())
    assert int((sc.max(dim=1).values > 0.05).sum()) == 1      # what the reference's filter keeps (running_main_v2.py:788-790)
    # -0.0 == +0.0: a tie, lower index first
    hz = head_from_boxes([A, Bx, [40, 40, 50, 50]], [0.0, -0.0, 0.0])
    hz[4:4 + NC][hz[4:4 + NC] == -1.0] = -5.0
    got = check_vs_torchvision(hz[None], -1.0, 0.6, 300, 64)
    assert got["det_keep"][0, :3].tolist() == [0, 1, 2]


def test_k5_k6_empty_input_and_output_format():
    h = head_from_boxes([A, Bx], [0.01, 0.02])[None]
    got = run_det(h, 0.05, 0.6, 300, 64)
    assert got["det_count"][0] == 0 and got["n_cand"][0] == 0 and (got["det_keep"][0] == -1).all() and not got["dets"].any()
    h = head_from_boxes([A, Bx, [40, 40, 50, 50]], [0.3, 0.9, 0.6])[None]
    got = check_vs_torchvision(h, 0.05, 0.6, 300, 64)
    assert got["det_keep"].dtype == np.int64 and got["det_keep"][0, :3].tolist() == [1, 2, 0]
    assert (np.diff(got["dets"][0, :3, 4]) <= 0).all()


def fuzz_heads(seed, n, S):
    """Quantised clusters: boxes on a 0.5 px grid around a few centres, scores on a 1/64 grid."""
    rng = np.random.default_rng(seed)
    n_obj = int(rng.integers(1, 6))
    ctr = rng.uniform(0.15, 0.85, (n_obj, 2)) * S
    size = rng.uniform(0.05, 0.4, (n_obj, 2)) * S
    which = rng.integers(0, n_obj + 1, n)                      # n_obj = background
    cx = np.where(which < n_obj, ctr[np.minimum(which, n_obj - 1), 0], rng.uniform(0, S, n))
    cy = np.where(which < n_obj, ctr[np.minimum(which, n_obj - 1), 1], rng.uniform(0, S, n))
    w = np.where(which < n_obj, size[np.minimum(which, n_obj - 1), 0], rng.uniform(2, 48, n))
    hh = np.where(which < n_obj, size[np.minimum(which, n_obj - 1), 1], rng.uniform(2, 48, n))
    jit = rng.normal(0, 0.04, (4, n)) * np.stack([w, hh, w, hh])
    q = lambda v: np.round(v * 2) / 2
    x1, y1 = q(cx - w / 2 + jit[0]), q(cy - hh / 2 + jit[1])
    x2, y2 = q(cx + w / 2 + jit[2]), q(cy + hh / 2 + jit[3])
    if seed % 5 == 0:                                          # exact duplicates and degenerate boxes
        d = rng.integers(0, n, n // 8)
        x1[d], y1[d], x2[d], y2[d] = x1[d - 1], y1[d - 1], x2[d - 1], y2[d - 1]
        z = rng.integers(0, n, 4)
        x2[z] = x1[z]
    boxes = np.stack([x1, y1, x2, y2], 1)
    scores = np.round(rng.uniform(0, 1, n) ** 2 * 64) / 64
    labels = rng.integers(0, NC, n)
    return head_from_boxes(boxes, scores, labels, n_pad=n)


@pytest.mark.parametrize("group", range(26))
def test_fuzz_against_torchvision(group):
    """26 groups x 8 seeds = 208 images; the IoU threshold cycles through both sides of the centre-cull switch."""
    iou = [0.5, 0.55, 0.6, 0.45, 0.7, 0.6][group % 6]
    n = [256, 512, 1024, 2048][group % 4]
    S = 640
    heads = np.stack([fuzz_heads(1000 + 8 * group + i, n, S) for i in range(8)])
    max_det = 300 if group % 3 else 100
    got = check_vs_torchvision(heads, 0.05, iou, max_det, S)
    # packaged rows = gathered candidates (running_main_v2.py:818-839)
    for b in range(8):
        bx, sc, lab = decoded(heads[b], 0.05, S)
        k = int(got["det_count"][b])
        kp = got["det_keep"][b, :k]
        assert got["dets"][b, :k, :4].tobytes() == bx[kp].tobytes()
        assert got["dets"][b, :k, 4].tobytes() == sc[kp].tobytes()
        np.testing.assert_array_equal(got["dets"][b, :k, 5], lab[kp].astype(np.float32))


@pytest.mark.parametrize("class_mode", [1, 2])
def test_fuzz_class_aware_against_oracle_and_torchvision(class_mode):
    """class-aware = torchvision nms inside each class on the raw coordinates, merged by descending score (mode 1);
    Ultralytics' coordinate offset cls * 7680 (mode 2): both against torchvision on the transformed problem."""
    S, n = 640, 1024
    heads = np.stack([fuzz_heads(5000 + i, n, S) for i in range(8)])
    got = run_det(heads, 0.05, 0.6, 300, S, class_mode=class_mode)
    for b in range(8):
        bx, sc, lab = decoded(heads[b], 0.05, S)
        tb = torch.from_numpy(bx)
        if class_mode == 2:
            want = torchvision.ops.nms(tb + torch.from_numpy(lab.astype(np.float32))[:, None] * 7680.0, torch.from_numpy(sc), 0.6)
        else:
            keep = torch.zeros(len(sc), dtype=torch.bool)
            for c in range(NC):
                idx = torch.nonzero(torch.from_numpy(lab == c))[:, 0]
                keep[idx[torchvision.ops.nms(tb[idx], torch.from_numpy(sc)[idx], 0.6)]] = True
            order = torch.sort(torch.from_numpy(sc), descending=True, stable=True).indices
            want = order[keep[order]]
        want = want[:300].numpy()
        k = int(got["det_count"][b])
        np.testing.assert_array_equal(got["det_keep"][b, :k], want)


@pytest.mark.parametrize("max_cand", [64, 300, 1000])
def test_max_cand_is_top_k_by_score_like_ultralytics(max_cand):
    """Ultralytics non_max_suppression: `x = x[x[:, 4].argsort(descending=True)[:max_nms]]` then torchvision nms
    (SURVEY.md A.5).  Scores are made distinct so that the (unstable) argsort of the original is well defined."""
    S, n = 640, 2048
    heads = np.stack([fuzz_heads(7000 + i, n, S) for i in range(4)])
    rng = np.random.default_rng(7)
    for b in range(4):                                        # distinct scores: a permutation of a 1/4096 grid
        lab = heads[b, 4:4 + NC].argmax(0)
        vals = (rng.permutation(n) + 1).astype(np.float32) / np.float32(4096)
        heads[b, 4:4 + NC] = -1.0
        heads[b, 4 + lab, np.arange(n)] = vals
    got = run_det(heads, 0.05, 0.6, 300, S, max_cand=max_cand)
    for b in range(4):
        bx, sc, _ = decoded(heads[b], 0.05, S)
        assert got["n_cand"][b] == len(sc)                    # every anchor that passed the filter is reported
        top = torch.from_numpy(sc).argsort(descending=True)[:max_cand]
        keep = torchvision.ops.nms(torch.from_numpy(bx)[top], torch.from_numpy(sc)[top], 0.6)[:300]
        want = top[keep].numpy()                              # indices into the filtered (anchor-ordered) list
        k = int(got["det_count"][b])
        np.testing.assert_array_equal(got["det_keep"][b, :k], want)


@pytest.mark.parametrize("nms_threads", [0, 512, 256])
@pytest.mark.parametrize("n", [700, 1024, 3000, 4096, 6000, 8400])
def test_sort_paths_against_torchvision(n, nms_threads):
    """The order the sweep consumes the candidates in, on every sort path of the kernel (bucket rank sort; radix /
    register / global bitonic fallbacks) and in each thread variant.  Image 0: every score equal (one bucket holds the whole
    list: the bucket sort hands over to the fallback, ties -> lower index decides everything); image 1: two score values;
    image 2: continuous scores in a narrow band (the bucket map stretches min..max); image 3: quantised clusters (buckets
    of dozens of ties); image 4: continuous scores with a block of 300 duplicates."""
    S = 640
    rng = np.random.default_rng(77 + n)
    heads = []
    for kind in range(5):
        h = fuzz_heads(5000 + n + kind, n, S)
        sc = h[4:4 + NC]
        live = sc.max(0) > 0.05
        lab = sc.argmax(0)
        if kind == 0:
            new = np.full(n, 0.5, np.float32)
        elif kind == 1:
            new = np.where(rng.uniform(size=n) < 0.5, np.float32(0.25), np.float32(0.75)).astype(np.float32)
        elif kind == 2:
            new = (0.3 + 1e-4 * rng.uniform(size=n)).astype(np.float32)
        elif kind == 3:
            new = sc.max(0)
        else:
            new = rng.uniform(0.06, 1.0, n).astype(np.float32)
            d = rng.integers(0, n, 300)
            new[d] = new[d[0]]
        if kind != 3:
            sc[:] = -1.0
            sc[lab, np.arange(n)] = new
            if kind in (0, 1):
                sc[:, ~live] = -1.0          # keep some anchors below the threshold so list positions != anchor indices
        heads.append(h)
    check_vs_torchvision(np.stack(heads), 0.05, 0.6, 300, S, nms_threads=nms_threads)
