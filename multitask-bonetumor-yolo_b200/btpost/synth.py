"""Counter-based synthetic BTXRD-style head outputs (SURVEY.md §8d "value distributions & seeds").

Every element is a pure function of ``(seed, tensor_id, image_idx, element_idx)`` through a
splitmix64 finaliser, and every floating-point step is a single correctly-rounded fp32 operation
(no transcendental functions), so the numpy generator here and the CUDA generator in
``csrc/synth.cu`` (``make_batch_device`` below, `include/btpost_synth.h`) materialise bit-identical
tensors on any shard without moving data (`tests/test_gpu_synth.py`).

Layouts produced (reference producers cited):
  * L2 ``head``  [B, 4+nc+nm, N]  = ``segment_preds_cat`` (`/root/reference/src/main_modelv2.py:367-375`):
    rows 0..3 xywh in pixels, rows 4..4+nc class scores (post-sigmoid), then nm mask coefficients.
  * ``protos``   [B, nm, S/4, S/4] = ``segment_protos`` (`main_modelv2.py:373`).
  * ``det_boxes_gt`` [G_total, 6]  = ``(batch_idx, cls, cx, cy, w, h)`` normalised, the
    ``collate_fn`` layout (`/root/reference/src/dataset_btxrdv2.py:261-284`).
  * ``masks_gt`` [B, 1, S, S] u8 {0,1} (reference keeps float32 0/1, `dataset_btxrdv2.py:164-166`).
  * L1 ``maps``  3 x [B, 4*reg_max+nc, H_l, W_l] raw logits (`running_main_v2.py:743-752`).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

MASK64 = (1 << 64) - 1
GOLD = 0x9E3779B97F4A7C15
K_TENSOR = 0xD1B54A32D192ED03
K_IMAGE = 0x8CB92BA72F3D8DD7

TID_OBJECTS, TID_HEAD, TID_PROTO, TID_L1 = 1, 2, 3, 4
HEAD_STREAMS_FIXED = 16  # streams 0..15 are per-anchor scalars; then nc score streams; then nm coeffs
MAX_OBJ = 3


def mix64_int(z: int) -> int:
    z &= MASK64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return z ^ (z >> 31)


def stream_key(seed: int, tensor_id: int, image: int) -> int:
    return mix64_int(seed * GOLD + tensor_id * K_TENSOR + image * K_IMAGE)


def mix64(z: np.ndarray) -> np.ndarray:
    z = z.astype(np.uint64, copy=True)
    z ^= z >> np.uint64(30)
    z *= np.uint64(0xBF58476D1CE4E5B9)
    z ^= z >> np.uint64(27)
    z *= np.uint64(0x94D049BB133111EB)
    z ^= z >> np.uint64(31)
    return z


def hash_elems(key: int, idx: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        return mix64(np.uint64(key) + idx.astype(np.uint64) * np.uint64(GOLD))


def u24(h: np.ndarray) -> np.ndarray:
    """Uniform [0,1) on a 24-bit grid (exact in fp32)."""
    return (h >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)


def gauss16(h: np.ndarray) -> np.ndarray:
    """Unit-variance, zero-mean Irwin-Hall(4) from the four 16-bit lanes of one hash."""
    m = np.uint64(0xFFFF)
    s = ((h & m) + ((h >> np.uint64(16)) & m) + ((h >> np.uint64(32)) & m) + (h >> np.uint64(48))).astype(np.int64)
    return (s - 131070).astype(np.float32) * np.float32(np.sqrt(3.0) / 65536.0)


@dataclass
class SynthConfig:
    batch: int = 16
    img_size: int = 640
    nc: int = 3
    nm: int = 32
    reg_max: int = 16
    seed: int = 20261
    image_offset: int = 0  # global index of image 0 of this batch (sharding)

    @property
    def levels(self):
        return [(self.img_size // s, self.img_size // s, float(s)) for s in (8, 16, 32)]

    @property
    def num_anchors(self):
        return sum(h * w for h, w, _ in self.levels)

    @property
    def proto_hw(self):
        return self.img_size // 4


def anchor_grid(cfg: SynthConfig):
    """Anchor centres in pixels and strides, level order P3,P4,P5, row-major (y*W+x)."""
    ax, ay, st = [], [], []
    for h, w, s in cfg.levels:
        yy, xx = np.meshgrid(np.arange(h, dtype=np.float32), np.arange(w, dtype=np.float32), indexing="ij")
        ax.append(((xx + np.float32(0.5)) * np.float32(s)).ravel())
        ay.append(((yy + np.float32(0.5)) * np.float32(s)).ravel())
        st.append(np.full(h * w, s, np.float32))
    return np.concatenate(ax), np.concatenate(ay), np.concatenate(st)


def object_table(cfg: SynthConfig) -> np.ndarray:
    """[B, MAX_OBJ, 6] fp32 rows (valid, cls, cx, cy, w, h) in pixels; G ~ U{1,2,3} objects."""
    S = np.float32(cfg.img_size)
    tab = np.zeros((cfg.batch, MAX_OBJ, 6), np.float32)
    for b in range(cfg.batch):
        key = stream_key(cfg.seed, TID_OBJECTS, cfg.image_offset + b)
        h = hash_elems(key, np.arange(1 + MAX_OBJ * 5))
        g = 1 + int(h[0] % np.uint64(3))
        u = u24(h[1:]).reshape(MAX_OBJ, 5)
        for o in range(g):
            tab[b, o, 0] = 1.0
            tab[b, o, 1] = np.float32(int(h[1 + o * 5] & np.uint64(0xFFFF)) % cfg.nc)
            tab[b, o, 2] = (np.float32(0.2) + np.float32(0.6) * u[o, 1]) * S
            tab[b, o, 3] = (np.float32(0.2) + np.float32(0.6) * u[o, 2]) * S
            tab[b, o, 4] = (np.float32(0.05) + np.float32(0.30) * u[o, 3]) * S
            tab[b, o, 5] = (np.float32(0.05) + np.float32(0.30) * u[o, 4]) * S
    return tab


def gt_from_objects(cfg: SynthConfig, tab: np.ndarray):
    """det_boxes_gt [G,6] normalised rows + masks_gt [B,1,S,S] u8 (union of inscribed ellipses)."""
    S = cfg.img_size
    rows = []
    masks = np.zeros((cfg.batch, 1, S, S), np.uint8)
    pc = np.arange(S, dtype=np.float32) + np.float32(0.5)
    for b in range(cfg.batch):
        for o in range(MAX_OBJ):
            valid, cls, cx, cy, w, h = tab[b, o]
            if valid == 0:
                continue
            inv = np.float32(1.0) / np.float32(S)
            rows.append([np.float32(b), cls, cx * inv, cy * inv, w * inv, h * inv])
            dx = (pc - cx) / (w * np.float32(0.5))
            dy = (pc - cy) / (h * np.float32(0.5))
            inside = (dy * dy)[:, None] + (dx * dx)[None, :] <= np.float32(1.0)
            masks[b, 0] |= inside.astype(np.uint8)
    gt = np.asarray(rows, np.float32).reshape(-1, 6)
    return gt, masks


def _pow96(u: np.ndarray) -> np.ndarray:
    u2 = u * u; u4 = u2 * u2; u8 = u4 * u4; u16 = u8 * u8; u32 = u16 * u16; u64 = u32 * u32
    return u64 * u32


def _anchor_fields(cfg: SynthConfig, b: int, tab: np.ndarray, tensor_id: int):
    """Shared by the L2 and L1 generators: per-anchor box (xywh, px) and class scores."""
    N, nc = cfg.num_anchors, cfg.nc
    K = HEAD_STREAMS_FIXED + cfg.nc + cfg.nm
    ax, ay, _ = anchor_grid(cfg)
    key = stream_key(cfg.seed, tensor_id, cfg.image_offset + b)
    n_idx = np.arange(N, dtype=np.uint64) * np.uint64(K)

    def U(s):
        return u24(hash_elems(key, n_idx + np.uint64(s)))

    f32 = np.float32
    # background default: anchor-centred box, wh U[4,64) px, low scores 0.002 + 0.6 u^96
    cx, cy = ax.copy(), ay.copy()
    w = f32(4.0) + f32(60.0) * U(1)
    h = f32(4.0) + f32(60.0) * U(2)
    scores = np.stack([f32(0.002) + f32(0.6) * _pow96(U(HEAD_STREAMS_FIXED + c)) for c in range(nc)], 0)
    assigned = np.zeros(N, bool)
    coin = U(0) < f32(0.5)
    j = [U(3) + U(4) - f32(1.0), U(5) + U(6) - f32(1.0), U(7) + U(8) - f32(1.0), U(9) + U(10) - f32(1.0)]
    s_obj = f32(0.3) + f32(0.65) * U(11)
    s_oth = [f32(0.02) * U(HEAD_STREAMS_FIXED + c) for c in range(nc)]
    for o in range(MAX_OBJ):
        valid, cls, ocx, ocy, ow, oh = tab[b, o]
        if valid == 0:
            continue
        inside = (np.abs(ax - ocx) < ow * f32(0.5)) & (np.abs(ay - ocy) < oh * f32(0.5)) & coin & ~assigned
        cx = np.where(inside, ocx + j[0] * (f32(0.06) * ow), cx)
        cy = np.where(inside, ocy + j[1] * (f32(0.06) * oh), cy)
        w = np.where(inside, ow + j[2] * (f32(0.06) * ow), w)
        h = np.where(inside, oh + j[3] * (f32(0.06) * oh), h)
        for c in range(nc):
            scores[c] = np.where(inside, s_obj if c == int(cls) else s_oth[c], scores[c])
        assigned |= inside
    _anchor_fields.last_assigned = assigned   # (used by the peaked L1 variant; plain attribute to keep the 4-tuple API)
    return key, n_idx, (cx, cy, w, h), scores


def make_head_l2(cfg: SynthConfig, tab: np.ndarray | None = None) -> np.ndarray:
    tab = object_table(cfg) if tab is None else tab
    N, nc, nm = cfg.num_anchors, cfg.nc, cfg.nm
    head = np.empty((cfg.batch, 4 + nc + nm, N), np.float32)
    for b in range(cfg.batch):
        key, n_idx, box, scores = _anchor_fields(cfg, b, tab, TID_HEAD)
        for r in range(4):
            head[b, r] = box[r]
        head[b, 4:4 + nc] = scores
        for m in range(nm):
            head[b, 4 + nc + m] = gauss16(hash_elems(key, n_idx + np.uint64(HEAD_STREAMS_FIXED + nc + m)))
    return head


def make_protos(cfg: SynthConfig) -> np.ndarray:
    P = cfg.proto_hw
    out = np.empty((cfg.batch, cfg.nm, P, P), np.float32)
    idx = np.arange(cfg.nm * P * P, dtype=np.uint64)
    for b in range(cfg.batch):
        key = stream_key(cfg.seed, TID_PROTO, cfg.image_offset + b)
        out[b] = gauss16(hash_elems(key, idx)).reshape(cfg.nm, P, P)
    return out


def make_maps_l1(cfg: SynthConfig, tab: np.ndarray | None = None, peaked: bool = False):
    """Raw per-level maps: box-bin logits 2*gauss over 4*reg_max channels, class logits chosen so
    that sigmoid(logit) follows the same recipe as the L2 scores (logit = log(s/(1-s)) in fp64,
    rounded once to fp32; host-only, the CUDA generator does not produce L1).

    ``peaked=True``: anchors that belong to an object get near-one-hot DFL logits (+12 on the bin nearest to the
    distance between the anchor and the edge of their jittered object box, -12 elsewhere), so the DECODED boxes sit
    on the objects: clusters for the NMS and many anchor<->GT confusion-matrix pairs, which random DFL logits
    almost never produce."""
    tab = object_table(cfg) if tab is None else tab
    nc, R = cfg.nc, cfg.reg_max
    maps, off = [], 0
    per_image, per_box, per_asg = [], [], []
    for b in range(cfg.batch):
        _, _, box, scores = _anchor_fields(cfg, b, tab, TID_HEAD)
        per_image.append(scores); per_box.append(box); per_asg.append(_anchor_fields.last_assigned.copy())
    ax, ay, _ = anchor_grid(cfg)
    for (h, w, s) in cfg.levels:
        m = np.empty((cfg.batch, 4 * R + nc, h, w), np.float32)
        idx = np.arange(4 * R * h * w, dtype=np.uint64)
        for b in range(cfg.batch):
            key = stream_key(cfg.seed, TID_L1 + int(s), cfg.image_offset + b)
            m[b, :4 * R] = (np.float32(2.0) * gauss16(hash_elems(key, idx))).reshape(4 * R, h, w)
            if peaked:
                cx, cy, bw, bh = (v[off:off + h * w] for v in per_box[b])
                asg = per_asg[b][off:off + h * w]
                axl, ayl = ax[off:off + h * w], ay[off:off + h * w]
                dist = np.stack([axl - (cx - bw * np.float32(0.5)), ayl - (cy - bh * np.float32(0.5)),
                                 (cx + bw * np.float32(0.5)) - axl, (cy + bh * np.float32(0.5)) - ayl]) / np.float32(s)
                k = np.clip(np.rint(dist), 0, R - 1).astype(np.int64)                      # [4, hw]
                dfl = m[b, :4 * R].reshape(4, R, h * w)
                onehot = np.full((4, R, h * w), np.float32(-12.0))
                np.put_along_axis(onehot, k[:, None, :], np.float32(12.0), axis=1)
                dfl[:, :, asg] = onehot[:, :, asg]
            sc = per_image[b][:, off:off + h * w].astype(np.float64)
            m[b, 4 * R:] = np.log(sc / (1.0 - sc)).astype(np.float32).reshape(nc, h, w)
        maps.append(m)
        off += h * w
    return maps


def make_batch(cfg: SynthConfig, l1: bool = False, l1_peaked: bool = False) -> dict:
    tab = object_table(cfg)
    gt, masks = gt_from_objects(cfg, tab)
    out = {"objects": tab, "head": make_head_l2(cfg, tab), "protos": make_protos(cfg),
           "det_boxes_gt": gt, "masks_gt": masks}
    if l1:
        out["maps"] = make_maps_l1(cfg, tab, peaked=l1_peaked)
        out["coeffs"] = np.ascontiguousarray(out["head"][:, 4 + cfg.nc:, :])   # Segment `mc` [B, nm, N]
    key = stream_key(cfg.seed, 99, 0)
    h = hash_elems(key, np.arange(cfg.nm + 1))
    out["proj_weight"] = gauss16(h[:cfg.nm]) * np.float32(0.25)
    out["proj_bias"] = gauss16(h[cfg.nm:])[0] * np.float32(0.1)
    return out


def prepare_batch_device(cfg: SynthConfig, device="cuda:0") -> dict:
    """Host part of the device generator: stream keys, object table (3 rows per image) and GT rows, uploaded once.
    `generate_into` then needs no host work, so a sweep can prepare all its batches up front."""
    import torch
    dev = torch.device(device)
    tab = object_table(cfg)
    inv = np.float32(1.0) / np.float32(cfg.img_size)
    rows = [[np.float32(b), tab[b, o, 1], tab[b, o, 2] * inv, tab[b, o, 3] * inv, tab[b, o, 4] * inv, tab[b, o, 5] * inv]
            for b in range(cfg.batch) for o in range(MAX_OBJ) if tab[b, o, 0] != 0]
    gt = np.asarray(rows, np.float32).reshape(-1, 6)
    keys = np.array([[stream_key(cfg.seed, TID_HEAD, cfg.image_offset + b), stream_key(cfg.seed, TID_PROTO, cfg.image_offset + b)]
                     for b in range(cfg.batch)], dtype=np.uint64)
    # uint64 has no torch dtype everywhere: ship the keys as int64 bit patterns
    return {"cfg": cfg, "objects": tab,
            "kh": torch.from_numpy(np.ascontiguousarray(keys[:, 0]).view(np.int64)).to(dev),
            "kp": torch.from_numpy(np.ascontiguousarray(keys[:, 1]).view(np.int64)).to(dev),
            "objs": torch.from_numpy(tab).to(dev), "det_boxes_gt": torch.from_numpy(gt).to(dev)}


def generate_into(prep: dict, head, protos, masks, stream=None):
    """Device part: fills caller-owned `head` [B,4+nc+nm,N] f32, `protos` [B,nm,S/4,S/4] f32, `masks` [B,1,S,S] u8."""
    import ctypes as C

    import torch

    from . import _lib
    cfg = prep["cfg"]
    dev = head.device
    assert head.dtype == torch.float32 and protos.dtype == torch.float32 and masks.dtype == torch.uint8
    assert tuple(head.shape) == (cfg.batch, 4 + cfg.nc + cfg.nm, cfg.num_anchors) and head.is_contiguous()
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        rc = _lib.load().btpost_synth_batch(cfg.batch, cfg.img_size, cfg.nc, cfg.nm, C.c_void_p(prep["kh"].data_ptr()),
                                           C.c_void_p(prep["kp"].data_ptr()), C.c_void_p(prep["objs"].data_ptr()),
                                           C.c_void_p(head.data_ptr()), C.c_void_p(protos.data_ptr()), C.c_void_p(masks.data_ptr()),
                                           C.c_void_p(st.cuda_stream))
    _lib.check(rc, "btpost_synth_batch")


def make_batch_device(cfg: SynthConfig, device="cuda:0", stream=None, out: dict | None = None) -> dict:
    """Same batch as ``make_batch`` (L2 layout), generated on the GPU by ``btpost_synth_batch``: only the stream
    keys, the object table (3 rows per image) and the GT rows are made on the host.  Returns torch tensors on
    ``device`` (``head``, ``protos``, ``masks_gt``, ``det_boxes_gt``, ``proj_weight``) and ``proj_bias``; with ``out``
    the three big tensors are generated straight into caller-owned buffers."""
    import torch
    dev = torch.device(device)
    prep = prepare_batch_device(cfg, dev)
    P = cfg.proto_hw
    if out is not None:
        head, protos, masks = out["head"], out["protos"], out["masks_gt"]
    else:
        head = torch.empty(cfg.batch, 4 + cfg.nc + cfg.nm, cfg.num_anchors, dtype=torch.float32, device=dev)
        protos = torch.empty(cfg.batch, cfg.nm, P, P, dtype=torch.float32, device=dev)
        masks = torch.empty(cfg.batch, 1, cfg.img_size, cfg.img_size, dtype=torch.uint8, device=dev)
    generate_into(prep, head, protos, masks, stream)
    key = stream_key(cfg.seed, 99, 0)
    h = hash_elems(key, np.arange(cfg.nm + 1))
    return {"objects": prep["objects"], "head": head, "protos": protos, "masks_gt": masks, "det_boxes_gt": prep["det_boxes_gt"],
            "proj_weight": torch.from_numpy(gauss16(h[:cfg.nm]) * np.float32(0.25)).to(dev),
            "proj_bias": float(gauss16(h[cfg.nm:])[0] * np.float32(0.1)), "_keepalive": prep}
