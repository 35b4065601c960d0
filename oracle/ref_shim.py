"""TEST INFRASTRUCTURE ONLY — loader that runs the *unmodified* reference code on CPU.

The reference (`/root/reference/src/running_main_v{2,3}.py`) imports packages that are not
installed in this image (pytorch_lightning, torchmetrics, matplotlib, seaborn, timm/ultralytics
via main_model, wandb via multitask_logging).  This module installs recorder stubs for exactly
those names in ``sys.modules`` and then imports the reference module *verbatim* from
``/root/reference/src``.  ``MultiTaskLitModel.validation_step`` then runs as shipped
(`running_main_v2.py:643-945`, `running_main_v3.py:447-599`) on synthetic L1 head outputs by
overriding ``forward``; whatever the reference would hand to ``val_map_iou50.update`` /
``val_seg_*.update`` is captured from the recorders.

It is only usable where ``/root/reference`` exists (the build container).  It is used by
``tests/golden/make_golden.py`` to produce the committed fixtures; nothing in the product,
in ``-m gpu`` tests, in ``smoke()`` or in ``bench.py`` imports it.
"""
from __future__ import annotations

import importlib
import sys
import types
from pathlib import Path

import torch
import torch.nn as nn

REFERENCE_SRC = Path("/root/reference/src")


class _Recorder:
    """Stands in for a torchmetrics metric object: stores every ``update`` call."""

    def __init__(self, *args, **kwargs):
        self.init_args, self.init_kwargs = args, kwargs
        self.calls = []

    def update(self, *args, **kwargs):
        self.calls.append((args, kwargs))

    def compute(self):
        return torch.tensor(0.0)

    def reset(self):
        self.calls = []

    def to(self, *a, **k):
        return self

    def __call__(self, *args, **kwargs):
        self.update(*args, **kwargs)
        return torch.tensor(0.0)


class _HParams(dict):
    __getattr__ = dict.__getitem__


class _LightningModule(nn.Module):
    """Minimal LightningModule: hparams, log sinks, epoch counters."""

    def __init__(self):
        super().__init__()
        self.hparams = _HParams()
        self.logged = []
        self._current_epoch = 0
        self.logger = None
        self.global_step = 0

    @property
    def current_epoch(self):
        return self._current_epoch

    def save_hyperparameters(self, *names):
        frame = sys._getframe(1)
        for n in names:
            self.hparams[n] = frame.f_locals[n]

    def log(self, *a, **k):
        self.logged.append((a, k))

    def log_dict(self, *a, **k):
        self.logged.append((a, k))


class _DummyDetect:
    reg_max = 16


class _DummyNet(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        self.detect = _DummyDetect()


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_stubs():
    if "pytorch_lightning" in sys.modules and getattr(sys.modules["pytorch_lightning"], "_bt_stub", False):
        return
    pl = _mod("pytorch_lightning", LightningModule=_LightningModule, LightningDataModule=object,
              Trainer=object, seed_everything=lambda *a, **k: None, _bt_stub=True)
    _mod("pytorch_lightning.loggers", WandbLogger=object)
    _mod("pytorch_lightning.callbacks", ModelCheckpoint=object, EarlyStopping=object,
         LearningRateMonitor=object)
    pl.loggers = sys.modules["pytorch_lightning.loggers"]
    pl.callbacks = sys.modules["pytorch_lightning.callbacks"]
    names = ["BinaryPrecision", "BinaryRecall", "BinaryAccuracy", "MulticlassAccuracy",
             "MulticlassConfusionMatrix"]
    _mod("torchmetrics", F1Score=_Recorder)
    _mod("torchmetrics.classification", **{n: _Recorder for n in names})
    _mod("torchmetrics.segmentation", DiceScore=_Recorder)
    _mod("torchmetrics.detection", MeanAveragePrecision=_Recorder)
    _mod("torchmetrics.functional")
    mpl = _mod("matplotlib", use=lambda *a, **k: None)
    mpl.pyplot = _mod("matplotlib.pyplot")
    _mod("seaborn")
    if "wandb" not in sys.modules:
        try:
            importlib.import_module("wandb")
        except Exception:
            _mod("wandb", Image=object)
    _mod("main_model", ConvNeXtBiFPNYOLO=_DummyNet, load_pretrained_heads=lambda *a, **k: None)
    _mod("multitask_logging", log_cls_metrics=lambda *a, **k: None,
         log_seg_examples=lambda *a, **k: None, log_det_examples=lambda *a, **k: None)
    _mod("dataset_btxrdv2", BTXRD=object, collate_fn=None)


def load_reference(version: str = "v3"):
    """Import ``running_main_<version>`` from /root/reference/src, unmodified."""
    if not REFERENCE_SRC.exists():
        raise RuntimeError("/root/reference is not present: the verbatim reference can only run "
                           "in the build container (fixtures under tests/golden/ are committed).")
    _install_stubs()
    if str(REFERENCE_SRC) not in sys.path:
        sys.path.insert(0, str(REFERENCE_SRC))
    mod = importlib.import_module(f"running_main_{version}")
    if version == "v2" and not hasattr(mod, "MAP_FULL_FREQ"):
        mod.MAP_FULL_FREQ = 1  # `__main__`-only global read inside validation_step (v2 :889,1264)
    return mod


def run_validation_step(version, det_maps, protos, det_boxes_gt, masks_gt, *, img_size,
                        nc_det=3, proj_weight=None, proj_bias=None, conf_th=None, nms_iou=None,
                        top_k=None, iou_match_thresh=0.5):
    """Run the reference's own ``validation_step`` on given L1 maps; return what it produced.

    det_maps: list of 3 fp32 [B, 64+nc, H_l, W_l]; protos [B,32,S/4,S/4];
    det_boxes_gt [G,6] (batch_idx, cls, cx, cy, w, h); masks_gt [B,1,S,S] float 0/1.
    """
    mod = load_reference(version)
    saved = (mod.CONF_TH, mod.NMS_IOU, mod.TOP_K)
    if conf_th is not None:
        mod.CONF_TH = conf_th
    if nms_iou is not None:
        mod.NMS_IOU = nms_iou
    if top_k is not None:
        mod.TOP_K = top_k
    try:
        torch.manual_seed(0)
        model = mod.MultiTaskLitModel(img_size=img_size, nc_det=nc_det, proto_ch=protos.shape[1],
                                      iou_match_thresh=iou_match_thresh)
        if proj_weight is not None:
            with torch.no_grad():
                model.seg_proto_projector.weight.copy_(proj_weight.view(1, -1, 1, 1))
                model.seg_proto_projector.bias.copy_(proj_bias.view(1))
        model.eval()
        B = protos.shape[0]
        img_logits = torch.zeros(B, 2)
        mc = torch.zeros(B, protos.shape[1], sum(m.shape[2] * m.shape[3] for m in det_maps))
        model.forward = lambda x, mode="train": (list(det_maps), (list(det_maps), mc, protos), img_logits)
        imgs = torch.zeros(B, 3, img_size, img_size)
        batch = (list(range(B)), imgs, det_boxes_gt, masks_gt, torch.zeros(B, dtype=torch.long))
        with torch.no_grad():
            model.validation_step(batch, 1)
        out = {
            "map_update": model.val_map_iou50.calls[-1][0] if model.val_map_iou50.calls else None,
            "seg_update": model.val_seg_f1.calls[-1][0] if model.val_seg_f1.calls else None,
            "dice_update": model.val_seg_dice.calls[-1][0] if model.val_seg_dice.calls else None,
            "cm_pairs": list(model.temp_matched_preds_for_cm),
            "seg_logits": getattr(model, "seg_logits_for_logging", None),
            "module": mod,
            "model": model,
        }
        return out
    finally:
        mod.CONF_TH, mod.NMS_IOU, mod.TOP_K = saved
