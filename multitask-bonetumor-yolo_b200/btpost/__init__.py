"""btpost — B200-native post-processing + evaluation hot path (host side).

Everything numeric runs in ``libbtpost.so`` (CUDA, sm_100a) through the C ABI in ``include/btpost.h``.
"""
from . import _lib  # noqa: F401
from .api import (CONF_TH, NMS_IOU, TOP_K, Pipeline, PostConfig, PostProcessor, map_iou_thresholds, num_anchors,  # noqa: F401
                  postprocess_infer, prepare_det_outputs_for_metrics_and_logging, unpack_infer_dict, unpack_seg_outputs)
from .sweep import DeviceSweep, SweepState  # noqa: F401
