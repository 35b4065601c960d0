"""CPU-only: libbtpost.so loads, exports every symbol include/btpost.h declares, and the ctypes
mirror of the POD structs has the layout the C compiler gives them (no compute calls)."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

from btpost import _lib

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "btpost.h"


def declared_symbols():
    txt = "".join(h.read_text() for h in sorted((ROOT / "include").glob("*.h")))   # every header under include/
    return re.findall(r"BTPOST_API\s+[\w\s\*]+?\b(btpost_\w+)\s*\(", txt)


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    names = declared_symbols()
    assert len(names) >= 9 and "btpost_synth_batch" in names and "btpost_masks_parts" in names
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/btpost.h but not exported"
    assert L.btpost_version() == 100
    assert L.btpost_error_string(0) == b"ok"
    assert b"workspace" in L.btpost_error_string(-3)


def test_struct_layout_matches_c(tmp_path):
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "btpost.h"\nint main(){'
                   'printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(BtParams), offsetof(BtParams, iou_thres),'
                   'offsetof(BtParams, iou_thrs), offsetof(BtParams, image_offset), sizeof(BtIO), offsetof(BtIO, dt_match),'
                   'offsetof(BtParams, drop_gt_no_cand), offsetof(BtParams, in_flight), offsetof(BtIO, inst_bits), offsetof(BtIO, sweep));return 0;}')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    want = [C.sizeof(_lib.BtParams), _lib.BtParams.iou_thres.offset, _lib.BtParams.iou_thrs.offset,
            _lib.BtParams.image_offset.offset, C.sizeof(_lib.BtIO), _lib.BtIO.dt_match.offset,
            _lib.BtParams.drop_gt_no_cand.offset, _lib.BtParams.in_flight.offset, _lib.BtIO.inst_bits.offset, _lib.BtIO.sweep.offset]
    assert got == want


def test_argument_validation_without_gpu():
    """Validation happens before any CUDA call, so these error paths run on a CPU-only box."""
    L = _lib.load()
    p, io = _lib.BtParams(), _lib.BtIO()
    n = C.c_size_t()
    assert L.btpost_workspace_bytes(None, C.byref(n)) == -1
    assert L.btpost_workspace_bytes(C.byref(p), C.byref(n)) == -1          # batch = 0
    p.batch, p.num_anchors, p.nc, p.nm, p.max_det, p.max_gt = 2, 8400, 3, 32, 300, 32
    p.img_h = p.img_w = 640
    p.proto_h = p.proto_w = 160
    assert L.btpost_workspace_bytes(C.byref(p), C.byref(n)) == 0 and n.value > 2 * 8400 * 28
    # the sort buffer holds the bitonic network's padded keys or the bucket sort's pairs + index list (12 B / candidate)
    q, m = _lib.BtParams(), C.c_size_t()
    C.memmove(C.byref(q), C.byref(p), C.sizeof(p))
    q.num_anchors = 16384                                                  # a power of two: pairs + indices need 1.5 x the keys
    assert L.btpost_workspace_bytes(C.byref(q), C.byref(m)) == 0
    assert m.value - n.value >= 2 * ((16384 - 8400) * 28 + (16384 * 12 - 16384 * 8)) - 4096   # 256-byte rounding of each buffer
    assert L.btpost_run(C.byref(p), C.byref(io), None, 0, None) == -1      # null workspace
    p.nm = 16
    assert L.btpost_run(C.byref(p), C.byref(io), C.c_void_p(256), n.value, None) == -2   # unsupported nm
    p.nm, p.proto_h = 32, 100
    assert L.btpost_run(C.byref(p), C.byref(io), C.c_void_p(256), n.value, None) == -2   # proto != img/4
    p.proto_h = 160
    assert L.btpost_run(C.byref(p), C.byref(io), C.c_void_p(260), n.value, None) == -4   # misaligned workspace
    assert L.btpost_run(C.byref(p), C.byref(io), C.c_void_p(256), 16, None) == -3        # workspace too small
    assert L.btpost_run(C.byref(p), C.byref(io), C.c_void_p(256), n.value, None) == -1   # null head
    for field, bad in (("nms_threads", 384), ("proto_dtype", 2), ("head_dtype", 7), ("in_flight", -1)):   # scheduling / dtype knobs
        setattr(p, field, bad)
        assert L.btpost_run(C.byref(p), C.byref(io), C.c_void_p(256), n.value, None) == -1, field
        setattr(p, field, 0)
    io.inst_masks = 4096                                                                   # dense instance masks need the bit planes
    io.head = io.protos = io.masks_gt = io.proj_weight = io.dets = io.det_count = io.det_coeff = 4096
    assert L.btpost_masks_parts(C.byref(p), C.byref(io), C.c_void_p(256), n.value, None, 4) == -1
    io = _lib.BtIO()
    assert L.btpost_masks_parts(C.byref(p), C.byref(io), C.c_void_p(256), n.value, None, 8) == -1   # unknown part bit / null inputs
    assert L.btpost_synth_batch(0, 640, 3, 32, None, None, None, None, None, None, None) == -1
    assert L.btpost_synth_batch(2, 640, 3, 32, None, None, None, C.c_void_p(256), None, None, None) == -1   # head without keys / objects
    with pytest.raises(ValueError):
        _lib.check(-2, "x")
    with pytest.raises(RuntimeError):
        _lib.check(-5, "x")
