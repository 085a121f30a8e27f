"""No-GPU tier: the C-ABI library builds for sm_100a, loads, and exports every symbol the
header declares; the Python mirror of the header constants is in sync; the product path
refuses to run without CUDA tensors (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from refinedet.pytorch_b200 import _ffi, build


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.exists(path)
    handle = ctypes.CDLL(path)
    declared = _ffi.declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(handle, name), 'missing export %s' % name
    assert set(declared) == set(_ffi._SIGNATURES), 'ctypes signatures out of sync with the header'
    assert _ffi.lib().rd_abi_version() == _ffi.RD_ABI_VERSION


def test_header_constants_match_python_mirror():
    text = open(_ffi.HEADER_PATH).read()
    defs = dict(re.findall(r'#define\s+(RD_\w+)\s+\(?(-?\d+)\)?', text))
    for name in ('RD_ABI_VERSION', 'RD_ERR_BAD_ARG', 'RD_ERR_ALIGNMENT', 'RD_ERR_UNSUPPORTED', 'RD_ERR_WORKSPACE',
                 'RD_MAX_NMS_BOXES', 'RD_MAX_GT', 'RD_NMS_NORMALISED', 'RD_NMS_PIXEL_PLUS1', 'RD_NMS_SUPPRESS_EQ',
                 'RD_ROW_BOX_SCORE', 'RD_ROW_SCORE_BOX', 'RD_INPUT_LOGITS'):
        assert int(defs[name]) == getattr(_ffi, name), name


def test_error_strings_and_argument_validation_without_gpu():
    L = _ffi.lib()
    assert L.rd_error_string(0) == b'success'
    assert b'bad argument' in L.rd_error_string(_ffi.RD_ERR_BAD_ARG)
    assert b'aligned' in L.rd_error_string(_ffi.RD_ERR_ALIGNMENT)
    # argument errors are detected before any CUDA call
    assert L.rd_decode(None, None, 0.1, 0.2, None, 4, None) == _ffi.RD_ERR_BAD_ARG
    assert L.rd_detect_fused(*([None] * 5), 1, 1, 1, 0.0, 0.0, 0.5, 1, 1, None, 0, 0, 0.1, 0.2, None, 0,
                             None, None, None, None) == _ffi.RD_ERR_BAD_ARG
    assert L.rd_detect(*([None] * 5), 1, 1, 1, 0.0, 0.0, 0.5, 1, 1, None, 0, 0, 0.1, 0.2, None, 0,
                       None, None, None, None) == _ffi.RD_ERR_BAD_ARG
    assert L.rd_decode_filter(*([None] * 5), 1, 1, 1, 0.0, 0.1, 0.2, None, None, None) == _ffi.RD_ERR_BAD_ARG
    assert L.rd_select_topk(None, 1, 1, 1, 0.0, 1, 1, None, None, None, None) == _ffi.RD_ERR_BAD_ARG
    assert L.rd_workspace_bytes(32, 16320, 81) == L.rd_detect_workspace_bytes(32, 16320, 81)
    assert L.rd_detect_workspace_bytes(32, 16320, 81) > 32 * 81 * 16320 * 8
    assert L.rd_nms_workspace_bytes(1000) >= 1000 * 8


def test_product_path_has_no_cpu_fallback():
    from refinedet.pytorch_b200 import Detect_RefineDet, box_utils
    x = torch.zeros(4, 4)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        box_utils.decode(x, x, [0.1, 0.2])
    det = Detect_RefineDet(3, 320, 0, 10, 0.01, 0.45, 0.01, 5)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        det.forward(torch.zeros(1, 4, 4), torch.zeros(1, 4, 2), torch.zeros(1, 4, 4), torch.zeros(1, 4, 3),
                    torch.zeros(4, 4))
    with pytest.raises(ValueError):
        Detect_RefineDet(3, 320, 0, 10, 0.01, 0.0, 0.01, 5)     # reference :21-22


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_ffi, '_lib', None)
    monkeypatch.setattr(_ffi, 'LIB_PATH', '/nonexistent/librefinedet_b200.so')
    with pytest.raises(RuntimeError, match='missing'):
        _ffi.lib()


def test_priorbox_matches_reference_layout(golden):
    import hashlib
    from refinedet.pytorch_b200 import PriorBox, REFINEDET_ANCHORS
    g = golden('priors.npz')
    for size in ('320', '512'):
        p = PriorBox(REFINEDET_ANCHORS[size]).forward().numpy()
        assert hashlib.sha256(p.tobytes()).hexdigest() == str(g['sha' + size])


def test_reference_gpu_nms_checker_loads():
    """oracle/_ref (the reference's own nms_kernel.cu, built unmodified by oracle/build_ref.py) loads
    and exports ``_nms`` under its C++ name; no compute without a GPU."""
    import ctypes
    from oracle import build_ref, ref_nms
    path = build_ref.build_ref()
    if path is None:
        pytest.skip('no reference checkout / nvcc and no prebuilt oracle/_ref')
    assert path == ref_nms.LIB
    assert hasattr(ctypes.CDLL(path), ref_nms._SYMBOL)
