"""ctypes binding of ``librefinedet_b200.so`` (the C ABI in ``include/refinedet_b200.h``).

This is the only place the package touches native code.  There is NO fallback: if the
library is missing or a call fails, a ``RuntimeError`` is raised (north_star: "no CPU
fallback"; the reference instead printed CUDA errors and carried on,
``utils/nms/nms_kernel.cu:12-19``).

Pointers come from ``tensor.data_ptr()``, the stream from
``torch.cuda.current_stream().cuda_stream``; ctypes releases the GIL during each call.
"""
import ctypes
import os
import re
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('RD_LIB_PATH') or os.path.join(_HERE, 'lib', 'librefinedet_b200.so')
HEADER_PATH = os.path.join(os.path.dirname(os.path.dirname(_HERE)), 'include', 'refinedet_b200.h')

c_int, c_float, c_void_p, c_size_t = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t

# constants mirrored from the header (checked against it by tests/test_abi.py)
RD_ABI_VERSION = 1
RD_ERR_BAD_ARG, RD_ERR_ALIGNMENT, RD_ERR_UNSUPPORTED, RD_ERR_WORKSPACE = -1, -2, -3, -4
RD_MAX_NMS_BOXES = 4096
RD_MAX_GT = 1024
RD_NMS_NORMALISED, RD_NMS_PIXEL_PLUS1, RD_NMS_SUPPRESS_EQ = 0, 1, 2
RD_INPUT_LOGITS = 4
RD_DEBUG_INSTANCE_SHIFT = 8
RD_DEBUG_INSTANCE_256, RD_DEBUG_INSTANCE_1024 = 1 << 8, 2 << 8
RD_ROW_BOX_SCORE, RD_ROW_SCORE_BOX = 0, 1

_P = c_void_p
_SIGNATURES = {
    'rd_abi_version': (c_int, []),
    'rd_error_string': (ctypes.c_char_p, [c_int]),
    'rd_launch_count': (ctypes.c_ulonglong, []),
    'rd_point_form': (c_int, [_P, _P, c_int, _P]),
    'rd_center_size': (c_int, [_P, _P, c_int, _P]),
    'rd_decode': (c_int, [_P, _P, c_float, c_float, _P, c_int, _P]),
    'rd_encode': (c_int, [_P, _P, c_float, c_float, _P, c_int, _P]),
    'rd_intersect': (c_int, [_P, _P, _P, c_int, c_int, _P]),
    'rd_jaccard': (c_int, [_P, _P, _P, c_int, c_int, _P]),
    'rd_detect_forward': (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_float, c_float, c_float,
                                  _P, _P, _P]),
    'rd_decode_filter': (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_float, c_float, c_float,
                                 _P, _P, _P]),
    'rd_arm_zero_rows': (c_int, [_P, _P, ctypes.c_longlong, c_int, c_float, _P]),
    'rd_select_topk': (c_int, [_P, c_int, c_int, c_int, c_float, c_int, c_int, _P, _P, _P, _P]),
    'rd_detect': (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_float, c_float, c_float,
                          c_int, c_int, _P, c_int, c_int, c_float, c_float, _P, c_size_t,
                          _P, _P, _P, _P]),
    'rd_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'rd_detect_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'rd_detect_workspace_reset': (c_int, [_P, c_size_t, _P]),
    'rd_detect_fused': (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_float, c_float, c_float,
                                c_int, c_int, _P, c_int, c_int, c_float, c_float, _P, c_size_t,
                                _P, _P, _P, _P]),
    'rd_detect_fused_timed': (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_float, c_float, c_float,
                                      c_int, c_int, _P, c_int, c_int, c_float, c_float, _P, c_size_t,
                                      _P, _P, _P, _P, _P]),
    'rd_detect_plan_create': (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_float, c_float, c_float,
                                      c_int, c_int, _P, c_int, c_int, c_float, c_float, _P, c_size_t,
                                      _P, _P, _P, _P]),
    'rd_plan_capture_begin': (c_int, [_P]),
    'rd_plan_capture_end': (c_int, [_P, _P]),
    'rd_detect_plan_launch': (c_int, [_P, _P]),
    'rd_detect_plan_destroy': (c_int, [_P]),
    'rd_pack_detections': (c_int, [_P, _P, c_int, c_int, c_int, _P, _P, c_int, _P]),
    'rd_coco_records': (c_int, [_P, _P, c_int, c_int, c_int, _P, _P, _P, _P, c_int, _P, _P]),
    'rd_exchange_slot_bytes': (c_size_t, [c_int, c_int, c_int]),
    'rd_pack_scatter': (c_int, [_P, _P, c_int, c_int, c_int, _P, _P, c_int, c_int, c_int, c_int, _P]),
    'rd_pack_scatter_ex': (c_int, [_P, _P, c_int, c_int, c_int, _P, _P, c_int, c_int, c_int, c_int, _P, c_int, _P]),
    'rd_exchange_ctrl_bytes': (c_size_t, []),
    'rd_exchange_round': (c_int, [_P, _P, c_int, c_int, c_int, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    'rd_nms_workspace_bytes': (c_size_t, [c_int]),
    'rd_nms': (c_int, [_P, _P, c_int, c_float, c_int, c_int, _P, c_size_t, _P, _P, _P]),
    'rd_nms_host': (c_int, [_P, _P, _P, c_int, c_int, c_float, c_int]),
    'rd_nms_host_ex': (c_int, [_P, _P, _P, c_int, c_int, c_float, c_int, c_int]),
    'rd_match_workspace_bytes': (c_size_t, [c_int, c_int]),
    'rd_refine_match': (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_float, c_float, c_float,
                                c_int, _P, c_size_t, _P, _P, _P, _P, _P]),
    'rd_hnm_select': (c_int, [_P, _P, c_int, c_int, c_int, _P, _P, _P]),
    'rd_pad_targets': (c_int, [_P, _P, c_int, c_int, _P, _P, _P, _P]),
    'rd_conf_loss': (c_int, [_P, _P, _P, c_float, ctypes.c_longlong, c_int, _P, _P, _P, _P]),
    'rd_multibox_loss_workspace_bytes': (c_size_t, [c_int]),
    'rd_multibox_loss_reduce': (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, _P, c_size_t, _P, _P, _P, _P]),
    'rd_multibox_criterion_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'rd_multibox_criterion': (c_int, [_P] * 8 + [c_int] * 4 + [c_float] * 3 + [c_int, c_float, c_int, _P, c_size_t] + [_P] * 9),
    'rd_multibox_loss_backward': (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, ctypes.c_longlong, c_int,
                                          _P, _P, _P]),
    'rd_criterion_state_bytes': (c_size_t, [c_int, c_int, c_int]),
    'rd_criterion_state_layout': (c_int, [c_int, c_int, c_int, _P]),
    'rd_multibox_criterion_pair': (c_int, [_P] * 8 + [c_int] * 4 + [c_float] * 4 + [c_int, c_float, c_int, c_int] +
                                   [_P, _P, c_size_t, _P, _P]),
    'rd_multibox_loss_backward_pair': (c_int, [_P] * 6 + [c_int] * 4 + [_P] * 10),
}

_lib = None
_lock = threading.Lock()


def declared_symbols(header_path=HEADER_PATH):
    """Names of every ``RD_API`` function the header declares."""
    with open(header_path) as f:
        text = f.read()
    return sorted(set(re.findall(r'RD_API\s+[\w\s\*]+?\b(rd_\w+)\s*\(', text)))


def lib():
    """Load (once) and return the shared library; raise loudly when it is unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                'refinedet.pytorch_b200: native library %s is missing. Build it with '
                '`python -m refinedet.pytorch_b200.build` (needs nvcc, targets sm_100a). '
                'There is no CPU / PyTorch fallback for this path.' % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError = ABI mismatch, also loud
            fn.restype = res
            fn.argtypes = args
        if handle.rd_abi_version() != RD_ABI_VERSION:
            raise RuntimeError('refinedet.pytorch_b200: ABI version mismatch (%d != %d); rebuild'
                               % (handle.rd_abi_version(), RD_ABI_VERSION))
        _lib = handle
    return _lib


def check(code, what):
    if code != 0:
        msg = lib().rd_error_string(int(code))
        raise RuntimeError('%s failed: [%d] %s' % (what, code, msg.decode() if msg else '?'))


try:                                    # raw accessors: torch.cuda.current_stream() / torch.cuda.device() cost
    _raw_stream = torch._C._cuda_getCurrentRawStream      # 10-20 us each on the host, more than a launch
    _raw_device = torch._C._cuda_getDevice
except AttributeError:                  # pragma: no cover
    _raw_stream = _raw_device = None


def _device_index(device):
    if device is None:
        return None
    if isinstance(device, int):
        return device
    return torch.device(device).index


def stream_ptr(device=None):
    """The current torch stream of ``device`` (default: the current device) as a ``cudaStream_t``."""
    if _raw_stream is None or not torch.cuda.is_initialized():
        return c_void_p(torch.cuda.current_stream(device).cuda_stream)
    idx = _device_index(device)
    return c_void_p(_raw_stream(_raw_device() if idx is None else idx))


class _NoSwitch(object):
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NO_SWITCH = _NoSwitch()


def on_device(device):
    """Context that makes ``device`` the current CUDA device; free when it already is."""
    idx = _device_index(device)
    if _raw_device is not None and torch.cuda.is_initialized() and (idx is None or idx == _raw_device()):
        return _NO_SWITCH
    return torch.cuda.device(device)


def ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def launch_count():
    return int(lib().rd_launch_count())


def require_cuda_f32(t, name, align=16):
    """Validate a tensor argument; returns a contiguous, aligned fp32 CUDA tensor.

    Non-contiguous or misaligned views are copied (``.contiguous()`` / ``.clone()``);
    anything that is not a CUDA tensor is an error — this package has no CPU path."""
    if not isinstance(t, torch.Tensor):
        raise TypeError('%s must be a torch.Tensor, got %r' % (name, type(t)))
    if not t.is_cuda:
        raise RuntimeError('%s must be a CUDA tensor: refinedet.pytorch_b200 has no CPU fallback '
                           '(got device %s)' % (name, t.device))
    if t.dtype != torch.float32:
        raise TypeError('%s must be float32, got %s' % (name, t.dtype))
    t = t.detach()
    if not t.is_contiguous():
        t = t.contiguous()
    if t.data_ptr() % align:
        t = t.clone()
    return t
