"""Drop-in for ``utils/nms_wrapper.py`` (reference :23-31) and the Cython glue
``utils/nms/gpu_nms.pyx:16-31``: ``nms(dets, thresh, force_cpu=False)`` -> list of kept row
indices of ``dets`` in score-descending order, pixel coordinates with the +1 convention.

``dets`` may be a float32 numpy array ``[n,5]`` (the reference's contract — host data goes
through ``rd_nms_host``, the replacement of ``_nms`` in ``utils/nms/gpu_nms.hpp``) or a CUDA
tensor ``[n,5]`` (stays on the device, ``rd_nms``).

``force_cpu=True`` selected the Cython ``cpu_nms`` in the reference, whose only observable
difference is suppressing on ``IoU >= thresh`` (utils/nms/cpu_nms.pyx:65) instead of ``>``.
There is no CPU path here: the flag selects that comparison on the GPU.
"""
import ctypes

import numpy as np
import torch

from .. import _ffi
from .._ffi import check, lib


def gpu_nms(dets, thresh, device_id=None, suppress_on_equal=False):
    """utils/nms/gpu_nms.pyx:16-31 over ``rd_nms_host``."""
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    if dets.ndim != 2 or dets.shape[1] < 5:
        raise ValueError('dets must be [n, >=5] (x1,y1,x2,y2,score)')
    n, dim = dets.shape
    if n == 0:
        return []
    if n > _ffi.RD_MAX_NMS_BOXES:
        raise RuntimeError('nms: %d boxes exceed the supported %d per call' % (n, _ffi.RD_MAX_NMS_BOXES))
    if device_id is None:
        device_id = torch.cuda.current_device()
    order = np.argsort(-dets[:, 4], kind='stable')            # score desc, lower index first on ties
    sorted_dets = np.ascontiguousarray(dets[order, :])
    keep = np.zeros(n, dtype=np.int32)
    num_out = ctypes.c_int(0)
    flags = _ffi.RD_NMS_PIXEL_PLUS1 | (_ffi.RD_NMS_SUPPRESS_EQ if suppress_on_equal else 0)
    check(lib().rd_nms_host_ex(keep.ctypes.data_as(ctypes.c_void_p), ctypes.cast(ctypes.byref(num_out), ctypes.c_void_p),
                               sorted_dets.ctypes.data_as(ctypes.c_void_p), n, dim, float(thresh), int(device_id),
                               flags), 'rd_nms_host')
    return order[keep[:num_out.value]].tolist()


def nms(dets, thresh, force_cpu=False):
    """Dispatch of utils/nms_wrapper.py:23-31."""
    if isinstance(dets, torch.Tensor):
        if dets.shape[0] == 0:
            return []
        from ..layers.box_utils import nms_device
        flags = _ffi.RD_NMS_PIXEL_PLUS1 | (_ffi.RD_NMS_SUPPRESS_EQ if force_cpu else 0)
        keep, count = nms_device(dets[:, :4].contiguous(), dets[:, 4].contiguous(), thresh, dets.shape[0], flags)
        return keep[:int(count.item())].tolist()
    if dets.shape[0] == 0:
        return []
    return gpu_nms(dets, thresh, suppress_on_equal=force_cpu)
