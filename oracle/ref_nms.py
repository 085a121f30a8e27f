"""TEST INFRASTRUCTURE — ctypes binding of the reference's own compiled GPU NMS (``oracle/_ref``).

``_nms`` (``utils/nms/gpu_nms.hpp:1-2``, defined at ``utils/nms/nms_kernel.cu:91-144``) is a C++
symbol (no ``extern "C"``), hence the mangled name.  :func:`gpu_nms` repeats what the Cython wrapper
``utils/nms/gpu_nms.pyx:16-31`` does around it: sort by score descending, call, map the kept
positions back through the order.  Only tests/, ``__graft_entry__.smoke()`` and bench.py's baseline
legs may use this module; the product never does.
"""
import ctypes
import os

import numpy as np

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref', 'libref_gpu_nms.so')
_SYMBOL = '_Z4_nmsPiS_PKfiifi'    # void _nms(int*, int*, const float*, int, int, float, int)
_lib = None


def available():
    return os.path.exists(LIB)


def _nms():
    global _lib
    if _lib is None:
        h = ctypes.CDLL(LIB)
        f = getattr(h, _SYMBOL)
        f.restype = None
        f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                      ctypes.c_float, ctypes.c_int]
        _lib = f
    return _lib


def gpu_nms_sorted(sorted_dets, thresh, device_id=0):
    """``_nms`` on rows already in score-descending order; returns the kept row positions (int32)."""
    sorted_dets = np.ascontiguousarray(sorted_dets, dtype=np.float32)
    n, dim = sorted_dets.shape
    keep = np.zeros(n, dtype=np.int32)
    num_out = ctypes.c_int(0)
    _nms()(keep.ctypes.data, ctypes.addressof(num_out), sorted_dets.ctypes.data, n, dim, float(thresh),
           int(device_id))
    return keep[:num_out.value]


def gpu_nms(dets, thresh, device_id=0):
    """``gpu_nms.pyx:16-31``: returns the list of kept row indices of ``dets[n,5]``."""
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    order = dets[:, 4].argsort()[::-1]
    keep = gpu_nms_sorted(dets[order], thresh, device_id)
    return list(order[keep])
