#!/usr/bin/env python
"""TEST INFRASTRUCTURE — builds ``oracle/_ref/libref_gpu_nms.so`` from the UNMODIFIED reference source.

The only compiled code on the reference's hot path is its GPU NMS:
``utils/nms/nms_kernel.cu`` (`nms_kernel` :34-76, `_nms` :91-144) + ``utils/nms/gpu_nms.hpp:1-2``.
It compiles from those two files alone, so this recipe runs ``nvcc`` on them WHERE THEY LIE under
``/root/reference`` (nothing is copied into the repo) with the reference's own flags
(``utils/build.py:131-135``: ``-O``-level default, no ``-fmad`` flag), only the architecture changed
from ``sm_52`` to ``sm_100a``.  The output goes to ``oracle/_ref/`` (git-ignored, not gpurun-ignored:
it travels to the GPU box like the product's own ``.so``).

The Cython siblings (``cpu_nms.pyx``, ``gpu_nms.pyx``) do not build unmodified with Cython 3 /
numpy 2, so they are not built; ``oracle/ref_nms.py`` binds ``_nms`` directly with
ctypes, passing exactly what ``gpu_nms.pyx:16-31`` passes.

    python oracle/build_ref.py            # no-op when /root/reference is absent (GPU box)
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')
REF_SRC = '/root/reference/utils/nms'
LIB = os.path.join(REF_DIR, 'libref_gpu_nms.so')


def build_ref(force=False):
    """Returns the library path, or None when the reference checkout (or nvcc) is not there."""
    src = os.path.join(REF_SRC, 'nms_kernel.cu')
    if not os.path.exists(src):
        return LIB if os.path.exists(LIB) else None
    if os.path.exists(LIB) and not force and os.path.getmtime(LIB) >= os.path.getmtime(src):
        return LIB
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        return None
    os.makedirs(REF_DIR, exist_ok=True)
    tmp = LIB + '.tmp.%d' % os.getpid()
    cmd = [nvcc, '-shared', '-Xcompiler', '-fPIC', '-gencode', 'arch=compute_100a,code=sm_100a',
           '-I', REF_SRC, src, '-o', tmp]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('reference nms_kernel.cu did not compile:\n%s\n%s' % (r.stdout, r.stderr))
    os.replace(tmp, LIB)
    return LIB


if __name__ == '__main__':
    print(build_ref(force='--force' in sys.argv))
