"""GPU tier: ``rd_nms_host`` (the replacement of ``_nms``, utils/nms/gpu_nms.hpp:1-2) against the
reference's OWN compiled GPU kernel — ``oracle/_ref/libref_gpu_nms.so``, built by
``oracle/build_ref.py`` from the unmodified ``utils/nms/nms_kernel.cu`` — and against the numpy
restatement of ``py_cpu_nms``.  Kept lists must be identical (index work: bit-exact).

The reference builds its kernel with nvcc's default ``-fmad=true``, so an IoU within
an ulp of the threshold could in principle round differently there than in its CPU siblings; the
seeded inputs below have no such pair (asserted through the three-way agreement with the oracle).
"""
import json
import os
import time

import numpy as np
import pytest

from oracle import box_oracle as bo
from oracle import ref_nms

pytestmark = pytest.mark.gpu

needs_ref = pytest.mark.skipif(not ref_nms.available(), reason='oracle/_ref/libref_gpu_nms.so not built '
                                                                '(python oracle/build_ref.py in the authoring container)')


@pytest.fixture(scope='module')
def rd():
    import refinedet.pytorch_b200 as rd
    rd._ffi.lib()
    return rd


def make_dets(seed, n, kind):
    rng = np.random.default_rng(seed)
    if kind == 'spread':                      # eval-like: boxes all over a 512 x 512 image
        xy = rng.uniform(0, 450, (n, 2))
        wh = rng.uniform(10, 90, (n, 2))
    elif kind == 'clusters':                  # detector-like: many near-duplicates around a few objects
        centres = rng.uniform(60, 450, (12, 2))
        xy = centres[rng.integers(0, 12, n)] + rng.normal(0, 6, (n, 2)) - 35
        wh = rng.uniform(60, 75, (n, 2))
    else:                                     # tiny boxes (SAR-ship like), hardly any overlap
        xy = rng.uniform(0, 500, (n, 2))
        wh = rng.uniform(4, 14, (n, 2))
    sc = rng.permutation(n).astype(np.float32) / n * 0.98 + 0.01          # pairwise distinct scores
    return np.concatenate([xy, xy + wh, sc[:, None]], 1).astype(np.float32)


@needs_ref
@pytest.mark.parametrize('n', [1, 2, 63, 64, 65, 500, 1000, 2000])
@pytest.mark.parametrize('kind', ['spread', 'clusters', 'tiny'])
@pytest.mark.parametrize('thr', [0.3, 0.45, 0.49])
def test_rd_nms_host_equals_reference_gpu_kernel(rd, n, kind, thr):
    dets = make_dets(1000 * n + int(100 * thr), n, kind)
    ref = [int(i) for i in ref_nms.gpu_nms(dets, thr)]
    ours = rd.nms_wrapper.nms(dets, thr)
    assert ours == ref
    assert ours == bo.nms_pixel(dets, thr)


@needs_ref
def test_rd_nms_host_vs_reference_gpu_kernel_eval_loop(rd):
    """The call pattern of eval_refinedet_coco.py:213-232: one call per (image, class) on <= top_k
    rows.  Same kept lists; the time per call of both is recorded in gpurun_out/ for profiles/."""
    calls = [make_dets(7 + i, n, kind) for i, (n, kind) in enumerate(
        [(1000, 'spread'), (1000, 'clusters'), (153, 'spread'), (435, 'tiny')] * 5)]
    for d in calls:
        assert rd.nms_wrapper.nms(d, 0.45) == [int(i) for i in ref_nms.gpu_nms(d, 0.45)]
    out = {}
    for name, fn in (('reference_nms_kernel_cu', lambda d: ref_nms.gpu_nms(d, 0.45)),
                     ('rd_nms_host', lambda d: rd.nms_wrapper.nms(d, 0.45))):
        for d in calls:
            fn(d)
        per = {}
        for tag, sel in (('1000_spread', 0), ('1000_clusters', 1), ('153_spread', 2), ('435_tiny', 3)):
            d = calls[sel]
            t0 = time.perf_counter()
            for _ in range(100):
                fn(d)
            per[tag] = (time.perf_counter() - t0) / 100 * 1e6
        out[name] = per
    out['unit'] = 'us per call, host numpy dets in, kept list out (argsort + native call + index map), 100 calls'
    os.makedirs('gpurun_out', exist_ok=True)
    with open(os.path.join('gpurun_out', 'ref_gpu_nms_timing.json'), 'w') as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))
