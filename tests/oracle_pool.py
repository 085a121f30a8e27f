"""The a4 oracle (``oracle.box_oracle.detect_stage_eval``: eval_refinedet_coco.py:205-232 with
utils/nms/py_cpu_nms.py:10-38) over a whole batch, images fanned out over a fork pool so that the
B = 32 headline configurations (2,560 (image, class) problems) finish in seconds.

TEST INFRASTRUCTURE: imported by ``tests/`` only."""
import multiprocessing as mp
import os

import numpy as np

from oracle import box_oracle as bo

_JOB = {}


def _one(b):
    j = _JOB
    scale = j['scale'][b] if j['scale'].ndim == 2 else j['scale']
    out, anc = bo.detect_stage_eval(j['boxes'][b], j['scores'][b], scale, j['conf_thr'], j['top_k'], j['nms_thr'],
                                    j['keep'], j['suppress_eq'])
    return b, out, anc


def detect_stage_eval_batch(boxes, scores, scale, conf_thr, top_k, nms_thr, keep, suppress_on_equal=False,
                            workers=None):
    """``boxes[B,P,4]``, ``scores[B,P,C]`` (numpy f32), ``scale[4]`` or ``[B,4]`` ->
    ``(counts[B,C] int32, anchors{(b,c): int64[n]}, rows{(b,c): f32[n,5]})``."""
    B, P, C = scores.shape
    _JOB.clear()
    _JOB.update(boxes=boxes, scores=scores, scale=np.asarray(scale, np.float32), conf_thr=conf_thr, top_k=top_k,
                nms_thr=nms_thr, keep=keep, suppress_eq=suppress_on_equal)
    if workers is None:
        workers = max(1, min(B, len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)))
    if workers == 1:
        results = [_one(b) for b in range(B)]
    else:
        # fork: the workers inherit the arrays (numpy only; they never touch CUDA)
        with mp.get_context('fork').Pool(workers) as pool:
            results = pool.map(_one, range(B), chunksize=1)
    _JOB.clear()
    counts = np.zeros((B, C), np.int32)
    anchors, rows = {}, {}
    for b, out, anc in results:
        for c in range(C):
            counts[b, c] = out[c].shape[0]
            anchors[b, c] = anc[c]
            rows[b, c] = out[c]
    return counts, anchors, rows


def assert_detections_equal(res, counts, anchors, rows, what=''):
    """``res`` (a ``Detections`` on the GPU) against the oracle triple: counts, anchor lists and rows bit-exact
    for EVERY (image, class) problem."""
    g_counts = res.counts.cpu().numpy()
    bad = np.argwhere(g_counts != counts)
    assert bad.size == 0, '%s: %d of %d problems differ in count, first (b,c)=%s gpu %d oracle %d' % (
        what, len(bad), counts.size, tuple(bad[0]), g_counts[tuple(bad[0])], counts[tuple(bad[0])])
    g_anchor = res.anchors.cpu().numpy()
    g_dets = res.dets.cpu().numpy()
    B, C = counts.shape
    for b in range(B):
        for c in range(C):
            n = counts[b, c]
            assert np.array_equal(g_anchor[b, c, :n], anchors[b, c]), (what, b, c)
            assert np.array_equal(g_dets[b, c, :n], rows[b, c].reshape(-1, 5)), (what, b, c)
    assert (g_counts[:, 0] == 0).all()
    return int(counts.sum())
