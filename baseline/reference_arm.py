"""Drives the UNMODIFIED reference staged under ``baseline/_ref/`` (see ``build_ref.py``) on the host cores.

Detect stage, one image per call:
    ``Detect_RefineDet.forward`` (layers/functions/detection_refinedet.py:27-65, the reference's own code) followed
    by the per-class loop of ``eval_refinedet_coco.py:205-232``.  That loop is script-level code inside ``test_net``
    (the script builds an argparse parser and a network at import), so it cannot be imported; it is restated here
    line for line around the reference's own ``py_cpu_nms`` (utils/nms/py_cpu_nms.py:10-38 — the ``nms`` the wrapper
    would dispatch to needs the Cython build, which fails with Cython 3 / numpy 2; ``py_cpu_nms`` has
    the GPU kernel's semantics).
Training step:
    the reference's ``RefineDetMultiBoxLoss`` (layers/modules/refinedet_multibox_loss.py:50-139), ARM + ODM criteria,
    forward + backward, exactly as train_refinedet.py:181-184,252-256 uses them.

TEST / BENCH INFRASTRUCTURE: only ``bench.py`` (``--impl reference``, ``cpu_baseline``) and ``tests/`` import this."""
import os
import sys
import time
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')
_REF = None


def available():
    return os.path.isfile(os.path.join(REF_DIR, 'layers', 'functions', 'detection_refinedet.py'))


def load_reference():
    """Import the staged reference (pycocotools stubbed: the eval-time dataset class is never touched)."""
    global _REF
    if _REF is not None:
        return _REF
    if not available():
        raise RuntimeError('baseline/_ref is missing: run __graft_entry__.build() where /root/reference exists')
    for n in ('pycocotools', 'pycocotools.coco', 'pycocotools.cocoeval'):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules['pycocotools.coco'].COCO = getattr(sys.modules['pycocotools.coco'], 'COCO', object)
    sys.modules['pycocotools.cocoeval'].COCOeval = getattr(sys.modules['pycocotools.cocoeval'], 'COCOeval', object)
    sys.dont_write_bytecode = True
    import warnings
    warnings.filterwarnings('ignore')
    sys.path.insert(0, REF_DIR)
    try:
        from layers.functions.detection_refinedet import Detect_RefineDet
        from layers.functions.prior_box import PriorBox
        from layers.modules.refinedet_multibox_loss import RefineDetMultiBoxLoss
        from data import coco_refinedet, voc_refinedet
        from utils.nms.py_cpu_nms import py_cpu_nms
    finally:
        sys.path.remove(REF_DIR)
    _REF = types.SimpleNamespace(Detect=Detect_RefineDet, PriorBox=PriorBox, Loss=RefineDetMultiBoxLoss,
                                 coco=coco_refinedet, voc=voc_refinedet, py_cpu_nms=py_cpu_nms)
    return _REF


def detect_one_image(R, detector, arm_loc, arm_conf, odm_loc, odm_conf, priors, scale, confidence_threshold, top_k,
                     nms_threshold, max_per_image):
    """The reference's detect stage for a batch of ONE image, as evaluated.  Tensors are CPU torch tensors;
    ``odm_conf`` is modified in place by the reference (:40-42), so callers pass a copy.  Returns ``all_boxes[j]``."""
    boxes, scores = detector.forward(arm_loc, arm_conf, odm_loc, odm_conf, priors)     # models/refinedet.py:141
    boxes = boxes[0]                                       # eval_refinedet_coco.py:205-211
    scores = scores[0]
    boxes *= scale
    boxes = boxes.cpu().numpy()
    scores = scores.cpu().numpy()
    num_classes = scores.shape[1]
    out = [np.empty([0, 5], dtype=np.float32)]
    for j in range(1, num_classes):                        # :213-232
        inds = np.where(scores[:, j] > confidence_threshold)[0]
        if len(inds) == 0:
            out.append(np.empty([0, 5], dtype=np.float32))
            continue
        c_bboxes = boxes[inds]
        c_scores = scores[inds, j]
        order = c_scores.argsort()[::-1][:top_k]
        c_bboxes = c_bboxes[order]
        c_scores = c_scores[order]
        c_dets = np.hstack((c_bboxes, c_scores[:, np.newaxis])).astype(np.float32, copy=False)
        keep = R.py_cpu_nms(c_dets, nms_threshold)
        c_dets = c_dets[keep, :]
        c_dets = c_dets[:max_per_image, :]
        out.append(c_dets)
    return out


# ---- worker-process entry points (fork pool; one image per worker per step) ---------------------------------
_W = {}


def worker_detect(args):
    """``(seed, kind, P, C, size, thresholds)``: first call generates the image's inputs (untimed, returns 0),
    later calls time one pass of the reference's detect stage over it and return the seconds."""
    seed, kind, P, C, size, conf_thr, top_k, nms_thr, keep, obj_thr = args
    key = args
    if _W.get('key') != key:
        torch.set_num_threads(1)
        from refinedet.pytorch_b200 import synthetic          # the shared seeded generator (inputs only)
        R = load_reference()
        a = synthetic.detect_inputs(seed, 1, P, C, kind)
        priors = R.PriorBox(R.coco[size] if size == '512' else R.voc[size]).forward()
        det = R.Detect(C, int(size), 0, top_k, conf_thr, nms_thr, obj_thr, keep)
        _W.clear()
        _W.update(key=key, a=a, priors=priors, det=det, R=R, scale=torch.tensor([float(size)] * 4))
        return 0.0
    a, R = _W['a'], _W['R']
    t0 = time.perf_counter()
    with torch.no_grad():
        detect_one_image(R, _W['det'], a[0], a[1], a[2], a[3].clone(), _W['priors'], _W['scale'], conf_thr, top_k,
                         nms_thr, keep)
    return time.perf_counter() - t0


def run_detect(seed0, kind, P, C, size, thresholds, steps, warmup, cores, budget_s=None):
    """Each step = ``cores`` images, one per worker process.  Returns ``(images/s, steps done, seconds)``."""
    import multiprocessing as mp
    jobs = [(seed0 + 7 * w, kind, P, C, size) + tuple(thresholds) for w in range(cores)]
    with mp.get_context('fork').Pool(cores) as pool:
        pool.map(worker_detect, jobs, chunksize=1)                # inputs + reference objects in the workers
        for _ in range(max(0, warmup)):
            pool.map(worker_detect, jobs, chunksize=1)
        t_begin = time.perf_counter()
        done = 0
        for _ in range(max(1, steps)):
            pool.map(worker_detect, jobs, chunksize=1)
            done += 1
            if budget_s is not None and time.perf_counter() - t_begin > budget_s:
                break
        elapsed = time.perf_counter() - t_begin
    return cores * done / elapsed, done, elapsed


def run_train_step(seed_pred, seed_tgt, B, P, C, G, steps=2, threads=None):
    """Reference ARM + ODM ``RefineDetMultiBoxLoss`` forward + backward on the CPU (``use_gpu=False``), all host
    threads given to torch.  Returns ``(seconds per step, losses of the last step)``."""
    from refinedet.pytorch_b200 import synthetic
    R = load_reference()
    if threads:
        torch.set_num_threads(int(threads))
    tp = synthetic.train_predictions(seed_pred, B, P, C)
    tg = synthetic.targets(seed_tgt, B, G, C)
    priors = R.PriorBox(R.coco['512']).forward()[:P]
    arm_crit = R.Loss(2, 0.5, True, 0, True, 3, 0.5, False, False)                  # train_refinedet.py:181-184
    odm_crit = R.Loss(C, 0.5, True, 0, True, 3, 0.5, False, False, use_ARM=True)
    leaves = [t.clone().requires_grad_(True) for t in tp]
    vals = None

    def one():
        preds = (leaves[0], leaves[1], leaves[2], leaves[3], priors)
        al, ac = arm_crit(preds, tg)                                               # :252-256
        ol, oc = odm_crit(preds, tg)
        (al + ac + ol + oc).backward()
        for t in leaves:
            t.grad = None
        return [float(al), float(ac), float(ol), float(oc)]
    one()
    t0 = time.perf_counter()
    for _ in range(steps):
        vals = one()
    return (time.perf_counter() - t0) / steps, vals
