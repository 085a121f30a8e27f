#!/usr/bin/env python
"""BASELINE.json config 1 fixture: the UNMODIFIED reference model (RefineDet320 VGG16, random init, VOC 21
classes) run on the CPU in the authoring container; stores the head outputs that enter
``Detect_RefineDet.forward`` and the reference's outputs for ONE image of the batch of 4 (size), plus
sha256 of the full-batch outputs.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_model.py
"""
import hashlib
import os
import sys
import types
import warnings

import numpy as np
import torch

REF = os.environ.get('RD_REFERENCE', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    warnings.filterwarnings('ignore')
    for n in ['pycocotools', 'pycocotools.coco', 'pycocotools.cocoeval']:
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules['pycocotools.coco'].COCO = object
    sys.modules['pycocotools.cocoeval'].COCOeval = object
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    from layers.functions.detection_refinedet import Detect_RefineDet
    from models.refinedet import build_refinedet
    from utils.nms.py_cpu_nms import py_cpu_nms

    captured = {}

    class Spy(Detect_RefineDet):          # records what models/refinedet.py:141-149 passes to the detector
        def forward(self, arm_loc, arm_conf, odm_loc, odm_conf, priors):
            captured['in'] = [t.detach().clone() for t in (arm_loc, arm_conf, odm_loc, odm_conf, priors)]
            return Detect_RefineDet.forward(self, arm_loc, arm_conf, odm_loc, odm_conf, priors)

    torch.manual_seed(0)
    top_k, conf_thr, nms_thr, obj_thr, keep = 1000, 0.01, 0.45, 0.01, 500
    det = Spy(21, 320, 0, top_k, conf_thr, nms_thr, obj_thr, keep)
    net = build_refinedet('test', 320, 21, detector=det)
    # random-init conv weights give near-uniform softmaxes; scale the head weights so that the scores
    # spread out (still random init, still the reference's code path end to end)
    for m in list(net.arm_conf) + list(net.odm_conf):
        m.weight.data.mul_(40.0)
    for m in list(net.arm_loc) + list(net.odm_loc):
        m.weight.data.mul_(8.0)
    net.eval()
    x = torch.randn(4, 3, 320, 320)
    with torch.no_grad():
        boxes, scores = net(x)
    arm_loc, arm_conf, odm_loc, odm_conf, priors = captured['in']
    sha = lambda t: hashlib.sha256(np.ascontiguousarray(t.numpy()).tobytes()).hexdigest()
    i = 1                                                   # the stored image
    sc = scores[i].numpy()
    bx = (boxes[i] * torch.tensor([320.0] * 4)).numpy()
    a4_counts = np.zeros(21, np.int32)
    a4_tie_free = np.zeros(21, np.bool_)
    a4 = np.zeros((21, keep, 5), np.float32)
    for j in range(1, 21):                                  # eval_refinedet_coco.py:213-232, reference py_cpu_nms
        inds = np.where(sc[:, j] > conf_thr)[0]
        if len(inds) == 0:
            continue
        c_scores = sc[inds, j]
        order = c_scores.argsort()[::-1][:top_k]
        a4_tie_free[j] = np.unique(c_scores).size == c_scores.size     # ties: reference order is unstable
        c_dets = np.hstack((bx[inds][order], c_scores[order][:, None])).astype(np.float32, copy=False)
        keep_idx = py_cpu_nms(c_dets, nms_thr)
        c_dets = c_dets[keep_idx, :][:keep, :]
        a4_counts[j] = c_dets.shape[0]
        a4[j, :c_dets.shape[0]] = c_dets
    print('passing ARM: %.3f, candidates/class: %.1f, kept/class: %.1f' % (
        float((arm_conf[i, :, 1] > obj_thr).float().mean()), np.mean([(sc[:, j] > conf_thr).sum() for j in range(1, 21)]),
        a4_counts[1:].mean()), 'tie-free classes:', int(a4_tie_free.sum()))
    np.savez_compressed(os.path.join(HERE, 'model_cfg1.npz'),
                        arm_loc=arm_loc[i].numpy(), arm_conf=arm_conf[i].numpy(), odm_loc=odm_loc[i].numpy(),
                        odm_conf=odm_conf[i].numpy(), priors=priors.numpy(), boxes=boxes[i].numpy(),
                        scores=scores[i].numpy(), a4_counts=a4_counts, a4_dets=a4, a4_tie_free=a4_tie_free,
                        params=np.array([21, top_k, keep, conf_thr, nms_thr, obj_thr], np.float64),
                        sha_boxes_batch=sha(boxes), sha_scores_batch=sha(scores))
    print('written', os.path.getsize(os.path.join(HERE, 'model_cfg1.npz')), 'bytes')


if __name__ == '__main__':
    main()
