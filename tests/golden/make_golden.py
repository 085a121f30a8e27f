#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the authoring container only (``/root/reference`` does not exist on the
GPU box):  ``PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py``

The reference has no tests or golden vectors of its own, so the
pin is the reference's code itself: this script imports ``layers.box_utils``,
``Detect_RefineDet``, ``RefineDetMultiBoxLoss``, ``PriorBox`` and
``utils/nms/py_cpu_nms.py`` from ``/root/reference`` (with a ``pycocotools`` stub,
SURVEY.md §8c), feeds them small seeded inputs and stores inputs + outputs as
``.npz``.  The only logic restated here is glue the reference keeps inside
script bodies that cannot be imported: the per-class loop of
``eval_refinedet_coco.py:213-232`` (calling the reference's own ``py_cpu_nms``)
and the mask expressions of ``refinedet_multibox_loss.py:96-123`` (to expose the
``pos``/``neg`` masks the module does not return).
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


def import_reference():
    for n in ['pycocotools', 'pycocotools.coco', 'pycocotools.cocoeval']:
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules['pycocotools.coco'].COCO = object
    sys.modules['pycocotools.cocoeval'].COCOeval = object
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    import layers.box_utils as bu
    from layers.functions.detection_refinedet import Detect_RefineDet
    from layers.functions.prior_box import PriorBox
    from layers.modules.refinedet_multibox_loss import RefineDetMultiBoxLoss
    from data import voc_refinedet, coco_refinedet
    from utils.nms.py_cpu_nms import py_cpu_nms
    return types.SimpleNamespace(bu=bu, Detect=Detect_RefineDet, PriorBox=PriorBox,
                                 Loss=RefineDetMultiBoxLoss, voc=voc_refinedet,
                                 coco=coco_refinedet, py_cpu_nms=py_cpu_nms)


def gen_detect_inputs(g, B, P, C, sparse):
    """SURVEY.md §8d generators (dense / sparse), on CPU with a seeded generator."""
    if sparse:
        loc_s = 1.0
        d = 2.0 * torch.randn(B, P, generator=g) - 3.0        # milder than -8 so small P keeps candidates
        arm_conf = torch.softmax(torch.stack([torch.zeros(B, P), d], -1), -1)
        logits = 1.5 * torch.randn(B, P, C, generator=g)
        logits[..., 0] += 2.0
        odm_conf = torch.softmax(logits, -1)
    else:
        loc_s = 0.5
        arm_conf = torch.softmax(3 * torch.randn(B, P, 2, generator=g), -1)
        odm_conf = torch.softmax(3 * torch.randn(B, P, C, generator=g), -1)
    arm_loc = loc_s * torch.randn(B, P, 4, generator=g)
    odm_loc = loc_s * torch.randn(B, P, 4, generator=g)
    return arm_loc, arm_conf, odm_loc, odm_conf


def gen_targets(g, B, G, num_classes, wh_lo=0.02, wh_hi=0.17):
    out = []
    for _ in range(B):
        xy = torch.rand(G, 2, generator=g) * 0.8
        wh = wh_lo + torch.rand(G, 2, generator=g) * (wh_hi - wh_lo)
        x2y2 = torch.clamp(xy + wh, max=1.0)
        lab = torch.randint(1, max(num_classes, 2), (G, 1), generator=g).float()
        out.append(torch.cat([xy, x2y2, lab], 1))
    return out


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    warnings.filterwarnings('ignore')
    R = import_reference()
    bu = R.bu
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(20261018)
    save = lambda name, **kw: np.savez_compressed(os.path.join(HERE, name), **kw)

    # ---- a0 priors -----------------------------------------------------------
    pri = {}
    for size in ('320', '512'):
        p = R.PriorBox(R.voc[size]).forward()
        p2 = R.PriorBox(R.coco[size]).forward()
        assert torch.equal(p, p2)
        pri[size] = p
    save('priors.npz',
         sum320=np.float64(pri['320'].double().sum().item()), sha320=sha(pri['320'].numpy()),
         rows320=pri['320'][[0, 1, 2, 4800, 6000, 6300, 6374]].numpy(),
         sub320=pri['320'][::25].numpy(),
         sum512=np.float64(pri['512'].double().sum().item()), sha512=sha(pri['512'].numpy()),
         rows512=pri['512'][[0, 1, 2, 12288, 15360, 16128, 16319]].numpy(),
         sub512=pri['512'][::64].numpy())

    # ---- box_utils elementwise ----------------------------------------------
    P = 96
    priors = pri['320'][::67][:P].contiguous()
    loc = torch.randn(P, 4, generator=g)
    var = [0.1, 0.2]
    dec = bu.decode(loc, priors, var)
    G = 7
    xy = torch.rand(G, 2, generator=g) * 0.7
    truths = torch.cat([xy, xy + 0.05 + 0.3 * torch.rand(G, 2, generator=g)], 1)
    matched = truths[torch.randint(0, G, (P,), generator=g)]
    x = 3 * torch.randn(40, 9, generator=g)
    save('box_utils.npz', priors=priors.numpy(), loc=loc.numpy(), truths=truths.numpy(),
         matched=matched.numpy(), x=x.numpy(),
         point_form=bu.point_form(priors).numpy(), center_size=bu.center_size(dec).numpy(),
         decode=dec.numpy(), encode=bu.encode(matched, priors, var).numpy(),
         intersect=bu.intersect(truths, bu.point_form(priors)).numpy(),
         jaccard=bu.jaccard(truths, bu.point_form(priors)).numpy(),
         jaccard_dec=bu.jaccard(truths, dec).numpy(),
         log_sum_exp=bu.log_sum_exp(x).numpy(),
         # Appendix B KAT
         kat_pri=np.array([[0.5, 0.5, 0.2, 0.1], [0.0125, 0.0125, 0.1, 0.1]], np.float32),
         kat_arm=np.array([[0.3, -0.2, 0.5, -0.4], [-1, 2, 0.1, 0]], np.float32),
         kat_odm=np.array([[-0.7, 0.9, -0.3, 0.25], [0.05, 0.05, 1.5, -1.5]], np.float32))

    # ---- box_utils.nms (a6) ---------------------------------------------------
    kat_boxes = torch.tensor([[.10, .10, .50, .50], [.12, .12, .52, .52], [.60, .60, .90, .90],
                              [.11, .09, .49, .51], [.61, .62, .88, .91], [.30, .30, .70, .70]])
    kat_scores = torch.tensor([.9, .8, .7, .95, .75, .6])
    k1, c1 = bu.nms(kat_boxes, kat_scores, 0.45, 200)
    k2, c2 = bu.nms(kat_boxes, kat_scores, 0.45, 3)
    empty = bu.nms(torch.zeros(0, 4), torch.zeros(0), 0.45, 200)
    assert isinstance(empty, torch.Tensor) and empty.numel() == 0
    n = 300
    cxy = torch.rand(n, 2, generator=g)
    wh = 0.05 + 0.25 * torch.rand(n, 2, generator=g)
    rb = torch.cat([cxy - wh / 2, cxy + wh / 2], 1)
    rs = torch.rand(n, generator=g)
    assert rs.unique().numel() == n
    r = {}
    for tag, thr, tk in (('a', 0.45, 200), ('b', 0.3, 50), ('c', 0.7, 1000)):
        k, c = bu.nms(rb, rs, thr, tk)
        r['keep_' + tag], r['count_' + tag] = k.numpy(), np.int64(c)
        r['args_' + tag] = np.array([thr, tk], np.float64)
    save('nms_box_utils.npz', kat_boxes=kat_boxes.numpy(), kat_scores=kat_scores.numpy(),
         kat_keep200=k1.numpy(), kat_count200=np.int64(c1), kat_keep3=k2.numpy(),
         kat_count3=np.int64(c2), boxes=rb.numpy(), scores=rs.numpy(), **r)

    # ---- py_cpu_nms (a7) ------------------------------------------------------
    n = 400
    cxy = 512 * torch.rand(n, 2, generator=g)
    wh = 20 + 120 * torch.rand(n, 2, generator=g)
    dets = torch.cat([cxy - wh / 2, cxy + wh / 2, torch.rand(n, 1, generator=g)], 1).numpy()
    assert np.unique(dets[:, 4]).size == n
    save('nms_pixel.npz', dets=dets,
         keep_045=np.array(R.py_cpu_nms(dets, 0.45), np.int64),
         keep_049=np.array(R.py_cpu_nms(dets, 0.49), np.int64),
         keep_070=np.array(R.py_cpu_nms(dets, 0.7), np.int64))

    # ---- Detect_RefineDet (a3, a4, a5) ---------------------------------------
    for tag, sparse, C in (('sparse', True, 6), ('dense', False, 4)):
        B = 2
        priors = pri['320'][::5].contiguous()                      # 1275 anchors, all 4 levels
        P = priors.shape[0]
        arm_loc, arm_conf, odm_loc, odm_conf = gen_detect_inputs(g, B, P, C, sparse)
        top_k, keep_top_k, conf_thr, nms_thr, obj_thr = 60, 25, 0.01, 0.45, 0.01
        det = R.Detect(C, 320, 0, top_k, conf_thr, nms_thr, obj_thr, keep_top_k)
        conf_in = odm_conf.clone()
        boxes, scores = det.forward(arm_loc, arm_conf, odm_loc, conf_in, priors)
        for b in range(B):
            for c in range(1, C):
                s = scores[b, :, c]
                s = s[s > conf_thr]
                assert s.unique().numel() == s.numel(), 'tied candidate scores'
        # a4: eval_refinedet_coco.py:205-232 glue, reference py_cpu_nms
        scale = torch.tensor([320., 320., 320., 320.])
        a4_counts = np.zeros((B, C), np.int32)
        a4_dets = np.zeros((B, C, keep_top_k, 5), np.float32)
        for b in range(B):
            bx = (boxes[b] * scale).numpy()
            sc = scores[b].numpy()
            for j in range(1, C):
                inds = np.where(sc[:, j] > conf_thr)[0]
                if len(inds) == 0:
                    continue
                c_bboxes, c_scores = bx[inds], sc[inds, j]
                order = c_scores.argsort()[::-1][:top_k]
                c_dets = np.hstack((c_bboxes[order], c_scores[order][:, None])).astype(np.float32, copy=False)
                keep = R.py_cpu_nms(c_dets, nms_thr)
                c_dets = c_dets[keep, :][:keep_top_k, :]
                a4_counts[b, j] = c_dets.shape[0]
                a4_dets[b, j, :c_dets.shape[0]] = c_dets
        # a5
        det5 = R.Detect(C, 320, 0, top_k, conf_thr, nms_thr, obj_thr, keep_top_k)
        out5 = det5.forward_python_nms(arm_loc, arm_conf, odm_loc, odm_conf.clone(), priors)
        save('detect_%s.npz' % tag, priors=priors.numpy(), arm_loc=arm_loc.numpy(),
             arm_conf=arm_conf.numpy(), odm_loc=odm_loc.numpy(), odm_conf=odm_conf.numpy(),
             params=np.array([C, top_k, keep_top_k, conf_thr, nms_thr, obj_thr], np.float64),
             boxes=boxes.numpy(), scores=scores.numpy(), conf_after=conf_in.numpy(),
             scale=scale.numpy(), a4_counts=a4_counts, a4_dets=a4_dets, a5_output=out5.numpy())

    # ---- refine_match / loss (a9-a11) -----------------------------------------
    B, C = 3, 6
    priors = pri['320'][::5].contiguous()
    P = priors.shape[0]
    targets = gen_targets(g, B, 9, C, 0.05, 0.4)
    # image 2: two identical GTs with different labels (forced-match "last j wins")
    targets[2][4, :4] = targets[2][1, :4]
    targets[2][1, 4], targets[2][4, 4] = 3.0, 5.0
    arm_loc = 0.1 * torch.randn(B, P, 4, generator=g)
    odm_loc = 0.1 * torch.randn(B, P, 4, generator=g)
    arm_conf = torch.randn(B, P, 2, generator=g)
    odm_conf = torch.randn(B, P, C, generator=g)
    rm = {}
    for mode in ('arm', 'odm'):
        loc_t = torch.zeros(B, P, 4)
        conf_t = torch.zeros(B, P, dtype=torch.long)
        for idx in range(B):
            truths, labels = targets[idx][:, :-1], targets[idx][:, -1]
            if mode == 'arm':
                bu.refine_match(0.5, truths, priors, var, labels >= 0, loc_t, conf_t, idx)
            else:
                bu.refine_match(0.5, truths, priors, var, labels, loc_t, conf_t, idx, arm_loc[idx])
        rm['loc_t_' + mode], rm['conf_t_' + mode] = loc_t.numpy(), conf_t.numpy()
    # SSD match() (labels + 1)
    loc_t = torch.zeros(B, P, 4)
    conf_t = torch.zeros(B, P, dtype=torch.long)
    for idx in range(B):
        bu.match(0.5, targets[idx][:, :-1], priors, var, targets[idx][:, -1] - 1, loc_t, conf_t, idx)
    rm['loc_t_ssd'], rm['conf_t_ssd'] = loc_t.numpy(), conf_t.numpy()

    preds = (arm_loc, arm_conf, odm_loc, odm_conf, priors)
    arm_crit = R.Loss(2, 0.5, True, 0, True, 3, 0.5, False, False)
    odm_crit = R.Loss(C, 0.5, True, 0, True, 3, 0.5, False, False, use_ARM=True)
    al, ac = arm_crit(preds, targets)
    ol, oc = odm_crit(preds, targets)
    # masks, restating refinedet_multibox_loss.py:96-123 on the reference's own targets
    masks = {}
    for mode, conf_data, nc in (('arm', arm_conf, 2), ('odm', odm_conf, C)):
        conf_t = torch.from_numpy(rm['conf_t_' + mode])
        pos = conf_t > 0
        if mode == 'odm':
            Pm = torch.softmax(arm_conf, 2)[:, :, 1]
            pos[(Pm <= 0.01)] = 0
        bc = conf_data.reshape(-1, nc)
        loss_c = bu.log_sum_exp(bc) - bc.gather(1, conf_t.view(-1, 1))
        loss_c[pos.view(-1, 1)] = 0
        loss_c = loss_c.view(B, -1)
        for b in range(B):
            assert loss_c[b][~pos[b]].unique().numel() == int((~pos[b]).sum()), 'tied losses'
        _, loss_idx = loss_c.sort(1, descending=True)
        _, idx_rank = loss_idx.sort(1)
        num_pos = pos.long().sum(1, keepdim=True)
        num_neg = torch.clamp(3 * num_pos, max=pos.size(1) - 1)
        neg = idx_rank < num_neg.expand_as(idx_rank)
        masks['pos_' + mode], masks['neg_' + mode] = pos.numpy(), neg.numpy()
        masks['loss_c_rows_' + mode] = loss_c.numpy()
    # ODM criterion when every positive is ARM-filtered -> zeros(1)
    arm_conf_off = arm_conf.clone()
    arm_conf_off[..., 0] += 50.0
    zl, zc = odm_crit((arm_loc, arm_conf_off, odm_loc, odm_conf, priors), targets)
    assert zl.shape == (1,) and float(zl) == 0.0 and float(zc) == 0.0
    save('match_loss.npz', priors=priors.numpy(), targets=torch.stack(targets).numpy(),
         arm_loc=arm_loc.numpy(), arm_conf=arm_conf.numpy(), odm_loc=odm_loc.numpy(),
         odm_conf=odm_conf.numpy(), arm_loss=np.array([al.item(), ac.item()], np.float64),
         odm_loss=np.array([ol.item(), oc.item()], np.float64), **rm, **masks)
    print('golden fixtures written to', HERE)
    for f in sorted(os.listdir(HERE)):
        if f.endswith('.npz'):
            print('  %-24s %7d B' % (f, os.path.getsize(os.path.join(HERE, f))))


if __name__ == '__main__':
    main()
