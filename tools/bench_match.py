#!/usr/bin/env python
"""Secondary measurement (BASELINE.json config 4/5): training-side matching and
hard-negative selection on one B200 — device time with CUDA events, L2 flushed between iterations,
against the algorithmic bytes (refine_match 40*P + 20*G per image, HNM 6*P per image).  (The CPU side of
this path is timed by bench.py's cpu_baseline leg only: oracle/ is test infrastructure.)

    python tools/bench_match.py [--G 50] [--classes 81] [--steps 30]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--G', type=int, default=50)
    ap.add_argument('--classes', type=int, default=81)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--batch', type=int, default=32)
    args = ap.parse_args()
    import refinedet.pytorch_b200 as rd
    from refinedet.pytorch_b200 import synthetic
    B, C, G, P = args.batch, args.classes, args.G, 16320
    dev = torch.device('cuda', 0)
    priors = rd.PriorBox(rd.REFINEDET_ANCHORS['512']).forward().to(dev)
    small = C == 2
    tg = synthetic.targets(5234, B, G, C, 0.01 if small else 0.02, 0.06 if small else 0.17)
    arm_loc, arm_conf, odm_loc, odm_conf = [t.to(dev) for t in synthetic.train_predictions(5235, B, P, C)]
    bu = rd.box_utils
    truths, labels, cnt = bu.pad_targets([t.to(dev) for t in tg], dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(
        os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0

    def timed(fn, flush_l2=True):
        """The memset before every call flushes the L2 AND keeps the GPU busy while the host prepares the launch, so the
        event pair brackets device time only (without it the same kernels read 12 - 14 us longer: launch latency)."""
        for _ in range(3):
            fn()
        ms = []
        for _ in range(args.steps):
            if flush_l2:
                flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); torch.cuda.synchronize()
            ms.append(s.elapsed_time(e))
        return float(np.median(ms))

    out = {'config': {'B': B, 'P': P, 'C': C, 'G': G}}
    ms = timed(lambda: bu.match_batch(0.5, truths, labels, cnt, priors, [0.1, 0.2], arm_loc, bu.LABEL_ODM))
    byts = B * (40 * P + 20 * G)
    out['refine_match_odm'] = {'ms': ms, 'images_per_s': B / ms * 1e3, 'algorithmic_GBs': byts / ms / 1e6,
                               'frac_of_hbm_peak': byts / ms / 1e6 / peak}
    loss = torch.rand(B, P, device=dev) * 8
    pos = torch.rand(B, P, device=dev) < 0.015
    ms = timed(lambda: bu.hnm_select(loss, pos, 3))
    byts = B * 6 * P
    out['hnm_select'] = {'ms': ms, 'images_per_s': B / ms * 1e3, 'algorithmic_GBs': byts / ms / 1e6,
                         'frac_of_hbm_peak': byts / ms / 1e6 / peak}
    crit = rd.RefineDetMultiBoxLoss(C, 0.5, True, 0, True, 3, 0.5, False, True, use_ARM=True)
    preds = (arm_loc, arm_conf, odm_loc, odm_conf, priors)
    tgd = [t.to(dev) for t in tg]
    ms = timed(lambda: crit(preds, tgd))
    out['odm_criterion_forward'] = {'ms': ms, 'images_per_s': B / ms * 1e3}
    # the loss tail kernel by kernel (f-4): forward conf loss (reads conf once), reduce, backward
    loc_t, conf_t = crit.match_targets(preds, tgd)
    conf_rot = [odm_conf, odm_conf.clone()]
    it = {'i': 0}

    def conf_loss_rot():
        it['i'] += 1
        return bu.conf_loss(conf_rot[it['i'] & 1], conf_t, arm_conf, 0.01)
    ms = timed(conf_loss_rot)
    byts = B * P * (4 * C + 8 + 9)                       # conf row + conf_t in; ce, lse, pos out
    out['conf_loss'] = {'ms': ms, 'algorithmic_GBs': byts / ms / 1e6, 'frac_of_hbm_peak': byts / ms / 1e6 / peak}
    ce, lse, pos_k = bu.conf_loss(odm_conf, conf_t, arm_conf, 0.01)
    neg_k, num_pos = bu.hnm_select(ce, pos_k, 3)
    ms = timed(lambda: bu.multibox_loss_reduce(odm_loc, loc_t, ce, pos_k, neg_k, num_pos))
    out['loss_reduce'] = {'ms': ms}
    one = torch.ones((), device=dev)
    n_dev = pos_k.sum().float()
    def backward_rot():
        it['i'] += 1
        return bu.multibox_loss_backward(odm_loc, loc_t, conf_rot[it['i'] & 1], conf_t, lse, pos_k, neg_k, one, one, n_dev)
    ms = timed(backward_rot)
    byts = B * P * (4 * C + 16 + 2)                      # grad_conf + grad_loc written, masks read
    out['loss_backward'] = {'ms': ms, 'algorithmic_GBs': byts / ms / 1e6, 'frac_of_hbm_peak': byts / ms / 1e6 / peak}
    p_loc = odm_loc.clone().requires_grad_(True)
    p_conf = odm_conf.clone().requires_grad_(True)
    preds_g = (arm_loc, arm_conf, p_loc, p_conf, priors)

    def fwd_bwd():
        l, c = crit(preds_g, tgd)
        (l + c).backward()
        p_loc.grad = None
        p_conf.grad = None
    ms = timed(fwd_bwd)
    out['odm_criterion_forward_backward'] = {'ms': ms, 'images_per_s': B / ms * 1e3}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
