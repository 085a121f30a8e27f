#!/usr/bin/env python
"""Secondary measurement: ``rd_select_topk`` on one B200 at the config-3 and config-2 shapes —
device time with CUDA events, L2 flushed between iterations, against the algorithmic bytes
(read ``4*P*C`` per image once; write 8 B per selected candidate + 4 B per (image, class)).

    python tools/bench_select.py [--steps 30]
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
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--batch', type=int, default=32)
    args = ap.parse_args()
    import refinedet.pytorch_b200 as rd
    from refinedet.pytorch_b200 import synthetic
    dev = torch.device('cuda', 0)
    B = args.batch
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(
        os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0
    out = {}
    for name, size, C, gen in (('cfg3_sparse', '512', 81, 'sparse'), ('cfg3_dense', '512', 81, 'dense'),
                               ('cfg2_sparse', '320', 21, 'sparse'), ('cfg5_dense', '512', 2, 'dense')):
        priors = rd.PriorBox(rd.REFINEDET_ANCHORS[size]).forward().to(dev)
        P = priors.shape[0]
        arm_loc, arm_conf, odm_loc, odm_conf = [t.to(dev) for t in synthetic.detect_inputs(4321, B, P, C, gen)]
        det = rd.Detect_RefineDet(C, int(size), 0, 1000, 0.01, 0.45, 0.01, 500)
        _, scores = det.forward(arm_loc, arm_conf, odm_loc, odm_conf, priors)      # ARM-filtered rows zeroed
        fn = lambda: rd.box_utils.select_topk(scores, 0.01, 1000)                  # noqa: E731
        for _ in range(3):
            idx, sc, counts = fn()
        ms = []
        for _ in range(args.steps):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); torch.cuda.synchronize()
            ms.append(s.elapsed_time(e))
        ms = float(np.median(ms))
        sel = int(counts.sum())
        byts = B * 4 * P * C + 8 * sel + 4 * B * C
        out[name] = {'B': B, 'P': P, 'C': C, 'ms': ms, 'selected': sel, 'images_per_s': B / ms * 1e3,
                     'algorithmic_GBs': byts / ms / 1e6, 'frac_of_hbm_peak': byts / ms / 1e6 / peak}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
