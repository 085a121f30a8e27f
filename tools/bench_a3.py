#!/usr/bin/env python
"""Device time of a3 (`rd_detect_forward`: Detect_RefineDet.forward with the in-place ARM zeroing) at BASELINE.json
config 3 (B=32, P=16320, C=81), CUDA events, 512 MiB memset before every call, a fresh copy of odm_conf per call (the
in-place zeroing is real work only the first time).

    python tools/bench_a3.py [sparse|dense] [n]
"""
import json
import os
import sys
from statistics import median

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import refinedet.pytorch_b200 as rd  # noqa: E402
from refinedet.pytorch_b200 import synthetic  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else 'sparse'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
B, P, C = 32, 16320, 81
dev = torch.device('cuda', 0)
priors = rd.PriorBox(rd.REFINEDET_ANCHORS['512']).forward().to(dev)
arm_loc, arm_conf, odm_loc, odm_conf = [t.to(dev) for t in synthetic.detect_inputs(7, B, P, C, kind)]
det = rd.Detect_RefineDet(C, 512, 0, 1000, 0.01, 0.45, 0.01, 500)
confs = [odm_conf.clone() for _ in range(n + 2)]
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for c in confs[:2]:
    det.forward(arm_loc, arm_conf, odm_loc, c, priors)
torch.cuda.synchronize()
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
for (a, b), c in zip(evs, confs[2:]):
    flush.zero_()
    a.record()
    det.forward(arm_loc, arm_conf, odm_loc, c, priors)
    b.record()
torch.cuda.synchronize()
ms = sorted(a.elapsed_time(b) for a, b in evs)
byts = B * (4 * P * (4 + 2 + 4 + C) + 4 * P * (4 + C))
peak = 6550.7
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    pass
print(json.dumps({'kind': kind, 'ms': round(median(ms), 5), 'min_ms': round(ms[0], 5),
                  'frac': round(byts / (median(ms) * 1e-3) / 1e9 / peak, 4)}))
