#!/usr/bin/env python
"""One pass of the fused detect stage (config 3 shape) between cudaProfilerStart/Stop: the script behind the ncu
captures of the stage on the generators other than the headline one.

    ncu --set full --clock-control none --import-source on --profile-from-start off \
        -k regex:'collect_kernel|graph_kernel|nms_small_kernel|nms_large_kernel' -o prof \
        python tools/stage_once.py dense | clustered-12 | -7.0 (sparse generator with that ARM logit shift)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import refinedet.pytorch_b200 as rd  # noqa: E402
from refinedet.pytorch_b200 import synthetic  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else 'dense'
B, P, C = 32, 16320, 81
priors = rd.PriorBox(rd.REFINEDET_ANCHORS['512']).forward().cuda()
scale = torch.tensor([512.] * 4).cuda().reshape(1, 4).expand(B, 4).contiguous()
if what == 'dense':
    a = [t.cuda() for t in synthetic.detect_inputs(77, B, P, C, 'dense')]
elif what.startswith('clustered'):
    a = [t.cuda() for t in synthetic.detect_inputs_clustered(77, B, priors.cpu(), C, n_obj=int(what.split('-')[1]))]
else:
    a = [t.cuda() for t in synthetic.detect_inputs(77, B, P, C, 'sparse', arm_shift=float(what))]
det = rd.Detect_RefineDet(C, 512, 0, 1000, 0.01, 0.45, 0.01, 500)
for i in range(2):
    det.detect(*a, priors, scale=scale)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
res = det.detect(*a, priors, scale=scale)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print(what, 'nodes/img', int((a[1][..., 1] > 0.01).sum()) // B, 'kept', int(res.counts.sum()))
