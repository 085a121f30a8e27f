#!/usr/bin/env python
"""Node-density sweep of the fused detect stage (DESIGN.md "Measured"): the sparse generator with the ARM logit
shift moved from -8 (4.4 % of the anchors pass, ~700 nodes per image) towards -6 (~3900 nodes), one batch of
32 images, L2 flushed, with the per-kernel breakdown of a serialised pass.

    python tools/density_sweep.py
"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import refinedet.pytorch_b200 as rd
from refinedet.pytorch_b200 import synthetic
B,P,C=32,16320,81
priors = rd.PriorBox(rd.REFINEDET_ANCHORS['512']).forward().cuda()
scale=torch.tensor([512.]*4).cuda().reshape(1,4).expand(B,4).contiguous()
flush = torch.empty(512<<20, dtype=torch.uint8, device='cuda')
det = rd.Detect_RefineDet(C,512,0,1000,0.01,0.45,0.01,500)
for shift in (-8.0,-7.5,-7.0,-6.5,-6.0,'clustered-6','clustered-12','clustered-24'):
    if isinstance(shift, str):   # trained-detector-like: n objects per image, dozens of overlapping boxes each
        a=[t.cuda() for t in synthetic.detect_inputs_clustered(77,B,priors.cpu(),C,n_obj=int(shift.split('-')[1]))]
    else:
        a=[t.cuda() for t in synthetic.detect_inputs(77,B,P,C,'sparse',arm_shift=shift)]
    nodes=int((a[1][...,1]>0.01).sum())//B
    for i in range(3): res=det.detect(*a,priors,scale=scale)
    torch.cuda.synchronize(); ms=[]
    for i in range(10):
        flush.zero_()
        s,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        s.record(); res=det.detect(*a,priors,scale=scale); e.record(); torch.cuda.synchronize(); ms.append(s.elapsed_time(e))
    ms.sort()
    print('arm_shift %s: %d nodes/img, median %.3f ms, kept %d, prof %s' % (shift,nodes,ms[len(ms)//2],int(res.counts.sum()), {k:round(v,4) for k,v in det.profile_stage([a],priors,scale,flush,steps=5).items()}))
