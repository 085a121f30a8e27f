#!/usr/bin/env python
"""One ODM-criterion forward + backward (BASELINE.json config 4 shape) between cudaProfilerStart/Stop: the script
behind the ncu captures of the training-side kernels.

    ncu --set full --clock-control none --import-source on --profile-from-start off \
        -k regex:'match_pass|hnm_|conf_loss|loss_reduce|loss_final|loss_backward' -o prof python tools/train_once.py [G] [C]
"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import refinedet.pytorch_b200 as rd
from refinedet.pytorch_b200 import synthetic
G=int(sys.argv[1]) if len(sys.argv)>1 else 50
C=int(sys.argv[2]) if len(sys.argv)>2 else 81
B,P=32,16320
dev=torch.device('cuda',0)
priors = rd.PriorBox(rd.REFINEDET_ANCHORS['512']).forward().to(dev)
small = C==2
tg=[t.to(dev) for t in synthetic.targets(5234,B,G,C,0.01 if small else 0.02,0.06 if small else 0.17)]
arm_loc,arm_conf,odm_loc,odm_conf=[t.to(dev) for t in synthetic.train_predictions(5235,B,P,C)]
crit = rd.RefineDetMultiBoxLoss(C,0.5,True,0,True,3,0.5,False,True,use_ARM=True)
p_loc=odm_loc.clone().requires_grad_(True); p_conf=odm_conf.clone().requires_grad_(True)
preds=(arm_loc,arm_conf,p_loc,p_conf,priors)
def step():
    l,c=crit(preds,tg); (l+c).backward(); p_loc.grad=None; p_conf.grad=None
for _ in range(2): step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
