"""One rd_select_topk call at the config-3 shape between cudaProfilerStart/Stop (ncu --profile-from-start off):
    python tools/select_once.py [sparse|dense]"""
import sys, os, torch
sys.path.insert(0, os.getcwd())
import refinedet.pytorch_b200 as rd
from refinedet.pytorch_b200 import synthetic
B,P,C=32,16320,81
gen = sys.argv[1] if len(sys.argv) > 1 else 'sparse'
priors = rd.PriorBox(rd.REFINEDET_ANCHORS['512']).forward().cuda()
a=[t.cuda() for t in synthetic.detect_inputs(4321,B,P,C,gen)]
det = rd.Detect_RefineDet(C,512,0,1000,0.01,0.45,0.01,500)
_, scores = det.forward(*a, priors)
for i in range(2): rd.box_utils.select_topk(scores, 0.01, 1000)
flush = torch.empty(512 << 20, dtype=torch.uint8, device='cuda'); flush.zero_()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
rd.box_utils.select_topk(scores, 0.01, 1000)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
