"""One GPU: does a trailing exchange (world = 1: pack + completion protocol, no NVLink) or a plain rd_pack_detections hide
behind the stage of the other lanes?  us per step of stage / stage + local exchange / stage + pack at 1 and 4 lanes.
(Result, round 2: +8.4 / +5.5 us per step at 4 lanes -- every extra kernel node costs ~2.7 us of pipelined throughput.)"""
import sys, os, ctypes, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import refinedet.pytorch_b200 as rd
from refinedet.pytorch_b200 import synthetic
from refinedet.pytorch_b200._ffi import check, lib, ptr, stream_ptr
from refinedet.pytorch_b200.layers.functions.detection_refinedet import DetectPlan
dev = torch.device('cuda', 0)
B, P, C = 32, 16320, 81
priors = rd.PriorBox(rd.REFINEDET_ANCHORS['512']).forward().to(dev)
det = rd.Detect_RefineDet(C, 512, 0, 1000, 0.01, 0.45, 0.01, 500)
sets = [[t.to(dev) for t in synthetic.detect_inputs(4234 + 100 * i, B, P, C, 'sparse')] for i in range(4)]
scale = torch.tensor([512.0] * 4, device=dev).reshape(1, 4).expand(B, 4).contiguous()
main_st = torch.cuda.current_stream(dev)
out = {}
for L in (1, 4):
    streams = [torch.cuda.Stream(dev) for _ in range(L)]
    lanes = [(det.new_workspace(B, P, dev), det.new_outputs(B, dev)) for _ in range(L)]
    cap = B * C * 500
    slot = int(lib().rd_exchange_slot_bytes(B, C, cap)); ctrl = int(lib().rd_exchange_ctrl_bytes())
    bufs = [torch.zeros(ctrl + 2 * slot, dtype=torch.uint8, device=dev) for _ in range(L)]
    bases = [(ctypes.c_void_p * 1)(b.data_ptr()) for b in bufs]
    def xround(l):
        def f(res):
            check(lib().rd_exchange_round(ptr(res.counts), ptr(res.dets), B, C, 500, bases[l], None, 1, 0, B, cap, 0, 2000, stream_ptr()), 'x')
        return f
    stage = [[det.plan(*a, priors, scale=scale, workspace=lanes[l][0], out=lanes[l][1]) for a in sets] for l in range(L)]
    both = [[det.plan(*a, priors, scale=scale, workspace=lanes[l][0], out=lanes[l][1], then=xround(l)) for a in sets] for l in range(L)]
    # pack only through rd_pack_detections (two kernels, no PDL)
    offs = [torch.empty(B*C+1, dtype=torch.int32, device=dev) for _ in range(L)]
    rows = [torch.empty(cap, 5, device=dev) for _ in range(L)]
    def packer(l):
        def f(res):
            check(lib().rd_pack_detections(ptr(res.counts), ptr(res.dets), B, C, 500, ptr(offs[l]), ptr(rows[l]), cap, stream_ptr()), 'p')
        return f
    both2 = [[det.plan(*a, priors, scale=scale, workspace=lanes[l][0], out=lanes[l][1], then=packer(l)) for a in sets] for l in range(L)]
    def timed(fn, K=200):
        for i in range(2 * L): fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main_st)
        for st in streams: st.wait_event(e0)
        for i in range(K): fn(i)
        for st in streams: main_st.wait_stream(st)
        e1.record(main_st); torch.cuda.synchronize()
        return round(e0.elapsed_time(e1) / K * 1e3, 2)
    out[L] = {'stage': timed(lambda i: stage[i % L][i % 4].launch(streams[i % L])),
              'stage+local_exchange(world=1)': timed(lambda i: both[i % L][i % 4].launch(streams[i % L])),
              'stage+pack_detections': timed(lambda i: both2[i % L][i % 4].launch(streams[i % L]))}
print(json.dumps(out))
