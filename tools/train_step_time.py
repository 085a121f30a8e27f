#!/usr/bin/env python
"""Wall-clock and device time of one RefineDet training-criterion step at BASELINE.json config 4 (B=32, P=16320,
C=81, 50 ground-truth boxes per image): ARM + ODM criteria forward, `loss.backward()`, losses read — as the
reference's two modules called one after the other (train_refinedet.py:252-261) and as one RefineDetCriterionPair.

    python tools/train_step_time.py [n] [--profile]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import refinedet.pytorch_b200 as rd  # noqa: E402
from refinedet.pytorch_b200 import synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 30
B, P, C, G = 32, 16320, 81, 50
dev = torch.device('cuda', 0)
priors = rd.PriorBox(rd.REFINEDET_ANCHORS['512']).forward().to(dev)
tp = [t.to(dev) for t in synthetic.train_predictions(40, B, P, C)]
tg = [t.to(dev) for t in synthetic.targets(41, B, G, C)]
leaves = [t.clone().requires_grad_(True) for t in tp]
preds = tuple(leaves) + (priors,)


def make(kind, sync_free, concurrent=True, read=True):
    a = rd.RefineDetMultiBoxLoss(2, 0.5, True, 0, True, 3, 0.5, False, True, sync_free=sync_free)
    o = rd.RefineDetMultiBoxLoss(C, 0.5, True, 0, True, 3, 0.5, False, True, use_ARM=True, sync_free=sync_free)
    pair = rd.RefineDetCriterionPair(a, o, concurrent=concurrent)

    def step():
        rd.box_utils.clear_pad_cache()
        if kind == 'pair':
            al, ac, ol, oc = pair(preds, tg)
        else:
            al, ac = a(preds, tg)
            ol, oc = o(preds, tg)
        (al + ac + ol + oc).backward()
        vals = torch.stack([al.detach().reshape(()), ac.detach().reshape(()), ol.detach().reshape(()),
                            oc.detach().reshape(())]).tolist() if (sync_free and read) else None
        for t in leaves:
            t.grad = None
        return vals
    return step


def wall(fn):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / n


def device_ms(fn, k=6):
    """Device time per step when the host is AHEAD of the GPU, as it is inside a training loop (the criterion is queued
    while the network's forward pass is still running): a spin kernel holds the stream while k steps are queued."""
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()
    torch.cuda.synchronize()
    torch.cuda._sleep(int(12e-3 * 1.9e9))                 # ~12 ms: k steps take the host 3-4 ms to queue
    a.record()
    for _ in range(k):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / k


out = {}
for name, kind, sf, conc in (('two_call', 'two', False, True), ('two_call_sync_free', 'two', True, True),
                             ('pair', 'pair', False, True), ('pair_sync_free', 'pair', True, True),
                             ('pair_sync_free_one_stream', 'pair', True, False)):
    out[name] = round(wall(make(kind, sf, conc)), 4)
out['device_two_call_sync_free'] = round(device_ms(make('two', True, read=False)), 4)
out['device_pair_sync_free'] = round(device_ms(make('pair', True, read=False)), 4)
out['device_pair_one_stream'] = round(device_ms(make('pair', True, False, read=False)), 4)
print(json.dumps(out))
if '--profile' in sys.argv:
    import cProfile
    import pstats
    step = make('pair', True)
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(50):
        step()
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats('cumulative').print_stats(35)
