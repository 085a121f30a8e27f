#!/usr/bin/env python
"""Device time of the two streaming kernels of the ODM criterion at BASELINE.json config 4 (B=32, P=16320, C=81):
`rd_conf_loss` (reads conf once) and `rd_multibox_loss_backward` (writes grad_conf once), CUDA events, a 512 MiB
memset before every call (L2 flush; it also hides the host's launch latency), inputs alternated between two copies.
The variant of the backward kernel is chosen by the library's env switches (RD_BWD=tile|regs, RD_BWD_PER_SM=n).

    python tools/bench_loss_kernels.py [C] [n]
"""
import json
import os
import sys
from statistics import median

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import refinedet.pytorch_b200 as rd  # noqa: E402
from refinedet.pytorch_b200 import _ffi, synthetic  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 81
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
B, P, G = 32, 16320, 50
dev = torch.device('cuda', 0)
bu = rd.box_utils
priors = rd.PriorBox(rd.REFINEDET_ANCHORS['512']).forward().to(dev)
tp = [t.to(dev) for t in synthetic.train_predictions(40, B, P, C)]
tg = [t.to(dev) for t in synthetic.targets(41, B, G, C)]
truths, labels, cnt = bu.pad_targets(tg, dev)
mode = bu.LABEL_ODM if C > 2 else bu.LABEL_ARM_BINARY
lt, ct = bu.match_batch(0.5, truths, labels, cnt, priors, [0.1, 0.2], tp[0] if C > 2 else None, mode)
conf = tp[3] if C > 2 else tp[1]
loc = tp[2] if C > 2 else tp[0]
ce, lse, pos = bu.conf_loss(conf, ct, tp[1] if C > 2 else None, 0.01)
neg, npos = bu.hnm_select(ce, pos, 3)
one, nn = torch.ones((), device=dev), pos.sum().float()
confs = [conf, conf.clone()]
gconfs = [torch.empty_like(conf) for _ in range(2)]
gloc = torch.empty_like(loc)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
peak = 6550.7
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    pass
i = [0]


def f_conf():
    i[0] += 1
    bu.conf_loss(confs[i[0] & 1], ct, tp[1] if C > 2 else None, 0.01)


def f_bwd():
    i[0] += 1
    _ffi.check(_ffi.lib().rd_multibox_loss_backward(
        _ffi.ptr(loc), _ffi.ptr(lt), _ffi.ptr(confs[i[0] & 1]), _ffi.ptr(ct), _ffi.ptr(lse), _ffi.ptr(pos), _ffi.ptr(neg),
        _ffi.ptr(one), _ffi.ptr(one), _ffi.ptr(nn), B * P, C, _ffi.ptr(gloc), _ffi.ptr(gconfs[i[0] & 1]),
        _ffi.stream_ptr()), 'rd_multibox_loss_backward')


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in evs:
        flush.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)
    return median(ms), ms[0]


out = {'C': C, 'variant': os.environ.get('RD_BWD', 'zero-stream'), 'per_sm': os.environ.get('RD_BWD_PER_SM')}
for name, fn, byts in (('conf_loss', f_conf, B * P * (4 * C + 17)), ('loss_backward', f_bwd, B * P * (4 * C + 18))):
    med, best = timed(fn)
    out[name] = {'ms': round(med, 5), 'min_ms': round(best, 5), 'frac': round(byts / (med * 1e-3) / 1e9 / peak, 4)}
# the gradient of the zero-stream kernel against the register kernel (bit-identical arithmetic)
f_bwd()
torch.cuda.synchronize()
out['grad_conf_nonzero_rows'] = int((gconfs[i[0] & 1].abs().sum(-1) > 0).sum())
out['selected_rows'] = int((pos | neg).sum())
print(json.dumps(out))
