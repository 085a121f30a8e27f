#!/usr/bin/env python
"""How well does the detection exchange pipeline with the stage?  (torchrun, one process per GPU)

For L lanes in {1, 2, 4, 8}: microseconds per step of (a) the stage alone, (b) the exchange alone, (c) stage + exchange
as one captured plan per lane, each launched round-robin over the lanes, CUDA events, max over ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 \
        tools/exchange_pipeline.py [p2p|multicast|auto]
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import refinedet.pytorch_b200 as rd  # noqa: E402
from refinedet.pytorch_b200 import dist as rdist, synthetic  # noqa: E402
from refinedet.pytorch_b200.layers.functions.detection_refinedet import DetectPlan  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else 'auto'
    ctas = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    B, P, C = 32, 16320, 81
    priors = rd.PriorBox(rd.REFINEDET_ANCHORS['512']).forward().to(dev)
    det = rd.Detect_RefineDet(C, 512, 0, 1000, 0.01, 0.45, 0.01, 500)
    sets = [[t.to(dev) for t in synthetic.detect_inputs(4234 + rank + 100 * i, B, P, C, 'sparse')] for i in range(4)]
    scale = torch.tensor([512.0] * 4, device=dev).reshape(1, 4).expand(B, 4).contiguous()
    out = {'world': world, 'mode': mode, 'copy_ctas': ctas, 'lanes': {}}
    main_st = torch.cuda.current_stream(dev)
    for L in (1, 4):
        streams = [torch.cuda.Stream(dev) for _ in range(L)]
        lanes = [(det.new_workspace(B, P, dev), det.new_outputs(B, dev)) for _ in range(L)]
        exs = [rdist.PeerExchange(B, C, lanes[l][1].dets.shape[2], dev, mode=None if mode == 'auto' else mode, copy_ctas=ctas) for l in range(L)]
        stage = [[det.plan(*a, priors, scale=scale, workspace=lanes[l][0], out=lanes[l][1]) for a in sets] for l in range(L)]
        both = [[det.plan(*a, priors, scale=scale, workspace=lanes[l][0], out=lanes[l][1], then=exs[l].exchange) for a in sets]
                for l in range(L)]
        for l in range(L):
            stage[l][0].launch(streams[l])
        torch.cuda.synchronize()
        xonly = [DetectPlan.capture(dev, (lambda l=l: exs[l].exchange(lanes[l][1])), lanes[l][1]) for l in range(L)]

        host_us = []

        def timed(fn, K=200):
            for i in range(2 * L):
                fn(i)
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(main_st)
            for st in streams:
                st.wait_event(e0)
            import time
            h0 = time.perf_counter()
            for i in range(K):
                fn(i)
            host_us.append(round((time.perf_counter() - h0) / K * 1e6, 2))
            for st in streams:
                main_st.wait_stream(st)
            e1.record(main_st)
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / K * 1e3], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return round(float(t), 2)
        rec = {'stage_us': timed(lambda i: stage[i % L][i % 4].launch(streams[i % L])),
               'exchange_us': timed(lambda i: xonly[i % L].launch(streams[i % L])),
               'stage_exchange_us': timed(lambda i: both[i % L][i % 4].launch(streams[i % L])),
               'exchange_mode': exs[0].mode}
        rec['host_issue_us'] = list(host_us)
        out['lanes'][L] = rec
        del stage, both, xonly, exs, lanes
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
