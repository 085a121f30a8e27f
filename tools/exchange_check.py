#!/usr/bin/env python
"""Multi-GPU check + timing of the detection gather, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/exchange_check.py

Every rank runs the config-3 detect stage on its own batch, then gathers the compact detections twice —
through NCCL (``dist.gather_packed``: header all_gather + padded all_gather_into_tensor) and through
``dist.PeerExchange`` (``rd_pack_scatter``: the pack kernel stores into every peer's buffer over NVLink) —
checks that both deliver identical counts and rows for every rank, and prints their times (max over ranks).
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import refinedet.pytorch_b200 as rd  # noqa: E402
from refinedet.pytorch_b200 import dist as rdist, synthetic  # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    B, P, C = 32, 16320, 81
    priors = rd.PriorBox(rd.REFINEDET_ANCHORS['512']).forward().to(dev)
    det = rd.Detect_RefineDet(C, 512, 0, 1000, 0.01, 0.45, 0.01, 500)
    a = [t.to(dev) for t in synthetic.detect_inputs(4234 + rank, B, P, C, 'sparse')]
    scale = torch.tensor([512.0] * 4, device=dev).reshape(1, 4).expand(B, 4).contiguous()
    res = det.detect(*a, priors, scale=scale)
    torch.cuda.synchronize()

    def timed(fn, n=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        dt = torch.tensor([(time.perf_counter() - t0) / n], device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return float(dt) * 1e3

    def timed_dev(fn, n=30):
        """device time per call (CUDA events on the current stream), max over ranks"""
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        dt = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return float(dt)

    counts_n, rows_n = rdist.gather_detections(res)
    ms_nccl = timed(lambda: rdist.gather_detections(res))
    out = {'world': world, 'rows_per_rank': int(rows_n[rank].shape[0]), 'nccl_gather_ms': ms_nccl, 'sweep': []}
    ok_all = True
    try:
        for mode in ('p2p', 'multicast'):
            for ctas in ((1, 2, 4, 8, 16, -24, -48) if mode == 'p2p' else (24, 48, 96, 148)):     # p2p: n > 0 TMA CTAs per peer, n < 0 LSU CTAs
                try:
                    ex = rdist.PeerExchange(B, C, res.dets.shape[2], dev, mode=mode, copy_ctas=ctas)
                except RuntimeError as e:
                    out['sweep'].append({'mode': mode, 'error': repr(e)[:120]})
                    break
                ex.exchange(res)
                counts_p, rows_p = ex.result()
                ok = all(torch.equal(counts_p[r], counts_n[r]) and torch.equal(rows_p[r], rows_n[r]) for r in range(world))
                ok_all = ok_all and ok
                rec = {'mode': ex.mode, 'copy_ctas': ctas, 'equals_nccl': bool(ok),
                       'device_ms': timed_dev(lambda: ex.exchange(res)),
                       'with_host_read_ms': timed(lambda: (ex.exchange(res), ex.result()))}
                if not out['sweep'] or 'barriers_only_device_ms' not in out:
                    out['barriers_only_device_ms'] = timed_dev(lambda: (ex.hdl.barrier(channel=0), ex.hdl.barrier(channel=1)))
                out['sweep'].append(rec)
                del ex
        out['peer_equals_nccl'] = bool(ok_all)
        good = [r for r in out['sweep'] if 'device_ms' in r]
        if good:
            best = min(good, key=lambda r: r['device_ms'])
            out['best'] = best
            bytes_out = out['rows_per_rank'] * 20 * (world - 1)
            out['egress_GBs_unicast_equiv'] = bytes_out / (best['device_ms'] * 1e-3) / 1e9
    except Exception as e:  # symmetric memory not available on this box
        out['peer_exchange_error'] = repr(e)[:300]
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()
    return 0 if out.get('peer_equals_nccl', True) else 1


if __name__ == '__main__':
    sys.exit(main())
