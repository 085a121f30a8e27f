"""GPU tier: packing fused with the multi-GPU gather (``rd_pack_scatter``).

On one GPU the peers are emulated by two exchange buffers of the same device: "rank 0" and "rank 1" each
scatter their detections into both buffers; afterwards both buffers must hold both ranks' counts and packed
rows, identical to ``Detections.packed()`` (pure copies: bit-exact).  The real thing — one process per GPU,
buffers mapped through symmetric memory, stores over NVLink — is ``tools/exchange_check.py`` under torchrun.
"""
import ctypes

import numpy as np
import pytest
import torch

from tests import gen

pytestmark = pytest.mark.gpu


def test_pack_scatter_two_emulated_ranks():
    import refinedet.pytorch_b200 as rd
    from refinedet.pytorch_b200 import _ffi, dist as rdist
    from refinedet.pytorch_b200._ffi import check, lib, ptr, stream_ptr
    C, size, top_k, keep = 21, '320', 1000, 500
    priors = rd.PriorBox(rd.REFINEDET_ANCHORS[size]).forward().cuda()
    det = rd.Detect_RefineDet(C, int(size), 0, top_k, 0.01, 0.45, 0.01, keep)
    scale = np.array([float(size)] * 4, np.float32)
    world, slot_B = 2, 3
    results = []
    for rank, B in enumerate((3, 2)):                       # shards may differ by one image
        a = [t.cuda() for t in gen.detect_inputs(900 + rank, B, priors.shape[0], C, 'sparse', arm_shift=-5.0)]
        results.append(det.detect(*a, priors, scale=scale))
    capacity = slot_B * C * keep
    slot_bytes = int(lib().rd_exchange_slot_bytes(slot_B, C, capacity))
    assert slot_bytes % 256 == 0
    bufs = [torch.zeros(world * slot_bytes, dtype=torch.uint8, device='cuda') for _ in range(world)]
    offsets = torch.empty(slot_B * C + 1, dtype=torch.int32, device='cuda')
    for rank, res in enumerate(results):
        peer = (ctypes.c_void_p * world)(*[b.data_ptr() + rank * slot_bytes for b in bufs])
        B = res.dets.shape[0]
        check(lib().rd_pack_scatter(ptr(res.counts), ptr(res.dets), B, C, keep, ptr(offsets), peer, world, rank,
                                    slot_B, capacity, stream_ptr()), 'rd_pack_scatter')
    torch.cuda.synchronize()
    for buf in bufs:
        counts_all, rows_all = rdist.decode_slots(buf, world, slot_bytes, slot_B * C)
        for rank, res in enumerate(results):
            _, rows = res.packed()
            assert torch.equal(counts_all[rank], res.counts)
            assert torch.equal(rows_all[rank], rows)
            assert rows.shape[0] == int(res.counts.sum()) > 0


def test_pack_scatter_capacity_and_arguments():
    from refinedet.pytorch_b200 import _ffi, dist as rdist
    from refinedet.pytorch_b200._ffi import lib, ptr, stream_ptr
    B, C, max_out = 1, 3, 4
    counts = torch.tensor([[0, 3, 2]], dtype=torch.int32, device='cuda')
    dets = torch.arange(B * C * max_out * 5, dtype=torch.float32, device='cuda').reshape(B, C, max_out, 5)
    offsets = torch.empty(B * C + 1, dtype=torch.int32, device='cuda')
    capacity = 4                                              # 5 rows wanted: the header reports the truncation
    slot_bytes = int(lib().rd_exchange_slot_bytes(B, C, capacity))
    buf = torch.zeros(slot_bytes, dtype=torch.uint8, device='cuda')
    peer = (ctypes.c_void_p * 1)(buf.data_ptr())
    assert lib().rd_pack_scatter(ptr(counts), ptr(dets), B, C, max_out, ptr(offsets), peer, 1, 0, B, capacity,
                                 stream_ptr()) == 0
    torch.cuda.synchronize()
    with pytest.raises(RuntimeError, match='capacity'):
        rdist.decode_slots(buf, 1, slot_bytes, B * C)
    hdr = buf[:16].view(torch.int32).cpu().tolist()
    assert hdr == [4, 1, 3, 5]
    rows = buf[256 + 256:256 + 256 + 4 * 20].view(torch.float32).view(4, 5).cpu()
    assert torch.equal(rows[:3], dets[0, 1, :3].cpu()) and torch.equal(rows[3], dets[0, 2, 0].cpu())
    # argument errors: world out of range, rank out of range, misaligned slot
    assert lib().rd_pack_scatter(ptr(counts), ptr(dets), B, C, max_out, ptr(offsets), peer, 0, 0, B, capacity,
                                 stream_ptr()) == _ffi.RD_ERR_BAD_ARG
    assert lib().rd_pack_scatter(ptr(counts), ptr(dets), B, C, max_out, ptr(offsets), peer, 1, 1, B, capacity,
                                 stream_ptr()) == _ffi.RD_ERR_BAD_ARG
    bad = (ctypes.c_void_p * 1)(buf.data_ptr() + 16)
    assert lib().rd_pack_scatter(ptr(counts), ptr(dets), B, C, max_out, ptr(offsets), bad, 1, 0, B, capacity,
                                 stream_ptr()) == _ffi.RD_ERR_ALIGNMENT


def test_exchange_round_single_rank_and_replay():
    """rd_exchange_round (pack + copy + folded rendezvous) with world = 1: rounds alternate between the two slot
    parities, the epoch lives on the device, so the launch pair can be replayed from a captured plan.  (The
    multi-rank rendezvous spins on flags other GPUs write: it is exercised by tools/exchange_check.py under torchrun,
    never by several emulated ranks on one GPU.)"""
    import refinedet.pytorch_b200 as rd
    from refinedet.pytorch_b200 import dist as rdist
    from refinedet.pytorch_b200._ffi import check, lib, ptr, stream_ptr
    from refinedet.pytorch_b200.layers.functions.detection_refinedet import DetectPlan
    C, size, keep = 21, '320', 500
    priors = rd.PriorBox(rd.REFINEDET_ANCHORS[size]).forward().cuda()
    det = rd.Detect_RefineDet(C, int(size), 0, 1000, 0.01, 0.45, 0.01, keep)
    scale = np.array([float(size)] * 4, np.float32)
    B = 3
    results = [det.detect(*[t.cuda() for t in gen.detect_inputs(910 + k, B, priors.shape[0], C, 'sparse', arm_shift=-5.0)],
                          priors, scale=scale) for k in range(3)]
    capacity = B * C * keep
    slot_bytes = int(lib().rd_exchange_slot_bytes(B, C, capacity))
    ctrl_bytes = int(lib().rd_exchange_ctrl_bytes())
    assert ctrl_bytes == 1024
    buf = torch.zeros(ctrl_bytes + 2 * slot_bytes, dtype=torch.uint8, device='cuda')
    bases = (ctypes.c_void_p * 1)(buf.data_ptr())

    def one_round(res):
        check(lib().rd_exchange_round(ptr(res.counts), ptr(res.dets), B, C, keep, bases, None, 1, 0, B, capacity, 0, 2000,
                                      stream_ptr()), 'rd_exchange_round')

    def latest():
        ctrl = buf[:ctrl_bytes].view(torch.int32).cpu()
        epoch = int(ctrl[64])
        assert int(ctrl[66]) == 0 and int(ctrl[0]) == epoch               # no timeout; own flag = epoch
        half = buf[ctrl_bytes + (epoch & 1) * slot_bytes:ctrl_bytes + ((epoch & 1) + 1) * slot_bytes]
        counts_all, rows_all = rdist.decode_slots(half, 1, slot_bytes, B * C)
        assert int(half[:32].view(torch.int32)[4]) == epoch
        return epoch, counts_all[0], rows_all[0]

    for k, res in enumerate(results):
        one_round(res)
        epoch, counts, rows = latest()
        assert epoch == k + 1
        _, exp_rows = res.packed()
        assert torch.equal(counts, res.counts) and torch.equal(rows, exp_rows)
    # captured once, replayed: the epoch advances on the device
    plan = DetectPlan.capture(torch.device('cuda', torch.cuda.current_device()), lambda: one_round(results[0]))
    for k in range(3):
        plan.launch()
        epoch, counts, rows = latest()
        assert epoch == 4 + k
        assert torch.equal(rows, results[0].packed()[1])
