"""GPU tier: ``rd_select_topk`` and the §8b-named aliases ``rd_decode_filter`` / ``rd_detect`` /
``rd_workspace_bytes`` against the numpy oracle and the golden fixtures.  Candidate lists are index work:
bit-exact, ties included (documented rule: lower anchor first)."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import box_oracle as bo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def rd():
    import refinedet.pytorch_b200 as rd
    rd._ffi.lib()
    return rd


def _check(rd, scores, thr, top_k, first_class=1):
    idx, sc, counts = rd.box_utils.select_topk(torch.as_tensor(scores).cuda(), thr, top_k, first_class)
    idx, sc, counts = idx.cpu().numpy(), sc.cpu().numpy(), counts.cpu().numpy()
    ref = bo.select_topk(scores, thr, top_k, first_class)
    B, P, C = scores.shape
    for b in range(B):
        for c in range(C):
            n = ref[b][c].shape[0]
            assert counts[b, c] == n, (b, c)
            assert np.array_equal(idx[b, c, :n], ref[b][c]), (b, c)
            assert np.array_equal(sc[b, c, :n], scores[b, ref[b][c], c]), (b, c)
    return counts


@pytest.mark.parametrize('B,P,C,top_k,thr', [
    (1, 1, 2, 10, 0.01),            # one anchor
    (2, 33, 3, 1000, 0.01),         # top_k > P
    (3, 1000, 5, 7, 0.01),          # deep select: 7 of ~200
    (2, 6375, 21, 1000, 0.01),      # config 2 shape, dense scores: every class saturates top_k
    (1, 16320, 81, 1000, 0.01),     # config 3 shape
    (2, 5000, 2, 4096, 0.001),      # SAR 2-class, 4096-key sort buffer
    (1, 16320, 2, 16384, -1.0),     # everything passes: n = P <= top_k, 128 KB of keys
    (2, 300, 4, 50, 2.0),           # nothing passes
    (1, 16320, 2, 1000, 0.01),      # config 5 dense: ~12 k candidates > list capacity -> radix-select fallback
    (2, 1500, 9, 1000, -1.0),       # top_k < n = 1500 <= list capacity 2048: shared-memory radix select, 1024-key sort
    (1, 70000, 3, 100, 0.3),        # P > 65536: 32-bit candidate lists
    (2, 3000, 11, 200, 0.001),      # 11 classes = one full group of 8 + a ragged one; every list overflows
])
def test_select_topk_vs_oracle(rd, B, P, C, top_k, thr):
    g = torch.Generator().manual_seed(B * 1000 + P + C + top_k)
    scores = torch.softmax(3 * torch.randn(B, P, C, generator=g), -1).numpy()
    counts = _check(rd, scores, thr, top_k)
    assert not counts[:, 0].any()
    if thr > 1.0:
        assert not counts.any()


def test_select_topk_ties_nan_and_first_class(rd):
    g = torch.Generator().manual_seed(5)
    scores = torch.rand(2, 700, 3, generator=g)
    scores = (scores * 16).round() / 16          # 17 distinct values: the select must cut through runs of ties
    scores[0, 5, 1] = float('nan')               # NaN > thr is false: never a candidate (eval :214)
    scores[1, 9, 2] = float('inf')
    s = scores.numpy()
    for top_k in (1, 10, 100, 300, 699, 700, 2000):       # 300: n ~ 560 in (top_k, capacity]: smem select through ties
        _check(rd, s, 0.2, top_k)
    counts = _check(rd, s, 0.2, 50, first_class=0)      # background column too
    assert counts[:, 0].all()
    _check(rd, s, 0.2, 50, first_class=3)               # nothing evaluated


def test_select_topk_matches_detect_golden(rd, golden):
    """On the reference-generated fixture: select_topk of the reference's own `scores` output gives the
    candidate list whose NMS (oracle, reference boxes) is the fixture's a4 result."""
    g = golden('detect_sparse.npz')
    C, top_k, keep_top_k, conf_thr, nms_thr, obj_thr = g['params']
    C, top_k = int(C), int(top_k)
    idx, sc, counts = rd.box_utils.select_topk(torch.as_tensor(g['scores']).cuda(), float(conf_thr), top_k)
    idx, sc, counts = idx.cpu().numpy(), sc.cpu().numpy(), counts.cpu().numpy()
    for b in range(g['scores'].shape[0]):
        boxes = g['boxes'][b] * g['scale'][None, :]
        for c in range(1, C):
            n = counts[b, c]
            dets = np.hstack([boxes[idx[b, c, :n]], sc[b, c, :n, None]]).astype(np.float32)
            keep = bo.nms_pixel(dets, float(nms_thr))[:int(keep_top_k)]
            m = int(g['a4_counts'][b, c])
            assert len(keep) == m
            assert np.array_equal(dets[keep], g['a4_dets'][b, c, :m])


def test_select_topk_argument_errors(rd):
    L = rd._ffi.lib()
    x = torch.zeros(1, 4, 2, device='cuda')
    out_i = torch.zeros(8, dtype=torch.int32, device='cuda')
    p = rd._ffi.ptr
    assert L.rd_select_topk(None, 1, 4, 2, 0.0, 4, 1, p(out_i), None, p(out_i), None) == rd._ffi.RD_ERR_BAD_ARG
    assert L.rd_select_topk(p(x), 1, 4, 2, 0.0, 0, 1, p(out_i), None, p(out_i), None) == rd._ffi.RD_ERR_BAD_ARG
    assert L.rd_select_topk(p(x), 1, 1 << 20, 2, 0.0, 1 << 20, 1, p(out_i), None, p(out_i), None) == \
        rd._ffi.RD_ERR_UNSUPPORTED
    with pytest.raises(ValueError):
        rd.box_utils.select_topk(torch.zeros(4, 2, device='cuda'), 0.0, 4)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        rd.box_utils.select_topk(torch.zeros(1, 4, 2), 0.0, 4)


def test_survey_8b_named_entry_points(rd, golden):
    """rd_decode_filter / rd_detect / rd_workspace_bytes are the §8b names of rd_detect_forward /
    rd_detect_fused / rd_detect_workspace_bytes: same arguments, identical results."""
    L, p, sp = rd._ffi.lib(), rd._ffi.ptr, rd._ffi.stream_ptr
    g = golden('detect_sparse.npz')
    C, top_k, keep_top_k, conf_thr, nms_thr, obj_thr = g['params']
    C, top_k, keep = int(C), int(top_k), int(keep_top_k)
    arm_loc, arm_conf, odm_loc, pri = (torch.as_tensor(g[k]).cuda() for k in ('arm_loc', 'arm_conf', 'odm_loc', 'priors'))
    B, P = arm_loc.shape[:2]
    assert L.rd_workspace_bytes(B, P, C) == L.rd_detect_workspace_bytes(B, P, C)
    conf = torch.as_tensor(g['odm_conf']).cuda()
    boxes = torch.empty(B, P, 4, device='cuda')
    scores = torch.empty(B, P, C, device='cuda')
    rd._ffi.check(L.rd_decode_filter(p(arm_loc), p(arm_conf), p(odm_loc), p(conf), p(pri), B, P, C, float(obj_thr),
                                     0.1, 0.2, p(boxes), p(scores), sp()), 'rd_decode_filter')
    np.testing.assert_allclose(boxes.cpu().numpy(), g['boxes'], rtol=1e-5, atol=1e-6)
    assert np.array_equal(scores.cpu().numpy(), g['scores'])
    assert np.array_equal(conf.cpu().numpy(), g['conf_after'])
    det = rd.Detect_RefineDet(C, 320, 0, top_k, float(conf_thr), float(nms_thr), float(obj_thr), keep)
    conf = torch.as_tensor(g['odm_conf']).cuda()
    scale = torch.as_tensor(g['scale']).cuda().reshape(1, 4).expand(B, 4).contiguous()
    want = det.detect(arm_loc, arm_conf, odm_loc, conf, pri, scale=scale)
    ws = det.new_workspace(B, P, 'cuda')
    out = det.new_outputs(B, 'cuda')
    rd._ffi.check(L.rd_detect(p(arm_loc), p(arm_conf), p(odm_loc), p(conf), p(pri), B, P, C, float(obj_thr),
                              float(conf_thr), float(nms_thr), top_k, out.dets.shape[2], p(scale),
                              rd._ffi.RD_NMS_PIXEL_PLUS1, rd._ffi.RD_ROW_BOX_SCORE, 0.1, 0.2, p(ws), ws.numel(),
                              p(out.counts), p(out.dets), p(out.anchors), sp()), 'rd_detect')
    assert torch.equal(out.counts, want.counts)
    assert np.array_equal(out.counts.cpu().numpy(), g['a4_counts'])
    cnt = out.counts.cpu().numpy()
    a, b = out.dets.cpu().numpy(), want.dets.cpu().numpy()
    for i in range(B):
        for c in range(C):
            assert np.array_equal(a[i, c, :cnt[i, c]], b[i, c, :cnt[i, c]])
