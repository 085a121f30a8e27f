"""Property tests (hypothesis) for greedy NMS and the matching rules: invariants of the oracle on the CPU
tier, oracle == kernel on generated inputs (duplicates, ties, degenerate boxes) on the GPU tier."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import box_oracle as bo

F32 = np.float32


@st.composite
def boxes_scores(draw, max_n=40, pixel=False):
    n = draw(st.integers(1, max_n))
    scale = 300.0 if pixel else 1.0
    grid = draw(st.integers(2, 12))                  # a coarse grid makes duplicates and exact ties likely
    cx = np.array(draw(st.lists(st.integers(0, grid), min_size=n, max_size=n)), F32) / grid
    cy = np.array(draw(st.lists(st.integers(0, grid), min_size=n, max_size=n)), F32) / grid
    w = np.array(draw(st.lists(st.integers(0, grid), min_size=n, max_size=n)), F32) / grid
    h = np.array(draw(st.lists(st.integers(0, grid), min_size=n, max_size=n)), F32) / grid
    boxes = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1).astype(F32) * F32(scale)
    scores = np.array(draw(st.lists(st.integers(1, 50), min_size=n, max_size=n)), F32) / F32(50)
    return boxes, scores


def _iou_pixel(a, b):
    w = max(F32(0), min(a[2], b[2]) - max(a[0], b[0]) + F32(1))
    h = max(F32(0), min(a[3], b[3]) - max(a[1], b[1]) + F32(1))
    inter = w * h
    sa = (a[2] - a[0] + F32(1)) * (a[3] - a[1] + F32(1))
    sb = (b[2] - b[0] + F32(1)) * (b[3] - b[1] + F32(1))
    return inter / (sa + sb - inter)


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(boxes_scores(pixel=True), st.sampled_from([0.3, 0.45, 0.7]))
def test_oracle_nms_pixel_invariants(bs, thr):
    boxes, scores = bs
    dets = np.concatenate([boxes, scores[:, None]], 1).astype(F32)
    keep = bo.nms_pixel(dets, thr)
    assert len(set(keep)) == len(keep) and len(keep) >= 1
    ks = [dets[k, 4] for k in keep]
    assert all(ks[i] >= ks[i + 1] for i in range(len(ks) - 1))               # score descending
    for a in range(len(keep)):                                                # kept boxes do not suppress each other
        for b in range(a + 1, len(keep)):
            assert _iou_pixel(dets[keep[a]], dets[keep[b]]) <= F32(thr)
    kept = set(keep)
    for j in range(dets.shape[0]):                                            # every dropped box has a kept suppressor
        if j not in kept:
            assert any(_iou_pixel(dets[k], dets[j]) > F32(thr) for k in keep)
    again = bo.nms_pixel(dets[keep], thr)                                     # idempotent
    assert again == list(range(len(keep)))


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(boxes_scores(max_n=30), st.integers(1, 6))
def test_oracle_refine_match_invariants(bs, G):
    priors_pt, _ = bs
    P = priors_pt.shape[0]
    priors = bo.center_size(priors_pt)
    priors[:, 2:] = np.maximum(priors[:, 2:], F32(0.05))
    rng = np.random.RandomState(P * 7 + G)
    xy = rng.rand(G, 2).astype(F32) * F32(0.7)
    truths = np.concatenate([xy, xy + F32(0.05) + rng.rand(G, 2).astype(F32) * F32(0.3)], 1).astype(F32)
    labels = rng.randint(1, 5, G).astype(F32)
    loc, conf, bti, bto = bo.refine_match(0.5, truths, priors, (0.1, 0.2), labels)
    assert conf.shape == (P,) and loc.shape == (P, 4)
    ov = bo.jaccard(truths, bo.point_form(priors))
    best_prior = ov.argmax(1)
    for j in range(G):                                    # every truth keeps its best prior unless a later truth claims it
        owners = [g for g in range(G) if best_prior[g] == best_prior[j]]
        assert bti[best_prior[j]] == max(owners)
        assert bto[best_prior[j]] == F32(2)
    assert np.all((conf == 0) | (bto >= F32(0.5)))         # positives need overlap >= threshold (or a forced match)
    assert set(np.unique(conf)) <= set([0]) | set(labels.astype(np.int64))


@pytest.mark.gpu
@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(boxes_scores(max_n=64, pixel=True), st.sampled_from([0.0, 0.3, 0.49, 0.9]), st.booleans())
def test_gpu_nms_pixel_equals_oracle(bs, thr, eq):
    import refinedet.pytorch_b200 as rd
    boxes, scores = bs
    dets = np.concatenate([boxes, scores[:, None]], 1).astype(F32)
    assert rd.nms_wrapper.nms(dets, thr, force_cpu=eq) == bo.nms_pixel(dets, thr, suppress_on_equal=eq)


@pytest.mark.gpu
@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(boxes_scores(max_n=64), st.sampled_from([0.0, 0.3, 0.45, 0.9]), st.sampled_from([1, 5, 200]))
def test_gpu_nms_normalised_equals_oracle(bs, thr, top_k):
    import torch
    import refinedet.pytorch_b200 as rd
    boxes, scores = bs
    ek, ec = bo.nms(boxes, scores, thr, top_k)
    keep, count = rd.box_utils.nms(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), thr, top_k)
    assert count == ec and np.array_equal(keep.cpu().numpy(), ek)


@st.composite
def score_tensor(draw):
    B, P, C = draw(st.integers(1, 2)), draw(st.integers(1, 90)), draw(st.integers(1, 11))
    levels = draw(st.integers(2, 40))                 # few distinct values: runs of exact ties cross the top_k cut
    vals = draw(st.lists(st.integers(0, levels), min_size=B * P * C, max_size=B * P * C))
    return (np.array(vals, F32) / F32(levels)).reshape(B, P, C)


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(score_tensor(), st.sampled_from([0.0, 0.01, 0.5]), st.sampled_from([1, 3, 16, 200]), st.integers(0, 2))
def test_oracle_select_topk_invariants(scores, thr, top_k, first_class):
    lists = bo.select_topk(scores, thr, top_k, first_class)
    B, P, C = scores.shape
    for b in range(B):
        for c in range(C):
            idx = lists[b][c]
            col = scores[b, :, c]
            if c < first_class:
                assert idx.size == 0
                continue
            n = int((col > F32(thr)).sum())
            assert idx.size == min(n, top_k) and len(set(idx.tolist())) == idx.size
            s = col[idx]
            assert (s > F32(thr)).all()
            # score descending, lower anchor first inside a run of equal scores
            assert all(s[i] > s[i + 1] or (s[i] == s[i + 1] and idx[i] < idx[i + 1]) for i in range(idx.size - 1))
            if idx.size:                               # nothing left out beats the last one taken
                rest = np.setdiff1d(np.where(col > F32(thr))[0], idx)
                assert all(col[r] < s[-1] or (col[r] == s[-1] and r > idx[-1]) for r in rest)


@pytest.mark.gpu
@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(score_tensor(), st.sampled_from([0.0, 0.01, 0.5]), st.sampled_from([1, 3, 16, 200]), st.integers(0, 2))
def test_gpu_select_topk_equals_oracle(scores, thr, top_k, first_class):
    import torch
    import refinedet.pytorch_b200 as rd
    idx, sc, counts = rd.box_utils.select_topk(torch.from_numpy(scores).cuda(), thr, top_k, first_class)
    idx, sc, counts = idx.cpu().numpy(), sc.cpu().numpy(), counts.cpu().numpy()
    ref = bo.select_topk(scores, thr, top_k, first_class)
    for b in range(scores.shape[0]):
        for c in range(scores.shape[2]):
            n = ref[b][c].size
            assert counts[b, c] == n
            assert np.array_equal(idx[b, c, :n], ref[b][c])
            assert np.array_equal(sc[b, c, :n], scores[b, ref[b][c], c])
