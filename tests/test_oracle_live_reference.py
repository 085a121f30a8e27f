"""Live pin of the numpy oracle: the UNMODIFIED reference, imported from ``/root/reference`` (authoring
container only — the checkout does not travel to the GPU box, where this file skips), is run beside
``oracle/box_oracle.py`` on fresh seeded inputs of varying shapes.  The committed fixtures
(tests/golden/*.npz, tests/test_oracle_golden.py) pin one draw per function; this file widens the pin to
many draws, ragged target counts and the edge cases the reference's control flow distinguishes
(empty class, one box, top_k cut, duplicated ground truth, no positives).

Bit-exact for indices, masks, labels and pure add/mul/compare arithmetic; 2e-6 relative where exp/log
differ between numpy and torch libm (same bounds as tests/test_oracle_golden.py)."""
import importlib.util
import os
import sys
import warnings

import numpy as np
import pytest
import torch

from oracle import box_oracle as bo

REF = os.environ.get('RD_REFERENCE', '/root/reference')
pytestmark = [pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'layers')),
                                 reason='reference checkout not present (GPU box): fixtures pin the oracle there'),
              pytest.mark.filterwarnings('ignore')]      # box_utils.py:263-266 out= resizing warns on torch 2.11
VAR = (0.1, 0.2)
RT = dict(rtol=2e-6, atol=1e-7)
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def ref():
    """The reference's modules (pycocotools stubbed) + the fixture generators of make_golden.py."""
    spec = importlib.util.spec_from_file_location('_make_golden', os.path.join(HERE, 'golden', 'make_golden.py'))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    saved_path, saved_mods = list(sys.path), set(sys.modules)
    warnings.filterwarnings('ignore')
    R = mg.import_reference()
    R.mg = mg
    yield R
    sys.path[:] = saved_path
    for name in set(sys.modules) - saved_mods:          # layers / data / utils of the reference
        if name.split('.')[0] in ('layers', 'data', 'utils', 'pycocotools'):
            sys.modules.pop(name, None)


def _priors(R, size, stride, n):
    return R.PriorBox(R.voc[size]).forward()[::stride][:n].contiguous()


@pytest.mark.parametrize('seed', range(6))
def test_elementwise_and_jaccard(ref, seed):
    g = torch.Generator().manual_seed(100 + seed)
    P, G = 64 + 37 * seed, 1 + 3 * seed
    pri = _priors(ref, '320', 11 + seed, P)
    P = pri.shape[0]
    loc = (0.3 + 0.3 * seed) * torch.randn(P, 4, generator=g)
    xy = torch.rand(G, 2, generator=g) * 0.7
    truths = torch.cat([xy, xy + 0.02 + 0.3 * torch.rand(G, 2, generator=g)], 1)
    bu = ref.bu
    dec = bu.decode(loc, pri, list(VAR))
    assert np.array_equal(bo.point_form(pri.numpy()), bu.point_form(pri).numpy())
    assert np.array_equal(bo.center_size(dec.numpy()), bu.center_size(dec).numpy())
    np.testing.assert_allclose(bo.decode(loc.numpy(), pri.numpy(), VAR), dec.numpy(), **RT)
    matched = truths[torch.randint(0, G, (P,), generator=g)]
    np.testing.assert_allclose(bo.encode(matched.numpy(), pri.numpy(), VAR),
                               bu.encode(matched, pri, list(VAR)).numpy(), **RT)
    assert np.array_equal(bo.intersect(truths.numpy(), dec.numpy()), bu.intersect(truths, dec).numpy())
    assert np.array_equal(bo.jaccard(truths.numpy(), dec.numpy()), bu.jaccard(truths, dec).numpy())
    x = (1 + seed) * torch.randn(33, 2 + seed, generator=g)
    # log(sum) + max cancels when the result is small against the global max: one ulp of the max in absolute terms
    np.testing.assert_allclose(bo.log_sum_exp(x.numpy()), bu.log_sum_exp(x).numpy(), rtol=1e-6, atol=1e-6)


def _boxes(g, n, spread, size):
    xy = torch.rand(n, 2, generator=g) * spread
    wh = size * (0.3 + torch.rand(n, 2, generator=g))
    return torch.cat([xy, xy + wh], 1)


@pytest.mark.parametrize('seed', range(8))
def test_box_utils_nms(ref, seed):
    g = torch.Generator().manual_seed(200 + seed)
    n = [1, 2, 17, 64, 150, 300, 301, 40][seed]
    top_k = [200, 1, 5, 200, 100, 200, 300, 7][seed]
    thr = [0.45, 0.45, 0.3, 0.5, 0.45, 0.6, 0.45, 0.1][seed]
    boxes = _boxes(g, n, 0.6, 0.25)                      # heavy overlap: long suppression chains
    scores = torch.rand(n, generator=g)
    assert scores.unique().numel() == n                  # tie-free by construction
    keep_r, count_r = ref.bu.nms(boxes.clone(), scores.clone(), thr, top_k)
    keep_o, count_o = bo.nms(boxes.numpy(), scores.numpy(), thr, top_k)
    assert count_o == int(count_r)
    assert np.array_equal(keep_o[:count_o], keep_r[:count_r].numpy())


def test_box_utils_nms_empty_returns_bare_tensor(ref):
    r = ref.bu.nms(torch.zeros(0, 4), torch.zeros(0), 0.45, 200)
    o = bo.nms(np.zeros((0, 4), np.float32), np.zeros((0,), np.float32), 0.45, 200)
    assert isinstance(r, torch.Tensor) and isinstance(o, np.ndarray)          # box_utils.py:235-236
    assert r.numel() == o.size == 0


@pytest.mark.parametrize('seed', range(8))
def test_py_cpu_nms(ref, seed):
    g = torch.Generator().manual_seed(300 + seed)
    n = [1, 2, 33, 64, 65, 400, 1000, 128][seed]
    thr = [0.45, 0.49, 0.3, 0.45, 0.49, 0.45, 0.49, 0.7][seed]
    side = [512.0, 320.0][seed % 2]
    boxes = _boxes(g, n, 0.7, 0.2) * side
    scores = torch.rand(n, 1, generator=g)
    dets = torch.cat([boxes, scores], 1).numpy()
    assert len(set(dets[:, 4].tolist())) == n
    assert bo.nms_pixel(dets, thr) == [int(i) for i in ref.py_cpu_nms(dets, thr)]


@pytest.mark.parametrize('seed,sparse', [(0, True), (1, False), (2, True), (3, False)])
def test_detect_forward_and_python_nms(ref, seed, sparse):
    g = torch.Generator().manual_seed(400 + seed)
    B, C, top_k = 2, 3 + seed, [40, 25, 60, 10][seed]
    pri = _priors(ref, '320', 23 + 5 * seed, 400)
    P = pri.shape[0]
    arm_loc, arm_conf, odm_loc, odm_conf = ref.mg.gen_detect_inputs(g, B, P, C, sparse)
    det = ref.Detect(C, 320, 0, top_k, 0.01, 0.45, 0.01, 500)
    conf_r = odm_conf.clone()
    boxes_r, scores_r = det.forward(arm_loc, arm_conf, odm_loc, conf_r, pri)
    conf_o = odm_conf.numpy().copy()
    boxes_o, scores_o = bo.detect_forward(arm_loc.numpy(), arm_conf.numpy(), odm_loc.numpy(), conf_o,
                                          pri.numpy(), 0.01, VAR)
    np.testing.assert_allclose(boxes_o, boxes_r.numpy(), rtol=1e-5, atol=1e-6)
    assert np.array_equal(scores_o, scores_r.numpy())
    assert np.array_equal(conf_o, conf_r.numpy())                    # in-place zeroing, :40-42
    # a5 — the oracle walks its own boxes; a 1-ulp exp difference could only matter at an IoU within
    # 1e-6 of the threshold, which the assert on the kept scores would expose
    out_r = det.forward_python_nms(arm_loc, arm_conf, odm_loc, odm_conf.clone(), pri).numpy()
    out_o, _ = bo.forward_python_nms(arm_loc.numpy(), arm_conf.numpy(), odm_loc.numpy(), odm_conf.numpy().copy(),
                                     pri.numpy(), C, top_k, 0.01, 0.45, 0.01, VAR)
    assert out_o.shape == out_r.shape
    assert np.array_equal(out_o[..., 0], out_r[..., 0])
    np.testing.assert_allclose(out_o[..., 1:], out_r[..., 1:], rtol=1e-5, atol=1e-6)
    assert not out_r[:, 0].any()


def _ragged_targets(ref, g, counts, num_classes):
    out = []
    for G in counts:
        out.extend(ref.mg.gen_targets(g, 1, G, num_classes, 0.05, 0.4))
    return out


@pytest.mark.parametrize('seed', range(5))
def test_refine_match_ragged(ref, seed):
    g = torch.Generator().manual_seed(500 + seed)
    C = 6
    pri = _priors(ref, '320', 7 + seed, 700)
    P = pri.shape[0]
    counts = [[1, 4, 9], [2, 2, 13], [7, 1, 3], [5, 20, 1], [3, 3, 3]][seed]
    targets = _ragged_targets(ref, g, counts, C)
    targets[1] = torch.cat([targets[1], targets[1][:1].clone()])     # duplicated GT: last label wins (:149-150)
    targets[1][-1, 4] = float(C - 1)
    arm_loc = 0.5 * torch.randn(len(counts), P, 4, generator=g)
    for b, t in enumerate(targets):
        truths, labels = t[:, :4], t[:, 4]
        for mode in ('arm', 'odm', 'ssd'):
            loc_t, conf_t = torch.zeros(1, P, 4), torch.zeros(1, P, dtype=torch.long)
            if mode == 'arm':
                ref.bu.refine_match(0.5, truths, pri, list(VAR), labels >= 0, loc_t, conf_t, 0)
                lo, co, _, _ = bo.refine_match(0.5, truths.numpy(), pri.numpy(), VAR, labels.numpy() >= 0)
                tol = RT
            elif mode == 'odm':
                ref.bu.refine_match(0.5, truths, pri, list(VAR), labels, loc_t, conf_t, 0, arm_loc[b])
                lo, co, _, _ = bo.refine_match(0.5, truths.numpy(), pri.numpy(), VAR, labels.numpy(),
                                               arm_loc[b].numpy())
                tol = dict(rtol=2e-5, atol=3e-5)         # 1-ulp exp difference amplified by 1/(0.1*w)
            else:
                ref.bu.match(0.5, truths, pri, list(VAR), labels - 1, loc_t, conf_t, 0)
                lo, co, _, _ = bo.refine_match(0.5, truths.numpy(), pri.numpy(), VAR, labels.numpy() - 1,
                                               label_offset=1)
                tol = RT
            assert np.array_equal(co, conf_t[0].numpy()), (mode, b)
            np.testing.assert_allclose(lo, loc_t[0].numpy(), **tol)


@pytest.mark.parametrize('seed', range(4))
def test_multibox_loss_values(ref, seed):
    g = torch.Generator().manual_seed(600 + seed)
    C = [6, 2, 9, 4][seed]
    pri = _priors(ref, '320', 5 + seed, 900)
    P = pri.shape[0]
    counts = [[3, 8, 1, 5], [6, 6], [2, 11, 4], [1, 1, 1, 1, 1]][seed]
    B = len(counts)
    targets = _ragged_targets(ref, g, counts, C)
    arm_loc = 0.1 * torch.randn(B, P, 4, generator=g)
    odm_loc = 0.1 * torch.randn(B, P, 4, generator=g)
    arm_conf = torch.randn(B, P, 2, generator=g)
    odm_conf = torch.randn(B, P, C, generator=g)
    preds = (arm_loc, arm_conf, odm_loc, odm_conf, pri)
    torch.set_default_dtype(torch.float32)
    for nc, use_arm in ((2, False), (C, True)):
        crit = ref.Loss(nc, 0.5, True, 0, True, 3, 0.5, False, False, use_ARM=use_arm)
        ll, lc = crit(preds, targets)
        r = bo.multibox_loss([t.numpy() for t in preds], [t.numpy() for t in targets], nc, 0.5, 3, 0.01,
                             use_arm, VAR)
        np.testing.assert_allclose([r['loss_l'], r['loss_c']], [float(ll), float(lc)], rtol=2e-5)


def test_multibox_loss_no_positives(ref):
    """N < 1 (refinedet_multibox_loss.py:134-139): every ground truth's forced match is ARM-filtered
    (theta above any softmax output), so the ODM criterion returns zeros."""
    g = torch.Generator().manual_seed(77)
    pri = _priors(ref, '320', 9, 500)
    P = pri.shape[0]
    targets = ref.mg.gen_targets(g, 2, 3, 4, 0.05, 0.4)
    preds = (0.1 * torch.randn(2, P, 4, generator=g), torch.randn(2, P, 2, generator=g),
             0.1 * torch.randn(2, P, 4, generator=g), torch.randn(2, P, 4, generator=g), pri)
    crit = ref.Loss(4, 0.5, True, 0, True, 3, 0.5, False, False, theta=1.5, use_ARM=True)
    ll, lc = crit(preds, targets)
    r = bo.multibox_loss([t.numpy() for t in preds], [t.numpy() for t in targets], 4, 0.5, 3, 1.5, True, VAR)
    assert float(ll) == float(lc) == 0.0 and tuple(ll.shape) == (1,)
    assert r['loss_l'] == r['loss_c'] == 0.0 and r['N'] == 0


def test_full_size_config3_and_config4(ref):
    """BASELINE.json sizes (P = 16,320 anchors of RefineDet512; C = 81; 50 ground-truth boxes per image) through
    the reference and the oracle: a3 forward, the a4 candidate + NMS loop on the sparse generator, and the ODM
    match targets — one image each (the reference's Python loops take seconds per image)."""
    pri = ref.PriorBox(ref.coco['512']).forward()
    P, C = pri.shape[0], 81
    assert P == 16320
    for seed in range(900, 920):                      # SURVEY 8d sparse generator; redraw until tie-free per class
        g = torch.Generator().manual_seed(seed)
        d = 2.0 * torch.randn(1, P, generator=g) - 8.0
        arm_conf = torch.softmax(torch.stack([torch.zeros(1, P), d], -1), -1)
        logits = 1.5 * torch.randn(1, P, C, generator=g)
        logits[..., 0] += 4.0
        odm_conf = torch.softmax(logits, -1)
        arm_loc, odm_loc = torch.randn(1, P, 4, generator=g), torch.randn(1, P, 4, generator=g)
        live = odm_conf[0][arm_conf[0, :, 1] > 0.01]
        if all(len(set(col[col > 0.01].tolist())) == int((col > 0.01).sum()) for col in live.t()[1:]):
            break
    else:
        pytest.fail('no tie-free draw in 20 seeds')
    det = ref.Detect(C, 512, 0, 1000, 0.01, 0.45, 0.01, 500)
    conf_r = odm_conf.clone()
    boxes_r, scores_r = det.forward(arm_loc, arm_conf, odm_loc, conf_r, pri)
    conf_o = odm_conf.numpy().copy()
    boxes_o, scores_o = bo.detect_forward(arm_loc.numpy(), arm_conf.numpy(), odm_loc.numpy(), conf_o, pri.numpy(),
                                          0.01, VAR)
    np.testing.assert_allclose(boxes_o, boxes_r.numpy(), rtol=1e-5, atol=1e-6)
    assert np.array_equal(scores_o, scores_r.numpy()) and np.array_equal(conf_o, conf_r.numpy())
    # a4 (eval_refinedet_coco.py:205-232) with the reference's own py_cpu_nms on the reference's boxes
    scale = np.array([512.0] * 4, np.float32)
    dets_o, _ = bo.detect_stage_eval(boxes_r[0].numpy(), scores_r[0].numpy(), scale, 0.01, 1000, 0.45, 500)
    px = boxes_r[0].numpy() * scale[None, :]
    sc = scores_r[0].numpy()
    n_rows = 0
    for j in range(1, C):
        inds = np.where(sc[:, j] > 0.01)[0]
        if len(inds) == 0:
            assert dets_o[j].shape[0] == 0
            continue
        c_scores = sc[inds, j]
        assert len(set(c_scores.tolist())) == len(c_scores)              # the tie-free draw chosen above
        order = c_scores.argsort()[::-1][:1000]
        c_dets = np.hstack((px[inds][order], c_scores[order][:, None])).astype(np.float32)
        keep = ref.py_cpu_nms(c_dets, 0.45)[:500]
        assert np.array_equal(dets_o[j], c_dets[keep, :]), j
        n_rows += len(keep)
    assert n_rows > 5000
    # config 4: ODM match targets, 50 ground-truth boxes
    t = ref.mg.gen_targets(g, 1, 50, C)[0]
    a = 0.1 * torch.randn(P, 4, generator=g)
    loc_t, conf_t = torch.zeros(1, P, 4), torch.zeros(1, P, dtype=torch.long)
    ref.bu.refine_match(0.5, t[:, :4], pri, list(VAR), t[:, 4], loc_t, conf_t, 0, a)
    lo, co, _, _ = bo.refine_match(0.5, t[:, :4].numpy(), pri.numpy(), VAR, t[:, 4].numpy(), a.numpy())
    assert np.array_equal(co, conf_t[0].numpy()) and (co > 0).sum() >= 50
    np.testing.assert_allclose(lo, loc_t[0].numpy(), rtol=2e-5, atol=3e-5)
