"""Pin the numpy oracle to outputs of the unmodified reference (tests/golden/*.npz,
written by tests/golden/make_golden.py).  Index / mask results are bit-exact; float
results involving exp/log are held to 2e-6 relative (numpy vs torch libm), everything
else is exact."""
import hashlib

import numpy as np
import pytest

from oracle import box_oracle as bo

VAR = (0.1, 0.2)
RT = dict(rtol=2e-6, atol=1e-7)


def test_priors(golden):
    g = golden('priors.npz')
    for size, rows in (('320', [0, 1, 2, 4800, 6000, 6300, 6374]),
                       ('512', [0, 1, 2, 12288, 15360, 16128, 16319])):
        p = bo.prior_box(bo.REFINEDET_CFG[size])
        assert hashlib.sha256(p.tobytes()).hexdigest() == str(g['sha' + size])
        assert np.array_equal(p[rows], g['rows' + size])
        assert abs(p.astype(np.float64).sum() - float(g['sum' + size])) < 1e-9
    assert bo.prior_box(bo.REFINEDET_CFG['320']).shape == (6375, 4)
    assert bo.prior_box(bo.REFINEDET_CFG['512']).shape == (16320, 4)


def test_box_utils_elementwise(golden):
    g = golden('box_utils.npz')
    pri, loc, truths, matched = g['priors'], g['loc'], g['truths'], g['matched']
    assert np.array_equal(bo.point_form(pri), g['point_form'])
    dec = bo.decode(loc, pri, VAR)
    np.testing.assert_allclose(dec, g['decode'], **RT)
    assert np.array_equal(bo.center_size(g['decode']), g['center_size'])
    np.testing.assert_allclose(bo.encode(matched, pri, VAR), g['encode'], **RT)
    assert np.array_equal(bo.intersect(truths, bo.point_form(pri)), g['intersect'])
    assert np.array_equal(bo.jaccard(truths, bo.point_form(pri)), g['jaccard'])
    assert np.array_equal(bo.jaccard(truths, g['decode']), g['jaccard_dec'])
    np.testing.assert_allclose(bo.log_sum_exp(g['x']), g['log_sum_exp'], rtol=1e-6)


def test_decode_kat_survey_appendix_b(golden):
    g = golden('box_utils.npz')
    d1 = bo.decode(g['kat_arm'], g['kat_pri'], VAR)
    np.testing.assert_allclose(d1, [[0.39548290, 0.45184419, 0.61651707, 0.54415584],
                                    [-0.04851007, -0.0175, 0.05351007, 0.0825]], rtol=1e-6)
    cs = bo.center_size(d1)
    np.testing.assert_allclose(cs, [[0.50599998, 0.49800003, 0.22103417, 0.09231165],
                                    [0.0025, 0.0325, 0.10202014, 0.1]], rtol=1e-6)
    d2 = bo.decode(g['kat_odm'], cs, VAR)
    np.testing.assert_allclose(d2, [[0.38644654, 0.45778579, 0.59460866, 0.55483037],
                                    [-0.06584629, -0.00404091, 0.07186650, 0.07004091]], rtol=2e-6)
    tr = np.array([[0.40, 0.42, 0.62, 0.58], [0, 0, 0.08, 0.07]], np.float32)
    np.testing.assert_allclose(bo.jaccard(tr, d1), [[0.56116617, 0], [0, 0.31068420]], rtol=1e-6)
    np.testing.assert_allclose(bo.encode(tr, cs, VAR),
                               [[0.18096785, 0.21665458, -0.02339852, 2.75004601],
                                [3.67574453, 0.24999975, -1.21565425, -1.78330326]], rtol=2e-6)


def test_nms_box_utils(golden):
    g = golden('nms_box_utils.npz')
    k, c = bo.nms(g['kat_boxes'], g['kat_scores'], 0.45, 200)
    assert c == 3 and list(k) == [3, 4, 5, 0, 0, 0]
    assert np.array_equal(k, g['kat_keep200']) and c == int(g['kat_count200'])
    k, c = bo.nms(g['kat_boxes'], g['kat_scores'], 0.45, 3)
    assert np.array_equal(k, g['kat_keep3']) and c == int(g['kat_count3']) == 1
    e = bo.nms(np.zeros((0, 4), np.float32), np.zeros((0,), np.float32), 0.45, 200)
    assert isinstance(e, np.ndarray) and e.size == 0        # bare array, box_utils.py:235-236
    for tag in 'abc':
        thr, tk = g['args_' + tag]
        k, c = bo.nms(g['boxes'], g['scores'], float(thr), int(tk))
        assert c == int(g['count_' + tag])
        assert np.array_equal(k, g['keep_' + tag])


def test_nms_pixel(golden):
    g = golden('nms_pixel.npz')
    for tag, thr in (('045', 0.45), ('049', 0.49), ('070', 0.7)):
        assert bo.nms_pixel(g['dets'], thr) == list(g['keep_' + tag])
    assert bo.nms_pixel(np.zeros((0, 5), np.float32), 0.5) == []
    # cpu_nms.pyx:65 flavour only differs at exact equality
    assert bo.nms_pixel(g['dets'], 0.45, suppress_on_equal=True) == list(g['keep_045'])
    two = np.array([[0, 0, 9, 9, .9], [0, 0, 9, 4, .8]], np.float32)     # IoU exactly 0.5
    assert bo.nms_pixel(two, 0.5) == [0, 1]
    assert bo.nms_pixel(two, 0.5, suppress_on_equal=True) == [0]


@pytest.mark.parametrize('tag', ['sparse', 'dense'])
def test_detect(golden, tag):
    g = golden('detect_%s.npz' % tag)
    C, top_k, keep_top_k, conf_thr, nms_thr, obj_thr = g['params']
    C, top_k, keep_top_k = int(C), int(top_k), int(keep_top_k)
    conf = g['odm_conf'].copy()
    boxes, scores = bo.detect_forward(g['arm_loc'], g['arm_conf'], g['odm_loc'], conf,
                                      g['priors'], obj_thr, VAR)
    np.testing.assert_allclose(boxes, g['boxes'], rtol=1e-5, atol=1e-6)
    assert np.array_equal(scores, g['scores'])
    assert np.array_equal(conf, g['conf_after'])                 # in-place side effect
    assert scores is not conf
    filtered = g['arm_conf'][:, :, 1] <= np.float32(obj_thr)
    assert filtered.any() and (~filtered).any()
    assert not conf[filtered].any()
    # a4 on the reference's boxes (so a 1-ulp exp difference cannot flip an IoU)
    for b in range(boxes.shape[0]):
        dets, _ = bo.detect_stage_eval(g['boxes'][b], g['scores'][b], g['scale'], conf_thr,
                                       top_k, nms_thr, keep_top_k)
        for j in range(C):
            n = int(g['a4_counts'][b, j])
            assert dets[j].shape[0] == n
            assert np.array_equal(dets[j], g['a4_dets'][b, j, :n])
    assert g['a4_counts'][:, 1:].max() == keep_top_k or tag == 'sparse'


@pytest.mark.parametrize('tag', ['sparse', 'dense'])
def test_forward_python_nms(golden, tag):
    g = golden('detect_%s.npz' % tag)
    C, top_k, keep_top_k, conf_thr, nms_thr, obj_thr = g['params']
    out, _ = bo.forward_python_nms(g['arm_loc'], g['arm_conf'], g['odm_loc'], g['odm_conf'].copy(),
                                   g['priors'], int(C), int(top_k), conf_thr, nms_thr, obj_thr, VAR)
    ref = g['a5_output']
    assert out.shape == ref.shape
    assert np.array_equal(out[..., 0], ref[..., 0])              # same kept scores, same order
    np.testing.assert_allclose(out[..., 1:], ref[..., 1:], rtol=1e-5, atol=1e-6)
    assert not ref[:, 0].any()


def test_refine_match(golden):
    g = golden('match_loss.npz')
    pri, tg = g['priors'], g['targets']
    for b in range(tg.shape[0]):
        truths, labels = tg[b, :, :4], tg[b, :, 4]
        loc, conf, _, _ = bo.refine_match(0.5, truths, pri, VAR, labels >= 0)
        assert np.array_equal(conf, g['conf_t_arm'][b])
        np.testing.assert_allclose(loc, g['loc_t_arm'][b], **RT)
        loc, conf, _, _ = bo.refine_match(0.5, truths, pri, VAR, labels, g['arm_loc'][b])
        assert np.array_equal(conf, g['conf_t_odm'][b])
        # ODM encode divides a centre difference by 0.1*w (w ~ 0.1): a 1-ulp exp difference in the
        # refined anchor (6e-8) is amplified ~100x in absolute terms
        np.testing.assert_allclose(loc, g['loc_t_odm'][b], rtol=2e-5, atol=3e-5)
        loc, conf, _, _ = bo.refine_match(0.5, truths, pri, VAR, labels - 1, label_offset=1)
        assert np.array_equal(conf, g['conf_t_ssd'][b])
        np.testing.assert_allclose(loc, g['loc_t_ssd'][b], **RT)
    # the duplicated GT of image 2: the shared best prior carries the LAST label (5)
    assert (g['conf_t_odm'][2] == 5).sum() >= 1


def test_multibox_loss(golden):
    g = golden('match_loss.npz')
    preds = (g['arm_loc'], g['arm_conf'], g['odm_loc'], g['odm_conf'], g['priors'])
    targets = list(g['targets'])
    for mode, nc, use_arm in (('arm', 2, False), ('odm', 6, True)):
        r = bo.multibox_loss(preds, targets, nc, 0.5, 3, 0.01, use_arm, VAR)
        assert np.array_equal(r['pos'], g['pos_' + mode])
        assert np.array_equal(r['neg'], g['neg_' + mode])
        np.testing.assert_allclose([r['loss_l'], r['loss_c']], g[mode + '_loss'], rtol=2e-5)
        np.testing.assert_allclose(r['loss_c_rows'], g['loss_c_rows_' + mode], rtol=1e-5, atol=1e-6)


def test_model_cfg1_fixture(golden):
    """BASELINE.json config 1: head outputs of the reference RefineDet320/VOC model (random init) and the
    reference's detect outputs for one image (tests/golden/make_golden_model.py)."""
    g = golden('model_cfg1.npz')
    C, top_k, keep, conf_thr, nms_thr, obj_thr = g['params']
    conf = g['odm_conf'][None].copy()
    boxes, scores = bo.detect_forward(g['arm_loc'][None], g['arm_conf'][None], g['odm_loc'][None], conf,
                                      g['priors'], float(obj_thr))
    np.testing.assert_allclose(boxes[0], g['boxes'], **RT)
    assert np.array_equal(scores[0], g['scores'])
    out, _ = bo.detect_stage_eval(g['boxes'], g['scores'], np.array([320.0] * 4, np.float32), float(conf_thr),
                                  int(top_k), float(nms_thr), int(keep))
    checked = 0
    for j in range(1, int(C)):
        if g['a4_tie_free'][j]:                       # with tied scores the reference's order is unstable
            n = int(g['a4_counts'][j])
            assert out[j].shape[0] == n
            assert np.array_equal(out[j], g['a4_dets'][j, :n])
            checked += 1
    assert checked >= 1


def test_oracle_loss_grads_match_autograd():
    """oracle.multibox_loss_grads (analytic) == torch autograd of the reference's loss tail
    (refinedet_multibox_loss.py:105-138) on the CPU, for given masks."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(5)
    B, P, C = 2, 300, 7
    loc = torch.randn(B, P, 4, generator=g) * 1.5
    loc_t = torch.randn(B, P, 4, generator=g)
    conf = torch.randn(B, P, C, generator=g) * 2
    conf_t = torch.randint(0, C, (B, P), generator=g)
    pos = (conf_t > 0) & (torch.rand(B, P, generator=g) < 0.2)
    neg = ~pos & (torch.rand(B, P, generator=g) < 0.3)
    N = float(pos.sum())
    lr, cr = loc.clone().requires_grad_(True), conf.clone().requires_grad_(True)
    pos_idx = pos.unsqueeze(2).expand_as(lr)
    loss_l = F.smooth_l1_loss(lr[pos_idx].view(-1, 4), loc_t[pos_idx].view(-1, 4), reduction='sum') / N
    sel = pos | neg
    loss_c = F.cross_entropy(cr[sel.unsqueeze(2).expand_as(cr)].view(-1, C), conf_t[sel], reduction='sum') / N
    (loss_l + loss_c).backward()
    g_loc, g_conf = bo.multibox_loss_grads(loc.numpy(), conf.numpy(), loc_t.numpy(), conf_t.numpy(), pos.numpy(),
                                           neg.numpy(), N)
    np.testing.assert_allclose(g_loc, lr.grad.numpy(), rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(g_conf, cr.grad.numpy(), rtol=1e-5, atol=1e-8)


def test_select_topk_oracle_is_the_front_half_of_the_eval_loop(golden):
    """``select_topk`` restates eval_refinedet_coco.py:214-222; NMS of its lists on the reference's own
    boxes / scores must give the fixture's a4 rows (outputs of the unmodified reference)."""
    g = golden('detect_dense.npz')
    C, top_k, keep_top_k, conf_thr, nms_thr, obj_thr = g['params']
    lists = bo.select_topk(g['scores'], conf_thr, int(top_k))
    for b in range(g['scores'].shape[0]):
        boxes = g['boxes'][b] * g['scale'][None, :]
        assert lists[b][0].size == 0
        for c in range(1, int(C)):
            idx = lists[b][c]
            assert idx.size <= int(top_k)
            sc = g['scores'][b, idx, c]
            assert (np.diff(sc) <= 0).all() and (sc > conf_thr).all()
            dets = np.hstack([boxes[idx], sc[:, None]]).astype(np.float32)
            keep = bo.nms_pixel(dets, float(nms_thr))[:int(keep_top_k)]
            m = int(g['a4_counts'][b, c])
            assert len(keep) == m and np.array_equal(dets[keep], g['a4_dets'][b, c, :m])
    # ties: lower anchor first; NaN is never a candidate
    s = np.array([[[0.0, 0.5], [0.0, 0.7], [0.0, 0.5], [0.0, np.nan], [0.0, 0.005]]], np.float32)
    assert bo.select_topk(s, 0.01, 10)[0][1].tolist() == [1, 0, 2]
    assert bo.select_topk(s, 0.01, 2)[0][1].tolist() == [1, 0]
