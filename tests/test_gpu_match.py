"""GPU tier: training-side matching (refine_match / match / encode), hard-negative mining and
the RefineDetMultiBoxLoss drop-in against the reference-generated golden fixtures and the
numpy oracle.  ``conf_t`` / ``pos`` / ``neg`` are bit-exact; ``loc_t`` (log) within 1e-5."""
import numpy as np
import pytest
import torch

from oracle import box_oracle as bo
from tests import gen

pytestmark = pytest.mark.gpu
VAR = [0.1, 0.2]
RTOL, ATOL = 1e-5, 2e-6
# ODM targets are encoded against center_size(decode(arm_loc, prior)): a 1-ulp difference of exp
# between CUDA and the CPU libm moves the refined anchor's centre/size by ~6e-8, and encode divides the
# centre offset by v0*w (w down to 0.03) and takes log((gt_w)/w): the error is amplified to ~2e-5
# absolute.  conf_t and best_truth_idx stay bit-exact.
ATOL_ODM = 1e-4


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope='module')
def rd():
    import refinedet.pytorch_b200 as rd
    rd._ffi.lib()
    return rd


def test_refine_match_golden(rd, golden):
    g = golden('match_loss.npz')
    bu = rd.box_utils
    priors, targets = cu(g['priors']), cu(g['targets'])
    B, P = targets.shape[0], priors.shape[0]
    arm_loc = cu(g['arm_loc'])
    for mode in ('arm', 'odm', 'ssd'):
        loc_t = torch.zeros(B, P, 4).cuda()
        conf_t = torch.zeros(B, P, dtype=torch.long).cuda()
        for idx in range(B):
            truths, labels = targets[idx][:, :-1], targets[idx][:, -1]
            if mode == 'arm':
                bu.refine_match(0.5, truths, priors, VAR, labels >= 0, loc_t, conf_t, idx)
            elif mode == 'odm':
                bu.refine_match(0.5, truths, priors, VAR, labels, loc_t, conf_t, idx, arm_loc[idx])
            else:
                bu.match(0.5, truths, priors, VAR, labels - 1, loc_t, conf_t, idx)
        assert np.array_equal(conf_t.cpu().numpy(), g['conf_t_' + mode]), mode
        np.testing.assert_allclose(loc_t.cpu().numpy(), g['loc_t_' + mode], rtol=RTOL,
                                   atol=ATOL_ODM if mode == 'odm' else ATOL, err_msg=mode)
    with pytest.raises(IndexError):                                  # G = 0: the reference raises too
        bu.refine_match(0.5, targets[0][:0, :-1], priors, VAR, targets[0][:0, -1],
                        torch.zeros(B, P, 4).cuda(), torch.zeros(B, P, dtype=torch.long).cuda(), 0)


def test_loss_golden(rd, golden):
    g = golden('match_loss.npz')
    C = g['odm_conf'].shape[-1]
    preds = tuple(cu(g[k]) for k in ('arm_loc', 'arm_conf', 'odm_loc', 'odm_conf', 'priors'))
    targets = [cu(t) for t in g['targets']]
    arm_crit = rd.RefineDetMultiBoxLoss(2, 0.5, True, 0, True, 3, 0.5, False, True)
    odm_crit = rd.RefineDetMultiBoxLoss(C, 0.5, True, 0, True, 3, 0.5, False, True, use_ARM=True)
    al, ac = arm_crit(preds, targets)
    ol, oc = odm_crit(preds, targets)
    np.testing.assert_allclose([al.item(), ac.item()], g['arm_loss'], rtol=2e-5)
    np.testing.assert_allclose([ol.item(), oc.item()], g['odm_loss'], rtol=2e-5)
    # masks: pos / neg bit-exact against the reference's sort-based ranking
    bu = rd.box_utils
    for mode, conf_key, nc in (('arm', 'arm_conf', 2), ('odm', 'odm_conf', C)):
        pos = cu(g['pos_' + mode])
        neg, num_pos = bu.hnm_select(cu(g['loss_c_rows_' + mode]), pos, 3)
        assert np.array_equal(neg.cpu().numpy(), g['neg_' + mode]), mode
        assert np.array_equal(num_pos.cpu().numpy(), g['pos_' + mode].sum(1))
    # every positive ARM-filtered -> (zeros(1), zeros(1)) like the reference (:135-136)
    arm_conf_off = preds[1].clone()
    arm_conf_off[..., 0] += 50.0
    zl, zc = odm_crit((preds[0], arm_conf_off, preds[2], preds[3], preds[4]), targets)
    assert zl.shape == (1,) and float(zl) == 0.0 and float(zc) == 0.0
    # sync_free extension: same values without the host read of N; N < 1 -> 0-dim zeros with zero gradients
    odm_sf = rd.RefineDetMultiBoxLoss(C, 0.5, True, 0, True, 3, 0.5, False, True, use_ARM=True, sync_free=True)
    sl, sc = odm_sf(preds, targets)
    assert sl.dim() == 0 and sl.is_cuda and float(sl) == float(ol) and float(sc) == float(oc)
    p0 = tuple(t.clone().requires_grad_(True) if i in (2, 3) else t for i, t in enumerate(preds))
    zl2, zc2 = odm_sf((p0[0], arm_conf_off, p0[2], p0[3], p0[4]), targets)
    assert zl2.dim() == 0 and float(zl2) == 0.0 and float(zc2) == 0.0
    (zl2 + zc2).backward()
    assert float(p0[2].grad.abs().sum()) == 0.0 and float(p0[3].grad.abs().sum()) == 0.0
    # gradients flow to the predictions through the stock-PyTorch loss tail
    p2 = tuple(t.clone().requires_grad_(True) if i < 4 else t for i, t in enumerate(preds))
    l, c = odm_crit(p2, targets)
    (l + c).backward()
    assert p2[2].grad is not None and float(p2[2].grad.abs().sum()) > 0
    assert p2[3].grad is not None and float(p2[3].grad.abs().sum()) > 0


@pytest.mark.parametrize('B,size,C,G,use_arm', [(4, '320', 21, 9, True), (4, '320', 21, 9, False),
                                                (3, '512', 81, 50, True), (2, '512', 2, 200, True)])
def test_match_batch_vs_oracle(rd, B, size, C, G, use_arm):
    priors = rd.PriorBox(rd.REFINEDET_ANCHORS[size]).forward()
    P = priors.shape[0]
    small = C == 2
    tg = gen.targets(9000 + G, B, G, C, 0.01 if small else 0.02, 0.06 if small else 0.17)
    # ragged ground truth: drop a different number of boxes per image
    tg = [t[:max(1, G - 3 * i)] for i, t in enumerate(tg)]
    arm_loc, arm_conf, odm_loc, odm_conf = gen.train_predictions(31 + G, B, P, C)
    bu = rd.box_utils
    truths, labels, cnt = bu.pad_targets([t.cuda() for t in tg], 'cuda')
    mode = bu.LABEL_ODM if use_arm else bu.LABEL_ARM_BINARY
    loc_t, conf_t, bt_idx, bt_ov = bu.match_batch(0.5, truths, labels, cnt, priors.cuda(), VAR,
                                                  arm_loc.cuda() if use_arm else None, mode, return_best=True)
    for i in range(B):
        t = tg[i].numpy()
        lab = t[:, 4] if use_arm else (t[:, 4] >= 0)
        l, c, bti, bto = bo.refine_match(0.5, t[:, :4], priors.numpy(), VAR, lab,
                                         arm_loc[i].numpy() if use_arm else None)
        assert np.array_equal(conf_t[i].cpu().numpy(), c), i
        assert np.array_equal(bt_idx[i].cpu().numpy(), bti), i
        pos = c > 0
        np.testing.assert_allclose(loc_t[i].cpu().numpy()[pos], l[pos], rtol=RTOL, atol=ATOL_ODM if use_arm else ATOL)
        if not use_arm:      # ARM branch has no exp/log ahead of jaccard: overlaps are bit-exact
            assert np.array_equal(bt_ov[i].cpu().numpy(), bto)
            np.testing.assert_allclose(loc_t[i].cpu().numpy(), l, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize('B,P', [(3, 1275), (32, 16320)])
def test_hnm_vs_oracle_and_properties(rd, B, P):
    g = torch.Generator().manual_seed(P)
    loss = torch.rand(B, P, generator=g) * 8
    pos = torch.rand(B, P, generator=g) < 0.015
    pos[0] = False                                   # num_pos = 0 -> no negatives
    if B > 2:
        pos[1] = torch.rand(P, generator=g) < 0.4    # 3*num_pos > P-1 -> clamp
    neg, num_pos = rd.box_utils.hnm_select(loss.cuda(), pos.cuda(), 3)
    e_neg, _ = bo.hnm_select(loss.numpy(), pos.numpy(), 3)
    assert np.array_equal(neg.cpu().numpy(), e_neg)
    assert np.array_equal(num_pos.cpu().numpy(), pos.sum(1).numpy())
    n_neg = neg.sum(1).cpu()
    assert torch.equal(n_neg, torch.clamp(3 * pos.sum(1), max=P - 1))
    # every selected loss >= every unselected loss of the row (positives count as zero)
    lz = loss.clone()
    lz[pos] = 0
    lz = lz.cuda()
    for b in range(B):
        if 0 < int(n_neg[b]) < P:
            assert float(lz[b][neg[b]].min()) >= float(lz[b][~neg[b]].max())


def test_full_size_loss_step(rd):
    """BASELINE.json config 4 shape: both criteria run, finite, N = sum(num_pos)."""
    B, P, C, G = 32, 16320, 81, 50
    priors = rd.PriorBox(rd.REFINEDET_ANCHORS['512']).forward().cuda()
    preds = tuple(t.cuda() for t in gen.train_predictions(1234 + 4000, B, P, C)) + (priors,)
    tg = [t.cuda() for t in gen.targets(55, B, G, C)]
    arm_crit = rd.RefineDetMultiBoxLoss(2, 0.5, True, 0, True, 3, 0.5, False, True)
    odm_crit = rd.RefineDetMultiBoxLoss(C, 0.5, True, 0, True, 3, 0.5, False, True, use_ARM=True)
    for crit in (arm_crit, odm_crit):
        l, c = crit(preds, tg)
        assert torch.isfinite(l) and torch.isfinite(c) and float(l) > 0 and float(c) > 0
    loc_t, conf_t = odm_crit.match_targets(preds, tg)
    # every ground truth owns at least one positive anchor (forced match), labels in range
    assert int(conf_t.min()) == 0 and int(conf_t.max()) <= C - 1
    assert int((conf_t > 0).sum(1).min()) >= 1
    l2, c2 = odm_crit.match_targets(preds, tg)
    assert torch.equal(c2, conf_t) and torch.equal(l2, loc_t)
    # the fused criterion call (rd_multibox_criterion) leaves the same targets / masks as the step-by-step entry points
    for crit, use_arm in ((arm_crit, False), (odm_crit, True)):
        lt, ct = crit.match_targets(preds, tg)
        f_lt, f_ct = crit.last_targets
        assert torch.equal(f_ct, ct) and torch.equal(f_lt, lt)
        conf = preds[3] if use_arm else preds[1]
        ce, lse, pos = rd.box_utils.conf_loss(conf, ct, preds[1] if use_arm else None, 0.01)
        neg, num_pos = rd.box_utils.hnm_select(ce, pos, 3)
        f_pos, f_neg = crit.last_masks
        assert torch.equal(f_pos, pos) and torch.equal(f_neg, neg)
        ll, lc, n = rd.box_utils.multibox_loss_reduce(preds[2] if use_arm else preds[0], lt, ce, pos, neg, num_pos)
        fl, fc = crit(preds, tg)
        assert abs(float(fl) - float(ll)) <= 2e-7 * abs(float(ll)) and abs(float(fc) - float(lc)) <= 2e-7 * abs(float(lc))
        assert float(n) == float(pos.sum())


@pytest.mark.parametrize('P', [16384, 20000, 300])
def test_criterion_tail_paths(rd, P):
    """The one-call criterion with the cluster mining kernel (rows up to 16,384 anchors) and with the single-CTA one
    (longer rows), on priors that are not a RefineDet grid: against the step-by-step entry points and stock PyTorch."""
    B, C, G = 3, 7, 6
    g = torch.Generator().manual_seed(P)
    cxcy = torch.rand(P, 2, generator=g)
    wh = 0.03 + 0.3 * torch.rand(P, 2, generator=g)
    priors = torch.cat([cxcy, wh], 1).cuda()
    arm_loc, arm_conf, odm_loc, odm_conf = [t.cuda() for t in gen.train_predictions(P + 1, B, P, C)]
    tg = [t.cuda() for t in gen.targets(P + 2, B, G, C, 0.05, 0.4)]
    crit = rd.RefineDetMultiBoxLoss(C, 0.5, True, 0, True, 3, 0.5, False, True, use_ARM=True)
    p_loc = odm_loc.clone().requires_grad_(True)
    p_conf = odm_conf.clone().requires_grad_(True)
    preds = (arm_loc, arm_conf, p_loc, p_conf, priors)
    fl, fc = crit(preds, tg)
    lt, ct = crit.match_targets(preds, tg)
    ce, lse, pos = rd.box_utils.conf_loss(odm_conf, ct, arm_conf, 0.01)
    neg, num_pos = rd.box_utils.hnm_select(ce, pos, 3)
    f_pos, f_neg = crit.last_masks
    assert torch.equal(f_pos, pos) and torch.equal(f_neg, neg) and int(pos.sum()) > 0
    ll, lc, n = rd.box_utils.multibox_loss_reduce(odm_loc, lt, ce, pos, neg, num_pos)
    assert abs(float(fl) - float(ll)) <= 2e-7 * abs(float(ll)) and abs(float(fc) - float(lc)) <= 2e-7 * abs(float(lc))
    rl, rc = _torch_tail(odm_loc, odm_conf, lt, ct, pos, neg)
    torch.testing.assert_close(fl, rl, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(fc, rc, rtol=1e-5, atol=1e-7)
    (fl + fc).backward()
    assert bool(torch.isfinite(p_conf.grad).all()) and float(p_conf.grad.abs().sum()) > 0


@pytest.mark.parametrize('P', [16320, 16384, 6375, 300, 20000])
def test_hnm_degenerate_rows(rd, P):
    """The interval-narrowing select of hnm_cluster_kernel (P <= 16,384; the radix kernel beyond) on rows built to
    defeat a value-based bucketing: all losses equal (ties broken by index only), two distinct values, a few huge
    outliers, losses that differ in their last bit, inf / tiny values, every anchor positive but one."""
    g = torch.Generator().manual_seed(P)
    rows = [torch.full((P,), 2.5), torch.zeros(P),
            torch.where(torch.rand(P, generator=g) < 0.5, torch.tensor(1.0), torch.tensor(3.0)),
            torch.cat([torch.rand(P - 3, generator=g), torch.tensor([1e30, 3e38, float('inf')])]),
            (torch.full((P,), 4.0).view(torch.int32) + torch.randint(0, 4, (P,), generator=g, dtype=torch.int32)).view(torch.float32),
            torch.rand(P, generator=g) * 1e-38, torch.rand(P, generator=g) * 8, torch.rand(P, generator=g)]
    loss = torch.stack(rows)
    B = loss.shape[0]
    pos = torch.rand(B, P, generator=g) < 0.02
    pos[-1] = True
    pos[-1, P // 2] = False                          # one negative, 3 * num_pos clamps to P - 1 ... of which one exists
    pos[-2] = torch.rand(P, generator=g) < 0.3       # clamp bites
    neg, num_pos = rd.box_utils.hnm_select(loss.cuda(), pos.cuda(), 3)
    e_neg, _ = bo.hnm_select(loss.numpy(), pos.numpy(), 3)
    assert np.array_equal(num_pos.cpu().numpy(), pos.sum(1).numpy())
    got = neg.cpu().numpy()
    for b in range(B):
        assert np.array_equal(got[b], e_neg[b]), (b, int(got[b].sum()), int(e_neg[b].sum()))


def _torch_tail(loc_data, conf_data, loc_t, conf_t, pos, neg):
    """Stock-PyTorch fp32 restatement of refinedet_multibox_loss.py:105-110,126-138 for GIVEN masks."""
    import torch.nn.functional as F
    C = conf_data.shape[-1]
    pos_idx = pos.unsqueeze(pos.dim()).expand_as(loc_data)
    loss_l = F.smooth_l1_loss(loc_data[pos_idx].view(-1, 4), loc_t[pos_idx].view(-1, 4), reduction='sum')
    sel = pos | neg
    conf_p = conf_data[sel.unsqueeze(2).expand_as(conf_data)].view(-1, C)
    loss_c = F.cross_entropy(conf_p, conf_t[sel], reduction='sum')
    N = pos.sum().float()
    return loss_l / N, loss_c / N


@pytest.mark.parametrize('B,size,C,G,use_arm', [(4, '320', 21, 9, True), (3, '320', 2, 12, False),
                                                (2, '512', 81, 50, True), (32, '512', 81, 50, True)])
def test_loss_tail_kernels(rd, B, size, C, G, use_arm):
    """f-4 / a11 / a12: rd_conf_loss, rd_multibox_loss_reduce and rd_multibox_loss_backward against stock
    PyTorch (fp32, tolerance 1e-5 relative — exp/log differ in the last bits and the sums are reduced in
    a different order) and, at the small sizes, against the oracle's analytic gradients."""
    bu = rd.box_utils
    priors = rd.PriorBox(rd.REFINEDET_ANCHORS[size]).forward().cuda()
    P = priors.shape[0]
    arm_loc, arm_conf, odm_loc, odm_conf = [t.cuda() for t in gen.train_predictions(77 + G, B, P, max(C, 2))]
    tg = [t.cuda() for t in gen.targets(9100 + G, B, G, C)]
    crit = rd.RefineDetMultiBoxLoss(C, 0.5, True, 0, True, 3, 0.5, False, True, use_ARM=use_arm)
    if use_arm:
        arm_conf[..., 1] -= 4.0        # theta gate: some matched anchors pass, some are filtered
        loc_data, conf_data = odm_loc, odm_conf
    else:
        loc_data, conf_data = arm_loc, arm_conf[..., :2].contiguous()
    preds = (arm_loc, arm_conf, odm_loc, odm_conf, priors) if use_arm else (arm_loc, conf_data, odm_loc, odm_conf, priors)
    loc_t, conf_t = crit.match_targets(preds, tg)
    # --- rd_conf_loss ---------------------------------------------------------------------------
    ce, lse, pos = bu.conf_loss(conf_data, conf_t, arm_conf if use_arm else None, 0.01)
    flat = conf_data.view(-1, conf_data.shape[-1])
    t_lse = torch.logsumexp(flat, 1)
    t_ce = (t_lse - flat.gather(1, conf_t.view(-1, 1)).squeeze(1)).view(B, P)
    torch.testing.assert_close(lse.view(-1), t_lse, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(ce, t_ce, rtol=1e-5, atol=2e-6)
    torch.testing.assert_close(bu.log_sum_exp(flat), t_lse.unsqueeze(1), rtol=1e-5, atol=1e-6)
    t_pos = conf_t > 0
    if use_arm:
        t_pos = t_pos & ~(torch.softmax(arm_conf, 2)[:, :, 1] <= 0.01)
        assert int((conf_t > 0).sum()) > int(t_pos.sum()) > 0          # the gate really removes some
    assert torch.equal(pos, t_pos)
    # --- forward / backward through the module --------------------------------------------------
    p_loc = loc_data.clone().requires_grad_(True)
    p_conf = conf_data.clone().requires_grad_(True)
    preds2 = (arm_loc, arm_conf, p_loc, p_conf, priors) if use_arm else (p_loc, p_conf, odm_loc, odm_conf, priors)
    l, c = crit(preds2, tg)
    (2.0 * l + 0.5 * c).backward()
    pos_k, neg_k = crit.last_masks
    assert torch.equal(pos_k, t_pos)
    assert torch.equal(neg_k.sum(1), torch.clamp(3 * pos_k.sum(1), max=P - 1))
    r_loc = loc_data.clone().requires_grad_(True)
    r_conf = conf_data.clone().requires_grad_(True)
    rl, rc = _torch_tail(r_loc, r_conf, loc_t, conf_t, pos_k, neg_k)
    (2.0 * rl + 0.5 * rc).backward()
    torch.testing.assert_close(l, rl, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(c, rc, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(p_loc.grad, r_loc.grad, rtol=1e-5, atol=1e-9)
    torch.testing.assert_close(p_conf.grad, r_conf.grad, rtol=2e-5, atol=1e-9)
    if B <= 4:
        N = float(pos_k.sum())
        g_loc, g_conf = bo.multibox_loss_grads(loc_data.cpu().numpy(), conf_data.cpu().numpy(), loc_t.cpu().numpy(),
                                               conf_t.cpu().numpy(), pos_k.cpu().numpy(), neg_k.cpu().numpy(), N)
        np.testing.assert_allclose(p_loc.grad.cpu().numpy(), 2.0 * g_loc, rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(p_conf.grad.cpu().numpy(), 0.5 * g_conf, rtol=2e-5, atol=1e-9)
    # only one of the two losses used: the other gradient is zero, nothing is left unwritten
    p_loc.grad = None
    p_conf.grad = None
    l2, c2 = crit(preds2, tg)
    c2.backward()
    assert float(p_loc.grad.abs().sum()) == 0.0 and float(p_conf.grad.abs().sum()) > 0
    assert bool(torch.isfinite(p_conf.grad).all())


@pytest.mark.parametrize('B,size,C,G', [(3, '320', 21, 9), (32, '512', 81, 50)])
@pytest.mark.parametrize('concurrent', [True, False])
def test_criterion_pair(rd, B, size, C, G, concurrent):
    """RefineDetCriterionPair (rd_multibox_criterion_pair / rd_multibox_loss_backward_pair: the two chains on two
    streams) == the two modules called one after the other as train_refinedet.py:252-256 does: losses, masks,
    targets and all four gradients BIT-identical (the same kernels on the same inputs)."""
    priors = rd.PriorBox(rd.REFINEDET_ANCHORS[size]).forward().cuda()
    P = priors.shape[0]
    tp = [t.cuda() for t in gen.train_predictions(300 + G, B, P, C)]
    tp[1][..., 1] -= 3.0                                    # the theta gate removes some ODM positives
    tg = [t.cuda() for t in gen.targets(9300 + G, B, G, C)]
    mk = lambda: [t.clone().requires_grad_(True) for t in tp]                       # noqa: E731
    arm_crit = rd.RefineDetMultiBoxLoss(2, 0.5, True, 0, True, 3, 0.5, False, True)
    odm_crit = rd.RefineDetMultiBoxLoss(C, 0.5, True, 0, True, 3, 0.5, False, True, use_ARM=True)
    ref = mk()
    preds = tuple(ref) + (priors,)
    al, ac = arm_crit(preds, tg)
    ol, oc = odm_crit(preds, tg)
    (1.5 * al + ac + ol + 0.25 * oc).backward()
    ref_masks = [m.clone() for m in arm_crit.last_masks + odm_crit.last_masks]
    ref_targets = [t.clone() for t in arm_crit.last_targets + odm_crit.last_targets]
    for sync_free in (False, True):
        a2 = rd.RefineDetMultiBoxLoss(2, 0.5, True, 0, True, 3, 0.5, False, True, sync_free=sync_free)
        o2 = rd.RefineDetMultiBoxLoss(C, 0.5, True, 0, True, 3, 0.5, False, True, use_ARM=True, sync_free=sync_free)
        pair = rd.RefineDetCriterionPair(a2, o2, concurrent=concurrent)
        got = mk()
        pal, pac, pol, poc = pair(tuple(got) + (priors,), tg)
        for x, y in ((pal, al), (pac, ac), (pol, ol), (poc, oc)):
            assert float(x) == float(y) and float(x) > 0
        (1.5 * pal + pac + pol + 0.25 * poc).backward()
        for g, r in zip(got, ref):
            assert torch.equal(g.grad, r.grad)
        for m, r in zip(a2.last_masks + o2.last_masks, ref_masks):
            assert torch.equal(m, r)
        for t, r in zip(a2.last_targets + o2.last_targets, ref_targets):
            assert torch.equal(t, r)
    # only the ODM losses used: the ARM gradients are zeros, nothing is left unwritten
    got = mk()
    _, _, pol, poc = pair(tuple(got) + (priors,), tg)
    (pol + poc).backward()
    assert float(got[0].grad.abs().sum()) == 0.0 and float(got[1].grad.abs().sum()) == 0.0
    assert float(got[3].grad.abs().sum()) > 0


def test_criterion_pair_empty_odm(rd):
    """Every matched anchor fails the ARM-theta gate: the ODM criterion has N = 0.  The pair returns the reference's
    (zeros(1), zeros(1)) for it (refinedet_multibox_loss.py:135-136) and the ARM losses unchanged; with sync_free the
    ODM losses are device zeros with zero gradients."""
    B, C, G = 2, 21, 5
    priors = rd.PriorBox(rd.REFINEDET_ANCHORS['320']).forward().cuda()
    P = priors.shape[0]
    tp = [t.cuda() for t in gen.train_predictions(812, B, P, C)]
    tp[1][..., 0] += 40.0                                   # softmax(arm_conf)[1] ~ 0 everywhere
    tg = [t.cuda() for t in gen.targets(813, B, G, C)]
    arm_crit = rd.RefineDetMultiBoxLoss(2, 0.5, True, 0, True, 3, 0.5, False, True)
    al, ac = arm_crit(tuple(tp) + (priors,), tg)
    for sync_free in (False, True):
        a2 = rd.RefineDetMultiBoxLoss(2, 0.5, True, 0, True, 3, 0.5, False, True, sync_free=sync_free)
        o2 = rd.RefineDetMultiBoxLoss(C, 0.5, True, 0, True, 3, 0.5, False, True, use_ARM=True, sync_free=sync_free)
        got = [t.clone().requires_grad_(True) for t in tp]
        pal, pac, pol, poc = rd.RefineDetCriterionPair(a2, o2)(tuple(got) + (priors,), tg)
        assert float(pal) == float(al) and float(pac) == float(ac)
        assert float(pol) == 0.0 and float(poc) == 0.0
        if sync_free:
            assert pol.is_cuda and pol.dim() == 0
            (pal + pac + pol + poc).backward()
            assert float(got[2].grad.abs().sum()) == 0.0 and float(got[3].grad.abs().sum()) == 0.0
            assert float(got[1].grad.abs().sum()) > 0
        else:
            assert not pol.is_cuda and tuple(pol.shape) == (1,)
    with pytest.raises(ValueError):
        rd.RefineDetCriterionPair(o2, a2)


@pytest.mark.parametrize('variant', ['zs', 'tile', 'regs'])
def test_loss_backward_variants(variant):
    """The library picks the backward kernel by the class count (zero-stream for C >= 48, tile for odd C, register
    stores otherwise); ``RD_BWD`` forces one of them for every shape.  The variable is read once per process, so
    the stock-PyTorch comparison above (C = 2 / 21 / 81, B up to 32) is re-run in a child process per variant."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RD_BWD=variant)
    r = subprocess.run([sys.executable, '-m', 'pytest', os.path.join(root, 'tests', 'test_gpu_match.py'), '-q', '-x',
                        '-k', 'test_loss_tail_kernels', '-p', 'no:cacheprovider'],
                       cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert '4 passed' in r.stdout, r.stdout[-500:]


def test_pad_targets_kernel(rd):
    """rd_pad_targets (detection_collate's ragged list -> padded batch): ragged counts incl. an empty image,
    host and device inputs, the ARM/ODM cache hit and its invalidation by an in-place edit."""
    bu = rd.box_utils
    tg = gen.targets(17, 5, 7, 21)
    tg = [tg[0], tg[1][:3], tg[2][:0], tg[3][:1], tg[4]]
    for src in ([t.cuda() for t in tg], tg):
        bu.clear_pad_cache()
        truths, labels, cnt = bu.pad_targets(src, 'cuda')
        assert truths.shape == (5, 7, 4) and cnt.tolist() == [7, 3, 0, 1, 7]
        for i, t in enumerate(tg):
            n = t.shape[0]
            assert torch.equal(truths[i, :n].cpu(), t[:, :4]) and torch.equal(labels[i, :n].cpu(), t[:, 4])
            assert float(truths[i, n:].abs().sum()) == 0.0 and float(labels[i, n:].abs().sum()) == 0.0
    dev = [t.cuda() for t in tg]
    a = bu.pad_targets(dev, 'cuda')
    assert bu.pad_targets(dev, 'cuda')[0] is a[0]                 # second criterion of the step: cached
    dev[1][0, 0] += 0.25                                           # in-place edit bumps the version
    b = bu.pad_targets(dev, 'cuda')
    assert b[0] is not a[0] and float(b[0][1, 0, 0]) == float(dev[1][0, 0])


def test_check_targets(rd):
    """SURVEY f-3: the per-coordinate validation loop of train_refinedet.py:240-245 as one reduction."""
    tg = [t.cuda() for t in gen.targets(3, 4, 6, 21)]
    truths, labels, cnt = rd.box_utils.check_targets(tg)
    assert truths.shape == (4, 6, 4) and cnt.tolist() == [6, 6, 6, 6]
    bad = [t.clone() for t in tg]
    bad[2][3, 1] = 1.0001
    with pytest.raises(StopIteration):
        rd.box_utils.check_targets(bad)
    bad[2][3, 1] = -1e-6
    with pytest.raises(StopIteration):
        rd.box_utils.check_targets(bad)
    ragged = [tg[0][:2], tg[1], tg[2][:1], tg[3]]           # padding rows (zeros) must not trip the check
    rd.box_utils.check_targets(ragged)
